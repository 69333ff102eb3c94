import os, sys
sys.path.insert(0, '.')
import numpy as np
from distgcn_b200 import engine as E
from tests import util
pb, w, _ = util.full_set('ba')
layers = util.load_layers('is4sat_l20_c32')
ctx = E.Context(0); model = E.Model(ctx, layers, E.gcn_dqn_acts(len(layers))); batch = E.DeviceBatch(ctx, pb)
for i in range(3): E.solve(ctx, model, batch, w)
os.environ['DG_FUSED_TIMING'] = '1'
E.reload_env()
E.solve(ctx, model, batch, w)
import os
if os.environ.get('DG_FUSED_TILE_DUMP'):
    d = np.loadtxt(os.environ['DG_FUSED_TILE_DUMP'])
    cyc, n, nnzp, ng = d[:, 0], d[:, 1], d[:, 2], d[:, 3]
    A = np.stack([nnzp, n, ng, np.ones_like(n)], axis=1)
    coef, *_ = np.linalg.lstsq(A, cyc, rcond=None)
    pred = A @ coef
    print('tiles', len(cyc), 'fit cycles = %.2f*nnz_padded + %.1f*n + %.0f*ng + %.0f' % tuple(coef),
          'rel rms err %.3f' % (np.sqrt(np.mean((pred - cyc) ** 2)) / cyc.mean()))
