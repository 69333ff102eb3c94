"""DG_TC_DEBUG probe: one tile (a 1-block and a 3-block graph) through tc_solve_kernel; with a -DDG_TC_TRACE build the
per-warp event clocks of the tile are printed, also after a protocol error."""
import os, sys
sys.path.insert(0, '.')
os.environ['DG_TC_DEBUG'] = '1'
import numpy as np, scipy.sparse as sp
from distgcn_b200 import engine as E
from distgcn_b200.batch import pack_graphs
from tests import util
rng = np.random.default_rng(5)
sizes = [int(x) for x in (sys.argv[1:] or ["100", "300"])]
adjs = []
for n in sizes:
    up = np.triu(rng.random((n, n)) < 8.0 / n, k=1)
    adjs.append(sp.csr_matrix((up | up.T).astype(np.float64)))
pb = pack_graphs(adjs)
w = rng.random(pb.n_nodes)
layers = util.load_layers('is4sat_l20_c32')
ctx = E.Context(0); model = E.Model(ctx, layers, E.gcn_dqn_acts(len(layers))); batch = E.DeviceBatch(ctx, pb)
try:
    m = E.solve(ctx, model, batch, w).member
    print("ok", int(np.asarray(m).sum()), ctx.last_kernel)
    os.environ['DG_DISABLE_TC'] = '1'
    E.reload_env()
    m2 = E.solve(ctx, model, batch, w).member
    print("same as the CUDA-core kernel:", bool(np.array_equal(np.asarray(m), np.asarray(m2))), ctx.last_kernel)
except Exception as e:
    print("ERR", e)
