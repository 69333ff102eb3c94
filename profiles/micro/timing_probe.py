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
E.solve(ctx, model, batch, w)
