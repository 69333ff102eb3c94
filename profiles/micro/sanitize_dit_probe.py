import os, sys
sys.path.insert(0, '.')
import numpy as np, scipy.sparse as sp
from distgcn_b200 import engine as E
from distgcn_b200.batch import pack_graphs
from tests import util
rng = np.random.default_rng(5)
sizes = [int(x) for x in sys.argv[2:]]
adjs = []
for n in sizes:
    up = np.triu(rng.random((n, n)) < 8.0 / n, k=1)
    adjs.append(sp.csr_matrix((up | up.T).astype(np.float64)))
pb = pack_graphs(adjs)
w = rng.random(pb.n_nodes)
layers = util.load_layers('is4sat_l20_c32')
ctx = E.Context(0); model = E.Model(ctx, layers, E.gcn_dqn_acts(len(layers))); batch = E.DeviceBatch(ctx, pb)
if sys.argv[1] == "dit":
    r = E.solve_dit(ctx, model, batch, w)
else:
    r = E.solve(ctx, model, batch, w)
print(sys.argv[1], sizes, int(r.member.sum()), ctx.last_kernel)
