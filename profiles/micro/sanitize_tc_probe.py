"""Small run of the tensor-core kernel and the dist-greedy kernel for compute-sanitizer."""
import sys
sys.path.insert(0, '.')
import numpy as np
from distgcn_b200 import engine as E
from tests import util
pb, w = util.small_graphs()
pb = pb.slice(20, 32)
w = w[: pb.n_nodes].copy()
w[::5] = 0.0
ctx = E.Context(0)
layers = util.load_layers("is4sat_l20_c32")
model = E.Model(ctx, layers, E.gcn_dqn_acts(len(layers)))
batch = E.DeviceBatch(ctx, pb)
r = E.solve(ctx, model, batch, w, want_score=True, want_util=True, want_steps=True)
print(ctx.last_kernel, int(r.member.sum()))
d = E.dist_greedy(ctx, batch, w, epsilon=0.1)
print("dist greedy", int(d.member.sum()), d.steps.tolist())
batch.close(); model.close(); ctx.close()
print("probe done")
