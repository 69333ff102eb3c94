"""tc_solve_kernel with an ODD number of hidden layers on a batch of more tiles than SMs (every CTA takes several tiles):
memberships and scores against the CUDA-core kernel.  (The per-block barriers complete an odd number of phases per tile
then: a re-arming that did not reset the phase would show here and nowhere in the shipped 20-layer model.)"""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
import numpy as np, scipy.sparse as sp
from distgcn_b200 import engine as E
from distgcn_b200.batch import pack_graphs
from distgcn_b200.ckpt import LayerWeights
rng = np.random.default_rng(3)
n_graphs = int(sys.argv[1]) if len(sys.argv) > 1 else 700
adjs = []
for k in range(n_graphs):
    n = int(rng.integers(100, 301))
    up = np.triu(rng.random((n, n)) < 6.0 / n, k=1)
    adjs.append(sp.csr_matrix((up | up.T).astype(np.float64)))
pb = pack_graphs(adjs)
w = rng.random(pb.n_nodes)
w[rng.random(pb.n_nodes) < 0.05] = 0.0
ctx = E.Context(0)
batch = E.DeviceBatch(ctx, pb)
for dims in ((1, 32, 32, 32, 32, 1), (1, 32, 32, 1), (1, 32, 32, 32, 32, 32, 32, 1), (1, 32, 32, 32, 1)):
    layers = [LayerWeights(weights=[(rng.standard_normal((ci, co)) / np.sqrt(ci + co)).astype(np.float32) for _ in range(2)])
              for ci, co in zip(dims[:-1], dims[1:])]
    acts = [1] * (len(dims) - 2) + [0]
    model = E.Model(ctx, layers, acts)
    os.environ.pop("DG_DISABLE_TC", None); E.reload_env()
    r = E.solve(ctx, model, batch, w, want_score=True)
    k1 = ctx.last_kernel
    os.environ["DG_DISABLE_TC"] = "1"; E.reload_env()
    r2 = E.solve(ctx, model, batch, w, want_score=True)
    k2 = ctx.last_kernel
    scale = np.abs(r2.score).max()
    print("hidden layers %d: %s vs %s, memberships equal %s, max score difference / scale %.2e" % (
        len(dims) - 3, k1, k2, bool(np.array_equal(r.member, r2.member)), float(np.abs(r.score - r2.score).max() / scale)))
    model.close()
# the iterative solve (solve_mwis_dit) with an odd number of hidden layers: completes, and agrees with the CUDA-core path up to
# near-tie flips
for dims in ((1, 32, 32, 32, 32, 1),):
    layers = [LayerWeights(weights=[(rng.standard_normal((ci, co)) / np.sqrt(ci + co)).astype(np.float32) for _ in range(2)])
              for ci, co in zip(dims[:-1], dims[1:])]
    acts = [1] * (len(dims) - 2) + [0]
    model = E.Model(ctx, layers, acts)
    os.environ.pop("DG_DISABLE_TC", None); E.reload_env()
    r = E.solve_dit(ctx, model, batch, w)
    k1 = ctx.last_kernel
    os.environ["DG_DISABLE_TC"] = "1"; E.reload_env()
    r2 = E.solve_dit(ctx, model, batch, w)
    print("iterative, hidden layers %d: %s vs %s, vertices differing %d of %d, members %d / %d" % (
        len(dims) - 3, k1, ctx.last_kernel, int((r.member != r2.member).sum()), pb.n_nodes, int(r.member.sum()), int(r2.member.sum())))
    model.close()
