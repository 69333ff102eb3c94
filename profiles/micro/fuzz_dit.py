"""One-off soak of the iterative solve (solve_mwis_dit): random 32-wide models and batches, tensor-core vs CUDA-core path.
Reports vertices whose membership differs (near-tie flips are possible: the two paths' scores differ at the 1e-6 level)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np, scipy.sparse as sp
from distgcn_b200 import engine as E
from distgcn_b200.batch import pack_graphs
from distgcn_b200.ckpt import LayerWeights
ctx = E.Context(0)
tot_diff = tot_n = 0
for seed in range(int(sys.argv[1]), int(sys.argv[2])):
    rng = np.random.default_rng(7000 + seed)
    adjs = []
    for k in range(int(rng.integers(100, 400))):
        n = int(rng.integers(1, 331))
        up = np.triu(rng.random((n, n)) < float(rng.choice([0.01, 0.03, 0.08])), k=1)
        adjs.append(sp.csr_matrix((up | up.T).astype(np.float64)))
    pb = pack_graphs(adjs)
    w = rng.random(pb.n_nodes) * (rng.random(pb.n_nodes) > 0.05)
    nh = int(rng.integers(1, 6))
    dims = (1,) + (32,) * (nh + 1) + (1,)
    layers = [LayerWeights(weights=[(rng.standard_normal((ci, co)) / np.sqrt(ci + co)).astype(np.float32) for _ in range(2)])
              for ci, co in zip(dims[:-1], dims[1:])]
    model = E.Model(ctx, layers, [1] * (len(dims) - 2) + [0])
    batch = E.DeviceBatch(ctx, pb)
    os.environ.pop("DG_DISABLE_TC", None); E.reload_env()
    a = E.solve_dit(ctx, model, batch, w, want_steps=True); ka = ctx.last_kernel
    os.environ["DG_DISABLE_TC"] = "1"; E.reload_env()
    b = E.solve_dit(ctx, model, batch, w, want_steps=True); kb = ctx.last_kernel
    d = int((a.member != b.member).sum())
    tot_diff += d; tot_n += pb.n_nodes
    print("seed %d: %d hidden, %d graphs, %s vs %s: %d of %d vertices differ, steps equal %s" % (
        seed, nh, pb.n_graphs, ka, kb, d, pb.n_nodes, bool(np.array_equal(a.steps, b.steps))), flush=True)
    batch.close(); model.close()
print("total: %d of %d vertices differ" % (tot_diff, tot_n))
