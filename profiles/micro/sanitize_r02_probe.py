"""Round-2 kernels under compute-sanitizer (memcheck / synccheck; racecheck with DG_DISABLE_TC=1 - racecheck does not
model the mbarrier-completed TMA copies of tc_solve_kernel): upper-triangle host format (tensor-core staging and the
device-side expansion), graph-staged streaming layer and SpMM, iterative solve on the tensor cores, cheb2 scalar network,
threshold greedy, the device-resident wireless slot loop."""
import os, sys
sys.path.insert(0, '.')
import numpy as np
from distgcn_b200 import engine as E
from distgcn_b200 import wireless as W
from tests import util
pb, w = util.small_graphs()
pb = pb.slice(20, 30)
w = w[: pb.n_nodes].copy()
w[::5] = 0.0
ctx = E.Context(0)
upper = pb.upper_compact()
for short in ("is4sat_l20_c32", "is4sat_l2_c64"):
    layers = util.load_layers(short)
    model = E.Model(ctx, layers, E.gcn_dqn_acts(len(layers)))
    m0, _ = E.solve_host(ctx, model, pb, w)
    m1, _ = E.solve_host(ctx, model, pb, w, upper=upper)
    os.environ["DG_DISABLE_TC"] = "1"; os.environ["DG_DISABLE_FUSED"] = "1"; E.reload_env()
    m2, _ = E.solve_host(ctx, model, pb, w, upper=upper)         # expansion + graph-staged streaming layers
    if not os.environ.get("SANITIZE_NO_TC"):
        del os.environ["DG_DISABLE_TC"]
    del os.environ["DG_DISABLE_FUSED"]; E.reload_env()
    batch = E.DeviceBatch(ctx, pb)
    r = E.solve_dit(ctx, model, batch, w)
    d = E.dist_greedy(ctx, batch, w, 0.1)
    z = np.random.default_rng(0).random((pb.n_nodes, 32)).astype(np.float32)
    y = E.spmm_laplacian(ctx, batch, z)
    print(short, bool(np.array_equal(m0, m1)), bool(np.array_equal(m0, m2)), int(r.member.sum()), int(d.member.sum()), float(np.abs(y).max()))
    batch.close(); model.close()
layers = util.layers_from_meta_fixture(util.load_npz("meta_activations.npz"), "is4sat_l2_c1_cheb2")
model = E.Model(ctx, layers, E.gcn_dqn_acts(len(layers)))
batch = E.DeviceBatch(ctx, pb)
print("cheb2", float(np.abs(E.gcn_forward(ctx, model, batch)).max()))
batch.close(); model.close()
insts = W.make_instances(n_networks=2, loads=[0.5], n_ch=3, timeslots=6, seed=3, n_nodes=40, area=120.0)
layers = util.load_layers("is4sat_l1")
model = E.Model(ctx, layers, E.gcn_dqn_acts(len(layers)))
for algo in ("DGCN-LGS", "DGCN-LGS-Seq", "Greedy-Th"):
    sim = W.BatchedScheduler(ctx, insts, algo, model)
    print(algo, float(sim.run().sum()))
    sim.close()
model.close()
ctx.close()
print("probe done")
