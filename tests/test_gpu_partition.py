"""Row-partitioned single-graph path, checked against the CPU oracle: scores within 1e-5 of
the float64 evaluation, membership exactly what oracle.lgs gives on the same utilities.  world_size 1 on one GPU;
test_partitioned_multi_rank spawns the 2-rank run (tests/run_partition_2gpu.py) when the box has two GPUs."""
import os

import numpy as np
import pytest
import scipy.sparse as sp

from tests import util

pytestmark = pytest.mark.gpu


def _big_graph(rng, n, deg):
    m = n * deg // 2
    u = rng.integers(0, n, m)
    v = rng.integers(0, n, m)
    ok = u != v
    a = sp.coo_matrix((np.ones(ok.sum()), (u[ok], v[ok])), shape=(n, n))
    a = ((a + a.T) > 0).astype(np.float64).tocsr()
    a.sort_indices()
    return a


class _ModelSpec:
    def __init__(self, layers, acts):
        self._layers, self._acts = layers, acts
        self.layers = [type("L", (), {"act_code": a})() for a in acts]

    def layers_as_weights(self):
        return self._layers


def oracle_check(a, w, layers, member, score, n):
    """The CPU oracle is the reference for the partitioned path: scores against the float64 evaluation of the network
    on the zero-weight-reduced graph (oracle.gcn_forward_fp64), membership against oracle.lgs run on the path's OWN
    utilities (fp32 score x fp64 weight), which makes the membership check exact whatever the last-bit rounding."""
    from oracle import gcn_oracle as G
    from oracle import lgs as L
    keep = np.where(w > 0)[0]
    ar = a[keep][:, keep].tocsr()
    feats = G.features_gen1(w[keep], layers[0].c_in)
    exact = np.zeros(n)
    exact[keep] = G.gcn_forward_fp64(feats, G.laplacian_supports(ar, 1), layers, "gcn_dqn")[:, 0]
    err = float(np.abs(score[:n] - exact).max()) / max(float(np.abs(exact).max()), 1e-30)
    o = L.run(a.indptr, a.indices, score[:n].astype(np.float64) * w, init_remain=(w != 0).astype(np.uint8))
    return err, bool(np.array_equal(o.member, member[:n])), o.steps


@pytest.mark.parametrize("short", ["is4sat_l2_c64", "is4sat_l3_c16", "is4sat_l20_c32"])
def test_partitioned_matches_oracle(gpu_ctx, short):
    """world size 1 exercises every slice kernel with row0 = 0; the result is checked against the CPU oracle, and
    the single-device path (streaming kernels + global-bitmap greedy search) must agree with it as well."""
    import torch
    from distgcn_b200 import engine as E
    from distgcn_b200.batch import pack_graphs
    from distgcn_b200.shard import RowPartitionedSolver, slice_csr
    rng = np.random.default_rng(5)
    n = 20011 if short != "is4sat_l20_c32" else 9001  # > 8192: the single-device run takes the generic path too
    a = _big_graph(rng, n, 12)
    w = rng.random(n)
    w[rng.random(n) < 0.1] = 0.0
    layers = util.load_layers(short)
    acts = E.gcn_dqn_acts(len(layers))
    rp, ci = slice_csr(a.indptr, a.indices, n, 0, 1)
    solver = RowPartitionedSolver(_ModelSpec(layers, acts), n, rp, ci, rank=0, world_size=1)
    member, score, rounds = solver.solve(w)
    solver.close()
    err, same, steps = oracle_check(a, w, layers, member, score, n)
    assert err <= 1e-5, "scores vs float64 oracle: %.3g" % err
    assert same, "membership differs from oracle.lgs on the same utilities"
    assert rounds == steps
    # the single-device path on the same graph, against the same oracle
    model = E.Model(gpu_ctx, layers, acts)
    batch = E.DeviceBatch(gpu_ctx, pack_graphs([a]))
    ref = E.solve(gpu_ctx, model, batch, w, want_score=True)
    batch.close()
    model.close()
    err1, same1, _ = oracle_check(a, w, layers, ref.member, ref.score[:, 0], n)
    assert err1 <= 1e-5 and same1
    torch.cuda.synchronize()


@pytest.mark.parametrize("world", [2])
def test_partitioned_multi_rank(world):
    """The real multi-rank run: `world` processes (one per GPU, NCCL rendezvous on 127.0.0.1), both exchange modes
    (NCCL all-gather and peer-arena stores fused into the kernels), three models, checked against the CPU oracle by
    tests/run_partition_2gpu.py.  Skipped on a box with fewer GPUs."""
    import os
    import subprocess
    import sys
    import torch
    if torch.cuda.device_count() < world:
        pytest.skip("needs %d GPUs" % world)
    script = os.path.join(util.ROOT, "tests", "run_partition_2gpu.py")
    env = dict(os.environ, DG_PART_N="60013")
    port = 29500 + (os.getpid() % 2000)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
           "--master-addr", "127.0.0.1", "--master-port", str(port), script]
    r = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=900)
    print(r.stdout[-4000:])
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "PARTITION_2GPU_OK" in r.stdout


@pytest.mark.gpu
def test_two_devices_in_one_process():
    """One process driving two GPUs: every kernel's opt-in shared-memory attribute belongs to the function ON A DEVICE, so
    the second device must get its own (they used to be set once per process).  The same batches through the tensor-core,
    CUDA-core (upper-triangle host format: the expansion kernel too) and per-layer paths and the threshold greedy on
    device 0, then on device 1: identical results."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    from distgcn_b200 import engine as E
    pb, w = util.small_graphs()
    pb = pb.slice(0, 24)
    w = w[: pb.n_nodes].copy()
    w[::6] = 0.0
    upper = pb.upper_compact()
    out = {}
    for dev in (0, 1):
        ctx = E.Context(dev)
        res = []
        for short in ("is4sat_l20_c32", "is4sat_l1", "is4sat_l3_c16"):
            layers = util.load_layers(short)
            model = E.Model(ctx, layers, E.gcn_dqn_acts(len(layers)))
            res.append(E.solve_host(ctx, model, pb, w, upper=upper)[0])
            os.environ["DG_DISABLE_TC"] = "1"
            os.environ["DG_DISABLE_FUSED"] = "1"
            E.reload_env()
            res.append(E.solve_host(ctx, model, pb, w)[0])
            del os.environ["DG_DISABLE_TC"], os.environ["DG_DISABLE_FUSED"]
            E.reload_env()
            model.close()
        batch = E.DeviceBatch(ctx, pb)
        res.append(E.dist_greedy(ctx, batch, w, 0.1).member)
        res.append(E.lgs(ctx, batch, w).member)
        batch.close()
        ctx.close()
        out[dev] = res
    for a, b in zip(out[0], out[1]):
        assert np.array_equal(a, b)
