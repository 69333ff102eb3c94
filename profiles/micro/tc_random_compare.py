"""Tensor-core kernel against the CUDA-core graph-resident kernel on random graphs of many sizes (1 .. 370 vertices,
several densities, zero weights sprinkled in): scores within 2e-5 of the score scale, memberships equal except where two
adjacent vertices' utilities are closer than the kernels' score difference (reported, expected to be none or a handful)."""
import os, subprocess, sys
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))


def make(seed, n_graphs):
    import scipy.sparse as sp
    from distgcn_b200.batch import pack_graphs
    rng = np.random.default_rng(seed)
    adjs = []
    for _ in range(n_graphs):
        n = int(rng.integers(1, 371))
        p = float(rng.choice([0.01, 0.05, 0.1, 0.3]))
        up = np.triu(rng.random((n, n)) < p, k=1)
        adjs.append(sp.csr_matrix((up | up.T).astype(np.float64)))
    pb = pack_graphs(adjs)
    w = rng.random(pb.n_nodes)
    w[rng.random(pb.n_nodes) < 0.05] = 0.0
    return pb, w


def run(seed, n_graphs):
    from distgcn_b200 import engine as E
    from tests import util
    pb, w = make(seed, n_graphs)
    layers = util.load_layers("is4sat_l20_c32")
    ctx = E.Context(0)
    model = E.Model(ctx, layers, E.gcn_dqn_acts(len(layers)))
    batch = E.DeviceBatch(ctx, pb)
    r = E.solve(ctx, model, batch, w, predict="mwis", remove_zero_weight=True, want_score=True, want_util=True)
    return ctx.last_kernel, r.member, r.score[:, 0], r.util, pb


if __name__ == "__main__":
    seed = int(sys.argv[1]) if len(sys.argv) > 1 else 0
    n_graphs = int(sys.argv[2]) if len(sys.argv) > 2 else 3000
    if os.environ.get("TC_COMPARE_CHILD"):
        k, m, s, u, pb = run(seed, n_graphs)
        np.savez(os.environ["TC_COMPARE_CHILD"], kernel=k, member=m, score=s, util=u)
        sys.exit(0)
    out = {}
    for tag, env in (("tc", {}), ("fused", {"DG_DISABLE_TC": "1"})):
        path = "/tmp/tc_compare_%s.npz" % tag
        e = dict(os.environ, TC_COMPARE_CHILD=path, **env)
        subprocess.check_call([sys.executable, os.path.abspath(__file__), str(seed), str(n_graphs)], env=e)
        out[tag] = np.load(path)
    a, b = out["tc"], out["fused"]
    print("kernels:", str(a["kernel"]), "vs", str(b["kernel"]))
    scale = np.abs(b["score"]).max()
    print("max |score difference| / scale = %.3g" % (np.abs(a["score"] - b["score"]).max() / scale))
    diff = np.flatnonzero(a["member"] != b["member"])
    print("vertices: %d, members %d / %d, memberships differing: %d" % (a["member"].shape[0], a["member"].sum(), b["member"].sum(), diff.shape[0]))
    if diff.shape[0]:
        pb, _ = make(seed, n_graphs)
        g_of = np.searchsorted(pb.graph_ptr, diff, side="right") - 1
        print("graphs involved:", np.unique(g_of)[:20])
