"""Host-side cost of one dg_solve_host_async submit (enqueue only) per workload."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
import bench
from distgcn_b200 import engine as E
for wl in (sys.argv[1:] or ["ba500", "er500"]):
    pb, w, layers, desc = bench.load_host_workload(wl, 0)
    ctx = E.Context(0)
    model = E.Model(ctx, layers, E.gcn_dqn_acts(len(layers)))
    h = {k: E.pinned_empty(a.shape, a.dtype) for k, a in (("gp", pb.graph_ptr), ("rp", pb.row_ptr), ("ci", pb.col_idx), ("w", w))}
    h["gp"][:], h["rp"][:], h["ci"][:], h["w"][:] = pb.graph_ptr, pb.row_ptr, pb.col_idx, w
    from distgcn_b200.batch import PackedBatch
    hp = PackedBatch(h["gp"], h["rp"], h["ci"])
    member = E.pinned_empty(pb.n_nodes, np.uint8); total = E.pinned_empty(pb.n_graphs, np.float64)
    rp_u, c16 = pb.upper_compact()
    hu = (E.pinned_empty(rp_u.shape, rp_u.dtype), E.pinned_empty(c16.shape, c16.dtype))
    hu[0][:], hu[1][:] = rp_u, c16
    for fmt, kw in (("int32 CSR", {}), ("upper", {"upper": hu})):
        for _ in range(5):
            E.solve_host(ctx, model, hp, h["w"], member=member, total=total, wait=False, **kw); ctx.synchronize()
        ts = []
        for _ in range(30):
            t0 = time.perf_counter()
            E.solve_host(ctx, model, hp, h["w"], member=member, total=total, wait=False, **kw)
            t1 = time.perf_counter()
            ctx.synchronize()
            t2 = time.perf_counter()
            ts.append((t1 - t0, t2 - t1))
        a = np.array(ts) * 1e3
        print(wl, fmt, ctx.last_kernel, "submit %.3f ms (min %.3f), then sync %.3f ms" % (a[:, 0].mean(), a[:, 0].min(), a[:, 1].mean()))
