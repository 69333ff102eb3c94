"""Where a `solve_mwis_batch(list of scipy matrices, list of weight vectors)` call spends its host time (measurement aid):
pointer tables (C helper), argument normalisation, the native call (pack + enqueue), the wait for the results."""
import os
import sys
import time

import numpy as np
import scipy.sparse as sp

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from distgcn_b200 import _pyingest, engine as E  # noqa: E402
from distgcn_b200.batch import GraphTables  # noqa: E402
from tests import util  # noqa: E402


def main():
    fam = sys.argv[1] if len(sys.argv) > 1 else "ba"
    pb, w, _ = util.full_set(fam)
    layers = util.load_layers("is4sat_l20_c32" if fam == "ba" else "is4sat_l1")
    sets = []
    for r in range(8):
        adjs = [sp.csc_matrix(pb.graph_adj(g)) for g in range(pb.n_graphs)]
        wl = [np.ascontiguousarray(w[pb.graph_ptr[g]:pb.graph_ptr[g + 1]]) for g in range(pb.n_graphs)]
        sets.append((adjs, wl))
    depth = int(os.environ.get("DEPTH", "2"))
    pipe = E.HostPipeline(0, layers, E.gcn_dqn_acts(len(layers)), depth=depth)
    member = [E.pinned_empty(pb.n_nodes, np.uint8) for _ in range(depth)]
    total = [E.pinned_empty(pb.n_graphs, np.float64) for _ in range(depth)]
    for i in range(6):
        pipe.submit_graphs(sets[i % 8][0], sets[i % 8][1], member[i % depth], total[i % depth])
    pipe.wait()
    acc = {"tables": 0.0, "weights": 0.0, "wait_slot": 0.0, "native": 0.0}
    n = 40
    t_all = time.perf_counter()
    for i in range(n):
        adjs, wl = sets[i % 8]
        slot = i % depth
        t0 = time.perf_counter()
        pipe.wait(slot)
        t1 = time.perf_counter()
        t = GraphTables(adjs, False)
        t2 = time.perf_counter()
        per_graph, lens, keep = _pyingest.pointers(wl, 8)
        t3 = time.perf_counter()
        E.check(pipe.ctxs[slot]._lib.dg_solve_graphs_host(
            pipe.ctxs[slot].handle, pipe.models[slot].handle, t.n_graphs, t.indptr, t.indices, t.data, t.n_rows_raw,
            per_graph, None, 0, 1, member[slot].ctypes.data, total[slot].ctypes.data, 0))
        t4 = time.perf_counter()
        pipe._busy[slot] = (t, keep, wl)
        acc["wait_slot"] += t1 - t0
        acc["tables"] += t2 - t1
        acc["weights"] += t3 - t2
        acc["native"] += t4 - t3
    pipe.wait()
    t_all = time.perf_counter() - t_all
    print("per step: %.0f us total; " % (1e6 * t_all / n) + ", ".join("%s %.0f us" % (k, 1e6 * v / n) for k, v in acc.items()))
    # the library call through engine.solve_graphs_host (what submit_graphs does)
    t0 = time.perf_counter()
    for i in range(n):
        pipe.submit_graphs(sets[i % 8][0], sets[i % 8][1], member[i % depth], total[i % depth])
    pipe.wait()
    print("submit_graphs: %.0f us per step" % (1e6 * (time.perf_counter() - t0) / n))
    # device time of one batch on one context (CUDA events on the context's stream), native list input vs compact arrays
    import ctypes as C
    ctx, model = pipe.ctxs[0], pipe.models[0]
    lib = ctx._lib
    ms = C.c_double()
    from distgcn_b200.batch import pack_graphs
    pbk = pack_graphs(sets[0][0], check_values=False)
    c16 = pbk.local_columns()
    hp = {k: E.pinned_empty(a.shape, a.dtype) for k, a in (("gp", pbk.graph_ptr), ("rp", pbk.row_ptr), ("c16", c16), ("w", w))}
    hp["gp"][:], hp["rp"][:], hp["c16"][:], hp["w"][:] = pbk.graph_ptr, pbk.row_ptr, c16, w
    from distgcn_b200.batch import PackedBatch
    for tag in ("compact", "native", "compact", "native"):
        ts = []
        for i in range(6):
            ctx.synchronize()
            t0 = time.perf_counter()
            E.check(lib.dg_timer_start(ctx.handle))
            if tag == "native":
                E.solve_graphs_host(ctx, model, sets[i % 8][0], sets[i % 8][1], member=member[0], total=total[0], wait=False)
            else:
                E.solve_host(ctx, model, PackedBatch(hp["gp"], hp["rp"], pbk.col_idx), hp["w"], member=member[0], total=total[0],
                             wait=False, col_local16=hp["c16"])
            t1 = time.perf_counter()
            E.check(lib.dg_timer_stop(ctx.handle, C.byref(ms)))
            t2 = time.perf_counter()
            ts.append((ms.value * 1e3, 1e6 * (t1 - t0), 1e6 * (t2 - t0)))
        print(tag, "device us / host enqueue us / host total us:", ["%.0f/%.0f/%.0f" % x for x in ts[1:]])
    pipe.close()


if __name__ == "__main__":
    main()
