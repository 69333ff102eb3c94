"""Where one `DQNAgent.solve_mwis(adj, wts)` call spends its time (measurement aid): the Python wrapper, the pointer
tables, the native call, and the device time of the call (CUDA events on the context's stream)."""
import ctypes as C
import os
import sys
import time

import numpy as np
import scipy.sparse as sp

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from distgcn_b200 import engine as E  # noqa: E402
from distgcn_b200.batch import GraphTables  # noqa: E402
from distgcn_b200.mwis_dqn_call import DQNAgent  # noqa: E402
from distgcn_b200.runtime_config import make_flags  # noqa: E402
from tests import util  # noqa: E402


def med(f, n=200):
    ts = []
    for _ in range(n):
        t0 = time.perf_counter()
        f()
        ts.append(time.perf_counter() - t0)
    return 1e6 * float(np.median(ts))


def main():
    for fam, ck, nl in (("ba", "is4sat_l20_c32", 20), ("er", "is4sat_l1", 1)):
        pb, w, _ = util.full_set(fam)
        agent = DQNAgent(1, 5000, flags=make_flags(feature_size=1, hidden1=32, num_layer=nl, diver_num=1, max_degree=1))
        agent.load(util.ckpt_dir(ck))
        agent.check_values = False
        g = 17
        a = sp.csc_matrix(pb.graph_adj(g))
        wg = np.ascontiguousarray(w[pb.graph_ptr[g]:pb.graph_ptr[g + 1]])
        ctx, model = agent.ctx, agent.model.compile(agent.ctx)
        for _ in range(20):
            agent.solve_mwis(a, wg)
        t_all = med(lambda: agent.solve_mwis(a, wg))
        t_engine = med(lambda: E.solve_graphs_host(ctx, model, [a], wg))
        t_tables = med(lambda: GraphTables([a], False))
        t = GraphTables([a], False)
        member = np.empty(t.n_nodes, np.uint8)
        total = np.empty(1, np.float64)
        lib = ctx._lib

        def raw():
            E.check(lib.dg_solve_graphs_host(ctx.handle, model.handle, 1, t.indptr, t.indices, None, t.n_rows_raw, None,
                                             wg.ctypes.data, 0, 1, member.ctypes.data, total.ctypes.data, 1))
        t_raw = med(raw)

        def enqueue_only():
            E.check(lib.dg_solve_graphs_host(ctx.handle, model.handle, 1, t.indptr, t.indices, None, t.n_rows_raw, None,
                                             wg.ctypes.data, 0, 1, member.ctypes.data, total.ctypes.data, 0))
        ts = []
        for _ in range(100):
            ctx.synchronize()
            t0 = time.perf_counter()
            enqueue_only()
            ts.append(time.perf_counter() - t0)
        ctx.synchronize()
        t_enq = 1e6 * float(np.median(ts))
        ms = C.c_double()
        dev = []
        for _ in range(50):
            ctx.synchronize()
            E.check(lib.dg_timer_start(ctx.handle))
            enqueue_only()
            E.check(lib.dg_timer_stop(ctx.handle, C.byref(ms)))
            dev.append(ms.value * 1e3)
        print("%s (n = %d, nnz = %d): solve_mwis %.0f us | engine.solve_graphs_host %.0f | tables %.1f | native call (wait) %.0f | "
              "native enqueue only %.0f | device time of the call %.0f us" % (ck, a.shape[0], a.nnz, t_all, t_engine, t_tables,
                                                                            t_raw, t_enq, float(np.median(dev))))


if __name__ == "__main__":
    main()
