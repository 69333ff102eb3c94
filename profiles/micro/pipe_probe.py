"""Where a HostPipeline step's host time goes (wait for the slot vs submit), per workload and depth."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
import bench
from distgcn_b200 import engine as E
from distgcn_b200.batch import PackedBatch
wl = sys.argv[1] if len(sys.argv) > 1 else "er500"
pb, w, layers, desc = bench.load_host_workload(wl, 0)
def pin(a):
    h = E.pinned_empty(a.shape, a.dtype); h[:] = a; return h
sets = []
NSETS = int(os.environ.get("PIPE_SETS", "8"))
rng = np.random.default_rng(1)
for r in range(NSETS):
    pbr, wr = (pb, w) if (r == 0 or not os.environ.get("PIPE_SHUFFLE")) else bench.shuffled_copy(pb, w, rng)[:2]
    rp_u, c16 = pbr.upper_compact()
    sets.append(dict(pb=PackedBatch(pin(pbr.graph_ptr), pin(pbr.row_ptr), pbr.col_idx), up=(pin(rp_u), pin(c16)), w=pin(wr),
                     m=E.pinned_empty(pbr.n_nodes, np.uint8), t=E.pinned_empty(pbr.n_graphs, np.float64)))
for depth in (1, 2, 4):
    pipe = E.HostPipeline(0, layers, E.gcn_dqn_acts(len(layers)), depth=depth)
    def step(i, acc):
        c = sets[i % NSETS]
        slot = pipe._next
        t0 = time.perf_counter()
        pipe.wait(slot)
        t1 = time.perf_counter()
        pipe.submit(c["pb"], c["w"], c["m"], c["t"], upper=c["up"])
        t2 = time.perf_counter()
        acc[0] += t1 - t0; acc[1] += t2 - t1
    acc = [0.0, 0.0]
    for i in range(20): step(i, acc)
    pipe.wait()
    acc = [0.0, 0.0]; n = 200
    t0 = time.perf_counter()
    for i in range(n): step(i, acc)
    pipe.wait()
    tot = time.perf_counter() - t0
    print("%s depth %d: %.1f us per step (wait %.1f, submit %.1f), %.2f M graphs/s" % (wl, depth, 1e6 * tot / n, 1e6 * acc[0] / n, 1e6 * acc[1] / n, pb.n_graphs * n / tot / 1e6))
    pipe.close()
