"""Slots/s of the batched multi-channel scheduler (BASELINE configs[2]) on cuda:0.  The per-instance CPU restatement's
timing lives in tests/probe_wireless_cpu.py (only tests/, smoke() and bench.py may run the oracle).  Measurement aid, not a test."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from distgcn_b200 import engine as E  # noqa: E402
from distgcn_b200 import wireless as W  # noqa: E402
from tests import util  # noqa: E402


def main():
    n_net, loads, T = 20, np.round(np.arange(0.1, 1.25, 0.1), 2), 200   # bash/twc_major_wireless_mc_test.sh
    insts = W.make_instances(n_net, loads, n_ch=3, timeslots=T, seed=0)
    out = {"instances": len(insts), "slots": T - 1, "mean_links": float(np.mean([i.nflows for i in insts])),
           "mean_joint_vertices": float(np.mean([i.adj_gK.shape[0] for i in insts])),
           "mean_joint_degree": float(np.mean([i.adj_gK.nnz / i.adj_gK.shape[0] for i in insts]))}
    ctx = E.Context(0)
    for ck in ("is4sat_l1", "is4sat_l20_c32"):
        layers = util.load_layers(ck)
        model = E.Model(ctx, layers, E.gcn_dqn_acts(len(layers)))
        for algo in W.ALGOS:
            if ck != "is4sat_l1" and not algo.startswith("DGCN"):
                continue
            sim = W.BatchedScheduler(ctx, insts, algo, model)
            sim.run(5)
            t0 = time.perf_counter()
            qs = sim.run()
            dt = time.perf_counter() - t0
            n = qs.shape[0]
            out["%s/%s" % (ck, algo)] = {"slots_per_s": n / dt, "instance_slots_per_s": n * len(insts) / dt,
                                         "graphs_per_s": n * sim.graphs_per_slot / dt, "mean_queue": float(qs.mean())}
            sim.close()
        model.close()
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
