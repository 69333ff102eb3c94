"""CPU restatement of the wireless slot loop (oracle/wireless_oracle.py), one instance at a time like the reference script, timed
on a few instances: the CPU figure quoted beside profiles/micro/wireless_probe.py's GPU numbers.  Not collected by pytest."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from distgcn_b200 import wireless as W  # noqa: E402
from oracle import wireless_oracle as WO  # noqa: E402
from tests import util  # noqa: E402


def main():
    import numpy as np
    insts = W.make_instances(20, np.round(np.arange(0.1, 1.25, 0.1), 2), n_ch=3, timeslots=200, seed=0)   # wireless_probe.py's instances
    out = {}
    layers = util.load_layers("is4sat_l1")
    for algo in ("Greedy", "DGCN-LGS"):
        t0 = time.perf_counter()
        k = 0
        for inst in insts[:6]:
            WO.run_instance(inst.adj_list, inst.adj_gK, inst.arrivals, inst.rates, algo, layers, n_slots=40)
            k += 40
        out["cpu_port/%s" % algo] = {"instance_slots_per_s": k / (time.perf_counter() - t0)}
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
