"""Stress / timing probe for the tensor-core solve kernel: back-to-back launches over shuffled copies of the BA test2
batch, membership compared with the graph-resident CUDA-core kernel (DG_DISABLE_TC=1) on the same inputs.
Usage: python profiles/micro/tc_stress.py [reps] [copies] [workload]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))


def main():
    reps = int(sys.argv[1]) if len(sys.argv) > 1 else 50
    ncopies = int(sys.argv[2]) if len(sys.argv) > 2 else 4
    workload = sys.argv[3] if len(sys.argv) > 3 else "ba500"
    import torch
    import bench
    from distgcn_b200 import engine as E
    pb0, w0, layers, desc = bench.load_host_workload(workload, 0)
    rng = np.random.default_rng(1)
    ctx = E.Context(0)
    model = E.Model(ctx, layers, E.gcn_dqn_acts(len(layers)))
    copies = []
    for r in range(ncopies):
        pb, w, _ = bench.shuffled_copy(pb0, w0, rng) if r else (pb0, w0, None)
        dev = E.DeviceBatch(ctx, pb)
        d_w = torch.from_numpy(w).to("cuda:0")
        d_member = torch.zeros(pb.n_nodes, dtype=torch.uint8, device="cuda:0")
        d_total = torch.zeros(pb.n_graphs, dtype=torch.float64, device="cuda:0")
        copies.append(dict(pb=pb, w=w, dev=dev, d_w=d_w, d_member=d_member, d_total=d_total))

    def run(c):
        E.solve_device(ctx, model, c["dev"], c["d_w"], c["d_member"], predict="mwis", remove_zero_weight=True,
                       total=c["d_total"])

    # reference memberships through the CUDA-core kernel
    os.environ["DG_DISABLE_TC"] = "1"
    refs = []
    for c in copies:
        run(c)
        ctx.synchronize()
        refs.append((c["d_member"].cpu().numpy().copy(), c["d_total"].cpu().numpy().copy()))
    del os.environ["DG_DISABLE_TC"]
    # one synchronised pass: membership equality
    for k, c in enumerate(copies):
        c["d_member"].zero_()
        run(c)
        ctx.synchronize()
        m = c["d_member"].cpu().numpy()
        t = c["d_total"].cpu().numpy()
        diff = int((m != refs[k][0]).sum())
        print("copy %d: %d vertices differ from the CUDA-core kernel, totals max rel diff %.3g"
              % (k, diff, float(np.abs(t - refs[k][1]).max() / max(1e-30, np.abs(refs[k][1]).max()))), flush=True)
    # back-to-back launches
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(reps):
        run(copies[i % ncopies])
    ctx.synchronize()
    dt = time.perf_counter() - t0
    print("%d back-to-back solves: %.3f ms per solve, %.0f graphs/s" % (reps, 1e3 * dt / reps, reps * pb0.n_graphs / dt),
          flush=True)
    for k, c in enumerate(copies):
        m = c["d_member"].cpu().numpy()
        print("copy %d after the loop: %d vertices differ" % (k, int((m != refs[k][0]).sum())), flush=True)


if __name__ == "__main__":
    main()
