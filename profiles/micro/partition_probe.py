"""Row-partitioned solve of ONE synthetic G(n, m = n*deg/2) graph at scale (BASELINE configs[4]), launched with
torchrun, one process per GPU: every rank builds the same graph on its GPU (same seed), keeps its row slice and
solves with both exchange modes; rank 0 also solves the whole graph alone and the memberships are compared on
the device.  DG_PART_N / DG_PART_DEG / DG_PART_CKPT select the instance.  Measurement aid, not a test."""
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from distgcn_b200 import engine as E  # noqa: E402
from distgcn_b200.shard import RowPartitionedSolver, row_slices  # noqa: E402
from profiles.micro.stream_probe import big_er_device  # noqa: E402
from tests import util  # noqa: E402
from tests.test_gpu_partition import _ModelSpec  # noqa: E402


def main():
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    n = int(os.environ.get("DG_PART_N", "50000000"))
    deg = int(os.environ.get("DG_PART_DEG", "16"))
    ck = os.environ.get("DG_PART_CKPT", "is4sat_l2_c64")
    reps = int(os.environ.get("DG_PART_REPS", "3"))
    t0 = time.perf_counter()
    gp, rp, ci = big_er_device(n, deg, 0, dev)
    g = torch.Generator(device=dev)
    g.manual_seed(1)
    w = torch.rand(n, dtype=torch.float64, device=dev, generator=g)
    w[torch.rand(n, device=dev, generator=g) < 0.05] = 0.0
    torch.cuda.synchronize()
    gen_s = time.perf_counter() - t0
    per, n_pad = row_slices(n, world)
    r0, r1 = rank * per, min(n, (rank + 1) * per)
    rp_l = torch.zeros(per + 1, dtype=torch.int32, device=dev)
    e0 = int(rp[r0]) if r1 > r0 else 0
    if r1 > r0:
        rp_l[: r1 - r0 + 1] = rp[r0:r1 + 1] - e0
        rp_l[r1 - r0 + 1:] = rp_l[r1 - r0]
        ci_l = ci[e0:int(rp[r1])].clone()
    else:
        ci_l = ci[:0].clone()
    layers = util.load_layers(ck)
    acts = E.gcn_dqn_acts(len(layers))
    out = {"n": n, "nnz": int(ci.numel()), "world": world, "ckpt": ck, "gen_s": gen_s}
    members = {}
    for ex in (("nccl", "p2p") if world > 1 else ("nccl",)):
        solver = RowPartitionedSolver(_ModelSpec(layers, acts), n, rp_l, ci_l, rank=rank, world_size=world,
                                      exchange=ex)
        solver.keep_on_device = True
        solver.solve(w[r0:r1])
        solver.exchanged_bytes = 0
        torch.cuda.synchronize()
        dist.barrier()
        t0 = time.perf_counter()
        for _ in range(reps):
            member, score, rounds = solver.solve(w[r0:r1])
        torch.cuda.synchronize()
        dt = torch.tensor([(time.perf_counter() - t0) / reps], device=dev, dtype=torch.float64)
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        out[ex] = {"ms_per_solve": 1e3 * float(dt), "rounds": rounds,
                   "exchanged_MB_per_rank_per_solve": solver.exchanged_bytes / reps / 1e6,
                   "nodes_per_s": n / float(dt)}
        full = torch.zeros(n_pad, dtype=torch.uint8, device=dev)
        dist.all_gather_into_tensor(full, member.contiguous())
        members[ex] = full[:n].clone()
        solver.close()
    if rank == 0:
        ctx = E.Context(local)
        model = E.Model(ctx, layers, acts)
        batch = E.DeviceBatch(ctx, graph_ptr=gp, row_ptr=rp, col_idx=ci)
        ref = torch.empty(n, dtype=torch.uint8, device=dev)
        E.solve_device(ctx, model, batch, w, ref)
        ctx.synchronize()
        t0 = time.perf_counter()
        for _ in range(reps):
            E.solve_device(ctx, model, batch, w, ref)
        ctx.synchronize()
        out["single_gpu"] = {"ms_per_solve": 1e3 * (time.perf_counter() - t0) / reps, "members": int(ref.sum())}
        for ex, mfull in members.items():
            out[ex]["membership_equal_to_single_gpu"] = bool(torch.equal(mfull, ref))
            out[ex]["differing_vertices"] = int((mfull != ref).sum())
        print(json.dumps(out))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
