"""torchrun --nproc-per-node 2 tests/run_partition_2gpu.py : row-partitioned solve on 2 (or more) GPUs, both exchange
modes, against the CPU oracle (float64 scores, oracle.lgs membership on the path's own utilities).  Collected by
pytest through tests/test_gpu_partition.py::test_partitioned_multi_rank."""
import os
import sys
import time

import numpy as np
import scipy.sparse as sp
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests import util  # noqa: E402
from tests.test_gpu_partition import _big_graph, _ModelSpec, oracle_check  # noqa: E402
from distgcn_b200 import engine as E  # noqa: E402
from distgcn_b200.batch import pack_graphs  # noqa: E402
from distgcn_b200.shard import RowPartitionedSolver, row_slices, slice_csr  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
    dist.init_process_group("nccl", device_id=torch.device("cuda", torch.cuda.current_device()))
    rng = np.random.default_rng(11)
    n = int(os.environ.get("DG_PART_N", "200003"))
    a = _big_graph(rng, n, 16)
    w = rng.random(n)
    w[rng.random(n) < 0.05] = 0.0
    ok = True
    cases = [(short, ex) for ex in ("nccl", "p2p") for short in ("is4sat_l2_c64", "is4sat_l3_c16", "is4sat_l20_c32")]
    for short, exchange in cases:
        if short == "is4sat_l20_c32" and n > 400000:
            continue
        layers = util.load_layers(short)
        acts = E.gcn_dqn_acts(len(layers))
        per, n_pad = row_slices(n, world)
        rp, ci = slice_csr(a.indptr, a.indices, n, rank, world)
        solver = RowPartitionedSolver(_ModelSpec(layers, acts), n, rp, ci, rank=rank, world_size=world,
                                      exchange=exchange)
        w_local = w[rank * per:min(n, (rank + 1) * per)]
        if exchange == "p2p":   # greedy rounds enqueued per count read: one at a time and speculatively give the same answer
            solver.check_every = 1
            m1, _, r1 = solver.solve(w_local)
            solver.check_every = 5
            m5, _, r5 = solver.solve(w_local)
            if not (np.array_equal(m1, m5) and r1 == r5):
                print("PART rank %d: check_every 1 / 5 disagree (rounds %d / %d)" % (rank, r1, r5))
                ok = False
            solver.check_every = 4
        solver.solve(w_local)   # warm-up (allocations, NCCL channels)
        solver.exchanged_bytes = 0
        torch.cuda.synchronize()
        dist.barrier()
        t0 = time.perf_counter()
        member, score, rounds = solver.solve(w_local)
        torch.cuda.synchronize()
        ms = 1e3 * (time.perf_counter() - t0)
        gathered = [None] * world
        dist.all_gather_object(gathered, (member, score))
        if rank == 0:
            full_m = np.concatenate([g[0] for g in gathered])[:n]
            full_s = np.concatenate([g[1] for g in gathered])[:n]
            err, same, steps = oracle_check(a, w, layers, full_m, full_s, n)
            print("PART %s %s n=%d world=%d rounds=%d ms=%.2f exchanged_MB=%.1f membership_equals_oracle=%s "
                  "score_err_vs_fp64=%.2e" % (short, exchange, n, world, rounds, ms, solver.exchanged_bytes / 1e6, same, err))
            ok = ok and same and err <= 1e-5 and rounds == steps
        solver.close()
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)   # any rank's failure fails the run
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0:
        print("PARTITION_2GPU_OK" if int(flag.item()) else "PARTITION_2GPU_FAILED")
    sys.exit(0 if int(flag.item()) else 1)


if __name__ == "__main__":
    main()
