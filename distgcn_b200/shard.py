"""Sharding layer: independent graph batches across GPUs with no data-path collective.

The reference is single-process ("Parallelism strategies"); batches of independent
conflict graphs shard trivially: every rank takes a contiguous range of graphs balanced by
``nnz + c * n_nodes``, solves it on its own GPU and the memberships are concatenated on the host.
``torch.distributed`` is used only for the rendezvous and for collecting results (gloo on CPU in the
tests, nccl's process group with CPU tensors via gather_object on the GPU box).

A single graph too large for one GPU would need a row partition with a per-layer halo exchange
(SURVEY.md 8e, config 5); that path is not built yet - see DESIGN.md "Multi-GPU".
"""
from __future__ import annotations

from typing import Callable, List, Optional, Tuple

import numpy as np

from .batch import PackedBatch, partition_by_work

SolveFn = Callable[[PackedBatch, np.ndarray], Tuple[np.ndarray, np.ndarray]]


def shard_ranges(packed: PackedBatch, world_size: int, node_cost: float = 8.0) -> List[Tuple[int, int]]:
    """Graph range [g0, g1) of every rank."""
    return partition_by_work(packed, world_size, node_cost)


def local_shard(packed: PackedBatch, wts: np.ndarray, rank: int, world_size: int):
    """(sub-batch, weights, (g0, g1), (v0, v1)) of this rank."""
    g0, g1 = shard_ranges(packed, world_size)[rank]
    v0, v1 = int(packed.graph_ptr[g0]), int(packed.graph_ptr[g1])
    return packed.slice(g0, g1), np.ascontiguousarray(wts[v0:v1]), (g0, g1), (v0, v1)


class ShardedSolver:
    """Solve a packed batch with `world_size` ranks.

    solve_fn(sub_batch, weights) -> (member uint8 [n], total float64 [g]) is what runs on each rank -
    normally ``DQNAgent.solve_mwis_batch`` bound to the rank's GPU.  ``solve`` returns the full
    (member, total) on rank `dst` (None elsewhere); ``solve_local`` returns only this rank's part.
    """

    def __init__(self, solve_fn: SolveFn, rank: Optional[int] = None, world_size: Optional[int] = None, group=None):
        self.solve_fn = solve_fn
        self.group = group
        if rank is None or world_size is None:
            import torch.distributed as dist
            if dist.is_available() and dist.is_initialized():
                rank, world_size = dist.get_rank(group), dist.get_world_size(group)
            else:
                rank, world_size = 0, 1
        self.rank, self.world_size = int(rank), int(world_size)

    def solve_local(self, packed: PackedBatch, wts: np.ndarray):
        sub, w, gr, vr = local_shard(packed, wts, self.rank, self.world_size)
        if sub.n_graphs == 0:
            return np.zeros(0, np.uint8), np.zeros(0, np.float64), gr, vr
        member, total = self.solve_fn(sub, w)
        return np.asarray(member, dtype=np.uint8), np.asarray(total, dtype=np.float64), gr, vr

    def solve(self, packed: PackedBatch, wts: np.ndarray, dst: int = 0):
        wts = np.asarray(wts, dtype=np.float64).reshape(-1)
        member, total, gr, vr = self.solve_local(packed, wts)
        if self.world_size == 1:
            return member, total
        import torch.distributed as dist
        payload = (gr, vr, member, total)
        gathered = [None] * self.world_size if self.rank == dst else None
        dist.gather_object(payload, gathered, dst=dst, group=self.group)
        if self.rank != dst:
            return None
        full_member = np.zeros(packed.n_nodes, dtype=np.uint8)
        full_total = np.zeros(packed.n_graphs, dtype=np.float64)
        for (g0, g1), (v0, v1), m, t in gathered:
            full_member[v0:v1] = m
            full_total[g0:g1] = t
        return full_member, full_total
