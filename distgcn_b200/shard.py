"""Sharding layer: independent graph batches across GPUs with no data-path collective.

The reference is single-process ("Parallelism strategies"); batches of independent
conflict graphs shard trivially: every rank takes a contiguous range of graphs balanced by
``nnz + c * n_nodes``, solves it on its own GPU and the memberships are concatenated on the host.
``torch.distributed`` is used only for the rendezvous and for collecting results (gloo on CPU in the
tests, nccl's process group with CPU tensors via gather_object on the GPU box).

A single graph too large for one GPU is row-partitioned with a per-layer halo exchange (SURVEY.md 8e,
config 5): ``RowPartitionedSolver`` below.
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


# ======================================================================================================
# One giant graph, row-partitioned (BASELINE config 5)
# ======================================================================================================
def row_slices(n_global: int, world_size: int) -> Tuple[int, int]:
    """(slice rows, padded global size): equal slices, multiples of 32 so bitmap words never straddle ranks."""
    per = -(-n_global // world_size)
    per = (per + 31) // 32 * 32
    return per, per * world_size


def slice_csr(indptr: np.ndarray, indices: np.ndarray, n_global: int, rank: int, world_size: int):
    """Local CSR (row_ptr starting at 0, GLOBAL column ids) of `rank`'s slice of a global CSR; rows past
    n_global are empty."""
    per, _ = row_slices(n_global, world_size)
    r0 = rank * per
    r1 = min(n_global, r0 + per)
    rp = np.zeros(per + 1, dtype=np.int64)
    if r1 > r0:
        seg = indptr[r0:r1 + 1].astype(np.int64)
        rp[: r1 - r0 + 1] = seg - seg[0]
        rp[r1 - r0 + 1:] = rp[r1 - r0]
        ci = indices[int(seg[0]):int(seg[-1])]
    else:
        ci = indices[:0]
    return rp.astype(np.int32), np.ascontiguousarray(ci, dtype=np.int32)


class RowPartitionedSolver:
    """GCN-scored local greedy MWIS of ONE graph whose rows are spread over `world_size` GPUs.

    Every rank holds rows [rank*per, (rank+1)*per) as a local CSR with global column ids and global-sized
    per-vertex arrays; each kernel writes the rank's rows and the rows are all-gathered (NCCL over NVLink)
    before the next kernel reads neighbours.  What crosses the links per solve: keep bytes, y (fp32), for each
    hidden layer the feature rows it reads, zs (fp32), utilities (fp64), and per greedy round two bitmaps
    of n/8 bytes plus one 8-byte count.  A model without hidden layers (c64 l2, config 5) exchanges scalars
    only - the rank-1 first layer and project-first last layer of DESIGN.md section 2.
    """

    def __init__(self, model, n_global: int, row_ptr_local: np.ndarray, col_idx_global: np.ndarray, rank: int = 0,
                 world_size: int = 1, group=None, device: Optional[int] = None, exchange: str = "nccl"):
        import ctypes as C

        import torch

        from . import _lib, engine
        self.torch, self.C, self.check = torch, C, _lib.check
        self.rank, self.world, self.group = int(rank), int(world_size), group
        self.n_global = int(n_global)
        self.per, self.n_pad = row_slices(self.n_global, self.world)
        self.row0 = self.rank * self.per
        self.device = torch.device("cuda", torch.cuda.current_device() if device is None else device)
        # One dedicated (non-default) stream for everything: the library enqueues its kernels on it and the
        # torch ops / NCCL collectives of solve() run under torch.cuda.stream(self.stream), so kernels and
        # all-gathers are ordered without host synchronisation.  (torch's default stream has handle 0,
        # which dg_context_create would take as "make a private stream".)
        self.stream = torch.cuda.Stream(device=self.device)
        self.ctx = engine.Context(self.device.index, stream=self.stream.cuda_stream)
        self.lib = self.ctx._lib
        self.model = engine.Model(self.ctx, model.layers_as_weights() if hasattr(model, "layers_as_weights") else model[0],
                                  [l.act_code for l in model.layers] if hasattr(model, "layers") else model[1])
        if self.model.n_layers < 2 or self.model.out_width != 1:
            raise NotImplementedError("row-partitioned path: >= 2 layers and one output column")
        if hasattr(row_ptr_local, "is_cuda"):   # int32 CUDA tensors: used in place
            if not (row_ptr_local.is_cuda and col_idx_global.is_cuda and row_ptr_local.dtype == torch.int32
                    and col_idx_global.dtype == torch.int32):
                raise TypeError("device CSR slices must be int32 CUDA tensors")
            self.d_rp, self.d_ci = row_ptr_local.contiguous(), col_idx_global.contiguous()
            torch.cuda.synchronize(self.device)
            n_rp, nnz_local = int(self.d_rp.numel()), int(self.d_ci.numel())
        else:
            rp = np.ascontiguousarray(row_ptr_local, dtype=np.int32)
            ci = np.ascontiguousarray(col_idx_global, dtype=np.int32)
            with torch.cuda.stream(self.stream):
                self.d_rp = torch.from_numpy(rp).to(self.device)
                self.d_ci = torch.from_numpy(ci).to(self.device)
            self.stream.synchronize()
            n_rp, nnz_local = int(rp.shape[0]), int(ci.shape[0])
        if n_rp != self.per + 1:
            raise ValueError("row_ptr_local must have %d entries" % (self.per + 1))
        h = C.c_void_p()
        self.check(self.lib.dg_part_create(self.ctx.handle, self.n_pad, self.row0, self.per, nnz_local,
                                           C.c_void_p(self.d_rp.data_ptr()), C.c_void_p(self.d_ci.data_ptr()),
                                           _lib.MEM_DEVICE, C.byref(h)))
        self.part = h
        self.exchanged_bytes = 0
        self.keep_on_device = False   # True: solve() returns CUDA tensors instead of numpy arrays
        self.check_every = 4          # p2p exchange: greedy rounds enqueued per read of the ranks' remaining counts
        if exchange not in ("nccl", "p2p"):
            raise ValueError("exchange must be 'nccl' or 'p2p'")
        self.exchange = exchange if self.world > 1 else "nccl"
        self.arena = None
        if self.exchange == "p2p":
            self._setup_arena()

    # ---- peer arenas: the exchange fused into the kernels (include/distgcn_b200.h, "peer arenas") ----------
    def _setup_arena(self):
        import torch.distributed as dist
        C, lib = self.C, self.lib
        L = self.model.n_layers
        cp = max([int(lib.dg_model_padded_width(self.model.handle, l)) for l in range(1, L - 1)] or [0])
        n_pad = self.n_pad
        off, lay = 0, {}

        def take(name, nbytes):
            nonlocal off
            lay[name] = off
            off += (nbytes + 255) // 256 * 256

        take("keep", n_pad)
        take("y", 4 * n_pad)
        take("pair2", 8 * n_pad)
        take("util", 8 * n_pad)
        take("remain", n_pad // 8)
        take("joined", n_pad // 8)
        if L > 2:
            take("dinv", 4 * n_pad)
            take("pair", 8 * n_pad)
        if L > 3:   # rows of hidden layers that a following hidden layer gathers
            take("hid0", 4 * n_pad * cp)
            take("hid1", 4 * n_pad * cp)
        take("flags", 4 * 8)
        take("counts", 8 * 8)
        self.arena_bytes, self.lay = off, lay
        base = C.c_void_p()
        handle = (C.c_uint8 * 64)()
        self.check(lib.dg_peer_alloc(self.ctx.handle, C.c_uint64(off), C.byref(base), handle))
        self.arena = int(base.value)
        handles = [None] * self.world
        dist.all_gather_object(handles, bytes(handle), group=self.group)
        bases = (C.c_void_p * self.world)()
        self._opened = []
        for r in range(self.world):
            if r == self.rank:
                bases[r] = self.arena
            else:
                q = C.c_void_p()
                hb = (C.c_uint8 * 64).from_buffer_copy(handles[r])
                self.check(lib.dg_peer_open(self.ctx.handle, hb, C.byref(q)))
                bases[r] = q.value
                self._opened.append(int(q.value))
        self.check(lib.dg_part_set_peers(self.part, self.world, self.rank, bases, C.c_uint64(off),
                                         C.c_uint64(lay["flags"]), C.c_uint64(lay["counts"])))
        dist.barrier(group=self.group)   # every rank has mapped every arena before anyone stores into one

    def _a(self, name):
        return self.C.c_void_p(self.arena + self.lay[name])

    def _solve_p2p(self, wts_local, predict, remove_zero_weight):
        torch, lib, part, m, C = self.torch, self.lib, self.part, self.model, self.C
        from . import engine
        dev, n_pad, per, r0 = self.device, self.n_pad, self.per, self.row0
        wts = torch.zeros(n_pad, dtype=torch.float64, device=dev)
        if hasattr(wts_local, "is_cuda"):
            wts[r0:r0 + wts_local.numel()] = wts_local.to(dev, torch.float64).reshape(-1)
        else:
            w_in = np.asarray(wts_local, dtype=np.float64).reshape(-1)
            wts[r0:r0 + w_in.shape[0]] = torch.from_numpy(w_in).to(dev)
        F, L = m.feature_size, m.n_layers
        barrier = lambda cnt=None: self.check(lib.dg_part_barrier(part, cnt))  # noqa: E731
        barrier()   # the previous solve has drained on every rank
        self.check(lib.dg_part_keep(part, self._p(wts), 1 if remove_zero_weight else 0, self.n_global, self._a("keep")))
        barrier()
        dinv_t = None
        if L > 2:
            dinv = self._a("dinv")
        else:
            dinv_t = torch.empty(n_pad, dtype=torch.float32, device=dev)
            dinv = self._p(dinv_t)
        self.check(lib.dg_part_prepare(part, F, self._a("keep"), None, dinv, self._a("y")))
        barrier()
        if L > 2:
            pair = self._a("pair")
        else:
            pair_t = torch.empty(n_pad, 2, dtype=torch.float32, device=dev)
            pair = self._p(pair_t)
        self.check(lib.dg_part_first(part, F, dinv, self._a("y"), self._a("keep"), None, pair))
        if L == 2:
            self.check(lib.dg_part_project(part, m.handle, dinv, pair, self._a("pair2")))
        else:
            barrier()   # (x0, s) pairs and dinv of the neighbours
            hin, keepalive = None, []
            for layer in range(1, L - 1):
                cp = int(lib.dg_model_padded_width(m.handle, layer))
                if layer < L - 2:
                    hout = self._a("hid%d" % (layer % 2))   # gathered by the next hidden layer: pushed to the peers
                else:
                    t = torch.empty(n_pad, cp, dtype=torch.float32, device=dev)   # only this rank's rows are read
                    keepalive.append(t)
                    hout = self._p(t)
                self.check(lib.dg_part_layer(part, m.handle, layer, dinv, pair, hin, hout))
                if layer < L - 2:
                    barrier()
                hin = hout
            self.check(lib.dg_part_tail(part, m.handle, dinv, hin, self._a("pair2")))
        barrier()   # zs
        score = torch.zeros(n_pad, dtype=torch.float32, device=dev)
        self.check(lib.dg_part_last(part, m.handle, dinv, self._a("pair2"), self._a("keep"), self._p(wts),
                                    engine.predict_code(predict), self._p(score), self._a("util")))
        barrier()   # utilities
        member = torch.zeros(n_pad, dtype=torch.uint8, device=dev)
        count = torch.zeros(1, dtype=torch.int64, device=dev)
        self.check(lib.dg_part_lgs_init(part, self._a("keep"), self._a("remain"), self._p(member), self._p(count)))
        barrier(self._p(count))
        # the rounds are driven natively: `check_every` rounds enqueued back to back per read of the ranks' remaining counts
        # (no Python, no torch kernel and no host synchronisation per round)
        r_out = C.c_int32(0)
        self.check(lib.dg_part_lgs_run(part, self._a("util"), self._a("remain"), self._a("joined"), self._p(member),
                                       self._p(count), int(self.check_every), C.byref(r_out)))
        rounds = int(r_out.value)
        self.ctx.synchronize()   # also surfaces a barrier time-out
        # bytes this rank stored into its peers' arenas (what an all-gather would have moved)
        per_solve = per * (1 + 4 + 4 + 8) + (rounds * 2 + 1) * (per // 8)
        if L > 2:
            per_solve += per * (4 + 8) + sum(per * 4 * int(lib.dg_model_padded_width(m.handle, l))
                                            for l in range(1, L - 2))
        self.exchanged_bytes += (self.world - 1) * per_solve
        if self.keep_on_device:
            return member[r0:r0 + per], score[r0:r0 + per], rounds
        return (member[r0:r0 + per].cpu().numpy(), score[r0:r0 + per].cpu().numpy(), rounds)

    def _p(self, t):
        return None if t is None else self.C.c_void_p(t.data_ptr())

    def _gather(self, full, elems_per_row=1):
        """Make this rank's rows of `full` (viewed as [n_pad * elems_per_row]) visible everywhere."""
        if self.world == 1:
            return
        import torch.distributed as dist
        flat = full.view(-1)
        chunk = self.per * elems_per_row
        mine = flat[self.rank * chunk:(self.rank + 1) * chunk].clone()
        dist.all_gather_into_tensor(flat, mine, group=self.group)
        self.exchanged_bytes += (self.world - 1) * chunk * full.element_size()

    def solve(self, wts_local: np.ndarray, predict="mwis", remove_zero_weight=True):
        """wts_local: this rank's `per` weights (rows past n_global ignored).  Returns (member_local uint8
        [per], score_local float32 [per], rounds)."""
        with self.torch.cuda.stream(self.stream):
            if self.exchange == "p2p":
                return self._solve_p2p(wts_local, predict, remove_zero_weight)
            return self._solve(wts_local, predict, remove_zero_weight)

    def _solve(self, wts_local, predict, remove_zero_weight):
        torch, lib, part, m = self.torch, self.lib, self.part, self.model
        from . import engine
        dev, n_pad, per, r0 = self.device, self.n_pad, self.per, self.row0
        wts = torch.zeros(n_pad, dtype=torch.float64, device=dev)
        if hasattr(wts_local, "is_cuda"):
            wts[r0:r0 + wts_local.numel()] = wts_local.to(dev, torch.float64).reshape(-1)
        else:
            w_in = np.asarray(wts_local, dtype=np.float64).reshape(-1)
            wts[r0:r0 + w_in.shape[0]] = torch.from_numpy(w_in).to(dev)
        keep = torch.zeros(n_pad, dtype=torch.uint8, device=dev)
        self.check(lib.dg_part_keep(part, self._p(wts), 1 if remove_zero_weight else 0, self.n_global, self._p(keep)))
        self._gather(keep)
        F = m.feature_size
        dinv = torch.zeros(n_pad, dtype=torch.float32, device=dev)
        y = torch.zeros(n_pad, dtype=torch.float32, device=dev)
        pair = torch.zeros(n_pad, 2, dtype=torch.float32, device=dev)
        pair2 = torch.zeros(2, n_pad, dtype=torch.float32, device=dev)   # planar: q plane, zs plane
        self.check(lib.dg_part_prepare(part, F, self._p(keep), None, self._p(dinv), self._p(y)))
        self._gather(y)
        self.check(lib.dg_part_first(part, F, self._p(dinv), self._p(y), self._p(keep), None, self._p(pair)))
        L = m.n_layers
        if L == 2:
            self.check(lib.dg_part_project(part, m.handle, self._p(dinv), self._p(pair), self._p(pair2)))
        else:
            self._gather(dinv)
            self._gather(pair, 2)
            hin = None
            for layer in range(1, L - 1):
                cp = int(lib.dg_model_padded_width(m.handle, layer))
                hout = torch.zeros(n_pad, cp, dtype=torch.float32, device=dev)
                self.check(lib.dg_part_layer(part, m.handle, layer, self._p(dinv), self._p(pair), self._p(hin),
                                             self._p(hout)))
                if layer < L - 2:
                    self._gather(hout, cp)  # the next hidden layer reads neighbours' rows
                hin = hout
            self.check(lib.dg_part_tail(part, m.handle, self._p(dinv), self._p(hin), self._p(pair2)))
        self._gather(pair2[1])   # only zs is read from other ranks' rows
        score = torch.zeros(n_pad, dtype=torch.float32, device=dev)
        util = torch.zeros(n_pad, dtype=torch.float64, device=dev)
        self.check(lib.dg_part_last(part, m.handle, self._p(dinv), self._p(pair2), self._p(keep), self._p(wts),
                                    engine.predict_code(predict), self._p(score), self._p(util)))
        self._gather(util)
        # greedy rounds
        words = n_pad // 32
        remain = torch.zeros(words, dtype=torch.int32, device=dev)
        joined = torch.zeros(words, dtype=torch.int32, device=dev)
        member = torch.zeros(n_pad, dtype=torch.uint8, device=dev)
        count = torch.zeros(1, dtype=torch.int64, device=dev)
        self.check(lib.dg_part_lgs_init(part, self._p(keep), self._p(remain), self._p(member), self._p(count)))
        rounds = 0
        while True:
            if self.world > 1:
                import torch.distributed as dist
                dist.all_reduce(count, group=self.group)
            if int(count.item()) == 0:
                break
            if rounds >= (1 << 20):
                raise RuntimeError("local greedy search did not converge (NaN utilities or self-loops?)")
            self._gather_words(remain)
            self.check(lib.dg_part_lgs_decide(part, self._p(util), self._p(remain), self._p(joined), self._p(member)))
            self._gather_words(joined)
            count.zero_()
            self.check(lib.dg_part_lgs_remove(part, self._p(joined), self._p(remain), self._p(count)))
            rounds += 1
        self.stream.synchronize()
        if self.keep_on_device:
            return member[r0:r0 + per], score[r0:r0 + per], rounds
        return (member[r0:r0 + per].cpu().numpy(), score[r0:r0 + per].cpu().numpy(), rounds)

    def _gather_words(self, words):
        if self.world == 1:
            return
        import torch.distributed as dist
        chunk = self.per // 32
        mine = words[self.rank * chunk:(self.rank + 1) * chunk].clone()
        dist.all_gather_into_tensor(words, mine, group=self.group)
        self.exchanged_bytes += (self.world - 1) * chunk * 4

    def close(self):
        if self.part is not None:
            self.lib.dg_part_destroy(self.part)
            self.part = None
        if self.arena is not None:
            import torch.distributed as dist
            self.ctx.synchronize()
            dist.barrier(group=self.group)   # nobody stores into an arena that is about to go away
            for q in self._opened:
                self.lib.dg_peer_close(self.ctx.handle, self.C.c_void_p(q))
            dist.barrier(group=self.group)
            self.lib.dg_peer_free(self.ctx.handle, self.C.c_void_p(self.arena))
            self.arena = None
        self.model.close()
        self.ctx.close()


class _RawCuda:
    def __init__(self, ptr, shape, typestr):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False),
                                         "version": 2, "strides": None}


def _device_view(ptr: int, shape, typestr: str, device):
    """torch tensor over raw device memory (no copy, no ownership)."""
    import torch
    return torch.as_tensor(_RawCuda(ptr, shape, typestr), device=device)
