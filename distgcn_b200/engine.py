"""Thin object layer over the C-ABI (include/distgcn_b200.h): Context, Model, DeviceBatch and the
operators.  numpy arrays travel as DG_MEM_HOST arguments; torch CUDA tensors (anything with
``data_ptr()`` and ``is_cuda``) travel zero-copy as DG_MEM_DEVICE arguments.  All arithmetic happens
in the CUDA library; nothing here computes on the CPU.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import List, Optional, Sequence

import numpy as np

from . import _lib
from ._lib import (ACT_IDENTITY, ACT_LEAKY_RELU, ACT_RELU, HEAD_LINEAR, HEAD_PAIR_SOFTMAX, MEM_DEVICE, MEM_HOST,
                   PREDICT_MIS, PREDICT_MWIS, check)
from .batch import GraphTables, PackedBatch

LEAKY_ALPHA = 0.2  # tf.nn.leaky_relu default, the alpha of every shipped checkpoint


def _is_device_tensor(x) -> bool:
    return hasattr(x, "data_ptr") and bool(getattr(x, "is_cuda", False))


def _np(x, dtype) -> np.ndarray:
    return np.ascontiguousarray(np.asarray(x), dtype=dtype)


def _ptr(a):
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        return a.ctypes.data_as(C.c_void_p)
    return C.c_void_p(a.data_ptr())


def predict_code(predict) -> int:
    if predict in (PREDICT_MWIS, "mwis"):
        return PREDICT_MWIS
    if predict in (PREDICT_MIS, "mis"):
        return PREDICT_MIS
    raise ValueError("predict must be 'mwis' or 'mis', got %r" % (predict,))


_LIVE_CONTEXTS = None   # weak set of live contexts (reload_env)


def reload_env() -> None:
    """Make every live context re-read the library's DG_* environment options (they are read once, when a context is
    created; tests and bench.py call this after changing os.environ)."""
    for ctx in list(_LIVE_CONTEXTS or ()):
        ctx.reload_env()


class Context:
    """One device + one stream + reusable scratch.  Replaces the reference's module-level
    ``tf.compat.v1.Session`` (mwis_dqn_call.py:336-344)."""

    def __init__(self, device: int = 0, stream: Optional[int] = None):
        global _LIVE_CONTEXTS
        self._lib = _lib.load()
        h = C.c_void_p()
        check(self._lib.dg_context_create(int(device), C.c_void_p(stream) if stream else None, C.byref(h)))
        self._h = h
        self.device = int(device)
        if _LIVE_CONTEXTS is None:
            import weakref
            _LIVE_CONTEXTS = weakref.WeakSet()
        _LIVE_CONTEXTS.add(self)

    def reload_env(self) -> None:
        if self._h is not None:
            check(self._lib.dg_context_reload_env(self._h))

    @property
    def handle(self):
        if self._h is None:
            raise RuntimeError("context is closed")
        return self._h

    def synchronize(self) -> None:
        check(self._lib.dg_context_synchronize(self.handle))

    @property
    def launch_count(self) -> int:
        return int(self._lib.dg_context_launch_count(self.handle))

    @property
    def last_kernel(self) -> str:
        """Kernel that carried the most recent solve on this context (which of the three GPU paths ran)."""
        return (self._lib.dg_context_last_kernel(self.handle) or b"").decode()

    def close(self) -> None:
        if self._h is not None:
            self._lib.dg_context_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Model:
    """GraphConvolution stack on the device.  ``layers`` is a sequence of objects with ``.weights``
    (list of [c_in, c_out] float32 arrays, one per support) and ``.bias`` (array or None), e.g. what
    ``distgcn_b200.ckpt.load_gcn_weights`` returns."""

    def __init__(self, ctx: Context, layers: Sequence, acts: Sequence[int], alpha: float = LEAKY_ALPHA,
                 head: int = HEAD_LINEAR):
        self.ctx = ctx
        self._lib = ctx._lib
        n_layers = len(layers)
        if n_layers == 0 or len(acts) != n_layers:
            raise ValueError("need one activation per layer")
        n_sup = len(layers[0].weights)
        keep: List[np.ndarray] = []
        wptrs = (C.c_void_p * (n_layers * n_sup))()
        bptrs = (C.c_void_p * n_layers)()
        c_in = np.zeros(n_layers, dtype=np.int32)
        c_out = np.zeros(n_layers, dtype=np.int32)
        for l, lw in enumerate(layers):
            if len(lw.weights) != n_sup:
                raise ValueError("layer %d has %d weight matrices, expected %d" % (l, len(lw.weights), n_sup))
            c_in[l], c_out[l] = lw.weights[0].shape
            for k, w in enumerate(lw.weights):
                w = _np(w, np.float32)
                if w.shape != (c_in[l], c_out[l]):
                    raise ValueError("layer %d support %d: shape %s" % (l, k, w.shape))
                keep.append(w)
                wptrs[l * n_sup + k] = w.ctypes.data_as(C.c_void_p)
            if lw.bias is not None:
                b = _np(lw.bias, np.float32).reshape(-1)
                if b.shape[0] != c_out[l]:
                    raise ValueError("layer %d: bias length %d" % (l, b.shape[0]))
                keep.append(b)
                bptrs[l] = b.ctypes.data_as(C.c_void_p)
            else:
                bptrs[l] = None
        acts_a = np.asarray(acts, dtype=np.int32)
        h = C.c_void_p()
        check(self._lib.dg_model_create(ctx.handle, n_layers, n_sup, _ptr(c_in), _ptr(c_out),
                                        C.cast(wptrs, C.c_void_p), C.cast(bptrs, C.c_void_p), _ptr(acts_a),
                                        C.c_float(alpha), int(head), C.byref(h)))
        self._h = h
        self.n_layers = n_layers
        self.c_in = [int(x) for x in c_in]
        self.c_out = [int(x) for x in c_out]
        self.acts = [int(a) for a in acts]
        self.head = int(head)
        self.alpha = float(alpha)

    @property
    def handle(self):
        if self._h is None:
            raise RuntimeError("model is closed")
        return self._h

    @property
    def out_width(self) -> int:
        return self.c_out[-1]

    @property
    def feature_size(self) -> int:
        return self.c_in[0]

    def close(self) -> None:
        if self._h is not None and self.ctx._h is not None:
            self._lib.dg_model_destroy(self._h)
        self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def gcn_dqn_acts(n_layers: int) -> List[int]:
    """GCN_DQN / GCN_DEEP_DIVER: leaky-ReLU on all but the last layer (gcn/models.py:536-573,411-434)."""
    return [ACT_LEAKY_RELU] * (n_layers - 1) + [ACT_IDENTITY]


def gcn2_dqn_acts(n_layers: int, act: int = ACT_LEAKY_RELU) -> List[int]:
    """GCN2_DQN: the configured activation on every layer (gcn/models.py:670-708)."""
    return [act] * n_layers


class DeviceBatch:
    """A packed batch resident on the device.  Replaces the ``support`` half of ``makestate``
    (mwis_dqn_call.py:136): only the pattern and fp32(deg^-1/2) are stored."""

    def __init__(self, ctx: Context, packed: Optional[PackedBatch] = None, *, graph_ptr=None, row_ptr=None,
                 col_idx=None):
        self.ctx = ctx
        self._lib = ctx._lib
        self._keepalive = None
        h = C.c_void_p()
        if packed is not None:
            gp, rp, ci = (_np(packed.graph_ptr, np.int32), _np(packed.row_ptr, np.int32), _np(packed.col_idx, np.int32))
            mem = MEM_HOST
            n_graphs, n_nodes, nnz = gp.shape[0] - 1, rp.shape[0] - 1, ci.shape[0]
        else:
            if not all(_is_device_tensor(t) for t in (graph_ptr, row_ptr, col_idx)):
                raise TypeError("device CSR arrays must be int32 CUDA tensors")
            gp, rp, ci = graph_ptr, row_ptr, col_idx
            mem = MEM_DEVICE
            n_graphs, n_nodes, nnz = gp.numel() - 1, rp.numel() - 1, ci.numel()
            self._keepalive = (gp, rp, ci)
        check(self._lib.dg_batch_create(ctx.handle, n_graphs, n_nodes, nnz, _ptr(gp), _ptr(rp), _ptr(ci), mem,
                                        C.byref(h)))
        self._h = h
        self.n_graphs, self.n_nodes, self.nnz = int(n_graphs), int(n_nodes), int(nnz)

    @property
    def handle(self):
        if self._h is None:
            raise RuntimeError("batch is closed")
        return self._h

    def set_keep(self, keep) -> None:
        if keep is None:
            check(self._lib.dg_batch_set_keep(self.handle, None, MEM_HOST))
        elif _is_device_tensor(keep):
            check(self._lib.dg_batch_set_keep(self.handle, _ptr(keep), MEM_DEVICE))
        else:
            k = _np(keep, np.uint8)
            if k.shape[0] != self.n_nodes:
                raise ValueError("keep mask length %d != n_nodes %d" % (k.shape[0], self.n_nodes))
            check(self._lib.dg_batch_set_keep(self.handle, _ptr(k), MEM_HOST))

    def set_keep_from_weights(self, wts) -> None:
        if _is_device_tensor(wts):
            check(self._lib.dg_batch_set_keep_from_weights(self.handle, _ptr(wts), MEM_DEVICE))
        else:
            w = _np(wts, np.float64).reshape(-1)
            if w.shape[0] != self.n_nodes:
                raise ValueError("weights length %d != n_nodes %d" % (w.shape[0], self.n_nodes))
            check(self._lib.dg_batch_set_keep_from_weights(self.handle, _ptr(w), MEM_HOST))

    def set_x0(self, x0) -> None:
        if x0 is None:
            check(self._lib.dg_batch_set_x0(self.handle, None, MEM_HOST))
        elif _is_device_tensor(x0):
            check(self._lib.dg_batch_set_x0(self.handle, _ptr(x0), MEM_DEVICE))
        else:
            x = _np(x0, np.float32).reshape(-1)
            if x.shape[0] != self.n_nodes:
                raise ValueError("x0 length %d != n_nodes %d" % (x.shape[0], self.n_nodes))
            check(self._lib.dg_batch_set_x0(self.handle, _ptr(x), MEM_HOST))

    def close(self) -> None:
        if self._h is not None and self.ctx._h is not None:
            self._lib.dg_batch_destroy(self._h)
        self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


@dataclass
class LgsResult:
    member: np.ndarray                 # uint8 [n_nodes]
    nb_is: Optional[np.ndarray] = None   # uint8 [n_nodes]
    steps: Optional[np.ndarray] = None   # int32 [n_graphs]
    p2p: Optional[np.ndarray] = None     # int64 [n_graphs]
    bst: Optional[np.ndarray] = None     # int64 [n_graphs]
    oh_vec: Optional[np.ndarray] = None  # float64 [n_nodes]


@dataclass
class SolveResult:
    member: np.ndarray
    score: Optional[np.ndarray] = None
    util: Optional[np.ndarray] = None
    total: Optional[np.ndarray] = None
    steps: Optional[np.ndarray] = None


# ----------------------------------------------------------------------------------------------
# operators (host arrays in / host arrays out; see *_device for the zero-copy forms)
# ----------------------------------------------------------------------------------------------
def graph_convolution(ctx: Context, batch: DeviceBatch, x, w0, w1, bias=None, act: int = ACT_LEAKY_RELU,
                      alpha: float = LEAKY_ALPHA) -> np.ndarray:
    w0 = _np(w0, np.float32)
    w1 = _np(w1, np.float32)
    c_in, c_out = w0.shape
    if w1.shape != w0.shape:
        raise ValueError("support weight shapes differ")
    b = None if bias is None else _np(bias, np.float32).reshape(-1)
    x = _np(x, np.float32)
    if x.shape != (batch.n_nodes, c_in):
        raise ValueError("x has shape %s, expected %s" % (x.shape, (batch.n_nodes, c_in)))
    y = np.empty((batch.n_nodes, c_out), dtype=np.float32)
    check(ctx._lib.dg_graph_convolution(ctx.handle, batch.handle, c_in, c_out, _ptr(w0), _ptr(w1), _ptr(b), int(act),
                                        C.c_float(alpha), _ptr(x), _ptr(y), MEM_HOST))
    return y


def gcn_forward(ctx: Context, model: Model, batch: DeviceBatch) -> np.ndarray:
    out = np.empty((batch.n_nodes, model.out_width), dtype=np.float32)
    check(ctx._lib.dg_gcn_forward(ctx.handle, model.handle, batch.handle, _ptr(out), MEM_HOST))
    return out


def utility(ctx: Context, batch: DeviceBatch, score, wts, predict="mwis") -> np.ndarray:
    s = _np(score, np.float32)
    stride = 1 if s.ndim == 1 else int(s.shape[1])
    w = None if wts is None else _np(wts, np.float64).reshape(-1)
    util = np.empty(batch.n_nodes, dtype=np.float64)
    check(ctx._lib.dg_utility(ctx.handle, batch.handle, _ptr(s), stride, _ptr(w), predict_code(predict), _ptr(util),
                              MEM_HOST))
    return util


def lgs(ctx: Context, batch: DeviceBatch, util, nstep: int = -1, want_nb_is: bool = False, want_steps: bool = True,
        want_stats: bool = False, want_overhead: bool = False) -> LgsResult:
    u = _np(util, np.float64).reshape(-1)
    if u.shape[0] != batch.n_nodes:
        raise ValueError("utility length %d != n_nodes %d" % (u.shape[0], batch.n_nodes))
    n, g = batch.n_nodes, batch.n_graphs
    res = LgsResult(member=np.zeros(n, dtype=np.uint8))
    if want_nb_is:
        res.nb_is = np.zeros(n, dtype=np.uint8)
    if want_steps:
        res.steps = np.zeros(g, dtype=np.int32)
    if want_stats or want_overhead:
        res.p2p = np.zeros(g, dtype=np.int64)
        res.bst = np.zeros(g, dtype=np.int64)
    if want_overhead:
        res.oh_vec = np.zeros(n, dtype=np.float64)
    check(ctx._lib.dg_lgs(ctx.handle, batch.handle, _ptr(u), int(nstep), _ptr(res.member), _ptr(res.nb_is),
                          _ptr(res.steps), _ptr(res.p2p), _ptr(res.bst), _ptr(res.oh_vec), MEM_HOST))
    return res


def dist_greedy(ctx: Context, batch: DeviceBatch, wts, epsilon: float = 0.5, want_steps: bool = True) -> LgsResult:
    """Threshold distributed greedy (heuristics.py:38-74) on every graph of the batch (dg_dist_greedy)."""
    w = _np(wts, np.float64).reshape(-1)
    if w.shape[0] != batch.n_nodes:
        raise ValueError("weights length %d != n_nodes %d" % (w.shape[0], batch.n_nodes))
    res = LgsResult(member=np.zeros(batch.n_nodes, dtype=np.uint8))
    if want_steps:
        res.steps = np.zeros(batch.n_graphs, dtype=np.int32)
    check(ctx._lib.dg_dist_greedy(ctx.handle, batch.handle, _ptr(w), float(epsilon), _ptr(res.member), _ptr(res.steps),
                                  MEM_HOST))
    return res


def member_weight(ctx: Context, batch: DeviceBatch, member, wts) -> np.ndarray:
    m = _np(member, np.uint8)
    w = _np(wts, np.float64).reshape(-1)
    total = np.zeros(batch.n_graphs, dtype=np.float64)
    check(ctx._lib.dg_member_weight(ctx.handle, batch.handle, _ptr(m), _ptr(w), _ptr(total), MEM_HOST))
    return total


def solve(ctx: Context, model: Model, batch: DeviceBatch, wts, predict="mwis", remove_zero_weight: bool = True,
          want_score: bool = False, want_util: bool = False, want_total: bool = True,
          want_steps: bool = False) -> SolveResult:
    """Fused GCN -> utility -> LGS on a resident batch with host weight / result arrays."""
    w = _np(wts, np.float64).reshape(-1)
    if w.shape[0] != batch.n_nodes:
        raise ValueError("weights length %d != n_nodes %d" % (w.shape[0], batch.n_nodes))
    n, g = batch.n_nodes, batch.n_graphs
    res = SolveResult(member=np.zeros(n, dtype=np.uint8))
    if want_score:
        res.score = np.zeros((n, model.out_width), dtype=np.float32)
    if want_util:
        res.util = np.zeros(n, dtype=np.float64)
    if want_total:
        res.total = np.zeros(g, dtype=np.float64)
    if want_steps:
        res.steps = np.zeros(g, dtype=np.int32)
    check(ctx._lib.dg_solve(ctx.handle, model.handle, batch.handle, _ptr(w), predict_code(predict),
                            1 if remove_zero_weight else 0, _ptr(res.member), _ptr(res.score), _ptr(res.util),
                            _ptr(res.total), _ptr(res.steps), MEM_HOST))
    return res


def solve_dit(ctx: Context, model: Model, batch: DeviceBatch, wts, predict="mwis", want_total: bool = True,
              want_steps: bool = False) -> SolveResult:
    """GCN embedded into the greedy iteration (dg_solve_dit; MWISSolver.solve_mwis_dit, mwis_gdpg_call.py:278-318)
    on a resident batch with host weight / result arrays."""
    w = _np(wts, np.float64).reshape(-1)
    if w.shape[0] != batch.n_nodes:
        raise ValueError("weights length %d != n_nodes %d" % (w.shape[0], batch.n_nodes))
    res = SolveResult(member=np.zeros(batch.n_nodes, dtype=np.uint8))
    if want_total:
        res.total = np.zeros(batch.n_graphs, dtype=np.float64)
    if want_steps:
        res.steps = np.zeros(batch.n_graphs, dtype=np.int32)
    check(ctx._lib.dg_solve_dit(ctx.handle, model.handle, batch.handle, _ptr(w), predict_code(predict), _ptr(res.member),
                                _ptr(res.total), _ptr(res.steps), MEM_HOST))
    return res


def spmm_laplacian(ctx: Context, batch: DeviceBatch, z, out=None):
    """``L . z`` with ``L = I - D^-1/2 A D^-1/2`` of the batch: the stand-alone form of the reference's
    ``tf.sparse_tensor_dense_matmul(support[1], pre_sup)`` (gcn/layers.py:206).  ``z``: [n_nodes, width] float32, numpy
    (copied) or a CUDA tensor (zero-copy; ``out`` must then be a CUDA tensor of the same shape)."""
    if _is_device_tensor(z):
        if out is None or not _is_device_tensor(out):
            raise TypeError("device input needs a device output tensor")
        check(ctx._lib.dg_spmm_laplacian(ctx.handle, batch.handle, int(z.shape[1]), _ptr(z), _ptr(out), MEM_DEVICE))
        return out
    zz = _np(z, np.float32)
    if zz.ndim != 2 or zz.shape[0] != batch.n_nodes:
        raise ValueError("z must be [n_nodes, width]")
    y = np.empty_like(zz)
    check(ctx._lib.dg_spmm_laplacian(ctx.handle, batch.handle, int(zz.shape[1]), _ptr(zz), _ptr(y), MEM_HOST))
    return y


def solve_device(ctx: Context, model: Model, batch: DeviceBatch, wts, member, predict="mwis",
                 remove_zero_weight: bool = True, score=None, util=None, total=None, steps=None) -> None:
    """Zero-copy form: every array is a CUDA tensor on the context's device; work is only enqueued."""
    for t in (wts, member, score, util, total, steps):
        if t is not None and not _is_device_tensor(t):
            raise TypeError("solve_device takes CUDA tensors")
    check(ctx._lib.dg_solve(ctx.handle, model.handle, batch.handle, _ptr(wts), predict_code(predict),
                            1 if remove_zero_weight else 0, _ptr(member), _ptr(score), _ptr(util), _ptr(total),
                            _ptr(steps), MEM_DEVICE))


def solve_host(ctx: Context, model: Model, packed: PackedBatch, wts, predict="mwis", remove_zero_weight: bool = True,
               member: Optional[np.ndarray] = None, total: Optional[np.ndarray] = None, wait: bool = True,
               col_local16: Optional[np.ndarray] = None, upper=None):
    """One-shot host CSR in / host membership out (dg_solve_host): H2D, kernels and D2H in one call.
    ``wait=False`` only enqueues (dg_solve_host_async): the arrays must stay alive (and should be pinned,
    see ``pinned_empty``) until ``ctx.synchronize()``.  ``col_local16`` (``PackedBatch.local_columns()``, uint16):
    the compact host format - it crosses PCIe instead of ``packed.col_idx`` (dg_solve_host_compact).  ``upper``
    (``PackedBatch.upper_compact()``: (row_ptr_upper, col_local_upper)): the upper-triangle format - half of that again
    (dg_solve_host_upper)."""
    gp, rp, ci = packed.graph_ptr, packed.row_ptr, packed.col_idx
    if upper is not None:
        rp_u, c_u = upper
        if rp_u.dtype != np.int32 or c_u.dtype != np.uint16 or not (rp_u.flags.c_contiguous and c_u.flags.c_contiguous) \
                or rp_u.shape[0] != packed.n_nodes + 1 or gp.dtype != np.int32:
            raise TypeError("upper = (row_ptr_upper int32 [n_nodes + 1], col_local_upper uint16), both contiguous")
        w = wts if (isinstance(wts, np.ndarray) and wts.dtype == np.float64 and wts.flags.c_contiguous) else _np(wts, np.float64)
        n, g = packed.n_nodes, packed.n_graphs
        if member is None:
            member = np.empty(n, dtype=np.uint8)
        if total is None:
            total = np.empty(g, dtype=np.float64)
        check(ctx._lib.dg_solve_host_upper(ctx.handle, model.handle, g, n, int(c_u.shape[0]), _ptr(gp), _ptr(rp_u), _ptr(c_u), _ptr(w),
                                           predict_code(predict), 1 if remove_zero_weight else 0, _ptr(member), _ptr(total),
                                           1 if wait else 0))
        return member, total
    for a in (gp, rp) + (() if col_local16 is not None else (ci,)):
        if a.dtype != np.int32 or not a.flags.c_contiguous:
            raise TypeError("PackedBatch arrays must be contiguous int32")
    w = wts if (isinstance(wts, np.ndarray) and wts.dtype == np.float64 and wts.flags.c_contiguous) else _np(wts, np.float64)
    n, g = packed.n_nodes, packed.n_graphs
    if member is None:
        member = np.empty(n, dtype=np.uint8)
    if total is None:
        total = np.empty(g, dtype=np.float64)
    if col_local16 is not None:
        if col_local16.dtype != np.uint16 or not col_local16.flags.c_contiguous or col_local16.shape[0] != packed.nnz:
            raise TypeError("col_local16 must be a contiguous uint16 array of nnz entries")
        check(ctx._lib.dg_solve_host_compact(ctx.handle, model.handle, g, n, packed.nnz, _ptr(gp), _ptr(rp), _ptr(col_local16),
                                             _ptr(w), predict_code(predict), 1 if remove_zero_weight else 0, _ptr(member),
                                             _ptr(total), 1 if wait else 0))
        return member, total
    fn = ctx._lib.dg_solve_host if wait else ctx._lib.dg_solve_host_async
    check(fn(ctx.handle, model.handle, g, n, packed.nnz, _ptr(gp), _ptr(rp), _ptr(ci), _ptr(w),
             predict_code(predict), 1 if remove_zero_weight else 0, _ptr(member), _ptr(total)))
    return member, total


def solve_graphs_host(ctx: Context, model: Model, graphs, wts, predict="mwis", remove_zero_weight: bool = True,
                      member: Optional[np.ndarray] = None, total: Optional[np.ndarray] = None, wait: bool = True,
                      check_values: bool = False):
    """``DQNAgent.solve_mwis`` for a LIST of per-graph matrices as the reference holds them (one scipy CSR/CSC matrix
    per graph, mwis_dqn_call.py:198), in one native call (dg_solve_graphs_host): the per-graph arrays are packed by the
    library's host threads into pinned staging, copied, solved and the membership (original vertex ids, graph after
    graph) copied back.  ``graphs``: list of matrices or a ``GraphTables``; ``wts``: one float64 array over all
    vertices, or a list of per-graph arrays.  ``check_values``: treat stored zeros as non-edges (costs a pass over the
    values; adjacency matrices loaded from the reference's .mat files hold only ones)."""
    t = graphs if isinstance(graphs, GraphTables) else GraphTables(graphs, check_values)
    per_graph = None
    keep = None
    w = None
    if isinstance(wts, (list, tuple)):
        from . import _pyingest
        try:     # float64 C-contiguous arrays (what the reference passes): pointer table built in C, no Python loop
            per_graph, lens, keep = _pyingest.pointers(wts, 8)
        except (TypeError, BufferError, ValueError):
            arrs = [_np(a, np.float64).reshape(-1) for a in wts]
            per_graph, lens, keep = _pyingest.pointers(arrs, 8)
        if lens != t.n_rows_raw:   # bytes compare: one int32 length per graph
            raise ValueError("one weight per vertex of every graph is needed")
    else:
        w = wts if (isinstance(wts, np.ndarray) and wts.dtype == np.float64 and wts.flags.c_contiguous) else _np(wts, np.float64)
        w = w.reshape(-1)
        if w.shape[0] != t.n_nodes:
            raise ValueError("weights have %d entries for %d vertices" % (w.shape[0], t.n_nodes))
    if member is None:
        member = np.empty(t.n_nodes, dtype=np.uint8)
    if total is None:
        total = np.empty(t.n_graphs, dtype=np.float64)
    check(ctx._lib.dg_solve_graphs_host(ctx.handle, model.handle, t.n_graphs, t.indptr, t.indices, t.data, t.n_rows_raw,
                                        per_graph, _ptr(w), predict_code(predict), 1 if remove_zero_weight else 0,
                                        _ptr(member), _ptr(total), 1 if wait else 0))
    if not wait:
        return member, total, (t, w, keep)   # what must stay alive until ctx.synchronize()
    return member, total


class HostPipeline:
    """Streams of host batches through ``depth`` contexts used in turn (the ingest side
    becomes the bottleneck once the kernels are fast): while one context's kernels run, the next batch's CSR
    is already crossing PCIe on the other context's stream and the finished membership is on its way back.
    ``submit`` returns the slot it used; results of a slot are valid after ``wait(slot)`` (``submit`` waits for
    the slot's previous batch itself).  Every array passed to ``submit`` must stay alive until then and should
    come from ``pinned_empty``."""

    def __init__(self, device: int, layers: Sequence, acts: Sequence[int], depth: int = 2, alpha: float = LEAKY_ALPHA,
                 head: int = HEAD_LINEAR):
        if depth < 1:
            raise ValueError("depth must be >= 1")
        self.ctxs = [Context(device) for _ in range(depth)]
        self.models = [Model(c, layers, acts, alpha, head) for c in self.ctxs]
        self._busy = [None] * depth
        self._next = 0

    @property
    def launch_count(self) -> int:
        return sum(c.launch_count for c in self.ctxs)

    def submit(self, packed: PackedBatch, wts, member: np.ndarray, total: Optional[np.ndarray] = None,
               predict="mwis", remove_zero_weight: bool = True, col_local16: Optional[np.ndarray] = None, upper=None) -> int:
        slot = self._next
        self._next = (slot + 1) % len(self.ctxs)
        self.wait(slot)
        if not (isinstance(wts, np.ndarray) and wts.dtype == np.float64 and wts.flags.c_contiguous):
            wts = _np(wts, np.float64)   # the converted copy is what the copy engine reads: it must outlive the call
        solve_host(self.ctxs[slot], self.models[slot], packed, wts, predict, remove_zero_weight, member, total,
                   wait=False, col_local16=col_local16, upper=upper)
        self._busy[slot] = (packed, wts, member, total, col_local16, upper)  # keep the arrays alive
        return slot

    def submit_graphs(self, graphs, wts, member: np.ndarray, total: Optional[np.ndarray] = None, predict="mwis",
                      remove_zero_weight: bool = True, check_values: bool = False, slot: Optional[int] = None) -> int:
        """The same for the reference's native input: a list of per-graph scipy matrices (or a GraphTables) and their
        weights; packing happens inside the call, in the library (dg_solve_graphs_host).  ``slot``: use this context
        instead of the next one in turn - one producer THREAD per slot may then submit concurrently (a context belongs to
        one host thread at a time; the native call releases the GIL, so one thread's pointer-table walk overlaps the
        other's packing)."""
        if slot is None:
            slot = self._next
            self._next = (slot + 1) % len(self.ctxs)
        self.wait(slot)
        _, _, alive = solve_graphs_host(self.ctxs[slot], self.models[slot], graphs, wts, predict, remove_zero_weight,
                                        member, total, wait=False, check_values=check_values)
        self._busy[slot] = (alive, member, total)
        return slot

    def wait(self, slot: Optional[int] = None) -> None:
        for s in (range(len(self.ctxs)) if slot is None else (slot,)):
            if self._busy[s] is not None:
                self._busy[s] = None
                self.ctxs[s].synchronize()

    def close(self) -> None:
        try:
            self.wait()
        finally:
            for m in self.models:
                m.close()
            for c in self.ctxs:
                c.close()


def pinned_empty(shape, dtype) -> np.ndarray:
    """numpy array backed by page-locked memory from dg_host_alloc (freed when the array dies)."""
    lib = _lib.load()
    dt = np.dtype(dtype)
    count = int(np.prod(shape)) if np.ndim(shape) else int(shape)
    nbytes = max(count * dt.itemsize, 1)
    p = lib.dg_host_alloc(nbytes)
    if not p:
        raise MemoryError("dg_host_alloc(%d) failed" % nbytes)

    class _Owner:
        def __init__(self, ptr):
            self.ptr = ptr

        def __del__(self):
            try:
                lib.dg_host_free(C.c_void_p(self.ptr))
            except Exception:
                pass

    owner = _Owner(p)
    buf = (C.c_char * nbytes).from_address(p)
    arr = np.frombuffer(buf, dtype=dt, count=count).reshape(shape)
    # views inherit the owner through __array_finalize__, so the pages live as long as any view
    arr = arr.view(_PinnedArray)
    arr._dg_owner = owner
    arr._dg_buf = buf
    return arr


class _PinnedArray(np.ndarray):
    _dg_owner = None
    _dg_buf = None

    def __array_finalize__(self, obj):
        if obj is not None:
            self._dg_owner = getattr(obj, "_dg_owner", None)
            self._dg_buf = getattr(obj, "_dg_buf", None)

