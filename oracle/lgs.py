"""ctypes wrapper around oracle/liblgs_oracle.so (lgs_oracle.c).  TEST INFRASTRUCTURE ONLY.

Return conventions mirror the reference's functions (heuristics.py:116,160,209,263,305), with the
vertex set given as a 0/1 uint8 membership vector instead of a Python set.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from typing import Optional

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liblgs_oracle.so")
_lib = None


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "lgs_oracle.c")
    if force or not os.path.isfile(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s", "liblgs_oracle.so"])
    return _LIB_PATH


def _load():
    global _lib
    if _lib is None:
        build()
        lib = ctypes.CDLL(_LIB_PATH)
        c = ctypes
        lib.lgs_oracle_run.restype = c.c_longlong
        lib.lgs_oracle_run.argtypes = [c.c_int, c.c_void_p, c.c_void_p, c.c_void_p, c.c_void_p, c.c_int,
                                       c.c_longlong, c.c_void_p, c.c_void_p, c.c_void_p, c.c_void_p,
                                       c.c_void_p, c.c_void_p]
        lib.lgs_oracle_run_batch.restype = c.c_longlong
        lib.lgs_oracle_run_batch.argtypes = [c.c_int, c.c_void_p, c.c_void_p, c.c_void_p, c.c_void_p,
                                             c.c_void_p, c.c_int, c.c_longlong, c.c_void_p, c.c_void_p,
                                             c.c_void_p, c.c_void_p, c.c_void_p, c.c_void_p]
        lib.greedy_oracle_run.restype = None
        lib.greedy_oracle_run.argtypes = [c.c_int, c.c_void_p, c.c_void_p, c.c_void_p, c.c_void_p]
        lib.dgs_oracle_run.restype = c.c_longlong
        lib.dgs_oracle_run.argtypes = [c.c_int, c.c_void_p, c.c_void_p, c.c_void_p, c.c_void_p, c.c_double,
                                       c.c_longlong, c.c_void_p, c.c_void_p]
        _lib = lib
    return _lib


def _ptr(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


class LgsResult:
    __slots__ = ("member", "nb_is", "steps", "p2p", "bst", "oh_vec")

    def __init__(self, member, nb_is, steps, p2p, bst, oh_vec):
        self.member, self.nb_is, self.steps, self.p2p, self.bst, self.oh_vec = member, nb_is, steps, p2p, bst, oh_vec


def run(row_ptr, col_idx, wts, init_remain=None, nstep: int = -1, max_rounds: int = -1) -> LgsResult:
    """One graph.  row_ptr/col_idx: CSR pattern; wts: any array-like, flattened to fp64
    (heuristics.py:84)."""
    lib = _load()
    rp = np.ascontiguousarray(row_ptr, dtype=np.int64)
    ci = np.ascontiguousarray(col_idx, dtype=np.int32)
    w = np.ascontiguousarray(np.asarray(wts, dtype=np.float64).reshape(-1))
    n = rp.shape[0] - 1
    assert w.shape[0] == n
    ir = None if init_remain is None else np.ascontiguousarray(init_remain, dtype=np.uint8)
    member = np.zeros(n, dtype=np.uint8)
    nb_is = np.zeros(n, dtype=np.uint8)
    oh = np.zeros(n, dtype=np.float64)
    steps = np.zeros(1, dtype=np.int64)
    p2p = np.zeros(1, dtype=np.int64)
    bst = np.zeros(1, dtype=np.int64)
    r = lib.lgs_oracle_run(n, _ptr(rp), _ptr(ci), _ptr(w), _ptr(ir), int(nstep), int(max_rounds),
                           _ptr(member), _ptr(nb_is), _ptr(steps), _ptr(p2p), _ptr(bst), _ptr(oh))
    if r < 0:
        raise RuntimeError("lgs_oracle_run failed with %d (-2 = did not converge)" % r)
    return LgsResult(member, nb_is, int(steps[0]), int(p2p[0]), int(bst[0]), oh)


def run_batch(graph_ptr, row_ptr, col_idx, wts, init_remain=None, nstep: int = -1, max_rounds: int = -1) -> LgsResult:
    """Packed batch of independent graphs; per-graph steps/p2p/bst arrays."""
    lib = _load()
    gp = np.ascontiguousarray(graph_ptr, dtype=np.int64)
    rp = np.ascontiguousarray(row_ptr, dtype=np.int64)
    ci = np.ascontiguousarray(col_idx, dtype=np.int32)
    w = np.ascontiguousarray(np.asarray(wts, dtype=np.float64).reshape(-1))
    g = gp.shape[0] - 1
    n = rp.shape[0] - 1
    ir = None if init_remain is None else np.ascontiguousarray(init_remain, dtype=np.uint8)
    member = np.zeros(n, dtype=np.uint8)
    nb_is = np.zeros(n, dtype=np.uint8)
    oh = np.zeros(n, dtype=np.float64)
    steps = np.zeros(g, dtype=np.int64)
    p2p = np.zeros(g, dtype=np.int64)
    bst = np.zeros(g, dtype=np.int64)
    r = lib.lgs_oracle_run_batch(g, _ptr(gp), _ptr(rp), _ptr(ci), _ptr(w), _ptr(ir), int(nstep), int(max_rounds),
                                 _ptr(member), _ptr(nb_is), _ptr(steps), _ptr(p2p), _ptr(bst), _ptr(oh))
    if r < 0:
        raise RuntimeError("lgs_oracle_run_batch failed with %d" % r)
    return LgsResult(member, nb_is, steps, p2p, bst, oh)


def greedy(row_ptr, col_idx, order) -> np.ndarray:
    lib = _load()
    rp = np.ascontiguousarray(row_ptr, dtype=np.int64)
    ci = np.ascontiguousarray(col_idx, dtype=np.int32)
    od = np.ascontiguousarray(order, dtype=np.int32)
    n = rp.shape[0] - 1
    member = np.zeros(n, dtype=np.uint8)
    lib.greedy_oracle_run(n, _ptr(rp), _ptr(ci), _ptr(od), _ptr(member))
    return member


def dist_greedy(row_ptr, col_idx, wts, epislon: float = 0.5, init_remain=None, max_rounds: int = -1):
    """heuristics.py:38-74 with the per-round scan in ascending vertex id (see dgs_oracle_run).
    -> (member uint8[n], rounds, order_free)"""
    lib = _load()
    rp = np.ascontiguousarray(row_ptr, dtype=np.int64)
    ci = np.ascontiguousarray(col_idx, dtype=np.int32)
    w = np.ascontiguousarray(np.asarray(wts, dtype=np.float64).reshape(-1))
    n = rp.shape[0] - 1
    assert w.shape[0] == n
    ir = None if init_remain is None else np.ascontiguousarray(init_remain, dtype=np.uint8)
    alpha = 1.0 + (epislon / 3.0)  # heuristics.py:46
    member = np.zeros(n, dtype=np.uint8)
    free_order = ctypes.c_int(1)
    r = lib.dgs_oracle_run(n, _ptr(rp), _ptr(ci), _ptr(w), _ptr(ir), float(alpha), int(max_rounds), _ptr(member),
                           ctypes.byref(free_order))
    if r < 0:
        raise RuntimeError("dgs_oracle_run failed with %d (-2 = did not converge)" % r)
    return member, int(r), bool(free_order.value)
