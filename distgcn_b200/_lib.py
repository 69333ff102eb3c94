"""ctypes binding of libdistgcn_b200.so (C-ABI declared in include/distgcn_b200.h).

There is no CPU fallback: importing this module without the built library, or creating a context
without a CUDA device, raises.  Build the library with ``python -m distgcn_b200.build`` (or
``__graft_entry__.build()``).
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("DISTGCN_B200_LIB") or os.path.join(HERE, "libdistgcn_b200.so")  # (override: experiments)

OK = 0
ERR_INVALID, ERR_CUDA, ERR_UNSUPPORTED, ERR_NOT_CONVERGED, ERR_NO_DEVICE = 1, 2, 3, 4, 5
MEM_HOST, MEM_DEVICE = 0, 1
ACT_IDENTITY, ACT_LEAKY_RELU, ACT_RELU = 0, 1, 2
PREDICT_MWIS, PREDICT_MIS = 0, 1
HEAD_LINEAR, HEAD_PAIR_SOFTMAX = 0, 1


class DistGCNError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__("distgcn_b200 error %d: %s" % (code, message))
        self.code = code


class LibraryMissingError(ImportError):
    pass


_p = C.c_void_p
_i32 = C.c_int32

# name -> (restype, argtypes); kept in one table so tests can check it against the header
SIGNATURES = {
    "dg_version": (C.c_int, []),
    "dg_last_error": (C.c_char_p, []),
    "dg_device_count": (C.c_int, []),
    "dg_context_create": (C.c_int, [C.c_int, _p, C.POINTER(_p)]),
    "dg_context_destroy": (None, [_p]),
    "dg_context_synchronize": (C.c_int, [_p]),
    "dg_context_reload_env": (C.c_int, [_p]),
    "dg_host_alloc": (_p, [C.c_uint64]),
    "dg_host_free": (None, [_p]),
    "dg_context_launch_count": (C.c_uint64, [_p]),
    "dg_context_last_kernel": (C.c_char_p, [_p]),
    "dg_context_wait": (C.c_int, [_p, _p]),
    "dg_timer_start": (C.c_int, [_p]),
    "dg_timer_stop": (C.c_int, [_p, _p]),
    "dg_profile_enable": (C.c_int, [_p, C.c_int]),
    "dg_profile_collect": (C.c_int, [_p, _p, _p, _p]),
    "dg_model_create": (C.c_int, [_p, C.c_int, C.c_int, _p, _p, _p, _p, _p, C.c_float, C.c_int, C.POINTER(_p)]),
    "dg_model_destroy": (None, [_p]),
    "dg_model_out_width": (C.c_int, [_p]),
    "dg_batch_create": (C.c_int, [_p, _i32, _i32, _i32, _p, _p, _p, C.c_int, C.POINTER(_p)]),
    "dg_batch_destroy": (None, [_p]),
    "dg_batch_set_keep": (C.c_int, [_p, _p, C.c_int]),
    "dg_batch_set_keep_from_weights": (C.c_int, [_p, _p, C.c_int]),
    "dg_batch_set_x0": (C.c_int, [_p, _p, C.c_int]),
    "dg_graph_convolution": (C.c_int, [_p, _p, _i32, _i32, _p, _p, _p, C.c_int, C.c_float, _p, _p, C.c_int]),
    "dg_gcn_forward": (C.c_int, [_p, _p, _p, _p, C.c_int]),
    "dg_spmm_laplacian": (C.c_int, [_p, _p, _i32, _p, _p, C.c_int]),
    "dg_utility": (C.c_int, [_p, _p, _p, _i32, _p, C.c_int, _p, C.c_int]),
    "dg_lgs": (C.c_int, [_p, _p, _p, _i32, _p, _p, _p, _p, _p, _p, C.c_int]),
    "dg_dist_greedy": (C.c_int, [_p, _p, _p, C.c_double, _p, _p, C.c_int]),
    "dg_member_weight": (C.c_int, [_p, _p, _p, _p, _p, C.c_int]),
    "dg_solve": (C.c_int, [_p, _p, _p, _p, C.c_int, C.c_int, _p, _p, _p, _p, _p, C.c_int]),
    "dg_part_create": (C.c_int, [_p, _i32, _i32, _i32, _i32, _p, _p, C.c_int, C.POINTER(_p)]),
    "dg_part_destroy": (None, [_p]),
    "dg_part_prepare": (C.c_int, [_p, _i32, _p, _p, _p, _p]),
    "dg_part_first": (C.c_int, [_p, _i32, _p, _p, _p, _p, _p]),
    "dg_part_project": (C.c_int, [_p, _p, _p, _p, _p]),
    "dg_part_layer": (C.c_int, [_p, _p, _i32, _p, _p, _p, _p]),
    "dg_part_tail": (C.c_int, [_p, _p, _p, _p, _p]),
    "dg_part_last": (C.c_int, [_p, _p, _p, _p, _p, _p, C.c_int, _p, _p]),
    "dg_model_padded_width": (C.c_int, [_p, _i32]),
    "dg_part_lgs_init": (C.c_int, [_p, _p, _p, _p, _p]),
    "dg_part_lgs_decide": (C.c_int, [_p, _p, _p, _p, _p]),
    "dg_part_lgs_remove": (C.c_int, [_p, _p, _p, _p]),
    "dg_part_lgs_run": (C.c_int, [_p, _p, _p, _p, _p, _p, C.c_int32, _p]),
    "dg_peer_alloc": (C.c_int, [_p, C.c_uint64, _p, _p]),
    "dg_peer_open": (C.c_int, [_p, _p, _p]),
    "dg_peer_close": (C.c_int, [_p, _p]),
    "dg_peer_free": (C.c_int, [_p, _p]),
    "dg_part_set_peers": (C.c_int, [_p, _i32, _i32, _p, C.c_uint64, C.c_uint64, C.c_uint64]),
    "dg_part_keep": (C.c_int, [_p, _p, C.c_int, _i32, _p]),
    "dg_part_barrier": (C.c_int, [_p, _p]),
    "dg_solve_host": (C.c_int, [_p, _p, _i32, _i32, _i32, _p, _p, _p, _p, C.c_int, C.c_int, _p, _p]),
    "dg_solve_dit": (C.c_int, [_p, _p, _p, _p, C.c_int, _p, _p, _p, C.c_int]),
    "dg_solve_host_async": (C.c_int, [_p, _p, _i32, _i32, _i32, _p, _p, _p, _p, C.c_int, C.c_int, _p, _p]),
    "dg_solve_host_compact": (C.c_int, [_p, _p, _i32, _i32, _i32, _p, _p, _p, _p, C.c_int, C.c_int, _p, _p, C.c_int]),
    "dg_pack_graphs_sizes": (C.c_int, [_i32, _p, _p, _p, _p, _p, _p]),
    "dg_pack_graphs_host": (C.c_int, [_i32, _p, _p, _p, _p, _p, _p, _p, _p, _i32]),
    "dg_pack_graphs_upper_host": (C.c_int, [_i32, _p, _p, _p, _p, _p, _p, _p, _i32]),
    "dg_solve_host_upper": (C.c_int, [_p, _p, _i32, _i32, _i32, _p, _p, _p, _p, C.c_int, C.c_int, _p, _p, C.c_int]),
    "dg_solve_graphs_host": (C.c_int, [_p, _p, _i32, _p, _p, _p, _p, _p, _p, C.c_int, C.c_int, _p, _p, C.c_int]),
    "dg_wireless_create": (C.c_int, [_p, _i32, _i32, _i32, _p, _p, _p, _p, _i32, C.POINTER(_p)]),
    "dg_wireless_destroy": (None, [_p]),
    "dg_wireless_buffers": (C.c_int, [_p, C.POINTER(_p), C.POINTER(_p), C.POINTER(_p)]),
    "dg_wireless_begin_slot": (C.c_int, [_p, _i32]),
    "dg_wireless_joint_weights": (C.c_int, [_p, _i32]),
    "dg_wireless_joint_serve": (C.c_int, [_p, _i32]),
    "dg_wireless_seq_weights": (C.c_int, [_p, _i32, _i32]),
    "dg_wireless_seq_serve": (C.c_int, [_p, _i32, _i32]),
    "dg_wireless_end_slot": (C.c_int, [_p, _i32]),
    "dg_wireless_run": (C.c_int, [_p, _p, _p, C.c_int32, C.c_int32, C.c_int32, C.c_int, C.c_int, C.c_double, C.c_int32, C.c_int32]),
    "dg_wireless_read_history": (C.c_int, [_p, _p]),
}

_lib = None


def load() -> C.CDLL:
    """Load the shared library (once).  Raises LibraryMissingError if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise LibraryMissingError(
            "%s not found: build it with `python -m distgcn_b200.build` (needs nvcc). "
            "distgcn_b200 has no CPU fallback." % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the library does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def last_error() -> str:
    msg = load().dg_last_error()
    return msg.decode("utf-8", "replace") if msg else ""


def check(status: int) -> None:
    if status != OK:
        raise DistGCNError(status, last_error())
