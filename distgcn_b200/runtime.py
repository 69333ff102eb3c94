"""Process-wide default CUDA contexts (one per device), created on first use."""
from __future__ import annotations

import os
import threading

from . import engine

_lock = threading.Lock()
_contexts = {}


def default_device() -> int:
    return int(os.environ.get("LOCAL_RANK", os.environ.get("DG_DEVICE", "0")))


def default_context(device=None) -> engine.Context:
    """The shared engine.Context of `device` (default: LOCAL_RANK / DG_DEVICE / 0).  Raises if there is
    no CUDA device - there is no CPU fallback."""
    dev = default_device() if device is None else int(device)
    with _lock:
        ctx = _contexts.get(dev)
        if ctx is None or ctx._h is None:
            ctx = engine.Context(dev)
            _contexts[dev] = ctx
        return ctx
