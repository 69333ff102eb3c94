"""Probe of the streaming (per-layer) kernels and of the fused kernel on large inputs.

    python profiles/micro/stream_probe.py er-batch 16384 [ckpt] [reps]     # G(N,0.1), N~U{100..300}
    python profiles/micro/stream_probe.py big-er 4000000 16 [ckpt] [reps]  # one G(n, m = n*deg/2) graph

Prints device time per solve through dg_solve (DG_MEM_DEVICE).  DG_DISABLE_FUSED=1 forces the per-layer
kernels.  Run under `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum` for the
per-kernel split.  Measurement aid, not a test."""
import ctypes as C
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from distgcn_b200 import engine as E  # noqa: E402
from tests import util  # noqa: E402


def er_batch_device(n_graphs, seed, dev, p=0.1, n_lo=100, n_hi=300):
    """Packed G(N, p) batch built with torch on the device (block-diagonal Bernoulli sampling)."""
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    sizes = torch.randint(n_lo, n_hi + 1, (n_graphs,), generator=g, device=dev)
    gp = torch.zeros(n_graphs + 1, dtype=torch.int64, device=dev)
    gp[1:] = torch.cumsum(sizes, 0)
    rows_l, cols_l = [], []
    chunk = 512
    for c0 in range(0, n_graphs, chunk):
        sz = sizes[c0:c0 + chunk]
        nmax = int(sz.max())
        r = torch.rand((sz.numel(), nmax, nmax), generator=g, device=dev) < p
        iu = torch.triu(torch.ones(nmax, nmax, dtype=torch.bool, device=dev), 1)
        valid = (torch.arange(nmax, device=dev)[None, :] < sz[:, None])
        r = r & iu[None] & valid[:, :, None] & valid[:, None, :]
        b, u, v = torch.nonzero(r, as_tuple=True)
        off = gp[c0:c0 + chunk][b]
        rows_l += [u + off, v + off]
        cols_l += [v + off, u + off]
    rows = torch.cat(rows_l)
    cols = torch.cat(cols_l)
    n = int(gp[-1])
    key = rows * n + cols
    key, _ = torch.sort(key)
    rows = key // n
    cols = (key % n).to(torch.int32)
    rp = torch.zeros(n + 1, dtype=torch.int64, device=dev)
    rp[1:] = torch.cumsum(torch.bincount(rows, minlength=n), 0)
    return gp.to(torch.int32), rp.to(torch.int32), cols.contiguous()


def big_er_device(n, deg, seed, dev):
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    m = n * deg // 2
    u = torch.randint(0, n, (m,), generator=g, device=dev)
    v = torch.randint(0, n, (m,), generator=g, device=dev)
    ok = u != v
    u, v = u[ok], v[ok]
    key = torch.cat([u * n + v, v * n + u])
    key = torch.unique(key)          # sorted, duplicates dropped
    rows = key // n
    cols = (key % n).to(torch.int32)
    rp = torch.zeros(n + 1, dtype=torch.int64, device=dev)
    rp[1:] = torch.cumsum(torch.bincount(rows, minlength=n), 0)
    gp = torch.tensor([0, n], dtype=torch.int32, device=dev)
    return gp, rp.to(torch.int32), cols.contiguous()


def main():
    kind = sys.argv[1]
    dev = torch.device("cuda", 0)
    if kind == "er-batch":
        n_graphs = int(sys.argv[2])
        ck = sys.argv[3] if len(sys.argv) > 3 else "is4sat_l20_c32"
        reps = int(sys.argv[4]) if len(sys.argv) > 4 else 5
        gp, rp, ci = er_batch_device(n_graphs, 0, dev)
    else:
        n, deg = int(sys.argv[2]), int(sys.argv[3])
        ck = sys.argv[4] if len(sys.argv) > 4 else "is4sat_l2_c64"
        reps = int(sys.argv[5]) if len(sys.argv) > 5 else 5
        gp, rp, ci = big_er_device(n, deg, 0, dev)
        n_graphs = 1
    torch.cuda.synchronize()
    n_nodes, nnz = rp.numel() - 1, ci.numel()
    layers = util.load_layers(ck)
    ctx = E.Context(0)
    model = E.Model(ctx, layers, E.gcn_dqn_acts(len(layers)))
    batch = E.DeviceBatch(ctx, graph_ptr=gp, row_ptr=rp, col_idx=ci)
    w = torch.rand(n_nodes, dtype=torch.float64, device=dev)
    member = torch.empty(n_nodes, dtype=torch.uint8, device=dev)
    total = torch.empty(n_graphs, dtype=torch.float64, device=dev)
    lib = ctx._lib
    for _ in range(2):
        E.solve_device(ctx, model, batch, w, member, total=total)
    ctx.synchronize()
    l0 = ctx.launch_count
    ms = C.c_double()
    E.check(lib.dg_timer_start(ctx.handle))
    for _ in range(reps):
        E.solve_device(ctx, model, batch, w, member, total=total)
    E.check(lib.dg_timer_stop(ctx.handle, C.byref(ms)))
    per = ms.value / reps
    out = {"kind": kind, "ckpt": ck, "n_graphs": n_graphs, "n_nodes": n_nodes, "nnz": nnz,
           "fused_disabled": bool(os.environ.get("DG_DISABLE_FUSED")), "ms_per_solve": per,
           "graphs_per_s": n_graphs / per * 1e3, "nodes_per_s": n_nodes / per * 1e3,
           "launches_per_solve": (ctx.launch_count - l0) / reps, "members": int(member.sum())}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
