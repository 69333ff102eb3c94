"""Single-GPU solve of one synthetic G(n, average degree 16) graph (config-5 shape) with the c64 l2 checkpoint: wall clock per
solve; under `ncu --metrics gpu__time_duration.sum` the launch list shows where the time goes.  DG_GIANT_N selects n."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np, torch
from distgcn_b200 import engine as E
from profiles.micro.stream_probe import big_er_device
from tests import util
n = int(os.environ.get("DG_GIANT_N", "12000000"))
reps = int(os.environ.get("DG_GIANT_REPS", "3"))
dev = torch.device("cuda", 0)
gp, rp, ci = big_er_device(n, 16, 0, dev)
g = torch.Generator(device=dev); g.manual_seed(1)
w = torch.rand(n, dtype=torch.float64, device=dev, generator=g)
w[torch.rand(n, device=dev, generator=g) < 0.05] = 0.0
layers = util.load_layers(os.environ.get("DG_GIANT_CKPT", "is4sat_l2_c64"))
ctx = E.Context(0)
model = E.Model(ctx, layers, E.gcn_dqn_acts(len(layers)))
batch = E.DeviceBatch(ctx, graph_ptr=gp, row_ptr=rp, col_idx=ci)
ref = torch.empty(n, dtype=torch.uint8, device=dev)
E.solve_device(ctx, model, batch, w, ref)
ctx.synchronize()
t0 = time.perf_counter()
for _ in range(reps):
    E.solve_device(ctx, model, batch, w, ref)
ctx.synchronize()
ms = 1e3 * (time.perf_counter() - t0) / reps
nnz = int(ci.numel())
print("n=%d nnz=%d: %.3f ms per solve, %.2f G vertices/s, members %d, kernel %s" % (n, nnz, ms, n / ms / 1e6, int(ref.sum()), ctx.last_kernel))
