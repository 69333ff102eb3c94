"""Throughput of tc_solve_kernel on synthetic graphs of n_lo..n_hi vertices (device-resident, CUDA events)."""
import os, sys, ctypes as C
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
import bench
from distgcn_b200 import engine as E
from tests import util
n_lo, n_hi, n_graphs = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
rng = np.random.default_rng(0)
pb, w = bench.synth_er_batch(rng, n_graphs, n_lo, n_hi, 0.1)
layers = util.load_layers("is4sat_l20_c32")
ctx = E.Context(0)
model = E.Model(ctx, layers, E.gcn_dqn_acts(len(layers)))
batch = E.DeviceBatch(ctx, pb)
for _ in range(3):
    r = E.solve(ctx, model, batch, w, predict="mwis", remove_zero_weight=True)
ref = r.member.copy()
import torch
d_w = torch.from_numpy(w).cuda(); d_m = torch.empty(pb.n_nodes, dtype=torch.uint8, device="cuda"); d_t = torch.empty(pb.n_graphs, dtype=torch.float64, device="cuda")
for _ in range(3):
    E.solve_device(ctx, model, batch, d_w, d_m, predict="mwis", remove_zero_weight=True, total=d_t)
t = bench.DeviceTimer(ctx); t.start()
K = 20
for _ in range(K):
    E.solve_device(ctx, model, batch, d_w, d_m, predict="mwis", remove_zero_weight=True, total=d_t)
ms = t.stop() / K
same = bool(np.array_equal(d_m.cpu().numpy(), ref))
print("%s n=%d..%d graphs=%d: %.4f ms/step, %.3f M graphs/s, members %d same=%s" % (ctx.last_kernel, n_lo, n_hi, n_graphs, ms, n_graphs / ms / 1e3, int(ref.sum()), same))
