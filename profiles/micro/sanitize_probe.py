"""Small end-to-end run for compute-sanitizer (memcheck / racecheck / synccheck)."""
import sys
sys.path.insert(0, '.')
import numpy as np
from distgcn_b200 import engine as E
from tests import util
pb, w = util.small_graphs()
pb = pb.slice(20, 32)
w = w[: pb.n_nodes].copy()
w[::5] = 0.0
ctx = E.Context(0)
for short in ("is4sat_l3_c16", "is4sat_l1", "is4sat_l2_c64"):
    layers = util.load_layers(short)
    model = E.Model(ctx, layers, E.gcn_dqn_acts(len(layers)))
    batch = E.DeviceBatch(ctx, pb)
    r = E.solve(ctx, model, batch, w, want_score=True, want_util=True, want_steps=True)
    out = E.gcn_forward(ctx, model, batch)
    l = E.lgs(ctx, batch, w, want_nb_is=True, want_overhead=True)
    print(short, int(r.member.sum()), float(np.abs(out).max()), int(l.member.sum()))
    batch.close(); model.close()
ctx.close()
print("probe done")
