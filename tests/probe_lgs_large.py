import sys
sys.path.insert(0, '.')
import numpy as np, scipy.sparse as sp
from distgcn_b200 import engine as E
from distgcn_b200.batch import pack_graphs
from oracle import lgs as L
rng = np.random.default_rng(3)
adjs = []
for n, deg in ((20000, 8), (9000, 3), (50, 4)):
    m = n * deg // 2
    u = rng.integers(0, n, m); v = rng.integers(0, n, m); ok = u != v
    a = sp.coo_matrix((np.ones(ok.sum()), (u[ok], v[ok])), shape=(n, n))
    adjs.append(((a + a.T) > 0).astype(np.float64).tocsr())
pb = pack_graphs(adjs)
ctx = E.Context(0)
batch = E.DeviceBatch(ctx, pb)
w = rng.random(pb.n_nodes)
o = L.run_batch(pb.graph_ptr, pb.row_ptr, pb.col_idx, w)
for rep in range(3):
    r = E.lgs(ctx, batch, w, want_nb_is=True, want_overhead=True)
    bad = np.flatnonzero(r.member != o.member)
    print("rep", rep, "mismatches", bad.shape[0], "first", bad[:5], "steps", r.steps, o.steps, "members", int(r.member.sum()), int(o.member.sum()))
    r = E.lgs(ctx, batch, w)
    bad = np.flatnonzero(r.member != o.member)
    print("   plain: mismatches", bad.shape[0], "steps", r.steps)
v = 21479
g = 1
print("vertex", v, "weight", w[v], "neighbours", pb.col_idx[pb.row_ptr[v]:pb.row_ptr[v + 1]], "their weights", w[pb.col_idx[pb.row_ptr[v]:pb.row_ptr[v + 1]]])
for kw in ({"want_nb_is": True}, {"want_overhead": True}, {"want_nb_is": True, "want_overhead": True}, {"want_count": True} if "want_count" in E.lgs.__code__.co_varnames else {}):
    r = E.lgs(ctx, batch, w, **kw)
    bad = np.flatnonzero(r.member != o.member)
    print(kw, "mismatches", bad.shape[0], bad[:4], "nb_is[v]", None if r.nb_is is None else int(r.nb_is[v]), "oracle nb_is", int(o.nb_is[v]), "member", int(r.member[v]), int(o.member[v]))
