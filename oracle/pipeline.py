"""CPU restatement of the whole hot path for one graph / a batch.  TEST INFRASTRUCTURE ONLY
(checker in tests/ and smoke(); timed as the CPU baseline / reference arm by bench.py).

Follows DQNAgent.solve_mwis (mwis_dqn_call.py:198-261): drop zero-weight vertices, build features and
supports (makestate, :129-138), run the GraphConvolution stack (predict, :140-143), multiply scores by
link weights (:230-235), run the local greedy search (:238), map back to original ids (:240-241).
The GCN half is oracle/gcn_oracle.py (numpy/scipy, as the reference's host code is), the greedy
half is oracle/lgs_oracle.c.
"""
from __future__ import annotations

import multiprocessing as mp
import os
from typing import List, Sequence, Tuple

import numpy as np
import scipy.sparse as sp

from . import gcn_oracle as G
from . import lgs as L


def solve_graph(adj, w, layers, predict: str = "mwis", kind: str = "gcn_dqn", remove_zero_weight: bool = True,
                generation: int = 1):
    """One graph.  Returns (score[N] fp32, util[N] fp64, member[N] uint8) indexed by ORIGINAL vertex
    ids; removed vertices carry zeros.  generation 2 = MWISSolver.solve_mwis (mwis_gdpg_call.py:200-235): no
    zero-weight removal, all-ones features (:82-96)."""
    w = np.asarray(w, dtype=np.float64).reshape(-1)
    n = w.shape[0]
    if generation == 2:
        remove_zero_weight = False
    if remove_zero_weight:
        keep = np.where(w > 0)[0]  # kp_nodes, mwis_dqn_call.py:204
    else:
        keep = np.arange(n)
    score = np.zeros(n, dtype=np.float32)
    util = np.zeros(n, dtype=np.float64)
    member = np.zeros(n, dtype=np.uint8)
    if keep.shape[0] == 0:
        return score, util, member
    a = sp.csr_matrix(adj)
    if keep.shape[0] != n:
        a = a[keep][:, keep].tocsr()
    wk = w[keep]
    feats = G.features_gen2(wk, layers[0].c_in, predict) if generation == 2 else G.features_gen1(wk, layers[0].c_in)
    sup = G.laplacian_supports(a, len(layers[0].weights) - 1)
    act = G.gcn_forward(feats, sup, layers, kind)
    u = G.utility(act[:, 0], wk, predict)
    r = L.run(a.indptr, a.indices, u)
    score[keep] = act[:, 0]
    util[keep] = u
    member[keep] = r.member
    return score, util, member


def solve_graph_dit(adj, w, layers, predict: str = "mwis", kind: str = "gcn_dqn", max_iter: int = 100000):
    """GCN embedded into the LGS iteration, MWISSolver.solve_mwis_dit (mwis_gdpg_call.py:278-318), for one graph.
    Generation-2 glue: no zero-weight removal, features from features_gen2.  Returns (member[N] uint8,
    best_IS_util, iterations)."""
    a0 = sp.csr_matrix(adj)
    wts = np.asarray(w, dtype=np.float64).reshape(-1)
    n = wts.shape[0]
    nis = -np.ones(n)                                   # nIS_vec, :287
    it = 0
    while (nis == -1).sum() > 0 and it < max_iter:      # :288
        remain = nis == -1                              # :290
        rev = np.flatnonzero(remain)                    # :291-292
        a = a0[remain, :][:, remain].tocsr()            # :293-295
        wk = wts[remain]
        if np.sum(wk) <= 0:                             # :298-299
            break
        it += 1
        feats = G.features_gen2(wk, layers[0].c_in, predict)     # makestate, :82-96
        sup = G.laplacian_supports(a, len(layers[0].weights) - 1)
        act = G.gcn_forward(feats, sup, layers, kind)
        u = G.utility(act[:, 0], wk, predict)           # :304-307
        r = L.run(a.indptr, a.indices, u, nstep=1)      # local_greedy_search_nstep(..., nstep=1), :309
        nis[rev[r.member == 1]] = 1                     # :311
        nis[rev[r.nb_is == 1]] = 0                      # :312
    member = (nis == 1).astype(np.uint8)
    return member, float(np.dot(nis, wts)), it          # :313


# ---- batch form with a process pool (the reference is single-threaded Python; graphs are
# independent, so a fan-out over all host cores is the fairest multi-core version of it) ----------
_POOL_STATE = {}


def _pool_init(graph_ptr, row_ptr, col_idx, wts, layers, predict):
    for var in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
        os.environ[var] = "1"
    try:
        from threadpoolctl import threadpool_limits
        _POOL_STATE["limit"] = threadpool_limits(limits=1)
    except Exception:
        pass
    _POOL_STATE.update(graph_ptr=graph_ptr, row_ptr=row_ptr, col_idx=col_idx, wts=wts, layers=layers, predict=predict)


def _graph_adj(graph_ptr, row_ptr, col_idx, g):
    v0, v1 = int(graph_ptr[g]), int(graph_ptr[g + 1])
    e0, e1 = int(row_ptr[v0]), int(row_ptr[v1])
    n = v1 - v0
    indptr = (row_ptr[v0:v1 + 1] - e0).astype(np.int32)
    indices = (col_idx[e0:e1] - v0).astype(np.int32)
    return sp.csr_matrix((np.ones(e1 - e0), indices, indptr), shape=(n, n)), v0, v1


def _pool_work(span):
    s = _POOL_STATE
    out = []
    for g in range(span[0], span[1]):
        adj, v0, v1 = _graph_adj(s["graph_ptr"], s["row_ptr"], s["col_idx"], g)
        _, _, member = solve_graph(adj, s["wts"][v0:v1], s["layers"], s["predict"])
        out.append(member)
    return span[0], (np.concatenate(out) if out else np.zeros(0, np.uint8))


class BatchSolver:
    """Keeps a worker pool alive across steps so that process start-up is not timed."""

    def __init__(self, graph_ptr, row_ptr, col_idx, wts, layers, predict="mwis", n_procs: int = 0):
        self.graph_ptr = np.asarray(graph_ptr)
        self.row_ptr = np.asarray(row_ptr)
        self.col_idx = np.asarray(col_idx)
        self.wts = np.asarray(wts, dtype=np.float64)
        self.layers = layers
        self.predict = predict
        self.n_procs = n_procs or (os.cpu_count() or 1)
        self.pool = None
        if self.n_procs > 1:
            ctx = mp.get_context("fork")
            self.pool = ctx.Pool(self.n_procs, initializer=_pool_init,
                                 initargs=(self.graph_ptr, self.row_ptr, self.col_idx, self.wts, layers, predict))
        else:
            _pool_init(self.graph_ptr, self.row_ptr, self.col_idx, self.wts, layers, predict)

    def solve(self, g0: int, g1: int) -> np.ndarray:
        """Membership of graphs g0 .. g1-1 (concatenated)."""
        n = g1 - g0
        if n <= 0:
            return np.zeros(0, np.uint8)
        if self.pool is None:
            return _pool_work((g0, g1))[1]
        chunks = max(1, min(n, self.n_procs * 4))
        bounds = np.linspace(g0, g1, chunks + 1).astype(int)
        spans = [(int(bounds[i]), int(bounds[i + 1])) for i in range(chunks) if bounds[i + 1] > bounds[i]]
        parts = sorted(self.pool.map(_pool_work, spans), key=lambda t: t[0])
        return np.concatenate([p[1] for p in parts])

    def close(self):
        if self.pool is not None:
            self.pool.close()
            self.pool.join()
            self.pool = None
