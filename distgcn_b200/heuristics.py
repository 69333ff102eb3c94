"""Drop-in for the local-greedy entry points of the reference's ``heuristics.py``.

Same names, arguments and return tuples (heuristics.py:77-305); the rounds run on the GPU (dg_lgs).
``adj`` may be any scipy sparse matrix / dense array (one graph), exactly like the reference, or an
``engine.DeviceBatch`` holding ONE graph that the caller keeps resident to skip the upload.

Batched forms (``*_batch``) take a PackedBatch / DeviceBatch with many graphs and return membership
vectors instead of Python sets: that is the form that runs at full speed.
"""
from __future__ import annotations

import numpy as np

from . import engine
from .batch import PackedBatch, pack_graphs
from .runtime import default_context


def _as_batch(adj, ctx=None):
    if isinstance(adj, engine.DeviceBatch):
        return adj, False
    ctx = ctx or default_context()
    if isinstance(adj, PackedBatch):
        return engine.DeviceBatch(ctx, adj), True
    return engine.DeviceBatch(ctx, pack_graphs([adj])), True


def _flat_weights(wts, n):
    w = np.ascontiguousarray(np.asarray(wts, dtype=np.float64).reshape(-1))  # np.array(wts).flatten(), heuristics.py:84
    if w.shape[0] != n:
        raise ValueError("weights have %d entries for %d vertices" % (w.shape[0], n))
    return w


def _member_set(member):
    return set(np.flatnonzero(member).tolist())


def _run_single(adj, wts, nstep=-1, **wants):
    batch, owned = _as_batch(adj)
    try:
        if batch.n_graphs != 1:
            raise ValueError("the per-graph entry points take one graph; use the *_batch forms for %d graphs"
                             % batch.n_graphs)
        w = _flat_weights(wts, batch.n_nodes)
        return engine.lgs(batch.ctx, batch, w, nstep=nstep, **wants), w
    finally:
        if owned:
            batch.close()


def local_greedy_search(adj, wts):
    """Return MWIS set and the total weights of MWIS (heuristics.py:77-116)."""
    res, w = _run_single(adj, wts, want_steps=False)
    mwis = _member_set(res.member)
    return mwis, np.sum(w[list(mwis)])


def local_greedy_search_count(adj, wts):
    """... and the number of rounds (heuristics.py:119-160)."""
    res, w = _run_single(adj, wts, want_steps=True)
    mwis = _member_set(res.member)
    return mwis, np.sum(w[list(mwis)]), int(res.steps[0])


def local_greedy_search_stats(adj, wts):
    """... and the message counts p2p, bst (heuristics.py:163-209)."""
    res, w = _run_single(adj, wts, want_steps=True, want_stats=True)
    mwis = _member_set(res.member)
    return mwis, np.sum(w[list(mwis)]), int(res.steps[0]), int(res.p2p[0]), int(res.bst[0])


def local_greedy_search_overhead(adj, wts):
    """... and the per-vertex overhead vector (heuristics.py:212-263)."""
    res, w = _run_single(adj, wts, want_steps=True, want_overhead=True)
    mwis = _member_set(res.member)
    return mwis, np.sum(w[list(mwis)]), int(res.steps[0]), int(res.p2p[0]), int(res.bst[0]), res.oh_vec


def local_greedy_search_nstep(adj, wts, nstep=1):
    """At most nstep rounds; also returns nb_is (heuristics.py:266-305).  A negative nstep never stops
    on the step counter in the reference (``while ... and step`` with step < 0); same here."""
    res, w = _run_single(adj, wts, nstep=int(nstep) if nstep >= 0 else -1, want_nb_is=True, want_steps=False)
    mwis = _member_set(res.member)
    return mwis, np.sum(w[list(mwis)]), _member_set(res.nb_is)


def greedy_search(adj, wts):
    """Centralised greedy (heuristics.py:13-35): visit vertices by descending weight, take a vertex unless
    a taken vertex is adjacent.  For DISTINCT weights this lexicographically-first independent set is
    exactly what the synchronous local greedy rounds converge to (a vertex that dominates its remaining
    neighbourhood is the next one the sorted scan would take), so it runs on the same kernel.  With equal
    weights the reference's own result depends on numpy's unstable argsort; here ties go to the smaller
    index, which is one of the orders the reference can produce."""
    return local_greedy_search(adj, wts)


def dist_greedy_search(adj, wts, epislon=0.5):
    """Threshold distributed greedy (heuristics.py:38-74; the misspelt keyword is the reference's).  Per round the
    remaining vertices within a factor alpha = 1 + epislon / 3 of their heaviest remaining neighbour are candidates,
    and a maximal independent subset of the candidates joins.  The reference picks that subset by walking a Python
    set; here the candidates are walked in ascending vertex id - the same result whenever no two candidates of a
    round are adjacent, one of the reference's possible results otherwise."""
    batch, owned = _as_batch(adj)
    try:
        if batch.n_graphs != 1:
            raise ValueError("the per-graph entry points take one graph; use dist_greedy_search_batch for %d graphs"
                             % batch.n_graphs)
        w = _flat_weights(wts, batch.n_nodes)
        res = engine.dist_greedy(batch.ctx, batch, w, epsilon=epislon, want_steps=False)
    finally:
        if owned:
            batch.close()
    mwis = _member_set(res.member)
    return mwis, np.sum(w[list(mwis)])


# ---- batched forms --------------------------------------------------------------------------------
def local_greedy_search_batch(graphs, wts, nstep=-1, stats=False, overhead=False, nb_is=False):
    """Many graphs at once.  `graphs`: PackedBatch, DeviceBatch or a list of adjacency matrices.
    Returns an engine.LgsResult (membership vector over the packed vertices + per-graph arrays)."""
    if isinstance(graphs, (list, tuple)):
        graphs = pack_graphs(graphs)
    batch, owned = _as_batch(graphs)
    try:
        w = _flat_weights(wts, batch.n_nodes)
        return engine.lgs(batch.ctx, batch, w, nstep=nstep, want_nb_is=nb_is, want_steps=True,
                          want_stats=stats or overhead, want_overhead=overhead)
    finally:
        if owned:
            batch.close()


def dist_greedy_search_batch(graphs, wts, epislon=0.5):
    """dist_greedy_search on many graphs at once -> engine.LgsResult (membership vector, rounds per graph)."""
    if isinstance(graphs, (list, tuple)):
        graphs = pack_graphs(graphs)
    batch, owned = _as_batch(graphs)
    try:
        w = _flat_weights(wts, batch.n_nodes)
        return engine.dist_greedy(batch.ctx, batch, w, epsilon=epislon, want_steps=True)
    finally:
        if owned:
            batch.close()
