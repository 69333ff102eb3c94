"""CPU restatement of one (network, load) run of the reference's multi-channel slot loop
(wireless_dqn_test_mc.py:225-366).  TEST INFRASTRUCTURE ONLY: the checker of distgcn_b200/wireless.py.

One instance at a time, one solver call per slot (or per channel), exactly as the reference script walks it;
the schedulers are the oracle's (oracle/pipeline.py for DQNAgent.solve_mwis, oracle/lgs.py for
heuristics.local_greedy_search).  Inputs (per-channel graphs, joint graph, arrivals, rates) are handed in, so
this file needs nothing from the product package.
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp

from . import lgs as L
from . import pipeline


def _lgs_set(adj, w):
    a = sp.csr_matrix(adj)
    return np.flatnonzero(L.run(a.indptr, a.indices, w).member)


def _dgcn_set(adj, w, layers, predict, generation=1):
    _, _, member = pipeline.solve_graph(adj, w, layers, predict, generation=generation)
    return np.flatnonzero(member)


def run_instance(adj_list, adj_gK, arrivals, rates, algo, layers=None, predict="mwis", n_slots=None, agent_generation=1):
    """Returns (queue matrix [T, nflows], list of per-slot schedules as sorted vertex arrays).  `agent_generation`: which
    agent DGCN-LGS calls on the joint graph - 1 = mwis_dqn_call.DQNAgent.solve_mwis (drops zero-weight vertices first),
    2 = mwis_gdpg_call.MWISSolver.solve_mwis (keeps them, all-ones features)."""
    T, nflows = arrivals.shape
    n_ch = len(adj_list)
    if n_slots is not None:
        T = min(T, n_slots + 1)
    queue = np.zeros((T, nflows))
    schedules = []
    for t in range(1, T):
        queue[t, :] = queue[t - 1, :] + arrivals[t, :]                                   # :227
        qa = queue[t, :][:, None] * np.ones((nflows, n_ch))                              # :228
        wts0 = qa * rates[t, :, :]                                                       # :230
        wts1 = np.reshape(wts0, nflows * n_ch, order="F")                                # :240
        if algo == "Greedy":
            mwis = _lgs_set(adj_gK, wts1)                                                # :244
        elif algo == "Greedy-Th":
            a = sp.csr_matrix(adj_gK)
            mwis = np.flatnonzero(L.dist_greedy(a.indptr, a.indices, wts1, 0.1)[0])      # :252 (ascending-id scan, lgs_oracle.c)
        elif algo == "DGCN-LGS":
            mwis = _dgcn_set(adj_gK, wts1, layers, predict, agent_generation)            # :289
        elif algo == "DGCN-LGS-it":
            mwis = np.flatnonzero(pipeline.solve_graph_dit(adj_gK, wts1, layers, predict)[0])   # :264
        elif algo in ("LGS-Seq", "DGCN-LGS-Seq"):
            parts = []
            for ic in range(n_ch):
                wts_ic = qa[:, ic] * rates[t, :, ic]                                     # :298
                wts_idx, = np.nonzero(wts_ic)                                            # :299
                adj_ii = sp.csr_matrix(adj_list[ic])[wts_idx, :][:, wts_idx]             # :300-301
                if wts_idx.shape[0] == 0:
                    sel = np.zeros(0, dtype=np.int64)
                elif algo == "LGS-Seq":
                    sel = _lgs_set(adj_ii, wts_ic[wts_idx])                              # :302
                else:
                    sel = _dgcn_set(adj_ii, wts_ic[wts_idx], layers, predict)            # :323
                mwis_ls = wts_idx[sel]
                parts.append(mwis_ls + ic * nflows)                                      # :303
                if ic + 1 < n_ch:
                    depart_est = np.minimum(qa[:, ic], rates[t, :, ic])                  # :307
                    qa[:, ic + 1] = qa[:, ic]
                    qa[mwis_ls, ic + 1] -= depart_est[mwis_ls]                           # :309
            mwis = np.concatenate(parts) if parts else np.zeros(0, dtype=np.int64)
        else:
            raise ValueError(algo)
        mwis = np.sort(np.asarray(mwis, dtype=np.int64))
        schedules.append(mwis)
        rates_ts = np.reshape(rates[t, :, :], nflows * n_ch, order="F")                  # :359
        capacity = np.zeros(nflows)
        capacity[mwis % nflows] = rates_ts[mwis]       # ascending vertex id: the highest channel wins (:360-363)
        dep = np.minimum(queue[t, :], capacity)                                          # :364
        queue[t, :] = queue[t, :] - dep                                                  # :365
    return queue, schedules
