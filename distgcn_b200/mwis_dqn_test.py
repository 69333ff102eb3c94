"""Drop-in for the evaluation loop of the reference's ``mwis_dqn_test.py`` (and the ingest half of SURVEY.md 8f
rank 1): read a directory of the reference's ``.mat`` graphs, solve them all on the GPU, report the
approximation ratio ``p = GCN-guided utility / greedy utility`` per file with the reference's CSV schema.

The reference walks the files one by one (mwis_dqn_test.py:304-348): ``sio.loadmat`` -> ``greedy_search`` on the
raw weights (the normaliser) -> GCN -> ``greedy_search`` on ``act * w`` (:243-256) -> ratio -> CSV ``data,p``.
Here the directory becomes ONE packed batch and two launches: the normaliser (the greedy set on the raw
weights; for distinct weights the synchronous local rounds converge to exactly the sorted scan's set, see
``heuristics.greedy_search``) and the GCN-guided solve.  ``search='local'`` is the scheduler of
``mwis_dqn_call.DQNAgent.solve_mwis`` (local greedy, with zero-weight removal), ``search='greedy'`` the one of the
test script (no zero-weight removal); on distinct utilities both give the same sets.

File format (Data_Generation.py:214-219): ``adj`` float64 CSC [N, N] symmetric 0/1 zero diagonal, ``weights``
float64 [1, N], plus ``greedy_utility`` / ``mwis_utility`` / ``N`` / ``p`` which are returned when present.
"""
from __future__ import annotations

import os
from typing import Dict, List, Optional, Tuple

import numpy as np

from . import engine
from .batch import PackedBatch, pack_graphs
from .heuristics import local_greedy_search_batch


def load_mat_dataset(datapath: str, names: Optional[List[str]] = None) -> Tuple[List[str], PackedBatch, np.ndarray, Dict]:
    """Every ``.mat`` of `datapath` (sorted, or the given names) as one packed batch.
    Returns (names, PackedBatch, weights [n_nodes], extras) where extras holds per-file arrays of the
    optional fields (``greedy_utility``, ``mwis_utility``)."""
    import scipy.io as sio
    if names is None:
        names = sorted(f for f in os.listdir(datapath) if f.endswith(".mat"))
    adjs, wts = [], []
    extras: Dict[str, List[float]] = {"greedy_utility": [], "mwis_utility": []}
    for name in names:
        m = sio.loadmat(os.path.join(datapath, name))
        adj = m["adj"]
        w = np.asarray(m["weights"], dtype=np.float64).reshape(-1)      # stored [1, N]; the script transposes it
        if adj.shape[0] != adj.shape[1] or adj.shape[0] != w.shape[0]:
            raise ValueError("%s: adj %s does not match %d weights" % (name, adj.shape, w.shape[0]))
        adjs.append(adj)
        wts.append(w)
        for k in extras:
            extras[k].append(float(np.asarray(m[k]).reshape(-1)[0]) if k in m else np.nan)
    packed = pack_graphs(adjs)
    return list(names), packed, (np.concatenate(wts) if wts else np.zeros(0)), {k: np.asarray(v) for k, v in extras.items()}


def evaluate(agent, packed: PackedBatch, wts: np.ndarray, search: str = "local"):
    """(p ratio per graph, GCN-guided utility per graph, greedy utility per graph, membership)."""
    if search not in ("local", "greedy"):
        raise ValueError("search must be 'local' or 'greedy'")
    wts = np.asarray(wts, dtype=np.float64).reshape(-1)
    gp = packed.graph_ptr
    base = local_greedy_search_batch(packed, wts)            # greedy_search(adj_0, wts), mwis_dqn_test.py:311
    greedy_util = np.add.reduceat(np.where(base.member == 1, wts, 0.0), gp[:-1]) if packed.n_nodes else np.zeros(0)
    if search == "local":
        member, total = agent.solve_mwis_batch(packed, wts)
    else:
        model = agent.model.compile(agent.ctx)
        batch = engine.DeviceBatch(agent.ctx, packed)
        try:
            r = engine.solve(agent.ctx, model, batch, wts, predict=agent.flags.predict, remove_zero_weight=False)
        finally:
            batch.close()
        member, total = r.member, r.total
    with np.errstate(divide="ignore", invalid="ignore"):
        p = total / greedy_util                                # mwis_dqn_test.py:321
    return p, total, greedy_util, member


def run(datapath: str, model_folder: str, output_csv: Optional[str] = None, search: str = "local", flags=None,
        agent=None):
    """The whole script: returns a list of {"data": file, "p": ratio} rows (and writes them as the reference's
    CSV, mwis_dqn_test.py:342-348, when `output_csv` is given)."""
    from .mwis_dqn_call import DQNAgent
    from .runtime_config import FLAGS
    flags = flags or FLAGS
    if agent is None:
        agent = DQNAgent(flags.feature_size, 5000, flags=flags)
        agent.load(model_folder)
    names, packed, wts, _ = load_mat_dataset(datapath)
    p, _, _, _ = evaluate(agent, packed, wts, search)
    rows = [{"data": n, "p": float(r)} for n, r in zip(names, p)]
    if output_csv:
        os.makedirs(os.path.dirname(os.path.abspath(output_csv)), exist_ok=True)
        with open(output_csv, "w") as f:
            f.write(",data,p\n")                               # pandas' to_csv of a ["data", "p"] frame with its index
            for i, row in enumerate(rows):
                f.write("%d,%s,%r\n" % (i, row["data"], row["p"]))
    return rows
