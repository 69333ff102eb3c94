"""Drop-in for the inference side of the reference's ``mwis_dqn_call.py`` (generation-1 agent).

    from distgcn_b200.mwis_dqn_call import dqn_agent          # lazily created singleton, as at :344
    dqn_agent.load(find_model_folder(FLAGS, 'dqn'))
    mwis, total_wt, _ = dqn_agent.solve_mwis(adj, wts)

``DQNAgent`` keeps ``load / makestate / predict / act / solve_mwis`` with the reference's signatures and
return values (mwis_dqn_call.py:104-261).  The whole body of ``solve_mwis`` - zero-weight removal,
supports, GraphConvolution stack, utility product, local greedy search, mapping back to original vertex
ids - is one call into the CUDA library (dg_solve).  ``solve_mwis_batch`` is the form that runs at
full speed.  The training members (memorize / replay / save, epsilon-greedy exploration with
train=True) are out of scope and raise NotImplementedError.
"""
from __future__ import annotations

import numpy as np

from . import ckpt, engine
from . import layers as L
from .batch import PackedBatch, pack_graphs
from .models import GCN_DQN
from .runtime import default_context
from .runtime_config import FLAGS, flags  # noqa: F401  (re-exported like the reference module does)


class State(dict):
    """What makestate returns.  The reference returns {"features": tuple, "support": [tuples]}
    (mwis_dqn_call.py:137); here the graph lives on the device: {"batch": DeviceBatch, "n": N}."""


class DQNAgent:
    def __init__(self, feature_size=32, memory_size=5000, flags=None, device=None):
        self.flags = flags or FLAGS
        self.feature_size = feature_size
        self.memory_size = memory_size
        self.smallconst = 0.000001
        self.gamma = 0.95
        self.epsilon = getattr(self.flags, "epsilon", 1.0)
        self.epsilon_min = getattr(self.flags, "epsilon_min", 0.001)
        self.epsilon_decay = 0.985
        self.learning_rate = self.flags.learning_rate
        self.ctx = default_context(device)
        # True: stored zeros in an adjacency matrix are not edges (np.nonzero semantics, heuristics.py:94) at the cost
        # of one pass over the stored values per call; the reference's datasets store only ones
        self.check_values = True
        self.placeholders = L.make_placeholders(1 + self.flags.max_degree, feature_size)
        self.model = self._build_model()

    def _build_model(self):
        return GCN_DQN(self.placeholders, input_dim=self.feature_size, logging=True, flags=self.flags)

    # ---- checkpoint ----------------------------------------------------------------------------------
    def load(self, name):
        """Restore from `name`/checkpoint -> model.ckpt.{index,data}; like the reference, a directory
        without a `checkpoint` state file is silently ignored (mwis_dqn_call.py:188-192)."""
        prefix = ckpt.checkpoint_prefix(name)
        if prefix:
            self.model.load(prefix)
            self.feature_size = self.model.input_dim
            print("loaded " + prefix)

    def save(self, name):
        raise NotImplementedError("saving checkpoints belongs to training, which is out of scope")

    def memorize(self, *args, **kwargs):
        raise NotImplementedError("experience replay belongs to training, which is out of scope")

    def replay(self, batch_size):
        raise NotImplementedError("experience replay belongs to training, which is out of scope")

    # ---- inference -----------------------------------------------------------------------------------
    def makestate(self, adj, wts_nn):
        """Upload the graph (mwis_dqn_call.py:129-138).  Features are w/||w|| row-normalised = 1/F on
        every vertex with a non-zero weight; zero-weight rows are empty."""
        w = np.asarray(wts_nn, dtype=np.float64).reshape(-1)
        batch = engine.DeviceBatch(self.ctx, pack_graphs([adj]))
        if batch.n_nodes != w.shape[0]:
            raise ValueError("weights have %d entries for %d vertices" % (w.shape[0], batch.n_nodes))
        if (w == 0).any():
            x0 = np.where(w != 0, np.float32(1.0 / self.feature_size), np.float32(0)).astype(np.float32)
            batch.set_x0(x0)
        return State(batch=batch, n=batch.n_nodes)

    def predict(self, state):
        """(act_values [N, diver_num] float32, action = argmax over the vertex axis), mwis_dqn_call.py:140-143."""
        act_values = self.model.run(state["batch"])
        return act_values, self.model.pred

    def act(self, state):
        act_values, action = self.predict(state)
        return action

    def solve_mwis(self, adj_0, wts_0, train=False):
        """GCN-scored local greedy MWIS of one graph (mwis_dqn_call.py:198-261).  Returns
        (set of original vertex ids, total weight, 1.0)."""
        if train:
            raise NotImplementedError("train=True (exploration + replay memory) is out of scope")
        wts_0 = np.asarray(wts_0, dtype=np.float64).reshape(-1)
        if (wts_0 < 0).any():
            raise ValueError("negative weights: the reference drops wts == 0 but keeps wts > 0 only "
                             "(mwis_dqn_call.py:203-204), so its behaviour is undefined here")
        # one native call: the matrix's own indptr / indices go to the library as they are (dg_solve_graphs_host)
        member, total = self._solve_graphs([adj_0], wts_0)
        return set(np.flatnonzero(member).tolist()), float(total[0]), 1.0

    def solve_mwis_batch(self, graphs, wts):
        """Many graphs in one launch.  `graphs`: PackedBatch, GraphTables or list of adjacency matrices (the
        reference's native form - packed inside the library, no Python loop); `wts`: one weight per vertex (one array,
        or a list of per-graph arrays).  Returns (member uint8 [n_nodes], total weight per graph)."""
        if isinstance(graphs, PackedBatch):
            return self._solve_packed(graphs, np.asarray(wts, dtype=np.float64).reshape(-1))
        return self._solve_graphs(graphs, wts)

    def _solve_graphs(self, graphs, wts):
        model = self.model.compile(self.ctx)
        if model.out_width != 1:
            raise NotImplementedError("solve_mwis needs diver_num == 1 (act_vals.flatten() * wts, mwis_dqn_call.py:232)")
        return engine.solve_graphs_host(self.ctx, model, graphs, wts, predict=self.flags.predict, remove_zero_weight=True,
                                        check_values=self.check_values)

    def _solve_packed(self, packed, wts):
        model = self.model.compile(self.ctx)
        if model.out_width != 1:
            raise NotImplementedError("solve_mwis needs diver_num == 1 (act_vals.flatten() * wts, mwis_dqn_call.py:232)")
        return engine.solve_host(self.ctx, model, packed, wts, predict=self.flags.predict, remove_zero_weight=True)


# ---- module-level singleton, created on first access (the reference builds it at import, :344) ----
_dqn_agent = None


def __getattr__(name):
    global _dqn_agent
    if name == "dqn_agent":
        if _dqn_agent is None:
            _dqn_agent = DQNAgent(FLAGS.feature_size, 5000)
        return _dqn_agent
    raise AttributeError(name)
