"""Drop-in for the inference entry points of the reference's generation-2 solver
(``mwis_gdpg_call.MWISSolver``: makestate / utility / schedule / solve_mwis, mwis_gdpg_call.py:82-97,147-235).

Differences from generation 1 that are reproduced: no zero-weight removal; features are all-ones
row-normalised when ``predict == 'mwis'`` and ``w / (max w + 1e-9)`` otherwise; ``solve_mwis`` accepts
``grd`` and returns (mwis, total_wt); the model is ``GCN2_DQN``; ``solve_mwis_dit`` (GCN inside the greedy
iteration) runs as one device-side loop.  Training, the target network, rollouts and the centralised
iterative variants are out of scope.
"""
from __future__ import annotations

import numpy as np

from . import ckpt, engine
from . import layers as L
from .batch import pack_graphs
from .models import GCN2_DQN
from .runtime import default_context
from .runtime_config import FLAGS


class MWISSolver(object):
    def __init__(self, input_flags=None, memory_size=5000, device=None, hidden_dim=None, num_layer=None, bias=True):
        self.flags = input_flags or FLAGS
        self.feature_size = self.flags.feature_size
        self.memory_size = memory_size
        self.ctx = default_context(device)
        self.placeholders = L.make_placeholders(1 + self.flags.max_degree, self.feature_size)
        self.model = GCN2_DQN(self.placeholders, hidden_dim=hidden_dim or self.flags.hidden1,
                              num_layer=num_layer or self.flags.num_layer, bias=bias, flags=self.flags)

    def load(self, name):
        prefix = ckpt.checkpoint_prefix(name)
        if prefix:
            self.model.load(prefix)
            print("loaded " + prefix)

    def makestate(self, adj, wts_nn):
        w = np.asarray(wts_nn, dtype=np.float64).reshape(-1)
        batch = engine.DeviceBatch(self.ctx, pack_graphs([adj]))
        if self.flags.predict != "mwis":
            # un-normalised w / (max w + 1e-9) in every feature column (mwis_gdpg_call.py:87-93)
            batch.set_x0((w / (np.amax(w) + 1e-9)).astype(np.float32))
        return {"batch": batch, "features_raw": None}

    def act(self, state, train=False):
        if train:
            raise NotImplementedError("exploration belongs to training, which is out of scope")
        out = self.model.run(state["batch"])
        return out, self.model.pred

    def predict(self, state):
        return self.act(state, False)

    def utility(self, adj_0, wts_0, train=False):
        state = self.makestate(adj_0, wts_0)
        act_vals, _ = self.act(state, train)
        return act_vals, state

    def _lgs_on(self, state, act_vals, wts):
        batch = state["batch"]
        util = engine.utility(self.ctx, batch, act_vals, wts, self.flags.predict)
        res = engine.lgs(self.ctx, batch, util, want_steps=False)
        return set(np.flatnonzero(res.member).tolist())

    def schedule(self, adj_0, wts_0, train=False):
        wts = np.asarray(wts_0, dtype=np.float64).reshape(-1)
        state = self.makestate(adj_0, wts)
        act_vals, _ = self.act(state, train)
        mwis = self._lgs_on(state, act_vals, wts)
        return mwis, np.sum(wts[list(mwis)]), state, act_vals

    def solve_mwis(self, adj_0, wts_0, train=False, grd=1.0):
        mwis, total_wt, _, _ = self.schedule(adj_0, wts_0, train)
        return mwis, total_wt


    def solve_mwis_dit(self, adj_0, wts_0, train=False, grd=1.0):
        """GCN embedded into the LGS iteration (mwis_gdpg_call.py:278-318): returns (mwis, best_IS_util).  The
        whole loop runs on the device (dg_solve_dit)."""
        if train:
            raise NotImplementedError("exploration belongs to training, which is out of scope")
        wts = np.asarray(wts_0, dtype=np.float64).reshape(-1)
        if (wts < 0).any():
            raise ValueError("negative weights: the loop's stopping rule (np.sum(wts_nn) <= 0, :296) is only "
                             "reproduced for non-negative weights")
        member, total = self.solve_mwis_dit_batch(pack_graphs([adj_0]), wts)
        return set(np.flatnonzero(member).tolist()), np.array([total[0]])

    def solve_mwis_dit_batch(self, graphs, wts):
        """Many graphs in one launch: (member uint8 [n_nodes], set weight per graph)."""
        from .batch import PackedBatch
        packed = graphs if isinstance(graphs, PackedBatch) else pack_graphs(graphs)
        if self.flags.predict != "mwis":
            raise NotImplementedError("solve_mwis_dit is built for predict == 'mwis' (constant input features)")
        model = self.model.compile(self.ctx)
        batch = engine.DeviceBatch(self.ctx, packed)
        try:
            r = engine.solve_dit(self.ctx, model, batch, np.asarray(wts, dtype=np.float64).reshape(-1),
                                 self.flags.predict)
        finally:
            batch.close()
        return r.member, r.total


DQNAgent = MWISSolver
