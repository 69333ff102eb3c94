"""Drop-in for the inference entry points of the reference's generation-2 solver
(``mwis_gdpg_call.MWISSolver``: makestate / utility / schedule / solve_mwis, mwis_gdpg_call.py:82-97,147-235).

Differences from generation 1 that are reproduced: no zero-weight removal; features are all-ones
row-normalised when ``predict == 'mwis'`` and ``w / (max w + 1e-9)`` otherwise; ``solve_mwis`` accepts
``grd`` and returns (mwis, total_wt); the model is ``GCN2_DQN``.  Training, the target network, rollouts
and the iterative variants are out of scope.
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


DQNAgent = MWISSolver
