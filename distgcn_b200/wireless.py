"""Batched multi-channel wireless scheduling on the GPU solver (BASELINE.json configs[2]; SURVEY.md 8d
"Config 3", 8f rank 4): the per-time-slot loop of the reference's ``wireless_dqn_test_mc.py`` with every
(network, load) instance of a sweep advanced together, one solver launch per slot (or per channel for the
sequential variants) instead of one Python call per instance.

What is reproduced from the reference (file:line relative to the reference root):

* joint conflict graph of K channels: K copies of the link conflict graph plus a K-clique per link (single
  radio), vertex ``k * nflows + link`` (wireless_rollout_test_flood.py:98-133);
* per-channel conflict graphs: every conflict edge of the base graph survives on a channel with probability
  ``p_overlap`` (wireless_rollout_test_flood.py:71-95);
* traffic: exponential inter-arrivals of rate ``50 * load`` cumulated into per-slot arrival counts, link rates
  ``clip(int(N(50, 25)), 0, 100)`` per (slot, link, channel), drawn from ``RandomState(treeseed)`` in the reference's
  call order (wireless_dqn_test_mc.py:178-204);
* per slot (wireless_dqn_test_mc.py:225-366): ``q += arrivals``; weights ``q * r`` flattened channel-major
  (``order='F'``); schedule by one of
    - ``"Greedy"``        local greedy search on the joint graph                               (:242-248)
    - ``"DGCN-LGS"``      DQNAgent.solve_mwis on the joint graph                               (:278-291)
    - ``"LGS-Seq"``       per channel: local greedy search on the non-zero-weight sub-graph,
                          queue estimate reduced between channels                             (:292-312)
    - ``"DGCN-LGS-Seq"``  the same with DQNAgent.solve_mwis per channel                        (:313-333)
  then ``capacity[link] = rate of its scheduled vertex``, ``departures = min(q, capacity)``, ``q -= departures``
  (:358-366).

The shipped repository has no ``data/wireless_test`` and no ``graph_util`` module, so the *networks* here are a
synthetic stand-in with the constants of wireless_dqn_test_mc.py:91-94 (100 nodes on a 250-area square, links =
node pairs within ``r_c``, two links conflict when they share a node or any two of their end points are within
``r_i``); everything downstream of the conflict graph follows the reference.

The schedules come from the CUDA library (dg_solve / dg_lgs on a resident packed batch).  ``run`` keeps the queue
bookkeeping on the device as well (dg_wireless_*: a sweep is a stream of kernel launches with one synchronisation at
the end); ``step`` / ``run_host`` do it in vectorised numpy with one synchronous solver call per slot.  Where the sequential variants
schedule one link on several channels the reference's ``capacity[schedule] = rates`` keeps whichever entry numpy
assigns last, in Python-set iteration order; here the highest channel wins (ascending vertex id, which is what
that iteration order is for small integers).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional, Sequence

import numpy as np
import scipy.sparse as sp

from . import engine
from .batch import PackedBatch, pack_graphs

SIM_AREA = 250.0      # wireless_dqn_test_mc.py:91
SIM_NODE = 100        # :92
SIM_RC = 1.0          # :93
SIM_RI = 4.0          # :94
P_OVERLAP = 0.8       # :96
RATE_HI = 100         # :98
RATE_LO = 0           # :99
ALGOS = ("Greedy", "Greedy-Th", "DGCN-LGS", "DGCN-LGS-it", "LGS-Seq", "DGCN-LGS-Seq")
GREEDY_TH_EPSILON = 0.1   # dist_greedy_search(adj_gK, wts, 0.1), wireless_dqn_test_mc.py:252


# ------------------------------------------------------------------------------------------------------
# networks
# ------------------------------------------------------------------------------------------------------
def poisson_link_conflict_graph(rng: np.random.Generator, n_nodes: int = SIM_NODE, area: float = SIM_AREA,
                                rc: float = SIM_RC, ri: float = SIM_RI):
    """Synthetic stand-in for one file of data/wireless_test: (links [nflows, 2], adj_i csr [nflows, nflows],
    xys).  Node placement as Data_Generation.py:61-69 (uniform on a sqrt(area) square)."""
    side = float(np.sqrt(area))
    xys = rng.uniform(0.0, side, (n_nodes, 2))
    d = np.sqrt(((xys[:, None, :] - xys[None, :, :]) ** 2).sum(-1))
    iu, ju = np.nonzero(np.triu(d <= rc, k=1))
    links = np.stack([iu, ju], axis=1).astype(np.int64)
    nflows = links.shape[0]
    if nflows == 0:
        return links, sp.csr_matrix((0, 0)), xys
    ends = links.reshape(-1)                       # [2 * nflows] end points
    near = d[np.ix_(ends, ends)] <= ri             # end-point proximity (shared node = distance 0)
    conf = near.reshape(nflows, 2, nflows, 2).any(axis=(1, 3))
    np.fill_diagonal(conf, False)
    return links, sp.csr_matrix(conf.astype(np.float64)), xys


def multichannel_conflict_simulate(adj_i, n_ch: int, p_overlap: float, rng: np.random.Generator) -> List[sp.csr_matrix]:
    """Per-channel conflict graphs: each undirected conflict edge is kept with probability p_overlap,
    independently per channel (wireless_rollout_test_flood.py:84-94)."""
    a = sp.triu(sp.csr_matrix(adj_i), k=1).tocoo()
    out = []
    for _ in range(n_ch):
        keep = rng.random(a.nnz) <= p_overlap      # the reference removes when rand() > p
        u, v = a.row[keep], a.col[keep]
        m = sp.coo_matrix((np.ones(2 * u.shape[0]), (np.concatenate([u, v]), np.concatenate([v, u]))),
                          shape=a.shape).tocsr()
        out.append(m)
    return out


def multichannel_conflict_graph(adj_list: Sequence) -> sp.csr_matrix:
    """Joint conflict graph adj_gK (wireless_rollout_test_flood.py:98-133): vertex k * nn + n is link n on
    channel k; the K copies of a link form a clique; channel k carries adj_list[k]."""
    nk = len(adj_list)
    nn = adj_list[0].shape[0]
    assert all(a.shape == (nn, nn) for a in adj_list)
    blocks = [[None] * nk for _ in range(nk)]
    eye = sp.identity(nn, format="csr")
    for k1 in range(nk):
        for k2 in range(nk):
            blocks[k1][k2] = sp.csr_matrix(adj_list[k1]) if k1 == k2 else eye
    m = sp.bmat(blocks, format="csr")
    m.data[:] = 1.0
    return m


def traffic(treeseed: int, load: float, nflows: int, n_ch: int, timeslots: int):
    """(arrival_pkts [T, nflows], link_rates [T, nflows, n_ch] int) exactly as wireless_dqn_test_mc.py:178-204
    draws them from the legacy global generator seeded with `treeseed`."""
    rs = np.random.RandomState(treeseed)
    arrival_rate = 0.5 * (RATE_LO + RATE_HI) * load
    interarrivals = rs.exponential(1.0 / arrival_rate, (nflows, int(2 * timeslots * arrival_rate)))
    arrival_time = np.cumsum(interarrivals, axis=1)
    acc_pkts = np.zeros((nflows, timeslots))
    for t in range(timeslots):
        acc_pkts[:, t] = np.count_nonzero(arrival_time < t, axis=1)
    arrival_pkts = np.diff(acc_pkts, prepend=0).transpose()
    link_rates = rs.normal(0.5 * (RATE_LO + RATE_HI), 0.25 * (RATE_HI - RATE_LO), size=[timeslots, nflows, n_ch])
    link_rates = link_rates.astype(int)
    link_rates[link_rates < RATE_LO] = RATE_LO
    link_rates[link_rates > RATE_HI] = RATE_HI
    return arrival_pkts, link_rates


@dataclass
class Instance:
    """One (network, load seed) run of the reference's double loop (wireless_dqn_test_mc.py:155,176)."""
    adj_list: List[sp.csr_matrix]     # per-channel conflict graphs
    adj_gK: sp.csr_matrix             # joint graph
    arrivals: np.ndarray              # [T, nflows]
    rates: np.ndarray                 # [T, nflows, n_ch]
    load: float
    treeseed: int

    @property
    def nflows(self) -> int:
        return self.adj_list[0].shape[0]

    @property
    def n_ch(self) -> int:
        return len(self.adj_list)


def make_instances(n_networks: int, loads: Sequence[float], n_ch: int = 3, timeslots: int = 200, seed: int = 0,
                   n_nodes: int = SIM_NODE, area: float = SIM_AREA) -> List[Instance]:
    rng = np.random.default_rng(seed)
    out = []
    for _ in range(n_networks):
        while True:
            _, adj_i, _ = poisson_link_conflict_graph(rng, n_nodes, area)
            if adj_i.shape[0] > 0:
                break
        adj_list = multichannel_conflict_simulate(adj_i, n_ch, P_OVERLAP, rng)
        adj_gK = multichannel_conflict_graph(adj_list)
        for i, load in enumerate(loads):
            arr, rates = traffic(i, float(load), adj_i.shape[0], n_ch, timeslots)
            out.append(Instance(adj_list, adj_gK, arr, rates, float(load), i))
    return out


# ------------------------------------------------------------------------------------------------------
# the batched slot loop
# ------------------------------------------------------------------------------------------------------
class BatchedScheduler:
    """All instances of a sweep advanced slot by slot; graphs stay resident on the device."""

    def __init__(self, ctx: engine.Context, instances: Sequence[Instance], algo: str = "DGCN-LGS",
                 model: Optional[engine.Model] = None, predict: str = "mwis", agent_generation: int = 1):
        """`agent_generation`: the agent "DGCN-LGS" stands for on the joint graph.  1 (default) =
        mwis_dqn_call.DQNAgent.solve_mwis, which drops zero-weight vertices before scoring (mwis_dqn_call.py:202-207) -
        the agent the shipped checkpoints belong to; 2 = mwis_gdpg_call.MWISSolver.solve_mwis(adj, wts, train, grd)
        (mwis_gdpg_call.py:200-235, the signature wireless_dqn_test_mc.py:289 calls at HEAD), which keeps them in the
        graph.  With empty queues the two see different graphs.  "DGCN-LGS-it" is generation 2 by definition; the
        sequential variants slice the non-zero sub-graph themselves (:299-301), so the generations coincide there."""
        if agent_generation not in (1, 2):
            raise ValueError("agent_generation must be 1 or 2")
        self.agent_generation = agent_generation
        if algo not in ALGOS:
            raise ValueError("algo must be one of %s" % (ALGOS,))
        if algo.startswith("DGCN") and model is None:
            raise ValueError("%s needs a model" % algo)
        self.ctx, self.algo, self.model, self.predict = ctx, algo, model, predict
        self.inst = list(instances)
        self.n_ch = self.inst[0].n_ch
        self.T = min(i.arrivals.shape[0] for i in self.inst)
        nfl = np.array([i.nflows for i in self.inst], dtype=np.int64)
        self.lp = np.concatenate([[0], np.cumsum(nfl)])            # link offsets
        self.n_links = int(self.lp[-1])
        self.q = np.zeros(self.n_links)                            # queue lengths, all instances
        self.seq = algo.endswith("-Seq")
        self._remove_zero = self.seq or agent_generation == 1
        if self.seq:
            # one packed batch per channel; vertex = link
            self.packed = [pack_graphs([i.adj_list[k] for i in self.inst]) for k in range(self.n_ch)]
        else:
            # joint graphs; vertex k * nflows + link  ->  (flat link id, channel)
            self.packed = [pack_graphs([i.adj_gK for i in self.inst])]
            gp = self.packed[0].graph_ptr.astype(np.int64)
            vloc = np.arange(gp[-1]) - np.repeat(gp[:-1], np.diff(gp))
            nf_v = np.repeat(nfl, np.diff(gp))
            self.v_ch = (vloc // nf_v).astype(np.int64)
            self.v_link = (np.repeat(self.lp[:-1], np.diff(gp)) + vloc % nf_v).astype(np.int64)
        self.batches = [engine.DeviceBatch(ctx, p) for p in self.packed]
        self.t = 0
        self._t_device = 0     # slots advanced by the device-resident loop
        self._ws = None        # (dg_wireless handle, weights buffer, membership buffer)
        self.solver_calls = 0
        self.last_weights = None   # what the solver saw in the last slot (for parity tests)
        self.last_member = None

    @property
    def graphs_per_slot(self) -> int:
        return len(self.inst) * (self.n_ch if self.seq else 1)

    def _rates_at(self, t: int) -> np.ndarray:
        """[n_links, n_ch] link rates of slot t for all instances."""
        return np.concatenate([i.rates[t] for i in self.inst], axis=0).astype(np.float64)

    def _solve(self, k: int, w: np.ndarray) -> np.ndarray:
        self.solver_calls += 1
        if self.algo == "DGCN-LGS-it":   # solve_mwis_dit on the joint graph (wireless_dqn_test_mc.py:264)
            return engine.solve_dit(self.ctx, self.model, self.batches[k], w, predict=self.predict, want_total=False).member
        if self.algo == "Greedy-Th":     # threshold distributed greedy on the joint graph (:252)
            return engine.dist_greedy(self.ctx, self.batches[k], w, epsilon=GREEDY_TH_EPSILON, want_steps=False).member
        if self.algo.startswith("DGCN"):
            return engine.solve(self.ctx, self.model, self.batches[k], w, predict=self.predict,
                                remove_zero_weight=self._remove_zero, want_total=False).member
        if self.seq:   # LGS on the sub-graph of non-zero weights (wireless_dqn_test_mc.py:300-303)
            self.batches[k].set_keep_from_weights(w)
        return engine.lgs(self.ctx, self.batches[k], w, want_steps=False).member

    def step(self) -> np.ndarray:
        """Advance one slot (t = 1 .. T-1 in the reference); returns the departures per link."""
        self.t += 1
        t = self.t
        if t >= self.T:
            raise StopIteration("all %d slots done" % self.T)
        self.q += np.concatenate([i.arrivals[t] for i in self.inst])           # :227
        r = self._rates_at(t)
        capacity = np.zeros(self.n_links)
        if not self.seq:
            w = self.q[self.v_link] * r[self.v_link, self.v_ch]                 # :230, :240 (order='F')
            member = self._solve(0, w)
            sel = np.flatnonzero(member)
            capacity[self.v_link[sel]] = r[self.v_link[sel], self.v_ch[sel]]    # :360-363
            self.last_weights, self.last_member = [w], [member]
        else:
            qest = self.q.copy()                                               # queue_mtx_algo[:, ic]
            self.last_weights, self.last_member = [], []
            for ic in range(self.n_ch):
                w = qest * r[:, ic]                                            # :298 / :319
                member = self._solve(ic, w)
                sel = np.flatnonzero(member)
                capacity[sel] = r[sel, ic]                                     # later channels overwrite
                self.last_weights.append(w)
                self.last_member.append(member)
                if ic + 1 < self.n_ch:
                    qest = qest.copy()
                    qest[sel] -= np.minimum(qest, r[:, ic])[sel]               # :306-309
        dep = np.minimum(self.q, capacity)                                     # :364
        self.q -= dep                                                          # :365
        return dep

    def run_host(self, n_slots: Optional[int] = None):
        """The slot loop with the queue bookkeeping in numpy and one synchronous solver call per slot (`step`)."""
        n = (self.T - 1 - self.t) if n_slots is None else min(n_slots, self.T - 1 - self.t)
        qs = np.zeros((n, self.n_links))
        for s in range(n):
            self.step()
            qs[s] = self.q
        return qs

    # ---- device-resident loop: queues, weights, capacities and the history stay on the GPU ------------------------
    def _device_state(self):
        if self._ws is not None:
            return self._ws
        import ctypes as C
        lib = self.ctx._lib
        arrivals = np.ascontiguousarray(np.concatenate([i.arrivals[:self.T] for i in self.inst], axis=1), dtype=np.float64)
        rates = np.ascontiguousarray(np.concatenate([i.rates[:self.T] for i in self.inst], axis=1), dtype=np.int32)
        if self.seq:
            v0 = nf = None
            n_vertices = self.n_links
        else:
            gp = self.packed[0].graph_ptr.astype(np.int64)
            nfl = np.diff(self.lp)
            v0 = np.ascontiguousarray(np.repeat(gp[:-1], nfl) + (np.arange(self.n_links) - np.repeat(self.lp[:-1], nfl)),
                                      dtype=np.int32)
            nf = np.ascontiguousarray(np.repeat(nfl, nfl), dtype=np.int32)
            n_vertices = int(gp[-1])
        h = C.c_void_p()
        engine.check(lib.dg_wireless_create(self.ctx.handle, self.n_links, self.n_ch, self.T, arrivals.ctypes.data,
                                            rates.ctypes.data, None if v0 is None else v0.ctypes.data,
                                            None if nf is None else nf.ctypes.data, n_vertices, C.byref(h)))
        w, member, q = C.c_void_p(), C.c_void_p(), C.c_void_p()
        engine.check(lib.dg_wireless_buffers(h, C.byref(w), C.byref(member), C.byref(q)))
        self._ws = (h, w, member)
        return self._ws

    def _solve_device(self, k: int, w, member) -> None:
        """The slot's scheduler on device buffers (DG_MEM_DEVICE: enqueue only)."""
        lib, ctx, b = self.ctx._lib, self.ctx, self.batches[k]
        self.solver_calls += 1
        pc = engine.predict_code(self.predict)
        if self.algo == "DGCN-LGS-it":
            engine.check(lib.dg_solve_dit(ctx.handle, self.model.handle, b.handle, w, pc, member, None, None, engine.MEM_DEVICE))
        elif self.algo == "Greedy-Th":
            import ctypes as C
            engine.check(lib.dg_dist_greedy(ctx.handle, b.handle, w, C.c_double(GREEDY_TH_EPSILON), member, None,
                                            engine.MEM_DEVICE))
        elif self.algo.startswith("DGCN"):
            engine.check(lib.dg_solve(ctx.handle, self.model.handle, b.handle, w, pc, 1 if self._remove_zero else 0, member,
                                      None, None, None, None, engine.MEM_DEVICE))
        else:
            if self.seq:
                engine.check(lib.dg_batch_set_keep_from_weights(b.handle, w, engine.MEM_DEVICE))
            engine.check(lib.dg_lgs(ctx.handle, b.handle, w, -1, member, None, None, None, None, None, engine.MEM_DEVICE))

    def run(self, n_slots: Optional[int] = None):
        """Run the remaining slots on the device - q += arrivals, weights, schedule, capacities, departures are all
        kernels on resident arrays, enqueued slot after slot without a host synchronisation - and return the
        queue-length matrix [slots, n_links] (one copy at the end).  Not to be mixed with `step` / `run_host`."""
        if self.t != self._t_device:
            raise RuntimeError("run() continues the device-resident loop; this scheduler has been stepped on the host")
        n = (self.T - 1 - self.t) if n_slots is None else min(n_slots, self.T - 1 - self.t)
        lib = self.ctx._lib
        h, w, member = self._device_state()
        t0 = self.t
        if n > 0:   # the whole sweep is one native call (dg_wireless_run): five enqueues per slot without Python in between
            import ctypes as C
            sched = (3 if self.algo == "DGCN-LGS-it" else 1 if self.algo == "Greedy-Th" else 2 if self.algo.startswith("DGCN")
                     else 0)
            handles = (C.c_void_p * len(self.batches))(*[b.handle for b in self.batches])
            engine.check(lib.dg_wireless_run(h, None if self.model is None else self.model.handle, handles, len(self.batches),
                                             sched, 1 if self.seq else 0, engine.predict_code(self.predict),
                                             1 if self._remove_zero else 0, C.c_double(GREEDY_TH_EPSILON), self.t + 1, n))
            self.solver_calls += n * (self.n_ch if self.seq else 1)
            self.t += n
        self._t_device = self.t
        hist = np.empty((self.T, self.n_links), dtype=np.float64)
        engine.check(lib.dg_wireless_read_history(h, hist.ctypes.data))
        self.q = hist[self.t].copy()
        return hist[t0 + 1:self.t + 1]

    def close(self) -> None:
        if self._ws is not None:
            self.ctx._lib.dg_wireless_destroy(self._ws[0])
            self._ws = None
        for b in self.batches:
            b.close()
