"""Multi-channel wireless scheduling (BASELINE configs[2]): graph construction on the CPU, the batched slot
loop on the GPU against the per-instance CPU restatement of wireless_dqn_test_mc.py:225-366."""
import numpy as np
import pytest
import scipy.sparse as sp

from tests import util


def test_multichannel_conflict_graph_structure():
    from distgcn_b200 import wireless as W
    rng = np.random.default_rng(3)
    links, adj_i, xys = W.poisson_link_conflict_graph(rng)
    nf = adj_i.shape[0]
    assert nf > 10 and links.shape == (nf, 2)
    assert (adj_i != adj_i.T).nnz == 0 and adj_i.diagonal().sum() == 0
    # links are node pairs within r_c; links sharing a node conflict
    d = np.linalg.norm(xys[links[:, 0]] - xys[links[:, 1]], axis=1)
    assert (d <= W.SIM_RC).all()
    a = adj_i.toarray()
    for i in range(nf):
        for j in range(i + 1, nf):
            if set(links[i]) & set(links[j]):
                assert a[i, j] == 1
    K = 3
    adj_list = W.multichannel_conflict_simulate(adj_i, K, 0.8, rng)
    kept = [m.nnz / max(adj_i.nnz, 1) for m in adj_list]
    assert all(0.6 < k < 0.95 for k in kept)
    for m in adj_list:
        assert (m != m.T).nnz == 0 and (m.multiply(adj_i) != m).nnz == 0   # a sub-graph of the base conflict graph
    gK = W.multichannel_conflict_graph(adj_list).toarray()
    assert gK.shape == (K * nf, K * nf) and (gK == gK.T).all() and np.trace(gK) == 0
    for k1 in range(K):
        for k2 in range(K):
            blk = gK[k1 * nf:(k1 + 1) * nf, k2 * nf:(k2 + 1) * nf]
            if k1 == k2:
                assert (blk == adj_list[k1].toarray()).all()       # wireless_rollout_test_flood.py:124-130
            else:
                assert (blk == np.eye(nf)).all()                    # single-radio clique, :114-122


def test_traffic_follows_the_reference_draw_order():
    from distgcn_b200 import wireless as W
    arr, rates = W.traffic(treeseed=4, load=0.3, nflows=7, n_ch=3, timeslots=50)
    assert arr.shape == (50, 7) and rates.shape == (50, 7, 3)
    assert rates.min() >= 0 and rates.max() <= 100 and rates.dtype.kind == "i"
    assert (arr >= 0).all() and (arr == np.round(arr)).all()
    # same numbers as drawing from the legacy global generator, as the reference does (:178-204)
    np.random.seed(4)
    ia = np.random.exponential(1.0 / 15.0, (7, int(2 * 50 * 15.0)))
    at = np.cumsum(ia, axis=1)
    acc = np.stack([np.count_nonzero(at < t, axis=1) for t in range(50)], axis=1)
    assert np.array_equal(arr, np.diff(acc, prepend=0).T)
    lr = np.random.normal(50.0, 25.0, size=[50, 7, 3]).astype(int)
    assert np.array_equal(rates, np.clip(lr, 0, 100))
    assert abs(arr[1:].mean() - 15.0) < 2.0


@pytest.mark.gpu
@pytest.mark.parametrize("algo", ["Greedy", "Greedy-Th", "DGCN-LGS", "DGCN-LGS-it", "LGS-Seq", "DGCN-LGS-Seq"])
@pytest.mark.parametrize("ck", ["is4sat_l1", "is4sat_l20_c32"])
def test_batched_slot_loop_matches_per_instance_restatement(algo, ck):
    from distgcn_b200 import engine as E
    from distgcn_b200 import wireless as W
    from oracle import wireless_oracle as WO
    if algo in ("Greedy", "Greedy-Th", "LGS-Seq") and ck != "is4sat_l1":
        pytest.skip("no model involved")
    n_slots = 14
    insts = W.make_instances(n_networks=3, loads=[0.2, 0.9], n_ch=3, timeslots=n_slots + 1, seed=11, n_nodes=60,
                             area=150.0)
    layers = util.load_layers(ck)
    ctx = E.Context(0)
    model = E.Model(ctx, layers, E.gcn_dqn_acts(len(layers)))
    sim = W.BatchedScheduler(ctx, insts, algo, model)
    qs = sim.run()                       # device-resident loop: one synchronisation at the end of the sweep
    assert qs.shape == (n_slots, sim.n_links) and sim.solver_calls == n_slots * (3 if sim.seq else 1)
    for k, inst in enumerate(insts):
        q_ref, _ = WO.run_instance(inst.adj_list, inst.adj_gK, inst.arrivals, inst.rates, algo, layers, n_slots=n_slots)
        l0, l1 = int(sim.lp[k]), int(sim.lp[k + 1])
        assert np.array_equal(qs[:, l0:l1], q_ref[1:]), "instance %d: queue trajectories differ" % k
    assert qs.sum() > 0
    sim.close()
    # the host-side loop (numpy bookkeeping, one synchronous solve per slot) walks the same trajectory; run() in two
    # pieces continues where it stopped
    sim_h = W.BatchedScheduler(ctx, insts, algo, model)
    assert np.array_equal(sim_h.run_host(), qs)
    sim_h.close()
    sim_2 = W.BatchedScheduler(ctx, insts, algo, model)
    first = sim_2.run(5)
    rest = sim_2.run()
    assert np.array_equal(np.concatenate([first, rest]), qs)
    with pytest.raises(RuntimeError):
        sim_2.t -= 1
        sim_2.run(1)
    sim_2.close()
    model.close()
    ctx.close()


@pytest.mark.gpu
def test_agent_generation_is_an_explicit_choice():
    """DGCN-LGS on the joint graph with the generation-2 agent (no zero-weight removal, all-ones features): trajectories
    equal the restatement's for that generation, and differ from generation 1 once queues run empty."""
    from distgcn_b200 import engine as E
    from distgcn_b200 import wireless as W
    from oracle import wireless_oracle as WO
    n_slots = 12
    insts = W.make_instances(n_networks=2, loads=[0.05, 0.3], n_ch=3, timeslots=n_slots + 1, seed=3, n_nodes=60, area=150.0)
    layers = util.load_layers("is4sat_l20_c32")
    ctx = E.Context(0)
    model = E.Model(ctx, layers, E.gcn_dqn_acts(len(layers)))
    out = {}
    for gen in (1, 2):
        sim = W.BatchedScheduler(ctx, insts, "DGCN-LGS", model, agent_generation=gen)
        qs = sim.run()
        for k, inst in enumerate(insts):
            q_ref, _ = WO.run_instance(inst.adj_list, inst.adj_gK, inst.arrivals, inst.rates, "DGCN-LGS", layers,
                                       n_slots=n_slots, agent_generation=gen)
            l0, l1 = int(sim.lp[k]), int(sim.lp[k + 1])
            assert np.array_equal(qs[:, l0:l1], q_ref[1:]), "generation %d, instance %d" % (gen, k)
        out[gen] = qs
        sim.close()
    assert not np.array_equal(out[1], out[2])
    with pytest.raises(ValueError):
        W.BatchedScheduler(ctx, insts, "DGCN-LGS", model, agent_generation=3)
    model.close()
    ctx.close()


def test_slot_loop_restatement_schedules_are_independent_sets():
    """CPU only: every scheduler of the per-instance restatement (oracle/wireless_oracle.py) returns conflict-free
    schedules on the joint graph and never serves more than a queue holds; Greedy-Th (dist_greedy_search, epsilon 0.1)
    and Greedy (local_greedy_search) both drain traffic."""
    from distgcn_b200 import wireless as W
    from oracle import wireless_oracle as WO
    layers = util.load_layers("is4sat_l1")
    inst = W.make_instances(n_networks=1, loads=[0.6], n_ch=3, timeslots=9, seed=5, n_nodes=40, area=100.0)[0]
    gK = inst.adj_gK.tocsr()
    for algo in ("Greedy", "Greedy-Th", "DGCN-LGS", "DGCN-LGS-it", "LGS-Seq", "DGCN-LGS-Seq"):
        q, sched = WO.run_instance(inst.adj_list, inst.adj_gK, inst.arrivals, inst.rates, algo, layers, n_slots=8)
        assert q.shape == (9, inst.nflows) and (q >= 0).all()
        served = 0
        for s in sched:
            if algo.endswith("-Seq"):
                for ic in range(inst.n_ch):       # per-channel schedules are independent in their channel's graph
                    loc = s[(s >= ic * inst.nflows) & (s < (ic + 1) * inst.nflows)] - ic * inst.nflows
                    a = inst.adj_list[ic].tocsr()
                    assert a[loc][:, loc].nnz == 0
            else:
                assert gK[s][:, s].nnz == 0, algo
            served += len(s)
        assert served > 0, algo
