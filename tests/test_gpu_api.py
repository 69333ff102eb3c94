"""GPU tests of the drop-in Python API (the reference's call sites): they read like the
reference's own usage - DQNAgent.load / solve_mwis, heuristics.local_greedy_search*, GraphConvolution."""
import numpy as np
import pytest
import scipy.sparse as sp

from tests import util

pytestmark = pytest.mark.gpu


def _flags(**kw):
    from distgcn_b200.runtime_config import make_flags
    base = dict(feature_size=1, hidden1=32, num_layer=20, diver_num=1, max_degree=1, predict="mwis")
    base.update(kw)
    return make_flags(**base)


def test_dqn_agent_solve_mwis_like_the_reference_scripts():
    from distgcn_b200.mwis_dqn_call import DQNAgent
    gold = util.load_npz("gcn_oracle_small.npz")
    pb, w = util.small_graphs()
    agent = DQNAgent(1, 5000, flags=_flags())
    agent.load(util.ckpt_dir("is4sat_l20_c32"))      # dqn_agent.load(find_model_folder(FLAGS, 'dqn'))
    for g in (0, 9, 30, 49):
        v0, v1 = int(pb.graph_ptr[g]), int(pb.graph_ptr[g + 1])
        adj = sp.csc_matrix(pb.graph_adj(g))          # the .mat files hold CSC matrices
        mwis, total_wt, one = agent.solve_mwis(adj, w[v0:v1])
        ref = set(np.flatnonzero(gold["is4sat_l20_c32_member"][v0:v1]).tolist())
        assert mwis == ref and one == 1.0
        assert abs(total_wt - w[v0:v1][sorted(ref)].sum()) < 1e-9
        # zero-weight vertices are removed and results come back in ORIGINAL vertex ids
        wz = gold["wz"][v0:v1]
        mwis_z, total_z, _ = agent.solve_mwis(adj, wz)
        ref_z = set(np.flatnonzero(gold["is4sat_l20_c32_wz_member"][v0:v1]).tolist())
        assert mwis_z == ref_z and all(wz[i] > 0 for i in mwis_z)
    member, total = agent.solve_mwis_batch(pb, w)
    assert np.array_equal(member, gold["is4sat_l20_c32_member"])
    with pytest.raises(NotImplementedError):
        agent.solve_mwis(sp.csr_matrix((2, 2)), np.ones(2), train=True)
    with pytest.raises(ValueError):
        agent.solve_mwis(sp.csr_matrix((2, 2)), np.array([1.0, -1.0]))


def test_dqn_agent_predict_and_makestate():
    from distgcn_b200.mwis_dqn_call import DQNAgent
    gold = util.load_npz("gcn_oracle_small.npz")
    pb, w = util.small_graphs()
    agent = DQNAgent(1, 5000, flags=_flags(num_layer=1))
    agent.load(util.ckpt_dir("is4sat_l1"))
    g = 12
    v0, v1 = int(pb.graph_ptr[g]), int(pb.graph_ptr[g + 1])
    state = agent.makestate(pb.graph_adj(g), w[v0:v1].reshape(-1, 1))
    act_values, action = agent.predict(state)
    assert act_values.shape == (v1 - v0, 1) and act_values.dtype == np.float32
    ref = gold["is4sat_l1_act"][v0:v1]
    assert np.abs(act_values[:, 0] - ref).max() <= 1e-5 * np.abs(ref).max()
    assert action.shape == (1,) and int(action[0]) == int(np.argmax(act_values[:, 0]))
    # a silent no-op when the directory has no `checkpoint` file, like the reference
    agent.load("/nonexistent/dir")


def test_heuristics_entry_points_match_reference_vectors():
    from distgcn_b200 import heuristics as H
    ref = util.load_npz("lgs_ref_small.npz")
    pb, w = util.small_graphs()
    for g in (1, 26, 44):
        v0, v1 = int(pb.graph_ptr[g]), int(pb.graph_ptr[g + 1])
        adj, wg = pb.graph_adj(g), w[v0:v1]
        want = set(np.flatnonzero(ref["member"][v0:v1]).tolist())
        mwis, total = H.local_greedy_search(adj, wg)
        assert mwis == want and abs(total - ref["total"][g]) < 1e-9
        mwis, total, step = H.local_greedy_search_count(adj, wg)
        assert mwis == want and step == ref["steps"][g]
        mwis, total, step, p2p, bst = H.local_greedy_search_stats(adj, wg)
        assert (step, p2p, bst) == (ref["steps"][g], ref["p2p"][g], ref["bst"][g])
        mwis, total, step, p2p, bst, oh = H.local_greedy_search_overhead(adj, wg)
        assert np.array_equal(oh, ref["oh_vec"][v0:v1])
        mwis1, _, nb1 = H.local_greedy_search_nstep(adj, wg, nstep=1)
        assert mwis1 == set(np.flatnonzero(ref["member_n1"][v0:v1]).tolist())
        assert nb1 == set(np.flatnonzero(ref["nbis_n1"][v0:v1]).tolist())
        # weights of any shape are flattened (heuristics.py:84)
        mwis2, _ = H.local_greedy_search(sp.csc_matrix(adj), wg.reshape(1, -1))
        assert mwis2 == want
    res = H.local_greedy_search_batch(pb, w, stats=True)
    assert np.array_equal(res.member, ref["member"]) and np.array_equal(res.p2p, ref["p2p"])
    # dist_greedy_search(adj, wts, epislon) as the wireless scripts call it (wireless_dqn_test_mc.py:252)
    from oracle import lgs as L
    dref = util.load_npz("dgs_ref.npz")
    dpb, dw = util.packed_from_npz(dref), dref["weights"]
    for g in (0, 3, 57):
        v0, v1 = int(dpb.graph_ptr[g]), int(dpb.graph_ptr[g + 1])
        sub = dpb.slice(g, g + 1)
        member, _, order_free = L.dist_greedy(sub.row_ptr, sub.col_idx, dw[v0:v1], 0.1)
        mwis, total = H.dist_greedy_search(dpb.graph_adj(g), dw[v0:v1], 0.1)
        assert mwis == set(np.flatnonzero(member).tolist())
        assert abs(total - float(dw[v0:v1][member.astype(bool)].sum())) < 1e-9
        if order_free:
            assert mwis == set(np.flatnonzero(dref["member_eps0p1"][v0:v1]).tolist())
    resd = H.dist_greedy_search_batch(dpb, dw, 0.5)
    assert resd.member.shape == (dpb.n_nodes,) and resd.steps.shape == (dpb.n_graphs,)


def test_graph_convolution_layer_call():
    from distgcn_b200 import engine as E
    from distgcn_b200 import layers as L
    from distgcn_b200.runtime import default_context
    from oracle import gcn_oracle as G
    pb, _ = util.small_graphs()
    sub = pb.slice(5, 8)
    ph = L.make_placeholders(2, 16)
    ph["batch"] = E.DeviceBatch(default_context(), sub)
    layer = L.GraphConvolution(input_dim=16, output_dim=24, placeholders=ph, act=lambda x: np.maximum(x, 0), bias=True,
                               flags=_flags(), rng=np.random.default_rng(3))
    layer.vars["bias"] = np.linspace(-0.1, 0.1, 24).astype(np.float32)
    x = np.random.default_rng(4).standard_normal((sub.n_nodes, 16)).astype(np.float32)
    y = layer(x)
    a = sp.csr_matrix((np.ones(sub.nnz), sub.col_idx, sub.row_ptr), shape=(sub.n_nodes, sub.n_nodes))
    sup = [G.to_fp32_csr(t) for t in G.laplacian_supports(a, 1)]
    ref = G.graph_convolution(x, sup, layer.weights, layer.vars["bias"], G.ACT_RELU)
    assert np.abs(y - ref).max() <= 1e-5 * np.abs(ref).max()
    ys = layer(sp.csr_matrix(x))  # sparse inputs are accepted like the first layer's
    assert np.array_equal(ys, y)


def test_gen2_solver_entry_points():
    from distgcn_b200.mwis_gdpg_call import MWISSolver
    from oracle import gcn_oracle as G
    from oracle import lgs as Lg
    pb, w = util.small_graphs()
    g = 33
    v0, v1 = int(pb.graph_ptr[g]), int(pb.graph_ptr[g + 1])
    adj, wg = pb.graph_adj(g), w[v0:v1]
    for predict in ("mwis", "mis"):
        solver = MWISSolver(_flags(feature_size=1, hidden1=16, num_layer=3, predict=predict), 5000, bias=True)
        rng = np.random.default_rng(7)
        for layer in solver.model.layers:
            layer.vars["bias"] = (0.05 * rng.standard_normal(layer.output_dim)).astype(np.float32)
        mwis, total = solver.solve_mwis(adj, wg, grd=1.0)
        feats = G.features_gen2(wg, 1, predict)
        sup = G.laplacian_supports(adj, 1)
        act = G.gcn_forward(feats, sup, solver.model.layers_as_weights(), "gcn2_dqn")
        util_ref = G.utility(act[:, 0], wg, predict)
        want = set(np.flatnonzero(Lg.run(adj.indptr, adj.indices, util_ref).member).tolist())
        assert mwis == want
        assert abs(total - wg[sorted(want)].sum()) < 1e-9
        act_vals, state = solver.utility(adj, wg)
        assert np.abs(act_vals - act).max() <= 1e-5 * max(np.abs(act).max(), 1e-30)


@pytest.mark.parametrize("fam", ["er", "ba"])
def test_greedy_search_reproduces_stored_greedy_utility(fam):
    """Every shipped .mat stores greedy_utility = greedy_search on the raw weights
    (Data_Generation.py:149-153,204); the weights are distinct U(0,1) draws."""
    from distgcn_b200 import heuristics as H
    pb, w, z = util.full_set(fam)
    res = H.local_greedy_search_batch(pb, w)
    tot = np.array([w[pb.graph_ptr[g]:pb.graph_ptr[g + 1]][res.member[pb.graph_ptr[g]:pb.graph_ptr[g + 1]] == 1].sum()
                    for g in range(pb.n_graphs)])
    assert np.allclose(tot, z["greedy_utility"], rtol=1e-12, atol=1e-12)
    g = 17
    v0, v1 = int(pb.graph_ptr[g]), int(pb.graph_ptr[g + 1])
    mwis, total = H.greedy_search(pb.graph_adj(g), w[v0:v1])
    assert abs(total - z["greedy_utility"][g]) < 1e-12


def test_host_pipeline_matches_one_call_at_a_time():
    """engine.HostPipeline (dg_solve_host_async on two contexts in turn) returns what dg_solve_host returns."""
    from distgcn_b200 import engine as E
    gold = util.load_npz("gcn_oracle_small.npz")
    pb, w = util.small_graphs()
    layers = util.load_layers("is4sat_l20_c32")
    pipe = E.HostPipeline(0, layers, E.gcn_dqn_acts(len(layers)), depth=2)
    subs = []
    for k, (g0, g1) in enumerate([(0, 50), (0, 17), (17, 50), (5, 6), (30, 50)]):
        sub = pb.slice(g0, g1)
        v0, v1 = int(pb.graph_ptr[g0]), int(pb.graph_ptr[g1])
        h = {n: E.pinned_empty(a.shape, a.dtype) for n, a in
             (("gp", sub.graph_ptr), ("rp", sub.row_ptr), ("ci", sub.col_idx), ("w", w[v0:v1]))}
        h["gp"][:], h["rp"][:], h["ci"][:], h["w"][:] = sub.graph_ptr, sub.row_ptr, sub.col_idx, w[v0:v1]
        member = E.pinned_empty(sub.n_nodes, np.uint8)
        total = E.pinned_empty(sub.n_graphs, np.float64)
        from distgcn_b200.batch import PackedBatch
        pipe.submit(PackedBatch(h["gp"], h["rp"], h["ci"]), h["w"], member, total)
        subs.append((v0, v1, g0, g1, member, total))
    pipe.wait()
    ref_member = gold["is4sat_l20_c32_member"]
    for v0, v1, g0, g1, member, total in subs:
        assert np.array_equal(np.asarray(member), ref_member[v0:v1])
        for g in range(g0, g1):
            a, b = int(pb.graph_ptr[g]), int(pb.graph_ptr[g + 1])
            assert abs(total[g - g0] - w[a:b][ref_member[a:b] != 0].sum()) < 1e-9
    assert pipe.launch_count >= len(subs)
    pipe.close()


def test_compact_host_format_matches_the_packed_one():
    """dg_solve_host_compact (16-bit graph-local column ids, expanded on the device) gives what dg_solve_host gives,
    one call at a time and through the pipeline; graphs above 65536 vertices are refused."""
    from distgcn_b200 import engine as E
    from distgcn_b200 import _lib
    from distgcn_b200.batch import PackedBatch, pack_graphs
    pb, w = util.small_graphs()
    c16 = pb.local_columns()
    assert c16.dtype == np.uint16 and c16.shape == pb.col_idx.shape
    base = np.repeat(pb.graph_ptr[:-1], pb.graph_nnz())
    assert np.array_equal(c16.astype(np.int64) + base, pb.col_idx)
    layers = util.load_layers("is4sat_l20_c32")
    ctx = E.Context(0)
    model = E.Model(ctx, layers, E.gcn_dqn_acts(len(layers)))
    m0, t0 = E.solve_host(ctx, model, pb, w)
    m1, t1 = E.solve_host(ctx, model, pb, w, col_local16=c16)
    assert np.array_equal(m0, m1) and np.array_equal(t0, t1)
    for short in ("is4sat_l1", "is4sat_l2_c64"):          # the other kernels read the expanded ids too
        lay = util.load_layers(short)
        mod = E.Model(ctx, lay, E.gcn_dqn_acts(len(lay)))
        a, _ = E.solve_host(ctx, mod, pb, w)
        b, _ = E.solve_host(ctx, mod, pb, w, col_local16=c16)
        assert np.array_equal(a, b)
        mod.close()
    pipe = E.HostPipeline(0, layers, E.gcn_dqn_acts(len(layers)), depth=2)
    outs = []
    for g0, g1 in ((0, 50), (3, 29), (29, 50)):
        sub = pb.slice(g0, g1)
        v0, v1 = int(pb.graph_ptr[g0]), int(pb.graph_ptr[g1])
        member = E.pinned_empty(sub.n_nodes, np.uint8)
        pipe.submit(sub, np.ascontiguousarray(w[v0:v1]), member, None, col_local16=sub.local_columns())
        outs.append((v0, v1, member))
    pipe.wait()
    for v0, v1, member in outs:
        assert np.array_equal(np.asarray(member), m0[v0:v1])
    pipe.close()
    with pytest.raises(TypeError):
        E.solve_host(ctx, model, pb, w, col_local16=c16.astype(np.int32))
    model.close()
    ctx.close()


def test_mwis_dqn_test_harness_ratios():
    """distgcn_b200.mwis_dqn_test.evaluate: the 500-file loop of mwis_dqn_test.py:304-348 as two launches; the
    normaliser equals the greedy_utility stored in the reference's own files and the ratios equal the ones
    from the reference-LGS memberships of the golden fixture."""
    from distgcn_b200.mwis_dqn_call import DQNAgent
    from distgcn_b200.mwis_dqn_test import evaluate
    pb, w, z = util.full_set("er")
    agent = DQNAgent(1, 5000, flags=_flags(num_layer=1))
    agent.load(util.ckpt_dir("is4sat_l1"))
    p, total, greedy_util, member = evaluate(agent, pb, w, search="local")
    assert np.allclose(greedy_util, z["greedy_utility"], rtol=1e-12, atol=1e-12)
    ref_member = np.unpackbits(z["member_gcn_lgs"])[:pb.n_nodes]
    ref_total = np.add.reduceat(np.where(ref_member == 1, w, 0.0), pb.graph_ptr[:-1])
    assert np.array_equal(member, ref_member)   # every graph, as test_full_config_sets asserts on the same set
    assert np.allclose(total, ref_total, rtol=1e-12)
    assert np.allclose(p, ref_total / z["greedy_utility"], rtol=1e-12)
    assert 1.0 < np.nanmean(p) < 1.12     # survey-time restatement: 1.041; Gurobi optimum of the set: 1.1197
    p2, total2, _, _ = evaluate(agent, pb, w, search="greedy")
    assert np.allclose(total2, total, rtol=1e-12)   # the weights of this set have no zeros: both searches solve the same graph


def test_native_ingest_list_of_scipy_matrices():
    """solve_mwis_batch(list of scipy matrices, weights): the reference's native input goes through
    dg_solve_graphs_host (packed by the library's host threads) and gives what the packed path gives - for one
    weight array or per-graph weight arrays, CSC / CSR / other formats, with zero weights (original vertex ids) and
    with stored zeros in a matrix; the streaming form (HostPipeline.submit_graphs) too."""
    from distgcn_b200 import engine as E
    from distgcn_b200.mwis_dqn_call import DQNAgent
    gold = util.load_npz("gcn_oracle_small.npz")
    pb, w = util.small_graphs()
    agent = DQNAgent(1, 5000, flags=_flags())
    agent.load(util.ckpt_dir("is4sat_l20_c32"))
    adjs = [sp.csc_matrix(pb.graph_adj(g)) for g in range(pb.n_graphs)]
    member, total = agent.solve_mwis_batch(adjs, w)
    assert np.array_equal(member, gold["is4sat_l20_c32_member"])
    ref_total = np.add.reduceat(np.where(member == 1, w, 0.0), pb.graph_ptr[:-1])
    assert np.allclose(total, ref_total, rtol=1e-12)
    # per-graph weight arrays, mixed matrix formats
    w_list = [w[pb.graph_ptr[g]:pb.graph_ptr[g + 1]] for g in range(pb.n_graphs)]
    mixed = [a.tocsr() if i % 3 == 0 else (a.tolil() if i % 3 == 1 else a) for i, a in enumerate(adjs)]
    member2, total2 = agent.solve_mwis_batch(mixed, w_list)
    assert np.array_equal(member2, member) and np.array_equal(total2, total)
    # zero weights
    wz = gold["wz"]
    member_z, _ = agent.solve_mwis_batch(adjs, wz)
    assert np.array_equal(member_z, gold["is4sat_l20_c32_wz_member"])
    # a stored zero is not an edge: same result as the matrix without those entries
    g = 7
    x = adjs[g].tocsr().copy()
    rows = np.repeat(np.arange(x.shape[0]), np.diff(x.indptr))
    drop = ((rows + x.indices) % 5 == 0)                  # a symmetric set of pairs
    x.data = np.where(drop, 0.0, x.data)
    y = x.copy()
    y.eliminate_zeros()
    assert 0 < y.nnz < x.nnz
    wg = w_list[g]
    m_x, t_x, _ = agent.solve_mwis(x, wg)
    m_y, t_y, _ = agent.solve_mwis(y, wg)
    m_full, _, _ = agent.solve_mwis(adjs[g], wg)
    assert m_x == m_y and t_x == t_y and m_x != m_full
    # streaming form
    pipe = E.HostPipeline(0, util.load_layers("is4sat_l20_c32"), E.gcn_dqn_acts(20), depth=2)
    outs = [(E.pinned_empty(pb.n_nodes, np.uint8), E.pinned_empty(pb.n_graphs, np.float64)) for _ in range(4)]
    for k in range(4):
        pipe.submit_graphs(adjs, w if k % 2 == 0 else wz, outs[k][0], outs[k][1])
    pipe.wait()
    for k in range(4):
        assert np.array_equal(np.asarray(outs[k][0]), member if k % 2 == 0 else member_z)
    # two producer threads, one per context of the pipeline (submit_graphs(slot = k))
    import threading
    outs2 = [(E.pinned_empty(pb.n_nodes, np.uint8), E.pinned_empty(pb.n_graphs, np.float64)) for _ in range(6)]

    def producer(k):
        for i in range(k, 6, 2):
            pipe.submit_graphs(adjs, w if i % 2 == 0 else wz, outs2[i][0], outs2[i][1], slot=k)
    threads = [threading.Thread(target=producer, args=(k,)) for k in range(2)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    pipe.wait()
    for i in range(6):
        assert np.array_equal(np.asarray(outs2[i][0]), member if i % 2 == 0 else member_z)
    pipe.close()
    # one graph per call, as the reference's scripts do (wireless_dqn_test_mc.py:289,323)
    for g in (3, 28):
        v0, v1 = int(pb.graph_ptr[g]), int(pb.graph_ptr[g + 1])
        mw, tw, _ = agent.solve_mwis(adjs[g], w_list[g])
        assert mw == set(np.flatnonzero(member[v0:v1]).tolist())


@pytest.mark.parametrize("short,nl", [("is4sat_l20_c32", 20), ("is4sat_l1", 1), ("is4sat_l2_c64", 2)])
def test_upper_host_format_matches_the_packed_one(short, nl, monkeypatch):
    """dg_solve_host_upper (column > row only) gives the memberships of dg_solve_host: through the tensor-core kernel
    (which builds its dense adjacency from the half lists) and through the device-side expansion to the full CSR (other
    models; DG_DISABLE_TC), with zero weights, empty graphs, a graph beyond the tensor-core kernel's limits, pipelined."""
    from distgcn_b200 import engine as E
    from distgcn_b200.batch import pack_graphs
    rng = np.random.default_rng(17)
    pb0, w0 = util.small_graphs()
    extra = []
    up = np.triu(rng.random((420, 420)) < 0.03, k=1)
    extra.append(sp.csr_matrix((up | up.T).astype(np.float64)))
    adjs = [pb0.graph_adj(g) for g in range(0, 50, 2)] + [sp.csr_matrix((0, 0)), sp.csr_matrix((4, 4))] + extra
    pb = pack_graphs(adjs)
    w = rng.random(pb.n_nodes)
    w[rng.random(pb.n_nodes) < 0.15] = 0.0
    upper = pb.upper_compact()
    layers = util.load_layers(short)
    ctx = E.Context(0)
    model = E.Model(ctx, layers, E.gcn_dqn_acts(len(layers)))
    ref_m, ref_t = E.solve_host(ctx, model, pb, w)
    kernel_full = ctx.last_kernel
    m1, t1 = E.solve_host(ctx, model, pb, w, upper=upper)
    assert np.array_equal(m1, ref_m) and np.allclose(t1, ref_t, rtol=1e-12)
    assert ctx.last_kernel == kernel_full
    monkeypatch.setenv("DG_DISABLE_TC", "1")       # every graph through the expansion + the CUDA-core kernel
    m2, _ = E.solve_host(ctx, model, pb, w, upper=upper)
    m2f, _ = E.solve_host(ctx, model, pb, w)
    assert np.array_equal(m2, m2f)
    monkeypatch.setenv("DG_DISABLE_FUSED", "1")    # ... and through the per-layer kernels
    m3, _ = E.solve_host(ctx, model, pb, w, upper=upper)
    m3f, _ = E.solve_host(ctx, model, pb, w)
    assert np.array_equal(m3, m3f)
    monkeypatch.delenv("DG_DISABLE_FUSED")
    monkeypatch.delenv("DG_DISABLE_TC")
    # a graph whose lists do not fit the shared-memory expansion (the global-memory one takes it), next to small ones
    nb = 6000
    src = rng.integers(0, nb, 30000)
    dst = rng.integers(0, nb, 30000)
    ok = src != dst
    big = sp.csr_matrix((np.ones(int(ok.sum())), (src[ok], dst[ok])), shape=(nb, nb))
    big = ((big + big.T) > 0).astype(np.float64).tocsr()
    pbb = pack_graphs([adjs[0], big, adjs[1]])
    wb = rng.random(pbb.n_nodes)
    wb[rng.random(pbb.n_nodes) < 0.1] = 0.0
    mb_full, _ = E.solve_host(ctx, model, pbb, wb)
    mb_up, _ = E.solve_host(ctx, model, pbb, wb, upper=pbb.upper_compact())
    assert np.array_equal(mb_full, mb_up)
    pipe = E.HostPipeline(0, layers, E.gcn_dqn_acts(len(layers)), depth=2)
    outs = [(E.pinned_empty(pb.n_nodes, np.uint8), E.pinned_empty(pb.n_graphs, np.float64)) for _ in range(4)]
    for k in range(4):
        pipe.submit(pb, w, outs[k][0], outs[k][1], upper=upper)
    pipe.wait()
    for k in range(4):
        assert np.array_equal(np.asarray(outs[k][0]), ref_m)
    pipe.close()
    model.close()
    ctx.close()


def test_native_ingest_upper_option(monkeypatch):
    """DG_INGEST_UPPER=1: dg_solve_graphs_host packs only the entries above the diagonal (filtered from the scipy arrays,
    stored zeros dropped) and gives the same memberships; an asymmetric matrix is refused instead of being mis-read."""
    from distgcn_b200 import _lib
    from distgcn_b200.mwis_dqn_call import DQNAgent
    pb, w = util.small_graphs()
    adjs = [sp.csc_matrix(pb.graph_adj(g)) for g in range(pb.n_graphs)]
    x = adjs[7].tocsr().copy()
    rows = np.repeat(np.arange(x.shape[0]), np.diff(x.indptr))
    x.data = np.where((rows + x.indices) % 5 == 0, 0.0, x.data)
    adjs[7] = x
    wz = w.copy()
    wz[::7] = 0.0
    for short in ("is4sat_l20_c32", "is4sat_l1"):
        agent = DQNAgent(1, 5000, flags=_flags(num_layer=20 if "l20" in short else 1))
        agent.load(util.ckpt_dir(short))
        ref = [agent.solve_mwis_batch(adjs, ww) for ww in (w, wz)]
        monkeypatch.setenv("DG_INGEST_UPPER", "1")
        for (m0, t0), ww in zip(ref, (w, wz)):
            m1, t1 = agent.solve_mwis_batch(adjs, ww)
            assert np.array_equal(m0, m1) and np.array_equal(t0, t1)
        skew = sp.csr_matrix(np.array([[0, 1, 1], [1, 0, 0], [0, 1, 0]], dtype=float))
        with pytest.raises(_lib.DistGCNError):
            agent.solve_mwis_batch([skew], np.ones(3))
        monkeypatch.delenv("DG_INGEST_UPPER")


def test_producer_threads_on_the_cuda_core_kernel():
    """Four producer threads, one per context, submit batches of different shapes through the CUDA-core graph-resident
    kernel at the same time (its shared-memory size depends on the batch: a per-call function attribute set by one thread
    used to undercut another thread's launch)."""
    import threading
    from distgcn_b200 import engine as E
    pb, w = util.small_graphs()
    layers = util.load_layers("is4sat_l1")
    acts = E.gcn_dqn_acts(len(layers))
    ctx = E.Context(0)
    model = E.Model(ctx, layers, acts)
    shapes = [(0, 50), (0, 7), (10, 40), (45, 50)]
    subs, refs = [], []
    for lo, hi in shapes:
        sub = pb.slice(lo, hi)
        ws = w[int(pb.graph_ptr[lo]):int(pb.graph_ptr[hi])].copy()
        subs.append(([sp.csr_matrix(sub.graph_adj(g)) for g in range(sub.n_graphs)], ws, sub))
        refs.append(E.solve_host(ctx, model, sub, ws)[0])
    pipe = E.HostPipeline(0, layers, acts, depth=4)
    errors = []
    outs = {}

    def producer(k):
        try:
            for i in range(12):
                j = (i + k) % len(shapes)
                adjs, ws, sub = subs[j]
                m = E.pinned_empty(sub.n_nodes, np.uint8)
                pipe.submit_graphs(adjs, ws, m, None, slot=k)
                pipe.wait(k)
                outs[(k, i)] = (j, np.asarray(m).copy())
        except Exception as e:   # noqa: BLE001
            errors.append(e)
    threads = [threading.Thread(target=producer, args=(k,)) for k in range(4)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    pipe.close()
    assert not errors, errors
    assert len(outs) == 48
    for (k, i), (j, m) in outs.items():
        assert np.array_equal(m, refs[j]), (k, i, j)
    model.close()
    ctx.close()
