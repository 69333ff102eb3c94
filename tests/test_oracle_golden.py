"""Pin the CPU oracle against vectors produced by the REFERENCE's own code
(tests/golden/make_golden.py imported /root/reference/heuristics.py and gcn/utils.py)."""
import numpy as np
import pytest
import scipy.sparse as sp

from oracle import gcn_oracle as G
from oracle import lgs as L
from tests import util


def _per_graph(pb, fn):
    for g in range(pb.n_graphs):
        v0, v1 = int(pb.graph_ptr[g]), int(pb.graph_ptr[g + 1])
        yield g, v0, v1, pb.slice(g, g + 1)


@pytest.mark.parametrize("fixture", ["lgs_ref_small.npz", "lgs_ref_ties.npz"])
def test_lgs_oracle_matches_reference(fixture):
    ref = util.load_npz(fixture)
    if fixture == "lgs_ref_small.npz":
        pb, w = util.small_graphs()
    else:
        pb, w = util.packed_from_npz(ref), ref["weights"]
    for g, v0, v1, sub in _per_graph(pb, None):
        r = L.run(sub.row_ptr, sub.col_idx, w[v0:v1])
        assert np.array_equal(r.member, ref["member"][v0:v1]), "membership, graph %d" % g
        assert r.steps == int(ref["steps"][g])
        assert r.p2p == int(ref["p2p"][g])
        assert r.bst == int(ref["bst"][g])
        assert np.array_equal(r.oh_vec, ref["oh_vec"][v0:v1])
        for k in (1, 2):
            rk = L.run(sub.row_ptr, sub.col_idx, w[v0:v1], nstep=k)
            assert np.array_equal(rk.member, ref["member_n%d" % k][v0:v1])
            assert np.array_equal(rk.nb_is, ref["nbis_n%d" % k][v0:v1])


def test_lgs_oracle_batch_equals_single():
    pb, w = util.small_graphs()
    rb = L.run_batch(pb.graph_ptr, pb.row_ptr, pb.col_idx, w)
    ref = util.load_npz("lgs_ref_small.npz")
    assert np.array_equal(rb.member, ref["member"])
    assert np.array_equal(rb.steps, ref["steps"])
    assert np.array_equal(rb.p2p, ref["p2p"])
    assert np.array_equal(rb.bst, ref["bst"])


def test_lgs_oracle_mask_equals_removal():
    """A keep mask must act exactly like deleting the vertices (mwis_dqn_call.py:202-207)."""
    pb, w = util.small_graphs()
    rng = np.random.default_rng(5)
    for g in (0, 7, 30, 49):
        sub = pb.slice(g, g + 1)
        v0, v1 = int(pb.graph_ptr[g]), int(pb.graph_ptr[g + 1])
        keep = rng.random(v1 - v0) < 0.7
        a = sub.graph_adj(0)
        idx = np.where(keep)[0]
        ar = a[idx][:, idx].tocsr()
        r_removed = L.run(ar.indptr, ar.indices, w[v0:v1][idx])
        r_masked = L.run(sub.row_ptr, sub.col_idx, w[v0:v1], init_remain=keep.astype(np.uint8))
        assert np.array_equal(r_masked.member[idx], r_removed.member)
        assert r_masked.member[~keep].sum() == 0
        assert (r_masked.steps, r_masked.p2p, r_masked.bst) == (r_removed.steps, r_removed.p2p, r_removed.bst)


def test_lgs_oracle_empty_and_trivial():
    r = L.run(np.zeros(1, np.int64), np.zeros(0, np.int32), np.zeros(0))
    assert r.member.shape == (0,) and r.steps == 0
    r = L.run(np.array([0, 0]), np.zeros(0, np.int32), np.array([3.0]))
    assert r.member.tolist() == [1] and r.steps == 1 and r.p2p == 0 and r.bst == 2


def test_lgs_oracle_nan_hits_round_cap():
    a = sp.csr_matrix(np.array([[0, 1], [1, 0]], dtype=float))
    with pytest.raises(RuntimeError):
        L.run(a.indptr, a.indices, np.array([np.nan, 1.0]), max_rounds=50)


def test_supports_and_features_match_reference():
    ref = util.load_npz("supports_ref.npz")
    pb, w = util.small_graphs()
    for idx in (0, 4, 20, 25, 29, 49):
        adj = pb.graph_adj(idx)
        for k in (1, 2):
            mine = G.laplacian_supports(adj, k)
            assert len(mine) == k + 1
            for i, t in enumerate(mine):
                m = sp.csr_matrix(t)
                m.sort_indices()
                key = "g%d_k%d_t%d_" % (idx, k, i)
                assert np.array_equal(m.indptr, ref[key + "indptr"])
                assert np.array_equal(m.indices, ref[key + "indices"])
                assert np.array_equal(m.data, ref[key + "data"])  # bit-exact fp64
        wz = ref["g%d_wz" % idx]
        for F in (1, 32):
            f = G.features_gen1(wz, F).tocoo()
            coords = np.vstack((f.row, f.col)).transpose()
            order_m = np.lexsort((coords[:, 1], coords[:, 0]))
            rc = ref["g%d_F%d_feat_coords" % (idx, F)]
            order_r = np.lexsort((rc[:, 1], rc[:, 0]))
            assert np.array_equal(coords[order_m], rc[order_r])
            rv = ref["g%d_F%d_feat_vals" % (idx, F)][order_r]
            # the row sum is accumulated in a different order (csr vs the reference's lil matrix), so
            # fp64 values may differ in the last bit; what TensorFlow is fed is the fp32 cast
            assert np.allclose(f.data[order_m], rv, rtol=1e-15, atol=0)
            assert np.array_equal(f.data[order_m].astype(np.float32), rv.astype(np.float32))
            # the property the CUDA path relies on: every stored entry rounds to fp32(1/F)
            assert np.all(f.data.astype(np.float32) == np.float32(1.0 / F))


@pytest.mark.parametrize("short", list(util.CKPTS))
def test_gcn_oracle_reproduces_golden(short):
    """The stored oracle activations are reproducible (guards against silent numpy/scipy drift) and
    the memberships stored beside them came from the reference's LGS on those utilities."""
    gold = util.load_npz("gcn_oracle_small.npz")
    pb, w = util.small_graphs()
    layers = util.load_layers(short)
    for g in (0, 13, 26, 49):
        v0, v1 = int(pb.graph_ptr[g]), int(pb.graph_ptr[g + 1])
        score, u, member = util.oracle_solve_graph(pb.graph_adj(g), w[v0:v1], layers)
        ref_act = gold[short + "_act"][v0:v1]
        scale = max(np.abs(ref_act).max(), 1e-30)
        assert np.abs(score - ref_act).max() <= 2e-6 * scale
        if np.array_equal(score, ref_act):
            assert np.array_equal(member, gold[short + "_member"][v0:v1])


def test_l1_model_closed_form():
    """For num_layer == 1 and feature_size == 1 the network is act_i = w0 + w1 * (L.1)_i
     - an independent check of the oracle's layer algebra."""
    layers = util.load_layers("is4sat_l1")
    w0 = float(layers[0].weights[0][0, 0])
    w1 = float(layers[0].weights[1][0, 0])
    pb, w = util.small_graphs()
    for g in (3, 40):
        v0, v1 = int(pb.graph_ptr[g]), int(pb.graph_ptr[g + 1])
        adj = pb.graph_adj(g)
        score, _, _ = util.oracle_solve_graph(adj, w[v0:v1], layers)
        lap = G.laplacian_supports(adj, 1)[1]
        s = np.asarray(lap @ np.ones(v1 - v0))
        assert np.allclose(score, w0 + w1 * s, rtol=0, atol=2e-6)


def _dgs_checks(pb, g, w, member, alpha):
    """Defining properties of a dist_greedy_search result (heuristics.py:38-74), whatever the scan order: the set is
    independent and maximal."""
    a = pb.graph_adj(g)
    m = member.astype(bool)
    assert a[m][:, m].nnz == 0, "graph %d: adjacent members" % g
    covered = m | (np.asarray(a[:, m].sum(axis=1)).reshape(-1) > 0)
    assert covered.all(), "graph %d: not maximal" % g


@pytest.mark.parametrize("eps,tag", [(0.1, "0p1"), (0.5, "0p5")])
def test_dist_greedy_oracle_matches_reference(eps, tag):
    """The reference scans every round's candidates in Python-set order, the restatement in ascending vertex id: on
    the instances where no round has adjacent candidates (order-free) the memberships must be those of the reference's
    own dist_greedy_search; on the others both must be maximal independent sets."""
    ref = util.load_npz("dgs_ref.npz")
    pb, w = util.packed_from_npz(ref), ref["weights"]
    want = ref["member_eps" + tag]
    n_free = 0
    for g, v0, v1, sub in _per_graph(pb, None):
        member, rounds, order_free = L.dist_greedy(sub.row_ptr, sub.col_idx, w[v0:v1], eps)
        _dgs_checks(pb, g, w[v0:v1], member, 1.0 + eps / 3.0)
        _dgs_checks(pb, g, w[v0:v1], want[v0:v1], 1.0 + eps / 3.0)
        if order_free:
            n_free += 1
            assert np.array_equal(member, want[v0:v1]), "graph %d (%s)" % (g, ref["names"][g])
            assert abs(float(w[v0:v1][member.astype(bool)].sum()) - float(ref["total_eps" + tag][g])) < 1e-9
    assert n_free >= 20, "too few order-free instances pin the restatement: %d" % n_free


def test_dist_greedy_oracle_spread_weights_equal_local_greedy():
    """Weights further apart than alpha on every edge: every round's candidates are exactly the vertices that dominate
    their remaining neighbourhood, i.e. the rounds of local_greedy_search (heuristics.py:77-116)."""
    pb, _ = util.small_graphs()
    rng = np.random.default_rng(3)
    for g in (0, 13, 27, 41):
        sub = pb.slice(g, g + 1)
        n = sub.n_nodes
        w = 1.2 ** (2.0 * rng.permutation(n))
        member, rounds, order_free = L.dist_greedy(sub.row_ptr, sub.col_idx, w, 0.5)
        r = L.run(sub.row_ptr, sub.col_idx, w)
        assert order_free and np.array_equal(member, r.member) and rounds == r.steps
