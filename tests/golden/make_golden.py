#!/usr/bin/env python
"""Generate the golden fixtures under tests/golden/ from the REFERENCE ITSELF.

Run in the build container, where /root/reference exists:

    python tests/golden/make_golden.py [--reference /root/reference] [--full]

The reference's ``heuristics.py`` and ``gcn/utils.py`` are imported unmodified (the three third-party
modules heuristics.py imports but the path never uses - dwave_networkx, igraph, pulp - are stubbed in
sys.modules).  Everything TensorFlow-bound cannot run here (TensorFlow is not installable), so the GCN
scores in the fixtures come from oracle/gcn_oracle.py and are marked as such ("parity unpinned", see
that module's header); the LGS memberships computed FROM those scores are produced by the
reference's own ``local_greedy_search``.

Nothing in tests/, smoke() or bench.py reads /root/reference at run time - only this script does.
"""
from __future__ import annotations

import argparse
import os
import shutil
import sys
import types
import warnings

import numpy as np
import scipy.io as sio
import scipy.sparse as sp

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

warnings.filterwarnings("ignore")

CKPTS = {
    # fixture name -> reference model directory
    "is4sat_l1": "result_IS4SAT_deep_ld1_c32_l1_cheb1_diver1_mwis_dqn",
    "is4sat_l20_c32": "result_IS4SAT_deep_ld1_c32_l20_cheb1_diver1_mwis_dqn",
    "is4sat_l2_c64": "result_IS4SAT_deep_ld1_c64_l2_cheb1_diver1_mwis_dqn",
    "dqnba_l20_c32": "result_DQNBA_deep_ld1_c32_l20_cheb1_diver1_mwis_dqn",
    "dqnmed_l1_bias": "result_DQNMED_deep_ld1_c16_l1_cheb1_diver1_mwis_dqn",
    "is4sat_ld32_l3_c32": "result_IS4SAT_deep_ld32_c32_l3_cheb1_diver1_mwis_dqn",
    "is4sat_l3_c16": "result_IS4SAT_deep_ld1_c16_l3_cheb1_diver1_mwis_dqn",
    "is4sat_l2_c8": "result_IS4SAT_deep_ld1_c8_l2_cheb1_diver1_mwis_dqn",
}


def import_reference(ref_root):
    for name in ("dwave_networkx", "igraph", "pulp"):
        if name not in sys.modules:
            mod = types.ModuleType(name)
            if name == "pulp":
                mod.GLPK = object
            sys.modules[name] = mod
    sys.path.insert(0, ref_root)
    import heuristics as ref_h  # noqa: E402  (the reference's own file)
    import gcn.utils as ref_u  # noqa: E402
    return ref_h, ref_u


def load_mat(path):
    m = sio.loadmat(path)
    adj = sp.csr_matrix(m["adj"])
    adj.sort_indices()
    w = np.asarray(m["weights"], dtype=np.float64).reshape(-1)
    return adj, w, float(m["greedy_utility"].reshape(-1)[0])


def pack(graphs):
    """graphs: list of (csr adjacency, weights).  Packed CSR with batch-global column ids."""
    gp = [0]
    rp = [np.zeros(1, dtype=np.int64)]
    ci = []
    ws = []
    nnz = 0
    for adj, w in graphs:
        n = adj.shape[0]
        rp.append(adj.indptr[1:].astype(np.int64) + nnz)
        ci.append(adj.indices.astype(np.int32) + gp[-1])
        ws.append(np.asarray(w, dtype=np.float64).reshape(-1))
        nnz += adj.nnz
        gp.append(gp[-1] + n)
    return dict(graph_ptr=np.asarray(gp, dtype=np.int64), row_ptr=np.concatenate(rp),
                col_idx=np.concatenate(ci) if ci else np.zeros(0, np.int32),
                weights=np.concatenate(ws) if ws else np.zeros(0))


def member_vec(s, n):
    v = np.zeros(n, dtype=np.uint8)
    if len(s):
        v[np.fromiter((int(x) for x in s), dtype=np.int64)] = 1
    return v


def ref_lgs_all(ref_h, adj, w):
    """Run every reference LGS variant on one graph; the spmatrix type is what the reference's
    row slicing expects."""
    a = sp.csr_matrix(adj)
    n = a.shape[0]
    s0, tot0 = ref_h.local_greedy_search(a, w)
    s1, tot1, steps = ref_h.local_greedy_search_count(a, w)
    s2, tot2, steps2, p2p, bst = ref_h.local_greedy_search_stats(a, w)
    s3, tot3, steps3, p2p3, bst3, oh = ref_h.local_greedy_search_overhead(a, w)
    assert s0 == s1 == s2 == s3 and steps == steps2 == steps3 and p2p == p2p3 and bst == bst3
    out = dict(member=member_vec(s0, n), total=float(tot0), steps=int(steps), p2p=int(p2p), bst=int(bst),
               oh_vec=np.asarray(oh, dtype=np.float64))
    for k in (1, 2):
        sk, totk, nbk = ref_h.local_greedy_search_nstep(a, w, nstep=k)
        out["member_n%d" % k] = member_vec(sk, n)
        out["nbis_n%d" % k] = member_vec(nbk, n)
    return out


def special_graphs():
    def from_edges(n, edges):
        if edges:
            r, c = zip(*edges)
        else:
            r, c = (), ()
        a = sp.coo_matrix((np.ones(len(edges)), (r, c)), shape=(n, n))
        a = ((a + a.T) > 0).astype(np.float64).tocsr()
        a.sort_indices()
        return a
    out = []
    out.append(("path7", from_edges(7, [(i, i + 1) for i in range(6)])))
    out.append(("star9", from_edges(9, [(0, i) for i in range(1, 9)])))
    out.append(("clique6", from_edges(6, [(i, j) for i in range(6) for j in range(i + 1, 6)])))
    out.append(("empty5", from_edges(5, [])))
    out.append(("single", from_edges(1, [])))
    out.append(("edge2", from_edges(2, [(0, 1)])))
    out.append(("cycle8", from_edges(8, [(i, (i + 1) % 8) for i in range(8)])))
    out.append(("bipart4x4", from_edges(8, [(i, 4 + j) for i in range(4) for j in range(4)])))
    return out



def make_dgs(ref_h):
    """dgs_ref.npz: the reference's own dist_greedy_search (heuristics.py:38-74) on the committed small
    graph set (file weights and tie-heavy variants) and the special graphs, epsilon 0.1 (the value every
    reference call site uses, e.g. wireless_dqn_test_mc.py:252) and 0.5 (the default).  The reference walks
    each round's candidate set in CPython set order; the fixtures keep every instance, the tests compare
    memberships on those the oracle reports as order-free and check the defining properties on the rest."""
    z = np.load(os.path.join(HERE, "graphs_small.npz"))
    gp, rp, ci, w_all = z["graph_ptr"], z["row_ptr"], z["col_idx"], z["weights"]
    rng = np.random.default_rng(77)
    inst = []
    names = []
    for g in range(len(gp) - 1):
        v0, v1 = int(gp[g]), int(gp[g + 1])
        e0, e1 = int(rp[v0]), int(rp[v1])
        n = v1 - v0
        adj = sp.csr_matrix((np.ones(e1 - e0), ci[e0:e1].astype(np.int64) - v0, rp[v0:v1 + 1].astype(np.int64) - e0),
                            shape=(n, n))
        w = w_all[v0:v1]
        inst.append((adj, w.copy()))
        names.append("g%d|file" % g)
        if g % 5 == 0:
            for vn, wv in (("int0to5", rng.integers(0, 6, n).astype(np.float64)),
                           ("withzeros", np.where(rng.random(n) < 0.25, 0.0, w)),
                           ("spread", np.exp(8.0 * rng.random(n))),
                           ("allequal", np.full(n, 0.75))):
                inst.append((adj, wv))
                names.append("g%d|%s" % (g, vn))
    for gname, adj in special_graphs():
        n = adj.shape[0]
        for vn, wv in (("equal", np.ones(n)), ("ramp", np.arange(n, dtype=np.float64)),
                       ("rramp", np.arange(n, 0, -1).astype(np.float64)), ("zeros", np.zeros(n)),
                       ("pow", 3.0 ** np.arange(n, dtype=np.float64))):
            inst.append((adj, wv))
            names.append("%s|%s" % (gname, vn))
    out = pack(inst)
    eps_list = (0.1, 0.5)
    for eps in eps_list:
        members, totals = [], []
        for adj, w in inst:
            s, tot = ref_h.dist_greedy_search(sp.csr_matrix(adj), w, eps)
            members.append(member_vec(s, adj.shape[0]))
            totals.append(float(tot))
        tag = ("%g" % eps).replace(".", "p")
        out["member_eps%s" % tag] = np.concatenate(members)
        out["total_eps%s" % tag] = np.asarray(totals)
    np.savez_compressed(os.path.join(HERE, "dgs_ref.npz"), names=np.asarray(names), eps=np.asarray(eps_list), **out)
    print("dgs_ref.npz: %d instances" % len(inst))


META_CKPTS = ("is4sat_l1", "is4sat_l2_c64", "is4sat_l20_c32", "dqnba_l20_c32", "dqnmed_l1_bias", "is4sat_ld32_l3_c32",
              "is4sat_l3_c16", "is4sat_l2_c8", "is4sat_l2_c1_cheb2", "is4sat_l1_c1_cheb2")
META_EXTRA = {  # checkpoints whose weights are not among the CKPTS data fixtures (cheb2: three supports [I, L, L^2])
    "is4sat_l2_c1_cheb2": "result_IS4SAT_deep_ld1_c1_l2_cheb2_diver1_mwis_dqn",
    "is4sat_l1_c1_cheb2": "result_IS4SAT_deep_ld1_c1_l1_cheb2_diver1_mwis_dqn",
}
META_COPY = ("is4sat_l1", "is4sat_l2_c64", "dqnmed_l1_bias", "is4sat_l3_c16")  # small .meta files kept as data fixtures
META_GRAPHS = (0, 6, 12, 18, 24, 25, 31, 37, 43, 49)


def make_meta(ref_u, ref_root):
    """meta_activations.npz: outputs of the reference's AS-TRAINED TensorFlow graphs (model/result_*/model.ckpt.meta),
    evaluated op by op with numpy by oracle/tf_meta.py - TensorFlow itself is not installable here.  Inputs are built
    exactly as DQNAgent.makestate does (mwis_dqn_call.py:129-138) with the reference's own preprocess_features /
    simple_polynomials, and fed through the reference's own construct_feed_dict4pred (gcn/utils.py:157-168).
    Per checkpoint and graph: model.outputs ([N, 1] float32), model.pred, the first layer's activation, and the
    wiring signature (compute ops in execution order).  Two weight variants per graph: the file's weights and the
    same with ~20 % zeros (empty feature rows, as makestate produces for zero weights)."""
    from distgcn_b200 import ckpt as ckpt_reader
    from oracle import tf_meta

    z = np.load(os.path.join(HERE, "graphs_small.npz"))
    gp, rp, ci, w_all = z["graph_ptr"], z["row_ptr"], z["col_idx"], z["weights"]
    rng = np.random.default_rng(4242)
    graphs = []
    for g in META_GRAPHS:
        v0, v1 = int(gp[g]), int(gp[g + 1])
        e0, e1 = int(rp[v0]), int(rp[v1])
        n = v1 - v0
        adj = sp.csr_matrix((np.ones(e1 - e0), ci[e0:e1].astype(np.int64) - v0, rp[v0:v1 + 1].astype(np.int64) - e0),
                            shape=(n, n))
        w = w_all[v0:v1].copy()
        wz = w.copy()
        wz[rng.random(n) < 0.2] = 0.0
        graphs.append((g, adj, w, wz))
    out = {"graphs": np.asarray(META_GRAPHS)}
    out["wz"] = np.concatenate([wz for _, _, _, wz in graphs])
    meta_dir = os.path.join(HERE, "meta")
    os.makedirs(meta_dir, exist_ok=True)
    for short in META_CKPTS:
        d = CKPTS.get(short) or META_EXTRA[short]
        src = os.path.join(ref_root, "model", d)
        sg = tf_meta.load_stored_graph(src, ckpt_reader.read_tensors)
        roles = sg.roles()
        n_sup = len(roles["support"])
        sig = sg.signature()
        first_act = [n.name for n in sg.forward_nodes() if n.op == "LeakyRelu"]
        first_act = first_act[0] if first_act else sg.outputs
        F = int(sg.variables[[k for k in sg.variables if k.endswith("graphconvolution_1_vars/weights_0")][0]].shape[0])
        outs, outs_z, preds, h1 = [], [], [], []
        for g, adj, w, wz in graphs:
            for which, wts_nn in (("w", w), ("wz", wz)):
                n = wts_nn.shape[0]
                # ---- DQNAgent.makestate, mwis_dqn_call.py:129-138 (reference utilities, unmodified)
                norm_wts = np.linalg.norm(wts_nn)
                features = np.multiply(np.ones([n, F]), wts_nn.reshape(n, 1) / norm_wts)
                features = sp.lil_matrix(features)
                features = ref_u.preprocess_features(features)
                support = ref_u.simple_polynomials(sp.csr_matrix(adj), n_sup - 1)
                # ---- DQNAgent.predict, mwis_dqn_call.py:140-143
                feed = ref_u.construct_feed_dict4pred(features, support, roles)
                o, p, h = sg.run([sg.outputs, sg.pred, first_act], tf_meta.expand_feed(feed))
                assert o.dtype == np.float32 and o.shape == (n, 1)
                # the dropout sub-graph is live in the stored graph; at rate 0 it must not depend on the draws
                o2, = sg.run([sg.outputs], tf_meta.expand_feed(feed), rng=np.random.default_rng(g + 1))
                assert np.array_equal(o, o2)
                if which == "w":
                    outs.append(o[:, 0]), preds.append(int(p[0])), h1.append(np.asarray(h, dtype=np.float32).reshape(n, -1))
                else:
                    outs_z.append(o[:, 0])
        out["%s_outputs" % short] = np.concatenate(outs)
        out["%s_outputs_wz" % short] = np.concatenate(outs_z)
        out["%s_pred" % short] = np.asarray(preds)
        out["%s_h1" % short] = np.concatenate(h1)
        out["%s_signature" % short] = np.asarray(sig)
        out["%s_alphas" % short] = np.asarray(sg.leaky_alphas())
        out["%s_dir" % short] = np.asarray(d)
        if short in META_EXTRA:
            for name, arr in sg.variables.items():
                if "/Adam" not in name and "graphconvolution" in name:
                    out["%s_var|%s" % (short, name)] = arr
        if short in META_COPY:
            shutil.copyfile(os.path.join(src, "model.ckpt.meta"), os.path.join(meta_dir, d + ".meta"))
        print("meta %-22s %3d ops  last: %s" % (short, len(sig), sig[-2:]))
    np.savez_compressed(os.path.join(HERE, "meta_activations.npz"), **out)
    print("meta_activations.npz written")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reference", default="/root/reference")
    ap.add_argument("--full", action="store_true", help="also write the full 500-graph ER/BA test2 fixtures")
    ap.add_argument("--only", default=None, choices=["dgs", "meta"], help="regenerate just one fixture family")
    args = ap.parse_args()
    ref_h, ref_u = import_reference(args.reference)
    if args.only == "dgs":
        make_dgs(ref_h)
        return
    if args.only == "meta":
        make_meta(ref_u, args.reference)
        return
    from distgcn_b200 import ckpt as ckpt_reader
    from oracle import gcn_oracle as G

    rng = np.random.default_rng(20261017)

    # ---------------------------------------------------------------- checkpoints (data fixtures)
    ck_dir = os.path.join(HERE, "ckpt")
    os.makedirs(ck_dir, exist_ok=True)
    layers = {}
    for short, d in CKPTS.items():
        src = os.path.join(args.reference, "model", d)
        dst = os.path.join(ck_dir, d)
        os.makedirs(dst, exist_ok=True)
        for fn in ("checkpoint", "model.ckpt.index", "model.ckpt.data-00000-of-00001"):
            shutil.copyfile(os.path.join(src, fn), os.path.join(dst, fn))
        layers[short] = ckpt_reader.load_gcn_weights(src)

    # ---------------------------------------------------------------- small graph set: b0 of each cell
    small = []
    names = []
    greedy_util = []
    for fam in ("ER", "BA"):
        d = os.path.join(args.reference, "data", "%s_Graph_Uniform_GEN21_test2" % fam)
        files = sorted(f for f in os.listdir(d) if f.endswith("_b0_uni.mat"))
        assert len(files) == 25
        for f in files:
            adj, w, gu = load_mat(os.path.join(d, f))
            small.append((adj, w))
            names.append(f)
            greedy_util.append(gu)
    packed = pack(small)
    np.savez_compressed(os.path.join(HERE, "graphs_small.npz"), names=np.asarray(names),
                        greedy_utility=np.asarray(greedy_util), **packed)

    # greedy_search on the raw weights must reproduce the stored greedy_utility
    for (adj, w), gu in zip(small[:10], greedy_util[:10]):
        s, tot = ref_h.greedy_search(sp.csr_matrix(adj), w)
        assert abs(tot - gu) < 1e-9

    # ---------------------------------------------------------------- B: reference LGS, raw weights
    keys = ("member", "oh_vec", "member_n1", "nbis_n1", "member_n2", "nbis_n2")
    acc = {k: [] for k in keys}
    per_graph = {k: [] for k in ("steps", "p2p", "bst", "total")}
    for adj, w in small:
        r = ref_lgs_all(ref_h, adj, w)
        for k in keys:
            acc[k].append(r[k])
        for k in per_graph:
            per_graph[k].append(r[k])
    np.savez_compressed(os.path.join(HERE, "lgs_ref_small.npz"),
                        **{k: np.concatenate(v) for k, v in acc.items()},
                        **{k: np.asarray(v) for k, v in per_graph.items()})

    # ---------------------------------------------------------------- C: tie / edge-case suites
    tie_graphs = []
    tie_names = []
    sel = [0, 6, 12, 18, 24, 25, 31, 37, 43, 49]
    for idx in sel:
        adj, w = small[idx]
        n = adj.shape[0]
        variants = {
            "int0to5": rng.integers(0, 6, n).astype(np.float64),
            "allequal": np.full(n, 0.75),
            "withzeros": np.where(rng.random(n) < 0.25, 0.0, w),
            "signedzero": np.where(rng.random(n) < 0.5, 0.0, -0.0),
            "negatives": w - 0.5,
            "twovalues": np.where(rng.random(n) < 0.5, 1.0, 2.0),
            "float32grid": np.round(w * 8) / 8,
        }
        for vn, wv in variants.items():
            tie_graphs.append((adj, wv))
            tie_names.append("%s|%s" % (names[idx], vn))
    for gname, adj in special_graphs():
        n = adj.shape[0]
        for vn, wv in (("equal", np.ones(n)), ("ramp", np.arange(n, dtype=np.float64)),
                       ("rramp", np.arange(n, 0, -1).astype(np.float64)), ("zeros", np.zeros(n))):
            tie_graphs.append((adj, wv))
            tie_names.append("%s|%s" % (gname, vn))
    tp = pack(tie_graphs)
    acc = {k: [] for k in keys}
    per_graph = {k: [] for k in ("steps", "p2p", "bst", "total")}
    for adj, w in tie_graphs:
        r = ref_lgs_all(ref_h, adj, w)
        for k in keys:
            acc[k].append(r[k])
        for k in per_graph:
            per_graph[k].append(r[k])
    np.savez_compressed(os.path.join(HERE, "lgs_ref_ties.npz"), names=np.asarray(tie_names), **tp,
                        **{k: np.concatenate(v) for k, v in acc.items()},
                        **{k: np.asarray(v) for k, v in per_graph.items()})

    # ---------------------------------------------------------------- D: reference supports / features
    sup = {}
    for j, idx in enumerate([0, 4, 20, 25, 29, 49]):
        adj, w = small[idx]
        wz = w.copy()
        if j % 2 == 1:
            wz[rng.random(w.shape[0]) < 0.3] = 0.0
        for k in (1, 2):
            tk = ref_u.simple_polynomials(sp.csr_matrix(adj), k)
            for i, (coords, vals, shape) in enumerate(tk):
                m = sp.coo_matrix((vals, (coords[:, 0], coords[:, 1])), shape=shape).tocsr()
                m.sort_indices()
                sup["g%d_k%d_t%d_indptr" % (idx, k, i)] = m.indptr
                sup["g%d_k%d_t%d_indices" % (idx, k, i)] = m.indices
                sup["g%d_k%d_t%d_data" % (idx, k, i)] = m.data
        for F in (1, 32):
            n = w.shape[0]
            feats = np.multiply(np.ones([n, F]), wz.reshape(n, 1) / np.linalg.norm(wz))  # mwis_dqn_call.py:131-132
            coords, vals, shape = ref_u.preprocess_features(sp.lil_matrix(feats))
            sup["g%d_F%d_feat_coords" % (idx, F)] = coords
            sup["g%d_F%d_feat_vals" % (idx, F)] = vals
        sup["g%d_wz" % idx] = wz
    np.savez_compressed(os.path.join(HERE, "supports_ref.npz"), **sup)

    # ---------------------------------------------------------------- E: GCN (oracle) -> util -> reference LGS
    def solve_gen1(adj, w, lws, predict="mwis"):
        """The body of DQNAgent.solve_mwis (mwis_dqn_call.py:198-261) with the oracle standing in for
        sess.run and the reference's own LGS for the heuristic."""
        keep = np.where(w > 0)[0]
        a = sp.csr_matrix(adj)[keep][:, keep]
        wk = w[keep]
        F = lws[0].c_in
        feats = G.features_gen1(wk, F)
        supports = G.laplacian_supports(a, len(lws[0].weights) - 1)
        act = G.gcn_forward(feats, supports, lws, "gcn_dqn")
        util = G.utility(act[:, 0], wk, predict)
        s, _ = ref_h.local_greedy_search(sp.csr_matrix(a), util)
        member = np.zeros(w.shape[0], dtype=np.uint8)
        member[keep[np.fromiter((int(x) for x in s), dtype=np.int64)]] = 1
        act_full = np.zeros(w.shape[0], dtype=np.float32)
        act_full[keep] = act[:, 0]
        util_full = np.zeros(w.shape[0], dtype=np.float64)
        util_full[keep] = util
        return act_full, util_full, member

    e = {}
    for short, lws in layers.items():
        acts, utils, members = [], [], []
        for adj, w in small:
            a_, u_, m_ = solve_gen1(adj, w, lws)
            acts.append(a_), utils.append(u_), members.append(m_)
        e["%s_act" % short] = np.concatenate(acts)
        e["%s_util" % short] = np.concatenate(utils)
        e["%s_member" % short] = np.concatenate(members)
    # zero-weight removal cases (degree renormalisation on the reduced graph)
    wz_all = []
    for adj, w in small:
        wz = w.copy()
        wz[rng.random(w.shape[0]) < 0.2] = 0.0
        wz_all.append(wz)
    e["wz"] = np.concatenate(wz_all)
    for short in ("is4sat_l1", "is4sat_l20_c32", "is4sat_l2_c64"):
        acts, utils, members = [], [], []
        for (adj, _), wz in zip(small, wz_all):
            a_, u_, m_ = solve_gen1(adj, wz, layers[short])
            acts.append(a_), utils.append(u_), members.append(m_)
        e["%s_wz_act" % short] = np.concatenate(acts)
        e["%s_wz_util" % short] = np.concatenate(utils)
        e["%s_wz_member" % short] = np.concatenate(members)
    np.savez_compressed(os.path.join(HERE, "gcn_oracle_small.npz"), **e)

    # ---------------------------------------------------------------- F: full config-1 / config-2 sets
    if args.full:
        for fam, short in (("ER", "is4sat_l1"), ("BA", "is4sat_l20_c32")):
            d = os.path.join(args.reference, "data", "%s_Graph_Uniform_GEN21_test2" % fam)
            files = sorted(f for f in os.listdir(d) if f.endswith(".mat"))
            graphs, gus = [], []
            for f in files:
                adj, w, gu = load_mat(os.path.join(d, f))
                graphs.append((adj, w))
                gus.append(gu)
            p = pack(graphs)
            # store the pattern compactly: upper-triangular edges as local uint16 pairs
            eu, ev = [], []
            for adj, _ in graphs:
                c = sp.triu(adj, k=1).tocoo()
                eu.append(c.row.astype(np.uint16)), ev.append(c.col.astype(np.uint16))
            acts, members, lgs_raw = [], [], []
            for adj, w in graphs:
                a_, u_, m_ = solve_gen1(adj, w, layers[short])
                acts.append(a_), members.append(m_)
                s, _ = ref_h.local_greedy_search(sp.csr_matrix(adj), w)
                lgs_raw.append(member_vec(s, w.shape[0]))
            np.savez_compressed(os.path.join(HERE, "%s_test2_full.npz" % fam.lower()),
                                names=np.asarray(files), graph_ptr=p["graph_ptr"],
                                edge_ptr=np.cumsum([0] + [len(x) for x in eu]).astype(np.int64),
                                edge_u=np.concatenate(eu), edge_v=np.concatenate(ev), weights=p["weights"],
                                greedy_utility=np.asarray(gus), ckpt=np.asarray(CKPTS[short]),
                                oracle_act=np.concatenate(acts),
                                member_gcn_lgs=np.packbits(np.concatenate(members)),
                                member_raw_lgs=np.packbits(np.concatenate(lgs_raw)))
    make_meta(ref_u, args.reference)
    print("golden fixtures written to", HERE)


if __name__ == "__main__":
    main()
