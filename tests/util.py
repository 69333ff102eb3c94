"""Shared helpers for the test-suite: golden fixture loading and oracle glue."""
from __future__ import annotations

import os

import numpy as np
import scipy.sparse as sp

from distgcn_b200 import ckpt
from distgcn_b200.batch import PackedBatch, from_edge_lists

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")

CKPTS = {
    "is4sat_l1": "result_IS4SAT_deep_ld1_c32_l1_cheb1_diver1_mwis_dqn",
    "is4sat_l20_c32": "result_IS4SAT_deep_ld1_c32_l20_cheb1_diver1_mwis_dqn",
    "is4sat_l2_c64": "result_IS4SAT_deep_ld1_c64_l2_cheb1_diver1_mwis_dqn",
    "dqnba_l20_c32": "result_DQNBA_deep_ld1_c32_l20_cheb1_diver1_mwis_dqn",
    "dqnmed_l1_bias": "result_DQNMED_deep_ld1_c16_l1_cheb1_diver1_mwis_dqn",
    "is4sat_ld32_l3_c32": "result_IS4SAT_deep_ld32_c32_l3_cheb1_diver1_mwis_dqn",
    "is4sat_l3_c16": "result_IS4SAT_deep_ld1_c16_l3_cheb1_diver1_mwis_dqn",
    "is4sat_l2_c8": "result_IS4SAT_deep_ld1_c8_l2_cheb1_diver1_mwis_dqn",
}


def ckpt_dir(short: str) -> str:
    return os.path.join(GOLDEN, "ckpt", CKPTS[short])


def load_layers(short: str):
    return ckpt.load_gcn_weights(ckpt_dir(short))


def load_npz(name: str):
    return np.load(os.path.join(GOLDEN, name), allow_pickle=False)


def packed_from_npz(z) -> PackedBatch:
    return PackedBatch(graph_ptr=z["graph_ptr"].astype(np.int32), row_ptr=z["row_ptr"].astype(np.int32),
                       col_idx=z["col_idx"].astype(np.int32))


def small_graphs():
    z = load_npz("graphs_small.npz")
    return packed_from_npz(z), z["weights"].astype(np.float64)


def full_set(fam: str):
    """('er'|'ba') -> (PackedBatch, weights, fixture dict)."""
    z = load_npz("%s_test2_full.npz" % fam)
    pb = from_edge_lists(z["graph_ptr"], z["edge_ptr"], z["edge_u"], z["edge_v"])
    return pb, z["weights"].astype(np.float64), z


def graph_csr(pb: PackedBatch, g: int) -> sp.csr_matrix:
    return pb.graph_adj(g)


def oracle_solve_graph(adj, w, layers, predict: str = "mwis", kind: str = "gcn_dqn"):
    """Oracle restatement of DQNAgent.solve_mwis for one graph (oracle/pipeline.py)."""
    from oracle import pipeline
    return pipeline.solve_graph(adj, w, layers, predict, kind)


def random_graph_batch(rng, n_graphs, n_lo, n_hi, p_lo=0.02, p_hi=0.15):
    """Synthetic ER batch (numpy only)."""
    from distgcn_b200.batch import pack_graphs
    adjs = []
    for _ in range(n_graphs):
        n = int(rng.integers(n_lo, n_hi + 1))
        p = float(rng.uniform(p_lo, p_hi))
        upper = np.triu(rng.random((n, n)) < p, k=1)
        a = sp.csr_matrix((upper | upper.T).astype(np.float64))
        adjs.append(a)
    return pack_graphs(adjs), adjs


def exact_scores(pb, w, layers, kind: str = "gcn_dqn", remove_zero_weight: bool = True) -> np.ndarray:
    """float64 evaluation of the network on every graph of a batch (oracle.gcn_forward_fp64), zeros on
    removed vertices - the yardstick the fp32 implementations are measured against."""
    from oracle import gcn_oracle as G
    out = np.zeros(pb.n_nodes, dtype=np.float64)
    for g in range(pb.n_graphs):
        v0, v1 = int(pb.graph_ptr[g]), int(pb.graph_ptr[g + 1])
        wg = w[v0:v1]
        keep = np.where(wg > 0)[0] if remove_zero_weight else np.arange(v1 - v0)
        if keep.shape[0] == 0:
            continue
        a = pb.graph_adj(g)
        if keep.shape[0] != v1 - v0:
            a = a[keep][:, keep].tocsr()
        feats = G.features_gen1(wg[keep], layers[0].c_in)
        sup = G.laplacian_supports(a, len(layers[0].weights) - 1)
        out[v0 + keep] = G.gcn_forward_fp64(feats, sup, layers, kind)[:, 0]
    return out


def layers_from_meta_fixture(z, short: str):
    """LayerWeights of a checkpoint whose variables travel inside meta_activations.npz (the cheb2 checkpoints, which
    are not among the tests/golden/ckpt data fixtures)."""
    import re
    by_layer = {}
    prefix = "%s_var|" % short
    for key in z.files:
        if not key.startswith(prefix):
            continue
        m = re.match(r"^[^/]+/graphconvolution_(\d+)_vars/(weights_(\d+)|bias)$", key[len(prefix):])
        if m:
            by_layer.setdefault(int(m.group(1)), {})[m.group(2)] = z[key]
    layers = []
    for lid in sorted(by_layer):
        v = by_layer[lid]
        ks = sorted(int(k.split("_")[1]) for k in v if k.startswith("weights_"))
        lw = ckpt.LayerWeights(weights=[np.ascontiguousarray(v["weights_%d" % k], dtype=np.float32) for k in ks])
        if "bias" in v:
            lw.bias = np.ascontiguousarray(v["bias"], dtype=np.float32).reshape(-1)
        layers.append(lw)
    return layers
