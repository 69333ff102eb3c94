"""CPU oracle for the GCN scoring half of the hot path.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / reference arm may import
this module.  Nothing under ``distgcn_b200/`` does: the product path is the CUDA library and fails
loudly without it.

It is a numpy/scipy *restatement* (not a copy) of what the reference feeds to and computes inside
TensorFlow.  Each function cites the reference lines it follows (paths relative to the reference
repository root).

Parity status:
* support / feature construction (``normalized_adjacency``, ``laplacian_supports``,
  ``row_normalised_features``): PINNED - checked against the reference's own ``gcn/utils.py``
  (importable without TensorFlow) in ``tests/test_oracle_golden.py::test_supports_and_features_match_reference`` through the golden
  vectors made by ``tests/golden/make_golden.py``.
* the layer / model forward (``graph_convolution``, ``gcn_forward``): PINNED TO THE REFERENCE'S STORED GRAPHS.
  The reference runs the forward inside TensorFlow, which is not installable here, and ships no stored
  activations - but every checkpoint directory holds ``model.ckpt.meta``, the as-trained TensorFlow graph
  (``SparseTensorDenseMatMul -> MatMul -> AddN -> LeakyRelu -> ArgMax``).  ``oracle/tf_meta.py`` evaluates
  that stored graph op by op on inputs built and fed by the reference's own ``gcn/utils.py``;
  ``tests/test_meta_pin.py`` checks this restatement against those activations (<= 1e-6 of the score scale
  on 10 checkpoints x 10 graphs x 2 weight variants; bit-identical on most) and the wiring signature
  against gcn/layers.py:198-216.  What stays unpinned is TensorFlow's own kernel arithmetic (summation
  order / FMA contraction inside its CPU ops), i.e. the last ~1e-6.
  One disagreement between the stored graphs and the source at HEAD is kept visible: 1-layer checkpoints
  were trained with ``LeakyRelu`` on their only layer, gcn/models.py:539-548 says identity (``acts=``).
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import numpy as np
import scipy.sparse as sp

LEAKY_ALPHA = np.float32(0.2)  # tf.nn.leaky_relu default alpha, used at gcn/models.py:553,562

ACT_IDENTITY = 0
ACT_LEAKY_RELU = 1
ACT_RELU = 2


# --------------------------------------------------------------------------------------
# inputs: supports and features (fp64 on the host, as in the reference)
# --------------------------------------------------------------------------------------
def normalized_adjacency(adj) -> sp.coo_matrix:
    """D^-1/2 A D^-1/2 in fp64, isolated vertices -> 0.   Follows gcn/utils.py:120-127: row sums,
    ``power(rowsum, -0.5)``, ``inf -> 0``, then ``(A . D) ^T . D``."""
    a = sp.coo_matrix(adj, dtype=np.float64)
    deg = np.asarray(a.sum(axis=1)).reshape(-1)
    with np.errstate(divide="ignore"):
        dis = np.power(deg, -0.5)
    dis[np.isinf(dis)] = 0.0
    dmat = sp.diags(dis)
    return ((a @ dmat).transpose() @ dmat).tocoo()


def laplacian_supports(adj, max_degree: int) -> List[sp.coo_matrix]:
    """[I, L, L^2, ...] with L = I - D^-1/2 A D^-1/2.   Follows gcn/utils.py:258-274
    (``simple_polynomials``): powers are plain sparse products of L, not the Chebyshev recurrence."""
    n = adj.shape[0]
    lap = (sp.eye(n, dtype=np.float64) - normalized_adjacency(adj)).tocsr()
    out = [sp.eye(n, dtype=np.float64).tocoo(), lap.tocoo()]
    for _ in range(2, max_degree + 1):
        out.append((out[-1].tocsr() @ lap).tocoo())
    return out


def row_normalised_features(features) -> sp.csr_matrix:
    """Row-normalise a feature matrix (rowsum^-1, inf -> 0).   Follows gcn/utils.py:98-106."""
    f = sp.csr_matrix(features, dtype=np.float64)
    rowsum = np.asarray(f.sum(axis=1)).reshape(-1)
    with np.errstate(divide="ignore"):
        rinv = np.power(rowsum, -1.0)
    rinv[np.isinf(rinv)] = 0.0
    return (sp.diags(rinv) @ f).tocsr()


def features_gen1(wts_nn, feature_size: int) -> sp.csr_matrix:
    """Feature matrix of the generation-1 agent.   Follows mwis_dqn_call.py:129-135: every column is
    ``w / ||w||_2``, then row normalisation - so each stored entry is 1/feature_size (up to one fp64
    rounding) and rows of zero-weight vertices are empty."""
    w = np.asarray(wts_nn, dtype=np.float64).reshape(-1, 1)
    norm = np.linalg.norm(w)
    with np.errstate(divide="ignore", invalid="ignore"):
        cols = np.ones((w.shape[0], feature_size)) * (w / norm)
    return row_normalised_features(sp.csr_matrix(cols))


def features_gen2(wts_nn, feature_size: int, predict: str) -> sp.csr_matrix:
    """Feature matrix of the generation-2 solver.   Follows mwis_gdpg_call.py:82-96: ``predict ==
    'mwis'`` uses all-ones features, row-normalised; otherwise ``w / (max w + 1e-9)`` un-normalised."""
    w = np.asarray(wts_nn, dtype=np.float64).reshape(-1, 1)
    if predict == "mwis":
        return row_normalised_features(sp.csr_matrix(np.ones((w.shape[0], feature_size))))
    return sp.csr_matrix(np.ones((w.shape[0], feature_size)) * (w / (np.amax(w) + 1e-9)))


def to_fp32_csr(mat) -> sp.csr_matrix:
    """The fp64 -> fp32 cast that happens when tuples are fed to float32 sparse placeholders
    (mwis_dqn_call.py:326-328 with gcn/utils.py:157-168).  Entry order inside a row is ascending
    column, the order scipy's COO conversion hands to TensorFlow."""
    m = sp.csr_matrix(mat).astype(np.float32)
    m.sort_indices()
    return m


# --------------------------------------------------------------------------------------
# the layer and the models (fp32 after the feed)
# --------------------------------------------------------------------------------------
def _activate(x: np.ndarray, act: int) -> np.ndarray:
    if act == ACT_IDENTITY:
        return x
    if act == ACT_LEAKY_RELU:
        return np.where(x >= 0, x, LEAKY_ALPHA * x).astype(np.float32)
    if act == ACT_RELU:
        return np.maximum(x, np.float32(0)).astype(np.float32)
    raise ValueError("unknown activation %r" % (act,))


def graph_convolution(x, supports32: Sequence[sp.csr_matrix], weights: Sequence[np.ndarray],
                      bias: Optional[np.ndarray], act: int) -> np.ndarray:
    """act( sum_i T_i . (X . W_i) + b ) in fp32.   Follows gcn/layers.py:198-216: project first
    (``dot(x, W_i)``, sparse or dense), then aggregate with support i, ``add_n`` the supports in
    order, add the bias if present, apply the activation.  Dropout is the identity at inference
    (rate placeholder defaults to 0, mwis_dqn_call.py:331)."""
    total = None
    for t, w in zip(supports32, weights):
        w32 = np.asarray(w, dtype=np.float32)
        if sp.issparse(x):
            pre = np.asarray(sp.csr_matrix(x, dtype=np.float32) @ w32, dtype=np.float32)
        else:
            pre = np.asarray(x, dtype=np.float32) @ w32
        part = np.asarray(t @ pre, dtype=np.float32)
        total = part if total is None else (total + part).astype(np.float32)
    if bias is not None:
        total = (total + np.asarray(bias, dtype=np.float32)).astype(np.float32)
    return _activate(total, act)


def layer_activations(num_layers: int, kind: str) -> List[int]:
    """Activation of each layer for the three model classes.
    * 'gcn_dqn'   (gcn/models.py:536-573): leaky-ReLU on all but the last layer, identity on the last
      (and on the single layer when num_layer == 1).
    * 'gcn_deep_diver' (gcn/models.py:411-434): same pattern, last layer is 2*diver_num wide.
    * 'gcn2_dqn'  (gcn/models.py:670-708): the configured activation (leaky-ReLU by default) on every
      layer, the last included."""
    if kind in ("gcn_dqn", "gcn_deep_diver"):
        return [ACT_LEAKY_RELU] * (num_layers - 1) + [ACT_IDENTITY]
    if kind == "gcn2_dqn":
        return [ACT_LEAKY_RELU] * num_layers
    raise ValueError(kind)


def gcn_forward(features, supports, layers, kind: str = "gcn_dqn", acts: Optional[Sequence[int]] = None) -> np.ndarray:
    """Run the GraphConvolution stack.  ``layers`` is a list of objects with ``.weights`` (list of
    [c_in, c_out] arrays, one per support) and ``.bias``.   Follows the sequential wiring at
    gcn/models.py:487-500 (each layer consumes the previous activation; the first layer takes the
    sparse features, gcn/models.py:544,555).  Returns ``outputs`` ([N, c_out_last], fp32)."""
    sup32 = [to_fp32_csr(t) for t in supports]
    if acts is None:
        acts = layer_activations(len(layers), kind)
    h = to_fp32_csr(features)
    for lw, act in zip(layers, acts):
        if len(lw.weights) != len(sup32):
            raise ValueError("layer has %d weight matrices but %d supports were given" % (len(lw.weights), len(sup32)))
        h = graph_convolution(h, sup32, lw.weights, lw.bias, act)
    return h


def gcn_forward_fp64(features, supports, layers, kind: str = "gcn_dqn", acts: Optional[Sequence[int]] = None) -> np.ndarray:
    """The same network evaluated in float64 on the fp32-rounded inputs TensorFlow is fed (supports,
    features and weights rounded to fp32, arithmetic exact to ~1e-16).  Not a restatement of the
    reference (which computes in fp32) but the yardstick for fp32 implementations: numpy's fp32
    forward above is itself ~1e-5 (relative to the score scale) away from it on the 20-layer
    checkpoints, which is why two correct fp32 implementations can differ by up to ~2e-5."""
    sup64 = [to_fp32_csr(t).astype(np.float64) for t in supports]
    if acts is None:
        acts = layer_activations(len(layers), kind)
    h = to_fp32_csr(features).astype(np.float64)
    alpha = np.float64(LEAKY_ALPHA)
    for lw, act in zip(layers, acts):
        total = None
        for t, w in zip(sup64, lw.weights):
            pre = h @ np.asarray(w, dtype=np.float32).astype(np.float64)
            pre = np.asarray(pre.todense()) if sp.issparse(pre) else np.asarray(pre)
            part = np.asarray(t @ pre)
            total = part if total is None else total + part
        if lw.bias is not None:
            total = total + np.asarray(lw.bias, dtype=np.float32).astype(np.float64)
        if act == ACT_LEAKY_RELU:
            total = np.where(total >= 0, total, alpha * total)
        elif act == ACT_RELU:
            total = np.maximum(total, 0.0)
        h = total
    return h


def pair_softmax(outputs: np.ndarray, diver_num: int) -> np.ndarray:
    """GCN_DEEP_DIVER head: softmax over each consecutive (neg, pos) column pair.   Follows
    gcn/models.py:399-401 with output_dim == 2."""
    out = np.empty_like(outputs, dtype=np.float32)
    for d in range(diver_num):
        z = outputs[:, 2 * d:2 * d + 2].astype(np.float32)
        z = z - z.max(axis=1, keepdims=True)
        e = np.exp(z).astype(np.float32)
        out[:, 2 * d:2 * d + 2] = e / e.sum(axis=1, keepdims=True)
    return out


def utility(act_vals: np.ndarray, wts, predict: str) -> np.ndarray:
    """fp32 score x fp64 link weight -> fp64 utility ('mwis'), or the score itself ('mis').
    Follows mwis_dqn_call.py:230-235."""
    a = np.asarray(act_vals).reshape(-1)
    if predict == "mwis":
        return np.multiply(a, np.asarray(wts, dtype=np.float64).reshape(-1))
    return a.astype(np.float64)
