"""Host-side mirror of the reference's ``gcn/layers.py`` for the GraphConvolution layer.

Same constructor, same variable names (``weights_0 .. weights_{K-1}``, ``bias``), same error
behaviour (assert on unknown kwargs, NameError on a bad ``wts_init``); ``__call__`` runs the layer on
the GPU through the C-ABI (dg_graph_convolution) instead of building TensorFlow ops.

What replaces the TensorFlow placeholders: ``placeholders`` is an ordinary dict.  ``placeholders
['support']`` only needs the right LENGTH (1 + max_degree, reference gcn/layers.py:165) - the support
matrices themselves are never materialised, the layer reads the graph from ``placeholders['batch']``,
an ``engine.DeviceBatch`` that the caller stores there before calling (the analogue of the feed_dict).
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp

from . import engine
from .runtime import default_context
from .runtime_config import FLAGS

# global unique layer ID dictionary for layer name assignment (gcn/layers.py:4-15)
_LAYER_UIDS = {}


def get_layer_uid(layer_name=""):
    if layer_name not in _LAYER_UIDS:
        _LAYER_UIDS[layer_name] = 1
        return 1
    _LAYER_UIDS[layer_name] += 1
    return _LAYER_UIDS[layer_name]


# ---- activations: the callables a caller may pass as ``act`` ------------------------------------
def leaky_relu(x, alpha=0.2):
    return np.where(x >= 0, x, alpha * x)


def relu(x):
    return np.maximum(x, 0)


def identity(x):
    return x


def act_code(act) -> int:
    """Map an activation callable to the library's code.  Known callables map directly; any other
    callable (e.g. the reference's ``lambda x: x`` or ``tf.nn.leaky_relu``) is probed on two points."""
    if act is None or act is identity:
        return engine.ACT_IDENTITY
    if act is leaky_relu:
        return engine.ACT_LEAKY_RELU
    if act is relu:
        return engine.ACT_RELU
    if isinstance(act, str):
        table = {"identity": engine.ACT_IDENTITY, "linear": engine.ACT_IDENTITY, "leaky_relu": engine.ACT_LEAKY_RELU,
                 "relu": engine.ACT_RELU}
        if act not in table:
            raise ValueError("unsupported activation %r" % act)
        return table[act]
    probe = np.asarray(act(np.array([-1.0, 2.0], dtype=np.float32)), dtype=np.float64).reshape(-1)
    if np.allclose(probe, [-1.0, 2.0]):
        return engine.ACT_IDENTITY
    if np.allclose(probe, [-0.2, 2.0]):
        return engine.ACT_LEAKY_RELU
    if np.allclose(probe, [0.0, 2.0]):
        return engine.ACT_RELU
    raise ValueError("unsupported activation: act(-1), act(2) = %s" % probe)


def glorot(shape, rng=None):
    """Glorot & Bengio uniform init (gcn/inits.py:17-21)."""
    rng = rng or np.random.default_rng()
    init_range = np.sqrt(6.0 / (shape[0] + shape[1]))
    return rng.uniform(-init_range, init_range, size=shape).astype(np.float32)


def zeros(shape):
    return np.zeros(shape, dtype=np.float32)


def make_placeholders(num_supports, feature_size=None):
    """The placeholder dict of mwis_dqn_call.py:325-333 without TensorFlow."""
    return {"support": [None] * num_supports, "features": None, "labels": None, "labels_mask": None, "dropout": 0.0,
            "num_features_nonzero": None, "batch": None, "feature_size": feature_size}


class Layer(object):
    """Base layer class (gcn/layers.py:56-99): name, vars, logging flag."""

    def __init__(self, **kwargs):
        allowed_kwargs = {"name", "logging"}
        for kwarg in kwargs.keys():
            assert kwarg in allowed_kwargs, "Invalid keyword argument: " + kwarg
        name = kwargs.get("name")
        if not name:
            layer = self.__class__.__name__.lower()
            name = layer + "_" + str(get_layer_uid(layer))
        self.name = name
        self.vars = {}
        self.logging = kwargs.get("logging", False)
        self.sparse_inputs = False

    def _call(self, inputs):
        return inputs

    def __call__(self, inputs):
        return self._call(inputs)


class GraphConvolution(Layer):
    """Graph convolution layer: act(sum_i T_i . (X . W_i) + b)  (gcn/layers.py:149-216)."""

    def __init__(self, input_dim, output_dim, placeholders, dropout=0., channel=0, num_channels=1,
                 sparse_inputs=False, act=relu, bias=False, featureless=False, flags=None, rng=None, **kwargs):
        super(GraphConvolution, self).__init__(**kwargs)
        fl = flags or FLAGS
        self.dropout = placeholders.get("dropout", 0.0) if dropout else 0.
        self.act = act
        self.act_code = act_code(act)
        self.sparse_inputs = sparse_inputs
        self.featureless = featureless
        self.bias = bias
        self.channel = int(channel)
        self.num_channels = int(num_channels)
        self.order = int(len(placeholders["support"]) / self.num_channels)
        self.placeholders = placeholders
        self.input_dim, self.output_dim = int(input_dim), int(output_dim)
        for i in range(self.order):
            if fl.wts_init == "random":
                self.vars["weights_" + str(i)] = glorot([input_dim, output_dim], rng)
            elif fl.wts_init == "zeros":
                self.vars["weights_" + str(i)] = zeros([input_dim, output_dim])
            else:
                raise NameError("Unsupported wts_init: {}".format(fl.wts_init))
        if self.bias:
            self.vars["bias"] = zeros([output_dim])

    # the pieces engine.Model wants
    @property
    def weights(self):
        return [self.vars["weights_" + str(i)] for i in range(self.order)]

    @property
    def bias_value(self):
        return self.vars.get("bias") if self.bias else None

    def _call(self, inputs):
        if self.featureless:
            raise NotImplementedError("featureless GraphConvolution is not on the accelerated path")
        if self.order != 2:
            raise NotImplementedError("only the cheb1 supports [I, L] are accelerated (got %d supports)" % self.order)
        batch = self.placeholders.get("batch")
        if batch is None:
            raise RuntimeError("placeholders['batch'] must hold the engine.DeviceBatch to convolve over")
        x = inputs
        if sp.issparse(x):
            x = np.asarray(x.todense())
        x = np.ascontiguousarray(x, dtype=np.float32)
        # dropout is the identity at inference (rate placeholder defaults to 0, mwis_dqn_call.py:331)
        return engine.graph_convolution(batch.ctx, batch, x, self.vars["weights_0"], self.vars["weights_1"],
                                        self.bias_value, act=self.act_code)


def device_batch_from_adj(adj, ctx=None):
    """One scipy adjacency matrix -> resident single-graph batch (helper for per-graph call sites)."""
    from .batch import pack_graphs
    return engine.DeviceBatch(ctx or default_context(), pack_graphs([adj]))
