"""Host-side mirror of the reference's GCN model classes (gcn/models.py) for inference.

``GCN_DQN`` (gcn/models.py:441-577), ``GCN_DEEP_DIVER`` (:301-438) and ``GCN2_DQN`` (:580-716) keep their
constructor signatures, layer wiring, variable names (``<scope>/graphconvolution_{l}_vars/weights_{i}``)
and output attributes; evaluation happens on the GPU.  Training members (loss, optimizer, opt_op) are
out of scope and absent.
"""
from __future__ import annotations

import numpy as np

from . import ckpt, engine
from . import layers as L
from .layers import _LAYER_UIDS, GraphConvolution
from .runtime import default_context
from .runtime_config import FLAGS


class Model(object):
    def __init__(self, **kwargs):
        allowed_kwargs = {"name", "logging", "concat", "flags"}
        for kwarg in kwargs.keys():
            assert kwarg in allowed_kwargs, "Invalid keyword argument: " + kwarg
        name = kwargs.get("name")
        if not name:
            name = self.__class__.__name__.lower()
        self.name = name
        self.flags = kwargs.get("flags") or FLAGS
        self.skip = self.flags.skip
        self.logging = kwargs.get("logging", False)
        self.vars = {}
        self.placeholders = {}
        self.layers = []
        self.activations = []
        self.inputs = None
        self.outputs = None
        self.outputs_softmax = None
        self.pred = None
        self.output_dim = None
        self.input_dim = None
        self._engine_model = None
        self.head = engine.HEAD_LINEAR

    def _build(self):
        raise NotImplementedError

    def build(self):
        """Wrapper for _build(): create the layers, collect the variables (gcn/models.py:63-82)."""
        if self.skip:
            raise NotImplementedError("FLAGS.skip (dense skip head, gcn/models.py:379-397) is not on the accelerated path")
        self._build()
        self._collect_vars()

    def _collect_vars(self):
        self.vars = {}
        for layer in self.layers:
            for vname, value in layer.vars.items():
                self.vars["%s/%s_vars/%s:0" % (self.name, layer.name, vname)] = value

    # ---- checkpoint -------------------------------------------------------------------------------
    def load(self, model_dir_or_prefix):
        """Restore the GraphConvolution variables from a reference checkpoint (TF bundle), as
        ``Saver.restore`` does at mwis_dqn_call.py:188-192.  Stored shapes win over the flags."""
        loaded = ckpt.load_gcn_weights(model_dir_or_prefix)
        if len(loaded) != len(self.layers):
            raise ValueError("checkpoint has %d layers, model has %d" % (len(loaded), len(self.layers)))
        for layer, lw in zip(self.layers, loaded):
            if len(lw.weights) != layer.order:
                raise ValueError("checkpoint layer has %d supports, model layer has %d" % (len(lw.weights), layer.order))
            for i, w in enumerate(lw.weights):
                layer.vars["weights_%d" % i] = w
            layer.input_dim, layer.output_dim = lw.c_in, lw.c_out
            if lw.bias is not None:
                layer.vars["bias"] = lw.bias
                layer.bias = True
        self.input_dim = self.layers[0].input_dim
        self._collect_vars()
        self._engine_model = None

    # ---- evaluation -------------------------------------------------------------------------------
    def compile(self, ctx=None) -> engine.Model:
        """The whole stack as one device-resident engine.Model (built lazily, rebuilt after load)."""
        ctx = ctx or default_context()
        if self._engine_model is None or self._engine_model.ctx is not ctx:
            acts = [layer.act_code for layer in self.layers]
            self._engine_model = engine.Model(ctx, self.layers_as_weights(), acts, head=self.head)
        return self._engine_model

    def layers_as_weights(self):
        return [ckpt.LayerWeights(weights=layer.weights, bias=layer.bias_value) for layer in self.layers]

    def run(self, batch):
        """Evaluate on a DeviceBatch; fills ``outputs``/``outputs_softmax``/``pred`` like sess.run does for
        the reference's tensors of the same names."""
        out = engine.gcn_forward(batch.ctx, self.compile(batch.ctx), batch)
        self.outputs_softmax = out
        self.outputs = out
        self.pred = np.argmax(out, axis=0)  # tf.argmax(outputs) reduces axis 0 (gcn/models.py:526)
        return out

    def predict(self):
        return self.outputs_softmax


class GCN_DQN(Model):
    def __init__(self, placeholders, input_dim, **kwargs):
        super(GCN_DQN, self).__init__(**kwargs)
        self.inputs = placeholders.get("features")
        self.input_dim = input_dim
        self.output_dim = 1  # placeholders['labels'] has one column (mwis_dqn_call.py:328)
        self.placeholders = placeholders
        self.build()

    def _build(self):
        fl = self.flags
        _LAYER_UIDS["graphconvolution"] = 0
        common = dict(placeholders=self.placeholders, dropout=True, logging=self.logging, flags=fl)
        if fl.num_layer == 1:
            last_act = getattr(fl, "last_act", "identity")
            if last_act not in ("identity", "leaky_relu"):
                raise ValueError("last_act must be 'identity' (source at HEAD) or 'leaky_relu' (as trained), got %r" % (last_act,))
            self.layers.append(GraphConvolution(input_dim=self.input_dim, output_dim=fl.diver_num,
                                                act=L.identity if last_act == "identity" else L.leaky_relu,
                                                sparse_inputs=True, **common))
        else:
            self.layers.append(GraphConvolution(input_dim=self.input_dim, output_dim=fl.hidden1, act=L.leaky_relu,
                                                sparse_inputs=True, **common))
            for _ in range(fl.num_layer - 2):
                self.layers.append(GraphConvolution(input_dim=fl.hidden1, output_dim=fl.hidden1, act=L.leaky_relu,
                                                    **common))
            self.layers.append(GraphConvolution(input_dim=fl.hidden1, output_dim=fl.diver_num, act=L.identity,
                                                **common))


class GCN_DEEP_DIVER(Model):
    """Same stack with a 2*diver_num wide last layer and a softmax over each (neg, pos) pair."""

    def __init__(self, placeholders, input_dim, **kwargs):
        super(GCN_DEEP_DIVER, self).__init__(**kwargs)
        self.inputs = placeholders.get("features")
        self.input_dim = input_dim
        self.output_dim = 2
        self.placeholders = placeholders
        self.head = engine.HEAD_PAIR_SOFTMAX
        self.build()

    def _build(self):
        fl = self.flags
        _LAYER_UIDS["graphconvolution"] = 0
        common = dict(placeholders=self.placeholders, dropout=True, logging=self.logging, flags=fl)
        self.layers.append(GraphConvolution(input_dim=self.input_dim, output_dim=fl.hidden1, act=L.leaky_relu,
                                            sparse_inputs=True, **common))
        for _ in range(fl.num_layer - 2):
            self.layers.append(GraphConvolution(input_dim=fl.hidden1, output_dim=fl.hidden1, act=L.leaky_relu, **common))
        self.layers.append(GraphConvolution(input_dim=fl.hidden1, output_dim=2 * fl.diver_num, act=L.identity, **common))

    def run(self, batch):
        out = engine.gcn_forward(batch.ctx, self.compile(batch.ctx), batch)
        self.outputs_softmax = out  # pair-softmax applied on the device (gcn/models.py:399-401)
        self.outputs = None         # the pre-softmax logits are not copied back
        self.pred = np.argmax(out, axis=0)
        return out


class GCN2_DQN(Model):
    """Explicit hyper-parameters, optional bias, the activation on every layer (gcn/models.py:580-716)."""

    def __init__(self, placeholders, hidden_dim, act=L.leaky_relu, num_layer=1, bias=False, learning_rate=0.00001,
                 learning_decay=1.0, weight_decay=5e-4, is_dual=False, is_noisy=False, **kwargs):
        super(GCN2_DQN, self).__init__(**kwargs)
        self.inputs = placeholders.get("features")
        self.input_dim = placeholders.get("feature_size") or self.flags.feature_size
        self.hidden_dim = hidden_dim
        self.output_dim = 1
        self.num_layer = num_layer
        self.placeholders = placeholders
        self.act = act
        self.bias = bias
        self.build()

    def _build(self):
        _LAYER_UIDS["graphconvolution"] = 0
        common = dict(placeholders=self.placeholders, dropout=True, logging=self.logging, act=self.act, bias=self.bias,
                      flags=self.flags)
        if self.num_layer == 1:
            self.layers.append(GraphConvolution(input_dim=self.input_dim, output_dim=self.output_dim, sparse_inputs=True,
                                                **common))
        else:
            self.layers.append(GraphConvolution(input_dim=self.input_dim, output_dim=self.hidden_dim, sparse_inputs=True,
                                                **common))
            for _ in range(self.num_layer - 2):
                self.layers.append(GraphConvolution(input_dim=self.hidden_dim, output_dim=self.hidden_dim, **common))
            self.layers.append(GraphConvolution(input_dim=self.hidden_dim, output_dim=self.output_dim, **common))
