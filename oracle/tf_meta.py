"""Interpreter for the reference's AS-TRAINED TensorFlow graphs.  TEST INFRASTRUCTURE ONLY.

Every shipped checkpoint directory ``model/result_*/`` holds ``model.ckpt.meta``: a serialized
``MetaGraphDef`` (producer TensorFlow 1.15.0) of the graph ``mwis_dqn_call.py`` built when the
weights were trained - the very ops ``DQNAgent.predict`` runs through ``sess.run([model.outputs_softmax,
model.pred])`` (mwis_dqn_call.py:140-143; graph built by gcn/models.py:487-526 + gcn/layers.py:189-216).
That file is a reference-held, executable description of the forward: which weight goes to which
``MatMul`` / ``SparseTensorDenseMatMul``, in which order supports are aggregated (``AddN``), where
``LeakyRelu`` sits, what ``ArgMax`` reduces over.  TensorFlow is not installable here, so this module

* decodes the protobuf wire format of ``MetaGraphDef -> GraphDef -> NodeDef`` (field numbers of
  tensorflow/core/framework/{graph,node_def,attr_value,tensor,tensor_shape}.proto, public and stable),
* evaluates the sub-graph that ``ArgMax`` (= ``model.pred``) depends on with numpy, one op at a time,
  in float32 where TensorFlow computes in float32, with the variables read from the checkpoint bundle,
* takes its inputs through the reference's own ``gcn/utils.construct_feed_dict4pred`` (imported
  unmodified by tests/golden/make_golden.py): placeholder stand-ins are hashable objects, a sparse
  placeholder expands to its (indices, values, dense_shape) nodes exactly as ``Session.run`` expands a
  ``SparseTensorValue`` feed.

Its outputs are committed as golden activations (tests/golden/meta_activations.npz) and pin
``oracle/gcn_oracle.py``'s restatement AND the CUDA path to the reference's stored graph.  Which
placeholder plays which role is inferred from the graph structure (``roles()``), not from names.

Numerics: op semantics follow the TensorFlow CPU kernels' documented behaviour; summation order inside
``SparseTensorDenseMatMul`` is the order of the fed indices (sequential ``out[row] += a * b[col]``), inside
``MatMul`` it is numpy's (Eigen's differs in the last bits) - compare at ~1e-6 relative, not bitwise.
Nothing under ``distgcn_b200/`` imports this module.
"""
from __future__ import annotations

import struct
from typing import Dict, List, Optional, Tuple

import numpy as np

# tensorflow/core/framework/types.proto
_DT = {1: np.float32, 2: np.float64, 3: np.int32, 9: np.int64, 10: np.bool_}


class MetaGraphError(RuntimeError):
    pass


# ---- protobuf wire format ---------------------------------------------------------------------------------
def _varint(buf: bytes, pos: int) -> Tuple[int, int]:
    out = 0
    shift = 0
    while True:
        b = buf[pos]
        pos += 1
        out |= (b & 0x7F) << shift
        if not b & 0x80:
            return out, pos
        shift += 7


def _fields(buf: bytes) -> List[Tuple[int, int, object]]:
    out = []
    pos = 0
    n = len(buf)
    while pos < n:
        tag, pos = _varint(buf, pos)
        num, wt = tag >> 3, tag & 7
        if wt == 0:
            val, pos = _varint(buf, pos)
        elif wt == 1:
            val = buf[pos:pos + 8]
            pos += 8
        elif wt == 2:
            ln, pos = _varint(buf, pos)
            val = buf[pos:pos + ln]
            pos += ln
        elif wt == 5:
            val = buf[pos:pos + 4]
            pos += 4
        else:
            raise MetaGraphError("unsupported wire type %d" % wt)
        out.append((num, wt, val))
    return out


def _signed(v: int) -> int:
    return v - (1 << 64) if v >= (1 << 63) else v


def _shape(buf: bytes) -> List[int]:
    """TensorShapeProto: repeated dim = 2 {size = 1}."""
    dims = []
    for num, _, val in _fields(buf):
        if num == 2:
            size = 0
            for n2, _, v2 in _fields(val):
                if n2 == 1:
                    size = _signed(v2)
            dims.append(size)
    return dims


def _tensor(buf: bytes) -> np.ndarray:
    """TensorProto: dtype = 1, tensor_shape = 2, tensor_content = 4, float_val = 5, double_val = 6,
    int_val = 7, int64_val = 10, bool_val = 11 (scalars may be given as one repeated value = splat)."""
    dtype = None
    shape: List[int] = []
    content = None
    vals: List = []
    for num, wt, val in _fields(buf):
        if num == 1:
            dtype = _DT[val]
        elif num == 2:
            shape = _shape(val)
        elif num == 4:
            content = val
        elif num == 5:
            if wt == 5:
                vals.append(struct.unpack("<f", val)[0])
            else:
                vals.extend(struct.unpack("<%df" % (len(val) // 4), val))
        elif num == 6:
            if wt == 1:
                vals.append(struct.unpack("<d", val)[0])
            else:
                vals.extend(struct.unpack("<%dd" % (len(val) // 8), val))
        elif num in (7, 10, 11):
            if wt == 0:
                vals.append(_signed(val))
            else:
                p = 0
                while p < len(val):
                    v, p = _varint(val, p)
                    vals.append(_signed(v))
    if dtype is None:
        raise MetaGraphError("tensor without dtype")
    count = int(np.prod(shape)) if shape else 1
    if content is not None:
        arr = np.frombuffer(content, dtype=np.dtype(dtype).newbyteorder("<")).astype(dtype)
    elif len(vals) == count:
        arr = np.asarray(vals, dtype=dtype)
    elif len(vals) == 1:
        arr = np.full(count, vals[0], dtype=dtype)
    elif not vals:
        arr = np.zeros(count, dtype=dtype)
    else:
        raise MetaGraphError("tensor has %d values for %d elements" % (len(vals), count))
    return arr.reshape(shape)


def _attr(buf: bytes):
    """AttrValue oneof: s = 2, i = 3, f = 4, b = 5, type = 6, shape = 7, tensor = 8, list = 1."""
    for num, wt, val in _fields(buf):
        if num == 2:
            return val
        if num == 3:
            return _signed(val)
        if num == 4:
            return np.float32(struct.unpack("<f", val)[0])
        if num == 5:
            return bool(val)
        if num == 6:
            return _DT.get(val, val)
        if num == 7:
            return _shape(val)
        if num == 8:
            return _tensor(val)
        if num == 1:
            return ("list", val)
    return None


class Node:
    __slots__ = ("name", "op", "inputs", "raw_attr", "_attr")

    def __init__(self):
        self.name = ""
        self.op = ""
        self.inputs: List[str] = []
        self.raw_attr: Dict[str, bytes] = {}
        self._attr: Dict[str, object] = {}

    def attr(self, key, default=None):
        if key not in self._attr:
            self._attr[key] = _attr(self.raw_attr[key]) if key in self.raw_attr else default
        return self._attr[key]


def parse_meta(path: str) -> Dict[str, Node]:
    """MetaGraphDef.graph_def (field 2) -> {node name: Node}.  NodeDef: name = 1, op = 2, input = 3, attr = 5
    (map entries: key = 1, value = 2)."""
    with open(path, "rb") as fh:
        buf = fh.read()
    graph_def = None
    for num, _, val in _fields(buf):
        if num == 2:
            graph_def = val
    if graph_def is None:
        raise MetaGraphError("%s: no graph_def" % path)
    nodes: Dict[str, Node] = {}
    for num, _, val in _fields(graph_def):
        if num != 1:
            continue
        nd = Node()
        for n2, _, v2 in _fields(val):
            if n2 == 1:
                nd.name = v2.decode()
            elif n2 == 2:
                nd.op = v2.decode()
            elif n2 == 3:
                nd.inputs.append(v2.decode())
            elif n2 == 5:
                key = None
                value = b""
                for n3, _, v3 in _fields(v2):
                    if n3 == 1:
                        key = v3.decode()
                    elif n3 == 2:
                        value = v3
                nd.raw_attr[key] = value
        nodes[nd.name] = nd
    return nodes


def _base(ref: str) -> str:
    ref = ref.lstrip("^")
    return ref.split(":")[0]


# ---- placeholders: stand-ins the reference's construct_feed_dict4pred can use as dict keys -------------------
class SparsePlaceholder:
    """What ``tf.compat.v1.sparse_placeholder`` returns, reduced to the three node names."""

    def __init__(self, indices: str, values: str, dense_shape: str):
        self.indices, self.values, self.dense_shape = indices, values, dense_shape

    def __repr__(self):
        return "SparsePlaceholder(%s, %s, %s)" % (self.indices, self.values, self.dense_shape)


class DensePlaceholder:
    def __init__(self, name: str):
        self.name = name

    def __repr__(self):
        return "DensePlaceholder(%s)" % self.name


def expand_feed(feed: dict) -> Dict[str, np.ndarray]:
    """{placeholder stand-in: value} -> {node name: array}; a (coords, values, shape) tuple fed to a sparse
    placeholder becomes three arrays, as Session.run does for SparseTensorValue."""
    out: Dict[str, np.ndarray] = {}
    for key, val in feed.items():
        if isinstance(key, SparsePlaceholder):
            coords, values, shape = val
            out[key.indices] = np.asarray(coords, dtype=np.int64)
            out[key.values] = np.asarray(values, dtype=np.float32)  # the placeholder's dtype: the fp64 -> fp32 cast
            out[key.dense_shape] = np.asarray(shape, dtype=np.int64)
        elif isinstance(key, DensePlaceholder):
            out[key.name] = np.asarray(val)
        else:
            raise MetaGraphError("unknown feed key %r" % (key,))
    return out


class StoredGraph:
    """The forward part of one checkpoint's stored graph."""

    def __init__(self, meta_path: str, variables: Dict[str, np.ndarray]):
        self.nodes = parse_meta(meta_path)
        self.variables = variables
        argmax = [n for n in self.nodes.values() if n.op == "ArgMax" and not n.name.startswith("gradients")]
        if len(argmax) != 1:
            raise MetaGraphError("expected one ArgMax (model.pred), found %d" % len(argmax))
        self.pred = argmax[0].name                      # model.pred = tf.argmax(outputs), gcn/models.py:526
        self.outputs = _base(argmax[0].inputs[0])       # model.outputs = activations[-1], gcn/models.py:499

    # -- structure ------------------------------------------------------------------------------------------
    def forward_nodes(self) -> List[Node]:
        """Topologically ordered nodes ``pred`` depends on."""
        order: List[Node] = []
        seen = set()

        def visit(name):
            stack = [(name, False)]
            while stack:
                nm, done = stack.pop()
                if done:
                    order.append(self.nodes[nm])
                    continue
                if nm in seen:
                    continue
                seen.add(nm)
                stack.append((nm, True))
                for inp in reversed(self.nodes[nm].inputs):
                    b = _base(inp)
                    if b not in seen:
                        stack.append((b, False))
        visit(self.pred)
        return order

    def _leaf_placeholders(self, name: str) -> List[str]:
        out = []
        seen = set()
        stack = [name]
        while stack:
            nm = stack.pop()
            if nm in seen:
                continue
            seen.add(nm)
            nd = self.nodes[nm]
            if nd.op == "Placeholder":
                out.append(nm)
            stack.extend(_base(i) for i in nd.inputs)
        return out

    def roles(self) -> dict:
        """The ``placeholders`` dict of mwis_dqn_call.py:323-333, recovered from the structure:
        * support[i]: the sparse operand of the i-th aggregation ``SparseTensorDenseMatMul`` feeding the first
          layer's ``AddN`` (its three inputs are placeholders);
        * features: the sparse operand of the first layer's projection, traced through ``sparse_retain``
          (GatherV2 of the fed indices / values) back to its placeholders;
        * num_features_nonzero: the shape input of the sparse-dropout ``RandomUniform``; dropout: the one
          ``PlaceholderWithDefault``; labels_mask / labels: the remaining placeholders (not on the forward)."""
        fwd = self.forward_nodes()
        addn = next(n for n in fwd if n.op == "AddN")
        supports = []
        feat = None
        for inp in addn.inputs:
            agg = self.nodes[_base(inp)]
            if agg.op != "SparseTensorDenseMatMul":
                raise MetaGraphError("AddN input %s is %s" % (agg.name, agg.op))
            idx, val, shp, dense = (_base(x) for x in agg.inputs)
            for nm in (idx, val, shp):
                if self.nodes[nm].op != "Placeholder":
                    raise MetaGraphError("support operand %s is not a placeholder" % nm)
            supports.append(SparsePlaceholder(idx, val, shp))
            proj = self.nodes[dense]
            if proj.op != "SparseTensorDenseMatMul":
                raise MetaGraphError("first-layer projection is %s" % proj.op)
            pidx, pval, pshp, _ = (_base(x) for x in proj.inputs)
            fi = [p for p in self._leaf_placeholders(pidx) if self.nodes[p].attr("dtype") is np.int64
                  and len(self.nodes[p].attr("shape") or []) == 2]
            fv = [p for p in self._leaf_placeholders(pval) if self.nodes[p].attr("dtype") is np.float32]
            fs = self._leaf_placeholders(pshp)
            cand = SparsePlaceholder(fi[0], fv[0], fs[0])
            if feat is not None and (feat.indices, feat.values, feat.dense_shape) != (cand.indices, cand.values, cand.dense_shape):
                raise MetaGraphError("projections of one layer read different features")
            feat = cand
        rnd = [n for n in fwd if n.op == "RandomUniform" and self.nodes[_base(n.inputs[0])].op == "Placeholder"]
        nnz = DensePlaceholder(_base(rnd[0].inputs[0])) if rnd else None
        pwd = [n for n in fwd if n.op == "PlaceholderWithDefault"]
        used = {feat.indices, feat.values, feat.dense_shape}
        for s in supports:
            used |= {s.indices, s.values, s.dense_shape}
        if nnz is not None:
            used.add(nnz.name)
        others = sorted((n for n, nd in self.nodes.items() if nd.op == "Placeholder" and n not in used),
                        key=lambda s: int(s.split("_")[-1]) if "_" in s else 0)
        out = {"support": supports, "features": feat, "num_features_nonzero": nnz,
               "dropout": DensePlaceholder(pwd[0].name) if pwd else None}
        # construct_feed_dict4pred also feeds labels_mask (gcn/utils.py:166); any remaining int32 placeholder will do
        masks = [n for n in others if self.nodes[n].attr("dtype") is np.int32]
        out["labels_mask"] = DensePlaceholder(masks[0]) if masks else DensePlaceholder("__unused_labels_mask")
        labels = [n for n in others if self.nodes[n].attr("dtype") is np.float32]
        out["labels"] = DensePlaceholder(labels[0]) if labels else None
        return out

    def signature(self) -> List[str]:
        """Compute ops of the forward in execution order with their layer scope - the wiring at a glance,
        e.g. ['graphconvolution_1:SparseTensorDenseMatMul(weights_0)', ..., 'graphconvolution_1:LeakyRelu', 'ArgMax']."""
        out = []
        for nd in self.forward_nodes():
            if nd.op not in ("SparseTensorDenseMatMul", "MatMul", "AddN", "LeakyRelu", "Relu", "ArgMax", "BiasAdd",
                             "Softmax", "AddV2", "Add"):
                continue
            scope = nd.name.split("/")[0]
            if nd.op in ("AddV2", "Add"):
                # only the bias add (an operand is a variable read); dropout / initializer arithmetic is skipped
                srcs = [self.nodes[_base(i)] for i in nd.inputs]
                if not any(s.op == "Identity" and self.nodes[_base(s.inputs[0])].op == "VariableV2" for s in srcs):
                    continue
            tag = nd.op
            for inp in nd.inputs:
                src = self.nodes[_base(inp)]
                if src.op == "Identity" and self.nodes[_base(src.inputs[0])].op == "VariableV2":
                    tag += "(%s)" % _base(src.inputs[0]).split("/")[-1]
                if nd.op == "SparseTensorDenseMatMul" and src.op == "Placeholder" and inp == nd.inputs[1]:
                    tag += "(support:%s)" % src.name
            out.append("%s:%s" % (scope, tag) if "/" in nd.name else tag)
        return out

    # -- evaluation -----------------------------------------------------------------------------------------
    def run(self, fetches: List[str], feed: Dict[str, np.ndarray], rng: Optional[np.random.Generator] = None):
        """Evaluate ``fetches`` (node names) given ``feed`` ({node name: array}).  ``RandomUniform`` draws from
        ``rng`` (default: a fixed generator) - with the dropout placeholder at its default 0 every draw keeps
        every entry (floor(1 + u) = 1, u >= 0), which the dropout test varies the generator to show."""
        rng = rng or np.random.default_rng(0)
        cache: Dict[str, np.ndarray] = {}

        def val(ref: str):
            return cache[_base(ref)]

        need = []
        seen = set()
        for f in fetches:
            stack = [(f, False)]
            while stack:
                nm, done = stack.pop()
                if done:
                    need.append(nm)
                    continue
                if nm in seen:
                    continue
                seen.add(nm)
                stack.append((nm, True))
                if nm in feed:
                    continue
                for inp in reversed(self.nodes[nm].inputs):
                    if inp.startswith("^"):
                        continue
                    stack.append((_base(inp), False))
        f32 = np.float32
        for nm in need:
            nd = self.nodes[nm]
            if nm in feed:
                cache[nm] = feed[nm]
                continue
            op = nd.op
            ins = [i for i in nd.inputs if not i.startswith("^")]
            if op == "Placeholder":
                raise MetaGraphError("placeholder %s needs a feed" % nm)
            elif op == "PlaceholderWithDefault":
                r = val(ins[0])
            elif op == "Const":
                r = nd.attr("value")
            elif op == "VariableV2":
                if nm not in self.variables:
                    raise MetaGraphError("variable %s is not in the checkpoint" % nm)
                r = np.asarray(self.variables[nm], dtype=nd.attr("dtype"))
                want = nd.attr("shape")
                if list(r.shape) != list(want):
                    raise MetaGraphError("variable %s: checkpoint shape %s, graph shape %s" % (nm, r.shape, want))
            elif op == "Identity":
                r = val(ins[0])
            elif op in ("Add", "AddV2"):
                r = val(ins[0]) + val(ins[1])
            elif op == "Sub":
                r = val(ins[0]) - val(ins[1])
            elif op == "Mul":
                r = val(ins[0]) * val(ins[1])
            elif op == "RealDiv":
                r = val(ins[0]) / val(ins[1])
            elif op == "Floor":
                r = np.floor(val(ins[0]))
            elif op == "Cast":
                r = np.asarray(val(ins[0])).astype(nd.attr("DstT"))
            elif op == "GreaterEqual":
                r = val(ins[0]) >= val(ins[1])
            elif op == "Shape":
                r = np.asarray(np.shape(val(ins[0])), dtype=nd.attr("out_type") or np.int32)
            elif op == "RandomUniform":
                shape = tuple(int(x) for x in np.asarray(val(ins[0])).reshape(-1))
                r = rng.random(shape, dtype=np.float32)
            elif op == "Where":
                r = np.argwhere(val(ins[0])).astype(np.int64)
            elif op == "Reshape":
                r = np.reshape(val(ins[0]), tuple(int(x) for x in np.asarray(val(ins[1])).reshape(-1)))
            elif op == "GatherV2":
                r = np.take(val(ins[0]), np.asarray(val(ins[1]), dtype=np.int64), axis=int(val(ins[2])))
            elif op == "SparseDenseCwiseMul":
                idx, v, shp, dense = (val(i) for i in ins)
                dense = np.asarray(dense)
                if dense.ndim != 0:
                    raise MetaGraphError("SparseDenseCwiseMul: only a scalar dense operand occurs in these graphs")
                r = np.asarray(v) * dense
            elif op == "SparseTensorDenseMatMul":
                if nd.attr("adjoint_a") or nd.attr("adjoint_b"):
                    raise MetaGraphError("adjoint SparseTensorDenseMatMul does not occur in these graphs")
                idx, v, shp, dense = (np.asarray(val(i)) for i in ins)
                if dense.dtype != f32 or v.dtype != f32:
                    raise MetaGraphError("%s: operands are %s, %s (float32 expected)" % (nm, v.dtype, dense.dtype))
                if int(shp[1]) != dense.shape[0]:
                    raise MetaGraphError("%s: sparse [%d, %d] x dense %s" % (nm, shp[0], shp[1], dense.shape))
                r = np.zeros((int(shp[0]), dense.shape[1]), dtype=f32)
                # the CPU kernel walks the nnz in the order they were fed: out[row] += a * b[col], float32
                np.add.at(r, idx[:, 0], v[:, None] * dense[idx[:, 1]])
            elif op == "MatMul":
                a, b = np.asarray(val(ins[0])), np.asarray(val(ins[1]))
                if nd.attr("transpose_a"):
                    a = a.T
                if nd.attr("transpose_b"):
                    b = b.T
                if a.dtype != f32 or b.dtype != f32:
                    raise MetaGraphError("%s: operands are %s, %s" % (nm, a.dtype, b.dtype))
                r = a @ b
            elif op == "AddN":
                r = val(ins[0])
                for i in ins[1:]:
                    r = r + val(i)
            elif op == "LeakyRelu":
                x = val(ins[0])
                alpha = nd.attr("alpha")
                alpha = f32(0.2) if alpha is None else f32(alpha)
                r = np.where(x > 0, x, alpha * x).astype(x.dtype)
            elif op == "Relu":
                r = np.maximum(val(ins[0]), 0)
            elif op == "ArgMax":
                r = np.argmax(val(ins[0]), axis=int(val(ins[1]))).astype(nd.attr("output_type") or np.int64)
            else:
                raise MetaGraphError("op %s (%s) is not implemented" % (op, nm))
            cache[nm] = r
        return [cache[f] for f in fetches]

    def leaky_alphas(self) -> List[float]:
        return [float(n.attr("alpha")) for n in self.forward_nodes() if n.op == "LeakyRelu"]


def load_stored_graph(model_dir: str, read_all_variables) -> StoredGraph:
    """``read_all_variables(prefix) -> {name: array}`` is the checkpoint reader (tests pass
    distgcn_b200.ckpt.read_bundle: the product's loader is itself under test that way)."""
    import os
    prefix = os.path.join(model_dir, "model.ckpt")
    return StoredGraph(prefix + ".meta", read_all_variables(prefix))
