"""Reader for the reference's saved checkpoints (TensorFlow "V2 bundle"), without TensorFlow.

Replaces ``tf.train.get_checkpoint_state`` + ``Saver.restore`` as used by
``DQNAgent.load`` (reference ``mwis_dqn_call.py:188-192``) for the variables the
inference path needs: ``<scope>/graphconvolution_{l}_vars/weights_{i}`` and the
optional ``.../bias`` (created at reference ``gcn/layers.py:174-186``).  Adam slot
variables and ``beta{1,2}_power`` are skipped.

File format ("Checkpoint / resume"):

* ``<prefix>.index`` is a LevelDB-style sorted table.  The last 48 bytes are the
  footer: two block handles (varint offset, varint size) for the metaindex and the
  index block, zero padding, and the magic ``57 fb 80 8b 24 75 47 db``.  A block is a
  run of prefix-compressed entries ``(varint shared, varint non_shared, varint
  value_len, key_suffix, value)`` followed by a restart array and ``uint32
  num_restarts``; on disk every block is followed by a 5-byte trailer (compression
  type, crc32c).  The index block maps "last key of data block" to that block's handle.
* key ``""`` holds a ``BundleHeaderProto``; every other key holds a
  ``BundleEntryProto`` with fields 1=dtype (1 = float32), 2=shape (repeated
  field 2 = dim, whose field 1 = size), 3=shard_id, 4=offset, 5=size, 6=crc32c.
* ``<prefix>.data-00000-of-00001`` is raw little-endian row-major data at those offsets.

This module is host-side product code (it is how weights reach the CUDA library); it
never touches ``oracle/``.
"""
from __future__ import annotations

import os
import re
import struct
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple

import numpy as np

_MAGIC = bytes([0x57, 0xFB, 0x80, 0x8B, 0x24, 0x75, 0x47, 0xDB])
_FOOTER_LEN = 48
_BLOCK_TRAILER_LEN = 5

_DT_FLOAT = 1
_DTYPES = {1: np.dtype("<f4"), 2: np.dtype("<f8"), 3: np.dtype("<i4"), 9: np.dtype("<i8")}


class CheckpointError(RuntimeError):
    pass


def _varint(buf: bytes, pos: int) -> Tuple[int, int]:
    result = 0
    shift = 0
    while True:
        if pos >= len(buf):
            raise CheckpointError("truncated varint")
        b = buf[pos]
        pos += 1
        result |= (b & 0x7F) << shift
        if not b & 0x80:
            return result, pos
        shift += 7
        if shift > 63:
            raise CheckpointError("varint too long")


def _read_block(data: bytes, offset: int, size: int) -> bytes:
    end = offset + size
    if end + _BLOCK_TRAILER_LEN > len(data):
        raise CheckpointError("block handle outside file")
    ctype = data[end]
    if ctype != 0:
        raise CheckpointError("compressed index blocks (type %d) are not supported" % ctype)
    return data[offset:end]


def _block_entries(block: bytes) -> List[Tuple[bytes, bytes]]:
    if len(block) < 4:
        raise CheckpointError("block too small")
    (num_restarts,) = struct.unpack_from("<I", block, len(block) - 4)
    limit = len(block) - 4 - 4 * num_restarts
    if limit < 0:
        raise CheckpointError("bad restart array")
    out = []
    pos = 0
    key = b""
    while pos < limit:
        shared, pos = _varint(block, pos)
        non_shared, pos = _varint(block, pos)
        vlen, pos = _varint(block, pos)
        if shared > len(key) or pos + non_shared + vlen > limit:
            raise CheckpointError("corrupt block entry")
        key = key[:shared] + block[pos:pos + non_shared]
        pos += non_shared
        out.append((key, block[pos:pos + vlen]))
        pos += vlen
    return out


def _proto_fields(buf: bytes) -> List[Tuple[int, int, object]]:
    """Minimal protobuf wire decoder: list of (field number, wire type, value)."""
    out = []
    pos = 0
    while pos < len(buf):
        tag, pos = _varint(buf, pos)
        fnum, wt = tag >> 3, tag & 7
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
            raise CheckpointError("unsupported protobuf wire type %d" % wt)
        out.append((fnum, wt, val))
    return out


@dataclass
class BundleEntry:
    name: str
    dtype: int = 0
    shape: Tuple[int, ...] = ()
    shard_id: int = 0
    offset: int = 0
    size: int = 0
    crc32c: Optional[int] = None


def _parse_entry(name: str, value: bytes) -> BundleEntry:
    e = BundleEntry(name=name)
    dims: List[int] = []
    for fnum, wt, val in _proto_fields(value):
        if fnum == 1 and wt == 0:
            e.dtype = int(val)
        elif fnum == 2 and wt == 2:
            for f2, w2, v2 in _proto_fields(val):  # TensorShapeProto
                if f2 == 2 and w2 == 2:  # dim
                    size = 0
                    for f3, w3, v3 in _proto_fields(v2):
                        if f3 == 1 and w3 == 0:
                            size = int(v3)
                    dims.append(size)
        elif fnum == 3 and wt == 0:
            e.shard_id = int(val)
        elif fnum == 4 and wt == 0:
            e.offset = int(val)
        elif fnum == 5 and wt == 0:
            e.size = int(val)
        elif fnum == 6 and wt == 5:
            (e.crc32c,) = struct.unpack("<I", val)
    e.shape = tuple(dims)
    return e


def read_index(index_path: str) -> Dict[str, BundleEntry]:
    """Parse ``<prefix>.index`` into ``{variable name: BundleEntry}``."""
    with open(index_path, "rb") as f:
        data = f.read()
    if len(data) < _FOOTER_LEN or data[-8:] != _MAGIC:
        raise CheckpointError("%s: not a TF bundle index (bad magic)" % index_path)
    footer = data[-_FOOTER_LEN:]
    pos = 0
    _meta_off, pos = _varint(footer, pos)
    _meta_size, pos = _varint(footer, pos)
    idx_off, pos = _varint(footer, pos)
    idx_size, pos = _varint(footer, pos)
    entries: Dict[str, BundleEntry] = {}
    for _last_key, handle in _block_entries(_read_block(data, idx_off, idx_size)):
        boff, p = _varint(handle, 0)
        bsize, p = _varint(handle, p)
        for key, value in _block_entries(_read_block(data, boff, bsize)):
            if key == b"":
                continue  # BundleHeaderProto
            name = key.decode("utf-8")
            entries[name] = _parse_entry(name, value)
    return entries


def checkpoint_prefix(model_dir: str) -> Optional[str]:
    """Resolve the bundle prefix the way ``tf.train.get_checkpoint_state`` does: read
    ``<dir>/checkpoint`` (text proto) and take ``model_checkpoint_path``.  Returns None when the
    directory holds no ``checkpoint`` file (the reference then silently loads nothing,
    ``mwis_dqn_call.py:189-192``)."""
    state = os.path.join(model_dir, "checkpoint")
    if not os.path.isfile(state):
        return None
    with open(state, "r") as f:
        text = f.read()
    m = re.search(r'^model_checkpoint_path:\s*"([^"]*)"', text, flags=re.M)
    if not m:
        return None
    path = m.group(1)
    if not os.path.isabs(path):
        path = os.path.join(model_dir, path)
    return path


def read_tensors(prefix: str, names=None) -> Dict[str, np.ndarray]:
    """Read tensors of a bundle.  ``names`` filters by exact name (default: all)."""
    entries = read_index(prefix + ".index")
    shards: Dict[int, bytes] = {}
    out: Dict[str, np.ndarray] = {}
    for name, e in entries.items():
        if names is not None and name not in names:
            continue
        if e.dtype not in _DTYPES:
            raise CheckpointError("%s: unsupported dtype enum %d" % (name, e.dtype))
        if e.shard_id not in shards:
            # single-shard bundles only: that is what tf.train.Saver writes
            shard_path = "%s.data-%05d-of-%05d" % (prefix, e.shard_id, 1)
            with open(shard_path, "rb") as f:
                shards[e.shard_id] = f.read()
        raw = shards[e.shard_id][e.offset:e.offset + e.size]
        dt = _DTYPES[e.dtype]
        count = int(np.prod(e.shape)) if e.shape else 1
        if len(raw) != e.size or count * dt.itemsize != e.size:
            raise CheckpointError("%s: size mismatch (shape %s, %d bytes)" % (name, e.shape, e.size))
        out[name] = np.frombuffer(raw, dtype=dt).reshape(e.shape).copy()
    return out


_VAR_RE = re.compile(r"^(?P<scope>[^/]+)/graphconvolution_(?P<layer>\d+)_vars/(?P<var>weights_(?P<k>\d+)|bias)$")


@dataclass
class LayerWeights:
    """One GraphConvolution layer: ``weights[k]`` is ``weights_k`` of shape ``[c_in, c_out]`` (one per
    support, reference ``gcn/layers.py:175-183``), ``bias`` is ``[c_out]`` or None."""
    weights: List[np.ndarray] = field(default_factory=list)
    bias: Optional[np.ndarray] = None

    @property
    def c_in(self) -> int:
        return int(self.weights[0].shape[0])

    @property
    def c_out(self) -> int:
        return int(self.weights[0].shape[1])


def load_gcn_weights(model_dir_or_prefix: str, scope: Optional[str] = None) -> List[LayerWeights]:
    """Load the GraphConvolution stack of a reference checkpoint as a list of LayerWeights ordered
    by layer id (1-based in the variable names because ``_build`` resets the layer uid counter,
    reference ``gcn/models.py:538``).  Stored shapes are authoritative (some directory names
    disagree with their contents)."""
    prefix = model_dir_or_prefix
    if os.path.isdir(model_dir_or_prefix):
        prefix = checkpoint_prefix(model_dir_or_prefix)
        if prefix is None:
            raise CheckpointError("%s: no 'checkpoint' state file" % model_dir_or_prefix)
    entries = read_index(prefix + ".index")
    wanted = {}
    for name in entries:
        m = _VAR_RE.match(name)
        if not m:
            continue
        if scope is not None and m.group("scope") != scope:
            continue
        wanted[name] = m
    if not wanted:
        raise CheckpointError("%s: no graphconvolution variables found" % prefix)
    scopes = {m.group("scope") for m in wanted.values()}
    if len(scopes) != 1:
        raise CheckpointError("%s: several model scopes %s, pass scope=" % (prefix, sorted(scopes)))
    tensors = read_tensors(prefix, names=set(wanted))
    by_layer: Dict[int, Dict[str, np.ndarray]] = {}
    for name, m in wanted.items():
        by_layer.setdefault(int(m.group("layer")), {})[m.group("var")] = tensors[name]
    layers: List[LayerWeights] = []
    for lid in sorted(by_layer):
        vars_ = by_layer[lid]
        ks = sorted(int(v.split("_")[1]) for v in vars_ if v.startswith("weights_"))
        if ks != list(range(len(ks))) or not ks:
            raise CheckpointError("layer %d: non-contiguous support weights %s" % (lid, ks))
        lw = LayerWeights(weights=[np.ascontiguousarray(vars_["weights_%d" % k], dtype=np.float32) for k in ks])
        for w in lw.weights:
            if w.ndim != 2 or w.shape != lw.weights[0].shape:
                raise CheckpointError("layer %d: inconsistent weight shapes" % lid)
        if "bias" in vars_:
            lw.bias = np.ascontiguousarray(vars_["bias"], dtype=np.float32).reshape(-1)
        layers.append(lw)
    for a, b in zip(layers[:-1], layers[1:]):
        if a.c_out != b.c_in:
            raise CheckpointError("layer widths do not chain: %d -> %d" % (a.c_out, b.c_in))
    return layers
