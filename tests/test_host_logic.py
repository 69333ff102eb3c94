"""CPU tests of the host-side logic: checkpoint reader, packing, sharding (gloo, world_size 2), the
API mirrors, and the C-ABI surface (symbols only - no compute without a GPU)."""
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import scipy.sparse as sp

from distgcn_b200 import batch as B
from distgcn_b200 import ckpt
from tests import util

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


# ---- checkpoints ------------------------------------------------------------------------------------
def test_checkpoint_reader_shapes_and_values():
    l1 = util.load_layers("is4sat_l1")
    assert len(l1) == 1 and l1[0].weights[0].shape == (1, 1) and len(l1[0].weights) == 2
    # the two scalars quoted in SURVEY.md section 5
    assert abs(float(l1[0].weights[0][0, 0]) - 0.66316652) < 1e-7
    assert abs(float(l1[0].weights[1][0, 0]) - 0.76819468) < 1e-7
    l20 = util.load_layers("is4sat_l20_c32")
    assert [(l.c_in, l.c_out) for l in l20] == [(1, 32)] + [(32, 32)] * 18 + [(32, 1)]
    assert all(l.bias is None for l in l20)
    bias = util.load_layers("dqnmed_l1_bias")
    assert bias[0].bias is not None and bias[0].bias.shape == (1,)
    ld32 = util.load_layers("is4sat_ld32_l3_c32")
    assert ld32[0].c_in == 32


def test_checkpoint_reader_rejects_garbage(tmp_path):
    p = tmp_path / "model.ckpt.index"
    p.write_bytes(b"not a table" * 10)
    with pytest.raises(ckpt.CheckpointError):
        ckpt.read_index(str(p))
    assert ckpt.checkpoint_prefix(str(tmp_path)) is None  # no `checkpoint` state file: silently nothing


def test_adam_slots_are_ignored():
    entries = ckpt.read_index(os.path.join(util.ckpt_dir("is4sat_l1"), "model.ckpt.index"))
    assert any("Adam" in k for k in entries) and "beta1_power" in entries
    assert all("Adam" not in n for n in ["weights_0", "weights_1"])


# ---- packing ----------------------------------------------------------------------------------------
def test_pack_slice_roundtrip():
    rng = np.random.default_rng(0)
    pb, adjs = util.random_graph_batch(rng, 7, 5, 40)
    pb.validate()
    assert pb.n_graphs == 7 and pb.n_nodes == sum(a.shape[0] for a in adjs)
    for g, a in enumerate(adjs):
        assert (pb.graph_adj(g) != a).nnz == 0
    sub = pb.slice(2, 5)
    sub.validate()
    assert sub.n_graphs == 3
    assert (sub.graph_adj(0) != adjs[2]).nnz == 0


def test_pack_ignores_explicit_zeros_and_accepts_dense():
    a = sp.csr_matrix((np.array([1.0, 0.0, 1.0]), (np.array([0, 0, 1]), np.array([1, 2, 0]))), shape=(3, 3))
    pb = B.pack_graphs([a, np.array([[0, 1], [1, 0]])])
    assert pb.nnz == 4 and pb.n_nodes == 5
    assert pb.col_idx.tolist() == [1, 0, 4, 3]


def test_validate_catches_bad_batches():
    bad = B.PackedBatch(np.array([0, 2], np.int32), np.array([0, 1, 1], np.int32), np.array([1], np.int32))
    with pytest.raises(ValueError):
        bad.validate()  # not symmetric
    loop = B.PackedBatch(np.array([0, 1], np.int32), np.array([0, 1], np.int32), np.array([0], np.int32))
    with pytest.raises(ValueError):
        loop.validate()


def test_edge_list_fixture_equals_small_fixture():
    pb_full, w_full, z = util.full_set("er")
    assert pb_full.n_graphs == 500 and pb_full.n_nodes == 100000
    pb_full.slice(0, 20).validate()
    pb_small, w_small = util.small_graphs()
    names_full = [str(x) for x in z["names"]]
    names_small = [str(x) for x in util.load_npz("graphs_small.npz")["names"]]
    g_small = 3
    g_full = names_full.index(names_small[g_small])
    assert (pb_full.graph_adj(g_full) != pb_small.graph_adj(g_small)).nnz == 0


def test_partition_by_work_balances_and_covers():
    pb, w, _ = util.full_set("ba")
    for parts in (1, 2, 3, 8):
        ranges = B.partition_by_work(pb, parts)
        assert len(ranges) == parts and ranges[0][0] == 0 and ranges[-1][1] == pb.n_graphs
        assert all(ranges[i][1] == ranges[i + 1][0] for i in range(parts - 1))
        work = [float(pb.row_ptr[pb.graph_ptr[b]] - pb.row_ptr[pb.graph_ptr[a]]) for a, b in ranges]
        assert max(work) <= 1.15 * (sum(work) / parts) + 12000


# ---- sharding over gloo -----------------------------------------------------------------------------
_WORKER = r'''
import os, sys
sys.path.insert(0, %(root)r)
import numpy as np
import torch.distributed as dist
from tests import util
from distgcn_b200.shard import ShardedSolver
from oracle import lgs as L

def oracle_solve(sub, w):  # CPU stand-in for the per-rank GPU solve: plain LGS on the raw weights
    r = L.run_batch(sub.graph_ptr, sub.row_ptr, sub.col_idx, w)
    tot = np.array([w[sub.graph_ptr[g]:sub.graph_ptr[g + 1]][r.member[sub.graph_ptr[g]:sub.graph_ptr[g + 1]] == 1].sum()
                    for g in range(sub.n_graphs)])
    return r.member, tot

dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%(port)d", rank=int(sys.argv[1]), world_size=2)
pb, w = util.small_graphs()
out = ShardedSolver(oracle_solve).solve(pb, w)
if dist.get_rank() == 0:
    ref = util.load_npz("lgs_ref_small.npz")
    assert np.array_equal(out[0], ref["member"]), "sharded membership differs"
    assert np.allclose(out[1], ref["total"])
    print("SHARD_OK")
else:
    assert out is None
dist.barrier()
dist.destroy_process_group()
'''


def test_sharded_solver_gloo_world2(tmp_path):
    import socket
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    script = tmp_path / "worker.py"
    script.write_text(_WORKER % {"root": ROOT, "port": port})
    procs = [subprocess.Popen([sys.executable, str(script), str(r)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
             for r in range(2)]
    outs = [p.communicate(timeout=240)[0].decode() for p in procs]
    assert all(p.returncode == 0 for p in procs), "\n".join(outs)
    assert "SHARD_OK" in outs[0]


# ---- API mirrors (no GPU needed to build them) ----------------------------------------------------------
def test_model_mirrors_build_and_load():
    from distgcn_b200 import layers as L
    from distgcn_b200 import models as M
    from distgcn_b200.directory import find_model_folder
    from distgcn_b200.runtime_config import make_flags
    fl = make_flags(feature_size=1, hidden1=32, num_layer=20, diver_num=1, training_set="IS4SAT")
    assert find_model_folder(fl, "dqn") == "./model/result_IS4SAT_deep_ld1_c32_l20_cheb1_diver1_mwis_dqn"
    m = M.GCN_DQN(L.make_placeholders(2, 1), input_dim=1, flags=fl)
    assert len(m.layers) == 20
    assert "gcn_dqn/graphconvolution_1_vars/weights_0:0" in m.vars
    assert [l.act_code for l in m.layers] == [1] * 19 + [0]
    m.load(util.ckpt_dir("is4sat_l20_c32"))
    ref = util.load_layers("is4sat_l20_c32")
    assert np.array_equal(m.layers[7].vars["weights_1"], ref[7].weights[1])
    d = M.GCN_DEEP_DIVER(L.make_placeholders(2, 1), input_dim=1, flags=make_flags(feature_size=1, hidden1=16, num_layer=3, diver_num=2))
    assert d.layers[-1].output_dim == 4
    g2 = M.GCN2_DQN(L.make_placeholders(2, 4), hidden_dim=8, num_layer=3, bias=True, flags=make_flags(feature_size=4))
    assert all(l.bias for l in g2.layers) and [l.act_code for l in g2.layers] == [1, 1, 1]
    with pytest.raises(AssertionError):
        L.GraphConvolution(1, 1, L.make_placeholders(2), bogus=1)
    with pytest.raises(NameError):
        L.GraphConvolution(1, 1, L.make_placeholders(2), flags=make_flags(wts_init="ones"))
    assert L.act_code(lambda x: x) == 0 and L.act_code(lambda x: np.maximum(x, 0)) == 2


# ---- C-ABI surface -----------------------------------------------------------------------------------
def test_library_exports_every_declared_symbol():
    from distgcn_b200 import _lib
    from distgcn_b200 import build as dg_build
    dg_build.build_library()
    header = open(os.path.join(ROOT, "include", "distgcn_b200.h")).read()
    declared = set(re.findall(r"DG_API[^;]*?\b(dg_[a-z0-9_]+)\s*\(", header))
    assert len(declared) >= 25
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    lib = _lib.load()  # resolves every symbol or raises AttributeError
    assert lib.dg_version() == 100
    nm = subprocess.check_output(["nm", "-D", "--defined-only", _lib.LIB_PATH]).decode()
    exported = set(re.findall(r" T (dg_[a-z0-9_]+)", nm))
    assert declared <= exported
    assert all(s.startswith("dg_") for s in exported), "internal symbols leak: %s" % (exported - declared)


def test_no_device_fails_loudly():
    from distgcn_b200 import _lib, engine
    if _lib.load().dg_device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(_lib.DistGCNError) as ei:
        engine.Context(0)
    assert ei.value.code == _lib.ERR_NO_DEVICE


def test_product_never_imports_oracle():
    """The product path must not route through the CPU oracle (nor read /root/reference)."""
    pkg = os.path.join(ROOT, "distgcn_b200")
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, fn)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), fn
                assert "liblgs_oracle" not in text, fn
                assert "/root/reference" not in text, fn


def test_mat_dataset_ingest(tmp_path):
    """load_mat_dataset reads the reference's .mat layout (Data_Generation.py:214-219): CSC float64 adj,
    weights [1, N]."""
    import scipy.io as sio
    import scipy.sparse as sp
    from distgcn_b200.mwis_dqn_test import load_mat_dataset
    rng = np.random.default_rng(0)
    ref = []
    for k, n in enumerate((7, 12, 1)):
        upper = np.triu(rng.random((n, n)) < 0.3, k=1)
        adj = sp.csc_matrix((upper | upper.T).astype(np.float64))
        w = rng.random((1, n))
        sio.savemat(str(tmp_path / ("ER_n%d_p0.3_b%d_uni.mat" % (n, k))),
                    {"adj": adj, "weights": w, "N": n, "p": 0.3, "greedy_utility": float(k), "mwis_utility": 1.0})
        ref.append((adj, w.reshape(-1)))
    names, packed, wts, extras = load_mat_dataset(str(tmp_path))
    assert names == sorted(names) and len(names) == 3
    order = [1, 2, 0]  # sorted file names: n12 (b1), n1 (b2), n7 (b0)
    assert packed.n_graphs == 3 and packed.n_nodes == 20
    off = 0
    for g, k in enumerate(order):
        adj, w = ref[k]
        n = adj.shape[0]
        assert (packed.graph_adj(g) != adj.tocsr()).nnz == 0
        assert np.array_equal(wts[off:off + n], w)
        off += n
    assert np.array_equal(extras["greedy_utility"], np.array([1.0, 2.0, 0.0]))
    packed.validate()


def test_local_columns_compact_format():
    """PackedBatch.local_columns(): 16-bit graph-local column ids (the compact host format of dg_solve_host_compact)."""
    pb, _ = util.small_graphs()
    c16 = pb.local_columns()
    assert c16.dtype == np.uint16 and c16.flags.c_contiguous and c16.shape == pb.col_idx.shape
    base = np.repeat(pb.graph_ptr[:-1].astype(np.int64), pb.graph_nnz())
    assert np.array_equal(c16.astype(np.int64) + base, pb.col_idx)
    sizes = np.repeat(pb.graph_sizes().astype(np.int64), pb.graph_nnz())
    assert (c16 < sizes).all()
    sub = pb.slice(7, 19)                      # slices re-base their ids: the local form is unchanged
    e0, e1 = int(pb.row_ptr[pb.graph_ptr[7]]), int(pb.row_ptr[pb.graph_ptr[19]])
    assert np.array_equal(sub.local_columns(), c16[e0:e1])
    from distgcn_b200.batch import PackedBatch
    empty = PackedBatch(np.zeros(1, np.int32), np.zeros(1, np.int32), np.zeros(0, np.int32))
    assert empty.local_columns().shape == (0,)
    big = PackedBatch(np.array([0, 70000], np.int32), np.zeros(70001, np.int32), np.zeros(0, np.int32))
    with pytest.raises(ValueError):
        big.local_columns()


# ---- native ingest (dg_pack_graphs_host; host code of the library, no GPU needed) ------------------------
def _random_adjs(rng, count, n_lo=1, n_hi=120):
    out = []
    for _ in range(count):
        n = int(rng.integers(n_lo, n_hi + 1))
        up = np.triu(rng.random((n, n)) < rng.uniform(0.02, 0.3), k=1)
        out.append(sp.csc_matrix((up | up.T).astype(np.float64)))     # the .mat files load as CSC
    return out


def test_native_packer_equals_python_packer():
    rng = np.random.default_rng(8)
    adjs = _random_adjs(rng, 300)
    a, b = B.pack_graphs(adjs), B.pack_graphs_python(adjs)
    for x, y in ((a.graph_ptr, b.graph_ptr), (a.row_ptr, b.row_ptr), (a.col_idx, b.col_idx)):
        assert x.dtype == np.int32 and np.array_equal(x, y)
    a.validate()
    # single thread, all threads, and the value check switched off give the same arrays
    for kw in (dict(n_threads=1), dict(n_threads=64), dict(check_values=False)):
        c = B.pack_graphs(adjs, **kw)
        assert np.array_equal(c.col_idx, b.col_idx) and np.array_equal(c.row_ptr, b.row_ptr)
    # a large batch (threads engaged: > 64 k entries)
    big = _random_adjs(rng, 60, 250, 300)
    assert np.array_equal(B.pack_graphs(big).col_idx, B.pack_graphs_python(big).col_idx)


def test_native_packer_input_forms_and_errors():
    from distgcn_b200 import _lib
    rng = np.random.default_rng(9)
    adjs = _random_adjs(rng, 6, 5, 40)
    mixed = [adjs[0].tolil(), adjs[1].toarray(), sp.csr_matrix((0, 0)), adjs[2].tocsr(), adjs[3].tocoo(),
             sp.csr_matrix((adjs[4].data, adjs[4].indices.astype(np.int64), adjs[4].indptr.astype(np.int64)),
                           shape=adjs[4].shape)]
    a, b = B.pack_graphs(mixed), B.pack_graphs_python(mixed)
    assert np.array_equal(a.graph_ptr, b.graph_ptr) and np.array_equal(a.row_ptr, b.row_ptr)
    assert np.array_equal(a.col_idx, b.col_idx)
    # stored zeros are not edges (np.nonzero(adj[v]), heuristics.py:94) ...
    x = adjs[5].tocsr().copy()
    x.data[::3] = 0.0
    a, b = B.pack_graphs([adjs[0], x, adjs[1]]), B.pack_graphs_python([adjs[0], x, adjs[1]])
    assert a.nnz == b.nnz < adjs[0].nnz + x.nnz + adjs[1].nnz and np.array_equal(a.col_idx, b.col_idx)
    assert np.array_equal(a.row_ptr, b.row_ptr)
    # ... unless the caller vouches for a pattern-only input
    assert B.pack_graphs([x], check_values=False).nnz == x.nnz
    # empty list, empty graphs
    e = B.pack_graphs([])
    assert e.n_graphs == 0 and e.n_nodes == 0 and e.nnz == 0
    e = B.pack_graphs([sp.csr_matrix((3, 3)), sp.csr_matrix((0, 0))])
    assert e.graph_ptr.tolist() == [0, 3, 3] and e.row_ptr.tolist() == [0, 0, 0, 0]
    # malformed patterns are refused with the graph's index
    bad = adjs[2].tocsr().copy()
    bad.indices = bad.indices.copy()
    bad.indices[0] = bad.shape[0] + 7
    with pytest.raises(_lib.DistGCNError) as ei:
        B.pack_graphs([adjs[0], bad])
    assert ei.value.code == _lib.ERR_INVALID and "graph 1" in str(ei.value)
    with pytest.raises(ValueError):
        B.pack_graphs([np.zeros((2, 3))])
    # the 16-bit compact form straight from the per-graph arrays equals PackedBatch.local_columns()
    import ctypes as C
    lib = _lib.load()
    t = B.GraphTables(adjs, True)
    pb = B.pack_graphs(adjs)
    gp = np.empty(t.n_graphs + 1, np.int32)
    rp = np.empty(pb.n_nodes + 1, np.int32)
    c16 = np.empty(pb.nnz, np.uint16)
    _lib.check(lib.dg_pack_graphs_host(t.n_graphs, t.indptr, t.indices, t.data, t.n_rows_raw, gp.ctypes.data, rp.ctypes.data,
                                       None, c16.ctypes.data, 0))
    assert np.array_equal(c16, pb.local_columns()) and np.array_equal(rp, pb.row_ptr) and np.array_equal(gp, pb.graph_ptr)
    with pytest.raises(_lib.DistGCNError):   # exactly one column output
        _lib.check(lib.dg_pack_graphs_host(t.n_graphs, t.indptr, t.indices, t.data, t.n_rows_raw, gp.ctypes.data,
                                           rp.ctypes.data, None, None, 0))


def test_upper_format_packer_and_python_form_agree():
    """dg_pack_graphs_upper_host (native, from the per-graph matrices) == PackedBatch.upper_compact() (numpy, from the packed
    batch); stored zeros dropped; asymmetric or diagonal-carrying input refused."""
    import ctypes as C
    from distgcn_b200 import _lib
    lib = _lib.load()
    rng = np.random.default_rng(31)
    adjs = _random_adjs(rng, 80, 1, 200) + [sp.csr_matrix((0, 0)), sp.csr_matrix((3, 3))]
    pb = B.pack_graphs(adjs)
    rp_ref, c_ref = pb.upper_compact()
    assert c_ref.shape[0] * 2 == pb.nnz
    t = B.GraphTables(adjs, True)
    gp = np.empty(t.n_graphs + 1, np.int32)
    rp = np.empty(pb.n_nodes + 1, np.int32)
    cu = np.empty(pb.nnz // 2, np.uint16)
    for threads in (1, 0):
        _lib.check(lib.dg_pack_graphs_upper_host(t.n_graphs, t.indptr, t.indices, t.data, t.n_rows_raw, gp.ctypes.data,
                                                 rp.ctypes.data, cu.ctypes.data, threads))
        assert np.array_equal(gp, pb.graph_ptr) and np.array_equal(rp, rp_ref) and np.array_equal(cu, c_ref)
    # stored zeros (a symmetric set of pairs) are not edges
    x = adjs[5].tocsr().copy()
    rows = np.repeat(np.arange(x.shape[0]), np.diff(x.indptr))
    x.data = np.where((rows + x.indices) % 3 == 0, 0.0, x.data)
    y = x.copy()
    y.eliminate_zeros()
    tx = B.GraphTables([x], True)
    rp1, cu1 = np.empty(x.shape[0] + 1, np.int32), np.empty(y.nnz // 2, np.uint16)
    gp1 = np.empty(2, np.int32)
    _lib.check(lib.dg_pack_graphs_upper_host(1, tx.indptr, tx.indices, tx.data, tx.n_rows_raw, gp1.ctypes.data, rp1.ctypes.data,
                                             cu1.ctypes.data, 1))
    rp_y, c_y = B.pack_graphs([y]).upper_compact()
    assert np.array_equal(rp1, rp_y) and np.array_equal(cu1, c_y)
    # not symmetric / diagonal entries: refused
    bad = sp.csr_matrix(np.array([[0, 1, 1], [0, 0, 0], [0, 0, 0]], dtype=float))    # two entries, both above the diagonal
    tb = B.GraphTables([bad], False)
    with pytest.raises(_lib.DistGCNError):
        _lib.check(lib.dg_pack_graphs_upper_host(1, tb.indptr, tb.indices, None, tb.n_rows_raw, gp1.ctypes.data,
                                                 np.empty(4, np.int32).ctypes.data, np.empty(4, np.uint16).ctypes.data, 1))
    skew = sp.csr_matrix(np.array([[0, 1, 1], [1, 0, 0], [0, 1, 0]], dtype=float))   # half above, half below, not symmetric
    ts = B.GraphTables([skew], False)
    with pytest.raises(_lib.DistGCNError):
        _lib.check(lib.dg_pack_graphs_upper_host(1, ts.indptr, ts.indices, None, ts.n_rows_raw, gp1.ctypes.data,
                                                 np.empty(4, np.int32).ctypes.data, np.empty(4, np.uint16).ctypes.data, 1))
    diag = sp.csr_matrix(np.array([[1, 1], [1, 0]], dtype=float))                     # odd count
    td = B.GraphTables([diag], False)
    with pytest.raises(_lib.DistGCNError):
        _lib.check(lib.dg_pack_graphs_upper_host(1, td.indptr, td.indices, None, td.n_rows_raw, gp1.ctypes.data,
                                                 np.empty(3, np.int32).ctypes.data, np.empty(4, np.uint16).ctypes.data, 1))
    with pytest.raises(ValueError):
        B.pack_graphs([bad]).upper_compact()
