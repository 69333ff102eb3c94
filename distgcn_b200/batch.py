"""Host-side packed-CSR container for batches of conflict graphs (the ingest format of the CUDA
library, see include/distgcn_b200.h "Data layout").

The reference handles one scipy matrix per call (``adj`` as loaded from the ``.mat`` files written by
Data_Generation.py:214-219: float64 CSC, symmetric, zero diagonal).  Here thousands of such graphs
are packed into one CSR so a single kernel launch covers all of them.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Iterable, List, Optional, Sequence, Tuple

import numpy as np
import scipy.sparse as sp


def _pattern_csr(adj) -> Tuple[np.ndarray, np.ndarray]:
    """(indptr, indices) of the non-zero pattern of one adjacency matrix - what
    ``np.nonzero(adj[v])`` enumerates in the reference (heuristics.py:94)."""
    if sp.issparse(adj):
        a = adj.tocsr()
        if a.nnz and (a.data == 0).any():
            a = a.copy()
            a.eliminate_zeros()
    else:
        a = sp.csr_matrix(np.asarray(adj))
    if a.shape[0] != a.shape[1]:
        raise ValueError("adjacency matrix must be square, got %s" % (a.shape,))
    return a.indptr, a.indices


@dataclass
class PackedBatch:
    """n_graphs graphs as one CSR; column indices are batch-global vertex ids."""
    graph_ptr: np.ndarray  # int32 [n_graphs + 1]
    row_ptr: np.ndarray    # int32 [n_nodes + 1]
    col_idx: np.ndarray    # int32 [nnz]

    @property
    def n_graphs(self) -> int:
        return int(self.graph_ptr.shape[0] - 1)

    @property
    def n_nodes(self) -> int:
        return int(self.row_ptr.shape[0] - 1)

    @property
    def nnz(self) -> int:
        return int(self.col_idx.shape[0])

    def graph_sizes(self) -> np.ndarray:
        return np.diff(self.graph_ptr)

    def graph_nnz(self) -> np.ndarray:
        return np.diff(self.row_ptr[self.graph_ptr])

    def local_columns(self) -> np.ndarray:
        """uint16 [nnz]: column ids local to their graph (what each per-graph scipy matrix of the reference holds) -
        the compact host format of ``engine.solve_host(..., col_local16=...)`` / dg_solve_host_compact, half the
        host->device bytes of ``col_idx``.  Graphs of at most 65536 vertices."""
        if self.n_graphs and int(self.graph_sizes().max()) > 65536:
            raise ValueError("16-bit local column ids need graphs of at most 65536 vertices")
        base = np.repeat(self.graph_ptr[:-1].astype(np.int64), self.graph_nnz())
        return (self.col_idx.astype(np.int64) - base).astype(np.uint16)

    def upper_compact(self):
        """(row_ptr_upper int32 [n_nodes + 1], col_local_upper uint16 [nnz / 2]): only the entries with column > row,
        graph-local ids - the UPPER host format of ``engine.solve_host(..., upper=...)`` / dg_solve_host_upper, a third of
        the host->device bytes of ``col_idx``.  The adjacency must be symmetric with a zero diagonal; graphs of at most
        8192 vertices."""
        if self.n_graphs and int(self.graph_sizes().max()) > 8192:
            raise ValueError("the upper host format takes graphs of at most 8192 vertices")
        rows = np.repeat(np.arange(self.n_nodes, dtype=np.int64), np.diff(self.row_ptr))
        cols = self.col_idx.astype(np.int64)
        keep = cols > rows
        if int(keep.sum()) * 2 != self.nnz:
            raise ValueError("adjacency pattern is not symmetric with a zero diagonal")
        base = np.repeat(self.graph_ptr[:-1].astype(np.int64), self.graph_sizes())[rows[keep]]
        rp_u = np.zeros(self.n_nodes + 1, dtype=np.int64)
        np.add.at(rp_u, rows[keep] + 1, 1)
        return np.cumsum(rp_u).astype(np.int32), (cols[keep] - base).astype(np.uint16)

    def slice(self, g0: int, g1: int) -> "PackedBatch":
        """Graphs g0 .. g1-1 as their own batch (vertex ids re-based)."""
        v0, v1 = int(self.graph_ptr[g0]), int(self.graph_ptr[g1])
        e0, e1 = int(self.row_ptr[v0]), int(self.row_ptr[v1])
        return PackedBatch(
            graph_ptr=(self.graph_ptr[g0:g1 + 1] - v0).astype(np.int32),
            row_ptr=(self.row_ptr[v0:v1 + 1] - e0).astype(np.int32),
            col_idx=(self.col_idx[e0:e1] - v0).astype(np.int32),
        )

    def graph_adj(self, g: int) -> sp.csr_matrix:
        """Graph g back as a scipy CSR matrix (float64 ones), for interoperability."""
        sub = self.slice(g, g + 1)
        n = sub.n_nodes
        return sp.csr_matrix((np.ones(sub.nnz), sub.col_idx, sub.row_ptr), shape=(n, n))

    def validate(self) -> None:
        """Structural checks the kernels rely on: in-range, same-graph, zero-diagonal, symmetric."""
        n = self.n_nodes
        if self.graph_ptr[0] != 0 or self.graph_ptr[-1] != n or (np.diff(self.graph_ptr) < 0).any():
            raise ValueError("graph_ptr must be non-decreasing from 0 to n_nodes")
        if self.row_ptr[0] != 0 or self.row_ptr[-1] != self.nnz or (np.diff(self.row_ptr) < 0).any():
            raise ValueError("row_ptr must be non-decreasing from 0 to nnz")
        if self.nnz == 0:
            return
        rows = np.repeat(np.arange(n, dtype=np.int64), np.diff(self.row_ptr))
        cols = self.col_idx.astype(np.int64)
        if cols.min() < 0 or cols.max() >= n:
            raise ValueError("col_idx out of range")
        gid = np.searchsorted(self.graph_ptr, rows, side="right") - 1
        if (cols < self.graph_ptr[gid]).any() or (cols >= self.graph_ptr[gid + 1]).any():
            raise ValueError("an edge crosses a graph boundary")
        if (rows == cols).any():
            raise ValueError("self-loops are not allowed (the reference's greedy search never terminates on them)")
        a = sp.csr_matrix((np.ones(self.nnz, dtype=np.int8), cols, self.row_ptr.astype(np.int64)), shape=(n, n))
        if (a != a.T).nnz:
            raise ValueError("adjacency pattern is not symmetric")


def _as_csr32(adj):
    """A matrix the native packer can read: CSR / CSC with int32 C-contiguous indptr / indices and float64 data.
    (The adjacency is symmetric, so the CSC arrays the .mat files load as are the CSR arrays as well.)"""
    if sp.issparse(adj) and adj.format in ("csr", "csc"):
        a = adj
    elif sp.issparse(adj):
        a = adj.tocsr()
    else:
        a = sp.csr_matrix(np.asarray(adj))
    if a.shape[0] != a.shape[1]:
        raise ValueError("adjacency matrix must be square, got %s" % (a.shape,))
    if a.indptr.dtype != np.int32 or a.indices.dtype != np.int32 or a.data.dtype != np.float64 or \
            not (a.indptr.flags.c_contiguous and a.indices.flags.c_contiguous and a.data.flags.c_contiguous):
        a = type(a)((np.ascontiguousarray(a.data, dtype=np.float64), np.ascontiguousarray(a.indices, dtype=np.int32),
                     np.ascontiguousarray(a.indptr, dtype=np.int32)), shape=a.shape)
    return a


class GraphTables:
    """Pointer tables of a list of per-graph matrices for the native ingest entry points (dg_pack_graphs_host,
    dg_solve_graphs_host): built by the C helper ``_pyingest.collect`` without a Python loop.  Holds the matrices'
    arrays alive.  ``check_values``: also pass the stored values so that stored zeros are not taken for edges."""

    def __init__(self, adjs, check_values: bool = True):
        from . import _pyingest
        adjs = adjs if isinstance(adjs, (list, tuple)) else list(adjs)
        try:
            tabs = _pyingest.collect(adjs, check_values)
        except TypeError:
            adjs = [_as_csr32(a) for a in adjs]     # other formats / index types: normalise once, in Python
            tabs = _pyingest.collect(adjs, check_values)
        self.indptr, self.indices, self.data, self.n_rows_raw, self._keep, self.n_nodes = tabs
        self._adjs = adjs
        self.n_graphs = len(adjs)
        self.n_rows = np.frombuffer(self.n_rows_raw, dtype=np.int32)
        for a in adjs[:1] + adjs[-1:]:
            shp = getattr(a, "shape", None)
            if shp is not None and len(shp) == 2 and shp[0] != shp[1]:
                raise ValueError("adjacency matrix must be square, got %s" % (shp,))


def pack_graphs(adjs: Iterable, check_values: bool = True, n_threads: int = 0) -> PackedBatch:
    """Pack scipy / dense adjacency matrices into one PackedBatch - natively: the per-graph arrays are handed to
    dg_pack_graphs_host as pointer tables and packed by the library's host threads (no per-graph Python work).
    Replaces the per-call networkx conversions of mwis_dqn_call.py:202-207."""
    import ctypes as C

    from . import _lib
    lib = _lib.load()
    t = adjs if isinstance(adjs, GraphTables) else GraphTables(adjs, check_values)
    n_nodes, nnz = C.c_int64(), C.c_int64()
    _lib.check(lib.dg_pack_graphs_sizes(t.n_graphs, t.indptr, t.data, t.n_rows_raw, C.byref(n_nodes), C.byref(nnz), None))
    gp = np.empty(t.n_graphs + 1, dtype=np.int32)
    rp = np.empty(n_nodes.value + 1, dtype=np.int32)
    ci = np.empty(nnz.value, dtype=np.int32)
    _lib.check(lib.dg_pack_graphs_host(t.n_graphs, t.indptr, t.indices, t.data, t.n_rows_raw, gp.ctypes.data, rp.ctypes.data,
                                       ci.ctypes.data if nnz.value else np.empty(1, np.int32).ctypes.data, None,
                                       int(n_threads)))
    return PackedBatch(graph_ptr=gp, row_ptr=rp, col_idx=ci)


def pack_graphs_python(adjs: Iterable) -> PackedBatch:
    """The pure numpy/scipy packer (one conversion per graph): kept as the independent check of the native one."""
    gp: List[int] = [0]
    rps: List[np.ndarray] = [np.zeros(1, dtype=np.int64)]
    cis: List[np.ndarray] = []
    nnz = 0
    for adj in adjs:
        indptr, indices = _pattern_csr(adj)
        n = indptr.shape[0] - 1
        rps.append(indptr[1:].astype(np.int64) + nnz)
        cis.append(indices.astype(np.int64) + gp[-1])
        nnz += int(indptr[-1])
        gp.append(gp[-1] + n)
    if gp[-1] >= 2 ** 31 or nnz >= 2 ** 31:
        raise ValueError("batch too large for int32 indexing: %d nodes, %d nnz" % (gp[-1], nnz))
    return PackedBatch(
        graph_ptr=np.asarray(gp, dtype=np.int32),
        row_ptr=np.concatenate(rps).astype(np.int32),
        col_idx=(np.concatenate(cis) if cis else np.zeros(0)).astype(np.int32),
    )


def from_edge_lists(graph_ptr: Sequence[int], edge_ptr: Sequence[int], edge_u: np.ndarray,
                    edge_v: np.ndarray) -> PackedBatch:
    """Build a PackedBatch from per-graph undirected edge lists with graph-local endpoints (the
    compact form of the tests/golden/*_full.npz fixtures)."""
    graph_ptr = np.asarray(graph_ptr, dtype=np.int64)
    edge_ptr = np.asarray(edge_ptr, dtype=np.int64)
    n = int(graph_ptr[-1])
    per_graph = np.diff(edge_ptr)
    off = np.repeat(graph_ptr[:-1], per_graph)
    u = edge_u.astype(np.int64) + off
    v = edge_v.astype(np.int64) + off
    rows = np.concatenate([u, v])
    cols = np.concatenate([v, u])
    a = sp.csr_matrix((np.ones(rows.shape[0], dtype=np.int8), (rows, cols)), shape=(n, n))
    a.sort_indices()
    return PackedBatch(graph_ptr=graph_ptr.astype(np.int32), row_ptr=a.indptr.astype(np.int32),
                       col_idx=a.indices.astype(np.int32))


def partition_by_work(batch: PackedBatch, n_parts: int, node_cost: float = 8.0) -> List[Tuple[int, int]]:
    """Contiguous graph ranges [g0, g1) with balanced ``nnz + node_cost * n_nodes`` (SURVEY.md 8e:
    independent graphs shard with no collective).  Always returns n_parts ranges (possibly empty)."""
    if n_parts < 1:
        raise ValueError("n_parts must be >= 1")
    work = batch.graph_nnz().astype(np.float64) + node_cost * batch.graph_sizes().astype(np.float64)
    csum = np.concatenate([[0.0], np.cumsum(work)])
    total = csum[-1]
    bounds = [0]
    for p in range(1, n_parts):
        target = total * p / n_parts
        g = int(np.searchsorted(csum, target, side="left"))
        # pick the boundary closest to the target
        if g > 0 and abs(csum[g - 1] - target) <= abs(csum[min(g, batch.n_graphs)] - target):
            g -= 1
        g = min(max(g, bounds[-1]), batch.n_graphs)
        bounds.append(g)
    bounds.append(batch.n_graphs)
    return [(bounds[i], bounds[i + 1]) for i in range(n_parts)]
