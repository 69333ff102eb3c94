#!/usr/bin/env python
"""Benchmark of the GCN-scored local-greedy MWIS path (BASELINE.json metric: graphs/sec).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME] [--main-only]

One "step" = one pass of the hot path (zero-weight removal -> GCN -> utility -> local greedy MWIS) over one batch of
conflict graphs.  The top-level keys of the printed JSON line are measured on BASELINE.json configs[1]: the 500 graphs of
BA_Graph_Uniform_GEN21_test2 (committed fixture tests/golden/ba_test2_full.npz) with the shipped checkpoint
result_IS4SAT_deep_ld1_c32_l20_cheb1_diver1_mwis_dqn.  With N GPUs every rank owns its own batch (graph batches shard
with no collective): weak scaling.

    value         graphs/s, inputs resident in HBM, CUDA events on the library's streams, exactly K steps
    e2e           graphs/s through the public host API (pinned host arrays in, membership out, copies timed);
                  e2e.from_reference_inputs: the same from a LIST OF SCIPY MATRICES (the reference's native input),
                  packing included (dg_solve_graphs_host)
    roofline      the dominant kernel against the roofline that bounds it (tensor pipe for tc_solve_kernel)
    cpu_baseline  the oracle port on this box's host cores (N = 1)
    configs       the other BASELINE configs, each measured the same way in the same run:
                  er500 (config 1), synth-er-<G> (a config-4 batch; carries roofline_streaming: the per-layer
                  streaming kernel on an input larger than L2 against the HBM roofline), per_graph_call (one graph per
                  call, the reference's call pattern), wireless (config 3: slots/s of the batched multi-channel slot
                  loop), partitioned (N >= 2: one large graph row-partitioned over the ranks, both exchange modes -
                  the path with real communication)

`--impl reference` times the CPU port of the reference path on the host cores with the same `config`.
"""
from __future__ import annotations

import argparse
import csv
import ctypes as C
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "graphs_per_sec_gcn_lgs"
UNIT = "graphs/s"
ROTATING_COPIES = 16  # distinct resident input sets cycled through the timed steps (> L2 in total)
NCU_TABLE = os.path.join(ROOT, "profiles", "ncu_traffic.csv")  # per (workload, kernel): ncu dram bytes / pipe figures
KERNEL_NOTES = {
    "tc_solve_kernel": "graph-resident, tcgen05: bf16-split projection + exact u8 aggregation MMAs, utility and greedy rounds "
                       "in one launch; adjacency, operands and accumulators stay in shared / tensor memory",
    "fused_solve_kernel": "graph-resident on CUDA cores: all GCN layers + utility + greedy rounds in one launch",
    "gs_layer_kernel": "streaming fused GraphConvolution layer, one graph per CTA iteration staged in shared memory",
    "gc_layer_kernel": "streaming fused GraphConvolution layer, warp per row, gathers through L1/L2",
}


# --------------------------------------------------------------------------------------------------
# workloads
# --------------------------------------------------------------------------------------------------
def shuffled_copy(pb, w, rng):
    """Same graphs in a different order (a distinct input set at distinct addresses)."""
    from distgcn_b200.batch import PackedBatch
    perm = rng.permutation(pb.n_graphs)
    sizes = pb.graph_sizes()[perm]
    gp = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
    rps = [np.zeros(1, np.int64)]
    cis, ws = [], []
    nnz = 0
    for k, g in enumerate(perm):
        v0, v1 = int(pb.graph_ptr[g]), int(pb.graph_ptr[g + 1])
        e0, e1 = int(pb.row_ptr[v0]), int(pb.row_ptr[v1])
        rps.append(pb.row_ptr[v0 + 1:v1 + 1].astype(np.int64) - e0 + nnz)
        cis.append(pb.col_idx[e0:e1].astype(np.int64) - v0 + gp[k])
        ws.append(w[v0:v1])
        nnz += e1 - e0
    return (PackedBatch(gp.astype(np.int32), np.concatenate(rps).astype(np.int32),
                        np.concatenate(cis).astype(np.int32)), np.concatenate(ws), perm)


def workload_info(name):
    """-> (checkpoint fixture, description, synthetic?)"""
    if name == "ba500":
        return "is4sat_l20_c32", ("BA_Graph_Uniform_GEN21_test2 (500 graphs, N=100-300) x IS4SAT c32 l20 checkpoint, "
                                  "local greedy MWIS"), False
    if name == "er500":
        return "is4sat_l1", "ER_Graph_Uniform_GEN21_test2 (500 graphs) x IS4SAT c32 l1 checkpoint, local greedy MWIS", False
    if name.startswith("synth-er-"):
        n_graphs = int(name.split("-")[-1])
        return "is4sat_l20_c32", ("synthetic G(N,0.1), N~U{100..300}, %d graphs x IS4SAT c32 l20 checkpoint (config-4 batch)"
                                  % n_graphs), True
    raise SystemExit("unknown workload %r" % name)


def load_host_workload(name, seed):
    """Fixture workloads (and small synthetic ones for the CPU arm) as host arrays."""
    from tests import util
    ck, desc, synth = workload_info(name)
    if name == "ba500":
        pb, w, _ = util.full_set("ba")
    elif name == "er500":
        pb, w, _ = util.full_set("er")
    else:
        pb, w = synth_er_batch_host(np.random.default_rng(seed), int(name.split("-")[-1]))
    return pb, w, util.load_layers(ck), desc


def synth_er_batch_host(rng, n_graphs, n_lo=100, n_hi=300, p=0.1):
    """Config-4 style synthetic G(N, p) graphs on the host (numpy only): the CPU arm's input."""
    from distgcn_b200.batch import pack_graphs
    import scipy.sparse as sp
    adjs = []
    for _ in range(n_graphs):
        n = int(rng.integers(n_lo, n_hi + 1))
        up = np.triu(rng.random((n, n)) < p, k=1)
        adjs.append(sp.csr_matrix((up | up.T).astype(np.float64)))
    pb = pack_graphs(adjs, check_values=False)
    return pb, rng.random(pb.n_nodes)


class GraphView:
    """Duck-typed per-graph matrix (indptr / indices / data / shape): what the native ingest reads of a scipy matrix."""
    __slots__ = ("indptr", "indices", "data", "shape", "format")

    def __init__(self, indptr, indices, n):
        self.indptr, self.indices, self.data, self.shape, self.format = indptr, indices, None, (n, n), "csr"


def per_graph_inputs(pb, w, scipy_objects):
    """The reference's native input for a packed batch: one matrix and one weight vector PER GRAPH.  For the fixture
    sets real scipy CSC matrices (what sio.loadmat returns); for large synthetic batches light views with the same
    attributes (building 16 k scipy objects would only time scipy)."""
    import scipy.sparse as sp
    local = pb.col_idx - np.repeat(pb.graph_ptr[:-1], pb.graph_nnz()).astype(np.int32)
    adjs, wl = [], []
    for g in range(pb.n_graphs):
        v0, v1 = int(pb.graph_ptr[g]), int(pb.graph_ptr[g + 1])
        e0, e1 = int(pb.row_ptr[v0]), int(pb.row_ptr[v1])
        ip = np.ascontiguousarray(pb.row_ptr[v0:v1 + 1] - e0, dtype=np.int32)
        ix = np.ascontiguousarray(local[e0:e1], dtype=np.int32)
        if scipy_objects:
            adjs.append(sp.csc_matrix((np.ones(e1 - e0), ix, ip), shape=(v1 - v0, v1 - v0)))
        else:
            adjs.append(GraphView(ip, ix, v1 - v0))
        wl.append(np.ascontiguousarray(w[v0:v1]))
    return adjs, wl


# --------------------------------------------------------------------------------------------------
# helpers
# --------------------------------------------------------------------------------------------------
class ClockSampler:
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device_index):
        self.device_index = device_index
        self.proc = None
        self.path = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(prefix="dg_clocks_", suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.device_index), "--query-gpu=" + self.QUERY,
                 "--format=csv,noheader,nounits", "-lms", "20"],
                stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        try:
            self.proc.terminate()
            self.proc.wait(timeout=5)
        except Exception:
            pass
        try:
            sm, mx = [], []
            reasons = set()
            for line in open(self.path):
                f = [x.strip() for x in line.split(",")]
                if len(f) < 9:
                    continue
                try:
                    sm.append(float(f[1]))
                    mx.append(float(f[2]))
                except ValueError:
                    continue
                for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"),
                                     f[5:9]):
                    if val.lower().startswith("active"):
                        reasons.add(name)
            if sm:
                out.update(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(mx)), reasons=sorted(reasons),
                           samples=len(sm))
        except Exception:
            pass
        finally:
            try:
                os.unlink(self.path)
            except Exception:
                pass
        return out


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            d = json.load(f)
        return {"hbm_gbs": float(d["hbm_gbs"]), "bf16_tflops": float(d["bf16_tflops"]),
                "source": "MEASURED_PEAKS.json (measured copy bandwidth / cuBLAS bf16 burst)"}
    except Exception:
        return {"hbm_gbs": 6650.0, "bf16_tflops": 1700.0,
                "source": "fallback of B200_PROFILING.md (MEASURED_PEAKS.json absent)"}


def ncu_row(workload, kernel):
    """Last committed ncu capture of `kernel` on `workload` (profiles/ncu_traffic.csv), or None."""
    try:
        hit = None
        with open(NCU_TABLE) as f:
            for r in csv.DictReader(l for l in f if not l.startswith("#")):
                if r["workload"] == workload and r["kernel"] == kernel:
                    hit = r
        return hit
    except Exception:
        return None


def tensor_roofline(pb, layers, launch_us, peaks, workload):
    """tc_solve_kernel against the tensor roofline: the dense MMA work the kernel ISSUES (bf16 3-term split projection,
    int8 digit aggregation over the dense adjacency; int8 counted at twice the bf16 rate) per launch / launch time,
    against the measured bf16 peak.  The sparse formulation's algorithmic flops are carried beside it."""
    n, nnz = float(pb.n_nodes), float(pb.nnz)
    hidden = [l for l in layers[1:-1]]
    alg = sum(2.0 * 2 * n * l.c_in * l.c_out + 2.0 * (nnz + n) * l.c_out + n * l.c_out for l in hidden)
    sizes = pb.graph_sizes().astype(np.int64)
    nb = (sizes + 127) // 128
    kp = (sizes + 31) // 32 * 32
    proj = float(nb.sum()) * 12 * 2 * 128 * 64 * 16 * len(hidden)          # 6 split products x 2 K steps, bf16
    agg = float((nb * kp).sum()) * 2 * 128 * 128 * len(hidden)             # u8 x s8 digits, N = 128
    t = launch_us * 1e-6
    achieved = (proj + agg / 2.0) / t / 1e12                               # bf16-equivalent TFLOP/s
    row = ncu_row(workload, "tc_solve_kernel")
    return {"bound": "tensor", "achieved": achieved, "peak": peaks["bf16_tflops"], "unit": "TFLOP/s",
            "frac": achieved / peaks["bf16_tflops"],
            "traffic": (float(row["dram_read_bytes"]) + float(row["dram_write_bytes"])) if row else None,
            "ncu": ({"tensor_pipe_active_pct": float(row["tensor_pipe_pct"]), "duration_us": float(row["duration_us"]),
                     "source": row["source"]} if row else None),
            "issued_bf16_tflops": proj / t / 1e12, "issued_int8_tops": agg / t / 1e12,
            "algorithmic_tflops_sparse_formulation": alg / t / 1e12,
            "note": "achieved = issued dense MMA work in bf16-equivalent TFLOP/s (int8 digits at twice the bf16 rate); the "
                    "kernel is bound by the dependent tensor-core / CUDA-core phases of a layer, not by HBM: its DRAM "
                    "traffic (`traffic`, ncu) is the CSR + model only"}


def dist_env():
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")),
            int(os.environ.get("WORLD_SIZE", "1")))


# --------------------------------------------------------------------------------------------------
# reference arm: the oracle port on the host cores
# --------------------------------------------------------------------------------------------------
def cpu_rate(pb, w, layers, n_graphs_sample, repeats, n_procs=0):
    """graphs/s of the CPU port on `n_graphs_sample` graphs, best of `repeats` passes."""
    from oracle import pipeline
    solver = pipeline.BatchSolver(pb.graph_ptr, pb.row_ptr, pb.col_idx, w, layers, "mwis", n_procs)
    try:
        solver.solve(0, min(pb.n_graphs, max(solver.n_procs, 8)))  # warm the workers
        best = None
        for _ in range(repeats):
            t0 = time.perf_counter()
            solver.solve(0, n_graphs_sample)
            dt = time.perf_counter() - t0
            best = dt if best is None or dt < best else best
        return n_graphs_sample / best, solver.n_procs
    finally:
        solver.close()


def bench_config(desc, n_graphs):
    """The `config` object, identical in both arms."""
    return {"workload": desc, "graphs_per_step_per_gpu": int(n_graphs)}


def run_reference(args):
    rank, _, world = dist_env()
    if rank != 0:
        return 0
    from oracle import pipeline
    name = args.workload
    if name.startswith("synth-er-") and int(name.split("-")[-1]) > 2000:
        name = "synth-er-2000"   # the CPU arm works on a bounded sample of the synthetic family
    pb, w, layers, desc = load_host_workload(name, args.seed)
    _, desc, _ = workload_info(args.workload)
    n_procs = os.cpu_count() or 1
    solver = pipeline.BatchSolver(pb.graph_ptr, pb.row_ptr, pb.col_idx, w, layers, "mwis", n_procs)
    try:
        # size the per-step sample so that the whole run stays within ~2 minutes
        probe = min(pb.n_graphs, 4 * n_procs)
        solver.solve(0, probe)
        t0 = time.perf_counter()
        solver.solve(0, probe)
        per_graph = (time.perf_counter() - t0) / probe
        budget = 120.0 / max(args.steps + args.warmup, 1)
        sample = int(max(n_procs, min(pb.n_graphs, budget / max(per_graph, 1e-9))))
        for _ in range(args.warmup):
            solver.solve(0, sample)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            solver.solve(0, sample)
        dt = time.perf_counter() - t0
    finally:
        solver.close()
    value = sample * args.steps / dt
    n_cfg = 500 if not args.workload.startswith("synth") else int(args.workload.split("-")[-1])
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32 scores, f64 utilities",
        "data": "reference dataset fixture (CPU port of the reference path; the reference itself is Python+TensorFlow and "
                "cannot run on this box)",
        "config": bench_config(desc, n_cfg),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": n_procs, "kind": "port",
                         "sample": "%d graphs per step, %d steps, %d worker processes (numpy/scipy GCN restatement + C local "
                                   "greedy search)" % (sample, args.steps, n_procs)},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit_line(line)
    return 0


# --------------------------------------------------------------------------------------------------
# our arm
# --------------------------------------------------------------------------------------------------
class Env:
    def __init__(self, args):
        import torch
        self.torch = torch
        self.args = args
        self.rank, self.local_rank, self.world = dist_env()
        if self.world != args.gpus and self.world > 1:
            raise SystemExit("--gpus %d but WORLD_SIZE=%d" % (args.gpus, self.world))
        self.use_dist = self.world > 1
        self.dist = None
        self.dev = torch.device("cuda", self.local_rank)
        torch.cuda.set_device(self.local_rank)
        if self.use_dist:
            import torch.distributed as dist
            self.dist = dist
            dist.init_process_group("nccl", device_id=self.dev)
        self.peaks = measured_peaks()

    def max_over_ranks(self, values):
        if not self.use_dist:
            return [float(v) for v in values]
        t = self.torch.tensor(list(values), dtype=self.torch.float64, device=self.dev)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return [float(x) for x in t.tolist()]

    def all_true(self, flag):
        if not self.use_dist:
            return bool(flag)
        t = self.torch.tensor([1 if flag else 0], dtype=self.torch.int32, device=self.dev)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MIN)
        return bool(int(t.item()))


def device_copies(env, name, seed, R):
    """R distinct input sets of the workload: resident on the device and in pinned host memory."""
    from distgcn_b200 import engine as E
    from distgcn_b200.batch import PackedBatch
    from tests import util
    torch = env.torch
    ck, desc, synth = workload_info(name)
    layers = util.load_layers(ck)
    sets = []
    if synth:
        from profiles.micro.stream_probe import er_batch_device
        n_graphs = int(name.split("-")[-1])
        for r in range(R):
            gp, rp, ci = er_batch_device(n_graphs, seed + 7919 * r + 104729 * env.rank, env.dev)
            g = torch.Generator(device=env.dev)
            g.manual_seed(seed + r)
            w = torch.rand(int(rp.numel()) - 1, dtype=torch.float64, device=env.dev, generator=g)
            torch.cuda.synchronize()
            pb = PackedBatch(gp.cpu().numpy(), rp.cpu().numpy(), ci.cpu().numpy())
            sets.append((pb, w.cpu().numpy()))
            del gp, rp, ci, w
    else:
        pb0, w0, _, _ = load_host_workload(name, seed)
        rng = np.random.default_rng(seed + 1000 * env.rank)
        for r in range(R):
            pb, w, _ = shuffled_copy(pb0, w0, rng) if r else (pb0, w0, None)
            sets.append((pb, w))
    return layers, desc, sets


def bench_batch_workload(env, name, steps, warmup, detail):
    """value / e2e / roofline of one batch workload.  `detail`: also the per-block statistics and the one-call-at-a-time
    e2e figure (the main workload)."""
    from distgcn_b200 import engine as E
    from distgcn_b200.batch import PackedBatch
    torch = env.torch
    args = env.args
    ck, desc, synth = workload_info(name)
    probe_nnz = 71e6 * int(name.split("-")[-1]) / 16384 if synth else 2e6
    R = 3 if probe_nnz * 4 > 100e6 else ROTATING_COPIES
    layers, desc, sets = device_copies(env, name, args.seed, R)
    # two contexts (streams) of the same GPU take the resident batches in turn, as engine.HostPipeline does for host
    # batches: the tail of one launch (SMs whose tiles are done) overlaps the head of the next
    NC = max(2, int(args.resident_contexts))
    ctxs = [E.Context(env.local_rank) for _ in range(NC)]
    ctx = ctxs[0]
    lib = ctx._lib
    models = [E.Model(c, layers, E.gcn_dqn_acts(len(layers))) for c in ctxs]
    copies = []
    input_bytes = 0
    for r, (pb, w) in enumerate(sets):
        dev_batch = E.DeviceBatch(ctxs[r % NC], pb)
        d_w = torch.from_numpy(w).to(env.dev)
        d_member = torch.empty(pb.n_nodes, dtype=torch.uint8, device=env.dev)
        d_total = torch.empty(pb.n_graphs, dtype=torch.float64, device=env.dev)
        # the upper host format: 16-bit graph-local column ids of the entries above the diagonal (dg_solve_host_upper)
        rp_u, c16 = pb.upper_compact()
        h = {k: E.pinned_empty(a.shape, a.dtype) for k, a in
             (("gp", pb.graph_ptr), ("rp", pb.row_ptr), ("w", w), ("c16", c16), ("rpu", rp_u))}
        h["gp"][:], h["rp"][:], h["w"][:], h["c16"][:], h["rpu"][:] = pb.graph_ptr, pb.row_ptr, w, c16, rp_u
        h["ci"] = pb.col_idx
        if detail:   # the packed int32 form, for the one-call-at-a-time figure
            h["ci"] = E.pinned_empty(pb.col_idx.shape, pb.col_idx.dtype)
            h["ci"][:] = pb.col_idx
        adjs, w_list = per_graph_inputs(pb, np.asarray(h["w"]), scipy_objects=not synth)
        copies.append(dict(pb=pb, w=w, dev=dev_batch, d_w=d_w, d_member=d_member, d_total=d_total,
                           h_pb=PackedBatch(h["gp"], h["rp"], h["ci"]), h_w=h["w"], h_upper=(h["rpu"], h["c16"]),
                           h_member=E.pinned_empty(pb.n_nodes, np.uint8), h_total=E.pinned_empty(pb.n_graphs, np.float64),
                           adjs=adjs, w_list=w_list))
        input_bytes += 4 * (pb.n_graphs + 1) + 4 * (pb.n_nodes + 1) + 4 * pb.nnz + 8 * pb.n_nodes
    n_graphs = sets[0][0].n_graphs
    c0 = sets[0][0]

    def barrier():
        for c in ctxs:
            c.synchronize()
        torch.cuda.synchronize()
        if env.use_dist:
            env.dist.barrier()

    def device_step(i):
        c = copies[i % R]
        k = (i % R) % NC
        E.solve_device(ctxs[k], models[k], c["dev"], c["d_w"], c["d_member"], predict="mwis", remove_zero_weight=True,
                       total=c["d_total"])

    def timed_block(k_steps, first):
        ev = DeviceTimer(ctx)
        ev.start()                                                    # start event on the first context's stream ...
        for c in ctxs[1:]:
            E.check(lib.dg_context_wait(c.handle, ctxs[0].handle))    # ... which the other contexts' work follows
        for i in range(k_steps):
            device_step(first + i)
        for c in ctxs[1:]:
            E.check(lib.dg_context_wait(ctxs[0].handle, c.handle))    # the stop event follows every stream's last kernel
        return ev.stop()  # synchronises

    # ---- value: inputs resident in HBM, CUDA events on the library's streams ----------------------------
    for i in range(max(warmup, R)):  # at least one untimed pass over every resident input set (tile plans are cached)
        device_step(i)
    barrier()
    launches0 = sum(c.launch_count for c in ctxs)
    t0 = time.perf_counter()
    dev_ms = timed_block(steps, warmup)
    wall_ms = 1e3 * (time.perf_counter() - t0)
    launches = sum(c.launch_count for c in ctxs) - launches0
    kernel_name = ctx.last_kernel
    barrier()
    rec = {}
    if detail:   # the K-step region is short: the same block repeated, so that its spread is on record
        blocks = []
        for b in range(5):
            blocks.append(timed_block(steps, warmup + (b + 1) * steps) / steps)
            barrier()
        blocks = env.max_over_ranks(blocks)
        rec["blocks"] = {"n": len(blocks), "steps_each": steps, "ms_per_step_median": float(np.median(blocks)),
                         "ms_per_step_min": float(min(blocks)), "ms_per_step_max": float(max(blocks)),
                         "note": "five more timed regions of K steps each after the headline one (max over ranks each)"}
    # roofline pass (not part of `value`): the same steps on ONE context, launches back to back without overlap, CUDA
    # events around every dominant-kernel launch (dg_profile_*)
    tot_ms, n_launch, alg_bytes = C.c_double(), C.c_uint64(), C.c_double()
    own = [r for r in range(R) if r % NC == 0]   # the input sets resident on the first context
    n_prof = max(3, min(steps, 50))
    lib.dg_profile_enable(ctx.handle, 1)
    ev = DeviceTimer(ctx)
    ev.start()
    for i in range(n_prof):
        device_step(own[i % len(own)])
    prof_ms = ev.stop()
    E.check(lib.dg_profile_collect(ctx.handle, C.byref(tot_ms), C.byref(n_launch), C.byref(alg_bytes)))
    lib.dg_profile_enable(ctx.handle, 0)
    barrier()

    # ---- e2e: public host API, pinned host buffers, copies inside the timed region ---------------------
    def host_step(i):
        c = copies[i % R]
        E.solve_host(ctx, models[0], c["h_pb"], c["h_w"], predict="mwis", remove_zero_weight=True,
                     member=c["h_member"], total=c["h_total"])

    e2e_sync_ms = None
    if detail:
        for i in range(3):
            host_step(i)
        barrier()
        t0 = time.perf_counter()
        for i in range(steps):
            host_step(i)
        e2e_sync_ms = 1e3 * (time.perf_counter() - t0)
        barrier()
    # the streaming form of the same API (engine.HostPipeline over dg_solve_host_compact): every step still copies its
    # own CSR + weights from pinned host memory and its membership + totals back, two contexts take turns so that one
    # batch's copies overlap the other's kernels.  This is the headline e2e number.
    pipe = E.HostPipeline(env.local_rank, layers, E.gcn_dqn_acts(len(layers)), depth=args.pipe_depth)

    def pipe_step(i):
        c = copies[i % R]
        pipe.submit(c["h_pb"], c["h_w"], c["h_member"], c["h_total"], predict="mwis", remove_zero_weight=True,
                    upper=c["h_upper"])

    def graphs_step(i):   # the reference's native input: a list of per-graph matrices + per-graph weight vectors
        c = copies[i % R]
        pipe.submit_graphs(c["adjs"], c["w_list"], c["h_member"], c["h_total"], predict="mwis", remove_zero_weight=True)

    def timed_pipe(step_fn):
        for i in range(max(4, warmup // 2, 3 * len(pipe.ctxs))):   # every context: buffers sized, tile-plan hint in place
            step_fn(i)
        pipe.wait()
        barrier()
        l0 = pipe.launch_count
        t0 = time.perf_counter()
        for i in range(steps):
            step_fn(i)
        pipe.wait()
        ms = 1e3 * (time.perf_counter() - t0)
        barrier()
        return ms, pipe.launch_count - l0

    e2e_ms, pipe_launches = timed_pipe(pipe_step)
    # results of the pipelined path equal the resident path's for every input set (not timed)
    same = True
    for r in range(min(R, steps)):
        device_step(r)
        barrier()
        same = same and bool(np.array_equal(copies[r]["d_member"].cpu().numpy(), np.asarray(copies[r]["h_member"])))
    ref_ms, ref_launches = timed_pipe(graphs_step)

    def timed_producers():
        """The same K steps from two producer threads, one per context of the pipeline (the documented threading model:
        a context belongs to one host thread): the Python-side walk over a list's scipy objects (GIL) overlaps the other
        thread's native packing (GIL released)."""
        import threading
        n_thr = len(pipe.ctxs)

        def run(k, lo, hi):
            for i in range(lo + k, hi, n_thr):
                c = copies[i % R]
                pipe.submit_graphs(c["adjs"], c["w_list"], c["h_member"], c["h_total"], predict="mwis",
                                   remove_zero_weight=True, slot=k)

        def go(lo, hi):
            ts = [threading.Thread(target=run, args=(k, lo, hi)) for k in range(n_thr)]
            for t in ts:
                t.start()
            for t in ts:
                t.join()
            pipe.wait()
        go(0, max(4, warmup // 2, 3 * n_thr))
        barrier()
        t0 = time.perf_counter()
        go(0, steps)
        ms = 1e3 * (time.perf_counter() - t0)
        barrier()
        return ms
    ref2_ms = timed_producers()
    # the same with the pointer tables of every input set built ONCE (batch.GraphTables): what a caller whose graphs stay
    # while the weights change - the wireless slot loop - pays per call
    from distgcn_b200.batch import GraphTables
    for c in copies:
        c["tables"] = GraphTables(c["adjs"], False)

    def tables_step(i):
        c = copies[i % R]
        pipe.submit_graphs(c["tables"], c["w_list"], c["h_member"], c["h_total"], predict="mwis", remove_zero_weight=True)
    ref3_ms, _ = timed_pipe(tables_step)
    for r in range(min(R, steps)):
        same = same and bool(np.array_equal(copies[r]["d_member"].cpu().numpy(), np.asarray(copies[r]["h_member"])))
    pipe.close()
    dev_ms, e2e_ms, ref_ms, wall_ms = env.max_over_ranks([dev_ms, e2e_ms, ref_ms, wall_ms])
    if e2e_sync_ms is not None:
        e2e_sync_ms, = env.max_over_ranks([e2e_sync_ms])
    same = env.all_true(same)

    world = env.world
    h2d = 4 * (c0.n_graphs + 1) + 4 * (c0.n_nodes + 1) + c0.nnz + 8 * c0.n_nodes  # upper format: nnz / 2 16-bit column ids
    d2h = c0.n_nodes + 8 * c0.n_graphs
    kern_launches = int(n_launch.value)
    avg_us = 1e3 * tot_ms.value / max(kern_launches, 1)
    rec.update({
        "workload": desc, "value": world * n_graphs * steps / (dev_ms / 1e3), "unit": UNIT,
        "ms_per_step": dev_ms / steps, "graphs_per_step_per_gpu": n_graphs, "nodes_per_step": int(c0.n_nodes),
        "nnz_per_step": int(c0.nnz), "gpu_launches": int(launches), "wall_ms_per_step": wall_ms / steps,
        "l2_policy": "inputs larger than L2: %d rotating resident input sets, %.0f MB in total" % (R, input_bytes / 1e6),
        "paths_agree": bool(same),
        "e2e": {"value": world * n_graphs * steps / (e2e_ms / 1e3), "unit": UNIT, "h2d_bytes_per_step": int(h2d),
                "d2h_bytes_per_step": int(d2h), "ms_per_step": e2e_ms / steps,
                "h2d_gbs_per_rank": h2d / (e2e_ms / steps) / 1e6,
                "api": "engine.HostPipeline.submit (dg_solve_host_upper, %d contexts in turn): pinned host arrays - row offsets and "
                       "16-bit graph-local column ids of the entries above the diagonal, weights - in, membership + totals out, "
                       "every step; wall clock over the K steps" % args.pipe_depth,
                "gpu_launches": int(pipe_launches),
                "from_reference_inputs": {
                    "value": world * n_graphs * steps / (ref_ms / 1e3), "unit": UNIT, "ms_per_step": ref_ms / steps,
                    "gpu_launches": int(ref_launches),
                    "api": "engine.HostPipeline.submit_graphs (dg_solve_graphs_host): a Python list of per-graph %s and "
                           "per-graph weight vectors in, membership out; packing (host threads of the library, into pinned "
                           "staging), copies and kernels all inside the timed region"
                           % ("scipy CSC matrices" if not synth else "CSR array pairs"),
                    "prebuilt_pointer_tables": {
                        "value": world * n_graphs * steps / (ref3_ms / 1e3), "unit": UNIT, "ms_per_step": ref3_ms / steps,
                        "api": "submit_graphs(batch.GraphTables built once per list, per-step weight vectors): packing, copies and "
                               "kernels per step, the walk over the scipy objects once"},
                    "two_producer_threads": {
                        "value": world * n_graphs * steps / (ref2_ms / 1e3), "unit": UNIT, "ms_per_step": ref2_ms / steps,
                        "api": "the same call from two host threads, one per context of the pipeline (submit_graphs(slot = k)): "
                               "one thread's walk over its list's scipy objects overlaps the other's native packing"}}},
    })
    if e2e_sync_ms is not None:
        rec["e2e"]["one_call_at_a_time"] = {"value": world * n_graphs * steps / (e2e_sync_ms / 1e3),
                                            "ms_per_step": e2e_sync_ms / steps, "api": "dg_solve_host (packed int32 column ids)"}
    # ---- roofline of the dominant kernel --------------------------------------------------------------
    if kernel_name == "tc_solve_kernel":
        roof = tensor_roofline(c0, layers, avg_us, env.peaks, name)
        roof["work_equivalent_hbm"] = {
            "achieved_gbs": (alg_bytes.value / 1e9) / (tot_ms.value / 1e3) if tot_ms.value > 0 else 0.0,
            "algorithmic_bytes_per_launch": alg_bytes.value / max(kern_launches, 1),
            "note": "SURVEY 8d bytes of the same layers as streaming passes / kernel time: a measure of work, NOT of HBM "
                    "pressure (the kernel moves `traffic` bytes)"}
    else:
        achieved = (alg_bytes.value / 1e9) / (tot_ms.value / 1e3) if tot_ms.value > 0 else 0.0
        row = ncu_row(name, kernel_name)
        roof = {"bound": "hbm", "achieved": achieved, "peak": env.peaks["hbm_gbs"], "unit": "GB/s",
                "frac": achieved / env.peaks["hbm_gbs"],
                "traffic": (float(row["dram_read_bytes"]) + float(row["dram_write_bytes"])) if row else None,
                "algorithmic_bytes_per_launch": alg_bytes.value / max(kern_launches, 1)}
    roof.update({"kernel": "%s (%s)" % (kernel_name, KERNEL_NOTES.get(kernel_name, "")), "launches_timed": kern_launches,
                 "avg_launch_us": avg_us, "share_of_step": tot_ms.value / prof_ms if prof_ms > 0 else None,
                 "measured_on": "%d launches back to back on one context after the timed region (the two-stream overlap of "
                                "`value` leaves no per-kernel duration); %.4f ms per step there" % (n_prof, prof_ms / n_prof),
                 "peak_source": env.peaks["source"]})
    rec["roofline"] = roof
    for m in models:
        m.close()
    rec["_keep"] = (copies, ctxs, layers, sets)   # the caller may reuse the resident inputs (streaming roofline)
    return rec


def bench_streaming(env, name, rec):
    """The per-layer STREAMING path (what SURVEY.md 8d's HBM roofline is about) on the config-4 batch - inputs far larger
    than L2, so every layer's feature rows really cross HBM.  DG_DISABLE_FUSED=1 routes the same dg_solve call through
    the per-layer kernels; CUDA events around every layer-kernel launch (dg_profile_*)."""
    from distgcn_b200 import engine as E
    copies, ctxs, layers, sets = rec["_keep"]
    ctx = ctxs[0]
    lib = ctx._lib
    model = E.Model(ctx, layers, E.gcn_dqn_acts(len(layers)))
    c = copies[0]
    os.environ["DG_DISABLE_FUSED"] = "1"
    E.reload_env()   # the library reads its options once per context, not on the solve path
    try:
        for _ in range(2):
            E.solve_device(ctx, model, c["dev"], c["d_w"], c["d_member"], predict="mwis", remove_zero_weight=True,
                           total=c["d_total"])
        ctx.synchronize()
        tot_ms, n_launch, alg_bytes = C.c_double(), C.c_uint64(), C.c_double()
        lib.dg_profile_enable(ctx.handle, 1)
        ev = DeviceTimer(ctx)
        ev.start()
        reps = 3
        for _ in range(reps):
            E.solve_device(ctx, model, c["dev"], c["d_w"], c["d_member"], predict="mwis", remove_zero_weight=True,
                           total=c["d_total"])
        solve_ms = ev.stop() / reps
        E.check(lib.dg_profile_collect(ctx.handle, C.byref(tot_ms), C.byref(n_launch), C.byref(alg_bytes)))
        lib.dg_profile_enable(ctx.handle, 0)
        kernel = ctx.last_kernel
    finally:
        del os.environ["DG_DISABLE_FUSED"]
        E.reload_env()
    model.close()
    pb = c["pb"]
    n_l = max(int(n_launch.value), 1)
    achieved = (alg_bytes.value / 1e9) / (tot_ms.value / 1e3) if tot_ms.value > 0 else 0.0
    row = ncu_row(name, kernel)
    return {"bound": "hbm", "kernel": "%s (%s)" % (kernel, KERNEL_NOTES.get(kernel, "")),
            "achieved": achieved, "peak": env.peaks["hbm_gbs"], "unit": "GB/s", "frac": achieved / env.peaks["hbm_gbs"],
            "traffic": (float(row["dram_read_bytes"]) + float(row["dram_write_bytes"])) if row else None,
            "ncu_source": row["source"] if row else None,
            "algorithmic_bytes_per_launch": alg_bytes.value / n_l, "avg_launch_us": 1e3 * tot_ms.value / n_l,
            "launches_timed": int(n_launch.value), "ms_per_solve_streaming_path": solve_ms,
            "share_of_solve": tot_ms.value / (solve_ms * reps) if solve_ms > 0 else None,
            "input": "%d graphs, %d vertices, %d nnz: %.0f MB of feature rows in + out per layer (L2: 126 MB)"
                     % (pb.n_graphs, pb.n_nodes, pb.nnz, 2 * 4 * 32 * pb.n_nodes / 1e6),
            "peak_source": env.peaks["source"]}


def bench_spmm(env, name, rec):
    """The stand-alone SpMM Y = L.Z (the reference's sparse_tensor_dense_matmul(support[1], pre_sup), 32 columns) on the
    config-4 batch: B_spmm of SURVEY.md 8d / kernel time, CUDA events around every launch."""
    from distgcn_b200 import engine as E
    torch = env.torch
    copies, ctxs, layers, sets = rec["_keep"]
    ctx = ctxs[0]
    lib = ctx._lib
    c = copies[0]
    pb = c["pb"]
    z = torch.randn(pb.n_nodes, 32, dtype=torch.float32, device=env.dev)
    y = torch.empty_like(z)
    torch.cuda.synchronize()
    for _ in range(3):
        E.spmm_laplacian(ctx, c["dev"], z, y)
    ctx.synchronize()
    tot_ms, n_launch, alg_bytes = C.c_double(), C.c_uint64(), C.c_double()
    lib.dg_profile_enable(ctx.handle, 1)
    for _ in range(20):
        E.spmm_laplacian(ctx, c["dev"], z, y)
    ctx.synchronize()
    E.check(lib.dg_profile_collect(ctx.handle, C.byref(tot_ms), C.byref(n_launch), C.byref(alg_bytes)))
    lib.dg_profile_enable(ctx.handle, 0)
    kernel = ctx.last_kernel
    n_l = max(int(n_launch.value), 1)
    achieved = (alg_bytes.value / 1e9) / (tot_ms.value / 1e3) if tot_ms.value > 0 else 0.0
    row = ncu_row(name, kernel)
    return {"bound": "hbm", "kernel": "%s (Y = L.Z alone, rows staged in shared memory and pre-scaled by dinv)" % kernel,
            "achieved": achieved, "peak": env.peaks["hbm_gbs"], "unit": "GB/s", "frac": achieved / env.peaks["hbm_gbs"],
            "traffic": (float(row["dram_read_bytes"]) + float(row["dram_write_bytes"])) if row else None,
            "ncu_source": row["source"] if row else None, "algorithmic_bytes_per_launch": alg_bytes.value / n_l,
            "avg_launch_us": 1e3 * tot_ms.value / n_l, "launches_timed": int(n_launch.value),
            "peak_source": env.peaks["source"]}


def bench_single_calls(env, n_calls=200):
    """The reference's real call pattern: ONE graph per call (wireless_dqn_test_mc.py:289,323 call
    dqn_agent.solve_mwis(adj, wts) per time slot).  Host wall clock per call, scipy CSC matrix in, Python set out."""
    import scipy.sparse as sp
    from distgcn_b200.mwis_dqn_call import DQNAgent
    from distgcn_b200.runtime_config import make_flags
    from tests import util
    out = {}
    for tag, fam, ck, nl in (("ba_l20", "ba", "is4sat_l20_c32", 20), ("er_l1", "er", "is4sat_l1", 1)):
        pb, w, _ = util.full_set(fam)
        agent = DQNAgent(1, 5000, flags=make_flags(feature_size=1, hidden1=32, num_layer=nl, diver_num=1, max_degree=1,
                                                   predict="mwis"), device=env.local_rank)
        agent.load(util.ckpt_dir(ck))
        agent.check_values = False
        k = min(n_calls, pb.n_graphs)
        graphs = [(sp.csc_matrix(pb.graph_adj(g)), np.ascontiguousarray(w[pb.graph_ptr[g]:pb.graph_ptr[g + 1]]))
                  for g in range(k)]
        for a, wg in graphs[:20]:
            agent.solve_mwis(a, wg)
        ts = []
        for a, wg in graphs:
            t0 = time.perf_counter()
            agent.solve_mwis(a, wg)
            ts.append(time.perf_counter() - t0)
        ts = np.asarray(ts) * 1e6
        out[tag] = {"calls": int(k), "us_per_call_median": float(np.median(ts)), "us_per_call_mean": float(ts.mean()),
                    "us_per_call_p95": float(np.quantile(ts, 0.95)), "graphs_per_s": float(1e6 / ts.mean())}
    out["api"] = "DQNAgent.solve_mwis(adj, wts) -> (set, total, 1.0), one scipy CSC matrix per call, host wall clock"
    return out


def bench_wireless(env, n_networks=20, timeslots=100):
    """BASELINE config 3: the multi-channel wireless slot loop (wireless_dqn_test_mc.py:225-366) on synthetic networks
    with the reference's constants, every (network, load) instance advanced together, queue bookkeeping resident on the
    device (distgcn_b200.wireless.BatchedScheduler.run).  Slots/s of the whole sweep, host wall clock."""
    from distgcn_b200 import engine as E
    from distgcn_b200 import wireless as W
    from tests import util
    loads = np.round(np.arange(0.1, 1.25, 0.1), 2)   # bash/twc_major_wireless_mc_test.sh
    insts = W.make_instances(n_networks, loads, n_ch=3, timeslots=timeslots, seed=0)
    out = {"workload": "synthetic stand-in for data/wireless_test: %d networks x %d loads = %d instances, 3 channels, %d slots, "
                       "joint conflict graphs of %.0f vertices on average"
                       % (n_networks, len(loads), len(insts), timeslots - 1, float(np.mean([i.adj_gK.shape[0] for i in insts]))),
           "instances": len(insts)}
    ctx = E.Context(env.local_rank)
    for ck in ("is4sat_l1", "is4sat_l20_c32"):
        layers = util.load_layers(ck)
        model = E.Model(ctx, layers, E.gcn_dqn_acts(len(layers)))
        for algo in ("DGCN-LGS", "DGCN-LGS-Seq") + (("Greedy",) if ck == "is4sat_l1" else ()):
            sim = W.BatchedScheduler(ctx, insts, algo, model)
            sim.run(5)
            t0 = time.perf_counter()
            qs = sim.run()
            dt = time.perf_counter() - t0
            out["%s/%s" % (ck, algo)] = {"slots_per_s": qs.shape[0] / dt, "instance_slots_per_s": qs.shape[0] * len(insts) / dt,
                                         "solver_graphs_per_s": qs.shape[0] * sim.graphs_per_slot / dt}
            sim.close()
        model.close()
    ctx.close()
    return out


def bench_partitioned(env, n, deg, reps=3):
    """ONE G(n, m = n*deg/2) graph row-partitioned over the ranks (SURVEY.md 8e, config 5 shape; c64 l2 checkpoint): both
    exchange modes, device-timed (max over ranks), membership compared with the single-GPU solve of the same graph."""
    from distgcn_b200 import engine as E
    from distgcn_b200.shard import RowPartitionedSolver, row_slices
    from profiles.micro.stream_probe import big_er_device
    from tests import util
    torch, dist = env.torch, env.dist
    dev, rank, world = env.dev, env.rank, env.world
    gp, rp, ci = big_er_device(n, deg, 0, dev)
    g = torch.Generator(device=dev)
    g.manual_seed(1)
    w = torch.rand(n, dtype=torch.float64, device=dev, generator=g)
    w[torch.rand(n, device=dev, generator=g) < 0.05] = 0.0
    torch.cuda.synchronize()
    per, n_pad = row_slices(n, world)
    r0, r1 = rank * per, min(n, (rank + 1) * per)
    rp_l = torch.zeros(per + 1, dtype=torch.int32, device=dev)
    e0 = int(rp[r0]) if r1 > r0 else 0
    if r1 > r0:
        rp_l[: r1 - r0 + 1] = rp[r0:r1 + 1] - e0
        rp_l[r1 - r0 + 1:] = rp_l[r1 - r0]
        ci_l = ci[e0:int(rp[r1])].clone()
    else:
        ci_l = ci[:0].clone()
    layers = util.load_layers("is4sat_l2_c64")
    acts = E.gcn_dqn_acts(len(layers))

    class Spec:
        def __init__(self):
            self.layers = [type("L", (), {"act_code": a})() for a in acts]

        def layers_as_weights(self):
            return layers
    out = {"workload": "one synthetic G(n = %d, average degree %d) graph, %d directed nnz, c64 l2 checkpoint, row-partitioned "
                       "over %d GPUs" % (n, deg, int(ci.numel()), world), "n": n, "nnz": int(ci.numel()), "world": world}
    members = {}
    for ex in ("nccl", "p2p"):
        solver = RowPartitionedSolver(Spec(), n, rp_l, ci_l, rank=rank, world_size=world, exchange=ex)
        solver.keep_on_device = True
        solver.solve(w[r0:r1])
        solver.exchanged_bytes = 0
        torch.cuda.synchronize()
        dist.barrier()
        t0 = time.perf_counter()
        for _ in range(reps):
            member, score, rounds = solver.solve(w[r0:r1])
        torch.cuda.synchronize()
        dt, = env.max_over_ranks([(time.perf_counter() - t0) / reps])
        out[ex] = {"ms_per_solve": 1e3 * dt, "rounds": int(rounds), "vertices_per_s": n / dt,
                   "exchanged_MB_per_rank_per_solve": solver.exchanged_bytes / reps / 1e6}
        full = torch.zeros(n_pad, dtype=torch.uint8, device=dev)
        dist.all_gather_into_tensor(full, member.contiguous())
        members[ex] = full[:n].clone()
        solver.close()
    flags = torch.zeros(4, dtype=torch.float64, device=dev)
    if rank == 0:
        ctx = E.Context(env.local_rank)
        model = E.Model(ctx, layers, acts)
        batch = E.DeviceBatch(ctx, graph_ptr=gp, row_ptr=rp, col_idx=ci)
        ref = torch.empty(n, dtype=torch.uint8, device=dev)
        E.solve_device(ctx, model, batch, w, ref)
        ctx.synchronize()
        t0 = time.perf_counter()
        for _ in range(reps):
            E.solve_device(ctx, model, batch, w, ref)
        ctx.synchronize()
        single_ms = 1e3 * (time.perf_counter() - t0) / reps
        out["single_gpu"] = {"ms_per_solve": single_ms}
        for ex in ("nccl", "p2p"):
            out[ex]["membership_equal_to_single_gpu"] = bool(torch.equal(members[ex], ref))
            out[ex]["speedup_vs_single_gpu"] = single_ms / out[ex]["ms_per_solve"]
            out[ex]["efficiency"] = single_ms / out[ex]["ms_per_solve"] / world
        batch.close()
        model.close()
        ctx.close()
    dist.barrier()
    del flags
    return out


def run_ours(args):
    env = Env(args)
    rank, world = env.rank, env.world
    sampler = ClockSampler(env.local_rank)  # samples from the warm-up to the end of the timed regions
    if rank == 0:
        sampler.start()
    main_rec = bench_batch_workload(env, args.workload, args.steps, args.warmup, detail=True)
    clocks = sampler.stop() if rank == 0 else None
    main_layers, main_sets = main_rec["_keep"][2], main_rec["_keep"][3]
    del main_rec["_keep"]
    configs = {}
    if not args.main_only:
        secondary = [w for w in ("ba500", "er500") if w != args.workload and not args.workload.startswith("synth")]
        secondary.append("synth-er-%d" % args.synth_graphs)
        for name in secondary:
            rec = bench_batch_workload(env, name, max(3, min(args.steps, 20)) if name.startswith("synth") else args.steps,
                                       args.warmup, detail=False)
            if name.startswith("synth"):
                rec["roofline_streaming"] = bench_streaming(env, name, rec)
                rec["roofline_spmm"] = bench_spmm(env, name, rec)
            keep = rec.pop("_keep")
            if name == "er500" and world == 1 and not args.no_cpu_baseline and rank == 0:
                pb0, w0 = keep[3][0]
                rate, cores = cpu_rate(pb0, w0, keep[2], min(pb0.n_graphs, 500), repeats=3)
                rec["cpu_baseline"] = {"value": rate, "unit": UNIT, "cores": cores, "kind": "port",
                                       "sample": "the 500 graphs of the workload, best of 3 passes, %d worker processes" % cores}
            del keep
            configs[name] = rec
        if rank == 0:
            configs["per_graph_call"] = bench_single_calls(env)
            configs["wireless"] = bench_wireless(env)
        if env.use_dist:
            env.dist.barrier()
            configs["partitioned"] = bench_partitioned(env, args.part_nodes or 6000000 * world, 16)
    if rank == 0:
        roof = main_rec.pop("roofline")
        line = {
            "metric": METRIC, "value": main_rec["value"], "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": main_rec["ms_per_step"], "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32 scores, f64 utilities",
            "data": "reference dataset fixture (real BA/ER test2 graphs + shipped checkpoint)" if not
                    args.workload.startswith("synth") else "synthetic",
            "config": bench_config(main_rec["workload"], main_rec["graphs_per_step_per_gpu"]),
            "run_info": {"nodes_per_step": main_rec["nodes_per_step"], "nnz_per_step": main_rec["nnz_per_step"],
                         "parallelism": "graph-batch sharding, no collectives",
                         "streams": "%d contexts of the library on the GPU take the resident steps in turn, %d the end-to-end ones"
                                    % (max(2, args.resident_contexts), args.pipe_depth),
                         "l2_policy": main_rec["l2_policy"], "paths_agree": main_rec["paths_agree"],
                         "wall_ms_per_step": main_rec["wall_ms_per_step"], "blocks": main_rec.get("blocks")},
            "e2e": main_rec["e2e"],
            "gpu_launches": main_rec["gpu_launches"],
            "roofline": roof,
            "clocks": clocks,
        }
        synth = configs.get("synth-er-%d" % args.synth_graphs)
        if synth and "roofline_streaming" in synth:
            line["roofline_streaming"] = synth["roofline_streaming"]
            line["roofline_spmm"] = synth.get("roofline_spmm")
        if not args.no_cpu_baseline and world == 1:
            pb0, w0 = main_sets[0]
            sample = min(pb0.n_graphs, 500)
            rate, cores = cpu_rate(pb0, w0, main_layers, sample, repeats=3)
            line["cpu_baseline"] = {"value": rate, "unit": UNIT, "cores": cores, "kind": "port",
                                    "sample": "%d graphs of the same workload, best of 3 passes, %d worker processes "
                                              "(numpy/scipy GCN restatement + C local greedy search; the reference's own "
                                              "Python/TensorFlow code cannot run on this box)" % (sample, cores)}
        if configs:
            line["configs"] = configs
        emit_line(line)
    if env.use_dist:
        env.dist.barrier()
        env.dist.destroy_process_group()
    return 0


class DeviceTimer:
    """CUDA events recorded on the library's own stream (dg_timer_start / dg_timer_stop): the kernels
    are launched on that private stream, which torch.cuda.Event would not see."""

    def __init__(self, ctx):
        self.ctx = ctx

    def start(self):
        from distgcn_b200 import engine as E
        E.check(self.ctx._lib.dg_timer_start(self.ctx.handle))

    def stop(self):
        from distgcn_b200 import engine as E
        ms = C.c_double()
        E.check(self.ctx._lib.dg_timer_stop(self.ctx.handle, C.byref(ms)))
        return float(ms.value)


_JSON_OUT = None


def claim_stdout():
    """Native libraries (NCCL prints its version line) write to file descriptor 1: from here on fd 1 is stderr and the
    one JSON line goes to the original stdout."""
    global _JSON_OUT
    if _JSON_OUT is None:
        sys.stdout.flush()
        _JSON_OUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def emit_line(line):
    out = _JSON_OUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="ba500")
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--main-only", action="store_true", help="skip the `configs` block (other workloads)")
    ap.add_argument("--resident-contexts", type=int, default=2,
                    help="contexts (streams) the resident batches are taken on in turn: one launch's tail overlaps the next one's head")
    ap.add_argument("--pipe-depth", type=int, default=4, help="contexts the end-to-end pipeline takes its steps on in turn")
    ap.add_argument("--synth-graphs", type=int, default=16384, help="graphs in the config-4 batch of `configs`")
    ap.add_argument("--part-nodes", type=int, default=0,
                    help="vertices of the row-partitioned graph (N >= 2); 0 = 6 M per GPU (48 M at N = 8: BASELINE config 5 is 50 M)")
    args = ap.parse_args()
    claim_stdout()
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
