/*
 * distgcn_b200 - C-ABI of the B200-native GCN-scored local-greedy MWIS path.
 *
 * This is the drop-in boundary.  The reference (zhongyuanzhao/distgcn) is pure Python on TensorFlow,
 * so it has no FFI of its own; each entry point below names the reference Python interface it
 * replaces (paths relative to the reference root) and INTEGRATION.md shows the ctypes stub a
 * maintainer of the reference would add.  No torch / C++ types cross this boundary: plain pointers,
 * sizes and opaque handles only.
 *
 * Memory spaces.  Every call that moves data takes `int mem`:
 *   DG_MEM_HOST   - all data pointers of the call are host pointers; the library stages them through
 *                   its own device buffers and the call returns after the stream has drained.
 *   DG_MEM_DEVICE - all data pointers are device pointers valid on the context's device; the call
 *                   only enqueues work on the context's stream (no synchronisation, zero copy).
 *
 * Data layout ("packed batch"): n_graphs independent graphs stored as one CSR over n_nodes rows.
 *   graph_ptr[g] .. graph_ptr[g+1]-1   are the vertices of graph g            (int32, n_graphs+1)
 *   row_ptr[v]   .. row_ptr[v+1]-1     index the neighbours of vertex v       (int32, n_nodes+1)
 *   col_idx[e]                         batch-global vertex id of a neighbour  (int32, nnz)
 * The pattern must be symmetric with a zero diagonal (the reference's conflict graphs,
 * Data_Generation.py:214-219); edge weights are implicit 1.  A vertex's index inside its graph is
 * v - graph_ptr[g]; ascending batch-global id == ascending local id, which is what the index
 * tie-break of the greedy heuristic needs.
 *
 * Threading: a dg_context is bound to one device and one stream; use one context per host thread.
 * Handles created from a context must be destroyed before the context.
 */
#ifndef DISTGCN_B200_H
#define DISTGCN_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DG_VERSION 100

#if defined(__GNUC__)
#define DG_API __attribute__((visibility("default")))
#else
#define DG_API
#endif

/* status codes */
#define DG_OK 0
#define DG_ERR_INVALID 1        /* bad argument (message in dg_last_error) */
#define DG_ERR_CUDA 2           /* a CUDA runtime call failed */
#define DG_ERR_UNSUPPORTED 3    /* shape outside what the kernels cover (e.g. width > 64) */
#define DG_ERR_NOT_CONVERGED 4  /* greedy rounds hit the round cap (NaN utilities / self-loops) */
#define DG_ERR_NO_DEVICE 5      /* no CUDA device: there is no CPU fallback */

#define DG_MEM_HOST 0
#define DG_MEM_DEVICE 1

/* activations (gcn/layers.py:216 applies self.act; gcn/models.py:541-573 picks it per layer) */
#define DG_ACT_IDENTITY 0
#define DG_ACT_LEAKY_RELU 1     /* tf.nn.leaky_relu, alpha passed to dg_model_create (0.2 in TF) */
#define DG_ACT_RELU 2

/* utility modes (FLAGS.predict, runtime_config.py:29; used at mwis_dqn_call.py:230-235) */
#define DG_PREDICT_MWIS 0       /* utility = fp64(score) * weight */
#define DG_PREDICT_MIS 1        /* utility = fp64(score) */

/* output heads */
#define DG_HEAD_LINEAR 0        /* GCN_DQN / GCN2_DQN: outputs_softmax = outputs (gcn/models.py:524) */
#define DG_HEAD_PAIR_SOFTMAX 1  /* GCN_DEEP_DIVER: softmax over each (2d, 2d+1) pair (gcn/models.py:399-401) */

typedef struct dg_context dg_context;
typedef struct dg_model dg_model;
typedef struct dg_batch dg_batch;
typedef struct dg_part dg_part;

/* ---- library ------------------------------------------------------------------------------- */
DG_API int dg_version(void);
/* Message of the last failing call on this thread ("" if none).  Replaces the reference's
 * exceptions / asserts (gcn/layers.py:72-74,182). */
DG_API const char *dg_last_error(void);
/* Number of visible CUDA devices (0 = none; every other call then fails with DG_ERR_NO_DEVICE). */
DG_API int dg_device_count(void);

/* ---- context: replaces the module-level tf.compat.v1.Session (mwis_dqn_call.py:336-344) ------ */
/* `stream` is a cudaStream_t to enqueue on (e.g. torch's current stream), or NULL for a private
 * non-blocking stream. */
DG_API int dg_context_create(int device, void *stream, dg_context **out);
DG_API void dg_context_destroy(dg_context *ctx);
DG_API int dg_context_synchronize(dg_context *ctx);
/* The library's options (DG_DISABLE_TC, DG_DISABLE_FUSED, DG_DISABLE_STAGED, DG_FUSED_MMA and the DG_*_TIMING / DG_*_DUMP
 * measurement aids) are read from the environment when a context is created - never on the solve path.  Call this after
 * changing them for a live context. */
DG_API int dg_context_reload_env(dg_context *ctx);
/* Pinned host memory for staging HOST-space calls at full PCIe speed. */
DG_API void *dg_host_alloc(uint64_t bytes);
DG_API void dg_host_free(void *p);
/* Number of kernel launches this context has enqueued so far (bench.py's gpu_launches). */
DG_API uint64_t dg_context_launch_count(const dg_context *ctx);
/* Name of the kernel that carried the most recent profiled launch on this context (the solve kernel of
 * dg_solve* / dg_gcn_forward: "tc_solve_kernel", "fused_solve_kernel" or "gc_layer_kernel"); "" before
 * the first launch.  The string is static. */
DG_API const char *dg_context_last_kernel(const dg_context *ctx);

/* CUDA-event stopwatch on the context's stream: start records an event, stop records a second one,
 * waits for it and returns the elapsed device time between the two in milliseconds. */
/* Work enqueued on `ctx` after this call starts only when everything enqueued on `other` so far has completed (an
 * event on other's stream, no host synchronisation).  For callers that drive two contexts of one device in turn. */
DG_API int dg_context_wait(dg_context *ctx, dg_context *other);
DG_API int dg_timer_start(dg_context *ctx);
DG_API int dg_timer_stop(dg_context *ctx, double *elapsed_ms);

/* Device timing of the dominant kernel (the fused GraphConvolution layer kernel): when enabled, every
 * launch is bracketed by CUDA events on the context's stream.  dg_profile_collect synchronises and
 * returns the summed duration, the number of launches and their algorithmic bytes (DESIGN.md
 * "B_layer") since the previous collect.  At most 8192 launches are kept between collects. */
DG_API int dg_profile_enable(dg_context *ctx, int on);
DG_API int dg_profile_collect(dg_context *ctx, double *total_ms, uint64_t *launches, double *algorithmic_bytes);

/* ---- model: replaces GCN_DQN / GCN_DEEP_DIVER / GCN2_DQN._build (gcn/models.py:536-573, --------
 * 411-434, 670-708) + Saver.restore (mwis_dqn_call.py:188-192).  `weights` holds
 * n_layers*n_supports host pointers, weights[l*n_supports + k] = "weights_k" of layer l, row-major
 * [c_in[l], c_out[l]] (gcn/layers.py:175-183); `bias` is NULL or n_layers host pointers with NULL
 * for bias-free layers.  Widths up to 64 and n_supports == 2 ("cheb1": T_0 = I, T_1 = L) are
 * supported. */
DG_API int dg_model_create(dg_context *ctx, int n_layers, int n_supports, const int32_t *c_in,
                    const int32_t *c_out, const float *const *weights, const float *const *bias,
                    const int32_t *act, float leaky_alpha, int head, dg_model **out);
DG_API void dg_model_destroy(dg_model *model);
DG_API int dg_model_out_width(const dg_model *model);

/* ---- batch: replaces makestate's support construction (mwis_dqn_call.py:129-138, --------------
 * gcn/utils.py:120-128,258-274): only the pattern is stored; L = I - D^-1/2 A D^-1/2 is applied
 * from the degree vector and never materialised. */
DG_API int dg_batch_create(dg_context *ctx, int32_t n_graphs, int32_t n_nodes, int32_t nnz,
                    const int32_t *graph_ptr, const int32_t *row_ptr, const int32_t *col_idx, int mem,
                    dg_batch **out);
DG_API void dg_batch_destroy(dg_batch *batch);
/* Vertex mask: keep[v] == 0 removes v (and its edges) from its graph, as the zero-weight removal of
 * mwis_dqn_call.py:202-207 does; degrees are recomputed on the kept sub-graph.  NULL keeps all. */
DG_API int dg_batch_set_keep(dg_batch *batch, const uint8_t *keep, int mem);
/* keep[v] = (wts[v] != 0), computed on the device (mwis_dqn_call.py:203-204). */
DG_API int dg_batch_set_keep_from_weights(dg_batch *batch, const double *wts, int mem);
/* Per-vertex input feature value x0[v] shared by all feature columns (what makestate produces:
 * mwis_dqn_call.py:131-135, mwis_gdpg_call.py:84-91).  NULL = 1/feature_size on kept vertices. */
DG_API int dg_batch_set_x0(dg_batch *batch, const float *x0, int mem);

/* ---- operators ------------------------------------------------------------------------------ */
/* y = L.z with L = I - D^-1/2 A D^-1/2 of the batch (values never stored: applied from dinv), z and y [n_nodes, width]
 * row-major float32.  Replaces tf.sparse_tensor_dense_matmul(support[1], pre_sup) of GraphConvolution._call
 * (gcn/layers.py:206) as a stand-alone operator.  Batches of small graphs with width 32 stage whole graphs in shared
 * memory (gs_spmm_kernel); everything else runs one warp per row. */
DG_API int dg_spmm_laplacian(dg_context *ctx, dg_batch *batch, int32_t width, const float *z, float *y, int mem);
/* One GraphConvolution layer on dense inputs: y = act(x.W_0 + L.(x.W_1) + b).
 * Replaces GraphConvolution.__call__ (gcn/layers.py:189-216).  x is [n_nodes, c_in] row-major,
 * y is [n_nodes, c_out]; W_k are host pointers ([c_in, c_out]); bias may be NULL. */
DG_API int dg_graph_convolution(dg_context *ctx, dg_batch *batch, int32_t c_in, int32_t c_out,
                         const float *w0, const float *w1, const float *bias, int act,
                         float leaky_alpha, const float *x, float *y, int mem);

/* The whole stack.  Replaces DQNAgent.predict's sess.run of model.outputs_softmax
 * (mwis_dqn_call.py:140-143).  out is [n_nodes, dg_model_out_width] row-major fp32; rows of removed
 * vertices are 0. */
DG_API int dg_gcn_forward(dg_context *ctx, const dg_model *model, dg_batch *batch, float *out, int mem);

/* util[v] = fp64(score[v*score_stride]) * wts[v]  (DG_PREDICT_MWIS) or fp64(score) (DG_PREDICT_MIS).
 * Replaces mwis_dqn_call.py:230-235. */
DG_API int dg_utility(dg_context *ctx, const dg_batch *batch, const float *score, int32_t score_stride,
               const double *wts, int predict, double *util, int mem);

/* Local greedy MWIS on every graph of the batch.  Replaces heuristics.local_greedy_search and its
 * _count/_stats/_overhead/_nstep variants (heuristics.py:77-305).
 *   util     fp64 utilities (n_nodes), compared exactly as IEEE doubles
 *   nstep    < 0: run to completion; otherwise stop after nstep rounds (heuristics.py:279)
 *   member   out, n_nodes bytes, 1 = in the set
 *   nb_is    out or NULL, n_nodes bytes, the reference's nb_is set (heuristics.py:305)
 *   steps    out or NULL, n_graphs int32, rounds executed per graph (heuristics.py:160)
 *   p2p,bst  out or NULL, n_graphs int64, message counts (heuristics.py:185,179,208)
 *   oh_vec   out or NULL, n_nodes fp64, per-vertex overhead (heuristics.py:238,249)
 * Vertices removed by dg_batch_set_keep start outside `remain`. */
DG_API int dg_lgs(dg_context *ctx, const dg_batch *batch, const double *util, int32_t nstep, uint8_t *member,
           uint8_t *nb_is, int32_t *steps, int64_t *p2p, int64_t *bst, double *oh_vec, int mem);

/* Threshold distributed greedy on every graph of the batch.  Replaces heuristics.dist_greedy_search
 * (heuristics.py:38-74; call sites wireless_dqn_test_mc.py:252, wireless_dqn_test.py:242 with epsilon 0.1).
 *   wts      fp64 weights (n_nodes); a vertex is a candidate of a round when it has no remaining neighbour
 *            or wts[v] >= max(remaining neighbours' wts) / (1 + epsilon / 3)
 *   member   out, n_nodes bytes;  steps  out or NULL, n_graphs int32, rounds executed
 * Each round's candidates are scanned in ascending vertex id (the reference scans them in Python-set
 * iteration order; identical whenever no two candidates of a round are adjacent).  One CTA per graph with the
 * round's four bitmaps in shared memory: graphs of up to ~460 000 vertices (DG_ERR_UNSUPPORTED beyond).  Vertices
 * removed by dg_batch_set_keep start outside `remain`. */
DG_API int dg_dist_greedy(dg_context *ctx, const dg_batch *batch, const double *wts, double epsilon, uint8_t *member,
                   int32_t *steps, int mem);

/* total[g] = sum of wts over the members of graph g (mwis_dqn_call.py:241). */
DG_API int dg_member_weight(dg_context *ctx, const dg_batch *batch, const uint8_t *member, const double *wts,
                     double *total, int mem);

/* Fused path: (optional) zero-weight removal -> GCN -> utility -> LGS -> totals.  Replaces
 * DQNAgent.solve_mwis(adj_0, wts_0, train=False) (mwis_dqn_call.py:198-261) for a whole batch.
 * score/util/total/steps may be NULL. */
DG_API int dg_solve(dg_context *ctx, const dg_model *model, dg_batch *batch, const double *wts, int predict,
             int remove_zero_weight, uint8_t *member, float *score, double *util, double *total,
             int32_t *steps, int mem);

/* GCN embedded into the greedy iteration.  Replaces MWISSolver.solve_mwis_dit (mwis_gdpg_call.py:278-318) for a
 * whole batch: while a graph has residual vertices whose weights sum to something positive, re-score the residual
 * graph (degrees renormalised on it), run ONE greedy round (heuristics.local_greedy_search_nstep, nstep = 1), add
 * the joined vertices to the set and drop them and their neighbours (nb_is).  No zero-weight removal (generation-2
 * semantics); a mask set with dg_batch_set_keep is the initial residual graph.  member: n_nodes bytes;
 * total (may be NULL): sum of wts over the set per graph; steps (may be NULL): iterations per graph.
 * Small graphs run the whole loop inside the graph-resident kernel (one launch). */
DG_API int dg_solve_dit(dg_context *ctx, const dg_model *model, dg_batch *batch, const double *wts, int predict,
                 uint8_t *member, double *total, int32_t *steps, int mem);

/* One-shot host form of dg_solve: host CSR in, host membership out; the batch lives in the
 * context's reusable device buffers.  This is what bench.py's e2e number calls. */
DG_API int dg_solve_host(dg_context *ctx, const dg_model *model, int32_t n_graphs, int32_t n_nodes, int32_t nnz,
                  const int32_t *graph_ptr, const int32_t *row_ptr, const int32_t *col_idx,
                  const double *wts, int predict, int remove_zero_weight, uint8_t *member,
                  double *total);
/* Enqueue-only form of dg_solve_host for streams of batches: the H2D copies, the kernels and the D2H copies are
 * queued on the context's stream and the call returns without waiting; dg_context_synchronize (which also reports a
 * DG_ERR_NOT_CONVERGED raised by the kernels) completes it.  The caller's arrays must stay valid until then and
 * should be pinned (dg_host_alloc) - pageable memory makes the copies synchronous.  Two contexts used alternately
 * overlap one batch's copies with the other's kernels (distgcn_b200.engine.HostPipeline; bench.py's e2e number). */
DG_API int dg_solve_host_async(dg_context *ctx, const dg_model *model, int32_t n_graphs, int32_t n_nodes, int32_t nnz,
                        const int32_t *graph_ptr, const int32_t *row_ptr, const int32_t *col_idx,
                        const double *wts, int predict, int remove_zero_weight, uint8_t *member,
                        double *total);

/* The same with the compact host format: col_local[e] = column id LOCAL to the graph (what every per-graph scipy
 * matrix of the reference holds, mwis_dqn_call.py:198), 16 bits - half the host->device bytes of the packed int32
 * form, which is what bounds the streaming path once the kernels are fast (PCIe).  Expanded to batch-global ids on
 * the device.  Graphs of at most 65536 vertices.  wait != 0: return after the results have landed (dg_solve_host),
 * wait == 0: enqueue only (dg_solve_host_async). */
DG_API int dg_solve_host_compact(dg_context *ctx, const dg_model *model, int32_t n_graphs, int32_t n_nodes, int32_t nnz,
                          const int32_t *graph_ptr, const int32_t *row_ptr, const uint16_t *col_local, const double *wts,
                          int predict, int remove_zero_weight, uint8_t *member, double *total, int wait);

/* ---- native ingest: the reference's per-graph inputs, as they are ----------------------------------------------
 * The reference passes every graph as its own scipy sparse matrix (mwis_dqn_call.py:198; the .mat files written by
 * Data_Generation.py:214-219 load as float64 CSC with int32 indptr / indices) and converts it on every call through
 * networkx (mwis_dqn_call.py:202-207), then maps the solution back (:240-241).  These entry points take TABLES OF
 * POINTERS to the per-graph arrays - indptr[g] (n_rows[g] + 1 entries starting at 0), indices[g] (column ids local to
 * the graph; the adjacency is symmetric, so CSC and CSR coincide), optionally data[g] (float64 stored values: a stored
 * zero is not an edge, np.nonzero(adj[v]) at heuristics.py:94; data or any data[g] may be NULL = all ones) - and pack
 * them with a pool of host threads: no per-graph work in Python.  Malformed input (decreasing indptr, a column id
 * outside [0, n_rows[g])) is DG_ERR_INVALID. */

/* Sizes of the packed batch (to allocate the outputs of dg_pack_graphs_host). */
DG_API int dg_pack_graphs_sizes(int32_t n_graphs, const int32_t *const *indptr, const double *const *data,
                         const int32_t *n_rows, int64_t *n_nodes, int64_t *nnz, int32_t *max_rows);
/* Pack into caller arrays: graph_ptr [n_graphs + 1], row_ptr [n_nodes + 1] and EXACTLY ONE of col_idx (int32,
 * batch-global ids: the packed form of dg_batch_create / dg_solve_host) and col_local16 (uint16, graph-local ids: the
 * compact form of dg_solve_host_compact; graphs of at most 65536 vertices).  n_threads <= 0: all pool threads.
 * Host-only: needs no GPU. */
DG_API int dg_pack_graphs_host(int32_t n_graphs, const int32_t *const *indptr, const int32_t *const *indices,
                        const double *const *data, const int32_t *n_rows, int32_t *graph_ptr, int32_t *row_ptr,
                        int32_t *col_idx, uint16_t *col_local16, int32_t n_threads);
/* The same in the UPPER format: only the entries with column > row (an adjacency matrix is symmetric with a zero
 * diagonal - self-loops make the reference's greedy search loop for ever - so half of it says everything): graph_ptr
 * [n_graphs + 1], row_ptr_upper [n_nodes + 1], col_local_upper (uint16, graph-local, sum(nnz) / 2 entries).  A graph
 * whose stored non-zeros are not exactly half above the diagonal is DG_ERR_INVALID. */
DG_API int dg_pack_graphs_upper_host(int32_t n_graphs, const int32_t *const *indptr, const int32_t *const *indices,
                              const double *const *data, const int32_t *n_rows, int32_t *graph_ptr, int32_t *row_ptr_upper,
                              uint16_t *col_local_upper, int32_t n_threads);
/* DQNAgent.solve_mwis for a list of graphs in one call (mwis_dqn_call.py:198-261 per graph): pack into the context's
 * pinned staging, copy, solve (dg_solve semantics: zero-weight removal, GCN, utility, local greedy search), copy the
 * results back.  Weights either per graph (wts_per_graph[g]: n_rows[g] doubles) or packed (wts_packed: one double per
 * vertex in graph order); exactly one of the two.  member: sum(n_rows) bytes in ORIGINAL vertex ids of each graph;
 * total (may be NULL): n_graphs doubles.  wait == 0: enqueue only - the per-graph arrays may be released on return
 * (they have been packed), member / total / wts_packed must stay valid until dg_context_synchronize. */
DG_API int dg_solve_graphs_host(dg_context *ctx, const dg_model *model, int32_t n_graphs, const int32_t *const *indptr,
                         const int32_t *const *indices, const double *const *data, const int32_t *n_rows,
                         const double *const *wts_per_graph, const double *wts_packed, int predict,
                         int remove_zero_weight, uint8_t *member, double *total, int wait);

/* ---- multi-channel wireless scheduling: the queue bookkeeping of the slot loop, resident on the device ----------
 * wireless_dqn_test_mc.py:225-366 advances one network one time slot at a time in numpy around one solver call.  A
 * dg_wireless holds the state of MANY (network, load) instances advanced together - queues, arrivals [T][n_links]
 * (float64), link rates [T][n_links][n_ch] (int32), the queue history - on the device, and its calls only ENQUEUE
 * kernels on the context's stream; with the solvers called in DG_MEM_DEVICE mode on the buffers of
 * dg_wireless_buffers a whole sweep runs without a host synchronisation until dg_wireless_read_history.
 * Joint-graph algorithms (Greedy, Greedy-Th, DGCN-LGS, DGCN-LGS-it; vertex link_v0[l] + k * link_nf[l] = link l on
 * channel k, wireless_rollout_test_flood.py:98-133), per slot t = 1 .. T-1:
 *     begin_slot(t): q += arrivals[t] (:227);  joint_weights(t): w[v] = q[link] * rate[t][link][channel] (:228-240);
 *     <solver on w -> member>;  joint_serve(t): capacity[link] = rate of its scheduled vertex (:358-363);
 *     end_slot(t): q -= min(q, capacity), history[t] = q (:364-365).
 * Sequential variants (LGS-Seq, DGCN-LGS-Seq; vertex = link, one graph batch per channel), per slot:
 *     begin_slot(t); for every channel: seq_weights(t, ic): w = queue estimate * rate[:, ic] (:298);  <solver>;
 *     seq_serve(t, ic): capacity of the scheduled links, estimate -= min(estimate, rate) on them (:306-309);  end_slot(t). */
typedef struct dg_wireless dg_wireless;
DG_API int dg_wireless_create(dg_context *ctx, int32_t n_links, int32_t n_ch, int32_t n_slots, const double *arrivals,
                       const int32_t *rates, const int32_t *link_v0, const int32_t *link_nf, int32_t n_vertices,
                       dg_wireless **out);
DG_API void dg_wireless_destroy(dg_wireless *s);
/* device buffers: w (weights for the solver, max(n_vertices, n_links) doubles), member (the solver writes its answer
 * here), q (current queue lengths, n_links doubles); any of the outputs may be NULL */
DG_API int dg_wireless_buffers(dg_wireless *s, double **w, uint8_t **member, double **q);
DG_API int dg_wireless_begin_slot(dg_wireless *s, int32_t t);
DG_API int dg_wireless_joint_weights(dg_wireless *s, int32_t t);
DG_API int dg_wireless_joint_serve(dg_wireless *s, int32_t t);
DG_API int dg_wireless_seq_weights(dg_wireless *s, int32_t t, int32_t channel);
DG_API int dg_wireless_seq_serve(dg_wireless *s, int32_t t, int32_t channel);
DG_API int dg_wireless_end_slot(dg_wireless *s, int32_t t);
/* The slot loop itself for slots t_first .. t_first + n_slots - 1 (wireless_dqn_test_mc.py:225-366), enqueued natively: one call
 * per sweep instead of five per slot.  `scheduler`: DG_WL_LGS = heuristics.local_greedy_search ("Greedy" / "LGS-Seq"),
 * DG_WL_DIST_GREEDY = dist_greedy_search(epsilon) ("Greedy-Th"), DG_WL_SOLVE = DQNAgent.solve_mwis ("DGCN-LGS" /
 * "DGCN-LGS-Seq"), DG_WL_SOLVE_DIT = MWISSolver.solve_mwis_dit ("DGCN-LGS-it").  `sequential` != 0: per channel, one batch
 * per channel (vertex = link); else one batch of joint graphs. */
#define DG_WL_LGS 0
#define DG_WL_DIST_GREEDY 1
#define DG_WL_SOLVE 2
#define DG_WL_SOLVE_DIT 3
DG_API int dg_wireless_run(dg_wireless *s, const dg_model *model, dg_batch *const *batches, int32_t n_batches, int32_t scheduler,
                           int32_t sequential, int predict, int remove_zero_weight, double epsilon, int32_t t_first,
                           int32_t n_slots);
/* synchronise and copy the queue history [n_slots][n_links] (row t = queues after slot t; row 0 zeros) to the host */
DG_API int dg_wireless_read_history(dg_wireless *s, double *history);

/* dg_solve_host_compact with the UPPER format (column > row only; see dg_pack_graphs_upper_host): a third of the
 * host->device bytes of the packed int32 form.  The tensor-core kernel builds its dense adjacency from the half lists
 * directly; other paths expand them to the full CSR on the device first (deterministically: rows in ascending column
 * order).  Graphs of at most 8192 vertices. */
DG_API int dg_solve_host_upper(dg_context *ctx, const dg_model *model, int32_t n_graphs, int32_t n_nodes, int32_t nnz_upper,
                        const int32_t *graph_ptr, const int32_t *row_ptr_upper, const uint16_t *col_local_upper,
                        const double *wts, int predict, int remove_zero_weight, uint8_t *member, double *total, int wait);

/* ---- one giant graph, row-partitioned over several GPUs (no reference counterpart: the
 * reference handles one 100-300 vertex graph per call) ------------------------------------------------
 * A dg_part is one rank's slice: rows row0 .. row0+n_local-1 of a graph with n_global vertices (row0 and
 * n_local multiples of 32), as a local CSR (row_ptr starts at 0) whose column ids are GLOBAL.  Every array
 * argument below is a DEVICE pointer to a GLOBAL-sized array indexed by global vertex id (float2 arrays as
 * 2*n_global floats, bitmaps as n_global/32 words).  A call may read any vertex's entry and writes only
 * this part's rows; between calls the caller makes the written rows visible to the other ranks
 * (all-gather over NVLink, e.g. torch.distributed / NCCL - see distgcn_b200/shard.py).  pair2 is PLANAR:
 * q = pair2[0 .. n_global), zs = pair2[n_global .. 2 n_global); only the zs plane is read from other ranks.
 * The sequence for a model without hidden layers (e.g. c64 l2) exchanges only scalars: y, zs, utilities,
 * bitmap words. */
DG_API int dg_part_create(dg_context *ctx, int32_t n_global, int32_t row0, int32_t n_local, int32_t nnz_local,
                          const int32_t *row_ptr_local, const int32_t *col_idx_global, int mem, dg_part **out);
DG_API void dg_part_destroy(dg_part *part);
/* dinv rows (degrees over kept neighbours; keep must be complete) and y = dinv * x0 rows */
DG_API int dg_part_prepare(dg_part *part, int32_t feature_size, const uint8_t *keep, const float *x0, float *dinv,
                           float *y);
/* first layer, rank-1: pair rows = (x0, s = (L x0)); reads y of the neighbours */
DG_API int dg_part_first(dg_part *part, int32_t feature_size, const float *dinv, const float *y, const uint8_t *keep,
                         const float *x0, float *pair);
/* two-layer models: pair2 rows = (q, zs) of the one-column last layer, from this part's pair rows */
DG_API int dg_part_project(dg_part *part, const dg_model *model, const float *dinv, const float *pair, float *pair2);
/* hidden layer `layer` (1 .. n_layers-2): hout rows [n_global, padded width]; reads neighbours' pair (layer 1)
 * or hin rows (later layers) and dinv */
DG_API int dg_part_layer(dg_part *part, const dg_model *model, int32_t layer, const float *dinv, const float *pair,
                         const float *hin, float *hout);
/* models with hidden layers: pair2 rows = (q, zs) from this part's rows of the last hidden layer's output */
DG_API int dg_part_tail(dg_part *part, const dg_model *model, const float *dinv, const float *hin, float *pair2);
/* last layer + utility: score / util rows; reads neighbours' pair2 */
DG_API int dg_part_last(dg_part *part, const dg_model *model, const float *dinv, const float *pair2,
                        const uint8_t *keep, const double *wts, int predict, float *score, double *util);
/* padded width of the features hidden layer `layer` writes (32 or 64) */
DG_API int dg_model_padded_width(const dg_model *model, int32_t layer);
/* greedy rounds: init writes this part's remain words and member rows and adds its remaining count to
 * *count; decide reads all remain words / utilities and writes this part's joined words and member rows;
 * remove reads all joined words, rewrites this part's remain words and adds its remaining count */
DG_API int dg_part_lgs_init(dg_part *part, const uint8_t *keep, uint32_t *remain, uint8_t *member, int64_t *count);
DG_API int dg_part_lgs_decide(dg_part *part, const double *util, const uint32_t *remain, uint32_t *joined,
                              uint8_t *member);
DG_API int dg_part_lgs_remove(dg_part *part, const uint32_t *joined, uint32_t *remain, int64_t *count);
/* All greedy rounds after dg_part_lgs_init + dg_part_barrier(count), driven natively (peer arenas required): `check_every`
 * rounds are enqueued back to back before the ranks' remaining counts are read once (rounds past the end change nothing);
 * *rounds_out = the rounds that started with a vertex left, i.e. what the reference's loop counts (heuristics.py:92-116).
 * No reference counterpart (the reference never partitions a graph). */
DG_API int dg_part_lgs_run(dg_part *part, const double *util, uint32_t *remain, uint32_t *joined, uint8_t *member,
                           int64_t *count, int32_t check_every, int32_t *rounds_out);

/* ---- exchange fused into the kernels: peer arenas over NVLink / NVSwitch (one process per GPU) ------------
 * Instead of all-gathering after every dg_part_* call, the ranks can share one "arena" each: a device
 * allocation of identical size and layout on every rank, exported with CUDA IPC and mapped by the others.
 * After dg_part_set_peers, every dg_part_* call that writes a quantity other ranks read next (keep, dinv, y,
 * (x0,s) pairs, hidden rows, zs, utilities, bitmap words) stores it to its own arena AND, at the same offset,
 * to every peer's arena from the kernel's epilogue, provided the output pointer lies inside the own arena.
 * dg_part_barrier then orders the ranks on their streams (flag exchange through the arenas, no host
 * round trip); with `count` it also publishes this rank's 8-byte count into slot [rank] of every arena's
 * count block, so that after the barrier each rank holds all ranks' counts (greedy-round termination). */
#define DG_PEER_HANDLE_BYTES 64
/* cudaMalloc + zero `bytes` on the context's device and export it: handle_out receives DG_PEER_HANDLE_BYTES
 * bytes to send to the other ranks (any byte transport, e.g. torch.distributed.all_gather_object). */
DG_API int dg_peer_alloc(dg_context *ctx, uint64_t bytes, void **dev_ptr, uint8_t *handle_out);
/* map another rank's arena from its handle / unmap it / free an own arena */
DG_API int dg_peer_open(dg_context *ctx, const uint8_t *handle, void **peer_ptr);
DG_API int dg_peer_close(dg_context *ctx, void *peer_ptr);
DG_API int dg_peer_free(dg_context *ctx, void *dev_ptr);
/* arena_bases[r] = rank r's arena as mapped in this process (arena_bases[rank] = the own allocation);
 * flags_offset / counts_offset: 4*world and 8*world bytes inside the arena reserved for the barrier
 * (zero-initialised, never touched by the caller).  world <= 8. */
DG_API int dg_part_set_peers(dg_part *part, int32_t world, int32_t rank, void *const *arena_bases,
                             uint64_t arena_bytes, uint64_t flags_offset, uint64_t counts_offset);
/* keep rows of this part: v < n_real (rows past the real vertex count are padding) and, with
 * remove_zero_weight, wts[v] != 0 (mwis_dqn_call.py:203-204); wts is global-indexed, own rows valid */
DG_API int dg_part_keep(dg_part *part, const double *wts, int remove_zero_weight, int32_t n_real, uint8_t *keep);
/* cross-rank barrier on the stream; count may be NULL.  A rank that does not arrive within ~4 s makes the
 * next synchronising call fail with DG_ERR_CUDA instead of hanging the GPUs. */
DG_API int dg_part_barrier(dg_part *part, const int64_t *count);

#ifdef __cplusplus
}
#endif

#endif /* DISTGCN_B200_H */
