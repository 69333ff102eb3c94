// Internal declarations shared by the translation units of libdistgcn_b200.so.
// The public surface is include/distgcn_b200.h; nothing here is exported.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <string>
#include <vector>

#include "../../include/distgcn_b200.h"

namespace dg {

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) once per DEVICE and call site (the attribute belongs to the function on a
// device: a process that drives two GPUs must set it on both).  `done` is a static bit mask owned by the call site.
template <typename Kernel>
inline cudaError_t smem_attr_once(Kernel kern, int device, int bytes, std::atomic<unsigned long long> *done) {
    const unsigned long long bit = 1ull << (device & 63);
    if (done->load(std::memory_order_acquire) & bit) return cudaSuccess;
    const cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (e == cudaSuccess) done->fetch_or(bit, std::memory_order_release);
    return e;
}

void set_error(const char *fmt, ...);
void clear_error();

#define DG_CUDA_CHECK(expr)                                                                    \
    do {                                                                                       \
        cudaError_t _e = (expr);                                                               \
        if (_e != cudaSuccess) {                                                               \
            dg::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__,    \
                          __LINE__);                                                           \
            return DG_ERR_CUDA;                                                                \
        }                                                                                      \
    } while (0)

#define DG_REQUIRE(cond, code, ...)      \
    do {                                 \
        if (!(cond)) {                   \
            dg::set_error(__VA_ARGS__);  \
            return (code);               \
        }                                \
    } while (0)

#define DG_TRY(expr)                 \
    do {                             \
        int _s = (expr);             \
        if (_s != DG_OK) return _s;  \
    } while (0)

constexpr int kMaxWidth = 64;      // widest layer the fused kernels cover
constexpr int kMaxLayers = 256;
constexpr int kLgsCtaMaxNodes = 8192;  // graphs up to this size run in the one-CTA-per-graph LGS kernel
constexpr int kLgsRoundCap = 1 << 20;  // NaN utilities / self-loops never converge in the reference

// Symmetric peer arenas of a row-partitioned solve (one process per GPU, NVLink / NVSwitch peer memory).
// Every rank allocates one arena of the same size and layout and maps the other ranks' arenas through CUDA
// IPC; base[r] is rank r's arena as seen from THIS process.  A kernel that produces a quantity other ranks
// read next stores it with peer_store(): to its own arena and, at the same offset, to every peer's - the
// exchange is fused into the producing kernel's epilogue instead of a separate all-gather.  world == 1 (or
// a pointer outside the arena) degenerates to a plain store.
constexpr int kMaxPeers = 8;
struct PeerMap {
    int world = 1;
    int rank = 0;
    unsigned long long bytes = 0;
    char *base[kMaxPeers] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
};

#ifdef __CUDACC__
template <typename T>
__device__ __forceinline__ void peer_store(const PeerMap &pm, T *own, T v) {
    *own = v;
    if (pm.world > 1) {
        const unsigned long long off = (unsigned long long)(reinterpret_cast<char *>(own) - pm.base[pm.rank]);
        if (off < pm.bytes) {
#pragma unroll
            for (int r = 0; r < kMaxPeers; ++r)
                if (r < pm.world && r != pm.rank) *reinterpret_cast<T *>(pm.base[r] + off) = v;
        }
    }
}
#endif

// grow-only device buffer
struct Buffer {
    void *ptr = nullptr;
    size_t cap = 0;
};

}  // namespace dg

// Options read from the environment ONCE per context (dg_context_create) and on dg_context_reload_env: the solve path
// never calls getenv.  DG_DISABLE_TC / DG_DISABLE_FUSED / DG_DISABLE_STAGED force the next kernel family down the
// list (tensor-core -> CUDA-core graph-resident -> per-layer; graph-staged -> warp-per-row), DG_FUSED_MMA selects the
// mma.sync projection of the fused kernel, the rest are measurement aids.
struct dg_env {
    bool disable_tc = false, disable_fused = false, disable_staged = false, fused_mma = false;
    bool fused_timing = false, tc_debug = false, ingest_upper = false, ingest_timing = false;
    int tile_rows = 0, tc_tiles = 0;
    std::string fused_tile_dump, tc_tile_dump;
};

struct dg_context {
    dg_env env;
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    int sm_count = 148;
    int max_smem_optin = 0;
    uint64_t launches = 0;
    std::vector<dg::Buffer> slots;  // named scratch, see dg::Slot
    cudaEvent_t order_ev = nullptr;  // dg_context_wait: marks this context's stream for another context to wait on
    bool tc_attr_set = false;       // tc_solve_kernel's shared-memory attribute has been set on this context's device
    int *h_flag = nullptr;          // pinned, 4 ints: [0] LGS round read-back, [2] copy of *d_status
    int *d_status = nullptr;        // device, sticky error status written by kernels
    dg_batch *host_batch = nullptr; // reusable batch of dg_solve_host
    // optional device timing of the dominant kernel (gc_layer_kernel), see dg_profile_*
    cudaEvent_t timer_ev[2] = {nullptr, nullptr};
    bool prof_on = false;
    std::vector<cudaEvent_t> prof_events;  // pairs: [2k] start, [2k+1] stop
    size_t prof_used = 0;                  // events recorded since the last collect
    double prof_bytes = 0.0;               // algorithmic bytes of the recorded launches
    const char *last_kernel = "";          // dominant kernel of the most recent solve (dg_context_last_kernel)
    void *ingest_staging = nullptr;        // pinned staging of dg_solve_graphs_host (dg_ingest.cu)
};

namespace dg {

enum Slot : int {
    kSlotFeatA = 0,   // layer ping
    kSlotFeatB,       // layer pong
    kSlotPair,        // per-node float2 (x0, s) of the rank-1 first layer / (q, zs) of the scalar tail
    kSlotPair2,
    kSlotY,           // dinv * x0
    kSlotScore,
    kSlotUtil,
    kSlotWts,
    kSlotMember,
    kSlotNbis,
    kSlotSteps,
    kSlotP2p,
    kSlotBst,
    kSlotOh,
    kSlotTotal,
    kSlotLgsWords,    // global-path bitmaps
    kSlotLgsCount,
    kSlotStageIn0,    // staging of HOST-space arguments
    kSlotStageIn1,
    kSlotStageIn2,
    kSlotStageOut0,
    kSlotStageOut1,
    kSlotLayerW,      // dg_graph_convolution's transient weights
    kSlotHostGraphPtr,
    kSlotHostRowPtr,
    kSlotHostColIdx,
    kSlotPartial,     // per-slice partial sums of member_weight
    kSlotSolveKeep,   // dg_solve's zero-weight keep mask on the per-layer path (the caller's mask is left alone)
    kSlotCount
};

int scratch(dg_context *ctx, int slot, size_t bytes, void **out);

template <typename T>
inline int scratch_as(dg_context *ctx, int slot, size_t count, T **out) {
    void *p = nullptr;
    int s = scratch(ctx, slot, count * sizeof(T), &p);
    *out = static_cast<T *>(p);
    return s;
}

}  // namespace dg

struct dg_layer_dev {
    int c_in = 0, c_out = 0;
    int cpi = 0, cpo = 0;        // padded widths (32 or 64)
    int act = 0;
    bool has_bias = false;
    float *wcat = nullptr;       // [2*cpi, cpo] row-major: rows 0..cpi-1 = W_0, rows cpi.. = W_1 (zero padded)
    float *bias = nullptr;       // [cpo] (zeros when absent)
    float *colsum0 = nullptr;    // [cpo]: sum over input rows of W_0 (rank-1 first layer)
    float *colsum1 = nullptr;    // [cpo]
};

struct dg_model {
    dg_context *ctx = nullptr;
    int n_layers = 0;
    int n_supports = 2;
    float alpha = 0.2f;
    int head = 0;
    std::vector<dg_layer_dev> layers;
    // host copies of the scalar-tail vectors (last layer when c_out == 1) live on the device too
    float *tail_w0 = nullptr;    // [cpi_last] W_0[:,0] of the last layer
    float *tail_w1 = nullptr;    // [cpi_last]
    float tail_bias = 0.f;
    // operands of the graph-resident fused kernel (dg_fused.cu); fused_cp == 0 when the model is not
    // eligible (several output columns, pair-softmax head, hidden layers wider than 32)
    int fused_cp = 0;
    float *fused_first = nullptr;  // [3*cp]: colsum(W_0), colsum(W_1), bias of layer 0
    float *fused_wall = nullptr;   // hidden layers 1..L-2, each [2*cp*cp + cp]: W_0 rows, W_1 rows, bias
    float *fused_wall_mma = nullptr;  // same layers for the tensor-core path: TF32 hi [64x40], lo [64x40], bias [32]
    float *fused_tail = nullptr;   // [2*cp]: W_0[:,0], W_1[:,0] of the last layer
    int *d_acts = nullptr;         // [n_layers]
    // operands of the tensor-core solve kernel (dg_tc.cu); null when the model is not eligible
    unsigned char *tc_wall = nullptr;  // hidden layers: bf16 terms of [W_0 | W_1 r], bias, 1/r, bound factor
    float tc_tail_norm = 0.f;          // sum |W_1[:,0]| of the last layer (fixed-point bound of its scalar aggregation)
    // Scalar network: every layer has ONE output column, any number of supports [I, L, L^2, ...] (the shipped cheb2
    // checkpoints: gcn/utils.py:258-274 with max_degree = 2).  Layer l is out = act(sum_k L^k (z * w[l][k]) + b[l]) on a
    // per-vertex scalar z; scalar_w[l * n_supports + k] is W_k[0,0], or for the first layer the column sum of W_k over the
    // (constant) input features.  Host copies: the values travel as kernel arguments.
    bool scalar_net = false;
    std::vector<float> scalar_w, scalar_b;
};

struct dg_batch {
    dg_context *ctx = nullptr;
    int n_graphs = 0, n_nodes = 0, nnz = 0;
    int max_graph_nodes = 0;
    bool owns_csr = false;
    int32_t *graph_ptr = nullptr, *row_ptr = nullptr, *col_idx = nullptr;  // device
    uint16_t *col16 = nullptr;  // device copy of graph-local 16-bit column ids (compact host format), cap_nnz entries
    bool cols_pending = false;  // col16 holds the batch's columns and col_idx has not been expanded from it yet
    // "upper" host format: col16 / row_ptr_u hold only the entries with column > row (the adjacency is symmetric, so half
    // of it says everything).  The tensor-core kernel builds its dense adjacency bytes from them directly; every other
    // path first expands them into the full row_ptr / col_idx on the device (batch_ensure_cols).
    int32_t *row_ptr_u = nullptr;
    size_t cap_row_ptr_u = 0;
    bool upper_pending = false;
    bool dinv_deferred = false;  // the degrees wait for that expansion
    std::vector<int32_t> h_graph_ptr;  // host copy (small) for launch planning
    float *dinv = nullptr;     // [n_nodes] fp32(deg^-1/2) on the kept sub-graph, 0 for isolated/removed
    uint8_t *keep = nullptr;   // [n_nodes] or nullptr = all kept
    float *x0 = nullptr;       // [n_nodes] or nullptr = 1/F
    size_t cap_nodes = 0, cap_nnz = 0, cap_graphs = 0;  // capacities when reused by dg_solve_host
    std::vector<int32_t> h_graph_e;  // row_ptr at the graph boundaries (host), n_graphs + 1
    int max_graph_nnz = 0;
    // tile table of the fused kernel, cached per (cap_n, cap_nnz)
    int *tiles_dev = nullptr;
    size_t tiles_cap = 0;
    int n_tiles = 0, tiles_cap_n = 0, tiles_cap_nnz = 0, tiles_cp = 0;
    bool tiles_hidden = false;
    size_t tiles_wblob = 0;
    bool tiles_valid = false;
    int tiles_hint_n = 0, tiles_hint_graphs = 0, tiles_hint_min_n = 0, tiles_hint_cp = 0;  // the previous plan's shape
    bool tiles_hint_hidden = false;
    unsigned tiles_hint_calls = 0;
    // tile table of the tensor-core kernel (dg_tc.cu): up to 4 graph ids per tile
    int *tc_tiles_dev = nullptr;
    size_t tc_tiles_cap = 0;
    int tc_n_tiles = 0;
    bool tc_tiles_valid = false;
    int tc_plan_hidden = -1;         // hidden-layer count the cached tile plan was made for (its cost model scales with depth)
    // tile table of the graph-staged streaming layer kernel (dg_stream.cu): row ranges aligned to graph boundaries
    int *gs_tiles_dev = nullptr;
    size_t gs_tiles_cap = 0;
    int gs_n_tiles = 0, gs_grid = 0, gs_grid_spmm = 0;
    bool gs_valid = false;
    bool tc_plan_ready = false;      // tc_tiles_host already holds the plan of the current batch (made ahead, on a pool thread)
    bool meta_ready = false;         // host metadata (h_graph_ptr, h_graph_e, maxima) already describe the batch being filled
    std::vector<int> tc_tiles_host;  // host copy of the table (scheduling diagnostics)
    std::vector<uint8_t> tc_skip;    // per graph: 1 = beyond the tensor-core kernel's limits (left to the CUDA-core kernel)
    int tc_n_skipped = 0;
    bool tc_ran_partial = false;     // the tensor-core kernel has just solved the batch except the tc_skip graphs
    bool tiles_subset = false;       // the fused kernel's cached tile table covers only the tc_skip graphs
};

struct dg_part {  // one rank's row slice of a single large graph (device CSR, global column ids)
    dg_context *ctx = nullptr;
    int n_global = 0, row0 = 0, n_local = 0, nnz = 0;
    bool owns = false;
    int32_t *row_ptr = nullptr, *col_idx = nullptr;
    dg::PeerMap peers;          // world == 1 until dg_part_set_peers
    unsigned long long flags_off = 0, counts_off = 0;  // arena offsets of the barrier flags / per-rank counts
    unsigned epoch = 0;         // barrier generation
    long long *h_counts = nullptr;  // pinned: the ranks' remaining-vertex counts + the round counter (dg_part_lgs_run)
    int *d_rounds = nullptr;        // device: rounds in which some rank still had a vertex left
};

namespace dg {

struct PartView {
    int n_global, row0, n_local, nnz;
    const int *row_ptr, *col_idx;
    PeerMap pm;
};

// row-slice drivers (dg_gcn.cu / dg_lgs.cu): every per-vertex array is GLOBAL sized, a call reads any
// vertex and writes only rows row0 .. row0+n_local-1
int part_prepare(dg_context *ctx, const PartView &pv, const uint8_t *keep, const float *x0, float x0val, float *dinv,
                 float *y);
int part_scale(dg_context *ctx, const PartView &pv, const uint8_t *keep, const float *x0, float x0val,
               const float *dinv, float *y);
int part_first(dg_context *ctx, const PartView &pv, const float *dinv, const float *y, const uint8_t *keep,
               const float *x0, float x0val, float2 *pair);
int part_project(dg_context *ctx, const PartView &pv, const dg_model *m, const float *dinv, const float2 *pair,
                 float *pair2);
int part_layer(dg_context *ctx, const PartView &pv, const dg_model *m, int layer, const float *dinv,
               const float2 *pair, const float *hin, float *hout);
int part_tail(dg_context *ctx, const PartView &pv, const dg_model *m, const float *dinv, const float *hin,
              float *pair2);
int part_last(dg_context *ctx, const PartView &pv, const dg_model *m, const float *dinv, const float *pair2,
              const uint8_t *keep, const double *wts, int predict, float *score, double *util);
int dit_filter_device(dg_context *ctx, const dg_batch *b, const double *wts, uint8_t *keep, int *flag, int *any_left);
int dit_steps_device(dg_context *ctx, int n_graphs, const int *flag, int *steps);
int dit_update_device(dg_context *ctx, int n, const uint8_t *joined, const uint8_t *nb_is, uint8_t *member,
                      uint8_t *keep);
int part_keep(dg_context *ctx, const PartView &pv, const double *wts, int remove_zero_weight, int n_real,
              uint8_t *keep);
int part_barrier(dg_context *ctx, const PeerMap &pm, unsigned long long flags_off, unsigned epoch,
                 const long long *count_src, unsigned long long counts_off);
int part_lgs_init(dg_context *ctx, const PartView &pv, const uint8_t *keep, uint32_t *remain, uint8_t *member,
                  long long *cnt);
int part_lgs_decide(dg_context *ctx, const PartView &pv, const double *util, const uint32_t *remain, uint32_t *joined,
                    uint8_t *member);
int part_lgs_remove(dg_context *ctx, const PartView &pv, const uint32_t *joined, uint32_t *remain, long long *cnt);
int part_tally(dg_context *ctx, const long long *counts, int world, long long *own_count, int *rounds);

// ---- graph-resident fused kernel (dg_fused.cu) -------------------------------------------------
constexpr int kFusedMaxTileGraphs = 64;

struct FusedSmemPlan {  // byte offsets into dynamic shared memory
    size_t feat_a, feat_b, wbuf, wblob_bytes, util, mbar, dinv, sa, sb, x0s, rp, words, gstart, gcnt, gsteps, gpos,
        col16, gid, vid, slotof, total;
    int n_words;
};

struct FusedParams {
    const int *tiles;  // 8 ints per tile: v0, n, e0, nnz, g0, ng, -, -
    int n_tiles;
    int *tile_counter;
    const int *graph_ptr, *row_ptr, *col_idx;
    const double *wts;
    const uint8_t *keep_in;
    const float *x0;
    float x0val;
    int remove_zero;
    int n_layers;
    const float *first;
    int first_act;
    const float *wall;
    int use_mma;       // hidden-layer projection on the tensor cores (split TF32), else FFMA
    int wblob_bytes;   // bytes of one hidden layer's weight blob in `wall`
    const int *acts;
    const float *tail;
    float tail_bias;
    int last_act;
    float alpha;
    int predict;
    int cap_n, cap_nnz;
    uint8_t *member;
    float *score;
    double *util;
    double *total;
    int *steps;
    int *status;
    int round_cap;
    int do_lgs;        // 0: stop after the scores (dg_gcn_forward)
    int dit;           // 1: GCN embedded into the greedy iteration (mwis_gdpg_call.py:278-318): re-score the residual
                       //    graph before every single greedy round
    long long *dbg;  // optional per-CTA phase timers (16 slots per CTA), see DG_FUSED_TIMING in bench tools
};

// Runs the whole solve in the fused kernel when model and batch are eligible; *handled tells.
int fused_try_solve(dg_context *ctx, const dg_model *m, dg_batch *b, const double *d_wts, int predict,
                    int remove_zero_weight, uint8_t *member, float *score, double *util, double *total,
                    int32_t *steps, bool *handled, bool dit = false);

// true when fused_try_solve would take this model and every graph of the batch fits one of its tiles (no launch)
bool fused_fits(dg_context *ctx, const dg_model *m, const dg_batch *b);

// ---- tensor-core solve kernel (dg_tc.cu) ---------------------------------------------------------
// hidden-layer operand blobs from the layers' weights ([c_in, c_out] row-major, widths <= 32)
void tc_build_weights(int n_hidden, const float *const *w0, const float *const *w1, const float *const *bias, const int *c_in,
                      const int *c_out, std::vector<unsigned char> *blob);
// host-only: plan the tensor-core kernel's tiles for the batch described by b's host metadata, ahead of the solve
void tc_plan_ahead(dg_context *ctx, const dg_model *m, dg_batch *b);
int tc_try_solve(dg_context *ctx, const dg_model *m, dg_batch *b, const double *d_wts, int predict,
                 int remove_zero_weight, uint8_t *member, float *score, double *util, double *total, int32_t *steps,
                 bool *handled, bool dit = false);

// arguments of one fused hidden GraphConvolution layer (gc_layer_kernel in dg_gcn.cu, gs_layer_kernel in dg_stream.cu)
struct LayerArgs {
    int n;
    int nnz;
    int row0;  // row-slice form: rows row0 .. row0+n-1 of a larger graph, per-vertex arrays global
    const int *row_ptr;
    const int *col_idx;
    const float *dinv;
    const float *hin;       // [n, CPI] (dense input)
    const float2 *pair_in;  // (x0, s)   (implicit input)
    const float *in_a0, *in_a1, *in_b;  // first layer's column sums / bias, [CPI]
    int in_act;
    const float *wcat;      // [2*CPI, CPO]
    const float *bias;      // [CPO]
    int act;
    float alpha;
    float *hout;            // [n, CPO]
    const float *tail_w0, *tail_w1;  // [CPO]
    float *tail_q, *tail_zs;  // (q, zs) planes
    PeerMap pm;             // row-partitioned runs: output rows / zs are also stored to the peers' arenas
};

// ---- graph-staged streaming layer kernel (dg_stream.cu): batches of small graphs, 32-wide layers ----
int gs_try_layer(dg_context *ctx, dg_batch *b, int cpi, int cpo, bool implicit_in, bool tail, const LayerArgs &a,
                 bool *handled);

// Y = L.Z for 32-wide rows on a batch of small graphs through the graph-staged kernel
int gs_try_spmm(dg_context *ctx, dg_batch *b, int width, const float *z, float *y, bool *handled);

// ---- kernels / drivers implemented in dg_gcn.cu ---------------------------------------------
int batch_compute_dinv(dg_batch *b);
// try_resident = false: go straight to the per-layer kernels (the caller has already tried the graph-resident ones)
int gcn_forward_device(dg_context *ctx, const dg_model *m, dg_batch *b, float *out /*device*/,
                       const double *wts /*device or null*/, int predict, double *util /*device or null*/,
                       bool try_resident = true);
int graph_convolution_device(dg_context *ctx, dg_batch *b, const dg_layer_dev &L, float alpha,
                             const float *x, int ldx, float *y, int ldy);
int utility_device(dg_context *ctx, int n, const float *score, int stride, const double *wts, int predict,
                   double *util);
int spmm_laplacian_device(dg_context *ctx, dg_batch *b, int width, const float *z, float *y);
int keep_from_weights_device(dg_context *ctx, int n, const double *wts, uint8_t *keep);

// ---- implemented in dg_lgs.cu ------------------------------------------------------------------
int lgs_device(dg_context *ctx, const dg_batch *b, const double *util, int nstep, uint8_t *member,
               uint8_t *nb_is, int32_t *steps, int64_t *p2p, int64_t *bst, double *oh_vec);
int dist_greedy_device(dg_context *ctx, const dg_batch *b, const double *wts, double alpha, uint8_t *member,
                       int32_t *steps);
int member_weight_device(dg_context *ctx, const dg_batch *b, const uint8_t *member, const double *wts,
                         double *total);

inline int pad_width(int c) { return c <= 32 ? 32 : 64; }

// host-CSR solve on the context's reusable batch (dg_api.cu); `copied` (optional) is recorded on the stream once the
// H2D copies of the inputs have been enqueued, so a caller may recycle its staging as soon as it has fired
int solve_host_staged(dg_context *ctx, const dg_model *m, int32_t n_graphs, int32_t n_nodes, int32_t nnz,
                      const int32_t *graph_ptr, const int32_t *row_ptr, const int32_t *col_idx, const double *wts,
                      int predict, int remove_zero_weight, uint8_t *member, double *total, bool wait,
                      const uint16_t *col_local16, cudaEvent_t copied, bool upper = false);
void ingest_staging_free(dg_context *ctx);  // dg_ingest.cu
void env_read(dg_env *env);                 // dg_api.cu
// set the host metadata of the context's reusable batch from per-graph vertex / edge offsets (n_graphs + 1 each) before
// its arrays exist; the following solve_host_staged(..) call then skips recomputing them
int host_batch_set_meta(dg_context *ctx, int32_t n_graphs, const int64_t *v0, const int64_t *e0, dg_batch **out);

// profiling helpers (dg_api.cu)
void prof_begin(dg_context *ctx);
void prof_end(dg_context *ctx, double algorithmic_bytes);

}  // namespace dg
