// C-ABI of libdistgcn_b200.so (include/distgcn_b200.h): handles, staging of HOST-space arguments,
// argument validation.  All arithmetic lives in dg_gcn.cu / dg_lgs.cu.
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <new>

#include <chrono>

#include "dg_common.cuh"

namespace dg {

static thread_local char g_error[1024] = "";

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_error, sizeof(g_error), fmt, ap);
    va_end(ap);
}

void clear_error() { g_error[0] = '\0'; }

int scratch(dg_context *ctx, int slot, size_t bytes, void **out) {
    Buffer &b = ctx->slots[slot];
    if (bytes < 256) bytes = 256;
    if (b.cap < bytes) {
        if (b.ptr) {
            // the old buffer may still be in use by enqueued work
            DG_CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
            DG_CUDA_CHECK(cudaFree(b.ptr));
            b.ptr = nullptr;
            b.cap = 0;
        }
        size_t cap = bytes + bytes / 4;
        DG_CUDA_CHECK(cudaMalloc(&b.ptr, cap));
        b.cap = cap;
    }
    *out = b.ptr;
    return DG_OK;
}

void env_read(dg_env *env) {
    auto on = [](const char *name) { return getenv(name) != nullptr; };
    auto num = [](const char *name) {
        const char *v = getenv(name);
        return v ? atoi(v) : 0;
    };
    auto str = [](const char *name) {
        const char *v = getenv(name);
        return std::string(v ? v : "");
    };
    env->disable_tc = on("DG_DISABLE_TC");
    env->disable_fused = on("DG_DISABLE_FUSED");
    env->disable_staged = on("DG_DISABLE_STAGED");
    env->fused_mma = on("DG_FUSED_MMA");
    env->fused_timing = on("DG_FUSED_TIMING");
    env->tc_debug = on("DG_TC_DEBUG");
    env->ingest_upper = on("DG_INGEST_UPPER");
    env->ingest_timing = on("DG_INGEST_TIMING");
    env->tile_rows = num("DG_TILE_ROWS");
    env->tc_tiles = num("DG_TC_TILES");
    env->fused_tile_dump = str("DG_FUSED_TILE_DUMP");
    env->tc_tile_dump = str("DG_TC_TILE_DUMP");
}

constexpr size_t kProfMaxPairs = 8192;

void prof_begin(dg_context *ctx) {
    if (!ctx->prof_on || ctx->prof_used / 2 >= kProfMaxPairs) return;
    if (ctx->prof_events.size() < ctx->prof_used + 2) {
        cudaEvent_t a, b;
        if (cudaEventCreate(&a) != cudaSuccess || cudaEventCreate(&b) != cudaSuccess) return;
        ctx->prof_events.push_back(a);
        ctx->prof_events.push_back(b);
    }
    cudaEventRecord(ctx->prof_events[ctx->prof_used], ctx->stream);
}

void prof_end(dg_context *ctx, double algorithmic_bytes) {
    if (!ctx->prof_on || ctx->prof_used / 2 >= kProfMaxPairs || ctx->prof_events.size() < ctx->prof_used + 2) return;
    cudaEventRecord(ctx->prof_events[ctx->prof_used + 1], ctx->stream);
    ctx->prof_used += 2;
    ctx->prof_bytes += algorithmic_bytes;
}

namespace {

struct DeviceGuard {
    int prev = -1;
    explicit DeviceGuard(int dev) {
        cudaGetDevice(&prev);
        if (prev != dev) cudaSetDevice(dev);
        else prev = -1;
    }
    ~DeviceGuard() {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

template <typename T>
int stage_in(dg_context *ctx, int slot, const T *host, size_t count, T **dev) {
    DG_TRY(scratch_as(ctx, slot, count, dev));
    if (count)
        DG_CUDA_CHECK(cudaMemcpyAsync(*dev, host, count * sizeof(T), cudaMemcpyHostToDevice, ctx->stream));
    return DG_OK;
}

template <typename T>
int copy_out(dg_context *ctx, T *host, const T *dev, size_t count) {
    if (count && host)
        DG_CUDA_CHECK(cudaMemcpyAsync(host, dev, count * sizeof(T), cudaMemcpyDeviceToHost, ctx->stream));
    return DG_OK;
}

// message for a status code a kernel left in the context's sticky status word
static void set_kernel_status_error(int code) {
    if (code == DG_ERR_NOT_CONVERGED)
        set_error("greedy search hit the round cap (NaN utilities or self-loops?)");
    else if (code == DG_ERR_INVALID)
        set_error("malformed graph: a column id points outside its graph");
    else if (code == DG_ERR_CUDA)
        set_error("a device-side wait gave up: the peer barrier of a row-partitioned solve timed out (a rank missing "
                  "or its arena not mapped?)");
    else
        set_error("a kernel reported status %d", code);
}

// synchronise and surface sticky kernel-side errors
int finish(dg_context *ctx) {
    // the kernels' sticky status word travels with the synchronisation, not with every launch
    DG_CUDA_CHECK(cudaMemcpyAsync(ctx->h_flag + 2, ctx->d_status, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    DG_CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    if (ctx->h_flag[2] != 0) {
        int code = ctx->h_flag[2];
        ctx->h_flag[2] = 0;
        cudaMemsetAsync(ctx->d_status, 0, sizeof(int), ctx->stream);
        set_kernel_status_error(code);
        return code;
    }
    return DG_OK;
}

int check_ctx(const dg_context *ctx) {
    DG_REQUIRE(ctx != nullptr, DG_ERR_INVALID, "null context");
    return DG_OK;
}

int batch_alloc_aux(dg_batch *b, size_t n_nodes) {
    if (b->cap_nodes >= n_nodes && b->dinv) return DG_OK;
    if (b->dinv) cudaFree(b->dinv);
    if (b->keep) cudaFree(b->keep);
    if (b->x0) cudaFree(b->x0);
    b->dinv = nullptr;
    b->keep = nullptr;
    b->x0 = nullptr;
    DG_CUDA_CHECK(cudaMalloc(&b->dinv, sizeof(float) * std::max<size_t>(n_nodes, 1)));
    b->cap_nodes = n_nodes;
    return DG_OK;
}

int validate_graph_ptr(const std::vector<int32_t> &gp, int n_graphs, int n_nodes, int *max_nodes) {
    DG_REQUIRE((int)gp.size() == n_graphs + 1, DG_ERR_INVALID, "graph_ptr length");
    DG_REQUIRE(gp[0] == 0 && gp[n_graphs] == n_nodes, DG_ERR_INVALID,
               "graph_ptr must start at 0 and end at n_nodes (got %d .. %d, n_nodes %d)", gp[0], gp[n_graphs],
               n_nodes);
    int mx = 0;
    for (int g = 0; g < n_graphs; ++g) {
        DG_REQUIRE(gp[g + 1] >= gp[g], DG_ERR_INVALID, "graph_ptr must be non-decreasing (graph %d)", g);
        mx = std::max(mx, gp[g + 1] - gp[g]);
    }
    *max_nodes = mx;
    return DG_OK;
}

}  // namespace
}  // namespace dg

using namespace dg;

extern "C" {

int dg_version(void) { return DG_VERSION; }

const char *dg_last_error(void) { return g_error; }

int dg_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

// -------------------------------------------------------------------------------------------------
int dg_context_create(int device, void *stream, dg_context **out) {
    clear_error();
    DG_REQUIRE(out != nullptr, DG_ERR_INVALID, "null out pointer");
    *out = nullptr;
    int ndev = dg_device_count();
    DG_REQUIRE(ndev > 0, DG_ERR_NO_DEVICE, "no CUDA device visible: distgcn_b200 has no CPU fallback");
    DG_REQUIRE(device >= 0 && device < ndev, DG_ERR_INVALID, "device %d out of range (0..%d)", device, ndev - 1);
    DG_CUDA_CHECK(cudaSetDevice(device));
    cudaDeviceProp prop;
    DG_CUDA_CHECK(cudaGetDeviceProperties(&prop, device));
    DG_REQUIRE(prop.major >= 10, DG_ERR_UNSUPPORTED,
               "device %d is sm_%d%d; this library is built for sm_100a (B200) only", device, prop.major,
               prop.minor);
    dg_context *ctx = new (std::nothrow) dg_context();
    DG_REQUIRE(ctx != nullptr, DG_ERR_INVALID, "out of host memory");
    ctx->device = device;
    env_read(&ctx->env);
    ctx->sm_count = prop.multiProcessorCount;
    ctx->max_smem_optin = (int)prop.sharedMemPerBlockOptin;
    ctx->slots.resize(kSlotCount);
    if (stream) {
        ctx->stream = (cudaStream_t)stream;
        ctx->own_stream = false;
    } else {
        cudaError_t e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking);
        if (e != cudaSuccess) {
            set_error("cudaStreamCreate failed: %s", cudaGetErrorString(e));
            delete ctx;
            return DG_ERR_CUDA;
        }
        ctx->own_stream = true;
    }
    if (cudaHostAlloc((void **)&ctx->h_flag, sizeof(int) * 4, cudaHostAllocDefault) != cudaSuccess ||
        cudaMalloc((void **)&ctx->d_status, sizeof(int) * 4) != cudaSuccess) {
        set_error("context allocation failed: %s", cudaGetErrorString(cudaGetLastError()));
        dg_context_destroy(ctx);
        return DG_ERR_CUDA;
    }
    memset(ctx->h_flag, 0, sizeof(int) * 4);
    cudaMemsetAsync(ctx->d_status, 0, sizeof(int) * 4, ctx->stream);
    *out = ctx;
    return DG_OK;
}

void dg_context_destroy(dg_context *ctx) {
    if (!ctx) return;
    DeviceGuard guard(ctx->device);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    if (ctx->host_batch) {
        dg_batch_destroy(ctx->host_batch);
        ctx->host_batch = nullptr;
    }
    ingest_staging_free(ctx);
    for (auto &b : ctx->slots)
        if (b.ptr) cudaFree(b.ptr);
    if (ctx->h_flag) cudaFreeHost(ctx->h_flag);
    if (ctx->d_status) cudaFree(ctx->d_status);
    for (auto e : ctx->prof_events) cudaEventDestroy(e);
    for (int k = 0; k < 2; ++k)
        if (ctx->timer_ev[k]) cudaEventDestroy(ctx->timer_ev[k]);
    if (ctx->order_ev) cudaEventDestroy(ctx->order_ev);
    if (ctx->own_stream && ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

int dg_context_reload_env(dg_context *ctx) {
    clear_error();
    DG_TRY(check_ctx(ctx));
    env_read(&ctx->env);
    return DG_OK;
}

int dg_context_synchronize(dg_context *ctx) {
    clear_error();
    DG_TRY(check_ctx(ctx));
    DeviceGuard guard(ctx->device);
    return finish(ctx);
}

void *dg_host_alloc(uint64_t bytes) {
    void *p = nullptr;
    if (cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocDefault) != cudaSuccess) {
        cudaGetLastError();
        return nullptr;
    }
    return p;
}

void dg_host_free(void *p) {
    if (p) cudaFreeHost(p);
}

uint64_t dg_context_launch_count(const dg_context *ctx) { return ctx ? ctx->launches : 0; }

const char *dg_context_last_kernel(const dg_context *ctx) { return ctx ? ctx->last_kernel : ""; }

int dg_timer_start(dg_context *ctx) {
    clear_error();
    DG_TRY(check_ctx(ctx));
    DeviceGuard guard(ctx->device);
    for (int k = 0; k < 2; ++k)
        if (!ctx->timer_ev[k]) DG_CUDA_CHECK(cudaEventCreate(&ctx->timer_ev[k]));
    DG_CUDA_CHECK(cudaEventRecord(ctx->timer_ev[0], ctx->stream));
    return DG_OK;
}

int dg_timer_stop(dg_context *ctx, double *elapsed_ms) {
    clear_error();
    DG_TRY(check_ctx(ctx));
    DG_REQUIRE(ctx->timer_ev[0] && ctx->timer_ev[1], DG_ERR_INVALID, "dg_timer_stop without dg_timer_start");
    DeviceGuard guard(ctx->device);
    DG_CUDA_CHECK(cudaEventRecord(ctx->timer_ev[1], ctx->stream));
    DG_CUDA_CHECK(cudaEventSynchronize(ctx->timer_ev[1]));
    float ms = 0.f;
    DG_CUDA_CHECK(cudaEventElapsedTime(&ms, ctx->timer_ev[0], ctx->timer_ev[1]));
    if (elapsed_ms) *elapsed_ms = ms;
    return DG_OK;
}

int dg_context_wait(dg_context *ctx, dg_context *other) {
    clear_error();
    DG_TRY(check_ctx(ctx));
    DG_TRY(check_ctx(other));
    DG_REQUIRE(ctx->device == other->device, DG_ERR_INVALID, "both contexts must live on the same device");
    if (ctx == other) return DG_OK;
    DeviceGuard guard(ctx->device);
    if (!other->order_ev) DG_CUDA_CHECK(cudaEventCreateWithFlags(&other->order_ev, cudaEventDisableTiming));
    DG_CUDA_CHECK(cudaEventRecord(other->order_ev, other->stream));
    DG_CUDA_CHECK(cudaStreamWaitEvent(ctx->stream, other->order_ev, 0));
    return DG_OK;
}

int dg_profile_enable(dg_context *ctx, int on) {
    clear_error();
    DG_TRY(check_ctx(ctx));
    ctx->prof_on = on != 0;
    return DG_OK;
}

int dg_profile_collect(dg_context *ctx, double *total_ms, uint64_t *launches, double *algorithmic_bytes) {
    clear_error();
    DG_TRY(check_ctx(ctx));
    DeviceGuard guard(ctx->device);
    DG_CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    double sum = 0.0;
    for (size_t k = 0; k + 1 < ctx->prof_used; k += 2) {
        float ms = 0.f;
        DG_CUDA_CHECK(cudaEventElapsedTime(&ms, ctx->prof_events[k], ctx->prof_events[k + 1]));
        sum += ms;
    }
    if (total_ms) *total_ms = sum;
    if (launches) *launches = ctx->prof_used / 2;
    if (algorithmic_bytes) *algorithmic_bytes = ctx->prof_bytes;
    ctx->prof_used = 0;
    ctx->prof_bytes = 0.0;
    return DG_OK;
}

// -------------------------------------------------------------------------------------------------
int dg_model_create(dg_context *ctx, int n_layers, int n_supports, const int32_t *c_in, const int32_t *c_out,
                    const float *const *weights, const float *const *bias, const int32_t *act,
                    float leaky_alpha, int head, dg_model **out) {
    clear_error();
    DG_TRY(check_ctx(ctx));
    DG_REQUIRE(out != nullptr, DG_ERR_INVALID, "null out pointer");
    *out = nullptr;
    DG_REQUIRE(n_layers >= 1 && n_layers <= kMaxLayers, DG_ERR_INVALID, "n_layers %d out of range", n_layers);
    DG_REQUIRE(c_in && c_out && weights && act, DG_ERR_INVALID, "null model argument");
    DG_REQUIRE(head == DG_HEAD_LINEAR || head == DG_HEAD_PAIR_SOFTMAX, DG_ERR_INVALID, "unknown head %d", head);
    if (n_supports != 2) {
        // Higher polynomial orders ([I, L, L^2, ..], simple_polynomials with max_degree >= 2): implemented for networks
        // whose layers all have ONE output column - the shape of the shipped cheb2 checkpoints (1 -> 1 -> 1 and 32 -> 1).
        bool scalar = n_supports >= 1 && n_supports <= 8 && head == DG_HEAD_LINEAR;
        for (int l = 0; l < n_layers && scalar; ++l) scalar = c_out[l] == 1 && c_in[l] >= 1 && (l == 0 || c_in[l] == 1);
        DG_REQUIRE(scalar, DG_ERR_UNSUPPORTED,
                   "n_supports %d: beyond the cheb1 supports [I, L] only networks whose layers all have one output column are "
                   "implemented", n_supports);
        DeviceGuard guard(ctx->device);
        dg_model *m = new (std::nothrow) dg_model();
        DG_REQUIRE(m != nullptr, DG_ERR_INVALID, "out of host memory");
        m->ctx = ctx;
        m->n_layers = n_layers;
        m->n_supports = n_supports;
        m->alpha = leaky_alpha;
        m->head = head;
        m->scalar_net = true;
        m->layers.resize(n_layers);
        for (int l = 0; l < n_layers; ++l) {
            DG_REQUIRE(act[l] >= DG_ACT_IDENTITY && act[l] <= DG_ACT_RELU, DG_ERR_INVALID, "layer %d: activation", l);
            m->layers[l].c_in = c_in[l];
            m->layers[l].c_out = 1;
            m->layers[l].cpo = pad_width(1);
            m->layers[l].act = act[l];
            for (int k = 0; k < n_supports; ++k) {
                const float *w = weights[(size_t)n_supports * l + k];
                if (!w) {
                    delete m;
                    set_error("layer %d: null weights", l);
                    return DG_ERR_INVALID;
                }
                float sum = 0.f;  // sequential fp32 sum over the constant input features (TF's sparse x dense order)
                for (int f = 0; f < c_in[l]; ++f) sum += w[f];
                m->scalar_w.push_back(sum);
            }
            m->scalar_b.push_back((bias && bias[l]) ? bias[l][0] : 0.f);
        }
        *out = m;
        return DG_OK;
    }
    for (int l = 0; l < n_layers; ++l) {
        DG_REQUIRE(c_in[l] >= 1 && c_out[l] >= 1, DG_ERR_INVALID, "layer %d: empty shape", l);
        DG_REQUIRE(c_out[l] <= kMaxWidth && (l == 0 || c_in[l] <= kMaxWidth), DG_ERR_UNSUPPORTED,
                   "layer %d: width %d -> %d exceeds %d", l, c_in[l], c_out[l], kMaxWidth);
        DG_REQUIRE(l == 0 || c_in[l] == c_out[l - 1], DG_ERR_INVALID, "layer %d: c_in %d != previous c_out %d", l,
                   c_in[l], c_out[l - 1]);
        DG_REQUIRE(act[l] >= DG_ACT_IDENTITY && act[l] <= DG_ACT_RELU, DG_ERR_INVALID, "layer %d: activation", l);
        DG_REQUIRE(weights[2 * l] && weights[2 * l + 1], DG_ERR_INVALID, "layer %d: null weights", l);
    }
    if (head == DG_HEAD_PAIR_SOFTMAX)
        DG_REQUIRE(c_out[n_layers - 1] % 2 == 0, DG_ERR_INVALID, "pair-softmax head needs an even output width");
    DeviceGuard guard(ctx->device);
    dg_model *m = new (std::nothrow) dg_model();
    DG_REQUIRE(m != nullptr, DG_ERR_INVALID, "out of host memory");
    m->ctx = ctx;
    m->n_layers = n_layers;
    m->n_supports = n_supports;
    m->alpha = leaky_alpha;
    m->head = head;
    m->layers.resize(n_layers);
    int status = DG_OK;
    auto upload = [&](float **dst, const std::vector<float> &src) -> int {
        DG_CUDA_CHECK(cudaMalloc((void **)dst, sizeof(float) * src.size()));
        DG_CUDA_CHECK(cudaMemcpy(*dst, src.data(), sizeof(float) * src.size(), cudaMemcpyHostToDevice));
        return DG_OK;
    };
    for (int l = 0; l < n_layers && status == DG_OK; ++l) {
        dg_layer_dev &L = m->layers[l];
        L.c_in = c_in[l];
        L.c_out = c_out[l];
        // layer 0 is applied in its rank-1 form (any input width); its padded matrix is only
        // built when it fits the fused kernel
        const bool fits = c_in[l] <= kMaxWidth;
        L.cpi = fits ? pad_width(c_in[l]) : 0;
        L.cpo = pad_width(c_out[l]);
        L.act = act[l];
        L.has_bias = bias && bias[l];
        const float *w0 = weights[2 * l], *w1 = weights[2 * l + 1];
        std::vector<float> hb(L.cpo, 0.f), cs0(L.cpo, 0.f), cs1(L.cpo, 0.f);
        if (L.has_bias)
            for (int c = 0; c < L.c_out; ++c) hb[c] = bias[l][c];
        for (int c = 0; c < L.c_out; ++c) {
            float s0 = 0.f, s1 = 0.f;  // sequential fp32 sums, the order TF's sparse x dense product uses
            for (int k = 0; k < L.c_in; ++k) {
                s0 += w0[(size_t)k * L.c_out + c];
                s1 += w1[(size_t)k * L.c_out + c];
            }
            cs0[c] = s0;
            cs1[c] = s1;
        }
        if (fits) {
            std::vector<float> wc((size_t)2 * L.cpi * L.cpo, 0.f);
            for (int k = 0; k < L.c_in; ++k)
                for (int c = 0; c < L.c_out; ++c) {
                    wc[(size_t)k * L.cpo + c] = w0[(size_t)k * L.c_out + c];
                    wc[(size_t)(L.cpi + k) * L.cpo + c] = w1[(size_t)k * L.c_out + c];
                }
            status = upload(&L.wcat, wc);
        }
        if (status == DG_OK) status = upload(&L.bias, hb);
        if (status == DG_OK) status = upload(&L.colsum0, cs0);
        if (status == DG_OK) status = upload(&L.colsum1, cs1);
    }
    if (status == DG_OK && n_layers >= 2 && c_out[n_layers - 1] == 1) {
        const int l = n_layers - 1;
        const int cpi = pad_width(c_in[l]);
        std::vector<float> t0(cpi, 0.f), t1(cpi, 0.f);
        for (int k = 0; k < c_in[l]; ++k) {
            t0[k] = weights[2 * l][k];
            t1[k] = weights[2 * l + 1][k];
        }
        status = upload(&m->tail_w0, t0);
        if (status == DG_OK) status = upload(&m->tail_w1, t1);
        m->tail_bias = (bias && bias[l]) ? bias[l][0] : 0.f;
    }
    // operands of the graph-resident fused kernel
    if (status == DG_OK) {
        std::vector<int> hacts(act, act + n_layers);
        status = [&]() -> int {
            DG_CUDA_CHECK(cudaMalloc((void **)&m->d_acts, sizeof(int) * n_layers));
            DG_CUDA_CHECK(cudaMemcpy(m->d_acts, hacts.data(), sizeof(int) * n_layers, cudaMemcpyHostToDevice));
            return DG_OK;
        }();
    }
    if (status == DG_OK && head == DG_HEAD_LINEAR && c_out[n_layers - 1] == 1) {
        int widest = 1;
        for (int l = 0; l + 1 < n_layers; ++l) widest = std::max(widest, (int)c_out[l]);
        const int cp = pad_width(widest);
        const bool has_hidden = n_layers >= 3;
        if (cp == 32 || !has_hidden) {
            const dg_layer_dev &L0 = m->layers[0];
            std::vector<float> first(3 * cp, 0.f), tail(2 * cp, 0.f);
            // layer 0 as the rank-1 map x0 * colsum(W_0) + s * colsum(W_1) + b (padded to cp columns)
            std::vector<float> cs0(L0.cpo), cs1(L0.cpo), b0(L0.cpo);
            status = [&]() -> int {
                DG_CUDA_CHECK(cudaMemcpy(cs0.data(), L0.colsum0, sizeof(float) * L0.cpo, cudaMemcpyDeviceToHost));
                DG_CUDA_CHECK(cudaMemcpy(cs1.data(), L0.colsum1, sizeof(float) * L0.cpo, cudaMemcpyDeviceToHost));
                DG_CUDA_CHECK(cudaMemcpy(b0.data(), L0.bias, sizeof(float) * L0.cpo, cudaMemcpyDeviceToHost));
                return DG_OK;
            }();
            const int ncopy = n_layers == 1 ? 1 : std::min(cp, L0.cpo);
            for (int c = 0; c < ncopy && c < L0.c_out; ++c) {
                first[c] = cs0[c];
                first[cp + c] = cs1[c];
                first[2 * cp + c] = b0[c];
            }
            if (n_layers >= 2) {
                const int l = n_layers - 1;
                for (int k = 0; k < c_in[l]; ++k) {
                    tail[k] = weights[2 * l][k];
                    tail[cp + k] = weights[2 * l + 1][k];
                }
            }
            if (status == DG_OK) status = upload(&m->fused_first, first);
            if (status == DG_OK) status = upload(&m->fused_tail, tail);
            if (status == DG_OK && has_hidden) {
                const size_t blob = (size_t)2 * cp * cp + cp;
                std::vector<float> wall(blob * (n_layers - 2), 0.f);
                for (int l = 1; l + 1 < n_layers; ++l) {
                    float *dst = &wall[blob * (l - 1)];
                    for (int k = 0; k < c_in[l]; ++k)
                        for (int c = 0; c < c_out[l]; ++c) {
                            dst[(size_t)k * cp + c] = weights[2 * l][(size_t)k * c_out[l] + c];
                            dst[(size_t)(cp + k) * cp + c] = weights[2 * l + 1][(size_t)k * c_out[l] + c];
                        }
                    if (bias && bias[l])
                        for (int c = 0; c < c_out[l]; ++c) dst[(size_t)2 * cp * cp + c] = bias[l][c];
                }
                status = upload(&m->fused_wall, wall);
                if (status == DG_OK && cp == 32) {
                    // tensor-core layout: every weight split into TF32 hi + lo (round to nearest, ties
                    // away, as cvt.rna does), rows padded to 40 words (conflict-free fragment loads)
                    auto tf32_rna = [](float x) -> float {
                        uint32_t b;
                        memcpy(&b, &x, 4);
                        if ((b & 0x7f800000u) == 0x7f800000u) return x;  // inf / nan unchanged
                        b += 0x00001000u;
                        b &= 0xffffe000u;
                        float y;
                        memcpy(&y, &b, 4);
                        return y;
                    };
                    const int ws = 40;
                    const size_t mblob = (size_t)2 * 64 * ws + 32;
                    std::vector<float> wm(mblob * (n_layers - 2), 0.f);
                    for (int l = 1; l + 1 < n_layers; ++l) {
                        const float *src = &wall[blob * (l - 1)];  // [64][32] rows: W_0 then W_1, then bias
                        float *dst = &wm[mblob * (l - 1)];
                        for (int k = 0; k < 64; ++k)
                            for (int c = 0; c < 32; ++c) {
                                const float v = src[(size_t)k * 32 + c];
                                const float hi = tf32_rna(v);
                                dst[(size_t)k * ws + c] = hi;
                                dst[(size_t)64 * ws + (size_t)k * ws + c] = tf32_rna(v - hi);
                            }
                        for (int c = 0; c < 32; ++c) dst[(size_t)2 * 64 * ws + c] = src[(size_t)2 * 32 * 32 + c];
                    }
                    status = upload(&m->fused_wall_mma, wm);
                }
            }
            if (status == DG_OK && has_hidden && cp == 32) {
                // tensor-core kernel: bf16-split projection operands and the fixed-point bound factors
                const int nh = n_layers - 2;
                std::vector<const float *> hw0(nh), hw1(nh), hb(nh);
                std::vector<int> hci(nh), hco(nh);
                for (int l = 1; l + 1 < n_layers; ++l) {
                    hw0[l - 1] = weights[2 * l];
                    hw1[l - 1] = weights[2 * l + 1];
                    hb[l - 1] = bias ? bias[l] : nullptr;
                    hci[l - 1] = c_in[l];
                    hco[l - 1] = c_out[l];
                }
                std::vector<unsigned char> blob;
                tc_build_weights(nh, hw0.data(), hw1.data(), hb.data(), hci.data(), hco.data(), &blob);
                status = [&]() -> int {
                    DG_CUDA_CHECK(cudaMalloc((void **)&m->tc_wall, blob.size()));
                    DG_CUDA_CHECK(cudaMemcpy(m->tc_wall, blob.data(), blob.size(), cudaMemcpyHostToDevice));
                    return DG_OK;
                }();
                double tn = 0.0;
                for (int k = 0; k < c_in[n_layers - 1]; ++k) tn += fabs((double)weights[2 * (n_layers - 1) + 1][k]);
                m->tc_tail_norm = (float)(tn * (1.0 + 1.0 / 1024.0));
            }
            if (status == DG_OK) {
                m->fused_cp = cp;
                if (n_layers == 1) m->tail_bias = 0.f;
            }
        }
    }
    if (status != DG_OK) {
        dg_model_destroy(m);
        return status;
    }
    *out = m;
    return DG_OK;
}

void dg_model_destroy(dg_model *m) {
    if (!m) return;
    DeviceGuard guard(m->ctx->device);
    cudaStreamSynchronize(m->ctx->stream);
    for (auto &L : m->layers) {
        if (L.wcat) cudaFree(L.wcat);
        if (L.bias) cudaFree(L.bias);
        if (L.colsum0) cudaFree(L.colsum0);
        if (L.colsum1) cudaFree(L.colsum1);
    }
    if (m->tail_w0) cudaFree(m->tail_w0);
    if (m->tail_w1) cudaFree(m->tail_w1);
    if (m->fused_first) cudaFree(m->fused_first);
    if (m->fused_wall) cudaFree(m->fused_wall);
    if (m->fused_wall_mma) cudaFree(m->fused_wall_mma);
    if (m->fused_tail) cudaFree(m->fused_tail);
    if (m->tc_wall) cudaFree(m->tc_wall);
    if (m->d_acts) cudaFree(m->d_acts);
    delete m;
}

int dg_model_out_width(const dg_model *m) { return m ? m->layers.back().c_out : 0; }

// -------------------------------------------------------------------------------------------------
// graph-local 16-bit column ids -> batch-global int32 ids, one CTA per graph (the compact host format of dg_solve_host_compact)
__global__ void __launch_bounds__(256)
expand_cols16_kernel(const int32_t *__restrict__ graph_ptr, const int32_t *__restrict__ row_ptr,
                     const uint16_t *__restrict__ col16, int32_t *__restrict__ col_idx) {
    const int v0 = graph_ptr[blockIdx.x], v1 = graph_ptr[blockIdx.x + 1];
    const int e1 = row_ptr[v1];
    for (int e = row_ptr[v0] + (int)threadIdx.x; e < e1; e += (int)blockDim.x) col_idx[e] = (int32_t)col16[e] + v0;
}

// Upper-triangle lists (column > row, graph-local 16-bit ids) -> the full symmetric CSR with batch-global ids, one CTA per
// graph: per-vertex counts with shared-memory atomics, an exclusive scan for the row offsets, the upper entries copied in
// place behind each row's lower part, the lower parts filled through per-row cursors and then sorted (rows come out in
// ascending column order whatever the order of the atomics: the result is deterministic).
constexpr int kSymMaxNodes = 8192;
__global__ void __launch_bounds__(256)
symmetrize_upper_kernel(const int32_t *__restrict__ graph_ptr, const int32_t *__restrict__ row_ptr_u,
                        const uint16_t *__restrict__ col_u, int32_t *__restrict__ row_ptr, int32_t *__restrict__ col_idx,
                        int n_graphs) {
    extern __shared__ int sym_sm[];
    const int g = blockIdx.x;
    const int v0 = graph_ptr[g], n = graph_ptr[g + 1] - v0;
    int *cnt = sym_sm;          // [n] degree, then exclusive offsets
    int *low = sym_sm + n;      // [n] entries below the diagonal (cursor while filling)
    __shared__ int part[256];
    const int tid = threadIdx.x;
    const int eu0 = row_ptr_u[v0];
    const int base = 2 * eu0;   // full offsets: every earlier graph holds twice its upper count
    for (int i = tid; i < n; i += 256) cnt[i] = 0, low[i] = 0;
    __syncthreads();
    for (int i = tid; i < n; i += 256) {
        const int b0 = row_ptr_u[v0 + i], b1 = row_ptr_u[v0 + i + 1];
        atomicAdd(&cnt[i], b1 - b0);
        for (int e = b0; e < b1; ++e) {
            const int j = col_u[e];
            atomicAdd(&cnt[j], 1);
            atomicAdd(&low[j], 1);
        }
    }
    __syncthreads();
    // exclusive scan of cnt: contiguous chunks per thread
    const int chunk = (n + 255) / 256;
    const int c0 = min(tid * chunk, n), c1 = min(c0 + chunk, n);
    int sum = 0;
    for (int i = c0; i < c1; ++i) sum += cnt[i];
    part[tid] = sum;
    __syncthreads();
    if (tid == 0) {
        int run = 0;
        for (int t = 0; t < 256; ++t) {
            const int v = part[t];
            part[t] = run;
            run += v;
        }
    }
    __syncthreads();
    int run = part[tid];
    for (int i = c0; i < c1; ++i) {
        const int v = cnt[i];
        cnt[i] = run;
        row_ptr[v0 + i] = base + run;
        run += v;
    }
    if (g == n_graphs - 1 && tid == 255) row_ptr[v0 + n] = 2 * row_ptr_u[v0 + n];
    __syncthreads();
    // upper parts in place (ascending as given), lower parts through the cursors
    for (int i = tid; i < n; i += 256) {
        const int b0 = row_ptr_u[v0 + i], b1 = row_ptr_u[v0 + i + 1];
        int *dst = col_idx + base + cnt[i] + low[i];
        for (int e = b0; e < b1; ++e) dst[e - b0] = v0 + (int)col_u[e];
    }
    __syncthreads();   // `low` becomes the fill cursor counting down
    for (int i = tid; i < n; i += 256) {
        const int b0 = row_ptr_u[v0 + i], b1 = row_ptr_u[v0 + i + 1];
        for (int e = b0; e < b1; ++e) {
            const int j = col_u[e];
            const int pos = atomicSub(&low[j], 1) - 1;
            col_idx[base + cnt[j] + pos] = v0 + i;
        }
    }
    __syncthreads();
    __threadfence_block();
    // sort every row's lower part (short lists: insertion sort, one thread per row)
    for (int i = tid; i < n; i += 256) {
        const int b0 = row_ptr_u[v0 + i], b1 = row_ptr_u[v0 + i + 1];
        const int end_i = (i + 1 < n) ? cnt[i + 1] : -1;
        const int deg = (end_i >= 0 ? end_i : (2 * (row_ptr_u[v0 + n] - eu0))) - cnt[i];
        const int nl = deg - (b1 - b0);
        int *a = col_idx + base + cnt[i];
        for (int x = 1; x < nl; ++x) {
            const int key = a[x];
            int y = x - 1;
            while (y >= 0 && a[y] > key) {
                a[y + 1] = a[y];
                --y;
            }
            a[y + 1] = key;
        }
    }
}


// The same for graphs whose lists fit in shared memory (every graph of the reference's sets): the upper lists are staged
// with coalesced loads, counted, scanned, expanded and sorted entirely in shared memory as 16-bit local ids, and the full
// rows leave with coalesced stores.  (The global-memory version above walks its rows thread by thread and insertion-sorts
// in global memory: 45 us for the 500-graph ER set against 52 us for the solve itself.)
__host__ __device__ inline size_t sym_smem_bytes(int n, int nnz_u) {
    // rp_u [n + 1] + cnt [n] + low [n] ints, then col_u [nnz_u] and full [2 nnz_u] 16-bit ids (each padded to 4 bytes)
    return sizeof(int) * (3 * (size_t)n + 2) + sizeof(uint16_t) * (3 * (size_t)nnz_u + 4);
}
__global__ void __launch_bounds__(256)
symmetrize_upper_smem_kernel(const int32_t *__restrict__ graph_ptr, const int32_t *__restrict__ row_ptr_u,
                             const uint16_t *__restrict__ col_u, int32_t *__restrict__ row_ptr, int32_t *__restrict__ col_idx,
                             int n_graphs) {
    extern __shared__ int sym_sm[];
    const int g = blockIdx.x, tid = threadIdx.x;
    const int v0 = graph_ptr[g], n = graph_ptr[g + 1] - v0;
    const int eu0 = row_ptr_u[v0], nu = row_ptr_u[v0 + n] - eu0;
    int *rp = sym_sm;                 // [n + 1] upper row offsets, local
    int *cnt = rp + n + 1;            // [n] degree, then exclusive offsets
    int *low = cnt + n;               // [n] entries below the diagonal (cursor while filling)
    uint16_t *cu = reinterpret_cast<uint16_t *>(low + n + 1);   // [nu] upper lists
    uint16_t *full = cu + ((nu + 1) & ~1);                       // [2 nu] full rows, local ids
    __shared__ int part[256];
    const int base = 2 * eu0;
    for (int i = tid; i <= n; i += 256) rp[i] = row_ptr_u[v0 + i] - eu0;
    for (int e = tid; e < nu; e += 256) cu[e] = col_u[eu0 + e];
    for (int i = tid; i < n; i += 256) cnt[i] = 0, low[i] = 0;
    __syncthreads();
    for (int i = tid; i < n; i += 256) {
        const int b0 = rp[i], b1 = rp[i + 1];
        atomicAdd(&cnt[i], b1 - b0);
        for (int e = b0; e < b1; ++e) {
            const int j = cu[e];
            atomicAdd(&cnt[j], 1);
            atomicAdd(&low[j], 1);
        }
    }
    __syncthreads();
    const int chunk = (n + 255) / 256;
    const int c0 = min(tid * chunk, n), c1 = min(c0 + chunk, n);
    int sum = 0;
    for (int i = c0; i < c1; ++i) sum += cnt[i];
    part[tid] = sum;
    __syncthreads();
    if (tid < 32) {   // exclusive scan of the 256 partial sums: 8 per lane, then a warp scan
        int loc[8], tot = 0;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            loc[k] = tot;
            tot += part[tid * 8 + k];
        }
        int inc = tot;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            const int up = __shfl_up_sync(0xffffffffu, inc, off);
            if (tid >= off) inc += up;
        }
        const int excl = inc - tot;
#pragma unroll
        for (int k = 0; k < 8; ++k) part[tid * 8 + k] = excl + loc[k];
    }
    __syncthreads();
    int run = part[tid];
    for (int i = c0; i < c1; ++i) {
        const int v = cnt[i];
        cnt[i] = run;
        run += v;
    }
    __syncthreads();
    for (int i = tid; i < n; i += 256) {
        const int b0 = rp[i], b1 = rp[i + 1];
        uint16_t *dst = full + cnt[i] + low[i];
        for (int e = b0; e < b1; ++e) dst[e - b0] = cu[e];
    }
    __syncthreads();   // `low` becomes the fill cursor counting down
    for (int i = tid; i < n; i += 256) {
        const int b0 = rp[i], b1 = rp[i + 1];
        for (int e = b0; e < b1; ++e) {
            const int j = cu[e];
            const int pos = atomicSub(&low[j], 1) - 1;
            full[cnt[j] + pos] = (uint16_t)i;
        }
    }
    __syncthreads();
    for (int i = tid; i < n; i += 256) {   // every row's lower part in ascending order (short lists: insertion sort)
        const int deg = ((i + 1 < n) ? cnt[i + 1] : 2 * nu) - cnt[i];
        const int nl = deg - (rp[i + 1] - rp[i]);
        uint16_t *a = full + cnt[i];
        for (int x = 1; x < nl; ++x) {
            const uint16_t key = a[x];
            int y = x - 1;
            while (y >= 0 && a[y] > key) {
                a[y + 1] = a[y];
                --y;
            }
            a[y + 1] = key;
        }
    }
    __syncthreads();
    for (int i = tid; i < n; i += 256) row_ptr[v0 + i] = base + cnt[i];
    if (g == n_graphs - 1 && tid == 0) row_ptr[v0 + n] = base + 2 * nu;
    for (int k = tid; k < 2 * nu; k += 256) col_idx[base + k] = v0 + (int)full[k];
}

// batch-global int32 column ids (and, for the upper format, the full symmetric CSR) of a batch that arrived in a compact
// host format (no-op otherwise)
static int batch_ensure_cols(dg_batch *b) {
    if (!b->cols_pending) return DG_OK;
    dg_context *ctx = b->ctx;
    if (b->upper_pending) {
        if (b->n_graphs > 0 && b->n_nodes > 0) {
            // (max_graph_nnz still counts upper entries here)
            const size_t smem_fast = sym_smem_bytes(std::max(b->max_graph_nodes, 1), std::max(b->max_graph_nnz, 0));
            if (smem_fast <= 64 * 1024) {   // several CTAs per SM: the reference's graphs need 10-30 KB
                static std::atomic<unsigned long long> fast_attr_done{0};
                DG_CUDA_CHECK(smem_attr_once(symmetrize_upper_smem_kernel, ctx->device, 64 * 1024, &fast_attr_done));
                symmetrize_upper_smem_kernel<<<b->n_graphs, 256, smem_fast, ctx->stream>>>(b->graph_ptr, b->row_ptr_u, b->col16,
                                                                                          b->row_ptr, b->col_idx, b->n_graphs);
                ctx->launches++;
                DG_CUDA_CHECK(cudaGetLastError());
            } else {
            const size_t smem = sizeof(int) * 2 * (size_t)std::max(b->max_graph_nodes, 1);
            static std::atomic<unsigned long long> attr_done{0};
            DG_CUDA_CHECK(smem_attr_once(symmetrize_upper_kernel, ctx->device, (int)(sizeof(int) * 2 * kSymMaxNodes), &attr_done));
            symmetrize_upper_kernel<<<b->n_graphs, 256, smem, ctx->stream>>>(b->graph_ptr, b->row_ptr_u, b->col16, b->row_ptr,
                                                                            b->col_idx, b->n_graphs);
            ctx->launches++;
            DG_CUDA_CHECK(cudaGetLastError());
            }
        }
        b->upper_pending = false;
        b->cols_pending = false;
        b->nnz *= 2;
        for (auto &e : b->h_graph_e) e *= 2;
        b->max_graph_nnz *= 2;
        b->tiles_valid = false;
        b->tc_tiles_valid = false;
        b->gs_valid = false;
        if (b->dinv_deferred) {
            b->dinv_deferred = false;
            DG_TRY(batch_compute_dinv(b));
        }
        return DG_OK;
    }
    if (b->n_graphs > 0 && b->nnz > 0) {
        expand_cols16_kernel<<<b->n_graphs, 256, 0, b->ctx->stream>>>(b->graph_ptr, b->row_ptr, b->col16, b->col_idx);
        b->ctx->launches++;
        DG_CUDA_CHECK(cudaGetLastError());
    }
    b->cols_pending = false;
    return DG_OK;
}

static int batch_fill(dg_batch *b, int32_t n_graphs, int32_t n_nodes, int32_t nnz, const int32_t *graph_ptr,
                      const int32_t *row_ptr, const int32_t *col_idx, int mem, bool compute_dinv = true,
                      const uint16_t *col_local16 = nullptr, bool upper = false) {
    dg_context *ctx = b->ctx;
    DG_REQUIRE(!upper || col_local16, DG_ERR_INVALID, "the upper-triangle format is a compact (16-bit) host format");
    b->upper_pending = false;
    b->dinv_deferred = false;
    DG_REQUIRE(n_graphs >= 0 && n_nodes >= 0 && nnz >= 0, DG_ERR_INVALID, "negative size");
    DG_REQUIRE(graph_ptr && row_ptr && (col_idx || col_local16 || nnz == 0), DG_ERR_INVALID, "null CSR pointer");
    DG_REQUIRE(!col_local16 || mem == DG_MEM_HOST, DG_ERR_INVALID, "the compact column format is a host format");
    b->cols_pending = false;
    const bool meta_ready = b->meta_ready && mem == DG_MEM_HOST && b->n_graphs == n_graphs && b->n_nodes == n_nodes &&
                            b->nnz == nnz;
    b->meta_ready = false;
    b->n_graphs = n_graphs;
    b->n_nodes = n_nodes;
    b->nnz = nnz;
    if (meta_ready) goto upload;  // host_batch_set_meta has described this very batch (and a tile plan may be waiting)
    b->tc_plan_ready = false;
    b->h_graph_ptr.resize((size_t)n_graphs + 1);
    if (mem == DG_MEM_HOST) {
        std::copy(graph_ptr, graph_ptr + n_graphs + 1, b->h_graph_ptr.begin());
        DG_REQUIRE(row_ptr[0] == 0 && row_ptr[n_nodes] == nnz, DG_ERR_INVALID,
                   "row_ptr must start at 0 and end at nnz (got %d .. %d, nnz %d)", row_ptr[0], row_ptr[n_nodes],
                   nnz);
    } else {
        DG_CUDA_CHECK(cudaMemcpyAsync(b->h_graph_ptr.data(), graph_ptr, sizeof(int32_t) * ((size_t)n_graphs + 1),
                                      cudaMemcpyDeviceToHost, ctx->stream));
        DG_CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    }
    DG_TRY(validate_graph_ptr(b->h_graph_ptr, n_graphs, n_nodes, &b->max_graph_nodes));
    // row_ptr at the graph boundaries (host copy): per-graph nnz for the fused kernel's tile table
    b->tiles_valid = false;
    b->tc_tiles_valid = false;
    b->gs_valid = false;
    b->h_graph_e.resize((size_t)n_graphs + 1);
    if (mem == DG_MEM_HOST) {
        for (int g = 0; g <= n_graphs; ++g) b->h_graph_e[g] = row_ptr[b->h_graph_ptr[g]];
    } else {
        for (int g = 0; g <= n_graphs; ++g)
            DG_CUDA_CHECK(cudaMemcpyAsync(&b->h_graph_e[g], row_ptr + b->h_graph_ptr[g], sizeof(int32_t),
                                          cudaMemcpyDeviceToHost, ctx->stream));
        DG_CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    }
    b->max_graph_nnz = 0;
    for (int g = 0; g < n_graphs; ++g) {
        DG_REQUIRE(b->h_graph_e[g + 1] >= b->h_graph_e[g], DG_ERR_INVALID, "row_ptr decreases at graph %d", g);
        b->max_graph_nnz = std::max(b->max_graph_nnz, b->h_graph_e[g + 1] - b->h_graph_e[g]);
    }
upload:
    if (mem == DG_MEM_HOST) {
        if (b->cap_graphs < (size_t)n_graphs + 1 || !b->graph_ptr) {
            if (b->graph_ptr) cudaFree(b->graph_ptr);
            b->graph_ptr = nullptr;
            b->cap_graphs = (size_t)n_graphs + 1 + (size_t)n_graphs / 4;
            DG_CUDA_CHECK(cudaMalloc((void **)&b->graph_ptr, sizeof(int32_t) * b->cap_graphs));
        }
        const bool grow_nodes = b->cap_nodes < (size_t)n_nodes || !b->row_ptr;
        if (grow_nodes) {
            if (b->row_ptr) cudaFree(b->row_ptr);
            b->row_ptr = nullptr;
            size_t cap = (size_t)n_nodes + (size_t)n_nodes / 4;
            DG_CUDA_CHECK(cudaMalloc((void **)&b->row_ptr, sizeof(int32_t) * (cap + 1)));
            DG_TRY(batch_alloc_aux(b, cap));  // sets cap_nodes
        }
        const size_t need_nnz = upper ? 2 * (size_t)nnz : (size_t)nnz;  // the expansion of the upper format doubles it
        if (b->cap_nnz < need_nnz || !b->col_idx) {
            if (b->col_idx) cudaFree(b->col_idx);
            if (b->col16) cudaFree(b->col16);
            b->col_idx = nullptr;
            b->col16 = nullptr;
            b->cap_nnz = need_nnz + need_nnz / 4 + 1;
            DG_CUDA_CHECK(cudaMalloc((void **)&b->col_idx, sizeof(int32_t) * b->cap_nnz));
        }
        if (col_local16 && !b->col16) DG_CUDA_CHECK(cudaMalloc((void **)&b->col16, sizeof(uint16_t) * b->cap_nnz));
        if (upper && (!b->row_ptr_u || b->cap_row_ptr_u < b->cap_nodes + 1)) {
            if (b->row_ptr_u) cudaFree(b->row_ptr_u);
            b->row_ptr_u = nullptr;
            b->cap_row_ptr_u = 0;
            DG_CUDA_CHECK(cudaMalloc((void **)&b->row_ptr_u, sizeof(int32_t) * (b->cap_nodes + 1)));
            b->cap_row_ptr_u = b->cap_nodes + 1;
        }
        b->owns_csr = true;
        DG_CUDA_CHECK(cudaMemcpyAsync(b->graph_ptr, graph_ptr, sizeof(int32_t) * ((size_t)n_graphs + 1),
                                      cudaMemcpyHostToDevice, ctx->stream));
        DG_CUDA_CHECK(cudaMemcpyAsync(upper ? b->row_ptr_u : b->row_ptr, row_ptr, sizeof(int32_t) * ((size_t)n_nodes + 1),
                                      cudaMemcpyHostToDevice, ctx->stream));
        if (upper) {
            DG_REQUIRE(b->max_graph_nodes <= kSymMaxNodes, DG_ERR_UNSUPPORTED,
                       "the upper-triangle host format takes graphs of at most %d vertices", kSymMaxNodes);
            b->upper_pending = true;
            b->dinv_deferred = compute_dinv;
            compute_dinv = false;   // the degrees need the full rows: computed when (and if) they are expanded
        }
        if (nnz && col_local16) {
            DG_REQUIRE(b->max_graph_nodes <= 65536, DG_ERR_INVALID, "16-bit column ids need graphs of at most 65536 vertices");
            DG_CUDA_CHECK(cudaMemcpyAsync(b->col16, col_local16, sizeof(uint16_t) * (size_t)nnz, cudaMemcpyHostToDevice,
                                          ctx->stream));
            b->cols_pending = true;  // the tensor-core kernel reads the 16-bit ids as they are; others expand them first
        } else if (nnz) {
            DG_CUDA_CHECK(cudaMemcpyAsync(b->col_idx, col_idx, sizeof(int32_t) * (size_t)nnz,
                                          cudaMemcpyHostToDevice, ctx->stream));
        }
    } else {
        b->owns_csr = false;
        b->graph_ptr = const_cast<int32_t *>(graph_ptr);
        b->row_ptr = const_cast<int32_t *>(row_ptr);
        b->col_idx = const_cast<int32_t *>(col_idx);
        DG_TRY(batch_alloc_aux(b, (size_t)n_nodes));
    }
    return compute_dinv ? batch_compute_dinv(b) : DG_OK;
}

int dg_batch_create(dg_context *ctx, int32_t n_graphs, int32_t n_nodes, int32_t nnz, const int32_t *graph_ptr,
                    const int32_t *row_ptr, const int32_t *col_idx, int mem, dg_batch **out) {
    clear_error();
    DG_TRY(check_ctx(ctx));
    DG_REQUIRE(out != nullptr, DG_ERR_INVALID, "null out pointer");
    *out = nullptr;
    DG_REQUIRE(mem == DG_MEM_HOST || mem == DG_MEM_DEVICE, DG_ERR_INVALID, "bad memory space %d", mem);
    DeviceGuard guard(ctx->device);
    dg_batch *b = new (std::nothrow) dg_batch();
    DG_REQUIRE(b != nullptr, DG_ERR_INVALID, "out of host memory");
    b->ctx = ctx;
    int s = batch_fill(b, n_graphs, n_nodes, nnz, graph_ptr, row_ptr, col_idx, mem);
    if (s == DG_OK && mem == DG_MEM_HOST) s = finish(ctx);  // the caller may free its arrays on return
    if (s != DG_OK) {
        dg_batch_destroy(b);
        return s;
    }
    *out = b;
    return DG_OK;
}

void dg_batch_destroy(dg_batch *b) {
    if (!b) return;
    DeviceGuard guard(b->ctx->device);
    cudaStreamSynchronize(b->ctx->stream);
    if (b->owns_csr) {
        if (b->graph_ptr) cudaFree(b->graph_ptr);
        if (b->row_ptr) cudaFree(b->row_ptr);
        if (b->col_idx) cudaFree(b->col_idx);
    }
    if (b->col16) cudaFree(b->col16);
    if (b->dinv) cudaFree(b->dinv);
    if (b->keep) cudaFree(b->keep);
    if (b->x0) cudaFree(b->x0);
    if (b->tiles_dev) cudaFree(b->tiles_dev);
    if (b->tc_tiles_dev) cudaFree(b->tc_tiles_dev);
    if (b->gs_tiles_dev) cudaFree(b->gs_tiles_dev);
    if (b->row_ptr_u) cudaFree(b->row_ptr_u);
    delete b;
}

int dg_batch_set_keep(dg_batch *b, const uint8_t *keep, int mem) {
    clear_error();
    DG_REQUIRE(b != nullptr, DG_ERR_INVALID, "null batch");
    dg_context *ctx = b->ctx;
    DeviceGuard guard(ctx->device);
    if (!keep) {
        if (b->keep) {
            DG_CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
            cudaFree(b->keep);
            b->keep = nullptr;
        }
    } else {
        if (!b->keep) DG_CUDA_CHECK(cudaMalloc((void **)&b->keep, std::max<size_t>(b->cap_nodes, 1)));
        DG_CUDA_CHECK(cudaMemcpyAsync(b->keep, keep, (size_t)b->n_nodes,
                                      mem == DG_MEM_HOST ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice,
                                      ctx->stream));
    }
    DG_TRY(batch_compute_dinv(b));
    return mem == DG_MEM_HOST ? finish(ctx) : DG_OK;
}

static int set_keep_from_weights_device(dg_batch *b, const double *d_wts) {
    if (!b->keep) DG_CUDA_CHECK(cudaMalloc((void **)&b->keep, std::max<size_t>(b->cap_nodes, 1)));
    DG_TRY(keep_from_weights_device(b->ctx, b->n_nodes, d_wts, b->keep));
    return batch_compute_dinv(b);
}

int dg_batch_set_keep_from_weights(dg_batch *b, const double *wts, int mem) {
    clear_error();
    DG_REQUIRE(b != nullptr && wts != nullptr, DG_ERR_INVALID, "null argument");
    dg_context *ctx = b->ctx;
    DeviceGuard guard(ctx->device);
    const double *d_wts = wts;
    if (mem == DG_MEM_HOST) {
        double *tmp = nullptr;
        DG_TRY(stage_in(ctx, kSlotWts, wts, (size_t)b->n_nodes, &tmp));
        d_wts = tmp;
    }
    DG_TRY(set_keep_from_weights_device(b, d_wts));
    return mem == DG_MEM_HOST ? finish(ctx) : DG_OK;
}

int dg_batch_set_x0(dg_batch *b, const float *x0, int mem) {
    clear_error();
    DG_REQUIRE(b != nullptr, DG_ERR_INVALID, "null batch");
    dg_context *ctx = b->ctx;
    DeviceGuard guard(ctx->device);
    if (!x0) {
        if (b->x0) {
            DG_CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
            cudaFree(b->x0);
            b->x0 = nullptr;
        }
        return DG_OK;
    }
    if (!b->x0) DG_CUDA_CHECK(cudaMalloc((void **)&b->x0, sizeof(float) * std::max<size_t>(b->cap_nodes, 1)));
    DG_CUDA_CHECK(cudaMemcpyAsync(b->x0, x0, sizeof(float) * (size_t)b->n_nodes,
                                  mem == DG_MEM_HOST ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice,
                                  ctx->stream));
    return mem == DG_MEM_HOST ? finish(ctx) : DG_OK;
}

// -------------------------------------------------------------------------------------------------
int dg_graph_convolution(dg_context *ctx, dg_batch *b, int32_t c_in, int32_t c_out, const float *w0,
                         const float *w1, const float *bias, int act, float leaky_alpha, const float *x, float *y,
                         int mem) {
    clear_error();
    DG_TRY(check_ctx(ctx));
    DG_REQUIRE(b && w0 && w1 && x && y, DG_ERR_INVALID, "null argument");
    DG_REQUIRE(c_in >= 1 && c_out >= 1, DG_ERR_INVALID, "empty layer shape");
    DG_REQUIRE(c_in <= kMaxWidth && c_out <= kMaxWidth, DG_ERR_UNSUPPORTED, "layer width %d -> %d exceeds %d", c_in,
               c_out, kMaxWidth);
    DG_REQUIRE(act >= DG_ACT_IDENTITY && act <= DG_ACT_RELU, DG_ERR_INVALID, "unknown activation %d", act);
    DeviceGuard guard(ctx->device);
    dg_layer_dev L;
    L.c_in = c_in;
    L.c_out = c_out;
    L.cpi = pad_width(c_in);
    L.cpo = pad_width(c_out);
    L.act = act;
    std::vector<float> wc((size_t)2 * L.cpi * L.cpo + L.cpo, 0.f);
    for (int k = 0; k < c_in; ++k)
        for (int c = 0; c < c_out; ++c) {
            wc[(size_t)k * L.cpo + c] = w0[(size_t)k * c_out + c];
            wc[(size_t)(L.cpi + k) * L.cpo + c] = w1[(size_t)k * c_out + c];
        }
    if (bias)
        for (int c = 0; c < c_out; ++c) wc[(size_t)2 * L.cpi * L.cpo + c] = bias[c];
    float *dw = nullptr;
    DG_TRY(scratch_as(ctx, kSlotLayerW, wc.size(), &dw));
    // synchronous copy: wc is a stack-lifetime staging vector
    DG_CUDA_CHECK(cudaMemcpyAsync(dw, wc.data(), sizeof(float) * wc.size(), cudaMemcpyHostToDevice, ctx->stream));
    DG_CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    L.wcat = dw;
    L.bias = dw + (size_t)2 * L.cpi * L.cpo;
    const size_t n = (size_t)b->n_nodes;
    const float *dx = x;
    float *dy = y;
    if (mem == DG_MEM_HOST) {
        float *tx = nullptr;
        DG_TRY(stage_in(ctx, kSlotStageIn0, x, n * c_in, &tx));
        dx = tx;
        DG_TRY(scratch_as(ctx, kSlotStageOut0, n * c_out, &dy));
    }
    DG_TRY(graph_convolution_device(ctx, b, L, leaky_alpha, dx, c_in, dy, c_out));
    if (mem == DG_MEM_HOST) {
        DG_TRY(copy_out(ctx, y, dy, n * c_out));
        return finish(ctx);
    }
    return DG_OK;
}

int dg_spmm_laplacian(dg_context *ctx, dg_batch *b, int32_t width, const float *z, float *y, int mem) {
    clear_error();
    DG_TRY(check_ctx(ctx));
    DG_REQUIRE(b && z && y, DG_ERR_INVALID, "null argument");
    DG_REQUIRE(width >= 1 && width <= 1024, DG_ERR_INVALID, "bad width %d", width);
    DeviceGuard guard(ctx->device);
    DG_TRY(batch_ensure_cols(b));
    const size_t count = (size_t)b->n_nodes * width;
    if (mem == DG_MEM_DEVICE) return spmm_laplacian_device(ctx, b, width, z, y);
    float *dz = nullptr, *dy = nullptr;
    DG_TRY(stage_in(ctx, kSlotStageIn0, z, count, &dz));
    DG_TRY(scratch_as(ctx, kSlotStageOut0, count, &dy));
    DG_TRY(spmm_laplacian_device(ctx, b, width, dz, dy));
    DG_TRY(copy_out(ctx, y, dy, count));
    return finish(ctx);
}

int dg_gcn_forward(dg_context *ctx, const dg_model *m, dg_batch *b, float *out, int mem) {
    clear_error();
    DG_TRY(check_ctx(ctx));
    DG_REQUIRE(m && b && out, DG_ERR_INVALID, "null argument");
    DeviceGuard guard(ctx->device);
    const size_t count = (size_t)b->n_nodes * m->layers.back().c_out;
    float *d_out = out;
    if (mem == DG_MEM_HOST) DG_TRY(scratch_as(ctx, kSlotScore, count, &d_out));
    DG_TRY(gcn_forward_device(ctx, m, b, d_out, nullptr, DG_PREDICT_MIS, nullptr));
    if (mem == DG_MEM_HOST) {
        DG_TRY(copy_out(ctx, out, d_out, count));
        return finish(ctx);
    }
    return DG_OK;
}

int dg_utility(dg_context *ctx, const dg_batch *b, const float *score, int32_t score_stride, const double *wts,
               int predict, double *util, int mem) {
    clear_error();
    DG_TRY(check_ctx(ctx));
    DG_REQUIRE(b && score && util, DG_ERR_INVALID, "null argument");
    DG_REQUIRE(predict == DG_PREDICT_MIS || wts != nullptr, DG_ERR_INVALID, "weights required for predict=mwis");
    DG_REQUIRE(score_stride >= 1, DG_ERR_INVALID, "score_stride must be >= 1");
    DeviceGuard guard(ctx->device);
    const size_t n = (size_t)b->n_nodes;
    const float *d_score = score;
    const double *d_wts = wts;
    double *d_util = util;
    if (mem == DG_MEM_HOST) {
        float *ts = nullptr;
        DG_TRY(stage_in(ctx, kSlotStageIn0, score, n * score_stride, &ts));
        d_score = ts;
        if (wts) {
            double *tw = nullptr;
            DG_TRY(stage_in(ctx, kSlotWts, wts, n, &tw));
            d_wts = tw;
        }
        DG_TRY(scratch_as(ctx, kSlotUtil, n, &d_util));
    }
    DG_TRY(utility_device(ctx, (int)n, d_score, score_stride, d_wts, predict, d_util));
    if (mem == DG_MEM_HOST) {
        DG_TRY(copy_out(ctx, util, d_util, n));
        return finish(ctx);
    }
    return DG_OK;
}

int dg_lgs(dg_context *ctx, const dg_batch *b, const double *util, int32_t nstep, uint8_t *member, uint8_t *nb_is,
           int32_t *steps, int64_t *p2p, int64_t *bst, double *oh_vec, int mem) {
    clear_error();
    DG_TRY(check_ctx(ctx));
    DG_REQUIRE(b && util && member, DG_ERR_INVALID, "null argument");
    DeviceGuard guard(ctx->device);
    const size_t n = (size_t)b->n_nodes, G = (size_t)b->n_graphs;
    if (mem == DG_MEM_DEVICE) return lgs_device(ctx, b, util, nstep, member, nb_is, steps, p2p, bst, oh_vec);
    double *d_util = nullptr;
    uint8_t *d_member = nullptr, *d_nbis = nullptr;
    int32_t *d_steps = nullptr;
    int64_t *d_p2p = nullptr, *d_bst = nullptr;
    double *d_oh = nullptr;
    DG_TRY(stage_in(ctx, kSlotUtil, util, n, &d_util));
    DG_TRY(scratch_as(ctx, kSlotMember, n, &d_member));
    if (nb_is) DG_TRY(scratch_as(ctx, kSlotNbis, n, &d_nbis));
    if (steps) DG_TRY(scratch_as(ctx, kSlotSteps, G, &d_steps));
    if (p2p) DG_TRY(scratch_as(ctx, kSlotP2p, G, &d_p2p));
    if (bst) DG_TRY(scratch_as(ctx, kSlotBst, G, &d_bst));
    if (oh_vec) DG_TRY(scratch_as(ctx, kSlotOh, n, &d_oh));
    DG_TRY(lgs_device(ctx, b, d_util, nstep, d_member, d_nbis, d_steps, d_p2p, d_bst, d_oh));
    DG_TRY(copy_out(ctx, member, d_member, n));
    DG_TRY(copy_out(ctx, nb_is, d_nbis, n));
    DG_TRY(copy_out(ctx, steps, d_steps, G));
    DG_TRY(copy_out(ctx, p2p, d_p2p, G));
    DG_TRY(copy_out(ctx, bst, d_bst, G));
    DG_TRY(copy_out(ctx, oh_vec, d_oh, n));
    return finish(ctx);
}

int dg_dist_greedy(dg_context *ctx, const dg_batch *b, const double *wts, double epsilon, uint8_t *member,
                   int32_t *steps, int mem) {
    clear_error();
    DG_TRY(check_ctx(ctx));
    DG_REQUIRE(b && wts && member, DG_ERR_INVALID, "null argument");
    DG_REQUIRE(epsilon > 0.0, DG_ERR_INVALID, "epsilon must be positive (heuristics.py:43)");
    DeviceGuard guard(ctx->device);
    const double alpha = 1.0 + (epsilon / 3.0);  // heuristics.py:46
    const size_t n = (size_t)b->n_nodes, G = (size_t)b->n_graphs;
    if (mem == DG_MEM_DEVICE) return dist_greedy_device(ctx, b, wts, alpha, member, steps);
    double *d_wts = nullptr;
    uint8_t *d_member = nullptr;
    int32_t *d_steps = nullptr;
    DG_TRY(stage_in(ctx, kSlotUtil, wts, n, &d_wts));
    DG_TRY(scratch_as(ctx, kSlotMember, n, &d_member));
    if (steps) DG_TRY(scratch_as(ctx, kSlotSteps, G, &d_steps));
    DG_TRY(dist_greedy_device(ctx, b, d_wts, alpha, d_member, d_steps));
    DG_TRY(copy_out(ctx, member, d_member, n));
    DG_TRY(copy_out(ctx, steps, d_steps, G));
    return finish(ctx);
}

int dg_member_weight(dg_context *ctx, const dg_batch *b, const uint8_t *member, const double *wts, double *total,
                     int mem) {
    clear_error();
    DG_TRY(check_ctx(ctx));
    DG_REQUIRE(b && member && wts && total, DG_ERR_INVALID, "null argument");
    DeviceGuard guard(ctx->device);
    if (mem == DG_MEM_DEVICE) return member_weight_device(ctx, b, member, wts, total);
    uint8_t *d_member = nullptr;
    double *d_wts = nullptr, *d_total = nullptr;
    DG_TRY(stage_in(ctx, kSlotMember, member, (size_t)b->n_nodes, &d_member));
    DG_TRY(stage_in(ctx, kSlotWts, wts, (size_t)b->n_nodes, &d_wts));
    DG_TRY(scratch_as(ctx, kSlotTotal, (size_t)b->n_graphs, &d_total));
    DG_TRY(member_weight_device(ctx, b, d_member, d_wts, d_total));
    DG_TRY(copy_out(ctx, total, d_total, (size_t)b->n_graphs));
    return finish(ctx);
}

// device-space body shared by dg_solve and dg_solve_host
static int solve_device(dg_context *ctx, const dg_model *m, dg_batch *b, const double *d_wts, int predict,
                        int remove_zero_weight, uint8_t *d_member, float *d_score, double *d_util, double *d_total,
                        int32_t *d_steps) {
    const size_t n = (size_t)b->n_nodes;
    // small graphs: everything in one graph-resident kernel (dg_fused.cu)
    bool handled = false;
    // ... on the tensor cores when the model has 32-wide hidden layers and every graph fits (dg_tc.cu)
    DG_TRY(tc_try_solve(ctx, m, b, d_wts, predict, remove_zero_weight, d_member, d_score, d_util, d_total, d_steps,
                        &handled));
    if (handled) return DG_OK;
    DG_TRY(batch_ensure_cols(b));
    DG_TRY(fused_try_solve(ctx, m, b, d_wts, predict, remove_zero_weight, d_member, d_score, d_util, d_total, d_steps,
                           &handled));
    if (handled) return DG_OK;
    b->tc_ran_partial = false;
    // Per-layer path.  Zero-weight removal (mwis_dqn_call.py:202-207) is a keep mask for THIS solve only: like the
    // graph-resident kernels it replaces the caller's mask for the duration of the call and leaves the batch (mask and
    // degrees) as it was found, so a later dg_lgs / dg_gcn_forward on the same batch sees the caller's graph.
    uint8_t *const caller_keep = b->keep;
    if (remove_zero_weight) {
        uint8_t *tmp = nullptr;
        DG_TRY(scratch_as(ctx, kSlotSolveKeep, std::max<size_t>(n, 1), &tmp));
        DG_TRY(keep_from_weights_device(ctx, b->n_nodes, d_wts, tmp));
        b->keep = tmp;
        const int st0 = batch_compute_dinv(b);
        if (st0 != DG_OK) {
            b->keep = caller_keep;
            return st0;
        }
    }
    int st = DG_OK;
    do {
        const int d_out = m->layers.back().c_out;
        if (!d_score && (st = scratch_as(ctx, kSlotScore, n * d_out, &d_score)) != DG_OK) break;
        if (!d_util && (st = scratch_as(ctx, kSlotUtil, n, &d_util)) != DG_OK) break;
        // with several output columns the utility uses column 0, as act_vals.flatten() does for diver_num == 1
        if ((st = gcn_forward_device(ctx, m, b, d_score, d_wts, predict, d_util, /*try_resident=*/false)) != DG_OK) break;
        if ((st = lgs_device(ctx, b, d_util, -1, d_member, nullptr, d_steps, nullptr, nullptr, nullptr)) != DG_OK) break;
        if (d_total) st = member_weight_device(ctx, b, d_member, d_wts, d_total);
    } while (false);
    if (remove_zero_weight) {
        b->keep = caller_keep;
        const int st2 = batch_compute_dinv(b);
        if (st == DG_OK) st = st2;
    }
    return st;
}

// GCN embedded into the greedy iteration (mwis_gdpg_call.py:278-318), device-space body
static int solve_dit_device(dg_context *ctx, const dg_model *m, dg_batch *b, const double *d_wts, int predict,
                            uint8_t *d_member, double *d_total, int32_t *d_steps) {
    const size_t n = (size_t)b->n_nodes, G = (size_t)b->n_graphs;
    if (n == 0 || G == 0) return DG_OK;
    bool handled = false;
    // small graphs: the whole iteration inside one graph-resident launch, on the tensor cores when the model allows
    DG_TRY(tc_try_solve(ctx, m, b, d_wts, predict, 0, d_member, nullptr, nullptr, d_total, d_steps, &handled, true));
    if (handled) return DG_OK;
    DG_TRY(batch_ensure_cols(b));
    DG_TRY(fused_try_solve(ctx, m, b, d_wts, predict, 0, d_member, nullptr, nullptr, d_total, d_steps, &handled, true));
    if (handled) return DG_OK;
    // generic path: a handful of launches per iteration over the per-layer kernels
    const int d_out = m->layers.back().c_out;
    float *d_score = nullptr;
    double *d_util = nullptr;
    uint8_t *d_joined = nullptr, *d_nbis = nullptr, *saved_keep = nullptr;
    int *d_flag = nullptr;
    DG_TRY(scratch_as(ctx, kSlotScore, n * d_out, &d_score));
    DG_TRY(scratch_as(ctx, kSlotUtil, n, &d_util));
    DG_TRY(scratch_as(ctx, kSlotStageIn1, n, &d_joined));
    DG_TRY(scratch_as(ctx, kSlotStageIn2, n, &d_nbis));
    DG_TRY(scratch_as(ctx, kSlotStageOut0, G + 1, &d_flag));
    int *d_any = d_flag + G;
    const bool had_keep = b->keep != nullptr;
    if (had_keep) {  // the caller's mask is the initial residual graph; put it back afterwards
        DG_TRY(scratch_as(ctx, kSlotStageOut1, n, &saved_keep));
        DG_CUDA_CHECK(cudaMemcpyAsync(saved_keep, b->keep, n, cudaMemcpyDeviceToDevice, ctx->stream));
    } else {
        DG_CUDA_CHECK(cudaMalloc((void **)&b->keep, std::max<size_t>(b->cap_nodes, 1)));
        DG_CUDA_CHECK(cudaMemsetAsync(b->keep, 1, n, ctx->stream));
    }
    DG_CUDA_CHECK(cudaMemsetAsync(d_member, 0, n, ctx->stream));
    if (d_steps) DG_CUDA_CHECK(cudaMemsetAsync(d_steps, 0, sizeof(int32_t) * G, ctx->stream));
    int st = DG_OK;
    for (int it = 0; st == DG_OK; ++it) {
        if (it >= kLgsRoundCap) {
            set_error("iterative solve hit the round cap (NaN utilities or self-loops?)");
            st = DG_ERR_NOT_CONVERGED;
            break;
        }
        if ((st = dit_filter_device(ctx, b, d_wts, b->keep, d_flag, d_any)) != DG_OK) break;
        if (cudaMemcpyAsync(ctx->h_flag + 1, d_any, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess ||
            cudaStreamSynchronize(ctx->stream) != cudaSuccess) {
            set_error("CUDA error in the iterative solve: %s", cudaGetErrorString(cudaGetLastError()));
            st = DG_ERR_CUDA;
            break;
        }
        if (ctx->h_flag[1] == 0) break;
        if ((st = dit_steps_device(ctx, (int)G, d_flag, d_steps)) != DG_OK) break;
        if ((st = batch_compute_dinv(b)) != DG_OK) break;
        if ((st = gcn_forward_device(ctx, m, b, d_score, d_wts, predict, d_util)) != DG_OK) break;
        if ((st = lgs_device(ctx, b, d_util, 1, d_joined, d_nbis, nullptr, nullptr, nullptr, nullptr)) != DG_OK) break;
        st = dit_update_device(ctx, (int)n, d_joined, d_nbis, d_member, b->keep);
    }
    // leave the batch as it was found
    if (had_keep) {
        cudaMemcpyAsync(b->keep, saved_keep, n, cudaMemcpyDeviceToDevice, ctx->stream);
    } else {
        cudaStreamSynchronize(ctx->stream);
        cudaFree(b->keep);
        b->keep = nullptr;
    }
    int st2 = batch_compute_dinv(b);
    if (st != DG_OK) return st;
    DG_TRY(st2);
    if (d_total) DG_TRY(member_weight_device(ctx, b, d_member, d_wts, d_total));
    return DG_OK;
}

int dg_solve_dit(dg_context *ctx, const dg_model *m, dg_batch *b, const double *wts, int predict, uint8_t *member,
                 double *total, int32_t *steps, int mem) {
    clear_error();
    DG_TRY(check_ctx(ctx));
    DG_REQUIRE(m && b && wts && member, DG_ERR_INVALID, "null argument");
    DG_REQUIRE(predict == DG_PREDICT_MWIS || predict == DG_PREDICT_MIS, DG_ERR_INVALID, "unknown predict mode");
    DG_REQUIRE(m->layers.back().c_out == 1 && m->head == DG_HEAD_LINEAR, DG_ERR_UNSUPPORTED,
               "the iterative solve needs a one-column linear head");
    DeviceGuard guard(ctx->device);
    if (mem == DG_MEM_DEVICE) return solve_dit_device(ctx, m, b, wts, predict, member, total, steps);
    const size_t n = (size_t)b->n_nodes, G = (size_t)b->n_graphs;
    double *d_wts = nullptr, *d_total = nullptr;
    uint8_t *d_member = nullptr;
    int32_t *d_steps = nullptr;
    DG_TRY(stage_in(ctx, kSlotWts, wts, n, &d_wts));
    DG_TRY(scratch_as(ctx, kSlotMember, n, &d_member));
    if (total) DG_TRY(scratch_as(ctx, kSlotTotal, G, &d_total));
    if (steps) DG_TRY(scratch_as(ctx, kSlotSteps, G, &d_steps));
    DG_TRY(solve_dit_device(ctx, m, b, d_wts, predict, d_member, d_total, d_steps));
    DG_TRY(copy_out(ctx, member, d_member, n));
    DG_TRY(copy_out(ctx, total, d_total, G));
    DG_TRY(copy_out(ctx, steps, d_steps, G));
    return finish(ctx);
}

int dg_solve(dg_context *ctx, const dg_model *m, dg_batch *b, const double *wts, int predict,
             int remove_zero_weight, uint8_t *member, float *score, double *util, double *total, int32_t *steps,
             int mem) {
    clear_error();
    DG_TRY(check_ctx(ctx));
    DG_REQUIRE(m && b && wts && member, DG_ERR_INVALID, "null argument");
    DG_REQUIRE(predict == DG_PREDICT_MWIS || predict == DG_PREDICT_MIS, DG_ERR_INVALID, "unknown predict mode");
    DeviceGuard guard(ctx->device);
    if (mem == DG_MEM_DEVICE)
        return solve_device(ctx, m, b, wts, predict, remove_zero_weight, member, score, util, total, steps);
    const size_t n = (size_t)b->n_nodes, G = (size_t)b->n_graphs;
    const int d_out = m->layers.back().c_out;
    double *d_wts = nullptr, *d_util = nullptr, *d_total = nullptr;
    uint8_t *d_member = nullptr;
    float *d_score = nullptr;
    int32_t *d_steps = nullptr;
    DG_TRY(stage_in(ctx, kSlotWts, wts, n, &d_wts));
    DG_TRY(scratch_as(ctx, kSlotMember, n, &d_member));
    DG_TRY(scratch_as(ctx, kSlotScore, n * d_out, &d_score));
    DG_TRY(scratch_as(ctx, kSlotUtil, n, &d_util));
    if (total) DG_TRY(scratch_as(ctx, kSlotTotal, G, &d_total));
    if (steps) DG_TRY(scratch_as(ctx, kSlotSteps, G, &d_steps));
    DG_TRY(solve_device(ctx, m, b, d_wts, predict, remove_zero_weight, d_member, d_score, d_util, d_total, d_steps));
    DG_TRY(copy_out(ctx, member, d_member, n));
    DG_TRY(copy_out(ctx, score, d_score, n * d_out));
    DG_TRY(copy_out(ctx, util, d_util, n));
    DG_TRY(copy_out(ctx, total, d_total, G));
    DG_TRY(copy_out(ctx, steps, d_steps, G));
    return finish(ctx);
}

static int solve_host_impl(dg_context *ctx, const dg_model *m, int32_t n_graphs, int32_t n_nodes, int32_t nnz,
                           const int32_t *graph_ptr, const int32_t *row_ptr, const int32_t *col_idx,
                           const double *wts, int predict, int remove_zero_weight, uint8_t *member, double *total,
                           bool wait, const uint16_t *col_local16 = nullptr, cudaEvent_t copied = nullptr,
                           bool upper = false) {
    clear_error();
    DG_TRY(check_ctx(ctx));
    DG_REQUIRE(m && wts && member, DG_ERR_INVALID, "null argument");
    DeviceGuard guard(ctx->device);
    if (!ctx->host_batch) {
        ctx->host_batch = new (std::nothrow) dg_batch();
        DG_REQUIRE(ctx->host_batch != nullptr, DG_ERR_INVALID, "out of host memory");
        ctx->host_batch->ctx = ctx;
    }
    dg_batch *b = ctx->host_batch;
    if (b->keep && !remove_zero_weight) {
        DG_CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
        cudaFree(b->keep);
        b->keep = nullptr;
    }
    if (b->keep && b->cap_nodes < (size_t)n_nodes) {  // will be re-allocated at the new capacity
        DG_CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
        cudaFree(b->keep);
        b->keep = nullptr;
    }
    // with zero-weight removal the degrees are computed once the keep mask is known (solve_device)
    const bool timing = ctx->env.ingest_timing;
    const auto t0 = std::chrono::steady_clock::now();
    DG_TRY(batch_fill(b, n_graphs, n_nodes, nnz, graph_ptr, row_ptr, col_idx, DG_MEM_HOST, !remove_zero_weight, col_local16,
                      upper));
    const size_t n = (size_t)n_nodes, G = (size_t)n_graphs;
    double *d_wts = nullptr, *d_total = nullptr;
    uint8_t *d_member = nullptr;
    DG_TRY(stage_in(ctx, kSlotWts, wts, n, &d_wts));
    if (copied) DG_CUDA_CHECK(cudaEventRecord(copied, ctx->stream));
    const auto t1 = std::chrono::steady_clock::now();
    DG_TRY(scratch_as(ctx, kSlotMember, n, &d_member));
    if (total) DG_TRY(scratch_as(ctx, kSlotTotal, G, &d_total));
    DG_TRY(solve_device(ctx, m, b, d_wts, predict, remove_zero_weight, d_member, nullptr, nullptr, d_total, nullptr));
    const auto t2 = std::chrono::steady_clock::now();
    DG_TRY(copy_out(ctx, member, d_member, n));
    DG_TRY(copy_out(ctx, total, d_total, G));
    if (timing) {
        const auto t3 = std::chrono::steady_clock::now();
        auto us = [](std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point c) {
            return std::chrono::duration<double, std::micro>(c - a).count();
        };
        fprintf(stderr, "[solve_host] batch_fill + weights %.0f us, plan + launch %.0f us, copy-out enqueue %.0f us\n", us(t0, t1),
                us(t1, t2), us(t2, t3));
    }
    return wait ? finish(ctx) : DG_OK;
}

extern "C++" {
namespace dg {
int host_batch_set_meta(dg_context *ctx, int32_t n_graphs, const int64_t *v0, const int64_t *e0, dg_batch **out) {
    if (!ctx->host_batch) {
        ctx->host_batch = new (std::nothrow) dg_batch();
        DG_REQUIRE(ctx->host_batch != nullptr, DG_ERR_INVALID, "out of host memory");
        ctx->host_batch->ctx = ctx;
    }
    dg_batch *b = ctx->host_batch;
    b->n_graphs = n_graphs;
    b->n_nodes = (int)v0[n_graphs];
    b->nnz = (int)e0[n_graphs];
    b->h_graph_ptr.resize((size_t)n_graphs + 1);
    b->h_graph_e.resize((size_t)n_graphs + 1);
    b->max_graph_nodes = 0;
    b->max_graph_nnz = 0;
    for (int g = 0; g <= n_graphs; ++g) {
        b->h_graph_ptr[(size_t)g] = (int32_t)v0[g];
        b->h_graph_e[(size_t)g] = (int32_t)e0[g];
        if (g) {
            b->max_graph_nodes = std::max(b->max_graph_nodes, (int)(v0[g] - v0[g - 1]));
            b->max_graph_nnz = std::max(b->max_graph_nnz, (int)(e0[g] - e0[g - 1]));
        }
    }
    b->tiles_valid = false;
    b->tc_tiles_valid = false;
    b->gs_valid = false;
    b->tc_plan_ready = false;
    b->meta_ready = true;
    *out = b;
    return DG_OK;
}

int solve_host_staged(dg_context *ctx, const dg_model *m, int32_t n_graphs, int32_t n_nodes, int32_t nnz,
                      const int32_t *graph_ptr, const int32_t *row_ptr, const int32_t *col_idx, const double *wts,
                      int predict, int remove_zero_weight, uint8_t *member, double *total, bool wait,
                      const uint16_t *col_local16, cudaEvent_t copied, bool upper) {
    return solve_host_impl(ctx, m, n_graphs, n_nodes, nnz, graph_ptr, row_ptr, col_idx, wts, predict, remove_zero_weight,
                           member, total, wait, col_local16, copied, upper);
}
}  // namespace dg
}  // extern "C++"

int dg_solve_host(dg_context *ctx, const dg_model *m, int32_t n_graphs, int32_t n_nodes, int32_t nnz,
                  const int32_t *graph_ptr, const int32_t *row_ptr, const int32_t *col_idx, const double *wts,
                  int predict, int remove_zero_weight, uint8_t *member, double *total) {
    return solve_host_impl(ctx, m, n_graphs, n_nodes, nnz, graph_ptr, row_ptr, col_idx, wts, predict,
                           remove_zero_weight, member, total, true);
}

int dg_solve_host_async(dg_context *ctx, const dg_model *m, int32_t n_graphs, int32_t n_nodes, int32_t nnz,
                        const int32_t *graph_ptr, const int32_t *row_ptr, const int32_t *col_idx, const double *wts,
                        int predict, int remove_zero_weight, uint8_t *member, double *total) {
    return solve_host_impl(ctx, m, n_graphs, n_nodes, nnz, graph_ptr, row_ptr, col_idx, wts, predict,
                           remove_zero_weight, member, total, false);
}

int dg_solve_host_compact(dg_context *ctx, const dg_model *m, int32_t n_graphs, int32_t n_nodes, int32_t nnz,
                          const int32_t *graph_ptr, const int32_t *row_ptr, const uint16_t *col_local, const double *wts,
                          int predict, int remove_zero_weight, uint8_t *member, double *total, int wait) {
    DG_REQUIRE(col_local || nnz == 0, DG_ERR_INVALID, "null column array");
    return solve_host_impl(ctx, m, n_graphs, n_nodes, nnz, graph_ptr, row_ptr, nullptr, wts, predict, remove_zero_weight,
                           member, total, wait != 0, col_local);
}

int dg_solve_host_upper(dg_context *ctx, const dg_model *m, int32_t n_graphs, int32_t n_nodes, int32_t nnz_upper,
                        const int32_t *graph_ptr, const int32_t *row_ptr_upper, const uint16_t *col_local_upper,
                        const double *wts, int predict, int remove_zero_weight, uint8_t *member, double *total, int wait) {
    DG_REQUIRE(col_local_upper || nnz_upper == 0, DG_ERR_INVALID, "null column array");
    return solve_host_impl(ctx, m, n_graphs, n_nodes, nnz_upper, graph_ptr, row_ptr_upper, nullptr, wts, predict,
                           remove_zero_weight, member, total, wait != 0, col_local_upper, nullptr, true);
}

// -------------------------------------------------------------------------------------------------
// row-partitioned giant graph
int dg_part_create(dg_context *ctx, int32_t n_global, int32_t row0, int32_t n_local, int32_t nnz_local,
                   const int32_t *row_ptr_local, const int32_t *col_idx_global, int mem, dg_part **out) {
    clear_error();
    DG_TRY(check_ctx(ctx));
    DG_REQUIRE(out != nullptr, DG_ERR_INVALID, "null out pointer");
    *out = nullptr;
    DG_REQUIRE(n_global >= 0 && row0 >= 0 && n_local >= 0 && nnz_local >= 0, DG_ERR_INVALID, "negative size");
    DG_REQUIRE(row0 % 32 == 0 && n_local % 32 == 0, DG_ERR_INVALID,
               "row0 (%d) and n_local (%d) must be multiples of 32 (bitmap words)", row0, n_local);
    DG_REQUIRE(row_ptr_local && (col_idx_global || nnz_local == 0), DG_ERR_INVALID, "null CSR pointer");
    DeviceGuard guard(ctx->device);
    dg_part *p = new (std::nothrow) dg_part();
    DG_REQUIRE(p != nullptr, DG_ERR_INVALID, "out of host memory");
    p->ctx = ctx;
    p->n_global = n_global;
    p->row0 = row0;
    p->n_local = n_local;
    p->nnz = nnz_local;
    int st = [&]() -> int {
        if (mem == DG_MEM_DEVICE) {
            p->row_ptr = const_cast<int32_t *>(row_ptr_local);
            p->col_idx = const_cast<int32_t *>(col_idx_global);
            return DG_OK;
        }
        DG_REQUIRE(row_ptr_local[0] == 0 && row_ptr_local[n_local] == nnz_local, DG_ERR_INVALID,
                   "row_ptr_local must run from 0 to nnz_local");
        p->owns = true;
        DG_CUDA_CHECK(cudaMalloc((void **)&p->row_ptr, sizeof(int32_t) * ((size_t)n_local + 1)));
        DG_CUDA_CHECK(cudaMalloc((void **)&p->col_idx, sizeof(int32_t) * std::max<size_t>((size_t)nnz_local, 1)));
        DG_CUDA_CHECK(cudaMemcpyAsync(p->row_ptr, row_ptr_local, sizeof(int32_t) * ((size_t)n_local + 1),
                                      cudaMemcpyHostToDevice, ctx->stream));
        if (nnz_local)
            DG_CUDA_CHECK(cudaMemcpyAsync(p->col_idx, col_idx_global, sizeof(int32_t) * (size_t)nnz_local,
                                          cudaMemcpyHostToDevice, ctx->stream));
        return finish(ctx);
    }();
    if (st != DG_OK) {
        dg_part_destroy(p);
        return st;
    }
    *out = p;
    return DG_OK;
}

void dg_part_destroy(dg_part *p) {
    if (!p) return;
    DeviceGuard guard(p->ctx->device);
    cudaStreamSynchronize(p->ctx->stream);
    if (p->owns) {
        if (p->row_ptr) cudaFree(p->row_ptr);
        if (p->col_idx) cudaFree(p->col_idx);
    }
    if (p->h_counts) cudaFreeHost(p->h_counts);
    if (p->d_rounds) cudaFree(p->d_rounds);
    delete p;
}

static inline PartView part_view(const dg_part *p) {
    return PartView{p->n_global, p->row0, p->n_local, p->nnz, p->row_ptr, p->col_idx, p->peers};
}

int dg_part_prepare(dg_part *p, int32_t feature_size, const uint8_t *keep, const float *x0, float *dinv, float *y) {
    clear_error();
    DG_REQUIRE(p && dinv && y && feature_size >= 1, DG_ERR_INVALID, "bad argument");
    DeviceGuard guard(p->ctx->device);
    return part_prepare(p->ctx, part_view(p), keep, x0, 1.0f / (float)feature_size, dinv, y);
}

int dg_part_first(dg_part *p, int32_t feature_size, const float *dinv, const float *y, const uint8_t *keep,
                  const float *x0, float *pair) {
    clear_error();
    DG_REQUIRE(p && dinv && y && pair && feature_size >= 1, DG_ERR_INVALID, "bad argument");
    DeviceGuard guard(p->ctx->device);
    return part_first(p->ctx, part_view(p), dinv, y, keep, x0, 1.0f / (float)feature_size,
                      reinterpret_cast<float2 *>(pair));
}

static int check_part_model(const dg_part *p, const dg_model *m) {
    DG_REQUIRE(p && m, DG_ERR_INVALID, "null argument");
    DG_REQUIRE(m->n_layers >= 2 && m->layers.back().c_out == 1 && m->head == DG_HEAD_LINEAR && m->tail_w0,
               DG_ERR_UNSUPPORTED, "the row-partitioned path needs >= 2 layers and a one-column linear head");
    return DG_OK;
}

int dg_part_project(dg_part *p, const dg_model *m, const float *dinv, const float *pair, float *pair2) {
    clear_error();
    DG_TRY(check_part_model(p, m));
    DG_REQUIRE(m->n_layers == 2, DG_ERR_INVALID, "dg_part_project is for two-layer models");
    DG_REQUIRE(dinv && pair && pair2, DG_ERR_INVALID, "null argument");
    DeviceGuard guard(p->ctx->device);
    return part_project(p->ctx, part_view(p), m, dinv, reinterpret_cast<const float2 *>(pair),
                        pair2);
}

int dg_part_layer(dg_part *p, const dg_model *m, int32_t layer, const float *dinv, const float *pair, const float *hin,
                  float *hout) {
    clear_error();
    DG_TRY(check_part_model(p, m));
    DG_REQUIRE(layer >= 1 && layer <= m->n_layers - 2, DG_ERR_INVALID, "layer %d is not a hidden layer", layer);
    DG_REQUIRE(dinv && hout && (layer == 1 ? pair != nullptr : hin != nullptr), DG_ERR_INVALID, "null argument");
    DG_REQUIRE(m->layers[layer].wcat != nullptr, DG_ERR_UNSUPPORTED, "layer %d is too wide", layer);
    DeviceGuard guard(p->ctx->device);
    return part_layer(p->ctx, part_view(p), m, layer, dinv, reinterpret_cast<const float2 *>(pair), hin, hout);
}

int dg_part_tail(dg_part *p, const dg_model *m, const float *dinv, const float *hin, float *pair2) {
    clear_error();
    DG_TRY(check_part_model(p, m));
    DG_REQUIRE(m->n_layers >= 3 && dinv && hin && pair2, DG_ERR_INVALID, "bad argument");
    DeviceGuard guard(p->ctx->device);
    return part_tail(p->ctx, part_view(p), m, dinv, hin, pair2);
}

int dg_part_last(dg_part *p, const dg_model *m, const float *dinv, const float *pair2, const uint8_t *keep,
                 const double *wts, int predict, float *score, double *util) {
    clear_error();
    DG_TRY(check_part_model(p, m));
    DG_REQUIRE(dinv && pair2 && (score || util), DG_ERR_INVALID, "null argument");
    DG_REQUIRE(predict == DG_PREDICT_MIS || wts != nullptr || util == nullptr, DG_ERR_INVALID,
               "weights required for predict=mwis");
    DeviceGuard guard(p->ctx->device);
    return part_last(p->ctx, part_view(p), m, dinv, pair2, keep, wts, predict, score,
                     util);
}

int dg_model_padded_width(const dg_model *m, int32_t layer) {
    if (!m || layer < 0 || layer >= m->n_layers) return 0;
    return m->layers[layer].cpo;
}

int dg_part_lgs_init(dg_part *p, const uint8_t *keep, uint32_t *remain, uint8_t *member, int64_t *count) {
    clear_error();
    DG_REQUIRE(p && remain && member && count, DG_ERR_INVALID, "null argument");
    DeviceGuard guard(p->ctx->device);
    return part_lgs_init(p->ctx, part_view(p), keep, remain, member, reinterpret_cast<long long *>(count));
}

int dg_part_lgs_decide(dg_part *p, const double *util, const uint32_t *remain, uint32_t *joined, uint8_t *member) {
    clear_error();
    DG_REQUIRE(p && util && remain && joined && member, DG_ERR_INVALID, "null argument");
    DeviceGuard guard(p->ctx->device);
    return part_lgs_decide(p->ctx, part_view(p), util, remain, joined, member);
}

// ---- peer arenas (CUDA IPC) and the fused exchange -------------------------------------------------------
int dg_peer_alloc(dg_context *ctx, uint64_t bytes, void **dev_ptr, uint8_t *handle_out) {
    clear_error();
    DG_TRY(check_ctx(ctx));
    DG_REQUIRE(dev_ptr && handle_out && bytes > 0, DG_ERR_INVALID, "bad argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == DG_PEER_HANDLE_BYTES, "IPC handle size");
    DeviceGuard guard(ctx->device);
    void *p = nullptr;
    DG_CUDA_CHECK(cudaMalloc(&p, bytes));
    cudaIpcMemHandle_t h;
    cudaError_t e = cudaIpcGetMemHandle(&h, p);
    if (e != cudaSuccess) {
        cudaFree(p);
        set_error("cudaIpcGetMemHandle failed: %s", cudaGetErrorString(e));
        return DG_ERR_CUDA;
    }
    DG_CUDA_CHECK(cudaMemsetAsync(p, 0, bytes, ctx->stream));
    DG_CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    memcpy(handle_out, &h, sizeof(h));
    *dev_ptr = p;
    return DG_OK;
}

int dg_peer_open(dg_context *ctx, const uint8_t *handle, void **peer_ptr) {
    clear_error();
    DG_TRY(check_ctx(ctx));
    DG_REQUIRE(handle && peer_ptr, DG_ERR_INVALID, "null argument");
    DeviceGuard guard(ctx->device);
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, sizeof(h));
    DG_CUDA_CHECK(cudaIpcOpenMemHandle(peer_ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return DG_OK;
}

int dg_peer_close(dg_context *ctx, void *peer_ptr) {
    clear_error();
    DG_TRY(check_ctx(ctx));
    DeviceGuard guard(ctx->device);
    if (peer_ptr) DG_CUDA_CHECK(cudaIpcCloseMemHandle(peer_ptr));
    return DG_OK;
}

int dg_peer_free(dg_context *ctx, void *dev_ptr) {
    clear_error();
    DG_TRY(check_ctx(ctx));
    DeviceGuard guard(ctx->device);
    if (dev_ptr) {
        DG_CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
        DG_CUDA_CHECK(cudaFree(dev_ptr));
    }
    return DG_OK;
}

int dg_part_set_peers(dg_part *p, int32_t world, int32_t rank, void *const *arena_bases, uint64_t arena_bytes,
                      uint64_t flags_offset, uint64_t counts_offset) {
    clear_error();
    DG_REQUIRE(p != nullptr, DG_ERR_INVALID, "null part");
    DG_REQUIRE(world >= 1 && world <= kMaxPeers && rank >= 0 && rank < world, DG_ERR_INVALID,
               "world %d / rank %d out of range (at most %d ranks)", world, rank, kMaxPeers);
    DG_REQUIRE(world == 1 || arena_bases != nullptr, DG_ERR_INVALID, "null arena table");
    DG_REQUIRE(flags_offset % 4 == 0 && counts_offset % 8 == 0 &&
                   flags_offset + 4ull * world <= arena_bytes && counts_offset + 8ull * world <= arena_bytes,
               DG_ERR_INVALID, "flag / count blocks outside the arena");
    PeerMap pm;
    pm.world = world;
    pm.rank = rank;
    pm.bytes = arena_bytes;
    for (int r = 0; r < world && world > 1; ++r) {
        DG_REQUIRE(arena_bases[r] != nullptr, DG_ERR_INVALID, "arena of rank %d is null", r);
        pm.base[r] = static_cast<char *>(arena_bases[r]);
    }
    p->peers = pm;
    p->flags_off = flags_offset;
    p->counts_off = counts_offset;
    p->epoch = 0;
    return DG_OK;
}

int dg_part_keep(dg_part *p, const double *wts, int remove_zero_weight, int32_t n_real, uint8_t *keep) {
    clear_error();
    DG_REQUIRE(p && keep && (wts || !remove_zero_weight), DG_ERR_INVALID, "null argument");
    DeviceGuard guard(p->ctx->device);
    return part_keep(p->ctx, part_view(p), wts, remove_zero_weight, n_real, keep);
}

int dg_part_barrier(dg_part *p, const int64_t *count) {
    clear_error();
    DG_REQUIRE(p != nullptr, DG_ERR_INVALID, "null part");
    DeviceGuard guard(p->ctx->device);
    if (p->peers.world <= 1) return DG_OK;
    p->epoch += 1;
    DG_TRY(part_barrier(p->ctx, p->peers, p->flags_off, p->epoch, reinterpret_cast<const long long *>(count),
                        p->counts_off));
    return DG_OK;
}

// The greedy rounds of the row-partitioned solve, driven from here instead of from Python: `check_every` rounds are enqueued
// back to back (decide, barrier, remove, barrier with the ranks' remaining counts) before the counts are read once; rounds
// past the end find nothing left and change nothing, and a one-thread kernel counts only the rounds that started with a vertex
// left, so *rounds_out is what one-round-at-a-time gives.  Every rank sees the same counts, so every rank stops at the same check.
int dg_part_lgs_run(dg_part *p, const double *util, uint32_t *remain, uint32_t *joined, uint8_t *member, int64_t *count,
                    int32_t check_every, int32_t *rounds_out) {
    clear_error();
    DG_REQUIRE(p && util && remain && joined && member && count && rounds_out, DG_ERR_INVALID, "null argument");
    DG_REQUIRE(p->peers.world >= 2, DG_ERR_INVALID, "dg_part_lgs_run needs the peer arenas (dg_part_set_peers)");
    DG_REQUIRE(check_every >= 1 && check_every <= 64, DG_ERR_INVALID, "check_every must be in 1 .. 64");
    dg_context *ctx = p->ctx;
    DeviceGuard guard(ctx->device);
    const int world = p->peers.world;
    if (!p->h_counts) DG_CUDA_CHECK(cudaHostAlloc((void **)&p->h_counts, sizeof(long long) * (kMaxPeers + 1), cudaHostAllocDefault));
    if (!p->d_rounds) DG_CUDA_CHECK(cudaMalloc((void **)&p->d_rounds, sizeof(int)));
    DG_CUDA_CHECK(cudaMemsetAsync(p->d_rounds, 0, sizeof(int), ctx->stream));
    const long long *counts = reinterpret_cast<const long long *>(p->peers.base[p->peers.rank] + p->counts_off);
    long long *cnt = reinterpret_cast<long long *>(count);
    const PartView pv = part_view(p);
    for (int checks = 0;; ++checks) {
        for (int i = 0; i < check_every; ++i) {
            DG_TRY(part_tally(ctx, counts, world, cnt, p->d_rounds));
            DG_TRY(part_lgs_decide(ctx, pv, util, remain, joined, member));
            p->epoch += 1;
            DG_TRY(part_barrier(ctx, p->peers, p->flags_off, p->epoch, nullptr, p->counts_off));
            DG_TRY(part_lgs_remove(ctx, pv, joined, remain, cnt));
            p->epoch += 1;
            DG_TRY(part_barrier(ctx, p->peers, p->flags_off, p->epoch, cnt, p->counts_off));
        }
        DG_CUDA_CHECK(cudaMemcpyAsync(p->h_counts, counts, sizeof(long long) * world, cudaMemcpyDeviceToHost, ctx->stream));
        DG_CUDA_CHECK(cudaMemcpyAsync(p->h_counts + kMaxPeers, p->d_rounds, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
        DG_TRY(finish(ctx));
        long long left = 0;
        for (int r = 0; r < world; ++r) left += p->h_counts[r];
        if (left == 0) break;
        DG_REQUIRE((checks + 1) * check_every < kLgsRoundCap, DG_ERR_NOT_CONVERGED,
                   "local greedy search hit the round cap (NaN utilities or self-loops?)");
    }
    *rounds_out = *reinterpret_cast<const int *>(p->h_counts + kMaxPeers);
    return DG_OK;
}

int dg_part_lgs_remove(dg_part *p, const uint32_t *joined, uint32_t *remain, int64_t *count) {
    clear_error();
    DG_REQUIRE(p && joined && remain && count, DG_ERR_INVALID, "null argument");
    DeviceGuard guard(p->ctx->device);
    return part_lgs_remove(p->ctx, part_view(p), joined, remain, reinterpret_cast<long long *>(count));
}

}  // extern "C"
