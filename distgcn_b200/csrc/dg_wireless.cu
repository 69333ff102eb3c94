// Device-resident queue bookkeeping of the multi-channel wireless slot loop (wireless_dqn_test_mc.py:225-366) for a set
// of (network, load) instances advanced together.  The schedulers are the library's solvers (dg_solve / dg_lgs /
// dg_dist_greedy / dg_solve_dit in DG_MEM_DEVICE mode); these kernels do what the reference does around them in
// numpy - q += arrivals (:227), weights q * r (:228-240, :298), capacity of the scheduled vertices (:358-363),
// departures min(q, capacity) and the queue update (:364-365), the sequential variants' queue estimate (:306-309) -
// on arrays that never leave the GPU, so a sweep of T slots is a stream of launches without a host synchronisation.
// All arithmetic in float64, like numpy's (rates are integers, queues integer-valued doubles: every product is exact).
#include <algorithm>
#include <initializer_list>
#include <new>

#include "dg_common.cuh"

struct dg_wireless {
    dg_context *ctx = nullptr;
    int n_links = 0, n_ch = 0, T = 0, n_vertices = 0;
    double *arrivals = nullptr;   // [T][n_links]
    int32_t *rates = nullptr;     // [T][n_links][n_ch]
    int32_t *link_v0 = nullptr;   // [n_links] joint-graph vertex of (link, channel 0)
    int32_t *link_nf = nullptr;   // [n_links] number of links of the link's instance (vertex stride between channels)
    double *q = nullptr, *qest = nullptr, *cap = nullptr;   // [n_links]
    double *history = nullptr;    // [T][n_links] queue lengths after every slot (row 0 = zeros)
    double *w = nullptr;          // [max(n_vertices, n_links)] weights handed to the solver
    uint8_t *member = nullptr;    // [max(n_vertices, n_links)] the solver's answer
};

namespace dg {
namespace {

__global__ void wl_begin_kernel(int n, const double *__restrict__ arr, double *__restrict__ q, double *__restrict__ qest,
                                double *__restrict__ cap) {
    const int l = blockIdx.x * blockDim.x + threadIdx.x;
    if (l >= n) return;
    const double v = q[l] + arr[l];   // queue_mtx_algo[t] = queue_mtx_algo[t-1] + arrival_pkts[t]   (:227)
    q[l] = v;
    qest[l] = v;
    cap[l] = 0.0;
}

// joint graph: vertex link_v0[l] + k * nf[l] is link l on channel k (wireless_rollout_test_flood.py:98-133);
// weight = q[l] * rate[t][l][k] (the order='F' reshape of :240)
__global__ void wl_joint_weights_kernel(int n_links, int n_ch, const double *__restrict__ q, const int32_t *__restrict__ rates,
                                        const int32_t *__restrict__ v0, const int32_t *__restrict__ nf, double *__restrict__ w) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_links * n_ch) return;
    const int l = i / n_ch, k = i - l * n_ch;
    w[v0[l] + k * nf[l]] = q[l] * (double)rates[(size_t)l * n_ch + k];
}

// capacity[link] = rate of its scheduled vertex (:358-363; the single-radio clique allows one channel per link; should a
// scheduler ever return two, the highest channel wins as in the reference's ascending-id assignment)
__global__ void wl_joint_serve_kernel(int n_links, int n_ch, const uint8_t *__restrict__ member, const int32_t *__restrict__ rates,
                                      const int32_t *__restrict__ v0, const int32_t *__restrict__ nf, double *__restrict__ cap) {
    const int l = blockIdx.x * blockDim.x + threadIdx.x;
    if (l >= n_links) return;
    double c = 0.0;
    for (int k = 0; k < n_ch; ++k)
        if (member[v0[l] + k * nf[l]]) c = (double)rates[(size_t)l * n_ch + k];
    cap[l] = c;
}

__global__ void wl_seq_weights_kernel(int n_links, int n_ch, int ic, const double *__restrict__ qest,
                                      const int32_t *__restrict__ rates, double *__restrict__ w) {
    const int l = blockIdx.x * blockDim.x + threadIdx.x;
    if (l >= n_links) return;
    w[l] = qest[l] * (double)rates[(size_t)l * n_ch + ic];   // wts_ic = queue estimate * rate   (:298 / :319)
}

// the channel's schedule: capacity (later channels overwrite) and the queue estimate for the next channel (:306-309)
__global__ void wl_seq_serve_kernel(int n_links, int n_ch, int ic, const uint8_t *__restrict__ member,
                                    const int32_t *__restrict__ rates, double *__restrict__ qest, double *__restrict__ cap) {
    const int l = blockIdx.x * blockDim.x + threadIdx.x;
    if (l >= n_links || !member[l]) return;
    const double r = (double)rates[(size_t)l * n_ch + ic];
    cap[l] = r;
    qest[l] -= fmin(qest[l], r);
}

__global__ void wl_end_kernel(int n, const double *__restrict__ cap, double *__restrict__ q, double *__restrict__ hist) {
    const int l = blockIdx.x * blockDim.x + threadIdx.x;
    if (l >= n) return;
    const double v = q[l] - fmin(q[l], cap[l]);   // departures = min(queue, capacity); queue -= departures   (:364-365)
    q[l] = v;
    hist[l] = v;
}

// Joint form, one kernel per slot boundary: close slot t (capacity of the scheduled vertex, departures, history row) and
// open slot t + 1 (arrivals, the joint graph's weights) for one link - the same arithmetic, in the same order, as
// wl_joint_serve_kernel + wl_end_kernel + wl_begin_kernel + wl_joint_weights_kernel.
__global__ void wl_joint_turn_kernel(int n_links, int n_ch, const uint8_t *__restrict__ member, const int32_t *__restrict__ rates_close,
                                     double *__restrict__ hist_close, const double *__restrict__ arr_open,
                                     const int32_t *__restrict__ rates_open, const int32_t *__restrict__ v0,
                                     const int32_t *__restrict__ nf, double *__restrict__ q, double *__restrict__ qest,
                                     double *__restrict__ cap, double *__restrict__ w) {
    const int l = blockIdx.x * blockDim.x + threadIdx.x;
    if (l >= n_links) return;
    const int base = v0[l], stride = nf[l];
    double v = q[l];
    if (rates_close) {
        double c = 0.0;
        for (int k = 0; k < n_ch; ++k)
            if (member[base + k * stride]) c = (double)rates_close[(size_t)l * n_ch + k];
        v = v - fmin(v, c);
        hist_close[l] = v;
        cap[l] = c;
    }
    if (arr_open) {
        v = v + arr_open[l];
        qest[l] = v;
        cap[l] = 0.0;
        for (int k = 0; k < n_ch; ++k) w[base + k * stride] = v * (double)rates_open[(size_t)l * n_ch + k];
    }
    q[l] = v;
}

// Sequential form, one kernel between two scheduler calls: serve channel ic_close (capacity, queue estimate), optionally end
// the slot (departures, history row) and begin the next one (arrivals), then the weights of channel ic_open - per link, in
// the order of wl_seq_serve_kernel, wl_end_kernel, wl_begin_kernel, wl_seq_weights_kernel.
__global__ void wl_seq_turn_kernel(int n_links, int n_ch, int ic_close, int ic_open, const uint8_t *__restrict__ member,
                                   const int32_t *__restrict__ rates_close, double *__restrict__ hist_end,
                                   const double *__restrict__ arr_begin, const int32_t *__restrict__ rates_open,
                                   double *__restrict__ q, double *__restrict__ qest, double *__restrict__ cap,
                                   double *__restrict__ w) {
    const int l = blockIdx.x * blockDim.x + threadIdx.x;
    if (l >= n_links) return;
    double qe = qest[l], c = cap[l], v = q[l];
    if (ic_close >= 0 && member[l]) {
        const double r = (double)rates_close[(size_t)l * n_ch + ic_close];
        c = r;
        qe -= fmin(qe, r);
    }
    if (hist_end) {
        v = v - fmin(v, c);
        hist_end[l] = v;
    }
    if (arr_begin) {
        v = v + arr_begin[l];
        qe = v;
        c = 0.0;
    }
    q[l] = v, qest[l] = qe, cap[l] = c;
    if (ic_open >= 0) w[l] = qe * (double)rates_open[(size_t)l * n_ch + ic_open];
}

inline int blocks(int n) { return (n + 255) / 256; }

template <typename T>
int dev_alloc(T **p, size_t count) {
    DG_CUDA_CHECK(cudaMalloc((void **)p, sizeof(T) * std::max<size_t>(count, 1)));
    return DG_OK;
}

}  // namespace
}  // namespace dg

using namespace dg;

extern "C" {

void dg_wireless_destroy(dg_wireless *s) {
    if (!s) return;
    if (s->ctx && s->ctx->stream) cudaStreamSynchronize(s->ctx->stream);
    for (void *p : {(void *)s->arrivals, (void *)s->rates, (void *)s->link_v0, (void *)s->link_nf, (void *)s->q, (void *)s->qest,
                    (void *)s->cap, (void *)s->history, (void *)s->w, (void *)s->member})
        if (p) cudaFree(p);
    delete s;
}

int dg_wireless_create(dg_context *ctx, int32_t n_links, int32_t n_ch, int32_t n_slots, const double *arrivals,
                       const int32_t *rates, const int32_t *link_v0, const int32_t *link_nf, int32_t n_vertices,
                       dg_wireless **out) {
    clear_error();
    DG_REQUIRE(ctx && out, DG_ERR_INVALID, "null argument");
    *out = nullptr;
    DG_REQUIRE(n_links >= 0 && n_ch >= 1 && n_slots >= 1 && n_vertices >= 0, DG_ERR_INVALID, "bad size");
    DG_REQUIRE(n_links == 0 || (arrivals && rates), DG_ERR_INVALID, "null traffic arrays");
    DG_REQUIRE((link_v0 != nullptr) == (link_nf != nullptr), DG_ERR_INVALID, "link_v0 and link_nf go together");
    int prev = -1;
    cudaGetDevice(&prev);
    if (prev != ctx->device) cudaSetDevice(ctx->device);
    dg_wireless *s = new (std::nothrow) dg_wireless();
    DG_REQUIRE(s != nullptr, DG_ERR_INVALID, "out of host memory");
    s->ctx = ctx;
    s->n_links = n_links, s->n_ch = n_ch, s->T = n_slots, s->n_vertices = n_vertices;
    const size_t nl = (size_t)n_links, nw = std::max<size_t>((size_t)n_vertices, nl);
    int st = [&]() -> int {
        DG_TRY(dev_alloc(&s->arrivals, nl * n_slots));
        DG_TRY(dev_alloc(&s->rates, nl * n_slots * n_ch));
        DG_TRY(dev_alloc(&s->q, nl));
        DG_TRY(dev_alloc(&s->qest, nl));
        DG_TRY(dev_alloc(&s->cap, nl));
        DG_TRY(dev_alloc(&s->history, nl * n_slots));
        DG_TRY(dev_alloc(&s->w, nw));
        DG_TRY(dev_alloc(&s->member, nw));
        if (nl) {
            DG_CUDA_CHECK(cudaMemcpyAsync(s->arrivals, arrivals, sizeof(double) * nl * n_slots, cudaMemcpyHostToDevice, ctx->stream));
            DG_CUDA_CHECK(cudaMemcpyAsync(s->rates, rates, sizeof(int32_t) * nl * n_slots * n_ch, cudaMemcpyHostToDevice,
                                          ctx->stream));
        }
        if (link_v0) {
            DG_TRY(dev_alloc(&s->link_v0, nl));
            DG_TRY(dev_alloc(&s->link_nf, nl));
            if (nl) {
                DG_CUDA_CHECK(cudaMemcpyAsync(s->link_v0, link_v0, sizeof(int32_t) * nl, cudaMemcpyHostToDevice, ctx->stream));
                DG_CUDA_CHECK(cudaMemcpyAsync(s->link_nf, link_nf, sizeof(int32_t) * nl, cudaMemcpyHostToDevice, ctx->stream));
            }
        }
        DG_CUDA_CHECK(cudaMemsetAsync(s->q, 0, sizeof(double) * std::max<size_t>(nl, 1), ctx->stream));
        DG_CUDA_CHECK(cudaMemsetAsync(s->qest, 0, sizeof(double) * std::max<size_t>(nl, 1), ctx->stream));
        DG_CUDA_CHECK(cudaMemsetAsync(s->cap, 0, sizeof(double) * std::max<size_t>(nl, 1), ctx->stream));
        DG_CUDA_CHECK(cudaMemsetAsync(s->history, 0, sizeof(double) * std::max<size_t>(nl * n_slots, 1), ctx->stream));
        DG_CUDA_CHECK(cudaMemsetAsync(s->w, 0, sizeof(double) * nw, ctx->stream));
        DG_CUDA_CHECK(cudaStreamSynchronize(ctx->stream));  // the caller may release its arrays
        return DG_OK;
    }();
    if (prev >= 0 && prev != ctx->device) cudaSetDevice(prev);
    if (st != DG_OK) {
        dg_wireless_destroy(s);
        return st;
    }
    *out = s;
    return DG_OK;
}

int dg_wireless_buffers(dg_wireless *s, double **w, uint8_t **member, double **q) {
    clear_error();
    DG_REQUIRE(s != nullptr, DG_ERR_INVALID, "null argument");
    if (w) *w = s->w;
    if (member) *member = s->member;
    if (q) *q = s->q;
    return DG_OK;
}

#define WL_PRELUDE(cond, msg)                                         \
    clear_error();                                                    \
    DG_REQUIRE(s != nullptr, DG_ERR_INVALID, "null argument");        \
    DG_REQUIRE(t >= 1 && t < s->T, DG_ERR_INVALID, "slot %d outside 1 .. %d", t, s->T - 1); \
    DG_REQUIRE(cond, DG_ERR_INVALID, msg);                            \
    if (s->n_links == 0) return DG_OK;                                \
    cudaStream_t st = s->ctx->stream;                                 \
    const int32_t *rates_t = s->rates + (size_t)t * s->n_links * s->n_ch;

int dg_wireless_begin_slot(dg_wireless *s, int32_t t) {
    WL_PRELUDE(true, "")
    (void)rates_t;
    wl_begin_kernel<<<blocks(s->n_links), 256, 0, st>>>(s->n_links, s->arrivals + (size_t)t * s->n_links, s->q, s->qest, s->cap);
    s->ctx->launches++;
    DG_CUDA_CHECK(cudaGetLastError());
    return DG_OK;
}

int dg_wireless_joint_weights(dg_wireless *s, int32_t t) {
    WL_PRELUDE(s->link_v0 != nullptr, "no joint-graph vertex map was given")
    wl_joint_weights_kernel<<<blocks(s->n_links * s->n_ch), 256, 0, st>>>(s->n_links, s->n_ch, s->q, rates_t, s->link_v0, s->link_nf,
                                                                          s->w);
    s->ctx->launches++;
    DG_CUDA_CHECK(cudaGetLastError());
    return DG_OK;
}

int dg_wireless_joint_serve(dg_wireless *s, int32_t t) {
    WL_PRELUDE(s->link_v0 != nullptr, "no joint-graph vertex map was given")
    wl_joint_serve_kernel<<<blocks(s->n_links), 256, 0, st>>>(s->n_links, s->n_ch, s->member, rates_t, s->link_v0, s->link_nf, s->cap);
    s->ctx->launches++;
    DG_CUDA_CHECK(cudaGetLastError());
    return DG_OK;
}

int dg_wireless_seq_weights(dg_wireless *s, int32_t t, int32_t channel) {
    WL_PRELUDE(channel >= 0 && channel < s->n_ch, "bad channel")
    wl_seq_weights_kernel<<<blocks(s->n_links), 256, 0, st>>>(s->n_links, s->n_ch, channel, s->qest, rates_t, s->w);
    s->ctx->launches++;
    DG_CUDA_CHECK(cudaGetLastError());
    return DG_OK;
}

int dg_wireless_seq_serve(dg_wireless *s, int32_t t, int32_t channel) {
    WL_PRELUDE(channel >= 0 && channel < s->n_ch, "bad channel")
    wl_seq_serve_kernel<<<blocks(s->n_links), 256, 0, st>>>(s->n_links, s->n_ch, channel, s->member, rates_t, s->qest, s->cap);
    s->ctx->launches++;
    DG_CUDA_CHECK(cudaGetLastError());
    return DG_OK;
}

int dg_wireless_end_slot(dg_wireless *s, int32_t t) {
    WL_PRELUDE(true, "")
    (void)rates_t;
    wl_end_kernel<<<blocks(s->n_links), 256, 0, st>>>(s->n_links, s->cap, s->q, s->history + (size_t)t * s->n_links);
    s->ctx->launches++;
    DG_CUDA_CHECK(cudaGetLastError());
    return DG_OK;
}

// The slot loop itself (wireless_dqn_test_mc.py:225-366), slots t_first .. t_first + n_slots - 1, enqueued from here instead
// of from Python: begin, weights, the slot's scheduler on the resident buffers, capacities, departures - per slot (joint
// graph) or per slot and channel (the sequential variants).  Nothing is synchronised; dg_wireless_read_history ends the sweep.
int dg_wireless_run(dg_wireless *s, const dg_model *model, dg_batch *const *batches, int32_t n_batches, int32_t scheduler,
                    int32_t sequential, int predict, int remove_zero_weight, double epsilon, int32_t t_first, int32_t n_slots) {
    clear_error();
    DG_REQUIRE(s && batches && n_batches >= 1, DG_ERR_INVALID, "null argument");
    DG_REQUIRE(scheduler >= DG_WL_LGS && scheduler <= DG_WL_SOLVE_DIT, DG_ERR_INVALID, "unknown scheduler %d", scheduler);
    DG_REQUIRE(scheduler < DG_WL_SOLVE || model != nullptr, DG_ERR_INVALID, "this scheduler needs a model");
    DG_REQUIRE(n_batches == (sequential ? s->n_ch : 1), DG_ERR_INVALID,
               "%d batches: the joint form takes one, the sequential form one per channel (%d)", n_batches, s->n_ch);
    DG_REQUIRE(n_slots >= 0 && t_first >= 1 && (n_slots == 0 || t_first + n_slots - 1 < s->T), DG_ERR_INVALID,
               "slots %d .. %d outside 1 .. %d", t_first, t_first + n_slots - 1, s->T - 1);
    dg_context *ctx = s->ctx;
    auto schedule = [&](dg_batch *b) -> int {
        switch (scheduler) {
        case DG_WL_LGS:
            if (sequential) DG_TRY(dg_batch_set_keep_from_weights(b, s->w, DG_MEM_DEVICE));   // the non-zero sub-graph (:299-301)
            return dg_lgs(ctx, b, s->w, -1, s->member, nullptr, nullptr, nullptr, nullptr, nullptr, DG_MEM_DEVICE);
        case DG_WL_DIST_GREEDY:
            return dg_dist_greedy(ctx, b, s->w, epsilon, s->member, nullptr, DG_MEM_DEVICE);
        case DG_WL_SOLVE:
            return dg_solve(ctx, model, b, s->w, predict, remove_zero_weight, s->member, nullptr, nullptr, nullptr, nullptr,
                            DG_MEM_DEVICE);
        default:
            return dg_solve_dit(ctx, model, b, s->w, predict, s->member, nullptr, nullptr, DG_MEM_DEVICE);
        }
    };
    if (!sequential && n_slots > 0 && s->n_links > 0) {   // joint form: the scheduler + ONE bookkeeping kernel per slot
        DG_REQUIRE(s->link_v0 != nullptr, DG_ERR_INVALID, "no joint-graph vertex map was given");
        const size_t row = (size_t)s->n_links;
        auto turn = [&](int32_t t_close, int32_t t_open) -> int {
            wl_joint_turn_kernel<<<blocks(s->n_links), 256, 0, ctx->stream>>>(
                s->n_links, s->n_ch, s->member, t_close ? s->rates + (size_t)t_close * row * s->n_ch : nullptr,
                t_close ? s->history + (size_t)t_close * row : nullptr, t_open ? s->arrivals + (size_t)t_open * row : nullptr,
                t_open ? s->rates + (size_t)t_open * row * s->n_ch : nullptr, s->link_v0, s->link_nf, s->q, s->qest, s->cap, s->w);
            ctx->launches++;
            DG_CUDA_CHECK(cudaGetLastError());
            return DG_OK;
        };
        DG_TRY(turn(0, t_first));
        for (int32_t t = t_first; t < t_first + n_slots; ++t) {
            DG_TRY(schedule(batches[0]));
            DG_TRY(turn(t, t + 1 < t_first + n_slots ? t + 1 : 0));
        }
        return DG_OK;
    }
    if (sequential && n_slots > 0 && s->n_links > 0) {   // sequential form: ONE bookkeeping kernel between two scheduler calls
        const size_t row = (size_t)s->n_links;
        auto rates_of = [&](int32_t t) { return s->rates + (size_t)t * row * s->n_ch; };
        auto turn = [&](int ic_close, int32_t t_close, bool end, int32_t t_begin, int ic_open, int32_t t_open) -> int {
            wl_seq_turn_kernel<<<blocks(s->n_links), 256, 0, ctx->stream>>>(
                s->n_links, s->n_ch, ic_close, ic_open, s->member, ic_close >= 0 ? rates_of(t_close) : nullptr,
                end ? s->history + (size_t)t_close * row : nullptr, t_begin ? s->arrivals + (size_t)t_begin * row : nullptr,
                ic_open >= 0 ? rates_of(t_open) : nullptr, s->q, s->qest, s->cap, s->w);
            ctx->launches++;
            DG_CUDA_CHECK(cudaGetLastError());
            return DG_OK;
        };
        DG_TRY(turn(-1, 0, false, t_first, 0, t_first));
        for (int32_t t = t_first; t < t_first + n_slots; ++t) {
            for (int32_t ic = 0; ic < s->n_ch; ++ic) {
                DG_TRY(schedule(batches[ic]));
                if (ic + 1 < s->n_ch) {
                    DG_TRY(turn(ic, t, false, 0, ic + 1, t));
                } else {
                    const bool more = t + 1 < t_first + n_slots;
                    DG_TRY(turn(ic, t, true, more ? t + 1 : 0, more ? 0 : -1, t + 1));
                }
            }
        }
        return DG_OK;
    }
    for (int32_t t = t_first; t < t_first + n_slots; ++t) {
        DG_TRY(dg_wireless_begin_slot(s, t));
        if (!sequential) {
            DG_TRY(dg_wireless_joint_weights(s, t));
            DG_TRY(schedule(batches[0]));
            DG_TRY(dg_wireless_joint_serve(s, t));
        } else {
            for (int32_t ic = 0; ic < s->n_ch; ++ic) {
                DG_TRY(dg_wireless_seq_weights(s, t, ic));
                DG_TRY(schedule(batches[ic]));
                DG_TRY(dg_wireless_seq_serve(s, t, ic));
            }
        }
        DG_TRY(dg_wireless_end_slot(s, t));
    }
    return DG_OK;
}

int dg_wireless_read_history(dg_wireless *s, double *history) {
    clear_error();
    DG_REQUIRE(s && history, DG_ERR_INVALID, "null argument");
    const size_t count = (size_t)s->T * s->n_links;
    if (count) DG_CUDA_CHECK(cudaMemcpyAsync(history, s->history, sizeof(double) * count, cudaMemcpyDeviceToHost, s->ctx->stream));
    return dg_context_synchronize(s->ctx);
}

}  // extern "C"
