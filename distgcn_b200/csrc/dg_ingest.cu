// Native ingest: the reference's per-graph inputs -> one packed batch, without a Python loop.
//
// The reference hands every graph to the solver as its own scipy sparse matrix (mwis_dqn_call.py:198; the .mat files
// of Data_Generation.py:214-219 load as float64 CSC with int32 indptr / indices) and converts it per call through
// networkx (mwis_dqn_call.py:202-207).  Here the caller passes the per-graph arrays as they are - tables of pointers
// to each matrix's indptr / indices (and optionally data: stored zeros are not edges, np.nonzero(adj[v]) at
// heuristics.py:94) - and a small pool of host threads writes the packed form (graph_ptr, row_ptr, 16-bit graph-local
// or 32-bit batch-global column ids) straight into pinned staging, from where the usual H2D copies and kernels run.
// Host code only; compiled into the same library as the kernels.
#include <atomic>
#include <chrono>
#include <cstdio>
#include <condition_variable>
#include <cstring>
#include <functional>
#include <mutex>
#include <thread>

#include "dg_common.cuh"

#if defined(__x86_64__)
#include <immintrin.h>
#endif

namespace dg {
namespace {

// ---- the two inner loops of the packer: column ids of one graph, narrowed to 16 bits or re-based to batch ids.
// Return non-zero when some id lies outside [0, n).
#if defined(__x86_64__)
// The outputs are pinned staging that the copy engine reads next: NON-TEMPORAL stores send them to memory instead of
// leaving dirty lines in the packing cores' caches (measured on the B200 box: a 5 MB H2D copy of staging written with
// ordinary stores by 8 threads takes 650 us, by 1 thread 120 us - the DMA reads snoop every writer's cache).
__attribute__((target("avx2"))) unsigned narrow16_avx2(const int32_t *ix, int64_t e, uint32_t n, uint16_t *out) {
    const __m256i lim = _mm256_set1_epi32((int)(n - 1));
    __m256i bad = _mm256_setzero_si256();
    unsigned acc = 0;
    int64_t k = 0;
    for (; k < e && ((uintptr_t)(out + k) & 31); ++k) {  // head: up to the first 32-byte boundary of the output
        const uint32_t c = (uint32_t)ix[k];
        acc |= (c >= n);
        out[k] = (uint16_t)c;
    }
    for (; k + 16 <= e; k += 16) {
        const __m256i a = _mm256_loadu_si256((const __m256i *)(ix + k));
        const __m256i b = _mm256_loadu_si256((const __m256i *)(ix + k + 8));
        // unsigned c > n - 1  <=>  min_u(c, n - 1) != c
        bad = _mm256_or_si256(bad, _mm256_xor_si256(_mm256_min_epu32(a, lim), a));
        bad = _mm256_or_si256(bad, _mm256_xor_si256(_mm256_min_epu32(b, lim), b));
        const __m256i p = _mm256_permute4x64_epi64(_mm256_packus_epi32(a, b), 0xD8);
        _mm256_stream_si256((__m256i *)(out + k), p);
    }
    acc |= _mm256_testz_si256(bad, bad) ? 0u : 1u;
    for (; k < e; ++k) {
        const uint32_t c = (uint32_t)ix[k];
        acc |= (c >= n);
        out[k] = (uint16_t)c;
    }
    return acc;
}
__attribute__((target("avx2"))) unsigned rebase32_avx2(const int32_t *ix, int64_t e, uint32_t n, uint32_t v0, int32_t *out) {
    const __m256i lim = _mm256_set1_epi32((int)(n - 1)), off = _mm256_set1_epi32((int)v0);
    __m256i bad = _mm256_setzero_si256();
    unsigned acc = 0;
    int64_t k = 0;
    for (; k < e && ((uintptr_t)(out + k) & 31); ++k) {
        const uint32_t c = (uint32_t)ix[k];
        acc |= (c >= n);
        out[k] = (int32_t)(c + v0);
    }
    for (; k + 8 <= e; k += 8) {
        const __m256i a = _mm256_loadu_si256((const __m256i *)(ix + k));
        bad = _mm256_or_si256(bad, _mm256_xor_si256(_mm256_min_epu32(a, lim), a));
        _mm256_stream_si256((__m256i *)(out + k), _mm256_add_epi32(a, off));
    }
    acc |= _mm256_testz_si256(bad, bad) ? 0u : 1u;
    for (; k < e; ++k) {
        const uint32_t c = (uint32_t)ix[k];
        acc |= (c >= n);
        out[k] = (int32_t)(c + v0);
    }
    return acc;
}
// row_ptr of one graph: out[r] = e0 + ip[r + 1]; returns non-zero when ip decreases somewhere
__attribute__((target("avx2"))) unsigned rowptr_avx2(const int32_t *ip, int n, int32_t e0, int32_t *out) {
    unsigned acc = 0;
    int r = 0;
    for (; r < n && ((uintptr_t)(out + r) & 31); ++r) {
        acc |= ip[r + 1] < ip[r];
        out[r] = e0 + ip[r + 1];
    }
    const __m256i off = _mm256_set1_epi32(e0);
    __m256i bad = _mm256_setzero_si256();
    for (; r + 8 <= n; r += 8) {
        const __m256i lo = _mm256_loadu_si256((const __m256i *)(ip + r));
        const __m256i hi = _mm256_loadu_si256((const __m256i *)(ip + r + 1));
        bad = _mm256_or_si256(bad, _mm256_cmpgt_epi32(lo, hi));
        _mm256_stream_si256((__m256i *)(out + r), _mm256_add_epi32(hi, off));
    }
    acc |= _mm256_testz_si256(bad, bad) ? 0u : 1u;
    for (; r < n; ++r) {
        acc |= ip[r + 1] < ip[r];
        out[r] = e0 + ip[r + 1];
    }
    return acc;
}
__attribute__((target("avx2"))) void copy_nt_avx2(const double *src, int64_t count, double *dst) {
    int64_t k = 0;
    for (; k < count && ((uintptr_t)(dst + k) & 31); ++k) dst[k] = src[k];
    for (; k + 4 <= count; k += 4) _mm256_stream_pd(dst + k, _mm256_loadu_pd(src + k));
    for (; k < count; ++k) dst[k] = src[k];
}
const bool kHaveAvx2 = __builtin_cpu_supports("avx2");
#else
const bool kHaveAvx2 = false;
#endif

unsigned narrow16(const int32_t *ix, int64_t e, uint32_t n, uint16_t *out) {
#if defined(__x86_64__)
    if (kHaveAvx2 && n > 0) return narrow16_avx2(ix, e, n, out);
#endif
    unsigned acc = 0;
    for (int64_t k = 0; k < e; ++k) {
        const uint32_t c = (uint32_t)ix[k];
        acc |= (c >= n);
        out[k] = (uint16_t)c;
    }
    return acc;
}
unsigned rebase32(const int32_t *ix, int64_t e, uint32_t n, uint32_t v0, int32_t *out) {
#if defined(__x86_64__)
    if (kHaveAvx2 && n > 0) return rebase32_avx2(ix, e, n, v0, out);
#endif
    unsigned acc = 0;
    for (int64_t k = 0; k < e; ++k) {
        const uint32_t c = (uint32_t)ix[k];
        acc |= (c >= n);
        out[k] = (int32_t)(c + v0);
    }
    return acc;
}

unsigned rowptr(const int32_t *ip, int n, int32_t e0, int32_t *out) {
#if defined(__x86_64__)
    if (kHaveAvx2) return rowptr_avx2(ip, n, e0, out);
#endif
    unsigned acc = 0;
    for (int r = 0; r < n; ++r) {
        acc |= ip[r + 1] < ip[r];
        out[r] = e0 + ip[r + 1];
    }
    return acc;
}
void copy_doubles(const double *src, int64_t count, double *dst) {
#if defined(__x86_64__)
    if (kHaveAvx2) return copy_nt_avx2(src, count, dst);
#endif
    memcpy(dst, src, sizeof(double) * (size_t)count);
}
inline void store_fence() {
#if defined(__x86_64__)
    _mm_sfence();
#endif
}

// A persistent pool: parallel_for(n, fn) runs fn(task) for task = 0..n-1 on the workers and the calling thread.
class Pool {
  public:
    static Pool &get() {
        static Pool p;
        return p;
    }
    int size() const { return (int)workers_.size() + 1; }
    void parallel_for(int n_tasks, int max_threads, const std::function<void(int)> &fn) {
        if (n_tasks <= 0) return;
        const int helpers = std::min({(int)workers_.size(), std::max(0, max_threads - 1), n_tasks - 1});
        if (helpers == 0) {
            for (int t = 0; t < n_tasks; ++t) fn(t);
            return;
        }
        std::unique_lock<std::mutex> run(run_mu_);  // one parallel_for at a time
        {
            std::lock_guard<std::mutex> lk(mu_);
            fn_ = &fn;
            n_tasks_ = n_tasks;
            next_.store(0);
            pending_ = helpers;
            wanted_ = helpers;
            ++epoch_;
        }
        cv_.notify_all();
        for (int t; (t = next_.fetch_add(1)) < n_tasks;) fn(t);
        std::unique_lock<std::mutex> lk(mu_);
        done_cv_.wait(lk, [&] { return pending_ == 0; });
        fn_ = nullptr;
    }

  private:
    Pool() {
        unsigned hc = std::thread::hardware_concurrency();
        int n = (int)std::min<unsigned>(hc ? hc : 4, 16) - 1;
        if (const char *e = getenv("DG_INGEST_THREADS")) n = std::max(0, atoi(e) - 1);
        for (int i = 0; i < n; ++i) workers_.emplace_back([this, i] { loop(i); });
    }
    ~Pool() {
        {
            std::lock_guard<std::mutex> lk(mu_);
            stop_ = true;
            ++epoch_;
        }
        cv_.notify_all();
        for (auto &w : workers_) w.join();
    }
    void loop(int id) {
        uint64_t seen = 0;
        for (;;) {
            const std::function<void(int)> *fn;
            int n;
            {
                std::unique_lock<std::mutex> lk(mu_);
                cv_.wait(lk, [&] { return epoch_ != seen; });
                seen = epoch_;
                if (stop_) return;
                if (id >= wanted_) continue;
                fn = fn_;
                n = n_tasks_;
            }
            for (int t; (t = next_.fetch_add(1)) < n;) (*fn)(t);
            {
                std::lock_guard<std::mutex> lk(mu_);
                --pending_;
            }
            done_cv_.notify_one();
        }
    }
    std::vector<std::thread> workers_;
    std::mutex mu_, run_mu_;
    std::condition_variable cv_, done_cv_;
    const std::function<void(int)> *fn_ = nullptr;
    int n_tasks_ = 0, pending_ = 0, wanted_ = 0;
    std::atomic<int> next_{0};
    uint64_t epoch_ = 0;
    bool stop_ = false;
};

bool pack_graph_upper(const int32_t *ip, const int32_t *ix, const double *d, int n, int64_t eu0, int64_t expect_u,
                      int32_t *rp_u, uint16_t *out);

struct PackPlan {
    std::vector<int64_t> v0, e0;  // vertex / edge offset of every graph, n_graphs + 1
    bool has_zeros = false;       // some stored value is zero: edge counts come from a counting pass
};

// sizes: vertex and edge offsets of every graph.  With `data`, stored zeros do not count.
int plan_sizes(int32_t n_graphs, const int32_t *const *indptr, const double *const *data, const int32_t *n_rows,
               int threads, PackPlan *plan, bool upper = false) {
    plan->v0.assign((size_t)n_graphs + 1, 0);
    plan->e0.assign((size_t)n_graphs + 1, 0);
    std::vector<int64_t> nnz((size_t)n_graphs, 0);
    std::atomic<int> bad{-1};
    std::atomic<bool> zeros{false};
    const int chunk = 64;
    const int n_tasks = (n_graphs + chunk - 1) / chunk;
    Pool::get().parallel_for(n_tasks, data ? threads : 1, [&](int t) {
        for (int g = t * chunk; g < std::min(n_graphs, (t + 1) * chunk); ++g) {
            const int n = n_rows[g];
            if (n < 0 || (n > 0 && (!indptr[g] || indptr[g][0] != 0 || indptr[g][n] < 0))) {
                bad.store(g);
                continue;
            }
            int64_t e = n > 0 ? indptr[g][n] : 0;
            if (data && data[g] && e > 0) {
                int64_t nz = 0;
                const double *d = data[g];
                for (int64_t k = 0; k < e; ++k) nz += d[k] != 0.0;
                if (nz != e) zeros.store(true);
                e = nz;
            }
            nnz[(size_t)g] = e;
        }
    });
    DG_REQUIRE(bad.load() < 0, DG_ERR_INVALID, "graph %d: indptr must start at 0 and have n_rows + 1 entries", bad.load());
    for (int g = 0; g < n_graphs; ++g) {
        plan->v0[(size_t)g + 1] = plan->v0[(size_t)g] + n_rows[g];
        plan->e0[(size_t)g + 1] = plan->e0[(size_t)g] + nnz[(size_t)g];
    }
    plan->has_zeros = zeros.load();
    if (upper) {   // symmetric, zero diagonal: exactly half of every graph's non-zeros lie above the diagonal
        for (int g = 0; g < n_graphs; ++g)
            DG_REQUIRE((nnz[(size_t)g] & 1) == 0, DG_ERR_INVALID,
                       "graph %d: %lld stored non-zeros - an adjacency matrix is symmetric with a zero diagonal", g,
                       (long long)nnz[(size_t)g]);
        for (int g = 0; g <= n_graphs; ++g) plan->e0[(size_t)g] /= 2;
    }
    DG_REQUIRE(plan->v0.back() < (1LL << 31) && plan->e0.back() < (1LL << 31), DG_ERR_INVALID,
               "batch too large for int32 indexing: %lld vertices, %lld edges", (long long)plan->v0.back(),
               (long long)plan->e0.back());
    return DG_OK;
}

// the parallel pack; exactly one of col_idx / col16 is written
int pack_into(int32_t n_graphs, const int32_t *const *indptr, const int32_t *const *indices, const double *const *data,
              const int32_t *n_rows, const PackPlan &plan, int threads, int32_t *graph_ptr, int32_t *row_ptr,
              int32_t *col_idx, uint16_t *col16, const std::function<void()> *side_task = nullptr, bool upper = false) {
    for (int g = 0; g <= n_graphs; ++g) graph_ptr[g] = (int32_t)plan.v0[(size_t)g];
    row_ptr[0] = 0;
    std::atomic<int> bad{-1};
    // tasks of roughly equal edge count: contiguous graph ranges
    const int64_t total_e = plan.e0.back() + plan.v0.back();
    const int want = std::max(1, std::min<int>(n_graphs, threads * 4));
    std::vector<int> cut{0};
    for (int k = 1; k < want; ++k) {
        const int64_t target = total_e * k / want;
        int lo = cut.back(), hi = n_graphs;
        while (lo < hi) {
            const int mid = (lo + hi) / 2;
            if (plan.e0[(size_t)mid] + plan.v0[(size_t)mid] < target) lo = mid + 1;
            else hi = mid;
        }
        if (lo > cut.back() && lo < n_graphs) cut.push_back(lo);
    }
    cut.push_back(n_graphs);
    if (side_task && threads <= 1) {   // a handful of graphs: waking a pool thread costs more than the plan itself
        (*side_task)();
        side_task = nullptr;
    }
    const int first = side_task ? 1 : 0;  // task 0 = the side task (the tile plan of the batch being packed)
    Pool::get().parallel_for((int)cut.size() - 1 + first, std::max(threads, first + 1), [&](int t_all) {
        if (t_all < first) {
            (*side_task)();
            return;
        }
        const int t = t_all - first;
        for (int g = cut[(size_t)t]; g < cut[(size_t)t + 1]; ++g) {
            const int n = n_rows[g];
            if (n == 0) continue;
            const int32_t *ip = indptr[g], *ix = indices[g];
            const double *d = (data && plan.has_zeros) ? data[g] : nullptr;
            const int64_t v0 = plan.v0[(size_t)g], e0 = plan.e0[(size_t)g];
            int32_t *rp = row_ptr + v0 + 1;
            bool ok = true;
            if (upper) {   // plan.e0 already counts upper entries (half of the stored non-zeros)
                if (!pack_graph_upper(ip, ix, (data && plan.has_zeros) ? data[g] : nullptr, n, e0,
                                      plan.e0[(size_t)g + 1] - e0, rp, col16 + e0))
                    bad.store(g);
                continue;
            }
            if (!d) {
                ok = rowptr(ip, n, (int32_t)e0, rp) == 0;
                const int64_t e = ip[n];
                if (!ok || (e > 0 && !ix)) {
                    bad.store(g);
                    continue;
                }
                const unsigned acc = col16 ? narrow16(ix, e, (uint32_t)n, col16 + e0)
                                           : rebase32(ix, e, (uint32_t)n, (uint32_t)v0, col_idx + e0);
                if (acc) bad.store(g);
            } else {  // stored zeros are dropped
                int64_t w = e0;
                for (int r = 0; r < n; ++r) {
                    if (ip[r + 1] < ip[r]) {
                        ok = false;
                        break;
                    }
                    for (int k = ip[r]; k < ip[r + 1]; ++k) {
                        if (d[k] == 0.0) continue;
                        const uint32_t c = (uint32_t)ix[k];
                        if (c >= (uint32_t)n) ok = false;
                        if (col16) col16[w] = (uint16_t)c;
                        else col_idx[w] = (int32_t)(c + (uint32_t)v0);
                        ++w;
                    }
                    rp[r] = (int32_t)w;
                }
                if (!ok || w != plan.e0[(size_t)g + 1]) bad.store(g);
            }
        }
        store_fence();  // the non-temporal stores of this task are globally visible before the task counts as done
    });
    DG_REQUIRE(bad.load() < 0, DG_ERR_INVALID,
               "graph %d: malformed pattern (indptr not non-decreasing, or a column id outside [0, n_rows))", bad.load());
    return DG_OK;
}

int default_threads(int32_t n_threads, int64_t work) {
    int t = n_threads > 0 ? n_threads : Pool::get().size();
    if (work < (1 << 16)) t = 1;  // a handful of graphs: the hand-off costs more than the copy
    return std::max(1, std::min(t, Pool::get().size()));
}

// ordinary memory -> pinned staging with non-temporal stores (head / tail up to the 32-byte boundaries with plain ones)
#if defined(__x86_64__)
__attribute__((target("avx2"))) void copy_nt_bytes_avx2(const void *src_v, size_t bytes, void *dst_v) {
    const char *src = (const char *)src_v;
    char *dst = (char *)dst_v;
    size_t k = 0;
    const size_t head = std::min(bytes, (size_t)((32 - ((uintptr_t)dst & 31)) & 31));
    if (head) memcpy(dst, src, head);
    for (k = head; k + 32 <= bytes; k += 32)
        _mm256_stream_si256((__m256i *)(dst + k), _mm256_loadu_si256((const __m256i *)(src + k)));
    if (k < bytes) memcpy(dst + k, src + k, bytes - k);
}
#endif
void copy_to_staging(const void *src, size_t bytes, void *dst) {
#if defined(__x86_64__)
    if (kHaveAvx2) return copy_nt_bytes_avx2(src, bytes, dst);
#endif
    memcpy(dst, src, bytes);
}

// Upper-triangle pack of one graph: entries with column > row, graph-local 16-bit ids; row_ptr_u[r + 1] = running count.
// Returns false when the pattern is malformed or not half above the diagonal (asymmetric, or with diagonal entries).
// The rows are filtered branch-free into a thread-local scratch (every stored entry is written, the cursor only advances
// past the ones that stay) and the result goes to the pinned staging in one non-temporal copy.
bool pack_graph_upper(const int32_t *ip, const int32_t *ix, const double *d, int n, int64_t eu0, int64_t expect_u,
                      int32_t *rp_u /* this graph's rows: n entries following row_ptr_u[v0] */, uint16_t *out /* at eu0 */) {
    static thread_local std::vector<uint16_t> cols;
    static thread_local std::vector<int32_t> rows;
    const int64_t stored = ip[n];
    if (stored < 0) return false;
    if (cols.size() < (size_t)stored + 1) cols.resize((size_t)stored + 1 + (size_t)stored / 4);
    if (rows.size() < (size_t)n) rows.resize((size_t)n + (size_t)n / 4);
    uint16_t *tmp = cols.data();
    int32_t *rt = rows.data();
    int64_t w = 0;
    unsigned bad = 0;
    // symmetry: every pair {i, j} must be stored once above and once below the diagonal - a signed sum of a hash of the
    // pair over all stored entries (+ above, - below) is zero for a symmetric pattern
    uint64_t balance = 0;
    auto pair_hash = [](uint32_t a, uint32_t b) {
        uint64_t h = ((uint64_t)std::min(a, b) << 32 | std::max(a, b)) * 0x9E3779B97F4A7C15ull;
        h ^= h >> 29;
        return h * 0xBF58476D1CE4E5B9ull;
    };
    for (int r = 0; r < n; ++r) {
        const int b0 = ip[r], b1 = ip[r + 1];
        if (b1 < b0 || b1 > stored) return false;
        if (!d) {
            for (int k = b0; k < b1; ++k) {
                const uint32_t c = (uint32_t)ix[k];
                bad |= (c >= (uint32_t)n) | (c == (uint32_t)r);
                tmp[w] = (uint16_t)c;
                const bool up = c > (uint32_t)r;
                w += up;
                const uint64_t h = pair_hash(c, (uint32_t)r);
                balance += up ? h : (uint64_t)0 - h;
            }
        } else {
            for (int k = b0; k < b1; ++k) {
                const uint32_t c = (uint32_t)ix[k];
                const bool nz = d[k] != 0.0;
                bad |= (c >= (uint32_t)n) | ((c == (uint32_t)r) & nz);
                tmp[w] = (uint16_t)c;
                const bool up = c > (uint32_t)r;
                w += up & nz;
                const uint64_t h = nz ? pair_hash(c, (uint32_t)r) : 0;
                balance += up ? h : (uint64_t)0 - h;
            }
        }
        rt[r] = (int32_t)(eu0 + w);
    }
    bad |= balance != 0;
    if (bad || w != expect_u) return false;
    copy_to_staging(tmp, sizeof(uint16_t) * (size_t)w, out);
    copy_to_staging(rt, sizeof(int32_t) * (size_t)n, rp_u);
    return true;
}

struct Staging {  // pinned, grow-only, one per context
    int32_t *gp = nullptr, *rp = nullptr, *c32 = nullptr;
    uint16_t *c16 = nullptr;
    double *w = nullptr;
    size_t cap_g = 0, cap_n = 0, cap_e16 = 0, cap_e32 = 0, cap_w = 0;
    cudaEvent_t copied = nullptr;  // the H2D copies of the previous call have read the staging
    bool in_flight = false;
};

template <typename T>
int grow_pinned(T **p, size_t *cap, size_t need) {
    if (*cap >= need && *p) return DG_OK;
    if (*p) cudaFreeHost(*p);
    *p = nullptr;
    const size_t c = need + need / 4 + 64;
    DG_CUDA_CHECK(cudaHostAlloc((void **)p, sizeof(T) * c, cudaHostAllocDefault));
    *cap = c;
    return DG_OK;
}

}  // namespace

void ingest_staging_free(dg_context *ctx) {
    Staging *s = static_cast<Staging *>(ctx->ingest_staging);
    if (!s) return;
    if (s->gp) cudaFreeHost(s->gp);
    if (s->rp) cudaFreeHost(s->rp);
    if (s->c32) cudaFreeHost(s->c32);
    if (s->c16) cudaFreeHost(s->c16);
    if (s->w) cudaFreeHost(s->w);
    if (s->copied) cudaEventDestroy(s->copied);
    delete s;
    ctx->ingest_staging = nullptr;
}

}  // namespace dg

using namespace dg;

extern "C" {

int dg_pack_graphs_sizes(int32_t n_graphs, const int32_t *const *indptr, const double *const *data,
                         const int32_t *n_rows, int64_t *n_nodes, int64_t *nnz, int32_t *max_rows) {
    clear_error();
    DG_REQUIRE(n_graphs >= 0 && (n_graphs == 0 || (indptr && n_rows)), DG_ERR_INVALID, "null argument");
    PackPlan plan;
    DG_TRY(plan_sizes(n_graphs, indptr, data, n_rows, default_threads(0, data ? (1 << 20) : 0), &plan));
    if (n_nodes) *n_nodes = plan.v0.back();
    if (nnz) *nnz = plan.e0.back();
    if (max_rows) {
        int32_t m = 0;
        for (int g = 0; g < n_graphs; ++g) m = std::max(m, n_rows[g]);
        *max_rows = m;
    }
    return DG_OK;
}

int dg_pack_graphs_host(int32_t n_graphs, const int32_t *const *indptr, const int32_t *const *indices,
                        const double *const *data, const int32_t *n_rows, int32_t *graph_ptr, int32_t *row_ptr,
                        int32_t *col_idx, uint16_t *col_local16, int32_t n_threads) {
    clear_error();
    DG_REQUIRE(n_graphs >= 0 && graph_ptr && row_ptr, DG_ERR_INVALID, "null argument");
    DG_REQUIRE(n_graphs == 0 || (indptr && indices && n_rows), DG_ERR_INVALID, "null argument");
    DG_REQUIRE((col_idx != nullptr) != (col_local16 != nullptr), DG_ERR_INVALID,
               "pass exactly one of col_idx (32-bit batch-global ids) and col_local16 (16-bit graph-local ids)");
    PackPlan plan;
    const int threads0 = default_threads(n_threads, 1 << 20);
    DG_TRY(plan_sizes(n_graphs, indptr, data, n_rows, threads0, &plan));
    if (col_local16)
        for (int g = 0; g < n_graphs; ++g)
            DG_REQUIRE(n_rows[g] <= 65536, DG_ERR_INVALID, "graph %d has %d vertices: 16-bit column ids need <= 65536", g,
                       n_rows[g]);
    const int threads = default_threads(n_threads, plan.e0.back() + plan.v0.back());
    return pack_into(n_graphs, indptr, indices, data, n_rows, plan, threads, graph_ptr, row_ptr, col_idx, col_local16);
}

int dg_pack_graphs_upper_host(int32_t n_graphs, const int32_t *const *indptr, const int32_t *const *indices,
                              const double *const *data, const int32_t *n_rows, int32_t *graph_ptr, int32_t *row_ptr_upper,
                              uint16_t *col_local_upper, int32_t n_threads) {
    clear_error();
    DG_REQUIRE(n_graphs >= 0 && graph_ptr && row_ptr_upper && col_local_upper, DG_ERR_INVALID, "null argument");
    DG_REQUIRE(n_graphs == 0 || (indptr && indices && n_rows), DG_ERR_INVALID, "null argument");
    for (int g = 0; g < n_graphs; ++g)
        DG_REQUIRE(n_rows[g] <= 65536, DG_ERR_INVALID, "graph %d has %d vertices: 16-bit column ids need <= 65536", g, n_rows[g]);
    PackPlan plan;
    DG_TRY(plan_sizes(n_graphs, indptr, data, n_rows, default_threads(n_threads, 1 << 20), &plan, true));
    const int threads = default_threads(n_threads, 2 * plan.e0.back() + plan.v0.back());
    DG_TRY(pack_into(n_graphs, indptr, indices, data, n_rows, plan, threads, graph_ptr, row_ptr_upper, nullptr, col_local_upper,
                     nullptr, true));
    return DG_OK;
}

int dg_solve_graphs_host(dg_context *ctx, const dg_model *m, int32_t n_graphs, const int32_t *const *indptr,
                         const int32_t *const *indices, const double *const *data, const int32_t *n_rows,
                         const double *const *wts_per_graph, const double *wts_packed, int predict,
                         int remove_zero_weight, uint8_t *member, double *total, int wait) {
    clear_error();
    DG_REQUIRE(ctx && m && member, DG_ERR_INVALID, "null argument");
    DG_REQUIRE(n_graphs >= 0 && (n_graphs == 0 || (indptr && indices && n_rows)), DG_ERR_INVALID, "null argument");
    DG_REQUIRE((wts_per_graph != nullptr) != (wts_packed != nullptr) || n_graphs == 0, DG_ERR_INVALID,
               "pass exactly one of wts_per_graph and wts_packed");
    int prev = -1;
    cudaGetDevice(&prev);
    if (prev != ctx->device) cudaSetDevice(ctx->device);
    struct Restore {
        int prev, dev;
        ~Restore() {
            if (prev >= 0 && prev != dev) cudaSetDevice(prev);
        }
    } restore{prev, ctx->device};
    if (!ctx->ingest_staging) {
        Staging *s = new (std::nothrow) Staging();
        DG_REQUIRE(s != nullptr, DG_ERR_INVALID, "out of host memory");
        if (cudaEventCreateWithFlags(&s->copied, ctx->env.ingest_timing ? cudaEventDefault : cudaEventDisableTiming) != cudaSuccess) {
            delete s;
            set_error("cudaEventCreate failed");
            return DG_ERR_CUDA;
        }
        ctx->ingest_staging = s;
    }
    Staging *s = static_cast<Staging *>(ctx->ingest_staging);
    const bool timing = ctx->env.ingest_timing;
    const auto t_begin = std::chrono::steady_clock::now();
    int32_t max_rows = 0;
    for (int g = 0; g < n_graphs; ++g) max_rows = std::max(max_rows, n_rows[g]);
    const bool narrow = max_rows <= 65536;
    // DG_INGEST_UPPER=1: small graphs travel as their upper triangle - half the column ids over PCIe, but the filtering
    // pack costs ~8x the straight narrowing one (2.9 vs 0.36 ns per stored entry and thread), so it only pays where the
    // link, not the host, is the limit.  Pre-packed callers use dg_pack_graphs_upper_host once + dg_solve_host_upper.
    const bool upper = narrow && max_rows <= 8192 && ctx->env.ingest_upper;
    PackPlan plan;
    DG_TRY(plan_sizes(n_graphs, indptr, data, n_rows, default_threads(0, 1 << 20), &plan, upper));
    const int64_t n = plan.v0.back(), e = plan.e0.back();
    // the previous call's copies must have read the staging before it is overwritten (its kernels may still run)
    const auto t_plan = std::chrono::steady_clock::now();
    if (s->in_flight) {
        DG_CUDA_CHECK(cudaEventSynchronize(s->copied));
        s->in_flight = false;
    }
    const auto t_wait = std::chrono::steady_clock::now();
    DG_TRY(grow_pinned(&s->gp, &s->cap_g, (size_t)n_graphs + 1));
    DG_TRY(grow_pinned(&s->rp, &s->cap_n, (size_t)n + 1));
    if (narrow) DG_TRY(grow_pinned(&s->c16, &s->cap_e16, (size_t)e + (size_t)max_rows + 16));  // (+ the filter's slack)
    else DG_TRY(grow_pinned(&s->c32, &s->cap_e32, (size_t)e + 1));
    const int threads = default_threads(0, e + n);
    // the tensor-core kernel's tile plan needs only the graphs' sizes: one pool thread makes it while the others pack
    dg_batch *hb = nullptr;
    DG_TRY(host_batch_set_meta(ctx, n_graphs, plan.v0.data(), plan.e0.data(), &hb));
    const std::function<void()> plan_tiles = [&] { tc_plan_ahead(ctx, m, hb); };
    DG_TRY(pack_into(n_graphs, indptr, indices, data, n_rows, plan, threads, s->gp, s->rp, narrow ? nullptr : s->c32,
                     narrow ? s->c16 : nullptr, &plan_tiles, upper));
    const double *w = wts_packed;
    if (wts_per_graph) {
        DG_TRY(grow_pinned(&s->w, &s->cap_w, (size_t)n + 1));
        for (int g = 0; g < n_graphs; ++g) {
            DG_REQUIRE(wts_per_graph[g] || n_rows[g] == 0, DG_ERR_INVALID, "graph %d: null weights", g);
            if (n_rows[g]) copy_doubles(wts_per_graph[g], n_rows[g], s->w + plan.v0[(size_t)g]);
        }
        store_fence();
        w = s->w;
    }
    static const double kNoWeights = 0.0;
    if (!w) w = &kNoWeights;  // empty batch
    const auto t_pack = std::chrono::steady_clock::now();
    cudaEvent_t ev_a = nullptr, ev_c = nullptr;
    if (timing) {
        cudaEventCreate(&ev_a);
        cudaEventCreate(&ev_c);
        cudaEventRecord(ev_a, ctx->stream);
    }
    const int st = solve_host_staged(ctx, m, n_graphs, (int32_t)n, (int32_t)e, s->gp, s->rp, narrow ? nullptr : s->c32, w,
                                     predict, remove_zero_weight, member, total, wait != 0, narrow ? s->c16 : nullptr,
                                     s->copied, upper);
    s->in_flight = st == DG_OK && wait == 0;
    if (timing) {
        const auto t_end = std::chrono::steady_clock::now();
        auto us = [](std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point b) {
            return std::chrono::duration<double, std::micro>(b - a).count();
        };
        float ab = 0.f, bc = 0.f;
        cudaEventRecord(ev_c, ctx->stream);
        cudaEventSynchronize(ev_c);
        cudaEventElapsedTime(&ab, ev_a, s->copied);
        cudaEventElapsedTime(&bc, s->copied, ev_c);
        cudaEventDestroy(ev_a);
        cudaEventDestroy(ev_c);
        fprintf(stderr, "[ingest] %d graphs %lld nnz, %d threads: plan %.0f us, wait staging %.0f us, pack %.0f us, enqueue %.0f us; "
                "device: copies %.0f us, kernels + copy-out %.0f us\n",
                n_graphs, (long long)e, threads, us(t_begin, t_plan), us(t_plan, t_wait), us(t_wait, t_pack), us(t_pack, t_end),
                1e3 * ab, 1e3 * bc);
    }
    return st;
}

}  // extern "C"
