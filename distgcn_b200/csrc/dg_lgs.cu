// Local greedy MWIS (synchronous-round distributed greedy) on a packed CSR batch.
//
// Reference: heuristics.py:77-116 (local_greedy_search) and the instrumented variants at
// heuristics.py:119-305.  Per round every vertex still in `remain` looks at its neighbours that are
// still in `remain` (round-start state: `remain` is rebound only after the vertex loop,
// heuristics.py:114) and joins the set iff it has none, or its weight is strictly larger than all of
// theirs, or it ties with the heaviest and its index is smaller than the smallest-index neighbour
// carrying that weight (heuristics.py:96-111).  That is exactly
//     join(v)  <=>  for all u in N(v) & remain:  w_v > w_u  or  (w_v == w_u and v < u)
// with IEEE-double comparisons (-0.0 == +0.0 is a tie).  Joined vertices and their remaining
// neighbours (`nb_is`) leave `remain`.  The rule only reads round-start state, so evaluating all
// vertices in parallel is the same computation as the reference's serial loop.
//
// `remain` and `joined` are bitmaps (frontier bitmaps); a warp owns one aligned 32-vertex word and
// assembles it with __ballot_sync, so no atomics touch the bitmaps.
//
//   small graphs (<= kLgsCtaMaxNodes vertices): one CTA per graph runs ALL rounds inside one launch
//       with the bitmaps in shared memory                                    lgs_cta_kernel
//   large graphs: one launch pair per round over all vertices with global bitmaps; the host reads
//       one int per round to detect convergence                              lgs_*_global kernels
#include <algorithm>

#include "dg_common.cuh"

namespace dg {

namespace {

constexpr int kLgsCtaThreads = 128;

__device__ __forceinline__ bool dominates(double wu, int u, double wv, int v) {
    // true when remaining neighbour u prevents v from joining: v needs w_v > w_u, or a tie that its
    // smaller index wins.  Written as a negation so that a NaN on either side blocks v, as
    // np.max / the float comparisons of heuristics.py:102-106 do.
    return !((wv > wu) || (wv == wu && v < u));
}

template <bool STATS>
__global__ void __launch_bounds__(kLgsCtaThreads)
lgs_cta_kernel(const int *__restrict__ graph_ptr, const int *__restrict__ row_ptr,
               const int *__restrict__ col_idx, const double *__restrict__ util,
               const uint8_t *__restrict__ keep, int nstep, int round_cap, int words_cap,
               uint8_t *__restrict__ member, uint8_t *__restrict__ nb_is, int *__restrict__ steps,
               long long *__restrict__ p2p, long long *__restrict__ bst, double *__restrict__ oh_vec,
               int *__restrict__ status) {
    extern __shared__ uint32_t lgs_words[];
    uint32_t *remain = lgs_words;
    uint32_t *joined = lgs_words + words_cap;
    __shared__ unsigned long long p2p_sm;
    __shared__ int member_sm;

    const int g = blockIdx.x;
    const int v0 = graph_ptr[g];
    const int n = graph_ptr[g + 1] - v0;
    const int lane = threadIdx.x & 31;
    const int span = ((n + 31) / 32) * 32;  // vertices rounded up to whole words

    if (threadIdx.x == 0) {
        p2p_sm = 0ull;
        member_sm = 0;
    }
    __syncthreads();
    int n_remain = 0;
    for (int base = 0; base < span; base += kLgsCtaThreads) {
        const int v = base + threadIdx.x;
        bool alive = false;
        if (v < n) {
            alive = keep ? keep[v0 + v] != 0 : true;
            member[v0 + v] = 0;
            if (nb_is) nb_is[v0 + v] = 0;
        }
        const uint32_t w = __ballot_sync(0xffffffffu, alive);
        if (lane == 0 && v < span) remain[v >> 5] = w;
        n_remain += __syncthreads_count(alive);
    }

    int rounds = 0;
    bool bad_col = false;
    long long bst_acc = 0;
    unsigned long long p2p_acc = 0;
    int joined_mine = 0;
    while (n_remain > 0 && (nstep < 0 || rounds < nstep)) {
        if (rounds >= round_cap) {
            if (threadIdx.x == 0) atomicExch(status, DG_ERR_NOT_CONVERGED);
            break;
        }
        bst_acc += n_remain;  // heuristics.py:179
        // ---- decide ------------------------------------------------------------------------------
        for (int base = 0; base < span; base += kLgsCtaThreads) {
            const int v = base + threadIdx.x;
            const bool active = v < span ? (remain[v >> 5] >> lane) & 1u : false;
            bool join = false;
            if (active) {
                const double wv = util[v0 + v];
                const int beg = row_ptr[v0 + v], end = row_ptr[v0 + v + 1];
                bool blocked = false;
                int cnt = 0;
                for (int e = beg; e < end; ++e) {
                    int u = col_idx[e] - v0;
                    // a column id outside its graph (malformed CSR) is reported and read as the vertex itself, never followed
                    const bool bad = (unsigned)u >= (unsigned)n;
                    bad_col |= bad;
                    u = bad ? v : u;
                    if ((remain[u >> 5] >> (u & 31)) & 1u) {
                        ++cnt;
                        if (dominates(util[v0 + u], u, wv, v)) {
                            blocked = true;
                            if (!STATS) break;
                        }
                    }
                }
                join = !blocked;
                if (join) {
                    member[v0 + v] = 1;
                    ++joined_mine;
                }
                if (STATS) {
                    p2p_acc += (unsigned)cnt;  // heuristics.py:185
                    // heuristics.py:238,249: neighbours heard from, +1 "mute" broadcast when joining
                    // with a non-empty neighbourhood
                    if (oh_vec) oh_vec[v0 + v] += (double)(cnt + ((join && cnt > 0) ? 1 : 0));
                }
            }
            const uint32_t jw = __ballot_sync(0xffffffffu, join);
            if (lane == 0 && v < span) joined[v >> 5] = jw;
        }
        __syncthreads();
        // ---- remove joined vertices and their remaining neighbours ---------------------------------
        n_remain = 0;
        for (int base = 0; base < span; base += kLgsCtaThreads) {
            const int v = base + threadIdx.x;
            bool still = false, excluded_v = false;
            if (v < span) {
                const bool active = (remain[v >> 5] >> lane) & 1u;
                const bool join = (joined[v >> 5] >> lane) & 1u;
                if (active && !join) {
                    const int beg = row_ptr[v0 + v], end = row_ptr[v0 + v + 1];
                    bool excluded = false;
                    for (int e = beg; e < end; ++e) {
                        int u = col_idx[e] - v0;
                        const bool bad = (unsigned)u >= (unsigned)n;
                        bad_col |= bad;
                        u = bad ? v : u;
                        if ((joined[u >> 5] >> (u & 31)) & 1u) {
                            excluded = true;
                            break;
                        }
                    }
                    excluded_v = excluded;
                    still = !excluded;
                }
            }
            const uint32_t rw = __ballot_sync(0xffffffffu, still);
            if (excluded_v && nb_is) nb_is[v0 + v] = 1;   // (after the warp has reconverged)
            __syncwarp();  // every lane has read this word (WAR within the warp)
            if (lane == 0 && v < span) remain[v >> 5] = rw;
            n_remain += __syncthreads_count(still);
        }
        ++rounds;
    }
    if (bad_col) atomicExch(status, DG_ERR_INVALID);
    if (STATS) {
        atomicAdd(&p2p_sm, p2p_acc);
        atomicAdd(&member_sm, joined_mine);
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        if (steps) steps[g] = rounds;
        if (STATS) {
            if (p2p) p2p[g] = (long long)p2p_sm;
            if (bst) bst[g] = bst_acc + member_sm;  // heuristics.py:208
        }
    }
}

// -------------------------------------------------------------------------------------------------
// Threshold distributed greedy (heuristics.py:38-74, dist_greedy_search), one CTA per graph.
// Per round: (1) seta = remaining vertices with no remaining neighbour, or whose weight reaches
// max(remaining neighbours' weights) / alpha (:54-63); (2) mis_i = the vertices of seta taken one
// after the other unless a neighbour was taken before (:64-69) - the reference walks seta in the
// iteration order of a Python set, here the order is ascending vertex id (the two agree whenever no
// two vertices of seta are adjacent), and that sequential scan is evaluated as
// the lexicographically-first maximal independent set of seta by parallel sub-rounds: an undecided
// vertex is dropped once a smaller-id neighbour is in, and joins once no smaller-id neighbour is
// undecided or in; (3) mis_i joins the result, mis_i and its neighbours leave `remain` (:70-71).
// A NaN among the weights compared keeps the vertex out of seta (np.max / >= semantics).
// -------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kLgsCtaThreads)
dgs_cta_kernel(const int *__restrict__ graph_ptr, const int *__restrict__ row_ptr, const int *__restrict__ col_idx,
               const double *__restrict__ wts, const uint8_t *__restrict__ keep, double alpha, int round_cap,
               int words_cap, uint8_t *__restrict__ member, int *__restrict__ steps, int *__restrict__ status) {
    extern __shared__ uint32_t lgs_words[];
    uint32_t *remain = lgs_words;
    uint32_t *seta = lgs_words + words_cap;       // this round's candidates
    uint32_t *und = lgs_words + 2 * words_cap;    // candidates not decided yet
    uint32_t *mis = lgs_words + 3 * words_cap;    // candidates taken this round

    const int g = blockIdx.x;
    const int v0 = graph_ptr[g];
    const int n = graph_ptr[g + 1] - v0;
    const int lane = threadIdx.x & 31;
    const int span = ((n + 31) / 32) * 32;
    auto bit = [](const uint32_t *w, int u) -> bool { return (w[u >> 5] >> (u & 31)) & 1u; };

    int n_remain = 0;
    for (int base = 0; base < span; base += kLgsCtaThreads) {
        const int v = base + threadIdx.x;
        bool alive = false;
        if (v < n) {
            alive = keep ? keep[v0 + v] != 0 : true;
            member[v0 + v] = 0;
        }
        const uint32_t w = __ballot_sync(0xffffffffu, alive);
        if (lane == 0 && v < span) remain[v >> 5] = w;
        n_remain += __syncthreads_count(alive);
    }

    int rounds = 0;
    while (n_remain > 0) {
        if (rounds >= round_cap) {
            if (threadIdx.x == 0) atomicExch(status, DG_ERR_NOT_CONVERGED);
            break;
        }
        // ---- (1) candidates ------------------------------------------------------------------------
        for (int base = 0; base < span; base += kLgsCtaThreads) {
            const int v = base + threadIdx.x;
            bool cand = false;
            if (v < span && bit(remain, v)) {
                const int beg = row_ptr[v0 + v], end = row_ptr[v0 + v + 1];
                bool any = false, has_nan = false;
                double w_bar = 0.0;
                for (int e = beg; e < end; ++e) {
                    const int u = col_idx[e] - v0;
                    if (!bit(remain, u)) continue;
                    const double wu = wts[v0 + u];
                    has_nan |= wu != wu;
                    if (!any || wu > w_bar) w_bar = wu;
                    any = true;
                }
                cand = !any || (!has_nan && wts[v0 + v] >= w_bar / alpha);
            }
            const uint32_t cw = __ballot_sync(0xffffffffu, cand);
            if (lane == 0 && v < span) {
                seta[v >> 5] = cw;
                und[v >> 5] = cw;
                mis[v >> 5] = 0u;
            }
        }
        __syncthreads();
        // ---- (2) ascending-id scan of seta as parallel sub-rounds -----------------------------------
        for (int sub = 0;; ++sub) {
            int left = 0;
            for (int base = 0; base < span; base += kLgsCtaThreads) {
                const int v = base + threadIdx.x;
                bool take = false, drop = false;
                if (v < span && bit(und, v)) {
                    const int beg = row_ptr[v0 + v], end = row_ptr[v0 + v + 1];
                    bool wait = false;
                    for (int e = beg; e < end; ++e) {
                        const int u = col_idx[e] - v0;
                        if (u >= v) continue;
                        if (bit(mis, u)) {
                            drop = true;
                            break;
                        }
                        wait |= bit(und, u);
                    }
                    take = !drop && !wait;
                }
                const uint32_t tw = __ballot_sync(0xffffffffu, take);
                const uint32_t dw = __ballot_sync(0xffffffffu, take || drop);
                // every thread of the CTA has read the round-start words before any of them changes
                const int undecided = __syncthreads_count(v < span && bit(und, v) && !(take || drop));
                if (lane == 0 && v < span) {
                    mis[v >> 5] |= tw;
                    und[v >> 5] &= ~dw;
                }
                left += undecided;
                __syncthreads();
            }
            if (left == 0) break;
            if (sub > n) {  // cannot happen: the smallest undecided vertex is decided in every sub-round
                if (threadIdx.x == 0) atomicExch(status, DG_ERR_NOT_CONVERGED);
                break;
            }
        }
        // ---- (3) the taken vertices join; they and their neighbours leave ---------------------------
        n_remain = 0;
        for (int base = 0; base < span; base += kLgsCtaThreads) {
            const int v = base + threadIdx.x;
            bool still = false;
            if (v < span && bit(remain, v)) {
                if (bit(mis, v)) {
                    member[v0 + v] = 1;
                } else {
                    const int beg = row_ptr[v0 + v], end = row_ptr[v0 + v + 1];
                    still = true;
                    for (int e = beg; e < end; ++e) {
                        if (bit(mis, col_idx[e] - v0)) {
                            still = false;
                            break;
                        }
                    }
                }
            }
            const uint32_t rw = __ballot_sync(0xffffffffu, still);
            __syncwarp();
            if (lane == 0 && v < span) remain[v >> 5] = rw;
            n_remain += __syncthreads_count(still);
        }
        ++rounds;
    }
    if (threadIdx.x == 0 && steps) steps[g] = rounds;
}

// -------------------------------------------------------------------------------------------------
// global path
// -------------------------------------------------------------------------------------------------
__device__ __forceinline__ int graph_of(const int *__restrict__ graph_ptr, int n_graphs, int v) {
    int lo = 0, hi = n_graphs;  // graph_ptr[lo] <= v < graph_ptr[hi]
    while (hi - lo > 1) {
        int mid = (lo + hi) >> 1;
        if (graph_ptr[mid] <= v) lo = mid; else hi = mid;
    }
    return lo;
}

__global__ void lgs_init_global(int n, int n_graphs, const int *__restrict__ graph_ptr,
                                const uint8_t *__restrict__ keep, uint32_t *__restrict__ remain,
                                uint8_t *__restrict__ member, uint8_t *__restrict__ nb_is,
                                double *__restrict__ oh_vec, int *__restrict__ cnt) {
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    bool alive = false;
    if (v < n) {
        alive = keep ? keep[v] != 0 : true;
        member[v] = 0;
        if (nb_is) nb_is[v] = 0;
        if (oh_vec) oh_vec[v] = 0.0;
    }
    const uint32_t w = __ballot_sync(0xffffffffu, alive);
    if (lane == 0 && (v >> 5) < (n + 31) / 32) remain[v >> 5] = w;
    if (n_graphs == 1) {  // one atomic per CTA: a giant graph would otherwise serialise on one address
        const int c = __syncthreads_count(alive);
        if (threadIdx.x == 0 && c) atomicAdd(&cnt[0], c);
    } else if (alive) {
        atomicAdd(&cnt[graph_of(graph_ptr, n_graphs, v)], 1);
    }
}

// per graph, at the start of a round: count the round if the graph still has remaining vertices
__global__ void lgs_round_begin_global(int n_graphs, int *__restrict__ cnt, int *__restrict__ steps,
                                       long long *__restrict__ bst, int *__restrict__ any_left) {
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n_graphs) return;
    const int c = cnt[g];
    if (c > 0) {
        if (steps) steps[g] += 1;
        if (bst) bst[g] += c;
        *any_left = 1;
    }
    cnt[g] = 0;
}

template <bool STATS>
__global__ void __launch_bounds__(256)
lgs_decide_global(int n, int n_graphs, const int *__restrict__ graph_ptr, const int *__restrict__ row_ptr,
                  const int *__restrict__ col_idx, const double *__restrict__ util,
                  const uint32_t *__restrict__ remain, uint32_t *__restrict__ joined,
                  uint8_t *__restrict__ member, long long *__restrict__ p2p, long long *__restrict__ bst,
                  double *__restrict__ oh_vec) {
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    const int n_words = (n + 31) / 32;
    const bool active = (v >> 5) < n_words ? (remain[v >> 5] >> lane) & 1u : false;
    bool join = false;
    int cnt = 0;
    if (active) {
        const double wv = util[v];
        const int beg = row_ptr[v], end = row_ptr[v + 1];
        bool blocked = false;
        for (int e = beg; e < end; ++e) {
            const int u = col_idx[e];
            if ((__ldg(remain + (u >> 5)) >> (u & 31)) & 1u) {
                ++cnt;
                if (dominates(util[u], u, wv, v)) {
                    blocked = true;
                    if (!STATS) break;
                }
            }
        }
        join = !blocked;
        if (join) member[v] = 1;
        if (STATS) {
            if (oh_vec) oh_vec[v] += (double)(cnt + ((join && cnt > 0) ? 1 : 0));
            const int g = n_graphs == 1 ? 0 : graph_of(graph_ptr, n_graphs, v);
            if (p2p && cnt) atomicAdd((unsigned long long *)&p2p[g], (unsigned long long)cnt);
            if (bst && join) atomicAdd((unsigned long long *)&bst[g], 1ull);  // the final + |mwis|
        }
    }
    const uint32_t jw = __ballot_sync(0xffffffffu, join);
    if (lane == 0 && (v >> 5) < n_words) joined[v >> 5] = jw;
}

__global__ void __launch_bounds__(256)
lgs_remove_global(int n, int n_graphs, const int *__restrict__ graph_ptr, const int *__restrict__ row_ptr,
                  const int *__restrict__ col_idx, const uint32_t *__restrict__ joined,
                  uint32_t *__restrict__ remain, uint8_t *__restrict__ nb_is, int *__restrict__ cnt) {
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    const int n_words = (n + 31) / 32;
    bool still = false, excluded_v = false;
    if ((v >> 5) < n_words) {
        const bool active = (remain[v >> 5] >> lane) & 1u;
        const bool join = (joined[v >> 5] >> lane) & 1u;
        if (active && !join) {
            const int beg = row_ptr[v], end = row_ptr[v + 1];
            bool excluded = false;
            for (int e = beg; e < end; ++e) {
                const int u = col_idx[e];
                if ((__ldg(joined + (u >> 5)) >> (u & 31)) & 1u) {
                    excluded = true;
                    break;
                }
            }
            excluded_v = excluded;
            still = !excluded;
        }
    }
    const uint32_t rw = __ballot_sync(0xffffffffu, still);
    if (excluded_v && nb_is) nb_is[v] = 1;   // (after the warp has reconverged)
    if (lane == 0 && (v >> 5) < n_words) remain[v >> 5] = rw;
    if (n_graphs == 1) {
        const int c = __syncthreads_count(still);
        if (threadIdx.x == 0 && c) atomicAdd(&cnt[0], c);
    } else if (still) {
        atomicAdd(&cnt[graph_of(graph_ptr, n_graphs, v)], 1);
    }
}

// ---- row-slice form (one giant graph partitioned by rows over several GPUs) ----------------------
// The launch covers rows row0 .. row0+n_local-1 (row0 and n_local multiples of 32, so a warp still owns
// one aligned bitmap word); row_ptr is the slice's local array, col_idx / util / bitmaps / member are
// indexed by GLOBAL vertex id.  Between the kernels the caller all-gathers the bitmap words it wrote.
__global__ void __launch_bounds__(256)
lgs_part_init(int n_local, int row0, int n_global, const uint8_t *__restrict__ keep, uint32_t *__restrict__ remain,
              uint8_t *__restrict__ member, long long *__restrict__ cnt, const PeerMap pm) {
    const int vl = blockIdx.x * blockDim.x + threadIdx.x;
    const int v = row0 + vl;
    const int lane = threadIdx.x & 31;
    bool alive = false;
    if (vl < n_local && v < n_global) {
        alive = keep ? keep[v] != 0 : true;
        member[v] = 0;
    }
    const uint32_t w = __ballot_sync(0xffffffffu, alive);
    if (lane == 0 && vl < n_local) peer_store(pm, remain + (v >> 5), w);
    const int c = __syncthreads_count(alive);
    if (threadIdx.x == 0 && c) atomicAdd((unsigned long long *)cnt, (unsigned long long)c);
}

__global__ void __launch_bounds__(256)
lgs_part_decide(int n_local, int row0, const int *__restrict__ row_ptr, const int *__restrict__ col_idx,
                const double *__restrict__ util, const uint32_t *__restrict__ remain,
                uint32_t *__restrict__ joined, uint8_t *__restrict__ member, const PeerMap pm) {
    const int vl = blockIdx.x * blockDim.x + threadIdx.x;
    const int v = row0 + vl;
    const int lane = threadIdx.x & 31;
    const bool active = vl < n_local ? (remain[v >> 5] >> lane) & 1u : false;
    bool join = false;
    if (active) {
        const double wv = util[v];
        const int beg = row_ptr[vl], end = row_ptr[vl + 1];
        join = true;
        for (int e = beg; e < end; ++e) {
            const int u = col_idx[e];
            if ((__ldg(remain + (u >> 5)) >> (u & 31)) & 1u) {
                if (dominates(util[u], u, wv, v)) {
                    join = false;
                    break;
                }
            }
        }
        if (join) member[v] = 1;
    }
    const uint32_t jw = __ballot_sync(0xffffffffu, join);
    if (lane == 0 && vl < n_local) peer_store(pm, joined + (v >> 5), jw);
}

__global__ void __launch_bounds__(256)
lgs_part_remove(int n_local, int row0, const int *__restrict__ row_ptr, const int *__restrict__ col_idx,
                const uint32_t *__restrict__ joined, uint32_t *__restrict__ remain, long long *__restrict__ cnt,
                const PeerMap pm) {
    const int vl = blockIdx.x * blockDim.x + threadIdx.x;
    const int v = row0 + vl;
    const int lane = threadIdx.x & 31;
    bool still = false;
    if (vl < n_local) {
        const bool active = (remain[v >> 5] >> lane) & 1u;
        const bool join = (joined[v >> 5] >> lane) & 1u;
        if (active && !join) {
            const int beg = row_ptr[vl], end = row_ptr[vl + 1];
            still = true;
            for (int e = beg; e < end; ++e) {
                const int u = col_idx[e];
                if ((__ldg(joined + (u >> 5)) >> (u & 31)) & 1u) {
                    still = false;
                    break;
                }
            }
        }
    }
    const uint32_t rw = __ballot_sync(0xffffffffu, still);
    __syncwarp();
    if (lane == 0 && vl < n_local) peer_store(pm, remain + (v >> 5), rw);
    const int c = __syncthreads_count(still);
    if (threadIdx.x == 0 && c) atomicAdd((unsigned long long *)cnt, (unsigned long long)c);
}

// Cross-rank barrier on the stream (one CTA): thread r publishes this rank's count into rank r's arena, raises
// this rank's flag there (release, system scope), then waits until rank r has raised its flag here (acquire).
// Everything the ranks stored to each other's arenas in earlier kernels of their streams is visible after
// it.  A rank that never arrives would spin forever, so the wait gives up after ~4 s and reports through *status.
__global__ void peer_barrier_kernel(const PeerMap pm, unsigned long long flags_off, unsigned epoch,
                                    const long long *__restrict__ count_src, unsigned long long counts_off,
                                    int *__restrict__ status) {
    const int r = threadIdx.x;
    if (r >= pm.world) return;
    if (count_src) {
        long long *slot = reinterpret_cast<long long *>(pm.base[r] + counts_off) + pm.rank;
        *slot = *count_src;
    }
    __threadfence_system();
    unsigned *remote = reinterpret_cast<unsigned *>(pm.base[r] + flags_off) + pm.rank;
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(remote), "r"(epoch) : "memory");
    const unsigned *local = reinterpret_cast<const unsigned *>(pm.base[pm.rank] + flags_off) + r;
    const long long t0 = clock64();
    for (;;) {
        unsigned seen;
        asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(seen) : "l"(local) : "memory");
        if ((int)(seen - epoch) >= 0) break;
        if (clock64() - t0 > 8000000000LL) {  // ~4 s at 2 GHz
            atomicExch(status, DG_ERR_CUDA);
            break;
        }
        __nanosleep(200);
    }
}

// total[g] = sum of wts over the members of graph g.  Graphs are cut into `splits` equal slices, CTA (g, s) sums
// slice s with a fixed tree and a second pass adds the slices in order: deterministic, and a giant graph is
// spread over the whole GPU instead of one CTA.
__global__ void __launch_bounds__(256)
member_weight_kernel(const int *__restrict__ graph_ptr, const uint8_t *__restrict__ member,
                     const double *__restrict__ wts, int splits, double *__restrict__ out) {
    __shared__ double part[256];
    const int g = blockIdx.x / splits, sidx = blockIdx.x - g * splits;
    const int v0 = graph_ptr[g], v1 = graph_ptr[g + 1];
    const int per = (v1 - v0 + splits - 1) / splits;
    const int a = v0 + sidx * per, b = min(v1, a + per);
    double acc = 0.0;
    for (int v = a + threadIdx.x; v < b; v += blockDim.x)
        if (member[v]) acc += wts[v];
    part[threadIdx.x] = acc;
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
        if (threadIdx.x < s) part[threadIdx.x] += part[threadIdx.x + s];
        __syncthreads();
    }
    if (threadIdx.x == 0) out[blockIdx.x] = part[0];
}

__global__ void member_weight_combine_kernel(int n_graphs, int splits, const double *__restrict__ partial,
                                             double *__restrict__ total) {
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n_graphs) return;
    double acc = 0.0;
    for (int s = 0; s < splits; ++s) acc += partial[(size_t)g * splits + s];
    total[g] = acc;
}

}  // namespace

int lgs_device(dg_context *ctx, const dg_batch *b, const double *util, int nstep, uint8_t *member,
               uint8_t *nb_is, int32_t *steps, int64_t *p2p, int64_t *bst, double *oh_vec) {
    const int n = b->n_nodes;
    const int G = b->n_graphs;
    if (G == 0) return DG_OK;
    cudaStream_t st = ctx->stream;
    const bool stats = (p2p != nullptr) || (bst != nullptr) || (oh_vec != nullptr);
    int *d_flag = nullptr;  // [1] any_left
    DG_TRY(scratch_as(ctx, kSlotLgsCount, (size_t)G + 4, &d_flag));
    int *d_cnt = d_flag + 4;

    if (b->max_graph_nodes <= kLgsCtaMaxNodes) {
        if (oh_vec && n) DG_CUDA_CHECK(cudaMemsetAsync(oh_vec, 0, sizeof(double) * (size_t)n, st));
        const int words_cap = (b->max_graph_nodes + 31) / 32 + 1;
        const size_t smem = sizeof(uint32_t) * 2 * (size_t)words_cap;
        if (stats) {
            lgs_cta_kernel<true><<<G, kLgsCtaThreads, smem, st>>>(
                b->graph_ptr, b->row_ptr, b->col_idx, util, b->keep, nstep, kLgsRoundCap, words_cap, member,
                nb_is, steps, (long long *)p2p, (long long *)bst, oh_vec, ctx->d_status);
        } else {
            lgs_cta_kernel<false><<<G, kLgsCtaThreads, smem, st>>>(
                b->graph_ptr, b->row_ptr, b->col_idx, util, b->keep, nstep, kLgsRoundCap, words_cap, member,
                nb_is, steps, nullptr, nullptr, nullptr, ctx->d_status);
        }
        ctx->launches++;
        DG_CUDA_CHECK(cudaGetLastError());
        // The status word is read back asynchronously and examined by the next synchronising call
        // (dg_context_synchronize or any HOST-space call): a non-converged graph can only come from
        // NaN utilities or self-loops, which the reference does not survive either.
        return DG_OK;
    }

    // ---- global path --------------------------------------------------------------------------
    const int n_words = (n + 31) / 32;
    uint32_t *words = nullptr;
    DG_TRY(scratch_as(ctx, kSlotLgsWords, (size_t)2 * n_words + 2, &words));
    uint32_t *remain = words, *joined = words + n_words;
    DG_CUDA_CHECK(cudaMemsetAsync(d_cnt, 0, sizeof(int) * (size_t)G, st));
    if (steps) DG_CUDA_CHECK(cudaMemsetAsync(steps, 0, sizeof(int32_t) * (size_t)G, st));
    if (p2p) DG_CUDA_CHECK(cudaMemsetAsync(p2p, 0, sizeof(int64_t) * (size_t)G, st));
    if (bst) DG_CUDA_CHECK(cudaMemsetAsync(bst, 0, sizeof(int64_t) * (size_t)G, st));
    const int vgrid = (n_words * 32 + 255) / 256;
    lgs_init_global<<<vgrid, 256, 0, st>>>(n, G, b->graph_ptr, b->keep, remain, member, nb_is, oh_vec, d_cnt);
    ctx->launches++;
    int rounds = 0;
    while (nstep < 0 || rounds < nstep) {
        DG_REQUIRE(rounds < kLgsRoundCap, DG_ERR_NOT_CONVERGED,
                   "local greedy search hit the round cap (NaN utilities or self-loops?)");
        DG_CUDA_CHECK(cudaMemsetAsync(d_flag + 1, 0, sizeof(int), st));
        lgs_round_begin_global<<<(G + 255) / 256, 256, 0, st>>>(G, d_cnt, steps, (long long *)bst, d_flag + 1);
        ctx->launches++;
        DG_CUDA_CHECK(cudaMemcpyAsync(ctx->h_flag, d_flag + 1, sizeof(int), cudaMemcpyDeviceToHost, st));
        DG_CUDA_CHECK(cudaStreamSynchronize(st));
        if (ctx->h_flag[0] == 0) break;
        if (stats) {
            lgs_decide_global<true><<<vgrid, 256, 0, st>>>(n, G, b->graph_ptr, b->row_ptr, b->col_idx, util,
                                                           remain, joined, member, (long long *)p2p,
                                                           (long long *)bst, oh_vec);
        } else {
            lgs_decide_global<false><<<vgrid, 256, 0, st>>>(n, G, b->graph_ptr, b->row_ptr, b->col_idx, util,
                                                            remain, joined, member, nullptr, nullptr, nullptr);
        }
        lgs_remove_global<<<vgrid, 256, 0, st>>>(n, G, b->graph_ptr, b->row_ptr, b->col_idx, joined, remain,
                                                 nb_is, d_cnt);
        ctx->launches += 2;
        DG_CUDA_CHECK(cudaGetLastError());
        ++rounds;
    }
    return DG_OK;
}

int dist_greedy_device(dg_context *ctx, const dg_batch *b, const double *wts, double alpha, uint8_t *member,
                       int32_t *steps) {
    const int G = b->n_graphs;
    if (G == 0) return DG_OK;
    // One CTA per graph with its four bitmaps in shared memory: 4 n / 8 bytes, so the opt-in limit of an SM (227 KB) holds
    // graphs of up to ~460 k vertices - far above the one-CTA limit of the local greedy search, whose per-round state is
    // larger.  (A graph that size runs on ONE SM: the ascending-id scan of a round's candidates is inherently a chain.)
    const int words_cap = (b->max_graph_nodes + 31) / 32 + 1;
    const size_t smem = sizeof(uint32_t) * 4 * (size_t)words_cap;
    DG_REQUIRE(smem <= (size_t)ctx->max_smem_optin, DG_ERR_UNSUPPORTED,
               "dist_greedy_search runs one CTA per graph: a graph of %d vertices needs %zu bytes of shared memory (limit %d, "
               "about %d vertices)", b->max_graph_nodes, smem, ctx->max_smem_optin, ctx->max_smem_optin * 2 - 64);
    static std::atomic<unsigned long long> attr_done{0};
    if (smem > 48 * 1024) DG_CUDA_CHECK(smem_attr_once(dgs_cta_kernel, ctx->device, ctx->max_smem_optin, &attr_done));
    dgs_cta_kernel<<<G, kLgsCtaThreads, smem, ctx->stream>>>(b->graph_ptr, b->row_ptr, b->col_idx, wts, b->keep, alpha,
                                                             kLgsRoundCap, words_cap, member, steps, ctx->d_status);
    ctx->launches++;
    DG_CUDA_CHECK(cudaGetLastError());
    return DG_OK;
}

int part_lgs_init(dg_context *ctx, const PartView &pv, const uint8_t *keep, uint32_t *remain, uint8_t *member,
                  long long *cnt) {
    if (pv.n_local == 0) return DG_OK;
    lgs_part_init<<<(pv.n_local + 255) / 256, 256, 0, ctx->stream>>>(pv.n_local, pv.row0, pv.n_global, keep, remain,
                                                                    member, cnt, pv.pm);
    ctx->launches++;
    DG_CUDA_CHECK(cudaGetLastError());
    return DG_OK;
}

int part_lgs_decide(dg_context *ctx, const PartView &pv, const double *util, const uint32_t *remain, uint32_t *joined,
                    uint8_t *member) {
    if (pv.n_local == 0) return DG_OK;
    lgs_part_decide<<<(pv.n_local + 255) / 256, 256, 0, ctx->stream>>>(pv.n_local, pv.row0, pv.row_ptr, pv.col_idx,
                                                                      util, remain, joined, member, pv.pm);
    ctx->launches++;
    DG_CUDA_CHECK(cudaGetLastError());
    return DG_OK;
}

int part_lgs_remove(dg_context *ctx, const PartView &pv, const uint32_t *joined, uint32_t *remain, long long *cnt) {
    if (pv.n_local == 0) return DG_OK;
    lgs_part_remove<<<(pv.n_local + 255) / 256, 256, 0, ctx->stream>>>(pv.n_local, pv.row0, pv.row_ptr, pv.col_idx,
                                                                      joined, remain, cnt, pv.pm);
    ctx->launches++;
    DG_CUDA_CHECK(cudaGetLastError());
    return DG_OK;
}

// ---- GCN embedded into the greedy iteration (mwis_gdpg_call.py:278-318), generic path --------------------
// flag[g] = 1 when graph g still has a residual vertex with a positive weight
__global__ void dit_flag_kernel(int n, int n_graphs, const int *__restrict__ graph_ptr, const uint8_t *__restrict__ keep,
                                const double *__restrict__ wts, int *__restrict__ flag) {
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= n) return;
    if (keep[v] && wts[v] > 0.0) flag[n_graphs == 1 ? 0 : graph_of(graph_ptr, n_graphs, v)] = 1;
}

// graphs without such a vertex stop: their residual vertices leave (:296-297); counts what is left
__global__ void dit_filter_kernel(int n, int n_graphs, const int *__restrict__ graph_ptr, const int *__restrict__ flag,
                                  uint8_t *__restrict__ keep, int *__restrict__ any_left) {
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    bool k = false;
    if (v < n && keep[v]) {
        k = flag[n_graphs == 1 ? 0 : graph_of(graph_ptr, n_graphs, v)] != 0;
        if (!k) keep[v] = 0;
    }
    if (__syncthreads_or(k) && threadIdx.x == 0) *any_left = 1;
}

// after one greedy round: taken vertices accumulate, taken and excluded ones leave the residual graph
__global__ void dit_update_kernel(int n, const uint8_t *__restrict__ joined, const uint8_t *__restrict__ nb_is,
                                  uint8_t *__restrict__ member, uint8_t *__restrict__ keep) {
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= n) return;
    if (joined[v]) member[v] = 1;
    if (joined[v] || nb_is[v]) keep[v] = 0;
}

int dit_filter_device(dg_context *ctx, const dg_batch *b, const double *wts, uint8_t *keep, int *flag, int *any_left) {
    const int n = b->n_nodes, G = b->n_graphs;
    if (n == 0) return DG_OK;
    DG_CUDA_CHECK(cudaMemsetAsync(flag, 0, sizeof(int) * (size_t)G, ctx->stream));
    DG_CUDA_CHECK(cudaMemsetAsync(any_left, 0, sizeof(int), ctx->stream));
    dit_flag_kernel<<<(n + 255) / 256, 256, 0, ctx->stream>>>(n, G, b->graph_ptr, keep, wts, flag);
    dit_filter_kernel<<<(n + 255) / 256, 256, 0, ctx->stream>>>(n, G, b->graph_ptr, flag, keep, any_left);
    ctx->launches += 2;
    DG_CUDA_CHECK(cudaGetLastError());
    return DG_OK;
}

__global__ void dit_steps_kernel(int n_graphs, const int *__restrict__ flag, int *__restrict__ steps) {
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g < n_graphs && flag[g]) steps[g] += 1;
}

int dit_steps_device(dg_context *ctx, int n_graphs, const int *flag, int *steps) {
    if (n_graphs == 0 || !steps) return DG_OK;
    dit_steps_kernel<<<(n_graphs + 255) / 256, 256, 0, ctx->stream>>>(n_graphs, flag, steps);
    ctx->launches++;
    DG_CUDA_CHECK(cudaGetLastError());
    return DG_OK;
}

int dit_update_device(dg_context *ctx, int n, const uint8_t *joined, const uint8_t *nb_is, uint8_t *member,
                      uint8_t *keep) {
    if (n == 0) return DG_OK;
    dit_update_kernel<<<(n + 255) / 256, 256, 0, ctx->stream>>>(n, joined, nb_is, member, keep);
    ctx->launches++;
    DG_CUDA_CHECK(cudaGetLastError());
    return DG_OK;
}

// one thread: a round is counted when some rank still had a vertex left at the last barrier; this rank's count restarts
__global__ void part_tally_kernel(const long long *__restrict__ counts, int world, long long *__restrict__ own_count,
                                  int *__restrict__ rounds) {
    long long left = 0;
    for (int r = 0; r < world; ++r) left += counts[r];
    if (left > 0) *rounds += 1;
    *own_count = 0;
}

int part_tally(dg_context *ctx, const long long *counts, int world, long long *own_count, int *rounds) {
    part_tally_kernel<<<1, 1, 0, ctx->stream>>>(counts, world, own_count, rounds);
    ctx->launches++;
    DG_CUDA_CHECK(cudaGetLastError());
    return DG_OK;
}

int part_barrier(dg_context *ctx, const PeerMap &pm, unsigned long long flags_off, unsigned epoch,
                 const long long *count_src, unsigned long long counts_off) {
    if (pm.world <= 1) return DG_OK;
    peer_barrier_kernel<<<1, 32, 0, ctx->stream>>>(pm, flags_off, epoch, count_src, counts_off, ctx->d_status);
    ctx->launches++;
    DG_CUDA_CHECK(cudaGetLastError());
    return DG_OK;
}

int member_weight_device(dg_context *ctx, const dg_batch *b, const uint8_t *member, const double *wts,
                         double *total) {
    if (b->n_graphs == 0) return DG_OK;
    // one slice per 16 Ki vertices of the largest graph, bounded so that the grid stays small for large batches
    long long splits = (b->max_graph_nodes + 16383) / 16384;
    splits = std::max<long long>(1, std::min<long long>(splits, 4096));
    while (splits > 1 && splits * b->n_graphs > (1LL << 22)) splits /= 2;
    if (splits == 1) {
        member_weight_kernel<<<b->n_graphs, 256, 0, ctx->stream>>>(b->graph_ptr, member, wts, 1, total);
        ctx->launches++;
    } else {
        double *partial = nullptr;
        DG_TRY(scratch_as(ctx, kSlotPartial, (size_t)splits * b->n_graphs, &partial));
        member_weight_kernel<<<(unsigned)(splits * b->n_graphs), 256, 0, ctx->stream>>>(b->graph_ptr, member, wts,
                                                                                       (int)splits, partial);
        member_weight_combine_kernel<<<(b->n_graphs + 255) / 256, 256, 0, ctx->stream>>>(b->n_graphs, (int)splits,
                                                                                        partial, total);
        ctx->launches += 2;
    }
    DG_CUDA_CHECK(cudaGetLastError());
    return DG_OK;
}

}  // namespace dg
