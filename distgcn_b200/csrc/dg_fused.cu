// Graph-resident fused solver: one CTA owns a tile of consecutive small graphs and runs the WHOLE hot
// path for them inside one launch - zero-weight removal, degrees, the rank-1 first layer, every hidden
// GraphConvolution layer, the one-column last layer, the fp64 utility product and all rounds of the
// local greedy search - with the feature matrix, the CSR pattern and the frontier bitmaps resident in
// shared memory.  HBM traffic per graph is the compulsory minimum: CSR + weights in, membership out.
//
// Why: the per-layer streaming kernel (dg_gcn.cu) gathers ~deg neighbour rows of 128 B per output
// row from L2; on the reference's 100-300 vertex graphs that gather (not HBM, not FP32) bounds it
// (profiles/r01_notes.md).  A whole graph (<= 512 x 32 fp32 = 64 KB) fits in one SM's shared memory,
// so the gather becomes shared-memory traffic and nothing is written back between layers.
//
// Reference semantics are those of dg_gcn.cu / dg_lgs.cu (gcn/layers.py:198-216,
// mwis_dqn_call.py:198-261, heuristics.py:77-116); this file only changes where the data lives.
//
// Shared-memory feature layout: row i starts at float offset i * (CP + 4).  The 16-byte pad per row
// makes both access patterns conflict-free without any index arithmetic beyond one multiply-add:
//   gather     - CP/4 lanes read the CP/4 consecutive chunks of ONE row   (a contiguous 128/256 B run)
//   projection - the 8 lanes of a quarter-warp read the SAME chunk of 8 consecutive rows: the row
//                stride of 36 (68) words puts them 4 banks apart.
// Hidden-layer weights stream in through the TMA engine (cp.async.bulk + mbarrier), double buffered,
// one layer ahead of the arithmetic.
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include <vector>

#include <algorithm>
#include <functional>

#include <chrono>

#include "dg_common.cuh"

namespace dg {

namespace {

constexpr int kThreads = 512;
constexpr int kWarps = kThreads / 32;
static_assert(kWarps == 16, "the row ownership pattern of the layer loop assumes 16 warps");

__device__ __forceinline__ float act_apply(float v, int act, float alpha) {
    if (act == DG_ACT_LEAKY_RELU) return v >= 0.f ? v : alpha * v;
    if (act == DG_ACT_RELU) return fmaxf(v, 0.f);
    return v;
}

// ---- mbarrier / bulk-copy helpers (PTX) ----------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_mbar_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    uint32_t done = 0;
    while (!done) {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(done)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
    }
}
// global -> shared bulk copy on the TMA engine; completion is signalled on `bar` as transaction bytes
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

}  // namespace

// ---- shared-memory plan, computed identically on host and device ----------------------------------
__host__ __device__ inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

__host__ __device__ inline FusedSmemPlan fused_smem_plan(int cp, int cap_n, int cap_nnz, bool has_hidden,
                                                        size_t wblob_bytes) {
    FusedSmemPlan p;
    size_t off = 0;
    // every per-vertex array has cap_n + 1 entries: slot cap_n is the all-zero dummy vertex that pads
    // each row's neighbour list to a multiple of 4 (see "stage the tile")
    p.feat_a = off;
    off += sizeof(float) * (size_t)(cap_n + 1) * (cp + 4);
    p.feat_b = off;
    if (has_hidden) off += sizeof(float) * (size_t)(cap_n + 1) * (cp + 4);
    p.wbuf = off;
    p.wblob_bytes = wblob_bytes;  // one hidden layer's weights; single buffer, refilled during the aggregation
    if (has_hidden) off += align_up(p.wblob_bytes, 16);
    p.util = off = align_up(off, 8);
    off += sizeof(double) * (size_t)cap_n;
    p.mbar = off;
    off += 2 * sizeof(uint64_t);
    p.dinv = off;
    off += sizeof(float) * (size_t)(cap_n + 1);
    p.sa = off;
    off += sizeof(float) * (size_t)(cap_n + 1);
    p.sb = off;
    off += sizeof(float) * (size_t)(cap_n + 1);
    p.x0s = off;
    off += sizeof(float) * (size_t)(cap_n + 1);
    p.rp = off;
    off += sizeof(int) * (size_t)(cap_n + 2);
    p.words = off;
    p.n_words = cap_n / 32 + 1;
    off += sizeof(uint32_t) * 4 * (size_t)p.n_words;
    p.gstart = off;
    off += sizeof(int) * (kFusedMaxTileGraphs + 1);
    p.gcnt = off;
    off += sizeof(int) * kFusedMaxTileGraphs;
    p.gsteps = off;
    off += sizeof(int) * kFusedMaxTileGraphs;
    p.gpos = off;
    off += sizeof(int) * kFusedMaxTileGraphs;
    p.col16 = off = align_up(off, 8);
    off += sizeof(uint16_t) * (size_t)cap_nnz;
    p.gid = off = align_up(off, 4);
    off += (size_t)cap_n + 4;
    p.vid = off = align_up(off, 4);
    off += sizeof(uint16_t) * (size_t)(cap_n + 2);
    p.slotof = off = align_up(off, 4);
    off += sizeof(uint16_t) * (size_t)(cap_n + 2);
    p.total = align_up(off, 16);
    return p;
}

namespace {

template <int CP>
__device__ __forceinline__ int swz(int row, int chunk) {
    // float offset of 16-byte chunk `chunk` of feature row `row` (padded row stride, see header)
    return row * (CP + 4) + (chunk << 2);
}

// DIT: the iterative variant (P.dit) is its own instantiation so that the one-shot kernel keeps its register
// allocation (the layer loop runs at the 128-register limit)
template <int CP, bool DIT, bool MMA>
__global__ void __launch_bounds__(kThreads, 1) fused_solve_kernel(const FusedParams P) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const bool has_hidden = P.n_layers >= 3;
    const FusedSmemPlan plan = fused_smem_plan(CP, P.cap_n, P.cap_nnz, has_hidden, (size_t)P.wblob_bytes);
    float *fa = reinterpret_cast<float *>(smem_raw + plan.feat_a);
    float *fb = reinterpret_cast<float *>(smem_raw + plan.feat_b);
    float *wbuf = reinterpret_cast<float *>(smem_raw + plan.wbuf);
    double *util_sm = reinterpret_cast<double *>(smem_raw + plan.util);
    uint64_t *mbar = reinterpret_cast<uint64_t *>(smem_raw + plan.mbar);
    float *dinv = reinterpret_cast<float *>(smem_raw + plan.dinv);
    float *sa = reinterpret_cast<float *>(smem_raw + plan.sa);
    float *sb = reinterpret_cast<float *>(smem_raw + plan.sb);
    float *x0s = reinterpret_cast<float *>(smem_raw + plan.x0s);
    int *rp = reinterpret_cast<int *>(smem_raw + plan.rp);
    uint32_t *keepw = reinterpret_cast<uint32_t *>(smem_raw + plan.words);
    uint32_t *remain = keepw + plan.n_words;
    uint32_t *joined = remain + plan.n_words;
    uint32_t *memb = joined + plan.n_words;
    int *gstart = reinterpret_cast<int *>(smem_raw + plan.gstart);
    int *gcnt = reinterpret_cast<int *>(smem_raw + plan.gcnt);
    int *gsteps = reinterpret_cast<int *>(smem_raw + plan.gsteps);
    int *gpos = reinterpret_cast<int *>(smem_raw + plan.gpos);
    uint16_t *col16 = reinterpret_cast<uint16_t *>(smem_raw + plan.col16);
    uint8_t *gid = reinterpret_cast<uint8_t *>(smem_raw + plan.gid);
    uint16_t *vid = reinterpret_cast<uint16_t *>(smem_raw + plan.vid);        // slot -> local vertex id
    uint16_t *slotof = reinterpret_cast<uint16_t *>(smem_raw + plan.slotof);  // local vertex id -> slot
    __shared__ int tile_sm;

    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int warp = tid >> 5;
    constexpr int CHUNKS = CP / 4;         // 16-byte chunks per feature row
    constexpr int LPR = CHUNKS;            // lanes covering one row in the gather
    constexpr int NG = 32 / LPR;           // neighbour rows in flight per gather step
    const uint32_t wblob_bytes = (uint32_t)plan.wblob_bytes;
    const int n_hidden = has_hidden ? P.n_layers - 2 : 0;

    if (tid == 0) {
        mbar_init(&mbar[0], 1);
        fence_mbar_init();
    }
    uint32_t wuse = 0;  // completed fills of the weight buffer (-> mbarrier phase parity)
    __syncthreads();
    // phase timers of thread 0 (only when P.dbg != nullptr): 0 stage, 1 gather, 2 wait after gather,
    // 3 weight wait, 4 projection, 5 wait after projection, 6 tail+score, 7 greedy rounds, 8 tiles, 9 total
    long long tm[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    const bool timing = P.dbg != nullptr && tid == 0;
    long long tk = timing ? clock64() : 0;
    const long long t_begin = tk;
#define DG_TICK(slot)                     \
    if (timing) {                         \
        const long long _now = clock64(); \
        tm[slot] += _now - tk;            \
        tk = _now;                        \
    }

    // the dummy vertex (slot cap_n): zero feature rows, dinv 0, every bitmap bit 0 - set once
    const int dummy = P.cap_n;
    if (tid < CP + 4) {
        fa[(size_t)dummy * (CP + 4) + tid] = 0.f;
        if (has_hidden) fb[(size_t)dummy * (CP + 4) + tid] = 0.f;
    }
    if (tid == 0) {
        dinv[dummy] = 0.f;
        sa[dummy] = 0.f;
        sb[dummy] = 0.f;
        x0s[dummy] = 0.f;
        util_sm[0] = 0.0;
    }
    for (int wd = tid; wd < 4 * plan.n_words; wd += kThreads) keepw[wd] = 0u;
    __shared__ int scan_sm[kWarps + 1];
    __syncthreads();

    for (;;) {
        if (tid == 0) tile_sm = atomicAdd(P.tile_counter, 1);
        __syncthreads();
        const int t = tile_sm;
        if (t >= P.n_tiles) break;
        const int *td = P.tiles + (size_t)t * 8;
        const int v0 = td[0], n = td[1], e0 = td[2], g0 = td[4], ng = td[5];
        const long long t_tile = timing ? clock64() : 0;
        const int span = ((n + 31) / 32) * 32;

        // weights of the first hidden layer start streaming in right away
        if (has_hidden && tid == 0) {
            mbar_expect_tx(&mbar[0], wblob_bytes);
            bulk_g2s(wbuf, P.wall, wblob_bytes, &mbar[0]);  // layer 1's weights; later layers: see the layer loop
        }

        // ---- 1. stage the tile -----------------------------------------------------------------------
        // The tile's vertices are renumbered into SLOTS sorted by descending (padded) degree; features,
        // bitmaps and the local CSR live in slot space.  Eight consecutive slots - one warp's concurrent
        // rows in the aggregation - then have near-equal degree (no idle lane groups), and dealing the
        // 8-slot groups to the warps in snake order balances the hubs.  Vertex ids only matter for
        // the greedy tie-break and the outputs (vid[]).
        // Every row's neighbour list is padded to a multiple of 4 entries pointing at the dummy vertex,
        // whose rows / weights / bits are all zero, so no loop below needs a bounds predicate and the
        // aggregation fetches 4 column ids with one 8-byte load.
        for (int g = tid; g <= ng; g += kThreads) gstart[g] = P.graph_ptr[g0 + g] - v0;
        for (int g = tid; g < ng; g += kThreads) {
            gcnt[g] = 0;
            gsteps[g] = 0;
        }
        int *key = reinterpret_cast<int *>(sb);  // scratch until stage 3
        for (int i = tid; i < n; i += kThreads) {
            const int len4 = (P.row_ptr[v0 + i + 1] - P.row_ptr[v0 + i] + 3) >> 2;
            key[i] = (len4 << 16) | (0xffff - i);  // unique: larger degree first, then smaller id
        }
        __syncthreads();
        for (int i = tid; i < n; i += kThreads) {  // rank by counting (deterministic, n <= a few hundred)
            const int mine = key[i];
            int r = 0;
            for (int j2 = 0; j2 < n; ++j2) r += key[j2] > mine;
            slotof[i] = (uint16_t)r;
            vid[r] = (uint16_t)i;
            int lo = 0, hi = ng;  // gstart[lo] <= i < gstart[hi]: the tile-local graph of vertex i
            while (hi - lo > 1) {
                const int mid = (lo + hi) >> 1;
                if (gstart[mid] <= i) lo = mid; else hi = mid;
            }
            gid[r] = (uint8_t)lo;
        }
        __syncthreads();
        {   // exclusive scan of the padded row lengths in slot order -> rp[0..n]
            int carry = 0;
            for (int base = 0; base < n; base += kThreads) {
                const int sl = base + tid;
                int len = 0;
                if (sl < n) len = (key[vid[sl]] >> 16) << 2;
                int incl = len;
#pragma unroll
                for (int off = 1; off < 32; off <<= 1) {
                    const int up = __shfl_up_sync(0xffffffffu, incl, off);
                    if (lane >= off) incl += up;
                }
                if (lane == 31) scan_sm[warp] = incl;
                __syncthreads();
                if (warp == 0) {
                    int ws = lane < kWarps ? scan_sm[lane] : 0;
#pragma unroll
                    for (int off = 1; off < kWarps; off <<= 1) {
                        const int up = __shfl_up_sync(0xffffffffu, ws, off);
                        if (lane >= off) ws += up;
                    }
                    if (lane < kWarps) scan_sm[lane] = ws;  // inclusive over warps
                }
                __syncthreads();
                const int warp_off = warp > 0 ? scan_sm[warp - 1] : 0;
                if (sl < n) rp[sl] = carry + warp_off + incl - len;
                carry += scan_sm[kWarps - 1];
                __syncthreads();
            }
            if (tid == 0) rp[n] = carry;
        }
        for (int base = 0; base < span; base += kThreads) {
            const int sl = base + tid;
            bool k = false;
            if (sl < n) {
                const int vtx = v0 + vid[sl];
                k = P.keep_in ? P.keep_in[vtx] != 0 : true;
                if (P.remove_zero) k = k && (P.wts[vtx] != 0.0);  // mwis_dqn_call.py:203
                if (P.member) P.member[vtx] = 0;
            }
            const uint32_t w = __ballot_sync(0xffffffffu, k);
            if (lane == 0 && sl < span) {
                keepw[sl >> 5] = w;
                memb[sl >> 5] = 0u;
            }
        }
        __syncthreads();
        {   // flat copy of the padded lists: every thread owns padded entries k, k + 512, ... and finds
            // its slot by binary search in rp (all loads independent)
            const int total = rp[n];
            for (int k = tid; k < total; k += kThreads) {
                int lo = 0, hi = n;  // rp[lo] <= k < rp[hi]
                while (hi - lo > 1) {
                    const int mid = (lo + hi) >> 1;
                    if (rp[mid] <= k) lo = mid; else hi = mid;
                }
                const int vtx = v0 + vid[lo];
                const int gb = P.row_ptr[vtx], ge = P.row_ptr[vtx + 1];
                const int src = gb + (k - rp[lo]);
                col16[k] = src < ge ? slotof[P.col_idx[src] - v0] : (uint16_t)dummy;
            }
        }
        (void)e0;
        __syncthreads();

        // With P.dit the stages 2-7 repeat: the residual graph (vertices neither taken nor excluded yet) is
        // re-scored by the network before every single greedy round (mwis_gdpg_call.py:278-318).
        for (int dit_iter = 0;; ++dit_iter) {
        if (DIT) {
            // a graph stops once its residual weights no longer sum to something positive (:296-297; weights are
            // non-negative, so: once no residual vertex has a positive weight)
            for (int g = tid; g < ng; g += kThreads) {
                gpos[g] = 0;
                gcnt[g] = 0;
            }
            __syncthreads();
            for (int i = tid; i < n; i += kThreads)
                if (((keepw[i >> 5] >> (i & 31)) & 1u) && P.wts[v0 + vid[i]] > 0.0) gpos[gid[i]] = 1;
            __syncthreads();
            int any_left = 0;
            for (int base = 0; base < span; base += kThreads) {
                const int sl = base + tid;
                const bool k = sl < n && ((keepw[sl >> 5] >> (sl & 31)) & 1u) && gpos[gid[sl]] != 0;
                const uint32_t w = __ballot_sync(0xffffffffu, k);
                __syncwarp();
                if (lane == 0 && sl < span) keepw[sl >> 5] = w;
                any_left |= k ? 1 : 0;
            }
            if (!__syncthreads_or(any_left)) {
                if (dit_iter == 0 && has_hidden) {  // consume the weight fill issued at the top of the tile
                    mbar_wait(&mbar[0], wuse & 1u);
                    ++wuse;
                }
                break;
            }
            if (dit_iter > 0 && has_hidden && tid == 0) {  // the first hidden layer's weights again
                mbar_expect_tx(&mbar[0], wblob_bytes);
                bulk_g2s(wbuf, P.wall, wblob_bytes, &mbar[0]);
            }
        }
        // ---- 2. degrees on the kept sub-graph -> dinv, y = dinv * x0 ----------------------------------
        for (int i = tid; i < n; i += kThreads) {
            const bool k = (keepw[i >> 5] >> (i & 31)) & 1u;
            const int vloc = vid[i];
            int deg = 0;
            if (k) {
                const int beg = rp[i], end = rp[i + 1];
                for (int e = beg; e < end; ++e) {
                    const int j = col16[e];
                    deg += (keepw[j >> 5] >> (j & 31)) & 1u;  // the dummy's bit is 0
                }
            }
            const float di = deg > 0 ? (float)(1.0 / sqrt((double)deg)) : 0.f;  // gcn/utils.py:122-125
            const float xi = k ? (P.x0 ? P.x0[v0 + vloc] : P.x0val) : 0.f;
            dinv[i] = di;
            x0s[i] = xi;
            sa[i] = di * xi;
            if (k) atomicAdd(&gcnt[gid[i]], 1);
        }
        __syncthreads();

        // ---- 3. first layer, rank-1: s = L.x0 ------------------------------------------------------------
        for (int i = tid; i < n; i += kThreads) {
            float acc = 0.f;
            const int beg = rp[i], end = rp[i + 1];
            for (int e = beg; e < end; ++e) acc += sa[col16[e]];
            sb[i] = x0s[i] - dinv[i] * acc;
        }
        __syncthreads();

        float *cur = fa, *nxt = fb;
        DG_TICK(0)
        if (P.n_layers >= 2) {
            // H1 = act(x0 * colsum(W_0) + s * colsum(W_1) + b), written straight into shared memory
            for (int idx = tid; idx < n * CHUNKS; idx += kThreads) {
                const int i = idx / CHUNKS, c = idx - i * CHUNKS;
                const float xi = x0s[i], si = sb[i];
                const float4 a0 = __ldg(reinterpret_cast<const float4 *>(P.first) + c);
                const float4 a1 = __ldg(reinterpret_cast<const float4 *>(P.first + CP) + c);
                const float4 b0 = __ldg(reinterpret_cast<const float4 *>(P.first + 2 * CP) + c);
                float4 h;
                h.x = act_apply(fmaf(si, a1.x, fmaf(xi, a0.x, b0.x)), P.first_act, P.alpha);
                h.y = act_apply(fmaf(si, a1.y, fmaf(xi, a0.y, b0.y)), P.first_act, P.alpha);
                h.z = act_apply(fmaf(si, a1.z, fmaf(xi, a0.z, b0.z)), P.first_act, P.alpha);
                h.w = act_apply(fmaf(si, a1.w, fmaf(xi, a0.w, b0.w)), P.first_act, P.alpha);
                *reinterpret_cast<float4 *>(cur + swz<CP>(i, c)) = h;
            }
            __syncthreads();

            // ---- 4. hidden layers ---------------------------------------------------------------------
            // A warp owns the 8-row groups {warp, warp + 16, warp + 32, warp + 48} (interleaved so that
            // the hubs of a graph spread over all warps).  Per layer it first aggregates its rows
            // (cur -> nxt: nxt_i = (L.H)_i), then projects the same rows in place
            // (nxt_i = act([cur_i | nxt_i] . [W_0 ; W_1] + b)).  Other warps only ever read `cur` and
            // their own rows of `nxt`, so the two steps need no CTA barrier between them: warps drift
            // apart and the shared-memory-bound aggregation of one overlaps the FFMA-bound projection of
            // another.  One barrier per layer publishes the new features.
            const int G8 = (n + 7) >> 3;
            for (int h = 0; h < n_hidden; ++h) {
                // this layer's weights stream in (TMA bulk copy) while the warps aggregate; the buffer is
                // free because the barrier that ended the previous layer follows its last projection read
                if (h > 0 && tid == 0) {
                    mbar_expect_tx(&mbar[0], wblob_bytes);
                    bulk_g2s(wbuf, reinterpret_cast<const unsigned char *>(P.wall) + (size_t)h * wblob_bytes, wblob_bytes,
                             &mbar[0]);
                }
                const float *w = wbuf;
                const float *bias = w + (MMA ? 2 * 2 * CP * (CP + 8) : 2 * CP * CP);
                const int act = P.acts[h + 1];
                bool w_ready = false;
                for (int pbase = 0; pbase < G8; pbase += 4 * kWarps) {
                    // Rows owned in this pass: from each block e of 128 consecutive slots the warp owns
                    // the 8 rows  own_row(e, k) = 128 (pbase/16 + e) + 16 k + ((warp + k) & 15),  k = 0..7.
                    // Slots are sorted by degree, so the heaviest rows go to 16 different warps (row-level
                    // round robin) and every warp's load is the same up to one row per stripe; the 8 rows
                    // have distinct residues mod 8, which keeps the projection's row loads conflict-free.
                    const int blk0 = pbase / kWarps;
                    int emask = 0;
#pragma unroll
                    for (int e = 0; e < 4; ++e)
                        if (128 * (blk0 + e) < n) emask |= 1 << e;
                    if (emask == 0) continue;
#define DG_OWN_ROW(e, k) (128 * (blk0 + (e)) + 16 * (k) + ((warp + (k)) & 15))
                    // -- aggregation: one lane group (CP/4 lanes) per row, NG rows in flight per warp,
                    //    4 neighbours per step, no predicates (padded lists), no cross-lane reduction
                    {
                        const int g = lane / LPR, q = lane % LPR;
                        const float *curq = cur + (q << 2);
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            if (!((emask >> e) & 1)) continue;
                            // the 8 rows of the group are handled as 8 / NG interleaved streams per lane
                            // group (rows g, g + NG, ...): independent dependency chains in flight
                            constexpr int NS = 8 / NG;
                            int pb[NS], cnt[NS];
                            float4 acc[NS], acc2[NS];
                            int trips = 0;
#pragma unroll
                            for (int sidx = 0; sidx < NS; ++sidx) {
                                const int row = DG_OWN_ROW(e, sidx * NG + g);
                                const bool valid = row < n;
                                pb[sidx] = valid ? rp[row] : 0;
                                cnt[sidx] = valid ? ((rp[row + 1] - pb[sidx]) >> 2) : 0;
                                trips = max(trips, cnt[sidx]);
                                acc[sidx] = make_float4(0.f, 0.f, 0.f, 0.f);
                                acc2[sidx] = make_float4(0.f, 0.f, 0.f, 0.f);
                            }
                            // Heavy rows (hubs) first, one at a time, by the WHOLE warp: the neighbour list is
                            // dealt to the NG x NS (lane group, stream) segments, partial sums are combined
                            // with shuffles.  This cuts the dependent-load chain of a degree-150 row from 38
                            // steps to 5; without it the hub rows are the critical path of every layer.
                            {
                                constexpr int kCoopTrips = 8;
                                unsigned hmask = 0;
#pragma unroll
                                for (int sidx = 0; sidx < NS; ++sidx) {
                                    const unsigned bal = __ballot_sync(0xffffffffu, cnt[sidx] > kCoopTrips);
#pragma unroll
                                    for (int gg = 0; gg < NG; ++gg)
                                        if ((bal >> (gg * LPR)) & 1u) hmask |= 1u << (sidx * NG + gg);
                                }
                                while (hmask) {
                                    const int k = __ffs(hmask) - 1;
                                    hmask &= hmask - 1;
                                    const int own_s = k / NG, own_g = k % NG;
                                    int sel_pb = 0, sel_cnt = 0;
#pragma unroll
                                    for (int sidx = 0; sidx < NS; ++sidx)
                                        if (sidx == own_s) sel_pb = pb[sidx], sel_cnt = cnt[sidx];
                                    const int pbk = __shfl_sync(0xffffffffu, sel_pb, own_g * LPR);
                                    const int cntk = __shfl_sync(0xffffffffu, sel_cnt, own_g * LPR);
                                    float4 part = make_float4(0.f, 0.f, 0.f, 0.f);
                                    for (int base_t = 0; base_t < cntk; base_t += NG * NS) {
#pragma unroll
                                        for (int ss = 0; ss < NS; ++ss) {
                                            const int tt = base_t + ss * NG + g;
                                            if (tt < cntk) {
                                                const uint2 cw = *reinterpret_cast<const uint2 *>(col16 + pbk + 4 * tt);
                                                const int j0 = cw.x & 0xffffu, j1 = cw.x >> 16;
                                                const int j2 = cw.y & 0xffffu, j3 = cw.y >> 16;
                                                const float d0 = dinv[j0], d1 = dinv[j1], d2 = dinv[j2], d3 = dinv[j3];
                                                const float4 x0v = *reinterpret_cast<const float4 *>(curq + j0 * (CP + 4));
                                                const float4 x1v = *reinterpret_cast<const float4 *>(curq + j1 * (CP + 4));
                                                const float4 x2v = *reinterpret_cast<const float4 *>(curq + j2 * (CP + 4));
                                                const float4 x3v = *reinterpret_cast<const float4 *>(curq + j3 * (CP + 4));
                                                part.x += fmaf(d0, x0v.x, d1 * x1v.x) + fmaf(d2, x2v.x, d3 * x3v.x);
                                                part.y += fmaf(d0, x0v.y, d1 * x1v.y) + fmaf(d2, x2v.y, d3 * x3v.y);
                                                part.z += fmaf(d0, x0v.z, d1 * x1v.z) + fmaf(d2, x2v.z, d3 * x3v.z);
                                                part.w += fmaf(d0, x0v.w, d1 * x1v.w) + fmaf(d2, x2v.w, d3 * x3v.w);
                                            }
                                        }
                                    }
#pragma unroll
                                    for (int off = LPR; off < 32; off <<= 1) {
                                        part.x += __shfl_xor_sync(0xffffffffu, part.x, off);
                                        part.y += __shfl_xor_sync(0xffffffffu, part.y, off);
                                        part.z += __shfl_xor_sync(0xffffffffu, part.z, off);
                                        part.w += __shfl_xor_sync(0xffffffffu, part.w, off);
                                    }
#pragma unroll
                                    for (int sidx = 0; sidx < NS; ++sidx) {
                                        if (sidx == own_s && g == own_g) {
                                            acc[sidx] = part;
                                            cnt[sidx] = 0;  // done: the streamed loop below skips it
                                        }
                                    }
                                }
                            }
                            trips = 0;
#pragma unroll
                            for (int sidx = 0; sidx < NS; ++sidx) trips = max(trips, cnt[sidx]);
                            trips = __reduce_max_sync(0xffffffffu, trips);
                            for (int tt = 0; tt < trips; ++tt) {
#pragma unroll
                                for (int sidx = 0; sidx < NS; ++sidx) {
                                    if (tt < cnt[sidx]) {
                                        const uint2 cw = *reinterpret_cast<const uint2 *>(col16 + pb[sidx] + 4 * tt);
                                        const int j0 = cw.x & 0xffffu, j1 = cw.x >> 16;
                                        const int j2 = cw.y & 0xffffu, j3 = cw.y >> 16;
                                        const float d0 = dinv[j0], d1 = dinv[j1], d2 = dinv[j2], d3 = dinv[j3];
                                        const float4 x0v = *reinterpret_cast<const float4 *>(curq + j0 * (CP + 4));
                                        const float4 x1v = *reinterpret_cast<const float4 *>(curq + j1 * (CP + 4));
                                        const float4 x2v = *reinterpret_cast<const float4 *>(curq + j2 * (CP + 4));
                                        const float4 x3v = *reinterpret_cast<const float4 *>(curq + j3 * (CP + 4));
                                        acc[sidx].x = fmaf(d0, x0v.x, acc[sidx].x);
                                        acc[sidx].y = fmaf(d0, x0v.y, acc[sidx].y);
                                        acc[sidx].z = fmaf(d0, x0v.z, acc[sidx].z);
                                        acc[sidx].w = fmaf(d0, x0v.w, acc[sidx].w);
                                        acc2[sidx].x = fmaf(d1, x1v.x, acc2[sidx].x);
                                        acc2[sidx].y = fmaf(d1, x1v.y, acc2[sidx].y);
                                        acc2[sidx].z = fmaf(d1, x1v.z, acc2[sidx].z);
                                        acc2[sidx].w = fmaf(d1, x1v.w, acc2[sidx].w);
                                        acc[sidx].x = fmaf(d2, x2v.x, acc[sidx].x);
                                        acc[sidx].y = fmaf(d2, x2v.y, acc[sidx].y);
                                        acc[sidx].z = fmaf(d2, x2v.z, acc[sidx].z);
                                        acc[sidx].w = fmaf(d2, x2v.w, acc[sidx].w);
                                        acc2[sidx].x = fmaf(d3, x3v.x, acc2[sidx].x);
                                        acc2[sidx].y = fmaf(d3, x3v.y, acc2[sidx].y);
                                        acc2[sidx].z = fmaf(d3, x3v.z, acc2[sidx].z);
                                        acc2[sidx].w = fmaf(d3, x3v.w, acc2[sidx].w);
                                    }
                                }
                            }
#pragma unroll
                            for (int sidx = 0; sidx < NS; ++sidx) {
                                const int row = DG_OWN_ROW(e, sidx * NG + g);
                                if (row < n) {
                                    const float di = dinv[row];
                                    const float4 hi = *reinterpret_cast<const float4 *>(curq + row * (CP + 4));
                                    float4 lh;
                                    lh.x = fmaf(-di, acc[sidx].x + acc2[sidx].x, hi.x);
                                    lh.y = fmaf(-di, acc[sidx].y + acc2[sidx].y, hi.y);
                                    lh.z = fmaf(-di, acc[sidx].z + acc2[sidx].z, hi.z);
                                    lh.w = fmaf(-di, acc[sidx].w + acc2[sidx].w, hi.w);
                                    *reinterpret_cast<float4 *>(nxt + row * (CP + 4) + (q << 2)) = lh;
                                }
                            }
                        }
                    }
                    __syncwarp();
                    DG_TICK(1)
                    if (!w_ready) {  // this layer's weights must have landed
                        mbar_wait(&mbar[0], wuse & 1u);
                        w_ready = true;
                    }
                    DG_TICK(3)
                    // -- projection of the same rows, in place.  Register-tiled FFMA GEMM: a lane owns
                    //    4 rows (one per owned group: row (gb0 + 16 e) * 8 + (lane & 7)) x 8 columns
                    //    ((lane >> 3) * 8 ..); per 4 values of k it needs 4 row loads and 8 weight loads.
                    if (CP == 32 && !MMA) {
                        const int rg = lane & 7, cg = lane >> 3;
                        float acc[4][8];
                        {
                            const float4 b0 = *reinterpret_cast<const float4 *>(bias + cg * 8);
                            const float4 b1 = *reinterpret_cast<const float4 *>(bias + cg * 8 + 4);
#pragma unroll
                            for (int e = 0; e < 4; ++e) {
                                acc[e][0] = b0.x, acc[e][1] = b0.y, acc[e][2] = b0.z, acc[e][3] = b0.w;
                                acc[e][4] = b1.x, acc[e][5] = b1.y, acc[e][6] = b1.z, acc[e][7] = b1.w;
                            }
                        }
                        int rows[4];
#pragma unroll
                        for (int e = 0; e < 4; ++e) rows[e] = min(DG_OWN_ROW(e, rg), n - 1);
#pragma unroll 1
                        for (int src = 0; src < 2; ++src) {
                            const float *feat = src == 0 ? cur : nxt;
                            const float *wk = w + (size_t)src * CP * CP + cg * 8;
#pragma unroll 2
                            for (int kc = 0; kc < CHUNKS; ++kc) {
                                float4 u[4];
#pragma unroll
                                for (int e = 0; e < 4; ++e)
                                    if ((emask >> e) & 1) u[e] = *reinterpret_cast<const float4 *>(feat + swz<CP>(rows[e], kc));
#pragma unroll
                                for (int kk = 0; kk < 4; ++kk) {
                                    const float4 wa = *reinterpret_cast<const float4 *>(wk + (kc * 4 + kk) * CP);
                                    const float4 wb = *reinterpret_cast<const float4 *>(wk + (kc * 4 + kk) * CP + 4);
#pragma unroll
                                    for (int e = 0; e < 4; ++e) {
                                        if ((emask >> e) & 1) {
                                            const float uv =
                                                kk == 0 ? u[e].x : kk == 1 ? u[e].y : kk == 2 ? u[e].z : u[e].w;
                                            acc[e][0] = fmaf(uv, wa.x, acc[e][0]);
                                            acc[e][1] = fmaf(uv, wa.y, acc[e][1]);
                                            acc[e][2] = fmaf(uv, wa.z, acc[e][2]);
                                            acc[e][3] = fmaf(uv, wa.w, acc[e][3]);
                                            acc[e][4] = fmaf(uv, wb.x, acc[e][4]);
                                            acc[e][5] = fmaf(uv, wb.y, acc[e][5]);
                                            acc[e][6] = fmaf(uv, wb.z, acc[e][6]);
                                            acc[e][7] = fmaf(uv, wb.w, acc[e][7]);
                                        }
                                    }
                                }
                            }
                        }
                        __syncwarp();  // every lane has finished reading the warp's rows
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const int row = DG_OWN_ROW(e, rg);
                            if (((emask >> e) & 1) && row < n) {
                                float4 o0, o1;
                                o0.x = act_apply(acc[e][0], act, P.alpha);
                                o0.y = act_apply(acc[e][1], act, P.alpha);
                                o0.z = act_apply(acc[e][2], act, P.alpha);
                                o0.w = act_apply(acc[e][3], act, P.alpha);
                                o1.x = act_apply(acc[e][4], act, P.alpha);
                                o1.y = act_apply(acc[e][5], act, P.alpha);
                                o1.z = act_apply(acc[e][6], act, P.alpha);
                                o1.w = act_apply(acc[e][7], act, P.alpha);
                                *reinterpret_cast<float4 *>(nxt + swz<CP>(row, cg * 2)) = o0;
                                *reinterpret_cast<float4 *>(nxt + swz<CP>(row, cg * 2 + 1)) = o1;
                            }
                        }
                    }
                    // -- projection on the tensor cores (mma.sync m16n8k8 TF32) with fp32-level accuracy:
                    //    every operand is split a = hi + lo with hi, lo exactly representable in TF32
                    //    (cvt.rna), and a.b ~ hi.hi + (lo.hi + hi.lo).  The dominant hi.hi products are exact
                    //    in fp32 and are accumulated OUTSIDE the tensor core with ordinary FADDs (the tensor
                    //    core's own accumulation truncates; chaining 8 k-steps through it would bias the
                    //    sum), the two small correction terms are chained inside.  Weights are pre-split on
                    //    the host and stored with a row stride of 40 words so that fragment loads are
                    //    conflict-free.  An m16 tile = two of the warp's 8-row groups.
                    if (CP == 32 && MMA) {
                        const int g = lane >> 2, t4 = lane & 3;
                        constexpr int WS = CP + 8;            // padded row stride of the weight matrices
                        const float *whi = w, *wlo = w + 2 * CP * WS;
                        int ra[2], rb[2];
#pragma unroll
                        for (int mt = 0; mt < 2; ++mt) {
                            ra[mt] = min(DG_OWN_ROW(2 * mt, g), n - 1) * (CP + 4);
                            rb[mt] = min(DG_OWN_ROW(2 * mt + 1, g), n - 1) * (CP + 4);
                        }
                        const int mmask = ((emask & 3) ? 1 : 0) | ((emask & 12) ? 2 : 0);
                        float accm[2][4][4], accc[2][4][4];
#pragma unroll
                        for (int mt = 0; mt < 2; ++mt)
#pragma unroll
                            for (int nt = 0; nt < 4; ++nt)
#pragma unroll
                                for (int c = 0; c < 4; ++c) accm[mt][nt][c] = 0.f, accc[mt][nt][c] = 0.f;
#pragma unroll 1
                        for (int src = 0; src < 2; ++src) {
                            const float *feat = src == 0 ? cur : nxt;
#pragma unroll 1
                            for (int k8 = 0; k8 < CP / 8; ++k8) {
                                const int kt = src * (CP / 8) + k8;  // k-tile of the concatenated [H | L.H]
                                uint32_t ahi[2][4], alo[2][4];
#pragma unroll
                                for (int mt = 0; mt < 2; ++mt) {
                                    if ((mmask >> mt) & 1) {
                                        float av[4];
                                        av[0] = feat[ra[mt] + k8 * 8 + t4];
                                        av[1] = feat[rb[mt] + k8 * 8 + t4];
                                        av[2] = feat[ra[mt] + k8 * 8 + t4 + 4];
                                        av[3] = feat[rb[mt] + k8 * 8 + t4 + 4];
#pragma unroll
                                        for (int c = 0; c < 4; ++c) {
                                            uint32_t hi, lo;
                                            asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hi) : "f"(av[c]));
                                            const float rem = av[c] - __uint_as_float(hi);
                                            asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(lo) : "f"(rem));
                                            ahi[mt][c] = hi;
                                            alo[mt][c] = lo;
                                        }
                                    }
                                }
#pragma unroll
                                for (int nt = 0; nt < 4; ++nt) {
                                    const int bo = (kt * 8 + t4) * WS + nt * 8 + g;
                                    const uint32_t bh0 = __float_as_uint(whi[bo]), bh1 = __float_as_uint(whi[bo + 4 * WS]);
                                    const uint32_t bl0 = __float_as_uint(wlo[bo]), bl1 = __float_as_uint(wlo[bo + 4 * WS]);
#pragma unroll
                                    for (int mt = 0; mt < 2; ++mt) {
                                        if ((mmask >> mt) & 1) {
                                            float d0, d1, d2, d3;
                                            asm volatile(
                                                "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 "
                                                "{%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%10,%10,%10,%10};"
                                                : "=f"(d0), "=f"(d1), "=f"(d2), "=f"(d3)
                                                : "r"(ahi[mt][0]), "r"(ahi[mt][1]), "r"(ahi[mt][2]), "r"(ahi[mt][3]),
                                                  "r"(bh0), "r"(bh1), "f"(0.f));
                                            accm[mt][nt][0] += d0;
                                            accm[mt][nt][1] += d1;
                                            accm[mt][nt][2] += d2;
                                            accm[mt][nt][3] += d3;
                                            asm volatile(
                                                "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 "
                                                "{%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                                                : "+f"(accc[mt][nt][0]), "+f"(accc[mt][nt][1]), "+f"(accc[mt][nt][2]),
                                                  "+f"(accc[mt][nt][3])
                                                : "r"(alo[mt][0]), "r"(alo[mt][1]), "r"(alo[mt][2]), "r"(alo[mt][3]),
                                                  "r"(bh0), "r"(bh1));
                                            asm volatile(
                                                "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 "
                                                "{%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                                                : "+f"(accc[mt][nt][0]), "+f"(accc[mt][nt][1]), "+f"(accc[mt][nt][2]),
                                                  "+f"(accc[mt][nt][3])
                                                : "r"(ahi[mt][0]), "r"(ahi[mt][1]), "r"(ahi[mt][2]), "r"(ahi[mt][3]),
                                                  "r"(bl0), "r"(bl1));
                                        }
                                    }
                                }
                            }
                        }
                        __syncwarp();  // every lane has finished reading the warp's rows
#pragma unroll
                        for (int mt = 0; mt < 2; ++mt) {
                            const int rowa = DG_OWN_ROW(2 * mt, g), rowb = DG_OWN_ROW(2 * mt + 1, g);
                            const bool oka = ((emask >> (2 * mt)) & 1) && rowa < n;
                            const bool okb = ((emask >> (2 * mt + 1)) & 1) && rowb < n;
#pragma unroll
                            for (int nt = 0; nt < 4; ++nt) {
                                const int col = nt * 8 + 2 * t4;
                                const float b0 = bias[col], b1 = bias[col + 1];
                                if (oka) {
                                    float2 o;
                                    o.x = act_apply(accm[mt][nt][0] + accc[mt][nt][0] + b0, act, P.alpha);
                                    o.y = act_apply(accm[mt][nt][1] + accc[mt][nt][1] + b1, act, P.alpha);
                                    *reinterpret_cast<float2 *>(nxt + rowa * (CP + 4) + col) = o;
                                }
                                if (okb) {
                                    float2 o;
                                    o.x = act_apply(accm[mt][nt][2] + accc[mt][nt][2] + b0, act, P.alpha);
                                    o.y = act_apply(accm[mt][nt][3] + accc[mt][nt][3] + b1, act, P.alpha);
                                    *reinterpret_cast<float2 *>(nxt + rowb * (CP + 4) + col) = o;
                                }
                            }
                        }
                    }
                    DG_TICK(4)
#undef DG_OWN_ROW
                }
                // every thread consumes this layer's weight-barrier phase exactly once
                if (!w_ready) mbar_wait(&mbar[0], wuse & 1u);
                ++wuse;
                __syncthreads();
                DG_TICK(5)
                float *tmp = cur;
                cur = nxt;
                nxt = tmp;
            }

            // ---- 5. last layer, projected first: q = H.w_0 + z, zs = dinv * z, z = H.w_1 -----------------
            for (int i = tid; i < n; i += kThreads) {
                float t0 = 0.f, t1 = 0.f;
#pragma unroll 4
                for (int c = 0; c < CHUNKS; ++c) {
                    const float4 hv = *reinterpret_cast<const float4 *>(cur + swz<CP>(i, c));
                    const float4 w0 = __ldg(reinterpret_cast<const float4 *>(P.tail) + c);
                    const float4 w1 = __ldg(reinterpret_cast<const float4 *>(P.tail + CP) + c);
                    t0 = fmaf(hv.x, w0.x, t0);
                    t0 = fmaf(hv.y, w0.y, t0);
                    t0 = fmaf(hv.z, w0.z, t0);
                    t0 = fmaf(hv.w, w0.w, t0);
                    t1 = fmaf(hv.x, w1.x, t1);
                    t1 = fmaf(hv.y, w1.y, t1);
                    t1 = fmaf(hv.z, w1.z, t1);
                    t1 = fmaf(hv.w, w1.w, t1);
                }
                sb[i] = t0 + t1;
                sa[i] = dinv[i] * t1;
            }
            __syncthreads();
        }

        // ---- 6. score and utility (mwis_dqn_call.py:230-235) -------------------------------------------
        for (int i = tid; i < n; i += kThreads) {
            const bool k = (keepw[i >> 5] >> (i & 31)) & 1u;
            float v;
            if (P.n_layers == 1) {
                v = act_apply(fmaf(sb[i], __ldg(P.first + CP), fmaf(x0s[i], __ldg(P.first), __ldg(P.first + 2 * CP))),
                              P.first_act, P.alpha);
            } else {
                float acc = 0.f;
                const int beg = rp[i], end = rp[i + 1];
                for (int e = beg; e < end; ++e) acc += sa[col16[e]];
                v = act_apply(sb[i] - dinv[i] * acc + P.tail_bias, P.last_act, P.alpha);
            }
            if (!k) v = 0.f;
            const int vtx = v0 + vid[i];
            const double u = (P.predict == DG_PREDICT_MWIS) ? (double)v * P.wts[vtx] : (double)v;
            util_sm[i] = u;
            if (P.score) P.score[vtx] = v;
            if (P.util) P.util[vtx] = u;
        }
        for (int wd = tid; wd < span / 32; wd += kThreads) remain[wd] = keepw[wd];
        __syncthreads();
        DG_TICK(6)

        // ---- 7. local greedy search rounds (heuristics.py:77-116), all graphs of the tile together ----
        int n_remain = 1;
        int rounds = 0;
        while (P.do_lgs) {
            if (DIT && rounds >= 1) break;  // one greedy round per re-scoring
            // per-graph round accounting (heuristics.py:119-160): a graph's step count grows while it
            // still has remaining vertices
            int any = 0;
            if (tid < ng) {
                const int c = gcnt[tid];
                if (c > 0) {
                    gsteps[tid] += 1;
                    any = 1;
                }
                gcnt[tid] = 0;
            }
            n_remain = __syncthreads_or(any);
            if (!n_remain) break;
            if (rounds >= P.round_cap) {
                if (tid == 0) atomicExch(P.status, DG_ERR_NOT_CONVERGED);
                break;
            }
            for (int base = 0; base < span; base += kThreads) {
                const int v = base + tid;
                const bool active = v < span ? (remain[v >> 5] >> lane) & 1u : false;
                bool join = false;
                if (active) {
                    const double wv = util_sm[v];
                    const int myid = vid[v];  // the index tie-break is on ORIGINAL vertex ids
                    const int beg = rp[v], end = rp[v + 1];
                    join = true;
                    for (int e = beg; e < end; ++e) {
                        const int u = col16[e];
                        if ((remain[u >> 5] >> (u & 31)) & 1u) {
                            const double wu = util_sm[u];
                            if (!((wv > wu) || (wv == wu && myid < (int)vid[u]))) {
                                join = false;
                                break;
                            }
                        }
                    }
                    if (join && P.member) P.member[v0 + myid] = 1;
                }
                const uint32_t jw = __ballot_sync(0xffffffffu, join);
                if (lane == 0 && v < span) {
                    joined[v >> 5] = jw;
                    memb[v >> 5] |= jw;
                }
            }
            __syncthreads();
            for (int base = 0; base < span; base += kThreads) {
                const int v = base + tid;
                bool still = false;
                if (v < span) {
                    const bool active = (remain[v >> 5] >> lane) & 1u;
                    const bool join = (joined[v >> 5] >> lane) & 1u;
                    if (active && !join) {
                        const int beg = rp[v], end = rp[v + 1];
                        still = true;
                        for (int e = beg; e < end; ++e) {
                            const int u = col16[e];
                            if ((joined[u >> 5] >> (u & 31)) & 1u) {
                                still = false;
                                break;
                            }
                        }
                    }
                }
                const uint32_t rw = __ballot_sync(0xffffffffu, still);
                __syncwarp();  // every lane has read this word (WAR within the warp)
                if (lane == 0 && v < span) remain[v >> 5] = rw;
                if (still) atomicAdd(&gcnt[gid[v]], 1);
            }
            ++rounds;
            __syncthreads();
        }
        if constexpr (!DIT) {
            break;
        } else {
            // the residual graph of the next iteration: what is still in `remain`
            for (int wd = tid; wd < span / 32; wd += kThreads) keepw[wd] = remain[wd];
            __syncthreads();
        }
        }  // dit_iter

        // ---- 8. per-graph outputs ---------------------------------------------------------------------
        if (P.steps)
            for (int g = tid; g < ng; g += kThreads) P.steps[g0 + g] = gsteps[g];
        if (P.total) {
            for (int g = warp; g < ng; g += kWarps) {
                double acc = 0.0;
                for (int i = gstart[g] + lane; i < gstart[g + 1]; i += 32) {  // vertex order: deterministic
                    const int sl = slotof[i];
                    if ((memb[sl >> 5] >> (sl & 31)) & 1u) acc += P.wts[v0 + i];  // mwis_dqn_call.py:241
                }
#pragma unroll
                for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
                if (lane == 0) P.total[g0 + g] = acc;
            }
        }
        __syncthreads();  // shared memory is recycled by the next tile
        DG_TICK(7)
        if (timing) {
            tm[8] += 1;
            P.dbg[(size_t)gridDim.x * 16 + (size_t)t * 4 + 0] = clock64() - t_tile;
            P.dbg[(size_t)gridDim.x * 16 + (size_t)t * 4 + 1] = n;
            P.dbg[(size_t)gridDim.x * 16 + (size_t)t * 4 + 2] = rp[n];
            P.dbg[(size_t)gridDim.x * 16 + (size_t)t * 4 + 3] = ng;
        }
    }
    if (timing) {
        tm[9] = clock64() - t_begin;
        for (int k = 0; k < 10; ++k) P.dbg[(size_t)blockIdx.x * 16 + k] = tm[k];
    }
#undef DG_TICK
    // the last CTA to leave re-arms the tile counter for the next launch on this context
    if (tid == 0) {
        __threadfence();
        const int done = atomicAdd(P.tile_counter + 1, 1);
        if (done == (int)gridDim.x - 1) {
            P.tile_counter[1] = 0;
            __threadfence();
            P.tile_counter[0] = 0;
        }
    }
}

}  // namespace

// ---- host side ---------------------------------------------------------------------------------------
namespace {

size_t fused_smem_bytes(int cp, int cap_n, int cap_nnz, bool has_hidden, size_t wblob_bytes) {
    return fused_smem_plan(cp, cap_n, cap_nnz, has_hidden, wblob_bytes).total;
}

template <int CP, bool DIT, bool MMA>
int launch_t(dg_context *ctx, const FusedParams &p, size_t smem, int n_tiles) {
    auto kern = fused_solve_kernel<CP, DIT, MMA>;
    // attribute and occupancy are queried once per (thread, device, instantiation, shared-memory size): they cost
    // microseconds of host time that the streaming path would pay on every call.  The attribute is always the device's
    // opt-in maximum: a per-call value would let one host thread lower it under another thread's launch (contexts of
    // several producer threads share the function).
    struct Cache {
        int dev = -1, dyn_max = 0, n = 0;
        size_t smem[16];
        int per_sm[16];
    };
    static thread_local Cache c;
    if (c.dev != ctx->device) {
        cudaFuncAttributes fa;
        DG_CUDA_CHECK(cudaFuncGetAttributes(&fa, kern));
        const int dyn_max = ctx->max_smem_optin - (int)fa.sharedSizeBytes;   // (static shared memory counts against the limit)
        DG_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, dyn_max));
        c.dev = ctx->device, c.dyn_max = dyn_max, c.n = 0;
    }
    DG_REQUIRE((long long)smem <= (long long)c.dyn_max, DG_ERR_UNSUPPORTED,
               "fused kernel needs %zu bytes of dynamic shared memory (limit %d)", smem, c.dyn_max);
    int per_sm = 0;
    for (int k = 0; k < c.n; ++k)
        if (c.smem[k] == smem) per_sm = c.per_sm[k];
    if (per_sm == 0) {   // (a stream of batches alternates between a few tile capacities: one entry per size)
        DG_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kThreads, smem));
        const int slot = c.n < 16 ? c.n++ : 15;
        c.smem[slot] = smem, c.per_sm[slot] = per_sm;
    }
    DG_REQUIRE(per_sm > 0, DG_ERR_UNSUPPORTED, "fused kernel does not fit on an SM with %zu bytes of shared memory",
               smem);
    int grid = ctx->sm_count * per_sm;
    if (grid > n_tiles) grid = n_tiles;
    kern<<<grid, kThreads, smem, ctx->stream>>>(p);
    ctx->launches++;
    DG_CUDA_CHECK(cudaGetLastError());
    return DG_OK;
}

struct Tile {
    int v0, n, e0, nnz, g0, ng;
    long long cost;   // tile_cost(), filled in when the tile is closed
};

// Cost model of a tile for scheduling decisions, in SM cycles for an 18-hidden-layer model, fitted to
// per-tile clock64() measurements on B200 (profiles/r01_notes.md, "tile cost fit", rms error 3 %):
// ~41 cycles per padded neighbour entry, ~300 per row, and a large constant - every warp streams the
// whole weight matrix through shared memory once per layer whatever the tile size, plus staging and the
// greedy rounds - which is why few large tiles beat many small ones.
inline long long tile_cost(const Tile &t) {
    return 41LL * ((long long)t.nnz + 3LL * t.n / 2) + 300LL * t.n + 200000LL;
}

// `subset`: only the graphs the tensor-core kernel left out (dg_batch::tc_skip); a tile never spans a gap
void pack_tiles(const dg_batch *b, int cap_n, int cap_nnz, bool subset, std::vector<Tile> *out) {
    out->clear();
    const auto &gp = b->h_graph_ptr;
    const auto &ge = b->h_graph_e;
    Tile cur{0, 0, 0, 0, 0, 0, 0};
    auto close = [&]() {
        cur.cost = tile_cost(cur);
        out->push_back(cur);
        cur = Tile{0, 0, 0, 0, 0, 0, 0};
    };
    for (int g = 0; g < b->n_graphs; ++g) {
        if (subset && !b->tc_skip[(size_t)g]) {
            if (cur.ng > 0) close();
            continue;
        }
        const int gn = gp[g + 1] - gp[g], gz = ge[g + 1] - ge[g];
        // cap_nnz bounds the PADDED neighbour lists (each row rounded up to a multiple of 4)
        if (cur.ng > 0 && (cur.n + gn > cap_n || cur.nnz + gz + 3 * (cur.n + gn) > cap_nnz ||
                           cur.ng >= kFusedMaxTileGraphs)) {
            close();
        }
        if (cur.ng == 0) {
            cur.v0 = gp[g];
            cur.e0 = ge[g];
            cur.g0 = g;
        }
        cur.n += gn;
        cur.nnz += gz;
        cur.ng += 1;
    }
    if (cur.ng > 0) close();
    std::stable_sort(out->begin(), out->end(), [](const Tile &a, const Tile &c) { return a.cost > c.cost; });
}

// makespan of heaviest-first list scheduling of the tiles on `workers` CTAs (what the kernel's
// atomic tile counter does)
long long simulate_makespan(const std::vector<Tile> &tiles, int workers) {
    std::vector<long long> load((size_t)workers, 0);
    std::make_heap(load.begin(), load.end(), std::greater<long long>());
    for (const Tile &t : tiles) {
        std::pop_heap(load.begin(), load.end(), std::greater<long long>());
        load.back() += t.cost;
        std::push_heap(load.begin(), load.end(), std::greater<long long>());
    }
    return *std::max_element(load.begin(), load.end());
}

// Choose the tile capacity and (re)build the tile table: consecutive graphs are packed greedily while
// they fit and tiles are ordered heaviest first for the dynamic scheduler.  The row capacity is picked
// among the feasible ones by simulating that schedule: with a few hundred graphs on 148 SMs the
// number of tiles per CTA is small and a capacity that leaves a few CTAs with one tile more than the
// others costs tens of percent.
int build_tiles(dg_context *ctx, dg_batch *b, int cp, bool has_hidden, size_t wblob, bool subset, bool *ok) {
    *ok = false;
    if (b->tiles_valid) {  // the plan depends only on the batch and on (cp, has_hidden, subset)
        if (b->tiles_cp == cp && b->tiles_hidden == has_hidden && b->tiles_wblob == wblob && b->tiles_subset == subset) {
            *ok = b->n_tiles > 0;
            return DG_OK;
        }
    }
    const size_t budget = (size_t)ctx->max_smem_optin - 1024;  // leave room for static shared memory
    const int min_n = ((std::max(b->max_graph_nodes, 32) + 31) / 32) * 32;
    const long long need_nnz = (long long)b->max_graph_nnz + 3LL * b->max_graph_nodes;  // padded worst case
    const int min_nnz = (int)((std::max<long long>(need_nnz, 64) + 63) / 64 * 64);
    if (min_n > 65535 - 1) return DG_OK;
    if (fused_smem_bytes(cp, min_n, min_nnz, has_hidden, wblob) > budget) return DG_OK;  // largest graph does not fit
    int forced_rows = 0;
    forced_rows = ctx->env.tile_rows;
    // candidate row capacities: multiples of 32 from the largest graph up to what shared memory allows
    std::vector<Tile> best_tiles, tiles;
    int best_n = 0, best_nnz = 0;
    long long best_span = -1;
    // A context that solves a stream of similar batches (dg_solve_host*) re-plans on every call: when the previous plan
    // was made for a batch of the same shape (graph count, largest graph), only its capacity and the two neighbouring
    // ones are simulated again.
    // (Only the hinted capacity itself: a plan costs ~20 us per candidate for 500 graphs in random order, most of it the
    // sort and the simulated schedule.  Every 64th call searches all capacities again, so a hint taken from an unlucky
    // first batch does not stick.)
    bool hinted = !subset && b->tiles_hint_n > 0 && b->tiles_hint_graphs == b->n_graphs &&
                  b->tiles_hint_min_n == min_n && b->tiles_hint_cp == cp && b->tiles_hint_hidden == has_hidden;
    if (hinted && (++b->tiles_hint_calls & 63) == 0) hinted = false;
    auto search = [&](bool use_hint) {
        for (int cap_n = min_n; cap_n <= 1024; cap_n += 32) {
            if (forced_rows > 0 && cap_n != std::max(min_n, (forced_rows + 31) / 32 * 32)) continue;
            if (forced_rows <= 0 && use_hint && cap_n != b->tiles_hint_n) continue;
            // give the neighbour lists whatever shared memory is left (bounded by 48 padded entries per row)
            const size_t fixed = fused_smem_bytes(cp, cap_n, 0, has_hidden, wblob);
            if (fixed + sizeof(uint16_t) * (size_t)min_nnz > budget) break;
            long long cap_nnz = (long long)((budget - fixed) / sizeof(uint16_t));
            cap_nnz = std::min<long long>(cap_nnz, std::max<long long>(min_nnz, 48LL * cap_n));
            cap_nnz = cap_nnz / 64 * 64;
            if (cap_nnz < min_nnz) break;
            pack_tiles(b, cap_n, (int)cap_nnz, subset, &tiles);
            // (a single hinted candidate is taken as it is: nothing to compare its simulated schedule with)
            const long long span = (use_hint && forced_rows <= 0) ? 0 : simulate_makespan(tiles, ctx->sm_count);
            if (best_span < 0 || span < best_span) {
                best_span = span;
                best_n = cap_n;
                best_nnz = (int)cap_nnz;
                best_tiles.swap(tiles);
            }
        }
    };
    search(hinted);
    if (best_span < 0 && hinted) search(false);   // the hinted capacity does not take this batch's largest graph
    if (best_span < 0) return DG_OK;
    if (!subset) {
        b->tiles_hint_n = best_n, b->tiles_hint_graphs = b->n_graphs, b->tiles_hint_min_n = min_n;
        b->tiles_hint_cp = cp, b->tiles_hint_hidden = has_hidden;
    }
    std::vector<int> flat(best_tiles.size() * 8, 0);
    for (size_t t = 0; t < best_tiles.size(); ++t) {
        int *d = &flat[t * 8];
        d[0] = best_tiles[t].v0, d[1] = best_tiles[t].n, d[2] = best_tiles[t].e0, d[3] = best_tiles[t].nnz,
        d[4] = best_tiles[t].g0, d[5] = best_tiles[t].ng;
    }
    if (b->tiles_cap < flat.size() || !b->tiles_dev) {
        if (b->tiles_dev) {
            DG_CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
            cudaFree(b->tiles_dev);
            b->tiles_dev = nullptr;
        }
        b->tiles_cap = flat.size() + flat.size() / 4 + 8;
        DG_CUDA_CHECK(cudaMalloc((void **)&b->tiles_dev, sizeof(int) * b->tiles_cap));
    }
    if (!flat.empty()) {
        // pageable source: cudaMemcpyAsync stages it before returning, so `flat` may die afterwards
        DG_CUDA_CHECK(cudaMemcpyAsync(b->tiles_dev, flat.data(), sizeof(int) * flat.size(), cudaMemcpyHostToDevice,
                                      ctx->stream));
    }
    b->n_tiles = (int)best_tiles.size();
    b->tiles_cap_n = best_n;
    b->tiles_cap_nnz = best_nnz;
    b->tiles_cp = cp;
    b->tiles_hidden = has_hidden;
    b->tiles_wblob = wblob;
    b->tiles_subset = subset;
    b->tiles_valid = true;
    if (ctx->env.fused_timing)
        fprintf(stderr, "[fused tiles] %d tiles, cap_n %d, cap_nnz %d, simulated makespan %lld (hinted %d: min_n %d, max nodes %d, "
                "max nnz %d)\n", b->n_tiles, best_n, best_nnz, best_span, (int)hinted, min_n, b->max_graph_nodes, b->max_graph_nnz);
    *ok = b->n_tiles > 0;
    return DG_OK;
}

}  // namespace

bool fused_fits(dg_context *ctx, const dg_model *m, const dg_batch *b) {
    if (ctx->env.disable_fused) return false;
    if (m->fused_cp == 0 || b->n_graphs == 0 || b->n_nodes == 0) return false;
    if ((int)b->h_graph_e.size() != b->n_graphs + 1) return false;
    const bool has_hidden = m->n_layers >= 3;
    const bool use_mma = m->fused_cp == 32 && has_hidden && m->fused_wall_mma && ctx->env.fused_mma;
    const size_t wblob = use_mma ? sizeof(float) * (size_t)(2 * 2 * 32 * 40 + 32)
                                 : sizeof(float) * (size_t)(2 * m->fused_cp * m->fused_cp + m->fused_cp);
    // the same feasibility test build_tiles starts with: the largest graph of the batch must fit a tile
    const size_t budget = (size_t)ctx->max_smem_optin - 1024;
    const int min_n = ((std::max(b->max_graph_nodes, 32) + 31) / 32) * 32;
    const long long need_nnz = (long long)b->max_graph_nnz * (b->upper_pending ? 2 : 1) + 3LL * b->max_graph_nodes;
    const int min_nnz = (int)((std::max<long long>(need_nnz, 64) + 63) / 64 * 64);
    if (min_n > 1024) return false;
    return fused_smem_bytes(m->fused_cp, min_n, min_nnz, has_hidden, wblob) <= budget;
}

int fused_try_solve(dg_context *ctx, const dg_model *m, dg_batch *b, const double *d_wts, int predict,
                    int remove_zero_weight, uint8_t *member, float *score, double *util, double *total,
                    int32_t *steps, bool *handled, bool dit) {
    *handled = false;
    // the tensor-core kernel has just solved this batch except the graphs beyond its limits: only those are left
    const bool subset = b->tc_ran_partial && !dit;
    b->tc_ran_partial = false;
    // member == nullptr: scores only (dg_gcn_forward); the greedy rounds are skipped
    if (member == nullptr && (d_wts == nullptr) && predict == DG_PREDICT_MWIS) return DG_OK;
    if (ctx->env.disable_fused) return DG_OK;
    if (m->fused_cp == 0 || b->n_graphs == 0 || b->n_nodes == 0) return DG_OK;
    if ((int)b->h_graph_e.size() != b->n_graphs + 1) return DG_OK;
    const bool has_hidden = m->n_layers >= 3;
    // The split-TF32 mma.sync projection is kept as an option (DG_FUSED_MMA=1): it is ~2x more accurate
    // than the FFMA chain but not faster on B200 - legacy mma.sync TF32 issues at about the FFMA rate
    // and the 3-term split triples the work (profiles/r01_notes.md).
    const bool use_mma = m->fused_cp == 32 && has_hidden && m->fused_wall_mma && ctx->env.fused_mma;
    const size_t wblob = use_mma ? sizeof(float) * (size_t)(2 * 2 * 32 * 40 + 32)
                                 : sizeof(float) * (size_t)(2 * m->fused_cp * m->fused_cp + m->fused_cp);
    bool ok = false;
    const auto t_plan0 = std::chrono::steady_clock::now();
    DG_TRY(build_tiles(ctx, b, m->fused_cp, has_hidden, wblob, subset, &ok));
    if (ctx->env.ingest_timing)
        fprintf(stderr, "[fused] tile plan %.0f us\n",
                std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t_plan0).count());
    if (!ok) return DG_OK;
    FusedParams p{};
    p.tiles = b->tiles_dev;
    p.n_tiles = b->n_tiles;
    p.tile_counter = ctx->d_status + 1;
    p.graph_ptr = b->graph_ptr;
    p.row_ptr = b->row_ptr;
    p.col_idx = b->col_idx;
    p.wts = d_wts;
    p.keep_in = remove_zero_weight ? nullptr : b->keep;
    p.x0 = b->x0;
    p.x0val = 1.0f / (float)m->layers[0].c_in;
    p.remove_zero = remove_zero_weight ? 1 : 0;
    p.n_layers = m->n_layers;
    p.first = m->fused_first;
    p.first_act = m->layers[0].act;
    p.wall = use_mma ? m->fused_wall_mma : m->fused_wall;
    p.use_mma = use_mma ? 1 : 0;
    p.wblob_bytes = (int)wblob;
    p.acts = m->d_acts;
    p.tail = m->fused_tail;
    p.tail_bias = m->tail_bias;
    p.last_act = m->layers.back().act;
    p.alpha = m->alpha;
    p.predict = predict;
    p.cap_n = b->tiles_cap_n;
    p.cap_nnz = b->tiles_cap_nnz;
    p.member = member;
    p.score = score;
    p.util = util;
    p.total = total;
    p.steps = steps;
    p.status = ctx->d_status;
    p.round_cap = kLgsRoundCap;
    p.do_lgs = member != nullptr ? 1 : 0;
    p.dit = (dit && member != nullptr) ? 1 : 0;
    p.dbg = nullptr;
    if (ctx->env.fused_timing) {
        long long *dbg = nullptr;
        DG_TRY(scratch_as(ctx, kSlotLgsWords, (size_t)ctx->sm_count * 4 * 16 + (size_t)p.n_tiles * 4, &dbg));
        DG_CUDA_CHECK(cudaMemsetAsync(dbg, 0, sizeof(long long) * ((size_t)ctx->sm_count * 4 * 16 + (size_t)p.n_tiles * 4),
                                      ctx->stream));
        p.dbg = dbg;
    }
    const size_t smem = fused_smem_bytes(m->fused_cp, p.cap_n, p.cap_nnz, has_hidden, wblob);
    // the tile counter is zero on entry: the last CTA of every launch resets it (no memset between launches)
    // Work-equivalent algorithmic bytes (SURVEY.md 8d / DESIGN.md): what the same layers would move if each
    // were a streaming pass - B_layer per hidden layer, a scalar SpMV pass for the first and the last layer,
    // one pass of the greedy search.  The fused kernel itself only reads CSR + weights and writes the
    // membership (see `traffic` in bench.py's roofline), so this figure measures work, not HBM pressure.
    {
        const double n = (double)b->n_nodes, nnz = (double)b->nnz, cp = (double)m->fused_cp;
        const double csr = 4.0 * (n + 1) + 4.0 * nnz;
        const double hidden = (double)std::max(0, m->n_layers - 2) * (csr + 4.0 * n + 8.0 * n * cp + 8.0 * cp * cp);
        const double scalar_passes = 2.0 * (csr + 12.0 * n);
        const double lgs = csr + 9.0 * n;
        ctx->last_kernel = "fused_solve_kernel";
        prof_begin(ctx);
        int st;
        if (m->fused_cp != 32)
            st = p.dit ? launch_t<64, true, false>(ctx, p, smem, p.n_tiles) : launch_t<64, false, false>(ctx, p, smem, p.n_tiles);
        else if (p.use_mma)
            st = p.dit ? launch_t<32, true, true>(ctx, p, smem, p.n_tiles) : launch_t<32, false, true>(ctx, p, smem, p.n_tiles);
        else
            st = p.dit ? launch_t<32, true, false>(ctx, p, smem, p.n_tiles) : launch_t<32, false, false>(ctx, p, smem, p.n_tiles);
        prof_end(ctx, hidden + scalar_passes + lgs);
        if (st != DG_OK) return st;
    }
    if (p.dbg) {  // debugging aid: print the phase timers of this launch
        std::vector<long long> h((size_t)ctx->sm_count * 4 * 16 + (size_t)p.n_tiles * 4);
        DG_CUDA_CHECK(cudaMemcpyAsync(h.data(), p.dbg, sizeof(long long) * h.size(), cudaMemcpyDeviceToHost, ctx->stream));
        DG_CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
        const char *names[10] = {"stage", "gather", "wait_g", "wait_w", "project", "wait_p", "tail", "lgs", "tiles", "total"};
        const int n_cta = std::min(p.n_tiles, ctx->sm_count * 4);
        for (int k = 0; k < 10; ++k) {
            long long mn = -1, mx = 0;
            double sum = 0;
            int cnt = 0;
            for (int c = 0; c < n_cta; ++c) {
                const long long v = h[(size_t)c * 16 + k];
                if (h[(size_t)c * 16 + 9] == 0) continue;
                mn = mn < 0 ? v : std::min(mn, v);
                mx = std::max(mx, v);
                sum += (double)v;
                ++cnt;
            }
            fprintf(stderr, "[fused timing] %-8s min %10lld avg %12.0f max %10lld (%d CTAs)\n", names[k], mn,
                    cnt ? sum / cnt : 0.0, mx, cnt);
        }
        if (const char *path = ctx->env.fused_tile_dump.empty() ? nullptr : ctx->env.fused_tile_dump.c_str()) {  // per-tile (cycles, rows, padded nnz, graphs)
            if (FILE *f = fopen(path, "w")) {
                const int grid = std::min(p.n_tiles, ctx->sm_count * 4);
                (void)grid;
                const size_t base = (size_t)std::min(p.n_tiles, ctx->sm_count) * 16;
                for (int t = 0; t < p.n_tiles; ++t)
                    fprintf(f, "%lld %lld %lld %lld\n", h[base + (size_t)t * 4], h[base + (size_t)t * 4 + 1],
                            h[base + (size_t)t * 4 + 2], h[base + (size_t)t * 4 + 3]);
                fclose(f);
            }
        }
    }
    *handled = true;
    return DG_OK;
}

}  // namespace dg
