// Graph-staged streaming GraphConvolution layer for BATCHES OF SMALL GRAPHS (32-wide layers).
//
//   H' = act( H.W_0 + L.(H.W_1) + b ),  (L Z)_i = Z_i - dinv_i * sum_{j in N(i)} dinv_j Z_j      gcn/layers.py:198-216
//
// evaluated as [H_i | (L H)_i] . [W_0 ; W_1] like gc_layer_kernel (dg_gcn.cu), but with the gather served from SHARED
// MEMORY: a tile is a run of consecutive graphs (rows v0 .. v0+n-1, all neighbours inside the run), its feature rows
// are staged once with cp.async (16-byte chunks, row stride 36 words so that both the row gather and the projection's
// operand reads are bank-conflict free), every neighbour row is then one 128-byte shared-memory wavefront instead of
// a trip to L2, and the output rows stream out with 128-bit stores.  HBM traffic per layer = the algorithmic
// B_layer of SURVEY.md 8d: CSR pattern + dinv + H in + H' out.
//
// One CTA = 8 warps, two CTAs per SM: while one CTA waits for its tile's rows the other computes (kGsBufs = 2 instead
// prefetches the next tile inside one 16-warp CTA; measured slower, see the constant).  A warp owns
// 16-row units: it gathers the unit (a lane group of 8 lanes per row, float4 per lane, four rows in flight, four
// partial sums per row combined pairwise - the summation tree of gc_layer_kernel), parks (L H)_i in its private
// scratch and projects the unit with a 4 x 4 register tile per lane (operands: 4 row chunks + 4 weight rows per 64
// FMAs).  Tiles are dealt to CTAs in contiguous, cost-balanced runs (host plan cached per batch).
//
// The kernel is bound by the shared-memory pipe, not HBM: per row it moves deg + ~13 wavefronts of 128 bytes through
// shared memory against 8 C + 4 deg bytes of HBM traffic (DESIGN.md section 3.2 has the arithmetic and the measurements).
#include <algorithm>

#include "dg_common.cuh"

namespace dg {

namespace {

constexpr int kGsThreads = 256;
constexpr int kGsWarps = kGsThreads / 32;
constexpr int kGsC = 32;            // feature width (padded)
constexpr int kGsStride = 36;       // words per staged row
constexpr int kGsUnit = 16;         // rows a warp gathers, then projects
constexpr int kGsMaxRows = 304;     // rows per tile (a multiple of kGsUnit)
constexpr int kGsMaxNnz = 10240;    // directed edges per tile (+ 4 of slack for the aligned copy)

struct GsArgs {
    const int4 *tiles;   // (v0, n, e0, nnz) per tile
    const int *cta_first;  // [grid + 1]: tiles of CTA c are cta_first[c] .. cta_first[c + 1] - 1
    LayerArgs a;
};

__device__ __forceinline__ float act_apply(float v, int act, float alpha) {
    if (act == DG_ACT_LEAKY_RELU) return v >= 0.f ? v : alpha * v;
    if (act == DG_ACT_RELU) return fmaxf(v, 0.f);
    return v;
}

__device__ __forceinline__ void cp_async16(void *smem, const void *gmem) {
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async4(void *smem, const void *gmem) {
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait_group() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory"); }

// one staged tile: feature rows, dinv, column ids, row_ptr
constexpr int kGsBufWords = kGsMaxRows * kGsStride + kGsMaxRows + (kGsMaxNnz + 8) + (kGsMaxRows + 8);
constexpr int kGsBufs = 1;   // 2 = prefetch the next tile inside the CTA (one 16-warp CTA per SM): measured SLOWER than two
                             // single-buffered 8-warp CTAs per SM (1290 vs 1106 us per layer on the config-4 batch): with one CTA
                             // the per-tile barrier idles the SM, with two the other CTA fills the gap
constexpr size_t gs_smem_bytes() {
    return sizeof(float) * ((size_t)kGsBufs * kGsBufWords
                            + (size_t)kGsWarps * kGsUnit * kGsStride  // per-warp (L H) scratch
                            + 2 * kGsC * kGsC + kGsC);                // [W_0 ; W_1], bias
}
// The stand-alone SpMM needs neither the scratch nor the weights, and keeps its column ids as 16-bit tile-local ids: 67 KB per
// CTA instead of 114 KB, i.e. THREE CTAs (24 warps) per SM instead of two - the kernel is latency-bound (barrier / shared-memory
// / global scoreboard stalls), so the extra warps are what it lacks.
constexpr int kGsSpmmCtas = 3;
constexpr int kGsSpmmBufWords = kGsMaxRows * kGsStride + kGsMaxRows + (kGsMaxNnz + 8) / 2 + (kGsMaxRows + 8);
constexpr size_t gs_spmm_smem_bytes() { return sizeof(float) * (size_t)kGsSpmmBufWords; }

// SPMM_ONLY: Y = L.Z alone (the reference's sparse_tensor_dense_matmul(support[1], pre_sup), gcn/layers.py:206) - no
// projection; the staged rows are pre-scaled by dinv once (G_j = dinv_j Z_j, two roundings per term like the unfused
// multiply-then-add of the reference's CPU kernel), so the gather is a plain sum of shared-memory rows, and the row's
// own value comes from global memory again (an L2 hit: the tile has just been read).
template <bool IMPLICIT_IN, bool TAIL, bool SPMM_ONLY = false>
__global__ void __launch_bounds__(kGsThreads, SPMM_ONLY ? kGsSpmmCtas : (kGsBufs == 1 ? 2 : 1)) gs_layer_kernel(const GsArgs P) {
    extern __shared__ __align__(16) unsigned char gs_smem[];
    float *buf0 = reinterpret_cast<float *>(gs_smem);                     // two staged tiles
    float *us_all = buf0 + kGsBufs * kGsBufWords;                         // [warps][16][36]
    float *ws = us_all + kGsWarps * kGsUnit * kGsStride;                  // [64][32]
    float *bs = ws + 2 * kGsC * kGsC;                                     // [32]

    const LayerArgs &a = P.a;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 3, q = lane & 7;   // lane group (row in flight) / 16-byte chunk of a row
    float *us = us_all + warp * (kGsUnit * kGsStride);

    if (!SPMM_ONLY) {
        for (int k = tid * 4; k < 2 * kGsC * kGsC; k += kGsThreads * 4)
            *reinterpret_cast<float4 *>(ws + k) = __ldg(reinterpret_cast<const float4 *>(a.wcat + k));
        if (tid < kGsC) bs[tid] = a.bias[tid];
    }

    // per-thread constants of the optional forms
    float4 ia0, ia1, ib;
    if (IMPLICIT_IN) {   // thread's chunk of the rank-1 first layer: H_j = act(x0_j a0 + s_j a1 + b)
        ia0 = __ldg(reinterpret_cast<const float4 *>(a.in_a0) + q);
        ia1 = __ldg(reinterpret_cast<const float4 *>(a.in_a1) + q);
        ib = __ldg(reinterpret_cast<const float4 *>(a.in_b) + q);
    }
    float4 tw0, tw1;
    if (TAIL) {
        tw0 = __ldg(reinterpret_cast<const float4 *>(a.tail_w0) + q);
        tw1 = __ldg(reinterpret_cast<const float4 *>(a.tail_w1) + q);
    }

    const int t_begin = P.cta_first[blockIdx.x], t_end = P.cta_first[blockIdx.x + 1];
    // stage tile t into buffer `which` (asynchronous copies; one commit group per tile)
    auto stage = [&](int t, int which) {
        float *hs = buf0 + which * (SPMM_ONLY ? kGsSpmmBufWords : kGsBufWords);
        float *dinv_s = hs + kGsMaxRows * kGsStride;
        int *cols = reinterpret_cast<int *>(dinv_s + kGsMaxRows);
        int *rp = cols + (SPMM_ONLY ? (kGsMaxNnz + 8) / 2 : kGsMaxNnz + 8);
        const int4 tile = __ldg(P.tiles + t);
        const int v0 = tile.x, n = tile.y, e0 = tile.z, nnz = tile.w;
        if (IMPLICIT_IN) {
            for (int item = tid; item < n * 8; item += kGsThreads) {   // item & 7 == q for every item of this thread
                const int r = item >> 3;
                const float2 p = __ldg(a.pair_in + v0 + r);
                float4 v;
                v.x = act_apply(fmaf(p.y, ia1.x, fmaf(p.x, ia0.x, ib.x)), a.in_act, a.alpha);
                v.y = act_apply(fmaf(p.y, ia1.y, fmaf(p.x, ia0.y, ib.y)), a.in_act, a.alpha);
                v.z = act_apply(fmaf(p.y, ia1.z, fmaf(p.x, ia0.z, ib.z)), a.in_act, a.alpha);
                v.w = act_apply(fmaf(p.y, ia1.w, fmaf(p.x, ia0.w, ib.w)), a.in_act, a.alpha);
                *reinterpret_cast<float4 *>(hs + r * kGsStride + 4 * q) = v;
            }
        } else {
            const float *src = a.hin + (size_t)v0 * kGsC;
            for (int item = tid; item < n * 8; item += kGsThreads)
                cp_async16(hs + (item >> 3) * kGsStride + 4 * (item & 7), src + (size_t)item * 4);
        }
        const int ea = e0 & ~3;            // 16-byte aligned start of the column window around [e0, e0 + nnz)
        const int total = nnz + (e0 - ea);
        const int chunks = total >> 2;     // whole 16-byte chunks; the last ids one by one (never past the end of col_idx)
        const int *csrc = a.col_idx + ea;
        if (SPMM_ONLY) {   // 16-bit tile-local ids (a tile has at most 304 rows): loaded, narrowed, stored
            uint16_t *c16 = reinterpret_cast<uint16_t *>(cols);
            for (int c = tid; c < chunks; c += kGsThreads) {
                const int4 v = __ldg(reinterpret_cast<const int4 *>(csrc) + c);
                // (ids of the previous tile's last rows in front of e0 fall below v0: never read, any value will do)
                *reinterpret_cast<uint2 *>(c16 + 4 * c) = make_uint2(((unsigned)(v.x - v0) & 0xffffu) | ((unsigned)(v.y - v0) << 16),
                                                                      ((unsigned)(v.z - v0) & 0xffffu) | ((unsigned)(v.w - v0) << 16));
            }
            for (int k = 4 * chunks + tid; k < total; k += kGsThreads) c16[k] = (uint16_t)(__ldg(csrc + k) - v0);
        } else {
            for (int c = tid; c < chunks; c += kGsThreads) cp_async16(cols + 4 * c, csrc + 4 * c);
            for (int k = 4 * chunks + tid; k < total; k += kGsThreads) cp_async4(cols + k, csrc + k);
        }
        for (int r = tid; r <= n; r += kGsThreads) cp_async4(rp + r, a.row_ptr + v0 + r);
        for (int r = tid; r < n; r += kGsThreads) cp_async4(dinv_s + r, a.dinv + v0 + r);
        cp_async_commit();
    };
    if (kGsBufs == 2 && t_begin < t_end) stage(t_begin, 0);
    for (int t = t_begin; t < t_end; ++t) {
        const int which = kGsBufs == 2 ? ((t - t_begin) & 1) : 0;
        if (kGsBufs == 2 && t + 1 < t_end) {
            stage(t + 1, which ^ 1);       // the other buffer was released by the barrier that ended the previous tile
            cp_async_wait_group<1>();      // everything but the group just committed has landed: tile t is here
        } else {
            if (kGsBufs == 1) stage(t, 0);
            cp_async_wait_group<0>();
        }
        __syncthreads();
        float *hs = buf0 + which * (SPMM_ONLY ? kGsSpmmBufWords : kGsBufWords);
        float *dinv_s = hs + kGsMaxRows * kGsStride;
        int *cols = reinterpret_cast<int *>(dinv_s + kGsMaxRows);
        const uint16_t *cols16 = reinterpret_cast<const uint16_t *>(cols);
        int *rp = cols + (SPMM_ONLY ? (kGsMaxNnz + 8) / 2 : kGsMaxNnz + 8);
        const int4 tile = __ldg(P.tiles + t);
        const int v0 = tile.x, n = tile.y, e0 = tile.z;
        const int ea = e0 & ~3;
        if (SPMM_ONLY) {   // G_j = dinv_j Z_j in place
            for (int item = tid; item < n * 8; item += kGsThreads) {
                const int r = item >> 3;
                float4 *p = reinterpret_cast<float4 *>(hs + r * kGsStride + 4 * (item & 7));
                const float d = dinv_s[r];
                float4 v = *p;
                v.x *= d, v.y *= d, v.z *= d, v.w *= d;
                *p = v;
            }
            __syncthreads();
        }

        // ---- units of 16 rows: gather, then project ---------------------------------------------------------------
        const int n_units = (n + kGsUnit - 1) / kGsUnit;
        for (int u = warp; u < n_units; u += kGsWarps) {
            const int ubase = u * kGsUnit;
#pragma unroll 1
            for (int pass = 0; pass < kGsUnit / 4; ++pass) {
                const int lr = pass * 4 + g;      // row inside the unit
                const int r = ubase + lr;
                const bool valid = r < n;
                int beg = 0, end = 0;
                if (valid) {
                    beg = rp[r] - ea;             // positions inside the staged window
                    end = rp[r + 1] - ea;
                }
                float4 acc[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) acc[k] = make_float4(0.f, 0.f, 0.f, 0.f);
                // the ids (and dinv) of the NEXT eight edges are fetched while the current eight rows are being summed
                int jl = 0;
                float d = 0.f;
                if (beg + q < end) {
                    jl = SPMM_ONLY ? (int)cols16[beg + q] : cols[beg + q] - v0;      // neighbour's row inside the tile
                    if (!SPMM_ONLY) d = dinv_s[jl];
                }
                for (int e = beg; __any_sync(0xffffffffu, e < end); e += 8) {
                    int jl_next = 0;
                    float d_next = 0.f;
                    if (e + 8 + q < end) {
                        jl_next = SPMM_ONLY ? (int)cols16[e + 8 + q] : cols[e + 8 + q] - v0;
                        if (!SPMM_ONLY) d_next = dinv_s[jl_next];
                    }
#pragma unroll
                    for (int tt = 0; tt < 8; ++tt) {
                        const int jj = __shfl_sync(0xffffffffu, jl, tt, 8);
                        float dj = 1.f;
                        if (!SPMM_ONLY) dj = __shfl_sync(0xffffffffu, d, tt, 8);
                        if (e + tt < end) {       // uniform inside a lane group
                            const float4 v = *reinterpret_cast<const float4 *>(hs + jj * kGsStride + 4 * q);
                            if (SPMM_ONLY) {
                                acc[tt & 3].x += v.x, acc[tt & 3].y += v.y, acc[tt & 3].z += v.z, acc[tt & 3].w += v.w;
                            } else {
                                acc[tt & 3].x = fmaf(dj, v.x, acc[tt & 3].x);
                                acc[tt & 3].y = fmaf(dj, v.y, acc[tt & 3].y);
                                acc[tt & 3].z = fmaf(dj, v.z, acc[tt & 3].z);
                                acc[tt & 3].w = fmaf(dj, v.w, acc[tt & 3].w);
                            }
                        }
                    }
                    jl = jl_next;
                    d = d_next;
                }
                acc[0].x += acc[1].x, acc[0].y += acc[1].y, acc[0].z += acc[1].z, acc[0].w += acc[1].w;
                acc[2].x += acc[3].x, acc[2].y += acc[3].y, acc[2].z += acc[3].z, acc[2].w += acc[3].w;
                acc[0].x += acc[2].x, acc[0].y += acc[2].y, acc[0].z += acc[2].z, acc[0].w += acc[2].w;
                float4 lh = make_float4(0.f, 0.f, 0.f, 0.f);
                if (valid) {
                    const float di = dinv_s[r];
                    const float4 hi = SPMM_ONLY ? __ldcg(reinterpret_cast<const float4 *>(a.hin + (size_t)(v0 + r) * kGsC) + q)
                                                : *reinterpret_cast<const float4 *>(hs + r * kGsStride + 4 * q);
                    lh = make_float4(fmaf(-di, acc[0].x, hi.x), fmaf(-di, acc[0].y, hi.y), fmaf(-di, acc[0].z, hi.z),
                                     fmaf(-di, acc[0].w, hi.w));
                    if (SPMM_ONLY) *reinterpret_cast<float4 *>(a.hout + (size_t)(v0 + r) * kGsC + 4 * q) = lh;
                }
                if (!SPMM_ONLY) *reinterpret_cast<float4 *>(us + lr * kGsStride + 4 * q) = lh;
            }
            if (SPMM_ONLY) continue;
            __syncwarp();
            // projection: lane (g, q) owns rows 4 j + g (j = 0..3) of the unit and columns 4 q .. 4 q + 3
            float4 out[4];
            {
                const float4 b4 = *reinterpret_cast<const float4 *>(bs + 4 * q);
#pragma unroll
                for (int j = 0; j < 4; ++j) out[j] = b4;
            }
            // rows past the end of the tile read row n - 1 (their results are never stored)
            int hrow[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) hrow[j] = min(ubase + 4 * j + g, n - 1) * kGsStride;
#pragma unroll 2
            for (int k4 = 0; k4 < 16; ++k4) {    // k = 4 k4 .. 4 k4 + 3 of [H_i | (L H)_i]
                float4 uu[4], ww[4];
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    uu[j] = k4 < 8 ? *reinterpret_cast<const float4 *>(hs + hrow[j] + 4 * k4)
                                   : *reinterpret_cast<const float4 *>(us + (4 * j + g) * kGsStride + 4 * (k4 - 8));
#pragma unroll
                for (int e = 0; e < 4; ++e) ww[e] = *reinterpret_cast<const float4 *>(ws + (4 * k4 + e) * kGsC + 4 * q);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    out[j].x = fmaf(uu[j].x, ww[0].x, out[j].x), out[j].y = fmaf(uu[j].x, ww[0].y, out[j].y);
                    out[j].z = fmaf(uu[j].x, ww[0].z, out[j].z), out[j].w = fmaf(uu[j].x, ww[0].w, out[j].w);
                    out[j].x = fmaf(uu[j].y, ww[1].x, out[j].x), out[j].y = fmaf(uu[j].y, ww[1].y, out[j].y);
                    out[j].z = fmaf(uu[j].y, ww[1].z, out[j].z), out[j].w = fmaf(uu[j].y, ww[1].w, out[j].w);
                    out[j].x = fmaf(uu[j].z, ww[2].x, out[j].x), out[j].y = fmaf(uu[j].z, ww[2].y, out[j].y);
                    out[j].z = fmaf(uu[j].z, ww[2].z, out[j].z), out[j].w = fmaf(uu[j].z, ww[2].w, out[j].w);
                    out[j].x = fmaf(uu[j].w, ww[3].x, out[j].x), out[j].y = fmaf(uu[j].w, ww[3].y, out[j].y);
                    out[j].z = fmaf(uu[j].w, ww[3].z, out[j].z), out[j].w = fmaf(uu[j].w, ww[3].w, out[j].w);
                }
            }
            // epilogue
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int r = ubase + 4 * j + g;
                float4 h;
                h.x = act_apply(out[j].x, a.act, a.alpha), h.y = act_apply(out[j].y, a.act, a.alpha);
                h.z = act_apply(out[j].z, a.act, a.alpha), h.w = act_apply(out[j].w, a.act, a.alpha);
                if (!TAIL) {
                    if (r < n) *reinterpret_cast<float4 *>(a.hout + (size_t)(v0 + r) * kGsC + 4 * q) = h;
                } else {
                    float t0 = h.x * tw0.x, t1 = h.x * tw1.x;
                    t0 = fmaf(h.y, tw0.y, t0), t1 = fmaf(h.y, tw1.y, t1);
                    t0 = fmaf(h.z, tw0.z, t0), t1 = fmaf(h.z, tw1.z, t1);
                    t0 = fmaf(h.w, tw0.w, t0), t1 = fmaf(h.w, tw1.w, t1);
#pragma unroll
                    for (int off = 4; off > 0; off >>= 1) {
                        t0 += __shfl_xor_sync(0xffffffffu, t0, off);
                        t1 += __shfl_xor_sync(0xffffffffu, t1, off);
                    }
                    if (q == 0 && r < n) {
                        a.tail_q[v0 + r] = t0 + t1;
                        a.tail_zs[v0 + r] = dinv_s[r] * t1;
                    }
                }
            }
            __syncwarp();  // the scratch is rewritten by the warp's next unit
        }
        __syncthreads();   // every warp is done with the staged tile
    }
}

// ---- host: tile plan -----------------------------------------------------------------------------------------------
int gs_build_plan(dg_context *ctx, dg_batch *b, bool *ok) {
    *ok = false;
    if (b->gs_valid) {
        *ok = b->gs_n_tiles > 0;
        return DG_OK;
    }
    b->gs_valid = true;
    b->gs_n_tiles = 0;
    if ((int)b->h_graph_e.size() != b->n_graphs + 1 || b->n_graphs == 0) return DG_OK;
    if (b->max_graph_nodes > kGsMaxRows || b->max_graph_nnz > kGsMaxNnz) return DG_OK;  // a graph does not fit a tile
    const auto &gp = b->h_graph_ptr;
    const auto &ge = b->h_graph_e;
    std::vector<int> flat;   // (v0, n, e0, nnz) per tile
    std::vector<long long> cost;
    flat.reserve((size_t)b->n_graphs * 4);
    int v0 = 0, e0 = 0, n = 0, nnz = 0;
    auto flush = [&]() {
        if (n > 0) {
            flat.push_back(v0), flat.push_back(n), flat.push_back(e0), flat.push_back(nnz);
            const int units = (n + kGsUnit - 1) / kGsUnit;
            // a warp round costs about a unit's gather + projection; the CTA's warps work in rounds
            cost.push_back((long long)nnz + 14LL * kGsUnit * ((units + kGsWarps - 1) / kGsWarps) * kGsWarps + 600);
        }
        n = 0, nnz = 0;
    };
    for (int g = 0; g < b->n_graphs; ++g) {
        const int gn = gp[g + 1] - gp[g], gz = ge[g + 1] - ge[g];
        if (gn == 0) continue;
        if (n > 0 && (n + gn > kGsMaxRows || nnz + gz > kGsMaxNnz)) flush();
        if (n == 0) v0 = gp[g], e0 = ge[g];
        n += gn, nnz += gz;
    }
    flush();
    const int n_tiles = (int)cost.size();
    if (n_tiles == 0) return DG_OK;
    std::vector<long long> prefix((size_t)n_tiles + 1, 0);
    for (int t = 0; t < n_tiles; ++t) prefix[(size_t)t + 1] = prefix[(size_t)t] + cost[(size_t)t];
    const long long total = prefix.back();
    // contiguous runs of tiles with balanced cost: CTA c starts at the first tile whose cost prefix reaches c / grid
    auto runs = [&](int grid, std::vector<int> *first) {
        first->assign((size_t)grid + 1, n_tiles);
        (*first)[0] = 0;
        for (int c = 1; c < grid; ++c) {
            const long long target = total * c / grid;
            int t = (int)(std::lower_bound(prefix.begin(), prefix.end(), target) - prefix.begin());
            // the boundary goes to the nearer side of the target
            if (t > 0 && t <= n_tiles && target - prefix[(size_t)t - 1] < prefix[(size_t)std::min(t, n_tiles)] - target) --t;
            (*first)[(size_t)c] = std::min(std::max(t, (*first)[(size_t)c - 1]), n_tiles);
        }
        (*first)[(size_t)grid] = n_tiles;
    };
    const int grid = std::min(n_tiles, ctx->sm_count * (kGsBufs == 1 ? 2 : 1));
    const int grid_spmm = std::min(n_tiles, ctx->sm_count * kGsSpmmCtas);   // the stand-alone SpMM runs three CTAs per SM
    std::vector<int> first, first_spmm;
    runs(grid, &first);
    runs(grid_spmm, &first_spmm);
    first.insert(first.end(), first_spmm.begin(), first_spmm.end());
    const size_t words = flat.size() + first.size();
    if (b->gs_tiles_cap < words || !b->gs_tiles_dev) {
        if (b->gs_tiles_dev) {
            DG_CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
            cudaFree(b->gs_tiles_dev);
            b->gs_tiles_dev = nullptr;
        }
        b->gs_tiles_cap = words + words / 4 + 16;
        DG_CUDA_CHECK(cudaMalloc((void **)&b->gs_tiles_dev, sizeof(int) * b->gs_tiles_cap));
    }
    DG_CUDA_CHECK(cudaMemcpyAsync(b->gs_tiles_dev, flat.data(), sizeof(int) * flat.size(), cudaMemcpyHostToDevice,
                                  ctx->stream));
    DG_CUDA_CHECK(cudaMemcpyAsync(b->gs_tiles_dev + flat.size(), first.data(), sizeof(int) * first.size(),
                                  cudaMemcpyHostToDevice, ctx->stream));
    DG_CUDA_CHECK(cudaStreamSynchronize(ctx->stream));  // the pageable sources die with this call
    b->gs_n_tiles = n_tiles;
    b->gs_grid = grid;
    b->gs_grid_spmm = grid_spmm;
    *ok = true;
    return DG_OK;
}

template <bool IMPLICIT_IN, bool TAIL, bool SPMM_ONLY = false>
int gs_launch(dg_context *ctx, dg_batch *b, const LayerArgs &a) {
    auto kern = gs_layer_kernel<IMPLICIT_IN, TAIL, SPMM_ONLY>;
    constexpr size_t smem = SPMM_ONLY ? gs_spmm_smem_bytes() : gs_smem_bytes();
    static std::atomic<unsigned long long> attr_done{0};
    DG_CUDA_CHECK(smem_attr_once(kern, ctx->device, (int)smem, &attr_done));
    GsArgs P;
    P.tiles = reinterpret_cast<const int4 *>(b->gs_tiles_dev);
    // two run tables follow the tiles: for the layers' grid, then for the stand-alone SpMM's
    P.cta_first = b->gs_tiles_dev + (size_t)b->gs_n_tiles * 4 + (SPMM_ONLY ? (size_t)b->gs_grid + 1 : 0);
    P.a = a;
    const double n = (double)a.n, nnz = (double)a.nnz;
    // B_layer / B_spmm of SURVEY.md 8d
    const double bytes = 4.0 * (n + 1) + 4.0 * nnz + 4.0 * n + 4.0 * n * (IMPLICIT_IN ? 2 : kGsC) +
                         4.0 * n * (TAIL ? 2 : kGsC) + (SPMM_ONLY ? 0.0 : 4.0 * (2 * kGsC * kGsC + kGsC));
    ctx->last_kernel = SPMM_ONLY ? "gs_spmm_kernel" : "gs_layer_kernel";
    prof_begin(ctx);
    kern<<<SPMM_ONLY ? b->gs_grid_spmm : b->gs_grid, kGsThreads, smem, ctx->stream>>>(P);
    prof_end(ctx, bytes);
    ctx->launches++;
    DG_CUDA_CHECK(cudaGetLastError());
    return DG_OK;
}

}  // namespace

int gs_try_spmm(dg_context *ctx, dg_batch *b, int width, const float *z, float *y, bool *handled) {
    *handled = false;
    if (width != kGsC || ctx->env.disable_staged || b->n_graphs < 8) return DG_OK;
    bool ok = false;
    DG_TRY(gs_build_plan(ctx, b, &ok));
    if (!ok) return DG_OK;
    LayerArgs a{};
    a.n = b->n_nodes;
    a.nnz = b->nnz;
    a.row_ptr = b->row_ptr;
    a.col_idx = b->col_idx;
    a.dinv = b->dinv;
    a.hin = z;
    a.hout = y;
    DG_TRY((gs_launch<false, false, true>(ctx, b, a)));
    *handled = true;
    return DG_OK;
}

int gs_try_layer(dg_context *ctx, dg_batch *b, int cpi, int cpo, bool implicit_in, bool tail, const LayerArgs &a,
                 bool *handled) {
    *handled = false;
    if (cpi != kGsC || cpo != kGsC || a.row0 != 0 || a.pm.world > 1) return DG_OK;
    if (ctx->env.disable_staged) return DG_OK;
    if (a.n != b->n_nodes || a.row_ptr != b->row_ptr) return DG_OK;
    // a handful of large graphs gain nothing from staging; the plan needs every graph to fit a tile
    if (b->n_graphs < 8) return DG_OK;
    bool ok = false;
    DG_TRY(gs_build_plan(ctx, b, &ok));
    if (!ok) return DG_OK;
    int st;
    if (implicit_in && tail) st = gs_launch<true, true>(ctx, b, a);
    else if (implicit_in) st = gs_launch<true, false>(ctx, b, a);
    else if (tail) st = gs_launch<false, true>(ctx, b, a);
    else st = gs_launch<false, false>(ctx, b, a);
    if (st != DG_OK) return st;
    *handled = true;
    return DG_OK;
}

}  // namespace dg
