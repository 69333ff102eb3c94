// GraphConvolution stack on a packed CSR batch - sm_100a kernels and their driver.
//
// What the reference computes per layer (gcn/layers.py:198-216, supports from gcn/utils.py:258-274):
//     H' = act( H.W_0 + L.(H.W_1) + b ),      L = I - D^-1/2 A D^-1/2
// L is never materialised here: with dinv = fp32(deg^-1/2) (0 for isolated / removed vertices)
//     (L.Z)_i = Z_i - dinv_i * sum_{j in N(i)} dinv_j * Z_j .
// The fused layer kernel aggregates first and projects second,
//     H'_i = act( [H_i | (L.H)_i] . [W_0 ; W_1] + b ),
// which is the same linear map as the reference's project-then-aggregate order (L.(H.W_1) ==
// (L.H).W_1) but reads and writes each feature row once (B_layer of SURVEY.md 8d).
//
// Structure of a forward pass (n_supports == 2, "cheb1"):
//   dinv (batch, once)                degree_kernel
//   layer 0, rank-1: every feature column of the reference's input equals x0_i, so
//       H1_i = act(x0_i * colsum(W_0) + s_i * colsum(W_1) + b),  s = L.x0
//     needs only the scalar SpMV s                                  first_scalar_kernel
//     and H1 is never written: consumers rebuild it from (x0_j, s_j) on the fly  (IMPLICIT_IN)
//   layers 1 .. : gc_layer_kernel<CPI, CPO, IMPLICIT_IN, TAIL>
//   last layer with one output column (diver_num == 1): projected BEFORE aggregation like the
//     reference does - the preceding kernel's epilogue emits q_i = H_i.w_0 + z_i and
//     zs_i = dinv_i * z_i (z = H.w_1) (TAIL), leaving the scalar SpMV     last_scalar_kernel
//     fused with the fp64 utility product of mwis_dqn_call.py:232.
#include <math.h>

#include <utility>

#include "dg_common.cuh"

namespace dg {

namespace {

constexpr int kWarpsPerCta = 8;
constexpr int kRowsPerWarp = 8;   // R: rows a warp aggregates, then projects together

__device__ __forceinline__ float act_apply(float v, int act, float alpha) {
    if (act == DG_ACT_LEAKY_RELU) return v >= 0.f ? v : alpha * v;
    if (act == DG_ACT_RELU) return fmaxf(v, 0.f);
    return v;
}

// ---------------------------------------------------------------------------------------------
// degrees -> dinv.  Follows gcn/utils.py:122-125 (rowsum^-0.5 in fp64, inf -> 0), rounded to fp32.
// With a keep mask the degree is taken on the kept sub-graph (mwis_dqn_call.py:202-207).
// ---------------------------------------------------------------------------------------------
// ---------------------------------------------------------------------------------------------
// CSR-stream scalar gather-sum: the shape of every scalar pass over the graph (degrees on a kept
// sub-graph, s = L.x0, the one-column last layer).  A CTA owns kStreamRows consecutive rows; their
// edges are one contiguous run of col_idx, which the CTA streams in tiles: every thread loads
// kStreamPer column ids with coalesced 4-byte loads and issues the kStreamPer gathers of the source
// value back to back (independent chains: the loads in flight per SM are what a random gather needs
// to approach the L2 / HBM sector rate), parks the values in shared memory, and after a barrier thread t
// sums the part of row t that lies inside the tile, in ascending edge order (the order in which the
// reference's COO SpMM accumulates, gcn/layers.py:206).  The shared array is skewed by one word per 32
// so that rows of equal length (stride = degree) do not collide on a bank.
// Row-slice form (row0 != 0): the launch covers rows row0 .. row0+n-1 of a larger graph; row_ptr is the
// slice's own (local) array, every per-vertex array is indexed by GLOBAL vertex id.
// ---------------------------------------------------------------------------------------------
constexpr int kStreamThreads = 256;
constexpr int kStreamRows = 256;
constexpr int kStreamPer = 8;
constexpr int kStreamTile = kStreamThreads * kStreamPer;

__device__ __forceinline__ int stream_skew(int i) { return i + (i >> 5); }

// acc = sum over the edges e of row (r0 + threadIdx.x) of fetch(col_idx[e]); T = float or int
template <typename T, typename Fetch>
__device__ __forceinline__ T csr_stream_sum(int n, const int *__restrict__ row_ptr, const int *__restrict__ col_idx,
                                            T *vals, Fetch fetch) {
    const int t = threadIdx.x;
    const int r0 = blockIdx.x * kStreamRows;
    const int nr = min(kStreamRows, n - r0);
    const int e0 = row_ptr[r0], e1 = row_ptr[r0 + nr];
    int rb = 0, re = 0;
    if (t < nr) {
        rb = row_ptr[r0 + t];
        re = row_ptr[r0 + t + 1];
    }
    T acc = T(0);
    for (int c0 = e0 & ~31; c0 < e1; c0 += kStreamTile) {
        int c[kStreamPer];
        T v[kStreamPer];
#pragma unroll
        for (int k = 0; k < kStreamPer; ++k) {
            const int e = c0 + k * kStreamThreads + t;
            c[k] = (e >= e0 && e < e1) ? __ldg(col_idx + e) : -1;
        }
#pragma unroll
        for (int k = 0; k < kStreamPer; ++k) v[k] = c[k] >= 0 ? fetch(c[k]) : T(0);
#pragma unroll
        for (int k = 0; k < kStreamPer; ++k) vals[stream_skew(k * kStreamThreads + t)] = v[k];
        __syncthreads();
        const int lo = max(rb, c0) - c0, hi = min(re, c0 + kStreamTile) - c0;
        for (int i = lo; i < hi; ++i) acc += vals[stream_skew(i)];
        __syncthreads();
    }
    return acc;
}

// degrees -> dinv.  Follows gcn/utils.py:122-125 (rowsum^-0.5 in fp64, inf -> 0), rounded to fp32.
// With a keep mask the degree is taken on the kept sub-graph (mwis_dqn_call.py:202-207).
__global__ void degree_kernel(int n, int row0, const int *__restrict__ row_ptr, float *__restrict__ dinv,
                              const PeerMap pm) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int deg = row_ptr[i + 1] - row_ptr[i];
    peer_store(pm, dinv + row0 + i, deg > 0 ? (float)(1.0 / sqrt((double)deg)) : 0.f);
}

__global__ void __launch_bounds__(kStreamThreads)
degree_keep_kernel(int n, int row0, const int *__restrict__ row_ptr, const int *__restrict__ col_idx,
                   const uint8_t *__restrict__ keep, float *__restrict__ dinv, const PeerMap pm) {
    __shared__ int vals[kStreamTile + kStreamTile / 32 + 1];
    const int deg = csr_stream_sum<int>(n, row_ptr, col_idx, vals, [&](int c) { return (int)(__ldg(keep + c) != 0); });
    const int i = blockIdx.x * kStreamRows + threadIdx.x;
    if (i < n) peer_store(pm, dinv + row0 + i, (deg > 0 && keep[row0 + i]) ? (float)(1.0 / sqrt((double)deg)) : 0.f);
}

__global__ void keep_from_weights_kernel(int n, const double *__restrict__ wts, uint8_t *__restrict__ keep) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) keep[i] = wts[i] != 0.0;  // rm_nodes = where(wts == 0), mwis_dqn_call.py:203
}

// row-partitioned runs: keep[v] = v is a real vertex (padding rows past n_real are not) and, with zero-weight
// removal, wts[v] != 0 (mwis_dqn_call.py:203); stored to every rank's arena
__global__ void part_keep_kernel(int n_local, int row0, int n_real, const double *__restrict__ wts, int remove_zero,
                                 uint8_t *__restrict__ keep, const PeerMap pm) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_local) return;
    const int v = row0 + i;
    const bool k = v < n_real && (!remove_zero || wts[v] != 0.0);
    peer_store(pm, keep + v, (uint8_t)(k ? 1 : 0));
}

// y_j = dinv_j * x0_j : the quantity the first layer's scalar SpMV gathers
__global__ void scaled_input_kernel(int n, const float *__restrict__ dinv, const uint8_t *__restrict__ keep,
                                    const float *__restrict__ x0, float x0val, float *__restrict__ y,
                                    const PeerMap pm) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float xi = (keep && !keep[i]) ? 0.f : (x0 ? x0[i] : x0val);
    peer_store(pm, y + i, dinv[i] * xi);
}

// s = L.x0 (scalar SpMV, CSR-stream).  Writes (x0_i, s_i).  Row-slice form: dinv / keep / x0 / pair_out are
// already offset to the slice's first row, y is the global array.
__global__ void __launch_bounds__(kStreamThreads)
first_scalar_kernel(int n, const int *__restrict__ row_ptr, const int *__restrict__ col_idx,
                    const float *__restrict__ dinv, const float *__restrict__ y,
                    const uint8_t *__restrict__ keep, const float *__restrict__ x0, float x0val,
                    float2 *__restrict__ pair_out, const PeerMap pm) {
    __shared__ float vals[kStreamTile + kStreamTile / 32 + 1];
    const float acc = csr_stream_sum<float>(n, row_ptr, col_idx, vals, [&](int c) { return __ldg(y + c); });
    const int row = blockIdx.x * kStreamRows + threadIdx.x;
    if (row < n) {
        float xi = (keep && !keep[row]) ? 0.f : (x0 ? x0[row] : x0val);
        peer_store(pm, pair_out + row, make_float2(xi, xi - dinv[row] * acc));
    }
}

// Single-layer models (num_layer == 1, gcn/models.py:539-548): out = act(x0*a0 + s*a1 + b).
__global__ void first_out_kernel(int n, int d_out, const float2 *__restrict__ pair,
                                 const float *__restrict__ a0, const float *__restrict__ a1,
                                 const float *__restrict__ bias, int act, float alpha,
                                 const uint8_t *__restrict__ keep, float *__restrict__ out) {
    int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n * d_out) return;
    int i = idx / d_out, d = idx - i * d_out;
    float2 p = pair[i];
    float v = act_apply(fmaf(p.y, a1[d], fmaf(p.x, a0[d], bias[d])), act, alpha);
    out[idx] = (keep && !keep[i]) ? 0.f : v;
}

// Two-layer models with one output (e.g. c64 l2): the hidden layer is the implicit rank-1 one, so the
// last layer's projections are per-vertex work: q_i = H1_i.w_0 + z_i, zs_i = dinv_i * z_i, z = H1.w_1.
__global__ void __launch_bounds__(256)
node_project_kernel(int n, int c, const float2 *__restrict__ pair, const float *__restrict__ a0,
                    const float *__restrict__ a1, const float *__restrict__ b0, int act, float alpha,
                    const float *__restrict__ w0, const float *__restrict__ w1,
                    const float *__restrict__ dinv, float *__restrict__ q_out, float *__restrict__ zs_out,
                    const PeerMap pm) {
    __shared__ __align__(16) float sm[5 * kMaxWidth];
    for (int k = threadIdx.x; k < kMaxWidth; k += blockDim.x) {  // columns past c carry zeros: no contribution
        const bool in = k < c;
        sm[k] = in ? a0[k] : 0.f;
        sm[kMaxWidth + k] = in ? a1[k] : 0.f;
        sm[2 * kMaxWidth + k] = in ? b0[k] : 0.f;
        sm[3 * kMaxWidth + k] = in ? w0[k] : 0.f;
        sm[4 * kMaxWidth + k] = in ? w1[k] : 0.f;
    }
    __syncthreads();
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float2 p = pair[i];
    float t0 = 0.f, t1 = 0.f;
    const float4 *s4 = reinterpret_cast<const float4 *>(sm);
    const int c4 = (c + 3) / 4;
    for (int k = 0; k < c4; ++k) {  // 128-bit broadcast loads: the loop is LSU-bound with scalar ones
        const float4 va0 = s4[k], va1 = s4[kMaxWidth / 4 + k], vb = s4[2 * (kMaxWidth / 4) + k];
        const float4 vw0 = s4[3 * (kMaxWidth / 4) + k], vw1 = s4[4 * (kMaxWidth / 4) + k];
        float h;
        h = act_apply(fmaf(p.y, va1.x, fmaf(p.x, va0.x, vb.x)), act, alpha);
        t0 = fmaf(h, vw0.x, t0), t1 = fmaf(h, vw1.x, t1);
        h = act_apply(fmaf(p.y, va1.y, fmaf(p.x, va0.y, vb.y)), act, alpha);
        t0 = fmaf(h, vw0.y, t0), t1 = fmaf(h, vw1.y, t1);
        h = act_apply(fmaf(p.y, va1.z, fmaf(p.x, va0.z, vb.z)), act, alpha);
        t0 = fmaf(h, vw0.z, t0), t1 = fmaf(h, vw1.z, t1);
        h = act_apply(fmaf(p.y, va1.w, fmaf(p.x, va0.w, vb.w)), act, alpha);
        t0 = fmaf(h, vw0.w, t0), t1 = fmaf(h, vw1.w, t1);
    }
    q_out[i] = t0 + t1;
    peer_store(pm, zs_out + i, dinv[i] * t1);
}

// Last layer with one output column: score_i = act(q_i - dinv_i * sum_j zs_j + b), then the utility
// product of mwis_dqn_call.py:230-235 in fp64 (CSR-stream).  q and zs are separate planes: only zs is
// gathered (and, row-partitioned, exchanged between ranks).
__global__ void __launch_bounds__(kStreamThreads)
last_scalar_kernel(int n, int row0, const int *__restrict__ row_ptr, const int *__restrict__ col_idx,
                   const float *__restrict__ dinv, const float *__restrict__ q, const float *__restrict__ zs,
                   float bias, int act, float alpha, const uint8_t *__restrict__ keep, float *__restrict__ score,
                   const double *__restrict__ wts, int predict, double *__restrict__ util, const PeerMap pm) {
    __shared__ float vals[kStreamTile + kStreamTile / 32 + 1];
    const float acc = csr_stream_sum<float>(n, row_ptr, col_idx, vals, [&](int c) { return __ldg(zs + c); });
    const int row = blockIdx.x * kStreamRows + threadIdx.x;
    if (row < n) {
        const int gr = row0 + row;
        float v = act_apply(q[gr] - dinv[gr] * acc + bias, act, alpha);
        if (keep && !keep[gr]) v = 0.f;
        if (score) score[gr] = v;
        if (util) peer_store(pm, util + gr, (predict == DG_PREDICT_MWIS) ? (double)v * wts[gr] : (double)v);
    }
}

__global__ void utility_kernel(int n, const float *__restrict__ score, int stride,
                               const double *__restrict__ wts, int predict, double *__restrict__ util) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double a = (double)score[(size_t)i * stride];
    util[i] = (predict == DG_PREDICT_MWIS) ? a * wts[i] : a;
}

// ---------------------------------------------------------------------------------------------
// The fused hidden-layer kernel.
//   CPI / CPO   padded input / output widths (32 or 64); padding columns carry zeros
//   IMPLICIT_IN input rows are rebuilt from (x0_j, s_j) and the first layer's column sums
//   TAIL        instead of writing H', emit (q, zs) for a following one-column last layer
// A warp owns kRowsPerWarp consecutive rows.  Aggregation: CPI/4 lanes cover one neighbour row with
// 128-bit loads, so 32/(CPI/4) neighbours are in flight per step; partial sums are combined with
// warp shuffles.  Projection: the warp's rows sit in shared memory, lane c owns output column c
// (and c+32), weights are read from shared memory once per 8 rows.
// ---------------------------------------------------------------------------------------------

template <int CPI, int CPO, bool IMPLICIT_IN, bool TAIL>
__global__ void __launch_bounds__(kWarpsPerCta * 32)
gc_layer_kernel(const LayerArgs a) {
    constexpr int R = kRowsPerWarp;
    constexpr int LPR = CPI / 4;    // lanes per feature row
    constexpr int NG = 32 / LPR;    // neighbour rows in flight per step
    constexpr int CO = CPO / 32;    // output columns per lane
    constexpr int KU = 2 * CPI;     // length of [H_i | (L.H)_i]

    extern __shared__ __align__(16) float smem[];
    float *w_sm = smem;                       // [KU, CPO]
    float *b_sm = w_sm + KU * CPO;            // [CPO]
    float *u_all = b_sm + CPO;                // [warps][R][KU]

    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    float *u_sm = u_all + warp * (R * KU);

    for (int k = threadIdx.x * 4; k < KU * CPO; k += blockDim.x * 4)
        *reinterpret_cast<float4 *>(w_sm + k) = __ldg(reinterpret_cast<const float4 *>(a.wcat + k));
    for (int k = threadIdx.x; k < CPO; k += blockDim.x) b_sm[k] = a.bias[k];
    __syncthreads();

    const int g = lane / LPR;
    const int q = lane % LPR;

    float ia0[4], ia1[4], ib[4];
    if (IMPLICIT_IN) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            ia0[e] = a.in_a0[4 * q + e];
            ia1[e] = a.in_a1[4 * q + e];
            ib[e] = a.in_b[4 * q + e];
        }
    }
    auto implicit_row = [&](float px, float py) {
        float4 v;
        v.x = act_apply(fmaf(py, ia1[0], fmaf(px, ia0[0], ib[0])), a.in_act, a.alpha);
        v.y = act_apply(fmaf(py, ia1[1], fmaf(px, ia0[1], ib[1])), a.in_act, a.alpha);
        v.z = act_apply(fmaf(py, ia1[2], fmaf(px, ia0[2], ib[2])), a.in_act, a.alpha);
        v.w = act_apply(fmaf(py, ia1[3], fmaf(px, ia0[3], ib[3])), a.in_act, a.alpha);
        return v;
    };

    float tw0[CO], tw1[CO];
    if (TAIL) {
#pragma unroll
        for (int cc = 0; cc < CO; ++cc) {
            tw0[cc] = a.tail_w0[lane + 32 * cc];
            tw1[cc] = a.tail_w1[lane + 32 * cc];
        }
    }

    const int n_chunks = (a.n + R - 1) / R;
    for (int chunk = blockIdx.x * kWarpsPerCta + warp; chunk < n_chunks; chunk += gridDim.x * kWarpsPerCta) {
        const int row0 = chunk * R;
        // ---- aggregation: u_r = [H_i | H_i - dinv_i * sum_j dinv_j H_j] --------------------------
        // One lane group (LPR lanes) per row, NG rows of the warp in flight.  Per step the group's lanes fetch
        // LPR consecutive column ids (one coalesced load) and the matching dinv, hand them round with
        // width-LPR shuffles and issue the LPR neighbour-row loads back to back: LPR independent 16-byte
        // loads per lane, which is what a gather from L2 / HBM needs to stay busy.  Lanes past the end of a row
        // re-read the row itself with weight 0, so the loop body carries no branch.
#pragma unroll 1
        for (int pass = 0; pass < R / NG; ++pass) {
            const int r = pass * NG + g;
            const int i = row0 + r;
            const bool valid = i < a.n;
            const int isafe = valid ? i : a.n - 1;
            int beg = 0, end = 0;
            if (valid) {
                beg = a.row_ptr[i];
                end = a.row_ptr[i + 1];
            }
            const int self = a.row0 + isafe;
            // NG partial sums per row, neighbour k of the row going to partial k mod NG, combined pairwise at the
            // end: the summation tree that keeps the 20-layer checkpoints inside the 1e-5 budget (a plain
            // two-way split is 1.02e-5 on one of them)
            float4 acc[NG];
#pragma unroll
            for (int k = 0; k < NG; ++k) acc[k] = make_float4(0.f, 0.f, 0.f, 0.f);
            for (int e = beg; __any_sync(0xffffffffu, e < end); e += LPR) {
                const bool in = e + q < end;
                const int c = in ? __ldg(a.col_idx + e + q) : self;
                const float d = in ? __ldg(a.dinv + c) : 0.f;
                float2 p = make_float2(0.f, 0.f);
                if (IMPLICIT_IN) p = __ldg(a.pair_in + c);
                if (IMPLICIT_IN) {
#pragma unroll
                    for (int t = 0; t < LPR; ++t) {
                        const float dj = __shfl_sync(0xffffffffu, d, t, LPR);
                        const float px = __shfl_sync(0xffffffffu, p.x, t, LPR);
                        const float py = __shfl_sync(0xffffffffu, p.y, t, LPR);
                        const float4 v = implicit_row(px, py);
                        acc[t % NG].x = fmaf(dj, v.x, acc[t % NG].x);
                        acc[t % NG].y = fmaf(dj, v.y, acc[t % NG].y);
                        acc[t % NG].z = fmaf(dj, v.z, acc[t % NG].z);
                        acc[t % NG].w = fmaf(dj, v.w, acc[t % NG].w);
                    }
                } else {
                    float4 v[LPR];
                    float dj[LPR];
#pragma unroll
                    for (int t = 0; t < LPR; ++t) {
                        const int jj = __shfl_sync(0xffffffffu, c, t, LPR);
                        dj[t] = __shfl_sync(0xffffffffu, d, t, LPR);
                        v[t] = __ldg(reinterpret_cast<const float4 *>(a.hin + (size_t)jj * CPI) + q);
                    }
#pragma unroll
                    for (int t = 0; t < LPR; ++t) {
                        acc[t % NG].x = fmaf(dj[t], v[t].x, acc[t % NG].x);
                        acc[t % NG].y = fmaf(dj[t], v[t].y, acc[t % NG].y);
                        acc[t % NG].z = fmaf(dj[t], v[t].z, acc[t % NG].z);
                        acc[t % NG].w = fmaf(dj[t], v[t].w, acc[t % NG].w);
                    }
                }
            }
            float4 hi = make_float4(0.f, 0.f, 0.f, 0.f);
            float di = 0.f;
            if (valid) {
                di = a.dinv[self];
                if (IMPLICIT_IN) {
                    const float2 p = a.pair_in[self];
                    hi = implicit_row(p.x, p.y);
                } else {
                    hi = __ldg(reinterpret_cast<const float4 *>(a.hin + (size_t)self * CPI) + q);
                }
            }
#pragma unroll
            for (int off = 1; off < NG; off <<= 1) {
#pragma unroll
                for (int k = 0; k + off < NG; k += 2 * off) {
                    acc[k].x += acc[k + off].x;
                    acc[k].y += acc[k + off].y;
                    acc[k].z += acc[k + off].z;
                    acc[k].w += acc[k + off].w;
                }
            }
            const float4 lh = make_float4(fmaf(-di, acc[0].x, hi.x), fmaf(-di, acc[0].y, hi.y),
                                          fmaf(-di, acc[0].z, hi.z), fmaf(-di, acc[0].w, hi.w));
            *reinterpret_cast<float4 *>(u_sm + r * KU + 4 * q) = hi;
            *reinterpret_cast<float4 *>(u_sm + r * KU + CPI + 4 * q) = lh;
        }
        __syncwarp();
        // ---- projection: out[r][c] = sum_k u[r][k] * Wcat[k][c] ----------------------------------
        float out[R][CO];
#pragma unroll
        for (int r = 0; r < R; ++r)
#pragma unroll
            for (int cc = 0; cc < CO; ++cc) out[r][cc] = b_sm[lane + 32 * cc];
#pragma unroll 2
        for (int k = 0; k < KU; k += 4) {
            float w[4][CO];
#pragma unroll
            for (int e = 0; e < 4; ++e)
#pragma unroll
                for (int cc = 0; cc < CO; ++cc) w[e][cc] = w_sm[(k + e) * CPO + lane + 32 * cc];
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const float4 u = *reinterpret_cast<const float4 *>(u_sm + r * KU + k);
#pragma unroll
                for (int cc = 0; cc < CO; ++cc) {
                    out[r][cc] = fmaf(u.x, w[0][cc], out[r][cc]);
                    out[r][cc] = fmaf(u.y, w[1][cc], out[r][cc]);
                    out[r][cc] = fmaf(u.z, w[2][cc], out[r][cc]);
                    out[r][cc] = fmaf(u.w, w[3][cc], out[r][cc]);
                }
            }
        }
        __syncwarp();  // u_sm is rewritten by the next chunk
        // ---- epilogue ----------------------------------------------------------------------------
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const int i = row0 + r;
            if (!TAIL) {
                if (i < a.n) {
#pragma unroll
                    for (int cc = 0; cc < CO; ++cc)
                        peer_store(a.pm, a.hout + (size_t)(a.row0 + i) * CPO + lane + 32 * cc,
                                   act_apply(out[r][cc], a.act, a.alpha));
                }
            } else {
                float t0 = 0.f, t1 = 0.f;
#pragma unroll
                for (int cc = 0; cc < CO; ++cc) {
                    const float h = act_apply(out[r][cc], a.act, a.alpha);
                    t0 = fmaf(h, tw0[cc], t0);
                    t1 = fmaf(h, tw1[cc], t1);
                }
#pragma unroll
                for (int off = 16; off > 0; off >>= 1) {
                    t0 += __shfl_xor_sync(0xffffffffu, t0, off);
                    t1 += __shfl_xor_sync(0xffffffffu, t1, off);
                }
                if (lane == 0 && i < a.n) {
                    a.tail_q[a.row0 + i] = t0 + t1;
                    peer_store(a.pm, a.tail_zs + a.row0 + i, a.dinv[a.row0 + i] * t1);
                }
            }
        }
    }
}

template <int CPI, int CPO>
constexpr size_t layer_smem_bytes() {
    return sizeof(float) * (size_t)(2 * CPI * CPO + CPO + kWarpsPerCta * kRowsPerWarp * 2 * CPI);
}

template <int CPI, int CPO, bool IMPLICIT_IN, bool TAIL>
int launch_layer_t(dg_context *ctx, const LayerArgs &args) {
    auto kern = gc_layer_kernel<CPI, CPO, IMPLICIT_IN, TAIL>;
    constexpr size_t smem = layer_smem_bytes<CPI, CPO>();
    static std::atomic<unsigned long long> attr_done{0};   // (the attribute: once per device)
    DG_CUDA_CHECK(smem_attr_once(kern, ctx->device, (int)smem, &attr_done));
    static int blocks_per_sm = 0;  // per instantiation; every context uses the same device kind
    if (blocks_per_sm == 0) {
        int nb = 0;
        DG_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, kWarpsPerCta * 32, smem));
        DG_REQUIRE(nb > 0, DG_ERR_CUDA, "gc_layer_kernel<%d,%d> does not fit on an SM", CPI, CPO);
        blocks_per_sm = nb;
    }
    const int n_chunks = (args.n + kRowsPerWarp - 1) / kRowsPerWarp;
    const int need = (n_chunks + kWarpsPerCta - 1) / kWarpsPerCta;
    int grid = ctx->sm_count * blocks_per_sm;
    if (grid > need) grid = need;
    if (grid < 1) grid = 1;
    // B_layer: CSR pattern + dinv + one read of the input rows + one write of the output
    // rows (two floats per vertex for the implicit / tail forms) + the weight matrix
    const double n = (double)args.n;
    const double nnz = (double)args.nnz;
    const double bytes = 4.0 * (n + 1) + 4.0 * nnz + 4.0 * n + 4.0 * n * (IMPLICIT_IN ? 2 : CPI) +
                         4.0 * n * (TAIL ? 2 : CPO) + 4.0 * (2 * CPI * CPO + CPO);
    ctx->last_kernel = "gc_layer_kernel";
        prof_begin(ctx);
    kern<<<grid, kWarpsPerCta * 32, smem, ctx->stream>>>(args);
    prof_end(ctx, bytes);
    ctx->launches++;
    DG_CUDA_CHECK(cudaGetLastError());
    return DG_OK;
}

int launch_layer(dg_context *ctx, int cpi, int cpo, bool implicit_in, bool tail, const LayerArgs &args) {
#define DG_DISPATCH(CI, CO_)                                                          \
    if (cpi == CI && cpo == CO_) {                                                    \
        if (implicit_in && tail) return launch_layer_t<CI, CO_, true, true>(ctx, args);   \
        if (implicit_in) return launch_layer_t<CI, CO_, true, false>(ctx, args);          \
        if (tail) return launch_layer_t<CI, CO_, false, true>(ctx, args);                 \
        return launch_layer_t<CI, CO_, false, false>(ctx, args);                          \
    }
    DG_DISPATCH(32, 32)
    DG_DISPATCH(32, 64)
    DG_DISPATCH(64, 32)
    DG_DISPATCH(64, 64)
#undef DG_DISPATCH
    set_error("unsupported padded layer shape %d -> %d", cpi, cpo);
    return DG_ERR_UNSUPPORTED;
}

// copy [n, c] (leading dimension ld_src) into a zero-padded [n, cp] buffer and back
__global__ void pad_rows_kernel(int n, int c, int cp, const float *__restrict__ src, int ld_src,
                                float *__restrict__ dst) {
    size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (size_t)n * cp) return;
    int i = (int)(idx / cp), k = (int)(idx - (size_t)i * cp);
    dst[idx] = k < c ? src[(size_t)i * ld_src + k] : 0.f;
}

// out[i, 0..d) = head(hpad[i, 0..d)), zero for removed vertices
__global__ void finalize_kernel(int n, int d_out, int cp, const float *__restrict__ hpad,
                                const uint8_t *__restrict__ keep, int head, float *__restrict__ out,
                                int ld_out) {
    int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n * d_out) return;
    int i = idx / d_out, d = idx - i * d_out;
    float v = hpad[(size_t)i * cp + d];
    if (head == DG_HEAD_PAIR_SOFTMAX) {
        // softmax over the (2p, 2p+1) pair that column d belongs to (gcn/models.py:399-401)
        int mate = d ^ 1;
        float o = mate < d_out ? hpad[(size_t)i * cp + mate] : v;
        float m = fmaxf(v, o);
        float ev = expf(v - m), eo = expf(o - m);
        v = ev / (ev + eo);
    }
    out[(size_t)i * ld_out + d] = (keep && !keep[i]) ? 0.f : v;
}

inline int grid_for(size_t n, int block) { return (int)((n + block - 1) / block); }

// (q, zs) of a one-column last layer from dense rows: q_i = H_i.w_0 + z_i, zs_i = dinv_i z_i, z = H.w_1
__global__ void __launch_bounds__(256)
tail_project_kernel(int n, int row0, int cp, const float *__restrict__ hin, const float *__restrict__ w0,
                    const float *__restrict__ w1, const float *__restrict__ dinv, float *__restrict__ q_out,
                    float *__restrict__ zs_out, const PeerMap pm) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 *row = reinterpret_cast<const float4 *>(hin + (size_t)(row0 + i) * cp);
    float t0 = 0.f, t1 = 0.f;
    for (int c = 0; c < cp / 4; ++c) {
        const float4 h = __ldg(row + c);
        const float4 a = __ldg(reinterpret_cast<const float4 *>(w0) + c);
        const float4 b = __ldg(reinterpret_cast<const float4 *>(w1) + c);
        t0 = fmaf(h.x, a.x, t0), t0 = fmaf(h.y, a.y, t0), t0 = fmaf(h.z, a.z, t0), t0 = fmaf(h.w, a.w, t0);
        t1 = fmaf(h.x, b.x, t1), t1 = fmaf(h.y, b.y, t1), t1 = fmaf(h.z, b.z, t1), t1 = fmaf(h.w, b.w, t1);
    }
    q_out[row0 + i] = t0 + t1;
    peer_store(pm, zs_out + row0 + i, dinv[row0 + i] * t1);
}

}  // namespace

// =================================================================================================
// row-slice drivers: one giant graph partitioned by rows over several GPUs.  Arrays are
// global sized and indexed by global vertex id; each call writes rows row0 .. row0+n_local-1 only and
// the caller all-gathers what the next call reads from other ranks' rows.
// =================================================================================================
int part_keep(dg_context *ctx, const PartView &pv, const double *wts, int remove_zero_weight, int n_real,
              uint8_t *keep) {
    if (pv.n_local == 0) return DG_OK;
    part_keep_kernel<<<grid_for(pv.n_local, 256), 256, 0, ctx->stream>>>(pv.n_local, pv.row0, n_real, wts,
                                                                        remove_zero_weight, keep, pv.pm);
    ctx->launches++;
    DG_CUDA_CHECK(cudaGetLastError());
    return DG_OK;
}

int part_prepare(dg_context *ctx, const PartView &pv, const uint8_t *keep, const float *x0, float x0val, float *dinv,
                 float *y) {
    if (pv.n_local == 0) return DG_OK;
    if (keep)
        degree_keep_kernel<<<grid_for(pv.n_local, kStreamRows), kStreamThreads, 0, ctx->stream>>>(
            pv.n_local, pv.row0, pv.row_ptr, pv.col_idx, keep, dinv, pv.pm);
    else
        degree_kernel<<<grid_for(pv.n_local, 256), 256, 0, ctx->stream>>>(pv.n_local, pv.row0, pv.row_ptr, dinv, pv.pm);
    ctx->launches++;
    DG_CUDA_CHECK(cudaGetLastError());
    return part_scale(ctx, pv, keep, x0, x0val, dinv, y);
}

int part_scale(dg_context *ctx, const PartView &pv, const uint8_t *keep, const float *x0, float x0val,
               const float *dinv, float *y) {
    if (pv.n_local == 0) return DG_OK;
    scaled_input_kernel<<<grid_for(pv.n_local, 256), 256, 0, ctx->stream>>>(
        pv.n_local, dinv + pv.row0, keep ? keep + pv.row0 : nullptr, x0 ? x0 + pv.row0 : nullptr, x0val, y + pv.row0,
        pv.pm);
    ctx->launches++;
    DG_CUDA_CHECK(cudaGetLastError());
    return DG_OK;
}

int part_first(dg_context *ctx, const PartView &pv, const float *dinv, const float *y, const uint8_t *keep,
               const float *x0, float x0val, float2 *pair) {
    if (pv.n_local == 0) return DG_OK;
    first_scalar_kernel<<<grid_for(pv.n_local, kStreamRows), kStreamThreads, 0, ctx->stream>>>(
        pv.n_local, pv.row_ptr, pv.col_idx, dinv + pv.row0, y, keep ? keep + pv.row0 : nullptr,
        x0 ? x0 + pv.row0 : nullptr, x0val, pair + pv.row0, pv.pm);
    ctx->launches++;
    DG_CUDA_CHECK(cudaGetLastError());
    return DG_OK;
}

int part_project(dg_context *ctx, const PartView &pv, const dg_model *m, const float *dinv, const float2 *pair,
                 float *pair2) {
    if (pv.n_local == 0) return DG_OK;
    const dg_layer_dev &first = m->layers[0];
    node_project_kernel<<<grid_for(pv.n_local, 256), 256, 0, ctx->stream>>>(
        pv.n_local, first.c_out, pair + pv.row0, first.colsum0, first.colsum1, first.bias, first.act, m->alpha,
        m->tail_w0, m->tail_w1, dinv + pv.row0, pair2 + pv.row0, pair2 + pv.n_global + pv.row0, pv.pm);
    ctx->launches++;
    DG_CUDA_CHECK(cudaGetLastError());
    return DG_OK;
}

int part_layer(dg_context *ctx, const PartView &pv, const dg_model *m, int layer, const float *dinv,
               const float2 *pair, const float *hin, float *hout) {
    if (pv.n_local == 0) return DG_OK;
    const dg_layer_dev &first = m->layers[0];
    const dg_layer_dev &ly = m->layers[layer];
    LayerArgs a{};
    a.n = pv.n_local;
    a.nnz = pv.nnz;
    a.row0 = pv.row0;
    a.row_ptr = pv.row_ptr;
    a.col_idx = pv.col_idx;
    a.dinv = dinv;
    const bool implicit_in = layer == 1;
    if (implicit_in) {
        a.pair_in = pair;
        a.in_a0 = first.colsum0;
        a.in_a1 = first.colsum1;
        a.in_b = first.bias;
        a.in_act = first.act;
    } else {
        a.hin = hin;
    }
    a.wcat = ly.wcat;
    a.bias = ly.bias;
    a.act = ly.act;
    a.alpha = m->alpha;
    a.hout = hout;
    a.pm = pv.pm;
    return launch_layer(ctx, ly.cpi, ly.cpo, implicit_in, false, a);
}

int part_tail(dg_context *ctx, const PartView &pv, const dg_model *m, const float *dinv, const float *hin,
              float *pair2) {
    if (pv.n_local == 0) return DG_OK;
    const int cp = m->layers[m->n_layers - 1].cpi;
    tail_project_kernel<<<grid_for(pv.n_local, 256), 256, 0, ctx->stream>>>(pv.n_local, pv.row0, cp, hin, m->tail_w0,
                                                                           m->tail_w1, dinv, pair2, pair2 + pv.n_global, pv.pm);
    ctx->launches++;
    DG_CUDA_CHECK(cudaGetLastError());
    return DG_OK;
}

int part_last(dg_context *ctx, const PartView &pv, const dg_model *m, const float *dinv, const float *pair2,
              const uint8_t *keep, const double *wts, int predict, float *score, double *util) {
    if (pv.n_local == 0) return DG_OK;
    const dg_layer_dev &last = m->layers[m->n_layers - 1];
    last_scalar_kernel<<<grid_for(pv.n_local, kStreamRows), kStreamThreads, 0, ctx->stream>>>(
        pv.n_local, pv.row0, pv.row_ptr, pv.col_idx, dinv, pair2, pair2 + pv.n_global, m->tail_bias, last.act, m->alpha, keep, score, wts,
        predict, util, pv.pm);
    ctx->launches++;
    DG_CUDA_CHECK(cudaGetLastError());
    return DG_OK;
}

// ---- scalar networks with any number of supports (the cheb2 checkpoints) ----------------------------------------
// first-layer input: z_i = x0_i on kept vertices (the row-normalised constant features, gcn/utils.py:98-106), else 0
__global__ void scalar_input_kernel(int n, const uint8_t *__restrict__ keep, const float *__restrict__ x0, float x0val,
                                    float *__restrict__ z) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const bool k = keep ? keep[i] != 0 : true;
    z[i] = k ? (x0 ? x0[i] : x0val) : 0.f;
}
__global__ void scalar_scale_kernel(int n, const float *__restrict__ z, float w, float *__restrict__ t) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) t[i] = z[i] * w;   // pre_sup = dot(x, W_k), gcn/layers.py:202
}
// t_out = L t_in, (L t)_i = t_i - dinv_i * sum_j dinv_j t_j; one thread per row, neighbours in ascending order
__global__ void scalar_lap_kernel(int n, const int *__restrict__ row_ptr, const int *__restrict__ col_idx,
                                  const float *__restrict__ dinv, const float *__restrict__ t_in, float *__restrict__ t_out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float acc = 0.f;
    for (int e = row_ptr[i]; e < row_ptr[i + 1]; ++e) {
        const int j = __ldg(col_idx + e);
        acc = fmaf(__ldg(dinv + j), __ldg(t_in + j), acc);
    }
    t_out[i] = fmaf(-dinv[i], acc, t_in[i]);
}
__global__ void scalar_add_kernel(int n, const float *__restrict__ t, float *__restrict__ acc, int first) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) acc[i] = first ? t[i] : acc[i] + t[i];   // tf.add_n over the supports, in order (gcn/layers.py:208)
}
__global__ void scalar_act_kernel(int n, const float *__restrict__ acc, float bias, int act, float alpha,
                                  const uint8_t *__restrict__ keep, int zero_removed, float *__restrict__ z) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float v = act_apply(acc[i] + bias, act, alpha);
    if (zero_removed && keep && !keep[i]) v = 0.f;
    z[i] = v;
}

// Y = L.Z, Z [n, width] row-major: the generic form (one warp per row, lanes over the feature columns, neighbours in
// ascending order); batches of small graphs with 32-wide rows take the graph-staged kernel of dg_stream.cu instead
__global__ void spmm_laplacian_kernel(int n, int width, const int *__restrict__ row_ptr, const int *__restrict__ col_idx,
                                      const float *__restrict__ dinv, const float *__restrict__ z, float *__restrict__ y) {
    const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (row >= n) return;
    const float di = dinv[row];
    for (int c = lane; c < width; c += 32) {
        float acc = 0.f;
        for (int e = row_ptr[row]; e < row_ptr[row + 1]; ++e) {
            const int j = __ldg(col_idx + e);
            acc = fmaf(__ldg(dinv + j), __ldg(z + (size_t)j * width + c), acc);
        }
        y[(size_t)row * width + c] = fmaf(-di, acc, z[(size_t)row * width + c]);
    }
}

// =================================================================================================
// drivers
// =================================================================================================
int batch_compute_dinv(dg_batch *b) {
    dg_context *ctx = b->ctx;
    if (b->n_nodes == 0) return DG_OK;
    if (b->keep)
        degree_keep_kernel<<<grid_for(b->n_nodes, kStreamRows), kStreamThreads, 0, ctx->stream>>>(
            b->n_nodes, 0, b->row_ptr, b->col_idx, b->keep, b->dinv, PeerMap{});
    else
        degree_kernel<<<grid_for(b->n_nodes, 256), 256, 0, ctx->stream>>>(b->n_nodes, 0, b->row_ptr, b->dinv, PeerMap{});
    ctx->launches++;
    DG_CUDA_CHECK(cudaGetLastError());
    return DG_OK;
}

int keep_from_weights_device(dg_context *ctx, int n, const double *wts, uint8_t *keep) {
    if (n == 0) return DG_OK;
    keep_from_weights_kernel<<<grid_for(n, 256), 256, 0, ctx->stream>>>(n, wts, keep);
    ctx->launches++;
    DG_CUDA_CHECK(cudaGetLastError());
    return DG_OK;
}

int utility_device(dg_context *ctx, int n, const float *score, int stride, const double *wts, int predict,
                   double *util) {
    if (n == 0) return DG_OK;
    utility_kernel<<<grid_for(n, 256), 256, 0, ctx->stream>>>(n, score, stride, wts, predict, util);
    ctx->launches++;
    DG_CUDA_CHECK(cudaGetLastError());
    return DG_OK;
}

int spmm_laplacian_device(dg_context *ctx, dg_batch *b, int width, const float *z, float *y) {
    const int n = b->n_nodes;
    if (n == 0) return DG_OK;
    bool staged = false;
    DG_TRY(gs_try_spmm(ctx, b, width, z, y, &staged));
    if (staged) return DG_OK;
    const double bytes = 4.0 * (n + 1.0) + 4.0 * b->nnz + 4.0 * n + 8.0 * (double)n * width;
    ctx->last_kernel = "spmm_laplacian_kernel";
    prof_begin(ctx);
    spmm_laplacian_kernel<<<grid_for((size_t)n * 32, 256), 256, 0, ctx->stream>>>(n, width, b->row_ptr, b->col_idx, b->dinv, z, y);
    prof_end(ctx, bytes);
    ctx->launches++;
    DG_CUDA_CHECK(cudaGetLastError());
    return DG_OK;
}

int graph_convolution_device(dg_context *ctx, dg_batch *b, const dg_layer_dev &L, float alpha, const float *x,
                             int ldx, float *y, int ldy) {
    const int n = b->n_nodes;
    if (n == 0) return DG_OK;
    float *fa = nullptr, *fb = nullptr;
    DG_TRY(scratch_as(ctx, kSlotFeatA, (size_t)n * kMaxWidth, &fa));
    DG_TRY(scratch_as(ctx, kSlotFeatB, (size_t)n * kMaxWidth, &fb));
    pad_rows_kernel<<<grid_for((size_t)n * L.cpi, 256), 256, 0, ctx->stream>>>(n, L.c_in, L.cpi, x, ldx, fa);
    ctx->launches++;
    DG_CUDA_CHECK(cudaGetLastError());
    LayerArgs a{};
    a.n = n;
    a.nnz = b->nnz;
    a.row_ptr = b->row_ptr;
    a.col_idx = b->col_idx;
    a.dinv = b->dinv;
    a.hin = fa;
    a.wcat = L.wcat;
    a.bias = L.bias;
    a.act = L.act;
    a.alpha = alpha;
    a.hout = fb;
    DG_TRY(launch_layer(ctx, L.cpi, L.cpo, false, false, a));
    finalize_kernel<<<grid_for((size_t)n * L.c_out, 256), 256, 0, ctx->stream>>>(n, L.c_out, L.cpo, fb, nullptr,
                                                                               DG_HEAD_LINEAR, y, ldy);
    ctx->launches++;
    DG_CUDA_CHECK(cudaGetLastError());
    return DG_OK;
}

int gcn_forward_device(dg_context *ctx, const dg_model *m, dg_batch *b, float *out, const double *wts,
                       int predict, double *util, bool try_resident) {
    const int n = b->n_nodes;
    if (n == 0) return DG_OK;
    if (try_resident) {   // small graphs with a one-column linear head: the graph-resident kernel, stopped after the scores
        bool handled = false;
        if (wts != nullptr || predict == DG_PREDICT_MIS) {
            DG_TRY(tc_try_solve(ctx, m, b, wts, predict, 0, nullptr, out, util, nullptr, nullptr, &handled));
            if (!handled)
                DG_TRY(fused_try_solve(ctx, m, b, wts, predict, 0, nullptr, out, util, nullptr, nullptr, &handled));
        }
        if (handled) return DG_OK;
    }
    if (m->scalar_net) {
        // every layer: acc = sum_k L^k (z * w_k) + b, z' = act(acc) - K (K - 1) / 2 scalar passes over the graph per layer
        cudaStream_t st = ctx->stream;
        const int K = m->n_supports, grid = grid_for(n, 256);
        float *z = nullptr, *acc = nullptr, *ta = nullptr, *tb = nullptr;
        DG_TRY(scratch_as(ctx, kSlotY, (size_t)n, &z));
        DG_TRY(scratch_as(ctx, kSlotFeatA, (size_t)n * kMaxWidth, &acc));
        ta = acc + n;
        tb = ta + n;
        scalar_input_kernel<<<grid, 256, 0, st>>>(n, b->keep, b->x0, 1.0f / (float)m->layers[0].c_in, z);
        ctx->launches++;
        for (int l = 0; l < m->n_layers; ++l) {
            for (int k = 0; k < K; ++k) {
                float *t = ta, *u = tb;
                scalar_scale_kernel<<<grid, 256, 0, st>>>(n, z, m->scalar_w[(size_t)l * K + k], t);
                for (int p = 0; p < k; ++p) {   // L^k as k applications of L (the reference multiplies by a materialised L^k)
                    scalar_lap_kernel<<<grid, 256, 0, st>>>(n, b->row_ptr, b->col_idx, b->dinv, t, u);
                    std::swap(t, u);
                    ctx->launches++;
                }
                scalar_add_kernel<<<grid, 256, 0, st>>>(n, t, acc, k == 0 ? 1 : 0);
                ctx->launches += 2;
            }
            const bool last_layer = l + 1 == m->n_layers;
            scalar_act_kernel<<<grid, 256, 0, st>>>(n, acc, m->scalar_b[(size_t)l], m->layers[l].act, m->alpha, b->keep,
                                                    last_layer ? 1 : 0, last_layer ? out : z);
            ctx->launches++;
        }
        DG_CUDA_CHECK(cudaGetLastError());
        ctx->last_kernel = "scalar_lap_kernel";
        if (util) DG_TRY(utility_device(ctx, n, out, 1, wts, predict, util));
        return DG_OK;
    }
    const int L = m->n_layers;
    const dg_layer_dev &first = m->layers[0];
    const dg_layer_dev &last = m->layers[L - 1];
    const int d_out = last.c_out;
    cudaStream_t st = ctx->stream;

    float *y = nullptr;
    float2 *pair = nullptr;
    float *tail_q = nullptr, *tail_zs = nullptr;  // (q, zs) planes of the one-column last layer
    DG_TRY(scratch_as(ctx, kSlotY, (size_t)n, &y));
    DG_TRY(scratch_as(ctx, kSlotPair, (size_t)n, &pair));
    const float x0val = 1.0f / (float)first.c_in;  // gcn/utils.py:98-106 on constant rows
    scaled_input_kernel<<<grid_for(n, 256), 256, 0, st>>>(n, b->dinv, b->keep, b->x0, x0val, y, PeerMap{});
    ctx->launches++;
    const int scalar_grid = grid_for(n, kStreamRows);
    first_scalar_kernel<<<scalar_grid, kStreamThreads, 0, st>>>(n, b->row_ptr, b->col_idx, b->dinv, y, b->keep, b->x0,
                                                                x0val, pair, PeerMap{});
    ctx->launches++;
    DG_CUDA_CHECK(cudaGetLastError());

    if (L == 1) {
        float *dst = out;
        if (m->head == DG_HEAD_PAIR_SOFTMAX) DG_TRY(scratch_as(ctx, kSlotFeatA, (size_t)n * kMaxWidth, &dst));
        first_out_kernel<<<grid_for((size_t)n * d_out, 256), 256, 0, st>>>(
            n, d_out, pair, first.colsum0, first.colsum1, first.bias, first.act, m->alpha, b->keep, dst);
        ctx->launches++;
        if (m->head == DG_HEAD_PAIR_SOFTMAX) {
            finalize_kernel<<<grid_for((size_t)n * d_out, 256), 256, 0, st>>>(n, d_out, d_out, dst, b->keep,
                                                                             m->head, out, d_out);
            ctx->launches++;
        }
        DG_CUDA_CHECK(cudaGetLastError());
        if (util) DG_TRY(utility_device(ctx, n, out, d_out, wts, predict, util));
        return DG_OK;
    }

    const bool scalar_tail = (d_out == 1 && m->head == DG_HEAD_LINEAR);
    float *fa = nullptr, *fb = nullptr;
    DG_TRY(scratch_as(ctx, kSlotFeatA, (size_t)n * kMaxWidth, &fa));
    DG_TRY(scratch_as(ctx, kSlotFeatB, (size_t)n * kMaxWidth, &fb));
    if (scalar_tail) {
        DG_TRY(scratch_as(ctx, kSlotPair2, (size_t)2 * n, &tail_q));
        tail_zs = tail_q + n;
    }

    const int last_fused = scalar_tail ? L - 2 : L - 1;  // last layer run through gc_layer_kernel
    const float *cur = nullptr;
    float *nxt = fa;
    for (int l = 1; l <= last_fused; ++l) {
        const dg_layer_dev &ly = m->layers[l];
        LayerArgs a{};
        a.n = n;
        a.nnz = b->nnz;
        a.row_ptr = b->row_ptr;
        a.col_idx = b->col_idx;
        a.dinv = b->dinv;
        const bool implicit_in = (l == 1);
        if (implicit_in) {
            a.pair_in = pair;
            a.in_a0 = first.colsum0;
            a.in_a1 = first.colsum1;
            a.in_b = first.bias;
            a.in_act = first.act;
        } else {
            a.hin = cur;
        }
        a.wcat = ly.wcat;
        a.bias = ly.bias;
        a.act = ly.act;
        a.alpha = m->alpha;
        const bool tail = scalar_tail && l == last_fused;
        if (tail) {
            a.tail_w0 = m->tail_w0;
            a.tail_w1 = m->tail_w1;
            a.tail_q = tail_q;
            a.tail_zs = tail_zs;
        } else {
            a.hout = nxt;
        }
        bool staged = false;  // batches of small graphs: the graph-staged kernel (dg_stream.cu), else warp per row
        DG_TRY(gs_try_layer(ctx, b, ly.cpi, ly.cpo, implicit_in, tail, a, &staged));
        if (!staged) DG_TRY(launch_layer(ctx, ly.cpi, ly.cpo, implicit_in, tail, a));
        cur = nxt;
        nxt = (nxt == fa) ? fb : fa;
    }

    if (scalar_tail) {
        if (L == 2) {
            node_project_kernel<<<grid_for(n, 256), 256, 0, st>>>(n, first.c_out, pair, first.colsum0,
                                                                  first.colsum1, first.bias, first.act, m->alpha,
                                                                  m->tail_w0, m->tail_w1, b->dinv, tail_q, tail_zs,
                                                                  PeerMap{});
            ctx->launches++;
        }
        last_scalar_kernel<<<scalar_grid, kStreamThreads, 0, st>>>(n, 0, b->row_ptr, b->col_idx, b->dinv, tail_q, tail_zs,
                                                                   m->tail_bias, last.act, m->alpha, b->keep, out, wts,
                                                                   predict, util, PeerMap{});
        ctx->launches++;
        DG_CUDA_CHECK(cudaGetLastError());
        return DG_OK;
    }

    finalize_kernel<<<grid_for((size_t)n * d_out, 256), 256, 0, st>>>(n, d_out, last.cpo, cur, b->keep, m->head,
                                                                     out, d_out);
    ctx->launches++;
    DG_CUDA_CHECK(cudaGetLastError());
    if (util) DG_TRY(utility_device(ctx, n, out, d_out, wts, predict, util));
    return DG_OK;
}

}  // namespace dg
