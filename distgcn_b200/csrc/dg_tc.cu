// Tensor-core solve kernel (tcgen05 / TMEM, sm_100a): the whole hot path for a tile of small graphs in one launch, like
// dg_fused.cu, but with both heavy steps of every hidden GraphConvolution layer on the 5th-generation tensor cores.
//
//   reference order (gcn/layers.py:198-216):  H' = act( H.W_0 + L.(H.W_1) + b ),  L = I - D^-1/2 A D^-1/2
//   here, per hidden layer and per 128-vertex block of a graph:
//     (1) projection   [P0 | P1] = H . [W_0 | W_1 r]        tcgen05.mma kind::f16 (bf16 x bf16 -> fp32 in TMEM), M128 N64 K32.
//         fp32 accuracy comes from a 3-term bf16 split of BOTH operands (x = hi + mid + lo, each term exact) and the 6
//         significant cross products chained into one accumulator, smallest first (measured: 2.1e-7 of the output scale
//         against 4.5e-7 for an fp32 FMA chain, profiles/micro/umma_probe.cu).
//     (2) aggregation  S = A . (dinv * P1)                    tcgen05.mma kind::i8 (u8 x u8 -> s32 in TMEM), M128 N128 K=n.
//         A is the graph's dense 0/1 adjacency, resident in shared memory as bytes for all layers; Y = dinv * P1 is
//         quantised per graph and layer to 31-bit fixed point (scale from a bound on max |Y|, below) and its four
//         base-256 digits are the operand columns, so the sum over neighbours is EXACT integer arithmetic.
//         Digits are balanced (signed bytes): sum_j A_ij v_j = sum_a 256^a D_a with |D_a| <= 128 deg_i.
//     (3) epilogue     H' = act(P0 + P1 + b - dinv_i * S_i)   CUDA cores, one thread per vertex (TMEM lane), result split
//         into bf16 terms and stored as the next projection's A operand.
//   Fixed-point scale: |Y_jf| <= max|H| * max_f sum_k |W_1[k,f] r_f|, with r_f powers of two (exact) that equalise the
//   column norms of W_1 (host, undone in the epilogue) and max|H| reduced per graph and layer in shared memory.  The
//   quantisation error of an aggregated element is <= deg * 2^-31 of that bound, below fp32 rounding of the same sum.
//   The scalar aggregations of the rank-1 first layer and of the project-first last layer use the same machinery with
//   16 operand columns.
//
// Work split: 16 warps, warpgroup w owns the TMEM lanes of block w (thread = vertex).  A graph is a synchronisation
// domain and there is no dedicated MMA warp: when a warp finishes writing an operand it bumps a shared-memory counter,
// and the warp that completes the count (the last of the block for a projection, the last of the graph for an
// aggregation) issues that step's tcgen05.mma chain from its lane 0 and commits it to the mbarrier the consumers wait
// on.  Weight blobs stream through a 2-deep TMA ring, refilled by whichever warp retires the previous user.  Blocks and
// graphs of a tile therefore drift apart and the tensor-core work of one overlaps the epilogue arithmetic of another.
//
// Reference semantics: dg_fused.cu / dg_gcn.cu / dg_lgs.cu (mwis_dqn_call.py:198-261, heuristics.py:77-116).
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <functional>
#include <type_traits>
#include <vector>

#include "dg_common.cuh"

namespace dg {

namespace {

#ifdef DG_TC_HALF  // experiment: two CTAs per SM, each with half the tensor memory / shared memory / warps
constexpr int kTcVertexThreads = 256;
constexpr int kTcThreads = 256;
constexpr int kTcMaxG = 2;
constexpr int kTcBlocks = 2;
constexpr int kTcCtasPerSm = 2;
#else
constexpr int kTcVertexThreads = 512;
constexpr int kTcThreads = 512;       // 16 warps: 128 registers per thread
constexpr int kTcMaxG = 4;            // graphs per tile (every graph owns >= 1 block)
constexpr int kTcBlocks = 4;          // 128-row blocks per tile = TMEM budget: 4 x 128 columns
constexpr int kTcCtasPerSm = 1;
#endif
constexpr int kTcTmemCols = 128 * kTcBlocks;
#ifndef DG_BEAT_UNROLL
#define DG_BEAT_UNROLL 4
#endif
constexpr int kBeatUnroll = DG_BEAT_UNROLL;
#ifndef DG_STAGE_EDGES
#define DG_STAGE_EDGES 8
#endif
constexpr int kStageEdges = DG_STAGE_EDGES;  // edges per thread and staging pass (global loads in flight)  // neighbours per pass of the greedy phase's 'who beats me' scan
// a partial last block reads up to 127 rows x 16 B past its graph's last region: the tile keeps that much of the pool free
constexpr int kTcOverread = 2048 + 128;
constexpr int kTcWBlob = 12560;       // one hidden layer: bf16 terms of [W_0 | W_1 r] (12288), bias[32], 1/r[32], bound, pad
// shared-memory map (bytes)
constexpr int kTcOffRing = 0;                      // 2 x kTcWBlob
constexpr int kTcOffUtil = 25216;                  // double[512]   (aliased by the staging copy of row_ptr)
constexpr int kTcOffDinv = kTcOffUtil + 4096;      // float[512]
constexpr int kTcOffBits = kTcOffDinv + 2048;      // uint32[4][16]: keep, remain, joined, member
constexpr int kTcOffGmax = kTcOffBits + 256;       // uint32[4][2] layer maxima, uint32[4] x0 maxima
constexpr int kTcOffBar = kTcOffGmax + 64;         // 22 mbarriers, see the kernel
constexpr int kTcOffMeta = kTcOffBar + 256;        // TcMeta[4] + tile scalars
// Pool, per graph of the tile (R = rows padded to 8, Kp = columns padded to 32):
//   adjacency  u8  [Kp/16 chunks][R rows][16 B]        K-major A operand of the aggregation  (R Kp bytes)
//   Y digits   u8  [8 runs][Kp/8 groups][8 k][16 B]     MN-major B operand of the aggregation (128 Kp bytes)
// A 128-row block of a graph whose last block is partial reads adjacency rows past R: those reads stay inside the pool (the
// next chunk / the Y region) and only feed accumulator rows nobody looks at.
// The projection's A operand (the bf16 terms of H) lives in TENSOR memory, see the column map in the kernel.
constexpr int kTcOffWatch = kTcOffMeta + 224;     // TcWatch (16 bytes)
constexpr int kTcOffPool = 32768;
// Tensor-memory columns of a block (128 of the 512; base = block * 128):
//   0..63    [P0 | P1]  fp32 accumulators of the projection                  (written by the projection, read by epilogue B)
//   0..127   S          s32 accumulators of the aggregation, column 4 f + a  (written after epilogue B, read by epilogue C)
//   64..111  H terms    bf16 A operand of the NEXT projection: term t of K step s (features 16 s .. 16 s + 15) in the 8
//            columns kTcHCol(t, s).  Epilogue C reads S in the feature order 16..31, 0..15 and writes the terms of the
//            features it has finished into S columns it has already consumed: features 16..31 (columns 64..127 of S) free
//            64..127, of which 64..87 take K step 1; K step 0 goes to 88..111 once features 24..31 are done.
__host__ __device__ constexpr uint32_t kTcHCol(int term, int kstep) { return (kstep ? 64u : 88u) + 8u * (uint32_t)term; }
static_assert(kTcOffMeta + 256 <= kTcOffPool, "shared-memory map overflows into the pool");

struct TcMeta {  // one graph of the tile; adj / hoff / yoff: byte offsets of its three regions in the pool
    int v0, nv, fb, nb, R, Kp, adj, hoff, yoff, e0, nnz, g;
};
struct TcTileInfo {
    int ng, nblocks, pool_used, blkg[4];
};

struct TcParams {
    const int *tiles;  // 32 ints per tile: ng, then per graph (g, v0, nv, e0, nnz, -)
    int n_tiles;
    int *tile_counter;
    const int *graph_ptr, *row_ptr, *col_idx;
    const uint16_t *col16;  // graph-local column ids (compact host format) or nullptr: then col_idx, batch-global
    int upper;              // the lists hold only the entries with column > row: every edge sets both adjacency bytes
    const double *wts;
    const uint8_t *keep_in;
    const float *x0;
    float x0val;
    int remove_zero;
    int n_hidden;
    const float *first;  // [3 * 32]: colsum(W_0), colsum(W_1), bias of layer 0
    int first_act;
    const unsigned char *wall;  // n_hidden blobs of kTcWBlob bytes
    const int *acts;
    const float *tail;  // [2 * 32]
    float tail_bias, tail_norm;
    int last_act;
    float alpha;
    int predict;
    uint8_t *member;
    float *score;
    double *util;
    double *total;
    int *steps;
    int *status;
    int round_cap;
    int do_lgs;
    long long *dbg;
    long long *trace;   // DG_TC_TRACE builds: per-warp event clocks of one tile [16 warps][24 layers][8 events]
    int trace_tile;
    int *watchdog;      // pinned host memory the wait site of a protocol error is left in
    int watchdog_wide;
};

// ---- PTX helpers ----------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t s32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_inval(uint64_t *bar) {
    asm volatile("mbarrier.inval.shared::cta.b64 [%0];" ::"r"(s32(bar)) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_test(uint64_t *bar, uint32_t parity) {
    uint32_t done;
    asm volatile(
        "{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
        : "=r"(done)
        : "r"(s32(bar)), "r"(parity)
        : "memory");
    return done != 0;
}
// Bounded wait: a protocol error must end the launch with an error, never hang the GPU.  The wait site that gave up is
// left in pinned host memory (dg_context::h_flag[3]) for the post-mortem.
// (the pointer and the "wide table" flag of DG_TC_DEBUG travel as kernel parameters and are parked in shared memory at
// kTcOffWatch, where the slow path of mbar_wait finds them: no per-launch symbol copies on the stream)
struct TcWatch {
    int *ptr;
    int wide;  // DG_TC_DEBUG: the pointer is a wide table, every stuck thread logs its site
};
// The fast path is a tight loop around try_wait (a suspend-time hint made wake-ups slower and bought nothing:
// profiles/r01_notes.md); only after ~64 K failed attempts does the instrumented loop run.
__device__ __forceinline__ bool mbar_wait_fast(uint64_t *bar, uint32_t parity) {
    uint32_t done;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        ".reg .u32 n;\n"
        "mov.u32 n, 0;\n"
        "TC_WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "@p bra TC_WAIT_DONE;\n"
        "add.u32 n, n, 1;\n"
        "setp.lt.u32 p, n, 65536;\n"
        "@p bra TC_WAIT_LOOP;\n"
        "setp.ne.u32 p, n, n;\n"
        "TC_WAIT_DONE:\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(done)
        : "r"(s32(bar)), "r"(parity)
        : "memory");
    return done != 0u;
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity, int site = 0) {
    if (mbar_wait_fast(bar, parity)) return;
    const long long t0 = clock64();
    bool logged = false;
    extern __shared__ __align__(1024) unsigned char tc_smem_base[];
    int *const g_tc_watchdog = reinterpret_cast<const TcWatch *>(tc_smem_base + kTcOffWatch)->ptr;
    const int g_tc_watchdog_wide = reinterpret_cast<const TcWatch *>(tc_smem_base + kTcOffWatch)->wide;
    for (uint32_t spins = 1;; ++spins) {
        if (mbar_test(bar, parity)) return;
        if ((spins & 1023u) == 0u) {
            const long long dt = clock64() - t0;
            if (!logged && dt > 1000000000LL && g_tc_watchdog && g_tc_watchdog_wide) {
                g_tc_watchdog[1 + blockIdx.x * 512 + threadIdx.x] = site;
                __threadfence_system();
                logged = true;
            }
            if (dt > 3000000000LL) {
                if (g_tc_watchdog) {
                    *g_tc_watchdog = site | ((int)threadIdx.x << 8) | ((int)blockIdx.x << 20);
                    __threadfence_system();
                }
                __trap();
            }
        }
    }
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(s32(dst)),
                 "l"(src), "r"(bytes), "r"(s32(bar))
                 : "memory");
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void bar_sync(int id, int count) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory");
}

// shared-memory matrix descriptor, no swizzle (cute::UMMA::SmemDescriptor): start >> 4, LBO >> 4 at bit 16, SBO >> 4 at
// bit 32, version 1 at bit 46.  K-major operand: core matrix = 8 rows x 16 B, SBO = stride between 8-row groups, LBO =
// stride between the 16-byte K chunks.  MN-major operand: core matrix = 8 k x 16 B (16 B = consecutive MN elements),
// LBO = stride between 8-k groups, SBO = stride between 16-byte MN runs.  (Both verified by profiles/micro/umma_probe.cu.)
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((saddr >> 4) & 0x3fffu) | ((uint64_t)((lbo_bytes >> 4) & 0x3fffu) << 16) |
           ((uint64_t)((sbo_bytes >> 4) & 0x3fffu) << 32) | ((uint64_t)1 << 46);
}
// instruction descriptors (cute::UMMA::InstrDescriptor): c_format bit 4, a/b format bits 7/10, b MN-major bit 16, N >> 3
// at bit 17, M >> 4 at bit 24
__host__ __device__ constexpr uint32_t idesc_bf16(int M, int N) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
// A unsigned 8-bit (the 0/1 adjacency), B signed 8-bit (balanced base-256 digits), B MN-major, s32 accumulators
__host__ __device__ constexpr uint32_t idesc_u8(int M, int N) {
    return (2u << 4) | (1u << 10) | (1u << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void mma_bf16(uint32_t d_tmem, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}" ::"r"(
                     d_tmem),
                 "l"(a), "l"(b), "r"(idesc), "r"(acc)
                 : "memory");
}
// A operand in tensor memory (lane = row, 8 columns per 16-element K step, element 2j / 2j+1 in the low / high half of
// column j: profiles/micro/umma_ts_probe.cu), B from shared memory
__device__ __forceinline__ void mma_bf16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n}" ::"r"(
                     d_tmem),
                 "r"(a_tmem), "l"(b), "r"(idesc), "r"(acc)
                 : "memory");
}
__device__ __forceinline__ void tmem_st4(uint32_t taddr, const uint4 &v) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1,%2,%3,%4};" ::"r"(taddr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w)
                 : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void mma_u8(uint32_t d_tmem, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n}" ::"r"(
                     d_tmem),
                 "l"(a), "l"(b), "r"(idesc), "r"(acc)
                 : "memory");
}
__device__ __forceinline__ void mma_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(s32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t *r) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%"
        "30,%31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
// split form: issue a 16-column load, do other work, then wait (the wait names the registers so that no use of them can
// be scheduled above it)
__device__ __forceinline__ void tmem_ld16_issue(uint32_t taddr, uint32_t *r) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld16_wait(uint32_t *r) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]),
                   "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15])
                 :
                 : "memory");
}
__device__ __forceinline__ void tmem_ld4(uint32_t taddr, uint32_t *r) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
                 : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// Every activation of the path is v >= 0 ? v : slope * v with slope = alpha (leaky ReLU), 0 (ReLU) or 1 (none); as
// max(v, slope v) it is branch-free and bit-identical to the select form for slope <= 1 (models with alpha > 1 are not
// taken by this kernel: tc_model_eligible).
// one lane of a converged warp (the lane that issues tcgen05.mma / commit for it)
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n.reg .pred p;\nelect.sync _|p, 0xffffffff;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(pred));
    return pred != 0u;
}
// a value every lane holds identically, handed to the compiler as warp-uniform (so that it can live in a uniform register
// and feed the MMA operands without a per-instruction broadcast)
__device__ __forceinline__ uint32_t uniform(uint32_t v) { return __shfl_sync(0xffffffffu, v, 0); }

__device__ __forceinline__ float tc_slope(int act, float alpha) {
    return act == DG_ACT_LEAKY_RELU ? alpha : act == DG_ACT_RELU ? 0.f : 1.f;
}
__device__ __forceinline__ float tc_act(float v, float slope) { return fmaxf(v, v * slope); }

// fixed-point scale for values bounded by `bound` (>= 0): q = 2^k with |v| * q < 2^30, and its inverse
__device__ __forceinline__ void tc_scale(float bound, float *q, float *inv_q) {
    int eb = (__float_as_int(bound) >> 23) & 0xff;  // bound < 2^(eb - 126)
    eb = min(max(eb, 30), 254);
    *q = __int_as_float((283 - eb) << 23);      // 2^(156 - eb)
    *inv_q = __int_as_float((eb - 29) << 23);   // 2^(eb - 156)
}
// fixed-point image v = rn(y q) (|v| < 2^30) as four balanced base-256 digits d_a in [-128, 127], v = sum_a 256^a d_a:
// the low three bytes of v + 0x808080 are d_0..d_2 offset by 128 (flip their top bits), its top byte is d_3 as it stands.
// The offset rides on the multiply: y q + 8421504 rounds to the same integer as rn(y q) + 8421504 (the constant is a
// multiple of every ulp involved), so the conversion is one FMA, one F2I and one XOR.
__device__ __forceinline__ uint32_t tc_digits(float y, float q) {
    return (uint32_t)__float2int_rn(fmaf(y, q, 8421504.f)) ^ 0x00808080u;
}
// sum_a 256^a D_a as a float: the two digit pairs are combined exactly in integers (|D_a| <= 128 deg), converted, and
// joined with one FMA
__device__ __forceinline__ float tc_combine(const uint32_t *d) {
    const int lo = (int)d[1] * 256 + (int)d[0];
    const int hi = (int)d[3] * 256 + (int)d[2];
    return fmaf((float)hi, 65536.f, (float)lo);
}

// Packed fp32 pairs (FFMA2 / FADD2 / FMUL2 of sm_100): the same IEEE result per element as the scalar instruction, half
// the issue slots - the epilogues are bound by instruction issue (profiles/r01_notes.md).
__device__ __forceinline__ uint64_t pk(float lo, float hi) {
    return (uint64_t)__float_as_uint(lo) | ((uint64_t)__float_as_uint(hi) << 32);
}
__device__ __forceinline__ float pk_lo(uint64_t v) { return __uint_as_float((uint32_t)v); }
__device__ __forceinline__ float pk_hi(uint64_t v) { return __uint_as_float((uint32_t)(v >> 32)); }
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) {
    uint64_t d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ uint64_t mul2(uint64_t a, uint64_t b) {
    uint64_t d;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ uint64_t add2(uint64_t a, uint64_t b) {
    uint64_t d;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ uint64_t sub2(uint64_t a, uint64_t b) {
    uint64_t d;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}

// 8 feature values -> three bf16 term vectors (hi, mid, lo), each 8 x bf16 = 16 B
__device__ __forceinline__ void tc_split8(const float *h, uint4 *hi, uint4 *mid, uint4 *lo) {
    uint32_t t0[4], t1[4], t2[4];
#pragma unroll
    for (int p = 0; p < 4; ++p) {
        uint64_t ab = pk(h[2 * p], h[2 * p + 1]);
        uint32_t w;
        asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(w) : "f"(pk_hi(ab)), "f"(pk_lo(ab)));
        t0[p] = w;
        ab = sub2(ab, (uint64_t)(w << 16) | ((uint64_t)(w & 0xffff0000u) << 32));
        asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(w) : "f"(pk_hi(ab)), "f"(pk_lo(ab)));
        t1[p] = w;
        ab = sub2(ab, (uint64_t)(w << 16) | ((uint64_t)(w & 0xffff0000u) << 32));
        asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(w) : "f"(pk_hi(ab)), "f"(pk_lo(ab)));
        t2[p] = w;
    }
    *hi = make_uint4(t0[0], t0[1], t0[2], t0[3]);
    *mid = make_uint4(t1[0], t1[1], t1[2], t1[3]);
    *lo = make_uint4(t2[0], t2[1], t2[2], t2[3]);
}

// 4 adjacency bytes (0/1 each) -> 4 bits (byte k -> bit k)
__device__ __forceinline__ uint32_t tc_bits4(uint32_t w) { return (w * 0x01020408u) >> 24; }

// DIT: the GCN embedded into the greedy iteration (MWISSolver.solve_mwis_dit, mwis_gdpg_call.py:278-318): the layers and ONE
// greedy round repeat on the residual graph (vertices neither taken nor excluded yet) until no graph of the tile has a
// residual vertex of positive weight.  The adjacency bytes are built once; a residual graph is the same bytes seen through
// the current keep mask (degrees count kept columns only, removed vertices have dinv = 0 and so contribute zero digits).
// All graphs of a tile iterate together (they share the weight ring); a finished graph idles through the layers with an
// empty mask.
#ifdef DG_TC_TRACE
#define TC_TRACE(h, e)                                                                                       \
    do {                                                                                                     \
        if (P.trace && t == P.trace_tile && lane == 0 && (h) < 24)                                           \
            P.trace[((size_t)warp * 24 + (h)) * 8 + (e)] = clock64();                                        \
    } while (0)
#else
#define TC_TRACE(h, e) do { } while (0)
#endif
template <bool DIT>
__global__ void __launch_bounds__(kTcThreads, kTcCtasPerSm) tc_solve_kernel(const TcParams P) {
    extern __shared__ __align__(1024) unsigned char smem[];
    unsigned char *wring = smem + kTcOffRing;
    double *util_sm = reinterpret_cast<double *>(smem + kTcOffUtil);
    uint32_t *keepw = reinterpret_cast<uint32_t *>(smem + kTcOffBits);
    uint32_t *remain = keepw + 16, *joined = keepw + 32, *memb = keepw + 48;
    uint32_t *gmax = reinterpret_cast<uint32_t *>(smem + kTcOffGmax);  // [4][2]
    uint32_t *gmax0 = gmax + 8;                                         // [4]
    unsigned char *kb = smem + kTcOffDinv;                              // [512] DIT: current keep mask as bytes, by tile slot
    uint32_t *gpos = reinterpret_cast<uint32_t *>(smem + kTcOffDinv + 512);  // [4] DIT: the graph still has positive residual weight
    uint64_t *bar_full = reinterpret_cast<uint64_t *>(smem + kTcOffBar);  // [2]  weight ring
    uint32_t *cnt_p = reinterpret_cast<uint32_t *>(bar_full + 22);     // [4] per block: warps whose H terms are in place
    uint32_t *cnt_a = cnt_p + 4;                                        // [4] per graph: warps whose Y digits are in place
    uint32_t *cnt_w = cnt_p + 8;                                        // [2] per ring buffer: blocks done with its layer
    uint32_t *base_da = cnt_p + 10;                                     // [4] per block: parity of bar_da's phases before this tile
    uint64_t *bar_dp = bar_full + 6;    // [4] per block: projection MMAs complete
    uint64_t *bar_da = bar_full + 10;   // [4] per block: aggregation MMAs complete
    uint64_t *bar_ra = bar_full + 14;   // [4] per graph: the aggregation's B operand (Y digits) is complete (one arrival per warp)
    uint64_t *bar_mx = bar_full + 18;   // [4] per graph: every vertex has published its max |H|
    TcMeta *meta = reinterpret_cast<TcMeta *>(smem + kTcOffMeta);
    TcTileInfo *tinfo = reinterpret_cast<TcTileInfo *>(smem + kTcOffMeta + sizeof(TcMeta) * kTcMaxG);
    unsigned char *pool = smem + kTcOffPool;
    __shared__ int tile_sm;
    __shared__ int trec[32];
    __shared__ uint32_t tmem_sm;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n_hidden = P.n_hidden;

    if (tid == 0) {
        TcWatch *w = reinterpret_cast<TcWatch *>(smem + kTcOffWatch);
        w->ptr = P.watchdog;
        w->wide = P.watchdog_wide;
        for (int i = 0; i < 2; ++i) mbar_init(&bar_full[i], 1);
        for (int i = 0; i < kTcBlocks; ++i) {
            mbar_init(&bar_dp[i], 1);
            mbar_init(&bar_da[i], 1);
            mbar_init(&bar_mx[i], 1);
            mbar_init(&bar_ra[i], 1);
        }
        fence_mbar_init();
    }
    if (warp == 0) {  // the whole tensor memory of the SM: 4 blocks x 128 accumulator columns
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s32(&tmem_sm)), "r"(kTcTmemCols));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_sm;
    uint32_t wseq = 0;  // hidden-layer weight fills of the tiles this CTA has finished (every thread keeps its own copy)
    // The per-block barriers bar_dp / bar_da are armed ONCE (prologue) and their phases run on across tiles: a thread always
    // belongs to the same block (tid >> 7), so it carries the parity of the phases its block's barriers have completed in
    // earlier tiles.  (Re-arming them per tile - mbarrier.inval + mbarrier.init, which ptxas lowers to an invalidate and a
    // plain 64-bit store when the address is not uniform - was seen to leave block 0's barriers in their OLD state: invisible
    // while a tile completes an even number of phases, i.e. with the shipped 20-layer model, a dead-lock with an odd number
    // of hidden layers; profiles/micro/tc_odd_depth_probe.py.)
    uint32_t ea_base = 0, ep_base = 0;

    const bool timing = P.dbg != nullptr && tid == 0;
    long long tm[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    const long long t_begin = timing ? clock64() : 0;
    unsigned long long ns_begin = 0;
    if (timing) asm volatile("mov.u64 %0, %globaltimer;" : "=l"(ns_begin));

    for (;;) {
        if (tid == 0) tile_sm = atomicAdd(P.tile_counter, 1);
        __syncthreads();
        const int t = tile_sm;
        if (t >= P.n_tiles) break;
        long long tk = timing ? clock64() : 0;
        const long long t_tile = tk;

        // ---- tile metadata; the barriers are re-armed with this tile's thread counts ---------------------------
        if (warp == 0) trec[lane] = __ldg(P.tiles + (size_t)t * 32 + lane);  // the tile's record in one coalesced load
        __syncwarp();
        if (tid == 0) {
            const int *td = trec;
            const int ng = td[0];
            int fb = 0, off = 0;
            for (int gi = 0; gi < ng; ++gi) {
                const int *gd = td + 1 + 6 * gi;  // (the host knows every graph's extent: no dependent loads here)
                TcMeta m;
                m.g = gd[0];
                m.v0 = gd[1];
                m.nv = gd[2];
                m.nb = (m.nv + 127) >> 7;
                m.fb = fb;
                m.R = (m.nv + 7) & ~7;
                m.Kp = (m.nv + 31) & ~31;
                m.adj = off;
                m.hoff = off + m.R * m.Kp;
                m.yoff = m.hoff;
                off = m.yoff + 128 * m.Kp;
                m.e0 = gd[3];
                m.nnz = gd[4];
                meta[gi] = m;
                for (int jb = 0; jb < m.nb; ++jb) tinfo->blkg[fb + jb] = gi;
                fb += m.nb;
                mbar_inval(&bar_mx[gi]);
                mbar_inval(&bar_ra[gi]);
                mbar_init(&bar_mx[gi], 4 * m.nb);
                mbar_init(&bar_ra[gi], 4 * m.nb);
            }
            tinfo->ng = ng;
            tinfo->nblocks = fb;
            tinfo->pool_used = off;
            for (int i = 0; i < 12; ++i) gmax[i] = 0u;
            for (int i = 0; i < 10; ++i) cnt_p[i] = 0u;
            fence_mbar_init();
        }
        __syncthreads();
        const int ng = tinfo->ng;
        const int nblocks = tinfo->nblocks;

        // the first two hidden layers' weights start streaming in (every MMA of the previous tile has completed)
        if (tid == 0) {
            for (int h = 0; h < n_hidden && h < 2; ++h) {
                const uint32_t buf = (wseq + h) & 1u;
                mbar_expect_tx(&bar_full[buf], kTcWBlob);
                bulk_g2s(wring + buf * kTcWBlob, P.wall + (size_t)h * kTcWBlob, kTcWBlob, &bar_full[buf]);
            }
        }

        // ---- which graph / block / row this thread is -----------------------------------------------------------
        int gi = -1, r = 0, v = 0;
        TcMeta G{};
        bool valid = false, keep = false;
        double wt_v = 0.0;
        if (tid < kTcVertexThreads) {
            const int b = tid >> 7;
            if ((tid & 127) == 0) base_da[b] = ea_base;   // read by the other blocks of the graph (agg_drained), two barriers on
            if (b < nblocks) gi = tinfo->blkg[b];
            if (gi >= 0) {
                G = meta[gi];
                r = (b - G.fb) * 128 + (tid & 127);
                valid = r < G.nv;
                v = G.v0 + r;
            }
            if (valid) {
                if (P.wts) wt_v = P.wts[v];
                keep = P.keep_in ? P.keep_in[v] != 0 : true;
                if (P.remove_zero) keep = keep && (wt_v != 0.0);  // mwis_dqn_call.py:203
                if (P.member) P.member[v] = 0;
                if (keep && P.x0) atomicMax(&gmax0[gi], __float_as_uint(fabsf(P.x0[v])));
            }
            const uint32_t kw = __ballot_sync(0xffffffffu, keep);
            if (lane == 0) {
                keepw[warp] = kw;
                memb[warp] = 0u;
            }
            for (int k = 0; k < ng; ++k) {
                const TcMeta m = meta[k];
                // clear the adjacency bytes
                uint4 *a16 = reinterpret_cast<uint4 *>(pool + m.adj);
                const int n16 = (m.R * m.Kp) >> 4;
                for (int i = tid; i < n16; i += kTcVertexThreads) a16[i] = make_uint4(0u, 0u, 0u, 0u);
            }
            // Edge -> row table of this thread's vertex, parked in the graph's Y region (free until the first layer):
            // the scatter below then needs one shared-memory load per edge instead of a binary search in row_ptr.
            if (valid) {
                uint16_t *rowof = reinterpret_cast<uint16_t *>(pool + G.hoff);
                if (2 * G.nnz <= 128 * G.Kp) {
                    const int beg = P.row_ptr[v] - G.e0, end = P.row_ptr[v + 1] - G.e0;
                    for (int e = beg; e < end; ++e) rowof[e] = (uint16_t)r;
                }
            }
        }
        __syncthreads();
        // ---- dense adjacency of the kept sub-graph: byte (row i, column j) = 1 for every edge with both ends kept ---
        if (tid < kTcVertexThreads) {
            for (int k = 0; k < ng; ++k) {
                const TcMeta m = meta[k];
                unsigned char *adj = pool + m.adj;
                const uint16_t *rowof = reinterpret_cast<const uint16_t *>(pool + m.hoff);
                const bool table = 2 * m.nnz <= 128 * m.Kp;
                // kStageEdges edges per thread and pass, their global loads issued back to back (the loop is latency bound)
                for (int eb = tid; eb < m.nnz; eb += kStageEdges * kTcVertexThreads) {
                    int jj[kStageEdges], ll[kStageEdges];
                    if (P.col16) {  // (uniform branch outside the loads: they must issue back to back)
#pragma unroll
                        for (int u = 0; u < kStageEdges; ++u) {
                            const int e = eb + u * kTcVertexThreads;
                            jj[u] = e < m.nnz ? (int)__ldg(P.col16 + m.e0 + e) : -1;
                        }
                    } else {
#pragma unroll
                        for (int u = 0; u < kStageEdges; ++u) {
                            const int e = eb + u * kTcVertexThreads;
                            jj[u] = e < m.nnz ? __ldg(P.col_idx + m.e0 + e) - m.v0 : -1;
                        }
                    }
#pragma unroll
                    for (int u = 0; u < kStageEdges; ++u) {
                        const int e = eb + u * kTcVertexThreads;
                        int lo = 0;
                        if (e < m.nnz) {
                            if (table) {
                                lo = rowof[e];
                            } else {  // (a nearly complete graph: the table does not fit)
                                int hi = m.nv;  // row_ptr[lo] <= e < row_ptr[hi]
                                while (hi - lo > 1) {
                                    const int mid = (lo + hi) >> 1;
                                    if (__ldg(P.row_ptr + m.v0 + mid) - m.e0 <= e) lo = mid; else hi = mid;
                                }
                            }
                        }
                        ll[u] = lo;
                    }
#pragma unroll
                    for (int u = 0; u < kStageEdges; ++u) {
                        const int j = jj[u], lo = ll[u];
                        if (j >= 0) {
                            const uint32_t ki = (keepw[m.fb * 4 + (lo >> 5)] >> (lo & 31)) & 1u;
                            const uint32_t kj = (keepw[m.fb * 4 + (j >> 5)] >> (j & 31)) & 1u;
                            if (ki & kj) {
                                adj[(size_t)(j >> 4) * m.R * 16 + lo * 16 + (j & 15)] = 1;
                                if (P.upper) adj[(size_t)(lo >> 4) * m.R * 16 + j * 16 + (lo & 15)] = 1;
                            }
                        }
                    }
                }
            }
            fence_async_smem();
        }
        __syncthreads();
        if (timing) {
            const long long now = clock64();
            tm[0] += now - tk;
            tk = now;
        }

        int dit_steps = 0;
        uint32_t ea = 0, ep = 0, em = 0;  // aggregation / projection / max events consumed -> barrier parities (per tile)
        for (int dit_iter = 0;; ++dit_iter) {
        if (DIT) {
            // residual bookkeeping of this iteration (whole CTA): a graph goes on while a residual vertex has positive weight
            if (tid < 4) gpos[tid] = 0u;
            if (tid < 8) gmax[tid] = 0u;     // the layer maxima (the previous iteration's last layer left one slot set)
            __syncthreads();
            if (gi >= 0 && valid && keep && wt_v > 0.0) gpos[gi] = 1u;
            __syncthreads();
            if (gi >= 0) keep = keep && gpos[gi] != 0u;
            kb[tid] = keep ? 1 : 0;
            const uint32_t kw = __ballot_sync(0xffffffffu, keep);
            if (lane == 0) keepw[warp] = kw;
            if (!__syncthreads_or(keep ? 1 : 0)) {
                if (dit_iter == 0) {   // nothing to solve in this tile: consume the weight fills issued at its top
                    const uint32_t issued = (uint32_t)min(n_hidden, 2);
                    if (tid == 0)
                        for (uint32_t h = 0; h < issued; ++h) mbar_wait(&bar_full[(wseq + h) & 1u], ((wseq + h) >> 1) & 1u, 11);
                    wseq += issued;
                }
                break;
            }
            if (gi >= 0 && gpos[gi]) ++dit_steps;
            if (dit_iter > 0 && tid == 0) {   // this iteration's first two weight fills (the previous iteration has drained)
                for (int h = 0; h < n_hidden && h < 2; ++h) {
                    const uint32_t buf = (wseq + h) & 1u;
                    mbar_expect_tx(&bar_full[buf], kTcWBlob);
                    bulk_g2s(wring + buf * kTcWBlob, P.wall + (size_t)h * kTcWBlob, kTcWBlob, &bar_full[buf]);
                }
            }
        }
        if (gi >= 0) {
            // ================= vertex threads =========================================================================
            const int b = tid >> 7;
            const int jb = b - G.fb;
            const int lastb = G.fb + G.nb - 1;
            const int dom_bar = 1 + G.fb, dom_cnt = 128 * G.nb;
            // (the same for every lane of the warp: handed over as warp-uniform so that it lives in a uniform register
            // instead of being recomputed from the thread id in front of every tensor-memory access)
            const uint32_t taddr = uniform(tmem + (uint32_t)b * 128u + ((uint32_t)((warp & 3) * 32) << 16));
            // a warp none of whose lanes is a vertex (the tail of a graph's last block) keeps every hand-off and barrier of
            // the protocol but skips the arithmetic and the tensor-memory traffic of the epilogues
            const bool live = __any_sync(0xffffffffu, valid);
            unsigned char *adj = pool + G.adj;
            unsigned char *yrow = pool + G.yoff + (r >> 3) * 128 + (r & 7) * 16;  // ... and in the Y digit runs

            // ---- hand-off to the tensor cores: the warp that completes an operand issues the MMAs that consume it ----
            const uint32_t pool_addr = s32(pool);
            // true for the warp whose arrival completes `total` (other warps' operand stores are then visible to it)
            auto last_warp = [&](uint32_t *cnt, uint32_t total) -> bool {
                __syncwarp();
                uint32_t old = 0u;
                if (lane == 0) {
                    __threadfence_block();
                    old = atomicAdd(cnt, 1u);
                }
                old = __shfl_sync(0xffffffffu, old, 0);
                if (old != total - 1u) return false;
                if (lane == 0) *cnt = 0u;  // next use follows the completion of the MMAs issued now
                __threadfence_block();
                return true;
            };
            // Aggregation D[block] = A[block rows, :] . Y, N = 16 (scalar) or 128 (hidden layer), for every block of the
            // graph, issued by the warp that completes Y (one issuer: the commits of the blocks then complete in order,
            // which is what agg_drained() relies on).
            auto hand_off_agg = [&](bool scalar) {
                if (!last_warp(&cnt_a[gi], 4u * (uint32_t)G.nb)) return;
                // the whole (converged) warp runs the loop on warp-uniform values; one elected lane issues
                tc_fence_after();
                const uint32_t u_r = uniform((uint32_t)G.R), u_kp = uniform((uint32_t)G.Kp), u_nb = uniform((uint32_t)G.nb);
                const uint32_t u_adj = uniform(pool_addr + (uint32_t)G.adj), u_y = uniform(pool_addr + (uint32_t)G.yoff);
                const uint32_t u_d = uniform(tmem + (uint32_t)G.fb * 128u), u_bar = uniform(s32(&bar_da[G.fb]));
                const uint32_t idesc = scalar ? idesc_u8(128, 16) : idesc_u8(128, 128);
                const uint32_t lbo_a = u_r * 16u;
                const uint64_t bdesc0 = umma_desc(u_y, 128u, scalar ? 128u : u_kp * 16u);
                const uint64_t adesc0 = umma_desc(u_adj, lbo_a, 128u);
                const uint32_t a_step = (2u * lbo_a) >> 4;  // one K-step (32 columns = 2 chunks) in descriptor units
                const uint32_t ksteps = u_kp >> 5;
                const bool leader = elect_one();
                for (uint32_t k = 0; k < u_nb; ++k) {
                    uint64_t adesc = adesc0 + (uint64_t)(k * 128u);  // 128 rows x 16 B >> 4
                    uint64_t bdesc = bdesc0;
                    const uint32_t d = u_d + k * 128u;
                    for (uint32_t s2 = 0; s2 < ksteps; ++s2) {
                        if (leader) mma_u8(d, adesc, bdesc, idesc, s2 > 0u);
                        adesc += a_step;
                        bdesc += 32u;  // 4 k-groups x 128 B >> 4
                    }
                    if (leader)
                        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(u_bar + k * 8u)
                                     : "memory");
                }
                __syncwarp();
            };
            // projection of this block, hidden layer h: [P0 | P1] = H . [W_0 | W_1 r], 6 products x 2 K-steps
            auto issue_proj = [&](int h) {
                const uint32_t seq = wseq + (uint32_t)h;
                if (lane == 0) {
                    if (h >= 1) {
                        // every thread of this block has finished layer h-1 (its epilogue reads bias / r from the weight
                        // buffer): when all blocks of the tile have, the buffer takes layer h+1
                        const uint32_t pbuf = (seq - 1u) & 1u;
                        if (atomicAdd(&cnt_w[pbuf], 1u) == (uint32_t)nblocks - 1u) {
                            cnt_w[pbuf] = 0u;
                            if (h + 1 < n_hidden) {
                                mbar_expect_tx(&bar_full[pbuf], kTcWBlob);
                                bulk_g2s(wring + pbuf * kTcWBlob, P.wall + (size_t)(h + 1) * kTcWBlob, kTcWBlob, &bar_full[pbuf]);
                            }
                        }
                    }
                    mbar_wait(&bar_full[seq & 1u], (seq >> 1) & 1u, 2);  // the layer's weights have landed
                }
                __syncwarp();
                tc_fence_after();
                // warp-uniform operands, one elected lane issues (see hand_off_agg)
                const uint32_t u_w = uniform(s32(wring) + (seq & 1u) * kTcWBlob);
                const uint32_t d = uniform(tmem + (uint32_t)b * 128u);
                const uint32_t u_bar = uniform(s32(&bar_dp[b]));
                const uint64_t b0 = umma_desc(u_w, 1024u, 128u);
                const uint32_t idesc = idesc_bf16(128, 64);
                // smallest products first: (lo,hi) (hi,lo) (mid,mid) (mid,hi) (hi,mid) (hi,hi); the H terms are this block's
                // tensor-memory columns kTcHCol(term, K step), W term t starts 4096 t bytes in, its second K step 2 chunks on
                const uint64_t bt[3] = {b0, b0 + 256u, b0 + 512u};
                const int ta[6] = {2, 0, 1, 1, 0, 0}, tb[6] = {0, 2, 1, 0, 1, 0};
                if (elect_one()) {
#pragma unroll
                    for (int pp = 0; pp < 6; ++pp) {
                        mma_bf16_ts(d, d + kTcHCol(ta[pp], 0), bt[tb[pp]], idesc, pp != 0);
                        mma_bf16_ts(d, d + kTcHCol(ta[pp], 1), bt[tb[pp]] + 128u, idesc, 1u);
                    }
                    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(u_bar) : "memory");
                }
                __syncwarp();
            };

            // degree on the kept sub-graph, dinv, x0, and the scalar operand of the rank-1 first layer
            unsigned deg = 0;
            if (valid && (!DIT || keep)) {
                const int nch = G.Kp >> 4;
                for (int c = 0; c < nch; ++c) {
                    const uint4 w = *reinterpret_cast<const uint4 *>(adj + (size_t)c * G.R * 16 + r * 16);
                    uint4 km = make_uint4(0x01010101u, 0x01010101u, 0x01010101u, 0x01010101u);
                    if (DIT) km = *reinterpret_cast<const uint4 *>(kb + G.fb * 128 + c * 16);  // kept columns only
                    deg = __dp4a(w.x, km.x, deg);
                    deg = __dp4a(w.y, km.y, deg);
                    deg = __dp4a(w.z, km.z, deg);
                    deg = __dp4a(w.w, km.w, deg);
                }
            }
            const float di = deg > 0 ? (float)(1.0 / sqrt((double)deg)) : 0.f;  // gcn/utils.py:122-125
            const float xi = keep ? (P.x0 ? P.x0[v] : P.x0val) : 0.f;
            float q0, iq0;
            tc_scale(P.x0 ? __uint_as_float(gmax0[gi]) : fabsf(P.x0val), &q0, &iq0);
            if (valid) *reinterpret_cast<uint32_t *>(yrow) = tc_digits(di * xi, q0);
            fence_async_smem();
            hand_off_agg(true);

            float hmax = 0.f, t0 = 0.f, t1 = 0.f;
            // consumes 8 new feature values of this vertex: operand terms for the next projection, or the last
            // layer's two dot products when no hidden layer follows
            auto emit = [&](int q, const float *hv, bool to_terms) {
                if (to_terms) {
                    // (rows past the graph's last vertex write whatever they hold: their outputs are ignored)
                    uint4 hi, mid, lo;
                    tc_split8(hv, &hi, &mid, &lo);
                    const uint32_t col = taddr + 4u * (uint32_t)(q & 1);
                    tmem_st4(col + kTcHCol(0, q >> 1), hi);
                    tmem_st4(col + kTcHCol(1, q >> 1), mid);
                    tmem_st4(col + kTcHCol(2, q >> 1), lo);
                } else {
                    const float4 wa = __ldg(reinterpret_cast<const float4 *>(P.tail) + 2 * q);
                    const float4 wb = __ldg(reinterpret_cast<const float4 *>(P.tail) + 2 * q + 1);
                    const float4 wc = __ldg(reinterpret_cast<const float4 *>(P.tail + 32) + 2 * q);
                    const float4 wd = __ldg(reinterpret_cast<const float4 *>(P.tail + 32) + 2 * q + 1);
                    t0 = fmaf(hv[0], wa.x, t0), t0 = fmaf(hv[1], wa.y, t0), t0 = fmaf(hv[2], wa.z, t0), t0 = fmaf(hv[3], wa.w, t0);
                    t0 = fmaf(hv[4], wb.x, t0), t0 = fmaf(hv[5], wb.y, t0), t0 = fmaf(hv[6], wb.z, t0), t0 = fmaf(hv[7], wb.w, t0);
                    t1 = fmaf(hv[0], wc.x, t1), t1 = fmaf(hv[1], wc.y, t1), t1 = fmaf(hv[2], wc.z, t1), t1 = fmaf(hv[3], wc.w, t1);
                    t1 = fmaf(hv[4], wd.x, t1), t1 = fmaf(hv[5], wd.y, t1), t1 = fmaf(hv[6], wd.z, t1), t1 = fmaf(hv[7], wd.w, t1);
                }
#pragma unroll
                for (int k = 0; k < 8; ++k) hmax = fmaxf(hmax, fabsf(hv[k]));
            };
            // publishes this vertex's contribution to the graph's max |H| (fixed-point bound of the next aggregation)
            auto publish_max = [&](int slot) {
                const uint32_t mine = (valid && keep) ? __float_as_uint(hmax) : 0u;
                const uint32_t wmax = __reduce_max_sync(0xffffffffu, mine);
                if (lane == 0) {
                    atomicMax(&gmax[gi * 2 + slot], wmax);
                    mbar_arrive(&bar_mx[gi]);   // one arrival per warp (release: the maximum above is visible to the waiters)
                }
                hmax = 0.f;
            };
            // the graph's max |H|, once every vertex of the graph has published
            auto graph_max = [&](int slot) -> float {
                mbar_wait(&bar_mx[gi], em & 1u, 3);
                ++em;
                return __uint_as_float(gmax[gi * 2 + slot]);
            };
            // Every block's MMAs of the graph's last aggregation have completed (commits complete in issue order, so the
            // graph's last block tells): the Y region may take the last layer's scalar operand.  (Between hidden layers no
            // such wait is needed: a block writes new Y digits only after its own next projection has completed, and the
            // tensor pipe runs the MMAs in issue order, so every aggregation issued before that projection is done.)
            auto agg_drained = [&]() {
                if (b != lastb) mbar_wait(&bar_da[lastb], (base_da[lastb] + ea - 1u) & 1u, 4);
            };

            // -- first layer (rank 1): s = (L x0)_i, H1 = act(x0 colsum(W_0) + s colsum(W_1) + b) -----------------
            mbar_wait(&bar_da[b], (ea + ea_base) & 1u, 6);
            ++ea;
            tc_fence_after();
            if (live) {
                uint32_t d4[4];
                tmem_ld4(taddr, d4);
                const float s_i = xi - di * (tc_combine(d4) * iq0);
                const float sl0 = tc_slope(P.first_act, P.alpha);
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    float hv[8];
#pragma unroll
                    for (int half = 0; half < 2; ++half) {
                        const float4 a0 = __ldg(reinterpret_cast<const float4 *>(P.first) + 2 * q + half);
                        const float4 a1 = __ldg(reinterpret_cast<const float4 *>(P.first + 32) + 2 * q + half);
                        const float4 b0 = __ldg(reinterpret_cast<const float4 *>(P.first + 64) + 2 * q + half);
                        hv[4 * half + 0] = tc_act(fmaf(s_i, a1.x, fmaf(xi, a0.x, b0.x)), sl0);
                        hv[4 * half + 1] = tc_act(fmaf(s_i, a1.y, fmaf(xi, a0.y, b0.y)), sl0);
                        hv[4 * half + 2] = tc_act(fmaf(s_i, a1.z, fmaf(xi, a0.z, b0.z)), sl0);
                        hv[4 * half + 3] = tc_act(fmaf(s_i, a1.w, fmaf(xi, a0.w, b0.w)), sl0);
                    }
                    emit(q, hv, n_hidden > 0);
                }
            }

            if (timing) {
                const long long now = clock64();
                tm[8] += now - tk;
                tk = now;
            }
            // -- hidden layers -----------------------------------------------------------------------------------------
            for (int h = 0; h < n_hidden; ++h) {
                // this block's H terms are in place (tensor memory): its projection may start
                tmem_st_wait();
                TC_TRACE(h, 0);
                tc_fence_before();
                if (last_warp(&cnt_p[b], 4u)) issue_proj(h);
                publish_max(h & 1);
                TC_TRACE(h, 1);
                const uint32_t seq = wseq + h;
                const unsigned char *wb = wring + (seq & 1u) * kTcWBlob;
                const float *bias = reinterpret_cast<const float *>(wb + 12288);
                const float *rinv = bias + 32;
                long long tq = timing ? clock64() : 0;
                mbar_wait(&bar_dp[b], (ep + ep_base) & 1u, 7);
                ++ep;
                tc_fence_after();
                TC_TRACE(h, 2);
                if (timing) {
                    const long long now = clock64();
                    tm[3] += now - tq;
                    tq = now;
                }
                mbar_wait(&bar_full[seq & 1u], (seq >> 1) & 1u, 8);  // (already complete: the projection has read it)
                float qs, iqs;
                tc_scale(graph_max(h & 1) * bias[64], &qs, &iqs);
                TC_TRACE(h, 3);
                const float dq = di * qs, ndq = -di * iqs;
                float c[32];
                uint32_t pa[16], pb[16];
                // TMEM loads run one 16-column piece ahead of the arithmetic: P1 (columns 32..63, scaled by r), then P0
                const uint64_t dq2 = pk(dq, dq);
                auto half_b = [&](const uint32_t *pp, int f0) {
#pragma unroll
                    for (int g4 = 0; g4 < 4; ++g4) {
                        const int f = f0 + 4 * g4;
                        const float4 ri = *reinterpret_cast<const float4 *>(rinv + f);
                        const float4 bi = *reinterpret_cast<const float4 *>(bias + f);
                        const uint64_t p01 = (uint64_t)pp[4 * g4] | ((uint64_t)pp[4 * g4 + 1] << 32);
                        const uint64_t p23 = (uint64_t)pp[4 * g4 + 2] | ((uint64_t)pp[4 * g4 + 3] << 32);
                        const uint64_t c01 = fma2(p01, pk(ri.x, ri.y), pk(bi.x, bi.y));
                        const uint64_t c23 = fma2(p23, pk(ri.z, ri.w), pk(bi.z, bi.w));
                        c[f + 0] = pk_lo(c01), c[f + 1] = pk_hi(c01), c[f + 2] = pk_lo(c23), c[f + 3] = pk_hi(c23);
                        // run f / 4 of the MN-major operand = features f .. f + 3, digit a of feature f at column 4 f + a
                        const uint64_t y01 = fma2(p01, dq2, pk(8421504.f, 8421504.f));  // tc_digits, two at a time
                        const uint64_t y23 = fma2(p23, dq2, pk(8421504.f, 8421504.f));
                        const uint4 dg4 = make_uint4((uint32_t)__float2int_rn(pk_lo(y01)) ^ 0x00808080u,
                                                     (uint32_t)__float2int_rn(pk_hi(y01)) ^ 0x00808080u,
                                                     (uint32_t)__float2int_rn(pk_lo(y23)) ^ 0x00808080u,
                                                     (uint32_t)__float2int_rn(pk_hi(y23)) ^ 0x00808080u);
                        if (valid) *reinterpret_cast<uint4 *>(yrow + (size_t)(f >> 2) * G.Kp * 16) = dg4;
                    }
                };
                if (live) {
                tmem_ld16_issue(taddr + 32, pa);
                tmem_ld16_wait(pa);
                tmem_ld16_issue(taddr + 48, pb);
                half_b(pa, 0);
                tmem_ld16_wait(pb);
                tmem_ld16_issue(taddr, pa);
                half_b(pb, 16);
                tmem_ld16_wait(pa);
                tmem_ld16_issue(taddr + 16, pb);
#pragma unroll
                for (int f = 0; f < 16; f += 2) {
                    const uint64_t t2 = add2(pk(c[f], c[f + 1]), (uint64_t)pa[f] | ((uint64_t)pa[f + 1] << 32));
                    c[f] = pk_lo(t2), c[f + 1] = pk_hi(t2);
                }
                tmem_ld16_wait(pb);
#pragma unroll
                for (int f = 0; f < 16; f += 2) {
                    const uint64_t t2 = add2(pk(c[16 + f], c[17 + f]), (uint64_t)pb[f] | ((uint64_t)pb[f + 1] << 32));
                    c[16 + f] = pk_lo(t2), c[17 + f] = pk_hi(t2);
                }
                }
                fence_async_smem();
                tc_fence_before();
                TC_TRACE(h, 4);
                hand_off_agg(false);
                TC_TRACE(h, 5);
                const float slope = tc_slope(__ldg(P.acts + h + 1), P.alpha);
                const bool more = h + 1 < n_hidden;
                if (timing) {
                    const long long now = clock64();
                    tm[4] += now - tq;
                    tq = now;
                }
                mbar_wait(&bar_da[b], (ea + ea_base) & 1u, 9);
                ++ea;
                tc_fence_after();
                TC_TRACE(h, 6);
                if (timing) {
                    const long long now = clock64();
                    tm[5] += now - tq;
                    tq = now;
                }
                if (r == 0) gmax[gi * 2 + (h & 1)] = 0u;  // every block of the graph has read it: all of Y was needed
                const uint64_t ndq2 = pk(ndq, ndq), slope2 = pk(slope, slope);
                auto half_c = [&](const uint32_t *pp, int f0, float *hv) {
                    const float4 ra = *reinterpret_cast<const float4 *>(rinv + f0);
                    const uint64_t rv2[2] = {pk(ra.x, ra.y), pk(ra.z, ra.w)};
#pragma unroll
                    for (int k = 0; k < 4; k += 2) {  // two features per packed instruction; per element as tc_combine / tc_act
                        const uint32_t *d = pp + 4 * k;
                        const int lo0 = (int)d[1] * 256 + (int)d[0], hi0 = (int)d[3] * 256 + (int)d[2];
                        const int lo1 = (int)d[5] * 256 + (int)d[4], hi1 = (int)d[7] * 256 + (int)d[6];
                        const uint64_t comb = fma2(pk((float)hi0, (float)hi1), pk(65536.f, 65536.f), pk((float)lo0, (float)lo1));
                        const uint64_t v2 = fma2(comb, mul2(ndq2, rv2[k >> 1]), pk(c[f0 + k], c[f0 + k + 1]));
                        const uint64_t t2 = mul2(v2, slope2);
                        hv[k] = fmaxf(pk_lo(v2), pk_lo(t2));
                        hv[k + 1] = fmaxf(pk_hi(v2), pk_hi(t2));
                    }
                };
                // feature groups in the order 2, 3, 0, 1 when H terms follow (see the column map); 0..3 for the last hidden
                // layer, whose dot products keep their summation order
                auto run_c = [&](auto qx_c) {
                    constexpr int qx = decltype(qx_c)::value;  // compile-time: c[] must stay in registers
                    tmem_ld16_issue(taddr + 32 * qx, pa);
#pragma unroll
                    for (int qi = 0; qi < 4; ++qi) {
                        const int q = qi ^ qx;
                        float hv[8];
                        tmem_ld16_wait(pa);
                        tmem_ld16_issue(taddr + 32 * q + 16, pb);
                        half_c(pa, 8 * q, hv);
                        tmem_ld16_wait(pb);
                        if (qi < 3) tmem_ld16_issue(taddr + 32 * ((qi + 1) ^ qx), pa);
                        half_c(pb, 8 * q + 4, hv + 4);
                        emit(q, hv, qx != 0);
                    }
                };
                if (live) {
                    if (more) run_c(std::integral_constant<int, 2>{}); else run_c(std::integral_constant<int, 0>{});
                }
                if (timing) tm[6] += clock64() - tq;
                TC_TRACE(h, 7);
            }

            // -- last layer, projected first: q = H.w_0 + z, z = H.w_1, score = act(q - dinv_i sum_j A_ij dinv_j z_j + b)
            float score;
            if (timing) tk = clock64();
            {
                publish_max(n_hidden & 1);
                float qs, iqs;
                tc_scale(graph_max(n_hidden & 1) * P.tail_norm, &qs, &iqs);
                agg_drained();
                if (valid) *reinterpret_cast<uint32_t *>(yrow) = tc_digits(di * t1, qs);
                fence_async_smem();
                tc_fence_before();
                hand_off_agg(true);
                mbar_wait(&bar_da[b], (ea + ea_base) & 1u, 10);
                ++ea;
                tc_fence_after();
                uint32_t d4[4];
                tmem_ld4(taddr, d4);
                tc_fence_before();
                const float sll = tc_slope(P.last_act, P.alpha);
                score = tc_act((t0 + t1) - di * (tc_combine(d4) * iqs) + P.tail_bias, sll);
                if (!keep) score = 0.f;
            }
            // -- utility (mwis_dqn_call.py:230-235) ------------------------------------------------------------------
            {
                const double u = !valid ? 0.0 : (P.predict == DG_PREDICT_MWIS) ? (double)score * wt_v : (double)score;
                util_sm[tid] = u;
                if (valid) {
                    if (P.score) P.score[v] = score;
                    if (P.util) P.util[v] = u;
                }
            }
            if (timing) {
                const long long now = clock64();
                tm[9] += now - tk;
                tk = now;
            }
            if (P.do_lgs) {
                // -- local greedy search (heuristics.py:77-116); neighbour sets as bit rows in registers -----------
                TC_TRACE(20, 0);
                uint32_t nbm[12];
#pragma unroll
                for (int w = 0; w < 12; ++w) nbm[w] = 0u;
                if (valid) {
                    const int nch = G.Kp >> 4;
#pragma unroll
                    for (int c = 0; c < 24; ++c) {
                        if (c < nch) {
                            const uint4 w = *reinterpret_cast<const uint4 *>(adj + (size_t)c * G.R * 16 + r * 16);
                            const uint32_t m16 = tc_bits4(w.x) | (tc_bits4(w.y) << 4) | (tc_bits4(w.z) << 8) | (tc_bits4(w.w) << 12);
                            nbm[c >> 1] |= m16 << ((c & 1) * 16);
                        }
                    }
                }
                uint32_t nbr[12];
#pragma unroll
                for (int w = 0; w < 12; ++w) nbr[w] = nbm[w];
                const int w0 = G.fb * 4, nw = G.nb * 4;
                if (lane == 0) remain[warp] = keepw[warp];
                TC_TRACE(20, 1);
                bar_sync(dom_bar, dom_cnt);
                TC_TRACE(20, 2);
                // nbm := the neighbours that beat this vertex (larger utility, or equal and smaller id - heuristics.py:
                // 96-111), found once; a round is then a handful of word operations: the vertex joins iff none of them
                // is still there.  NaN utilities beat and are beaten by nobody's rule: they block, as np.max does.
                {
                    const double wv = util_sm[tid];
#pragma unroll
                    for (int w = 0; w < 12; ++w) {
                        if (w < nw) {
                            uint32_t cand = nbm[w] & remain[w0 + w], beat = 0u;
                            const int base = (w0 + w) * 32;  // the word's first slot
                            while (cand) {
                                // up to kBeatUnroll neighbours per pass, their utilities loaded back to back (the loop is bound
                                // by the shared-memory round trip, and a warp runs as many passes as its busiest lane)
                                int bit[kBeatUnroll];
                                double wu[kBeatUnroll];
#pragma unroll
                                for (int k = 0; k < kBeatUnroll; ++k) {
                                    bit[k] = cand ? __ffs(cand) - 1 : -1;
                                    cand &= cand - 1;  // (0 stays 0)
                                }
#pragma unroll
                                for (int k = 0; k < kBeatUnroll; ++k) wu[k] = util_sm[bit[k] >= 0 ? base + bit[k] : tid];
#pragma unroll
                                for (int k = 0; k < kBeatUnroll; ++k) {
                                    const int us = base + bit[k];
                                    if (bit[k] >= 0 && !((wv > wu[k]) || (wv == wu[k] && tid < us))) beat |= 1u << bit[k];
                                }
                            }
                            nbm[w] = beat;
                        }
                    }
                }
                // nbr := all kept neighbours (for the removal of a joined vertex's neighbourhood)
                TC_TRACE(20, 3);
                int rounds = 0, steps = 0;
                for (;;) {
                    if (DIT && rounds >= 1) break;   // one greedy round per re-scoring
                    uint32_t any = 0u;
#pragma unroll
                    for (int w = 0; w < 12; ++w)
                        if (w < nw) any |= remain[w0 + w];
                    if (!any) break;
                    ++steps;
                    if (rounds >= P.round_cap) {
                        if (r == 0) atomicExch(P.status, DG_ERR_NOT_CONVERGED);
                        break;
                    }
                    const bool active = (remain[warp] >> lane) & 1u;
                    bool join = active;
#pragma unroll
                    for (int w = 0; w < 12; ++w)
                        if (w < nw && (nbm[w] & remain[w0 + w])) join = false;
                    if (join && P.member) P.member[v] = 1;
                    const uint32_t jw = __ballot_sync(0xffffffffu, join);
                    if (lane == 0) {
                        joined[warp] = jw;
                        memb[warp] |= jw;
                    }
                    bar_sync(dom_bar, dom_cnt);
                    bool still = active && !join;
                    if (still) {
#pragma unroll
                        for (int w = 0; w < 12; ++w)
                            if (w < nw && (nbr[w] & joined[w0 + w])) still = false;
                    }
                    const uint32_t rw = __ballot_sync(0xffffffffu, still);
                    if (lane == 0) remain[warp] = rw;
                    ++rounds;
                    bar_sync(dom_bar, dom_cnt);
                }
                TC_TRACE(20, 4);
#ifdef DG_TC_TRACE
                if (lane == 0 && P.trace && t == P.trace_tile) P.trace[((size_t)warp * 24 + 20) * 8 + 7] = rounds;   // (a count)
#endif
                if (DIT) {
                    keep = valid && ((remain[warp] >> lane) & 1u);   // the residual graph of the next iteration
                    bar_sync(dom_bar, dom_cnt);                      // (every thread of the graph has read `remain`)
                }
                if (!DIT && r == 0 && P.steps) P.steps[G.g] = steps;
                if (!DIT && P.total) {  // member weights through shared memory, summed in the order dg_fused.cu uses
                    util_sm[tid] = (valid && ((memb[warp] >> lane) & 1u)) ? wt_v : 0.0;  // mwis_dqn_call.py:241
                    bar_sync(dom_bar, dom_cnt);
                    if (jb == 0 && (warp & 3) == 0) {
                        double acc = 0.0;
                        for (int i = lane; i < G.nv; i += 32) acc += util_sm[G.fb * 128 + i];
#pragma unroll
                        for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
                        if (lane == 0) P.total[G.g] = acc;
                    }
                }
                TC_TRACE(20, 5);
            }
        }
        if (timing) {
            const long long now = clock64();
            tm[10] += now - tk;
            tk = now;
        }
        wseq += (uint32_t)n_hidden;
        if (!DIT) break;
        tc_fence_before();
        __syncthreads();  // the next iteration reuses tensor memory, the Y regions and the weight ring
        tc_fence_after();
        }  // dit_iter
        if (gi >= 0) {   // the phases this tile has added to the block's barriers
            ea_base = (ea_base + ea) & 1u;
            ep_base = (ep_base + ep) & 1u;
        }
        if (DIT && gi >= 0) {   // per-graph outputs of the iterative solve
            const TcMeta G2 = meta[gi];
            const int b2 = tid >> 7;
            if (r == 0 && P.steps) P.steps[G2.g] = dit_steps;
            if (P.total) {
                util_sm[tid] = (valid && ((memb[warp] >> lane) & 1u)) ? wt_v : 0.0;  // mwis_gdpg_call.py:313
                bar_sync(1 + G2.fb, 128 * G2.nb);
                if (b2 == G2.fb && (warp & 3) == 0) {
                    double acc = 0.0;
                    for (int i = lane; i < G2.nv; i += 32) acc += util_sm[G2.fb * 128 + i];
#pragma unroll
                    for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
                    if (lane == 0) P.total[G2.g] = acc;
                }
            }
        }
        tc_fence_before();
        __syncthreads();  // shared memory, tensor memory and the barriers are recycled by the next tile
        tc_fence_after();
        if (timing) {
            const long long now = clock64();
            tm[11] += now - tk;
            tm[2] += 1;
            long long *td2 = P.dbg + (size_t)gridDim.x * 12 + (size_t)t * 2;
            td2[0] = now - t_tile;
            td2[1] = t;
        }
    }
    if (timing) {
        tm[7] = clock64() - t_begin;
        unsigned long long ns_end;
        asm volatile("mov.u64 %0, %globaltimer;" : "=l"(ns_end));
        tm[1] = (long long)(ns_end - ns_begin);  // wall time of the same interval: tm[7] / tm[1] = SM clock in GHz
        for (int k = 0; k < 12; ++k) P.dbg[(size_t)blockIdx.x * 12 + k] = tm[k];
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(kTcTmemCols));
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

// ---- host side --------------------------------------------------------------------------------------------------------
uint16_t bf16_rne(float x) {
    uint32_t b;
    memcpy(&b, &x, 4);
    if ((b & 0x7f800000u) == 0x7f800000u) return (uint16_t)(b >> 16);
    b += 0x7fffu + ((b >> 16) & 1u);
    return (uint16_t)(b >> 16);
}
float bf16_to_f(uint16_t h) {
    const uint32_t b = (uint32_t)h << 16;
    float f;
    memcpy(&f, &b, 4);
    return f;
}

struct TcGraph {
    int g, nv, nb;
    size_t bytes;   // adjacency + Y digits
    long long cost;
};

}  // namespace

// Operand blobs of the hidden layers (dg_model_create): per layer the bf16 terms of [W_0 | W_1 r] in the K-major core
// matrix layout of the projection's B operand, then bias[32], 1/r[32] and the layer's fixed-point bound factor.
void tc_build_weights(int n_hidden, const float *const *w0, const float *const *w1, const float *const *bias, const int *c_in,
                      const int *c_out, std::vector<unsigned char> *blob) {
    blob->assign((size_t)n_hidden * kTcWBlob, 0);
    for (int h = 0; h < n_hidden; ++h) {
        unsigned char *dst = blob->data() + (size_t)h * kTcWBlob;
        float *fb = reinterpret_cast<float *>(dst + 12288);
        float norm[32], rscale[32];
        float maxnorm = 0.f;
        for (int f = 0; f < 32; ++f) {
            double s = 0.0;
            if (f < c_out[h])
                for (int k = 0; k < c_in[h]; ++k) s += fabs((double)w1[h][(size_t)k * c_out[h] + f]);
            norm[f] = (float)s;
            maxnorm = std::max(maxnorm, norm[f]);
        }
        for (int f = 0; f < 32; ++f) {
            rscale[f] = 1.f;
            if (norm[f] > 0.f && maxnorm > 0.f) {
                int e;
                frexp((double)maxnorm / (double)norm[f], &e);  // ratio = m 2^e, m in [0.5, 1)
                rscale[f] = (float)ldexp(1.0, e - 1);           // largest power of two <= ratio
            }
        }
        for (int n = 0; n < 64; ++n)
            for (int k = 0; k < 32; ++k) {
                float w = 0.f;
                const int f = n & 31;
                if (f < c_out[h] && k < c_in[h])
                    w = n < 32 ? w0[h][(size_t)k * c_out[h] + f] : w1[h][(size_t)k * c_out[h] + f] * rscale[f];
                const uint16_t t0 = bf16_rne(w);
                const float r1 = w - bf16_to_f(t0);
                const uint16_t t1 = bf16_rne(r1);
                const float r2 = r1 - bf16_to_f(t1);
                const uint16_t t2 = bf16_rne(r2);
                const uint16_t terms[3] = {t0, t1, t2};
                for (int term = 0; term < 3; ++term)
                    memcpy(dst + term * 4096 + (k >> 3) * 1024 + n * 16 + (k & 7) * 2, &terms[term], 2);
            }
        for (int f = 0; f < 32; ++f) {
            fb[f] = (bias && bias[h] && f < c_out[h]) ? bias[h][f] : 0.f;
            fb[32 + f] = 1.f / rscale[f];
        }
        fb[64] = maxnorm * (1.f + 1.f / 1024.f);
    }
}

namespace {

// tiles: first-fit over a window of open tiles, graphs in order of decreasing cost
// dynamic shared memory of one CTA
inline size_t tc_smem_bytes(const dg_context *ctx) {
    return kTcCtasPerSm == 1 ? (size_t)ctx->max_smem_optin - 1024 : (size_t)112 * 1024;
}

struct Open {  // a tile being filled
    int ng, blocks, g[kTcMaxG];
    size_t bytes;
    long long cost;
};

// Cost model of a tile in cycles.  Fitted on the 20-layer model (profiles/r01_notes.md r01-j: 1610 tiles of the BA set in five
// packings, rms error 2.2 %): 92.9 k per tile - the dependent hand-offs of the layers, whatever the tile holds - plus, per
// graph, 8.6 k per 128-row block + 43.6 per block and adjacency column + 2.3 per edge.  Everything but the per-edge term
// (staging, greedy rounds) and ~7 k of set-up is work per LAYER, so for a model with `layers` = n_hidden + 2 layers the
// constants scale: 7.0 k + 4.3 k per layer and tile, 429 per layer and block, 2.18 per layer, block and adjacency column.
struct TcCost {
    long long fixed, per_block, per_block_col_x100;
    explicit TcCost(int n_hidden) {
        const long long layers = n_hidden + 2;
        fixed = 7000 + 4295 * layers;
        per_block = 429 * layers;
        per_block_col_x100 = 218 * layers;
    }
    long long graph(int nb, int Kp, int nnz) const {
        return per_block * nb + (per_block_col_x100 * nb * Kp) / 100 + (23LL * nnz) / 10;
    }
};

// `bins` tiles filled longest-processing-time first: every graph, heaviest first, goes to the lightest tile that can take it
bool tc_pack_lpt(const std::vector<TcGraph> &gs, int bins, size_t pool, std::vector<Open> *out) {
    std::vector<const TcGraph *> order;
    order.reserve(gs.size());
    for (const TcGraph &t : gs)
        if (t.nb > 0) order.push_back(&t);
    std::sort(order.begin(), order.end(), [](const TcGraph *a, const TcGraph *c) { return a->cost != c->cost ? a->cost > c->cost : a->g < c->g; });
    std::vector<Open> tiles((size_t)bins, Open{});
    typedef std::pair<long long, int> Key;  // (cost so far, tile)
    std::vector<Key> heap;
    heap.reserve((size_t)bins);
    for (int i = 0; i < bins; ++i) heap.push_back(Key(0, i));
    auto cmp = std::greater<Key>();
    std::vector<Key> aside;
    for (const TcGraph *c : order) {
        int best = -1;
        aside.clear();
        while (!heap.empty()) {
            std::pop_heap(heap.begin(), heap.end(), cmp);
            const Key k = heap.back();
            heap.pop_back();
            const Open &o = tiles[(size_t)k.second];
            if (o.ng < kTcMaxG && o.blocks + c->nb <= kTcBlocks && o.bytes + c->bytes + kTcOverread <= pool) {
                best = k.second;
                break;
            }
            if (o.ng < kTcMaxG && o.blocks < kTcBlocks) aside.push_back(k);  // (a full tile never comes back)
        }
        for (const Key &k : aside) {
            heap.push_back(k);
            std::push_heap(heap.begin(), heap.end(), cmp);
        }
        if (best < 0) return false;
        Open &o = tiles[(size_t)best];
        o.g[o.ng++] = c->g;
        o.blocks += c->nb;
        o.bytes += c->bytes;
        o.cost += c->cost;
        heap.push_back(Key(o.cost, best));
        std::push_heap(heap.begin(), heap.end(), cmp);
    }
    tiles.erase(std::remove_if(tiles.begin(), tiles.end(), [](const Open &o) { return o.ng == 0; }), tiles.end());
    out->swap(tiles);
    return true;
}

// makespan of heaviest-first list scheduling on `workers` CTAs (what the kernel's atomic tile counter does); `tiles` sorted
long long tc_makespan(const std::vector<Open> &tiles, int workers, long long tile_fixed) {
    std::vector<long long> load((size_t)workers, 0);
    for (const Open &t : tiles) {
        std::pop_heap(load.begin(), load.end(), std::greater<long long>());
        load.back() += t.cost + tile_fixed;
        std::push_heap(load.begin(), load.end(), std::greater<long long>());
    }
    return *std::max_element(load.begin(), load.end());
}

// host half of the tile plan: fills b->tc_tiles_host / tc_skip / tc_n_tiles from the batch's host metadata
// (h_graph_ptr, h_graph_e).  No CUDA calls: dg_solve_graphs_host runs it on a pool thread while the others pack.
void tc_plan_tiles_host(dg_context *ctx, dg_batch *b, int n_hidden) {
    const TcCost cost_model(n_hidden);
    b->tc_plan_hidden = n_hidden;
    b->tc_n_tiles = 0;
    b->tc_tiles_host.clear();
    const size_t pool = tc_smem_bytes(ctx) - kTcOffPool;
    const auto &gp = b->h_graph_ptr;
    const auto &ge = b->h_graph_e;
    std::vector<TcGraph> gs((size_t)b->n_graphs);
    b->tc_skip.assign((size_t)b->n_graphs, 0);
    b->tc_n_skipped = 0;
    for (int g = 0; g < b->n_graphs; ++g) {
        TcGraph t;
        t.g = g;
        t.nv = gp[g + 1] - gp[g];
        t.nb = std::max(1, (t.nv + 127) / 128);
        const size_t R = (size_t)((t.nv + 7) & ~7), Kp = (size_t)((t.nv + 31) & ~31);
        t.bytes = R * Kp + 128 * Kp;
        if (t.nv <= 0 || t.nb > kTcBlocks || t.bytes + kTcOverread > pool) {
            // beyond this kernel's limits (or empty): the graph is left to the CUDA-core kernel, the rest of the batch stays here
            t.nb = 0;
            t.cost = 0;
            b->tc_skip[(size_t)g] = 1;
            b->tc_n_skipped++;
            gs[(size_t)g] = t;
            continue;
        }
        t.cost = cost_model.graph(t.nb, (int)Kp, ge[g + 1] - ge[g]);
        gs[(size_t)g] = t;
    }
    // Largest graph first; every tile is then topped up with the largest remaining graphs that still fit (blocks, bytes
    // and cost all grow with the vertex count, so "largest that fits" is a lookup in the size-ordered remainder).
    std::vector<Open> tiles;
    // Unplaced graphs as one stack per vertex count (LIFO: the last graph of a size goes first); prev_size[s] leads to the
    // largest non-empty size <= s (removals only: pointer jumping with path compression), so a fill is O(1) - this planner
    // runs on the host for every streamed batch, 16 k graphs at a time in the large configurations.
    const int max_size = kTcBlocks * 128;
    std::vector<int> head((size_t)max_size + 1, -1), next_same((size_t)b->n_graphs, -1), prev_size((size_t)max_size + 1);
    if (b->tc_n_skipped == b->n_graphs) return;  // nothing for this kernel
    for (const TcGraph &t : gs) {
        if (t.nb == 0) continue;
        next_same[(size_t)t.g] = head[(size_t)t.nv];
        head[(size_t)t.nv] = t.g;
    }
    prev_size[0] = 0;
    for (int sz = 1; sz <= max_size; ++sz) prev_size[(size_t)sz] = head[(size_t)sz] >= 0 ? sz : prev_size[(size_t)sz - 1];
    auto largest_at_most = [&](int sz) {
        int r = sz;
        while (r > 0 && prev_size[(size_t)r] != r) r = prev_size[(size_t)r];
        for (int q = sz; q > 0 && prev_size[(size_t)q] != q;) {  // path compression
            const int nq = prev_size[(size_t)q];
            prev_size[(size_t)q] = r;
            q = nq;
        }
        return r;
    };
    int remaining = b->n_graphs - b->tc_n_skipped;
    while (remaining > 0) {
        Open o{};
        while (o.ng < kTcMaxG && o.blocks < kTcBlocks) {
            // the largest graph that fits into what is left of the tile: blocks bound its size, then the pool bytes
            int sz = largest_at_most((kTcBlocks - o.blocks) * 128);
            while (sz > 0 && o.bytes + gs[(size_t)head[(size_t)sz]].bytes + kTcOverread > pool) sz = largest_at_most(sz - 1);
            if (sz == 0) break;
            const TcGraph *c = &gs[(size_t)head[(size_t)sz]];
            head[(size_t)sz] = next_same[(size_t)c->g];
            if (head[(size_t)sz] < 0) prev_size[(size_t)sz] = sz - 1;
            o.g[o.ng++] = c->g;
            o.blocks += c->nb;
            o.bytes += c->bytes;
            o.cost += c->cost;
            --remaining;
        }
        tiles.push_back(o);
    }
    std::stable_sort(tiles.begin(), tiles.end(), [](const Open &a, const Open &c) { return a.cost > c.cost; });
    // With few tiles per SM the dense packing leaves some SMs a whole tile short (250 tiles on 148 SMs: 102 run two, 46 one).
    // A tile costs a large fixed time whatever it holds, so a multiple of the SM count of lighter tiles, balanced by
    // cost, finishes earlier; both plans are simulated and the shorter one is kept.  (Large batches: no difference, skipped.)
    {
        const int sms = ctx->sm_count * kTcCtasPerSm;
        const int k = ((int)tiles.size() + sms - 1) / sms;
        int want = k * sms;
        if (ctx->env.tc_tiles > 0) want = ctx->env.tc_tiles;  // (experiments)
        if (k <= 8 && want > (int)tiles.size() && want <= b->n_graphs - b->tc_n_skipped) {
            std::vector<Open> lpt;
            if (tc_pack_lpt(gs, want, pool, &lpt)) {
                std::stable_sort(lpt.begin(), lpt.end(), [](const Open &a, const Open &c) { return a.cost > c.cost; });
                if (tc_makespan(lpt, sms, cost_model.fixed) < tc_makespan(tiles, sms, cost_model.fixed)) tiles.swap(lpt);
            }
        }
    }
    std::vector<int> flat(tiles.size() * 32, 0);
    for (size_t i = 0; i < tiles.size(); ++i) {
        flat[i * 32] = tiles[i].ng;
        for (int k = 0; k < tiles[i].ng; ++k) {
            const int g = tiles[i].g[k];
            int *d = &flat[i * 32 + 1 + 6 * k];
            d[0] = g, d[1] = gp[g], d[2] = gp[g + 1] - gp[g], d[3] = ge[g], d[4] = ge[g + 1] - ge[g];
        }
    }
    b->tc_n_tiles = (int)tiles.size();
    b->tc_tiles_host.swap(flat);
}

int tc_build_tiles(dg_context *ctx, dg_batch *b, int n_hidden, bool *ok) {
    *ok = false;
    if (b->tc_tiles_valid && b->tc_plan_hidden == n_hidden) {   // the plan depends on the batch and on the model's depth
        *ok = b->tc_n_tiles > 0;
        return DG_OK;
    }
    if (!b->tc_plan_ready || b->tc_plan_hidden != n_hidden) tc_plan_tiles_host(ctx, b, n_hidden);
    b->tc_plan_ready = false;
    b->tc_tiles_valid = true;
    const std::vector<int> &flat = b->tc_tiles_host;
    if (b->tc_tiles_cap < flat.size() || !b->tc_tiles_dev) {
        if (b->tc_tiles_dev) {
            DG_CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
            cudaFree(b->tc_tiles_dev);
            b->tc_tiles_dev = nullptr;
        }
        b->tc_tiles_cap = flat.size() + flat.size() / 4 + 8;
        DG_CUDA_CHECK(cudaMalloc((void **)&b->tc_tiles_dev, sizeof(int) * b->tc_tiles_cap));
    }
    if (!flat.empty())
        DG_CUDA_CHECK(cudaMemcpyAsync(b->tc_tiles_dev, flat.data(), sizeof(int) * flat.size(), cudaMemcpyHostToDevice,
                                      ctx->stream));
    if (ctx->env.fused_timing) fprintf(stderr, "[tc tiles] %d tiles for %d graphs\n", b->tc_n_tiles, b->n_graphs);
    *ok = b->tc_n_tiles > 0;
    return DG_OK;
}

}  // namespace

bool tc_model_eligible(const dg_context *ctx, const dg_model *m) {
    return !(ctx->env.disable_tc || ctx->env.disable_fused) && m->tc_wall && m->n_layers >= 3 && m->alpha <= 1.0f;
}

void tc_plan_ahead(dg_context *ctx, const dg_model *m, dg_batch *b) {
    if (!tc_model_eligible(ctx, m) || b->n_graphs == 0 || b->n_nodes == 0) return;
    tc_plan_tiles_host(ctx, b, m->n_layers - 2);
    b->tc_plan_ready = true;
}

static int *g_wide_buf = nullptr;  // DG_TC_DEBUG: pinned table the stuck threads log their wait sites into

int tc_try_solve(dg_context *ctx, const dg_model *m, dg_batch *b, const double *d_wts, int predict, int remove_zero_weight,
                 uint8_t *member, float *score, double *util, double *total, int32_t *steps, bool *handled, bool dit) {
    *handled = false;
    if (dit && (member == nullptr || d_wts == nullptr)) return DG_OK;
    if (ctx->env.disable_tc || ctx->env.disable_fused) return DG_OK;
    if (!m->tc_wall || m->n_layers < 3 || !(m->alpha <= 1.0f) || b->n_graphs == 0 || b->n_nodes == 0) return DG_OK;
    if ((int)b->h_graph_e.size() != b->n_graphs + 1) return DG_OK;
    if (member == nullptr && d_wts == nullptr && predict == DG_PREDICT_MWIS) return DG_OK;
    bool ok = false;
    DG_TRY(tc_build_tiles(ctx, b, m->n_layers - 2, &ok));
    if (!ok) return DG_OK;
    // Graphs beyond this kernel's limits go to the CUDA-core graph-resident kernel in a second launch.  If that kernel
    // cannot take them either, the whole batch belongs to the per-layer path: decline BEFORE launching anything.
    if (b->tc_n_skipped > 0 && (dit || !fused_fits(ctx, m, b))) return DG_OK;
    TcParams p{};
    p.tiles = b->tc_tiles_dev;
    p.n_tiles = b->tc_n_tiles;
    p.tile_counter = ctx->d_status + 1;
    p.graph_ptr = b->graph_ptr;
    p.row_ptr = b->upper_pending ? b->row_ptr_u : b->row_ptr;
    p.col_idx = b->col_idx;
    p.col16 = b->cols_pending ? b->col16 : nullptr;
    p.upper = b->upper_pending ? 1 : 0;
    p.wts = d_wts;
    p.keep_in = remove_zero_weight ? nullptr : b->keep;
    p.x0 = b->x0;
    p.x0val = 1.0f / (float)m->layers[0].c_in;
    p.remove_zero = remove_zero_weight ? 1 : 0;
    p.n_hidden = m->n_layers - 2;
    p.first = m->fused_first;
    p.first_act = m->layers[0].act;
    p.wall = m->tc_wall;
    p.acts = m->d_acts;
    p.tail = m->fused_tail;
    p.tail_bias = m->tail_bias;
    p.tail_norm = m->tc_tail_norm;
    p.last_act = m->layers.back().act;
    p.alpha = m->alpha;
    p.predict = predict;
    p.member = member;
    p.score = score;
    p.util = util;
    p.total = total;
    p.steps = steps;
    p.status = ctx->d_status;
    p.round_cap = kLgsRoundCap;
    p.do_lgs = member != nullptr ? 1 : 0;
    p.dbg = nullptr;
    const size_t smem = tc_smem_bytes(ctx);
    if (ctx->env.fused_timing) {
        long long *dbg = nullptr;
        DG_TRY(scratch_as(ctx, kSlotLgsWords, (size_t)ctx->sm_count * kTcCtasPerSm * 12 + (size_t)p.n_tiles * 2, &dbg));
        DG_CUDA_CHECK(cudaMemsetAsync(dbg, 0, sizeof(long long) * ((size_t)ctx->sm_count * kTcCtasPerSm * 12 + (size_t)p.n_tiles * 2), ctx->stream));
        p.dbg = dbg;
#ifdef DG_TC_TRACE
        DG_TRY(scratch_as(ctx, kSlotLgsWords, (size_t)ctx->sm_count * kTcCtasPerSm * 12 + (size_t)p.n_tiles * 2 + 16 * 24 * 8, &dbg));
        DG_CUDA_CHECK(cudaMemsetAsync(dbg, 0, sizeof(long long) * ((size_t)ctx->sm_count * kTcCtasPerSm * 12 + (size_t)p.n_tiles * 2 + 16 * 24 * 8), ctx->stream));
        p.dbg = dbg;
        p.trace = dbg + (size_t)ctx->sm_count * kTcCtasPerSm * 12 + (size_t)p.n_tiles * 2;
        p.trace_tile = getenv("DG_TC_TRACE_TILE") ? atoi(getenv("DG_TC_TRACE_TILE")) : 0;
#endif
    }
    if (!ctx->tc_attr_set) {  // once per context (the attribute is per device)
        DG_CUDA_CHECK(cudaFuncSetAttribute(tc_solve_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        DG_CUDA_CHECK(cudaFuncSetAttribute(tc_solve_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        ctx->tc_attr_set = true;
    }
    const int grid = std::min(ctx->sm_count * kTcCtasPerSm, p.n_tiles);
    {
        int *wd = ctx->h_flag + 3;  // pinned host memory: readable after a device-side trap
        int wide = 0;
        if (ctx->env.tc_debug) {
            if (!g_wide_buf) DG_CUDA_CHECK(cudaHostAlloc((void **)&g_wide_buf, sizeof(int) * (1 + 148 * 512), cudaHostAllocDefault));
            memset(g_wide_buf, 0, sizeof(int) * (1 + 148 * 512));
            wd = g_wide_buf;
            wide = 1;
        }
        p.watchdog = wd;
        p.watchdog_wide = wide;
    }
    {
        // work-equivalent algorithmic bytes, the same figure dg_fused.cu reports (SURVEY.md 8d / DESIGN.md)
        const double n = (double)b->n_nodes, nnz = (double)b->nnz * (b->upper_pending ? 2.0 : 1.0), cp = 32.0;
        const double csr = 4.0 * (n + 1) + 4.0 * nnz;
        const double hidden = (double)p.n_hidden * (csr + 4.0 * n + 8.0 * n * cp + 8.0 * cp * cp);
        const double scalar_passes = 2.0 * (csr + 12.0 * n);
        const double lgs = csr + 9.0 * n;
        ctx->last_kernel = "tc_solve_kernel";
        prof_begin(ctx);
        if (dit) tc_solve_kernel<true><<<grid, kTcThreads, smem, ctx->stream>>>(p);
        else tc_solve_kernel<false><<<grid, kTcThreads, smem, ctx->stream>>>(p);
        ctx->launches++;
        prof_end(ctx, hidden + scalar_passes + lgs);
        DG_CUDA_CHECK(cudaGetLastError());
    }
    if (ctx->env.tc_debug) {
        const cudaError_t e = cudaStreamSynchronize(ctx->stream);
        if (e != cudaSuccess && g_wide_buf) {
            const int w = g_wide_buf[0];
            const int cta = w >> 20;
            fprintf(stderr, "[tc debug] %s; first to give up: site %d, thread %d, CTA %d; stuck threads of that CTA (site x count):",
                    cudaGetErrorString(e), w & 0xff, (w >> 8) & 0xfff, cta);
            for (int wp = 0; wp < 16; ++wp) {
                fprintf(stderr, "\n  warp %2d:", wp);
                for (int l = 0; l < 32; ++l) fprintf(stderr, " %d", g_wide_buf[1 + cta * 512 + wp * 32 + l]);
            }
            fprintf(stderr, "\n");
        }
    }
    if (p.dbg) {
        std::vector<long long> h((size_t)ctx->sm_count * kTcCtasPerSm * 12 + (size_t)p.n_tiles * 2);
        DG_CUDA_CHECK(cudaMemcpyAsync(h.data(), p.dbg, sizeof(long long) * h.size(), cudaMemcpyDeviceToHost, ctx->stream));
        DG_CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
        const char *names[12] = {"stage", "wall_ns", "tiles", "w_proj", "phaseB", "w_agg", "phaseC", "total", "first", "tail", "greedy", "endwait"};
        for (int k : {0, 8, 3, 4, 5, 6, 9, 10, 11, 2, 7, 1}) {
            long long mn = -1, mx = 0;
            double sum = 0;
            int cnt = 0;
            for (int c = 0; c < grid; ++c) {
                const long long vv = h[(size_t)c * 12 + k];
                mn = mn < 0 ? vv : std::min(mn, vv);
                mx = std::max(mx, vv);
                sum += (double)vv;
                ++cnt;
            }
            fprintf(stderr, "[tc timing] %-6s min %10lld avg %12.0f max %10lld (%d CTAs)\n", names[k], mn, cnt ? sum / cnt : 0.0,
                    mx, cnt);
        }
#ifdef DG_TC_TRACE
        {
            std::vector<long long> tr(16 * 24 * 8);
            DG_CUDA_CHECK(cudaMemcpy(tr.data(), p.trace, sizeof(long long) * tr.size(), cudaMemcpyDeviceToHost));
            const int *td = &b->tc_tiles_host[(size_t)p.trace_tile * 32];
            fprintf(stderr, "[tc trace] tile %d:", p.trace_tile);
            for (int k = 0; k < td[0]; ++k) fprintf(stderr, " nv=%d nnz=%d", td[1 + 6 * k + 2], td[1 + 6 * k + 4]);
            fprintf(stderr, "\n");
            long long t0 = 0;
            for (int w = 0; w < 16; ++w) if (tr[(size_t)w * 24 * 8] && (!t0 || tr[(size_t)w * 24 * 8] < t0)) t0 = tr[(size_t)w * 24 * 8];
            for (int h = 0; h < 24; ++h)
                for (int w = 0; w < 16; ++w) {
                    const long long *e = &tr[((size_t)w * 24 + h) * 8];
                    if (!e[0]) continue;
                    fprintf(stderr, "[tc trace] h %2d warp %2d:", h, w);
                    for (int k = 0; k < 8; ++k) fprintf(stderr, " %7lld", (h == 20 && k == 7) ? e[k] : e[k] ? e[k] - t0 : -1);
                    fprintf(stderr, "\n");
                }
        }
#endif
        if (const char *path = ctx->env.tc_tile_dump.empty() ? nullptr : ctx->env.tc_tile_dump.c_str()) {  // per tile: cycles, then the vertex counts of its graphs
            if (FILE *f = fopen(path, "w")) {
                for (int t = 0; t < p.n_tiles; ++t) {
                    const int *td = &b->tc_tiles_host[(size_t)t * 32];
                    fprintf(f, "%lld", h[(size_t)grid * 12 + (size_t)t * 2]);
                    for (int k = 0; k < td[0]; ++k) fprintf(f, " %d:%d", td[1 + 6 * k + 2], td[1 + 6 * k + 4]);
                    fprintf(f, "\n");
                }
                fclose(f);
            }
        }
    }
    // graphs beyond this kernel's limits are solved next by the CUDA-core kernel (fused_try_solve reads the flag)
    b->tc_ran_partial = b->tc_n_skipped > 0;
    *handled = b->tc_n_skipped == 0;
    return DG_OK;
}

}  // namespace dg
