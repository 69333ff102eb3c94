// Probe for the tcgen05 building blocks of the tensor-core solve kernel (dg_tc.cu) on B200:
//   (1) kind::f16 (bf16 x bf16 -> fp32), A [128 x 32] K-major and B [64 x 32] K-major, no swizzle, 3-term split of both
//       operands, 6 products chained into one TMEM accumulator  (the hidden-layer projection [H].[W0 | W1]);
//   (2) kind::i8  (u8 x u8 -> s32), A = 0/1 adjacency [128 x K] K-major, B = base-256 digits [K x 128] MN-major
//       (the neighbour aggregation  A . Y  in exact integer arithmetic);
//   (3) tcgen05.ld 32x32b lane/column mapping, (4) issue-to-completion cycles of MMA chains.
// Each test runs with the descriptor's LBO / SBO fields in both orders and reports which one reproduces the CPU result.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o umma_probe umma_probe.cu ; run on the GPU box.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>
#include <cuda_bf16.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3fffu);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3fffu) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3fffu) << 32;
    d |= (uint64_t)1 << 46;  // descriptor version (Blackwell)
    return d;                // base offset 0, layout type 0 = no swizzle
}

__device__ __forceinline__ void mma_f16(uint32_t d_tmem, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}"
                 :: "r"(d_tmem), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_i8(uint32_t d_tmem, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n}"
                 :: "r"(d_tmem), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    uint32_t done = 0;
    while (!done) {
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                     : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    }
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t *r) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// instruction descriptors (cute::UMMA::InstrDescriptor bit layout)
__host__ __device__ constexpr uint32_t idesc_f16_bf16(int M, int N) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__host__ __device__ constexpr uint32_t idesc_i8_u8(int M, int N, int b_mn_major) {
    return (2u << 4) | ((uint32_t)b_mn_major << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

constexpr int KAGG = 256;  // K of the aggregation test

struct Params {
    const uint8_t *a_f16;   // 3 terms x [4 chunks][128 rows][16 B]            = 24576 B
    const uint8_t *b_f16;   // 3 terms x [4 chunks][64 rows][16 B]             = 12288 B
    const uint8_t *a_i8;    // [KAGG/16 chunks][128 rows][16 B]                = 32768 B
    const uint8_t *b_i8;    // [8 runs][KAGG/8 groups][8][16 B]                = 32768 B
    float *d_f16;           // 2 variants x [128][64]
    int *d_i8;              // 2 variants x [128][128]
    long long *cycles;      // [4]
    int reps;
};

__global__ void __launch_bounds__(128, 1) probe_kernel(Params P) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t *sa_f16 = smem;                 // 24576
    uint8_t *sb_f16 = smem + 24576;         // 12288
    uint8_t *sa_i8 = smem + 36864;          // 32768
    uint8_t *sb_i8 = smem + 69632;          // 32768
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base_sm;
    const int tid = threadIdx.x, warp = tid >> 5;

    for (int i = tid; i < 24576 / 16; i += 128) reinterpret_cast<uint4 *>(sa_f16)[i] = reinterpret_cast<const uint4 *>(P.a_f16)[i];
    for (int i = tid; i < 12288 / 16; i += 128) reinterpret_cast<uint4 *>(sb_f16)[i] = reinterpret_cast<const uint4 *>(P.b_f16)[i];
    for (int i = tid; i < 32768 / 16; i += 128) reinterpret_cast<uint4 *>(sa_i8)[i] = reinterpret_cast<const uint4 *>(P.a_i8)[i];
    for (int i = tid; i < 32768 / 16; i += 128) reinterpret_cast<uint4 *>(sb_i8)[i] = reinterpret_cast<const uint4 *>(P.b_i8)[i];
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(&bar)), "r"(1));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&tmem_base_sm)), "r"(256));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_base_sm;
    uint32_t phase = 0;
    const uint32_t lane_base = (uint32_t)(warp * 32) << 16;

    // ---- (1) projection, two descriptor conventions -------------------------------------------------------
    for (int variant = 0; variant < 2; ++variant) {
        if (tid == 0) {
            const uint32_t idesc = idesc_f16_bf16(128, 64);
            // products ordered small to large: (lo,hi) (hi,lo) (mid,mid) (mid,hi) (hi,mid) (hi,hi)
            const int ta[6] = {2, 0, 1, 1, 0, 0}, tb[6] = {0, 2, 1, 0, 1, 0};
            uint32_t acc = 0;
            for (int p = 0; p < 6; ++p)
                for (int s = 0; s < 2; ++s) {
                    const uint32_t aaddr = smem_u32(sa_f16) + ta[p] * 8192 + s * 2 * 2048;
                    const uint32_t baddr = smem_u32(sb_f16) + tb[p] * 4096 + s * 2 * 1024;
                    const uint64_t ad = variant == 0 ? make_desc(aaddr, 2048, 128) : make_desc(aaddr, 128, 2048);
                    const uint64_t bd = variant == 0 ? make_desc(baddr, 1024, 128) : make_desc(baddr, 128, 1024);
                    mma_f16(tmem, ad, bd, idesc, acc);
                    acc = 1;
                }
            mma_commit(&bar);
        }
        mbar_wait(&bar, phase);
        phase ^= 1;
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        uint32_t r[32];
        for (int half = 0; half < 2; ++half) {
            tmem_ld32(tmem + lane_base + half * 32, r);
            for (int c = 0; c < 32; ++c) P.d_f16[(size_t)variant * 128 * 64 + (size_t)tid * 64 + half * 32 + c] = __uint_as_float(r[c]);
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    }
    // ---- (2) aggregation, two descriptor conventions for the MN-major operand -----------------------------
    for (int variant = 0; variant < 2; ++variant) {
        if (tid == 0) {
            const uint32_t idesc = idesc_i8_u8(128, 128, 1);
            uint32_t acc = 0;
            for (int s = 0; s < KAGG / 32; ++s) {
                const uint32_t aaddr = smem_u32(sa_i8) + s * 2 * 2048;
                const uint32_t baddr = smem_u32(sb_i8) + s * 4 * 128;
                const uint64_t ad = make_desc(aaddr, 2048, 128);
                const uint64_t bd = variant == 0 ? make_desc(baddr, 128, KAGG * 16) : make_desc(baddr, KAGG * 16, 128);
                mma_i8(tmem + 64, ad, bd, idesc, acc);
                acc = 1;
            }
            mma_commit(&bar);
        }
        mbar_wait(&bar, phase);
        phase ^= 1;
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        uint32_t r[32];
        for (int q = 0; q < 4; ++q) {
            tmem_ld32(tmem + 64 + lane_base + q * 32, r);
            for (int c = 0; c < 32; ++c) P.d_i8[(size_t)variant * 128 * 128 + (size_t)tid * 128 + q * 32 + c] = (int)r[c];
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    }
    // ---- (4) timing of MMA chains ---------------------------------------------------------------------------
    for (int test = 0; test < 4; ++test) {
        long long t0 = 0;
        if (tid == 0) {
            t0 = clock64();
            for (int rep = 0; rep < P.reps; ++rep) {
                if (test == 0) {  // projection: 12 MMAs M128 N64 K16
                    const uint32_t idesc = idesc_f16_bf16(128, 64);
                    for (int p = 0; p < 12; ++p)
                        mma_f16(tmem, make_desc(smem_u32(sa_f16) + (p % 6) * 4096, 2048, 128),
                                make_desc(smem_u32(sb_f16) + (p % 6) * 2048, 1024, 128), idesc, 1);
                } else if (test == 1) {  // aggregation: 8 MMAs M128 N128 K32
                    const uint32_t idesc = idesc_i8_u8(128, 128, 1);
                    for (int s = 0; s < 8; ++s)
                        mma_i8(tmem + 64, make_desc(smem_u32(sa_i8) + s * 4096, 2048, 128),
                               make_desc(smem_u32(sb_i8) + s * 512, 128, KAGG * 16), idesc, 1);
                } else if (test == 2) {  // aggregation with N = 96
                    const uint32_t idesc = idesc_i8_u8(128, 96, 1);
                    for (int s = 0; s < 8; ++s)
                        mma_i8(tmem + 64, make_desc(smem_u32(sa_i8) + s * 4096, 2048, 128),
                               make_desc(smem_u32(sb_i8) + s * 512, 128, KAGG * 16), idesc, 1);
                } else {  // bf16 aggregation shape M128 N96 K16 (for comparison)
                    const uint32_t idesc = idesc_f16_bf16(128, 96);
                    for (int s = 0; s < 8; ++s)
                        mma_f16(tmem + 64, make_desc(smem_u32(sa_i8) + s * 4096, 2048, 128),
                                make_desc(smem_u32(sb_i8) + s * 512, 1024, 128), idesc, 1);
                }
            }
            mma_commit(&bar);
        }
        mbar_wait(&bar, phase);
        phase ^= 1;
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (tid == 0) P.cycles[test] = clock64() - t0;
        __syncthreads();
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"(256));
}

static uint16_t bf16_rn(float x) {
    uint32_t b;
    memcpy(&b, &x, 4);
    b += 0x7fffu + ((b >> 16) & 1u);
    return (uint16_t)(b >> 16);
}
static float bf16_f(uint16_t h) {
    uint32_t b = (uint32_t)h << 16;
    float f;
    memcpy(&f, &b, 4);
    return f;
}

int main() {
    srand(1);
    // (1) projection operands
    std::vector<float> A(128 * 32), B(64 * 32);
    for (auto &v : A) v = (float)rand() / RAND_MAX * 2.f - 1.f;
    for (auto &v : B) v = ((float)rand() / RAND_MAX * 2.f - 1.f) * 0.3f;
    std::vector<uint8_t> a_f16(24576, 0), b_f16(12288, 0);
    auto split3 = [](float x, uint16_t t[3]) {
        t[0] = bf16_rn(x);
        float r1 = x - bf16_f(t[0]);
        t[1] = bf16_rn(r1);
        float r2 = r1 - bf16_f(t[1]);
        t[2] = bf16_rn(r2);
    };
    for (int r = 0; r < 128; ++r)
        for (int k = 0; k < 32; ++k) {
            uint16_t t[3];
            split3(A[r * 32 + k], t);
            for (int term = 0; term < 3; ++term)
                memcpy(&a_f16[term * 8192 + (k / 8) * 2048 + r * 16 + (k % 8) * 2], &t[term], 2);
        }
    for (int n = 0; n < 64; ++n)
        for (int k = 0; k < 32; ++k) {
            uint16_t t[3];
            split3(B[n * 32 + k], t);
            for (int term = 0; term < 3; ++term)
                memcpy(&b_f16[term * 4096 + (k / 8) * 1024 + n * 16 + (k % 8) * 2], &t[term], 2);
        }
    // (2) aggregation operands
    std::vector<uint8_t> adj(128 * KAGG), dig(KAGG * 128);
    for (auto &v : adj) v = (rand() % 10) == 0;
    for (auto &v : dig) v = (uint8_t)(rand() & 255);
    std::vector<uint8_t> a_i8(32768, 0), b_i8(32768, 0);
    for (int r = 0; r < 128; ++r)
        for (int k = 0; k < KAGG; ++k) a_i8[(k / 16) * 2048 + r * 16 + (k % 16)] = adj[r * KAGG + k];
    for (int k = 0; k < KAGG; ++k)
        for (int n = 0; n < 128; ++n) b_i8[(n / 16) * (KAGG * 16) + (k / 8) * 128 + (k % 8) * 16 + (n % 16)] = dig[k * 128 + n];

    Params P{};
    uint8_t *d1, *d2, *d3, *d4;
    CK(cudaMalloc(&d1, 24576)); CK(cudaMalloc(&d2, 12288)); CK(cudaMalloc(&d3, 32768)); CK(cudaMalloc(&d4, 32768));
    CK(cudaMemcpy(d1, a_f16.data(), 24576, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d2, b_f16.data(), 12288, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d3, a_i8.data(), 32768, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d4, b_i8.data(), 32768, cudaMemcpyHostToDevice));
    float *df; int *di; long long *dc;
    CK(cudaMalloc(&df, sizeof(float) * 2 * 128 * 64)); CK(cudaMalloc(&di, sizeof(int) * 2 * 128 * 128)); CK(cudaMalloc(&dc, 8 * 4));
    P.a_f16 = d1; P.b_f16 = d2; P.a_i8 = d3; P.b_i8 = d4; P.d_f16 = df; P.d_i8 = di; P.cycles = dc; P.reps = 64;
    const int smem = 36864 + 65536;
    CK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    probe_kernel<<<1, 128, smem>>>(P);
    CK(cudaDeviceSynchronize());
    std::vector<float> hf(2 * 128 * 64);
    std::vector<int> hi(2 * 128 * 128);
    long long cyc[4];
    CK(cudaMemcpy(hf.data(), df, hf.size() * 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(hi.data(), di, hi.size() * 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(cyc, dc, 32, cudaMemcpyDeviceToHost));
    for (int variant = 0; variant < 2; ++variant) {
        double maxerr = 0, maxref = 0, maxerr32 = 0;
        for (int r = 0; r < 128; ++r)
            for (int n = 0; n < 64; ++n) {
                double ref = 0;
                float f32 = 0.f;
                for (int k = 0; k < 32; ++k) {
                    ref += (double)A[r * 32 + k] * (double)B[n * 32 + k];
                    f32 = fmaf(A[r * 32 + k], B[n * 32 + k], f32);
                }
                maxerr = fmax(maxerr, fabs(ref - hf[variant * 128 * 64 + r * 64 + n]));
                maxerr32 = fmax(maxerr32, fabs(ref - f32));
                maxref = fmax(maxref, fabs(ref));
            }
        printf("f16 projection variant %d (%s): max abs err %.3e (fp32 FMA chain: %.3e), max |ref| %.3f\n", variant,
               variant == 0 ? "LBO=K-chunk stride, SBO=8-row stride" : "swapped", maxerr, maxerr32, maxref);
    }
    for (int variant = 0; variant < 2; ++variant) {
        long long bad = 0;
        for (int r = 0; r < 128; ++r)
            for (int n = 0; n < 128; ++n) {
                int ref = 0;
                for (int k = 0; k < KAGG; ++k) ref += (int)adj[r * KAGG + k] * (int)dig[k * 128 + n];
                bad += ref != hi[variant * 128 * 128 + r * 128 + n];
            }
        printf("i8 aggregation variant %d (%s): %lld of 16384 entries differ\n", variant,
               variant == 0 ? "LBO=8-k group stride, SBO=16-n run stride" : "swapped", bad);
    }
    const char *names[4] = {"12 x f16 M128 N64 K16", "8 x i8 M128 N128 K32", "8 x i8 M128 N96 K32", "8 x f16 M128 N96 K16"};
    const int per[4] = {12, 8, 8, 8};
    for (int t = 0; t < 4; ++t)
        printf("timing %-24s: %lld cycles for %d reps -> %.1f cycles per MMA\n", names[t], cyc[t], 64,
               (double)cyc[t] / (64.0 * per[t]));
    return 0;
}
