// Probe: tcgen05.mma kind::f16 with the A operand in TENSOR MEMORY (the "TS" form) on B200.
//   (1) layout check: A [128 x 16] bf16 written by tcgen05.st.32x32b.x8 (thread = row = TMEM lane, register j = elements
//       2j (low half) and 2j+1 (high half)), B [64 x 16] bf16 K-major in shared memory; D = A.B^T against the CPU;
//   (2) a K = 32 chain with two A column groups, accumulating;
//   (3) timing: 12 x (M128 N64 K16) with A from TMEM vs A from shared memory.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o umma_ts_probe umma_ts_probe.cu ; run on the GPU box.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((saddr >> 4) & 0x3fffu) | ((uint64_t)((lbo_bytes >> 4) & 0x3fffu) << 16) |
           ((uint64_t)((sbo_bytes >> 4) & 0x3fffu) << 32) | ((uint64_t)1 << 46);
}
__host__ __device__ constexpr uint32_t idesc_bf16(int M, int N) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void mma_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}"
                 :: "r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a_tmem, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n}"
                 :: "r"(d), "r"(a_tmem), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    uint32_t done = 0;
    while (!done)
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                     : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t *r) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                 :: "r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t *r) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

struct Params {
    const uint16_t *a;   // [128][32] bf16 row-major
    const uint8_t *b;    // K-major core-matrix layout [4 chunks][64 rows][16 B] (K = 32)
    const uint8_t *a_sm; // the same A in the K-major core-matrix layout [4 chunks][128 rows][16 B]
    float *d;            // [3][128][64]
    long long *cycles;   // [4]
    int reps;
};

__global__ void __launch_bounds__(128, 1) probe(Params P) {
    __shared__ __align__(1024) uint8_t sb[4096];
    __shared__ __align__(1024) uint8_t sa[8192];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_sm;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < 4096 / 16; i += 128) reinterpret_cast<uint4 *>(sb)[i] = reinterpret_cast<const uint4 *>(P.b)[i];
    for (int i = tid; i < 8192 / 16; i += 128) reinterpret_cast<uint4 *>(sa)[i] = reinterpret_cast<const uint4 *>(P.a_sm)[i];
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(&bar)), "r"(1));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&tmem_sm)), "r"(256));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_sm;
    const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
    uint32_t phase = 0;

    // A rows into TMEM columns 64..79: K-step s at columns 64 + 8 s, register j = elements (2j, 2j+1) of the step
    for (int s = 0; s < 2; ++s) {
        uint32_t r[8];
        for (int j = 0; j < 8; ++j)
            r[j] = (uint32_t)P.a[tid * 32 + s * 16 + 2 * j] | ((uint32_t)P.a[tid * 32 + s * 16 + 2 * j + 1] << 16);
        tmem_st8(tmem + lane_base + 64 + 8 * s, r);
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

    const uint32_t idesc = idesc_bf16(128, 64);
    for (int test = 0; test < 3; ++test) {
        if (tid == 0) {
            if (test == 0) {  // K = 16, TS
                mma_ts(tmem, tmem + 64, make_desc(smem_u32(sb), 1024, 128), idesc, 0);
            } else if (test == 1) {  // K = 32, TS, two steps
                mma_ts(tmem, tmem + 64, make_desc(smem_u32(sb), 1024, 128), idesc, 0);
                mma_ts(tmem, tmem + 72, make_desc(smem_u32(sb) + 2048, 1024, 128), idesc, 1);
            } else {  // K = 32, SS (known-good reference path)
                mma_ss(tmem, make_desc(smem_u32(sa), 2048, 128), make_desc(smem_u32(sb), 1024, 128), idesc, 0);
                mma_ss(tmem, make_desc(smem_u32(sa) + 4096, 2048, 128), make_desc(smem_u32(sb) + 2048, 1024, 128), idesc, 1);
            }
            mma_commit(&bar);
        }
        mbar_wait(&bar, phase);
        phase ^= 1;
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        uint32_t r[16];
        for (int q = 0; q < 4; ++q) {
            tmem_ld16(tmem + lane_base + q * 16, r);
            for (int c = 0; c < 16; ++c) P.d[(size_t)test * 128 * 64 + (size_t)tid * 64 + q * 16 + c] = __uint_as_float(r[c]);
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    }
    // the whole (converged) warp 0 runs the loop on warp-uniform operands, one elected lane issues (as dg_tc.cu does): a
    // single thread in a divergent branch needs ~17 instructions per MMA and measures its own issue rate instead
    for (int test = 0; test < 4; ++test) {  // 0: TS, 1: SS, 2: TS alternating between two accumulators, 3: SS likewise
        long long t0 = 0;
        if (warp == 0) {
            const uint32_t sbu = __shfl_sync(0xffffffffu, smem_u32(sb), 0), sau = __shfl_sync(0xffffffffu, smem_u32(sa), 0);
            const uint32_t tm = __shfl_sync(0xffffffffu, tmem, 0);
            const uint64_t bd0 = make_desc(sbu, 1024, 128), bd1 = make_desc(sbu + 2048, 1024, 128);
            const uint64_t ad0 = make_desc(sau, 2048, 128), ad1 = make_desc(sau + 4096, 2048, 128);
            const uint32_t at0 = tm + 64, at1 = tm + 72;
            const uint32_t d1 = (test >= 2) ? tm + 128 : tm;
            uint32_t leader;
            asm volatile("{\n.reg .pred p;\nelect.sync _|p, 0xffffffff;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(leader));
            t0 = clock64();
            for (int rep = 0; rep < P.reps; ++rep) {
#pragma unroll
                for (int p = 0; p < 6; ++p) {
                    if ((test & 1) == 0) {
                        if (leader) mma_ts(tm, at0, bd0, idesc, 1);
                        if (leader) mma_ts(d1, at1, bd1, idesc, 1);
                    } else {
                        if (leader) mma_ss(tm, ad0, bd0, idesc, 1);
                        if (leader) mma_ss(d1, ad1, bd1, idesc, 1);
                    }
                }
            }
            if (leader) mma_commit(&bar);
            __syncwarp();
        }
        mbar_wait(&bar, phase);
        phase ^= 1;
        if (tid == 0) P.cycles[test] = clock64() - t0;
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    }
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"(256));
}

static uint16_t f2bf(float x) { uint32_t b; memcpy(&b, &x, 4); b += 0x7fffu + ((b >> 16) & 1u); return (uint16_t)(b >> 16); }
static float bf2f(uint16_t h) { uint32_t b = (uint32_t)h << 16; float f; memcpy(&f, &b, 4); return f; }

int main() {
    std::vector<uint16_t> a(128 * 32), bm(64 * 32);
    srand(1);
    for (auto &x : a) x = f2bf((float)rand() / RAND_MAX - 0.5f);
    for (auto &x : bm) x = f2bf((float)rand() / RAND_MAX - 0.5f);
    std::vector<uint8_t> bl(4096), al(8192);
    for (int n = 0; n < 64; ++n)
        for (int k = 0; k < 32; ++k) memcpy(&bl[(k / 8) * 1024 + n * 16 + (k % 8) * 2], &bm[n * 32 + k], 2);
    for (int m = 0; m < 128; ++m)
        for (int k = 0; k < 32; ++k) memcpy(&al[(k / 8) * 2048 + m * 16 + (k % 8) * 2], &a[m * 32 + k], 2);
    Params P;
    uint16_t *da; uint8_t *db, *dal; float *dd; long long *dc;
    CK(cudaMalloc(&da, a.size() * 2)); CK(cudaMalloc(&db, 4096)); CK(cudaMalloc(&dal, 8192));
    CK(cudaMalloc(&dd, 3 * 128 * 64 * 4)); CK(cudaMalloc(&dc, 32));
    CK(cudaMemcpy(da, a.data(), a.size() * 2, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(db, bl.data(), 4096, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dal, al.data(), 8192, cudaMemcpyHostToDevice));
    P.a = da; P.b = db; P.a_sm = dal; P.d = dd; P.cycles = dc; P.reps = 50;
    probe<<<1, 128>>>(P);
    CK(cudaDeviceSynchronize());
    std::vector<float> d(3 * 128 * 64);
    long long cyc[4];
    CK(cudaMemcpy(d.data(), dd, d.size() * 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(cyc, dc, 32, cudaMemcpyDeviceToHost));
    for (int test = 0; test < 3; ++test) {
        const int K = test == 0 ? 16 : 32;
        double worst = 0;
        for (int m = 0; m < 128; ++m)
            for (int n = 0; n < 64; ++n) {
                double ref = 0;
                for (int k = 0; k < K; ++k) ref += (double)bf2f(a[m * 32 + k]) * bf2f(bm[n * 32 + k]);
                worst = fmax(worst, fabs(ref - d[(size_t)test * 8192 + m * 64 + n]));
            }
        printf("test %d (%s, K = %d): max |error| = %.3g %s\n", test, test == 2 ? "A in shared memory" : "A in tensor memory", K, worst,
               worst < 1e-5 ? "OK" : "MISMATCH");
    }
    printf("M128 N64 K16 chain into one accumulator: A in tensor memory %.1f cycles / MMA, A in shared memory %.1f\n",
           cyc[0] / (50.0 * 12), cyc[1] / (50.0 * 12));
    printf("alternating between two accumulators:    A in tensor memory %.1f cycles / MMA, A in shared memory %.1f\n",
           cyc[2] / (50.0 * 12), cyc[3] / (50.0 * 12));
    return 0;
}
