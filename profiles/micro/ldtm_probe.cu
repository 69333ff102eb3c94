// TMEM read throughput probe (B200): back-to-back tcgen05.ld.32x32b.x32 from 1 / 4 / 8 / 16 warps of one CTA.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ldtm_probe ldtm_probe.cu
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void __launch_bounds__(512, 1) k(long long *cyc, uint32_t *sink, int iters) {
    __shared__ uint32_t tmem_sm;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_sm)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_sm;
    const uint32_t taddr = tmem + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(warp >> 2) * 128u;
    for (int nw = 1; nw <= 16; nw *= 2) {
        __syncthreads();
        uint32_t acc = 0;
        const long long t0 = clock64();
        if (warp < nw) {
            for (int it = 0; it < iters; ++it) {
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    uint32_t r[32];
                    asm volatile(
                        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
                          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
                          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                        : "r"(taddr + q * 32));
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                    for (int c = 0; c < 32; ++c) acc ^= r[c];
                }
            }
        }
        const long long t1 = clock64();
        __syncthreads();
        if (threadIdx.x == 0) cyc[__ffs(nw) - 1] = t1 - t0;
        sink[threadIdx.x] = acc;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
}

int main() {
    long long *cyc;
    uint32_t *sink;
    cudaMalloc(&cyc, 64);
    cudaMalloc(&sink, 4 * 512);
    const int iters = 256;
    k<<<1, 512>>>(cyc, sink, iters);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
    long long h[8];
    cudaMemcpy(h, cyc, 64, cudaMemcpyDeviceToHost);
    for (int i = 0; i < 5; ++i) {
        const int nw = 1 << i;
        const double loads = (double)iters * 4;  // x32 loads per warp
        printf("%2d warps: %lld cycles, %.1f cycles per x32 load per warp, %.1f B/clk for the SM\n", nw, h[i], h[i] / loads,
               nw * loads * 4096.0 / h[i]);
    }
    return 0;
}
