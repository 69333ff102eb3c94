// Microbenchmark: shared-memory wavefront cost of LDS.128 / LDS.32 under broadcast patterns on B200.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o lds_patterns lds_patterns.cu ; run on the GPU box.
#include <cstdio>
#include <cuda_runtime.h>

__global__ void k128(const int *offs, float *out, long long *cyc, int iters) {
    __shared__ __align__(16) float sm[8192];
    for (int i = threadIdx.x; i < 8192; i += blockDim.x) sm[i] = (float)i;
    __syncthreads();
    int o = offs[threadIdx.x & 31];
    float4 acc = make_float4(0, 0, 0, 0);
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 16; ++u) {
            float4 v;
            unsigned a = (unsigned)__cvta_generic_to_shared(&sm[(o + u * 128) & 8191]);
            asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a));
            acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        }
    }
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc.x + acc.y + acc.z + acc.w;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

__global__ void k32(const int *offs, float *out, long long *cyc, int iters) {
    __shared__ __align__(16) float sm[8192];
    for (int i = threadIdx.x; i < 8192; i += blockDim.x) sm[i] = (float)i;
    __syncthreads();
    int o = offs[threadIdx.x & 31];
    float acc = 0;
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 16; ++u) {
            float v;
            unsigned a = (unsigned)__cvta_generic_to_shared(&sm[(o + u * 128) & 8191]);
            asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a));
            acc += v;
        }
    }
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

int main() {
    const char *names[] = {"all lanes same 16B", "2 distinct (half-warps)", "4 distinct (one per quarter-warp)",
                           "8 distinct in each quarter, same across quarters", "16 distinct x2 (lane&15)",
                           "32 distinct contiguous", "8 lanes/row x 4 rows (gather pattern)",
                           "4 distinct interleaved (lane&3)"};
    int pat[8][32];
    for (int l = 0; l < 32; ++l) {
        pat[0][l] = 0;
        pat[1][l] = (l >> 4) * 64;
        pat[2][l] = (l >> 3) * 8;
        pat[3][l] = (l & 7) * 4;
        pat[4][l] = (l & 15) * 4;
        pat[5][l] = l * 4;
        pat[6][l] = (l >> 3) * 1024 + (l & 7) * 4;
        pat[7][l] = (l & 3) * 8;
    }
    int *d_offs; float *d_out; long long *d_cyc;
    cudaMalloc(&d_offs, 32 * sizeof(int)); cudaMalloc(&d_out, 1024 * 4 * sizeof(float)); cudaMalloc(&d_cyc, 64 * sizeof(long long));
    const int iters = 2000;
    for (int warps : {1, 4, 16}) {
        printf("--- %d warps per CTA, 1 CTA ---\n", warps);
        for (int p = 0; p < 8; ++p) {
            cudaMemcpy(d_offs, pat[p], sizeof(pat[p]), cudaMemcpyHostToDevice);
            long long c128, c32;
            k128<<<1, warps * 32>>>(d_offs, d_out, d_cyc, iters); cudaDeviceSynchronize();
            cudaMemcpy(&c128, d_cyc, sizeof(long long), cudaMemcpyDeviceToHost);
            k32<<<1, warps * 32>>>(d_offs, d_out, d_cyc, iters); cudaDeviceSynchronize();
            cudaMemcpy(&c32, d_cyc, sizeof(long long), cudaMemcpyDeviceToHost);
            double n = (double)iters * 16 * warps;
            printf("%-52s LDS.128: %.2f cyc/warp-instr   LDS.32: %.2f cyc/warp-instr\n", names[p], c128 / n, c32 / n);
        }
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
