// DMMA.8x8x4 throughput against resident warps per SM sub-partition (k_gram runs 4: 254 registers).
// 28 independent accumulator tiles per warp, 7 distinct operand fragments -- k_gram's inner step.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/dmma_occupancy tools/dmma_occupancy.cu
#include <cstdio>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s\n", cudaGetErrorString(e)); return 1; } } while (0)

__global__ void k(double* out, int iters, double av) {
    double c[28][2], f[7];
#pragma unroll
    for (int i = 0; i < 28; i++) { c[i][0] = i; c[i][1] = threadIdx.x; }
#pragma unroll
    for (int i = 0; i < 7; i++) f[i] = av + i * 1e-6 + threadIdx.x * 1e-9;
    for (int it = 0; it < iters; it++) {
        int t = 0;
#pragma unroll
        for (int i = 0; i < 7; i++)
#pragma unroll
            for (int j = 0; j <= i; j++, t++)
                asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                             : "+d"(c[t][0]), "+d"(c[t][1]) : "d"(f[i]), "d"(f[j]));
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 28; i++) s += c[i][0] + c[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main() {
    cudaDeviceProp pr; CK(cudaGetDeviceProperties(&pr, 0));
    const int sms = pr.multiProcessorCount;
    double* out; CK(cudaMalloc(&out, sizeof(double) * sms * 1024));
    const int iters = 2000;
    for (int warps_per_sm : {4, 8, 16, 32}) {
        const int threads = warps_per_sm * 32;   // one CTA per SM
        k<<<sms, threads>>>(out, 10, 1.0);
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        cudaEventRecord(e0);
        k<<<sms, threads>>>(out, iters, 1.0);
        cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        const double flop = 2.0 * 8 * 8 * 4 * 28 * iters * (double)sms * warps_per_sm;
        printf("{\"warps_per_smsp\": %d, \"ms\": %.3f, \"tflops\": %.2f}\n", warps_per_sm / 4, ms, flop / ms * 1e-9);
    }
    return 0;
}
