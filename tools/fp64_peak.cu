// Measures the fp64 peaks the ALS Gram kernel is bounded by (not in MEASURED_PEAKS.json):
// DFMA (vector pipe) and DMMA (mma.sync f64 tensor path) throughput on the whole chip.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/fp64_peak tools/fp64_peak.cu
#include <cuda_runtime.h>
#include <cstdio>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

__global__ void __launch_bounds__(256) k_dfma(double* out, int iters, double b, double c) {
    double a[16];
#pragma unroll
    for (int i = 0; i < 16; i++) a[i] = threadIdx.x * 1e-3 + i;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 16; i++) a[i] = fma(a[i], b, c);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 16; i++) s += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void __launch_bounds__(256) k_ffma(float* out, int iters, float b, float c) {
    float a[16];
#pragma unroll
    for (int i = 0; i < 16; i++) a[i] = threadIdx.x * 1e-3f + i;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 16; i++) a[i] = fmaf(a[i], b, c);
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < 16; i++) s += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// mma.sync.m8n8k4 f64: 8 independent accumulator tiles per warp.
__global__ void __launch_bounds__(256) k_dmma884(double* out, int iters, double av, double bv) {
    double c[8][2];
#pragma unroll
    for (int i = 0; i < 8; i++) { c[i][0] = i; c[i][1] = threadIdx.x; }
    double a = av + threadIdx.x * 1e-6, b = bv;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) s += c[i][0] + c[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// mma.sync.m16n8k16 f64 (sm_90+): A 8 regs, B 4 regs, C 4 regs per lane.
__global__ void __launch_bounds__(256) k_dmma16816(double* out, int iters, double av, double bv) {
    double c[4][4];
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) c[i][j] = i + j + threadIdx.x;
    double a[8], b[4];
#pragma unroll
    for (int i = 0; i < 8; i++) a[i] = av + i * 1e-6;
#pragma unroll
    for (int i = 0; i < 4; i++) b[i] = bv + i * 1e-6;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 4; i++)
            asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, "
                         "{%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};"
                         : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3])
                         : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]),
                           "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) s += c[i][j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F>
static float time_ms(F launch) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    launch();  // warm-up
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < 5; r++) {
        cudaEventRecord(e0);
        launch();
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    return best;
}

int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    const int sms = p.multiProcessorCount;
    const int blocks = sms * 8, threads = 256;
    double* out; CK(cudaMalloc(&out, sizeof(double) * blocks * threads));
    const int iters = 20000;
    float ms = time_ms([&] { k_dfma<<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); });
    double flop = 2.0 * 16 * iters * (double)blocks * threads;
    printf("{\"kernel\": \"dfma\", \"sms\": %d, \"ms\": %.3f, \"tflops\": %.2f}\n", sms, ms, flop / ms * 1e-9);
    ms = time_ms([&] { k_ffma<<<blocks, threads>>>((float*)out, iters, 1.0000001f, 1e-9f); });
    printf("{\"kernel\": \"ffma\", \"sms\": %d, \"ms\": %.3f, \"tflops\": %.2f}\n", sms, ms, flop / ms * 1e-9);
    const int miters = 4000;
    ms = time_ms([&] { k_dmma884<<<blocks, threads>>>(out, miters, 1.0, 1e-9); });
    flop = 2.0 * 8 * 8 * 4 * 8 * miters * (double)blocks * (threads / 32);
    printf("{\"kernel\": \"dmma_m8n8k4\", \"ms\": %.3f, \"tflops\": %.2f}\n", ms, flop / ms * 1e-9);
    ms = time_ms([&] { k_dmma16816<<<blocks, threads>>>(out, miters, 1.0, 1e-9); });
    flop = 2.0 * 16 * 8 * 16 * 4 * miters * (double)blocks * (threads / 32);
    printf("{\"kernel\": \"dmma_m16n8k16\", \"ms\": %.3f, \"tflops\": %.2f}\n", ms, flop / ms * 1e-9);
    CK(cudaGetLastError());
    CK(cudaDeviceSynchronize());
    return 0;
}
