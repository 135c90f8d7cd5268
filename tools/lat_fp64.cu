// Dependent-chain latencies of the instructions on k_gram's Cholesky critical path (one warp).
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o lat_fp64 lat_fp64.cu ; run: ./lat_fp64
#include <cstdio>
#include <cuda_runtime.h>

#define N 512
__device__ __forceinline__ double shfl_double(double v, int src) {
    int lo = __double2loint(v), hi = __double2hiint(v);
    lo = __shfl_sync(0xffffffffu, lo, src);
    hi = __shfl_sync(0xffffffffu, hi, src);
    return __hiloint2double(hi, lo);
}
__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__device__ __forceinline__ double fast_rsqrt(double d) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(d));
    const double e = fma(-d, y * y, 1.0);
    return fma(fma(e, 0.375, 0.5), y * e, y);
}
__global__ void k(double* out, long long* cyc, double seed) {
    const int lane = threadIdx.x;
    double x = seed + lane * 1e-9, y = 1.0000001, c0 = 0, c1 = 0;
    long long t[16];
    int i = 0;
    t[i++] = clock64();
    asm volatile("" ::: "memory");
#pragma unroll 16
    for (int j = 0; j < N; j++) asm volatile("fma.rn.f64 %0, %0, %1, %1;" : "+d"(x) : "d"(y));
    t[i++] = clock64();
#pragma unroll 16
    for (int j = 0; j < N; j++) asm volatile("mul.rn.f64 %0, %0, %1;" : "+d"(x) : "d"(y));
    t[i++] = clock64();
#pragma unroll 16
    for (int j = 0; j < N; j++) asm volatile("add.rn.f64 %0, %0, %1;" : "+d"(x) : "d"(y));
    t[i++] = clock64();
#pragma unroll 16
    for (int j = 0; j < N; j++) { double r; asm volatile("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x)); x = r; }
    t[i++] = clock64();
    x = 2.0 + lane;
#pragma unroll 16
    for (int j = 0; j < N; j++) x = fast_rsqrt(x) + 2.0;
    t[i++] = clock64();
#pragma unroll 16
    for (int j = 0; j < N; j++) x = shfl_double(x, (lane + 1) & 31);
    t[i++] = clock64();
#pragma unroll 16
    for (int j = 0; j < N; j++) dmma(c0, c1, y, y);          // accumulator chain
    t[i++] = clock64();
#pragma unroll 16
    for (int j = 0; j < N; j++) { double d0 = 0, d1 = 0; dmma(d0, d1, x, y); x = d0 + 1e-30; }   // result -> A operand (+DADD)
    t[i++] = clock64();
#pragma unroll 16
    for (int j = 0; j < N; j++) x = x > 0.5 ? x : y;          // select
    t[i++] = clock64();
    double z = x;
#pragma unroll 16
    for (int j = 0; j < N; j++) z = 1.0 / z + 1.0;            // IEEE division
    t[i++] = clock64();
#pragma unroll 16
    for (int j = 0; j < N; j++) z = sqrt(z) + 1.0;            // IEEE sqrt
    t[i++] = clock64();
    out[lane] = x + c0 + c1 + z;
    if (lane == 0) for (int j = 0; j + 1 < i; j++) cyc[j] = t[j + 1] - t[j];
}
int main() {
    double* out; long long* cyc;
    cudaMalloc(&out, 32 * 8); cudaMalloc(&cyc, 16 * 8);
    for (int r = 0; r < 2; r++) k<<<1, 32>>>(out, cyc, 1.0);
    long long h[16];
    cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    const char* names[] = {"DFMA", "DMUL", "DADD", "MUFU.RSQ64H (rsqrt.approx.f64)", "fast_rsqrt + DADD", "shfl_double (2 SHFL)",
                           "DMMA accumulator chain", "DMMA -> DADD -> A operand", "select (FSEL pair)", "1/x + DADD", "sqrt + DADD"};
    for (int j = 0; j < 11; j++) printf("%-36s %.1f cycles per dependent step\n", names[j], double(h[j]) / N);
    return cudaGetLastError() != cudaSuccess;
}
