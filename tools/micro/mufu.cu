// MUFU throughput microbenchmark (B200): ops per clock per SM for ex2.approx, rcp.approx and the 5:2 mix of the LSTM cell.
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ float ex2a(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float rcpa(float x) { float y; asm volatile("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
template <int MODE> __global__ void k(float* out, int iters) {
    float v[8];
    for (int i = 0; i < 8; ++i) v[i] = 0.001f * (threadIdx.x + i);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (MODE == 0) v[i] = ex2a(v[i]);
            else if (MODE == 1) v[i] = rcpa(v[i] + 1.5f);
            else { // 5 ex2 + 2 rcp
                v[i] = ex2a(v[i]); v[i] = ex2a(-v[i]); v[i] = rcpa(v[i] + 1.5f); v[i] = ex2a(v[i]); v[i] = ex2a(-v[i]); v[i] = rcpa(v[i] + 1.5f); v[i] = ex2a(-v[i]);
            }
        }
    }
    float s = 0; for (int i = 0; i < 8; ++i) s += v[i];
    if (s == 123.456f) out[0] = s;
}
int main() {
    float* d; cudaMalloc(&d, 4);
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    for (int mode = 0; mode < 3; ++mode) for (int warps = 4; warps <= 32; warps *= 2) {
        const int iters = 20000, per = mode == 2 ? 7 : 1;
        cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
        auto launch = [&]() { if (mode == 0) k<0><<<p.multiProcessorCount, warps * 32>>>(d, iters); else if (mode == 1) k<1><<<p.multiProcessorCount, warps * 32>>>(d, iters); else k<2><<<p.multiProcessorCount, warps * 32>>>(d, iters); };
        launch(); cudaDeviceSynchronize();
        cudaEventRecord(a); launch(); cudaEventRecord(b); cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b);
        double ops = (double)p.multiProcessorCount * warps * 32 * 8.0 * iters * per;
        printf("mode %d warps/SM %2d: %.2f ms  %.1f Gops/s  %.2f ops/clk/SM at %d MHz nominal\n", mode, warps, ms, ops / ms / 1e6, ops / (ms * 1e-3) / p.multiProcessorCount / (clk * 1e3), clk / 1000);
    }
    return 0;
}
