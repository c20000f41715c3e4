// Microbenchmark: issue throughput of FFMA (3-reg), fma.rn.f32x2 (FFMA2), FMNMX and an FFMA2+FMNMX mix on sm_100a.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ffma2_bench ffma2_bench.cu
#include <cstdio>
#include <cuda_runtime.h>
#define ITERS 4096
__device__ __forceinline__ void fma2(float2& d, const float2& a, const float2& b) {
    unsigned long long dd = *reinterpret_cast<unsigned long long*>(&d);
    const unsigned long long aa = *reinterpret_cast<const unsigned long long*>(&a);
    const unsigned long long bb = *reinterpret_cast<const unsigned long long*>(&b);
    asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(dd) : "l"(aa), "l"(bb));
    d = *reinterpret_cast<float2*>(&dd);
}
template <int MODE> __global__ void __launch_bounds__(256) k(float* out, float s) {
    float a[16]; float2 p[8];
    for (int i = 0; i < 16; ++i) a[i] = threadIdx.x * 0.001f + i;
    for (int i = 0; i < 8; ++i) p[i] = make_float2(threadIdx.x * 0.001f + i, i);
    const float w0 = s, w1 = s * 0.5f; const float2 w = make_float2(s, s * 0.5f), x = make_float2(0.25f, 0.5f);
    for (int it = 0; it < ITERS; ++it) {
        if (MODE == 0) {
#pragma unroll
            for (int i = 0; i < 16; ++i) a[i] = fmaf(a[i], w0, w1);
        } else if (MODE == 1) {
#pragma unroll
            for (int i = 0; i < 8; ++i) fma2(p[i], x, w);
        } else if (MODE == 2) {
#pragma unroll
            for (int i = 0; i < 16; ++i) a[i] = fmaxf(a[i], w0 + i);
        } else if (MODE == 3) {  // 8 FFMA2 + 8 FMNMX
#pragma unroll
            for (int i = 0; i < 8; ++i) fma2(p[i], x, w);
#pragma unroll
            for (int i = 0; i < 8; ++i) a[i] = fmaxf(a[i], w0 + i);
        } else if (MODE == 4) {  // 8 FFMA + 8 FMNMX
#pragma unroll
            for (int i = 0; i < 8; ++i) a[i] = fmaf(a[i], w0, w1);
#pragma unroll
            for (int i = 8; i < 16; ++i) a[i] = fmaxf(a[i], w0 + i);
        }
    }
    float r = 0; for (int i = 0; i < 16; ++i) r += a[i]; for (int i = 0; i < 8; ++i) r += p[i].x + p[i].y;
    if (r == 123.456f) out[0] = r;
}
template <int MODE> void run(const char* name, double inst_per_iter, double flop_per_iter) {
    float* d; cudaMalloc(&d, 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int blocks = 148 * 8;
    k<MODE><<<blocks, 256>>>(d, 1.0001f); cudaDeviceSynchronize();
    cudaEventRecord(e0); k<MODE><<<blocks, 256>>>(d, 1.0001f); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double threads = (double)blocks * 256;
    const double winst = threads / 32 * ITERS * inst_per_iter;
    printf("%-22s %8.3f ms  warp-inst/clk/SMSP(@1.965GHz) %.3f   TFLOP/s %.1f\n", name, ms,
           winst / (ms * 1e-3) / (148 * 4 * 1.965e9), threads * ITERS * flop_per_iter / (ms * 1e-3) / 1e12);
}
int main() {
    run<0>("FFMA x16", 16, 32);
    run<1>("FFMA2 x8", 8, 32);
    run<2>("FMNMX x16", 16, 0);
    run<3>("FFMA2 x8 + FMNMX x8", 16, 32);
    run<4>("FFMA x8 + FMNMX x8", 16, 16);
    return 0;
}
