// Micro-benchmark: FFMA vs FFMA2 (fma.rn.f32x2) issue throughput on sm_100a.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fma_peak fma_peak.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void __launch_bounds__(256) k(float2 *out, float2 s, int iters)
{
    float2 a[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = make_float2(threadIdx.x * 0.001f + i, i * 0.5f);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 8; ++r) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                if (MODE == 0) { a[i].x = fmaf(a[i].x, s.x, s.y); a[i].y = fmaf(a[i].y, s.x, s.y); }
                else a[i] = __ffma2_rn(a[i], s, make_float2(s.y, s.y));
            }
        }
    }
    float2 r = make_float2(0.f, 0.f);
#pragma unroll
    for (int i = 0; i < 8; ++i) { r.x += a[i].x; r.y += a[i].y; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

int main()
{
    float2 *out;
    const int blocks = 148 * 8, iters = 4096;
    cudaMalloc(&out, sizeof(float2) * blocks * 256);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int mode = 0; mode < 2; ++mode) {
        for (int rep = 0; rep < 3; ++rep) {
            cudaEventRecord(e0);
            if (mode == 0) k<0><<<blocks, 256>>>(out, make_float2(0.999f, 0.001f), iters);
            else k<1><<<blocks, 256>>>(out, make_float2(0.999f, 0.001f), iters);
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            double fma = (double)blocks * 256 * iters * 8 * 8 * 2;   // scalar FMAs
            printf("mode %s rep %d: %.3f ms  %.2f TFMA/s (%.2f TFLOP/s)\n", mode ? "FFMA2" : "FFMA ", rep, ms,
                   fma / ms * 1e-9, 2 * fma / ms * 1e-9);
        }
    }
    return 0;
}
