// Micro-benchmark: FFMA2 throughput in the shape the wavelet kernels use it -- acc = fma2(window value (register),
// tap (kernel parameter -> uniform register / constant bank), acc) -- as a function of warps per SM and independent
// accumulation chains per thread.  Occupancy is limited through dynamic shared memory.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fma_occ fma_occ.cu
#include <cstdio>
#include <cuda_runtime.h>

struct Taps { float2 lo[16], hi[16]; };

template <int CH>
__global__ void __launch_bounds__(128) k(float2 *out, const Taps tp, int iters)
{
    extern __shared__ float2 sm[];
    float2 w[16 + CH];
#pragma unroll
    for (int i = 0; i < 16 + CH; ++i) w[i] = make_float2(threadIdx.x * 0.001f + i, i * 0.5f);
    float2 acc[CH], ach[CH];
#pragma unroll
    for (int i = 0; i < CH; ++i) { acc[i] = make_float2(0.f, 0.f); ach[i] = make_float2(0.f, 0.f); }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int kk = 0; kk < 16; ++kk)
#pragma unroll
            for (int r = 0; r < CH; ++r) {
                acc[r] = __ffma2_rn(w[r + kk], tp.lo[kk], acc[r]);
                ach[r] = __ffma2_rn(w[r + kk], tp.hi[kk], ach[r]);
            }
#pragma unroll
        for (int i = 0; i < CH; ++i) w[i] = acc[i];      // keeps the window live and the loop honest
    }
    float2 r = make_float2(0.f, 0.f);
#pragma unroll
    for (int i = 0; i < CH; ++i) { r.x += acc[i].x + ach[i].x; r.y += acc[i].y + ach[i].y; }
    if (threadIdx.x == 9999) sm[0] = r;
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

template <int CH>
static void run(float2 *out, int ctas_per_sm)
{
    Taps tp;
    for (int i = 0; i < 16; ++i) { tp.lo[i] = make_float2(0.01f * i, 0.01f * i); tp.hi[i] = make_float2(-0.02f * i, -0.02f * i); }
    const int smem = (227 * 1024) / ctas_per_sm - 1024 - 2048;
    cudaFuncSetAttribute(k<CH>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    const int blocks = 148 * ctas_per_sm * 8, iters = 2048 / CH;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e9f;
    for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(e0);
        k<CH><<<blocks, 128, smem>>>(out, tp, iters);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    const double fma = (double)blocks * 128 * iters * 16 * CH * 2 * 2;   // scalar FMAs
    printf("chains/thread %2d (x2 lo,hi)  warps/SM %2d : %.3f ms  %.2f TFMA/s  err=%s\n", 2 * CH, 4 * ctas_per_sm, best,
           fma / best * 1e-9, cudaGetErrorString(cudaGetLastError()));
}

int main()
{
    float2 *out;
    cudaMalloc(&out, sizeof(float2) * 148 * 16 * 8 * 128);
    for (int c : {1, 2, 3, 4, 6, 8, 12, 16}) {
        run<1>(out, c);
        run<2>(out, c);
        run<4>(out, c);
    }
    return 0;
}
