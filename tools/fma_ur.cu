// Micro-benchmark: does an FFMA2 whose tap operand lives in a UNIFORM register cost more than one whose tap lives in
// a regular register, and does the order (same tap for G consecutive FFMA2 vs a new tap every instruction) matter?
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fma_ur fma_ur.cu ; inspect with cuobjdump -sass
#include <cstdio>
#include <cuda_runtime.h>

struct Taps { float2 t[32]; };

// MODE 0: taps from the kernel parameters (uniform registers); MODE 1: taps loaded per thread from global memory
// (regular registers).  ORDER 0: chain-major inside a tap pair (lo, hi, lo, hi ...: a new tap every instruction);
// ORDER 1: tap-major (all CH chains with tap lo[k], then all with hi[k]).
template <int MODE, int ORDER, int CH>
__global__ void __launch_bounds__(128) k(float2 *out, const Taps tp, const float2 *gt, int iters)
{
    float2 t[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) t[i] = MODE ? gt[threadIdx.x * 32 + i] : tp.t[i];
    float2 w[16 + CH];
#pragma unroll
    for (int i = 0; i < 16 + CH; ++i) w[i] = make_float2(threadIdx.x * 0.001f + i, i * 0.5f);
    float2 acc[CH], ach[CH];
#pragma unroll
    for (int i = 0; i < CH; ++i) { acc[i] = make_float2(0.f, 0.f); ach[i] = make_float2(0.f, 0.f); }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int kk = 0; kk < 16; ++kk) {
            if (ORDER == 0) {
#pragma unroll
                for (int r = 0; r < CH; ++r) {
                    acc[r] = __ffma2_rn(w[r + kk], t[kk], acc[r]);
                    ach[r] = __ffma2_rn(w[r + kk], t[16 + kk], ach[r]);
                }
            } else {
#pragma unroll
                for (int r = 0; r < CH; ++r) acc[r] = __ffma2_rn(w[r + kk], t[kk], acc[r]);
#pragma unroll
                for (int r = 0; r < CH; ++r) ach[r] = __ffma2_rn(w[r + kk], t[16 + kk], ach[r]);
            }
        }
#pragma unroll
        for (int i = 0; i < CH; ++i) w[i] = acc[i];
    }
    float2 r = make_float2(0.f, 0.f);
#pragma unroll
    for (int i = 0; i < CH; ++i) { r.x += acc[i].x + ach[i].x; r.y += acc[i].y + ach[i].y; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

template <int MODE, int ORDER, int CH>
static void run(float2 *out, const float2 *gt)
{
    Taps tp;
    for (int i = 0; i < 32; ++i) tp.t[i] = make_float2(0.01f * i - 0.1f, 0.01f * i - 0.1f);
    const int blocks = 148 * 4 * 8, iters = 2048 / CH;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e9f;
    for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(e0);
        k<MODE, ORDER, CH><<<blocks, 128>>>(out, tp, gt, iters);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    const double fma = (double)blocks * 128 * iters * 16 * CH * 2 * 2;
    printf("taps in %s registers, %s, %d chains: %.3f ms  %.2f TFMA/s  (%s)\n", MODE ? "regular" : "uniform",
           ORDER ? "tap-major " : "chain-major", 2 * CH, best, fma / best * 1e-9, cudaGetErrorString(cudaGetLastError()));
}

int main()
{
    float2 *out, *gt;
    cudaMalloc(&out, sizeof(float2) * 148 * 4 * 8 * 128);
    cudaMalloc(&gt, sizeof(float2) * 128 * 32);
    cudaMemset(gt, 0, sizeof(float2) * 128 * 32);
    run<0, 0, 1>(out, gt); run<0, 0, 2>(out, gt); run<0, 0, 4>(out, gt);
    run<0, 1, 2>(out, gt); run<0, 1, 4>(out, gt);
    run<1, 0, 1>(out, gt); run<1, 0, 2>(out, gt); run<1, 0, 4>(out, gt);
    run<1, 1, 2>(out, gt); run<1, 1, 4>(out, gt);
    return 0;
}
