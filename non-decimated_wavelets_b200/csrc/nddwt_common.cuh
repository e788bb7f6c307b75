// nddwt_common.cuh -- shared device/host helpers for libnddwt_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/nddwt_b200.h"

#define NDDWT_MAXL 20   // db10

namespace nddwt {

// ---- element traits: T = storage element, R = real scalar the taps are held in ----------
template <typename T> struct Elem;
template <> struct Elem<float>   { using R = float;  static constexpr bool cplx = false; };
template <> struct Elem<double>  { using R = double; static constexpr bool cplx = false; };
template <> struct Elem<float2>  { using R = float;  static constexpr bool cplx = true;  };
template <> struct Elem<double2> { using R = double; static constexpr bool cplx = true;  };

__device__ __forceinline__ float   zero_of(float)   { return 0.f; }
__device__ __forceinline__ double  zero_of(double)  { return 0.0; }
__device__ __forceinline__ float2  zero_of(float2)  { return make_float2(0.f, 0.f); }
__device__ __forceinline__ double2 zero_of(double2) { return make_double2(0.0, 0.0); }

// acc += g * v  (real tap times real/complex sample)
__device__ __forceinline__ void mac(float &acc, float g, float v)    { acc = fmaf(g, v, acc); }
__device__ __forceinline__ void mac(double &acc, double g, double v) { acc = fma(g, v, acc); }
__device__ __forceinline__ void mac(float2 &acc, float g, float2 v)  { acc.x = fmaf(g, v.x, acc.x); acc.y = fmaf(g, v.y, acc.y); }
__device__ __forceinline__ void mac(double2 &acc, double g, double2 v) { acc.x = fma(g, v.x, acc.x); acc.y = fma(g, v.y, acc.y); }

__device__ __forceinline__ float   add(float a, float b)     { return a + b; }
__device__ __forceinline__ double  add(double a, double b)   { return a + b; }
__device__ __forceinline__ float2  add(float2 a, float2 b)   { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ double2 add(double2 a, double2 b) { return make_double2(a.x + b.x, a.y + b.y); }

// ---- soft threshold (fused coefficient shrink): real  sign(v) max(|v| - t, 0);  complex  v max(0, 1 - t/|v|).
// t = 0 is the identity (scale exactly 1).
// Branch-free (the analysis tile kernel sits at its register cap; control flow in its store loop spills).
__device__ __forceinline__ float shrink1(float v, float t) { return copysignf(fmaxf(fabsf(v) - t, 0.f), v); }
__device__ __forceinline__ double shrink1(double v, double t) { return copysign(fmax(fabs(v) - t, 0.0), v); }
__device__ __forceinline__ float2 shrink1(float2 v, float t)
{
    const float m2 = fmaxf(v.x * v.x + v.y * v.y, 1e-37f);
    const float sc = fmaxf(1.f - t * rsqrtf(m2), 0.f);
    return make_float2(v.x * sc, v.y * sc);
}
__device__ __forceinline__ double2 shrink1(double2 v, double t)
{
    const double m2 = fmax(v.x * v.x + v.y * v.y, 1e-300);
    const double sc = fmax(1.0 - t * rsqrt(m2), 0.0);
    return make_double2(v.x * sc, v.y * sc);
}

// ---- taps for one dimension, passed by value in kernel parameters (constant bank) --------
template <typename R>
struct DimTaps {
    R lo[NDDWT_MAXL];
    R hi[NDDWT_MAXL];
};

template <typename R>
struct AllTaps {
    DimTaps<R> d[NDDWT_MAX_DIMS];
};

// positive modulo for possibly negative m and offsets larger than n
__device__ __forceinline__ int64_t wrap(int64_t m, int64_t n)
{
    if (m >= n) { m -= n; if (m >= n) m %= n; }
    else if (m < 0) { m += n; if (m < 0) { m %= n; if (m < 0) m += n; } }
    return m;
}

}  // namespace nddwt
