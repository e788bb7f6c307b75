// nddwt_fused1d.cu -- 1-D multi-level CASCADE kernels: a CTA loads one tile of a signal (plus the
// halo all J levels need, periodic wrap folded into the load) into shared memory once, runs the J
// analysis levels there (approximation ping-pongs between two shared buffers), and writes the J
// detail bands and the final approximation once -- 1 read + (J+1) writes per sample, the compulsory
// traffic, instead of the 3 J passes of a level-by-level scheme.  Synthesis mirrors it.
// Replaces dec / rec / level_1_dec / level_1_rec of Functions/nd_dwt_1D.m:136-318 and
// nd_dwt_dec / nd_dwt_rec (mex/nddwt.c:189-292) for num_dims == 1, for one signal or a batch of
// signals (batch extension; BASELINE configs[1]: 4096 signals x 65536 samples, db8, 6 levels).
#include "nddwt_plan.h"

namespace nddwt {

__device__ __forceinline__ int64_t wrap1(int64_t m, int64_t n)
{
    m %= n;
    return m < 0 ? m + n : m;
}

// complex single: taps duplicated (t, t) so that one FFMA2 updates (re, im)
template <typename T> struct Tap1Of { using type = typename Elem<T>::R; };
template <> struct Tap1Of<float2> { using type = float2; };
__device__ __forceinline__ void mac1(float &a, float g, float v) { a = fmaf(g, v, a); }
__device__ __forceinline__ void mac1(double &a, double g, double v) { a = fma(g, v, a); }
__device__ __forceinline__ void mac1(float2 &a, float2 g, float2 v) { a = __ffma2_rn(g, v, a); }
__device__ __forceinline__ void mac1(double2 &a, double g, double2 v) { a.x = fma(g, v.x, a.x); a.y = fma(g, v.y, a.y); }
static inline float mk1(float, double v) { return (float)v; }
static inline double mk1(double, double v) { return v; }
static inline float2 mk1(float2, double v) { return make_float2((float)v, (float)v); }

template <typename T, int L>
struct Taps1 {
    typename Tap1Of<T>::type lo[L];
    typename Tap1Of<T>::type hi[L];
    typename Elem<T>::R thr[NDDWT_MAX_LEVELS];   // fused coefficient shrink: soft threshold of d_j (0 = keep), analysis only
};

// contiguous periodic copy global -> shared: one wrap computation per thread, then increments
template <typename T, int NT>
__device__ __forceinline__ void load_wrapped(T *dst, const T *src, int64_t start, int count, int64_t n1, int tid)
{
    int64_t g = wrap1(start + tid, n1);
    const int64_t step = NT % n1;
    for (int i = tid; i < count; i += NT) {
        dst[i] = __ldg(src + g);
        g += step;
        if (g >= n1) g -= n1;
    }
}


template <typename T, int VEC>
__device__ __forceinline__ void ld16(const T *p, T *dst)
{
    union { uint4 u; T t[VEC]; } cv;
    cv.u = *reinterpret_cast<const uint4 *>(p);
#pragma unroll
    for (int i = 0; i < VEC; ++i) dst[i] = cv.t[i];
}
template <typename T, int VEC>
__device__ __forceinline__ void st16(T *p, const T *src)
{
    union { uint4 u; T t[VEC]; } cv;
#pragma unroll
    for (int i = 0; i < VEC; ++i) cv.t[i] = src[i];
    *reinterpret_cast<uint4 *>(p) = cv.u;
}

// coeffs: [n1][batch][J+1] column-major => band s of signal b starts at (s * batch + b) * n1
template <typename T, int L, int TILE, int NT, int CPT>
__global__ void __launch_bounds__(NT)
k_dec1_cascade(const T *__restrict__ x, T *__restrict__ coeffs, int64_t n1, int64_t batch, int J,
               const Taps1<T, L> tp)
{
    constexpr int HLs = L / 2 - 1;                 // analysis reads n-(L/2-1) .. n+L/2
    extern __shared__ __align__(16) unsigned char smem1_raw[];
    const int W0 = (TILE + J * (L - 1) + 4 * (16 / (int)sizeof(T)) + 3) & ~3;   // slack for whole-chunk accesses
    T *buf0 = reinterpret_cast<T *>(smem1_raw);
    T *buf1 = buf0 + W0;
    const int tid = threadIdx.x;
    const int64_t b = blockIdx.y;
    const int64_t t0 = (int64_t)blockIdx.x * TILE;
    const T *xs = x + b * n1;
    // level-0 buffer: origin global t0 - J*(L/2-1)
    load_wrapped<T, NT>(buf0, xs, t0 - (int64_t)J * HLs, TILE + J * (L - 1), n1, tid);
    __syncthreads();
    T *src = buf0, *dst = buf1;
    for (int j = 1; j <= J; ++j) {
        const int Wj = TILE + (J - j) * (L - 1);       // valid outputs of this level
        const int c0 = (J - j) * HLs;                   // buffer index of global t0 at this level
        T *band = coeffs + ((int64_t)(J - j + 1) * batch + b) * n1;     // detail d_j lives in slot J-j+1
        // each thread produces CPT 16-byte chunks (R consecutive outputs) from NCH chunk loads; the tap loop is
        // outermost so that the 2 R accumulations are independent FFMA2 chains (two chains per thread stall on
        // the FFMA2 latency: profiles/r01_rows_kernel.md)
        constexpr int VEC = 16 / (int)sizeof(T), R = CPT * VEC, NCH = (R + L - 1 + VEC - 1) / VEC;
        for (int o = tid * R; o < Wj; o += NT * R) {
            T v[NCH * VEC];
#pragma unroll
            for (int q = 0; q < NCH; ++q) ld16<T, VEC>(src + o + q * VEC, v + q * VEC);
            T lo[R], hi[R];
#pragma unroll
            for (int r = 0; r < R; ++r) { lo[r] = zero_of(T()); hi[r] = zero_of(T()); }
#pragma unroll
            for (int k = 0; k < L; ++k)
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    mac1(lo[r], tp.lo[k], v[r + (L - 1) - k]);
                    mac1(hi[r], tp.hi[k], v[r + (L - 1) - k]);
                }
#pragma unroll
            for (int c = 0; c < CPT; ++c) st16<T, VEC>(dst + o + c * VEC, lo + c * VEC);
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const int64_t g = t0 + (o + r - c0);
                if (o + r >= c0 && o + r < c0 + TILE && g < n1) band[g] = shrink1(hi[r], tp.thr[j - 1]);
            }
        }
        __syncthreads();
        T *t = src; src = dst; dst = t;
    }
    T *approx = coeffs + b * n1;                        // slot 0 = a_J
    for (int o = tid; o < TILE; o += NT)
        if (t0 + o < n1) approx[t0 + o] = src[o];
}

template <typename T, int L, int TILE, int NT, int CPT>
__global__ void __launch_bounds__(NT)
k_rec1_cascade(const T *__restrict__ coeffs, T *__restrict__ x, int64_t n1, int64_t batch, int J,
               const Taps1<T, L> tp)
{
    constexpr int HLr = L / 2;                     // synthesis reads n-L/2 .. n+L/2-1
    extern __shared__ __align__(16) unsigned char smem1_raw[];
    const int WJ = (TILE + J * (L - 1) + 4 * (16 / (int)sizeof(T)) + 3) & ~3;
    T *a0 = reinterpret_cast<T *>(smem1_raw);
    T *a1 = a0 + WJ;
    T *dd = a1 + WJ;
    const int tid = threadIdx.x;
    const int64_t b = blockIdx.y;
    const int64_t t0 = (int64_t)blockIdx.x * TILE;
    // a_J on [t0 - J*L/2, t0 + TILE + J*(L/2-1))
    {
        const T *aJ = coeffs + b * n1;
        load_wrapped<T, NT>(a0, aJ, t0 - (int64_t)J * HLr, TILE + J * (L - 1), n1, tid);
    }
    T *src = a0, *dst = a1;
    for (int j = J; j >= 1; --j) {
        const int Wj = TILE + j * (L - 1);             // extent of a_j / d_j needed at this level
        const T *dj = coeffs + ((int64_t)(J - j + 1) * batch + b) * n1;
        load_wrapped<T, NT>(dd, dj, t0 - (int64_t)j * HLr, Wj, n1, tid);
        __syncthreads();
        const int Wo = Wj - (L - 1);                   // outputs a_{j-1}
        constexpr int VEC = 16 / (int)sizeof(T), R = CPT * VEC, NCH = (R + L - 1 + VEC - 1) / VEC;
        for (int o = tid * R; o < Wo; o += NT * R) {
            T va[NCH * VEC], vd[NCH * VEC];
#pragma unroll
            for (int q = 0; q < NCH; ++q) {
                ld16<T, VEC>(src + o + q * VEC, va + q * VEC);
                ld16<T, VEC>(dd + o + q * VEC, vd + q * VEC);
            }
            // approximation and detail parts in separate accumulators, tap loop outermost: 2 R independent chains
            T acc[R], acd[R];
#pragma unroll
            for (int r = 0; r < R; ++r) { acc[r] = zero_of(T()); acd[r] = zero_of(T()); }
#pragma unroll
            for (int k = 0; k < L; ++k)
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    mac1(acc[r], tp.lo[k], va[r + k]);
                    mac1(acd[r], tp.hi[k], vd[r + k]);
                }
#pragma unroll
            for (int r = 0; r < R; ++r) acc[r] = add(acc[r], acd[r]);
#pragma unroll
            for (int c = 0; c < CPT; ++c) st16<T, VEC>(dst + o + c * VEC, acc + c * VEC);
        }
        __syncthreads();
        T *t = src; src = dst; dst = t;
    }
    T *xs = x + b * n1;
    for (int o = tid; o < TILE; o += NT)
        if (t0 + o < n1) xs[t0 + o] = src[o];
}

template <typename T, int L>
static Taps1<T, L> make_taps1(const nddwt_plan *p, bool rec)
{
    using TT = typename Tap1Of<T>::type;
    Taps1<T, L> t;
    const AllTaps<double> &src = rec ? p->rec_d : p->dec_d;
    for (int k = 0; k < L; ++k) {
        t.lo[k] = mk1(TT(), src.d[0].lo[k]);
        t.hi[k] = mk1(TT(), src.d[0].hi[k]);
    }
    for (int j = 0; j < NDDWT_MAX_LEVELS; ++j)
        t.thr[j] = (typename Elem<T>::R)((!rec && p->shrink_mode) ? p->shrink_thr[j][1] : 0.0);
    return t;
}

template <typename T, int L, int TILE>
static int launch1(nddwt_plan *p, bool rec, const void *in, void *out, int J, cudaStream_t s)
{
    constexpr int NT = 256, CPT = 1;   // 16-byte chunks per thread and level (2: measured slower, 7.1 + 10.4 ms vs 5.4 + 9.6 ms on cfg2)
    const int64_t n1 = p->dims[0], batch = p->batch;
    const size_t W = ((size_t)TILE + (size_t)J * (L - 1) + 4 * (16 / sizeof(T)) + 3) & ~(size_t)3;
    const size_t smem = (rec ? 3 : 2) * W * sizeof(T);
    if (smem > 200 * 1024 || batch > 65535) return 1;
    dim3 grid((unsigned)((n1 + TILE - 1) / TILE), (unsigned)batch);
    if (rec) {
        auto kern = k_rec1_cascade<T, L, TILE, NT, CPT>;
        NDDWT_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        LaunchTimer lt(p, KIND_REC3, s);
        kern<<<grid, NT, smem, s>>>(reinterpret_cast<const T *>(in), reinterpret_cast<T *>(out), n1, batch, J,
                                    make_taps1<T, L>(p, true));
    } else {
        auto kern = k_dec1_cascade<T, L, TILE, NT, CPT>;
        NDDWT_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        LaunchTimer lt(p, KIND_DEC3, s);
        kern<<<grid, NT, smem, s>>>(reinterpret_cast<const T *>(in), reinterpret_cast<T *>(out), n1, batch, J,
                                    make_taps1<T, L>(p, false));
    }
    p->launches++;
    NDDWT_CUDA(cudaGetLastError());
    return 0;
}

template <typename T, int L>
static int launch1_tile(nddwt_plan *p, bool rec, const void *in, void *out, int J, cudaStream_t s)
{
    if (p->dims[0] >= 1024) return launch1<T, L, 2048>(p, rec, in, out, J, s);
    return launch1<T, L, 256>(p, rec, in, out, J, s);
}

#define NDDWT1_L_SWITCH(L_, CALL)                          \
    switch (L_) {                                          \
        case 2: { constexpr int LL = 2; return CALL; }     \
        case 4: { constexpr int LL = 4; return CALL; }     \
        case 6: { constexpr int LL = 6; return CALL; }     \
        case 8: { constexpr int LL = 8; return CALL; }     \
        case 10: { constexpr int LL = 10; return CALL; }   \
        case 12: { constexpr int LL = 12; return CALL; }   \
        case 14: { constexpr int LL = 14; return CALL; }   \
        case 16: { constexpr int LL = 16; return CALL; }   \
        case 18: { constexpr int LL = 18; return CALL; }   \
        case 20: { constexpr int LL = 20; return CALL; }   \
        default: return 1;                                 \
    }

template <typename T>
static int dispatch1(nddwt_plan *p, bool rec, const void *in, void *out, int J, cudaStream_t s)
{
    NDDWT1_L_SWITCH(p->L[0], (launch1_tile<T, LL>(p, rec, in, out, J, s)));
}

// whole multi-level transform of a 1-D plan in one launch; returns 1 when not applicable
int fused1d_transform(nddwt_plan *p, bool rec, const void *in, void *out, int level, cudaStream_t s)
{
    if (p->ndims != 1 || p->kernel_mode != 0) return 1;
    for (int j = 0; j < level; ++j)
        if (p->dil[j] != 1) return 1;
    switch (p->dtype) {
        case NDDWT_F32: return dispatch1<float>(p, rec, in, out, level, s);
        case NDDWT_F64: return dispatch1<double>(p, rec, in, out, level, s);
        case NDDWT_C64: return dispatch1<float2>(p, rec, in, out, level, s);
        case NDDWT_C128: return dispatch1<double2>(p, rec, in, out, level, s);
    }
    return 1;
}

}  // namespace nddwt
