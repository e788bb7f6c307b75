// nddwt_generic.cu -- generic separable kernels: one circular lo/hi filtering pass along one
// dimension, any tap length (db1..db10, mixed per dim), any dilation, any of the four element
// types.  This is the always-available GPU path (mixed wavelets, a-trous dilation, odd shapes
// the fused kernels do not instantiate).  It is NOT a CPU fallback: everything runs on the device.
//
// Reference semantics replaced (per level): nd_dwt_dec_1level / nd_dwt_rec_1level
// (mex/nddwt.c:98-186) and level_1_dec / level_1_rec of Functions/nd_dwt_{1,2,3,4}D.m, restated
// in the spatial domain:
//   analysis  along a dim:  y_g[n] = sum_k g[k] x[(n - (k - L/2) dil) mod N]
//   synthesis along a dim:  x[n]   = sum_k lo[k] cL[(n + (k - L/2) dil) mod N] + hi[k] cH[...]
#include "nddwt_plan.h"

namespace nddwt {

template <typename T>
__global__ void __launch_bounds__(256)
k_dec_dim(const T *__restrict__ in, const T *__restrict__ halo_lo, const T *__restrict__ halo_hi,
          T *__restrict__ out_lo, T *__restrict__ out_hi, int64_t inner, int64_t n, int64_t total,
          int L, int dil, int below, DimTaps<typename Elem<T>::R> taps)
{
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const bool slab = (halo_lo != nullptr) || (halo_hi != nullptr);
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += stride) {
        const int64_t i = idx % inner;
        const int64_t t = idx / inner;
        const int64_t p = t % n;
        const int64_t o = t / n;
        const int64_t base = i + inner * n * o;
        T alo = zero_of(T()), ahi = zero_of(T());
        for (int k = 0; k < L; ++k) {
            int64_t m = p - (int64_t)(k - L / 2) * dil;
            T v;
            if (!slab) {
                m = wrap(m, n);
                v = in[base + m * inner];
            } else if (m < 0) {
                v = halo_lo[i + (m + below) * inner];
            } else if (m >= n) {
                v = halo_hi[i + (m - n) * inner];
            } else {
                v = in[base + m * inner];
            }
            mac(alo, taps.lo[k], v);
            mac(ahi, taps.hi[k], v);
        }
        out_lo[idx] = alo;
        out_hi[idx] = ahi;
    }
}

template <typename T>
__global__ void __launch_bounds__(256)
k_rec_dim(const T *__restrict__ in_lo, const T *__restrict__ in_hi, const T *__restrict__ halo_lo,
          const T *__restrict__ halo_hi, T *__restrict__ out, int64_t inner, int64_t n, int64_t total,
          int L, int dil, int below, int above, DimTaps<typename Elem<T>::R> taps)
{
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const bool slab = (halo_lo != nullptr) || (halo_hi != nullptr);
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += stride) {
        const int64_t i = idx % inner;
        const int64_t t = idx / inner;
        const int64_t p = t % n;
        const int64_t o = t / n;
        const int64_t base = i + inner * n * o;
        T acc = zero_of(T());
        for (int k = 0; k < L; ++k) {
            int64_t m = p + (int64_t)(k - L / 2) * dil;
            T vl, vh;
            if (!slab) {
                m = wrap(m, n);
                vl = in_lo[base + m * inner];
                vh = in_hi[base + m * inner];
            } else if (m < 0) {   // halo_lo = [u_lo planes (below)] [u_hi planes (below)]
                vl = halo_lo[i + (m + below) * inner];
                vh = halo_lo[i + (m + below + below) * inner];
            } else if (m >= n) {  // halo_hi = [u_lo planes (above)] [u_hi planes (above)]
                vl = halo_hi[i + (m - n) * inner];
                vh = halo_hi[i + (m - n + above) * inner];
            } else {
                vl = in_lo[base + m * inner];
                vh = in_hi[base + m * inner];
            }
            mac(acc, taps.lo[k], vl);
            mac(acc, taps.hi[k], vh);
        }
        out[idx] = acc;
    }
}

template <typename T>
__global__ void __launch_bounds__(256) k_shrink_band(T *__restrict__ band, int64_t n, typename Elem<T>::R thr)
{
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) band[i] = shrink1(band[i], thr);
}

static inline int grid_for(int64_t total)
{
    int64_t b = (total + 255) / 256;
    const int64_t cap = 148 * 16;
    return (int)(b < cap ? (b < 1 ? 1 : b) : cap);
}

template <typename R> static const AllTaps<R> &dec_taps(const nddwt_plan *p);
template <> const AllTaps<float> &dec_taps<float>(const nddwt_plan *p) { return p->dec_f; }
template <> const AllTaps<double> &dec_taps<double>(const nddwt_plan *p) { return p->dec_d; }
template <typename R> static const AllTaps<R> &rec_taps(const nddwt_plan *p);
template <> const AllTaps<float> &rec_taps<float>(const nddwt_plan *p) { return p->rec_f; }
template <> const AllTaps<double> &rec_taps<double>(const nddwt_plan *p) { return p->rec_d; }

static void dim_geometry(const nddwt_plan *p, int k, int64_t &inner, int64_t &n, int64_t &outer)
{
    inner = 1;
    outer = 1;
    for (int i = 0; i < k; ++i) inner *= p->dims[i];
    n = p->dims[k];
    for (int i = k + 1; i < p->ndims; ++i) outer *= p->dims[i];
}

template <typename T>
static int launch_dec_dim(nddwt_plan *p, int k, int dil, const T *in, const T *halo_lo, const T *halo_hi,
                          T *out_lo, T *out_hi, cudaStream_t s)
{
    using R = typename Elem<T>::R;
    int64_t inner, n, outer;
    dim_geometry(p, k, inner, n, outer);
    const int L = p->L[k];
    const int below = (L / 2 - 1) * dil;
    {
        LaunchTimer lt(p, KIND_GENERIC, s);
        k_dec_dim<T><<<grid_for(p->numel), 256, 0, s>>>(in, halo_lo, halo_hi, out_lo, out_hi, inner, n, p->numel, L,
                                                        dil, below, dec_taps<R>(p).d[k]);
    }
    p->launches++;
    NDDWT_CUDA(cudaGetLastError());
    return 0;
}

template <typename T>
static int launch_rec_dim(nddwt_plan *p, int k, int dil, const T *in_lo, const T *in_hi, const T *halo_lo,
                          const T *halo_hi, T *out, cudaStream_t s)
{
    using R = typename Elem<T>::R;
    int64_t inner, n, outer;
    dim_geometry(p, k, inner, n, outer);
    const int L = p->L[k];
    const int below = (L / 2) * dil, above = (L / 2 - 1) * dil;
    {
        LaunchTimer lt(p, KIND_GENERIC, s);
        k_rec_dim<T><<<grid_for(p->numel), 256, 0, s>>>(in_lo, in_hi, halo_lo, halo_hi, out, inner, n, p->numel, L,
                                                        dil, below, above, rec_taps<R>(p).d[k]);
    }
    p->launches++;
    NDDWT_CUDA(cudaGetLastError());
    return 0;
}

// scratch layout: for k = 1..d-1, two arrays (lo, hi) each of numel elements
template <typename T>
static T *scratch_at(nddwt_plan *p, int k, int which)
{
    return reinterpret_cast<T *>(p->gen_scratch) + (int64_t)(2 * (k - 1) + which) * p->numel;
}

template <typename T>
static int dec_recurse(nddwt_plan *p, int k, int dil, const T *in, const LevelIO &io, int idx,
                       void *const *out_bands, cudaStream_t s)
{
    const bool last = (k == p->ndims - 1);
    const T *hl = last ? reinterpret_cast<const T *>(io.halo_lo) : nullptr;
    const T *hh = last ? reinterpret_cast<const T *>(io.halo_hi) : nullptr;
    // hybrid: dims 1 and 2 of every remaining plane in one fused 2-D launch (4 bands written once) instead of two more
    // generic passes with their intermediates; the outer dims were filtered by the passes above
    if (k == 1 && dil >= 1 && dil <= 10 && p->kernel_mode == 0 && !(last && (hl || hh))) {
        const int64_t planes = p->numel / (p->dims[0] * p->dims[1]);
        DilScope ds(p, dil);                   // a-trous levels: stretched taps in the 2-D kernels (nddwt_plan.h)
        const int rc = fused2d_dec_planes(p, in, out_bands + idx * 4, planes, idx * 4, s);
        if (rc <= 0) { p->hybrid_used = true; return rc; }
    }
    if (k == 0) {
        return launch_dec_dim<T>(p, 0, dil, in, hl, hh, reinterpret_cast<T *>(out_bands[idx * 2 + 0]),
                                 reinterpret_cast<T *>(out_bands[idx * 2 + 1]), s);
    }
    T *tlo = scratch_at<T>(p, k, 0), *thi = scratch_at<T>(p, k, 1);
    int rc = launch_dec_dim<T>(p, k, dil, in, hl, hh, tlo, thi, s);
    if (rc) return rc;
    LevelIO none;
    rc = dec_recurse<T>(p, k - 1, dil, tlo, none, idx * 2 + 0, out_bands, s);
    if (rc) return rc;
    return dec_recurse<T>(p, k - 1, dil, thi, none, idx * 2 + 1, out_bands, s);
}

// S(k, idx): synthesis over dims 0..k-1 of the band group whose bits for dims >= k are idx.
// Result goes to `dst` (scratch or the final output).
template <typename T>
static int rec_recurse(nddwt_plan *p, int k, int dil, int idx, const void *const *bands, T *dst, cudaStream_t s)
{
    if (k == 2 && dil >= 1 && dil <= 10 && p->kernel_mode == 0) {      // hybrid: dims 1, 2 by the fused 2-D kernels
        const int64_t planes = p->numel / (p->dims[0] * p->dims[1]);
        DilScope ds(p, dil);
        const int rc = fused2d_rec_planes(p, bands + idx * 4, dst, planes, s);
        if (rc <= 0) { p->hybrid_used = true; return rc; }
    }
    // operands S(k-1, 2 idx) and S(k-1, 2 idx + 1)
    const T *lo, *hi;
    if (k == 1) {
        lo = reinterpret_cast<const T *>(bands[idx * 2 + 0]);
        hi = reinterpret_cast<const T *>(bands[idx * 2 + 1]);
    } else {
        T *slo = scratch_at<T>(p, k - 1, 0), *shi = scratch_at<T>(p, k - 1, 1);
        int rc = rec_recurse<T>(p, k - 1, dil, idx * 2 + 0, bands, slo, s);
        if (rc) return rc;
        rc = rec_recurse<T>(p, k - 1, dil, idx * 2 + 1, bands, shi, s);
        if (rc) return rc;
        lo = slo;
        hi = shi;
    }
    return launch_rec_dim<T>(p, k - 1, dil, lo, hi, nullptr, nullptr, dst, s);
}

#define NDDWT_DISPATCH(p, CALL)                                        \
    switch ((p)->dtype) {                                              \
        case NDDWT_F32: { using T = float; return CALL; }              \
        case NDDWT_F64: { using T = double; return CALL; }             \
        case NDDWT_C64: { using T = float2; return CALL; }             \
        case NDDWT_C128: { using T = double2; return CALL; }           \
        default: set_error("bad dtype"); return NDDWT_ERR_ARG;         \
    }

template <typename T>
static int shrink_launch(nddwt_plan *p, T *band, int64_t n, double thr, cudaStream_t s)
{
    {
        LaunchTimer lt(p, KIND_GENERIC, s);
        k_shrink_band<T><<<grid_for(n), 256, 0, s>>>(band, n, (typename Elem<T>::R)thr);
    }
    p->launches++;
    NDDWT_CUDA(cudaGetLastError());
    return 0;
}

int shrink_band(nddwt_plan *p, void *band, int64_t nelem, double thr, cudaStream_t s)
{
    if (thr <= 0.0) return 0;
    NDDWT_DISPATCH(p, shrink_launch<T>(p, reinterpret_cast<T *>(band), nelem, thr, s));
}

int generic_dec_level(nddwt_plan *p, int dil, const void *a_in, const LevelIO &io, void *const *out_bands,
                      cudaStream_t s)
{
    int rc = ensure_scratch(p);
    if (rc) return rc;
    NDDWT_DISPATCH(p, dec_recurse<T>(p, p->ndims - 1, dil, reinterpret_cast<const T *>(a_in), io, 0, out_bands, s));
}

int generic_rec_level(nddwt_plan *p, int dil, const void *const *in_bands, void *a_out, cudaStream_t s)
{
    int rc = ensure_scratch(p);
    if (rc) return rc;
    NDDWT_DISPATCH(p, rec_recurse<T>(p, p->ndims, dil, 0, in_bands, reinterpret_cast<T *>(a_out), s));
}

template <typename T>
static int rec_stage1(nddwt_plan *p, int dil, const void *const *bands, T *u_lo, T *u_hi, cudaStream_t s)
{
    const int d = p->ndims;
    if (d == 1) {  // nothing to synthesise locally: u = the two bands themselves
        NDDWT_CUDA(cudaMemcpyAsync(u_lo, bands[0], p->numel * p->esize, cudaMemcpyDeviceToDevice, s));
        NDDWT_CUDA(cudaMemcpyAsync(u_hi, bands[1], p->numel * p->esize, cudaMemcpyDeviceToDevice, s));
        return 0;
    }
    int rc = rec_recurse<T>(p, d - 1, dil, 0, bands, u_lo, s);
    if (rc) return rc;
    return rec_recurse<T>(p, d - 1, dil, 1, bands, u_hi, s);
}

int generic_rec_stage1(nddwt_plan *p, int dil, const void *const *in_bands, void *u_lo, void *u_hi, cudaStream_t s)
{
    int rc = ensure_scratch(p);
    if (rc) return rc;
    NDDWT_DISPATCH(p, rec_stage1<T>(p, dil, in_bands, reinterpret_cast<T *>(u_lo), reinterpret_cast<T *>(u_hi), s));
}

int generic_rec_stage2(nddwt_plan *p, int dil, const void *u_lo, const void *u_hi, const LevelIO &io, void *a_out,
                       cudaStream_t s)
{
    NDDWT_DISPATCH(p, launch_rec_dim<T>(p, p->ndims - 1, dil, reinterpret_cast<const T *>(u_lo),
                                        reinterpret_cast<const T *>(u_hi), reinterpret_cast<const T *>(io.halo_lo),
                                        reinterpret_cast<const T *>(io.halo_hi), reinterpret_cast<T *>(a_out), s));
}

}  // namespace nddwt
