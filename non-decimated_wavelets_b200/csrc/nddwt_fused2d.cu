// nddwt_fused2d.cu -- fused per-level kernels for 2-D arrays: one launch reads the band once
// (haloed 32 x 16 tile staged in shared memory, periodic wrap folded into the load) and writes the
// four subbands once; synthesis reads the four subbands once and writes the band once.
// Replaces level_1_dec / level_1_rec of Functions/nd_dwt_2D.m:312-337 (and harr_nddwt_2D.m:250-323)
// and nd_dwt_dec_1level / nd_dwt_rec_1level (mex/nddwt.c:98-186) for num_dims == 2.
// Any tap length (db1..db10), any of the four element types, any (odd) sizes; mixed wavelets run with the
// longer tap length, the shorter filter zero-padded symmetrically (same phase, same result).
#include <algorithm>
#include "nddwt_plan.h"

namespace nddwt {

template <typename T, int L>
struct Taps2 {
    typename Elem<T>::R lo[2][L];
    typename Elem<T>::R hi[2][L];
    typename Elem<T>::R thr[4];   // fused coefficient shrink: soft threshold per subband (0 = keep), analysis only
};

__device__ __forceinline__ int wrap2(int m, int n)
{
    m %= n;
    return m < 0 ? m + n : m;
}

template <typename T, int L, int TX, int TY, int NT>
__global__ void __launch_bounds__(NT)
k_dec2_fused(const T *__restrict__ in, T *__restrict__ o0, T *__restrict__ o1, T *__restrict__ o2,
             T *__restrict__ o3, int n1, int n2, const Taps2<T, L> tp)
{
    {   // blockIdx.z: independent planes of a 3-D / 4-D / batched array (hybrid path: outer dims by the generic passes)
        const int64_t po = (int64_t)blockIdx.z * n1 * n2;
        in += po; o0 += po; o1 += po; o2 += po; o3 += po;
    }
    constexpr int H = L - 1, HB = L / 2 - 1;          // analysis reads n-(L/2-1) .. n+L/2
    constexpr int W1 = TX + H, W2 = TY + H, P = W1 | 1;   // odd pitch
    extern __shared__ __align__(16) unsigned char smem2_raw[];
    T *IN = reinterpret_cast<T *>(smem2_raw);         // [W2][P]
    T *S = IN + W2 * P;                               // [2][TY][P]   lo2 / hi2, columns still haloed
    const int tid = threadIdx.x;
    const int a1 = blockIdx.x * TX, a2 = blockIdx.y * TY;
    for (int q = tid; q < W1 * W2; q += NT) {
        const int r = q / W1, c = q - r * W1;
        IN[r * P + c] = __ldg(in + (int64_t)wrap2(a2 - HB + r, n2) * n1 + wrap2(a1 - HB + c, n1));
    }
    __syncthreads();
    // dim 2: y[j] = sum_k g[k] x[j + (L-1) - k] in tile rows
    for (int q = tid; q < W1 * TY; q += NT) {
        const int j = q / W1, c = q - j * W1;
        T lo = zero_of(T()), hi = zero_of(T());
#pragma unroll
        for (int k = 0; k < L; ++k) {
            const T v = IN[(j + H - k) * P + c];
            mac(lo, tp.lo[1][k], v);
            mac(hi, tp.hi[1][k], v);
        }
        S[j * P + c] = lo;
        S[(TY + j) * P + c] = hi;
    }
    __syncthreads();
    // dim 1 and the four subband stores (band = b1 + 2 b2)
    for (int q = tid; q < 2 * TX * TY; q += NT) {
        const int c = q % TX, rest = q / TX;
        const int j = rest % TY, b2 = rest / TY;
        const T *row = S + (b2 * TY + j) * P + c;
        T lo = zero_of(T()), hi = zero_of(T());
#pragma unroll
        for (int k = 0; k < L; ++k) {
            const T v = row[H - k];
            mac(lo, tp.lo[0][k], v);
            mac(hi, tp.hi[0][k], v);
        }
        const int g1 = a1 + c, g2 = a2 + j;
        if (g1 < n1 && g2 < n2) {
            const int64_t idx = (int64_t)g2 * n1 + g1;
            (b2 ? o2 : o0)[idx] = shrink1(lo, tp.thr[2 * b2]);
            (b2 ? o3 : o1)[idx] = shrink1(hi, tp.thr[2 * b2 + 1]);
        }
    }
}

template <typename T, int L, int TX, int TY, int NT>
__global__ void __launch_bounds__(NT)
k_rec2_fused(const T *__restrict__ c0, const T *__restrict__ c1, const T *__restrict__ c2,
             const T *__restrict__ c3, T *__restrict__ out, int n1, int n2, const Taps2<T, L> tp)
{
    {
        const int64_t po = (int64_t)blockIdx.z * n1 * n2;
        c0 += po; c1 += po; c2 += po; c3 += po; out += po;
    }
    constexpr int H = L - 1, HB = L / 2;              // synthesis reads n-L/2 .. n+L/2-1
    constexpr int W1 = TX + H, W2 = TY + H, P = W1 | 1;
    extern __shared__ __align__(16) unsigned char smem2_raw[];
    T *IN = reinterpret_cast<T *>(smem2_raw);         // [4][W2][P]
    T *U = IN + 4 * W2 * P;                           // [2][TY][P]   (b1), dim 2 synthesised
    const int tid = threadIdx.x;
    const int a1 = blockIdx.x * TX, a2 = blockIdx.y * TY;
    const T *bands[4] = {c0, c1, c2, c3};
    for (int q = tid; q < 4 * W1 * W2; q += NT) {
        const int b = q / (W1 * W2), rem = q - b * (W1 * W2);
        const int r = rem / W1, c = rem - r * W1;
        IN[(b * W2 + r) * P + c] =
            __ldg(bands[b] + (int64_t)wrap2(a2 - HB + r, n2) * n1 + wrap2(a1 - HB + c, n1));
    }
    __syncthreads();
    // dim 2: u_b1[j] = sum_k lo2[k] c_{b1}[j + k] + hi2[k] c_{b1 + 2}[j + k]
    for (int q = tid; q < 2 * W1 * TY; q += NT) {
        const int c = q % W1, rest = q / W1;
        const int j = rest % TY, b1 = rest / TY;
        T acc = zero_of(T());
#pragma unroll
        for (int k = 0; k < L; ++k) {
            mac(acc, tp.lo[1][k], IN[(b1 * W2 + j + k) * P + c]);
            mac(acc, tp.hi[1][k], IN[((b1 + 2) * W2 + j + k) * P + c]);
        }
        U[(b1 * TY + j) * P + c] = acc;
    }
    __syncthreads();
    for (int q = tid; q < TX * TY; q += NT) {
        const int c = q % TX, j = q / TX;
        T acc = zero_of(T());
#pragma unroll
        for (int k = 0; k < L; ++k) {
            mac(acc, tp.lo[0][k], U[j * P + c + k]);
            mac(acc, tp.hi[0][k], U[(TY + j) * P + c + k]);
        }
        const int g1 = a1 + c, g2 = a2 + j;
        if (g1 < n1 && g2 < n2) out[(int64_t)g2 * n1 + g1] = acc;
    }
}

template <typename T, int L>
static Taps2<T, L> make_taps2(const nddwt_plan *p, bool rec, int band0 = 0)
{
    using R = typename Elem<T>::R;
    Taps2<T, L> t;
    const AllTaps<double> &src = rec ? p->rec_d : p->dec_d;
    for (int d = 0; d < 2; ++d)
        for (int k = 0; k < L; ++k) {
            t.lo[d][k] = (R)padded_tap(src.d[d].lo, p->L[d], L, k, p->cur_dil);     // mixed wavelets: zero-padded; a-trous levels: stretched
            t.hi[d][k] = (R)padded_tap(src.d[d].hi, p->L[d], L, k, p->cur_dil);
        }
    const int j = p->cur_level >= 1 && p->cur_level <= NDDWT_MAX_LEVELS ? p->cur_level : 1;
    for (int b = 0; b < 4; ++b) t.thr[b] = (R)((!rec && p->shrink_mode) ? p->shrink_thr[j - 1][band0 + b] : 0.0);
    return t;
}

template <typename T, int L>
static int launch_dec2(nddwt_plan *p, const void *a_in, void *const *out_bands, cudaStream_t s, int64_t planes = 1,
                       int band0 = 0)
{
    constexpr int TX = 32, TY = 16, NT = 256, H = L - 1, P = (TX + H) | 1;
    const int n1 = (int)p->dims[0], n2 = (int)p->dims[1];
    const size_t smem = (size_t)((TY + H) * P + 2 * TY * P) * sizeof(T);
    auto kern = k_dec2_fused<T, L, TX, TY, NT>;
    NDDWT_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));   // per-device attribute: set on every launch
    const Taps2<T, L> tp = make_taps2<T, L>(p, false, band0);
    const int64_t ps = (int64_t)n1 * n2;
    for (int64_t z0 = 0; z0 < planes; z0 += 65535) {        // grid.z limit
        const unsigned nz = (unsigned)std::min<int64_t>(65535, planes - z0);
        dim3 grid((n1 + TX - 1) / TX, (n2 + TY - 1) / TY, nz);
        {
            LaunchTimer lt(p, KIND_DEC3, s);
            kern<<<grid, NT, smem, s>>>(reinterpret_cast<const T *>(a_in) + z0 * ps, reinterpret_cast<T *>(out_bands[0]) + z0 * ps,
                                        reinterpret_cast<T *>(out_bands[1]) + z0 * ps, reinterpret_cast<T *>(out_bands[2]) + z0 * ps,
                                        reinterpret_cast<T *>(out_bands[3]) + z0 * ps, n1, n2, tp);
        }
        p->launches++;
        NDDWT_CUDA(cudaGetLastError());
    }
    return 0;
}

template <typename T, int L>
static int launch_rec2(nddwt_plan *p, const void *const *in_bands, void *a_out, cudaStream_t s, int64_t planes = 1)
{
    constexpr int TX = 32, TY = 16, NT = 256, H = L - 1, P = (TX + H) | 1;
    const int n1 = (int)p->dims[0], n2 = (int)p->dims[1];
    const size_t smem = (size_t)(4 * (TY + H) * P + 2 * TY * P) * sizeof(T);
    auto kern = k_rec2_fused<T, L, TX, TY, NT>;
    NDDWT_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));   // per-device attribute: set on every launch
    const Taps2<T, L> tp = make_taps2<T, L>(p, true);
    const int64_t ps = (int64_t)n1 * n2;
    for (int64_t z0 = 0; z0 < planes; z0 += 65535) {
        const unsigned nz = (unsigned)std::min<int64_t>(65535, planes - z0);
        dim3 grid((n1 + TX - 1) / TX, (n2 + TY - 1) / TY, nz);
        {
            LaunchTimer lt(p, KIND_REC3, s);
            kern<<<grid, NT, smem, s>>>(reinterpret_cast<const T *>(in_bands[0]) + z0 * ps, reinterpret_cast<const T *>(in_bands[1]) + z0 * ps,
                                        reinterpret_cast<const T *>(in_bands[2]) + z0 * ps, reinterpret_cast<const T *>(in_bands[3]) + z0 * ps,
                                        reinterpret_cast<T *>(a_out) + z0 * ps, n1, n2, tp);
        }
        p->launches++;
        NDDWT_CUDA(cudaGetLastError());
    }
    return 0;
}

#define NDDWT2_L_SWITCH(L_, CALL)                          \
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

static int taps2d(const nddwt_plan *p) { return (p->L[0] > p->L[1] ? p->L[0] : p->L[1]) * p->cur_dil; }   // dims 1, 2 only; stretched on a-trous levels

template <typename T>
static int dispatch_dec2(nddwt_plan *p, const void *a_in, void *const *out_bands, cudaStream_t s, int64_t planes = 1,
                         int band0 = 0)
{
    NDDWT2_L_SWITCH(taps2d(p), (launch_dec2<T, LL>(p, a_in, out_bands, s, planes, band0)));
}
template <typename T>
static int dispatch_rec2(nddwt_plan *p, const void *const *in_bands, void *a_out, cudaStream_t s, int64_t planes = 1)
{
    NDDWT2_L_SWITCH(taps2d(p), (launch_rec2<T, LL>(p, in_bands, a_out, s, planes)));
}

static bool ok2d_geometry(const nddwt_plan *p)
{
    if (p->ndims < 2 || taps2d(p) > 20) return false;
    if (p->dims[0] < taps2d(p) || p->dims[1] < taps2d(p)) return false;
    if (p->dims[0] > 0x7fffffff - 64 || p->dims[1] > (int64_t)65535 * 16) return false;
    return p->dims[0] * p->dims[1] < ((int64_t)1 << 31);
}

// Hybrid path of 3-D / 4-D / batched arrays whose outer dimensions the tile kernels do not take (db5..db10, batches):
// `planes` independent (dim 1, dim 2) planes through the fused 2-D kernels; the caller (nddwt_generic.cu) has filtered
// the outer dimensions with the generic passes.  band0: index of out_bands[0] within the level (thresholds).
int fused2d_dec_planes(nddwt_plan *p, const void *a_in, void *const *out_bands, int64_t planes, int band0, cudaStream_t s)
{
    if (!ok2d_geometry(p)) return 1;
    switch (p->dtype) {
        case NDDWT_F32: return dispatch_dec2<float>(p, a_in, out_bands, s, planes, band0);
        case NDDWT_F64: return dispatch_dec2<double>(p, a_in, out_bands, s, planes, band0);
        case NDDWT_C64: return dispatch_dec2<float2>(p, a_in, out_bands, s, planes, band0);
        case NDDWT_C128: return dispatch_dec2<double2>(p, a_in, out_bands, s, planes, band0);
    }
    return 1;
}

int fused2d_rec_planes(nddwt_plan *p, const void *const *in_bands, void *a_out, int64_t planes, cudaStream_t s)
{
    if (!ok2d_geometry(p)) return 1;
    switch (p->dtype) {
        case NDDWT_F32: return dispatch_rec2<float>(p, in_bands, a_out, s, planes);
        case NDDWT_F64: return dispatch_rec2<double>(p, in_bands, a_out, s, planes);
        case NDDWT_C64: return dispatch_rec2<float2>(p, in_bands, a_out, s, planes);
        case NDDWT_C128: return dispatch_rec2<double2>(p, in_bands, a_out, s, planes);
    }
    return 1;
}

static bool ok2d(const nddwt_plan *p, int dil, const LevelIO *io)
{
    if (dil != p->cur_dil || p->ndims != 2 || p->batch != 1) return false;
    if (taps2d(p) > 20) return false;                          // stretched taps beyond the longest instantiation
    if (p->dims[0] < taps2d(p) || p->dims[1] < taps2d(p)) return false;
    if (io && (io->halo_lo || io->halo_hi)) return false;      // slabs of 2-D arrays use the generic kernels
    // launch geometry: dims are ints in the kernels, grid.y = ceil(n2 / 16) must stay <= 65535
    if (p->dims[0] > 0x7fffffff - 64 || p->dims[1] > (int64_t)65535 * 16) return false;
    return p->dims[0] * p->dims[1] < ((int64_t)1 << 40);
}

int fused2d_dec_level(nddwt_plan *p, int dil, const void *a_in, const LevelIO &io, void *const *out_bands,
                      cudaStream_t s)
{
    if (dil < 1 || dil > 10) return 1;
    DilScope ds(p, dil);
    if (!ok2d(p, dil, &io)) return 1;
    switch (p->dtype) {
        case NDDWT_F32: return dispatch_dec2<float>(p, a_in, out_bands, s);
        case NDDWT_F64: return dispatch_dec2<double>(p, a_in, out_bands, s);
        case NDDWT_C64: return dispatch_dec2<float2>(p, a_in, out_bands, s);
        case NDDWT_C128: return dispatch_dec2<double2>(p, a_in, out_bands, s);
    }
    return 1;
}

int fused2d_rec_level(nddwt_plan *p, int dil, const void *const *in_bands, void *a_out, cudaStream_t s)
{
    if (dil < 1 || dil > 10) return 1;
    DilScope ds(p, dil);
    if (!ok2d(p, dil, nullptr)) return 1;
    switch (p->dtype) {
        case NDDWT_F32: return dispatch_rec2<float>(p, in_bands, a_out, s);
        case NDDWT_F64: return dispatch_rec2<double>(p, in_bands, a_out, s);
        case NDDWT_C64: return dispatch_rec2<float2>(p, in_bands, a_out, s);
        case NDDWT_C128: return dispatch_rec2<double2>(p, in_bands, a_out, s);
    }
    return 1;
}

}  // namespace nddwt
