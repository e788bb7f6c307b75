// nddwt_fused1d.cu -- 1-D multi-level CASCADE kernels: a CTA loads one tile of a signal (plus the
// halo all J levels need, periodic wrap folded into the load) into shared memory once, runs the J
// analysis levels there (approximation ping-pongs between two shared buffers), and writes the J
// detail bands and the final approximation once -- 1 read + (J+1) writes per sample, the compulsory
// traffic, instead of the 3 J passes of a level-by-level scheme.  Synthesis mirrors it.
// Replaces dec / rec / level_1_dec / level_1_rec of Functions/nd_dwt_1D.m:136-318 and
// nd_dwt_dec / nd_dwt_rec (mex/nddwt.c:189-292) for num_dims == 1, for one signal or a batch of
// signals (batch extension; BASELINE configs[1]: 4096 signals x 65536 samples, db8, 6 levels).
#include <cstdlib>
#include "nddwt_plan.h"

namespace nddwt {

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

// ---------------------------------------------------------------------------------------------
// Run kernels (round 2).  The first version of the cascade (256 threads, 2 outputs per thread and iteration,
// 2048-sample tiles) took 5.5 + 8.9 ms on BASELINE configs[1] (4096 x 65536, db8, J6, complex single) against an
// FP32-issue floor of 2.9 + 2.9 ms: 9 / 18 16-byte shared-memory loads per 2 outputs, a tile of 2048 + halo
// samples = 4.2 iterations of the CTA (i.e. 5), element-wise detail stores, a 64-bit modulo per level and thread.
//   * a thread produces a RUN of R = CPT x VEC consecutive outputs per level from a sliding window of the source
//     held in a statically indexed register ring (one 16-byte load per VEC outputs); CPT is odd, so the lanes'
//     16-byte accesses (stride CPT chunks) are bank-conflict-free;
//   * the tile length is chosen on the host so that the widest level is exactly `kit` full iterations of the CTA;
//   * every level's buffer origin moves by the halo rounded up to whole 16-byte chunks (the window then starts
//     SA / SR elements into its first chunk), so shared and global accesses stay chunk-aligned at every level;
//   * analysis: the detail outputs of a level are staged in shared memory (double-buffered) and leave as coalesced
//     16-byte streaming stores while the next level computes;  synthesis: d_{j-1} is fetched by cp.async into the
//     other detail buffer while level j computes (no exposed global latency per level);
//   * the soft threshold runs only on levels whose threshold is non-zero.
// Measured steps (cfg2, ms per launch, analysis + synthesis): 5.58 + 8.94 -> 5.38 + 4.85 (runs) -> 4.57 + 5.02
// (threshold skipped, lo / hi chains interleaved) -> 4.32 + 4.47 (no integer division in the staging loops, leaner
// store loops): pair 14.45 -> 8.8 ms, 36 % -> 60 % of the HBM roofline of the pair (profiles/r02_variants.md).
template <typename T, int L>
struct Casc {
    static constexpr int VEC = 16 / (int)sizeof(T);
    static constexpr int HLS = L / 2 - 1;                               // analysis reads n-(L/2-1) .. n+L/2
    static constexpr int HLSP = (HLS + VEC - 1) / VEC * VEC, SA = HLSP - HLS;
    static constexpr int HALO_A = HLSP + L / 2;                         // buffer growth per analysis level
    static constexpr int HLR = L / 2;                                   // synthesis reads n-L/2 .. n+L/2-1
    static constexpr int HLRP = (HLR + VEC - 1) / VEC * VEC, SR = HLRP - HLR;
    static constexpr int HALO_R = HLRP + L / 2 - 1;
    static constexpr int WCH_A = (VEC - 1 + HALO_A) / VEC + 1;          // chunks under the window of one output chunk
    static constexpr int WCH_R = (VEC - 1 + HALO_R) / VEC + 1;
};

template <int B>
__device__ __forceinline__ void cp_async_sa(unsigned smem_addr, const void *gmem)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], %2;" ::"r"(smem_addr), "l"(__cvta_generic_to_global(gmem)), "n"(B) : "memory");
}
template <int B>
__device__ __forceinline__ void cp_async(void *smem, const void *gmem)
{
    cp_async_sa<B>((unsigned)__cvta_generic_to_shared(smem), gmem);
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// `cnt` pieces of B bytes, piece i of thread tid at byte offset (tid + i NT) B of both sides: constant strides, so the
// unrolled body is one LDGSTS per piece with immediate offsets
template <int B, int NT>
__device__ __forceinline__ void stage_linear(void *dst, const void *src, int cnt, int tid)
{
    unsigned sa = (unsigned)__cvta_generic_to_shared(dst) + (unsigned)tid * B;
    const char *g = reinterpret_cast<const char *>(src) + (size_t)tid * B;
    int i = tid;
    for (; i + 3 * NT < cnt; i += 4 * NT) {
#pragma unroll
        for (int u = 0; u < 4; ++u) cp_async_sa<B>(sa + u * NT * B, g + u * NT * B);
        sa += 4 * NT * B;
        g += 4 * NT * B;
    }
    for (; i < cnt; i += NT) {
        cp_async_sa<B>(sa, g);
        sa += NT * B;
        g += NT * B;
    }
}

// asynchronous periodic copy global -> shared of `count` elements starting at global index `start` (any sign).
// vec: start and n1 are multiples of VEC and src is 16-byte aligned -> whole chunks, none straddles the wrap.
// All but the first / last tiles of a signal take the linear path (the window lies inside the signal); the periodic
// path runs without integer division unless the window is longer than the signal (a 64-bit modulo is ~100
// instructions and this runs once per level in the synthesis cascade: 22 % of the stall samples of the first run-kernel
// form; the per-piece wrap bookkeeping was still 28 % of the instructions of the second, profiles/r02_cfg2_ncu.txt).
template <typename T, int NT>
__device__ __forceinline__ void stage_wrapped(T *dst, const T *src, int64_t start, int count, int64_t n1, int tid, bool vec)
{
    constexpr int VEC = 16 / (int)sizeof(T);
    if (vec) {
        const int cnt = (count + VEC - 1) / VEC;
        if (start >= 0 && start + (int64_t)cnt * VEC <= n1) {
            stage_linear<16, NT>(dst, src + start, cnt, tid);
            return;
        }
        const int64_t nc = n1 / VEC;                    // (a shift: VEC is a power of two and n1 >= 0)
        int64_t g = wrap(start / VEC + tid, nc);
        const int64_t step = (NT < nc) ? NT : NT % nc;
        for (int i = tid; i < cnt; i += NT) {
            cp_async<16>(dst + (int64_t)i * VEC, src + g * VEC);
            g += step;
            if (g >= nc) g -= nc;
        }
    } else {
        if (start >= 0 && start + count <= n1) {
            stage_linear<(int)sizeof(T), NT>(dst, src + start, count, tid);
            return;
        }
        int64_t g = wrap(start + tid, n1);
        const int64_t step = (NT < n1) ? NT : NT % n1;
        for (int i = tid; i < count; i += NT) {
            cp_async<(int)sizeof(T)>(dst + i, src + g);
            g += step;
            if (g >= n1) g -= n1;
        }
    }
}

// shared -> global copy of one tile's band (coalesced 16-byte streaming stores when aligned); `lim` = samples of the
// tile inside the signal (a multiple of VEC when vec).  Constant strides: one LDS + one STG per piece when unrolled.
template <typename T, int NT>
__device__ __forceinline__ void store_tile(T *__restrict__ band_tile, const T *__restrict__ buf, int lim, int tid, bool vec)
{
    constexpr int VEC = 16 / (int)sizeof(T);
    if (vec) {
        uint4 *g = reinterpret_cast<uint4 *>(band_tile) + tid;
        const uint4 *sm = reinterpret_cast<const uint4 *>(buf) + tid;
        const int nch = lim / VEC;
        int i = tid;
        for (; i + 3 * NT < nch; i += 4 * NT) {
            uint4 v[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) v[u] = sm[u * NT];
#pragma unroll
            for (int u = 0; u < 4; ++u) __stcs(g + u * NT, v[u]);
            sm += 4 * NT;
            g += 4 * NT;
        }
        for (; i < nch; i += NT) {
            __stcs(g, *sm);
            sm += NT;
            g += NT;
        }
    } else {
#pragma unroll 4
        for (int i = tid; i < lim; i += NT) band_tile[i] = buf[i];
    }
}

// coeffs: [n1][batch][J+1] column-major => band s of signal b starts at (s * batch + b) * n1
// SPS = output chunks produced together (2 * SPS * VEC independent accumulation chains; the taps are fetched once per
// group).  Source chunk q of a run lives in ring slot q % NS, NS = WCH + SPS - 1; a group loads the SPS chunks it needs
// beyond the previous group's window and walks the taps from the oldest source element to the newest, so the loads
// have the whole group to land.
template <typename T, int L, int NT, int CPT, int SPS>
__global__ void __launch_bounds__(NT)
k_dec1_runs(const T *__restrict__ x, T *__restrict__ coeffs, int64_t n1, int64_t batch, int J, int tile, int wb, int vec,
            const Taps1<T, L> tp)
{
    using C = Casc<T, L>;
    constexpr int VEC = C::VEC, R = CPT * VEC, WCH = C::WCH_A, NS = WCH + SPS - 1, D = C::HALO_A;
    extern __shared__ __align__(16) unsigned char smem1_raw[];
    T *buf0 = reinterpret_cast<T *>(smem1_raw);
    T *buf1 = buf0 + wb;
    T *stg0 = buf1 + wb, *stg1 = stg0 + tile;
    T *dump = stg1 + tile;                             // one chunk: where the detail outputs of halo samples go
    const int tid = threadIdx.x;
    const int64_t b = blockIdx.y;
    const int64_t t0 = (int64_t)blockIdx.x * tile;
    const int lim = (int)((n1 - t0 < tile) ? n1 - t0 : tile);   // samples of this tile inside the signal
    // level-0 buffer: origin global t0 - J * HLSP
    stage_wrapped<T, NT>(buf0, x + b * n1, t0 - (int64_t)J * C::HLSP, tile + J * D, n1, tid, vec != 0);
    cp_async_commit();
    cp_async_wait<0>();
    __syncthreads();
    T *src = buf0, *dst = buf1;
    for (int j = 1; j <= J; ++j) {
        const int Wj = tile + (J - j) * D;             // valid outputs of this level
        const int c0 = (J - j) * C::HLSP;              // buffer index of global t0 at this level (chunk-aligned)
        T *stg = (j & 1) ? stg1 : stg0;
        const typename Elem<T>::R thr = tp.thr[j - 1];
        for (int o = tid * R; o < Wj; o += NT * R) {
            T w[NS * VEC];
#pragma unroll
            for (int q = 0; q < NS; ++q) ld16<T, VEC>(src + o + q * VEC, w + q * VEC);
#pragma unroll
            for (int c = 0; c < CPT; c += SPS) {
                const int n = (CPT - c < SPS) ? CPT - c : SPS;          // compile-time after unrolling
                if (c > 0) {
#pragma unroll
                    for (int q = 0; q < SPS; ++q)
                        if (q < n) ld16<T, VEC>(src + o + (c + WCH - 1 + q) * VEC, w + ((c + WCH - 1 + q) % NS) * VEC);
                }
                T lo[SPS * VEC], hi[SPS * VEC];
#pragma unroll
                for (int r = 0; r < SPS * VEC; ++r) { lo[r] = zero_of(T()); hi[r] = zero_of(T()); }
#pragma unroll
                for (int k = L - 1; k >= 0; --k)
#pragma unroll
                    for (int r = 0; r < SPS * VEC; ++r)
                        if (r < n * VEC) {
                            const int idx = c * VEC + r + D - k;        // element of the run's source window
                            const T e = w[((idx / VEC) % NS) * VEC + idx % VEC];
                            mac1(lo[r], tp.lo[k], e);
                            mac1(hi[r], tp.hi[k], e);
                        }
                if (thr > 0) {                                          // fused coefficient shrink (uniform per level)
#pragma unroll
                    for (int r = 0; r < SPS * VEC; ++r)
                        if (r < n * VEC) hi[r] = shrink1(hi[r], thr);
                }
#pragma unroll
                for (int q = 0; q < SPS; ++q)
                    if (q < n) {
                        st16<T, VEC>(dst + o + (c + q) * VEC, lo + q * VEC);
                        // detail d_j: only the tile's own samples leave (the others go to the dump chunk, so that the
                        // lo and hi chains stay interleaved instead of hi being sunk behind a branch)
                        const int oc = o + (c + q) * VEC - c0;
                        st16<T, VEC>((oc >= 0 && oc < tile) ? stg + oc : dump, hi + q * VEC);
                    }
            }
        }
        __syncthreads();
        // d_j (slot J-j+1) leaves while the next level computes; its staging buffer is rewritten two levels later
        store_tile<T, NT>(coeffs + ((int64_t)(J - j + 1) * batch + b) * n1 + t0, stg, lim, tid, vec != 0);
        T *t = src; src = dst; dst = t;
    }
    store_tile<T, NT>(coeffs + b * n1 + t0, src, lim, tid, vec != 0);   // slot 0 = a_J
}

template <typename T, int L, int NT, int CPT, int SPS>
__global__ void __launch_bounds__(NT)
k_rec1_runs(const T *__restrict__ coeffs, T *__restrict__ x, int64_t n1, int64_t batch, int J, int tile, int wb, int vec,
            const Taps1<T, L> tp)
{
    using C = Casc<T, L>;
    constexpr int VEC = C::VEC, R = CPT * VEC, WCH = C::WCH_R, NS = WCH + SPS - 1, D = C::HALO_R, SR = C::SR;
    extern __shared__ __align__(16) unsigned char smem1_raw[];
    T *a0 = reinterpret_cast<T *>(smem1_raw);
    T *a1 = a0 + wb;
    T *dd0 = a1 + wb, *dd1 = dd0 + wb;
    const int tid = threadIdx.x;
    const int64_t b = blockIdx.y;
    const int64_t t0 = (int64_t)blockIdx.x * tile;
    // level-j buffers: origin global t0 - j * HLRP, extent tile + j * D
    stage_wrapped<T, NT>(a0, coeffs + b * n1, t0 - (int64_t)J * C::HLRP, tile + J * D, n1, tid, vec != 0);
    stage_wrapped<T, NT>((J & 1) ? dd1 : dd0, coeffs + (batch + b) * n1, t0 - (int64_t)J * C::HLRP, tile + J * D, n1, tid,
                         vec != 0);
    cp_async_commit();
    T *src = a0, *dst = a1;
    for (int j = J; j >= 1; --j) {
        // d_{j-1} (slot J-j+2) travels into the other detail buffer while this level computes
        if (j > 1)
            stage_wrapped<T, NT>(((j - 1) & 1) ? dd1 : dd0, coeffs + ((int64_t)(J - j + 2) * batch + b) * n1,
                                 t0 - (int64_t)(j - 1) * C::HLRP, tile + (j - 1) * D, n1, tid, vec != 0);
        cp_async_commit();
        cp_async_wait<1>();                            // everything but the group just committed has landed
        __syncthreads();
        const T *dj = (j & 1) ? dd1 : dd0;
        const int Wo = tile + (j - 1) * D;             // outputs a_{j-1}
        for (int o = tid * R; o < Wo; o += NT * R) {
            T wa[NS * VEC], wd[NS * VEC];
#pragma unroll
            for (int q = 0; q < NS; ++q) {
                ld16<T, VEC>(src + o + q * VEC, wa + q * VEC);
                ld16<T, VEC>(dj + o + q * VEC, wd + q * VEC);
            }
#pragma unroll
            for (int c = 0; c < CPT; c += SPS) {
                const int n = (CPT - c < SPS) ? CPT - c : SPS;
                if (c > 0) {
#pragma unroll
                    for (int q = 0; q < SPS; ++q)
                        if (q < n) {
                            ld16<T, VEC>(src + o + (c + WCH - 1 + q) * VEC, wa + ((c + WCH - 1 + q) % NS) * VEC);
                            ld16<T, VEC>(dj + o + (c + WCH - 1 + q) * VEC, wd + ((c + WCH - 1 + q) % NS) * VEC);
                        }
                }
                T acc[SPS * VEC], acd[SPS * VEC];
#pragma unroll
                for (int r = 0; r < SPS * VEC; ++r) { acc[r] = zero_of(T()); acd[r] = zero_of(T()); }
#pragma unroll
                for (int k = 0; k < L; ++k)            // ascending taps = oldest source element first
#pragma unroll
                    for (int r = 0; r < SPS * VEC; ++r)
                        if (r < n * VEC) {
                            const int idx = c * VEC + r + SR + k;
                            mac1(acc[r], tp.lo[k], wa[((idx / VEC) % NS) * VEC + idx % VEC]);
                            mac1(acd[r], tp.hi[k], wd[((idx / VEC) % NS) * VEC + idx % VEC]);
                        }
#pragma unroll
                for (int r = 0; r < SPS * VEC; ++r)
                    if (r < n * VEC) acc[r] = add(acc[r], acd[r]);
#pragma unroll
                for (int q = 0; q < SPS; ++q)
                    if (q < n) st16<T, VEC>(dst + o + (c + q) * VEC, acc + q * VEC);
            }
        }
        __syncthreads();
        T *t = src; src = dst; dst = t;
    }
    store_tile<T, NT>(x + b * n1 + t0, src, (int)((n1 - t0 < tile) ? n1 - t0 : tile), tid, vec != 0);
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

#ifdef NDDWT_TUNING
static int tuning1_env(const char *name, int dflt)
{
    const char *e = getenv(name);
    return e ? atoi(e) : dflt;
}
#endif

// run kernels: NT threads, runs of CPT chunks; the tile is sized so that the widest level of the cascade is
// exactly `kit` iterations of the CTA
template <typename T, int L, int NT, int CPT, int SPS>
static int launch1_runs(nddwt_plan *p, bool rec, const void *in, void *out, int J, cudaStream_t s)
{
    using C = Casc<T, L>;
    constexpr int VEC = C::VEC, R = CPT * VEC;
    static_assert(CPT % 2 == 1, "odd chunk stride between lanes: conflict-free 16-byte shared-memory accesses");
    const int64_t n1 = p->dims[0], batch = p->batch;
    const int halo = rec ? C::HALO_R : C::HALO_A, wch = rec ? C::WCH_R : C::WCH_A;
    const int64_t grow = (int64_t)(J - 1) * halo;          // the widest level computes tile + grow outputs
    int kit = 1;
    while (kit < 8 && (int64_t)kit * NT * R - grow < 2 * grow + VEC) ++kit;
    int64_t tile = ((int64_t)kit * NT * R - grow) / VEC * VEC;
    if (tile < VEC) return 1;
    const int64_t n1r = (n1 + VEC - 1) / VEC * VEC;
    if (tile > n1r) tile = n1r;
    // buffers: a level's extent + what the last run may touch beyond it (its outputs, its window, the chunk ahead)
    const int64_t wb = (tile + (int64_t)J * halo + R + (int64_t)(wch + SPS + 1) * VEC + VEC - 1) / VEC * VEC;
    const size_t smem = (size_t)(rec ? 4 * wb : 2 * wb + 2 * tile + VEC) * sizeof(T);
    if (smem > 200 * 1024 || batch > 65535 || (n1 + tile - 1) / tile > 0x7fffffff) return 1;
    const int vec = (n1 % VEC == 0 && (uintptr_t)in % 16 == 0 && (uintptr_t)out % 16 == 0) ? 1 : 0;
    dim3 grid((unsigned)((n1 + tile - 1) / tile), (unsigned)batch);
    if (rec) {
        auto kern = k_rec1_runs<T, L, NT, CPT, SPS>;
        NDDWT_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        LaunchTimer lt(p, KIND_REC3, s);
        kern<<<grid, NT, smem, s>>>(reinterpret_cast<const T *>(in), reinterpret_cast<T *>(out), n1, batch, J, (int)tile,
                                    (int)wb, vec, make_taps1<T, L>(p, true));
    } else {
        auto kern = k_dec1_runs<T, L, NT, CPT, SPS>;
        NDDWT_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        LaunchTimer lt(p, KIND_DEC3, s);
        kern<<<grid, NT, smem, s>>>(reinterpret_cast<const T *>(in), reinterpret_cast<T *>(out), n1, batch, J, (int)tile,
                                    (int)wb, vec, make_taps1<T, L>(p, false));
    }
    p->launches++;
    NDDWT_CUDA(cudaGetLastError());
    return 0;
}

template <typename T, int L>
static int launch1_tile(nddwt_plan *p, bool rec, const void *in, void *out, int J, cudaStream_t s)
{
#ifdef NDDWT_TUNING
    if constexpr (L == 16 && sizeof(T) == 8 && Elem<T>::cplx) {
        switch (tuning1_env("NDDWT_CASC", 0)) {
            case 2: return launch1_runs<T, L, 128, 7, 2>(p, rec, in, out, J, s);   // two chunks per group: 4.33 + 4.56 ms (cfg2)
            case 5: return launch1_runs<T, L, 128, 5, 2>(p, rec, in, out, J, s);   // 5 CTAs per SM: 4.35 + 4.88 ms
            default: break;
        }
    }
#endif
    // 128 threads x runs of 7 chunks, one chunk per group: 4.32 + 4.47 ms on cfg2 (256 threads: 4.9 + 6.0; runs of 9 / 11
    // chunks, groups of 3 / 4 chunks: within 5 %, profiles/r02_variants.md)
    return launch1_runs<T, L, 128, 7, 1>(p, rec, in, out, J, s);
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
