// nddwt_fused.cu -- fused per-level kernels for 3-D and 4-D arrays on sm_100a: one launch reads the
// approximation band once and writes all 2^3 subbands once (analysis), or reads the subbands once
// and writes the reconstructed band once (synthesis).  Direct separable circular lo/hi filtering;
// no FFT, no stored Fourier-domain filters, no tensor cores (short-filter stencil).
//
// Replaces, per level: nd_dwt_dec_1level / nd_dwt_rec_1level (mex/nddwt.c:98-186) and
// level_1_dec / level_1_rec of Functions/nd_dwt_3D.m:345-393 and nd_dwt_4D.m:394-467.
//
// k_dec3_fused  (analysis tile kernel, also the back end of the 4-D path)
//   * a CTA owns a T1 x T2 column of the (dim1, dim2) plane and marches along dim 3;
//   * stage A (dim 3): each thread keeps an L-deep ring of input planes in REGISTERS for its
//     (T1+L-1) x (T2+L-1) haloed positions; the next plane is loaded a full step ahead straight
//     from global memory (coalesced along dim 1, periodic wrap folded into precomputed offsets);
//   * stage B (dim 1): 16-byte shared-memory chunks, lanes run along rows, odd chunk pitch =>
//     conflict-free; each thread produces 2*VEC consecutive outputs for lo1 / hi1;
//   * stage C (dim 2): lanes run along dim 1, sliding register window down the rows, the 8
//     subbands leave through 16-byte st.global.cs streaming stores;
//   * complex-single arithmetic is FFMA2 (fma.rn.f32x2) on (re, im) pairs with duplicated taps
//     taken from the kernel-parameter constant bank; all index math is hoisted out of the march.
// k_rec3_bulk   (synthesis tile kernel): TMA-staged (cp.async.bulk.tensor / cp.async.bulk + mbarrier)
//   haloed subband tiles, stages RA (dim 2) / RB (dim 1) / RC (dim 3, scatter ring of partial sums).
// k_rec3_rows   (synthesis, full-row tiles for 4-D batches of 8-byte elements): contiguous band-pair stages in an
//   mbarrier ring, 4.65 -> 3.47 ms per launch on cfg5 (profiles/r01_tile_probe.md, r01_rows_kernel.md).
// k_rec3_fused  : the same pipeline with direct ld.global.nc loads (rows narrower than a staged tile).
// k_dec_last / k_rec_last[_scatter] : dim-4 passes of the 4-D path and slab-exchange points (multi-GPU).
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <cuda.h>
#include "nddwt_plan.h"

namespace nddwt {

// ---------------------------------------------------------------------------------------------
// tap storage in kernel parameters: complex single keeps each tap duplicated (t, t) so that one
// FFMA2 updates (re, im) at once.
template <typename T> struct TapOf { using type = typename Elem<T>::R; };
template <> struct TapOf<float2> { using type = float2; };

__device__ __forceinline__ void macp(float &acc, float g, float v) { acc = fmaf(g, v, acc); }
__device__ __forceinline__ void macp(double &acc, double g, double v) { acc = fma(g, v, acc); }
__device__ __forceinline__ void macp(float2 &acc, float2 g, float2 v) { acc = __ffma2_rn(g, v, acc); }
__device__ __forceinline__ void macp(double2 &acc, double g, double2 v)
{
    acc.x = fma(g, v.x, acc.x);
    acc.y = fma(g, v.y, acc.y);
}

static inline float mk_tap(float, double v) { return (float)v; }
static inline double mk_tap(double, double v) { return v; }
static inline float2 mk_tap(float2, double v) { return make_float2((float)v, (float)v); }

template <typename T, int L>
struct FusedTaps {
    typename TapOf<T>::type lo[3][L];
    typename TapOf<T>::type hi[3][L];
};

template <typename T>
struct Dec3Params {
    const T *in[2];     // 3-D: in[0] = local planes [n3][n2][n1]; 4-D back end: lo4 / hi4 arrays
    const T *halo_lo;   // planes below the slab (ascending), or nullptr   (3-D slabs only)
    const T *halo_hi;   // planes above the slab, or nullptr
    T *out[16];         // subband bases (8 per input array)
    int n1, n2, n3;
    int64_t s3;         // plane stride (elements)
    int64_t s4;         // hyperplane stride (4-D batches)
    int nhyp;           // hyperplanes per input array (1 for 3-D)
    int tiles1, tiles2, zc, nchunks;
    int halo_below;     // planes held by halo_lo
    int band0 = 0;      // band index of out[0] within the level (8 when a part-wise 4-D launch handles the hi4 half)
    int zbase = 0, zcount = 0;   // dim-3 sub-range produced by this launch (zcount 0: all n3 planes); the
                                 // ring still reads its L-1 neighbour planes around the range, periodic in n3
    typename Elem<T>::R thr[16]; // SHR instantiations: soft threshold per output band (0 = keep)
};

__device__ __forceinline__ int wrapi(int m, int n)
{
    m %= n;
    return m < 0 ? m + n : m;
}

template <typename T> __device__ __forceinline__ T ldg_stream(const T *p) { return __ldg(p); }

template <typename T> __device__ __forceinline__ void st_stream(T *p, T v) { __stcs(p, v); }

template <typename T, int VEC>
__device__ __forceinline__ void ld_chunk(const T *p, T *dst)
{
    const uint4 q = *reinterpret_cast<const uint4 *>(p);
    union { uint4 u; T t[VEC]; } cv;
    cv.u = q;
#pragma unroll
    for (int i = 0; i < VEC; ++i) dst[i] = cv.t[i];
}

template <typename T, int VEC>
__device__ __forceinline__ void st_chunk(T *p, const T *src)
{
    union { uint4 u; T t[VEC]; } cv;
#pragma unroll
    for (int i = 0; i < VEC; ++i) cv.t[i] = src[i];
    *reinterpret_cast<uint4 *>(p) = cv.u;
}

// ---- compile-time geometry ---------------------------------------------------------------
template <typename T, int L, int T2>
struct Geo {
    static constexpr int VEC = 16 / (int)sizeof(T);   // elements per 16-byte chunk
    static constexpr int R1 = 2 * VEC;                // dim-1 outputs per stage-B item (two chunks)
    static constexpr int T1 = 8 * R1;                 // tile width = 16 chunks
    static constexpr int H = L - 1, HB = L / 2 - 1, HA = L / 2;
    static constexpr int W1 = T1 + H, W2 = T2 + H;
    static constexpr int W2P = (W2 + 7) & ~7;         // rows padded to whole quarter-warps (stage B lanes run along rows)
    static constexpr int PAC = ((W1 + VEC - 1) / VEC) | 1;   // SA row pitch in chunks, odd -> conflict-free column access
    static constexpr int PA = PAC * VEC;
    static constexpr int PBC = 17;                    // SB row pitch in chunks (16 + 1 pad, odd)
    static constexpr int PB = PBC * VEC;
    static constexpr int NPOS = W1 * W2;
    static constexpr int NCH = (R1 + L - 1 + VEC - 1) / VEC;   // chunks read per stage-B item
    static constexpr size_t SMEM = (size_t)(2 * W2 * PA + 4 * W2 * PB) * sizeof(T);
};

// ---- stage A for ring phase U (all ring indices static -> the ring stays in registers) ----
template <typename T, int L, int PPT, int U>
__device__ __forceinline__ void dec_stage_a(T (&ring)[PPT][L], const typename TapOf<T>::type *lo,
                                            const typename TapOf<T>::type *hi, T *sa_lo, T *sa_hi,
                                            const int (&sa_off)[PPT], const int (&g_off)[PPT], int npos_mask,
                                            const T *next_plane)
{
#pragma unroll
    for (int k = 0; k < PPT; ++k) {
        T alo = zero_of(T()), ahi = zero_of(T());
#pragma unroll
        for (int j = 0; j < L; ++j) {
            const T v = ring[k][(U + j) % L];
            macp(alo, lo[L - 1 - j], v);
            macp(ahi, hi[L - 1 - j], v);
        }
        if (npos_mask & (1 << k)) {
            sa_lo[sa_off[k]] = alo;
            sa_hi[sa_off[k]] = ahi;
        }
    }
    if (next_plane != nullptr) {
#pragma unroll
        for (int k = 0; k < PPT; ++k)
            if (npos_mask & (1 << k)) ring[k][U] = ldg_stream(next_plane + g_off[k]);
    }
}

template <typename T, int L, int PPT, int U>
struct DispatchA {
    __device__ __forceinline__ static void run(int u, T (&ring)[PPT][L], const typename TapOf<T>::type *lo,
                                               const typename TapOf<T>::type *hi, T *sa_lo, T *sa_hi,
                                               const int (&sa_off)[PPT], const int (&g_off)[PPT], int mask,
                                               const T *next_plane)
    {
        if (u == U) dec_stage_a<T, L, PPT, U>(ring, lo, hi, sa_lo, sa_hi, sa_off, g_off, mask, next_plane);
        else DispatchA<T, L, PPT, U + 1>::run(u, ring, lo, hi, sa_lo, sa_hi, sa_off, g_off, mask, next_plane);
    }
};
template <typename T, int L, int PPT>
struct DispatchA<T, L, PPT, L> {
    __device__ __forceinline__ static void run(int, T (&)[PPT][L], const typename TapOf<T>::type *,
                                               const typename TapOf<T>::type *, T *, T *, const int (&)[PPT],
                                               const int (&)[PPT], int, const T *) {}
};

template <typename T, int L, int T2, int NT, int R2, int MINB, int CWSEL = 0, int RBM = 1, int ZINC = 0, int SHR = 0>
__global__ void __launch_bounds__(NT, MINB)
k_dec3_fused(const Dec3Params<T> p, const FusedTaps<T, L> tp)
{
    using G = Geo<T, L, T2>;
    constexpr int VEC = G::VEC, T1 = G::T1, HB = G::HB, HA = G::HA;
    constexpr int R1 = RBM * G::R1;                         // dim-1 outputs per stage-B item (RBM = 2: half the items, 2/3 of the shared-memory reads)
    constexpr int NCB = T1 / R1;                            // stage-B items per row
    constexpr int NCH = (R1 + L - 1 + VEC - 1) / VEC;       // chunks read per stage-B item
    constexpr int W1 = G::W1, W2 = G::W2, W2P = G::W2P, PA = G::PA, PB = G::PB, NPOS = G::NPOS;
    constexpr int PPT = (NPOS + NT - 1) / NT;
    constexpr int NRUN = T2 / R2;
    constexpr int NB_ITEMS = 2 * NCB * W2P, KB = (NB_ITEMS + NT - 1) / NT;
    static_assert((NCB - 1) * R1 + NCH * VEC <= PA, "stage-B reads stay inside the SA row pitch");
    constexpr int CW = (CWSEL == 1) ? 1 : VEC;              // columns per stage-C item (16-byte chunk or one element)
    constexpr int CPR = T1 / CW;                            // stage-C items per row
    constexpr int NC_ITEMS = 4 * NRUN * CPR, KC = (NC_ITEMS + NT - 1) / NT;
    static_assert(T2 % R2 == 0, "R2 must divide T2");

    extern __shared__ __align__(128) unsigned char smem_raw[];
    T *SA = reinterpret_cast<T *>(smem_raw);     // [2][W2][PA]   lo3 / hi3 of the haloed tile
    T *SB = SA + 2 * W2 * PA;                    // [4][W2][PB]   (b1 + 2 b3), rows still haloed
    // fused coefficient shrink: thresholds in shared memory (indexing the parameter block with a register would
    // force a local-memory copy of it); first read after the two barriers of the first plane
    __shared__ typename Elem<T>::R s_thr[SHR ? 16 : 1];
    if (SHR && threadIdx.x < 16) s_thr[threadIdx.x] = p.thr[threadIdx.x];

    const int tid = threadIdx.x;
    int bid = blockIdx.x;
    const int t1 = bid % p.tiles1;
    bid /= p.tiles1;
    const int t2 = bid % p.tiles2;
    bid /= p.tiles2;
    const int chunk = bid % p.nchunks;
    const int batch = bid / p.nchunks;
    const int a1 = t1 * T1, a2 = t2 * T2;
    const int z0 = p.zbase + chunk * p.zc;
    const int z1 = min(z0 + p.zc, p.zbase + p.zcount);
    const int n1 = p.n1, n2 = p.n2, n3 = p.n3;
    const int64_t s3 = p.s3;
    const bool slab = (p.halo_lo != nullptr) || (p.halo_hi != nullptr);
    // batch (4-D back end): input array `bsel` at hyperplane `bhyp`; output bands 8*bsel.. at the same hyperplane
    const int bsel = batch / p.nhyp, bhyp = batch - bsel * p.nhyp;
    const T *in_base = p.in[bsel] + (int64_t)bhyp * p.s4;
    const int64_t out_boff = (int64_t)bhyp * p.s4;

    auto plane_ptr = [&](int zi) -> const T * {
        if (!slab) return in_base + (int64_t)wrapi(zi, n3) * s3;
        if (zi < 0) return p.halo_lo + (int64_t)(zi + p.halo_below) * s3;
        if (zi >= n3) return p.halo_hi + (int64_t)(zi - n3) * s3;
        return in_base + (int64_t)zi * s3;
    };

    // ---- per-thread constants, hoisted out of the marching loop -------------------------
    // stage A: haloed positions (global in-plane offset with the periodic wrap folded in, SA slot)
    int g_off[PPT], sa_off[PPT];
    int mask = 0;
#pragma unroll
    for (int k = 0; k < PPT; ++k) {
        const int q = tid + k * NT;
        const int r = q / W1, c = q - r * W1;
        if (q < NPOS) mask |= 1 << k;
        g_off[k] = wrapi(a2 - HB + r, n2) * n1 + wrapi(a1 - HB + c, n1);
        sa_off[k] = r * PA + c;
    }
    // stage B: item = (array a, chunk pair cb, row r), lanes run along rows
    int b_src[KB], b_dst[KB];
    int b_mask = 0;
#pragma unroll
    for (int k = 0; k < KB; ++k) {
        const int it = tid + k * NT;
        const int r = it % W2P, g = it / W2P;
        const int cb = g % NCB, a = g / NCB;
        if (it < NB_ITEMS && r < W2) b_mask |= 1 << k;
        b_src[k] = (a * W2 + r) * PA + cb * R1;
        b_dst[k] = (2 * a * W2 + r) * PB + cb * R1;
    }
    // stage C: item = (array m, run, chunk column cp); lanes run along dim 1
    int c_src[KC];
    T *c_lo[KC], *c_hi[KC];
    int c_rows[KC];
#pragma unroll
    for (int k = 0; k < KC; ++k) {
        const int it = tid + k * NT;
        const int cp = it % CPR, rest = it / CPR;
        const int run = rest % NRUN, m = rest / NRUN;
        c_src[k] = (m * W2 + run * R2) * PB + cp * CW;
        const int g1 = a1 + cp * CW, g2 = a2 + run * R2;
        int rows = min(R2, n2 - g2);
        if (it >= NC_ITEMS || g1 >= n1 || rows < 0) rows = 0;
        c_rows[k] = rows;
        const int mm = (it < NC_ITEMS) ? m : 0;
        const int64_t off = out_boff + (int64_t)z0 * s3 + (int64_t)g2 * n1 + g1;
        c_lo[k] = p.out[8 * bsel + (mm & 1) + 4 * (mm >> 1)] + off;
        c_hi[k] = p.out[8 * bsel + (mm & 1) + 4 * (mm >> 1) + 2] + off;
        // fused coefficient shrink: the item's band index rides in the upper bits of c_rows (no extra live register);
        // the thresholds are read from the parameter bank at the stores
        if (SHR) c_rows[k] |= (8 * bsel + (mm & 1) + 4 * (mm >> 1)) << 8;
    }

    // warm-up: planes z0-HB .. z0+HA fill ring slots 0..L-1
    T ring[PPT][L];
#pragma unroll
    for (int j = 0; j < L; ++j) {
        const T *pl = plane_ptr(z0 - HB + j);
#pragma unroll
        for (int k = 0; k < PPT; ++k) ring[k][j] = (mask & (1 << k)) ? ldg_stream(pl + g_off[k]) : zero_of(T());
    }

    T *sa_lo = SA, *sa_hi = SA + W2 * PA;
    int u = 0;
    // ZINC: the plane to prefetch advances by pointer increments (no integer modulo per plane; stage A spends
    // ~60 of its 181 instructions per plane and warp on plane addressing and loop control, ncu source view)
    int zn = (ZINC && !slab) ? wrapi(z0 + 1 + HA, n3) : 0;
    const T *np = in_base + (int64_t)zn * s3;
    for (int z = z0; z < z1; ++z) {
        const T *next_plane;
        if (ZINC && !slab) {
            next_plane = (z + 1 < z1) ? np : nullptr;
            if (++zn == n3) { zn = 0; np = in_base; } else np += s3;
        } else {
            next_plane = (z + 1 < z1) ? plane_ptr(z + 1 + HA) : nullptr;
        }
        DispatchA<T, L, PPT, 0>::run(u, ring, tp.lo[2], tp.hi[2], sa_lo, sa_hi, sa_off, g_off, mask, next_plane);
        u = (u + 1 == L) ? 0 : u + 1;
        __syncthreads();

        // ---- stage B: dim 1, SA[a][r][:] -> SB[b1 + 2a][r][:]
#pragma unroll
        for (int k = 0; k < KB; ++k) {
            if (b_mask & (1 << k)) {
                // SHR instantiations recompute the item offsets instead of keeping them live across the march (the
                // kernel sits at its 128-register cap and the shrink epilogue needs the room)
                int bs = b_src[k], bd = b_dst[k];
                if (SHR) {
                    int it = tid + k * NT;
                    asm volatile("" : "+r"(it));      // keeps the recomputation inside the loop (no re-hoisting)
                    const int r = it % W2P, g = it / W2P;
                    const int cb = g % NCB, a = g / NCB;
                    bs = (a * W2 + r) * PA + cb * R1;
                    bd = (2 * a * W2 + r) * PB + cb * R1;
                }
                const T *row = SA + bs;
                T v[NCH * VEC];
#pragma unroll
                for (int j = 0; j < NCH; ++j) ld_chunk<T, VEC>(row + j * VEC, v + j * VEC);
                T *d0 = SB + bd;
                // lo1 and hi1 together, tap loop outermost: 2 * R1 independent FFMA2 chains
                T acc[R1], ach[R1];
#pragma unroll
                for (int o = 0; o < R1; ++o) { acc[o] = zero_of(T()); ach[o] = zero_of(T()); }
#pragma unroll
                for (int j = 0; j < L; ++j)
#pragma unroll
                    for (int o = 0; o < R1; ++o) {
                        macp(acc[o], tp.lo[0][L - 1 - j], v[o + j]);
                        macp(ach[o], tp.hi[0][L - 1 - j], v[o + j]);
                    }
#pragma unroll
                for (int c = 0; c < R1 / VEC; ++c) st_chunk<T, VEC>(d0 + c * VEC, acc + c * VEC);
#pragma unroll
                for (int c = 0; c < R1 / VEC; ++c) st_chunk<T, VEC>(d0 + W2 * PB + c * VEC, ach + c * VEC);
            }
        }
        __syncthreads();

        // ---- stage C: dim 2, SB[m][rows][cols] -> subbands in global memory (streaming stores).
        // Sliding L-row register window per item (row r lives in slot r % L, the next row is loaded one
        // output ahead), so the window costs the same registers for any run length R2.
#pragma unroll
        for (int k = 0; k < KC; ++k) {
            const int rows = SHR ? (c_rows[k] & 0xff) : c_rows[k];
            if (rows > 0) {
                int csrc = c_src[k];
                if (SHR) {
                    int it = tid + k * NT;
                    asm volatile("" : "+r"(it));
                    const int cp = it % CPR, rest = it / CPR;
                    csrc = ((rest / NRUN) * W2 + (rest % NRUN) * R2) * PB + cp * CW;
                }
                const T *col = SB + csrc;
                T w[L][CW];
#pragma unroll
                for (int j = 0; j < L; ++j) {
                    if (CW == 1) w[j][0] = col[j * PB];
                    else ld_chunk<T, CW>(col + j * PB, w[j]);
                }
                T *plo = c_lo[k], *phi = c_hi[k];
#pragma unroll
                for (int o = 0; o < R2; ++o) {
                    T lo[CW], hi[CW];
#pragma unroll
                    for (int e = 0; e < CW; ++e) {
                        lo[e] = zero_of(T());
                        hi[e] = zero_of(T());
#pragma unroll
                        for (int j = 0; j < L; ++j) {
                            macp(lo[e], tp.lo[1][L - 1 - j], w[(o + j) % L][e]);
                            macp(hi[e], tp.hi[1][L - 1 - j], w[(o + j) % L][e]);
                        }
                    }
                    if (o + 1 < R2) {   // row o + L replaces row o (its last use was this output)
                        if (CW == 1) w[o % L][0] = col[(o + L) * PB];
                        else ld_chunk<T, CW>(col + (o + L) * PB, w[o % L]);
                    }
                    if (o < rows) {
                        if (SHR) {
                            const typename Elem<T>::R tl = s_thr[c_rows[k] >> 8], th = s_thr[(c_rows[k] >> 8) + 2];
#pragma unroll
                            for (int e = 0; e < CW; ++e) {     // one element at a time: the kernel sits at its register cap
                                lo[e] = shrink1(lo[e], tl);
                                asm volatile("" ::: "memory");
                                hi[e] = shrink1(hi[e], th);
                                asm volatile("" ::: "memory");
                            }
                        }
                        if (CW == 1) {
                            st_stream(plo, lo[0]);
                            st_stream(phi, hi[0]);
                        } else {
                            union { uint4 q; T t[CW]; } a, b;
#pragma unroll
                            for (int e = 0; e < CW; ++e) { a.t[e] = lo[e]; b.t[e] = hi[e]; }
                            __stcs(reinterpret_cast<uint4 *>(plo), a.q);
                            __stcs(reinterpret_cast<uint4 *>(phi), b.q);
                        }
                    }
                    plo += n1;
                    phi += n1;
                    asm volatile("" ::: "memory");   // keep the window loads of later rows from being hoisted (register pressure)
                }
            }
            c_lo[k] += s3;
            c_hi[k] += s3;
        }
        // SA is rewritten by the next stage A only after the barrier that follows stage B of this
        // step; SB is rewritten after the barrier that follows the next stage A: no third barrier.
    }
}


// ---------------------------------------------------------------------------------------------
// Synthesis kernel (3-D, also the back end of the 4-D path): reads the 8 subbands once, writes the
// reconstructed band once; the band sum and the 2^-d normalisation (folded into the taps) are fused.
//   stage RA (dim 2): straight from global memory -- lanes run along dim 1 (coalesced), each thread
//            slides a register window down R2 rows for the (b2 = 0, b2 = 1) pair of one (b1, b3) group;
//   stage RB (dim 1): 16-byte shared-memory chunks, lanes run along rows (odd chunk pitch);
//   stage RC (dim 3): scatter form -- each thread owns one 16-byte output chunk and an L-deep ring of
//            partial sums in registers; a plane is stored (st.global.cs) when its last tap has arrived.
template <typename T>
struct Rec3Params {
    const T *in[16];    // subbands (8 per output array)
    T *out[2];          // 3-D: out[0]; 4-D back end: u_lo / u_hi
    int n1, n2, n3;
    int64_t s3, s4;
    int nhyp;
    int tiles1, tiles2, zc, nchunks;
    int prefetch;       // 0 none, 1 prefetch.global.L1 of the next plane's footprint, 2 prefetch.global.L2
    int cl1, cl2;       // thread-block cluster shape in tiles (cl1 x cl2 CTAs, 1 x 1 = no cluster)
    int hint;           // 1: stage the subband tiles with an L2 evict_last policy
    int zbase = 0, zcount = 0;   // dim-3 sub-range of OUTPUT planes produced by this launch (zcount 0: all n3)
};

template <typename T, int L, int T2>
struct GeoR {
    static constexpr int VEC = 16 / (int)sizeof(T);
    static constexpr int R1 = 2 * VEC;
    static constexpr int T1 = 8 * R1;
    static constexpr int H = L - 1, HB = L / 2, HA = L / 2 - 1;   // synthesis reads n-L/2 .. n+L/2-1
    static constexpr int W1 = T1 + H, W2 = T2 + H;
    static constexpr int PUC = ((W1 + VEC - 1) / VEC) | 1;
    static constexpr int PU = PUC * VEC;          // U pitch (elements), odd chunk count
    static constexpr int PV = 17 * VEC;           // V pitch
    static constexpr int NCH = (R1 + L - 1 + VEC - 1) / VEC;
    static constexpr size_t SMEM = (size_t)(4 * T2 * PU + 2 * T2 * PV) * sizeof(T);
};

template <typename T, int L, int VEC, int U>
__device__ __forceinline__ void rec_stage_c(T (&acc)[L][VEC], const T (&v0)[VEC], const T (&v1)[VEC],
                                            const typename TapOf<T>::type *lo, const typename TapOf<T>::type *hi,
                                            T *dst, bool store, bool rmw = false, int nel = -1)
{
    // nel >= 0: rows that are not 16-byte multiples -- the completed chunk leaves element by element, nel of them
    // coefficient plane t (phase U = t mod L) feeds output planes n = z0 + t - k, slot (U - k) mod L
#pragma unroll
    for (int k = 0; k < L; ++k) {
#pragma unroll
        for (int e = 0; e < VEC; ++e) {
            macp(acc[(U - k + L) % L][e], lo[k], v0[e]);
            macp(acc[(U - k + L) % L][e], hi[k], v1[e]);
        }
    }
    // plane n = z0 + t - (L-1) is complete: slot (U + 1) mod L
    if (store && nel >= 0) {
#pragma unroll
        for (int e = 0; e < VEC; ++e)
            if (e < nel) st_stream(dst + e, acc[(U + 1) % L][e]);
    } else if (store) {
        union { uint4 q; T t[VEC]; } a;
#pragma unroll
        for (int e = 0; e < VEC; ++e) a.t[e] = acc[(U + 1) % L][e];
        if (rmw) {   // periodic closure: add the partial sum this thread stored at the start of the march
            union { uint4 q; T t[VEC]; } b;
            b.q = __ldcg(reinterpret_cast<const uint4 *>(dst));   // L2-coherent load of this thread's own earlier store
#pragma unroll
            for (int e = 0; e < VEC; ++e) a.t[e] = add(a.t[e], b.t[e]);
        }
        __stcs(reinterpret_cast<uint4 *>(dst), a.q);
    }
#pragma unroll
    for (int e = 0; e < VEC; ++e) acc[(U + 1) % L][e] = zero_of(T());
}

template <typename T, int L, int VEC, int U>
struct DispatchC {
    __device__ __forceinline__ static void run(int u, T (&acc)[L][VEC], const T (&v0)[VEC], const T (&v1)[VEC],
                                               const typename TapOf<T>::type *lo, const typename TapOf<T>::type *hi,
                                               T *dst, bool store, bool rmw = false, int nel = -1)
    {
        if (u == U) rec_stage_c<T, L, VEC, U>(acc, v0, v1, lo, hi, dst, store, rmw, nel);
        else DispatchC<T, L, VEC, U + 1>::run(u, acc, v0, v1, lo, hi, dst, store, rmw, nel);
    }
};
template <typename T, int L, int VEC>
struct DispatchC<T, L, VEC, L> {
    __device__ __forceinline__ static void run(int, T (&)[L][VEC], const T (&)[VEC], const T (&)[VEC],
                                               const typename TapOf<T>::type *, const typename TapOf<T>::type *, T *,
                                               bool, bool = false, int = -1) {}
};

template <typename T, int L, int T2, int NT, int R2, int MINB, int EDGE = 0>
__global__ void __launch_bounds__(NT, MINB)
k_rec3_fused(const Rec3Params<T> p, const FusedTaps<T, L> tp)
{
    using G = GeoR<T, L, T2>;
    constexpr int VEC = G::VEC, R1 = G::R1, T1 = G::T1, HB = G::HB, H = G::H;
    constexpr int W1 = G::W1, PU = G::PU, PV = G::PV, NCH = G::NCH;
    constexpr int NRUN = T2 / R2;
    constexpr int NA_ITEMS = 4 * NRUN * W1, KA = (NA_ITEMS + NT - 1) / NT;
    constexpr int NB_ITEMS = 2 * 8 * T2, KB = (NB_ITEMS + NT - 1) / NT;
    constexpr int NC_ITEMS = T2 * 16, KC = (NC_ITEMS + NT - 1) / NT;
    static_assert(T2 % R2 == 0 && T2 % 8 == 0, "tile rows");

    extern __shared__ __align__(128) unsigned char smem_raw[];
    T *SU = reinterpret_cast<T *>(smem_raw);     // [4][T2][PU]  (b1 + 2 b3), dim 2 synthesised
    T *SV = SU + 4 * T2 * PU;                    // [2][T2][PV]  (b3), dims 1,2 synthesised

    const int tid = threadIdx.x;
    int bid = blockIdx.x;
    const int t1 = bid % p.tiles1;
    bid /= p.tiles1;
    const int t2 = bid % p.tiles2;
    bid /= p.tiles2;
    const int chunk = bid % p.nchunks;
    const int batch = bid / p.nchunks;
    const int a1 = t1 * T1, a2 = t2 * T2;
    const int z0 = p.zbase + chunk * p.zc;
    const int z1 = min(z0 + p.zc, p.zbase + p.zcount);
    const int n1 = p.n1, n2 = p.n2, n3 = p.n3;
    const int64_t s3 = p.s3;
    const int bsel = batch / p.nhyp, bhyp = batch - bsel * p.nhyp;
    const int64_t boff = (int64_t)bhyp * p.s4;
    const bool rows_interior = (a2 - HB >= 0) && (a2 - HB + T2 + H <= n2);

    // ---- hoisted per-thread constants ----
    // stage RA: item = (q = b1 + 2 b3, run, haloed column c)
    const T *a_lo[KA], *a_hi[KA];
    int a_dst[KA], a_row0[KA];
    int a_mask = 0;
#pragma unroll
    for (int k = 0; k < KA; ++k) {
        const int it = tid + k * NT;
        const int c = it % W1, g = it / W1;
        const int run = g % NRUN, q = (it < NA_ITEMS) ? g / NRUN : 0;
        if (it < NA_ITEMS) a_mask |= 1 << k;
        const int blo = 8 * bsel + (q & 1) + 4 * (q >> 1);
        const int gcol = wrapi(a1 - HB + c, n1);
        a_row0[k] = a2 - HB + run * R2;                 // first (unwrapped) global row of the window
        a_lo[k] = p.in[blo] + boff + gcol;
        a_hi[k] = p.in[blo + 2] + boff + gcol;
        a_dst[k] = (q * T2 + run * R2) * PU + c;
    }
    // stage RB: item = (b3, chunk pair cb, row j), lanes run along rows
    int b_src[KB], b_dst[KB];
    int b_mask = 0;
#pragma unroll
    for (int k = 0; k < KB; ++k) {
        const int it = tid + k * NT;
        const int j = it % T2, g = it / T2;
        const int cb = g & 7, b3 = (g >> 3) & 1;
        if (it < NB_ITEMS) b_mask |= 1 << k;
        b_src[k] = (2 * b3 * T2 + j) * PU + cb * R1;
        b_dst[k] = (b3 * T2 + j) * PV + cb * R1;
    }
    // stage RC: item = output chunk (row j, chunk column cp)
    int c_src[KC];
    T *c_out[KC];
    bool c_ok[KC];
    int c_nel[EDGE ? KC : 1];
    T acc[KC][L][VEC];
#pragma unroll
    for (int k = 0; k < KC; ++k) {
        const int it = tid + k * NT;
        const int cp = it & 15, j = it >> 4;
        c_src[k] = j * PV + cp * VEC;
        const int g1 = a1 + cp * VEC, g2 = a2 + j;
        c_ok[k] = (it < NC_ITEMS) && g1 < n1 && g2 < n2;
        if (EDGE) c_nel[k] = min(VEC, n1 - g1);     // rows that are not 16-byte multiples: element-wise, guarded stores
        // pointer to output plane (z0 - (L-1)): advanced by s3 per step, first stored at step t = L-1
        c_out[k] = p.out[bsel] + boff + ((int64_t)z0 - (L - 1)) * s3 + (int64_t)g2 * n1 + g1;
#pragma unroll
        for (int s = 0; s < L; ++s)
#pragma unroll
            for (int e = 0; e < VEC; ++e) acc[k][s][e] = zero_of(T());
    }

    // next-plane prefetch: one 128-byte line per (band, haloed row, segment), spread over the threads
    constexpr int SEG = (W1 * (int)sizeof(T) + 127) / 128 + 1;
    constexpr int NPF = 8 * (T2 + H) * SEG, KP = (NPF + NT - 1) / NT;
    const T *pf_ptr[KP];
#pragma unroll
    for (int k = 0; k < KP; ++k) {
        const int it = tid + k * NT;
        const int sg = it % SEG, g = it / SEG;
        const int r = g % (T2 + H), b = (it < NPF) ? g / (T2 + H) : 0;
        const int col = wrapi(a1 - HB + sg * (128 / (int)sizeof(T)), n1);
        pf_ptr[k] = (it < NPF && sg * (128 / (int)sizeof(T)) < W1 + (128 / (int)sizeof(T)))
                        ? p.in[8 * bsel + b] + boff + (int64_t)wrapi(a2 - HB + r, n2) * n1 + col
                        : nullptr;
    }

    const int nsteps = (z1 - z0) + L - 1;
    int u = 0;
    for (int t = 0; t < nsteps; ++t) {
        const int zc = wrapi(z0 - HB + t, n3);
        const int64_t zoff = (int64_t)zc * s3;
        if (p.prefetch && t + 1 < nsteps) {
            const int64_t znext = (int64_t)wrapi(z0 - HB + t + 1, n3) * s3;
#pragma unroll
            for (int k = 0; k < KP; ++k) {
                if (pf_ptr[k] != nullptr) {
                    if (p.prefetch == 1) asm volatile("prefetch.global.L1 [%0];" ::"l"(pf_ptr[k] + znext));
                    else asm volatile("prefetch.global.L2 [%0];" ::"l"(pf_ptr[k] + znext));
                }
            }
        }

        // ---- stage RA: dim 2 from global memory
#pragma unroll
        for (int k = 0; k < KA; ++k) {
            if (a_mask & (1 << k)) {
                T o[R2];
#pragma unroll
                for (int i = 0; i < R2; ++i) o[i] = zero_of(T());
#pragma unroll
                for (int hb = 0; hb < 2; ++hb) {
                    const T *src = (hb ? a_hi[k] : a_lo[k]) + zoff;
                    const typename TapOf<T>::type *g = hb ? tp.hi[1] : tp.lo[1];
                    T w[R2 + L - 1];
                    if (rows_interior) {
#pragma unroll
                        for (int i = 0; i < R2 + L - 1; ++i) w[i] = ldg_stream(src + (int64_t)(a_row0[k] + i) * n1);
                    } else {
#pragma unroll
                        for (int i = 0; i < R2 + L - 1; ++i) w[i] = ldg_stream(src + (int64_t)wrapi(a_row0[k] + i, n2) * n1);
                    }
#pragma unroll
                    for (int i = 0; i < R2; ++i)
#pragma unroll
                        for (int kk = 0; kk < L; ++kk) macp(o[i], g[kk], w[i + kk]);
                }
                T *dst = SU + a_dst[k];
#pragma unroll
                for (int i = 0; i < R2; ++i) dst[i * PU] = o[i];
            }
        }
        __syncthreads();

        // ---- stage RB: dim 1, SU[b1 + 2 b3][j][:] -> SV[b3][j][:]
#pragma unroll
        for (int k = 0; k < KB; ++k) {
            if (b_mask & (1 << k)) {
                T o[R1];
#pragma unroll
                for (int i = 0; i < R1; ++i) o[i] = zero_of(T());
#pragma unroll
                for (int hb = 0; hb < 2; ++hb) {
                    const T *row = SU + b_src[k] + hb * T2 * PU;
                    const typename TapOf<T>::type *g = hb ? tp.hi[0] : tp.lo[0];
                    T v[NCH * VEC];
#pragma unroll
                    for (int j = 0; j < NCH; ++j) ld_chunk<T, VEC>(row + j * VEC, v + j * VEC);
#pragma unroll
                    for (int i = 0; i < R1; ++i)
#pragma unroll
                        for (int kk = 0; kk < L; ++kk) macp(o[i], g[kk], v[i + kk]);
                }
                st_chunk<T, VEC>(SV + b_dst[k], o);
                st_chunk<T, VEC>(SV + b_dst[k] + VEC, o + VEC);
            }
        }
        __syncthreads();

        // ---- stage RC: dim 3, scatter into the register ring of partial sums
        const bool store = (t >= L - 1);
#pragma unroll
        for (int k = 0; k < KC; ++k) {
            if (tid + k * NT < NC_ITEMS) {
                T v0[VEC], v1[VEC];
                ld_chunk<T, VEC>(SV + c_src[k], v0);
                ld_chunk<T, VEC>(SV + c_src[k] + T2 * PV, v1);
                DispatchC<T, L, VEC, 0>::run(u, acc[k], v0, v1, tp.lo[2], tp.hi[2], c_out[k], store && c_ok[k], false,
                                             EDGE ? c_nel[k] : -1);
                c_out[k] += s3;
            }
        }
        u = (u + 1 == L) ? 0 : u + 1;
        // SU is rewritten by the next stage RA only after both barriers of this step; SV after the
        // barrier that follows the next stage RA.
    }
}


// ---------------------------------------------------------------------------------------------
// Synthesis kernel, bulk-copy (TMA engine) staged variant.  The haloed tiles of the 8 subbands of
// the NEXT coefficient plane are brought into shared memory by cp.async.bulk (UBLKCP) row copies
// that complete on an mbarrier, while stages RB / RC of the current plane run.  The periodic
// boundary costs nothing: a row that crosses the array edge is issued as two copies.  Stage RA then
// slides its window over shared memory with immediate offsets (no address arithmetic, no exposed
// global-load latency).
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE;\n"
        "bra WAIT_LOOP;\n"
        "DONE:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ uint64_t l2_policy_evict_last()
{
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ void bulk_g2s_hint(void *dst, const void *src, uint32_t bytes, uint64_t *bar, uint64_t pol)
{
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
            smem_u32(dst)),
        "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(pol)
        : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

template <typename T, int L, int T2>
struct GeoRB {
    static constexpr int VEC = 16 / (int)sizeof(T);
    static constexpr int R1 = 2 * VEC;
    static constexpr int T1 = 8 * R1;
    static constexpr int H = L - 1, HB = L / 2, HA = L / 2 - 1;
    static constexpr int HBAL = (HB + VEC - 1) / VEC * VEC;              // staged rows start 16-byte aligned
    static constexpr int SHIFT = HBAL - HB;                              // column of the first needed element
    static constexpr int W1S = (HBAL + T1 + HA + VEC - 1) / VEC * VEC;   // staged row width (elements)
    static constexpr int W1 = T1 + H, W2 = T2 + H;
    static constexpr int PUC = ((W1 + VEC - 1) / VEC) | 1;
    static constexpr int PU = PUC * VEC;
    static constexpr int PV = 17 * VEC;
    static constexpr int NCH = (R1 + L - 1 + VEC - 1) / VEC;
    static constexpr int BP = (int)(((size_t)W2 * W1S * sizeof(T) + 127) / 128 * 128 / sizeof(T));   // band pitch, 128-byte multiple (TMA destination)
    static constexpr size_t RAW_ELEMS = (size_t)8 * BP;
    static constexpr uint32_t PLANE_BYTES = (uint32_t)((size_t)8 * W2 * W1S * sizeof(T));
    static constexpr size_t SMEM = (RAW_ELEMS + 4 * T2 * PU + 2 * T2 * PV) * sizeof(T) + 16;
};

struct TmaMaps {
    CUtensorMap m[16];      // one 3-D map (n1, n2, planes) per subband array; box = (W1S, W2, 1)
};

__device__ __forceinline__ void tma_load_3d_hint(void *dst, const CUtensorMap *map, int x, int y, int z, uint64_t *bar,
                                                 uint64_t pol)
{
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%2, %3, %4}], [%5], %6;" ::"r"(
            smem_u32(dst)),
        "l"(map), "r"(x), "r"(y), "r"(z), "r"(smem_u32(bar)), "l"(pol)
        : "memory");
}
__device__ __forceinline__ void tma_load_3d(void *dst, const CUtensorMap *map, int x, int y, int z, uint64_t *bar)
{
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(
            smem_u32(dst)),
        "l"(map), "r"(x), "r"(y), "r"(z), "r"(smem_u32(bar))
        : "memory");
}

template <typename T, int L, int T2, int NT, int R2, int MINB>
__global__ void __launch_bounds__(NT, MINB)
k_rec3_bulk(const Rec3Params<T> p, const FusedTaps<T, L> tp, const __grid_constant__ TmaMaps maps)
{
    using G = GeoRB<T, L, T2>;
    constexpr int VEC = G::VEC, R1 = G::R1, T1 = G::T1, HB = G::HB, HBAL = G::HBAL, SHIFT = G::SHIFT;
    constexpr int W1 = G::W1, W2 = G::W2, W1S = G::W1S, PU = G::PU, PV = G::PV, NCH = G::NCH;
    constexpr int NRUN = T2 / R2;
    constexpr int NA_ITEMS = 4 * NRUN * W1, KA = (NA_ITEMS + NT - 1) / NT;
    constexpr int NB_ITEMS = 2 * 8 * T2, KB = (NB_ITEMS + NT - 1) / NT;
    constexpr int NC_ITEMS = T2 * 16, KC = (NC_ITEMS + NT - 1) / NT;
    constexpr int NROWS = 8 * W2, KR = (NROWS + NT - 1) / NT;       // staged rows per plane
    constexpr uint32_t PLANE_BYTES = G::PLANE_BYTES;
    constexpr int BP = G::BP;
    static_assert(T2 % R2 == 0 && T2 % 8 == 0, "tile rows");

    extern __shared__ __align__(128) unsigned char smem_raw[];
    T *RAW = reinterpret_cast<T *>(smem_raw);          // [8][W2][W1S]  haloed subband tiles of one plane
    T *SU = RAW + G::RAW_ELEMS;                         // [4][T2][PU]
    T *SV = SU + 4 * T2 * PU;                           // [2][T2][PV]
    uint64_t *bar = reinterpret_cast<uint64_t *>(SV + 2 * T2 * PV);

    const int tid = threadIdx.x;
    // CTAs of one cluster (cl1 x cl2 neighbouring tiles) are consecutive block ids; they march in
    // lockstep (split-phase cluster barrier per plane) so that halo rows/columns shared between
    // neighbouring tiles are fetched from DRAM once and hit L2 for the neighbours
    int bid = blockIdx.x;
    const int csz = p.cl1 * p.cl2;
    const int crank = bid % csz;
    bid /= csz;
    const int ct1 = p.tiles1 / p.cl1;
    const int t1 = (bid % ct1) * p.cl1 + crank % p.cl1;
    bid /= ct1;
    const int ct2 = p.tiles2 / p.cl2;
    const int t2 = (bid % ct2) * p.cl2 + crank / p.cl1;
    bid /= ct2;
    const int chunk = bid % p.nchunks;
    const int batch = bid / p.nchunks;
    const int a1 = t1 * T1, a2 = t2 * T2;
    const int z0 = p.zbase + chunk * p.zc;
    const int z1 = min(z0 + p.zc, p.zbase + p.zcount);
    const int n1 = p.n1, n2 = p.n2, n3 = p.n3;
    const int64_t s3 = p.s3;
    const int bsel = batch / p.nhyp, bhyp = batch - bsel * p.nhyp;
    const int64_t boff = (int64_t)bhyp * p.s4;

    // interior tiles (no periodic wrap inside the haloed box) are staged with ONE tensor TMA per
    // subband, issued by a single thread; edge tiles fall back to per-row bulk copies
    const uint64_t pol = p.hint ? l2_policy_evict_last() : 0;
    const bool use_tma = p.prefetch == 3 && (a1 - HBAL >= 0) && (a1 - HBAL + W1S <= n1) && (a2 - HB >= 0) &&
                         (a2 - HB + W2 <= n2);
    // ---- hoisted per-thread constants ----
    // row copies: (band b, haloed row r) -> up to two contiguous segments (periodic wrap along dim 1)
    const T *r_src[KR];
    int r_dst[KR], r_len0[KR];          // first segment length (elements); the second one is W1S - len0
    int r_col1[KR];                     // global column where the second segment starts
#pragma unroll
    for (int k = 0; k < KR; ++k) {
        // copy index spread round-robin over the warps: the per-lane UBLKCP issue is serialised
        // inside a warp, so every warp should carry as few copies as possible
        constexpr int NW = NT / 32;
        const int slot = tid + k * NT;
        const int it = (slot % 32) * NW + (slot / 32) % NW + (slot / NT) * NT;
        const int r = it % W2, b = (it < NROWS) ? it / W2 : 0;
        const int grow = wrapi(a2 - HB + r, n2);
        const int gc0 = wrapi(a1 - HBAL, n1);
        r_src[k] = p.in[8 * bsel + b] + boff + (int64_t)grow * n1;
        r_dst[k] = (it < NROWS) ? b * BP + r * W1S : -1;
        r_len0[k] = min(W1S, n1 - gc0);
        r_col1[k] = gc0;                // segment 0 starts at gc0, segment 1 (if any) at column 0
    }
    // stage RA: item = (q = b1 + 2 b3, run, haloed column c)
    int a_src[KA], a_dst[KA];
    int a_mask = 0;
#pragma unroll
    for (int k = 0; k < KA; ++k) {
        const int it = tid + k * NT;
        // warp-aligned mapping: the first 32 columns of every (q, run) group sit on whole warps
        // (conflict-free shared-memory rows), the W1-32 tail columns of all groups are packed after them
        int c, g;
        constexpr int NG = 4 * NRUN, TAIL = W1 - 32;
        if (W1 > 32 && W1 <= 64) {
            if (it < NG * 32) { g = it >> 5; c = it & 31; }
            else { const int tt = it - NG * 32; g = tt / TAIL; c = 32 + tt % TAIL; }
        } else { c = it % W1; g = it / W1; }
        const int run = g % NRUN, q = (it < NA_ITEMS) ? g / NRUN : 0;
        if (it < NA_ITEMS) a_mask |= 1 << k;
        const int blo = (q & 1) + 4 * (q >> 1);
        a_src[k] = blo * BP + run * R2 * W1S + SHIFT + c;
        a_dst[k] = (q * T2 + run * R2) * PU + c;
    }
    int b_src[KB], b_dst[KB];
    int b_mask = 0;
#pragma unroll
    for (int k = 0; k < KB; ++k) {
        const int it = tid + k * NT;
        const int j = it % T2, g = it / T2;
        const int cb = g & 7, b3 = (g >> 3) & 1;
        if (it < NB_ITEMS) b_mask |= 1 << k;
        b_src[k] = (2 * b3 * T2 + j) * PU + cb * R1;
        b_dst[k] = (b3 * T2 + j) * PV + cb * R1;
    }
    int c_src[KC];
    T *c_out[KC];
    bool c_ok[KC];
    T acc[KC][L][VEC];
#pragma unroll
    for (int k = 0; k < KC; ++k) {
        const int it = tid + k * NT;
        const int cp = it & 15, j = it >> 4;
        c_src[k] = j * PV + cp * VEC;
        const int g1 = a1 + cp * VEC, g2 = a2 + j;
        c_ok[k] = (it < NC_ITEMS) && g1 < n1 && g2 < n2;
        c_out[k] = p.out[bsel] + boff + ((int64_t)z0 - (L - 1)) * s3 + (int64_t)g2 * n1 + g1;
#pragma unroll
        for (int s = 0; s < L; ++s)
#pragma unroll
            for (int e = 0; e < VEC; ++e) acc[k][s][e] = zero_of(T());
    }

    if (tid == 0) {
        mbar_init(bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    // the expect_tx of thread 0 must precede every complete_tx of the same phase: warp 0 issues it
    // first inside issue_plane; the other warps' copies may only start after it -> barrier
    // single chunk = the whole periodic dimension: every coefficient plane is staged exactly once; the
    // L-1 output planes that wrap around get the partial sum of the first steps stored early and the
    // rest added by L-1 flush steps at the end (periodic closure) instead of re-reading L-1 planes
    const bool closed = (p.nchunks == 1 && p.zcount == p.n3);
    const int nsteps = closed ? (z1 - z0) : (z1 - z0) + L - 1;
    if (tid < 32) { if (tid == 0) mbar_expect_tx(bar, PLANE_BYTES); }
    __syncthreads();
    if (use_tma) {
        if (tid == 0) {
            const int zp = bhyp * n3 + wrapi(z0 - HB, n3);
#pragma unroll
            for (int b = 0; b < 8; ++b) { if (p.hint) tma_load_3d_hint(RAW + b * BP, &maps.m[8 * bsel + b], a1 - HBAL, a2 - HB, zp, bar, pol); else tma_load_3d(RAW + b * BP, &maps.m[8 * bsel + b], a1 - HBAL, a2 - HB, zp, bar); }
        }
    } else {
        const int64_t zoff = (int64_t)wrapi(z0 - HB, n3) * s3;
#pragma unroll
        for (int k = 0; k < KR; ++k) {
            if (r_dst[k] >= 0) {
                const T *src = r_src[k] + zoff;
                T *dst = RAW + r_dst[k];
                const int len0 = r_len0[k];
                if (p.hint) {
                    bulk_g2s_hint(dst, src + r_col1[k], (uint32_t)(len0 * sizeof(T)), bar, pol);
                    if (len0 < W1S) bulk_g2s_hint(dst + len0, src, (uint32_t)((W1S - len0) * sizeof(T)), bar, pol);
                } else {
                    bulk_g2s(dst, src + r_col1[k], (uint32_t)(len0 * sizeof(T)), bar);
                    if (len0 < W1S) bulk_g2s(dst + len0, src, (uint32_t)((W1S - len0) * sizeof(T)), bar);
                }
            }
        }
    }

    int u = 0;
    uint32_t parity = 0;
    for (int t = 0; t < nsteps; ++t) {
        mbar_wait(bar, parity);
        parity ^= 1;
        if (csz > 1) asm volatile("barrier.cluster.arrive.relaxed.aligned;" ::: "memory");

        // ---- stage RA: dim 2 out of the staged tiles
#pragma unroll
        for (int k = 0; k < KA; ++k) {
            if (a_mask & (1 << k)) {
                T o[R2];
#pragma unroll
                for (int i = 0; i < R2; ++i) o[i] = zero_of(T());
#pragma unroll
                for (int hb = 0; hb < 2; ++hb) {
                    const T *src = RAW + a_src[k] + hb * 2 * BP;
                    const typename TapOf<T>::type *g = hb ? tp.hi[1] : tp.lo[1];
                    T w[R2 + L - 1];
#pragma unroll
                    for (int i = 0; i < R2 + L - 1; ++i) w[i] = src[i * W1S];
#pragma unroll
                    for (int i = 0; i < R2; ++i)
#pragma unroll
                        for (int kk = 0; kk < L; ++kk) macp(o[i], g[kk], w[i + kk]);
                }
                T *dst = SU + a_dst[k];
#pragma unroll
                for (int i = 0; i < R2; ++i) dst[i * PU] = o[i];
            }
        }
        __syncthreads();   // RAW fully consumed, SU complete

        // ---- stage the next coefficient plane while RB / RC run
        if (t + 1 < nsteps) {
            if (tid == 0) mbar_expect_tx(bar, PLANE_BYTES);
            if (use_tma) {
                if (tid == 0) {
                    const int zp = bhyp * n3 + wrapi(z0 - HB + t + 1, n3);
#pragma unroll
                    for (int b = 0; b < 8; ++b)
                        { if (p.hint) tma_load_3d_hint(RAW + b * BP, &maps.m[8 * bsel + b], a1 - HBAL, a2 - HB, zp, bar, pol); else tma_load_3d(RAW + b * BP, &maps.m[8 * bsel + b], a1 - HBAL, a2 - HB, zp, bar); }
                }
            } else {
                const int64_t zoff = (int64_t)wrapi(z0 - HB + t + 1, n3) * s3;
#pragma unroll
                for (int k = 0; k < KR; ++k) {
                    if (r_dst[k] >= 0) {
                        const T *src = r_src[k] + zoff;
                        T *dst = RAW + r_dst[k];
                        const int len0 = r_len0[k];
                        if (p.hint) {
                            bulk_g2s_hint(dst, src + r_col1[k], (uint32_t)(len0 * sizeof(T)), bar, pol);
                            if (len0 < W1S) bulk_g2s_hint(dst + len0, src, (uint32_t)((W1S - len0) * sizeof(T)), bar, pol);
                        } else {
                            bulk_g2s(dst, src + r_col1[k], (uint32_t)(len0 * sizeof(T)), bar);
                            if (len0 < W1S) bulk_g2s(dst + len0, src, (uint32_t)((W1S - len0) * sizeof(T)), bar);
                        }
                    }
                }
            }
        }

        // ---- stage RB: dim 1
#pragma unroll
        for (int k = 0; k < KB; ++k) {
            if (b_mask & (1 << k)) {
                T o[R1];
#pragma unroll
                for (int i = 0; i < R1; ++i) o[i] = zero_of(T());
#pragma unroll
                for (int hb = 0; hb < 2; ++hb) {
                    const T *row = SU + b_src[k] + hb * T2 * PU;
                    const typename TapOf<T>::type *g = hb ? tp.hi[0] : tp.lo[0];
                    T v[NCH * VEC];
#pragma unroll
                    for (int j = 0; j < NCH; ++j) ld_chunk<T, VEC>(row + j * VEC, v + j * VEC);
#pragma unroll
                    for (int i = 0; i < R1; ++i)
#pragma unroll
                        for (int kk = 0; kk < L; ++kk) macp(o[i], g[kk], v[i + kk]);
                }
                st_chunk<T, VEC>(SV + b_dst[k], o);
                st_chunk<T, VEC>(SV + b_dst[k] + VEC, o + VEC);
            }
        }
        __syncthreads();

        // ---- stage RC: dim 3 scatter ring
        const bool store = closed || (t >= L - 1);
        const int64_t wrap_off = (closed && t < L - 1) ? (int64_t)n3 * s3 : 0;   // early partials of the wrapped planes
#pragma unroll
        for (int k = 0; k < KC; ++k) {
            if (tid + k * NT < NC_ITEMS) {
                T v0[VEC], v1[VEC];
                ld_chunk<T, VEC>(SV + c_src[k], v0);
                ld_chunk<T, VEC>(SV + c_src[k] + T2 * PV, v1);
                DispatchC<T, L, VEC, 0>::run(u, acc[k], v0, v1, tp.lo[2], tp.hi[2], c_out[k] + wrap_off,
                                             store && c_ok[k]);
                c_out[k] += s3;
            }
        }
        u = (u + 1 == L) ? 0 : u + 1;
        if (csz > 1) asm volatile("barrier.cluster.wait.aligned;" ::: "memory");
    }
    if (closed) {   // flush: planes n3-L+1 .. n3-1 = early partial (already in memory) + what is left in the ring
        for (int f = 0; f < L - 1; ++f) {
#pragma unroll
            for (int k = 0; k < KC; ++k) {
                if (tid + k * NT < NC_ITEMS) {
                    T v0[VEC], v1[VEC];
#pragma unroll
                    for (int e = 0; e < VEC; ++e) { v0[e] = zero_of(T()); v1[e] = zero_of(T()); }
                    DispatchC<T, L, VEC, 0>::run(u, acc[k], v0, v1, tp.lo[2], tp.hi[2], c_out[k], c_ok[k], true);
                    c_out[k] += s3;
                }
            }
            u = (u + 1 == L) ? 0 : u + 1;
        }
    }
}


// ---------------------------------------------------------------------------------------------
// Synthesis tile kernel, FULL-ROW variant (8-byte elements, rows of at most KC*NT*VEC/T2 elements: 192 with 384
// threads, 256 with 512).
// tools/tile_probe.cu (profiles/r01_tile_probe.md) showed that the memory access pattern of the
// 32-column tiles is itself the limit of k_rec3_bulk: a copy-only kernel with that geometry reaches
// 4.4 TB/s, the same copy with full contiguous rows 6.3 TB/s.  Here a CTA owns ALL n1 columns of
// T2 rows and marches along dim 3:
//   * staging unit = one (b2 = 0, 1) band pair of one (b1, b3) group: 2 x (T2+L-1) whole rows, which
//     are CONTIGUOUS in memory -> one cp.async.bulk per band (two when the rows wrap around dim 2),
//     issued by one thread into an NSTG-deep ring of stages with one mbarrier each; the ring runs
//     ahead of the arithmetic across group and plane boundaries;
//   * stage RA (dim 2): item = (column, half of the tile rows), so that all warps work; the RH outputs of
//     both bands are 2*RH independent FFMA2 chains (tap loop outermost) over one window of RH+L-1 staged
//     rows per band; the periodic wrap of dim 1 is materialised as HB + HA pad columns of SU;
//   * stage RB (dim 1): 2*VEC outputs per item, lo/hi sums as separate chains; stage RC: the dim-3
//     scatter ring of k_rec3_bulk, two 16-byte chunks per thread.
struct RowsGeo {
    int pu, pv;            // SU / SV row pitch (elements), odd chunk counts
    int nstg;              // stages in the ring
    int stage_elems;       // 2 * (T2+L-1) * n1
};

template <typename T, int L>
__host__ __device__ constexpr int rows_pu(int n1) { return (((n1 + L - 1 + 16 / (int)sizeof(T) - 1) / (16 / (int)sizeof(T))) | 1) * (16 / (int)sizeof(T)); }
template <typename T>
__host__ __device__ constexpr int rows_pv(int n1) { return ((n1 / (16 / (int)sizeof(T))) | 1) * (16 / (int)sizeof(T)); }

// N1 > 0: row length known at compile time (every shared-memory offset becomes an immediate);
// N1 == 0: taken from the parameters.
template <typename T, int L, int T2, int NT, int KC, int N1, int RASPLIT = 2, int RBM = 1>
__global__ void __launch_bounds__(NT, 1)
k_rec3_rows(const Rec3Params<T> p, const FusedTaps<T, L> tp, const RowsGeo g)
{
    constexpr int VEC = 16 / (int)sizeof(T);
    constexpr int HB = L / 2, HA = L / 2 - 1, W2 = T2 + L - 1;
    // stage-RA rows per item: RASPLIT items per column.  tools/tile_model.py: with TMA writes counted, the kernel
    // runs at ~85 % of the shared-memory pipe; RASPLIT = 1 and RBM = 2 move 19 % fewer wavefronts but keep only
    // half of the warps busy in those stages (variants 7xx / 8xx, not yet timed)
    constexpr int RH = T2 / RASPLIT;
    constexpr int R1B = 2 * RBM * VEC, NCHB = (R1B + L - 1 + VEC - 1) / VEC;
    static_assert(sizeof(T) == 8 && T2 % 2 == 0, "8-byte elements, even tile height");

    const int n1 = N1 ? N1 : p.n1;
    const int pu = N1 ? rows_pu<T, L>(N1) : g.pu;
    const int pv = N1 ? rows_pv<T>(N1) : g.pv;
    const int stage_elems = 2 * W2 * n1;
    const int nstg = g.nstg;

    extern __shared__ __align__(128) unsigned char smem_raw[];
    T *STG = reinterpret_cast<T *>(smem_raw);                    // [nstg][2][W2][n1]
    T *SU = STG + (size_t)nstg * stage_elems;                     // [4][T2][pu]   (q = b1 + 2 b3), dim 2 synthesised, wrap pads
    T *SV = SU + 4 * T2 * pu;                                     // [2][T2][pv]   (b3), dims 1,2 synthesised
    uint64_t *bar = reinterpret_cast<uint64_t *>(SV + 2 * T2 * pv);

    const int tid = threadIdx.x;
    int bid = blockIdx.x;
    const int t2 = bid % p.tiles2;
    bid /= p.tiles2;
    const int chunk = bid % p.nchunks;
    const int batch = bid / p.nchunks;
    const int a2 = t2 * T2;
    const int z0 = p.zbase + chunk * p.zc;
    const int z1 = min(z0 + p.zc, p.zbase + p.zcount);
    const int n2 = p.n2, n3 = p.n3;
    const int64_t s3 = p.s3;
    const int bsel = batch / p.nhyp, bhyp = batch - bsel * p.nhyp;
    const int64_t boff = (int64_t)bhyp * p.s4;

    const bool closed = (p.nchunks == 1 && p.zcount == p.n3);
    const int nsteps = closed ? (z1 - z0) : (z1 - z0) + L - 1;
    const int total = 4 * nsteps;                                 // pair steps
    const uint32_t stage_bytes = (uint32_t)(stage_elems * sizeof(T));

    // staging: rows a2-HB .. a2-HB+W2-1 (mod n2) of a band plane are one contiguous run (two at the dim-2
    // wrap): thread 0 issues one or two cp.async.bulk per band.  (One copy per row from many lanes costs
    // 25 % more instructions over the whole kernel and is slower: profiles/r01_rows_kernel.md.)
    const int r0 = wrapi(a2 - HB, n2);
    const int rows0 = min(W2, n2 - r0);
    int iz = wrapi(z0 - HB, n3);                                  // plane of the next stage to issue
    int iq = 0;                                                   // group of the next stage to issue
    auto issue = [&](int stg) {   // stages are issued in order
        if (tid == 0) {
            mbar_expect_tx(bar + stg, stage_bytes);
            const int blo = 8 * bsel + (iq & 1) + 4 * (iq >> 1);
            T *dst = STG + (size_t)stg * stage_elems;
#pragma unroll
            for (int hb = 0; hb < 2; ++hb) {
                const T *src = p.in[blo + 2 * hb] + boff + (int64_t)iz * s3;
                T *d = dst + hb * W2 * n1;
                bulk_g2s(d, src + (int64_t)r0 * n1, (uint32_t)(rows0 * n1 * sizeof(T)), bar + stg);
                if (rows0 < W2) bulk_g2s(d + rows0 * n1, src, (uint32_t)((W2 - rows0) * n1 * sizeof(T)), bar + stg);
            }
        }
        if (++iq == 4) { iq = 0; if (++iz == n3) iz = 0; }
    };

    // stage RA: item = (column c, half h of the tile rows)
    // stage RC: thread owns KC 16-byte output chunks and their L-deep rings of partial sums
    const int CPR = n1 / VEC, NC_ITEMS = T2 * CPR;
    int c_src[KC];
    T *c_out[KC];
    bool c_ok[KC];
    T acc[KC][L][VEC];
#pragma unroll
    for (int k = 0; k < KC; ++k) {
        const int it = tid + k * NT;
        const int j = (it / CPR) % T2, cp = it - (it / CPR) * CPR;
        c_src[k] = j * pv + cp * VEC;
        c_ok[k] = (it < NC_ITEMS) && (a2 + j < n2);
        c_out[k] = p.out[bsel] + boff + ((int64_t)z0 - (L - 1)) * s3 + (int64_t)(a2 + j) * n1 + cp * VEC;
#pragma unroll
        for (int s = 0; s < L; ++s)
#pragma unroll
            for (int e = 0; e < VEC; ++e) acc[k][s][e] = zero_of(T());
    }

    if (tid == 0) {
        for (int i = 0; i < nstg; ++i) mbar_init(bar + i, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    int issued = 0;
    for (; issued < nstg && issued < total; ++issued) issue(issued);

    const int NA_ITEMS = RASPLIT * n1;
    const int NB_ITEMS = T2 * (n1 / R1B);
    int u = 0, stg = 0;
    uint32_t parity = 0;
    for (int t = 0; t < nsteps; ++t) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            mbar_wait(bar + stg, parity);
            // ---- stage RA: dim 2 for group q, sliding windows over both b2 bands
            {
                const T *S = STG + (size_t)stg * stage_elems;
                for (int it = tid; it < NA_ITEMS; it += NT) {
                    const int h = it / n1, c = it - h * n1;
                    const T *s0 = S + h * RH * n1 + c;
                    const T *s1 = s0 + W2 * n1;
                    T *dst = SU + (q * T2 + h * RH) * pu + HB + c;
                    const int wl = (c >= n1 - HB) ? -n1 : 0;      // left-pad copy of columns n1-HB .. n1-1
                    const int wr = (c < HA) ? n1 : 0;             // right-pad copy of columns 0 .. HA-1
                    // all RH outputs in flight at once: 2 * RH independent FFMA2 chains (two chains stall on the
                    // FFMA2 latency: "wait" was the top stall reason)
                    T w0[RH + L - 1], w1[RH + L - 1];
#pragma unroll
                    for (int i = 0; i < RH + L - 1; ++i) {
                        w0[i] = s0[i * n1];
                        w1[i] = s1[i * n1];
                    }
                    T o0[RH], o1[RH];
#pragma unroll
                    for (int i = 0; i < RH; ++i) { o0[i] = zero_of(T()); o1[i] = zero_of(T()); }
#pragma unroll
                    for (int kk = 0; kk < L; ++kk)
#pragma unroll
                        for (int i = 0; i < RH; ++i) {
                            macp(o0[i], tp.lo[1][kk], w0[i + kk]);
                            macp(o1[i], tp.hi[1][kk], w1[i + kk]);
                        }
#pragma unroll
                    for (int i = 0; i < RH; ++i) {
                        o0[i] = add(o0[i], o1[i]);
                        dst[i * pu] = o0[i];
                    }
                    if (wl | wr) {
                        const int wo = wl ? wl : wr;
#pragma unroll
                        for (int i = 0; i < RH; ++i) dst[i * pu + wo] = o0[i];
                    }
                }
            }
            __syncthreads();   // stage consumed, SU[q] complete
            if (issued < total) { issue(stg); ++issued; }
            if (++stg == nstg) { stg = 0; parity ^= 1; }

            // ---- stage RB: dim 1 for b3 = q >> 1 once both of its groups are in SU
            if (q & 1) {
                const int b3 = q >> 1;
                for (int it = tid; it < NB_ITEMS; it += NT) {
                    const int j = it % T2, cb = it / T2;
                    const T *ra = SU + (2 * b3 * T2 + j) * pu + cb * R1B;
                    const T *rb = ra + T2 * pu;
                    T va[NCHB * VEC], vb[NCHB * VEC];
#pragma unroll
                    for (int c = 0; c < NCHB; ++c) {
                        ld_chunk<T, VEC>(ra + c * VEC, va + c * VEC);
                        ld_chunk<T, VEC>(rb + c * VEC, vb + c * VEC);
                    }
                    T o[R1B], ob[R1B];
#pragma unroll
                    for (int i = 0; i < R1B; ++i) { o[i] = zero_of(T()); ob[i] = zero_of(T()); }
#pragma unroll
                    for (int kk = 0; kk < L; ++kk)
#pragma unroll
                        for (int i = 0; i < R1B; ++i) {
                            macp(o[i], tp.lo[0][kk], va[i + kk]);
                            macp(ob[i], tp.hi[0][kk], vb[i + kk]);
                        }
#pragma unroll
                    for (int i = 0; i < R1B; ++i) o[i] = add(o[i], ob[i]);
                    T *d = SV + (b3 * T2 + j) * pv + cb * R1B;
#pragma unroll
                    for (int c = 0; c < R1B / VEC; ++c) st_chunk<T, VEC>(d + c * VEC, o + c * VEC);
                }
            }
        }
        __syncthreads();   // SV complete

        // ---- stage RC: dim 3 scatter ring
        const bool store = closed || (t >= L - 1);
        const int64_t wrap_off = (closed && t < L - 1) ? (int64_t)n3 * s3 : 0;   // early partials of the wrapped planes
#pragma unroll
        for (int k = 0; k < KC; ++k) {
            if (tid + k * NT < NC_ITEMS) {
                T v0[VEC], v1[VEC];
                ld_chunk<T, VEC>(SV + c_src[k], v0);
                ld_chunk<T, VEC>(SV + c_src[k] + T2 * pv, v1);
                DispatchC<T, L, VEC, 0>::run(u, acc[k], v0, v1, tp.lo[2], tp.hi[2], c_out[k] + wrap_off,
                                             store && c_ok[k]);
                c_out[k] += s3;
            }
        }
        u = (u + 1 == L) ? 0 : u + 1;
    }
    if (closed) {   // flush: planes n3-L+1 .. n3-1 = early partial (already in memory) + what is left in the ring
        for (int f = 0; f < L - 1; ++f) {
#pragma unroll
            for (int k = 0; k < KC; ++k) {
                if (tid + k * NT < NC_ITEMS) {
                    T v0[VEC], v1[VEC];
#pragma unroll
                    for (int e = 0; e < VEC; ++e) { v0[e] = zero_of(T()); v1[e] = zero_of(T()); }
                    DispatchC<T, L, VEC, 0>::run(u, acc[k], v0, v1, tp.lo[2], tp.hi[2], c_out[k], c_ok[k], true);
                    c_out[k] += s3;
                }
            }
            u = (u + 1 == L) ? 0 : u + 1;
        }
    }
}


// ---------------------------------------------------------------------------------------------
// Last-dimension passes of the 4-D path (and the slab-exchange points of the multi-GPU path):
// one thread owns one 16-byte chunk of the (dim 1..3) hyperplane and marches along dim 4 with an
// L-deep register ring, so every input hyperplane is read exactly once.
template <typename T, int L>
struct LastTaps {
    typename TapOf<T>::type lo[L];
    typename TapOf<T>::type hi[L];
};

template <typename T, int L>
__global__ void __launch_bounds__(256)
k_dec_last(const T *__restrict__ in, const T *__restrict__ halo_lo, const T *__restrict__ halo_hi,
           T *__restrict__ out_lo, T *__restrict__ out_hi, int64_t nchunks, int n4, int64_t s4, int below,
           const LastTaps<T, L> tp)
{
    constexpr int VEC = 16 / (int)sizeof(T), HB = L / 2 - 1, HA = L / 2;
    const int64_t ci = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (ci >= nchunks) return;
    const bool slab = (halo_lo != nullptr) || (halo_hi != nullptr);
    auto plane_ptr = [&](int zi) -> const T * {
        if (!slab) return in + (int64_t)wrapi(zi, n4) * s4;
        if (zi < 0) return halo_lo + (int64_t)(zi + below) * s4;
        if (zi >= n4) return halo_hi + (int64_t)(zi - n4) * s4;
        return in + (int64_t)zi * s4;
    };
    const int64_t off = ci * VEC;
    T ring[L][VEC];
#pragma unroll
    for (int j = 0; j < L; ++j) ld_chunk<T, VEC>(plane_ptr(j - HB) + off, ring[j]);
    for (int zb = 0; zb < n4; zb += L) {
#pragma unroll
        for (int u = 0; u < L; ++u) {
            const int z = zb + u;
            if (z < n4) {
                T lo[VEC], hi[VEC];
#pragma unroll
                for (int e = 0; e < VEC; ++e) {
                    lo[e] = zero_of(T());
                    hi[e] = zero_of(T());
#pragma unroll
                    for (int j = 0; j < L; ++j) {
                        macp(lo[e], tp.lo[L - 1 - j], ring[(u + j) % L][e]);
                        macp(hi[e], tp.hi[L - 1 - j], ring[(u + j) % L][e]);
                    }
                }
                if (z + 1 < n4) ld_chunk<T, VEC>(plane_ptr(z + 1 + HA) + off, ring[u]);
                st_chunk<T, VEC>(out_lo + (int64_t)z * s4 + off, lo);
                st_chunk<T, VEC>(out_hi + (int64_t)z * s4 + off, hi);
            }
        }
    }
}

template <typename T, int L>
__global__ void __launch_bounds__(256)
k_rec_last(const T *__restrict__ u_lo, const T *__restrict__ u_hi, const T *__restrict__ halo_lo,
           const T *__restrict__ halo_hi, T *__restrict__ out, int64_t nchunks, int n4, int64_t s4, int below,
           int above, const LastTaps<T, L> tp)
{
    constexpr int VEC = 16 / (int)sizeof(T), HB = L / 2;
    const int64_t ci = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (ci >= nchunks) return;
    const bool slab = (halo_lo != nullptr) || (halo_hi != nullptr);
    const int64_t off = ci * VEC;
    T acc[L][VEC];
#pragma unroll
    for (int j = 0; j < L; ++j)
#pragma unroll
        for (int e = 0; e < VEC; ++e) acc[j][e] = zero_of(T());
    const int nsteps = n4 + L - 1;
    for (int tb = 0; tb < nsteps; tb += L) {
#pragma unroll
        for (int u = 0; u < L; ++u) {
            const int t = tb + u;
            if (t < nsteps) {
                const int zi = t - HB;          // coefficient hyperplane (may lie in the halo)
                const T *pl, *ph;
                if (!slab) {
                    const int64_t o = (int64_t)wrapi(zi, n4) * s4;
                    pl = u_lo + o;
                    ph = u_hi + o;
                } else if (zi < 0) {
                    pl = halo_lo + (int64_t)(zi + below) * s4;
                    ph = halo_lo + (int64_t)(zi + 2 * below) * s4;
                } else if (zi >= n4) {
                    pl = halo_hi + (int64_t)(zi - n4) * s4;
                    ph = halo_hi + (int64_t)(zi - n4 + above) * s4;
                } else {
                    pl = u_lo + (int64_t)zi * s4;
                    ph = u_hi + (int64_t)zi * s4;
                }
                T v0[VEC], v1[VEC];
                ld_chunk<T, VEC>(pl + off, v0);
                ld_chunk<T, VEC>(ph + off, v1);
                // output hyperplane n = t - (L-1) completes at this step
                DispatchC<T, L, VEC, 0>::run(u, acc, v0, v1, tp.lo, tp.hi, out + (int64_t)(t - (L - 1)) * s4 + off,
                                             t >= L - 1);
            }
        }
    }
}

// Scatter form of the last-dim synthesis for slabs: only the LOCAL u planes are read; the partial
// sums that fall outside the slab (L/2-1 planes below, L/2 above) go to two overhang buffers which
// the caller sends to the owning ranks (one array instead of the two u arrays of the gather form).
template <typename T, int L>
__global__ void __launch_bounds__(256)
k_rec_last_scatter(const T *__restrict__ u_lo, const T *__restrict__ u_hi, T *__restrict__ out,
                   T *__restrict__ over_lo, T *__restrict__ over_hi, int64_t nchunks, int n4, int64_t s4,
                   const LastTaps<T, L> tp)
{
    constexpr int VEC = 16 / (int)sizeof(T), HB = L / 2, BELOW = L / 2 - 1;
    const int64_t ci = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (ci >= nchunks) return;
    const int64_t off = ci * VEC;
    T acc[L][VEC];
#pragma unroll
    for (int j = 0; j < L; ++j)
#pragma unroll
        for (int e = 0; e < VEC; ++e) acc[j][e] = zero_of(T());
    // local coefficient plane zi = t - HB, t = HB .. HB + n4 - 1; then L-1 flush steps
    const int nsteps = n4 + L - 1;
    for (int tb = 0; tb < nsteps; tb += L) {
#pragma unroll
        for (int u = 0; u < L; ++u) {
            const int tt = tb + u;
            if (tt < nsteps) {
                const int zi = tt;                       // 0 .. n4 + L - 2 (>= n4: flush with zeros)
                T v0[VEC], v1[VEC];
                if (zi < n4) {
                    ld_chunk<T, VEC>(u_lo + (int64_t)zi * s4 + off, v0);
                    ld_chunk<T, VEC>(u_hi + (int64_t)zi * s4 + off, v1);
                } else {
#pragma unroll
                    for (int e = 0; e < VEC; ++e) { v0[e] = zero_of(T()); v1[e] = zero_of(T()); }
                }
                // this step completes output plane m = zi - L/2 + 1
                const int m = zi - HB + 1;
                T *dst = (m < 0) ? over_lo + (int64_t)(m + BELOW) * s4
                                 : (m >= n4 ? over_hi + (int64_t)(m - n4) * s4 : out + (int64_t)m * s4);
                DispatchC<T, L, VEC, 0>::run(u, acc, v0, v1, tp.lo, tp.hi, dst + off, true);
            }
        }
    }
}

template <typename R>
__global__ void __launch_bounds__(256) k_accumulate(R *__restrict__ dst, const R *__restrict__ src, int64_t n)
{
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) dst[i] += src[i];
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
template <typename T, int L>
static FusedTaps<T, L> make_taps(const nddwt_plan *p, bool rec)
{
    FusedTaps<T, L> t;
    const AllTaps<double> &src = rec ? p->rec_d : p->dec_d;
    for (int d = 0; d < 3; ++d)
        for (int k = 0; k < L; ++k) {
            const int dd = d < p->ndims ? d : 0;
            t.lo[d][k] = mk_tap(typename TapOf<T>::type(), padded_tap(src.d[dd].lo, p->L[dd], L, k, p->cur_dil));
            t.hi[d][k] = mk_tap(typename TapOf<T>::type(), padded_tap(src.d[dd].hi, p->L[dd], L, k, p->cur_dil));
        }
    return t;
}

static int pick_zc(int n3, int tiles, int H, int ctas_per_wave)
{
    // balanced chunks of the marching dimension: enough CTAs to fill the machine in whole waves,
    // while keeping the ring warm-up (L planes re-read per chunk, loads only) a small fraction
    int best = n3;
    double best_cost = 1e30;
    for (int nch = 1; nch <= n3; ++nch) {
        const int zc = (n3 + nch - 1) / nch;
        if (zc < 4 && nch > 1) break;
        const int real_nch = (n3 + zc - 1) / zc;
        const double ctas = (double)tiles * real_nch;
        const double waves = ctas / ctas_per_wave;
        const double eff = waves / (double)((int64_t)(waves + 0.999999));   // tail efficiency
        const double work = (double)(zc + 0.35 * (H + 1)) / zc;              // warm-up overhead (loads only)
        const double cost = work / eff;
        if (cost < best_cost - 1e-9) { best_cost = cost; best = zc; }
    }
    return best;
}

template <typename T, int L, int T2, int NT, int R2, int MINB, int CWSEL = 0, int RBM = 1, int ZINC = 0, int SHR = 0>
static int launch_dec3_v(nddwt_plan *p, const Dec3Params<T> &base, cudaStream_t s)
{
    using G = Geo<T, L, T2>;
    Dec3Params<T> prm = base;
    {   // thresholds of this level's bands; prm.out[b] == nullptr marks bands this launch does not produce
        const int j = p->cur_level >= 1 && p->cur_level <= NDDWT_MAX_LEVELS ? p->cur_level : 1;
        const int half = prm.band0;
        for (int b = 0; b < 16; ++b)
            prm.thr[b] = (typename Elem<T>::R)((SHR && b + half < (1 << p->ndims)) ? p->shrink_thr[j - 1][b + half] : 0.0);
    }
    prm.tiles1 = (prm.n1 + G::T1 - 1) / G::T1;
    prm.tiles2 = (prm.n2 + T2 - 1) / T2;
    const int batches = prm.nhyp * (prm.in[1] ? 2 : 1);
    if (prm.zcount <= 0) { prm.zbase = 0; prm.zcount = prm.n3; }
    prm.zc = pick_zc(prm.zcount, prm.tiles1 * prm.tiles2 * batches, L - 1, 148 * MINB);
    prm.nchunks = (prm.zcount + prm.zc - 1) / prm.zc;
    prm.halo_below = (L / 2 - 1);
    auto kern = k_dec3_fused<T, L, T2, NT, R2, MINB, CWSEL, RBM, ZINC, SHR>;
    NDDWT_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)G::SMEM));   // per-device attribute: set on every launch
    const FusedTaps<T, L> tp = make_taps<T, L>(p, false);
    const int64_t grid = (int64_t)prm.tiles1 * prm.tiles2 * prm.nchunks * batches;
    {
        LaunchTimer lt(p, KIND_DEC3, s);
        kern<<<(unsigned)grid, NT, G::SMEM, s>>>(prm, tp);
    }
    p->launches++;
    NDDWT_CUDA(cudaGetLastError());
    return 0;
}

// Experiment switches exist only in tuning builds (make TUNING=1 -> -DNDDWT_TUNING): the release library
// reads no environment variables.  What each rejected variant measured is in profiles/.
#ifdef NDDWT_TUNING
static int tuning_env(const char *name, int dflt)
{
    const char *e = getenv(name);
    return e ? atoi(e) : dflt;
}
static int tuning_variant()
{
    static int v = -1;
    if (v < 0) v = tuning_env("NDDWT_VARIANT", 0);
    return v;
}
#else
static constexpr int tuning_env(const char *, int dflt) { return dflt; }
static constexpr int tuning_variant() { return 0; }
#endif

// tile-kernel configuration shared by the 3-D path and the 4-D back end
template <typename T, int L>
static int launch_dec3_any(nddwt_plan *p, const Dec3Params<T> &prm, cudaStream_t s)
{
#ifdef NDDWT_TUNING
    if constexpr (L == 8 && sizeof(T) == 8 && Elem<T>::cplx) {
        switch (tuning_variant() % 10) {
            case 1: return launch_dec3_v<T, L, 16, 256, 2, 2>(p, prm, s);   // plane pointer by integer modulo (round-1 default): 3.74 ms vs 3.53 ms (cfg5)
            default: break;
        }
    }
#endif
    // rows that are not 16-byte multiples: element-wise stage-C columns and stores (plans with the fused shrink never get
    // here with such rows: dec_rows_ok)
    if (prm.n1 % (16 / (int)sizeof(T)) != 0) return launch_dec3_v<T, L, 16, 256, 8, 2, 1, 1, 1>(p, prm, s);
    // incremental plane pointer (ZINC): no integer modulo per plane in stage A; SHR: soft threshold fused into the stores
    // shrink epilogue: one-row stage-C runs keep it spill-free (two-row runs: 260 B of spills, 6.8 vs 4.4 ms on cfg5)
    if (p->shrink_mode) return launch_dec3_v<T, L, 16, 256, 1, 2, 0, 1, 1, 1>(p, prm, s);
    // (stage B shared evenly by all warps -- every thread half an item in the last half-iteration -- measured
    // slower: 3.9-4.1 vs 3.65-3.7 ms; 32 x 32 tiles with one 512-thread CTA per SM (ring of 3 positions per thread,
    // -11 % FMAs, -28 % shared-memory wavefronts): 3.95-3.99 ms with 2-row stage-C runs, 5.0-5.2 ms with 4-row runs
    // (spills) vs 3.51-3.53 ms: profiles/r02_variants.md)
    return launch_dec3_v<T, L, 16, 256, 2, 2, 0, 1, 1>(p, prm, s);
}

template <typename T, int L>
static int launch_dec3(nddwt_plan *p, const void *a_in, const LevelIO &io, void *const *out_bands, cudaStream_t s)
{
    Dec3Params<T> prm;
    prm.in[0] = reinterpret_cast<const T *>(a_in);
    prm.in[1] = nullptr;
    prm.s4 = 0;
    prm.nhyp = 1;
    prm.halo_lo = reinterpret_cast<const T *>(io.halo_lo);
    prm.halo_hi = reinterpret_cast<const T *>(io.halo_hi);
    for (int b = 0; b < 8; ++b) prm.out[b] = reinterpret_cast<T *>(out_bands[b]);
    for (int b = 8; b < 16; ++b) prm.out[b] = nullptr;
    prm.n1 = (int)p->dims[0];
    prm.n2 = (int)p->dims[1];
    prm.n3 = (int)p->dims[2];
    prm.s3 = p->dims[0] * p->dims[1];
    return launch_dec3_any<T, L>(p, prm, s);
}

template <typename T>
static int dispatch_dec3(nddwt_plan *p, int L, const void *a_in, const LevelIO &io, void *const *out_bands,
                         cudaStream_t s)
{
    switch (L) {
        case 2: return launch_dec3<T, 2>(p, a_in, io, out_bands, s);
        case 4: return launch_dec3<T, 4>(p, a_in, io, out_bands, s);
        case 6: return launch_dec3<T, 6>(p, a_in, io, out_bands, s);
        case 8: return launch_dec3<T, 8>(p, a_in, io, out_bands, s);
        default: return 1;
    }
}


static int pick_zc_rec(int n3, int units_per_plane_chunk, int H, int slots, bool closure = false)
{
    // synthesis pays the full pipeline for the L-1 warm-up planes of every chunk: minimise
    // ceil(units / slots) * (zc + H)
    int best = n3;
    double best_cost = 1e30;
    for (int nch = 1; nch <= n3; ++nch) {
        const int zc = (n3 + nch - 1) / nch;
        if (zc < 4 && nch > 1) break;
        const int64_t units = (int64_t)units_per_plane_chunk * ((n3 + zc - 1) / zc);
        const double cost = (double)((units + slots - 1) / slots) * (zc + ((nch == 1 && closure) ? 0 : H));
        if (cost < best_cost - 1e-9) { best_cost = cost; best = zc; }
    }
    return best;
}

template <typename T, int L, int T2, int NT, int R2, int MINB, int EDGE = 0>
static int launch_rec3_v(nddwt_plan *p, const Rec3Params<T> &base, cudaStream_t s)
{
    using G = GeoR<T, L, T2>;
    Rec3Params<T> prm = base;
    prm.tiles1 = (prm.n1 + G::T1 - 1) / G::T1;
    prm.tiles2 = (prm.n2 + T2 - 1) / T2;
    const int batches = prm.nhyp * (prm.out[1] ? 2 : 1);
    if (prm.zcount <= 0) { prm.zbase = 0; prm.zcount = prm.n3; }
    prm.zc = pick_zc_rec(prm.zcount, prm.tiles1 * prm.tiles2 * batches, L - 1, 148 * MINB);
    prm.nchunks = (prm.zcount + prm.zc - 1) / prm.zc;
    prm.prefetch = tuning_env("NDDWT_PREFETCH", 0);   // prefetch.global.L1/L2 of the next plane: no gain (profiles/)
    p->last_rec_kernel = 1;
    auto kern = k_rec3_fused<T, L, T2, NT, R2, MINB, EDGE>;
    NDDWT_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)G::SMEM));   // per-device attribute: set on every launch
    const FusedTaps<T, L> tp = make_taps<T, L>(p, true);
    const int64_t grid = (int64_t)prm.tiles1 * prm.tiles2 * prm.nchunks * batches;
    {
        LaunchTimer lt(p, KIND_REC3, s);
        kern<<<(unsigned)grid, NT, G::SMEM, s>>>(prm, tp);
    }
    p->launches++;
    NDDWT_CUDA(cudaGetLastError());
    return 0;
}

// cuTensorMapEncodeTiled through the runtime's driver entry point (no link-time libcuda dependency)
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_fn()
{
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void *ptr = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(ptr);
    }
    return fn;
}

static CUtensorMapL2promotion l2_promotion()
{
    const int v = tuning_env("NDDWT_L2PROMO", 0);   // 128B / 256B promotion: no effect (profiles/)
    return v == 2 ? CU_TENSOR_MAP_L2_PROMOTION_L2_256B : v == 1 ? CU_TENSOR_MAP_L2_PROMOTION_L2_128B
                                                               : CU_TENSOR_MAP_L2_PROMOTION_NONE;
}

// 3-D map over one subband array of 8-byte elements: (n1, n2, planes), box (bw, bh, 1)
static bool encode_band_map(CUtensorMap *map, const void *base, int n1, int n2, int64_t planes, int bw, int bh)
{
    EncodeTiledFn fn = encode_fn();
    if (!fn || (reinterpret_cast<uintptr_t>(base) & 15) != 0) return false;
    const cuuint64_t gdim[3] = {(cuuint64_t)n1, (cuuint64_t)n2, (cuuint64_t)planes};
    const cuuint64_t gstr[2] = {(cuuint64_t)n1 * 8, (cuuint64_t)n1 * n2 * 8};
    const cuuint32_t box[3] = {(cuuint32_t)bw, (cuuint32_t)bh, 1};
    const cuuint32_t estr[3] = {1, 1, 1};
    return fn(map, CU_TENSOR_MAP_DATA_TYPE_UINT64, 3, const_cast<void *>(base), gdim, gstr, box, estr,
              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, l2_promotion(),
              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

template <typename T, int L, int T2, int NT, int R2, int MINB>
static int launch_rec3_bulk(nddwt_plan *p, const Rec3Params<T> &base, cudaStream_t s)
{
    using G = GeoRB<T, L, T2>;
    Rec3Params<T> prm = base;
    prm.tiles1 = (prm.n1 + G::T1 - 1) / G::T1;
    prm.tiles2 = (prm.n2 + T2 - 1) / T2;
    const int batches = prm.nhyp * (prm.out[1] ? 2 : 1);
    if (prm.zcount <= 0) { prm.zbase = 0; prm.zcount = prm.n3; }
    prm.zc = pick_zc_rec(prm.zcount, prm.tiles1 * prm.tiles2 * batches, L - 1, 148 * MINB, prm.zcount == prm.n3);
    prm.nchunks = (prm.zcount + prm.zc - 1) / prm.zc;
    prm.prefetch = 0;
    prm.cl1 = prm.cl2 = 1;
    prm.hint = tuning_env("NDDWT_L2HINT", 0);         // evict_last hints on the staged tiles: no effect (profiles/)
    TmaMaps maps;
    memset(&maps, 0, sizeof maps);
    {
        const int use = tuning_env("NDDWT_TMA", 1);
        if (use && sizeof(T) == 8 && prm.n1 >= G::W1S && prm.n2 >= G::W2) {
            const int nb = prm.out[1] ? 16 : 8;
            bool ok = true;
            for (int b = 0; b < nb && ok; ++b)
                ok = encode_band_map(&maps.m[b], prm.in[b], prm.n1, prm.n2, (int64_t)prm.n3 * prm.nhyp, G::W1S, G::W2);
            if (ok) prm.prefetch = 3;   // kernel flag: tensor maps valid
        }
    }
    p->last_rec_kernel = 2;
    auto kern = k_rec3_bulk<T, L, T2, NT, R2, MINB>;
    NDDWT_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)G::SMEM));   // per-device attribute: set on every launch
    const FusedTaps<T, L> tp = make_taps<T, L>(p, true);
    const int64_t grid = (int64_t)prm.tiles1 * prm.tiles2 * prm.nchunks * batches;
    {
        const int cl = tuning_env("NDDWT_CLUSTER", 1);   // measured: lockstep clusters do not pay (profiles/)
        prm.cl1 = 1;
        prm.cl2 = 1;
        if (cl == 8 && prm.tiles1 % 2 == 0 && prm.tiles2 % 4 == 0) { prm.cl1 = 2; prm.cl2 = 4; }
        else if (cl >= 4 && prm.tiles2 % 4 == 0) { prm.cl1 = 1; prm.cl2 = 4; }
        else if (cl >= 2 && prm.tiles2 % 2 == 0) { prm.cl1 = 1; prm.cl2 = 2; }
    }
    {
        LaunchTimer lt(p, KIND_REC3, s);
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((unsigned)grid);
        cfg.blockDim = dim3(NT);
        cfg.dynamicSmemBytes = G::SMEM;
        cfg.stream = s;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = (unsigned)(prm.cl1 * prm.cl2);
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        NDDWT_CUDA(cudaLaunchKernelEx(&cfg, kern, prm, tp, maps));
    }
    p->launches++;
    NDDWT_CUDA(cudaGetLastError());
    return 0;
}

// full-row synthesis kernel: returns -1 when the geometry does not fit (caller falls back)
template <typename T, int L, int T2, int NT, int KC, int N1, int RASPLIT = 2, int RBM = 1>
static int launch_rec3_rows_n(nddwt_plan *p, const Rec3Params<T> &base, cudaStream_t s)
{
    constexpr int VEC = 16 / (int)sizeof(T), W2 = T2 + L - 1;
    Rec3Params<T> prm = base;
    const int n1 = prm.n1;
    if (sizeof(T) != 8 || (N1 && n1 != N1) || n1 % (2 * RBM * VEC) != 0 || n1 < 4 * VEC || T2 * (n1 / VEC) > KC * NT ||
        prm.n2 < W2)
        return -1;
    RowsGeo g;
    g.pu = rows_pu<T, L>(n1);
    g.pv = rows_pv<T>(n1);
    g.stage_elems = 2 * W2 * n1;
    const size_t fixed = ((size_t)4 * T2 * g.pu + (size_t)2 * T2 * g.pv) * sizeof(T) + 64;
    const size_t stage_bytes = (size_t)g.stage_elems * sizeof(T);
    const size_t cap = 227 * 1024;
    if (fixed + 2 * stage_bytes > cap) return -1;
    g.nstg = (int)std::min<size_t>(4, (cap - fixed) / stage_bytes);
    const size_t smem = fixed + (size_t)g.nstg * stage_bytes;
    prm.tiles1 = 1;
    prm.tiles2 = (prm.n2 + T2 - 1) / T2;
    const int batches = prm.nhyp * (prm.out[1] ? 2 : 1);
    if (prm.zcount <= 0) { prm.zbase = 0; prm.zcount = prm.n3; }
    prm.zc = pick_zc_rec(prm.zcount, prm.tiles2 * batches, L - 1, 148, prm.zcount == prm.n3);
    prm.nchunks = (prm.zcount + prm.zc - 1) / prm.zc;
    {
        // one CTA per SM: needs enough CTAs (row blocks x z-chunks x batches) to fill most of the machine once
        // (one rank's half-level of cfg4 on 8 GPUs is 128 blocks; a 256^3 volume is 32 row blocks x 4 chunks of
        // 64+7 planes: 1.37 -> 0.87 ms for the three synthesis levels of cfg3).  Tests lower the bound through
        // nddwt_plan_set_param("rows_min_ctas") to reach small shapes.
        const int64_t blocks = (int64_t)prm.tiles2 * batches * prm.nchunks;
        if (blocks < p->rows_min_ctas) return -1;
    }
    prm.prefetch = 0;
    prm.cl1 = prm.cl2 = 1;
    prm.hint = 0;
    p->last_rec_kernel = 4;
    auto kern = k_rec3_rows<T, L, T2, NT, KC, N1, RASPLIT, RBM>;
    NDDWT_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));   // per-device attribute: set on every launch
    const FusedTaps<T, L> tp = make_taps<T, L>(p, true);
    const int64_t grid = (int64_t)prm.tiles2 * prm.nchunks * batches;
    {
        LaunchTimer lt(p, KIND_REC3, s);
        kern<<<(unsigned)grid, NT, smem, s>>>(prm, tp, g);
    }
    p->launches++;
    NDDWT_CUDA(cudaGetLastError());
    return 0;
}

template <typename T, int L>
static int launch_rec3_rows(nddwt_plan *p, const Rec3Params<T> &prm, cudaStream_t s)
{
    // RASPLIT = 1 / RBM = 2 (19 % fewer shared-memory wavefronts, but half of the warps idle in those stages)
    // measured slower in round 2: 3.87 / 3.70 / 3.93 ms vs 3.64 ms (profiles/r02_variants.md)
    // A paired-group form (both b1 groups of a b3 at once, full-height stage RA that reads every staged element
    // once, 4-stage ring, stage RC in halves: -32 % stage-RA wavefronts, 4 barriers per plane instead of 5) was
    // correct on every test shape and slower: 3.86 vs 3.75 ms on cfg5 (profiles/r02_variants.md); removed.
    if (prm.n1 == 192) return launch_rec3_rows_n<T, L, 8, 384, 2, 192>(p, prm, s);
    if (prm.n1 == 256) return launch_rec3_rows_n<T, L, 8, 512, 2, 256>(p, prm, s);   // 2-stage ring (224 KB), 128-register cap
    if (prm.n1 > 192) return launch_rec3_rows_n<T, L, 8, 512, 2, 0>(p, prm, s);      // rows of 196..252 elements
    return launch_rec3_rows_n<T, L, 8, 384, 2, 0>(p, prm, s);
}

// tile-shape variants of the bulk synthesis kernel (tuning builds: NDDWT_VARIANT / 100 selects; 0 = default)
template <typename T, int L>
static int launch_rec3_bulk_any(nddwt_plan *p, const Rec3Params<T> &prm, cudaStream_t s)
{
#ifdef NDDWT_TUNING
    if constexpr (L == 8 && sizeof(T) == 8 && Elem<T>::cplx) {
        switch (tuning_variant() / 100 % 10) {
            case 1: return launch_rec3_bulk<T, L, 32, 640, 8, 1>(p, prm, s);
            case 2: return launch_rec3_bulk<T, L, 8, 192, 8, 4>(p, prm, s);
            case 3: return launch_rec3_bulk<T, L, 8, 160, 8, 4>(p, prm, s);
            case 4: return launch_rec3_bulk<T, L, 8, 256, 8, 3>(p, prm, s);
            default: break;
        }
    }
#endif
    return launch_rec3_bulk<T, L, 16, 320, 8, 2>(p, prm, s);
}

// synthesis tile kernel choice: full rows (8-byte elements, rows up to 256 elements, enough CTAs to fill the
// machine; tuning builds: NDDWT_VARIANT=9xx switches it off), else TMA-staged 32-column tiles, else direct
// loads (rows narrower than a staged tile)
template <typename T, int L>
static int launch_rec3_any(nddwt_plan *p, const Rec3Params<T> &prm, cudaStream_t s)
{
    // rows that are not 16-byte multiples (odd n1: 131 x 128 x 30 of mex/mex_test.m:84): direct element loads and
    // element-wise guarded stores -- the staged kernels need 16-byte aligned rows for their bulk copies
    if (prm.n1 % (16 / (int)sizeof(T)) != 0) return launch_rec3_v<T, L, 16, 320, 8, 2, 1>(p, prm, s);
    if constexpr (sizeof(T) == 8) {
        if (tuning_variant() / 100 % 10 != 9) {
            const int rc = launch_rec3_rows<T, L>(p, prm, s);
            if (rc >= 0) return rc;
        }
    }
    if (prm.n1 >= GeoRB<T, L, 16>::W1S) return launch_rec3_bulk_any<T, L>(p, prm, s);
    return launch_rec3_v<T, L, 16, 320, 8, 2>(p, prm, s);
}

template <typename T, int L>
static int launch_rec3(nddwt_plan *p, const void *const *in_bands, void *a_out, cudaStream_t s)
{
    Rec3Params<T> prm;
    for (int b = 0; b < 8; ++b) prm.in[b] = reinterpret_cast<const T *>(in_bands[b]);
    for (int b = 8; b < 16; ++b) prm.in[b] = nullptr;
    prm.out[0] = reinterpret_cast<T *>(a_out);
    prm.out[1] = nullptr;
    prm.n1 = (int)p->dims[0];
    prm.n2 = (int)p->dims[1];
    prm.n3 = (int)p->dims[2];
    prm.s3 = p->dims[0] * p->dims[1];
    prm.s4 = 0;
    prm.nhyp = 1;
#ifdef NDDWT_TUNING
    if constexpr (L == 8 && sizeof(T) == 8 && Elem<T>::cplx) {
        switch (tuning_variant() / 10 % 10) {
            case 1: return launch_rec3_v<T, L, 16, 256, 8, 2>(p, prm, s);
            case 2: return launch_rec3_v<T, L, 16, 320, 4, 2>(p, prm, s);
            case 3: return launch_rec3_v<T, L, 16, 256, 4, 2>(p, prm, s);
            case 4: return launch_rec3_v<T, L, 16, 320, 8, 1>(p, prm, s);
            case 5: if (prm.n1 >= GeoRB<T, L, 16>::W1S) return launch_rec3_bulk<T, L, 16, 320, 8, 2>(p, prm, s); break;
            case 6: if (prm.n1 >= GeoRB<T, L, 16>::W1S) return launch_rec3_bulk<T, L, 16, 256, 8, 2>(p, prm, s); break;
            case 7: if (prm.n1 >= GeoRB<T, L, 16>::W1S) return launch_rec3_bulk<T, L, 16, 320, 4, 2>(p, prm, s); break;
            default: break;
        }
    }
#endif
    return launch_rec3_any<T, L>(p, prm, s);
}

template <typename T>
static int dispatch_rec3(nddwt_plan *p, int L, const void *const *in_bands, void *a_out, cudaStream_t s)
{
    switch (L) {
        case 2: return launch_rec3<T, 2>(p, in_bands, a_out, s);
        case 4: return launch_rec3<T, 4>(p, in_bands, a_out, s);
        case 6: return launch_rec3<T, 6>(p, in_bands, a_out, s);
        case 8: return launch_rec3<T, 8>(p, in_bands, a_out, s);
        default: return 1;
    }
}

static bool uniform_taps(const nddwt_plan *p) { return plan_uniform_taps(p); }

// tap length the fused 3-D / 4-D kernels run with (0: no instantiation): the longest filter of the plan, at most
// db4 (the register ring); slabs (halo planes counted for the TRUE filter of the last dim) need uniform taps
static int fused_L(const nddwt_plan *p, bool slab)
{
    if (slab && (!plan_uniform_taps(p) || p->cur_dil != 1)) return 0;
    const int L = plan_max_taps(p);            // includes the dilation of the level in flight (stretched taps)
    if (!plan_uniform_taps(p) || p->cur_dil != 1)
        for (int i = 0; i < p->ndims; ++i)
            if (p->dims[i] < L) return 0;      // a padded / stretched filter longer than the dimension: generic kernels
    return (L == 2 || L == 4 || L == 6 || L == 8) ? L : 0;
}
// ------------------------------- 4-D path ---------------------------------------------------
template <typename T, int L>
static LastTaps<T, L> make_last_taps(const nddwt_plan *p, bool rec)
{
    LastTaps<T, L> t;
    const AllTaps<double> &src = rec ? p->rec_d : p->dec_d;
    const int d = p->ndims - 1;
    for (int k = 0; k < L; ++k) {
        t.lo[k] = mk_tap(typename TapOf<T>::type(), padded_tap(src.d[d].lo, p->L[d], L, k, p->cur_dil));
        t.hi[k] = mk_tap(typename TapOf<T>::type(), padded_tap(src.d[d].hi, p->L[d], L, k, p->cur_dil));
    }
    return t;
}

static int ensure_fused_scratch(nddwt_plan *p, size_t bytes)
{
    if (p->fused_scratch_bytes < bytes) {
        if (p->fused_scratch) { cudaFree(p->fused_scratch); p->fused_scratch = nullptr; p->fused_scratch_bytes = 0; }
        NDDWT_CUDA(cudaMalloc(&p->fused_scratch, bytes));
        p->fused_scratch_bytes = bytes;
    }
    return 0;
}

// element sub-range of the hyperplane for the passes that are pointwise in dims 1..d-1 (multi-GPU pipelining)
static void plane_range(const nddwt_plan *p, const ZRange &zr, int64_t s4, int64_t *e0, int64_t *en)
{
    *e0 = 0;
    *en = s4;
    if (zr.zn > 0 && p->ndims == 4) {
        const int64_t s3 = p->dims[0] * p->dims[1];
        *e0 = (int64_t)zr.z0 * s3;
        *en = (int64_t)zr.zn * s3;
    }
}

template <typename T, int L>
static int launch_dec_last(nddwt_plan *p, const T *in, const LevelIO &io, T *out_lo, T *out_hi, cudaStream_t s,
                           const ZRange &zr = ZRange())
{
    constexpr int VEC = 16 / (int)sizeof(T);
    const int d = p->ndims - 1;
    int64_t s4 = 1;
    for (int i = 0; i < d; ++i) s4 *= p->dims[i];
    int64_t e0, en;
    plane_range(p, zr, s4, &e0, &en);
    const int64_t nchunks = en / VEC;
    const int n4 = (int)p->dims[d];
    const unsigned grid = (unsigned)((nchunks + 255) / 256);
    const T *hl = reinterpret_cast<const T *>(io.halo_lo), *hh = reinterpret_cast<const T *>(io.halo_hi);
    {
        LaunchTimer lt(p, KIND_DEC_LAST, s);
        k_dec_last<T, L><<<grid, 256, 0, s>>>(in + e0, hl ? hl + e0 : nullptr, hh ? hh + e0 : nullptr, out_lo + e0,
                                             out_hi + e0, nchunks, n4, s4, L / 2 - 1, make_last_taps<T, L>(p, false));
    }
    p->launches++;
    NDDWT_CUDA(cudaGetLastError());
    return 0;
}

template <typename T, int L>
static int launch_rec_last(nddwt_plan *p, const T *u_lo, const T *u_hi, const LevelIO &io, T *out, cudaStream_t s)
{
    constexpr int VEC = 16 / (int)sizeof(T);
    const int d = p->ndims - 1;
    int64_t s4 = 1;
    for (int i = 0; i < d; ++i) s4 *= p->dims[i];
    const int64_t nchunks = s4 / VEC;
    const int n4 = (int)p->dims[d];
    const unsigned grid = (unsigned)((nchunks + 255) / 256);
    {
        LaunchTimer lt(p, KIND_REC_LAST, s);
        k_rec_last<T, L><<<grid, 256, 0, s>>>(u_lo, u_hi, reinterpret_cast<const T *>(io.halo_lo),
                                             reinterpret_cast<const T *>(io.halo_hi), out, nchunks, n4, s4, L / 2,
                                             L / 2 - 1, make_last_taps<T, L>(p, true));
    }
    p->launches++;
    NDDWT_CUDA(cudaGetLastError());
    return 0;
}

template <typename T, int L>
static int launch_rec_last_scatter(nddwt_plan *p, const T *u_lo, const T *u_hi, T *out, T *over_lo, T *over_hi,
                                   cudaStream_t s, const ZRange &zr = ZRange())
{
    constexpr int VEC = 16 / (int)sizeof(T);
    const int d = p->ndims - 1;
    int64_t s4 = 1;
    for (int i = 0; i < d; ++i) s4 *= p->dims[i];
    int64_t e0, en;
    plane_range(p, zr, s4, &e0, &en);
    const int64_t nchunks = en / VEC;
    const int n4 = (int)p->dims[d];
    const unsigned grid = (unsigned)((nchunks + 255) / 256);
    {
        LaunchTimer lt(p, KIND_REC_LAST, s);
        k_rec_last_scatter<T, L><<<grid, 256, 0, s>>>(u_lo + e0, u_hi + e0, out + e0, over_lo + e0, over_hi + e0, nchunks,
                                                     n4, s4, make_last_taps<T, L>(p, true));
    }
    p->launches++;
    NDDWT_CUDA(cudaGetLastError());
    return 0;
}

template <typename T>
static int dispatch_rec_last_scatter(nddwt_plan *p, const void *u_lo, const void *u_hi, void *out, void *over_lo,
                                     void *over_hi, cudaStream_t s, const ZRange &zr);

// part: 0 = whole level; 1 = dim-4 pass only (the one that needs the slab halos);
//       2 = tile pass on the lo4 half (bands 0..7, incl. the approximation); 3 = tile pass on the hi4 half
template <typename T, int L>
static int dec4_level(nddwt_plan *p, const void *a_in, const LevelIO &io, void *const *out_bands, cudaStream_t s,
                      int part = 0, const ZRange &zr = ZRange())
{
    const size_t band_bytes = (size_t)p->numel * p->esize;
    int rc = ensure_fused_scratch(p, 2 * band_bytes);
    if (rc) return rc;
    T *lo4 = reinterpret_cast<T *>(p->fused_scratch);
    T *hi4 = lo4 + p->numel;
    if (part == 0 || part == 1) {
        rc = launch_dec_last<T, L>(p, reinterpret_cast<const T *>(a_in), io, lo4, hi4, s, zr);
        if (rc || part == 1) return rc;
    }
    Dec3Params<T> prm;
    if (zr.zn > 0) { prm.zbase = zr.z0; prm.zcount = zr.zn; }
    prm.halo_lo = nullptr;
    prm.halo_hi = nullptr;
    prm.n1 = (int)p->dims[0];
    prm.n2 = (int)p->dims[1];
    prm.n3 = (int)p->dims[2];
    prm.s3 = p->dims[0] * p->dims[1];
    prm.s4 = prm.s3 * p->dims[2];
    prm.nhyp = (int)p->dims[3];
    if (part == 0) {
        prm.in[0] = lo4;
        prm.in[1] = hi4;
        for (int b = 0; b < 16; ++b) prm.out[b] = reinterpret_cast<T *>(out_bands[b]);
    } else {
        prm.in[0] = (part == 2) ? lo4 : hi4;
        prm.in[1] = nullptr;
        prm.band0 = (part == 2) ? 0 : 8;
        for (int b = 0; b < 8; ++b) prm.out[b] = reinterpret_cast<T *>(out_bands[b + (part == 2 ? 0 : 8)]);
        for (int b = 8; b < 16; ++b) prm.out[b] = nullptr;
    }
    return launch_dec3_any<T, L>(p, prm, s);
}

// part: 0 = both halves; 1 = u_lo half (bands 0..7, needs the approximation band); 2 = u_hi half (bands 8..15)
template <typename T, int L>
static int rec4_stage1(nddwt_plan *p, const void *const *in_bands, T *u_lo, T *u_hi, cudaStream_t s, int part = 0,
                       const ZRange &zr = ZRange())
{
    Rec3Params<T> prm;
    if (zr.zn > 0) { prm.zbase = zr.z0; prm.zcount = zr.zn; }
    if (part == 0) {
        for (int b = 0; b < 16; ++b) prm.in[b] = reinterpret_cast<const T *>(in_bands[b]);
        prm.out[0] = u_lo;
        prm.out[1] = u_hi;
    } else {
        for (int b = 0; b < 8; ++b) prm.in[b] = reinterpret_cast<const T *>(in_bands[b + (part == 1 ? 0 : 8)]);
        for (int b = 8; b < 16; ++b) prm.in[b] = nullptr;
        prm.out[0] = (part == 1) ? u_lo : u_hi;
        prm.out[1] = nullptr;
    }
    prm.n1 = (int)p->dims[0];
    prm.n2 = (int)p->dims[1];
    prm.n3 = (int)p->dims[2];
    prm.s3 = p->dims[0] * p->dims[1];
    prm.s4 = prm.s3 * p->dims[2];
    prm.nhyp = (int)p->dims[3];
    return launch_rec3_any<T, L>(p, prm, s);
}

template <typename T, int L>
static int rec4_level(nddwt_plan *p, const void *const *in_bands, void *a_out, cudaStream_t s)
{
    const size_t band_bytes = (size_t)p->numel * p->esize;
    int rc = ensure_fused_scratch(p, 2 * band_bytes);
    if (rc) return rc;
    T *u_lo = reinterpret_cast<T *>(p->fused_scratch);
    T *u_hi = u_lo + p->numel;
    rc = rec4_stage1<T, L>(p, in_bands, u_lo, u_hi, s);
    if (rc) return rc;
    LevelIO none;
    return launch_rec_last<T, L>(p, u_lo, u_hi, none, reinterpret_cast<T *>(a_out), s);
}

#define NDDWT_L_SWITCH(L_, CALL)                \
    switch (L_) {                               \
        case 2: { constexpr int LL = 2; return CALL; } \
        case 4: { constexpr int LL = 4; return CALL; } \
        case 6: { constexpr int LL = 6; return CALL; } \
        case 8: { constexpr int LL = 8; return CALL; } \
        default: return 1;                      \
    }

template <typename T>
static int dispatch_dec4(nddwt_plan *p, const void *a_in, const LevelIO &io, void *const *out_bands, cudaStream_t s,
                         int part = 0, const ZRange &zr = ZRange())
{
    NDDWT_L_SWITCH(plan_max_taps(p), (dec4_level<T, LL>(p, a_in, io, out_bands, s, part, zr)));
}
template <typename T>
static int dispatch_rec4(nddwt_plan *p, const void *const *in_bands, void *a_out, cudaStream_t s)
{
    NDDWT_L_SWITCH(plan_max_taps(p), (rec4_level<T, LL>(p, in_bands, a_out, s)));
}
template <typename T>
static int dispatch_rec4_stage1(nddwt_plan *p, const void *const *in_bands, void *u_lo, void *u_hi, cudaStream_t s,
                                int part = 0, const ZRange &zr = ZRange())
{
    NDDWT_L_SWITCH(plan_max_taps(p), (rec4_stage1<T, LL>(p, in_bands, reinterpret_cast<T *>(u_lo), reinterpret_cast<T *>(u_hi), s, part, zr)));
}
template <typename T>
static int dispatch_rec_last(nddwt_plan *p, const void *u_lo, const void *u_hi, const LevelIO &io, void *a_out,
                             cudaStream_t s)
{
    NDDWT_L_SWITCH(p->L[p->ndims - 1], (launch_rec_last<T, LL>(p, reinterpret_cast<const T *>(u_lo),
                                                                 reinterpret_cast<const T *>(u_hi), io,
                                                                 reinterpret_cast<T *>(a_out), s)));
}

template <typename T>
static int dispatch_rec_last_scatter(nddwt_plan *p, const void *u_lo, const void *u_hi, void *out, void *over_lo,
                                     void *over_hi, cudaStream_t s, const ZRange &zr)
{
    NDDWT_L_SWITCH(p->L[p->ndims - 1],
                   (launch_rec_last_scatter<T, LL>(p, reinterpret_cast<const T *>(u_lo), reinterpret_cast<const T *>(u_hi),
                                                   reinterpret_cast<T *>(out), reinterpret_cast<T *>(over_lo),
                                                   reinterpret_cast<T *>(over_hi), s, zr)));
}

// int-range guards: the tile kernels index planes with ints and put every CTA in grid.x
static bool fused_ranges_ok(const nddwt_plan *p)
{
    if (p->dims[0] * p->dims[1] >= (int64_t)1 << 30) return false;
    if (p->dims[2] > (1 << 24)) return false;
    if (p->ndims == 4 && p->dims[3] > (1 << 20)) return false;
    // CTAs: tiles x z-chunks x (2 x hyperplanes) must fit grid.x
    const int64_t tiles = ((p->dims[0] + 31) / 32) * ((p->dims[1] + 7) / 8);
    const int64_t batches = p->ndims == 4 ? 2 * p->dims[3] : 1;
    return tiles * ((p->dims[2] + 3) / 4) * batches < (int64_t)0x7fffffff;
}

static bool rows_unaligned(const nddwt_plan *p) { return p->dims[0] % (16 / (int64_t)p->esize) != 0; }
// analysis with rows that are not 16-byte multiples has no shrink epilogue: such plans take the generic kernels + in-place pass
static bool dec_rows_ok(const nddwt_plan *p) { return !(p->shrink_mode && rows_unaligned(p)); }

static bool fused_geometry_ok(const nddwt_plan *p)
{
    if (p->ndims < 3 || p->batch != 1) return false;
    if (!fused_ranges_ok(p)) return false;
    const int FL = fused_L(p, false);
    if (FL == 0 || p->dims[0] < FL || p->dims[1] < FL) return false;
    // 4-D: the last-dim passes move 16-byte chunks of the hyperplane; rows themselves may be odd (element-wise
    // tile-kernel instantiations)
    if (p->ndims == 4 && (p->dims[0] * p->dims[1] * p->dims[2]) % (16 / (int64_t)p->esize) != 0) return false;
    return true;
}

#define NDDWT_T_SWITCH(p_, CALL)                                   \
    switch ((p_)->dtype) {                                         \
        case NDDWT_C64: { using TT = float2; return CALL; }        \
        case NDDWT_F32: { using TT = float; return CALL; }         \
        case NDDWT_F64: { using TT = double; return CALL; }        \
        case NDDWT_C128: { using TT = double2; return CALL; }      \
        default: return 1;                                         \
    }

int fused_rec_stage1(nddwt_plan *p, int dil, const void *const *in_bands, void *u_lo, void *u_hi, cudaStream_t s,
                     int part, const ZRange &zr)
{
    if (dil != 1 || !uniform_taps(p) || p->ndims != 4 || !fused_geometry_ok(p)) return 1;
    NDDWT_T_SWITCH(p, (dispatch_rec4_stage1<TT>(p, in_bands, u_lo, u_hi, s, part, zr)));
}

int fused_rec_stage2_scatter(nddwt_plan *p, int dil, const void *u_lo, const void *u_hi, void *out, void *over_lo,
                             void *over_hi, cudaStream_t s, const ZRange &zr)
{
    if (dil != 1 || p->ndims != 4 || !fused_geometry_ok(p)) return 1;
    NDDWT_T_SWITCH(p, (dispatch_rec_last_scatter<TT>(p, u_lo, u_hi, out, over_lo, over_hi, s, zr)));
}

int accumulate_elems(nddwt_plan *p, void *dst, const void *src, int64_t nelem, cudaStream_t s)
{
    // element type is irrelevant for an elementwise add: count real scalars
    const bool dbl = (p->dtype == NDDWT_F64 || p->dtype == NDDWT_C128);
    const int64_t n = nelem * (int64_t)(p->esize / (dbl ? 8 : 4));
    const unsigned grid = (unsigned)std::min<int64_t>((n + 255) / 256, 148 * 16);
    {
        LaunchTimer lt(p, KIND_REC_LAST, s);
        if (dbl) k_accumulate<double><<<grid, 256, 0, s>>>(reinterpret_cast<double *>(dst), reinterpret_cast<const double *>(src), n);
        else k_accumulate<float><<<grid, 256, 0, s>>>(reinterpret_cast<float *>(dst), reinterpret_cast<const float *>(src), n);
    }
    p->launches++;
    NDDWT_CUDA(cudaGetLastError());
    return 0;
}

// dst plane += sum of up to ACC_MAXS source planes, for several planes in one launch (the adds of the
// scatter-form synthesis exchange: every plane is read and written once however many overhangs land on it)
__device__ __forceinline__ void acc_add(float &a, float b) { a += b; }
__device__ __forceinline__ void acc_add(double &a, double b) { a += b; }
__device__ __forceinline__ void acc_add(float4 &a, float4 b) { a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w; }
__device__ __forceinline__ void acc_add(double2 &a, double2 b) { a.x += b.x; a.y += b.y; }

template <typename V>
__global__ void __launch_bounds__(256) k_accumulate_planes(const __grid_constant__ AccParams prm, int64_t nvec)
{
    const AccItem &it = prm.item[blockIdx.y];      // indexed straight out of the parameter bank
    V *dst = reinterpret_cast<V *>(it.dst);
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += stride) {
        V a = dst[i];
        for (int k = 0; k < it.ns; ++k) {
            acc_add(a, reinterpret_cast<const V *>(it.src[k])[i]);
        }
        dst[i] = a;
    }
}

int accumulate_planes(nddwt_plan *p, const AccParams &prm, int64_t plane_elems, cudaStream_t s)
{
    if (prm.n < 1) return 0;
    const bool dbl = (p->dtype == NDDWT_F64 || p->dtype == NDDWT_C128);
    const int64_t bytes = plane_elems * (int64_t)p->esize;
    bool vec = bytes % 16 == 0;
    for (int i = 0; i < prm.n && vec; ++i) {
        vec = (reinterpret_cast<uintptr_t>(prm.item[i].dst) & 15) == 0;
        for (int k = 0; k < prm.item[i].ns && vec; ++k) vec = (reinterpret_cast<uintptr_t>(prm.item[i].src[k]) & 15) == 0;
    }
    const int64_t n = vec ? bytes / 16 : bytes / (dbl ? 8 : 4);
    const unsigned gx = (unsigned)std::max<int64_t>(1, std::min<int64_t>((n + 255) / 256, (148 * 16 + prm.n - 1) / prm.n));
    const dim3 grid(gx, (unsigned)prm.n);
    {
        LaunchTimer lt(p, KIND_REC_LAST, s);
        if (vec) {
            if (dbl) k_accumulate_planes<double2><<<grid, 256, 0, s>>>(prm, n);   // 16 bytes = (x, y)
            else k_accumulate_planes<float4><<<grid, 256, 0, s>>>(prm, n);
        } else {
            if (dbl) k_accumulate_planes<double><<<grid, 256, 0, s>>>(prm, n);
            else k_accumulate_planes<float><<<grid, 256, 0, s>>>(prm, n);
        }
    }
    p->launches++;
    NDDWT_CUDA(cudaGetLastError());
    return 0;
}

bool fused_is_separable(const nddwt_plan *p)
{
    if (p->kernel_mode != 0 || p->batch != 1 || p->ndims != 4 || !uniform_taps(p) || !fused_geometry_ok(p)) return false;
    if (rows_unaligned(p)) return false;      // multi-GPU part-wise schedule: aligned rows only
    for (int j = 0; j < NDDWT_MAX_LEVELS; ++j)
        if (p->dil[j] != 1) return false;
    return p->L[0] == 2 || p->L[0] == 4 || p->L[0] == 6 || p->L[0] == 8;
}

// 4-D analysis level in parts (multi-GPU overlap); returns 1 when the plan has no fused 4-D path
int fused_dec_level_part(nddwt_plan *p, int dil, int part, const void *a_in, const LevelIO &io,
                         void *const *out_bands, cudaStream_t s, const ZRange &zr)
{
    if (dil != 1 || !uniform_taps(p) || p->ndims != 4 || !fused_geometry_ok(p) || !dec_rows_ok(p)) return 1;
    NDDWT_T_SWITCH(p, (dispatch_dec4<TT>(p, a_in, io, out_bands, s, part, zr)));
}

int fused_rec_stage2(nddwt_plan *p, int dil, const void *u_lo, const void *u_hi, const LevelIO &io, void *a_out,
                     cudaStream_t s)
{
    if (dil != 1 || p->ndims != 4 || !fused_geometry_ok(p)) return 1;
    NDDWT_T_SWITCH(p, (dispatch_rec_last<TT>(p, u_lo, u_hi, io, a_out, s)));
}

int fused_dec_level(nddwt_plan *p, int dil, const void *a_in, const LevelIO &io, void *const *out_bands,
                    cudaStream_t s)
{
    if (dil < 1 || dil > 4) return 1;
    DilScope ds(p, dil);                       // a-trous levels: the same kernels with stretched taps (nddwt_plan.h)
    const int FL = fused_L(p, io.halo_lo != nullptr || io.halo_hi != nullptr);
    if (FL == 0 || !dec_rows_ok(p)) return 1;
    if (p->ndims == 3) {
        if (!fused_ranges_ok(p)) return 1;
        if (p->dims[0] < FL || p->dims[1] < FL) return 1;
        switch (p->dtype) {
            case NDDWT_C64: return dispatch_dec3<float2>(p, FL, a_in, io, out_bands, s);
            case NDDWT_F32: return dispatch_dec3<float>(p, FL, a_in, io, out_bands, s);
            case NDDWT_F64: return dispatch_dec3<double>(p, FL, a_in, io, out_bands, s);
            case NDDWT_C128: return dispatch_dec3<double2>(p, FL, a_in, io, out_bands, s);
        }
    }
    if (p->ndims == 4 && fused_geometry_ok(p)) {
        NDDWT_T_SWITCH(p, (dispatch_dec4<TT>(p, a_in, io, out_bands, s)));
    }
    return 1;
}

int fused_rec_level(nddwt_plan *p, int dil, const void *const *in_bands, void *a_out, cudaStream_t s)
{
    if (dil < 1 || dil > 4) return 1;
    DilScope ds(p, dil);
    const int FL = fused_L(p, false);
    if (FL == 0) return 1;
    if (p->ndims == 3) {
        if (!fused_ranges_ok(p)) return 1;
        if (p->dims[0] < FL || p->dims[1] < FL) return 1;
        switch (p->dtype) {
            case NDDWT_C64: return dispatch_rec3<float2>(p, FL, in_bands, a_out, s);
            case NDDWT_F32: return dispatch_rec3<float>(p, FL, in_bands, a_out, s);
            case NDDWT_F64: return dispatch_rec3<double>(p, FL, in_bands, a_out, s);
            case NDDWT_C128: return dispatch_rec3<double2>(p, FL, in_bands, a_out, s);
        }
    }
    if (p->ndims == 4 && fused_geometry_ok(p)) {
        NDDWT_T_SWITCH(p, (dispatch_rec4<TT>(p, in_bands, a_out, s)));
    }
    return 1;
}

}  // namespace nddwt
