/*
 * nddwt_b200.h -- C ABI of the B200-native non-decimated wavelet library (libnddwt_b200.so).
 *
 * This is the drop-in boundary for the hot path of arg-min-x/Non-Decimated_Wavelets:
 * it replaces mex/nddwt.h (the native core behind nd_dwt_mex) and the parts of the MATLAB
 * classes that build and apply the stored filters.  Plain C: pointers, sizes, int return codes.
 * All arrays are dense COLUMN-MAJOR (MATLAB order, dim 1 contiguous); complex numbers are
 * interleaved (re, im) pairs.  The coefficient stack is [dims..., nb] with
 * nb = 1 + level*(2^ndims - 1) bands, deepest level first (mex/nddwt.c:209-210,226;
 * mex/nd_dwt_mex.c:79-83): slot 0 = a_J, slots 1..2^d-1 = d_J, ..., last 2^d-1 = d_1;
 * band b = sum_i b_i 2^(i-1), b_i = 1 for the high-pass along dim i.
 *
 * Every function returns 0 on success or a negative nddwt_status; nddwt_last_error() returns
 * the message of the calling thread's last failure.  There is no CPU path: every entry point
 * that computes needs a CUDA device and fails with NDDWT_ERR_CUDA without one.
 */
#ifndef NDDWT_B200_H
#define NDDWT_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define NDDWT_API __attribute__((visibility("default")))
#else
#define NDDWT_API
#endif

#define NDDWT_MAX_DIMS 4
#define NDDWT_MAX_LEVELS 16

typedef struct nddwt_plan nddwt_plan;

/* element types: the reference's 'precision' x real/complex (nd_dwt_1D.m:124-126,144-148) */
typedef enum {
    NDDWT_F32 = 0,   /* real single    */
    NDDWT_F64 = 1,   /* real double    */
    NDDWT_C64 = 2,   /* complex single */
    NDDWT_C128 = 3   /* complex double */
} nddwt_dtype;

typedef enum {
    NDDWT_OK = 0,
    NDDWT_ERR_ARG = -1,       /* bad argument (sizes, level, dtype, null pointer) */
    NDDWT_ERR_WAVELET = -2,   /* "Unknown Wavelet Name"           (wave_filters.m:158-159) */
    NDDWT_ERR_SHORT_DIM = -3, /* data dim shorter than the filter (nd_dwt_2D.m:271-277)     */
    NDDWT_ERR_CUDA = -4,      /* CUDA runtime failure / no device */
    NDDWT_ERR_NOMEM = -5,
    NDDWT_ERR_SIZE = -6       /* "FIlter size and image size not consistant" (nd_dwt_mex.c:36-51,124-127) */
} nddwt_status;

NDDWT_API const char *nddwt_last_error(void);
NDDWT_API const char *nddwt_version(void);

/* [low_d, hi_d] = wave_filters(wname)   -- replaces Functions/wave_filters.m:1-174.
 * Writes *len taps (2..20) into low_d and hi_d (each must hold 20 doubles). */
NDDWT_API int nddwt_wave_filters(const char *wname, double *low_d, double *hi_d, int *len);

/* nb = 2^d + (2^d - 1)(level - 1)       -- mex/nd_dwt_mex.c:83 */
NDDWT_API int64_t nddwt_num_bands(int ndims, int level);
/* level from the band count             -- nd_dwt_1D.m:213, nd_dwt_2D.m:215, nd_dwt_3D.m:217, nd_dwt_4D.m:213.
 * Returns 0 when nb matches no level. */
NDDWT_API int nddwt_infer_level(int ndims, int64_t nb);

/* Plan = the "stored filter" object: replaces the constructor + get_filters of
 * nd_dwt_{1,2,3,4}D.m (nd_dwt_1D.m:79-133,257-290 ... nd_dwt_4D.m:79-134,255-391) and
 * init_fftw_plan (mex/nddwt.c:15-61).  Holds per-dim taps (device-ready), scratch for the
 * intermediate approximation bands, and is reused across any number of dec/rec calls.
 *   ndims    1..4;  dims[ndims] sizes in MATLAB order
 *   wnames   ndims strings "db1".."db10" (one per dim; harr_nddwt_* == "db1")
 *   dtype    nddwt_dtype of x and of the coefficients
 *   pres_l2_norm  0/1: scale 2^(-d/2) per level so that ||coeffs|| = ||x|| (nd_dwt_2D.m:295-299)
 *   device   CUDA device ordinal */
NDDWT_API int nddwt_plan_create(nddwt_plan **plan, int ndims, const int64_t *dims,
                      const char *const *wnames, int dtype, int pres_l2_norm, int device);
NDDWT_API int nddwt_plan_destroy(nddwt_plan *plan);

/* Opt-in a-trous mode: dilation of the taps at level j (1-based) is dil[j-1].  Default (and
 * reference parity, nd_dwt_2D.m:183 / nddwt.c:214-228): 1 at every level.
 * A dilated level runs the same fused kernels with its taps stretched about their centre (an L-tap filter at
 * dilation s = an L s-tap filter with zeros in between): 3-D / 4-D tile kernels while L s <= 8 (Haar at 1, 2, 4;
 * db2 at 1, 2), 2-D kernels and the hybrid level while L s <= 20, the generic separable kernels beyond that, for
 * slabs, and in the 1-D cascade. */
NDDWT_API int nddwt_plan_set_dilations(nddwt_plan *plan, const int *dil, int nlevels);

/* Extension (the reference has no batch API): the arrays carry one extra trailing
 * dimension of `batch` independent signals/images: x is [dims..., batch], the coefficient stack
 * [dims..., batch, nb].  1-D batches run the one-launch cascade kernel; batched 2-D..4-D plans run the
 * generic separable kernels. */
NDDWT_API int nddwt_plan_set_batch(nddwt_plan *plan, int64_t batch);

/* Fused coefficient-domain shrink: the step between dec and rec of an iterative
 * compressed-sensing loop, x <- rec(shrink(dec(x))) (README.md:2 "such as in an iterative algorithm").
 * mode 1 = soft threshold: every later dec of this plan (any entry point: nddwt_dec, *_host, slab levels, the
 * multi-GPU plan) applies  c <- c * max(0, 1 - t/|c|)  (real: sign(c) max(|c| - t, 0))  to the DETAIL bands as
 * the analysis kernels store them, so the coefficient stack is written once, already thresholded -- instead
 * of a separate pass that reads and writes all nb bands again.  thr is a table [nlevels][2^ndims]:
 * thr[(j-1) * 2^ndims + b] is the threshold of band b of level j (j = 1 finest); entries b = 0 are ignored
 * (the approximation band is never thresholded); levels beyond nlevels use 0.  mode 0 switches it off. */
NDDWT_API int nddwt_plan_set_shrink(nddwt_plan *plan, int mode, const double *thr, int nlevels);

/* Selects the kernel family: 0 = auto (fused kernels where an instantiation exists, generic
 * otherwise), 1 = force the generic separable kernels.  Both run on the GPU. */
NDDWT_API int nddwt_plan_set_kernel_mode(nddwt_plan *plan, int mode);
/* Named integer parameters of a plan.  "rows_min_ctas": the full-row synthesis kernel (one CTA per SM) is
 * chosen when a level offers at least this many CTAs (default 118); the tests set 0 to reach it on small shapes. */
NDDWT_API int nddwt_plan_set_param(nddwt_plan *plan, const char *name, int64_t value);
/* Number of kernel launches issued by the plan so far (bench.py's gpu_launches). */
NDDWT_API int64_t nddwt_plan_launch_count(const nddwt_plan *plan);
/* Per-kernel timing with CUDA events on the launch stream (bench.py's roofline): switch on, run any
 * number of dec/rec calls, then read the accumulated device time and launch count of one kernel
 * kind (0 analysis 3-D tile kernel, 1 synthesis 3-D tile kernel, 2 analysis last-dim pass,
 * 3 synthesis last-dim pass incl. the adds of the multi-GPU exchange, 4 generic separable pass, 5 halo pushes of
 * a multi-GPU plan).  Reading synchronises and clears that kind. */
NDDWT_API int nddwt_plan_profile(nddwt_plan *plan, int on);
NDDWT_API int nddwt_plan_kernel_time(nddwt_plan *plan, int kind, double *total_ms, int64_t *count);
/* Kernel family of the last dec/rec level of this plan: 1 = fused level kernels, 2 = hybrid (generic separable passes
 * along the outer dimensions, the fused 2-D kernels over all (dim 1, dim 2) planes: db5..db10 in 3-D/4-D, batched 2-D..4-D
 * arrays), 0 = generic separable passes only (a-trous dilation, forced by nddwt_plan_set_kernel_mode). */
NDDWT_API int nddwt_plan_last_path(const nddwt_plan *plan);
/* Which synthesis tile kernel the last fused 3-D/4-D level of this plan launched (tests and profiles):
 * 0 none yet, 1 direct-load tiles (k_rec3_fused), 2 TMA-staged 32-column tiles (k_rec3_bulk),
 * 4 full-row tiles (k_rec3_rows). */
NDDWT_API int nddwt_plan_last_synthesis_kernel(const nddwt_plan *plan);

/* y = dec(x, level)   -- replaces nd_dwt_dec / nd_dwt_dec_1level (mex/nddwt.c:98-139,189-239)
 *                        together with the fftn the MATLAB side does first (nd_dwt_2D.m:156).
 * x_dev: prod(dims) elements; coeffs_dev: prod(dims)*nb elements; both DEVICE pointers on the
 * plan's device.  Asynchronous on `stream` (a cudaStream_t, may be NULL).  x is never written. */
NDDWT_API int nddwt_dec(nddwt_plan *plan, const void *x_dev, void *coeffs_dev, int level, void *stream);

/* The unfused form of the shrink: applies the plan's threshold table (nddwt_plan_set_shrink) in place to the
 * detail bands of an existing coefficient stack (one extra read + write of every thresholded band). */
NDDWT_API int nddwt_shrink(nddwt_plan *plan, void *coeffs_dev, int level, void *stream);

/* x = rec(y)          -- replaces nd_dwt_rec / nd_dwt_rec_1level (mex/nddwt.c:142-186,242-292).
 * Unlike the reference (nddwt.c:163,264-265) the coefficient stack is never modified.
 * nddwt_dec / nddwt_rec / nddwt_shrink only launch kernels on `stream` once the plan's scratch exists (after the
 * first call of that kind): no allocation, no synchronisation, no host round trip -- they can be captured into a
 * CUDA graph and replayed (tests/test_gpu_parity.py::test_cuda_graph_capture_and_replay). */
NDDWT_API int nddwt_rec(nddwt_plan *plan, const void *coeffs_dev, void *x_dev, int level, void *stream);

/* Same two calls with HOST buffers: the shape of nd_dwt_mex(x, f, dir, level, pres_l2)
 * (mex/nd_dwt_mex.c:8-153) for host mxArrays.  Copies in, runs the kernels, copies out;
 * synchronous.  Host memory may be pageable or pinned.
 * 2-D ... 4-D arrays with two or more levels are LEVEL-STREAMED: the device holds x, the approximation
 * ping-pong and two buffers of 2^d - 1 detail bands (not the whole coefficient stack), and the bands of
 * one level cross PCIe while the next level computes -- (3 + 2 (2^d - 1)) N e of device memory (+ 2 N e of
 * scratch in 4-D) instead of (3 + nb) N e. */
NDDWT_API int nddwt_dec_host(nddwt_plan *plan, const void *x_host, void *coeffs_host, int level);
NDDWT_API int nddwt_rec_host(nddwt_plan *plan, const void *coeffs_host, void *x_host, int level);

/* ---- slab interface (multi-GPU: one process per GPU, slabs along the LAST dimension) ----
 * A slab plan is created with nddwt_plan_create_slab: local_dims[ndims-1] = number of LOCAL planes,
 * global_last_dim = the full extent of that dimension (the filter-length check of
 * nd_dwt_2D.m:271-277 applies to the global extent; a slab may be thinner than the filter).  One level of analysis of
 * the local planes needs (L/2 - 1)*dil planes below and (L/2)*dil planes above the slab
 * (L = taps of the last dim); the caller supplies them (neighbour exchange) as two dense
 * buffers.  halo_lo holds the planes immediately below the slab in ascending order, halo_hi the
 * planes immediately above.  Passing NULL for both means "periodic within the slab" (1 GPU). */
NDDWT_API int nddwt_plan_create_slab(nddwt_plan **plan, int ndims, const int64_t *local_dims, int64_t global_last_dim,
                           const char *const *wnames, int dtype, int pres_l2_norm, int device);
NDDWT_API int nddwt_halo_planes(const nddwt_plan *plan, int level_index /*1-based*/, int *below, int *above);

/* One analysis level: a_in (local planes) -> 2^d bands.  out_bands[b] are 2^d device pointers,
 * each to a dense local-slab-sized array. */
NDDWT_API int nddwt_dec_level_slab(nddwt_plan *plan, int level_index, const void *a_in,
                         const void *halo_lo, const void *halo_hi,
                         void *const *out_bands, void *stream);

/* The same level in parts, so that the caller can overlap the halo exchange of the NEXT level with
 * compute (part 0 = whole level; 1 = the last-dim pass, the only one that reads the halos;
 * 2 = the tile pass that produces bands 0..2^(d-1)-1, including the approximation band the next
 * level's exchange needs; 3 = the tile pass for the remaining bands).  On plans without a
 * separable fused path part 1 runs the whole level and parts 2/3 are no-ops. */
/* 1 when the part-wise calls below really split the level (fused 4-D path), else 0. */
NDDWT_API int nddwt_plan_is_separable(const nddwt_plan *plan);
NDDWT_API int nddwt_dec_level_slab_part(nddwt_plan *plan, int level_index, int part, const void *a_in,
                              const void *halo_lo, const void *halo_hi, void *const *out_bands, void *stream);

/* One synthesis level, split at the slab dimension (the adjoint of the analysis exchange):
 *  stage 1 (local):  u_lo/u_hi[local planes] = synthesis over dims 1..d-1 of the 2^d bands,
 *                    summed over their band bits (no halo needed);
 *  stage 2:          a_out = synthesis along the last dim from u_lo/u_hi, which needs
 *                    (L/2)*dil planes of u below and (L/2 - 1)*dil above (halo buffers hold
 *                    u_lo planes then u_hi planes, each group ascending). */
NDDWT_API int nddwt_rec_level_slab_stage1(nddwt_plan *plan, int level_index, const void *const *in_bands,
                                void *u_lo, void *u_hi, void *stream);
/* stage 1 in halves: part 0 = both; 1 = u_lo (bands 0..2^(d-1)-1, needs the approximation band);
 * 2 = u_hi (pure detail bands -- can run before the previous level has finished). */
NDDWT_API int nddwt_rec_level_slab_stage1_part(nddwt_plan *plan, int level_index, int part,
                                     const void *const *in_bands, void *u_lo, void *u_hi, void *stream);
NDDWT_API int nddwt_rec_level_slab_stage2(nddwt_plan *plan, int level_index, const void *u_lo, const void *u_hi,
                                const void *halo_lo, const void *halo_hi, void *a_out, void *stream);

/* Scatter form of stage 2 (separable plans only): reads only the local u planes, writes the local
 * partial result to a_out and the partial sums that belong to neighbouring slabs to over_lo
 * (L/2-1 planes: global planes start-(L/2-1)..start-1) and over_hi (L/2 planes: end..end+L/2-1).
 * The caller sends the overhangs to their owners, which add them with nddwt_accumulate.  Halves the
 * synthesis halo traffic compared with exchanging u_lo and u_hi. */
NDDWT_API int nddwt_rec_level_slab_stage2_scatter(nddwt_plan *plan, int level_index, const void *u_lo, const void *u_hi,
                                        void *a_out, void *over_lo, void *over_hi, void *stream);
/* dst[i] += src[i] for nelem elements of the plan's dtype (device pointers). */
NDDWT_API int nddwt_accumulate(nddwt_plan *plan, void *dst, const void *src, int64_t nelem, void *stream);

/* ---- multi-GPU plan: the plan owns the peer mapping and the halo exchange --------
 * The array is split in contiguous slabs along the LAST dimension (at most +-1 plane ragged); rank r holds
 * its planes of x ([dims[0..d-2], count_r]) and of every subband (coefficient slab [dims[0..d-2], count_r, nb],
 * band = slowest dim, same band order as nddwt_dec).  Per level the ranks push the halo planes their
 * neighbours need straight into peer memory with copy-engine peer copies over NVLink (analysis: L/2-1 planes
 * below, L/2 above of the approximation band; synthesis: the L-1 overhang planes of partial sums of the
 * scatter-form last-dim pass), overlapped with the tile pass that does not feed the exchange.  No NCCL.
 *
 * (1) one process drives all GPUs (C / MATLAB callers): */
typedef struct nddwt_mplan nddwt_mplan;
NDDWT_API int nddwt_mplan_create(nddwt_mplan **mplan, int ndims, const int64_t *dims, const char *const *wnames,
                                 int dtype, int pres_l2_norm, int ngpus, const int *devices /* NULL: 0..ngpus-1 */);
/* (2) one process per GPU (torchrun): create the rank's plan, all-gather the export blobs
 *     (nddwt_mplan_export_size() bytes each, rank order) with any host-side transport, import them.  The
 *     blobs carry CUDA IPC handles of the halo inboxes: all ranks must run on the same node. */
NDDWT_API int nddwt_mplan_create_rank(nddwt_mplan **mplan, int ndims, const int64_t *dims, const char *const *wnames,
                                      int dtype, int pres_l2_norm, int rank, int world, int device);
NDDWT_API int64_t nddwt_mplan_export_size(void);
NDDWT_API int nddwt_mplan_export(nddwt_mplan *mplan, void *blob);
NDDWT_API int nddwt_mplan_import(nddwt_mplan *mplan, const void *blobs);
NDDWT_API int nddwt_mplan_destroy(nddwt_mplan *mplan);

NDDWT_API int nddwt_mplan_world(const nddwt_mplan *mplan);
NDDWT_API int nddwt_mplan_num_local(const nddwt_mplan *mplan);   /* ranks this process drives: ngpus or 1 */
NDDWT_API int nddwt_mplan_slab(const nddwt_mplan *mplan, int rank, int64_t *start, int64_t *count);
NDDWT_API int nddwt_mplan_is_separable(const nddwt_mplan *mplan); /* 1: overlapped scatter schedule (fused 4-D path) */
NDDWT_API int nddwt_mplan_set_dilations(nddwt_mplan *mplan, const int *dil, int nlevels);  /* a-trous: halos (L-1)*dil */
NDDWT_API int nddwt_mplan_set_kernel_mode(nddwt_mplan *mplan, int mode);
NDDWT_API int nddwt_mplan_set_shrink(nddwt_mplan *mplan, int mode, const double *thr, int nlevels);   /* nddwt_plan_set_shrink on every rank */
/* "z_chunks" (0..64, default 0 = chosen from the halo : slab ratio): fused 4-D plans issue a level in that many
 * chunks of dim 3, so that the halo planes of one chunk travel while the next chunk computes; "comm_streams" (1..4,
 * default 1): every pushed run of planes is cut in that many pieces which travel on different streams / copy
 * engines at once; other names are forwarded to the per-rank plans. */
NDDWT_API int nddwt_mplan_set_param(nddwt_mplan *mplan, const char *name, int64_t value);

/* y = dec(x, level) / x = rec(y) on slabs.  x_slabs / coeff_slabs: one DEVICE pointer per local rank (in
 * rank order) on that rank's device; streams: one cudaStream_t per local rank (a NULL entry is the legacy
 * default stream), or a NULL array for plan-owned streams.  Asynchronous; nddwt_mplan_sync waits for every local rank.  In the one-process-per-GPU form
 * every rank must make the same calls in the same order. */
NDDWT_API int nddwt_mplan_dec(nddwt_mplan *mplan, const void *const *x_slabs, void *const *coeff_slabs, int level,
                              void *const *streams);
NDDWT_API int nddwt_mplan_rec(nddwt_mplan *mplan, const void *const *coeff_slabs, void *const *x_slabs, int level,
                              void *const *streams);
NDDWT_API int nddwt_mplan_sync(nddwt_mplan *mplan);
/* The same two calls with whole HOST arrays (one-process plans): the shape of nd_dwt_mex(x, f, dir, level, l2)
 * for a caller that owns every GPU of the box.  The slabs of a column-major array along its last dimension are
 * contiguous blocks, so every GPU copies its own part of x and of every band over its own PCIe link.
 * Synchronous. */
NDDWT_API int nddwt_mplan_dec_host(nddwt_mplan *mplan, const void *x_host, void *coeffs_host, int level);
NDDWT_API int nddwt_mplan_rec_host(nddwt_mplan *mplan, const void *coeffs_host, void *x_host, int level);
NDDWT_API int64_t nddwt_mplan_launch_count(const nddwt_mplan *mplan);
NDDWT_API int64_t nddwt_mplan_halo_bytes(const nddwt_mplan *mplan);     /* bytes pushed to peers so far (local ranks) */
/* nddwt_plan_profile / nddwt_plan_kernel_time of one local rank's plan; kind 5 = the pushes of a level
 * (flag wait + peer copies + flag signal) timed on the rank's comm stream. */
NDDWT_API int nddwt_mplan_profile(nddwt_mplan *mplan, int on);
NDDWT_API int nddwt_mplan_kernel_time(nddwt_mplan *mplan, int local_index, int kind, double *total_ms, int64_t *count);
NDDWT_API int nddwt_mplan_wait_timeouts(const nddwt_mplan *mplan);      /* peer flag waits that gave up (0 when healthy) */

/* Host-only routing query (no device): the planes rank `rank` of `world` needs below (which = 0) or above
 * (which = 1) its slab for a halo of (below, above) planes of a periodic last dimension of n_last planes, as
 * runs {owner rank, first local plane, count, first halo slot} written to runs4[4 * k ..]; returns the
 * number of runs (halos wider than a slab come from several ranks, possibly from the rank itself). */
NDDWT_API int nddwt_slab_route(int64_t n_last, int world, int rank, int which, int64_t below, int64_t above,
                               int64_t *runs4, int cap);

#ifdef __cplusplus
}
#endif
#endif /* NDDWT_B200_H */
