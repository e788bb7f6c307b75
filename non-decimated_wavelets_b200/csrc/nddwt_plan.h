// nddwt_plan.h -- internal plan object of libnddwt_b200 (not part of the C ABI).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>
#include <vector>
#include "nddwt_common.cuh"

struct nddwt_plan {
    int ndims = 0;
    int64_t dims[NDDWT_MAX_DIMS] = {1, 1, 1, 1};
    int L[NDDWT_MAX_DIMS] = {0, 0, 0, 0};
    int dtype = NDDWT_C64;
    int pres_l2 = 0;
    int device = 0;
    int64_t numel = 0;     // prod(dims) * batch
    int64_t batch = 1;     // trailing batch dimension (extension: the reference has no batch API)
    size_t esize = 0;      // bytes per element

    // unscaled wave_filters output per dim
    double lo[NDDWT_MAX_DIMS][NDDWT_MAXL];
    double hi[NDDWT_MAX_DIMS][NDDWT_MAXL];
    // per-dim taps with the level scale folded in: analysis x 2^(-1/2) per dim when pres_l2_norm,
    // synthesis x 2^(-1/2) (pres_l2_norm) or x 1/2 (otherwise) per dim  (nd_dwt_2D.m:230-232,295-299)
    nddwt::AllTaps<double> dec_d, rec_d;
    nddwt::AllTaps<float> dec_f, rec_f;

    int dil[NDDWT_MAX_LEVELS];
    int kernel_mode = 0;   // 0 auto, 1 generic only
    int last_path = 0;     // 1 fused, 0 generic, 2 hybrid (generic passes for the outer dims + fused 2-D kernels over the planes)
    bool hybrid_used = false;
    int last_rec_kernel = 0;   // synthesis tile kernel of the last 3-D/4-D fused level: 1 direct-load, 2 bulk (32-column tiles), 4 full rows
    int64_t launches = 0;
    // fused coefficient shrink (nddwt_plan_set_shrink): soft threshold applied to the detail bands as the
    // analysis kernels store them; thr[j-1][b] for level j, band b (b = 0, the approximation, is exempt)
    int shrink_mode = 0;
    double shrink_thr[NDDWT_MAX_LEVELS][1 << NDDWT_MAX_DIMS];
    int cur_level = 1;         // level index of the level call in flight (selects the threshold row)
    int cur_dil = 1;           // dilation of the level call in flight when it runs the fused kernels (taps stretched, see padded_tap)
    int rows_min_ctas = 118;   // full-row synthesis kernel needs at least this many CTAs (nddwt_plan_set_param)

    // device scratch owned by the plan (allocated on first use, reused across calls)
    void *approx[2] = {nullptr, nullptr};      // ping-pong intermediate approximation bands
    void *gen_scratch = nullptr;               // generic path temporaries, 2*(d-1) arrays
    void *fused_scratch = nullptr;             // fused 4-D path intermediates
    size_t fused_scratch_bytes = 0;
    void *host_x = nullptr;                    // device staging for the *_host entry points
    void *host_c = nullptr;
    size_t host_c_bytes = 0;
    cudaStream_t host_stream = nullptr;
    // level-streamed host transforms (2-D ... 4-D, level >= 2): host_c then holds TWO level buffers of 2^d - 1 detail
    // bands instead of the whole stack; the copies run on host_copy, ordered against the kernels by these events
    cudaStream_t host_copy = nullptr;
    cudaEvent_t host_ev_k[2] = {nullptr, nullptr};   // kernels of the level using level buffer i are done
    cudaEvent_t host_ev_c[2] = {nullptr, nullptr};   // copy out of / into level buffer i is done

    // optional per-kernel CUDA-event timing (nddwt_plan_profile): event pairs recorded on the launch
    // stream around every kernel, grouped by kind
    bool profiling = false;
    struct Timed { int kind; cudaEvent_t e0, e1; };
    std::vector<Timed> timed;
    std::vector<cudaEvent_t> event_pool;
};

namespace nddwt {

void set_error(const std::string &msg);
int cuda_fail(cudaError_t e, const char *what);

#define NDDWT_CUDA(call)                                                          \
    do {                                                                          \
        cudaError_t e__ = (call);                                                 \
        if (e__ != cudaSuccess) return nddwt::cuda_fail(e__, #call);              \
    } while (0)

// geometry of one analysis/synthesis level on a slab (or the whole array)
struct LevelIO {
    const void *halo_lo = nullptr;   // planes below the slab along the last dim (nullptr: periodic)
    const void *halo_hi = nullptr;   // planes above
};

// Mixed wavelets in the fused kernels: every dimension runs with the LONGEST tap length of the plan; a shorter
// filter is zero-padded symmetrically (m = (L - L_i)/2 zeros on each side), which leaves its phase L_i/2 and
// therefore the result unchanged: sum_k g[k] x[n - k + L_i/2] == sum_k' g'[k'] x[n - k' + L/2].
// A-trous levels (dilation s > 1) run the same kernels the same way: the taps are STRETCHED about their centre,
// g'[(k - L_i/2) s + L/2] = g[k] and zero elsewhere, so that sum_k' g'[k'] x[n - (k' - L/2)] ==
// sum_k g[k] x[n - (k - L_i/2) s] -- an effective filter of L_i s taps (Haar at dilations 1, 2, 4 = the 2-, 4- and 8-tap
// instantiations of the tile kernels; up to 20 taps in the 2-D kernels).
inline int plan_max_taps(const nddwt_plan *p)
{
    int m = 0;
    for (int i = 0; i < p->ndims; ++i) m = p->L[i] > m ? p->L[i] : m;
    return m * p->cur_dil;
}
inline double padded_tap(const double *taps, int Li, int L, int k, int s = 1)
{
    const int c = k - L / 2;                 // offset from the centre, in samples
    if (c % s != 0) return 0.0;
    const int kk = c / s + Li / 2;
    return (kk >= 0 && kk < Li) ? taps[kk] : 0.0;
}
// sets the dilation of the level in flight for the duration of a fused-path attempt
struct DilScope {
    nddwt_plan *p;
    DilScope(nddwt_plan *plan, int dil) : p(plan) { p->cur_dil = dil; }
    ~DilScope() { p->cur_dil = 1; }
};
inline bool plan_uniform_taps(const nddwt_plan *p)
{
    for (int i = 1; i < p->ndims; ++i)
        if (p->L[i] != p->L[0]) return false;
    return true;
}

// dim-3 sub-range of a 4-D level (zn == 0: everything).  The multi-GPU plan issues the parts of a level per
// z-chunk so that the halo planes of one chunk travel while the next chunk computes (nddwt_multi.cu).
struct ZRange {
    int z0 = 0, zn = 0;
};

// ---- generic separable kernels (nddwt_generic.cu) ----
int generic_dec_level(nddwt_plan *p, int dil, const void *a_in, const LevelIO &io,
                      void *const *out_bands, cudaStream_t s);
int generic_rec_level(nddwt_plan *p, int dil, const void *const *in_bands, void *a_out, cudaStream_t s);
int generic_rec_stage1(nddwt_plan *p, int dil, const void *const *in_bands, void *u_lo, void *u_hi,
                       cudaStream_t s);
int generic_rec_stage2(nddwt_plan *p, int dil, const void *u_lo, const void *u_hi, const LevelIO &io,
                       void *a_out, cudaStream_t s);

// ---- fused kernels (nddwt_fused.cu); return 1 if the case has no fused instantiation ----
int fused_dec_level(nddwt_plan *p, int dil, const void *a_in, const LevelIO &io,
                    void *const *out_bands, cudaStream_t s);
int fused_rec_level(nddwt_plan *p, int dil, const void *const *in_bands, void *a_out, cudaStream_t s);
int fused_rec_stage1(nddwt_plan *p, int dil, const void *const *in_bands, void *u_lo, void *u_hi, cudaStream_t s,
                     int part = 0, const ZRange &zr = ZRange());
int fused_dec_level_part(nddwt_plan *p, int dil, int part, const void *a_in, const LevelIO &io,
                         void *const *out_bands, cudaStream_t s, const ZRange &zr = ZRange());
int fused_rec_stage2(nddwt_plan *p, int dil, const void *u_lo, const void *u_hi, const LevelIO &io, void *a_out,
                     cudaStream_t s);

int fused2d_dec_level(nddwt_plan *p, int dil, const void *a_in, const LevelIO &io, void *const *out_bands,
                      cudaStream_t s);
int fused2d_rec_level(nddwt_plan *p, int dil, const void *const *in_bands, void *a_out, cudaStream_t s);
int fused2d_dec_planes(nddwt_plan *p, const void *a_in, void *const *out_bands, int64_t planes, int band0, cudaStream_t s);
int fused2d_rec_planes(nddwt_plan *p, const void *const *in_bands, void *a_out, int64_t planes, cudaStream_t s);
int fused1d_transform(nddwt_plan *p, bool rec, const void *in, void *out, int level, cudaStream_t s);
bool fused_is_separable(const nddwt_plan *p);
int fused_rec_stage2_scatter(nddwt_plan *p, int dil, const void *u_lo, const void *u_hi, void *out, void *over_lo,
                             void *over_hi, cudaStream_t s, const ZRange &zr = ZRange());
int accumulate_elems(nddwt_plan *p, void *dst, const void *src, int64_t nelem, cudaStream_t s);
// in-place soft threshold of one band (paths whose analysis kernels have no fused epilogue)
int shrink_band(nddwt_plan *p, void *band, int64_t nelem, double thr, cudaStream_t s);
// several "plane += planes" in one launch (multi-GPU synthesis exchange)
constexpr int ACC_MAXS = 8, ACC_MAXP = 32;   // kernel parameter block stays under 4 KB
struct AccItem { void *dst; const void *src[ACC_MAXS]; int ns; };
struct AccParams { AccItem item[ACC_MAXP]; int n; };
int accumulate_planes(nddwt_plan *p, const AccParams &prm, int64_t plane_elems, cudaStream_t s);
int ensure_scratch(nddwt_plan *p);

// kernel kinds for nddwt_plan_kernel_time
enum { KIND_DEC3 = 0, KIND_REC3 = 1, KIND_DEC_LAST = 2, KIND_REC_LAST = 3, KIND_GENERIC = 4, KIND_COMM = 5, KIND_COUNT = 6 };

// RAII event bracket: records (kind, e0, e1) around a launch when profiling is on
struct LaunchTimer {
    nddwt_plan *p;
    cudaStream_t s;
    cudaEvent_t e1 = nullptr;
    LaunchTimer(nddwt_plan *plan, int kind, cudaStream_t stream) : p(plan), s(stream)
    {
        if (!p->profiling) return;
        cudaEvent_t ev[2];
        for (int i = 0; i < 2; ++i) {
            if (!p->event_pool.empty()) { ev[i] = p->event_pool.back(); p->event_pool.pop_back(); }
            else cudaEventCreate(&ev[i]);
        }
        cudaEventRecord(ev[0], s);
        e1 = ev[1];
        p->timed.push_back({kind, ev[0], ev[1]});
    }
    ~LaunchTimer() { if (e1) cudaEventRecord(e1, s); }
};

}  // namespace nddwt
