// nddwt_capi.cu -- the extern "C" boundary declared in include/nddwt_b200.h.
// Host-side logic that replaces mexFunction's marshalling and the level loops of
// nd_dwt_dec / nd_dwt_rec (mex/nd_dwt_mex.c:8-153, mex/nddwt.c:189-292): slot arithmetic,
// intermediate approximation ping-pong, scale folding, argument checks.
#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>
#include <strings.h>
#include "nddwt_plan.h"
#include "nddwt_taps.h"

namespace nddwt {

static thread_local std::string g_err;

void set_error(const std::string &msg) { g_err = msg; }

int cuda_fail(cudaError_t e, const char *what)
{
    g_err = std::string("CUDA error: ") + cudaGetErrorString(e) + " in " + what;
    return e == cudaErrorMemoryAllocation ? NDDWT_ERR_NOMEM : NDDWT_ERR_CUDA;
}

static size_t elem_size(int dtype)
{
    switch (dtype) {
        case NDDWT_F32: return 4;
        case NDDWT_F64: return 8;
        case NDDWT_C64: return 8;
        case NDDWT_C128: return 16;
    }
    return 0;
}

int ensure_scratch(nddwt_plan *p)
{
    NDDWT_CUDA(cudaSetDevice(p->device));
    if (!p->gen_scratch && p->ndims > 1) {
        NDDWT_CUDA(cudaMalloc(&p->gen_scratch, (size_t)(2 * (p->ndims - 1)) * p->numel * p->esize));
    }
    return 0;
}

static int ensure_approx(nddwt_plan *p, int which)
{
    if (!p->approx[which]) NDDWT_CUDA(cudaMalloc(&p->approx[which], (size_t)p->numel * p->esize));
    return 0;
}

static int dec_level(nddwt_plan *p, int dil, const void *a_in, const LevelIO &io, void *const *out_bands,
                     cudaStream_t s)
{
    if (p->batch > 1 && (io.halo_lo || io.halo_hi)) { set_error("slabs of batched plans are not supported"); return NDDWT_ERR_ARG; }
    if (p->kernel_mode == 0 && p->batch == 1) {
        int rc = fused_dec_level(p, dil, a_in, io, out_bands, s);
        if (rc <= 0) { p->last_path = 1; return rc; }
        rc = fused2d_dec_level(p, dil, a_in, io, out_bands, s);
        if (rc <= 0) { p->last_path = 1; return rc; }
    }
    p->hybrid_used = false;
    int rc = generic_dec_level(p, dil, a_in, io, out_bands, s);
    p->last_path = p->hybrid_used ? 2 : 0;
    if (rc || !p->shrink_mode || p->hybrid_used) return rc;      // the hybrid path thresholds in the 2-D kernels' stores
    // no fused epilogue on this path: threshold the detail bands in place
    const int j = p->cur_level >= 1 && p->cur_level <= NDDWT_MAX_LEVELS ? p->cur_level : 1;
    for (int b = 1; b < (1 << p->ndims) && !rc; ++b) rc = shrink_band(p, out_bands[b], p->numel, p->shrink_thr[j - 1][b], s);
    return rc;
}

static int rec_level(nddwt_plan *p, int dil, const void *const *in_bands, void *a_out, cudaStream_t s)
{
    if (p->kernel_mode == 0 && p->batch == 1) {
        int rc = fused_rec_level(p, dil, in_bands, a_out, s);
        if (rc <= 0) { p->last_path = 1; return rc; }
        rc = fused2d_rec_level(p, dil, in_bands, a_out, s);
        if (rc <= 0) { p->last_path = 1; return rc; }
    }
    p->hybrid_used = false;
    const int rc = generic_rec_level(p, dil, in_bands, a_out, s);
    p->last_path = p->hybrid_used ? 2 : 0;
    return rc;
}

static int check_level(const nddwt_plan *p, int level)
{
    if (!p) { set_error("null plan"); return NDDWT_ERR_ARG; }
    if (level < 1 || level > NDDWT_MAX_LEVELS) { set_error("level must be in 1..16"); return NDDWT_ERR_ARG; }
    return 0;
}

}  // namespace nddwt

using namespace nddwt;

extern "C" {

const char *nddwt_last_error(void) { return g_err.c_str(); }
const char *nddwt_version(void) { return "nddwt_b200 0.1 (sm_100a)"; }

int nddwt_wave_filters(const char *wname, double *low_d, double *hi_d, int *len)
{
    if (!wname || !low_d || !hi_d || !len) { set_error("null argument"); return NDDWT_ERR_ARG; }
    int p = 0;
    if (strncasecmp(wname, "db", 2) == 0) {
        const char *q = wname + 2;
        if (*q) {
            char *end = nullptr;
            long v = strtol(q, &end, 10);
            if (end && *end == '\0' && v >= 1 && v <= NDDWT_NUM_WAVELETS) p = (int)v;
        }
    }
    if (!p) { set_error("Unknown Wavelet Name"); return NDDWT_ERR_WAVELET; }
    const int L = 2 * p;
    const double *h = NDDWT_DB_TAPS[p - 1];
    // low_d[k] = h[L-1-k]; hi_d[k] = (-1)^(k+1) h[k]   (wave_filters.m:164-172)
    for (int k = 0; k < L; ++k) {
        low_d[k] = h[L - 1 - k];
        hi_d[k] = (k & 1) ? h[k] : -h[k];
    }
    *len = L;
    return 0;
}

int64_t nddwt_num_bands(int ndims, int level)
{
    if (ndims < 1 || ndims > NDDWT_MAX_DIMS || level < 1) return 0;
    const int64_t nd = (int64_t)1 << ndims;
    return nd + (nd - 1) * (level - 1);
}

int nddwt_infer_level(int ndims, int64_t nb)
{
    if (ndims < 1 || ndims > NDDWT_MAX_DIMS) return 0;
    const int64_t nd = (int64_t)1 << ndims;
    if (nb < nd || (nb - nd) % (nd - 1) != 0) return 0;
    return (int)(1 + (nb - nd) / (nd - 1));
}

static int plan_create_impl(nddwt_plan **plan, int ndims, const int64_t *dims, int64_t global_last,
                            const char *const *wnames, int dtype, int pres_l2_norm, int device)
{
    if (!plan || !dims || !wnames) { set_error("null argument"); return NDDWT_ERR_ARG; }
    *plan = nullptr;
    if (ndims < 1 || ndims > NDDWT_MAX_DIMS) { set_error("ndims must be 1..4"); return NDDWT_ERR_ARG; }
    if (elem_size(dtype) == 0) { set_error("bad dtype"); return NDDWT_ERR_ARG; }
    nddwt_plan *p = new nddwt_plan();
    p->ndims = ndims;
    p->dtype = dtype;
    p->pres_l2 = pres_l2_norm ? 1 : 0;
    p->device = device;
    p->esize = elem_size(dtype);
    p->numel = 1;
    for (int i = 0; i < ndims; ++i) {
        if (dims[i] < 1) { delete p; set_error("sizes must be positive"); return NDDWT_ERR_ARG; }
        p->dims[i] = dims[i];
        p->numel *= dims[i];
        int len = 0;
        int rc = nddwt_wave_filters(wnames[i], p->lo[i], p->hi[i], &len);
        if (rc) { delete p; return rc; }
        p->L[i] = len;
        const int64_t extent = (i == ndims - 1 && global_last > 0) ? global_last : dims[i];
        if ((int64_t)len > extent) {
            delete p;
            char msg[128];
            snprintf(msg, sizeof msg, "Dimension %d of Data is shorter than the wavelet filter being used", i + 1);
            set_error(msg);
            return NDDWT_ERR_SHORT_DIM;
        }
    }
    const double sd = p->pres_l2 ? 1.0 / std::sqrt(2.0) : 1.0;   // analysis, per dim
    const double sr = p->pres_l2 ? 1.0 / std::sqrt(2.0) : 0.5;   // synthesis (scale * norm), per dim
    memset(&p->dec_d, 0, sizeof p->dec_d);
    memset(&p->rec_d, 0, sizeof p->rec_d);
    memset(&p->dec_f, 0, sizeof p->dec_f);
    memset(&p->rec_f, 0, sizeof p->rec_f);
    for (int i = 0; i < ndims; ++i)
        for (int k = 0; k < p->L[i]; ++k) {
            p->dec_d.d[i].lo[k] = sd * p->lo[i][k];
            p->dec_d.d[i].hi[k] = sd * p->hi[i][k];
            p->rec_d.d[i].lo[k] = sr * p->lo[i][k];
            p->rec_d.d[i].hi[k] = sr * p->hi[i][k];
            p->dec_f.d[i].lo[k] = (float)p->dec_d.d[i].lo[k];
            p->dec_f.d[i].hi[k] = (float)p->dec_d.d[i].hi[k];
            p->rec_f.d[i].lo[k] = (float)p->rec_d.d[i].lo[k];
            p->rec_f.d[i].hi[k] = (float)p->rec_d.d[i].hi[k];
        }
    for (int j = 0; j < NDDWT_MAX_LEVELS; ++j) p->dil[j] = 1;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || device < 0 || device >= ndev) {
        delete p;
        set_error(std::string("no usable CUDA device (this library has no CPU path): ") +
                  (e != cudaSuccess ? cudaGetErrorString(e) : "device ordinal out of range"));
        return NDDWT_ERR_CUDA;
    }
    *plan = p;
    return 0;
}

int nddwt_plan_create(nddwt_plan **plan, int ndims, const int64_t *dims, const char *const *wnames, int dtype,
                      int pres_l2_norm, int device)
{
    return plan_create_impl(plan, ndims, dims, 0, wnames, dtype, pres_l2_norm, device);
}

int nddwt_plan_create_slab(nddwt_plan **plan, int ndims, const int64_t *local_dims, int64_t global_last_dim,
                           const char *const *wnames, int dtype, int pres_l2_norm, int device)
{
    if (global_last_dim < 1 || !local_dims || ndims < 1 || ndims > NDDWT_MAX_DIMS ||
        local_dims[ndims - 1] > global_last_dim) {
        set_error("bad slab geometry");
        return NDDWT_ERR_ARG;
    }
    return plan_create_impl(plan, ndims, local_dims, global_last_dim, wnames, dtype, pres_l2_norm, device);
}

int nddwt_plan_destroy(nddwt_plan *p)
{
    if (!p) return 0;
    cudaSetDevice(p->device);
    for (int i = 0; i < 2; ++i) if (p->approx[i]) cudaFree(p->approx[i]);
    if (p->gen_scratch) cudaFree(p->gen_scratch);
    if (p->fused_scratch) cudaFree(p->fused_scratch);
    if (p->host_x) cudaFree(p->host_x);
    if (p->host_c) cudaFree(p->host_c);
    if (p->host_stream) cudaStreamDestroy(p->host_stream);
    if (p->host_copy) cudaStreamDestroy(p->host_copy);
    for (int i = 0; i < 2; ++i) {
        if (p->host_ev_k[i]) cudaEventDestroy(p->host_ev_k[i]);
        if (p->host_ev_c[i]) cudaEventDestroy(p->host_ev_c[i]);
    }
    for (size_t i = 0; i < p->timed.size(); ++i) { cudaEventDestroy(p->timed[i].e0); cudaEventDestroy(p->timed[i].e1); }
    for (size_t i = 0; i < p->event_pool.size(); ++i) cudaEventDestroy(p->event_pool[i]);
    delete p;
    return 0;
}

int nddwt_plan_set_dilations(nddwt_plan *p, const int *dil, int nlevels)
{
    if (!p || !dil || nlevels < 1 || nlevels > NDDWT_MAX_LEVELS) { set_error("bad dilation list"); return NDDWT_ERR_ARG; }
    for (int j = 0; j < nlevels; ++j) {
        if (dil[j] < 1) { set_error("dilation must be >= 1"); return NDDWT_ERR_ARG; }
        p->dil[j] = dil[j];
    }
    return 0;
}

int nddwt_plan_set_batch(nddwt_plan *p, int64_t batch)
{
    if (!p || batch < 1) { set_error("batch must be >= 1"); return NDDWT_ERR_ARG; }
    if (batch == p->batch) return 0;
    cudaSetDevice(p->device);
    // scratch is sized by numel: drop it, it is re-allocated on the next call
    for (int i = 0; i < 2; ++i) if (p->approx[i]) { cudaFree(p->approx[i]); p->approx[i] = nullptr; }
    if (p->gen_scratch) { cudaFree(p->gen_scratch); p->gen_scratch = nullptr; }
    if (p->fused_scratch) { cudaFree(p->fused_scratch); p->fused_scratch = nullptr; p->fused_scratch_bytes = 0; }
    if (p->host_x) { cudaFree(p->host_x); p->host_x = nullptr; }
    if (p->host_c) { cudaFree(p->host_c); p->host_c = nullptr; p->host_c_bytes = 0; }
    p->numel = p->numel / p->batch * batch;
    p->batch = batch;
    return 0;
}

int nddwt_plan_set_shrink(nddwt_plan *p, int mode, const double *thr, int nlevels)
{
    if (!p || mode < 0 || mode > 1) { set_error("shrink mode must be 0 (off) or 1 (soft threshold)"); return NDDWT_ERR_ARG; }
    const int nd = 1 << p->ndims;
    memset(p->shrink_thr, 0, sizeof p->shrink_thr);
    p->shrink_mode = 0;
    if (mode == 0) return 0;
    if (!thr || nlevels < 1 || nlevels > NDDWT_MAX_LEVELS) { set_error("bad threshold table"); return NDDWT_ERR_ARG; }
    for (int j = 0; j < nlevels; ++j)
        for (int b = 1; b < nd; ++b) {      // b = 0 (approximation) is exempt
            const double t = thr[(size_t)j * nd + b];
            if (!(t >= 0.0)) { set_error("thresholds must be >= 0"); return NDDWT_ERR_ARG; }
            p->shrink_thr[j][b] = t;
        }
    p->shrink_mode = 1;
    return 0;
}

int nddwt_plan_set_kernel_mode(nddwt_plan *p, int mode)
{
    if (!p || mode < 0 || mode > 1) { set_error("bad kernel mode"); return NDDWT_ERR_ARG; }
    p->kernel_mode = mode;
    return 0;
}

int nddwt_plan_set_param(nddwt_plan *p, const char *name, int64_t value)
{
    if (!p || !name) { set_error("null argument"); return NDDWT_ERR_ARG; }
    if (strcmp(name, "rows_min_ctas") == 0) {
        if (value < 0 || value > (1 << 30)) { set_error("rows_min_ctas out of range"); return NDDWT_ERR_ARG; }
        p->rows_min_ctas = (int)value;
        return 0;
    }
    set_error(std::string("unknown plan parameter: ") + name);
    return NDDWT_ERR_ARG;
}

int64_t nddwt_plan_launch_count(const nddwt_plan *p) { return p ? p->launches : 0; }

int nddwt_plan_profile(nddwt_plan *p, int on)
{
    if (!p) { set_error("null plan"); return NDDWT_ERR_ARG; }
    p->profiling = on != 0;
    if (on) {   // a new profiling session starts empty: records nobody read go back to the pool
        for (size_t i = 0; i < p->timed.size(); ++i) { p->event_pool.push_back(p->timed[i].e0); p->event_pool.push_back(p->timed[i].e1); }
        p->timed.clear();
    }
    return 0;
}

int nddwt_plan_kernel_time(nddwt_plan *p, int kind, double *total_ms, int64_t *count)
{
    if (!p || !total_ms || !count || kind < 0 || kind >= KIND_COUNT) { set_error("bad argument"); return NDDWT_ERR_ARG; }
    NDDWT_CUDA(cudaSetDevice(p->device));
    double tot = 0.0;
    int64_t n = 0;
    // entries of this kind leave `timed` whatever happens (a failed read must not leave an event both in
    // `timed` and in the pool); their events go back to the pool
    std::vector<nddwt_plan::Timed> keep, mine;
    for (size_t i = 0; i < p->timed.size(); ++i) (p->timed[i].kind == kind ? mine : keep).push_back(p->timed[i]);
    p->timed.swap(keep);
    cudaError_t err = cudaSuccess;
    for (size_t i = 0; i < mine.size(); ++i) {
        float ms = 0.f;
        if (err == cudaSuccess) err = cudaEventSynchronize(mine[i].e1);
        if (err == cudaSuccess) err = cudaEventElapsedTime(&ms, mine[i].e0, mine[i].e1);
        if (err == cudaSuccess) { tot += ms; ++n; }
        p->event_pool.push_back(mine[i].e0);
        p->event_pool.push_back(mine[i].e1);
    }
    if (err != cudaSuccess) return cuda_fail(err, "reading kernel timing events");
    *total_ms = tot;
    *count = n;
    return 0;
}
int nddwt_plan_last_path(const nddwt_plan *p) { return p ? p->last_path : 0; }
int nddwt_plan_last_synthesis_kernel(const nddwt_plan *p) { return p ? p->last_rec_kernel : 0; }

int nddwt_dec(nddwt_plan *p, const void *x_dev, void *coeffs_dev, int level, void *stream)
{
    int rc = check_level(p, level);
    if (rc) return rc;
    if (!x_dev || !coeffs_dev) { set_error("null device pointer"); return NDDWT_ERR_ARG; }
    NDDWT_CUDA(cudaSetDevice(p->device));
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    if (p->ndims == 1) {   // 1-D: the whole multi-level cascade is one kernel
        rc = fused1d_transform(p, false, x_dev, coeffs_dev, level, s);
        if (rc <= 0) { p->last_path = 1; return rc; }
    }
    const int nd = 1 << p->ndims;
    char *c = reinterpret_cast<char *>(coeffs_dev);
    const size_t band_bytes = (size_t)p->numel * p->esize;
    const void *a_in = x_dev;
    LevelIO io;
    for (int j = 1; j <= level; ++j) {
        // bands of level j live in slots (nd-1)(level-j) .. +nd-1 (mex/nddwt.c:209-210,226); the
        // approximation of a non-final level goes to plan scratch instead of a slot that the
        // next level overwrites (the reference copies it out, nddwt.c:216-219).
        const int64_t start = (int64_t)(nd - 1) * (level - j);
        void *bands[1 << NDDWT_MAX_DIMS];
        for (int b = 1; b < nd; ++b) bands[b] = c + (size_t)(start + b) * band_bytes;
        if (j == level) {
            bands[0] = c;
        } else {
            rc = ensure_approx(p, j & 1);
            if (rc) return rc;
            bands[0] = p->approx[j & 1];
        }
        p->cur_level = j;
        rc = dec_level(p, p->dil[j - 1], a_in, io, bands, s);
        if (rc) return rc;
        a_in = bands[0];
    }
    return 0;
}

int nddwt_shrink(nddwt_plan *p, void *coeffs_dev, int level, void *stream)
{
    int rc = check_level(p, level);
    if (rc) return rc;
    if (!coeffs_dev) { set_error("null device pointer"); return NDDWT_ERR_ARG; }
    if (!p->shrink_mode) { set_error("no threshold table: call nddwt_plan_set_shrink first"); return NDDWT_ERR_ARG; }
    NDDWT_CUDA(cudaSetDevice(p->device));
    const int nd = 1 << p->ndims;
    char *c = reinterpret_cast<char *>(coeffs_dev);
    const size_t band_bytes = (size_t)p->numel * p->esize;
    for (int j = 1; j <= level && !rc; ++j) {
        const int64_t start = (int64_t)(nd - 1) * (level - j);
        for (int b = 1; b < nd && !rc; ++b)
            rc = shrink_band(p, c + (size_t)(start + b) * band_bytes, p->numel, p->shrink_thr[j - 1][b],
                             reinterpret_cast<cudaStream_t>(stream));
    }
    return rc;
}

int nddwt_rec(nddwt_plan *p, const void *coeffs_dev, void *x_dev, int level, void *stream)
{
    int rc = check_level(p, level);
    if (rc) return rc;
    if (!x_dev || !coeffs_dev) { set_error("null device pointer"); return NDDWT_ERR_ARG; }
    NDDWT_CUDA(cudaSetDevice(p->device));
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    if (p->ndims == 1) {
        rc = fused1d_transform(p, true, coeffs_dev, x_dev, level, s);
        if (rc <= 0) { p->last_path = 1; return rc; }
    }
    const int nd = 1 << p->ndims;
    const char *c = reinterpret_cast<const char *>(coeffs_dev);
    const size_t band_bytes = (size_t)p->numel * p->esize;
    const void *a = c;   // slot 0 = deepest approximation
    for (int j = level; j >= 1; --j) {
        const int64_t start = (int64_t)(nd - 1) * (level - j);
        const void *bands[1 << NDDWT_MAX_DIMS];
        bands[0] = a;
        for (int b = 1; b < nd; ++b) bands[b] = c + (size_t)(start + b) * band_bytes;
        void *out;
        if (j == 1) {
            out = x_dev;
        } else {
            rc = ensure_approx(p, j & 1);
            if (rc) return rc;
            out = p->approx[j & 1];
        }
        rc = rec_level(p, p->dil[j - 1], bands, out, s);
        if (rc) return rc;
        a = out;
    }
    return 0;
}

// Host-pointer entry points.  `streamed`: the device holds x, the approximation ping-pong and TWO level buffers of
// 2^d - 1 detail bands; whole stack otherwise (1-D: the cascade kernel wants the stack; one level: nothing to stream).
static bool host_streamed(const nddwt_plan *p, int level) { return p->ndims >= 2 && level >= 2; }

static int ensure_host_staging(nddwt_plan *p, int level)
{
    NDDWT_CUDA(cudaSetDevice(p->device));
    const size_t band_bytes = (size_t)p->numel * p->esize;
    const size_t nd = (size_t)1 << p->ndims;
    const size_t need = host_streamed(p, level) ? 2 * (nd - 1) * band_bytes
                                                : band_bytes * (size_t)nddwt_num_bands(p->ndims, level);
    if (!p->host_stream) NDDWT_CUDA(cudaStreamCreateWithFlags(&p->host_stream, cudaStreamNonBlocking));
    if (!p->host_copy) NDDWT_CUDA(cudaStreamCreateWithFlags(&p->host_copy, cudaStreamNonBlocking));
    for (int i = 0; i < 2; ++i) {
        if (!p->host_ev_k[i]) NDDWT_CUDA(cudaEventCreateWithFlags(&p->host_ev_k[i], cudaEventDisableTiming));
        if (!p->host_ev_c[i]) NDDWT_CUDA(cudaEventCreateWithFlags(&p->host_ev_c[i], cudaEventDisableTiming));
    }
    if (!p->host_x) NDDWT_CUDA(cudaMalloc(&p->host_x, band_bytes));
    if (p->host_c_bytes < need) {
        if (p->host_c) { cudaFree(p->host_c); p->host_c = nullptr; p->host_c_bytes = 0; }
        NDDWT_CUDA(cudaMalloc(&p->host_c, need));
        p->host_c_bytes = need;
    }
    return 0;
}

// Level-streamed analysis: level j writes its detail bands into level buffer j & 1; they leave for their (contiguous)
// slots of the host stack on the copy stream while level j + 1 computes into the other buffer.  Device memory:
// (1 + 2 (2^d - 1) + 2) N e instead of (1 + nb + 2) N e -- BASELINE configs[3] (46 bands of 4.3 GB) fits one 180 GB
// GPU this way (35 N e = 150 GB), which the whole-stack form (197.6 GB) does not.
static int dec_host_streamed_issue(nddwt_plan *p, const void *x_host, void *coeffs_host, int level)
{
    const size_t band_bytes = (size_t)p->numel * p->esize;
    const int nd = 1 << p->ndims;
    const size_t lvl_bytes = (size_t)(nd - 1) * band_bytes;
    cudaStream_t ks = p->host_stream, cs = p->host_copy;
    char *hc = reinterpret_cast<char *>(coeffs_host);
    NDDWT_CUDA(cudaMemcpyAsync(p->host_x, x_host, band_bytes, cudaMemcpyHostToDevice, ks));
    const void *a_in = p->host_x;
    LevelIO io;
    for (int j = 1; j <= level; ++j) {
        const int w = j & 1;
        char *lvl = reinterpret_cast<char *>(p->host_c) + (size_t)w * lvl_bytes;
        int rc = ensure_approx(p, w);
        if (rc) return rc;
        void *bands[1 << NDDWT_MAX_DIMS];
        bands[0] = p->approx[w];
        for (int b = 1; b < nd; ++b) bands[b] = lvl + (size_t)(b - 1) * band_bytes;
        if (j > 2) NDDWT_CUDA(cudaStreamWaitEvent(ks, p->host_ev_c[w], 0));   // level j - 2 has left this buffer
        p->cur_level = j;
        rc = dec_level(p, p->dil[j - 1], a_in, io, bands, ks);
        if (rc) return rc;
        NDDWT_CUDA(cudaEventRecord(p->host_ev_k[w], ks));
        NDDWT_CUDA(cudaStreamWaitEvent(cs, p->host_ev_k[w], 0));
        // bands of level j live in slots (nd-1)(level-j)+1 .. +nd-1 (mex/nddwt.c:209-210,226): one contiguous block
        const size_t start = (size_t)(nd - 1) * (size_t)(level - j);
        NDDWT_CUDA(cudaMemcpyAsync(hc + (start + 1) * band_bytes, lvl, lvl_bytes, cudaMemcpyDeviceToHost, cs));
        if (j == level) NDDWT_CUDA(cudaMemcpyAsync(hc, p->approx[w], band_bytes, cudaMemcpyDeviceToHost, cs));
        NDDWT_CUDA(cudaEventRecord(p->host_ev_c[w], cs));
        a_in = bands[0];
    }
    return 0;
}

// both streams are drained whatever happened: no copy may still be writing into the caller's arrays after the return
static int host_streams_drain(nddwt_plan *p, int rc)
{
    const cudaError_t e1 = cudaStreamSynchronize(p->host_copy), e2 = cudaStreamSynchronize(p->host_stream);
    if (rc) return rc;
    if (e1 != cudaSuccess) return cuda_fail(e1, "cudaStreamSynchronize(copy stream)");
    if (e2 != cudaSuccess) return cuda_fail(e2, "cudaStreamSynchronize(kernel stream)");
    return 0;
}

static int dec_host_streamed(nddwt_plan *p, const void *x_host, void *coeffs_host, int level)
{
    return host_streams_drain(p, dec_host_streamed_issue(p, x_host, coeffs_host, level));
}

// Level-streamed synthesis: the detail bands of level j - 1 arrive in the other level buffer while level j computes.
static int rec_host_streamed_issue(nddwt_plan *p, const void *coeffs_host, void *x_host, int level)
{
    const size_t band_bytes = (size_t)p->numel * p->esize;
    const int nd = 1 << p->ndims;
    const size_t lvl_bytes = (size_t)(nd - 1) * band_bytes;
    cudaStream_t ks = p->host_stream, cs = p->host_copy;
    const char *hc = reinterpret_cast<const char *>(coeffs_host);
    int rc = ensure_approx(p, 0);
    if (!rc) rc = ensure_approx(p, 1);
    if (rc) return rc;
    // a_J goes where level J + 1 would have written it
    void *a = p->approx[(level + 1) & 1];
    NDDWT_CUDA(cudaMemcpyAsync(a, hc, band_bytes, cudaMemcpyHostToDevice, cs));
    for (int j = level; j >= 1; --j) {
        const int w = j & 1;
        char *lvl = reinterpret_cast<char *>(p->host_c) + (size_t)w * lvl_bytes;
        if (j == level) {
            const size_t start = (size_t)(nd - 1) * (size_t)(level - j);
            NDDWT_CUDA(cudaMemcpyAsync(lvl, hc + (start + 1) * band_bytes, lvl_bytes, cudaMemcpyHostToDevice, cs));
            NDDWT_CUDA(cudaEventRecord(p->host_ev_c[w], cs));
        }
        if (j > 1) {   // prefetch level j - 1 into the other buffer (level j + 1 read it last)
            const int wn = (j - 1) & 1;
            char *nxt = reinterpret_cast<char *>(p->host_c) + (size_t)wn * lvl_bytes;
            if (j < level) NDDWT_CUDA(cudaStreamWaitEvent(cs, p->host_ev_k[wn], 0));
            const size_t start = (size_t)(nd - 1) * (size_t)(level - (j - 1));
            NDDWT_CUDA(cudaMemcpyAsync(nxt, hc + (start + 1) * band_bytes, lvl_bytes, cudaMemcpyHostToDevice, cs));
            NDDWT_CUDA(cudaEventRecord(p->host_ev_c[wn], cs));
        }
        // level j: its details (and, for j == level, a_J: same stream, earlier) have landed -- host_ev_c[w] was
        // recorded one iteration earlier, or just above for j == level
        NDDWT_CUDA(cudaStreamWaitEvent(ks, p->host_ev_c[w], 0));
        const void *bands[1 << NDDWT_MAX_DIMS];
        bands[0] = a;
        for (int b = 1; b < nd; ++b) bands[b] = lvl + (size_t)(b - 1) * band_bytes;
        void *out = (j == 1) ? p->host_x : p->approx[w];
        p->cur_level = j;
        rc = rec_level(p, p->dil[j - 1], bands, out, ks);
        if (rc) return rc;
        NDDWT_CUDA(cudaEventRecord(p->host_ev_k[w], ks));
        a = out;
    }
    NDDWT_CUDA(cudaMemcpyAsync(x_host, p->host_x, band_bytes, cudaMemcpyDeviceToHost, ks));
    return 0;
}

static int rec_host_streamed(nddwt_plan *p, const void *coeffs_host, void *x_host, int level)
{
    return host_streams_drain(p, rec_host_streamed_issue(p, coeffs_host, x_host, level));
}

int nddwt_dec_host(nddwt_plan *p, const void *x_host, void *coeffs_host, int level)
{
    int rc = check_level(p, level);
    if (rc) return rc;
    if (!x_host || !coeffs_host) { set_error("null host pointer"); return NDDWT_ERR_ARG; }
    rc = ensure_host_staging(p, level);
    if (rc) return rc;
    if (host_streamed(p, level)) return dec_host_streamed(p, x_host, coeffs_host, level);
    const size_t band_bytes = (size_t)p->numel * p->esize;
    const size_t nb = (size_t)nddwt_num_bands(p->ndims, level);
    NDDWT_CUDA(cudaMemcpyAsync(p->host_x, x_host, band_bytes, cudaMemcpyHostToDevice, p->host_stream));
    rc = nddwt_dec(p, p->host_x, p->host_c, level, p->host_stream);
    if (rc) return rc;
    NDDWT_CUDA(cudaMemcpyAsync(coeffs_host, p->host_c, band_bytes * nb, cudaMemcpyDeviceToHost, p->host_stream));
    NDDWT_CUDA(cudaStreamSynchronize(p->host_stream));
    return 0;
}

int nddwt_rec_host(nddwt_plan *p, const void *coeffs_host, void *x_host, int level)
{
    int rc = check_level(p, level);
    if (rc) return rc;
    if (!x_host || !coeffs_host) { set_error("null host pointer"); return NDDWT_ERR_ARG; }
    rc = ensure_host_staging(p, level);
    if (rc) return rc;
    if (host_streamed(p, level)) return rec_host_streamed(p, coeffs_host, x_host, level);
    const size_t band_bytes = (size_t)p->numel * p->esize;
    const size_t nb = (size_t)nddwt_num_bands(p->ndims, level);
    NDDWT_CUDA(cudaMemcpyAsync(p->host_c, coeffs_host, band_bytes * nb, cudaMemcpyHostToDevice, p->host_stream));
    rc = nddwt_rec(p, p->host_c, p->host_x, level, p->host_stream);
    if (rc) return rc;
    NDDWT_CUDA(cudaMemcpyAsync(x_host, p->host_x, band_bytes, cudaMemcpyDeviceToHost, p->host_stream));
    NDDWT_CUDA(cudaStreamSynchronize(p->host_stream));
    return 0;
}

// ------------------------------------------------------------------------------------------
// slab interface
// ------------------------------------------------------------------------------------------
int nddwt_halo_planes(const nddwt_plan *p, int level_index, int *below, int *above)
{
    if (!p || level_index < 1 || level_index > NDDWT_MAX_LEVELS || !below || !above) {
        set_error("bad argument");
        return NDDWT_ERR_ARG;
    }
    const int L = p->L[p->ndims - 1], dil = p->dil[level_index - 1];
    *below = (L / 2 - 1) * dil;
    *above = (L / 2) * dil;
    return 0;
}

int nddwt_dec_level_slab(nddwt_plan *p, int level_index, const void *a_in, const void *halo_lo,
                         const void *halo_hi, void *const *out_bands, void *stream)
{
    int rc = check_level(p, level_index);
    if (rc) return rc;
    if (!a_in || !out_bands) { set_error("null pointer"); return NDDWT_ERR_ARG; }
    NDDWT_CUDA(cudaSetDevice(p->device));
    LevelIO io;
    io.halo_lo = halo_lo;
    io.halo_hi = halo_hi;
    p->cur_level = level_index;
    return dec_level(p, p->dil[level_index - 1], a_in, io, out_bands, reinterpret_cast<cudaStream_t>(stream));
}

int nddwt_rec_level_slab_stage2_scatter(nddwt_plan *p, int level_index, const void *u_lo, const void *u_hi,
                                        void *a_out, void *over_lo, void *over_hi, void *stream)
{
    int rc = check_level(p, level_index);
    if (rc) return rc;
    if (!u_lo || !u_hi || !a_out || !over_lo || !over_hi) { set_error("null pointer"); return NDDWT_ERR_ARG; }
    NDDWT_CUDA(cudaSetDevice(p->device));
    rc = fused_rec_stage2_scatter(p, p->dil[level_index - 1], u_lo, u_hi, a_out, over_lo, over_hi,
                                  reinterpret_cast<cudaStream_t>(stream));
    if (rc > 0) { set_error("scatter-form synthesis needs a separable (fused 4-D) plan"); return NDDWT_ERR_ARG; }
    return rc;
}

int nddwt_accumulate(nddwt_plan *p, void *dst, const void *src, int64_t nelem, void *stream)
{
    if (!p || !dst || !src || nelem < 0) { set_error("bad argument"); return NDDWT_ERR_ARG; }
    NDDWT_CUDA(cudaSetDevice(p->device));
    return accumulate_elems(p, dst, src, nelem, reinterpret_cast<cudaStream_t>(stream));
}

int nddwt_plan_is_separable(const nddwt_plan *p) { return (p && fused_is_separable(p)) ? 1 : 0; }

int nddwt_dec_level_slab_part(nddwt_plan *p, int level_index, int part, const void *a_in, const void *halo_lo,
                              const void *halo_hi, void *const *out_bands, void *stream)
{
    int rc = check_level(p, level_index);
    if (rc) return rc;
    if (!a_in || !out_bands || part < 0 || part > 3) { set_error("bad argument"); return NDDWT_ERR_ARG; }
    NDDWT_CUDA(cudaSetDevice(p->device));
    LevelIO io;
    io.halo_lo = halo_lo;
    io.halo_hi = halo_hi;
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    p->cur_level = level_index;
    if (p->kernel_mode == 0) {
        rc = fused_dec_level_part(p, p->dil[level_index - 1], part, a_in, io, out_bands, s);
        if (rc <= 0) { p->last_path = 1; return rc; }
    }
    // no separable parts on this plan: the whole level runs as part 0/1, parts 2 and 3 are no-ops
    if (part >= 2) return 0;
    return dec_level(p, p->dil[level_index - 1], a_in, io, out_bands, s);
}

int nddwt_rec_level_slab_stage1_part(nddwt_plan *p, int level_index, int part, const void *const *in_bands,
                                     void *u_lo, void *u_hi, void *stream)
{
    int rc = check_level(p, level_index);
    if (rc) return rc;
    if (!in_bands || !u_lo || !u_hi || part < 0 || part > 2) { set_error("bad argument"); return NDDWT_ERR_ARG; }
    NDDWT_CUDA(cudaSetDevice(p->device));
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    if (p->kernel_mode == 0) {
        rc = fused_rec_stage1(p, p->dil[level_index - 1], in_bands, u_lo, u_hi, s, part);
        if (rc <= 0) { p->last_path = 1; return rc; }
    }
    if (part == 2) return 0;   // not separable: everything is done by part 0/1
    p->last_path = 0;
    return generic_rec_stage1(p, p->dil[level_index - 1], in_bands, u_lo, u_hi, s);
}

int nddwt_rec_level_slab_stage1(nddwt_plan *p, int level_index, const void *const *in_bands, void *u_lo,
                                void *u_hi, void *stream)
{
    int rc = check_level(p, level_index);
    if (rc) return rc;
    if (!in_bands || !u_lo || !u_hi) { set_error("null pointer"); return NDDWT_ERR_ARG; }
    NDDWT_CUDA(cudaSetDevice(p->device));
    if (p->kernel_mode == 0) {
        rc = fused_rec_stage1(p, p->dil[level_index - 1], in_bands, u_lo, u_hi, reinterpret_cast<cudaStream_t>(stream));
        if (rc <= 0) { p->last_path = 1; return rc; }
    }
    p->last_path = 0;
    return generic_rec_stage1(p, p->dil[level_index - 1], in_bands, u_lo, u_hi,
                              reinterpret_cast<cudaStream_t>(stream));
}

int nddwt_rec_level_slab_stage2(nddwt_plan *p, int level_index, const void *u_lo, const void *u_hi,
                                const void *halo_lo, const void *halo_hi, void *a_out, void *stream)
{
    int rc = check_level(p, level_index);
    if (rc) return rc;
    if (!u_lo || !u_hi || !a_out) { set_error("null pointer"); return NDDWT_ERR_ARG; }
    NDDWT_CUDA(cudaSetDevice(p->device));
    LevelIO io;
    io.halo_lo = halo_lo;
    io.halo_hi = halo_hi;
    if (p->kernel_mode == 0) {
        rc = fused_rec_stage2(p, p->dil[level_index - 1], u_lo, u_hi, io, a_out, reinterpret_cast<cudaStream_t>(stream));
        if (rc <= 0) { p->last_path = 1; return rc; }
    }
    p->last_path = 0;
    return generic_rec_stage2(p, p->dil[level_index - 1], u_lo, u_hi, io, a_out,
                              reinterpret_cast<cudaStream_t>(stream));
}

}  // extern "C"
