// nd_dwt_mex.cpp -- MEX gateway over libnddwt_b200.so.  Drop-in for mex/nd_dwt_mex.c of
// arg-min-x/Non-Decimated_Wavelets: same function name, same five-argument call
//
//     y = nd_dwt_mex(x, f, dir, level, pres_l2_norm)
//
// with these differences (the spatial kernels need no Fourier-domain filters):
//   * x  is the SPATIAL array (dir == 0) or the coefficient stack [sizes, nb] (dir ~= 0), single or double,
//        real or complex (the reference accepts only complex double, nd_dwt_mex.c:23-30, and wants fftn(x)),
//        a host array OR A gpuArray: a gpuArray stays on the device -- its device pointer goes straight to
//        nddwt_dec / nddwt_rec and the result is created as a gpuArray (no gather / re-upload, which is what
//        the reference's 'gpu_off' mode does on every call, nd_dwt_1D.m:139-141,192-194);
//   * f  is a PLAN HANDLE (uint64 scalar from nd_dwt_mex('plan', ...), stored in the object by the nd_dwt_*D
//        classes of this directory: the "stored filters" of the reference, a few bytes instead of 2^d*numel
//        complex values), or -- compatibility -- the descriptor struct (wname, sizes), looked up by key;
//   * inputs are never modified (the reference overwrites prhs[0], nddwt.c:163,264-265).
// Other entries:
//     [lo, hi] = nd_dwt_mex('taps', wname)                                   wave_filters.m
//     h = nd_dwt_mex('plan', f, is_single, is_complex, pres_l2_norm [, ngpus])   uint64 handle; ngpus > 1: the
//                                                                            multi-GPU plan (host arrays only)
//     nd_dwt_mex('shrink', h, table)      soft-threshold table [J x 2^d] fused into later dec calls ([] = off)
//     nd_dwt_mex('dilations', h, dil)     a-trous mode: dilation of the taps per level (default 1 everywhere = reference)
//     nd_dwt_mex('release', h) / nd_dwt_mex('release')
// Build:  mex -R2018a nd_dwt_mex.cpp -I../../include -L.. -lnddwt_b200                  (host arrays)
//         mexcuda -R2018a -DNDDWT_MEX_GPU nd_dwt_mex.cpp -I../../include -L.. -lnddwt_b200   (+ gpuArray)
//         add -DNDDWT_MEX_SPLIT_COMPLEX and drop -R2018a for the legacy split-complex API (mxGetPr / mxGetPi,
//         what the reference gateway uses, nd_dwt_mex.c:55-58): complex host data is then interleaved on the way.
// No C++ object with a destructor is alive when an error is raised (mexErrMsgIdAndTxt does not return).
#include <stdio.h>
#include <string.h>
#include "mex.h"
#ifdef NDDWT_MEX_GPU
#include "gpu/mxGPUArray.h"
#endif
#include "../../include/nddwt_b200.h"

namespace {

struct Entry {
    bool used;
    nddwt_plan *plan;        // single-GPU plan
    nddwt_mplan *mplan;      // multi-GPU plan (ngpus > 1)
    int d, dtype, l2, ngpus;
    int64_t dims[NDDWT_MAX_DIMS];
    char key[160];           // descriptor-struct compatibility path
};
const int MAX_ENTRIES = 256;
Entry g_tab[MAX_ENTRIES];
bool g_exit_registered = false;
char g_msg[512];

void release_all(void)
{
    for (int i = 0; i < MAX_ENTRIES; ++i)
        if (g_tab[i].used) {
            if (g_tab[i].plan) nddwt_plan_destroy(g_tab[i].plan);
            if (g_tab[i].mplan) nddwt_mplan_destroy(g_tab[i].mplan);
            g_tab[i].used = false;
        }
}

void fail(const char *msg) { mexErrMsgIdAndTxt("MATLAB:FFT2mx:invalidNumInputs", "%s", msg); }
void fail_lib(void)
{
    snprintf(g_msg, sizeof g_msg, "%s", nddwt_last_error());
    fail(g_msg);
}

struct Desc {
    int d;
    int64_t dims[NDDWT_MAX_DIMS];
    char names[NDDWT_MAX_DIMS][16];
    const char *cnames[NDDWT_MAX_DIMS];
};

// descriptor struct (wname: cell of d strings, sizes: 1 x d double) -> Desc; returns an error text or NULL
const char *parse_desc(const mxArray *f, Desc *out)
{
    if (!mxIsStruct(f)) return "FIlter size and image size not consistant";
    const mxArray *fw = mxGetField(f, 0, "wname"), *fs = mxGetField(f, 0, "sizes");
    if (!fw || !fs || !mxIsCell(fw) || !mxIsDouble(fs)) return "FIlter size and image size not consistant";
    const int d = (int)mxGetNumberOfElements(fs);
    if (d < 1 || d > NDDWT_MAX_DIMS || (int)mxGetNumberOfElements(fw) != d) return "FIlter size and image size not consistant";
    out->d = d;
    const double *sz = mxGetDoubles(fs);
    for (int i = 0; i < d; ++i) {
        out->dims[i] = (int64_t)sz[i];
        char *s = mxArrayToString(mxGetCell(fw, i));
        if (!s) return "Unknown Wavelet Name";
        snprintf(out->names[i], sizeof out->names[i], "%s", s);
        mxFree(s);
        out->cnames[i] = out->names[i];
    }
    return NULL;
}

int new_entry(const Desc &ds, int dtype, int l2, int ngpus, const char *key)
{
    int slot = -1;
    for (int i = 0; i < MAX_ENTRIES && slot < 0; ++i)
        if (!g_tab[i].used) slot = i;
    if (slot < 0) fail("too many nd_dwt plans alive: nd_dwt_mex('release')");
    Entry &e = g_tab[slot];
    memset(&e, 0, sizeof e);
    int rc;
    if (ngpus > 1) rc = nddwt_mplan_create(&e.mplan, ds.d, ds.dims, ds.cnames, dtype, l2, ngpus, NULL);
    else rc = nddwt_plan_create(&e.plan, ds.d, ds.dims, ds.cnames, dtype, l2, 0);
    if (rc) fail_lib();
    e.used = true;
    e.d = ds.d;
    e.dtype = dtype;
    e.l2 = l2;
    e.ngpus = ngpus;
    for (int i = 0; i < ds.d; ++i) e.dims[i] = ds.dims[i];
    snprintf(e.key, sizeof e.key, "%s", key ? key : "");
    if (!g_exit_registered) { mexAtExit(release_all); g_exit_registered = true; }
    return slot;
}

Entry *entry_of_handle(const mxArray *h)
{
    if (!mxIsUint64(h) || mxGetNumberOfElements(h) != 1) return NULL;
    const uint64_t v = *reinterpret_cast<const uint64_t *>(mxGetData(h));
    if (v < 1 || v > (uint64_t)MAX_ENTRIES || !g_tab[v - 1].used) return NULL;
    return &g_tab[v - 1];
}

int dtype_code(bool is_single, bool is_complex)
{
    if (is_single) return is_complex ? NDDWT_C64 : NDDWT_F32;
    return is_complex ? NDDWT_C128 : NDDWT_F64;
}

// compatibility: descriptor struct in the f position -> plan cached under a key
Entry *entry_of_struct(const mxArray *f, int dtype, int l2)
{
    Desc ds;
    const char *err = parse_desc(f, &ds);
    if (err) fail(err);
    char key[160];
    int n = snprintf(key, sizeof key, "%d:%d", dtype, l2);
    for (int i = 0; i < ds.d && n < (int)sizeof key; ++i)
        n += snprintf(key + n, sizeof key - n, ":%s/%lld", ds.names[i], (long long)ds.dims[i]);
    for (int i = 0; i < MAX_ENTRIES; ++i)
        if (g_tab[i].used && g_tab[i].key[0] && strcmp(g_tab[i].key, key) == 0) return &g_tab[i];
    return &g_tab[new_entry(ds, dtype, l2, 1, key)];
}

void out_dims(const Entry &e, bool forward, int64_t nb, mwSize *odims, mwSize *nd_out)
{
    for (int i = 0; i < e.d; ++i) odims[i] = (mwSize)e.dims[i];
    if (forward) {                                   // [sizes, nb]                       nd_dwt_mex.c:79-88
        odims[e.d] = (mwSize)nb;
        *nd_out = (mwSize)e.d + 1;
    } else {                                         // sizes (1-D: column vector)        nd_dwt_mex.c:136-138
        *nd_out = (mwSize)e.d;
        if (e.d == 1) { odims[1] = 1; *nd_out = 2; }
    }
}

}  // namespace

extern "C" void mexFunction(int nlhs, mxArray *plhs[], int nrhs, const mxArray *prhs[])
{
    (void)nlhs;
    // ---------------------------------------------------------------- string commands
    if (nrhs >= 1 && mxIsChar(prhs[0])) {
        char cmd[16] = {0};
        {
            char *c = mxArrayToString(prhs[0]);
            if (c) { snprintf(cmd, sizeof cmd, "%s", c); mxFree(c); }
        }
        if (strcmp(cmd, "taps") == 0 && nrhs == 2) {            // [lo, hi] = nd_dwt_mex('taps', wname)
            char name[32] = {0};
            char *nm = mxArrayToString(prhs[1]);
            if (nm) { snprintf(name, sizeof name, "%s", nm); mxFree(nm); }
            double lo[20], hi[20];
            int len = 0;
            if (nddwt_wave_filters(name, lo, hi, &len)) mexErrMsgIdAndTxt("nddwt:wavelet", "Unknown Wavelet Name");
            plhs[0] = mxCreateDoubleMatrix(1, (mwSize)len, mxREAL);
            plhs[1] = mxCreateDoubleMatrix(1, (mwSize)len, mxREAL);
            memcpy(mxGetDoubles(plhs[0]), lo, sizeof(double) * len);
            memcpy(mxGetDoubles(plhs[1]), hi, sizeof(double) * len);
            return;
        }
        if (strcmp(cmd, "plan") == 0 && nrhs >= 5) {            // h = nd_dwt_mex('plan', f, is_single, is_complex, l2 [, ngpus])
            Desc ds;
            const char *err = parse_desc(prhs[1], &ds);
            if (err) fail(err);
            const int dtype = dtype_code(mxGetScalar(prhs[2]) != 0, mxGetScalar(prhs[3]) != 0);
            const int ngpus = nrhs >= 6 ? (int)mxGetScalar(prhs[5]) : 1;
            const int slot = new_entry(ds, dtype, mxGetScalar(prhs[4]) != 0, ngpus < 1 ? 1 : ngpus, NULL);
            plhs[0] = mxCreateNumericMatrix(1, 1, mxUINT64_CLASS, mxREAL);
            *reinterpret_cast<uint64_t *>(mxGetData(plhs[0])) = (uint64_t)slot + 1;
            return;
        }
        if (strcmp(cmd, "shrink") == 0 && nrhs == 3) {          // nd_dwt_mex('shrink', h, table)
            Entry *e = entry_of_handle(prhs[1]);
            if (!e) fail("not a plan handle");
            int rc;
            if (mxIsEmpty(prhs[2])) {
                rc = e->plan ? nddwt_plan_set_shrink(e->plan, 0, NULL, 0) : nddwt_mplan_set_shrink(e->mplan, 0, NULL, 0);
            } else {
                const int nd = 1 << e->d;
                const mwSize *td = mxGetDimensions(prhs[2]);
                if (!mxIsDouble(prhs[2]) || mxIsComplex(prhs[2]) || mxGetNumberOfDimensions(prhs[2]) != 2 || (int)td[1] != nd ||
                    td[0] < 1 || td[0] > NDDWT_MAX_LEVELS)
                    fail("threshold table must be a real double [levels x 2^d] matrix");
                // MATLAB is column-major: table(j, b) -> row-major [j][b] for the library
                double tab[NDDWT_MAX_LEVELS * (1 << NDDWT_MAX_DIMS)];
                const double *src = mxGetDoubles(prhs[2]);
                const int J = (int)td[0];
                for (int j = 0; j < J; ++j)
                    for (int b = 0; b < nd; ++b) tab[j * nd + b] = src[(size_t)b * J + j];
                rc = e->plan ? nddwt_plan_set_shrink(e->plan, 1, tab, J) : nddwt_mplan_set_shrink(e->mplan, 1, tab, J);
            }
            if (rc) fail_lib();
            return;
        }
        if (strcmp(cmd, "dilations") == 0 && nrhs == 3) {       // nd_dwt_mex('dilations', h, dil)   a-trous: dilation per level
            Entry *e = entry_of_handle(prhs[1]);
            if (!e) fail("not a plan handle");
            const int n = (int)mxGetNumberOfElements(prhs[2]);
            if (!mxIsDouble(prhs[2]) || n < 1 || n > NDDWT_MAX_LEVELS) fail("dilations: 1..16 positive values");
            int dil[NDDWT_MAX_LEVELS];
            const double *src = mxGetDoubles(prhs[2]);
            for (int j = 0; j < n; ++j) dil[j] = (int)src[j];
            const int rc = e->plan ? nddwt_plan_set_dilations(e->plan, dil, n) : nddwt_mplan_set_dilations(e->mplan, dil, n);
            if (rc) fail_lib();
            return;
        }
        if (strcmp(cmd, "release") == 0) {
            if (nrhs == 1) { release_all(); return; }
            Entry *e = entry_of_handle(prhs[1]);
            if (e) {
                if (e->plan) nddwt_plan_destroy(e->plan);
                if (e->mplan) nddwt_mplan_destroy(e->mplan);
                e->used = false;
            }
            return;
        }
        fail("unknown nd_dwt_mex command");
    }

    // ---------------------------------------------------------------- y = nd_dwt_mex(x, f, dir, level, pres_l2_norm)
    if (nrhs < 5) fail("Four Inputs Required");                         // nd_dwt_mex.c:19-22
    const mxArray *x = prhs[0], *f = prhs[1];
    const int dir = (int)mxGetScalar(prhs[2]);
    const int level = (int)mxGetScalar(prhs[3]);
    const int l2 = mxGetScalar(prhs[4]) != 0;
    if (level < 1 || level > NDDWT_MAX_LEVELS) fail("level must be in 1..16");

#ifdef NDDWT_MEX_GPU
    if (mxIsGPUArray(x)) {
        // ---- device-resident branch: no host round trip
        if (mxInitGPU() != MX_GPU_SUCCESS) fail("could not initialise the MATLAB GPU API");
        const mxGPUArray *xg = mxGPUCreateFromMxArray(x);
        const mxClassID xcls = mxGPUGetClassID(xg);
        const bool is_cx = mxGPUGetComplexity(xg) == mxCOMPLEX;
        const int64_t nx = (int64_t)mxGPUGetNumberOfElements(xg);
        const char *err = NULL;
        Entry *e = NULL;
        if (xcls != mxDOUBLE_CLASS && xcls != mxSINGLE_CLASS) err = "Arrays must be double or single";
        const int dtype = dtype_code(xcls == mxSINGLE_CLASS, is_cx);
        if (!err) {
            e = entry_of_handle(f);
            if (!e && mxIsStruct(f)) {
                mxGPUDestroyGPUArray(xg);                 // entry_of_struct may raise: nothing of ours may be alive then
                e = entry_of_struct(f, dtype, l2);
                xg = mxGPUCreateFromMxArray(x);
            }
            if (!e) err = "FIlter size and image size not consistant";
            else if (e->dtype != dtype) err = "plan and data differ in class or complexity";
            else if (!e->plan) err = "multi-GPU plans take host arrays (every GPU loads its own slab); gather(x) first";
        }
        int64_t nb = 0, numel = 1;
        if (!err) {
            nb = nddwt_num_bands(e->d, level);
            for (int i = 0; i < e->d; ++i) numel *= e->dims[i];
            if ((dir == 0 && nx != numel) || (dir != 0 && nx != numel * nb)) err = "FIlter size and image size not consistant";
        }
        if (err) { mxGPUDestroyGPUArray(xg); fail(err); }
        mwSize odims[NDDWT_MAX_DIMS + 2], nd_out;
        out_dims(*e, dir == 0, nb, odims, &nd_out);
        mxGPUArray *yg = mxGPUCreateGPUArray(nd_out, odims, xcls, is_cx ? mxCOMPLEX : mxREAL, MX_GPU_DO_NOT_INITIALIZE);
        const int rc = dir == 0 ? nddwt_dec(e->plan, mxGPUGetDataReadOnly(xg), mxGPUGetData(yg), level, NULL)
                                : nddwt_rec(e->plan, mxGPUGetDataReadOnly(xg), mxGPUGetData(yg), level, NULL);
        if (rc == 0) plhs[0] = mxGPUCreateMxArrayOnGPU(yg);
        mxGPUDestroyGPUArray(xg);
        mxGPUDestroyGPUArray(yg);
        if (rc) fail_lib();
        return;
    }
#endif

    // ---- host arrays
    if (!mxIsDouble(x) && !mxIsSingle(x)) fail("Arrays must be double or single");   // nd_dwt_mex.c:23-26 (single is new)
    const bool is_single = mxIsSingle(x), is_cx = mxIsComplex(x);
    const int dtype = dtype_code(is_single, is_cx);
    Entry *e = entry_of_handle(f);
    if (!e) e = entry_of_struct(f, dtype, l2);          // raises on a malformed descriptor
    if (e->dtype != dtype) fail("plan and data differ in class or complexity");
    const int64_t nb = nddwt_num_bands(e->d, level);
    int64_t numel = 1;
    for (int i = 0; i < e->d; ++i) numel *= e->dims[i];
    const int64_t nx = (int64_t)mxGetNumberOfElements(x);
    if ((dir == 0 && nx != numel) || (dir != 0 && nx != numel * nb))
        fail("FIlter size and image size not consistant");              // nd_dwt_mex.c:36-51,124-127
    mwSize odims[NDDWT_MAX_DIMS + 2], nd_out;
    out_dims(*e, dir == 0, nb, odims, &nd_out);
    plhs[0] = mxCreateNumericArray(nd_out, odims, is_single ? mxSINGLE_CLASS : mxDOUBLE_CLASS, is_cx ? mxCOMPLEX : mxREAL);
    const int64_t ny = dir == 0 ? numel * nb : numel;
    (void)ny;
    const void *xin = mxGetData(x);
    void *yout = mxGetData(plhs[0]);
#ifdef NDDWT_MEX_SPLIT_COMPLEX
    // legacy API: real and imaginary parts live in two arrays (mxGetPr / mxGetPi, nd_dwt_mex.c:55-58);
    // the library wants interleaved pairs.  mxMalloc memory is released by MATLAB if an error is raised.
    void *xi = NULL, *yi = NULL;
    if (is_cx) {
        const size_t es = is_single ? sizeof(float) : sizeof(double);
        xi = mxMalloc((size_t)nx * 2 * es);
        yi = mxMalloc((size_t)ny * 2 * es);
        const char *re = reinterpret_cast<const char *>(mxGetData(x)), *im = reinterpret_cast<const char *>(mxGetImagData(x));
        for (int64_t i = 0; i < nx; ++i) {
            memcpy(reinterpret_cast<char *>(xi) + (size_t)(2 * i) * es, re + (size_t)i * es, es);
            memcpy(reinterpret_cast<char *>(xi) + (size_t)(2 * i + 1) * es, im + (size_t)i * es, es);
        }
        xin = xi;
        yout = yi;
    }
#endif
    int rc;
    if (e->plan) rc = dir == 0 ? nddwt_dec_host(e->plan, xin, yout, level) : nddwt_rec_host(e->plan, xin, yout, level);
    else rc = dir == 0 ? nddwt_mplan_dec_host(e->mplan, xin, yout, level) : nddwt_mplan_rec_host(e->mplan, xin, yout, level);
    if (rc) fail_lib();
#ifdef NDDWT_MEX_SPLIT_COMPLEX
    if (is_cx) {
        const size_t es = is_single ? sizeof(float) : sizeof(double);
        char *re = reinterpret_cast<char *>(mxGetData(plhs[0])), *im = reinterpret_cast<char *>(mxGetImagData(plhs[0]));
        for (int64_t i = 0; i < ny; ++i) {
            memcpy(re + (size_t)i * es, reinterpret_cast<const char *>(yi) + (size_t)(2 * i) * es, es);
            memcpy(im + (size_t)i * es, reinterpret_cast<const char *>(yi) + (size_t)(2 * i + 1) * es, es);
        }
        mxFree(xi);
        mxFree(yi);
    }
#endif
}
