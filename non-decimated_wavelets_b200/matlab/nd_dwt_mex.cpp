// nd_dwt_mex.cpp -- MEX gateway over libnddwt_b200.so.  Drop-in for mex/nd_dwt_mex.c of
// arg-min-x/Non-Decimated_Wavelets: same function name, same five-argument call
//
//     y = nd_dwt_mex(x, f, dir, level, pres_l2_norm)
//
// with these differences (the spatial kernels need no Fourier-domain filters):
//   * x  is the SPATIAL array (dir == 0) or the coefficient stack [sizes, nb] (dir ~= 0),
//        single or double, real or complex (the reference accepts only complex double,
//        nd_dwt_mex.c:23-30, and wants fftn(x));
//   * f  is a struct with fields  wname (cell of d strings)  and  sizes (1 x d double) built by the
//        nd_dwt_*D classes in this directory, in place of the 2^d*numel stored filters f_dec;
//   * inputs are never modified (the reference overwrites prhs[0], nddwt.c:163,264-265).
// Extra entry used by wave_filters.m:  [lo, hi] = nd_dwt_mex('taps', wname).
// Build (needs MATLAB R2018a+):  mex -R2018a nd_dwt_mex.cpp -I../../include -L.. -lnddwt_b200
#include <string.h>
#include <vector>
#include <string>
#include "mex.h"
#include "../../include/nddwt_b200.h"

static void fail(const char *msg) { mexErrMsgIdAndTxt("MATLAB:FFT2mx:invalidNumInputs", "%s", msg); }

static void check(int rc)
{
    if (rc != 0) fail(nddwt_last_error());
}

// plans are the "stored filters": keep them across calls (iterative algorithms call dec/rec hundreds
// of times with the same geometry), release them when the MEX file is cleared
struct CachedPlan { std::string key; nddwt_plan *plan; };
static std::vector<CachedPlan> g_plans;
static void release_plans(void)
{
    for (size_t i = 0; i < g_plans.size(); ++i) nddwt_plan_destroy(g_plans[i].plan);
    g_plans.clear();
}
static nddwt_plan *get_plan(int d, const int64_t *dims, const char *const *names, int dtype, int l2)
{
    std::string key = std::to_string(dtype) + ":" + std::to_string(l2);
    for (int i = 0; i < d; ++i) key += std::string(":") + names[i] + "/" + std::to_string((long long)dims[i]);
    for (size_t i = 0; i < g_plans.size(); ++i)
        if (g_plans[i].key == key) return g_plans[i].plan;
    nddwt_plan *plan = nullptr;
    check(nddwt_plan_create(&plan, d, dims, names, dtype, l2, 0));
    if (g_plans.empty()) mexAtExit(release_plans);
    g_plans.push_back(CachedPlan{key, plan});
    return plan;
}

static int dtype_of(const mxArray *a)
{
    if (mxIsDouble(a)) return mxIsComplex(a) ? NDDWT_C128 : NDDWT_F64;
    if (mxIsSingle(a)) return mxIsComplex(a) ? NDDWT_C64 : NDDWT_F32;
    fail("Arrays must be double or single");
    return -1;
}

extern "C" void mexFunction(int nlhs, mxArray *plhs[], int nrhs, const mxArray *prhs[])
{
    (void)nlhs;
    if (nrhs == 2 && mxIsChar(prhs[0])) {          // [lo, hi] = nd_dwt_mex('taps', wname)
        char *name = mxArrayToString(prhs[1]);
        double lo[20], hi[20];
        int len = 0;
        int rc = nddwt_wave_filters(name, lo, hi, &len);
        mxFree(name);
        if (rc) mexErrMsgIdAndTxt("nddwt:wavelet", "Unknown Wavelet Name");
        plhs[0] = mxCreateDoubleMatrix(1, (mwSize)len, mxREAL);
        plhs[1] = mxCreateDoubleMatrix(1, (mwSize)len, mxREAL);
        memcpy(mxGetDoubles(plhs[0]), lo, sizeof(double) * len);
        memcpy(mxGetDoubles(plhs[1]), hi, sizeof(double) * len);
        return;
    }
    if (nrhs < 5) fail("Four Inputs Required");                        // nd_dwt_mex.c:19-22
    const mxArray *x = prhs[0], *f = prhs[1];
    if (!mxIsStruct(f)) fail("FIlter size and image size not consistant");
    const mxArray *fw = mxGetField(f, 0, "wname"), *fs = mxGetField(f, 0, "sizes");
    if (!fw || !fs || !mxIsCell(fw)) fail("FIlter size and image size not consistant");
    const int d = (int)mxGetNumberOfElements(fs);
    if (d < 1 || d > NDDWT_MAX_DIMS || (int)mxGetNumberOfElements(fw) != d)
        fail("FIlter size and image size not consistant");
    int64_t dims[NDDWT_MAX_DIMS];
    const double *sz = mxGetDoubles(fs);
    std::vector<std::string> names(d);
    const char *cnames[NDDWT_MAX_DIMS];
    int64_t numel = 1;
    for (int i = 0; i < d; ++i) {
        dims[i] = (int64_t)sz[i];
        numel *= dims[i];
        char *s = mxArrayToString(mxGetCell(fw, i));
        names[i] = s;
        mxFree(s);
        cnames[i] = names[i].c_str();
    }
    const int dir = (int)mxGetScalar(prhs[2]);
    const int level = (int)mxGetScalar(prhs[3]);
    const int l2 = (int)mxGetScalar(prhs[4]);
    const int dtype = dtype_of(x);
    const int64_t nb = nddwt_num_bands(d, level);
    const int64_t nx = (int64_t)mxGetNumberOfElements(x);
    if ((dir == 0 && nx != numel) || (dir != 0 && nx != numel * nb))
        fail("FIlter size and image size not consistant");             // nd_dwt_mex.c:36-51,124-127

    nddwt_plan *plan = get_plan(d, dims, cnames, dtype, l2);
    const mxClassID cls = (dtype == NDDWT_F32 || dtype == NDDWT_C64) ? mxSINGLE_CLASS : mxDOUBLE_CLASS;
    const mxComplexity cx = (dtype == NDDWT_C64 || dtype == NDDWT_C128) ? mxCOMPLEX : mxREAL;
    mwSize odims[NDDWT_MAX_DIMS + 2];
    for (int i = 0; i < d; ++i) odims[i] = (mwSize)dims[i];
    int rc;
    if (dir == 0) {                                                     // forward: [sizes, nb]
        mwSize nd_out = (mwSize)d + 1;
        odims[d] = (mwSize)nb;
        if (d == 1) { odims[0] = (mwSize)dims[0]; odims[1] = (mwSize)nb; nd_out = 2; }
        plhs[0] = mxCreateNumericArray(nd_out, odims, cls, cx);         // nd_dwt_mex.c:79-88
        rc = nddwt_dec_host(plan, mxGetData(x), mxGetData(plhs[0]), level);
    } else {                                                            // inverse: sizes
        mwSize nd_out = (mwSize)d;
        if (d == 1) { odims[1] = 1; nd_out = 2; }
        plhs[0] = mxCreateNumericArray(nd_out, odims, cls, cx);         // nd_dwt_mex.c:136-138
        rc = nddwt_rec_host(plan, mxGetData(x), mxGetData(plhs[0]), level);
    }
    check(rc);
}
