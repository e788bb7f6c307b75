// mock_mex.cpp -- a small stand-in for MATLAB's MEX / C Matrix API runtime (and, with -DNDDWT_MEX_GPU, of the mxGPU
// API of the Parallel Computing Toolbox), TEST INFRASTRUCTURE ONLY: MATLAB is not installed in this image, so this is
// what lets tests/test_mex_gateway.py EXECUTE non-decimated_wavelets_b200/matlab/nd_dwt_mex.cpp -- the gateway a MATLAB
// user would build with mex / mexcuda -- instead of only compile-checking it.  It implements exactly the functions
// declared in matlab/stub/mex.h and matlab/stub/gpu/mxGPUArray.h with the documented semantics the gateway relies on:
// column-major numeric arrays (interleaved complex, or split real / imaginary storage with
// -DNDDWT_MEX_SPLIT_COMPLEX like the pre-R2018a API the reference gateway uses, mex/nd_dwt_mex.c:55-58), char arrays,
// cell arrays, 1 x 1 structs, mxMalloc / mxFree, mexAtExit, and mexErrMsgIdAndTxt that does NOT return (here: a C++
// exception caught in mock_call, where MATLAB would longjmp back into the interpreter).
// The mock_* functions at the bottom are the driver interface the Python test binds with ctypes.
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <utility>
#include <vector>
#include "mex.h"
#ifdef NDDWT_MEX_GPU
#include <cuda_runtime_api.h>
#include "gpu/mxGPUArray.h"
#endif

struct DeviceBuf {
    void *ptr = nullptr;
    int refs = 0;
};

struct mxArray_tag {
    mxClassID cls = mxDOUBLE_CLASS;
    bool cplx = false;
    std::vector<mwSize> dims;
    void *re = nullptr;   // interleaved build: all the data; split build: the real parts
    void *im = nullptr;   // split build only
    std::string chars;    // mxCHAR_CLASS
    std::vector<mxArray *> cells;
    std::vector<std::pair<std::string, mxArray *>> fields;
    DeviceBuf *dev = nullptr;   // gpuArray: the data live on the device
};

namespace {

struct MockMexError {
    std::string id, msg;
};

void (*g_atexit)(void) = nullptr;
int g_live_mallocs = 0;

size_t elem_size(mxClassID c)
{
    switch (c) {
        case mxDOUBLE_CLASS: case mxINT64_CLASS: case mxUINT64_CLASS: return 8;
        case mxSINGLE_CLASS: case mxINT32_CLASS: case mxUINT32_CLASS: return 4;
        case mxINT16_CLASS: case mxUINT16_CLASS: case mxCHAR_CLASS: return 2;
        default: return 1;
    }
}

size_t numel(const mxArray *a)
{
    if (a->cls == mxCHAR_CLASS) return a->chars.size();
    if (a->cls == mxCELL_CLASS) return a->cells.size();
    if (a->cls == mxSTRUCT_CLASS) return 1;
    size_t n = 1;
    for (mwSize d : a->dims) n *= d;
    return a->dims.empty() ? 0 : n;
}

mxArray *new_numeric(mwSize ndim, const mwSize *dims, mxClassID cls, mxComplexity c, bool allocate)
{
    mxArray *a = new mxArray_tag;
    a->cls = cls;
    a->cplx = (c == mxCOMPLEX);
    a->dims.assign(dims, dims + ndim);
    while (a->dims.size() < 2) a->dims.push_back(1);
    if (allocate) {
        const size_t bytes = numel(a) * elem_size(cls);
#ifdef NDDWT_MEX_SPLIT_COMPLEX
        a->re = calloc(bytes ? bytes : 1, 1);
        if (a->cplx) a->im = calloc(bytes ? bytes : 1, 1);
#else
        a->re = calloc((a->cplx ? 2 : 1) * bytes + 1, 1);
#endif
    }
    return a;
}

}  // namespace

extern "C" {

int mexAtExit(void (*fn)(void)) { g_atexit = fn; return 0; }
void mexLock(void) {}
void mexErrMsgIdAndTxt(const char *id, const char *fmt, ...)
{
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    throw MockMexError{id ? id : "", buf};
}

bool mxIsDouble(const mxArray *a) { return a && a->cls == mxDOUBLE_CLASS; }
bool mxIsSingle(const mxArray *a) { return a && a->cls == mxSINGLE_CLASS; }
bool mxIsComplex(const mxArray *a) { return a && a->cplx; }
bool mxIsChar(const mxArray *a) { return a && a->cls == mxCHAR_CLASS; }
bool mxIsCell(const mxArray *a) { return a && a->cls == mxCELL_CLASS; }
bool mxIsStruct(const mxArray *a) { return a && a->cls == mxSTRUCT_CLASS; }
bool mxIsUint64(const mxArray *a) { return a && a->cls == mxUINT64_CLASS; }
bool mxIsEmpty(const mxArray *a) { return !a || numel(a) == 0; }
mxClassID mxGetClassID(const mxArray *a) { return a->cls; }
double mxGetScalar(const mxArray *a)
{
    if (!a || numel(a) == 0 || !a->re) return 0.0;
    switch (a->cls) {
        case mxDOUBLE_CLASS: return *static_cast<const double *>(a->re);
        case mxSINGLE_CLASS: return *static_cast<const float *>(a->re);
        case mxUINT64_CLASS: return (double)*static_cast<const uint64_t *>(a->re);
        case mxINT32_CLASS: return *static_cast<const int32_t *>(a->re);
        default: return 0.0;
    }
}
mwSize mxGetNumberOfDimensions(const mxArray *a) { return a->dims.size(); }
const mwSize *mxGetDimensions(const mxArray *a) { return a->dims.data(); }
size_t mxGetNumberOfElements(const mxArray *a) { return numel(a); }
void *mxGetData(const mxArray *a) { return a->re; }
void *mxGetImagData(const mxArray *a) { return a->im; }
double *mxGetPr(const mxArray *a) { return static_cast<double *>(a->re); }
double *mxGetPi(const mxArray *a) { return static_cast<double *>(a->im); }
double *mxGetDoubles(const mxArray *a) { return static_cast<double *>(a->re); }
mxArray *mxGetField(const mxArray *a, mwSize index, const char *name)
{
    if (!a || a->cls != mxSTRUCT_CLASS || index != 0) return nullptr;
    for (const auto &f : a->fields)
        if (f.first == name) return f.second;
    return nullptr;
}
mxArray *mxGetCell(const mxArray *a, mwSize index)
{
    if (!a || a->cls != mxCELL_CLASS || index >= a->cells.size()) return nullptr;
    return a->cells[index];
}
void *mxMalloc(size_t n) { ++g_live_mallocs; return malloc(n ? n : 1); }
void mxFree(void *p) { if (p) { --g_live_mallocs; free(p); } }
char *mxArrayToString(const mxArray *a)
{
    if (!a || a->cls != mxCHAR_CLASS) return nullptr;
    char *s = static_cast<char *>(mxMalloc(a->chars.size() + 1));
    memcpy(s, a->chars.c_str(), a->chars.size() + 1);
    return s;
}
mxArray *mxCreateNumericArray(mwSize ndim, const mwSize *dims, mxClassID cls, mxComplexity c)
{
    return new_numeric(ndim, dims, cls, c, true);
}
mxArray *mxCreateNumericMatrix(mwSize m, mwSize n, mxClassID cls, mxComplexity c)
{
    const mwSize d[2] = {m, n};
    return new_numeric(2, d, cls, c, true);
}
mxArray *mxCreateDoubleMatrix(mwSize m, mwSize n, mxComplexity c) { return mxCreateNumericMatrix(m, n, mxDOUBLE_CLASS, c); }

#ifdef NDDWT_MEX_GPU
// ---- mxGPU API: a gpuArray is an mxArray whose data live in device memory (reference-counted buffer)
struct mxGPUArray_tag {
    mxClassID cls;
    bool cplx;
    std::vector<mwSize> dims;
    DeviceBuf *dev;
};
static int g_live_gpu_handles = 0;

int mxInitGPU(void)
{
    int n = 0;
    return (cudaGetDeviceCount(&n) == cudaSuccess && n > 0) ? MX_GPU_SUCCESS : 1;
}
int mxIsGPUArray(const mxArray *a) { return a && a->dev != nullptr; }
const mxGPUArray *mxGPUCreateFromMxArray(const mxArray *a)
{
    mxGPUArray *g = new mxGPUArray_tag{a->cls, a->cplx, a->dims, a->dev};
    g->dev->refs++;
    ++g_live_gpu_handles;
    return g;
}
mxGPUArray *mxGPUCreateGPUArray(mwSize ndims, const mwSize *dims, mxClassID cls, mxComplexity c, mxGPUInitialize init)
{
    mxGPUArray *g = new mxGPUArray_tag{cls, c == mxCOMPLEX, std::vector<mwSize>(dims, dims + ndims), new DeviceBuf};
    while (g->dims.size() < 2) g->dims.push_back(1);
    size_t n = 1;
    for (mwSize d : g->dims) n *= d;
    const size_t bytes = n * elem_size(cls) * (g->cplx ? 2 : 1);
    if (cudaMalloc(&g->dev->ptr, bytes ? bytes : 1) != cudaSuccess) mexErrMsgIdAndTxt("mock:gpu", "cudaMalloc failed");
    if (init == MX_GPU_INITIALIZE_VALUES) cudaMemset(g->dev->ptr, 0, bytes);
    g->dev->refs = 1;
    ++g_live_gpu_handles;
    return g;
}
mxClassID mxGPUGetClassID(const mxGPUArray *a) { return a->cls; }
mxComplexity mxGPUGetComplexity(const mxGPUArray *a) { return a->cplx ? mxCOMPLEX : mxREAL; }
mwSize mxGPUGetNumberOfElements(const mxGPUArray *a)
{
    size_t n = 1;
    for (mwSize d : a->dims) n *= d;
    return n;
}
const void *mxGPUGetDataReadOnly(const mxGPUArray *a) { return a->dev->ptr; }
void *mxGPUGetData(mxGPUArray *a) { return a->dev->ptr; }
mxArray *mxGPUCreateMxArrayOnGPU(const mxGPUArray *g)
{
    mxArray *a = new mxArray_tag;
    a->cls = g->cls;
    a->cplx = g->cplx;
    a->dims = g->dims;
    a->dev = g->dev;
    a->dev->refs++;
    return a;
}
static void release_dev(DeviceBuf *d)
{
    if (d && --d->refs == 0) {
        cudaFree(d->ptr);
        delete d;
    }
}
void mxGPUDestroyGPUArray(const mxGPUArray *g)
{
    if (!g) return;
    release_dev(g->dev);
    --g_live_gpu_handles;
    delete g;
}
#endif

// ------------------------------------------------------------------------------------------------------------
// driver interface (ctypes)
void mexFunction(int nlhs, mxArray *plhs[], int nrhs, const mxArray *prhs[]);

mxArray *mock_create_numeric(int ndim, const uint64_t *dims, int cls, int cplx)
{
    std::vector<mwSize> d(dims, dims + ndim);
    return new_numeric((mwSize)ndim, d.data(), (mxClassID)cls, cplx ? mxCOMPLEX : mxREAL, true);
}
mxArray *mock_create_string(const char *s)
{
    mxArray *a = new mxArray_tag;
    a->cls = mxCHAR_CLASS;
    a->chars = s;
    a->dims = {1, (mwSize)a->chars.size()};
    return a;
}
mxArray *mock_create_cell(int n)
{
    mxArray *a = new mxArray_tag;
    a->cls = mxCELL_CLASS;
    a->cells.assign((size_t)n, nullptr);
    a->dims = {1, (mwSize)n};
    return a;
}
void mock_set_cell(mxArray *c, int i, mxArray *v) { c->cells[(size_t)i] = v; }
mxArray *mock_create_struct(void)
{
    mxArray *a = new mxArray_tag;
    a->cls = mxSTRUCT_CLASS;
    a->dims = {1, 1};
    return a;
}
void mock_set_field(mxArray *s, const char *name, mxArray *v) { s->fields.emplace_back(name, v); }
void *mock_real(mxArray *a) { return a->re; }
void *mock_imag(mxArray *a) { return a->im; }
int mock_class(const mxArray *a) { return (int)a->cls; }
int mock_is_complex(const mxArray *a) { return a->cplx ? 1 : 0; }
int mock_is_gpu(const mxArray *a) { return a->dev ? 1 : 0; }
int mock_ndims(const mxArray *a) { return (int)a->dims.size(); }
uint64_t mock_dim(const mxArray *a, int i) { return (uint64_t)a->dims[(size_t)i]; }
int mock_split_complex(void)
{
#ifdef NDDWT_MEX_SPLIT_COMPLEX
    return 1;
#else
    return 0;
#endif
}
int mock_live_mallocs(void) { return g_live_mallocs; }

void mock_destroy(mxArray *a)   // deep: owns its cells and fields
{
    if (!a) return;
    for (mxArray *c : a->cells) mock_destroy(c);
    for (auto &f : a->fields) mock_destroy(f.second);
    free(a->re);
    free(a->im);
#ifdef NDDWT_MEX_GPU
    release_dev(a->dev);
#endif
    delete a;
}

#ifdef NDDWT_MEX_GPU
// gpuArray(x): a device copy of a host array (interleaved storage);  gather(g): the host copy of a gpuArray
mxArray *mock_gpu_array(const mxArray *h)
{
    mxArray *a = new mxArray_tag;
    a->cls = h->cls;
    a->cplx = h->cplx;
    a->dims = h->dims;
    a->dev = new DeviceBuf;
    a->dev->refs = 1;
    const size_t bytes = numel(h) * elem_size(h->cls) * (h->cplx ? 2 : 1);
    if (cudaMalloc(&a->dev->ptr, bytes ? bytes : 1) != cudaSuccess ||
        cudaMemcpy(a->dev->ptr, h->re, bytes, cudaMemcpyHostToDevice) != cudaSuccess) {
        mock_destroy(a);
        return nullptr;
    }
    return a;
}
mxArray *mock_gather(const mxArray *g)
{
    mxArray *h = new_numeric(g->dims.size(), g->dims.data(), g->cls, g->cplx ? mxCOMPLEX : mxREAL, true);
    const size_t bytes = numel(h) * elem_size(h->cls) * (h->cplx ? 2 : 1);
    if (cudaMemcpy(h->re, g->dev->ptr, bytes, cudaMemcpyDeviceToHost) != cudaSuccess) {
        mock_destroy(h);
        return nullptr;
    }
    return h;
}
int mock_live_gpu_handles(void) { return g_live_gpu_handles; }
#endif

// y = nd_dwt_mex(...): returns 0, or 1 with the error identifier and text when the gateway raised
int mock_call(int nlhs, mxArray **plhs, int nrhs, mxArray **prhs, char *err_id, char *err_msg, int err_len)
{
    try {
        mexFunction(nlhs, plhs, nrhs, const_cast<const mxArray **>(prhs));
    } catch (const MockMexError &e) {
        snprintf(err_id, (size_t)err_len, "%s", e.id.c_str());
        snprintf(err_msg, (size_t)err_len, "%s", e.msg.c_str());
        return 1;
    }
    return 0;
}
void mock_run_atexit(void)
{
    if (g_atexit) g_atexit();
}

}  // extern "C"
