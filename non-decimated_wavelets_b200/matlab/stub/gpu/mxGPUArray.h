/* Declarations-only stand-in for MATLAB's gpu/mxGPUArray.h (Parallel Computing Toolbox), just enough to
 * COMPILE-CHECK the gpuArray branch of nd_dwt_mex.cpp in this image.  Names and signatures follow the
 * documented mxGPU API; nothing here is linked into anything that runs. */
#ifndef NDDWT_STUB_MXGPUARRAY_H
#define NDDWT_STUB_MXGPUARRAY_H
#include "mex.h"
#ifdef __cplusplus
extern "C" {
#endif
typedef struct mxGPUArray_tag mxGPUArray;
typedef enum { MX_GPU_DO_NOT_INITIALIZE = 0, MX_GPU_INITIALIZE_VALUES = 1 } mxGPUInitialize;
#define MX_GPU_SUCCESS 0
int mxInitGPU(void);
int mxIsGPUArray(const mxArray *a);
const mxGPUArray *mxGPUCreateFromMxArray(const mxArray *a);
mxGPUArray *mxGPUCreateGPUArray(mwSize ndims, const mwSize *dims, mxClassID cls, mxComplexity c, mxGPUInitialize init);
mxClassID mxGPUGetClassID(const mxGPUArray *a);
mxComplexity mxGPUGetComplexity(const mxGPUArray *a);
mwSize mxGPUGetNumberOfElements(const mxGPUArray *a);
const void *mxGPUGetDataReadOnly(const mxGPUArray *a);
void *mxGPUGetData(mxGPUArray *a);
mxArray *mxGPUCreateMxArrayOnGPU(const mxGPUArray *a);
void mxGPUDestroyGPUArray(const mxGPUArray *a);
#ifdef __cplusplus
}
#endif
#endif
