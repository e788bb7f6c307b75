/* Declarations-only stand-in for MATLAB's mex.h / matrix.h, just enough to COMPILE-CHECK
 * nd_dwt_mex.cpp in this image (no MATLAB installed).  It follows the documented C Matrix API
 * (R2018a interleaved-complex names plus the legacy split-complex mxGetPr / mxGetPi / mxGetImagData used
 * behind NDDWT_MEX_SPLIT_COMPLEX); it is not linked into anything that runs. */
#ifndef NDDWT_STUB_MEX_H
#define NDDWT_STUB_MEX_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif
typedef struct mxArray_tag mxArray;
typedef size_t mwSize;
typedef enum { mxUNKNOWN_CLASS = 0, mxCELL_CLASS, mxSTRUCT_CLASS, mxLOGICAL_CLASS, mxCHAR_CLASS, mxVOID_CLASS,
               mxDOUBLE_CLASS, mxSINGLE_CLASS, mxINT8_CLASS, mxUINT8_CLASS, mxINT16_CLASS, mxUINT16_CLASS,
               mxINT32_CLASS, mxUINT32_CLASS, mxINT64_CLASS, mxUINT64_CLASS } mxClassID;
typedef enum { mxREAL = 0, mxCOMPLEX } mxComplexity;
int mexAtExit(void (*fn)(void));
void mexErrMsgIdAndTxt(const char *id, const char *fmt, ...);
void mexLock(void);
bool mxIsDouble(const mxArray *a);
bool mxIsSingle(const mxArray *a);
bool mxIsComplex(const mxArray *a);
bool mxIsChar(const mxArray *a);
bool mxIsCell(const mxArray *a);
bool mxIsStruct(const mxArray *a);
bool mxIsUint64(const mxArray *a);
bool mxIsEmpty(const mxArray *a);
mxClassID mxGetClassID(const mxArray *a);
double mxGetScalar(const mxArray *a);
mwSize mxGetNumberOfDimensions(const mxArray *a);
const mwSize *mxGetDimensions(const mxArray *a);
size_t mxGetNumberOfElements(const mxArray *a);
void *mxGetData(const mxArray *a);
void *mxGetImagData(const mxArray *a);      /* legacy split-complex storage */
double *mxGetPr(const mxArray *a);
double *mxGetPi(const mxArray *a);
mxArray *mxGetField(const mxArray *a, mwSize index, const char *name);
mxArray *mxGetCell(const mxArray *a, mwSize index);
char *mxArrayToString(const mxArray *a);
void *mxMalloc(size_t n);
void mxFree(void *p);
mxArray *mxCreateNumericArray(mwSize ndim, const mwSize *dims, mxClassID cls, mxComplexity c);
mxArray *mxCreateNumericMatrix(mwSize m, mwSize n, mxClassID cls, mxComplexity c);
mxArray *mxCreateDoubleMatrix(mwSize m, mwSize n, mxComplexity c);
double *mxGetDoubles(const mxArray *a);
#ifdef __cplusplus
}
#endif
#endif
