/* mex_stub.h -- the minimal subset of MATLAB's mex.h / gpu/mxGPUArray.h the shims in this directory use,
 * declared here ONLY so that the shims can be compile-checked in an image without MATLAB
 * (`gcc -fsyntax-only -DXEMO_MEX_STUB -I../include mex/vl_nnconv_mex.c`).  Under MATLAB, build with `mex` and the
 * real headers; this file is then not included. */
#ifndef XEMO_MEX_STUB_H_
#define XEMO_MEX_STUB_H_
#include <stddef.h>
typedef struct mxArray_tag mxArray;
typedef struct mxGPUArray_tag mxGPUArray;
typedef size_t mwSize;
typedef enum { mxSINGLE_CLASS = 7, mxUINT8_CLASS = 9 } mxClassID;
typedef enum { mxREAL = 0 } mxComplexity;
typedef enum { MX_GPU_DO_NOT_INITIALIZE = 0, MX_GPU_INITIALIZE_VALUES = 1 } mxGPUInitialize;
int mxIsGPUArray(const mxArray*);
int mxIsEmpty(const mxArray*);
int mxIsChar(const mxArray*);
int mxIsSingle(const mxArray*);
mwSize mxGetNumberOfDimensions(const mxArray*);
const mwSize* mxGetDimensions(const mxArray*);
size_t mxGetNumberOfElements(const mxArray*);
void* mxGetData(const mxArray*);
double* mxGetPr(const mxArray*);
double mxGetScalar(const mxArray*);
char* mxArrayToString(const mxArray*);
void mxFree(void*);
mxArray* mxCreateNumericArray(mwSize, const mwSize*, mxClassID, mxComplexity);
mxArray* mxCreateDoubleScalar(double);
int mxInitGPU(void);
const mxGPUArray* mxGPUCreateFromMxArray(const mxArray*);
const void* mxGPUGetDataReadOnly(const mxGPUArray*);
void* mxGPUGetData(mxGPUArray*);
mxGPUArray* mxGPUCreateGPUArray(mwSize, const mwSize*, mxClassID, mxComplexity, mxGPUInitialize);
mxArray* mxGPUCreateMxArrayOnGPU(const mxGPUArray*);
void mxGPUDestroyGPUArray(const mxGPUArray*);
void mexErrMsgIdAndTxt(const char*, const char*, ...);
int mexAtExit(void (*)(void));
#endif
