/* xemo_mex_common.h -- shared plumbing of the MEX shims: one persistent xemo_ctx per MATLAB process (MATLAB
 * runs one GPU per process, also under spmd), mxArray / gpuArray -> xemo_array, name/value option parsing,
 * status code -> mexErrMsgIdAndTxt.  NOT EXECUTED UNDER MATLAB in this repository (no MATLAB in the image):
 * compile-checked against mex_stub.h only; all logic below the ABI is exercised through the same C entry
 * points from Python (tests/test_gpu_vl_ops.py). */
#ifndef XEMO_MEX_COMMON_H_
#define XEMO_MEX_COMMON_H_
#ifdef XEMO_MEX_STUB
#include "mex_stub.h"
#else
#include "mex.h"
#include "gpu/mxGPUArray.h"
#endif
#include <string.h>

#include "xemo.h"

static xemo_ctx* g_ctx = NULL;
static void xm_at_exit(void) { if (g_ctx) { xemo_destroy(g_ctx); g_ctx = NULL; } }

static xemo_ctx* xm_ctx(void) {
  if (!g_ctx) {
    int dev = 0;
    mxInitGPU();
    /* The device MATLAB selected (gpuDevice(opts.gpus(labindex)) under cnn_train_dag's spmd) is the CUDA runtime's
     * current device of this thread.  The context runs on MATLAB's own stream -- the legacy default stream, which is what
     * gpuArray arithmetic uses -- so that the library's kernels are ordered after the producers of its gpuArray inputs and
     * before the consumers of its gpuArray outputs without any extra synchronisation; calls that return CPU arrays
     * synchronise inside the library. */
    xemo_current_device(&dev);
    if (xemo_create(dev, XEMO_STREAM_LEGACY, &g_ctx))
      mexErrMsgIdAndTxt("xemo:nodevice", "libxemo needs an sm_100 (B200) device; there is no CPU fallback");
    mexAtExit(xm_at_exit);
  }
  return g_ctx;
}

static void xm_check(int status) {
  if (status) mexErrMsgIdAndTxt("xemo:error", "%s", xemo_last_error(g_ctx));
}

/* view of a single array (CPU or gpuArray) as xemo_array; *keep receives the mxGPUArray to destroy after the call */
static xemo_array xm_in(const mxArray* a, const mxGPUArray** keep) {
  xemo_array r;
  const mwSize* d;
  mwSize nd;
  *keep = NULL;
  memset(&r, 0, sizeof(r));
  if (!a || mxIsEmpty(a)) return r;
  nd = mxGetNumberOfDimensions(a);
  d = mxGetDimensions(a);
  r.h = (int64_t)d[0];
  r.w = nd > 1 ? (int64_t)d[1] : 1;
  r.c = nd > 2 ? (int64_t)d[2] : 1;
  r.n = nd > 3 ? (int64_t)d[3] : 1;
  if (mxIsGPUArray(a)) {
    *keep = mxGPUCreateFromMxArray(a);
    r.data = (void*)mxGPUGetDataReadOnly(*keep);
  } else {
    if (!mxIsSingle(a)) mexErrMsgIdAndTxt("xemo:type", "arrays must be single");
    r.data = mxGetData(a);
  }
  return r;
}

/* the same for network inputs, which may also be uint8 (grey faces, teacher/ferplus_baselines.m:181) */
static xemo_array xm_in_any(const mxArray* a, const mxGPUArray** keep) {
  xemo_array r;
  const mwSize* d;
  mwSize nd;
  *keep = NULL;
  memset(&r, 0, sizeof(r));
  if (!a || mxIsEmpty(a)) return r;
  nd = mxGetNumberOfDimensions(a);
  d = mxGetDimensions(a);
  r.h = (int64_t)d[0];
  r.w = nd > 1 ? (int64_t)d[1] : 1;
  r.c = nd > 2 ? (int64_t)d[2] : 1;
  r.n = nd > 3 ? (int64_t)d[3] : 1;
  if (mxIsGPUArray(a)) {
    *keep = mxGPUCreateFromMxArray(a);
    r.data = (void*)mxGPUGetDataReadOnly(*keep);
  } else {
    r.data = mxGetData(a);
  }
  return r;
}

/* allocate an output where the input lives (gpuArray in -> gpuArray out), return its xemo_array view */
static xemo_array xm_out(mxArray** plhs, int on_gpu, int64_t h, int64_t w, int64_t c, int64_t n) {
  xemo_array r;
  mwSize dims[4];
  dims[0] = (mwSize)h; dims[1] = (mwSize)w; dims[2] = (mwSize)c; dims[3] = (mwSize)n;
  r.h = h; r.w = w; r.c = c; r.n = n;
  if (on_gpu) {
    mxGPUArray* g = mxGPUCreateGPUArray(4, dims, mxSINGLE_CLASS, mxREAL, MX_GPU_DO_NOT_INITIALIZE);
    r.data = mxGPUGetData(g);
    *plhs = mxGPUCreateMxArrayOnGPU(g);
    mxGPUDestroyGPUArray(g);
  } else {
    *plhs = mxCreateNumericArray(4, dims, mxSINGLE_CLASS, mxREAL);
    r.data = mxGetData(*plhs);
  }
  return r;
}

/* 'pad' ([t b l r] or scalar) / 'stride' ([sy sx] or scalar) / 'pool' values */
static void xm_ints(const mxArray* v, int* out, int n) {
  size_t k = mxGetNumberOfElements(v), i;
  const double* p = mxGetPr(v);
  if (k != 1 && k != (size_t)n) mexErrMsgIdAndTxt("xemo:opt", "option must be a scalar or have %d elements", n);
  for (i = 0; i < (size_t)n; ++i) out[i] = (int)p[k == 1 ? 0 : i];
}
#endif
