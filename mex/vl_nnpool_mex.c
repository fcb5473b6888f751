/* vl_nnpool MEX gateway over libxemo.so -- replaces MatConvNet's matlab/src/vl_nnpool.cu gateway.
 *   Y = vl_nnpool(X, POOL, 'pad', P, 'stride', S, 'method', M) ;  DX = vl_nnpool(X, POOL, DZDY, ...)
 * (dagnn.Pooling; `pool6` is re-sized at emoVoxCeleb/emoVoxZoo.m:263-269.)  Source-only: see xemo_mex_common.h. */
#include "xemo_mex_common.h"

void mexFunction(int nlhs, mxArray* plhs[], int nrhs, const mxArray* prhs[]) {
  int pad[4] = {0, 0, 0, 0}, stride[2] = {1, 1}, pool[2], method = 0, next = 2, i, on_gpu;
  const mxGPUArray *kx, *kd = NULL;
  xemo_array x, dzdy, out;
  int64_t oh, ow;
  xemo_ctx* ctx = xm_ctx();
  (void)nlhs;
  if (nrhs < 2) mexErrMsgIdAndTxt("xemo:args", "vl_nnpool needs X and POOL");
  x = xm_in(prhs[0], &kx);
  xm_ints(prhs[1], pool, 2);
  memset(&dzdy, 0, sizeof(dzdy));
  if (nrhs > 2 && !mxIsChar(prhs[2])) { dzdy = xm_in(prhs[2], &kd); next = 3; }
  for (i = next; i + 1 < nrhs; i += 2) {
    char* name = mxArrayToString(prhs[i]);
    if (!strcmp(name, "pad")) xm_ints(prhs[i + 1], pad, 4);
    else if (!strcmp(name, "stride")) xm_ints(prhs[i + 1], stride, 2);
    else if (!strcmp(name, "method")) { char* m = mxArrayToString(prhs[i + 1]); method = !strcmp(m, "avg"); if (!method && strcmp(m, "max")) mexErrMsgIdAndTxt("xemo:opt", "unknown pooling method"); mxFree(m); }
    mxFree(name);
  }
  on_gpu = mxIsGPUArray(prhs[0]);
  if (!dzdy.data) {
    if (xemo_out_size(x.h, x.w, pool[0], pool[1], pad, stride, &oh, &ow)) mexErrMsgIdAndTxt("xemo:size", "pooling window larger than the padded input");
    out = xm_out(&plhs[0], on_gpu, oh, ow, x.c, x.n);
    xm_check(xemo_vl_nnpool(ctx, &x, pool, NULL, pad, stride, method, &out, NULL));
  } else {
    out = xm_out(&plhs[0], on_gpu, x.h, x.w, x.c, x.n);
    xm_check(xemo_vl_nnpool(ctx, &x, pool, &dzdy, pad, stride, method, &out, NULL));
  }
  if (kx) mxGPUDestroyGPUArray(kx);
  if (kd) mxGPUDestroyGPUArray(kd);
}
