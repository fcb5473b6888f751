/* vl_nnsoftmaxceloss gateway over libxemo.so -- replaces mcnExtraLayers' vl_nnsoftmaxceloss.m (pure MATLAB
 * gpuArray code upstream), the loss dagnn.SoftmaxCELoss('temperature', 2, 'logitTargets', true) calls
 * (emoVoxCeleb/emoVoxZoo.m:151-157).
 *   Y  = vl_nnsoftmaxceloss(X, P, 'temperature', T, 'logitTargets', tf, 'instanceWeights', W)
 *   DX = vl_nnsoftmaxceloss(X, P, DZDY, ...)
 * Source-only: see xemo_mex_common.h. */
#include "xemo_mex_common.h"

void mexFunction(int nlhs, mxArray* plhs[], int nrhs, const mxArray* prhs[]) {
  const mxGPUArray *kx, *kp, *kw = NULL;
  xemo_array x, p, w, dx;
  float T = 1.f, dz = 0.f, loss = 0.f;
  int lt = 0, next = 2, backward = 0, i;
  mwSize one[2] = {1, 1};
  xemo_ctx* ctx = xm_ctx();
  (void)nlhs;
  if (nrhs < 2) mexErrMsgIdAndTxt("xemo:args", "vl_nnsoftmaxceloss needs X and P");
  x = xm_in(prhs[0], &kx);
  p = xm_in(prhs[1], &kp);
  memset(&w, 0, sizeof(w));
  if (nrhs > 2 && !mxIsChar(prhs[2])) { if (!mxIsEmpty(prhs[2])) { dz = (float)mxGetScalar(prhs[2]); backward = 1; } next = 3; }
  for (i = next; i + 1 < nrhs; i += 2) {
    char* name = mxArrayToString(prhs[i]);
    if (!strcmp(name, "temperature")) T = (float)mxGetScalar(prhs[i + 1]);
    else if (!strcmp(name, "logitTargets")) lt = mxGetScalar(prhs[i + 1]) != 0;
    else if (!strcmp(name, "instanceWeights")) w = xm_in(prhs[i + 1], &kw);
    mxFree(name);
  }
  if (!backward) {
    xm_check(xemo_vl_nnsoftmaxceloss(ctx, &x, &p, NULL, T, lt, (const float*)w.data, &loss, NULL));
    plhs[0] = mxCreateNumericArray(2, one, mxSINGLE_CLASS, mxREAL);
    *(float*)mxGetData(plhs[0]) = loss;
  } else {
    dx = xm_out(&plhs[0], mxIsGPUArray(prhs[0]), x.h, x.w, x.c, x.n);
    xm_check(xemo_vl_nnsoftmaxceloss(ctx, &x, &p, &dz, T, lt, (const float*)w.data, NULL, &dx));
  }
  if (kx) mxGPUDestroyGPUArray(kx);
  if (kp) mxGPUDestroyGPUArray(kp);
  if (kw) mxGPUDestroyGPUArray(kw);
}
