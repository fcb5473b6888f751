/* vl_nnbnorm MEX gateway over libxemo.so -- replaces MatConvNet's matlab/src/vl_nnbnorm.cu gateway, the operator
 * dagnn.BatchNorm calls in train mode under cnn_train_dag (emoVoxCeleb/run_distillation.m:170) and in test mode
 * (dag.mode = 'test', emoVoxCeleb/fetch_emovoxceleb_imdb.m:107, external/compute_audio_feats.m:106).
 *   [Y, MOMENTS]          = vl_nnbnorm(X, G, B, 'epsilon', E, 'moments', M)
 *   [DX, DG, DB, MOMENTS] = vl_nnbnorm(X, G, B, DZDY, 'epsilon', E, 'moments', M)
 * G, B: C x 1; MOMENTS: C x 2 = [mu sigma]; function default epsilon = 1e-4 (dagnn.BatchNorm passes its own 1e-5).
 * Source-only: see xemo_mex_common.h. */
#include "xemo_mex_common.h"

void mexFunction(int nlhs, mxArray* plhs[], int nrhs, const mxArray* prhs[]) {
  const mxGPUArray *kx, *kg, *kb, *kd = NULL, *km = NULL;
  xemo_array x, g, b, dzdy, mom_in, out, dg, db, mom_out;
  float eps = 1e-4f;
  int next = 3, i, on_gpu;
  xemo_ctx* ctx = xm_ctx();
  if (nrhs < 3) mexErrMsgIdAndTxt("xemo:args", "vl_nnbnorm needs X, G and B");
  x = xm_in(prhs[0], &kx);
  g = xm_in(prhs[1], &kg);
  b = xm_in(prhs[2], &kb);
  memset(&dzdy, 0, sizeof(dzdy));
  memset(&mom_in, 0, sizeof(mom_in));
  if (nrhs > 3 && !mxIsChar(prhs[3])) { dzdy = xm_in(prhs[3], &kd); next = 4; }
  for (i = next; i + 1 < nrhs; i += 2) {
    char* name = mxArrayToString(prhs[i]);
    if (!strcmp(name, "epsilon")) eps = (float)mxGetScalar(prhs[i + 1]);
    else if (!strcmp(name, "moments")) mom_in = xm_in(prhs[i + 1], &km);
    mxFree(name);
  }
  if (g.h * g.w * g.c * g.n != x.c || b.h * b.w * b.c * b.n != x.c)
    mexErrMsgIdAndTxt("xemo:size", "G and B must have one element per channel of X");
  if (mom_in.data && mom_in.h * mom_in.w * mom_in.c * mom_in.n != 2 * x.c)
    mexErrMsgIdAndTxt("xemo:size", "MOMENTS must be C x 2");
  on_gpu = mxIsGPUArray(prhs[0]);
  out = xm_out(&plhs[0], on_gpu, x.h, x.w, x.c, x.n);
  if (!dzdy.data) {
    /* forward: MOMENTS is the optional second output */
    memset(&mom_out, 0, sizeof(mom_out));
    if (nlhs > 1) mom_out = xm_out(&plhs[1], on_gpu, x.c, 2, 1, 1);
    xm_check(xemo_vl_nnbnorm(ctx, &x, (const float*)g.data, (const float*)b.data, NULL, eps, (const float*)mom_in.data, &out,
                             NULL, NULL, (float*)mom_out.data));
  } else {
    mxArray* tmp[3] = {NULL, NULL, NULL};
    dg = xm_out(nlhs > 1 ? &plhs[1] : &tmp[0], on_gpu, x.c, 1, 1, 1);
    db = xm_out(nlhs > 2 ? &plhs[2] : &tmp[1], on_gpu, x.c, 1, 1, 1);
    mom_out = xm_out(nlhs > 3 ? &plhs[3] : &tmp[2], on_gpu, x.c, 2, 1, 1);
    xm_check(xemo_vl_nnbnorm(ctx, &x, (const float*)g.data, (const float*)b.data, &dzdy, eps, (const float*)mom_in.data, &out,
                             (float*)dg.data, (float*)db.data, (float*)mom_out.data));
  }
  if (kx) mxGPUDestroyGPUArray(kx);
  if (kg) mxGPUDestroyGPUArray(kg);
  if (kb) mxGPUDestroyGPUArray(kb);
  if (kd) mxGPUDestroyGPUArray(kd);
  if (km) mxGPUDestroyGPUArray(km);
}
