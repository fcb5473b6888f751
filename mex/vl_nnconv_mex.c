/* vl_nnconv MEX gateway over libxemo.so -- replaces MatConvNet's matlab/src/vl_nnconv.cu gateway.
 *   Y = vl_nnconv(X, F, B, 'pad', P, 'stride', S)
 *   [DX, DF, DB] = vl_nnconv(X, F, B, DZDY, 'pad', P, 'stride', S)
 * Reached from dagnn.Conv.forward / backward inside dag.eval (emoVoxCeleb/fetch_emovoxceleb_imdb.m:129,
 * external/compute_visual_feats.m:90, external/compute_audio_feats.m:126, cnn_train_dag at
 * emoVoxCeleb/run_distillation.m:170).  Source-only in this repository: see xemo_mex_common.h. */
#include "xemo_mex_common.h"

void mexFunction(int nlhs, mxArray* plhs[], int nrhs, const mxArray* prhs[]) {
  int pad[4] = {0, 0, 0, 0}, stride[2] = {1, 1}, next = 3, backward = 0, i, on_gpu;
  const mxGPUArray *kx, *kf, *kb, *kd = NULL;
  xemo_array x, f, b, dzdy, y, dx, df, db;
  int64_t oh, ow;
  xemo_ctx* ctx = xm_ctx();
  if (nrhs < 3) mexErrMsgIdAndTxt("xemo:args", "vl_nnconv needs X, F and B");
  x = xm_in(prhs[0], &kx);
  f = xm_in(prhs[1], &kf);
  b = xm_in(prhs[2], &kb);
  memset(&dzdy, 0, sizeof(dzdy));
  if (nrhs > 3 && !mxIsChar(prhs[3])) { dzdy = xm_in(prhs[3], &kd); backward = dzdy.data != NULL; next = 4; }
  for (i = next; i + 1 < nrhs; i += 2) {
    char* name = mxArrayToString(prhs[i]);
    if (!strcmp(name, "pad") || !strcmp(name, "Pad")) xm_ints(prhs[i + 1], pad, 4);
    else if (!strcmp(name, "stride") || !strcmp(name, "Stride")) xm_ints(prhs[i + 1], stride, 2);
    else if (!strcmp(name, "dilate") || !strcmp(name, "Dilate")) { int d[2]; xm_ints(prhs[i + 1], d, 2); if (d[0] != 1 || d[1] != 1) mexErrMsgIdAndTxt("xemo:opt", "dilation is not supported on this path"); }
    /* cuDNN / verbosity switches of the upstream gateway are accepted and ignored */
    mxFree(name);
  }
  on_gpu = mxIsGPUArray(prhs[0]);
  if (!backward) {
    if (xemo_out_size(x.h, x.w, (int)f.h, (int)f.w, pad, stride, &oh, &ow)) mexErrMsgIdAndTxt("xemo:size", "filters larger than the padded input");
    y = xm_out(&plhs[0], on_gpu, oh, ow, f.n, x.n);
    xm_check(xemo_vl_nnconv(ctx, &x, &f, b.data ? &b : NULL, NULL, pad, stride, &y, NULL, NULL, NULL));
  } else {
    mxArray* tmp[3];
    dx = xm_out(&tmp[0], on_gpu, x.h, x.w, x.c, x.n);
    df = xm_out(&tmp[1], on_gpu, f.h, f.w, f.c, f.n);
    memset(&db, 0, sizeof(db));
    if (b.data) db = xm_out(&tmp[2], on_gpu, f.n, 1, 1, 1); else tmp[2] = mxCreateNumericArray(0, NULL, mxSINGLE_CLASS, mxREAL);
    xm_check(xemo_vl_nnconv(ctx, &x, &f, b.data ? &b : NULL, &dzdy, pad, stride, NULL, &dx, &df, b.data ? &db : NULL));
    for (i = 0; i < 3 && (i == 0 || i < nlhs); ++i) plhs[i] = tmp[i];
  }
  if (kx) mxGPUDestroyGPUArray(kx);
  if (kf) mxGPUDestroyGPUArray(kf);
  if (kb) mxGPUDestroyGPUArray(kb);
  if (kd) mxGPUDestroyGPUArray(kd);
}
