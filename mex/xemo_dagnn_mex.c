/* xemo_dagnn MEX gateway over the graph-level entry points of libxemo.so (include/xemo.h section C): whole networks behind
 * one MEX function, so that a MATLAB host reaches the fast path -- device-resident network, kernels sequenced and
 * captured by the library -- instead of one MEX call (and one layout conversion) per MatConvNet operator.
 *
 *   h = xemo_dagnn('create', kind, batch, size, inputMode, numOutputs)   kind 0 resnet50 / 1 senet50 / 2 vggvox
 *   xemo_dagnn('set_param', h, name, value)        value: single, MatConvNet layout (net.params(i).value as stored)
 *   xemo_dagnn('finalize', h)
 *   y = xemo_dagnn('teacher_forward', h, x)        dag.eval({'data', x}) ; gather(squeeze(dag.vars(end).value))'
 *                                                  (emoVoxCeleb/fetch_emovoxceleb_imdb.m:129-131, external/compute_visual_feats.m:90-92)
 *   y = xemo_dagnn('student_forward', h, x, train) dag.eval at external/compute_audio_feats.m:126 (train = 0)
 *   xemo_dagnn('train_step', h, x, logitTarget [, instanceWeights])      forward + loss + backward of one cnn_train_dag
 *                                                  iteration (emoVoxCeleb/run_distillation.m:170-182, emoVoxZoo.m:137-157)
 *   xemo_dagnn('sgd_step', h, lr, momentum, weightDecay, batchSize)      accumulateGradients
 *   m = xemo_dagnn('metrics', h)                   [objective classerror correct(1:K) count(1:K) nonfinite skipped]
 *   v = xemo_dagnn('get', h, which, name)          which 0 value / 1 gradient / 2 momentum, as a MatConvNet array
 *   xemo_dagnn('destroy', h)
 * x / y live where the caller's array lives (gpuArray in -> the copy is device-to-device on MATLAB's stream).
 * Source-only, like the other shims: compile-checked against mex_stub.h, exercised through ctypes (net.py) instead. */
#include "xemo_mex_common.h"

static xemo_net* xm_net(const mxArray* a) { return (xemo_net*)(size_t)mxGetScalar(a); }

static size_t xm_numel(const xemo_array* a) { return (size_t)(a->h * a->w * a->c * a->n); }

void mexFunction(int nlhs, mxArray* plhs[], int nrhs, const mxArray* prhs[]) {
  char* cmd;
  xemo_ctx* ctx = xm_ctx();
  (void)nlhs;
  if (nrhs < 1 || !mxIsChar(prhs[0])) mexErrMsgIdAndTxt("xemo:args", "xemo_dagnn(command, ...)");
  cmd = mxArrayToString(prhs[0]);
  if (!strcmp(cmd, "create")) {
    xemo_net* net = NULL;
    if (nrhs < 6) mexErrMsgIdAndTxt("xemo:args", "create needs kind, batch, size, inputMode, numOutputs");
    xm_check(xemo_net_create(ctx, (int)mxGetScalar(prhs[1]), (int)mxGetScalar(prhs[2]), (int)mxGetScalar(prhs[3]), (int)mxGetScalar(prhs[4]),
                             (int)mxGetScalar(prhs[5]), &net));
    plhs[0] = mxCreateDoubleScalar((double)(size_t)net);
  } else if (!strcmp(cmd, "set_param")) {
    const mxGPUArray* keep;
    char* name = mxArrayToString(prhs[2]);
    xemo_array v = xm_in(prhs[3], &keep);
    if (keep) mexErrMsgIdAndTxt("xemo:args", "parameters are handed over as CPU arrays (gather them first)");
    xm_check(xemo_net_set_param(xm_net(prhs[1]), name, (const float*)v.data, xm_numel(&v)));
    mxFree(name);
  } else if (!strcmp(cmd, "finalize")) {
    xm_check(xemo_net_finalize(xm_net(prhs[1])));
  } else if (!strcmp(cmd, "teacher_forward") || !strcmp(cmd, "student_forward")) {
    const mxGPUArray* keep;
    xemo_net* net = xm_net(prhs[1]);
    xemo_array x = xm_in_any(prhs[2], &keep), y;
    size_t bytes = 0;
    int64_t dims[4];
    const int teacher = cmd[0] == 't';
    xm_check(xemo_net_input_bytes(net, &bytes));
    xm_check(xemo_net_set_input(net, x.data, bytes));
    xm_check(xemo_net_param_dims(net, teacher ? "classifierb" : "fc8b", dims));        /* K */
    y = xm_out(&plhs[0], mxIsGPUArray(prhs[2]), 1, 1, dims[0], xemo_net_batch(net));
    if (teacher) xm_check(xemo_teacher_forward(net, (float*)y.data));
    else xm_check(xemo_student_forward(net, nrhs > 3 ? (int)mxGetScalar(prhs[3]) : 0, (float*)y.data));
    if (keep) mxGPUDestroyGPUArray(keep);
  } else if (!strcmp(cmd, "train_step")) {
    const mxGPUArray *kx, *kt, *kw = NULL;
    xemo_net* net = xm_net(prhs[1]);
    xemo_array x = xm_in(prhs[2], &kx), t = xm_in(prhs[3], &kt), w;
    size_t bytes = 0;
    memset(&w, 0, sizeof(w));
    if (nrhs > 4) w = xm_in(prhs[4], &kw);
    xm_check(xemo_net_input_bytes(net, &bytes));
    xm_check(xemo_net_set_input(net, x.data, bytes));
    xm_check(xemo_net_set_target(net, (const float*)t.data, (const float*)w.data));
    xm_check(xemo_student_train_step(net, NULL));
    if (!kx || !kt) xm_check(xemo_sync(ctx));     /* CPU inputs are copied asynchronously: keep them alive until done */
    if (kx) mxGPUDestroyGPUArray(kx);
    if (kt) mxGPUDestroyGPUArray(kt);
    if (kw) mxGPUDestroyGPUArray(kw);
  } else if (!strcmp(cmd, "sgd_step")) {
    if (nrhs < 6) mexErrMsgIdAndTxt("xemo:args", "sgd_step needs lr, momentum, weightDecay, batchSize");
    xm_check(xemo_sgd_step(xm_net(prhs[1]), (float)mxGetScalar(prhs[2]), (float)mxGetScalar(prhs[3]), (float)mxGetScalar(prhs[4]),
                           (int)mxGetScalar(prhs[5])));
  } else if (!strcmp(cmd, "metrics")) {
    xemo_net* net = xm_net(prhs[1]);
    int64_t dims[4];
    xemo_array m;
    xm_check(xemo_net_param_dims(net, "fc8b", dims));
    m = xm_out(&plhs[0], 0, 1, 4 + 2 * dims[0], 1, 1);
    xm_check(xemo_net_metrics(net, (float*)m.data, (int)(4 + 2 * dims[0])));
  } else if (!strcmp(cmd, "get")) {
    xemo_net* net = xm_net(prhs[1]);
    char* name = mxArrayToString(prhs[3]);
    int64_t d[4];
    xemo_array v;
    xm_check(xemo_net_param_dims(net, name, d));
    v = xm_out(&plhs[0], 0, d[0], d[1], d[2], d[3]);
    xm_check(xemo_net_get_tensor(net, (int)mxGetScalar(prhs[2]), name, (float*)v.data, xm_numel(&v)));
    mxFree(name);
  } else if (!strcmp(cmd, "destroy")) {
    xemo_net_destroy(xm_net(prhs[1]));
  } else {
    mexErrMsgIdAndTxt("xemo:args", "unknown command %s", cmd);
  }
  mxFree(cmd);
}
