/* xemo.h -- C ABI of libxemo.so: the B200-native (sm_100a) replacement for the MatConvNet /
 * mcnExtraLayers operator back-end that albanie/mcnCrossModalEmotions reaches through
 * dagnn.DagNN.eval and cnn_train_dag.
 *
 * The reference repository contains no native code; the interface this ABI replaces is the MEX
 * gateway of the un-vendored upstream operators (vl_nnconv, vl_nnbnorm, vl_nnpool, vl_nnrelu,
 * vl_nnsigmoid, vl_nnsoftmaxceloss, ...), reached from the reference at
 *   emoVoxCeleb/fetch_emovoxceleb_imdb.m:129      dag.eval (teacher forward)
 *   external/compute_visual_feats.m:90            dag.eval (teacher forward)
 *   external/compute_audio_feats.m:126            dag.eval (student forward)
 *   emoVoxCeleb/run_distillation.m:170-182        cnn_train_dag (student forward/backward/update)
 *   emoVoxCeleb/emoVoxZoo.m:151-157               dagnn.SoftmaxCELoss(T=2, logitTargets)
 *
 * Two layers of entry points:
 *   (A) xemo_vl_*   MatConvNet-boundary operators.  Arrays are `single`, H x W x C x N, column-major
 *                   (MATLAB layout), in host OR device memory (detected per pointer); outputs are
 *                   caller-allocated and live where the caller put them.  Forward when `dzdy` is
 *                   NULL, backward otherwise -- the MATLAB calling convention.  These are what the
 *                   MEX shims in mex/ forward to.
 *   (B) xemo_op_*   device-native building blocks on NHWC fp16 activations (fp32 where stated) that
 *                   the host-side graph compiler (mcncrossmodalemotions_b200/dagnn.py) strings
 *                   together into fused teacher / student / distillation-step programs, captured
 *                   as CUDA graphs (xemo_capture_*).
 *
 * Conventions: every function returns 0 on success or an xemo_status error code; the message is
 * available from xemo_last_error().  No exceptions cross the ABI.  A context is bound to one device
 * and one stream and is not thread-safe; all work is asynchronous on that stream unless the call
 * returns data to host memory.  The caller owns every buffer it passes in.  There is no CPU
 * fallback: without an sm_100 device xemo_create fails.
 */
#ifndef XEMO_H_
#define XEMO_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct xemo_ctx xemo_ctx;
typedef struct xemo_graph xemo_graph;

typedef enum {
  XEMO_OK = 0,
  XEMO_ERR_INVALID = 1,     /* bad argument / unsupported shape */
  XEMO_ERR_CUDA = 2,        /* CUDA runtime or driver error */
  XEMO_ERR_NO_DEVICE = 3,   /* no sm_100 device / TMA entry points unavailable */
  XEMO_ERR_NOMEM = 4
} xemo_status;

/* single-precision MATLAB array: dims H x W x C x N, column-major (index h + H*(w + W*(c + C*n))) */
typedef struct {
  void* data;
  int64_t h, w, c, n;
} xemo_array;

/* ------------------------------------------------------------------ context */
int xemo_version(void);
/* `cuda_stream` is a cudaStream_t (NULL: the context creates its own non-blocking stream).  A host whose other GPU work
 * runs on the legacy default stream (MATLAB's gpuArray arithmetic) passes XEMO_STREAM_LEGACY (== cudaStreamLegacy) so that
 * the library's kernels are stream-ordered with it; xemo_current_device reports the CUDA runtime's current device. */
#define XEMO_STREAM_LEGACY ((void*)0x1)
int xemo_current_device(int* device);
int xemo_create(int device, void* cuda_stream, xemo_ctx** out);
void xemo_destroy(xemo_ctx* ctx);
const char* xemo_last_error(xemo_ctx* ctx);
int xemo_sync(xemo_ctx* ctx);
/* The boundary operators (section B) take their scratch from a stream-ordered pool owned by the context and keep it
 * between calls; xemo_trim synchronises and hands the unused part back to the driver. */
int xemo_trim(xemo_ctx* ctx);
int xemo_num_sms(xemo_ctx* ctx);
/* Operand precision of xemo_vl_nnconv (the reference computes in `single` end to end: cnn_train_dag at
 * emoVoxCeleb/run_distillation.m:170-182 on gpuArray(single) batches, getBatchEmoVoxCeleb.m:197).
 *   XEMO_CONV_F16   (default) fp16 operands, fp32 accumulation: ~2^-11 relative operand rounding
 *   XEMO_CONV_F32X3 split operands: each single-precision operand is staged as hi + lo fp16 halves and every product
 *                   as the three leading terms hi*hi + lo*hi + hi*lo on tcgen05 (one convolution over a 3x longer
 *                   reduction), fp32 accumulation -- fp32-equivalent results (~2^-22 per product) at 3x the MMA work. */
enum { XEMO_CONV_F16 = 0, XEMO_CONV_F32X3 = 1 };
/* Determinism option of the training step (also XEMO_DETERMINISTIC=1 at xemo_create): the filter-gradient kernel does
 * not split its pixel reduction (every dF element has one producer instead of several `red.global.add` contributors),
 * the bias column sums use one slab, the loss sums its warps in a fixed order.  The BatchNorm reductions add fp32 block
 * partials with fp64 atomics (exact for these magnitudes, hence order-free).  Two runs of a step then produce bit-identical
 * gradients; costs idle SMs on layers with few filter-gradient tiles. */
int xemo_set_deterministic(xemo_ctx* ctx, int on);
int xemo_get_deterministic(xemo_ctx* ctx);
int xemo_set_conv_precision(xemo_ctx* ctx, int mode);
int xemo_get_conv_precision(xemo_ctx* ctx);
/* number of kernels this context has launched (graph replays count their kernel nodes) */
uint64_t xemo_launch_count(xemo_ctx* ctx);
/* async copies on the context stream (host memory should be pinned for true asynchrony) */
int xemo_h2d(xemo_ctx* ctx, void* dst_dev, const void* src_host, size_t bytes);
int xemo_d2h(xemo_ctx* ctx, void* dst_host, const void* src_dev, size_t bytes);
int xemo_memset(xemo_ctx* ctx, void* dst_dev, int byte, size_t bytes);

/* Multi-stream programs.  xemo_set_stream redirects subsequent launches to `cuda_stream` (NULL = back to the
 * primary stream the context was created on).  xemo_stream_wait(waiter, signal) makes `waiter` wait for the work
 * enqueued so far on `signal` (NULL = primary); inside a capture this forks / joins the second stream. */
int xemo_set_stream(xemo_ctx* ctx, void* cuda_stream);
int xemo_stream_wait(xemo_ctx* ctx, void* waiter_stream, void* signal_stream);

/* CUDA-graph capture of a sequence of xemo_op_* calls (begins / ends on the primary stream) */
int xemo_capture_begin(xemo_ctx* ctx);
int xemo_capture_end(xemo_ctx* ctx, xemo_graph** out);
int xemo_graph_launch(xemo_ctx* ctx, xemo_graph* g);
int xemo_graph_num_kernels(xemo_graph* g);
void xemo_graph_destroy(xemo_graph* g);

/* ------------------------------------------------------------------ (A) MatConvNet-boundary operators */
/* output size rule shared by vl_nnconv / vl_nnpool: floor((H + pt + pb - FH)/sy) + 1 */
int xemo_out_size(int64_t h, int64_t w, int fh, int fw, const int pad[4], const int stride[2], int64_t* oh,
                  int64_t* ow);

/* Y = vl_nnconv(X, F, B, 'pad', pad, 'stride', stride)            (dzdy == NULL; y written)
 * [DX, DF, DB] = vl_nnconv(X, F, B, DZDY, ...)                    (dx / df / db written; each may be NULL)
 * F is FH x FW x FC x K with FC == C (no groups on this path); B is K x 1 or NULL. */
int xemo_vl_nnconv(xemo_ctx* ctx, const xemo_array* x, const xemo_array* f, const xemo_array* b,
                   const xemo_array* dzdy, const int pad[4], const int stride[2], xemo_array* y, xemo_array* dx,
                   xemo_array* df, xemo_array* db);

/* method: 0 = max (padding = -inf; backward to the first maximum of the memory-order scan), 1 = avg
 * (divides by the in-bounds window area).  `argmax` (optional, forward max only) receives the
 * window-local index dw*PH + dh as uint8, OH x OW x C x N column-major, same memory space as y. */
int xemo_vl_nnpool(xemo_ctx* ctx, const xemo_array* x, const int pool[2], const xemo_array* dzdy, const int pad[4],
                   const int stride[2], int method, xemo_array* y_or_dx, uint8_t* argmax);

/* [Y, MOMENTS] = vl_nnbnorm(X, G, B, 'epsilon', e [, 'moments', M])  /  [DX, DG, DB, MOMENTS] = ...(.., DZDY, ..)
 * g, b: C x 1; moments: C x 2 = [mu sigma] (sigma, not variance).  moments_in NULL = batch statistics. */
int xemo_vl_nnbnorm(xemo_ctx* ctx, const xemo_array* x, const float* g, const float* b, const xemo_array* dzdy,
                    float epsilon, const float* moments_in, xemo_array* y_or_dx, float* dg, float* db,
                    float* moments_out);

int xemo_vl_nnrelu(xemo_ctx* ctx, const xemo_array* x, const xemo_array* dzdy, float leak, xemo_array* y_or_dx);
int xemo_vl_nnsigmoid(xemo_ctx* ctx, const xemo_array* x, const xemo_array* dzdy, xemo_array* y_or_dx);
/* softmax along dim 3 (channels) of x / temperature */
int xemo_vl_nnsoftmaxt(xemo_ctx* ctx, const xemo_array* x, float temperature, xemo_array* y);
/* Y = vl_nnsoftmaxceloss(X, P, 'temperature', T, 'logitTargets', lt, 'instanceWeights', W): scalar *loss;
 * DX = vl_nnsoftmaxceloss(X, P, DZDY, ...): dx written.  x, p: 1 x 1 x C x N with C <= 16. */
int xemo_vl_nnsoftmaxceloss(xemo_ctx* ctx, const xemo_array* x, const xemo_array* p, const float* dzdy,
                            float temperature, int logit_targets, const float* instance_weights, float* loss,
                            xemo_array* dx);
/* vl_nnloss(X, c, [], 'loss', 'classerror'): labels are 1-based (float, as MATLAB passes them) */
int xemo_vl_nnloss_classerror(xemo_ctx* ctx, const xemo_array* x, const float* labels, float* nerr);
/* mcnExtraLayers global average pooling (SE squeeze): H x W x C x N -> 1 x 1 x C x N */
int xemo_vl_nnglobalpool(xemo_ctx* ctx, const xemo_array* x, const xemo_array* dzdy, xemo_array* y_or_dx);
/* mcnExtraLayers Axpy (SE excite + shortcut): out = a (.) x + y, a is 1 x 1 x C x N */
int xemo_vl_nnaxpy(xemo_ctx* ctx, const xemo_array* a, const xemo_array* x, const xemo_array* y, xemo_array* out);

/* ------------------------------------------------------------------ (B) device-native building blocks
 * Activations: NHWC fp16, channel count a multiple of 16 for convolution operands, 8 otherwise.
 * Filters: [Kout][R][S][Cin] fp16 ("KRSC").  All pointers are device pointers. */

/* layout conversion at graph edges */
int xemo_op_hwcn_to_nhwc(xemo_ctx* ctx, const float* src, int H, int W, int C, int N, void* dst, int Cp, int dst_f32);
int xemo_op_nhwc_to_hwcn(xemo_ctx* ctx, const void* src, int src_f32, int H, int W, int C, int N, int Cp, float* dst);
int xemo_op_filters_to_krsc(xemo_ctx* ctx, const float* f, int FH, int FW, int FC, int K, void* dst16, int Kp, int Cp,
                            int flip_transpose);
/* teacher stem staging: faces H x W x C x N (fp32 HWCN) -> [N][H][OW][32] fp16 row-im2col (7x7/2 -> 7x1, 32 ch) */
int xemo_op_face_rows_im2col(xemo_ctx* ctx, const float* faces, int H, int W, int C, int N, int S, int stride_w,
                             int pad_l, int OW, void* dst16);
/* the same staging fused with the reference's face preprocessing (emoVoxCeleb/fetch_emovoxceleb_imdb.m:175-193,
 * teacher/ferplus_baselines.m:203-213): uint8 grey IH x IW x N (column-major) -> 3 channels, single, minus
 * mean3[c] (device, fp32[3]), corner-aligned bilinear resize to OHt x OWt -> [N][OHt][OW][32] fp16 */
int xemo_op_face_u8_rows_im2col(xemo_ctx* ctx, const uint8_t* faces, int IH, int IW, int N, int OHt, int OWt,
                                const float* mean3, int S, int stride_w, int pad_l, int OW, void* dst16);
/* student stem staging: spectrograms H x W x 1 x N -> [N][HP][OW][16] fp16 space-to-depth (7x7/2 -> 4x1, 16 ch) */
int xemo_op_spec_s2d(xemo_ctx* ctx, const float* spec, int H, int W, int N, int pad_t, int pad_l, int HP, int OW,
                     void* dst16);

/* student front-end (VGGVox runSpec called at emoVoxCeleb/getBatchEmoVoxCeleb.m:162,
 * external/compute_audio_feats.m:176; constants emoVoxCeleb/run_distillation.m:109-117): wav [N][L] fp32 ->
 * pre-emphasis (alpha) -> W frames of Nw samples every Ns -> Hamming -> |FFT_nfft| -> spec nfft x W x 1 x N
 * column-major fp32; spec_rownorm = per-row (x - mean)/std over time, N-1 normalised (getBatchEmoVoxCeleb.m:164-169) */
int xemo_op_spectrogram(xemo_ctx* ctx, const float* wav, int N, int L, int Nw, int Ns, int nfft, float alpha, float scale,
                        int W, float* spec);
int xemo_op_spec_rownorm(xemo_ctx* ctx, float* spec, int H, int W, int N);

/* implicit-GEMM convolution on tcgen05: out = act(scale[k]*conv(x,w) + shift[k] + residual).
 * out16 (fp16) and/or out32 (fp32) receive [N*OH*OW][ldc]; scale/shift/residual may be NULL. */
int xemo_op_conv_fwd(xemo_ctx* ctx, const void* x16, int N, int H, int W, int Cin, const void* w16, int Kout, int R,
                     int S, int sh, int sw, int pt, int pb, int pl, int pr, const float* scale, const float* shift,
                     const void* residual16, int relu, void* out16, float* out32, int ldc);
/* data gradient: dx[N][H][W][Cin] from dy[N][OH][OW][Kout] and the packed filters produced by
 * xemo_op_pack_dgrad_filters (one flipped/transposed sub-filter per output parity class). */
size_t xemo_dgrad_pack_elems(int Cin, int Kout, int R, int S, int sh, int sw);
int xemo_op_pack_dgrad_filters(xemo_ctx* ctx, const void* w16_krsc, int Kout, int R, int S, int Cin, int sh, int sw,
                               int pt, int pl, void* packed16);
int xemo_op_conv_dgrad(xemo_ctx* ctx, const void* dy16, int N, int H, int W, int Cin, const void* packed16, int Kout,
                       int R, int S, int sh, int sw, int pt, int pb, int pl, int pr, void* dx16);
/* the same for a FULL-HEIGHT filter (R == H, S == 1, stride 1, no padding: one output row -- the student's fc6): the data
 * gradient as a plain GEMM over (n, w) rows, without the R - 1 zero taps per pixel of the general form.  `packed16` has the
 * size xemo_dgrad_pack_elems gives; Cin <= 256. */
int xemo_op_pack_dgrad_filters_fullheight(xemo_ctx* ctx, const void* w16_krsc, int Kout, int R, int Cin, void* packed16);
int xemo_op_conv_dgrad_fullheight(xemo_ctx* ctx, const void* dy16, int N, int H, int W, int Cin, const void* packed16, int Kout,
                                  void* dx16);
/* filter gradient, accumulated (+=) into dF[Kout][R][S][Cin] fp32, scaled by `scale` */
int xemo_op_conv_wgrad(xemo_ctx* ctx, const void* x16, int N, int H, int W, int Cin, const void* dy16, int ldy,
                       int Kout, int R, int S, int sh, int sw, int pt, int pb, int pl, int pr, float* dF, float scale);
/* host-side shape planning of the two tcgen05 kernels without a device (no tensor maps are encoded): the tile shape,
 * pipeline depth and work split chosen for a geometry, for CPU-side invariant checks.  out receives 12 ints --
 * conv : bk, block_n, num_m_tiles, num_n_tiles, num_stages, epi_bufs, b_resident, use_tma_store, smem, grid, epi_cw, k_iters
 * wgrad: chunk_a, chunk_b, block_c, c_tiles, T, mt, pix, groups, splits, num_stages, smem, grid */
int xemo_debug_conv_plan(int N, int H, int W, int Cin, int Kout, int R, int S, int sh, int sw, int pt, int pb, int pl, int pr,
                         int num_sms, int* out);
int xemo_debug_wgrad_plan(int N, int H, int W, int Cin, int ldy, int Kout, int R, int S, int sh, int sw, int pt, int pb, int pl,
                          int pr, int num_sms, int* out);
/* CTA pairs (tcgen05.mma.cta_group::2, clusters of two CTAs that split the filter tile) in the forward / data-gradient
 * convolution: -1 = follow XEMO_CONV_2CTA / the planner's rule (default), 0 = never, 1 = the rule, 2 = whenever legal.
 * xemo_debug_conv_plan reports the choice as out[12] when asked for 13 ints through xemo_debug_conv_plan2. */
int xemo_debug_set_conv_pair_mode(int mode);
int xemo_debug_conv_plan2(int N, int H, int W, int Cin, int Kout, int R, int S, int sh, int sw, int pt, int pb, int pl, int pr,
                          int num_sms, int* out13);
/* Host-side planners of two memory-bound kernel families (no device needed; `-m "not gpu"` tests).
 * xemo_debug_se_gate_plan: the SE gate's launch for N faces -- out[0] = cluster size K, out[1] = ranges of hidden units in
 *   the last phase's sum (a function of C and Cr only), out[2] = thread groups per channel, out[3] = dynamic shared
 *   memory in bytes, out[4] = grid size.  lin = 1: the form by linearity from mean_hw(t2) (Cm channels).
 * xemo_debug_fixed_channel_grid: grid of a grid-stride kernel whose threads keep one channel group (C8 = C / 8 groups,
 *   `threads` per block) when at most per_sm blocks per SM are resident. */
int xemo_debug_se_gate_plan(int N, int C, int Cm, int Cr, int lin, int num_sms, int* out5);
int xemo_debug_fixed_channel_grid(long long items, int C8, int threads, int num_sms, int per_sm);
/* bias gradient: out[c] = scale * sum_p dy[p][c] */
int xemo_op_colsum(xemo_ctx* ctx, const void* dy16, size_t P, int ld, int C, float scale, float* out);

/* pooling (fp16 NHWC).  maxpool_fwd optionally applies z = relu(a[c]*x + b[c]) on the fly (BN+ReLU folded into
 * the pooling read) and emits the uint8 window-local arg-max dw*PH + dh that maxpool_bwd consumes. */
int xemo_op_maxpool_fwd(xemo_ctx* ctx, const void* x16, int N, int H, int W, int C, int PH, int PW, int sh, int sw,
                        int pt, int pb, int pl, int pr, const float* a, const float* b, void* y16, uint8_t* argmax);
/* the same, also recording the raw (pre-affine) value of each window's winner in xwin16 (optional; 3x3 / 5x3 windows).
 * pooled_ld (0 = C): channel pitch of y16 / argmax / xwin16, so that the pooled tensor can carry zero-padded channels
 * for its consumer (the student's conv2 reads 128-byte TMA rows from a 96 -> 128 channel pitch); maxpool_bwd_ld is the
 * matching backward (3x3 / stride 2 / pad 0), reading dy16 / argmax with that pitch */
int xemo_op_maxpool_fwd_win(xemo_ctx* ctx, const void* x16, int N, int H, int W, int C, int PH, int PW, int sh, int sw,
                            int pt, int pb, int pl, int pr, const float* a, const float* b, void* y16, uint8_t* argmax,
                            void* xwin16, int pooled_ld);
int xemo_op_maxpool_bwd_ld(xemo_ctx* ctx, const void* dy16, const uint8_t* argmax, int N, int H, int W, int C, int PH,
                           int PW, int sh, int sw, int pt, int pb, int pl, int pr, void* dx16, int pooled_ld);
int xemo_op_maxpool_bwd(xemo_ctx* ctx, const void* dy16, const uint8_t* argmax, int N, int H, int W, int C, int PH,
                        int PW, int sh, int sw, int pt, int pb, int pl, int pr, void* dx16);
int xemo_op_avgpool_fwd(xemo_ctx* ctx, const void* x16, int N, int H, int W, int C, int PH, int PW, int sh, int sw,
                        int pt, int pb, int pl, int pr, void* y16);
int xemo_op_avgpool_bwd(xemo_ctx* ctx, const void* dy16, int N, int H, int W, int C, int PH, int PW, int sh, int sw,
                        int pt, int pb, int pl, int pr, void* dx16);

/* batch normalisation over P = N*H*W rows of C channels.
 * bn_train: batch statistics -> moments[2C] = [mu | sigma], a = g/sigma, b = beta - a*mu (ws: 2C doubles).
 * bn_test : a, b from given moments; conv_bias (optional) is folded into b (b += a*conv_bias) so that the BN of a
 *           biased convolution can run as that convolution's scale/shift epilogue.  affine_act: y = a*x + b (+ReLU).
 * bn_bwd  : dz = dy*[a*x+b > 0] (relu_mask); dg = sum dz*xhat, db = sum dz (scaled by inv_grad_scale, fp32);
 *           dx = a*(dz - db/P - xhat*dg/P)  (train)  or  a*dz  (test_mode).  dconv_bias (optional, train mode,
 *           C <= 8192) receives inv_grad_scale * sum_rows dx: the bias gradient of the convolution feeding the BN.
 * bn_bwd_pool: the same for a BN+ReLU followed by a max pool whose windows overlap at most 2 x 2: dz is gathered
 *           from the POOLED gradient dpool16 [N][OH][OW][C] through the recorded arg-max, so that the
 *           full-resolution gradient is never materialised between the pooling and the BN backward. */
int xemo_op_bn_train(xemo_ctx* ctx, const void* x16, size_t P, int C, const float* g, const float* beta, float eps,
                     double* ws, float* moments, float* a, float* b);
int xemo_op_bn_test(xemo_ctx* ctx, const float* moments, int C, const float* g, const float* beta, const float* conv_bias,
                    float* a, float* b);
int xemo_op_affine_act(xemo_ctx* ctx, const void* x16, size_t P, int C, const float* a, const float* b, int relu,
                       void* y16);
int xemo_op_bn_bwd(xemo_ctx* ctx, const void* x16, const void* dy16, size_t P, int C, const float* moments,
                   const float* a, const float* b, int relu_mask, int test_mode, double* ws, void* dx16, float* dg,
                   float* db, float* dconv_bias, float inv_grad_scale);
int xemo_op_bn_bwd_pool(xemo_ctx* ctx, const void* x16, const void* dpool16, const uint8_t* argmax, int N, int H, int W,
                        int C, int PH, int PW, int sh, int sw, int pt, int pb, int pl, int pr, const float* moments,
                        const float* a, const float* b, double* ws, void* dx16, float* dg, float* db, float* dconv_bias,
                        float inv_grad_scale);
int xemo_op_relu_bwd(xemo_ctx* ctx, const void* y16, const void* dy16, size_t n, void* dx16);
int xemo_op_add_act(xemo_ctx* ctx, const void* a16, const void* b16, size_t n, int relu, void* y16);

/* student stem (conv1 of the VGGVox graph built at emoVoxCeleb/emoVoxZoo.m:50-62 on the 512 x W x 1 batch of
 * getBatchEmoVoxCeleb.m:197, followed by train-mode BN + ReLU + 3x3/2 max pool under cnn_train_dag,
 * run_distillation.m:170).  With one input channel the layer is linear in the 64-entry space-to-depth patch, so
 * the BN statistics and the BN / filter gradients follow from the patch autocorrelation R (64 x 64) and patch sum S
 * instead of passes over the 1.85 GB conv1 activation (csrc/stem_kernels.cuh).
 * stem_autocorr      : s2d16 [N][HP][OW][16] (xemo_op_spec_s2d), HP == OH + 3 -> ws (xemo_stem_ws_doubles() doubles):
 *                      row-pair products then the assembled [R | S]
 * stem_bn_train      : as bn_train for x = conv(s2d, w16 [C][64]) + bias, from ws
 * stem_pool_bn_reduce: xwin16 / dpool16 [P][ld] (ld = 0: C) at the POOLED resolution (xemo_op_maxpool_fwd_win); masks dpool16
 *                      in place with [a*xwin+b > 0] and accumulates acc[2C] = {sum dz, sum dz*xhat} (doubles)
 * stem_wgrad_finalize: dW [C][64] holds G1 = inv_grad_scale * sum_p dz[p,.] patch[p] (xemo_op_conv_wgrad on the dz
 *                      that xemo_op_maxpool_bwd scatters from the masked dpool16) and is overwritten with the
 *                      filter gradient A*G1 - D*(R w + bias*S) + E*S; dgamma/dbeta = inv_grad_scale*acc; dbias = 0.
 *                      g1_pair (optional): G1 in pixel-pair form [2C][4][32] instead (wgrad on the 32-channel view of
 *                      s2d16 and the [P/2][2C] view of dz -- fewer TMA rows per pixel); its diagonal blocks are summed
 * stem_pair_filter   : w16 [C][4][1][16] -> w2 [2C][4][1][32], w2[(e,k)][j][e'*16+c] = [e==e'] w[k][j][c]: the filter
 *                      of the pixel-pair form (s2d viewed as [N][HP][OW/2][32], output as [N][OH][OW/2][2C]; OW even)
 * tile_f32           : dst[r*C + c] = src[c] (or `fill` when src is NULL), r < reps */
size_t xemo_stem_ws_doubles(void);
int xemo_op_stem_pair_filter(xemo_ctx* ctx, const void* w16, int C, void* w2_16);
int xemo_op_tile_f32(xemo_ctx* ctx, const float* src, int C, int reps, float fill, float* dst);
int xemo_op_stem_autocorr(xemo_ctx* ctx, const void* s2d16, int N, int HP, int OW, int OH, double* ws);
int xemo_op_stem_bn_train(xemo_ctx* ctx, const double* ws, const void* w16, const float* bias, size_t P, int C,
                          const float* g, const float* beta, float eps, float* moments, float* a, float* b);
int xemo_op_stem_pool_bn_reduce(xemo_ctx* ctx, const void* xwin16, void* dpool16, size_t P, int C, int ld,
                                const float* moments, const float* a, const float* b, double* acc);
int xemo_op_stem_wgrad_finalize(xemo_ctx* ctx, const double* ws, const void* w16, const float* bias, const double* acc,
                                size_t P, int C, const float* moments, const float* a, float inv_grad_scale, float* dW,
                                float* dbias, float* dgamma, float* dbeta, const float* g1_pair);

/* squeeze-and-excitation: s = mean_hw(u) ; gate = sigmoid(W2 relu(W1 s + b1) + b2) ; y = relu(gate*u + shortcut).
 * w1 is [Cr][C]; w2 is passed TRANSPOSED, [Cr][C] (w2t[j][c] = W2[c][j]), so that the gate kernel reads it coalesced. */
int xemo_op_se_squeeze(xemo_ctx* ctx, const void* u16, int N, int HW, int C, float* s);
int xemo_op_se_gate(xemo_ctx* ctx, const float* s, int N, int C, int Cr, const float* w1, const float* b1,
                    const float* w2, const float* b2, float* gate);
/* The SE block by linearity (measured in round 2: the default on the teacher's 56 x 56 / 28 x 28 stages).  The
 * squeeze is linear in the bottleneck's 3x3 output t2, s = a3 * (W3 . mean_hw t2) + b3, so the gate is known before the
 * expand convolution runs and the excite folds into that convolution's epilogue: m2 = se_squeeze(t2) ([N][Cm]) ->
 * se_gate_lin -> nc_scale = gate*a3, nc_shift = gate*b3 ([N][C]) -> conv_fwd_nc: out = act(nc_scale[n,k]*conv + nc_shift[n,k]
 * + residual).  The expand output u is never written or re-read. */
int xemo_op_se_gate_lin(xemo_ctx* ctx, const float* m2, int N, int C, int Cm, int Cr, const void* w3_16, const float* a3,
                        const float* b3, const float* w1, const float* b1, const float* w2t, const float* b2,
                        float* nc_scale, float* nc_shift);
int xemo_op_conv_fwd_nc(xemo_ctx* ctx, const void* x16, int N, int H, int W, int Cin, const void* w16, int Kout, int R, int S,
                        int sh, int sw, int pt, int pb, int pl, int pr, const float* nc_scale, const float* nc_shift,
                        const void* residual16, int relu, void* out16);
int xemo_op_se_excite(xemo_ctx* ctx, const void* u16, const float* gate, const void* shortcut16, int N, int HW, int C,
                      int relu, void* y16);

/* teacher -> student coupling (emoVoxCeleb/getBatchEmoVoxCeleb.m:133-159,179-188): per clip n aggregate frame
 * logits rows [start[n], end[n]) with max (use_mean = 0) or mean into target[N][num_pred] */
int xemo_op_logit_aggregate(xemo_ctx* ctx, const float* frame_logits, int ldl, const int* start, const int* end, int N,
                            int num_pred, int use_mean, float* target);
/* fused T-softmax CE + backward + metric layers.  x16: [N][ldx] fp16 student logits; t: [N][ldt] fp32 teacher
 * logits (or distributions); dx16 (optional) = grad_scale*dzdy*w_n*(q-p)/T; scalars[0] += loss, scalars[1] +=
 * classerror; class_stats[0..C) += correct per class, [C..2C) += count per class; max_label[N] (1-based). */
int xemo_op_softmaxce(xemo_ctx* ctx, const void* x16, int ldx, const float* t, int ldt, const float* w, int N, int C,
                      float T, int logit_targets, float dzdy, float grad_scale, void* dx16, float* scalars,
                      float* class_stats, int* max_label);
/* the same kernel behind every `lossType` of emoVoxCeleb/emoVoxZoo.m:137-157 that compares {prediction, logitTarget}:
 * loss_type 0 = softmax CE as above (logit_targets = 0, T = 1: dagnn.Loss('softmaxlog') on one-hot rows),
 * 1 = dagnn.EuclideanLoss: loss = sum_n w_n/2 |x_n - t_n|^2, dx = dzdy w_n (x - t)          (emoVoxZoo.m:138-144)
 * 2 = dagnn.HuberLoss('sigma', T): smooth-L1, linear where |x - t| > 1/sigma^2              (emoVoxZoo.m:145-147)
 * x / dx are fp32 when x_f32 / dx_f32 are set (fp16 otherwise), with row pitches ldx / lddx; metrics as above. */
int xemo_op_loss(xemo_ctx* ctx, const void* x, int x_f32, int ldx, const float* t, int ldt, const float* w, int N, int C,
                 int loss_type, float T, int logit_targets, float dzdy, float grad_scale, void* dx, int dx_f32, int lddx,
                 float* scalars, float* class_stats, int* max_label);
/* cnn_train_dag update: m <- mu*m - (wd*w + g*inv_grad_scale/B); w <- w + lr*m; optional fp16 copy refresh.
 * hyper (device, fp32[4]) = {lr, momentum, weight_decay, 1/B}; lr_mult / wd_mult are per-parameter multipliers */
int xemo_op_sgd_momentum(xemo_ctx* ctx, float* w, float* m, const float* g, size_t n, const float* hyper, float lr_mult,
                         float wd_mult, float inv_grad_scale, void* w16);
/* dagnn.BatchNorm moments parameter: moments <- (1-rate)*moments + rate*batch_moments */
int xemo_op_moments_average(xemo_ctx* ctx, float* moments, const float* batch_moments, int n, float rate);
/* Overflow guard of the fp16 gradient chain (the fast programs carry activation gradients in fp16 under a fixed loss
 * scale; the reference trains in single and cannot overflow there).  grad_guard scans the flat fp32 gradient: state
 * (device, int[3], zero-initialised by the caller) gets state[0] = 1 if any element is inf / NaN this step (else 0) and
 * state[1] += state[0] (steps skipped so far).  The *_guarded updates are no-ops while state[0] is set, so a poisoned
 * gradient never reaches the master weights, the momentum, the fp16 mirror or the BN moments. */
int xemo_op_grad_guard(xemo_ctx* ctx, const float* g, size_t n, int* state);
int xemo_op_sgd_momentum_guarded(xemo_ctx* ctx, float* w, float* m, const float* g, size_t n, const float* hyper,
                                 float lr_mult, float wd_mult, float inv_grad_scale, void* w16, const int* guard);
/* bm_scale: 1 / ranks when batch_moments holds the sum over data-parallel ranks (1 otherwise) */
int xemo_op_moments_average_guarded(xemo_ctx* ctx, float* moments, const float* batch_moments, int n, float rate,
                                    float bm_scale, const int* guard);
int xemo_op_cast_f32_f16(xemo_ctx* ctx, const float* src, size_t n, void* dst16);
int xemo_op_cast_f16_f32(xemo_ctx* ctx, const void* src16, size_t n, float scale, float* dst);
/* dst[o*outer_stride + inner_off + i] = value for o < outer, i < inner (masks structurally-zero filter slots) */
int xemo_op_fill_strided_f32(xemo_ctx* ctx, float* dst, int outer, size_t outer_stride, size_t inner_off, int inner,
                             float value);

/* ------------------------------------------------------------------ (C) graph-level entry points
 * Whole networks behind the ABI: what dagnn.DagNN.eval (emoVoxCeleb/fetch_emovoxceleb_imdb.m:129,
 * external/compute_visual_feats.m:90, external/compute_audio_feats.m:126) and cnn_train_dag
 * (emoVoxCeleb/run_distillation.m:170-182: forward, loss emoVoxZoo.m:137-157, backward, accumulateGradients, gradient
 * sum over the labs of 'gpus', opts.gpus) do, as single calls.  The library owns the device-resident network, sequences
 * the kernels, captures the sequence in CUDA graphs and replays them; the host (MATLAB through mex/xemo_dagnn_mex.c,
 * Python through net.py) moves parameters in, inputs in and logits / metrics out.
 *
 * Parameters are named as in the zoo (`conv1f`, `conv1b`, `bn1m`, `bn1b`, `bn1x`, `s2b1_c1f`, ..., `classifierf`) and
 * travel in MatConvNet layouts: filters FH x FW x FC x K column-major, vectors K x 1, BatchNorm moments C x 2 = [mu sigma]
 * column-major.  xemo_net_param_name / _dims enumerate them in graph order. */
typedef struct xemo_net xemo_net;
typedef struct xemo_comm xemo_comm;
enum { XEMO_NET_RESNET50 = 0, XEMO_NET_SENET50 = 1, XEMO_NET_VGGVOX = 2 };
enum { XEMO_INPUT_F32 = 0, XEMO_INPUT_U8 = 1 };
enum { XEMO_LOSS_SOFTMAXCE = 0, XEMO_LOSS_SOFTMAXLOG = 1, XEMO_LOSS_EUCLIDEAN = 2, XEMO_LOSS_HUBER = 3 };
enum { XEMO_TENSOR_PARAM = 0, XEMO_TENSOR_GRAD = 1, XEMO_TENSOR_MOMENTUM = 2 };

/* kind XEMO_NET_RESNET50 / _SENET50 (teachers, test-mode BN folded at finalize): `size` is the face size for
 * XEMO_INPUT_U8 (uint8 grey size x size x N column-major; normalizeFace + bilinear resize to 224 run on the device,
 * emoVoxCeleb/fetch_emovoxceleb_imdb.m:175-193) and ignored for XEMO_INPUT_F32 (224 x 224 x 3 x N single, normalised).
 * kind XEMO_NET_VGGVOX (student): `size` is the spectrogram width W (input 512 x W x 1 x N single, row-normalised);
 * W must reduce to a 1 x 1 output (the width buckets 100 ... 1000 of emoVoxZoo.m:258-259 do). */
int xemo_net_create(xemo_ctx* ctx, int kind, int batch, int size, int input_mode, int num_outputs, xemo_net** out);
void xemo_net_destroy(xemo_net* net);
int xemo_net_batch(xemo_net* net);
int xemo_net_num_params(xemo_net* net);
const char* xemo_net_param_name(xemo_net* net, int index);
int xemo_net_param_dims(xemo_net* net, const char* name, int64_t dims[4]);
/* host pointers; every parameter must be set before xemo_net_finalize, which builds the device state.  A finalized
 * student accepts further set_param calls (checkpoint restore); a teacher's parameters are folded at finalize. */
int xemo_net_set_param(xemo_net* net, const char* name, const float* data, size_t numel);
int xemo_net_finalize(xemo_net* net);
/* student: current value / gradient of the last step (BN `x` entries: the batch moments) / momentum, MatConvNet layout */
int xemo_net_get_tensor(xemo_net* net, int which, const char* name, float* out, size_t numel);
int xemo_net_set_momentum(xemo_net* net, const char* name, const float* data, size_t numel);
/* inputs (host or device memory, stream-ordered copies): faces / spectrograms; student targets logitTarget
 * (1 x 1 x K x N, i.e. [N][K]; one-hot rows for XEMO_LOSS_SOFTMAXLOG) and instanceWeights (N), either may be NULL */
int xemo_net_input_bytes(xemo_net* net, size_t* bytes);
int xemo_net_set_input(xemo_net* net, const void* data, size_t bytes);
int xemo_net_set_target(xemo_net* net, const float* target, const float* weights);
/* loss of the training step (emoVoxZoo.m:137-157), its temperature (hot-cross-ent; huber: sigma is 1) and the loss scale
 * of the fp16 gradient chain */
int xemo_net_set_loss(xemo_net* net, int loss_type, float temperature, float grad_scale);
int xemo_net_set_hyper(xemo_net* net, float lr, float momentum, float weight_decay, int batch_size);
/* named device buffers for zero-copy hosts ("faces", "spec", "target", "logits", "pred32", "grad", ...) */
void* xemo_net_buffer(xemo_net* net, const char* name);
size_t xemo_net_grad_elems(xemo_net* net);
int xemo_net_num_kernels(xemo_net* net);

/* dag.eval: logits_out / pred_out (optional, host or device) receive 1 x 1 x K x N */
int xemo_teacher_forward(xemo_net* net, float* logits_out);
int xemo_student_forward(xemo_net* net, int train_mode, float* pred_out);
/* one cnn_train_dag iteration up to the gradients: forward (train-mode BN) + loss + backward; with a communicator of
 * more than one rank the flat gradient is summed across ranks inside the captured sequence (fc6..fc8 bucket on a forked
 * stream while conv5..conv1 are differentiated).  xemo_sgd_step: accumulateGradients (guarded against non-finite values). */
int xemo_student_train_step(xemo_net* net, xemo_comm* comm);
int xemo_sgd_step(xemo_net* net, float lr, float momentum, float weight_decay, int batch_size);
int xemo_allreduce_grads(xemo_net* net, xemo_comm* comm);
/* teacher -> student coupling on the device; start / end: device int[N_student] half-open row ranges into the teacher's
 * frame logits, or NULL / NULL for the windows stored by xemo_distill_set_windows (host arrays) */
int xemo_distill_set_windows(xemo_net* student, const int* start, const int* end);
int xemo_distill_couple(xemo_net* teacher, xemo_net* student, const int* start, const int* end, int use_mean);
/* the full distillation step (BASELINE.json's headline configuration) as ONE graph replay */
int xemo_distill_step(xemo_net* teacher, xemo_net* student, xemo_comm* comm, const int* start, const int* end, int use_mean,
                      float lr, float momentum, float weight_decay, int batch_size);
/* concurrency inside the captured student / distillation step: 1 = the teacher forward runs beside the student forward and
 * the filter gradients beside the data-gradient chain (forked streams inside the graph); 0 = one stream; -1 (default) =
 * on when the per-GPU batch is <= 64, where single kernels leave most SMs idle */
int xemo_net_set_overlap(xemo_net* student, int mode);
int xemo_net_reset_metrics(xemo_net* net);
/* out[0] objective, out[1] classerror of the last step; out[2..2+K) correct and out[2+K..2+2K) count per class since the
 * reset (ErrorStats); out[2+2K] non-finite gradient in the last update, out[3+2K] updates skipped so far */
int xemo_net_metrics(xemo_net* net, float* out, int n_out);

/* data-parallel communicator: NCCL resolved at run time (dlsym; libnccl.so.2), one rank per context.  Rank 0 creates the
 * 128-byte unique id and the host distributes it (MATLAB: labBroadcast; Python: torch.distributed / a file). */
int xemo_comm_unique_id(void* id128);
int xemo_comm_create(xemo_ctx* ctx, const void* id128, int rank, int world, xemo_comm** out);
/* Destroy every network whose captured step used the communicator FIRST: the captured graphs hold NCCL nodes, and
 * ncclCommDestroy waits until those graphs are gone. */
void xemo_comm_destroy(xemo_comm* comm);
int xemo_comm_allreduce_f32(xemo_comm* comm, float* device_buf, size_t n);

#ifdef __cplusplus
}
#endif
#endif /* XEMO_H_ */
