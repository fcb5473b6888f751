// xemo_ops.cu -- context management, CUDA-graph capture and the device-native building blocks
// (xemo_op_*) of libxemo.so.  See include/xemo.h for the contract of every entry point.
#include "xemo_internal.h"

#include "conv_launch.cuh"
#include "hbm_kernels_extra.cuh"
#include "se_gate.cuh"
#include "stem_kernels.cuh"

using namespace xemo;

// ================================================================================================
// context
extern "C" int xemo_version(void) { return 100; }

// (XEMO_STREAM_LEGACY in xemo.h is cudaStreamLegacy, 0x1, spelled without the CUDA headers; checked in xemo_create)
extern "C" int xemo_current_device(int* device) {
  if (!device) return XEMO_ERR_INVALID;
  return cudaGetDevice(device) == cudaSuccess ? XEMO_OK : XEMO_ERR_NO_DEVICE;
}

// grid-stride kernels: `waves` whole waves of the blocks that fit (XEMO_GRID_WAVES, default 1)
static int grid_waves() {
  static const int w = [] { const char* e = getenv("XEMO_GRID_WAVES"); const int v = e ? atoi(e) : 1; return v >= 1 && v <= 64 ? v : 1; }();
  return w;
}
template <typename K>
static int per_sm_of(K kernel, int threads = 256, size_t smem = 0) { return resident_blocks(kernel, threads, smem) * grid_waves(); }

extern "C" int xemo_create(int device, void* cuda_stream, xemo_ctx** out) {
  if (!out) return XEMO_ERR_INVALID;
  *out = nullptr;
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess || device < 0 || device >= count) return XEMO_ERR_NO_DEVICE;
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return XEMO_ERR_NO_DEVICE;
  if (prop.major != 10) return XEMO_ERR_NO_DEVICE;  // sm_100a only: there is no fallback path
  if (cudaSetDevice(device) != cudaSuccess) return XEMO_ERR_CUDA;
  if (!tma_api().ok) return XEMO_ERR_NO_DEVICE;
  xemo_ctx* ctx = new xemo_ctx();
  if (const char* e = getenv("XEMO_DETERMINISTIC")) ctx->deterministic = e[0] == '1';

  ctx->device = device;
  ctx->num_sms = prop.multiProcessorCount;
  if (cuda_stream) {
    if (cuda_stream == XEMO_STREAM_LEGACY) cuda_stream = reinterpret_cast<void*>(cudaStreamLegacy);
    ctx->stream = ctx->primary = reinterpret_cast<cudaStream_t>(cuda_stream);
  } else {
    if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) {
      delete ctx;
      return XEMO_ERR_CUDA;
    }
    ctx->own_stream = true;
    ctx->primary = ctx->stream;
  }
  *out = ctx;
  return XEMO_OK;
}

extern "C" void xemo_destroy(xemo_ctx* ctx) {
  if (!ctx) return;
  if (ctx->pool) {
    cudaStreamSynchronize(ctx->stream);
    cudaMemPoolDestroy(ctx->pool);
  }
  if (ctx->own_stream) cudaStreamDestroy(ctx->primary);
  delete ctx;
}

extern "C" const char* xemo_last_error(xemo_ctx* ctx) { return ctx ? ctx->err.c_str() : "null context"; }

extern "C" int xemo_trim(xemo_ctx* ctx) {
  XEMO_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  if (ctx->pool) XEMO_CUDA(ctx, cudaMemPoolTrimTo(ctx->pool, 0));
  return XEMO_OK;
}

extern "C" int xemo_sync(xemo_ctx* ctx) {
  XEMO_CUDA(ctx, cudaStreamSynchronize(ctx->primary));
  return XEMO_OK;
}
extern "C" int xemo_num_sms(xemo_ctx* ctx) { return ctx ? ctx->num_sms : 0; }
extern "C" int xemo_set_conv_precision(xemo_ctx* ctx, int mode) {
  XEMO_REQUIRE(ctx, ctx && (mode == XEMO_CONV_F16 || mode == XEMO_CONV_F32X3), "set_conv_precision: mode must be 0 (fp16 operands) or 1 (split fp16 x 3)");
  ctx->conv_precision = mode;
  return XEMO_OK;
}
extern "C" int xemo_get_conv_precision(xemo_ctx* ctx) { return ctx ? ctx->conv_precision : -1; }
extern "C" int xemo_set_deterministic(xemo_ctx* ctx, int on) {
  if (!ctx) return XEMO_ERR_INVALID;
  ctx->deterministic = on ? 1 : 0;
  return XEMO_OK;
}
extern "C" int xemo_get_deterministic(xemo_ctx* ctx) { return ctx ? ctx->deterministic : -1; }
extern "C" uint64_t xemo_launch_count(xemo_ctx* ctx) { return ctx ? ctx->launches : 0; }

extern "C" int xemo_h2d(xemo_ctx* ctx, void* dst, const void* src, size_t bytes) {
  XEMO_CUDA(ctx, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, ctx->stream));
  return XEMO_OK;
}
extern "C" int xemo_d2h(xemo_ctx* ctx, void* dst, const void* src, size_t bytes) {
  XEMO_CUDA(ctx, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, ctx->stream));
  return XEMO_OK;
}
extern "C" int xemo_memset(xemo_ctx* ctx, void* dst, int byte, size_t bytes) {
  XEMO_CUDA(ctx, cudaMemsetAsync(dst, byte, bytes, ctx->stream));
  return XEMO_OK;
}

// Multi-stream programs: launches go to the context's *current* stream; xemo_stream_wait makes `waiter` wait for
// everything enqueued so far on `signal` (event record + wait), which also pulls `waiter` into an ongoing
// capture of `signal` (fork) or joins it back.  NULL designates the context's primary stream.
extern "C" int xemo_set_stream(xemo_ctx* ctx, void* cuda_stream) {
  ctx->stream = cuda_stream ? reinterpret_cast<cudaStream_t>(cuda_stream) : ctx->primary;
  return XEMO_OK;
}
extern "C" int xemo_stream_wait(xemo_ctx* ctx, void* waiter_stream, void* signal_stream) {
  cudaStream_t w = waiter_stream ? reinterpret_cast<cudaStream_t>(waiter_stream) : ctx->primary;
  cudaStream_t s = signal_stream ? reinterpret_cast<cudaStream_t>(signal_stream) : ctx->primary;
  cudaEvent_t ev;
  XEMO_CUDA(ctx, cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
  XEMO_CUDA(ctx, cudaEventRecord(ev, s));
  XEMO_CUDA(ctx, cudaStreamWaitEvent(w, ev, 0));
  XEMO_CUDA(ctx, cudaEventDestroy(ev));  // released once the recorded work has completed
  return XEMO_OK;
}

// ================================================================================================
// CUDA-graph capture
extern "C" int xemo_capture_begin(xemo_ctx* ctx) {
  XEMO_REQUIRE(ctx, !ctx->capturing, "capture already in progress");
  XEMO_REQUIRE(ctx, ctx->stream == ctx->primary, "capture must begin on the primary stream");
  XEMO_CUDA(ctx, cudaStreamBeginCapture(ctx->primary, cudaStreamCaptureModeThreadLocal));
  ctx->capturing = true;
  ctx->capture_mark = ctx->launches;
  return XEMO_OK;
}

extern "C" int xemo_capture_end(xemo_ctx* ctx, xemo_graph** out) {
  XEMO_REQUIRE(ctx, ctx->capturing && out, "no capture in progress");
  ctx->capturing = false;
  cudaGraph_t graph = nullptr;
  ctx->stream = ctx->primary;
  XEMO_CUDA(ctx, cudaStreamEndCapture(ctx->primary, &graph));
  xemo_graph* g = new xemo_graph();
  g->graph = graph;
  g->num_kernels = int(ctx->launches - ctx->capture_mark);
  ctx->launches = ctx->capture_mark;  // captured launches did not execute
  cudaError_t e = cudaGraphInstantiate(&g->exec, graph, 0);
  if (e != cudaSuccess) {
    cudaGraphDestroy(graph);
    delete g;
    return fail(ctx, XEMO_ERR_CUDA, "cudaGraphInstantiate failed: %s", cudaGetErrorString(e));
  }
  *out = g;
  return XEMO_OK;
}

extern "C" int xemo_graph_launch(xemo_ctx* ctx, xemo_graph* g) {
  XEMO_REQUIRE(ctx, g && g->exec, "null graph");
  XEMO_CUDA(ctx, cudaGraphLaunch(g->exec, ctx->primary));
  ctx->launches += uint64_t(g->num_kernels);
  return XEMO_OK;
}
extern "C" int xemo_graph_num_kernels(xemo_graph* g) { return g ? g->num_kernels : 0; }
extern "C" void xemo_graph_destroy(xemo_graph* g) {
  if (!g) return;
  if (g->exec) cudaGraphExecDestroy(g->exec);
  if (g->graph) cudaGraphDestroy(g->graph);
  delete g;
}

// ================================================================================================
// layout conversion
extern "C" int xemo_op_hwcn_to_nhwc(xemo_ctx* ctx, const float* src, int H, int W, int C, int N, void* dst, int Cp,
                                    int dst_f32) {
  XEMO_REQUIRE(ctx, src && dst && Cp >= C && H > 0 && W > 0 && W <= 65535 && C > 0 && N > 0, "hwcn_to_nhwc: bad arguments");
  dim3 block(32, 8), grid((H + 31) / 32, (Cp + 31) / 32, 1);
  const int per = 65535 / W;  // gridDim.z <= 65535: whole images per launch
  for (int n0 = 0; n0 < N; n0 += per) {
    const int nn = N - n0 < per ? N - n0 : per;
    grid.z = unsigned(nn) * W;
    const float* s = src + size_t(n0) * H * W * C;
    if (dst_f32)
      hwcn_f32_to_nhwc_kernel<float><<<grid, block, 0, ctx->stream>>>(s, H, W, C, nn,
                                                                    static_cast<float*>(dst) + size_t(n0) * H * W * Cp, Cp);
    else
      hwcn_f32_to_nhwc_kernel<__half><<<grid, block, 0, ctx->stream>>>(s, H, W, C, nn,
                                                                     static_cast<__half*>(dst) + size_t(n0) * H * W * Cp, Cp);
    XEMO_LAUNCHED(ctx, 1);
  }
  return XEMO_OK;
}

extern "C" int xemo_op_nhwc_to_hwcn(xemo_ctx* ctx, const void* src, int src_f32, int H, int W, int C, int N, int Cp,
                                    float* dst) {
  XEMO_REQUIRE(ctx, src && dst && Cp >= C && H > 0 && W > 0 && W <= 65535 && C > 0 && N > 0, "nhwc_to_hwcn: bad arguments");
  dim3 block(32, 8), grid((H + 31) / 32, (C + 31) / 32, 1);
  const int per = 65535 / W;
  for (int n0 = 0; n0 < N; n0 += per) {
    const int nn = N - n0 < per ? N - n0 : per;
    grid.z = unsigned(nn) * W;
    float* d = dst + size_t(n0) * H * W * C;
    if (src_f32)
      nhwc_to_hwcn_f32_kernel<float><<<grid, block, 0, ctx->stream>>>(static_cast<const float*>(src) + size_t(n0) * H * W * Cp, H,
                                                                    W, C, nn, Cp, d);
    else
      nhwc_to_hwcn_f32_kernel<__half><<<grid, block, 0, ctx->stream>>>(static_cast<const __half*>(src) + size_t(n0) * H * W * Cp,
                                                                     H, W, C, nn, Cp, d);
    XEMO_LAUNCHED(ctx, 1);
  }
  return XEMO_OK;
}

extern "C" int xemo_op_filters_to_krsc(xemo_ctx* ctx, const float* f, int FH, int FW, int FC, int K, void* dst16, int Kp,
                                       int Cp, int flip_transpose) {
  XEMO_REQUIRE(ctx, f && dst16 && Kp >= K && Cp >= FC, "filters_to_krsc: bad arguments");
  const size_t total = size_t(Kp) * FH * FW * Cp;
  filters_to_krsc_f16_kernel<<<grid_for(total, 256, ctx->num_sms), 256, 0, ctx->stream>>>(
      f, FH, FW, FC, K, static_cast<__half*>(dst16), Kp, Cp, flip_transpose);
  XEMO_LAUNCHED(ctx, 1);
  return XEMO_OK;
}

extern "C" int xemo_op_face_rows_im2col(xemo_ctx* ctx, const float* faces, int H, int W, int C, int N, int S, int stride_w,
                                        int pad_l, int OW, void* dst16) {
  XEMO_REQUIRE(ctx, faces && dst16 && C <= 4 && S * 4 <= 32, "face_rows_im2col: needs C <= 4 and S <= 8");
  const size_t total = size_t(N) * H * OW;
  rows_im2col_from_hwcn_kernel<<<grid_for(total, 256, ctx->num_sms, 16), 256, 0, ctx->stream>>>(
      faces, H, W, C, N, S, stride_w, pad_l, OW, static_cast<__half*>(dst16));
  XEMO_LAUNCHED(ctx, 1);
  return XEMO_OK;
}

extern "C" int xemo_op_face_u8_rows_im2col(xemo_ctx* ctx, const uint8_t* faces, int IH, int IW, int N, int OHt, int OWt,
                                           const float* mean3, int S, int stride_w, int pad_l, int OW, void* dst16) {
  XEMO_REQUIRE(ctx, faces && mean3 && dst16 && S * 4 <= 32 && IH > 0 && IW > 0, "face_u8_rows_im2col: bad arguments");
  const size_t total = size_t(N) * OHt * OW;
  face_u8_rows_im2col_kernel<<<grid_for(total, 256, ctx->num_sms, 2 * per_sm_of(face_u8_rows_im2col_kernel)), 256, 0, ctx->stream>>>(
      faces, IH, IW, N, OHt, OWt, mean3, S, stride_w, pad_l, OW, static_cast<__half*>(dst16));
  XEMO_LAUNCHED(ctx, 1);
  return XEMO_OK;
}

extern "C" int xemo_op_spec_s2d(xemo_ctx* ctx, const float* spec, int H, int W, int N, int pad_t, int pad_l, int HP, int OW,
                                void* dst16) {
  XEMO_REQUIRE(ctx, spec && dst16, "spec_s2d: null pointer");
  const size_t total = size_t(N) * HP * OW;
  spec_s2d_from_hwcn_kernel<<<grid_for(total, 256, ctx->num_sms, 2 * per_sm_of(spec_s2d_from_hwcn_kernel)), 256, 0, ctx->stream>>>(
      spec, H, W, N, pad_t, pad_l, HP, OW, static_cast<__half*>(dst16));
  XEMO_LAUNCHED(ctx, 1);
  return XEMO_OK;
}

extern "C" int xemo_op_spectrogram(xemo_ctx* ctx, const float* wav, int N, int L, int Nw, int Ns, int nfft, float alpha,
                                   float scale, int W, float* spec) {
  XEMO_REQUIRE(ctx, wav && spec && N > 0 && Nw > 1 && Ns > 0 && nfft >= Nw && nfft <= 4096 && (nfft & (nfft - 1)) == 0,
               "spectrogram: nfft must be a power of two in [Nw, 4096]");
  XEMO_REQUIRE(ctx, W > 0 && W <= 65535 * 4 && size_t(W - 1) * Ns + Nw <= size_t(L), "spectrogram: %d frames need %zu samples, clip has %d", W,
               size_t(W - 1) * Ns + Nw, L);
  int log2n = 0;
  while ((1 << log2n) < nfft) ++log2n;
  dim3 grid(W, N);
  spectrogram_kernel<<<grid, 256, size_t(nfft) * sizeof(float2), ctx->stream>>>(wav, L, Nw, Ns, nfft, log2n, alpha, scale, W, spec);
  XEMO_LAUNCHED(ctx, 1);
  return XEMO_OK;
}

extern "C" int xemo_op_spec_rownorm(xemo_ctx* ctx, float* spec, int H, int W, int N) {
  XEMO_REQUIRE(ctx, spec && H > 0 && W > 1 && N > 0, "spec_rownorm: bad arguments");
  spec_rownorm_kernel<<<N, 512, 0, ctx->stream>>>(spec, H, W);
  XEMO_LAUNCHED(ctx, 1);
  return XEMO_OK;
}

// ================================================================================================
// convolution
static int run_fprop(xemo_ctx* ctx, const ConvGeom& g, const __half* x, const __half* w, const ConvEpilogue& e) {
  ConvPlan plan;
  if (!conv_fprop_plan(&plan, g, x, w, e, ctx->num_sms))
    return fail(ctx, XEMO_ERR_INVALID,
                "conv: unsupported geometry N=%d H=%d W=%d Cin=%d Kout=%d R=%d S=%d stride=%d,%d pad=%d,%d,%d,%d", g.N, g.H,
                g.W, g.Cin, g.Kout, g.R, g.S, g.sh, g.sw, g.pt, g.pb, g.pl, g.pr);
  cudaError_t err = conv_fprop_run(plan, ctx->stream);
  if (err != cudaSuccess) return fail(ctx, XEMO_ERR_CUDA, "conv_fprop launch failed: %s", cudaGetErrorString(err));
  ctx->launches += 1;
  return XEMO_OK;
}

extern "C" int xemo_op_conv_fwd(xemo_ctx* ctx, const void* x16, int N, int H, int W, int Cin, const void* w16, int Kout, int R,
                                int S, int sh, int sw, int pt, int pb, int pl, int pr, const float* scale, const float* shift,
                                const void* residual16, int relu, void* out16, float* out32, int ldc) {
  XEMO_REQUIRE(ctx, x16 && w16 && (out16 || out32), "conv_fwd: null pointer");
  XEMO_REQUIRE(ctx, Cin % 16 == 0 && Kout % 16 == 0, "conv_fwd: Cin=%d and Kout=%d must be multiples of 16", Cin, Kout);
  XEMO_REQUIRE(ctx, pl <= 127 && pt <= 127 && R <= 256 && S <= 256, "conv_fwd: pad / filter out of TMA range");
  ConvGeom g{N, H, W, Cin, Kout, R, S, sh, sw, pt, pb, pl, pr};
  XEMO_REQUIRE(ctx, g.OH() > 0 && g.OW() > 0, "conv_fwd: empty output");
  ConvEpilogue e;
  e.scale = scale; e.shift = shift; e.residual = static_cast<const __half*>(residual16); e.relu = relu;
  e.out = static_cast<__half*>(out16); e.out_f32 = out32; e.ldc = ldc ? ldc : Kout;
  XEMO_REQUIRE(ctx, e.ldc >= Kout && e.ldc % 8 == 0, "conv_fwd: ldc=%d must be >= Kout and a multiple of 8", e.ldc);
  return run_fprop(ctx, g, static_cast<const __half*>(x16), static_cast<const __half*>(w16), e);
}

// Shape planning without a device (no tensor-map encoding): what tile shape / pipeline depth / work split the host code
// picks for a convolution.  out[] (12 ints):
//   fprop: bk, block_n, num_m_tiles, num_n_tiles, num_stages, epi_bufs, b_resident, use_tma_store, smem bytes, grid, epi_cw, k_iters
//   wgrad: chunk_a, chunk_b, block_c, c_tiles, T, mt, pix, groups, splits, num_stages, smem bytes, grid
extern "C" int xemo_debug_conv_plan(int N, int H, int W, int Cin, int Kout, int R, int S, int sh, int sw, int pt, int pb, int pl,
                                    int pr, int num_sms, int* out) {
  if (!out) return XEMO_ERR_INVALID;
  ConvGeom g{N, H, W, Cin, Kout, R, S, sh, sw, pt, pb, pl, pr};
  ConvEpilogue e;
  e.out = reinterpret_cast<__half*>(uintptr_t(16));   // never dereferenced: only "an fp16 output exists"
  e.ldc = Kout;
  ConvPlan plan;
  if (!conv_fprop_plan(&plan, g, nullptr, nullptr, e, num_sms, 0, false)) return XEMO_ERR_INVALID;
  const ConvFpropParams& p = plan.p;
  const int v[12] = {plan.bk, p.block_n, p.num_m_tiles, p.num_n_tiles, p.num_stages, p.epi_bufs, p.b_resident, p.use_tma_store,
                     plan.smem, plan.grid, p.epi_cw, p.R * p.S * p.kc_blocks};
  for (int i = 0; i < 12; ++i) out[i] = v[i];
  return XEMO_OK;
}

extern "C" int xemo_debug_conv_plan2(int N, int H, int W, int Cin, int Kout, int R, int S, int sh, int sw, int pt, int pb, int pl,
                                     int pr, int num_sms, int* out13) {
  if (!out13) return XEMO_ERR_INVALID;
  const int rc = xemo_debug_conv_plan(N, H, W, Cin, Kout, R, S, sh, sw, pt, pb, pl, pr, num_sms, out13);
  if (rc) return rc;
  ConvGeom g{N, H, W, Cin, Kout, R, S, sh, sw, pt, pb, pl, pr};
  ConvEpilogue e;
  e.out = reinterpret_cast<__half*>(uintptr_t(16));
  e.ldc = Kout;
  ConvPlan plan;
  if (!conv_fprop_plan(&plan, g, nullptr, nullptr, e, num_sms, 0, false)) return XEMO_ERR_INVALID;
  out13[12] = plan.ctas;
  out13[9] = plan.grid;
  return XEMO_OK;
}

extern "C" int xemo_debug_set_conv_pair_mode(int mode) {
  conv_pair_mode_override() = mode;
  return XEMO_OK;
}

extern "C" int xemo_debug_wgrad_plan(int N, int H, int W, int Cin, int ldy, int Kout, int R, int S, int sh, int sw, int pt, int pb,
                                     int pl, int pr, int num_sms, int* out) {
  if (!out) return XEMO_ERR_INVALID;
  ConvGeom g{N, H, W, Cin, Kout, R, S, sh, sw, pt, pb, pl, pr};
  WgradPlan plan;
  if (!conv_wgrad_plan(&plan, g, nullptr, nullptr, ldy, nullptr, 1.f, num_sms, false)) return XEMO_ERR_INVALID;
  const ConvWgradParams& p = plan.p;
  const int v[12] = {p.chunk_a, p.chunk_b, p.block_c, p.c_tiles, p.T, p.mt, p.pix, p.groups, p.splits, p.num_stages, plan.smem, plan.grid};
  for (int i = 0; i < 12; ++i) out[i] = v[i];
  return XEMO_OK;
}

// parity classes of the data gradient (see hbm_kernels_extra.cuh)
static int dgrad_classes(int Cin, int Kout, int R, int S, int sh, int sw, int pt, int pl, DgradPackParams* p) {
  p->num_classes = 0;
  p->Kout = Kout; p->R = R; p->S = S; p->Cin = Cin; p->sh = sh; p->sw = sw;
  long long off = 0;
  if (sh * sw > 16) return -1;
  for (int ph = 0; ph < sh; ++ph)
    for (int pw = 0; pw < sw; ++pw) {
      DgradClass& c = p->cls[p->num_classes++];
      c.r0 = (ph + pt) % sh;
      c.s0 = (pw + pl) % sw;
      c.Jh = c.r0 < R ? (R - c.r0 + sh - 1) / sh : 0;
      c.Jw = c.s0 < S ? (S - c.s0 + sw - 1) / sw : 0;
      c.offset = off;
      off += (long long)Cin * c.Jh * c.Jw * Kout;
    }
  return 0;
}

extern "C" size_t xemo_dgrad_pack_elems(int Cin, int Kout, int R, int S, int sh, int sw) {
  (void)sh; (void)sw;
  return size_t(Cin) * R * S * Kout;  // every filter tap belongs to exactly one parity class
}

extern "C" int xemo_op_pack_dgrad_filters(xemo_ctx* ctx, const void* w16_krsc, int Kout, int R, int S, int Cin, int sh,
                                          int sw, int pt, int pl, void* packed16) {
  XEMO_REQUIRE(ctx, w16_krsc && packed16, "pack_dgrad_filters: null pointer");
  DgradPackParams p;
  XEMO_REQUIRE(ctx, dgrad_classes(Cin, Kout, R, S, sh, sw, pt, pl, &p) == 0, "pack_dgrad_filters: stride too large");
  dim3 grid(grid_for(size_t(Cin) * R * S * Kout / p.num_classes + 1, 256, ctx->num_sms), p.num_classes);
  dgrad_pack_kernel<<<grid, 256, 0, ctx->stream>>>(static_cast<const __half*>(w16_krsc), p, static_cast<__half*>(packed16));
  XEMO_LAUNCHED(ctx, 1);
  return XEMO_OK;
}

// shared by xemo_op_conv_dgrad (fp16 out) and the vl_nnconv boundary (fp32 out)
int xemo_conv_dgrad_impl(xemo_ctx* ctx, const void* dy16, int N, int H, int W, int Cin, const void* packed16, int Kout, int R,
                         int S, int sh, int sw, int pt, int pb, int pl, int pr, void* dx16, float* dx32,
                         const float* out_scale = nullptr) {
  XEMO_REQUIRE(ctx, dy16 && packed16 && (dx16 || dx32), "conv_dgrad: null pointer");
  XEMO_REQUIRE(ctx, Cin % 16 == 0 && Kout % 16 == 0, "conv_dgrad: Cin=%d and Kout=%d must be multiples of 16", Cin, Kout);
  const int OH = (H + pt + pb - R) / sh + 1, OW = (W + pl + pr - S) / sw + 1;
  XEMO_REQUIRE(ctx, OH > 0 && OW > 0, "conv_dgrad: empty output");
  DgradPackParams p;
  XEMO_REQUIRE(ctx, dgrad_classes(Cin, Kout, R, S, sh, sw, pt, pl, &p) == 0, "conv_dgrad: stride too large");
  bool any_empty = false;
  for (int i = 0; i < p.num_classes; ++i) any_empty |= (p.cls[i].Jh == 0 || p.cls[i].Jw == 0);
  // rows / columns of x that no output window reaches keep a zero gradient only through taps that
  // fall outside dY, which the TMA zero-fills -- but classes without any tap are never written
  if (any_empty) {
    if (dx16) XEMO_CUDA(ctx, cudaMemsetAsync(dx16, 0, size_t(N) * H * W * Cin * 2, ctx->stream));
    if (dx32) XEMO_CUDA(ctx, cudaMemsetAsync(dx32, 0, size_t(N) * H * W * Cin * 4, ctx->stream));
  }
  int ci = 0;
  for (int ph = 0; ph < sh; ++ph)
    for (int pw = 0; pw < sw; ++pw, ++ci) {
      const DgradClass& c = p.cls[ci];
      if (c.Jh == 0 || c.Jw == 0) continue;
      const int sub_h = (H - ph + sh - 1) / sh, sub_w = (W - pw + sw - 1) / sw;
      if (sub_h <= 0 || sub_w <= 0) continue;
      const int qh = (ph + pt) / sh, qw = (pw + pl) / sw;
      const int pad_t = c.Jh - 1 - qh, pad_l = c.Jw - 1 - qw;
      // a negative value (forward padding larger than the sub-filter reach) simply starts the TMA
      // im2col bounding box inside dY
      XEMO_REQUIRE(ctx, pad_t >= -127 && pad_l >= -127 && pad_t <= 127 && pad_l <= 127, "conv_dgrad: padding out of TMA range");
      ConvGeom g{N, OH, OW, Kout, Cin, c.Jh, c.Jw, 1, 1, pad_t, 0, pad_l, 0};
      g.oh_override = sub_h;
      g.ow_override = sub_w;
      ConvEpilogue e;
      e.scale = out_scale;   // per-input-channel factor (split-operand mode of the boundary operator)
      const size_t base = (size_t(ph) * W + pw) * Cin;
      if (sh == 1 && sw == 1) {
        e.out = static_cast<__half*>(dx16);
        e.out_f32 = dx32;
        e.ldc = Cin;
      } else {
        e.out = dx16 ? static_cast<__half*>(dx16) + base : nullptr;
        e.out_f32 = dx32 ? dx32 + base : nullptr;
        e.strided_out = 1;
        e.out_sn = (long long)H * W * Cin;
        e.out_sh = (long long)sh * W * Cin;
        e.out_sw = (long long)sw * Cin;
        e.ldc = Cin;
      }
      int rc = run_fprop(ctx, g, static_cast<const __half*>(dy16), static_cast<const __half*>(packed16) + c.offset, e);
      if (rc) return rc;
    }
  return XEMO_OK;
}

extern "C" int xemo_op_conv_dgrad(xemo_ctx* ctx, const void* dy16, int N, int H, int W, int Cin, const void* packed16,
                                  int Kout, int R, int S, int sh, int sw, int pt, int pb, int pl, int pr, void* dx16) {
  return xemo_conv_dgrad_impl(ctx, dy16, N, H, W, Cin, packed16, Kout, R, S, sh, sw, pt, pb, pl, pr, dx16, nullptr);
}

// Full-height filters (the student's fc6: 9 x 1 over a 9 x W map): the parity-decomposed form above walks R taps of which
// R - 1 fall outside the one-row dY for every output pixel (196 TFLOP/s of algorithmic work); as a GEMM over (n, w) rows
// with the R * Cin columns scattered to the R rows of dX by N tile it runs without the zero taps.
extern "C" int xemo_op_pack_dgrad_filters_fullheight(xemo_ctx* ctx, const void* w16_krsc, int Kout, int R, int Cin, void* packed16) {
  XEMO_REQUIRE(ctx, w16_krsc && packed16, "pack_dgrad_filters_fullheight: null pointer");
  XEMO_REQUIRE(ctx, R >= 1 && R <= 65535, "pack_dgrad_filters_fullheight: filter height out of range");
  dgrad_pack_fullheight_kernel<<<dim3((Cin + 31) / 32, (Kout + 31) / 32, R), dim3(32, 8), 0, ctx->stream>>>(
      static_cast<const __half*>(w16_krsc), Kout, R, Cin, static_cast<__half*>(packed16));
  XEMO_LAUNCHED(ctx, 1);
  return XEMO_OK;
}

extern "C" int xemo_op_conv_dgrad_fullheight(xemo_ctx* ctx, const void* dy16, int N, int H, int W, int Cin, const void* packed16, int Kout,
                                             void* dx16) {
  XEMO_REQUIRE(ctx, dy16 && packed16 && dx16, "conv_dgrad_fullheight: null pointer");
  XEMO_REQUIRE(ctx, Cin % 16 == 0 && Cin <= 256 && Kout % 16 == 0, "conv_dgrad_fullheight: Cin=%d must be a multiple of 16 up to 256, Kout=%d a multiple of 16", Cin, Kout);
  ConvGeom g{N, 1, W, Kout, H * Cin, 1, 1, 1, 1, 0, 0, 0, 0};     // dY as a 1 x W map with Kout channels; H * Cin output columns
  ConvEpilogue e;
  e.out = static_cast<__half*>(dx16);
  e.strided_out = 1;
  e.out_sn = (long long)H * W * Cin;    // image
  e.out_sh = 0;
  e.out_sw = Cin;                       // pixel (n, w) of row 0; N tile h lands on row h
  e.out_stile = (long long)W * Cin;
  e.ldc = Cin;
  ConvPlan plan;
  if (!conv_fprop_plan(&plan, g, static_cast<const __half*>(dy16), static_cast<const __half*>(packed16), e, ctx->num_sms, Cin))
    return fail(ctx, XEMO_ERR_INVALID, "conv_dgrad_fullheight: unsupported geometry N=%d H=%d W=%d Cin=%d Kout=%d", N, H, W, Cin, Kout);
  cudaError_t err = conv_fprop_run(plan, ctx->stream);
  if (err != cudaSuccess) return fail(ctx, XEMO_ERR_CUDA, "conv_fprop launch failed: %s", cudaGetErrorString(err));
  ctx->launches += 1;
  return XEMO_OK;
}

extern "C" int xemo_op_conv_wgrad(xemo_ctx* ctx, const void* x16, int N, int H, int W, int Cin, const void* dy16, int ldy,
                                  int Kout, int R, int S, int sh, int sw, int pt, int pb, int pl, int pr, float* dF,
                                  float scale) {
  XEMO_REQUIRE(ctx, x16 && dy16 && dF, "conv_wgrad: null pointer");
  ConvGeom g{N, H, W, Cin, Kout, R, S, sh, sw, pt, pb, pl, pr};
  WgradPlan plan;
  if (!conv_wgrad_plan(&plan, g, static_cast<const __half*>(x16), static_cast<const __half*>(dy16), ldy, dF, scale,
                       ctx->num_sms, true, ctx->deterministic != 0))
    return fail(ctx, XEMO_ERR_INVALID, "conv_wgrad: unsupported geometry Cin=%d Kout=%d ldy=%d R=%d S=%d", Cin, Kout, ldy, R, S);
  cudaError_t err = conv_wgrad_run(plan, ctx->stream);
  if (err != cudaSuccess) return fail(ctx, XEMO_ERR_CUDA, "conv_wgrad launch failed: %s", cudaGetErrorString(err));
  ctx->launches += 1;
  return XEMO_OK;
}

extern "C" int xemo_op_colsum(xemo_ctx* ctx, const void* dy16, size_t P, int ld, int C, float scale, float* out) {
  XEMO_REQUIRE(ctx, dy16 && out && C <= ld, "colsum: bad arguments");
  XEMO_CUDA(ctx, cudaMemsetAsync(out, 0, size_t(C) * 4, ctx->stream));
  int row_blocks = int((P + 511) / 512);
  const int cap = ctx->num_sms * 4;
  if (row_blocks > cap) row_blocks = cap;
  if (row_blocks < 1 || ctx->deterministic) row_blocks = 1;   // (one slab: no fp32 atomics between slabs)
  dim3 grid((C + 31) / 32, row_blocks), block(32, 8);
  colsum_kernel<__half><<<grid, block, 0, ctx->stream>>>(static_cast<const __half*>(dy16), P, ld, C, scale, out);
  XEMO_LAUNCHED(ctx, 1);
  return XEMO_OK;
}

// ================================================================================================
// pooling
static bool pool_geom(PoolGeom* g, int N, int H, int W, int C, int PH, int PW, int sh, int sw, int pt, int pb, int pl, int pr) {
  g->N = N; g->H = H; g->W = W; g->C = C; g->PH = PH; g->PW = PW; g->sh = sh; g->sw = sw; g->pt = pt; g->pl = pl;
  g->OH = (H + pt + pb - PH) / sh + 1;
  g->OW = (W + pl + pr - PW) / sw + 1;
  g->OC = C;
  return g->OH > 0 && g->OW > 0 && C % 8 == 0 && PH * PW <= 255;
}

static int maxpool_fwd_impl(xemo_ctx* ctx, const void* x16, int N, int H, int W, int C, int PH, int PW, int sh, int sw,
                            int pt, int pb, int pl, int pr, const float* a, const float* b, void* y16, uint8_t* argmax,
                            __half* xwin, int pooled_ld) {
  PoolGeom g;
  XEMO_REQUIRE(ctx, x16 && y16 && pool_geom(&g, N, H, W, C, PH, PW, sh, sw, pt, pb, pl, pr), "maxpool_fwd: bad geometry");
  if (pooled_ld) {
    XEMO_REQUIRE(ctx, pooled_ld >= C && pooled_ld % 8 == 0 && ((PH == 3 && PW == 3) || (PH == 5 && PW == 3)),
                 "maxpool_fwd: a pooled-side pitch needs the 3x3 / 5x3 fast paths and ld >= C, ld %% 8 == 0");
    g.OC = pooled_ld;
  }
  const size_t total = size_t(N) * g.OH * g.OW * (C / 8);
  XEMO_REQUIRE(ctx, size_t(N) * g.OH * g.OW < (size_t(1) << 31), "maxpool_fwd: tensor too large for 32-bit pixel indices");
  const __half* xp = static_cast<const __half*>(x16);
  __half* yp = static_cast<__half*>(y16);
#define XEMO_POOL_GRID(K) fixed_channel_grid(total, C / 8, 256, ctx->num_sms, per_sm_of(K))
#define XEMO_POOL_FWD(AFF, PHc, PWc)                                                                                      \
  maxpool_fwd_kernel<__half, AFF, PHc, PWc><<<XEMO_POOL_GRID((maxpool_fwd_kernel<__half, AFF, PHc, PWc>)), 256, 0, ctx->stream>>>(xp, g, a, b, yp, argmax)
#define XEMO_POOL_FWD_H2(AFF, PHc, PWc)                                                                                   \
  do {                                                                                                                     \
    if (nopad) maxpool_fwd_h2_kernel<AFF, PHc, PWc, true><<<XEMO_POOL_GRID((maxpool_fwd_h2_kernel<AFF, PHc, PWc, true>)), 256, 0, ctx->stream>>>(xp, g, a, b, yp, argmax, xwin);    \
    else maxpool_fwd_h2_kernel<AFF, PHc, PWc, false><<<XEMO_POOL_GRID((maxpool_fwd_h2_kernel<AFF, PHc, PWc, false>)), 256, 0, ctx->stream>>>(xp, g, a, b, yp, argmax, xwin);         \
  } while (0)
  const bool nopad = (pt == 0 && pl == 0 && (g.OH - 1) * sh + PH <= H && (g.OW - 1) * sw + PW <= W);
  XEMO_REQUIRE(ctx, !xwin || (PH == 3 && PW == 3) || (PH == 5 && PW == 3), "maxpool_fwd_win: only the 3x3 and 5x3 windows record the winner");
  if (PH == 3 && PW == 3) { if (a) XEMO_POOL_FWD_H2(true, 3, 3); else XEMO_POOL_FWD_H2(false, 3, 3); }
  else if (PH == 5 && PW == 3) { if (a) XEMO_POOL_FWD_H2(true, 5, 3); else XEMO_POOL_FWD_H2(false, 5, 3); }
  else { if (a) XEMO_POOL_FWD(true, 0, 0); else XEMO_POOL_FWD(false, 0, 0); }
#undef XEMO_POOL_FWD_H2
#undef XEMO_POOL_FWD
#undef XEMO_POOL_GRID
  XEMO_LAUNCHED(ctx, 1);
  return XEMO_OK;
}

extern "C" int xemo_op_maxpool_fwd(xemo_ctx* ctx, const void* x16, int N, int H, int W, int C, int PH, int PW, int sh, int sw,
                                   int pt, int pb, int pl, int pr, const float* a, const float* b, void* y16,
                                   uint8_t* argmax) {
  return maxpool_fwd_impl(ctx, x16, N, H, W, C, PH, PW, sh, sw, pt, pb, pl, pr, a, b, y16, argmax, nullptr, 0);
}

extern "C" int xemo_op_maxpool_fwd_win(xemo_ctx* ctx, const void* x16, int N, int H, int W, int C, int PH, int PW, int sh,
                                       int sw, int pt, int pb, int pl, int pr, const float* a, const float* b, void* y16,
                                       uint8_t* argmax, void* xwin16, int pooled_ld) {
  return maxpool_fwd_impl(ctx, x16, N, H, W, C, PH, PW, sh, sw, pt, pb, pl, pr, a, b, y16, argmax, static_cast<__half*>(xwin16),
                          pooled_ld);
}

static int maxpool_bwd_impl(xemo_ctx* ctx, const void* dy16, const uint8_t* argmax, int N, int H, int W, int C, int PH, int PW,
                            int sh, int sw, int pt, int pb, int pl, int pr, void* dx16, int pooled_ld) {
  PoolGeom g;
  XEMO_REQUIRE(ctx, dy16 && argmax && dx16 && pool_geom(&g, N, H, W, C, PH, PW, sh, sw, pt, pb, pl, pr),
               "maxpool_bwd: bad geometry");
  if (pooled_ld) {
    XEMO_REQUIRE(ctx, pooled_ld >= C && pooled_ld % 8 == 0 && PH == 3 && PW == 3 && sh == 2 && sw == 2 && pt == 0 && pl == 0,
                 "maxpool_bwd: a pooled-side pitch needs the 3x3 / stride 2 / pad 0 path and ld >= C, ld %% 8 == 0");
    g.OC = pooled_ld;
  }
  const size_t total = size_t(N) * H * W * (C / 8);
  XEMO_REQUIRE(ctx, size_t(N) * H * W < (size_t(1) << 31), "maxpool_bwd: tensor too large for 32-bit pixel indices");
  if (PH == 3 && PW == 3 && sh == 2 && sw == 2 && pt == 0 && pl == 0) {
    const size_t cells = size_t(N) * ((H + 1) / 2) * ((W + 1) / 2) * (C / 8);
    // (this one gathers through the arg-max bytes and writes 3x what it reads: it wants several waves -- measured at
    // 1 / 2 / 4 / 8 / 16 waves of the 6 blocks that fit: 0.691 / 0.574 / 0.527 / 0.505 / 0.500 ms on the student's first
    // pooling layer, where the BN / pooling-forward kernels are fastest at exactly one wave; XEMO_POOLBWD_WAVES)
    static const int pool_waves = [] { const char* e = getenv("XEMO_POOLBWD_WAVES"); const int v = e ? atoi(e) : 8; return v >= 1 && v <= 64 ? v : 8; }();
    const int per_sm = resident_blocks(maxpool_bwd_3x3s2_h2_kernel, 256) * pool_waves;
    maxpool_bwd_3x3s2_h2_kernel<<<fixed_channel_grid(cells, C / 8, 256, ctx->num_sms, per_sm), 256, 0, ctx->stream>>>(
        static_cast<const __half*>(dy16), argmax, g, static_cast<__half*>(dx16));
  } else if ((PH + sh - 1) / sh == 2 && (PW + sw - 1) / sw == 2)
    maxpool_bwd_h2_kernel<2, 2><<<fixed_channel_grid(total, C / 8, 256, ctx->num_sms, per_sm_of(maxpool_bwd_h2_kernel<2, 2>)), 256, 0, ctx->stream>>>(
        static_cast<const __half*>(dy16), argmax, g, static_cast<__half*>(dx16));
  else
    maxpool_bwd_kernel<__half, 0, 0><<<fixed_channel_grid(total, C / 8, 256, ctx->num_sms, per_sm_of(maxpool_bwd_kernel<__half, 0, 0>)), 256, 0, ctx->stream>>>(
        static_cast<const __half*>(dy16), argmax, g, static_cast<__half*>(dx16));
  XEMO_LAUNCHED(ctx, 1);
  return XEMO_OK;
}

extern "C" int xemo_op_maxpool_bwd(xemo_ctx* ctx, const void* dy16, const uint8_t* argmax, int N, int H, int W, int C, int PH,
                                   int PW, int sh, int sw, int pt, int pb, int pl, int pr, void* dx16) {
  return maxpool_bwd_impl(ctx, dy16, argmax, N, H, W, C, PH, PW, sh, sw, pt, pb, pl, pr, dx16, 0);
}

extern "C" int xemo_op_maxpool_bwd_ld(xemo_ctx* ctx, const void* dy16, const uint8_t* argmax, int N, int H, int W, int C, int PH,
                                      int PW, int sh, int sw, int pt, int pb, int pl, int pr, void* dx16, int pooled_ld) {
  return maxpool_bwd_impl(ctx, dy16, argmax, N, H, W, C, PH, PW, sh, sw, pt, pb, pl, pr, dx16, pooled_ld);
}

extern "C" int xemo_op_avgpool_fwd(xemo_ctx* ctx, const void* x16, int N, int H, int W, int C, int PH, int PW, int sh, int sw,
                                   int pt, int pb, int pl, int pr, void* y16) {
  PoolGeom g;
  XEMO_REQUIRE(ctx, x16 && y16 && pool_geom(&g, N, H, W, C, PH, PW, sh, sw, pt, pb, pl, pr), "avgpool_fwd: bad geometry");
  const size_t total = size_t(N) * g.OH * g.OW * (C / 8);
  avgpool_fwd_kernel<__half><<<grid_for(total, 256, ctx->num_sms, 16), 256, 0, ctx->stream>>>(
      static_cast<const __half*>(x16), g, static_cast<__half*>(y16));
  XEMO_LAUNCHED(ctx, 1);
  return XEMO_OK;
}

extern "C" int xemo_op_avgpool_bwd(xemo_ctx* ctx, const void* dy16, int N, int H, int W, int C, int PH, int PW, int sh, int sw,
                                   int pt, int pb, int pl, int pr, void* dx16) {
  PoolGeom g;
  XEMO_REQUIRE(ctx, dy16 && dx16 && pool_geom(&g, N, H, W, C, PH, PW, sh, sw, pt, pb, pl, pr), "avgpool_bwd: bad geometry");
  const size_t total = size_t(N) * H * W * (C / 8);
  avgpool_bwd_kernel<__half><<<grid_for(total, 256, ctx->num_sms, 16), 256, 0, ctx->stream>>>(
      static_cast<const __half*>(dy16), g, static_cast<__half*>(dx16));
  XEMO_LAUNCHED(ctx, 1);
  return XEMO_OK;
}

// ================================================================================================
// batch normalisation
template <typename T>
int bn_stats_launch(xemo_ctx* ctx, const T* x, size_t P, int C, double* ws) {
  XEMO_CUDA(ctx, cudaMemsetAsync(ws, 0, size_t(2) * C * sizeof(double), ctx->stream));
  const BnGrid bg = bn_grid(P, C, ctx->num_sms);
  dim3 grid(bg.slabs_x, bg.slabs_y);
  bn_stats_kernel<T><<<grid, kBnThreads, 0, ctx->stream>>>(x, P, C, bg.lanes, bg.rows_par, ws);
  XEMO_LAUNCHED(ctx, 1);
  return XEMO_OK;
}
template int bn_stats_launch<float>(xemo_ctx*, const float*, size_t, int, double*);

extern "C" int xemo_op_bn_train(xemo_ctx* ctx, const void* x16, size_t P, int C, const float* g, const float* beta, float eps,
                                double* ws, float* moments, float* a, float* b) {
  XEMO_REQUIRE(ctx, x16 && g && beta && ws && moments && a && b && C % 8 == 0 && P > 0, "bn_train: bad arguments");
  int rc = bn_stats_launch<__half>(ctx, static_cast<const __half*>(x16), P, C, ws);
  if (rc) return rc;
  bn_finalize_kernel<<<(C + 127) / 128, 128, 0, ctx->stream>>>(ws, P, C, g, beta, eps, moments, a, b);
  XEMO_LAUNCHED(ctx, 1);
  return XEMO_OK;
}

extern "C" int xemo_op_bn_test(xemo_ctx* ctx, const float* moments, int C, const float* g, const float* beta,
                               const float* conv_bias, float* a, float* b) {
  XEMO_REQUIRE(ctx, moments && g && beta && a && b, "bn_test: null pointer");
  bn_affine_from_moments_kernel<<<(C + 127) / 128, 128, 0, ctx->stream>>>(moments, C, g, beta, conv_bias, a, b);
  XEMO_LAUNCHED(ctx, 1);
  return XEMO_OK;
}

extern "C" int xemo_op_affine_act(xemo_ctx* ctx, const void* x16, size_t P, int C, const float* a, const float* b, int relu,
                                  void* y16) {
  XEMO_REQUIRE(ctx, x16 && y16 && C % 8 == 0, "affine_act: bad arguments");
  affine_act_kernel<__half><<<fixed_channel_grid(P * (C / 8), C / 8, 256, ctx->num_sms, per_sm_of(affine_act_kernel<__half>)), 256, 0, ctx->stream>>>(
      static_cast<const __half*>(x16), P, C, a, b, relu, static_cast<__half*>(y16));
  XEMO_LAUNCHED(ctx, 1);
  return XEMO_OK;
}

// shared by xemo_op_bn_bwd (dy at the BN resolution) and xemo_op_bn_bwd_pool (dy gathered through a max pool)
static int bn_bwd_impl(xemo_ctx* ctx, const __half* x, const __half* dy, size_t P, int C, const float* moments, const float* a,
                       const float* b, int relu_mask, int test_mode, double* ws, __half* dx, float* dg, float* db,
                       float* dconv_bias, float inv_grad_scale, const uint8_t* idx, const PoolGeom* pg) {
  XEMO_CUDA(ctx, cudaMemsetAsync(ws, 0, size_t(2) * C * sizeof(double), ctx->stream));
  if (dconv_bias) XEMO_CUDA(ctx, cudaMemsetAsync(dconv_bias, 0, size_t(C) * 4, ctx->stream));
  // The bias of a convolution that feeds a train-mode BN has an identically zero gradient (sum_rows dx = 0: BN removes the
  // mean); the fused column sum only collects the fp16 rounding residue of dx, through order-dependent fp32 atomics.  The
  // deterministic mode writes the exact value instead.
  if (ctx->deterministic && !test_mode) dconv_bias = nullptr;
  const int C8 = C / 8;
  const BnGrid bg = bn_grid(P, C, ctx->num_sms);
  dim3 grid(bg.slabs_x, bg.slabs_y);
  PoolGeom g0;
  memset(&g0, 0, sizeof(g0));
  if (pg)
    bn_bwd_reduce_kernel<__half, true><<<grid, kBnThreads, 0, ctx->stream>>>(x, dy, P, C, bg.lanes, bg.rows_par, moments, a, b, relu_mask, ws, idx, *pg);
  else
    bn_bwd_reduce_kernel<__half, false><<<grid, kBnThreads, 0, ctx->stream>>>(x, dy, P, C, bg.lanes, bg.rows_par, moments, a, b, relu_mask, ws, nullptr, g0);
  XEMO_LAUNCHED(ctx, 1);
  const size_t cs_smem = dconv_bias ? size_t(C) * 4 : 0;
  if (test_mode) {
    XEMO_REQUIRE(ctx, !pg && !dconv_bias, "bn_bwd: test mode does not support the fused pool / bias-gradient variants");
    const int egrid = fixed_channel_grid(P * C8, C8, 256, ctx->num_sms, per_sm_of(bn_bwd_test_kernel<__half>));
    bn_bwd_test_kernel<__half><<<egrid, 256, 0, ctx->stream>>>(x, dy, P, C, a, b, relu_mask, dx);
  } else if (pg) {
    const int egrid = fixed_channel_grid(P * C8, C8, 256, ctx->num_sms, per_sm_of(bn_bwd_apply_kernel<__half, true>, 256, cs_smem));
    bn_bwd_apply_kernel<__half, true><<<egrid, 256, cs_smem, ctx->stream>>>(x, dy, P, C, moments, a, b, relu_mask, ws, dx, idx, *pg, dconv_bias, inv_grad_scale);
  } else {
    const int egrid = fixed_channel_grid(P * C8, C8, 256, ctx->num_sms, per_sm_of(bn_bwd_apply_kernel<__half, false>, 256, cs_smem));
    bn_bwd_apply_kernel<__half, false><<<egrid, 256, cs_smem, ctx->stream>>>(x, dy, P, C, moments, a, b, relu_mask, ws, dx, nullptr, g0, dconv_bias, inv_grad_scale);
  }
  XEMO_LAUNCHED(ctx, 1);
  if (dg && db) {
    bn_bwd_params_kernel<<<(C + 127) / 128, 128, 0, ctx->stream>>>(ws, C, inv_grad_scale, dg, db);
    XEMO_LAUNCHED(ctx, 1);
  }
  return XEMO_OK;
}

extern "C" int xemo_op_bn_bwd(xemo_ctx* ctx, const void* x16, const void* dy16, size_t P, int C, const float* moments,
                              const float* a, const float* b, int relu_mask, int test_mode, double* ws, void* dx16,
                              float* dg, float* db, float* dconv_bias, float inv_grad_scale) {
  XEMO_REQUIRE(ctx, x16 && dy16 && moments && a && b && ws && dx16 && C % 8 == 0 && C <= 8192, "bn_bwd: bad arguments");
  return bn_bwd_impl(ctx, static_cast<const __half*>(x16), static_cast<const __half*>(dy16), P, C, moments, a, b, relu_mask,
                     test_mode, ws, static_cast<__half*>(dx16), dg, db, dconv_bias, inv_grad_scale, nullptr, nullptr);
}

extern "C" int xemo_op_bn_bwd_pool(xemo_ctx* ctx, const void* x16, const void* dpool16, const uint8_t* argmax, int N, int H,
                                   int W, int C, int PH, int PW, int sh, int sw, int pt, int pb, int pl, int pr,
                                   const float* moments, const float* a, const float* b, double* ws, void* dx16, float* dg,
                                   float* db, float* dconv_bias, float inv_grad_scale) {
  PoolGeom g;
  XEMO_REQUIRE(ctx, x16 && dpool16 && argmax && moments && a && b && ws && dx16 && C <= 8192 &&
                        pool_geom(&g, N, H, W, C, PH, PW, sh, sw, pt, pb, pl, pr),
               "bn_bwd_pool: bad arguments");
  XEMO_REQUIRE(ctx, (PH + sh - 1) / sh <= 2 && (PW + sw - 1) / sw <= 2 && size_t(N) * H * W < (size_t(1) << 31),
               "bn_bwd_pool: at most 2 x 2 windows may cover one position (use maxpool_bwd + bn_bwd otherwise)");
  return bn_bwd_impl(ctx, static_cast<const __half*>(x16), static_cast<const __half*>(dpool16), size_t(N) * H * W, C, moments, a, b,
                     1, 0, ws, static_cast<__half*>(dx16), dg, db, dconv_bias, inv_grad_scale, argmax, &g);
}

extern "C" int xemo_op_relu_bwd(xemo_ctx* ctx, const void* y16, const void* dy16, size_t n, void* dx16) {
  XEMO_REQUIRE(ctx, y16 && dy16 && dx16 && n % 8 == 0, "relu_bwd: bad arguments");
  relu_bwd_kernel<__half><<<grid_for(n / 8, 256, ctx->num_sms, 16), 256, 0, ctx->stream>>>(
      static_cast<const __half*>(y16), static_cast<const __half*>(dy16), n / 8, static_cast<__half*>(dx16));
  XEMO_LAUNCHED(ctx, 1);
  return XEMO_OK;
}

extern "C" int xemo_op_add_act(xemo_ctx* ctx, const void* a16, const void* b16, size_t n, int relu, void* y16) {
  XEMO_REQUIRE(ctx, a16 && b16 && y16 && n % 8 == 0, "add_act: bad arguments");
  add_act_kernel<__half><<<grid_for(n / 8, 256, ctx->num_sms, 16), 256, 0, ctx->stream>>>(
      static_cast<const __half*>(a16), static_cast<const __half*>(b16), n / 8, relu, static_cast<__half*>(y16));
  XEMO_LAUNCHED(ctx, 1);
  return XEMO_OK;
}

// ================================================================================================
// squeeze-and-excitation
extern "C" int xemo_op_se_squeeze(xemo_ctx* ctx, const void* u16, int N, int HW, int C, float* s) {
  XEMO_REQUIRE(ctx, u16 && s && C % 8 == 0, "se_squeeze: bad arguments");
  const int C8 = C / 8;
  int lx = 1;
  while (lx < 32 && lx * 2 <= C8) lx *= 2;   // min(32, largest power of two <= C/8)
  // Virtual pixel-stride lanes (a function of the map and the channel count only, so that the mean of a face does not
  // depend on the batch): 1024 / lx, but no more than HW / 8 -- a lane with fewer than ~8 pixels spends its time in the
  // one-load-at-a-time tail (14 x 14 maps on 32 lanes: 0.037 ms per launch, on 16 lanes 0.027) -- and 8 for the 7 x 7 maps.
  // A full 1024-lane block becomes 512 threads owning two lanes each once the grid has a block per SM (same summation
  // order, bit-identical result; measured, teacher forward at 256 faces: 4.91 ms with 1024-thread blocks, 4.80 with 512,
  // 4.90 with 256; at 32 faces 1.25 / 1.30 / 1.32 ms: few blocks want all the threads).
  const int gx = (C8 + lx - 1) / lx;
  int ly = 8;
  while (ly * 2 <= 1024 / lx && ly * 2 <= HW / 8) ly *= 2;
  const bool two = lx * ly == 1024 && gx * N >= ctx->num_sms;
  dim3 grid(gx, N), block(lx, two ? ly / 2 : ly);
  if (two) se_squeeze_kernel<__half, 2><<<grid, block, 0, ctx->stream>>>(static_cast<const __half*>(u16), HW, C, s);
  else se_squeeze_kernel<__half, 1><<<grid, block, 0, ctx->stream>>>(static_cast<const __half*>(u16), HW, C, s);
  XEMO_LAUNCHED(ctx, 1);
  return XEMO_OK;
}

extern "C" int xemo_debug_se_gate_plan(int N, int C, int Cm, int Cr, int lin, int num_sms, int* out5) {
  if (!out5 || N < 1 || C < 1 || Cr < 1) return XEMO_ERR_INVALID;
  const int groups = (N + kSeSpb - 1) / kSeSpb;
  const int K = se_gate_cluster_size(groups, C, Cr, num_sms);
  const int pc = se_gate_ranges(C, Cr), nout = C / K;
  const int tg = se_gate_groups(pc, nout);
  const size_t red = tg > 1 ? size_t(pc) * kSeSpb * nout : 0;
  out5[0] = K;
  out5[1] = pc;
  out5[2] = tg;
  out5[3] = int((size_t(kSeSpb) * ((lin ? Cm : 0) + C + Cr) + red) * sizeof(float));
  out5[4] = groups * K;
  return XEMO_OK;
}

extern "C" int xemo_debug_fixed_channel_grid(long long items, int C8, int threads, int num_sms, int per_sm) {
  if (items < 0 || C8 < 1 || threads < 1 || num_sms < 1 || per_sm < 1) return -1;
  return fixed_channel_grid(size_t(items), C8, threads, num_sms, per_sm);
}

extern "C" int xemo_op_se_gate(xemo_ctx* ctx, const float* s, int N, int C, int Cr, const float* w1, const float* b1,
                               const float* w2, const float* b2, float* gate) {
  XEMO_REQUIRE(ctx, s && w1 && w2 && gate && kSeSpb * (C + Cr) * 4 <= 48 * 1024 && C % 128 == 0, "se_gate: C must be a multiple of 128 and (C + Cr) <= 6144");
  // (a two-launch form -- hidden units over a (samples, units) grid, then gates over a (samples, channels) grid -- was
  // measured in round 2: 5.23 vs 5.12 ms teacher forward at 256 faces, 1.43 vs 1.34 ms at 32: every thread still walks a
  // dependent chain of L2 loads; removed)
  static const bool cluster_form = !(getenv("XEMO_SE_GATE_CLUSTER") && getenv("XEMO_SE_GATE_CLUSTER")[0] == '0');
  if (cluster_form) {
    SeGateParams p{s, N, C, 0, Cr, nullptr, nullptr, nullptr, w1, b1, w2, b2, gate, nullptr};
    XEMO_CUDA(ctx, se_gate_cluster_launch<false>(p, ctx->num_sms, ctx->stream));
    ctx->launches += 1;
    return XEMO_OK;
  }
  const int threads = C <= 512 ? 512 : 1024;  // latency-bound: many warps keep enough weight loads in flight
  se_gate_kernel<<<(N + kSeSpb - 1) / kSeSpb, threads, size_t(kSeSpb) * (C + Cr) * 4, ctx->stream>>>(s, N, C, Cr, w1, b1, w2, b2, gate);
  XEMO_LAUNCHED(ctx, 1);
  return XEMO_OK;
}

extern "C" int xemo_op_se_gate_lin(xemo_ctx* ctx, const float* m2, int N, int C, int Cm, int Cr, const void* w3_16, const float* a3,
                                   const float* b3, const float* w1, const float* b1, const float* w2t, const float* b2,
                                   float* nc_scale, float* nc_shift) {
  XEMO_REQUIRE(ctx, m2 && w3_16 && a3 && b3 && w1 && w2t && nc_scale && nc_shift && C % 128 == 0 && Cm % 64 == 0 &&
                        size_t(kSeSpb) * (Cm + C + Cr) * 4 <= 48 * 1024,
               "se_gate_lin: C must be a multiple of 128, Cm of 64, and (Cm + C + Cr) <= 6144");
  static const bool cluster_form = !(getenv("XEMO_SE_GATE_CLUSTER") && getenv("XEMO_SE_GATE_CLUSTER")[0] == '0');
  if (cluster_form && (Cm == 64 || Cm == 128 || Cm == 256 || Cm == 512)) {
    SeGateParams p{m2, N, C, Cm, Cr, static_cast<const __half*>(w3_16), a3, b3, w1, b1, w2t, b2, nc_scale, nc_shift};
    XEMO_CUDA(ctx, se_gate_cluster_launch<true>(p, ctx->num_sms, ctx->stream));
    ctx->launches += 1;
    return XEMO_OK;
  }
  const int threads = C <= 512 ? 512 : 1024;
  se_gate_lin_kernel<<<(N + kSeSpb - 1) / kSeSpb, threads, size_t(kSeSpb) * (Cm + C + Cr) * 4, ctx->stream>>>(
      m2, N, C, Cm, Cr, static_cast<const __half*>(w3_16), a3, b3, w1, b1, w2t, b2, nc_scale, nc_shift);
  XEMO_LAUNCHED(ctx, 1);
  return XEMO_OK;
}

// 1x1 / general convolution whose epilogue applies per-(image, channel) scale / shift (SE blocks by linearity: the default on
// the teacher's 56 x 56 / 28 x 28 stages)
// ([N][Kout] fp32) + residual + ReLU:  out = act(nc_scale[n,k]*conv(x,w) + nc_shift[n,k] + residual)
extern "C" int xemo_op_conv_fwd_nc(xemo_ctx* ctx, const void* x16, int N, int H, int W, int Cin, const void* w16, int Kout, int R,
                                   int S, int sh, int sw, int pt, int pb, int pl, int pr, const float* nc_scale,
                                   const float* nc_shift, const void* residual16, int relu, void* out16) {
  XEMO_REQUIRE(ctx, x16 && w16 && out16 && nc_scale && nc_shift, "conv_fwd_nc: null pointer");
  XEMO_REQUIRE(ctx, Cin % 64 == 0 && Kout % 16 == 0, "conv_fwd_nc: Cin=%d must be a multiple of 64 and Kout=%d of 16", Cin, Kout);
  ConvGeom g{N, H, W, Cin, Kout, R, S, sh, sw, pt, pb, pl, pr};
  XEMO_REQUIRE(ctx, g.OH() > 0 && g.OW() > 0 && g.OH() * g.OW() >= 43, "conv_fwd_nc: needs at least 43 output pixels per image");
  ConvEpilogue e;
  e.nc_scale = nc_scale; e.nc_shift = nc_shift; e.residual = static_cast<const __half*>(residual16); e.relu = relu;
  e.out = static_cast<__half*>(out16); e.ldc = Kout;
  return run_fprop(ctx, g, static_cast<const __half*>(x16), static_cast<const __half*>(w16), e);
}

extern "C" int xemo_op_se_excite(xemo_ctx* ctx, const void* u16, const float* gate, const void* shortcut16, int N, int HW,
                                 int C, int relu, void* y16) {
  XEMO_REQUIRE(ctx, u16 && gate && y16 && C % 8 == 0, "se_excite: bad arguments");
  const size_t total8 = size_t(N) * HW * (C / 8);
  se_excite_kernel<__half><<<grid_for(total8, 256, ctx->num_sms, 2 * per_sm_of(se_excite_kernel<__half>)), 256, 0, ctx->stream>>>(
      static_cast<const __half*>(u16), gate, static_cast<const __half*>(shortcut16), HW, C, total8, relu,
      static_cast<__half*>(y16));
  XEMO_LAUNCHED(ctx, 1);
  return XEMO_OK;
}

// ================================================================================================
// student stem (conv1 on the one-channel spectrogram): BN statistics and BN/filter gradients by linearity
extern "C" size_t xemo_stem_ws_doubles(void) { return size_t(kAcAccDoubles) + size_t(kStemRS); }

extern "C" int xemo_op_stem_autocorr(xemo_ctx* ctx, const void* s2d16, int N, int HP, int OW, int OH, double* ws) {
  XEMO_REQUIRE(ctx, s2d16 && ws && N > 0 && OW > 0 && OH >= 3 && HP == OH + 3, "stem_autocorr: needs HP == OH + 3 (4 x 1 taps) and OH >= 3");
  static bool attr = false;
  if (!attr) {
    XEMO_CUDA(ctx, cudaFuncSetAttribute(stem_autocorr_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kAcSmemBytes));
    attr = true;
  }
  XEMO_CUDA(ctx, cudaMemsetAsync(ws, 0, size_t(kAcAccDoubles) * sizeof(double), ctx->stream));
  const int strips = (HP + kAcRows - 1) / kAcRows, chunks = (OW + kAcOw - 1) / kAcOw;
  const long units = long(N) * strips * chunks;
  const int grid = int(units < 2L * ctx->num_sms ? units : 2L * ctx->num_sms);
  stem_autocorr_kernel<<<grid, kAcThreads, kAcSmemBytes, ctx->stream>>>(static_cast<const __half*>(s2d16), N, HP, OW, OH, ws);
  XEMO_LAUNCHED(ctx, 1);
  stem_assemble_kernel<<<1, 1024, 0, ctx->stream>>>(ws, ws + kAcAccDoubles);
  XEMO_LAUNCHED(ctx, 1);
  return XEMO_OK;
}

extern "C" int xemo_op_stem_pair_filter(xemo_ctx* ctx, const void* w16, int C, void* w2_16) {
  XEMO_REQUIRE(ctx, w16 && w2_16 && C > 0, "stem_pair_filter: bad arguments");
  stem_pair_filter_kernel<<<grid_for(size_t(2) * C * 128, 256, ctx->num_sms), 256, 0, ctx->stream>>>(
      static_cast<const __half*>(w16), C, static_cast<__half*>(w2_16));
  XEMO_LAUNCHED(ctx, 1);
  return XEMO_OK;
}

extern "C" int xemo_op_tile_f32(xemo_ctx* ctx, const float* src, int C, int reps, float fill, float* dst) {
  XEMO_REQUIRE(ctx, dst && C > 0 && reps > 0, "tile_f32: bad arguments");
  tile_f32_kernel<<<grid_for(size_t(C) * reps, 256, ctx->num_sms), 256, 0, ctx->stream>>>(src, C, reps, fill, dst);
  XEMO_LAUNCHED(ctx, 1);
  return XEMO_OK;
}

extern "C" int xemo_op_stem_bn_train(xemo_ctx* ctx, const double* ws, const void* w16, const float* bias, size_t P, int C,
                                     const float* g, const float* beta, float eps, float* moments, float* a, float* b) {
  XEMO_REQUIRE(ctx, ws && w16 && bias && g && beta && moments && a && b && C > 0 && P > 0, "stem_bn_train: bad arguments");
  stem_bn_stats_kernel<<<C, kStemT, 0, ctx->stream>>>(ws + kAcAccDoubles, static_cast<const __half*>(w16), bias, double(P), C, g,
                                                     beta, eps, moments, a, b);
  XEMO_LAUNCHED(ctx, 1);
  return XEMO_OK;
}

extern "C" int xemo_op_stem_pool_bn_reduce(xemo_ctx* ctx, const void* xwin16, void* dpool16, size_t P, int C, int ld,
                                           const float* moments, const float* a, const float* b, double* acc) {
  if (ld == 0) ld = C;
  XEMO_REQUIRE(ctx, xwin16 && dpool16 && moments && a && b && acc && C % 8 == 0 && ld % 8 == 0 && ld >= C && P > 0,
               "stem_pool_bn_reduce: bad arguments");
  XEMO_CUDA(ctx, cudaMemsetAsync(acc, 0, size_t(2) * C * sizeof(double), ctx->stream));
  const BnGrid bg = bn_grid(P, C, ctx->num_sms, per_sm_of(stem_pool_bn_reduce_kernel, kBnThreads));
  dim3 grid(bg.slabs_x, bg.slabs_y);
  stem_pool_bn_reduce_kernel<<<grid, kBnThreads, 0, ctx->stream>>>(static_cast<const __half*>(xwin16), static_cast<__half*>(dpool16),
                                                                  P, C, ld, bg.lanes, bg.rows_par, moments, a, b, acc);
  XEMO_LAUNCHED(ctx, 1);
  return XEMO_OK;
}

extern "C" int xemo_op_stem_wgrad_finalize(xemo_ctx* ctx, const double* ws, const void* w16, const float* bias,
                                           const double* acc, size_t P, int C, const float* moments, const float* a,
                                           float inv_grad_scale, float* dW, float* dbias, float* dgamma, float* dbeta,
                                           const float* g1_pair) {
  XEMO_REQUIRE(ctx, ws && w16 && bias && acc && moments && a && dW && dgamma && dbeta && C > 0 && P > 0, "stem_wgrad_finalize: bad arguments");
  stem_wgrad_finalize_kernel<<<C, kStemT, 0, ctx->stream>>>(ws + kAcAccDoubles, static_cast<const __half*>(w16), bias, acc,
                                                           double(P), C, moments, a, inv_grad_scale, dW, dbias, dgamma, dbeta, g1_pair);
  XEMO_LAUNCHED(ctx, 1);
  return XEMO_OK;
}

// ================================================================================================
// coupling, loss, update
extern "C" int xemo_op_logit_aggregate(xemo_ctx* ctx, const float* frame_logits, int ldl, const int* start, const int* end,
                                       int N, int num_pred, int use_mean, float* target) {
  XEMO_REQUIRE(ctx, frame_logits && start && end && target && num_pred <= ldl, "logit_aggregate: bad arguments");
  logit_aggregate_kernel<<<(N * num_pred + 127) / 128, 128, 0, ctx->stream>>>(frame_logits, ldl, start, end, N, num_pred,
                                                                             use_mean, target);
  XEMO_LAUNCHED(ctx, 1);
  return XEMO_OK;
}

extern "C" int xemo_op_loss(xemo_ctx* ctx, const void* x, int x_f32, int ldx, const float* t, int ldt, const float* w, int N, int C,
                            int loss_type, float T, int logit_targets, float dzdy, float grad_scale, void* dx, int dx_f32, int lddx,
                            float* scalars, float* class_stats, int* max_label) {
  XEMO_REQUIRE(ctx, x && t && scalars && C >= 1 && C <= kLossMaxC && C <= ldx && C <= ldt && (!dx || C <= lddx), "loss: bad arguments");
  XEMO_REQUIRE(ctx, loss_type >= 0 && loss_type <= 2 && T > 0.f, "loss: loss_type must be 0 (softmax CE), 1 (euclidean) or 2 (huber), T / sigma > 0");
  // (deterministic: one block of up to 1024 samples, whose warps' partial sums are added in a fixed order)
  const int threads = ctx->deterministic ? 1024 : 128;
  const int grid = (N + threads - 1) / threads;
#define XEMO_LOSS(TX, TDX)                                                                                                     \
  loss_fused_kernel<TX, TDX><<<grid, threads, 0, ctx->stream>>>(static_cast<const TX*>(x), ldx, t, ldt, w, N, C, loss_type, T,     \
                                                           logit_targets, dzdy, grad_scale, static_cast<TDX*>(dx), lddx, scalars, \
                                                           class_stats, max_label)
  if (x_f32 && dx_f32) XEMO_LOSS(float, float);
  else if (x_f32) XEMO_LOSS(float, __half);
  else if (dx_f32) XEMO_LOSS(__half, float);
  else XEMO_LOSS(__half, __half);
#undef XEMO_LOSS
  XEMO_LAUNCHED(ctx, 1);
  return XEMO_OK;
}

extern "C" int xemo_op_softmaxce(xemo_ctx* ctx, const void* x16, int ldx, const float* t, int ldt, const float* w, int N, int C,
                                 float T, int logit_targets, float dzdy, float grad_scale, void* dx16, float* scalars,
                                 float* class_stats, int* max_label) {
  return xemo_op_loss(ctx, x16, 0, ldx, t, ldt, w, N, C, 0, T, logit_targets, dzdy, grad_scale, dx16, 0, ldx, scalars, class_stats,
                      max_label);
}

extern "C" int xemo_op_grad_guard(xemo_ctx* ctx, const float* g, size_t n, int* state) {
  XEMO_REQUIRE(ctx, g && state, "grad_guard: null pointer");
  grad_guard_scan_kernel<<<grid_for(n, 256, ctx->num_sms, 8), 256, 0, ctx->stream>>>(g, n, state);
  grad_guard_publish_kernel<<<1, 1, 0, ctx->stream>>>(state);
  XEMO_LAUNCHED(ctx, 2);
  return XEMO_OK;
}

extern "C" int xemo_op_sgd_momentum_guarded(xemo_ctx* ctx, float* w, float* m, const float* g, size_t n, const float* hyper,
                                            float lr_mult, float wd_mult, float inv_grad_scale, void* w16, const int* guard) {
  XEMO_REQUIRE(ctx, w && m && g && hyper, "sgd_momentum: null pointer");
  sgd_momentum_dev_kernel<<<grid_for(n, 256, ctx->num_sms, 8), 256, 0, ctx->stream>>>(w, m, g, n, hyper, lr_mult, wd_mult,
                                                                                     inv_grad_scale,
                                                                                     static_cast<__half*>(w16), guard);
  XEMO_LAUNCHED(ctx, 1);
  return XEMO_OK;
}

extern "C" int xemo_op_sgd_momentum(xemo_ctx* ctx, float* w, float* m, const float* g, size_t n, const float* hyper,
                                    float lr_mult, float wd_mult, float inv_grad_scale, void* w16) {
  return xemo_op_sgd_momentum_guarded(ctx, w, m, g, n, hyper, lr_mult, wd_mult, inv_grad_scale, w16, nullptr);
}

extern "C" int xemo_op_moments_average_guarded(xemo_ctx* ctx, float* moments, const float* batch_moments, int n, float rate,
                                               float bm_scale, const int* guard) {
  XEMO_REQUIRE(ctx, moments && batch_moments, "moments_average: null pointer");
  moments_average_kernel<<<(n + 127) / 128, 128, 0, ctx->stream>>>(moments, batch_moments, n, rate, guard, bm_scale);
  XEMO_LAUNCHED(ctx, 1);
  return XEMO_OK;
}

extern "C" int xemo_op_moments_average(xemo_ctx* ctx, float* moments, const float* batch_moments, int n, float rate) {
  return xemo_op_moments_average_guarded(ctx, moments, batch_moments, n, rate, 1.f, nullptr);
}

extern "C" int xemo_op_cast_f32_f16(xemo_ctx* ctx, const float* src, size_t n, void* dst16) {
  XEMO_REQUIRE(ctx, src && dst16, "cast: null pointer");
  f32_to_f16_kernel<<<grid_for(n, 256, ctx->num_sms, 8), 256, 0, ctx->stream>>>(src, n, static_cast<__half*>(dst16));
  XEMO_LAUNCHED(ctx, 1);
  return XEMO_OK;
}
extern "C" int xemo_op_cast_f16_f32(xemo_ctx* ctx, const void* src16, size_t n, float scale, float* dst) {
  XEMO_REQUIRE(ctx, src16 && dst, "cast: null pointer");
  f16_to_f32_kernel<<<grid_for(n, 256, ctx->num_sms, 8), 256, 0, ctx->stream>>>(static_cast<const __half*>(src16), n, scale,
                                                                               dst);
  XEMO_LAUNCHED(ctx, 1);
  return XEMO_OK;
}
extern "C" int xemo_op_fill_strided_f32(xemo_ctx* ctx, float* dst, int outer, size_t outer_stride, size_t inner_off, int inner,
                                        float value) {
  XEMO_REQUIRE(ctx, dst && outer > 0 && inner > 0, "fill_strided: bad arguments");
  fill_strided_f32_kernel<<<grid_for(size_t(outer) * inner, 256, ctx->num_sms, 4), 256, 0, ctx->stream>>>(
      dst, outer, outer_stride, inner_off, inner, value);
  XEMO_LAUNCHED(ctx, 1);
  return XEMO_OK;
}
