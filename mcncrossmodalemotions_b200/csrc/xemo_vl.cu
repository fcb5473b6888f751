// xemo_vl.cu -- the MatConvNet-boundary operators (xemo_vl_*): `single` H x W x C x N column-major
// arrays in host or device memory, forward when dzdy == NULL and backward otherwise -- the calling
// convention of the upstream vl_nn* MEX gateways that dagnn blocks invoke from dag.eval
// (/root/reference/emoVoxCeleb/fetch_emovoxceleb_imdb.m:129, external/compute_visual_feats.m:90,
// external/compute_audio_feats.m:126) and from cnn_train_dag (emoVoxCeleb/run_distillation.m:170).
//
// Each call stages its arrays on the device (when they are host arrays), converts to the
// device-native NHWC layout, runs the same kernels the fused graphs use (convolutions on tcgen05 with
// fp16 operands / fp32 accumulation; every other operator in fp32), converts back and -- when an
// output lives in host memory -- synchronises.  This is the per-operator parity surface; the fused
// graph programs avoid the per-call layout conversions.
#include "xemo_internal.h"

#include <string.h>

#include "hbm_kernels_extra.cuh"

using namespace xemo;

int xemo_conv_dgrad_impl(xemo_ctx* ctx, const void* dy16, int N, int H, int W, int Cin, const void* packed16, int Kout, int R,
                         int S, int sh, int sw, int pt, int pb, int pl, int pr, void* dx16, float* dx32,
                         const float* out_scale = nullptr);
template <typename T>
int bn_stats_launch(xemo_ctx* ctx, const T* x, size_t P, int C, double* ws);

namespace {

bool is_device_ptr(const void* p) {
  cudaPointerAttributes at;
  if (cudaPointerGetAttributes(&at, p) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  return at.type == cudaMemoryTypeDevice || at.type == cudaMemoryTypeManaged;
}

// stream-ordered scratch memory + staging of host arrays for one boundary call
struct Arena {
  xemo_ctx* ctx;
  std::vector<void*> bufs;
  struct Out { void* user; void* dev; size_t bytes; };
  std::vector<Out> outs;
  bool failed = false;
  explicit Arena(xemo_ctx* c) : ctx(c) {}
  ~Arena() {
    for (void* b : bufs) cudaFreeAsync(b, ctx->stream);
  }
  void* alloc(size_t bytes, bool zero = false) {
    void* p = nullptr;
    if (bytes == 0) bytes = 16;
    if (!ctx->pool) {
      // the context's own pool with an unbounded release threshold: the device's default pool hands freed blocks back
      // to the driver at every synchronisation, and a chain of boundary calls then re-maps gigabytes of scratch per step
      cudaMemPoolProps props = {};
      props.allocType = cudaMemAllocationTypePinned;
      props.location.type = cudaMemLocationTypeDevice;
      props.location.id = ctx->device;
      if (cudaMemPoolCreate(&ctx->pool, &props) != cudaSuccess) { ctx->pool = nullptr; failed = true; return nullptr; }
      uint64_t keep = UINT64_MAX;
      cudaMemPoolSetAttribute(ctx->pool, cudaMemPoolAttrReleaseThreshold, &keep);
    }
    if (cudaMallocFromPoolAsync(&p, bytes, ctx->pool, ctx->stream) != cudaSuccess) { failed = true; return nullptr; }
    bufs.push_back(p);
    if (zero) cudaMemsetAsync(p, 0, bytes, ctx->stream);
    return p;
  }
  template <typename T>
  T* alloc_n(size_t n, bool zero = false) { return static_cast<T*>(alloc(n * sizeof(T), zero)); }
  // device view of an input array (copies host arrays)
  const void* in(const void* p, size_t bytes) {
    if (!p) return nullptr;
    if (is_device_ptr(p)) return p;
    void* d = alloc(bytes);
    if (d && cudaMemcpyAsync(d, p, bytes, cudaMemcpyHostToDevice, ctx->stream) != cudaSuccess) failed = true;
    return d;
  }
  // device buffer standing for an output array (the user's own when it is device memory)
  void* out(void* p, size_t bytes) {
    if (!p) return nullptr;
    if (is_device_ptr(p)) return p;
    void* d = alloc(bytes);
    outs.push_back({p, d, bytes});
    return d;
  }
  int finish() {
    if (failed) return fail(ctx, XEMO_ERR_NOMEM, "device staging allocation / copy failed");
    for (const Out& o : outs)
      XEMO_CUDA(ctx, cudaMemcpyAsync(o.user, o.dev, o.bytes, cudaMemcpyDeviceToHost, ctx->stream));
    if (!outs.empty()) XEMO_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return XEMO_OK;
  }
};

size_t numel(const xemo_array* a) { return size_t(a->h) * a->w * a->c * a->n; }

// copy `n` floats (host or device source) into a zero-initialised device vector of `np` floats
float* padded_vec(Arena& ar, const float* src, int n, int np, float fill = 0.f) {
  float* d = ar.alloc_n<float>(np, true);
  if (!d) return nullptr;
  if (fill != 0.f && np > n) {
    fill_strided_f32_kernel<<<1, 256, 0, ar.ctx->stream>>>(d, 1, 0, size_t(n), np - n, fill);
    ar.ctx->launches++;
  }
  if (src && cudaMemcpyAsync(d, src, size_t(n) * 4, cudaMemcpyDefault, ar.ctx->stream) != cudaSuccess) ar.failed = true;
  return d;
}

// ---- split-operand (fp32-equivalent) staging, see hbm_kernels_extra.cuh
unsigned* absmax_of(Arena& ar, const float* x, size_t n) {
  unsigned* a = ar.alloc_n<unsigned>(1, true);
  if (!a) return nullptr;
  absmax_f32_kernel<<<grid_for(n, 256, ar.ctx->num_sms, 8), 256, 0, ar.ctx->stream>>>(x, n, a);
  ar.ctx->launches++;
  return a;
}

// HWCN fp32 -> (hi|lo|hi) [lo_mask = 2] or (hi|hi|lo) [lo_mask = 4] fp16 NHWC, concatenated along channels or images
int stage_split(xemo_ctx* ctx, const float* src, int H, int W, int C, int N, __half* dst, int Cp, bool image_concat, int lo_mask,
                const unsigned* amax) {
  dim3 block(32, 8), grid((H + 31) / 32, (Cp + 31) / 32, 1);
  const int per = 65535 / W;
  const int row_pitch = image_concat ? Cp : 3 * Cp;
  const size_t slot_stride = image_concat ? size_t(N) * H * W * Cp : size_t(Cp);
  for (int n0 = 0; n0 < N; n0 += per) {
    const int nn = N - n0 < per ? N - n0 : per;
    grid.z = unsigned(nn) * W;
    hwcn_f32_to_nhwc_split_kernel<<<grid, block, 0, ctx->stream>>>(src + size_t(n0) * H * W * C, H, W, C, nn,
                                                                   dst + size_t(n0) * H * W * row_pitch, Cp, row_pitch, slot_stride,
                                                                   lo_mask, amax);
    XEMO_LAUNCHED(ctx, 1);
  }
  return XEMO_OK;
}

// vl_nnconv with split operands: every product carries ~22 mantissa bits, accumulation in fp32 (TMEM) -- the arithmetic
// of the reference's single-precision path (MatConvNet CPU: im2row + SGEMM) up to summation order.
int vl_nnconv_split(xemo_ctx* ctx, Arena& ar, const float* xd, const float* fd, const xemo_array* b, const xemo_array* dzdy,
                    const int pad[4], const int stride[2], xemo_array* y, xemo_array* dx, xemo_array* df, int H, int W, int C, int N,
                    int FH, int FW, int K, int OH, int OW) {
  const int Cp = pad_to(C, 16), Kp = pad_to(K, 16);
  const size_t nx = size_t(H) * W * C * N, nf = size_t(FH) * FW * C * K, ny = size_t(OH) * OW * K * N;
  int rc;
  unsigned* ax = absmax_of(ar, xd, nx);
  unsigned* af = absmax_of(ar, fd, nf);
  if (ar.failed) return ar.finish();
  if (!dzdy || !dzdy->data) {
    __half* x3 = ar.alloc_n<__half>(size_t(N) * H * W * 3 * Cp);
    __half* w3 = ar.alloc_n<__half>(size_t(Kp) * FH * FW * 3 * Cp);
    float* unscale = ar.alloc_n<float>(Kp);
    const float* bias = (b && b->data) ? padded_vec(ar, static_cast<const float*>(b->data), K, Kp) : nullptr;
    float* out32 = ar.alloc_n<float>(size_t(N) * OH * OW * Kp);
    float* yd = static_cast<float*>(ar.out(y->data, ny * 4));
    if (ar.failed) return ar.finish();
    if ((rc = stage_split(ctx, xd, H, W, C, N, x3, Cp, false, 2, ax))) return rc;
    filters_to_krsc_split_kernel<<<grid_for(size_t(3) * Kp * FH * FW * Cp, 256, ctx->num_sms), 256, 0, ctx->stream>>>(fd, FH, FW, C, K, w3, Kp, Cp, 0, af);
    split_unscale_kernel<<<1, 256, 0, ctx->stream>>>(ax, af, Kp, unscale);
    XEMO_LAUNCHED(ctx, 2);
    if ((rc = xemo_op_conv_fwd(ctx, x3, N, H, W, 3 * Cp, w3, Kp, FH, FW, stride[0], stride[1], pad[0], pad[1], pad[2], pad[3], unscale,
                               bias, nullptr, 0, nullptr, out32, Kp)))
      return rc;
    if ((rc = xemo_op_nhwc_to_hwcn(ctx, out32, 1, OH, OW, K, N, Kp, yd))) return rc;
    return ar.finish();
  }
  const float* dyd = static_cast<const float*>(ar.in(dzdy->data, ny * 4));
  if (ar.failed) return ar.finish();
  unsigned* ay = absmax_of(ar, dyd, ny);
  if (ar.failed) return ar.finish();
  if (dx && dx->data) {
    // dX = corr(dY, flipped F^T): reduction over the 3*Kp split output channels of dY (hi|lo|hi) x F (hi|hi|lo along K)
    __half* dy3 = ar.alloc_n<__half>(size_t(N) * OH * OW * 3 * Kp);
    __half* w3k = ar.alloc_n<__half>(size_t(3) * Kp * FH * FW * Cp);
    __half* packed = ar.alloc_n<__half>(xemo_dgrad_pack_elems(Cp, 3 * Kp, FH, FW, stride[0], stride[1]));
    float* unscale = ar.alloc_n<float>(Cp);
    float* dx32 = ar.alloc_n<float>(size_t(N) * H * W * Cp);
    float* dxd = static_cast<float*>(ar.out(dx->data, nx * 4));
    if (ar.failed) return ar.finish();
    if ((rc = stage_split(ctx, dyd, OH, OW, K, N, dy3, Kp, false, 2, ay))) return rc;
    filters_to_krsc_split_kernel<<<grid_for(size_t(3) * Kp * FH * FW * Cp, 256, ctx->num_sms), 256, 0, ctx->stream>>>(fd, FH, FW, C, K, w3k, Kp, Cp, 1, af);
    split_unscale_kernel<<<1, 256, 0, ctx->stream>>>(ay, af, Cp, unscale);
    XEMO_LAUNCHED(ctx, 2);
    if ((rc = xemo_op_pack_dgrad_filters(ctx, w3k, 3 * Kp, FH, FW, Cp, stride[0], stride[1], pad[0], pad[2], packed))) return rc;
    if ((rc = xemo_conv_dgrad_impl(ctx, dy3, N, H, W, Cp, packed, 3 * Kp, FH, FW, stride[0], stride[1], pad[0], pad[1], pad[2], pad[3],
                                   nullptr, dx32, unscale)))
      return rc;
    if ((rc = xemo_op_nhwc_to_hwcn(ctx, dx32, 1, H, W, C, N, Cp, dxd))) return rc;
  }
  if (df && df->data) {
    // dF = sum over pixels of x (x) dY: the three split terms as 3N images, x (hi;lo;hi) against dY (hi;hi;lo)
    __half* x3 = ar.alloc_n<__half>(size_t(3) * N * H * W * Cp);
    __half* dy3 = ar.alloc_n<__half>(size_t(3) * N * OH * OW * Kp);
    float* dF = ar.alloc_n<float>(size_t(Kp) * FH * FW * Cp, true);
    float* unscale = ar.alloc_n<float>(1);
    float* dfd = static_cast<float*>(ar.out(df->data, nf * 4));
    if (ar.failed) return ar.finish();
    if ((rc = stage_split(ctx, xd, H, W, C, N, x3, Cp, true, 2, ax))) return rc;
    if ((rc = stage_split(ctx, dyd, OH, OW, K, N, dy3, Kp, true, 4, ay))) return rc;
    split_unscale_kernel<<<1, 32, 0, ctx->stream>>>(ax, ay, 1, unscale);
    XEMO_LAUNCHED(ctx, 1);
    if ((rc = xemo_op_conv_wgrad(ctx, x3, 3 * N, H, W, Cp, dy3, Kp, Kp, FH, FW, stride[0], stride[1], pad[0], pad[1], pad[2], pad[3],
                                 dF, 1.f)))
      return rc;
    krsc_f32_to_filters_kernel<<<grid_for(nf, 256, ctx->num_sms), 256, 0, ctx->stream>>>(dF, FH, FW, C, K, Cp, dfd, unscale);
    XEMO_LAUNCHED(ctx, 1);
  }
  return XEMO_OK;
}

}  // namespace

extern "C" int xemo_out_size(int64_t h, int64_t w, int fh, int fw, const int pad[4], const int stride[2], int64_t* oh,
                             int64_t* ow) {
  if (!pad || !stride || !oh || !ow || stride[0] <= 0 || stride[1] <= 0) return XEMO_ERR_INVALID;
  const int64_t nh = h + pad[0] + pad[1] - fh, nw = w + pad[2] + pad[3] - fw;
  if (nh < 0 || nw < 0) return XEMO_ERR_INVALID;
  *oh = nh / stride[0] + 1;
  *ow = nw / stride[1] + 1;
  return XEMO_OK;
}

// ================================================================================================
extern "C" int xemo_vl_nnconv(xemo_ctx* ctx, const xemo_array* x, const xemo_array* f, const xemo_array* b,
                              const xemo_array* dzdy, const int pad[4], const int stride[2], xemo_array* y, xemo_array* dx,
                              xemo_array* df, xemo_array* db) {
  XEMO_REQUIRE(ctx, x && x->data && f && f->data && pad && stride, "vl_nnconv: X, F, pad and stride are required");
  XEMO_REQUIRE(ctx, f->c == x->c, "vl_nnconv: filter depth %lld != input depth %lld (groups are not on this path)",
               (long long)f->c, (long long)x->c);
  XEMO_REQUIRE(ctx, !b || !b->data || int64_t(numel(b)) == f->n, "vl_nnconv: bias must have K elements");
  const int H = int(x->h), W = int(x->w), C = int(x->c), N = int(x->n);
  const int FH = int(f->h), FW = int(f->w), K = int(f->n);
  int64_t OH64, OW64;
  XEMO_REQUIRE(ctx, xemo_out_size(H, W, FH, FW, pad, stride, &OH64, &OW64) == 0, "vl_nnconv: filter larger than padded input");
  const int OH = int(OH64), OW = int(OW64);
  const int Cp = pad_to(C, 16), Kp = pad_to(K, 16);
  Arena ar(ctx);
  const float* xd = static_cast<const float*>(ar.in(x->data, numel(x) * 4));
  const float* fd = static_cast<const float*>(ar.in(f->data, numel(f) * 4));
  const bool backward = dzdy && dzdy->data;
  if (backward) {
    XEMO_REQUIRE(ctx, dzdy->h == OH && dzdy->w == OW && dzdy->c == K && dzdy->n == N, "vl_nnconv: DZDY must be %d x %d x %d x %d",
                 OH, OW, K, N);
    XEMO_REQUIRE(ctx, !(dx && dx->data) || (dx->h == H && dx->w == W && dx->c == C && dx->n == N), "vl_nnconv: DX must have the size of X");
    XEMO_REQUIRE(ctx, !(df && df->data) || (df->h == FH && df->w == FW && df->c == C && df->n == K), "vl_nnconv: DF must have the size of F");
  } else {
    XEMO_REQUIRE(ctx, y && y->data, "vl_nnconv: forward needs Y");
    XEMO_REQUIRE(ctx, y->h == OH && y->w == OW && y->c == K && y->n == N, "vl_nnconv: Y must be %d x %d x %d x %d", OH, OW, K, N);
  }
  const bool split = ctx->conv_precision == 1;
  if (split) {
    if (ar.failed) return ar.finish();
    const int rc = vl_nnconv_split(ctx, ar, xd, fd, b, dzdy, pad, stride, y, dx, df, H, W, C, N, FH, FW, K, OH, OW);
    if (rc || !backward) return rc;   // (backward: the bias gradient below is shared -- it is fp32 arithmetic on dzdy)
  }
  int rc;
  const float* dyd = backward ? static_cast<const float*>(ar.in(dzdy->data, numel(dzdy) * 4)) : nullptr;
  if (!split) {
    __half* x16 = ar.alloc_n<__half>(size_t(N) * H * W * Cp);
    __half* w16 = ar.alloc_n<__half>(size_t(Kp) * FH * FW * Cp);
    if (ar.failed) return ar.finish();
    if ((rc = xemo_op_hwcn_to_nhwc(ctx, xd, H, W, C, N, x16, Cp, 0))) return rc;
    if ((rc = xemo_op_filters_to_krsc(ctx, fd, FH, FW, C, K, w16, Kp, Cp, 0))) return rc;

    if (!backward) {
      const float* bias = (b && b->data) ? padded_vec(ar, static_cast<const float*>(b->data), K, Kp) : nullptr;
      float* out32 = ar.alloc_n<float>(size_t(N) * OH * OW * Kp);
      float* yd = static_cast<float*>(ar.out(y->data, numel(y) * 4));
      if (ar.failed) return ar.finish();
      if ((rc = xemo_op_conv_fwd(ctx, x16, N, H, W, Cp, w16, Kp, FH, FW, stride[0], stride[1], pad[0], pad[1], pad[2], pad[3],
                                 nullptr, bias, nullptr, 0, nullptr, out32, Kp)))
        return rc;
      if ((rc = xemo_op_nhwc_to_hwcn(ctx, out32, 1, OH, OW, K, N, Kp, yd))) return rc;
      return ar.finish();
    }

    __half* dy16 = ar.alloc_n<__half>(size_t(N) * OH * OW * Kp);
    if (ar.failed) return ar.finish();
    if ((rc = xemo_op_hwcn_to_nhwc(ctx, dyd, OH, OW, K, N, dy16, Kp, 0))) return rc;
    if (dx && dx->data) {
      __half* packed = ar.alloc_n<__half>(xemo_dgrad_pack_elems(Cp, Kp, FH, FW, stride[0], stride[1]));
      float* dx32 = ar.alloc_n<float>(size_t(N) * H * W * Cp);
      float* dxd = static_cast<float*>(ar.out(dx->data, numel(dx) * 4));
      if (ar.failed) return ar.finish();
      if ((rc = xemo_op_pack_dgrad_filters(ctx, w16, Kp, FH, FW, Cp, stride[0], stride[1], pad[0], pad[2], packed))) return rc;
      if ((rc = xemo_conv_dgrad_impl(ctx, dy16, N, H, W, Cp, packed, Kp, FH, FW, stride[0], stride[1], pad[0], pad[1], pad[2],
                                     pad[3], nullptr, dx32)))
        return rc;
      if ((rc = xemo_op_nhwc_to_hwcn(ctx, dx32, 1, H, W, C, N, Cp, dxd))) return rc;
    }
    if (df && df->data) {
      float* dF = ar.alloc_n<float>(size_t(Kp) * FH * FW * Cp, true);
      float* dfd = static_cast<float*>(ar.out(df->data, numel(df) * 4));
      if (ar.failed) return ar.finish();
      if ((rc = xemo_op_conv_wgrad(ctx, x16, N, H, W, Cp, dy16, Kp, Kp, FH, FW, stride[0], stride[1], pad[0], pad[1], pad[2],
                                   pad[3], dF, 1.f)))
        return rc;
      const size_t total = size_t(FH) * FW * C * K;
      krsc_f32_to_filters_kernel<<<grid_for(total, 256, ctx->num_sms), 256, 0, ctx->stream>>>(dF, FH, FW, C, K, Cp, dfd);
      XEMO_LAUNCHED(ctx, 1);
    }
  }
  if (db && db->data) {
    XEMO_REQUIRE(ctx, int64_t(numel(db)) == K, "vl_nnconv: DB must have K elements");
    float* dbd = static_cast<float*>(ar.out(db->data, size_t(K) * 4));
    if (ar.failed) return ar.finish();
    // bias gradient from the fp32 dzdy itself (no fp16 rounding): column sums of the NHWC fp32 view
    float* dy32 = ar.alloc_n<float>(size_t(N) * OH * OW * Kp);
    if (ar.failed) return ar.finish();
    if ((rc = xemo_op_hwcn_to_nhwc(ctx, dyd, OH, OW, K, N, dy32, Kp, 1))) return rc;
    XEMO_CUDA(ctx, cudaMemsetAsync(dbd, 0, size_t(K) * 4, ctx->stream));
    const size_t P = size_t(N) * OH * OW;
    int row_blocks = int((P + 511) / 512);
    if (row_blocks > ctx->num_sms * 4) row_blocks = ctx->num_sms * 4;
    dim3 grid((K + 31) / 32, row_blocks), block(32, 8);
    colsum_kernel<float><<<grid, block, 0, ctx->stream>>>(dy32, P, Kp, K, 1.f, dbd);
    XEMO_LAUNCHED(ctx, 1);
  }
  return ar.finish();
}

// ================================================================================================
extern "C" int xemo_vl_nnpool(xemo_ctx* ctx, const xemo_array* x, const int pool[2], const xemo_array* dzdy, const int pad[4],
                              const int stride[2], int method, xemo_array* y_or_dx, uint8_t* argmax) {
  XEMO_REQUIRE(ctx, x && x->data && pool && pad && stride && y_or_dx && y_or_dx->data, "vl_nnpool: missing argument");
  XEMO_REQUIRE(ctx, method == 0 || method == 1, "vl_nnpool: method must be 0 (max) or 1 (avg)");
  const int H = int(x->h), W = int(x->w), C = int(x->c), N = int(x->n);
  int64_t OH64, OW64;
  XEMO_REQUIRE(ctx, xemo_out_size(H, W, pool[0], pool[1], pad, stride, &OH64, &OW64) == 0, "vl_nnpool: window larger than input");
  XEMO_REQUIRE(ctx, pool[0] * pool[1] <= 255, "vl_nnpool: window too large for uint8 indices");
  const int OH = int(OH64), OW = int(OW64), Cp = pad_to(C, 8);
  PoolGeom g{N, H, W, Cp, pool[0], pool[1], stride[0], stride[1], pad[0], pad[2], OH, OW, Cp};
  Arena ar(ctx);
  const float* xd = static_cast<const float*>(ar.in(x->data, numel(x) * 4));
  float* xn = ar.alloc_n<float>(size_t(N) * H * W * Cp);
  float* yn = ar.alloc_n<float>(size_t(N) * OH * OW * Cp);
  uint8_t* idx = method == 0 ? ar.alloc_n<uint8_t>(size_t(N) * OH * OW * Cp) : nullptr;
  if (ar.failed) return ar.finish();
  int rc;
  if ((rc = xemo_op_hwcn_to_nhwc(ctx, xd, H, W, C, N, xn, Cp, 1))) return rc;
  const size_t out8 = size_t(N) * OH * OW * (Cp / 8), in8 = size_t(N) * H * W * (Cp / 8);
  const bool backward = dzdy && dzdy->data;
  if (method == 0) {
    maxpool_fwd_kernel<float, false, 0, 0><<<fixed_channel_grid(out8, Cp / 8, 256, ctx->num_sms, 16), 256, 0, ctx->stream>>>(xn, g, nullptr, nullptr, yn, idx);
    XEMO_LAUNCHED(ctx, 1);
  } else if (!backward) {
    avgpool_fwd_kernel<float><<<grid_for(out8, 256, ctx->num_sms, 16), 256, 0, ctx->stream>>>(xn, g, yn);
    XEMO_LAUNCHED(ctx, 1);
  }
  if (!backward) {
    XEMO_REQUIRE(ctx, y_or_dx->h == OH && y_or_dx->w == OW && y_or_dx->c == C && y_or_dx->n == N,
                 "vl_nnpool: Y must be %d x %d x %d x %d", OH, OW, C, N);
    float* yd = static_cast<float*>(ar.out(y_or_dx->data, numel(y_or_dx) * 4));
    if (ar.failed) return ar.finish();
    if ((rc = xemo_op_nhwc_to_hwcn(ctx, yn, 1, OH, OW, C, N, Cp, yd))) return rc;
    if (argmax && method == 0) {
      const size_t total = size_t(OH) * OW * C * N;
      uint8_t* ad = static_cast<uint8_t*>(ar.out(argmax, total));
      if (ar.failed) return ar.finish();
      nhwc_to_hwcn_u8_kernel<<<grid_for(total, 256, ctx->num_sms), 256, 0, ctx->stream>>>(idx, OH, OW, C, N, Cp, ad);
      XEMO_LAUNCHED(ctx, 1);
    }
    return ar.finish();
  }
  XEMO_REQUIRE(ctx, dzdy->h == OH && dzdy->w == OW && dzdy->c == C && dzdy->n == N, "vl_nnpool: DZDY has the wrong size");
  XEMO_REQUIRE(ctx, y_or_dx->h == H && y_or_dx->w == W && y_or_dx->c == C && y_or_dx->n == N, "vl_nnpool: DX must have the size of X");
  const float* dyd = static_cast<const float*>(ar.in(dzdy->data, numel(dzdy) * 4));
  float* dyn = ar.alloc_n<float>(size_t(N) * OH * OW * Cp);
  float* dxn = ar.alloc_n<float>(size_t(N) * H * W * Cp);
  float* dxd = static_cast<float*>(ar.out(y_or_dx->data, numel(y_or_dx) * 4));
  if (ar.failed) return ar.finish();
  if ((rc = xemo_op_hwcn_to_nhwc(ctx, dyd, OH, OW, C, N, dyn, Cp, 1))) return rc;
  if (method == 0)
    maxpool_bwd_kernel<float, 0, 0><<<fixed_channel_grid(in8, Cp / 8, 256, ctx->num_sms, 16), 256, 0, ctx->stream>>>(dyn, idx, g, dxn);
  else
    avgpool_bwd_kernel<float><<<grid_for(in8, 256, ctx->num_sms, 16), 256, 0, ctx->stream>>>(dyn, g, dxn);
  XEMO_LAUNCHED(ctx, 1);
  if ((rc = xemo_op_nhwc_to_hwcn(ctx, dxn, 1, H, W, C, N, Cp, dxd))) return rc;
  return ar.finish();
}

// ================================================================================================
extern "C" int xemo_vl_nnbnorm(xemo_ctx* ctx, const xemo_array* x, const float* g, const float* b, const xemo_array* dzdy,
                               float epsilon, const float* moments_in, xemo_array* y_or_dx, float* dg, float* db,
                               float* moments_out) {
  XEMO_REQUIRE(ctx, x && x->data && g && b && y_or_dx && y_or_dx->data, "vl_nnbnorm: missing argument");
  XEMO_REQUIRE(ctx, y_or_dx->h == x->h && y_or_dx->w == x->w && y_or_dx->c == x->c && y_or_dx->n == x->n,
               "vl_nnbnorm: output must have the size of X");
  const int H = int(x->h), W = int(x->w), C = int(x->c), N = int(x->n), Cp = pad_to(C, 8);
  const size_t P = size_t(N) * H * W;
  XEMO_REQUIRE(ctx, P > 0, "vl_nnbnorm: empty input");
  Arena ar(ctx);
  const float* xd = static_cast<const float*>(ar.in(x->data, numel(x) * 4));
  float* xn = ar.alloc_n<float>(P * Cp);
  float* gd = padded_vec(ar, g, C, Cp, 1.f);
  float* bd = padded_vec(ar, b, C, Cp);
  float* mom = ar.alloc_n<float>(size_t(2) * Cp, true);
  float* av = ar.alloc_n<float>(Cp);
  float* bv = ar.alloc_n<float>(Cp);
  double* ws = ar.alloc_n<double>(size_t(2) * Cp);
  float* outn = ar.alloc_n<float>(P * Cp);
  float* outd = static_cast<float*>(ar.out(y_or_dx->data, numel(x) * 4));
  if (ar.failed) return ar.finish();
  int rc;
  if ((rc = xemo_op_hwcn_to_nhwc(ctx, xd, H, W, C, N, xn, Cp, 1))) return rc;
  if (moments_in) {
    // C x 2 column-major [mu sigma] -> pitch Cp; padded sigmas = 1
    fill_strided_f32_kernel<<<1, 256, 0, ctx->stream>>>(mom, 1, 0, size_t(Cp), Cp, 1.f);
    XEMO_LAUNCHED(ctx, 1);
    XEMO_CUDA(ctx, cudaMemcpyAsync(mom, moments_in, size_t(C) * 4, cudaMemcpyDefault, ctx->stream));
    XEMO_CUDA(ctx, cudaMemcpyAsync(mom + Cp, moments_in + C, size_t(C) * 4, cudaMemcpyDefault, ctx->stream));
    bn_affine_from_moments_kernel<<<(Cp + 127) / 128, 128, 0, ctx->stream>>>(mom, Cp, gd, bd, nullptr, av, bv);
    XEMO_LAUNCHED(ctx, 1);
  } else {
    if ((rc = bn_stats_launch<float>(ctx, xn, P, Cp, ws))) return rc;
    bn_finalize_kernel<<<(Cp + 127) / 128, 128, 0, ctx->stream>>>(ws, P, Cp, gd, bd, epsilon, mom, av, bv);
    XEMO_LAUNCHED(ctx, 1);
  }
  const int C8 = Cp / 8;
  const int egrid = fixed_channel_grid(P * C8, C8, 256, ctx->num_sms, 16);
  if (!dzdy || !dzdy->data) {
    affine_act_kernel<float><<<egrid, 256, 0, ctx->stream>>>(xn, P, Cp, av, bv, 0, outn);
    XEMO_LAUNCHED(ctx, 1);
  } else {
    XEMO_REQUIRE(ctx, dzdy->h == x->h && dzdy->w == x->w && dzdy->c == x->c && dzdy->n == x->n, "vl_nnbnorm: DZDY must have the size of X");
    const float* dyd = static_cast<const float*>(ar.in(dzdy->data, numel(x) * 4));
    float* dyn = ar.alloc_n<float>(P * Cp);
    if (ar.failed) return ar.finish();
    if ((rc = xemo_op_hwcn_to_nhwc(ctx, dyd, H, W, C, N, dyn, Cp, 1))) return rc;
    XEMO_CUDA(ctx, cudaMemsetAsync(ws, 0, size_t(2) * Cp * sizeof(double), ctx->stream));
    const BnGrid bg = bn_grid(P, Cp, ctx->num_sms);
    dim3 grid(bg.slabs_x, bg.slabs_y);
    PoolGeom g0;
    memset(&g0, 0, sizeof(g0));
    bn_bwd_reduce_kernel<float, false><<<grid, kBnThreads, 0, ctx->stream>>>(xn, dyn, P, Cp, bg.lanes, bg.rows_par, mom, av, bv, 0, ws, nullptr, g0);
    XEMO_LAUNCHED(ctx, 1);
    if (moments_in)
      bn_bwd_test_kernel<float><<<egrid, 256, 0, ctx->stream>>>(xn, dyn, P, Cp, av, bv, 0, outn);
    else
      bn_bwd_apply_kernel<float, false><<<egrid, 256, 0, ctx->stream>>>(xn, dyn, P, Cp, mom, av, bv, 0, ws, outn, nullptr, g0, nullptr, 1.f);
    XEMO_LAUNCHED(ctx, 1);
    if (dg || db) {
      float* dgp = ar.alloc_n<float>(Cp);
      float* dbp = ar.alloc_n<float>(Cp);
      if (ar.failed) return ar.finish();
      bn_bwd_params_kernel<<<(Cp + 127) / 128, 128, 0, ctx->stream>>>(ws, Cp, 1.f, dgp, dbp);
      XEMO_LAUNCHED(ctx, 1);
      if (dg) XEMO_CUDA(ctx, cudaMemcpyAsync(dg, dgp, size_t(C) * 4, cudaMemcpyDefault, ctx->stream));
      if (db) XEMO_CUDA(ctx, cudaMemcpyAsync(db, dbp, size_t(C) * 4, cudaMemcpyDefault, ctx->stream));
    }
  }
  if ((rc = xemo_op_nhwc_to_hwcn(ctx, outn, 1, H, W, C, N, Cp, outd))) return rc;
  if (moments_out) {
    XEMO_CUDA(ctx, cudaMemcpyAsync(moments_out, mom, size_t(C) * 4, cudaMemcpyDefault, ctx->stream));
    XEMO_CUDA(ctx, cudaMemcpyAsync(moments_out + C, mom + Cp, size_t(C) * 4, cudaMemcpyDefault, ctx->stream));
  }
  rc = ar.finish();
  if (rc) return rc;
  // dg / db / moments_out may be host arrays written by cudaMemcpyDefault: make them visible
  if ((dg && !is_device_ptr(dg)) || (db && !is_device_ptr(db)) || (moments_out && !is_device_ptr(moments_out)))
    XEMO_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return XEMO_OK;
}

// ================================================================================================
static int elementwise(xemo_ctx* ctx, const xemo_array* x, const xemo_array* dzdy, xemo_array* out, int kind, float leak) {
  XEMO_REQUIRE(ctx, x && x->data && out && out->data, "missing argument");
  XEMO_REQUIRE(ctx, numel(out) == numel(x) && (!dzdy || !dzdy->data || numel(dzdy) == numel(x)), "size mismatch");
  const size_t n = numel(x);
  Arena ar(ctx);
  const float* xd = static_cast<const float*>(ar.in(x->data, n * 4));
  const float* dyd = (dzdy && dzdy->data) ? static_cast<const float*>(ar.in(dzdy->data, n * 4)) : nullptr;
  float* od = static_cast<float*>(ar.out(out->data, n * 4));
  if (ar.failed) return ar.finish();
  const int grid = grid_for(n, 256, ctx->num_sms, 8);
  if (kind == 0) relu_f32_kernel<<<grid, 256, 0, ctx->stream>>>(xd, dyd, n, leak, od);
  else sigmoid_f32_kernel<<<grid, 256, 0, ctx->stream>>>(xd, dyd, n, od);
  XEMO_LAUNCHED(ctx, 1);
  return ar.finish();
}

extern "C" int xemo_vl_nnrelu(xemo_ctx* ctx, const xemo_array* x, const xemo_array* dzdy, float leak, xemo_array* y_or_dx) {
  return elementwise(ctx, x, dzdy, y_or_dx, 0, leak);
}
extern "C" int xemo_vl_nnsigmoid(xemo_ctx* ctx, const xemo_array* x, const xemo_array* dzdy, xemo_array* y_or_dx) {
  return elementwise(ctx, x, dzdy, y_or_dx, 1, 0.f);
}

extern "C" int xemo_vl_nnsoftmaxt(xemo_ctx* ctx, const xemo_array* x, float temperature, xemo_array* y) {
  XEMO_REQUIRE(ctx, x && x->data && y && y->data && numel(x) == numel(y) && temperature > 0.f, "vl_nnsoftmaxt: bad arguments");
  const size_t n = numel(x);
  const int HW = int(x->h * x->w), C = int(x->c), N = int(x->n);
  Arena ar(ctx);
  const float* xd = static_cast<const float*>(ar.in(x->data, n * 4));
  float* yd = static_cast<float*>(ar.out(y->data, n * 4));
  if (ar.failed) return ar.finish();
  const size_t cols = size_t(HW) * N;
  softmaxt_hwcn_kernel<<<unsigned((cols + 127) / 128), 128, 0, ctx->stream>>>(xd, HW, C, N, 1.f / temperature, yd);
  XEMO_LAUNCHED(ctx, 1);
  return ar.finish();
}

extern "C" int xemo_vl_nnsoftmaxceloss(xemo_ctx* ctx, const xemo_array* x, const xemo_array* p, const float* dzdy,
                                       float temperature, int logit_targets, const float* instance_weights, float* loss,
                                       xemo_array* dx) {
  XEMO_REQUIRE(ctx, x && x->data && p && p->data && numel(x) == numel(p), "vl_nnsoftmaxceloss: X and P must have the same size");
  XEMO_REQUIRE(ctx, x->h == 1 && x->w == 1 && x->c <= kLossMaxC && x->c >= 1, "vl_nnsoftmaxceloss: X must be 1 x 1 x C x N, C <= %d", kLossMaxC);
  XEMO_REQUIRE(ctx, temperature > 0.f, "vl_nnsoftmaxceloss: temperature must be positive");
  XEMO_REQUIRE(ctx, (dzdy && dx && dx->data) || (!dzdy && loss), "vl_nnsoftmaxceloss: forward needs loss, backward needs DZDY and DX");
  const int C = int(x->c), N = int(x->n);
  Arena ar(ctx);
  const float* xd = static_cast<const float*>(ar.in(x->data, numel(x) * 4));
  const float* pd = static_cast<const float*>(ar.in(p->data, numel(p) * 4));
  const float* wd = instance_weights ? static_cast<const float*>(ar.in(instance_weights, size_t(N) * 4)) : nullptr;
  const float* dzd = dzdy ? static_cast<const float*>(ar.in(dzdy, 4)) : nullptr;
  float* dxd = dzdy ? static_cast<float*>(ar.out(dx->data, numel(x) * 4)) : nullptr;
  float* ld = !dzdy ? ar.alloc_n<float>(1, true) : nullptr;
  if (ar.failed) return ar.finish();
  softmaxce_f32_kernel<<<(N + 127) / 128, 128, 0, ctx->stream>>>(xd, pd, wd, N, C, temperature, logit_targets, dzd, dxd, ld);
  XEMO_LAUNCHED(ctx, 1);
  if (ld) {
    XEMO_CUDA(ctx, cudaMemcpyAsync(loss, ld, 4, cudaMemcpyDefault, ctx->stream));
    if (!is_device_ptr(loss)) XEMO_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  }
  return ar.finish();
}

extern "C" int xemo_vl_nnloss_classerror(xemo_ctx* ctx, const xemo_array* x, const float* labels, float* nerr) {
  XEMO_REQUIRE(ctx, x && x->data && labels && nerr && x->h == 1 && x->w == 1, "vl_nnloss_classerror: X must be 1 x 1 x C x N");
  const int C = int(x->c), N = int(x->n);
  Arena ar(ctx);
  const float* xd = static_cast<const float*>(ar.in(x->data, numel(x) * 4));
  const float* ld = static_cast<const float*>(ar.in(labels, size_t(N) * 4));
  float* ed = ar.alloc_n<float>(1, true);
  if (ar.failed) return ar.finish();
  classerror_kernel<<<(N + 127) / 128, 128, 0, ctx->stream>>>(xd, ld, C, N, ed);
  XEMO_LAUNCHED(ctx, 1);
  XEMO_CUDA(ctx, cudaMemcpyAsync(nerr, ed, 4, cudaMemcpyDefault, ctx->stream));
  if (!is_device_ptr(nerr)) XEMO_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return ar.finish();
}

extern "C" int xemo_vl_nnglobalpool(xemo_ctx* ctx, const xemo_array* x, const xemo_array* dzdy, xemo_array* y_or_dx) {
  XEMO_REQUIRE(ctx, x && x->data && y_or_dx && y_or_dx->data, "vl_nnglobalpool: missing argument");
  const int HW = int(x->h * x->w);
  const size_t planes = size_t(x->c) * x->n;
  Arena ar(ctx);
  if (!dzdy || !dzdy->data) {
    XEMO_REQUIRE(ctx, numel(y_or_dx) == planes, "vl_nnglobalpool: Y must be 1 x 1 x C x N");
    const float* xd = static_cast<const float*>(ar.in(x->data, numel(x) * 4));
    float* yd = static_cast<float*>(ar.out(y_or_dx->data, planes * 4));
    if (ar.failed) return ar.finish();
    globalpool_fwd_hwcn_kernel<<<unsigned((planes * 32 + 255) / 256), 256, 0, ctx->stream>>>(xd, HW, planes, yd);
    XEMO_LAUNCHED(ctx, 1);
    return ar.finish();
  }
  XEMO_REQUIRE(ctx, numel(dzdy) == planes && numel(y_or_dx) == numel(x), "vl_nnglobalpool: size mismatch");
  const float* dyd = static_cast<const float*>(ar.in(dzdy->data, planes * 4));
  float* dxd = static_cast<float*>(ar.out(y_or_dx->data, numel(x) * 4));
  if (ar.failed) return ar.finish();
  globalpool_bwd_hwcn_kernel<<<grid_for(numel(x), 256, ctx->num_sms, 8), 256, 0, ctx->stream>>>(dyd, HW, numel(x), dxd);
  XEMO_LAUNCHED(ctx, 1);
  return ar.finish();
}

extern "C" int xemo_vl_nnaxpy(xemo_ctx* ctx, const xemo_array* a, const xemo_array* x, const xemo_array* y, xemo_array* out) {
  XEMO_REQUIRE(ctx, a && a->data && x && x->data && y && y->data && out && out->data, "vl_nnaxpy: missing argument");
  XEMO_REQUIRE(ctx, numel(a) == size_t(x->c) * x->n && numel(y) == numel(x) && numel(out) == numel(x), "vl_nnaxpy: size mismatch");
  const size_t n = numel(x);
  Arena ar(ctx);
  const float* ad = static_cast<const float*>(ar.in(a->data, numel(a) * 4));
  const float* xd = static_cast<const float*>(ar.in(x->data, n * 4));
  const float* yd = static_cast<const float*>(ar.in(y->data, n * 4));
  float* od = static_cast<float*>(ar.out(out->data, n * 4));
  if (ar.failed) return ar.finish();
  axpy_hwcn_kernel<<<grid_for(n, 256, ctx->num_sms, 8), 256, 0, ctx->stream>>>(ad, xd, yd, int(x->h * x->w), n, od);
  XEMO_LAUNCHED(ctx, 1);
  return ar.finish();
}
