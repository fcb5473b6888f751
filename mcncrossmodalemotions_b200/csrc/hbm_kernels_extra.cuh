// hbm_kernels_extra.cuh -- the remaining memory-bound pieces of the hot path: bias gradient, dgrad
// filter packing, boundary-only element-wise operators (sigmoid, leaky ReLU, softmax, class error,
// global pooling backward, Axpy), layout conversion of filters / uint8 indices, and the
// device-hyper-parameter SGD update that CUDA-graph replays need.
#pragma once
#include "hbm_kernels.cuh"

namespace xemo {

// ------------------------------------------------------------------------------------------------
// bias gradient of vl_nnconv: out[c] (+)= scale * sum_p dy[p][c].  dy: [P][ld] of T.
// grid (ceil(C/32), row_blocks): each block reduces a slab of rows for 32 channels; fp32 atomics
// combine slabs (out must be zeroed by the caller; the wrapper does it).
template <typename T>
__global__ void colsum_kernel(const T* __restrict__ dy, size_t P, int ld, int C, float scale, float* __restrict__ out) {
  __shared__ float part[8][33];
  const int c = blockIdx.x * 32 + threadIdx.x;
  const size_t rows_per = (P + gridDim.y - 1) / gridDim.y;
  const size_t r0 = size_t(blockIdx.y) * rows_per;
  const size_t r1 = r0 + rows_per < P ? r0 + rows_per : P;
  float acc = 0.f;
  if (c < C)
    for (size_t r = r0 + threadIdx.y; r < r1; r += 8) acc += float(dy[r * ld + c]);
  part[threadIdx.y][threadIdx.x] = acc;
  __syncthreads();
  if (threadIdx.y == 0 && c < C) {
    float t = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) t += part[j][threadIdx.x];
    atomicAdd(out + c, t * scale);
  }
}

// ------------------------------------------------------------------------------------------------
// Data-gradient filter packing.  Source: KRSC fp16 [Kout][R][S][Cin].  For every output parity class
// (ph, pw) of a stride-(sh, sw) convolution the data gradient is a stride-1 correlation of dY with
//   G[c][j'][i'][k] = F[k][r0 + sh*(Jh-1-j')][s0 + sw*(Jw-1-i')][c],  r0 = (ph+pt) mod sh, s0 = (pw+pl) mod sw
// stored class after class (the class tables are recomputed identically on the host).
struct DgradClass {
  int r0, s0, Jh, Jw;
  long long offset;  // element offset of this class inside the packed buffer
};
struct DgradPackParams {
  DgradClass cls[16];
  int num_classes;
  int Kout, R, S, Cin, sh, sw;
};
static __global__ void dgrad_pack_kernel(const __half* __restrict__ w, DgradPackParams p, __half* __restrict__ dst) {
  const DgradClass c = p.cls[blockIdx.y];
  const size_t total = size_t(p.Cin) * c.Jh * c.Jw * p.Kout;
  for (size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x; i < total; i += size_t(gridDim.x) * blockDim.x) {
    const int k = int(i % p.Kout);
    const int i2 = int((i / p.Kout) % c.Jw);
    const int j2 = int((i / (size_t(p.Kout) * c.Jw)) % c.Jh);
    const int ch = int(i / (size_t(p.Kout) * c.Jw * c.Jh));
    const int r = c.r0 + p.sh * (c.Jh - 1 - j2);
    const int s = c.s0 + p.sw * (c.Jw - 1 - i2);
    dst[c.offset + i] = w[((size_t(k) * p.R + r) * p.S + s) * p.Cin + ch];
  }
}

// Data gradient of a FULL-HEIGHT filter (R == H, one output row, S == 1, no padding: the student's fc6, a 9 x 1 filter over
// a 9 x W map): dx[(n, w), (h, c)] = sum_k dy[(n, w), k] * F[k][h][0][c] is a plain GEMM.  Its filter operand, one row per
// output column (h, c) with k contiguous:  G[h * Cin + c][k] = F[k][h][0][c].
// (for every h a [Kout][Cin] -> [Cin][Kout] transpose through a 32 x 33 shared-memory tile: coalesced on both sides)
static __global__ void dgrad_pack_fullheight_kernel(const __half* __restrict__ w, int Kout, int R, int Cin, __half* __restrict__ dst) {
  __shared__ __half tile[32][33];
  const int h = blockIdx.z, k0 = blockIdx.y * 32, c0 = blockIdx.x * 32;
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const int k = k0 + j, c = c0 + threadIdx.x;
    tile[j][threadIdx.x] = (k < Kout && c < Cin) ? w[(size_t(k) * R + h) * Cin + c] : __float2half(0.f);
  }
  __syncthreads();
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const int c = c0 + j, k = k0 + threadIdx.x;
    if (c < Cin && k < Kout) dst[(size_t(h) * Cin + c) * Kout + k] = tile[threadIdx.x][j];
  }
}

// dF [Kp][R][S][Cp] fp32 (device layout) -> FH x FW x FC x K column-major fp32 (MatConvNet)
static __global__ void krsc_f32_to_filters_kernel(const float* __restrict__ src, int FH, int FW, int FC, int K, int Cp,
                                           float* __restrict__ dst, const float* __restrict__ mul = nullptr) {
  const size_t total = size_t(FH) * FW * FC * K;
  const float m = mul ? *mul : 1.f;
  for (size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x; i < total; i += size_t(gridDim.x) * blockDim.x) {
    const int r = int(i % FH);
    const int s = int((i / FH) % FW);
    const int c = int((i / (size_t(FH) * FW)) % FC);
    const int k = int(i / (size_t(FH) * FW * FC));
    dst[i] = m * src[((size_t(k) * FH + r) * FW + s) * Cp + c];
  }
}

// ------------------------------------------------------------------------------------------------
// Split-operand ("f32x3") convolution staging: the fp32-equivalent mode of xemo_vl_nnconv
// (xemo_set_conv_precision).  A single-precision operand v is written as v*s = hi + lo with hi = fp16(v*s),
// lo = fp16(v*s - hi) (s a power of two taken from the tensor's max |v|, so that nothing under- or overflows in
// fp16); the product of two such operands keeps the three leading terms
//     a*b ~= a_hi*b_hi + a_lo*b_hi + a_hi*b_lo                        (dropped: a_lo*b_lo ~ 2^-22 |a b|)
// which the tcgen05 kernels compute as ONE convolution over a 3x longer reduction axis: the activation-side
// operand is staged as (hi | lo | hi), the filter-side operand as (hi | hi | lo) -- concatenated along the
// channels for the forward / data-gradient convolutions and along the images for the filter gradient (whose
// reduction runs over pixels).  Accumulation is the kernels' usual fp32 in TMEM.
__device__ __forceinline__ float split_pow2_scale(unsigned amax_bits, bool inverse) {
  // s = 2^(13 - floor(log2 amax)): max |v|*s lies in [2^13, 2^14); 1 for an all-zero / denormal / non-finite tensor
  const int e = int((amax_bits >> 23) & 0xffu);
  if (e <= 13 || e == 255) return 1.f;                 // (the same guard for s and 1/s)
  const int se = inverse ? e - 13 : 254 - e + 13;      // biased exponents of 2^(e-127-13) / 2^(127-e+13)
  return __uint_as_float(unsigned(se) << 23);
}
__device__ __forceinline__ void split_f16(float v, __half* hi, __half* lo) {
  const __half h = __float2half_rn(v);
  *hi = h;
  *lo = __float2half_rn(v - __half2float(h));
}

static __global__ void absmax_f32_kernel(const float* __restrict__ x, size_t n, unsigned* __restrict__ amax) {
  float m = 0.f;
  for (size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x; i < n; i += size_t(gridDim.x) * blockDim.x) m = fmaxf(m, fabsf(x[i]));
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0 && m > 0.f) atomicMax(amax, __float_as_uint(m));
}

// HWCN fp32 -> three fp16 NHWC slots: dst[((n*H+h)*W+w)*row_pitch + c + j*slot_stride], j = 0..2, slot j holding the lo
// part when bit j of lo_mask is set.  Channel concatenation: row_pitch = 3*Cp, slot_stride = Cp; image concatenation:
// row_pitch = Cp, slot_stride = N*H*W*Cp.
static __global__ void hwcn_f32_to_nhwc_split_kernel(const float* __restrict__ src, int H, int W, int C, int N, __half* __restrict__ dst,
                                              int Cp, int row_pitch, size_t slot_stride, int lo_mask,
                                              const unsigned* __restrict__ amax) {
  __shared__ float tile[32][33];
  const float sc = split_pow2_scale(*amax, false);
  const int n = blockIdx.z / W, w = blockIdx.z % W;
  const int h0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const int c = c0 + j, h = h0 + threadIdx.x;
    tile[j][threadIdx.x] = (c < C && h < H) ? src[h + size_t(H) * (w + size_t(W) * (c + size_t(C) * n))] * sc : 0.f;
  }
  __syncthreads();
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const int h = h0 + j, c = c0 + threadIdx.x;
    if (h < H && c < Cp) {
      __half hi, lo;
      split_f16(tile[threadIdx.x][j], &hi, &lo);
      __half* d = dst + ((size_t(n) * H + h) * W + w) * row_pitch + c;
#pragma unroll
      for (int s = 0; s < 3; ++s) d[s * slot_stride] = ((lo_mask >> s) & 1) ? lo : hi;
    }
  }
}

// filters FH x FW x FC x K (column-major fp32) -> (hi | hi | lo) fp16 KRSC: along the channels ([Kp][FH][FW][3*Cp],
// forward / filter-side operand of a 3*Cp-channel convolution) or along the output channels ([3*Kp][FH][FW][Cp], what
// the data gradient reduces over)
static __global__ void filters_to_krsc_split_kernel(const float* __restrict__ f, int FH, int FW, int FC, int K, __half* __restrict__ dst,
                                             int Kp, int Cp, int along_k, const unsigned* __restrict__ amax) {
  const float sc = split_pow2_scale(*amax, false);
  const int Cd = along_k ? Cp : 3 * Cp;
  const size_t total = size_t(3) * Kp * FH * FW * Cp;
  for (size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x; i < total; i += size_t(gridDim.x) * blockDim.x) {
    int c = int(i % Cd);
    const int s = int((i / Cd) % FW);
    const int r = int((i / (size_t(Cd) * FW)) % FH);
    int k = int(i / (size_t(Cd) * FW * FH));
    int slot;
    if (along_k) { slot = k / Kp; k %= Kp; } else { slot = c / Cp; c %= Cp; }
    __half hi = __float2half_rn(0.f), lo = hi;
    if (c < FC && k < K) split_f16(f[r + size_t(FH) * (s + size_t(FW) * (c + size_t(FC) * k))] * sc, &hi, &lo);
    dst[i] = slot == 2 ? lo : hi;
  }
}

// v[i] = 1 / (s_a * s_b): undoes the two operand scales (per-output-channel epilogue vector, or one scalar)
static __global__ void split_unscale_kernel(const unsigned* __restrict__ amax_a, const unsigned* __restrict__ amax_b, int n,
                                     float* __restrict__ v) {
  const float inv = split_pow2_scale(*amax_a, true) * split_pow2_scale(*amax_b, true);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) v[i] = inv;
}

// generic (slow-path, boundary only) NHWC <-> HWCN conversion for uint8 index tensors
static __global__ void nhwc_to_hwcn_u8_kernel(const uint8_t* __restrict__ src, int H, int W, int C, int N, int Cp,
                                       uint8_t* __restrict__ dst) {
  const size_t total = size_t(H) * W * C * N;
  for (size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x; i < total; i += size_t(gridDim.x) * blockDim.x) {
    const int h = int(i % H);
    const int w = int((i / H) % W);
    const int c = int((i / (size_t(H) * W)) % C);
    const int n = int(i / (size_t(H) * W * C));
    dst[i] = src[((size_t(n) * H + h) * W + w) * Cp + c];
  }
}

// ------------------------------------------------------------------------------------------------
// boundary element-wise operators on flat fp32 arrays (layout-agnostic)
static __global__ void relu_f32_kernel(const float* __restrict__ x, const float* __restrict__ dy, size_t n, float leak,
                                float* __restrict__ out) {
  for (size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x; i < n; i += size_t(gridDim.x) * blockDim.x) {
    const float v = x[i];
    if (dy) out[i] = dy[i] * (v > 0.f ? 1.f : leak);
    else out[i] = v > 0.f ? v : leak * v;
  }
}
static __global__ void sigmoid_f32_kernel(const float* __restrict__ x, const float* __restrict__ dy, size_t n,
                                   float* __restrict__ out) {
  for (size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x; i < n; i += size_t(gridDim.x) * blockDim.x) {
    const float y = 1.f / (1.f + expf(-x[i]));
    out[i] = dy ? dy[i] * y * (1.f - y) : y;
  }
}
// softmax over the channel dimension of an H x W x C x N column-major array: one thread per (h,w,n)
static __global__ void softmaxt_hwcn_kernel(const float* __restrict__ x, int HW, int C, int N, float invT,
                                     float* __restrict__ y) {
  const size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x;
  if (i >= size_t(HW) * N) return;
  const size_t n = i / HW, hw = i % HW;
  const float* xp = x + n * size_t(C) * HW + hw;
  float* yp = y + n * size_t(C) * HW + hw;
  float m = -INFINITY;
  for (int c = 0; c < C; ++c) m = fmaxf(m, xp[size_t(c) * HW] * invT);
  float s = 0.f;
  for (int c = 0; c < C; ++c) s += expf(xp[size_t(c) * HW] * invT - m);
  for (int c = 0; c < C; ++c) yp[size_t(c) * HW] = expf(xp[size_t(c) * HW] * invT - m) / s;
}
// vl_nnloss 'classerror' on 1 x 1 x C x N logits (column-major == [N][C] row-major); labels 1-based
static __global__ void classerror_kernel(const float* __restrict__ x, const float* __restrict__ labels, int C, int N,
                                  float* __restrict__ nerr) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  float e = 0.f;
  if (n < N) {
    float m = -INFINITY;
    int a = 0;
    for (int c = 0; c < C; ++c) {
      const float v = x[size_t(n) * C + c];
      if (v > m) { m = v; a = c; }
    }
    e = (a + 1 != int(labels[n])) ? 1.f : 0.f;
  }
  e = warp_sum(e);
  if ((threadIdx.x & 31) == 0 && e != 0.f) atomicAdd(nerr, e);
}
// T-softmax CE on fp32 [N][C] logits (boundary variant of softmaxce_fused_kernel: fp32 in / out)
static __global__ void softmaxce_f32_kernel(const float* __restrict__ x, const float* __restrict__ t,
                                     const float* __restrict__ w, int N, int C, float T, int logit_targets,
                                     const float* __restrict__ dzdy, float* __restrict__ dx, float* __restrict__ loss) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  float l = 0.f;
  if (n < N) {
    float xv[kLossMaxC], tv[kLossMaxC];
    float xm = -INFINITY, tm = -INFINITY;
    for (int c = 0; c < C; ++c) {
      xv[c] = x[size_t(n) * C + c];
      tv[c] = t[size_t(n) * C + c];
      xm = fmaxf(xm, xv[c]);
      tm = fmaxf(tm, tv[c]);
    }
    const float invT = 1.f / T;
    float xs = 0.f, ts = 0.f;
    for (int c = 0; c < C; ++c) {
      xv[c] = (xv[c] - xm) * invT;
      xs += expf(xv[c]);
      if (logit_targets) { tv[c] = expf((tv[c] - tm) * invT); ts += tv[c]; }
    }
    const float lse = logf(xs);
    const float wn = w ? w[n] : 1.f;
    const float dz = dzdy ? dzdy[0] : 1.f;
    for (int c = 0; c < C; ++c) {
      const float p = logit_targets ? tv[c] / ts : tv[c];
      const float logq = xv[c] - lse;
      l -= p * logq;
      if (dx) dx[size_t(n) * C + c] = dz * wn * (expf(logq) - p) * invT;
    }
    l *= wn;
  }
  l = warp_sum(l);
  if (loss && (threadIdx.x & 31) == 0) atomicAdd(loss, l);
}

// global average pooling backward: dx[h,w,c,n] = dy[c,n] / (H*W)   (HWCN column-major, flat)
static __global__ void globalpool_bwd_hwcn_kernel(const float* __restrict__ dy, int HW, size_t total, float* __restrict__ dx) {
  const float inv = 1.f / float(HW);
  for (size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x; i < total; i += size_t(gridDim.x) * blockDim.x)
    dx[i] = dy[i / HW] * inv;
}
// global average pooling forward on HWCN column-major: one warp per (c,n) plane
static __global__ void globalpool_fwd_hwcn_kernel(const float* __restrict__ x, int HW, size_t planes, float* __restrict__ y) {
  const size_t warp = (blockIdx.x * size_t(blockDim.x) + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= planes) return;
  float acc = 0.f;
  for (int i = lane; i < HW; i += 32) acc += x[warp * HW + i];
  acc = warp_sum(acc);
  if (lane == 0) y[warp] = acc / float(HW);
}
// Axpy on HWCN column-major: out = a[c,n] * x + y
static __global__ void axpy_hwcn_kernel(const float* __restrict__ a, const float* __restrict__ x, const float* __restrict__ y,
                                 int HW, size_t total, float* __restrict__ out) {
  for (size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x; i < total; i += size_t(gridDim.x) * blockDim.x)
    out[i] = fmaf(a[i / HW], x[i], y[i]);
}

static __global__ void fill_strided_f32_kernel(float* __restrict__ dst, int outer, size_t outer_stride, size_t inner_off,
                                        int inner, float value) {
  const size_t total = size_t(outer) * inner;
  for (size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x; i < total; i += size_t(gridDim.x) * blockDim.x)
    dst[(i / inner) * outer_stride + inner_off + (i % inner)] = value;
}

// ------------------------------------------------------------------------------------------------
// Face preprocessing fused into the teacher stem staging (emoVoxCeleb/fetch_emovoxceleb_imdb.m:175-193,
// teacher/ferplus_baselines.m:203-213): uint8 grey IH x IW x N (column-major) -> replicate to 3 channels,
// single, subtract averageImage[c], bilinear resize to OHt x OWt (corner-aligned grid, as the identity
// vl_nnaffinegrid + vl_nnbilinearsampler pair) -> row-im2col tensor Xr[n][h][ow][s*4+c] fp16 for the
// 7x1 tcgen05 stem.  One thread per (n, h, ow).
static __global__ void face_u8_rows_im2col_kernel(const uint8_t* __restrict__ src, int IH, int IW, int N, int OHt, int OWt,
                                                  const float* __restrict__ mean, int S, int stride_w, int pad_l, int OW,
                                                  __half* __restrict__ dst) {
  const size_t total = size_t(N) * OHt * OW;
  const float sy = OHt > 1 ? float(IH - 1) / float(OHt - 1) : 0.f;
  const float sx = OWt > 1 ? float(IW - 1) / float(OWt - 1) : 0.f;
  const float m0 = mean[0], m1 = mean[1], m2 = mean[2];
  for (size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x; i < total; i += size_t(gridDim.x) * blockDim.x) {
    const int ow = int(i % OW);
    const int h = int((i / OW) % OHt);
    const int n = int(i / (size_t(OW) * OHt));
    const uint8_t* img = src + size_t(n) * IH * IW;
    const float ys = h * sy;
    int y0 = int(floorf(ys));
    y0 = min(max(y0, 0), IH - 1);
    const int y1 = min(y0 + 1, IH - 1);
    const float wy = ys - float(y0);
    __align__(16) __half v[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = __float2half_rn(0.f);
#pragma unroll
    for (int s = 0; s < 8; ++s) {   // S <= 8 (host check); compile-time trip count keeps v[] in registers
      const int w = ow * stride_w + s - pad_l;
      if (s >= S || w < 0 || w >= OWt) continue;
      const float xs = w * sx;
      int x0 = int(floorf(xs));
      x0 = min(max(x0, 0), IW - 1);
      const int x1 = min(x0 + 1, IW - 1);
      const float wx = xs - float(x0);
      const float top = float(img[y0 + IH * x0]) * (1.f - wx) + float(img[y0 + IH * x1]) * wx;
      const float bot = float(img[y1 + IH * x0]) * (1.f - wx) + float(img[y1 + IH * x1]) * wx;
      const float g = top * (1.f - wy) + bot * wy;
      v[s * 4 + 0] = __float2half_rn(g - m0);
      v[s * 4 + 1] = __float2half_rn(g - m1);
      v[s * 4 + 2] = __float2half_rn(g - m2);
    }
    uint4* o = reinterpret_cast<uint4*>(dst + ((size_t(n) * OHt + h) * OW + ow) * 32);
    const uint4* vi = reinterpret_cast<const uint4*>(v);
    o[0] = vi[0]; o[1] = vi[1]; o[2] = vi[2]; o[3] = vi[3];
  }
}

// ------------------------------------------------------------------------------------------------
// Spectrogram front-end (VGGVox `runSpec`, called at emoVoxCeleb/getBatchEmoVoxCeleb.m:162 and
// external/compute_audio_feats.m:176 with the constants of emoVoxCeleb/run_distillation.m:109-117):
//   y[n] = x[n] - alpha * x[n-1]  (pre-emphasis, y[0] = x[0]) ; frames of Nw samples every Ns samples, no padding ;
//   symmetric Hamming window ; |FFT_nfft| with nfft = 2^nextpow2(Nw) -> nfft rows (the full spectrum).
// One block per (frame, clip): the windowed frame is zero-padded to nfft in shared memory and transformed by an
// in-place radix-2 FFT (fp32).  Output layout: nfft x W x 1 x N column-major (what the student graph consumes).
static __global__ void spectrogram_kernel(const float* __restrict__ wav, int L, int Nw, int Ns, int nfft, int log2n,
                                          float alpha, float scale, int W, float* __restrict__ spec) {
  extern __shared__ float2 fft_buf[];  // [nfft]
  const int t = blockIdx.x, n = blockIdx.y;
  const float* x = wav + size_t(n) * L + size_t(t) * Ns;
  for (int i = threadIdx.x; i < nfft; i += blockDim.x) {
    float v = 0.f;
    if (i < Nw) {
      const size_t g = size_t(t) * Ns + i;  // index into the clip
      const float cur = x[i] * scale;
      const float prev = g > 0 ? x[i - 1] * scale : 0.f;
      const float wgt = 0.54f - 0.46f * cospif(2.f * float(i) / float(Nw - 1));
      v = (cur - alpha * prev) * wgt;
    }
    fft_buf[__brev(unsigned(i)) >> (32 - log2n)] = make_float2(v, 0.f);  // bit-reversed order for the DIT butterflies
  }
  __syncthreads();
  for (int s = 1; s <= log2n; ++s) {
    const int half = 1 << (s - 1);
    for (int k = threadIdx.x; k < nfft / 2; k += blockDim.x) {
      const int j = k & (half - 1);
      const int i0 = ((k >> (s - 1)) << s) + j, i1 = i0 + half;
      float sn, cs;
      sincospif(-float(j) / float(half), &sn, &cs);  // exp(-2 pi i j / 2^s)
      const float2 a = fft_buf[i0], b = fft_buf[i1];
      const float2 tw = make_float2(b.x * cs - b.y * sn, b.x * sn + b.y * cs);
      fft_buf[i0] = make_float2(a.x + tw.x, a.y + tw.y);
      fft_buf[i1] = make_float2(a.x - tw.x, a.y - tw.y);
    }
    __syncthreads();
  }
  float* out = spec + (size_t(n) * W + t) * nfft;
  for (int i = threadIdx.x; i < nfft; i += blockDim.x) out[i] = sqrtf(fft_buf[i].x * fft_buf[i].x + fft_buf[i].y * fft_buf[i].y);
}

// per frequency row (x - mean) / std over time, std with the N-1 normalisation (getBatchEmoVoxCeleb.m:164-169).
// One block per clip, one thread per row (coalesced across rows), two-pass variance.  In place.
static __global__ void spec_rownorm_kernel(float* __restrict__ spec, int H, int W) {
  float* s = spec + size_t(blockIdx.x) * H * W;
  for (int f = threadIdx.x; f < H; f += blockDim.x) {
    float sum = 0.f;
    for (int t = 0; t < W; ++t) sum += s[f + size_t(t) * H];
    const float mu = sum / float(W);
    float ss = 0.f;
    for (int t = 0; t < W; ++t) { const float d = s[f + size_t(t) * H] - mu; ss = fmaf(d, d, ss); }
    const float inv = rsqrtf(ss / float(W - 1));
    for (int t = 0; t < W; ++t) s[f + size_t(t) * H] = (s[f + size_t(t) * H] - mu) * inv;
  }
}

// ------------------------------------------------------------------------------------------------
// cnn_train_dag update with the hyper-parameters in device memory (hyper = {lr, momentum, wd, 1/B}),
// so that a captured CUDA graph follows the learning-rate schedule without re-capture.
static __global__ void sgd_momentum_dev_kernel(float* __restrict__ w, float* __restrict__ m, const float* __restrict__ g,
                                        size_t n, const float* __restrict__ hyper, float lr_mult, float wd_mult,
                                        float inv_grad_scale, __half* __restrict__ w16, const int* __restrict__ guard = nullptr) {
  if (guard && guard[0]) return;   // non-finite gradient this step (grad_guard_kernel): leave w, m and the fp16 mirror alone
  const float lr = hyper[0] * lr_mult, momentum = hyper[1], wd = hyper[2] * wd_mult, inv_batch = hyper[3];
  for (size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x; i < n; i += size_t(gridDim.x) * blockDim.x) {
    const float wi = w[i];
    const float mi = momentum * m[i] - (wd * wi + g[i] * inv_grad_scale * inv_batch);
    const float wn = wi + lr * mi;
    m[i] = mi;
    w[i] = wn;
    if (w16) w16[i] = __float2half_rn(wn);
  }
}

// Overflow guard of the fp16 gradient chain (fixed loss scale): state[0] <- 1 when any element of the flat gradient is
// inf / NaN (0 otherwise), state[1] += that (skipped steps so far).  Two launches: scan, then publish.
static __global__ void grad_guard_scan_kernel(const float* __restrict__ g, size_t n, int* __restrict__ state) {
  int bad = 0;
  for (size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x; i < n; i += size_t(gridDim.x) * blockDim.x)
    bad |= !isfinite(g[i]);
  if (__any_sync(0xffffffffu, bad) && (threadIdx.x & 31) == 0) atomicOr(&state[2], 1);
}
static __global__ void grad_guard_publish_kernel(int* __restrict__ state) {
  state[0] = state[2];
  state[1] += state[2];
  state[2] = 0;
}

// test-mode BN backward: dx = a * dz, dz = dy * [a*x+b > 0] when relu_mask
template <typename T>
__global__ void bn_bwd_test_kernel(const T* __restrict__ x, const T* __restrict__ dy, size_t P, int C,
                                   const float* __restrict__ a, const float* __restrict__ b, int relu_mask,
                                   T* __restrict__ dx) {
  const int C8 = C >> 3;
  const size_t total = P * C8;
  const size_t tid = blockIdx.x * size_t(blockDim.x) + threadIdx.x;
  const int c8 = int(tid % C8);  // constant along the loop (fixed_channel_grid)
  float av[8], bv[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) { av[k] = a[c8 * 8 + k]; bv[k] = b[c8 * 8 + k]; }
  for (size_t i = tid; i < total; i += size_t(gridDim.x) * blockDim.x) {
    Vec8<T> vx, vd;
    vx.load(x + i * 8);
    vd.load(dy + i * 8);
    float fx[8], fd[8];
    vx.to_float(fx);
    vd.to_float(fd);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      float dz = fd[k];
      if (relu_mask && !(fmaf(av[k], fx[k], bv[k]) > 0.f)) dz = 0.f;
      fd[k] = av[k] * dz;
    }
    vd.from_float(fd);
    vd.store(dx + i * 8);
  }
}

}  // namespace xemo
