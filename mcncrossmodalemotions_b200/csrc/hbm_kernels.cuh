// hbm_kernels.cuh -- the memory-bound operators of the distillation hot path as coalesced,
// vectorised NHWC kernels (fp16 storage, fp32 math, warp-shuffle reductions).
//
// Operator semantics follow MatConvNet / mcnExtraLayers; the reference reaches
// them through dagnn blocks inside dag.eval (emoVoxCeleb/fetch_emovoxceleb_imdb.m:129,
// external/compute_visual_feats.m:90, external/compute_audio_feats.m:126) and cnn_train_dag
// (emoVoxCeleb/run_distillation.m:170).
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace xemo {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Eight consecutive channels of storage type T (fp16 inside the fused graphs, fp32 at the MatConvNet
// boundary operators, where results must not be re-quantised).
template <typename T>
struct Vec8;
template <>
struct Vec8<float> {
  float4 a, b;
  __device__ __forceinline__ void load(const float* p) {
    a = *reinterpret_cast<const float4*>(p);
    b = *reinterpret_cast<const float4*>(p + 4);
  }
  __device__ __forceinline__ void store(float* p) const {
    *reinterpret_cast<float4*>(p) = a;
    *reinterpret_cast<float4*>(p + 4) = b;
  }
  __device__ __forceinline__ void to_float(float (&f)[8]) const {
    f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
  }
  __device__ __forceinline__ void from_float(const float (&f)[8]) {
    a = make_float4(f[0], f[1], f[2], f[3]);
    b = make_float4(f[4], f[5], f[6], f[7]);
  }
};
template <>
struct Vec8<__half> {
  uint4 u;
  __device__ __forceinline__ void load(const __half* p) { u = *reinterpret_cast<const uint4*>(p); }
  __device__ __forceinline__ void store(__half* p) const { *reinterpret_cast<uint4*>(p) = u; }
  __device__ __forceinline__ void to_float(float (&f)[8]) const {
    const __half2* h = reinterpret_cast<const __half2*>(&u);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float2 t = __half22float2(h[i]);
      f[2 * i] = t.x;
      f[2 * i + 1] = t.y;
    }
  }
  __device__ __forceinline__ void from_float(const float (&f)[8]) {
    __half2* h = reinterpret_cast<__half2*>(&u);
#pragma unroll
    for (int i = 0; i < 4; ++i) h[i] = __floats2half2_rn(f[2 * i], f[2 * i + 1]);
  }
};
typedef Vec8<__half> Half8;
// value as it would be stored in T (the fused pooling read must compare what a separate pass stores)
template <typename T>
__device__ __forceinline__ float round_to(float v);
template <>
__device__ __forceinline__ float round_to<float>(float v) { return v; }
template <>
__device__ __forceinline__ float round_to<__half>(float v) { return __half2float(__float2half_rn(v)); }
template <typename T>
__device__ __forceinline__ T from_f32(float v);
template <>
__device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <>
__device__ __forceinline__ __half from_f32<__half>(float v) { return __float2half_rn(v); }

// ============================================================================================
// Layout conversion at the MatConvNet boundary.  MatConvNet: H x W x C x N column-major fp32
// (index h + H*(w + W*(c + C*n))).  Device-native: NHWC fp16, channel pitch Cp >= C (zero padded).
// ============================================================================================
template <typename T>
__global__ void hwcn_f32_to_nhwc_kernel(const float* __restrict__ src, int H, int W, int C, int N,
                                            T* __restrict__ dst, int Cp) {
  // tile transpose over (h, c) for fixed (n, w): reads coalesced along h, writes coalesced along c
  __shared__ float tile[32][33];
  const int n = blockIdx.z / W, w = blockIdx.z % W;
  const int h0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const int c = c0 + j, h = h0 + threadIdx.x;
    tile[j][threadIdx.x] = (c < C && h < H) ? src[h + size_t(H) * (w + size_t(W) * (c + size_t(C) * n))] : 0.f;
  }
  __syncthreads();
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const int h = h0 + j, c = c0 + threadIdx.x;
    if (h < H && c < Cp) dst[((size_t(n) * H + h) * W + w) * Cp + c] = from_f32<T>(tile[threadIdx.x][j]);
  }
}

template <typename TSrc>
__global__ void nhwc_to_hwcn_f32_kernel(const TSrc* __restrict__ src, int H, int W, int C, int N, int Cp,
                                        float* __restrict__ dst) {
  __shared__ float tile[32][33];
  const int n = blockIdx.z / W, w = blockIdx.z % W;
  const int h0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const int h = h0 + j, c = c0 + threadIdx.x;
    tile[j][threadIdx.x] = (h < H && c < C) ? float(src[((size_t(n) * H + h) * W + w) * Cp + c]) : 0.f;
  }
  __syncthreads();
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const int c = c0 + j, h = h0 + threadIdx.x;
    if (c < C && h < H) dst[h + size_t(H) * (w + size_t(W) * (c + size_t(C) * n))] = tile[threadIdx.x][j];
  }
}

// filters FH x FW x FC x K (column-major fp32) -> [Kp][FH][FW][Cp] fp16 (zero padded), optionally
// flipped+transposed for dgrad: dst[c][FH-1-r][FW-1-s][k].
static __global__ void filters_to_krsc_f16_kernel(const float* __restrict__ f, int FH, int FW, int FC, int K,
                                           __half* __restrict__ dst, int Kp, int Cp, int flip_transpose) {
  const size_t total = flip_transpose ? size_t(Cp) * FH * FW * Kp : size_t(Kp) * FH * FW * Cp;
  for (size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x; i < total; i += size_t(gridDim.x) * blockDim.x) {
    float v = 0.f;
    if (!flip_transpose) {
      const int c = int(i % Cp);
      const int s = int((i / Cp) % FW);
      const int r = int((i / (size_t(Cp) * FW)) % FH);
      const int k = int(i / (size_t(Cp) * FW * FH));
      if (c < FC && k < K) v = f[r + size_t(FH) * (s + size_t(FW) * (c + size_t(FC) * k))];
    } else {
      const int k = int(i % Kp);
      const int s2 = int((i / Kp) % FW);
      const int r2 = int((i / (size_t(Kp) * FW)) % FH);
      const int c = int(i / (size_t(Kp) * FW * FH));
      const int r = FH - 1 - r2, s = FW - 1 - s2;
      if (c < FC && k < K) v = f[r + size_t(FH) * (s + size_t(FW) * (c + size_t(FC) * k))];
    }
    dst[i] = __float2half_rn(v);
  }
}

// ============================================================================================
// Teacher input: faces 224x224x3xN (HWCN fp32, already mean-subtracted) -> "row-im2col" tensor
// Xr[n][h][ow][32] with Xr[.., s*4+c] = x[h, 2*ow + s - pad_l, c]  (s < S, c < C; zero elsewhere), so
// that the 7x7/2 stem becomes a 7x1 convolution with 32 input channels for the tcgen05 kernel.
// Generic in (S, stride_w, pad_l, C<=4).
// ============================================================================================
static __global__ void rows_im2col_from_hwcn_kernel(const float* __restrict__ src, int H, int W, int C, int N, int S,
                                             int stride_w, int pad_l, int OW, __half* __restrict__ dst) {
  // one thread per (n, h, ow): writes 32 halves (64 B)
  const size_t total = size_t(N) * H * OW;
  for (size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x; i < total; i += size_t(gridDim.x) * blockDim.x) {
    const int h = int(i % H);  // h fastest so that the HWCN source is read coalesced
    const int ow = int((i / H) % OW);
    const int n = int(i / (size_t(H) * OW));
    __align__(16) __half v[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = __float2half_rn(0.f);
#pragma unroll
    for (int s = 0; s < 8; ++s) {   // S <= 8, C <= 4 (host check): compile-time trip counts keep v[] in registers
      const int w = ow * stride_w + s - pad_l;
      if (s >= S || w < 0 || w >= W) continue;
#pragma unroll
      for (int c = 0; c < 4; ++c)
        if (c < C) v[s * 4 + c] = __float2half_rn(src[h + size_t(H) * (w + size_t(W) * (c + size_t(C) * n))]);
    }
    uint4* o = reinterpret_cast<uint4*>(dst + ((size_t(n) * H + h) * OW + ow) * 32);
    const uint4* vi = reinterpret_cast<const uint4*>(v);
    o[0] = vi[0]; o[1] = vi[1]; o[2] = vi[2]; o[3] = vi[3];
  }
}

// Student input: spectrograms H x W x 1 x N (HWCN fp32) -> space-to-depth tensor
// Y[n][hp][ow][16] with Y[.., dr*8+s] = x[2*hp + dr - pad_t, 2*ow + s - pad_l]  (dr<2, s<7), so that the
// 7x7/2 stem becomes a 4x1 stride-1 convolution with 16 input channels (K = 64).
static __global__ void spec_s2d_from_hwcn_kernel(const float* __restrict__ src, int H, int W, int N, int pad_t, int pad_l,
                                          int HP, int OW, __half* __restrict__ dst) {
  const size_t total = size_t(N) * HP * OW;
  for (size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x; i < total; i += size_t(gridDim.x) * blockDim.x) {
    const int hp = int(i % HP);
    const int ow = int((i / HP) % OW);
    const int n = int(i / (size_t(HP) * OW));
    __align__(16) __half v[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] = __float2half_rn(0.f);
#pragma unroll
    for (int dr = 0; dr < 2; ++dr) {
      const int h = 2 * hp + dr - pad_t;
      if (h < 0 || h >= H) continue;
#pragma unroll
      for (int s = 0; s < 7; ++s) {
        const int w = 2 * ow + s - pad_l;
        if (w < 0 || w >= W) continue;
        v[dr * 8 + s] = __float2half_rn(src[h + size_t(H) * (w + size_t(W) * size_t(n))]);
      }
    }
    uint4* o = reinterpret_cast<uint4*>(dst + ((size_t(n) * HP + hp) * OW + ow) * 16);
    const uint4* vi = reinterpret_cast<const uint4*>(v);
    o[0] = vi[0]; o[1] = vi[1];
  }
}

// ============================================================================================
// Max pooling (vl_nnpool 'max'), NHWC fp16, 8 channels per thread.  Optional fused input transform
// z = relu(a[c]*x + b[c]) (train/test-mode BN + ReLU folded into the pooling read).  Emits the
// window-local arg-max idx = dw*PH + dh (first maximum in MatConvNet's memory-order scan: w outer,
// h inner, strict '>'), which the backward pass consumes -- the "bit-exact pooling indices".
// Padding acts as -inf.
// ============================================================================================
struct PoolGeom {
  int N, H, W, C;       // input
  int PH, PW, sh, sw, pt, pl;
  int OH, OW;
  int OC;               // channel pitch of the pooled-side tensors (y / arg-max / winner / dy): C, or larger when the
                        // consumer wants zero-padded channels (fp16 fast paths only).  Last member: aggregate
                        // initialisers that stop at OW leave it 0 and only reach kernels that ignore it.
};

// Grid for kernels whose threads keep per-channel coefficients in registers: the total thread count is
// a multiple of C8 (= C/8 channel groups), so that a grid-stride loop over (row, channel-group) items
// always revisits the same channel group.
inline int gcd_int(int a, int b) { return b ? gcd_int(b, a % b) : a; }
inline int fixed_channel_grid(size_t items, int C8, int threads, int num_sms, int per_sm) {
  size_t blocks = (items + threads - 1) / threads;
  const size_t cap = size_t(num_sms) * per_sm;
  const int m = C8 / gcd_int(C8, threads);
  // capped grids round DOWN to the multiple: callers pass the number of blocks that are resident at once, and two blocks
  // past it (592 -> 594 on the 96-channel stem) form a second wave that costs 20 % of the kernel
  if (blocks > cap) return int(cap >= size_t(m) ? cap / m * m : m);
  if (blocks < 1) blocks = 1;
  return int((blocks + m - 1) / m * m);
}

// PH / PW == 0 selects the run-time window of `g` (generic path); the 3x3 and 5x3 windows of the two
// networks are compile-time so that the window loads are issued together.
template <typename T, bool kAffineRelu, int PHc, int PWc>
__global__ void maxpool_fwd_kernel(const T* __restrict__ x, PoolGeom g, const float* __restrict__ a,
                                   const float* __restrict__ b, T* __restrict__ y, uint8_t* __restrict__ idx) {
  const int PH = PHc ? PHc : g.PH, PW = PWc ? PWc : g.PW;
  // 32-bit index arithmetic: thread -> fixed channel group c8, grid-stride loop over output pixels
  const uint32_t C8 = uint32_t(g.C >> 3);
  const uint32_t npix = uint32_t(g.N) * uint32_t(g.OH) * uint32_t(g.OW);
  const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t c8 = tid % C8;  // constant along the loop (fixed_channel_grid)
  const uint32_t pstride = (gridDim.x * blockDim.x) / C8;
  float av[8], bv[8];
  if (kAffineRelu) {
#pragma unroll
    for (int k = 0; k < 8; ++k) { av[k] = a[c8 * 8 + k]; bv[k] = b[c8 * 8 + k]; }
  }
  for (uint32_t pp = tid / C8; pp < npix; pp += pstride) {
    const uint32_t t1 = pp / uint32_t(g.OW);
    const int ow = int(pp - t1 * uint32_t(g.OW));
    const int n = int(t1 / uint32_t(g.OH));
    const int oh = int(t1 - uint32_t(n) * uint32_t(g.OH));
    float best[8];
    uint8_t arg[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) { best[k] = -INFINITY; arg[k] = 0; }
    const T* xn = x + size_t(n) * g.H * g.W * g.C + c8 * 8;
    const int h0 = oh * g.sh - g.pt, w0 = ow * g.sw - g.pl;
#pragma unroll
    for (int dw = 0; dw < PW; ++dw) {
      const int w = w0 + dw;
#pragma unroll
      for (int dh = 0; dh < PH; ++dh) {
        const int h = h0 + dh;
        if (w < 0 || w >= g.W || h < 0 || h >= g.H) continue;
        Vec8<T> v;
        v.load(xn + (size_t(h) * g.W + w) * g.C);
        float f[8];
        v.to_float(f);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          float z = f[k];
          // the window maximum is taken over the fp32 normalised activations (as the CPU path does) and
          // rounded once on store; ties -- the all-zero windows ReLU produces -- stay exact ties
          if (kAffineRelu) z = fmaxf(fmaf(av[k], z, bv[k]), 0.f);
          if (z > best[k]) { best[k] = z; arg[k] = uint8_t(dw * PH + dh); }
        }
      }
    }
    Vec8<T> o;
    o.from_float(best);
    const size_t off = ((size_t(n) * g.OH + oh) * g.OW + ow) * g.C + c8 * 8;
    o.store(y + off);
    if (idx) {
      uint2 pk;
      pk.x = arg[0] | (arg[1] << 8) | (arg[2] << 16) | (uint32_t(arg[3]) << 24);
      pk.y = arg[4] | (arg[5] << 8) | (arg[6] << 16) | (uint32_t(arg[7]) << 24);
      *reinterpret_cast<uint2*>(idx + off) = pk;
    }
  }
}

// Backward of max pooling as a gather: every input position collects dy from the windows whose
// recorded arg-max points at it (no atomics, deterministic).  MH / MW = max windows covering one
// input position along h / w (ceil(P/stride)); 0 selects run-time loops.
template <typename T, int MH, int MW>
__global__ void maxpool_bwd_kernel(const T* __restrict__ dy, const uint8_t* __restrict__ idx, PoolGeom g,
                                   T* __restrict__ dx) {
  const uint32_t C8 = uint32_t(g.C >> 3);
  const uint32_t npix = uint32_t(g.N) * uint32_t(g.H) * uint32_t(g.W);
  const int mh = MH ? MH : (g.PH + g.sh - 1) / g.sh, mw = MW ? MW : (g.PW + g.sw - 1) / g.sw;
  const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t c8 = tid % C8;  // constant along the loop (fixed_channel_grid)
  const uint32_t pstride = (gridDim.x * blockDim.x) / C8;
  for (uint32_t pp = tid / C8; pp < npix; pp += pstride) {
    const uint32_t t1 = pp / uint32_t(g.W);
    const int w = int(pp - t1 * uint32_t(g.W));
    const int n = int(t1 / uint32_t(g.H));
    const int h = int(t1 - uint32_t(n) * uint32_t(g.H));
    float acc[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[k] = 0.f;
    // windows (oh, ow) covering (h, w): oh*sh - pt <= h < oh*sh - pt + PH
    const int oh_hi = (h + g.pt) / g.sh, ow_hi = (w + g.pl) / g.sw;
#pragma unroll
    for (int ia = 0; ia < (MH ? MH : mh); ++ia) {
      const int oh = oh_hi - ia;
      const int dh = h + g.pt - oh * g.sh;
      if (oh < 0 || oh >= g.OH || dh >= g.PH) continue;
#pragma unroll
      for (int ib = 0; ib < (MW ? MW : mw); ++ib) {
        const int ow = ow_hi - ib;
        const int dw = w + g.pl - ow * g.sw;
        if (ow < 0 || ow >= g.OW || dw >= g.PW) continue;
        const size_t off = ((size_t(n) * g.OH + oh) * g.OW + ow) * g.C + c8 * 8;
        const uint2 pk = *reinterpret_cast<const uint2*>(idx + off);
        Vec8<T> v;
        v.load(dy + off);
        float f[8];
        v.to_float(f);
        const uint32_t me = uint32_t(dw * g.PH + dh);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const uint32_t a = ((k < 4 ? pk.x : pk.y) >> (8 * (k & 3))) & 0xFF;
          if (a == me) acc[k] += f[k];
        }
      }
    }
    Vec8<T> o;
    o.from_float(acc);
    o.store(dx + ((size_t(n) * g.H + h) * g.W + w) * g.C + c8 * 8);
  }
}

// ------------------------------------------------------------------------------------------------
// fp16 fast paths of the two pooling kernels (the generic templates above are instruction-bound: ~700
// instructions per 8-channel item).  Forward: z = relu(a*x + b) is monotone in x (increasing for a > 0,
// decreasing for a < 0), so the window arg-max is found on sign-adjusted raw fp16 values with packed
// half2 compares, the affine is applied once to the winner, and the ties ReLU creates (all-zero window)
// resolve to the first in-bounds element as in MatConvNet's scan.  Backward: the uint8 arg-max bytes are
// compared four at a time and the selected gradients accumulated in half2.
template <bool kAffineRelu, int PH, int PW, bool kNoPad>
__global__ void maxpool_fwd_h2_kernel(const __half* __restrict__ x, PoolGeom g, const float* __restrict__ a,
                                      const float* __restrict__ b, __half* __restrict__ y, uint8_t* __restrict__ idx,
                                      __half* __restrict__ xwin) {
  const uint32_t C8 = uint32_t(g.C >> 3);
  const uint32_t npix = uint32_t(g.N) * uint32_t(g.OH) * uint32_t(g.OW);
  const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t c8 = tid % C8;  // constant along the loop (fixed_channel_grid)
  const uint32_t pstride = (gridDim.x * blockDim.x) / C8;
  float av[8], bv[8];
  __half2 sgn[4];
  if (kAffineRelu) {
#pragma unroll
    for (int k = 0; k < 8; ++k) { av[k] = a[c8 * 8 + k]; bv[k] = b[c8 * 8 + k]; }
#pragma unroll
    for (int k = 0; k < 4; ++k) sgn[k] = __floats2half2_rn(av[2 * k] < 0.f ? -1.f : 1.f, av[2 * k + 1] < 0.f ? -1.f : 1.f);
  }
  const __half2 ninf = __floats2half2_rn(-INFINITY, -INFINITY);
  for (uint32_t pp = tid / C8; pp < npix; pp += pstride) {
    const uint32_t t1 = pp / uint32_t(g.OW);
    const int ow = int(pp - t1 * uint32_t(g.OW));
    const int n = int(t1 / uint32_t(g.OH));
    const int oh = int(t1 - uint32_t(n) * uint32_t(g.OH));
    const __half* xn = x + size_t(n) * g.H * g.W * g.C + c8 * 8;
    const int h0 = oh * g.sh - g.pt, w0 = ow * g.sw - g.pl;
    __half2 best[4] = {ninf, ninf, ninf, ninf};
    uint32_t arg[4] = {0u, 0u, 0u, 0u};  // two 16-bit lanes per word
    uint32_t first_valid = 0xFFFFu;
#pragma unroll
    for (int dw = 0; dw < PW; ++dw) {
      const int w = w0 + dw;
#pragma unroll
      for (int dh = 0; dh < PH; ++dh) {
        const int h = h0 + dh;
        // kNoPad: every window lies inside the image (pad = 0 with vl_nnpool's floor output size), so the
        // window loads carry no control dependence and are issued together
        if (!kNoPad && (w < 0 || w >= g.W || h < 0 || h >= g.H)) continue;
        const uint32_t pos = uint32_t(dw * PH + dh);
        if (kNoPad) first_valid = 0u;
        else if (first_valid == 0xFFFFu) first_valid = pos;
        const uint4 v = __ldg(reinterpret_cast<const uint4*>(xn + (size_t(h) * g.W + w) * g.C));
        const __half2* v2 = reinterpret_cast<const __half2*>(&v);
        const uint32_t pos2 = pos * 0x00010001u;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const __half2 xv = kAffineRelu ? __hmul2(v2[k], sgn[k]) : v2[k];
          const uint32_t m = __hgt2_mask(xv, best[k]);
          arg[k] = (arg[k] & ~m) | (pos2 & m);
          best[k] = __hmax2(best[k], xv);
        }
      }
    }
    uint4 o, xo;
    __half2* o2 = reinterpret_cast<__half2*>(&o);
    __half2* xo2 = reinterpret_cast<__half2*>(&xo);
    uint32_t ab[8];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      xo2[k] = kAffineRelu ? __hmul2(best[k], sgn[k]) : best[k];   // the raw (pre-affine) value of the winner
      float2 f = __half22float2(xo2[k]);
      ab[2 * k] = arg[k] & 0xFFFFu;
      ab[2 * k + 1] = arg[k] >> 16;
      if (kAffineRelu) {
        f.x = fmaxf(fmaf(av[2 * k], f.x, bv[2 * k]), 0.f);
        f.y = fmaxf(fmaf(av[2 * k + 1], f.y, bv[2 * k + 1]), 0.f);
        if (!(f.x > 0.f)) ab[2 * k] = first_valid;       // all-zero window: the first scanned element wins
        if (!(f.y > 0.f)) ab[2 * k + 1] = first_valid;
      }
      o2[k] = __floats2half2_rn(f.x, f.y);
    }
    const size_t off = ((size_t(n) * g.OH + oh) * g.OW + ow) * g.OC + c8 * 8;
    *reinterpret_cast<uint4*>(y + off) = o;
    if (xwin) *reinterpret_cast<uint4*>(xwin + off) = xo;   // consumed by stem_pool_bn_reduce_kernel
    if (idx) {
      uint2 pk;
      pk.x = ab[0] | (ab[1] << 8) | (ab[2] << 16) | (ab[3] << 24);
      pk.y = ab[4] | (ab[5] << 8) | (ab[6] << 16) | (ab[7] << 24);
      *reinterpret_cast<uint2*>(idx + off) = pk;
    }
  }
}

template <int MH, int MW>
__global__ void maxpool_bwd_h2_kernel(const __half* __restrict__ dy, const uint8_t* __restrict__ idx, PoolGeom g,
                                      __half* __restrict__ dx) {
  const uint32_t C8 = uint32_t(g.C >> 3);
  const uint32_t npix = uint32_t(g.N) * uint32_t(g.H) * uint32_t(g.W);
  const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t c8 = tid % C8;
  const uint32_t pstride = (gridDim.x * blockDim.x) / C8;
  for (uint32_t pp = tid / C8; pp < npix; pp += pstride) {
    const uint32_t t1 = pp / uint32_t(g.W);
    const int w = int(pp - t1 * uint32_t(g.W));
    const int n = int(t1 / uint32_t(g.H));
    const int h = int(t1 - uint32_t(n) * uint32_t(g.H));
    __half2 acc[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) acc[k] = __floats2half2_rn(0.f, 0.f);
    const int oh_hi = (h + g.pt) / g.sh, ow_hi = (w + g.pl) / g.sw;
#pragma unroll
    for (int ia = 0; ia < MH; ++ia) {
      const int oh = oh_hi - ia;
      const int dh = h + g.pt - oh * g.sh;
      if (oh < 0 || oh >= g.OH || dh >= g.PH) continue;
#pragma unroll
      for (int ib = 0; ib < MW; ++ib) {
        const int ow = ow_hi - ib;
        const int dw = w + g.pl - ow * g.sw;
        if (ow < 0 || ow >= g.OW || dw >= g.PW) continue;
        const size_t off = ((size_t(n) * g.OH + oh) * g.OW + ow) * g.C + c8 * 8;
        const uint2 pk = __ldg(reinterpret_cast<const uint2*>(idx + off));
        const uint4 v = __ldg(reinterpret_cast<const uint4*>(dy + off));
        const uint32_t me4 = uint32_t(dw * g.PH + dh) * 0x01010101u;
        const uint32_t m0 = __vcmpeq4(pk.x, me4), m1 = __vcmpeq4(pk.y, me4);  // 0xFF per matching channel
        // byte masks -> 16-bit lane masks, AND with the packed gradients, accumulate in half2
        const uint32_t w0 = v.x & __byte_perm(m0, 0, 0x1100), w1 = v.y & __byte_perm(m0, 0, 0x3322);
        const uint32_t w2 = v.z & __byte_perm(m1, 0, 0x1100), w3 = v.w & __byte_perm(m1, 0, 0x3322);
        acc[0] = __hadd2(acc[0], *reinterpret_cast<const __half2*>(&w0));
        acc[1] = __hadd2(acc[1], *reinterpret_cast<const __half2*>(&w1));
        acc[2] = __hadd2(acc[2], *reinterpret_cast<const __half2*>(&w2));
        acc[3] = __hadd2(acc[3], *reinterpret_cast<const __half2*>(&w3));
      }
    }
    uint4 o;
    __half2* o2 = reinterpret_cast<__half2*>(&o);
#pragma unroll
    for (int k = 0; k < 4; ++k) o2[k] = acc[k];
    *reinterpret_cast<uint4*>(dx + ((size_t(n) * g.H + h) * g.W + w) * g.C + c8 * 8) = o;
  }
}

// 3x3 / stride 2 / pad 0 max-pool backward (both overlapping student pools): one thread owns the 2x2 input
// cell (2i..2i+1, 2j..2j+1) of one channel group.  The cell is covered by exactly the four windows
// (i-a, j-b), a, b in {0,1}; each window's (arg-max, gradient) pair is loaded once and routed to the (at
// most) nine (position, window) combinations -- 24 bytes read per 16 bytes written, no redundant loads.
static __global__ void maxpool_bwd_3x3s2_h2_kernel(const __half* __restrict__ dy, const uint8_t* __restrict__ idx, PoolGeom g,
                                                   __half* __restrict__ dx) {
  const uint32_t C8 = uint32_t(g.C >> 3);
  const uint32_t HC = uint32_t(g.H + 1) >> 1, WC = uint32_t(g.W + 1) >> 1;
  const uint32_t ncell = uint32_t(g.N) * HC * WC;
  const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t c8 = tid % C8;
  const uint32_t pstride = (gridDim.x * blockDim.x) / C8;
  for (uint32_t pp = tid / C8; pp < ncell; pp += pstride) {
    const uint32_t t1 = pp / WC;
    const int j = int(pp - t1 * WC);
    const int n = int(t1 / HC);
    const int i = int(t1 - uint32_t(n) * HC);
    uint2 pk[2][2];
    uint4 v[2][2];
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
      for (int b = 0; b < 2; ++b) {
        const int oh = i - a, ow = j - b;
        const bool valid = oh >= 0 && oh < g.OH && ow >= 0 && ow < g.OW;
        const size_t off = ((size_t(n) * g.OH + (valid ? oh : 0)) * g.OW + (valid ? ow : 0)) * g.OC + c8 * 8;
        pk[a][b] = __ldg(reinterpret_cast<const uint2*>(idx + off));
        v[a][b] = __ldg(reinterpret_cast<const uint4*>(dy + off));
        if (!valid) pk[a][b] = make_uint2(0xFFFFFFFFu, 0xFFFFFFFFu);  // matches no window-local index
      }
#pragma unroll
    for (int py = 0; py < 2; ++py)
#pragma unroll
      for (int px = 0; px < 2; ++px) {
        const int h = 2 * i + py, w = 2 * j + px;
        if (h >= g.H || w >= g.W) continue;
        __half2 acc[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) acc[k] = __floats2half2_rn(0.f, 0.f);
#pragma unroll
        for (int a = 0; a < 2; ++a) {
          if (a == 1 && py == 1) continue;  // window i-1 reaches row 2i only (dh = 2)
#pragma unroll
          for (int b = 0; b < 2; ++b) {
            if (b == 1 && px == 1) continue;
            const uint32_t me4 = uint32_t((px + 2 * b) * 3 + (py + 2 * a)) * 0x01010101u;  // dw * PH + dh
            const uint32_t m0 = __vcmpeq4(pk[a][b].x, me4), m1 = __vcmpeq4(pk[a][b].y, me4);
            const uint32_t w0 = v[a][b].x & __byte_perm(m0, 0, 0x1100), w1 = v[a][b].y & __byte_perm(m0, 0, 0x3322);
            const uint32_t w2 = v[a][b].z & __byte_perm(m1, 0, 0x1100), w3 = v[a][b].w & __byte_perm(m1, 0, 0x3322);
            acc[0] = __hadd2(acc[0], *reinterpret_cast<const __half2*>(&w0));
            acc[1] = __hadd2(acc[1], *reinterpret_cast<const __half2*>(&w1));
            acc[2] = __hadd2(acc[2], *reinterpret_cast<const __half2*>(&w2));
            acc[3] = __hadd2(acc[3], *reinterpret_cast<const __half2*>(&w3));
          }
        }
        uint4 o;
        __half2* o2 = reinterpret_cast<__half2*>(&o);
#pragma unroll
        for (int k = 0; k < 4; ++k) o2[k] = acc[k];
        *reinterpret_cast<uint4*>(dx + ((size_t(n) * g.H + h) * g.W + w) * g.C + c8 * 8) = o;
      }
  }
}

// Average pooling (vl_nnpool 'avg'): divides by the number of in-bounds window elements.
template <typename T>
__global__ void avgpool_fwd_kernel(const T* __restrict__ x, PoolGeom g, T* __restrict__ y) {
  const int C8 = g.C >> 3;
  const size_t total = size_t(g.N) * g.OH * g.OW * C8;
  for (size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x; i < total; i += size_t(gridDim.x) * blockDim.x) {
    const int c8 = int(i % C8);
    const int ow = int((i / C8) % g.OW);
    const int oh = int((i / (size_t(C8) * g.OW)) % g.OH);
    const int n = int(i / (size_t(C8) * g.OW * g.OH));
    float acc[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[k] = 0.f;
    int cnt = 0;
    for (int dh = 0; dh < g.PH; ++dh) {
      const int h = oh * g.sh + dh - g.pt;
      if (h < 0 || h >= g.H) continue;
      for (int dw = 0; dw < g.PW; ++dw) {
        const int w = ow * g.sw + dw - g.pl;
        if (w < 0 || w >= g.W) continue;
        Vec8<T> v;
        v.load(x + ((size_t(n) * g.H + h) * g.W + w) * g.C + c8 * 8);
        float f[8];
        v.to_float(f);
#pragma unroll
        for (int k = 0; k < 8; ++k) acc[k] += f[k];
        ++cnt;
      }
    }
    const float inv = 1.f / float(cnt);
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[k] *= inv;
    Vec8<T> o;
    o.from_float(acc);
    o.store(y + ((size_t(n) * g.OH + oh) * g.OW + ow) * g.C + c8 * 8);
  }
}

template <typename T>
__global__ void avgpool_bwd_kernel(const T* __restrict__ dy, PoolGeom g, T* __restrict__ dx) {
  const int C8 = g.C >> 3;
  const size_t total = size_t(g.N) * g.H * g.W * C8;
  for (size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x; i < total; i += size_t(gridDim.x) * blockDim.x) {
    const int c8 = int(i % C8);
    const int w = int((i / C8) % g.W);
    const int h = int((i / (size_t(C8) * g.W)) % g.H);
    const int n = int(i / (size_t(C8) * g.W * g.H));
    float acc[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[k] = 0.f;
    const int oh_hi = min((h + g.pt) / g.sh, g.OH - 1);
    const int ow_hi = min((w + g.pl) / g.sw, g.OW - 1);
    for (int oh = oh_hi; oh >= 0; --oh) {
      if (h + g.pt - oh * g.sh >= g.PH) break;
      for (int ow = ow_hi; ow >= 0; --ow) {
        if (w + g.pl - ow * g.sw >= g.PW) break;
        // in-bounds element count of window (oh, ow)
        const int h0 = max(oh * g.sh - g.pt, 0), h1 = min(oh * g.sh - g.pt + g.PH, g.H);
        const int w0 = max(ow * g.sw - g.pl, 0), w1 = min(ow * g.sw - g.pl + g.PW, g.W);
        const float inv = 1.f / float((h1 - h0) * (w1 - w0));
        Vec8<T> v;
        v.load(dy + ((size_t(n) * g.OH + oh) * g.OW + ow) * g.C + c8 * 8);
        float f[8];
        v.to_float(f);
#pragma unroll
        for (int k = 0; k < 8; ++k) acc[k] += f[k] * inv;
      }
    }
    Vec8<T> o;
    o.from_float(acc);
    o.store(dx + ((size_t(n) * g.H + h) * g.W + w) * g.C + c8 * 8);
  }
}

// ============================================================================================
// Batch normalisation (vl_nnbnorm).  x: [P = N*H*W][C] fp16.
//   stats pass : per-channel sum / sum of squares in fp32 -> (double) atomics into acc[2*C]
//   finalize   : mu = S1/P, var = S2/P - mu^2 (biased), sigma = sqrt(var + eps); moments = [mu sigma];
//                a = g/sigma, b = beta - a*mu  (the affine the apply / pooling kernels consume)
//   apply      : y = a*x + b (+ReLU)
//   backward   : reduce db = sum(dz), dg = sum(dz * xhat) with dz = dy * [y>0] when ReLU is fused;
//                dx = a * (dz - db/P - xhat*dg/P)
// ============================================================================================
// Reduction layout shared by the statistics and the backward-reduce kernels: a block of 256 threads
// covers `lanes` = min(C/8, 256) channel groups x `rows_par` = 256 / lanes rows at a time and strides
// over the rows of its slab in registers (fp32); the rows_par partial sums are combined through
// shared memory and leave the block as ONE double atomic per channel and statistic.  The grid is
// (row slabs ~ 4 per SM, channel slabs), so the atomic traffic is a few thousand adds per launch.
constexpr int kBnThreads = 256;

struct BnGrid {
  int lanes, rows_par, slabs_x, slabs_y;
};
inline BnGrid bn_grid(size_t P, int C, int num_sms, int per_sm = 4) {
  BnGrid g;
  const int C8 = C >> 3;
  g.lanes = C8 < kBnThreads ? C8 : kBnThreads;
  g.rows_par = kBnThreads / g.lanes;
  g.slabs_y = (C8 + g.lanes - 1) / g.lanes;
  size_t want = size_t(num_sms) * per_sm / g.slabs_y;
  if (want < 1) want = 1;
  const size_t max_slabs = (P + size_t(g.rows_par) * 4 - 1) / (size_t(g.rows_par) * 4);  // >= 4 rows per thread
  g.slabs_x = int(want < max_slabs ? want : (max_slabs < 1 ? 1 : max_slabs));
  return g;
}

template <typename T>
__global__ void bn_stats_kernel(const T* __restrict__ x, size_t P, int C, int lanes, int rows_par,
                                double* __restrict__ acc) {
  __shared__ float red[2][kBnThreads][8];
  const int C8 = C >> 3;
  const int rl = threadIdx.x / lanes;
  const int cl = threadIdx.x - rl * lanes;
  const int c8 = blockIdx.y * lanes + cl;
  const bool active = rl < rows_par && c8 < C8;
  float s1[8], s2[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) { s1[k] = 0.f; s2[k] = 0.f; }
  if (active) {
    const size_t stride = size_t(gridDim.x) * rows_par;
    size_t r = size_t(blockIdx.x) * rows_par + rl;
    for (; r + 3 * stride < P; r += 4 * stride) {  // four independent 16-byte loads in flight per thread
      Vec8<T> v0, v1, v2, v3;
      v0.load(x + r * C + c8 * 8);
      v1.load(x + (r + stride) * C + c8 * 8);
      v2.load(x + (r + 2 * stride) * C + c8 * 8);
      v3.load(x + (r + 3 * stride) * C + c8 * 8);
      float f0[8], f1[8], f2[8], f3[8];
      v0.to_float(f0); v1.to_float(f1); v2.to_float(f2); v3.to_float(f3);
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        s1[k] += (f0[k] + f1[k]) + (f2[k] + f3[k]);
        s2[k] = fmaf(f0[k], f0[k], fmaf(f1[k], f1[k], fmaf(f2[k], f2[k], fmaf(f3[k], f3[k], s2[k]))));
      }
    }
    for (; r < P; r += stride) {
      Vec8<T> v;
      v.load(x + r * C + c8 * 8);
      float f[8];
      v.to_float(f);
#pragma unroll
      for (int k = 0; k < 8; ++k) { s1[k] += f[k]; s2[k] = fmaf(f[k], f[k], s2[k]); }
    }
  }
#pragma unroll
  for (int k = 0; k < 8; ++k) { red[0][threadIdx.x][k] = s1[k]; red[1][threadIdx.x][k] = s2[k]; }
  __syncthreads();
  if (rl == 0 && c8 < C8) {
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      double t1 = 0.0, t2 = 0.0;
      for (int j = 0; j < rows_par; ++j) { t1 += double(red[0][j * lanes + cl][k]); t2 += double(red[1][j * lanes + cl][k]); }
      atomicAdd(&acc[c8 * 8 + k], t1);
      atomicAdd(&acc[C + c8 * 8 + k], t2);
    }
  }
}

static __global__ void bn_finalize_kernel(const double* __restrict__ acc, size_t P, int C, const float* __restrict__ g,
                                   const float* __restrict__ beta, float eps, float* __restrict__ moments,
                                   float* __restrict__ a, float* __restrict__ b) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const double mu = acc[c] / double(P);
  double var = acc[C + c] / double(P) - mu * mu;
  if (var < 0) var = 0;
  const double sigma = sqrt(var + double(eps));
  moments[c] = float(mu);
  moments[C + c] = float(sigma);  // C x 2 column-major: [mu | sigma]
  const double aa = double(g[c]) / sigma;
  a[c] = float(aa);
  b[c] = float(double(beta[c]) - aa * mu);
}

// test mode: moments given (C x 2 = [mu sigma]).  conv_bias (optional) folds the bias of the convolution that
// feeds the BN into the shift, so that y = a * conv_nobias(x) + b can run in the convolution's epilogue.
static __global__ void bn_affine_from_moments_kernel(const float* __restrict__ moments, int C, const float* __restrict__ g,
                                              const float* __restrict__ beta, const float* __restrict__ conv_bias,
                                              float* __restrict__ a, float* __restrict__ b) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const float aa = g[c] / moments[C + c];
  a[c] = aa;
  b[c] = beta[c] - aa * moments[c] + (conv_bias ? aa * conv_bias[c] : 0.f);
}

template <typename T>
__global__ void affine_act_kernel(const T* __restrict__ x, size_t P, int C, const float* __restrict__ a,
                                  const float* __restrict__ b, int relu, T* __restrict__ y) {
  const int C8 = C >> 3;
  const size_t total = P * C8;
  const size_t tid = blockIdx.x * size_t(blockDim.x) + threadIdx.x;
  const int c8 = int(tid % C8);  // constant along the loop (fixed_channel_grid)
  float av[8], bv[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) { av[k] = a ? a[c8 * 8 + k] : 1.f; bv[k] = a ? b[c8 * 8 + k] : 0.f; }
  for (size_t i = tid; i < total; i += size_t(gridDim.x) * blockDim.x) {
    Vec8<T> v;
    v.load(x + i * 8);
    float f[8];
    v.to_float(f);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      float z = fmaf(av[k], f[k], bv[k]);
      if (relu) z = fmaxf(z, 0.f);
      f[k] = z;
    }
    v.from_float(f);
    v.store(y + i * 8);
  }
}

// dz of one (n, h, w, channel-group) position gathered through a max-pooling layer that follows the
// BN+ReLU (the pooled tensor's gradient `dout` and the recorded arg-max): the sum over the (at most
// 2 x 2) windows covering the position whose arg-max points at it.  Lets the BN backward read the
// 4x smaller pooled gradient instead of a materialised full-resolution one.
template <typename T>
__device__ __forceinline__ void pool_gather8(const T* __restrict__ dout, const uint8_t* __restrict__ idx, const PoolGeom& g,
                                             int n, int h, int w, int c8, float (&acc)[8]) {
#pragma unroll
  for (int k = 0; k < 8; ++k) acc[k] = 0.f;
  const int oh_hi = (h + g.pt) / g.sh, ow_hi = (w + g.pl) / g.sw;
#pragma unroll
  for (int ia = 0; ia < 2; ++ia) {
    const int oh = oh_hi - ia;
    const int dh = h + g.pt - oh * g.sh;
    if (oh < 0 || oh >= g.OH || dh >= g.PH) continue;
#pragma unroll
    for (int ib = 0; ib < 2; ++ib) {
      const int ow = ow_hi - ib;
      const int dw = w + g.pl - ow * g.sw;
      if (ow < 0 || ow >= g.OW || dw >= g.PW) continue;
      const size_t off = ((size_t(n) * g.OH + oh) * g.OW + ow) * g.C + c8 * 8;
      const uint2 pk = *reinterpret_cast<const uint2*>(idx + off);
      Vec8<T> v;
      v.load(dout + off);
      float f[8];
      v.to_float(f);
      const uint32_t me = uint32_t(dw * g.PH + dh);
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const uint32_t am = ((k < 4 ? pk.x : pk.y) >> (8 * (k & 3))) & 0xFF;
        if (am == me) acc[k] += f[k];
      }
    }
  }
}

// backward reduce: acc[0..C) += sum dz ; acc[C..2C) += sum dz * xhat, xhat = (x - mu)/sigma.
// dz = dy * [a*x+b > 0] when relu_mask.  Same block layout as bn_stats_kernel.  kPool: dy is gathered
// through the following max-pooling layer (dy = pooled gradient, idx = its arg-max, g = pool geometry).
template <typename T, bool kPool>
__global__ void bn_bwd_reduce_kernel(const T* __restrict__ x, const T* __restrict__ dy, size_t P, int C, int lanes,
                                     int rows_par, const float* __restrict__ moments, const float* __restrict__ a,
                                     const float* __restrict__ b, int relu_mask, double* __restrict__ acc,
                                     const uint8_t* __restrict__ idx, PoolGeom g) {
  __shared__ float red[2][kBnThreads][8];
  const int C8 = C >> 3;
  const int rl = threadIdx.x / lanes;
  const int cl = threadIdx.x - rl * lanes;
  const int c8 = blockIdx.y * lanes + cl;
  const bool active = rl < rows_par && c8 < C8;
  float s1[8], s2[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) { s1[k] = 0.f; s2[k] = 0.f; }
  if (active) {
    float mu[8], isg[8], av[8], bv[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      mu[k] = moments[c8 * 8 + k];
      isg[k] = 1.f / moments[C + c8 * 8 + k];
      av[k] = a[c8 * 8 + k];
      bv[k] = b[c8 * 8 + k];
    }
    const size_t stride = size_t(gridDim.x) * rows_par;
    size_t r = size_t(blockIdx.x) * rows_par + rl;
    if (!kPool) {
      for (; r + stride < P; r += 2 * stride) {  // two rows (four 16-byte loads) in flight per thread
        Vec8<T> vx0, vd0, vx1, vd1;
        vx0.load(x + r * C + c8 * 8);
        vd0.load(dy + r * C + c8 * 8);
        vx1.load(x + (r + stride) * C + c8 * 8);
        vd1.load(dy + (r + stride) * C + c8 * 8);
        float fx0[8], fd0[8], fx1[8], fd1[8];
        vx0.to_float(fx0); vd0.to_float(fd0); vx1.to_float(fx1); vd1.to_float(fd1);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          float dz0 = fd0[k], dz1 = fd1[k];
          if (relu_mask && !(fmaf(av[k], fx0[k], bv[k]) > 0.f)) dz0 = 0.f;
          if (relu_mask && !(fmaf(av[k], fx1[k], bv[k]) > 0.f)) dz1 = 0.f;
          s1[k] += dz0 + dz1;
          s2[k] = fmaf(dz0, (fx0[k] - mu[k]) * isg[k], fmaf(dz1, (fx1[k] - mu[k]) * isg[k], s2[k]));
        }
      }
    }
    for (; r < P; r += stride) {
      Vec8<T> vx;
      vx.load(x + r * C + c8 * 8);
      float fx[8], fd[8];
      vx.to_float(fx);
      if (kPool) {
        const uint32_t r32 = uint32_t(r);
        const uint32_t t = r32 / uint32_t(g.W);
        pool_gather8<T>(dy, idx, g, int(t / uint32_t(g.H)), int(t % uint32_t(g.H)), int(r32 - t * uint32_t(g.W)), c8, fd);
      } else {
        Vec8<T> vd;
        vd.load(dy + r * C + c8 * 8);
        vd.to_float(fd);
      }
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        float dz = fd[k];
        if (relu_mask && !(fmaf(av[k], fx[k], bv[k]) > 0.f)) dz = 0.f;
        s1[k] += dz;
        s2[k] = fmaf(dz, (fx[k] - mu[k]) * isg[k], s2[k]);
      }
    }
  }
#pragma unroll
  for (int k = 0; k < 8; ++k) { red[0][threadIdx.x][k] = s1[k]; red[1][threadIdx.x][k] = s2[k]; }
  __syncthreads();
  if (rl == 0 && c8 < C8) {
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      double t1 = 0.0, t2 = 0.0;
      for (int j = 0; j < rows_par; ++j) { t1 += double(red[0][j * lanes + cl][k]); t2 += double(red[1][j * lanes + cl][k]); }
      atomicAdd(&acc[c8 * 8 + k], t1);
      atomicAdd(&acc[C + c8 * 8 + k], t2);
    }
  }
}

// dx = a * (dz - db/P - xhat * dg/P) = A*dz - D*x + E with the per-channel constants
//   A = a = g/sigma,  D = a*dg/(P*sigma),  E = mu*D - a*db/P     (held in registers per thread).
// Optionally also accumulates colsum[c] += scale * sum_rows dx (the bias gradient of the convolution that
// feeds this BN layer) so that no separate pass over dx is needed.
template <typename T, bool kPool>
__global__ void bn_bwd_apply_kernel(const T* __restrict__ x, const T* __restrict__ dy, size_t P, int C,
                                    const float* __restrict__ moments, const float* __restrict__ a,
                                    const float* __restrict__ b, int relu_mask, const double* __restrict__ acc,
                                    T* __restrict__ dx, const uint8_t* __restrict__ idx, PoolGeom g,
                                    float* __restrict__ colsum, float colsum_scale) {
  extern __shared__ float cs_smem[];  // [C] when colsum
  const int C8 = C >> 3;
  const size_t total = P * C8;
  const size_t tid = blockIdx.x * size_t(blockDim.x) + threadIdx.x;
  const int c8 = int(tid % C8);  // constant along the loop (fixed_channel_grid)
  float A[8], B[8], D[8], E[8], cs[8];
  const double invP = 1.0 / double(P);
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int c = c8 * 8 + k;
    const double av = double(a[c]), mu = double(moments[c]), sg = double(moments[C + c]);
    const double d = av * acc[C + c] * invP / sg;
    A[k] = float(av);
    B[k] = b[c];
    D[k] = float(d);
    E[k] = float(mu * d - av * acc[c] * invP);
    cs[k] = 0.f;
  }
  if (colsum) {
    for (int c = threadIdx.x; c < C; c += blockDim.x) cs_smem[c] = 0.f;
    __syncthreads();
  }
  const size_t gstride = size_t(gridDim.x) * blockDim.x;
  size_t i = tid;
  if (!kPool) {
    for (; i + gstride < total; i += 2 * gstride) {  // two items (four 16-byte loads) in flight per thread
      Vec8<T> vx0, vd0, vx1, vd1;
      vx0.load(x + i * 8);
      vd0.load(dy + i * 8);
      vx1.load(x + (i + gstride) * 8);
      vd1.load(dy + (i + gstride) * 8);
      float fx0[8], fd0[8], fx1[8], fd1[8];
      vx0.to_float(fx0); vd0.to_float(fd0); vx1.to_float(fx1); vd1.to_float(fd1);
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        float dz0 = fd0[k], dz1 = fd1[k];
        if (relu_mask && !(fmaf(A[k], fx0[k], B[k]) > 0.f)) dz0 = 0.f;
        if (relu_mask && !(fmaf(A[k], fx1[k], B[k]) > 0.f)) dz1 = 0.f;
        fd0[k] = fmaf(A[k], dz0, fmaf(-D[k], fx0[k], E[k]));
        fd1[k] = fmaf(A[k], dz1, fmaf(-D[k], fx1[k], E[k]));
      }
      vd0.from_float(fd0);
      vd1.from_float(fd1);
      vd0.store(dx + i * 8);
      vd1.store(dx + (i + gstride) * 8);
      if (colsum) {
        float fr0[8], fr1[8];
        vd0.to_float(fr0); vd1.to_float(fr1);  // sum what was actually stored
#pragma unroll
        for (int k = 0; k < 8; ++k) cs[k] += fr0[k] + fr1[k];
      }
    }
  }
  for (; i < total; i += gstride) {
    Vec8<T> vx, vd;
    vx.load(x + i * 8);
    float fx[8], fd[8];
    vx.to_float(fx);
    if (kPool) {
      const uint32_t r32 = uint32_t(i / C8);
      const uint32_t t = r32 / uint32_t(g.W);
      pool_gather8<T>(dy, idx, g, int(t / uint32_t(g.H)), int(t % uint32_t(g.H)), int(r32 - t * uint32_t(g.W)), c8, fd);
    } else {
      vd.load(dy + i * 8);
      vd.to_float(fd);
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      float dz = fd[k];
      if (relu_mask && !(fmaf(A[k], fx[k], B[k]) > 0.f)) dz = 0.f;
      fd[k] = fmaf(A[k], dz, fmaf(-D[k], fx[k], E[k]));
    }
    vd.from_float(fd);
    vd.store(dx + i * 8);
    if (colsum) {
      float fr[8];
      vd.to_float(fr);  // sum what was actually stored
#pragma unroll
      for (int k = 0; k < 8; ++k) cs[k] += fr[k];
    }
  }
  if (colsum) {
#pragma unroll
    for (int k = 0; k < 8; ++k) atomicAdd(&cs_smem[c8 * 8 + k], cs[k]);
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += blockDim.x) atomicAdd(colsum + c, cs_smem[c] * colsum_scale);
  }
}

static __global__ void bn_bwd_params_kernel(const double* __restrict__ acc, int C, float inv_grad_scale,
                                     float* __restrict__ dg, float* __restrict__ db) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  db[c] = float(acc[c]) * inv_grad_scale;
  dg[c] = float(acc[C + c]) * inv_grad_scale;
}

// plain ReLU backward: dx = dy * [y > 0] (y = ReLU output)
template <typename T>
__global__ void relu_bwd_kernel(const T* __restrict__ y, const T* __restrict__ dy, size_t n8,
                                T* __restrict__ dx) {
  for (size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x; i < n8; i += size_t(gridDim.x) * blockDim.x) {
    Vec8<T> vy, vd;
    vy.load(y + i * 8);
    vd.load(dy + i * 8);
    float fy[8], fd[8];
    vy.to_float(fy);
    vd.to_float(fd);
#pragma unroll
    for (int k = 0; k < 8; ++k) fd[k] = fy[k] > 0.f ? fd[k] : 0.f;
    vd.from_float(fd);
    vd.store(dx + i * 8);
  }
}

// ============================================================================================
// Squeeze-and-excitation (mcnExtraLayers GlobalPooling / Conv 1x1 / Sigmoid / Axpy).
//   squeeze : s[n][c] = mean_{h,w} u[n,h,w,c]                       (fp32 out)
//   gate    : a = sigmoid(W2 * relu(W1 * s + b1) + b2)              (one block per sample)
//   excite  : y = relu(a[n][c] * u + shortcut)                      (vectorised pass)
// ============================================================================================
template <typename T, int F = 1>
__global__ void se_squeeze_kernel(const T* __restrict__ u, int HW, int C, float* __restrict__ s) {
  // grid: (ceil(C/8/LX), N); x -> channel group, y -> pixel stride.  The sum is defined over VLY = F * blockDim.y VIRTUAL
  // pixel-stride lanes (lane v takes pixels v, v + VLY, ... in groups of four, then singly; the lanes are added in order):
  // a thread owns the F lanes y, y + blockDim.y, ..., so that a block of 1024 / F threads produces bit-identical means
  // -- the launcher picks F by grid size, and the same face must give the same logits in any batch.
  // LX = min(32, C/8) so that narrow tensors still fill the block; LX * VLY <= 1024.
  const int LX = blockDim.x, LYt = blockDim.y, VLY = LYt * F;
  const int n = blockIdx.y;
  const int c8 = blockIdx.x * LX + threadIdx.x;
  __shared__ float part[1024 * 9];
  float acc[F][8];
#pragma unroll
  for (int f = 0; f < F; ++f)
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[f][k] = 0.f;
  if (c8 * 8 < C) {
    const T* base = u + size_t(n) * HW * C + c8 * 8;
    auto quad = [&](int p, float (&a)[8]) {
      Vec8<T> v0, v1, v2, v3;
      v0.load(base + size_t(p) * C);
      v1.load(base + size_t(p + VLY) * C);
      v2.load(base + size_t(p + 2 * VLY) * C);
      v3.load(base + size_t(p + 3 * VLY) * C);
      float f0[8], f1[8], f2[8], f3[8];
      v0.to_float(f0); v1.to_float(f1); v2.to_float(f2); v3.to_float(f3);
#pragma unroll
      for (int k = 0; k < 8; ++k) a[k] += (f0[k] + f1[k]) + (f2[k] + f3[k]);
    };
    auto single = [&](int p, float (&a)[8]) {
      Vec8<T> v;
      v.load(base + size_t(p) * C);
      float f[8];
      v.to_float(f);
#pragma unroll
      for (int k = 0; k < 8; ++k) a[k] += f[k];
    };
    int p0 = threadIdx.y;
    if constexpr (F == 2) {
      // both lanes' groups of four in one iteration: eight 16-byte loads in flight
      int p1 = p0 + LYt;
      for (; p1 + 3 * VLY < HW; p0 += 4 * VLY, p1 += 4 * VLY) {
        Vec8<T> a0, a1, a2, a3, b0, b1, b2, b3;
        a0.load(base + size_t(p0) * C);
        a1.load(base + size_t(p0 + VLY) * C);
        a2.load(base + size_t(p0 + 2 * VLY) * C);
        a3.load(base + size_t(p0 + 3 * VLY) * C);
        b0.load(base + size_t(p1) * C);
        b1.load(base + size_t(p1 + VLY) * C);
        b2.load(base + size_t(p1 + 2 * VLY) * C);
        b3.load(base + size_t(p1 + 3 * VLY) * C);
        float f0[8], f1[8], f2[8], f3[8];
        a0.to_float(f0); a1.to_float(f1); a2.to_float(f2); a3.to_float(f3);
#pragma unroll
        for (int k = 0; k < 8; ++k) acc[0][k] += (f0[k] + f1[k]) + (f2[k] + f3[k]);
        b0.to_float(f0); b1.to_float(f1); b2.to_float(f2); b3.to_float(f3);
#pragma unroll
        for (int k = 0; k < 8; ++k) acc[F - 1][k] += (f0[k] + f1[k]) + (f2[k] + f3[k]);
      }
      for (; p1 < HW; p1 += VLY) single(p1, acc[F - 1]);   // (lane y + blockDim.y has no group of four left)
    }
    for (; p0 + 3 * VLY < HW; p0 += 4 * VLY) quad(p0, acc[0]);
    for (; p0 < HW; p0 += VLY) single(p0, acc[0]);
  }
#pragma unroll
  for (int f = 0; f < F; ++f) {
    float* mine = part + ((threadIdx.y + f * LYt) * LX + threadIdx.x) * 9;
#pragma unroll
    for (int k = 0; k < 8; ++k) mine[k] = acc[f][k];
  }
  __syncthreads();
  // LX x 8 (channel-group, k) outputs, each summed over the VLY lanes by four threads (a quarter of the lanes each, in
  // order), combined as (p0 + p1) + (p2 + p3): a fixed order for a given VLY, and a serial tail of VLY / 4 instead of VLY
  const int nthreads = LX * LYt;
  const int per = VLY >> 2;   // VLY is a multiple of 4 (1024 / LX or 8)
  const unsigned lanes = __activemask();   // (whole warps, or the one partial warp of a tensor with fewer than 32 channels;
                                           //  the trip count below is uniform within a warp: LX * 32 and nthreads are multiples of 8)
  for (int t = threadIdx.y * LX + threadIdx.x; t < LX * 32; t += nthreads) {
    const int o = t >> 2, quarter = t & 3;
    const int cg = o >> 3, k = o & 7;
    float ps = 0.f;
    for (int j = quarter * per; j < (quarter + 1) * per; ++j) ps += part[(j * LX + cg) * 9 + k];
    ps += __shfl_xor_sync(lanes, ps, 1);
    ps += __shfl_xor_sync(lanes, ps, 2);
    const int c = (blockIdx.x * LX + cg) * 8 + k;
    if (quarter == 0 && c < C) s[size_t(n) * C + c] = ps / float(HW);
  }
}

// kSeSpb samples per block (256 threads) so that every weight element fetched from L2 feeds kSeSpb FMAs.
// w1: [Cr][C] fp32, w2t: [Cr][C] fp32 (the second FC transposed: consecutive threads read consecutive
// addresses).  Dynamic smem: kSeSpb * (C + Cr) floats.
constexpr int kSeSpb = 2;
static __global__ void se_gate_kernel(const float* __restrict__ s, int N, int C, int Cr, const float* __restrict__ w1,
                               const float* __restrict__ b1, const float* __restrict__ w2t,
                               const float* __restrict__ b2, float* __restrict__ gate) {
  extern __shared__ float sm[];
  float* sv = sm;                  // [kSeSpb][C]
  float* hid = sm + kSeSpb * C;    // [kSeSpb][Cr]
  const int n0 = blockIdx.x * kSeSpb;
  for (int i = threadIdx.x; i < kSeSpb * C; i += blockDim.x) {
    const int sidx = i / C, c = i - sidx * C;
    sv[i] = (n0 + sidx < N) ? s[size_t(n0 + sidx) * C + c] : 0.f;
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  for (int j = warp; j < Cr; j += nwarps) {
    float t[kSeSpb];
#pragma unroll
    for (int q = 0; q < kSeSpb; ++q) t[q] = 0.f;
    const float* wr = w1 + size_t(j) * C;
#pragma unroll 4
    for (int c = lane * 4; c < C; c += 128) {  // C is a multiple of 128 on this path (checked by the wrapper)
      const float4 w = *reinterpret_cast<const float4*>(wr + c);
#pragma unroll
      for (int q = 0; q < kSeSpb; ++q) {
        const float4 v = *reinterpret_cast<const float4*>(sv + q * C + c);
        t[q] = fmaf(w.x, v.x, fmaf(w.y, v.y, fmaf(w.z, v.z, fmaf(w.w, v.w, t[q]))));
      }
    }
#pragma unroll
    for (int q = 0; q < kSeSpb; ++q) {
      const float r = warp_sum(t[q]);
      if (lane == 0) hid[q * Cr + j] = fmaxf(r + (b1 ? b1[j] : 0.f), 0.f);
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float t[kSeSpb];
#pragma unroll
    for (int q = 0; q < kSeSpb; ++q) t[q] = b2 ? b2[c] : 0.f;
#pragma unroll 8
    for (int j = 0; j < Cr; ++j) {
      const float w = w2t[size_t(j) * C + c];
#pragma unroll
      for (int q = 0; q < kSeSpb; ++q) t[q] = fmaf(w, hid[q * Cr + j], t[q]);
    }
#pragma unroll
    for (int q = 0; q < kSeSpb; ++q)
      if (n0 + q < N) gate[size_t(n0 + q) * C + c] = 1.f / (1.f + __expf(-t[q]));
  }
}

// SE blocks by linearity (default on the 56 x 56 / 28 x 28 stages).  The SE squeeze is linear in the bottleneck's
// 3x3 output t2:  s[n,c] = mean_hw(a3[c] * (W3 t2)[n,.,c] + b3[c]) = a3[c] * (W3[c,:] . mean_hw t2[n,:]) + b3[c],
// so the expand convolution's output u never has to exist before the gate is known.  This kernel takes
// m2 = mean_hw(t2) ([N][Cm], se_squeeze_kernel on the C/4-channel tensor), forms s, runs the two gate FCs and emits the
// per-(image, channel) epilogue vectors of the fused expand+excite convolution (conv_fprop_kernel<64, true>):
//   nc_scale = gate * a3,  nc_shift = gate * b3    ->   y = relu(nc_scale * acc + nc_shift + shortcut).
// w3: [C][Cm] fp16 (the KRSC 1x1 filter); w1, w2t as in se_gate_kernel.  Dynamic smem: kSeSpb * (Cm + C + Cr) floats.
static __global__ void se_gate_lin_kernel(const float* __restrict__ m2, int N, int C, int Cm, int Cr,
                                          const __half* __restrict__ w3, const float* __restrict__ a3,
                                          const float* __restrict__ b3, const float* __restrict__ w1,
                                          const float* __restrict__ b1, const float* __restrict__ w2t,
                                          const float* __restrict__ b2, float* __restrict__ nc_scale,
                                          float* __restrict__ nc_shift) {
  extern __shared__ float sm[];
  float* mv = sm;                       // [kSeSpb][Cm]
  float* sv = mv + kSeSpb * Cm;         // [kSeSpb][C]
  float* hid = sv + kSeSpb * C;         // [kSeSpb][Cr]
  const int n0 = blockIdx.x * kSeSpb;
  for (int i = threadIdx.x; i < kSeSpb * Cm; i += blockDim.x) {
    const int sidx = i / Cm, j = i - sidx * Cm;
    mv[i] = (n0 + sidx < N) ? m2[size_t(n0 + sidx) * Cm + j] : 0.f;
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  for (int c = warp; c < C; c += nwarps) {          // s = a3 * (W3 . m2) + b3, one warp per output channel
    float t[kSeSpb];
#pragma unroll
    for (int q = 0; q < kSeSpb; ++q) t[q] = 0.f;
    const __half2* wr = reinterpret_cast<const __half2*>(w3 + size_t(c) * Cm);
    for (int j = lane; j < Cm / 2; j += 32) {        // Cm is a multiple of 64 (checked by the wrapper)
      const float2 w = __half22float2(wr[j]);
#pragma unroll
      for (int q = 0; q < kSeSpb; ++q) t[q] = fmaf(w.x, mv[q * Cm + 2 * j], fmaf(w.y, mv[q * Cm + 2 * j + 1], t[q]));
    }
#pragma unroll
    for (int q = 0; q < kSeSpb; ++q) {
      const float r = warp_sum(t[q]);
      if (lane == 0) sv[q * C + c] = fmaf(a3[c], r, b3[c]);
    }
  }
  __syncthreads();
  for (int j = warp; j < Cr; j += nwarps) {          // hidden = relu(W1 s + b1)
    float t[kSeSpb];
#pragma unroll
    for (int q = 0; q < kSeSpb; ++q) t[q] = 0.f;
    const float* wr = w1 + size_t(j) * C;
    for (int c = lane * 4; c < C; c += 128) {
      const float4 w = *reinterpret_cast<const float4*>(wr + c);
#pragma unroll
      for (int q = 0; q < kSeSpb; ++q) {
        const float4 v = *reinterpret_cast<const float4*>(sv + q * C + c);
        t[q] = fmaf(w.x, v.x, fmaf(w.y, v.y, fmaf(w.z, v.z, fmaf(w.w, v.w, t[q]))));
      }
    }
#pragma unroll
    for (int q = 0; q < kSeSpb; ++q) {
      const float r = warp_sum(t[q]);
      if (lane == 0) hid[q * Cr + j] = fmaxf(r + (b1 ? b1[j] : 0.f), 0.f);
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {   // gate = sigmoid(W2 hidden + b2) -> epilogue vectors
    float t[kSeSpb];
#pragma unroll
    for (int q = 0; q < kSeSpb; ++q) t[q] = b2 ? b2[c] : 0.f;
    for (int j = 0; j < Cr; ++j) {
      const float w = w2t[size_t(j) * C + c];
#pragma unroll
      for (int q = 0; q < kSeSpb; ++q) t[q] = fmaf(w, hid[q * Cr + j], t[q]);
    }
#pragma unroll
    for (int q = 0; q < kSeSpb; ++q)
      if (n0 + q < N) {
        const float gate = 1.f / (1.f + __expf(-t[q]));
        nc_scale[size_t(n0 + q) * C + c] = gate * a3[c];
        nc_shift[size_t(n0 + q) * C + c] = gate * b3[c];
      }
  }
}

template <typename T>
__global__ void se_excite_kernel(const T* __restrict__ u, const float* __restrict__ gate,
                                 const T* __restrict__ shortcut, int HW, int C, size_t total8, int relu,
                                 T* __restrict__ y) {
  const int C8 = C >> 3;
  for (size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x; i < total8; i += size_t(gridDim.x) * blockDim.x) {
    const int c8 = int(i % C8);
    const size_t n = i / (size_t(C8) * HW);
    Vec8<T> vu, vs;
    vu.load(u + i * 8);
    float fu[8], fs[8];
    vu.to_float(fu);
    if (shortcut) { vs.load(shortcut + i * 8); vs.to_float(fs); }
    const float4 g0 = *reinterpret_cast<const float4*>(gate + n * C + c8 * 8);
    const float4 g1 = *reinterpret_cast<const float4*>(gate + n * C + c8 * 8 + 4);
    const float gv[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      float z = fmaf(gv[k], fu[k], shortcut ? fs[k] : 0.f);
      if (relu) z = fmaxf(z, 0.f);
      fu[k] = z;
    }
    vu.from_float(fu);
    vu.store(y + i * 8);
  }
}

// element-wise add (+ReLU): dagnn.Sum for the cases not fused into a conv epilogue
template <typename T>
__global__ void add_act_kernel(const T* __restrict__ a, const T* __restrict__ b, size_t n8, int relu,
                               T* __restrict__ y) {
  for (size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x; i < n8; i += size_t(gridDim.x) * blockDim.x) {
    Vec8<T> va, vb;
    va.load(a + i * 8);
    vb.load(b + i * 8);
    float fa[8], fb[8];
    va.to_float(fa);
    vb.to_float(fb);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      float z = fa[k] + fb[k];
      fa[k] = relu ? fmaxf(z, 0.f) : z;
    }
    va.from_float(fa);
    va.store(y + i * 8);
  }
}

// ============================================================================================
// Teacher -> student coupling (emoVoxCeleb/getBatchEmoVoxCeleb.m:133-159,179-188,210-214): per clip,
// aggregate ('max' default, or 'mean') the frame logits lgts[start:end, :] over the frames selected for
// the audio crop.  frame_logits: [sum F_i][ldl] fp32 (row-major per frame), first numPred classes used.
// ============================================================================================
static __global__ void logit_aggregate_kernel(const float* __restrict__ frame_logits, int ldl, const int* __restrict__ start,
                                       const int* __restrict__ end, int N, int numPred, int use_mean,
                                       float* __restrict__ target /* [N][numPred] */) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N * numPred) return;
  const int n = i / numPred, c = i - n * numPred;
  float acc = use_mean ? 0.f : -INFINITY;
  for (int f = start[n]; f < end[n]; ++f) {
    const float v = frame_logits[size_t(f) * ldl + c];
    acc = use_mean ? acc + v : fmaxf(acc, v);
  }
  if (use_mean) acc /= float(end[n] - start[n]);
  target[i] = acc;
}

// ============================================================================================
// Distillation loss (mcnExtraLayers vl_nnsoftmaxceloss wired at emoVoxCeleb/emoVoxZoo.m:151-157) fused
// with its backward and with the reference's metric layers (classerror / ErrorStats,
// emoVoxZoo.m:160-169).  One thread per sample, C <= 16 classes.
//   p = softmax(t/T) (logitTargets) ; q = softmax(x/T) ; loss = sum_n w_n * -sum_c p log q
//   dx = dzdy * w_n * (q - p)/T   (stored fp16, scaled by grad_scale, row pitch ldx)
//   maxLabel = argmax_c t (first max, 1-based) ; nerr += [argmax x != maxLabel] ; per-class counters.
// out_scalars: [0]=loss [1]=classerror ; class_stats: [C] correct, [C] count.
// ============================================================================================
constexpr int kLossMaxC = 16;
// loss_type: 0 = temperature-softmax cross-entropy (mcnExtraLayers vl_nnsoftmaxceloss; logit_targets = 0 with T = 1 is
// dagnn.Loss('softmaxlog') against a one-hot row), 1 = dagnn.EuclideanLoss (1/2 w_n |x - t|^2, dx = dzdy w_n (x - t)),
// 2 = dagnn.HuberLoss('sigma', sigma): smooth-L1 per element, linear where |x - t| > 1/sigma^2 (`T` carries sigma) --
// the three `lossType`s of emoVoxCeleb/emoVoxZoo.m:137-157 that take {prediction, logitTarget [, instanceWeights]}.
// TX / TDX: storage type of the logits and of their gradient (fp16 in the fast programs, fp32 otherwise).
template <typename TX, typename TDX>
__global__ void loss_fused_kernel(const TX* __restrict__ x, int ldx, const float* __restrict__ t, int ldt,
                                  const float* __restrict__ w, int N, int C, int loss_type, float T, int logit_targets,
                                  float dzdy, float grad_scale, TDX* __restrict__ dx, int lddx,
                                  float* __restrict__ out_scalars, float* __restrict__ class_stats,
                                  int* __restrict__ max_label) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  float loss = 0.f, err = 0.f;
  if (n < N) {
    float xv[kLossMaxC], tv[kLossMaxC];
    float xm = -INFINITY, tm = -INFINITY;
    int xa = 0, ta = 0;
    for (int c = 0; c < C; ++c) {
      xv[c] = float(x[size_t(n) * ldx + c]);
      tv[c] = t[size_t(n) * ldt + c];
      if (xv[c] > xm) { xm = xv[c]; xa = c; }
      if (tv[c] > tm) { tm = tv[c]; ta = c; }
    }
    const float wn = w ? w[n] : 1.f;
    const float gsc = grad_scale * dzdy * wn;
    if (loss_type == 0) {
      const float invT = 1.f / T;
      float xs = 0.f, ts = 0.f;
      for (int c = 0; c < C; ++c) {
        xv[c] = (xv[c] - xm) * invT;
        xs += expf(xv[c]);
        if (logit_targets) { tv[c] = expf((tv[c] - tm) * invT); ts += tv[c]; }
      }
      const float lse = logf(xs);
      for (int c = 0; c < C; ++c) {
        const float p = logit_targets ? tv[c] / ts : tv[c];
        const float logq = xv[c] - lse;
        loss -= p * logq;
        if (dx) dx[size_t(n) * lddx + c] = from_f32<TDX>(gsc * (expf(logq) - p) * invT);
      }
    } else if (loss_type == 1) {
      for (int c = 0; c < C; ++c) {
        const float d = xv[c] - tv[c];
        loss += 0.5f * d * d;
        if (dx) dx[size_t(n) * lddx + c] = from_f32<TDX>(gsc * d);
      }
    } else {
      const float s2 = T * T, knee = 1.f / s2;
      for (int c = 0; c < C; ++c) {
        const float d = xv[c] - tv[c], ad = fabsf(d);
        const bool lin = ad > knee;
        loss += lin ? ad - 0.5f * knee : 0.5f * s2 * ad * ad;
        if (dx) dx[size_t(n) * lddx + c] = from_f32<TDX>(gsc * (lin ? (d > 0.f ? 1.f : -1.f) : s2 * d));
      }
    }
    loss *= wn;
    err = (xa != ta) ? 1.f : 0.f;
    if (max_label) max_label[n] = ta + 1;
    if (class_stats) {
      atomicAdd(&class_stats[C + ta], 1.f);
      if (xa == ta) atomicAdd(&class_stats[ta], 1.f);
    }
  }
  // block-level sum in a fixed order (warp partials -> thread 0), then one atomic per block: a launch of one or two blocks
  // gives a bit-reproducible objective
  __shared__ float part[2][32];
  loss = warp_sum(loss);
  err = warp_sum(err);
  const int warp = threadIdx.x >> 5;
  if ((threadIdx.x & 31) == 0) { part[0][warp] = loss; part[1][warp] = err; }
  __syncthreads();
  if (threadIdx.x == 0) {
    float l = 0.f, e = 0.f;
    for (int wv = 0; wv < int(blockDim.x + 31) / 32; ++wv) { l += part[0][wv]; e += part[1][wv]; }
    atomicAdd(&out_scalars[0], l);
    atomicAdd(&out_scalars[1], e);
  }
}

// ============================================================================================
// cnn_train_dag's SGD-momentum update (defaults inherited at emoVoxCeleb/run_distillation.m:170-182):
//   m <- mu*m - (lambda*wd_mult*w + g/B) ;  w <- w + lr*lr_mult*m
// fused with the refresh of the fp16 KRSC filter copy the tcgen05 kernels read.  `g` may carry the
// loss scale (inv_grad_scale undoes it).  One launch per parameter tensor (flat fp32 master copy).
// ============================================================================================
static __global__ void sgd_momentum_kernel(float* __restrict__ w, float* __restrict__ m, const float* __restrict__ g, size_t n,
                                    float lr, float momentum, float wd, float inv_batch, float inv_grad_scale,
                                    __half* __restrict__ w16) {
  for (size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x; i < n; i += size_t(gridDim.x) * blockDim.x) {
    const float wi = w[i];
    const float mi = momentum * m[i] - (wd * wi + g[i] * inv_grad_scale * inv_batch);
    const float wn = wi + lr * mi;
    m[i] = mi;
    w[i] = wn;
    if (w16) w16[i] = __float2half_rn(wn);
  }
}

// BN moments moving average (dagnn.BatchNorm moments param: trainMethod 'average', learningRate 0.1):
//   moments <- (1 - rate) * moments + rate * batch_moments
static __global__ void moments_average_kernel(float* __restrict__ moments, const float* __restrict__ batch_moments, int n,
                                       float rate, const int* __restrict__ guard = nullptr, float bm_scale = 1.f) {
  if (guard && guard[0]) return;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  // bm_scale = 1 / ranks when batch_moments holds the SUM over the data-parallel ranks (MatConvNet's parameter server sums
  // the labs' moment "derivatives", each weighted by its sub-batch size, and divides by the global batch)
  if (i < n) moments[i] = (1.f - rate) * moments[i] + rate * bm_scale * batch_moments[i];
}

static __global__ void f32_to_f16_kernel(const float* __restrict__ s, size_t n, __half* __restrict__ d) {
  for (size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x; i < n; i += size_t(gridDim.x) * blockDim.x)
    d[i] = __float2half_rn(s[i]);
}
static __global__ void f16_to_f32_kernel(const __half* __restrict__ s, size_t n, float scale, float* __restrict__ d) {
  for (size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x; i < n; i += size_t(gridDim.x) * blockDim.x)
    d[i] = __half2float(s[i]) * scale;
}

}  // namespace xemo
