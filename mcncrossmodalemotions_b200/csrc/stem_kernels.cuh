// stem_kernels.cuh -- the student's first layer (conv1 7x7/2 on a ONE-channel spectrogram + train-mode BN + ReLU +
// 3x3/2 max-pool) without the five full passes over its 1.85 GB activation that the generic BN forward statistics
// and BN backward cost (emoVoxCeleb/run_distillation.m:170: cnn_train_dag -> dagnn.Conv / BatchNorm / Pooling
// forward+backward on the 512 x W x 1 x N batch of getBatchEmoVoxCeleb.m:197).
//
// With C_in = 1 the layer is linear in a 64-entry patch vector  xs[p] = (s2d[n, oh+j, ow, c])_{j<4, c<16}
// (hbm_kernels.cuh: spec_s2d_from_hwcn_kernel), x[p,k] = w_k . xs[p] + b_k, so everything BN needs reduces to the
// 64 x 64 patch autocorrelation  R = sum_p xs[p] xs[p]^T  and the patch sum  S = sum_p xs[p]  (input-only, 155 MB read):
//   forward  : mean_k = w_k.S/P + b_k ; var_k = w_k^T R w_k / P - (w_k.S/P)^2                     (vl_nnbnorm moments)
//   backward : dY = A dz - D x + E (bn_bwd_apply_kernel's form) is never materialised: the filter gradient is linear
//              in dY,  dW_k = A_k G1_k - D_k (R w_k + b_k S) + E_k S  with  G1 = sum_p dz[p,.] xs[p]^T  (the tcgen05
//              wgrad kernel run on dz), and the two BN reductions  sum dz, sum dz*xhat  only touch the positions the
//              max-pool routed a gradient to -- they are taken at the POOLED resolution from the pooled gradient and
//              the raw winner value the pooling forward records (4x fewer bytes), which is also where the ReLU mask
//              is applied.
// R has lag structure: R[(j',c'),(j,c)] = sum over rows h1 in [j', j'+OH) of P_{j-j'}[h1][c',c] with the row-pair
// products P_d[h1] = X[h1]^T X[h1+d] (16 x 16, reduced over n and ow), so one pass computes the four totals T_d and
// the six boundary rows (h1 < 3, h1 >= OH) separately: 21.6 GFLOP instead of 79.  That pass is a small K-reduction
// with M = 16: it runs on the legacy warp-level tensor path (mma.sync m16n8k16, ldmatrix) -- 0.03 % of the step's
// FLOPs, bound by the 155 MB read.
#pragma once
#include <cuda_fp16.h>
#include <stdint.h>

#include "hbm_kernels.cuh"

namespace xemo {

constexpr int kAcRows = 16;                 // h1 rows (tasks) per work unit
constexpr int kAcOw = 160;                  // ow positions (GEMM-K) per work unit, multiple of 16
constexpr int kAcThreads = 256;
constexpr int kAcClasses = 7;               // 0: all rows; 1..3: h1 = 0,1,2; 4..6: h1 = OH, OH+1, OH+2
constexpr int kAcPElems = 16 * 4 * 16;      // [c'][d][c]
constexpr int kAcAccDoubles = kAcClasses * (kAcPElems + 16);   // products then column sums, per class
constexpr int kAcSmemBytes = (kAcRows + 3) * kAcOw * 32;
constexpr int kStemT = 64;                  // patch entries: 4 taps x 16 s2d channels
constexpr int kStemRS = kStemT * kStemT + kStemT;   // assembled R (64 x 64) followed by S (64), doubles

__device__ __forceinline__ uint32_t ac_off(int row, int ow, int half) {
  // 32-byte (16-channel) cells; the two 16-byte halves are swapped on every other group of four ow so that the
  // eight 16-byte rows an ldmatrix phase reads (32-byte pitch) fall into distinct banks
  return uint32_t(row) * uint32_t(kAcOw * 32) + uint32_t(ow) * 32u + uint32_t((half ^ ((ow >> 2) & 1)) << 4);
}

__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(addr));
}

__device__ __forceinline__ void mma_m16n8k16(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// acc: [kAcClasses][kAcPElems + 16] doubles, zeroed by the caller.
//   acc[cls][(c'*4 + d)*16 + c] += sum_{n, ow, h1 in cls} X[n,h1,ow,c'] * X[n,h1+d,ow,c]   (rows >= HP read as zero)
//   acc[cls][kAcPElems + c]     += sum_{n, ow, h1 in cls} X[n,h1,ow,c]
__global__ void __launch_bounds__(kAcThreads, 2)
stem_autocorr_kernel(const __half* __restrict__ x, int N, int HP, int OW, int OH, double* __restrict__ acc) {
  extern __shared__ __align__(128) uint8_t ac_smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  const int strips = (HP + kAcRows - 1) / kAcRows;
  const int chunks = (OW + kAcOw - 1) / kAcOw;
  const int units = N * strips * chunks;
  const uint32_t sbase = uint32_t(__cvta_generic_to_shared(ac_smem));

  float tot[8][4];      // running totals of this warp (class 0)
  float cs_tot[2] = {0.f, 0.f};
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) tot[i][j] = 0.f;

  for (int unit = blockIdx.x; unit < units; unit += gridDim.x) {
    const int chunk = unit % chunks;
    const int strip = (unit / chunks) % strips;
    const int n = unit / (chunks * strips);
    const int h0 = strip * kAcRows, ow0 = chunk * kAcOw;
    __syncthreads();  // the previous unit's fragments have been read
    // ---- stage rows h0 .. h0+18 (zero outside the tensor) : 16-byte pieces, coalesced along (ow, half)
#pragma unroll 4
    for (int i = threadIdx.x; i < (kAcRows + 3) * kAcOw * 2; i += kAcThreads) {
      const int half = i & 1;
      const int ow = (i >> 1) % kAcOw;
      const int row = i / (2 * kAcOw);
      const int h = h0 + row, w = ow0 + ow;
      uint4 v = make_uint4(0u, 0u, 0u, 0u);
      if (h < HP && w < OW) v = __ldg(reinterpret_cast<const uint4*>(x + ((size_t(n) * HP + h) * OW + w) * 16) + half);
      *reinterpret_cast<uint4*>(ac_smem + ac_off(row, ow, half)) = v;
    }
    __syncthreads();
    // ---- each warp: tasks (h1 rows) warp, warp + 8
    for (int task = warp; task < kAcRows; task += 8) {
      const int h1 = h0 + task;
      if (h1 >= HP) break;
      float accd[8][4];
      float cs[2] = {0.f, 0.f};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) accd[i][j] = 0.f;
      const int mid = lane >> 3, mi = lane & 7;
#pragma unroll 2
      for (int kk = 0; kk < kAcOw; kk += 16) {
        uint32_t a[4];
        // A = X[h1]^T (16 c' x 16 ow): matrices (c' 0-7, ow 0-7), (c' 8-15, ow 0-7), (c' 0-7, ow 8-15), (c' 8-15, ow 8-15)
        ldmatrix_x4_trans(sbase + ac_off(task, kk + ((mid & 2) ? 8 : 0) + mi, mid & 1), a);
#pragma unroll
        for (int d = 0; d < 4; ++d) {
          uint32_t b[4];
          // B = X[h1+d] (16 ow x 16 c): matrices (ow 0-7, c 0-7), (ow 8-15, c 0-7), (ow 0-7, c 8-15), (ow 8-15, c 8-15)
          ldmatrix_x4_trans(sbase + ac_off(task + d, kk + ((mid & 1) ? 8 : 0) + mi, mid >> 1), b);
          mma_m16n8k16(accd[2 * d], a, b[0], b[1]);
          mma_m16n8k16(accd[2 * d + 1], a, b[2], b[3]);
          if (d == 0) {
            const float2 f0 = __half22float2(*reinterpret_cast<const __half2*>(&b[0]));
            const float2 f1 = __half22float2(*reinterpret_cast<const __half2*>(&b[1]));
            const float2 f2 = __half22float2(*reinterpret_cast<const __half2*>(&b[2]));
            const float2 f3 = __half22float2(*reinterpret_cast<const __half2*>(&b[3]));
            cs[0] += (f0.x + f0.y) + (f1.x + f1.y);   // column c = g       over this lane's four ow
            cs[1] += (f2.x + f2.y) + (f3.x + f3.y);   // column c = 8 + g
          }
        }
      }
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        cs[i] += __shfl_xor_sync(0xffffffffu, cs[i], 1);
        cs[i] += __shfl_xor_sync(0xffffffffu, cs[i], 2);
        cs_tot[i] += cs[i];
      }
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) tot[i][j] += accd[i][j];
      const int cls = h1 < 3 ? 1 + h1 : (h1 >= OH ? 4 + (h1 - OH) : 0);
      if (cls > 0 && cls < kAcClasses) {   // boundary row: its products also go to their own bin
        double* dst = acc + size_t(cls) * (kAcPElems + 16);
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
          const int d = nt >> 1, c = (nt & 1) * 8 + 2 * t;
          atomicAdd(dst + ((g * 4 + d) * 16 + c), double(accd[nt][0]));
          atomicAdd(dst + ((g * 4 + d) * 16 + c + 1), double(accd[nt][1]));
          atomicAdd(dst + (((g + 8) * 4 + d) * 16 + c), double(accd[nt][2]));
          atomicAdd(dst + (((g + 8) * 4 + d) * 16 + c + 1), double(accd[nt][3]));
        }
        if (t == 0) {
          atomicAdd(dst + kAcPElems + g, double(cs[0]));
          atomicAdd(dst + kAcPElems + 8 + g, double(cs[1]));
        }
      }
    }
  }
  // ---- block reduction of the class-0 totals (8 warps -> 1) through shared memory, then one double atomic per entry
  __syncthreads();
  float* red = reinterpret_cast<float*>(ac_smem);   // [8 warps][kAcPElems + 16]
  {
    float* mine = red + warp * (kAcPElems + 16);
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      const int d = nt >> 1, c = (nt & 1) * 8 + 2 * t;
      mine[(g * 4 + d) * 16 + c] = tot[nt][0];
      mine[(g * 4 + d) * 16 + c + 1] = tot[nt][1];
      mine[((g + 8) * 4 + d) * 16 + c] = tot[nt][2];
      mine[((g + 8) * 4 + d) * 16 + c + 1] = tot[nt][3];
    }
    if (t == 0) {
      mine[kAcPElems + g] = cs_tot[0];
      mine[kAcPElems + 8 + g] = cs_tot[1];
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < kAcPElems + 16; i += kAcThreads) {
    double s = 0.0;
#pragma unroll
    for (int w = 0; w < 8; ++w) s += double(red[w * (kAcPElems + 16) + i]);
    atomicAdd(acc + i, s);
  }
}

// acc (stem_autocorr_kernel) -> rs = [R (64 x 64) | S (64)] in doubles.  One block of 64 x 16 threads.
//   R[(j',c'),(j,c)] = T_d[c',c] - sum_{h1 < j'} P_d[h1][c',c] - sum_{h1 >= OH + j'} P_d[h1][c',c],  d = j - j' >= 0
//   (mirrored for j < j'),  S[(j,c)] = cs_T[c] - sum_{h < j} cs[h][c] - sum_{h >= OH + j} cs[h][c].
static __global__ void stem_assemble_kernel(const double* __restrict__ acc, double* __restrict__ rs) {
  const int stride = kAcPElems + 16;
  for (int i = threadIdx.x; i < kStemT * kStemT; i += blockDim.x) {
    int tp = i / kStemT, tq = i % kStemT;
    int jp = tp >> 4, cp = tp & 15, jq = tq >> 4, cq = tq & 15;
    if (jq < jp) { int s; s = jp; jp = jq; jq = s; s = cp; cp = cq; cq = s; }
    const int d = jq - jp;
    const int e = (cp * 4 + d) * 16 + cq;
    double v = acc[e];
    for (int h = 0; h < jp; ++h) v -= acc[(1 + h) * stride + e];          // top rows h1 = 0 .. j'-1
    for (int h = jp; h < 3; ++h) v -= acc[(4 + h) * stride + e];          // bottom rows h1 = OH + j' .. OH + 2
    rs[i] = v;
  }
  for (int i = threadIdx.x; i < kStemT; i += blockDim.x) {
    const int j = i >> 4, c = i & 15;
    double v = acc[kAcPElems + c];
    for (int h = 0; h < j; ++h) v -= acc[(1 + h) * stride + kAcPElems + c];
    for (int h = j; h < 3; ++h) v -= acc[(4 + h) * stride + kAcPElems + c];
    rs[kStemT * kStemT + i] = v;
  }
}

// Train-mode BN statistics of conv1's output from (R, S): one block of 64 threads per output channel.
//   w16: [C][64] fp16 (the filter the tensor cores read), bias: [C] fp32.  Outputs as bn_finalize_kernel.
static __global__ void stem_bn_stats_kernel(const double* __restrict__ rs, const __half* __restrict__ w16,
                                            const float* __restrict__ bias, double P, int C, const float* __restrict__ gamma,
                                            const float* __restrict__ beta, float eps, float* __restrict__ moments,
                                            float* __restrict__ a, float* __restrict__ b) {
  __shared__ double wv[kStemT];
  __shared__ double red[2][kStemT];
  const int k = blockIdx.x, t = threadIdx.x;
  wv[t] = double(__half2float(w16[size_t(k) * kStemT + t]));
  __syncthreads();
  double rw = 0.0;
  for (int u = 0; u < kStemT; ++u) rw += rs[u * kStemT + t] * wv[u];   // (R w)_t  (R symmetric)
  red[0][t] = rw * wv[t];
  red[1][t] = rs[kStemT * kStemT + t] * wv[t];
  __syncthreads();
  if (t == 0) {
    double q = 0.0, ws = 0.0;
    for (int u = 0; u < kStemT; ++u) { q += red[0][u]; ws += red[1][u]; }
    const double m0 = ws / P;                         // mean without the bias
    double var = q / P - m0 * m0;
    if (var < 0) var = 0;
    const double mu = m0 + double(bias[k]);
    const double sigma = sqrt(var + double(eps));
    moments[k] = float(mu);
    moments[C + k] = float(sigma);
    const double aa = double(gamma[k]) / sigma;
    a[k] = float(aa);
    b[k] = float(double(beta[k]) - aa * mu);
  }
}

// BN-backward reductions at the pooled resolution + ReLU masking of the pooled gradient (in place).
//   xw  : [P][ld] raw (pre-BN) value of each pooling window's winner (maxpool_fwd_h2_kernel's `xwin`), C <= ld valid
//   g   : [P][ld] gradient w.r.t. the pooled output; on exit g *= [a*xw + b > 0]
//   acc : [2C] doubles, acc[c] += sum g_masked, acc[C+c] += sum g_masked * (xw - mu)/sigma
// Every conv-resolution position that receives gradient is the winner of the windows that route to it, so these sums
// equal the full-resolution sums of bn_bwd_reduce_kernel.  Same block layout as bn_stats_kernel.
static __global__ void stem_pool_bn_reduce_kernel(const __half* __restrict__ xw, __half* __restrict__ g, size_t P, int C,
                                                  int ld, int lanes, int rows_par, const float* __restrict__ moments,
                                                  const float* __restrict__ a, const float* __restrict__ b,
                                                  double* __restrict__ acc) {
  __shared__ float red[2][256][8];
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
    for (size_t r = size_t(blockIdx.x) * rows_par + rl; r < P; r += 2 * stride) {
      const bool two = r + stride < P;
      const size_t o0 = r * ld + c8 * 8, o1 = (two ? r + stride : r) * ld + c8 * 8;
      uint4 vx0 = __ldg(reinterpret_cast<const uint4*>(xw + o0));
      uint4 vg0 = *reinterpret_cast<const uint4*>(g + o0);
      uint4 vx1 = __ldg(reinterpret_cast<const uint4*>(xw + o1));
      uint4 vg1 = *reinterpret_cast<const uint4*>(g + o1);
      const __half2* x0 = reinterpret_cast<const __half2*>(&vx0);
      const __half2* x1 = reinterpret_cast<const __half2*>(&vx1);
      __half2* g0 = reinterpret_cast<__half2*>(&vg0);
      __half2* g1 = reinterpret_cast<__half2*>(&vg1);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float2 fx0 = __half22float2(x0[k]), fx1 = __half22float2(x1[k]);
        float2 fg0 = __half22float2(g0[k]), fg1 = __half22float2(g1[k]);
        if (!(fmaf(av[2 * k], fx0.x, bv[2 * k]) > 0.f)) fg0.x = 0.f;
        if (!(fmaf(av[2 * k + 1], fx0.y, bv[2 * k + 1]) > 0.f)) fg0.y = 0.f;
        if (!(fmaf(av[2 * k], fx1.x, bv[2 * k]) > 0.f)) fg1.x = 0.f;
        if (!(fmaf(av[2 * k + 1], fx1.y, bv[2 * k + 1]) > 0.f)) fg1.y = 0.f;
        g0[k] = __floats2half2_rn(fg0.x, fg0.y);   // exact: the values are either the stored fp16 or zero
        g1[k] = __floats2half2_rn(fg1.x, fg1.y);
        s1[2 * k] += fg0.x;
        s1[2 * k + 1] += fg0.y;
        s2[2 * k] = fmaf(fg0.x, (fx0.x - mu[2 * k]) * isg[2 * k], s2[2 * k]);
        s2[2 * k + 1] = fmaf(fg0.y, (fx0.y - mu[2 * k + 1]) * isg[2 * k + 1], s2[2 * k + 1]);
        if (two) {
          s1[2 * k] += fg1.x;
          s1[2 * k + 1] += fg1.y;
          s2[2 * k] = fmaf(fg1.x, (fx1.x - mu[2 * k]) * isg[2 * k], s2[2 * k]);
          s2[2 * k + 1] = fmaf(fg1.y, (fx1.y - mu[2 * k + 1]) * isg[2 * k + 1], s2[2 * k + 1]);
        }
      }
      *reinterpret_cast<uint4*>(g + o0) = vg0;
      if (two) *reinterpret_cast<uint4*>(g + o1) = vg1;
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

// Filter / BN-parameter gradients of the stem from G1 (in dW, written by the wgrad kernel on dz with scale
// 1/grad_scale), (R, S), the BN reductions `acc` (scaled by grad_scale) and the batch moments.  One block of 64
// threads per output channel; dW is overwritten (G1 is read from it unless g1_pair is given), structurally-zero s2d slots are cleared
// (column 7 / 15 of every tap, and rows 8..15 of tap 3 -- programs.py: student_conv1_to_s2d).
//   A = a, D = a*dg/(P*sigma), E = mu*D - a*db/P  (bn_bwd_apply_kernel);  dW = A*G1 - D*(R w + bias*S) + E*S
static __global__ void stem_wgrad_finalize_kernel(const double* __restrict__ rs, const __half* __restrict__ w16,
                                                  const float* __restrict__ bias, const double* __restrict__ acc, double P,
                                                  int C, const float* __restrict__ moments, const float* __restrict__ a,
                                                  float inv_grad_scale, float* __restrict__ dW, float* __restrict__ dbias,
                                                  float* __restrict__ dgamma, float* __restrict__ dbeta,
                                                  const float* __restrict__ g1_pair) {
  __shared__ double wv[kStemT];
  const int k = blockIdx.x, t = threadIdx.x;
  wv[t] = double(__half2float(w16[size_t(k) * kStemT + t]));
  __syncthreads();
  double rw = 0.0;
  for (int u = 0; u < kStemT; ++u) rw += rs[u * kStemT + t] * wv[u];
  const double St = rs[kStemT * kStemT + t];
  const double g2 = rw + double(bias[k]) * St;
  const double av = double(a[k]), mu = double(moments[k]), sg = double(moments[C + k]);
  const double db = acc[k] * double(inv_grad_scale), dg = acc[C + k] * double(inv_grad_scale);
  const double D = av * dg / (P * sg);
  const double E = mu * D - av * db / P;
  const int j = t >> 4, c = t & 15;
  const bool structural_zero = (c & 7) == 7 || (j == 3 && c >= 8);
  const size_t o = size_t(k) * kStemT + t;
  // G1: in place, or the two diagonal blocks of the pixel-pair form [2C][4][32] (the wgrad kernel run on the 32-channel
  // view of the s2d tensor and the [P/2][2C] view of dz): G1[k][j][c] = pair[(0,k)][j][c] + pair[(1,k)][j][16 + c]
  const double g1 = g1_pair ? double(g1_pair[(size_t(k) * 4 + j) * 32 + c]) + double(g1_pair[(size_t(C + k) * 4 + j) * 32 + 16 + c])
                            : double(dW[o]);
  dW[o] = structural_zero ? 0.f : float(av * g1 - D * g2 + E * St);
  if (t == 0) {
    dgamma[k] = float(dg);
    dbeta[k] = float(db);
    if (dbias) dbias[k] = 0.f;   // a bias ahead of train-mode BN has an identically zero gradient
  }
}

// Pixel-pair form of the stem convolution: the s2d tensor [N][HP][OW][16] viewed as [N][HP][OW/2][32] and the
// output [N][OH][OW][C] viewed as [N][OH][OW/2][2C] make conv1 a 4 x 1 convolution with 32 input channels and a
// block-diagonal filter  w2[(e,k)][j][e'*16 + c] = [e == e'] w[k][j][c]  -- same bytes in HBM, half the TMA row
// requests per output pixel (64-byte instead of 32-byte rows), at the price of 2x (free) tensor-core work.
static __global__ void stem_pair_filter_kernel(const __half* __restrict__ w16, int C, __half* __restrict__ w2) {
  const int total = 2 * C * 4 * 32;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int cc = i & 31, j = (i >> 5) & 3, ek = i >> 7;
    const int e = ek / C, k = ek - e * C;
    w2[i] = (cc >> 4) == e ? w16[(k * 4 + j) * 16 + (cc & 15)] : __float2half_rn(0.f);
  }
}

// dst[r*C + c] = src[c] (NULL src -> fill) for r < reps: per-channel epilogue vectors of the pixel-pair form
static __global__ void tile_f32_kernel(const float* __restrict__ src, int C, int reps, float fill, float* __restrict__ dst) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < C * reps; i += gridDim.x * blockDim.x) dst[i] = src ? src[i % C] : fill;
}

}  // namespace xemo
