// se_gate.cuh -- the squeeze-and-excitation gate (mcnExtraLayers: GlobalPooling -> Conv 1x1 -> ReLU -> Conv 1x1 -> Sigmoid)
// as ONE kernel per SE block, written for latency: per block of kSeSpb samples the gate is three dependent matrix-vector
// products whose weights (64 KB ... 2 MB) stream from L2, so what counts is how many loads each SM keeps in flight and
// how many SMs pull weights at once.
//
//   * every phase issues its weight loads in batches of >= 8 independent 16-byte loads per thread before the FMAs;
//   * a thread-block CLUSTER of K CTAs (1, 2, 4; chosen so that groups * K stays within the SMs) shares one group of
//     samples: CTA r computes 1/K of each phase's outputs from 1/K of the weights and writes them into the shared
//     memory of ALL K CTAs (st.shared::cluster through mapa), with a cluster barrier between phases.  At 256 faces
//     K = 1 (128 groups); at the 32 faces per GPU of the 8-GPU strong-scaling point K = 4 turns 16 busy SMs into 64.
//     Measured (teacher forward graph, ms, K capped at 1 / 2 / 4 / 8): 32 faces 1.207 / 1.152 / 1.124 / 1.245,
//     64 faces 1.728 / 1.673 / 1.645 / --, 128 faces 2.765 / 2.711 / -- / --; the round-1 kernels: 1.349 at 32 faces.
//     Clusters of 8 CTAs x 1024 threads lose more to placement and barriers than they gain: K <= 4.
//
//   kLin = false : s [N][C] (se_squeeze of u)               -> gate [N][C]
//   kLin = true  : m2 [N][Cm] (se_squeeze of the 3x3 output) -> s = a3 * (W3 m2) + b3 -> gate -> nc_scale = gate * a3,
//                  nc_shift = gate * b3, the per-(image, channel) epilogue vectors of conv_fprop_kernel<64, true>
//                  (SE by linearity)
// w3: [C][Cm] fp16 (the KRSC 1x1 expand filter); w1: [Cr][C] fp32; w2t: [Cr][C] fp32 (second FC transposed).
#pragma once
#include <stdlib.h>

#include "hbm_kernels.cuh"
#include "xemo_ptx.cuh"

namespace xemo {

struct SeGateParams {
  const float* in;   // kLin ? m2 : s
  int N, C, Cm, Cr;
  const __half* w3;
  const float* a3;
  const float* b3;
  const float* w1;
  const float* b1;
  const float* w2t;
  const float* b2;
  float* out0;       // kLin ? nc_scale : gate
  float* out1;       // kLin ? nc_shift : unused
};

__device__ __forceinline__ uint32_t cluster_num_ctas() {
  uint32_t n;
  asm volatile("mov.u32 %0, %%cluster_nctaid.x;" : "=r"(n));
  return n;
}
// store into the same shared-memory offset of CTA `rank` of the cluster
__device__ __forceinline__ void st_cluster_f32(float* local, uint32_t rank, float v) {
  uint32_t remote;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(smem_u32(local)), "r"(rank));
  asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(remote), "f"(v) : "memory");
}

constexpr int kSeGateThreads = 1024;

// ranges of hidden units in the last phase's sum (see there) and thread groups that share one channel's ranges
__host__ __device__ inline int se_gate_ranges(int C, int Cr) {
  int pc = 1;
  while (pc * 2 <= Cr / 8 && pc * 2 <= 4096 / C) pc *= 2;
  return pc;
}
__host__ __device__ inline int se_gate_groups(int pc, int nout) {
  int groups = 1;
  while (groups * 2 <= pc && groups * 2 * nout <= kSeGateThreads) groups *= 2;
  return groups;
}

// dynamic smem (floats): kSeSpb * ((kLin ? Cm : 0) + C + Cr) + (groups > 1 ? ranges * kSeSpb * C / K : 0)
template <bool kLin, int CM64>
static __global__ void __launch_bounds__(kSeGateThreads, 1) se_gate_cluster_kernel(SeGateParams p) {
  extern __shared__ __align__(16) float se_sm[];
  const int K = int(cluster_num_ctas()), rank = int(cluster_ctarank());
  const int n0 = (blockIdx.x / K) * kSeSpb;
  const int C = p.C, Cr = p.Cr, Cm = kLin ? p.Cm : 0;
  float* mv = se_sm;                       // [kSeSpb][Cm]
  float* sv = mv + kSeSpb * Cm;            // [kSeSpb][C]
  float* hid = sv + kSeSpb * C;            // [kSeSpb][Cr]
  float* red = hid + kSeSpb * Cr;          // [ranges][kSeSpb][C / K] partial sums of the last phase
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  constexpr int nwarps = kSeGateThreads / 32;

  {
    const int len = kLin ? Cm : C;
    float* dst = kLin ? mv : sv;
    for (int i = tid; i < kSeSpb * len; i += kSeGateThreads) {
      const int q = i / len, j = i - q * len;
      dst[i] = (n0 + q < p.N) ? __ldg(p.in + size_t(n0 + q) * len + j) : 0.f;
    }
  }
  // (a cluster: the peers' shared memory may only be written once they run; alone: the staged vector must be visible)
  if (K > 1) cluster_sync_all(); else __syncthreads();

  if constexpr (kLin) {
    // s[c] = a3[c] * (W3[c, :] . m2) + b3[c] for this CTA's C / K channels: 8 lanes per filter row (8 x 16 B = one
    // 128-byte line per load), 4 rows per warp and round, kRB rounds' loads in flight
    constexpr int kRB = 8 / CM64;
    const int nout = C / K, c_lo = rank * nout;
    const int sub = lane >> 3, l8 = lane & 7;
    for (int base = 0; base < nout; base += nwarps * 4 * kRB) {
      uint4 w[kRB][CM64];
      float av[kRB], bv[kRB];
#pragma unroll
      for (int b = 0; b < kRB; ++b) {
        const int o = base + (b * nwarps + warp) * 4 + sub;
        const bool ok = o < nout;
        const int c = c_lo + (ok ? o : 0);
        const uint4* row = reinterpret_cast<const uint4*>(p.w3 + size_t(c) * Cm);
#pragma unroll
        for (int i = 0; i < CM64; ++i) w[b][i] = ok ? __ldg(row + i * 8 + l8) : make_uint4(0, 0, 0, 0);
        av[b] = __ldg(p.a3 + c);
        bv[b] = __ldg(p.b3 + c);
      }
#pragma unroll
      for (int b = 0; b < kRB; ++b) {
        const int o = base + (b * nwarps + warp) * 4 + sub;
        float t[kSeSpb];
#pragma unroll
        for (int q = 0; q < kSeSpb; ++q) t[q] = 0.f;
#pragma unroll
        for (int i = 0; i < CM64; ++i) {
          const __half2* h2 = reinterpret_cast<const __half2*>(&w[b][i]);
          const float2 f0 = __half22float2(h2[0]), f1 = __half22float2(h2[1]), f2 = __half22float2(h2[2]), f3 = __half22float2(h2[3]);
#pragma unroll
          for (int q = 0; q < kSeSpb; ++q) {
            const float4 m0 = *reinterpret_cast<const float4*>(mv + q * Cm + (i * 8 + l8) * 8);
            const float4 m1 = *reinterpret_cast<const float4*>(mv + q * Cm + (i * 8 + l8) * 8 + 4);
            t[q] = fmaf(f0.x, m0.x, fmaf(f0.y, m0.y, fmaf(f1.x, m0.z, fmaf(f1.y, m0.w, t[q]))));
            t[q] = fmaf(f2.x, m1.x, fmaf(f2.y, m1.y, fmaf(f3.x, m1.z, fmaf(f3.y, m1.w, t[q]))));
          }
        }
#pragma unroll
        for (int q = 0; q < kSeSpb; ++q) {
          t[q] += __shfl_xor_sync(0xffffffffu, t[q], 4);
          t[q] += __shfl_xor_sync(0xffffffffu, t[q], 2);
          t[q] += __shfl_xor_sync(0xffffffffu, t[q], 1);
        }
        if (l8 == 0 && o < nout) {
#pragma unroll
          for (int q = 0; q < kSeSpb; ++q) {
            const float v = fmaf(av[b], t[q], bv[b]);
            float* dst = sv + q * C + c_lo + o;
            if (K == 1) *dst = v;
            else
              for (int r = 0; r < K; ++r) st_cluster_f32(dst, uint32_t(r), v);
          }
        }
      }
    }
    if (K > 1) cluster_sync_all(); else __syncthreads();
  }

  {
    // hidden[j] = relu(W1[j, :] . s + b1[j]) for this CTA's Cr / K units: one warp per unit, 8 float4 loads in flight
    const int nout = Cr / K, j_lo = rank * nout;
    for (int jj = warp; jj < nout; jj += nwarps) {
      const int j = j_lo + jj;
      const float* wr = p.w1 + size_t(j) * C;
      float t[kSeSpb];
#pragma unroll
      for (int q = 0; q < kSeSpb; ++q) t[q] = 0.f;
#pragma unroll 8
      for (int c = lane * 4; c < C; c += 128) {   // C is a multiple of 128 (checked by the wrapper)
        const float4 w = __ldg(reinterpret_cast<const float4*>(wr + c));
#pragma unroll
        for (int q = 0; q < kSeSpb; ++q) {
          const float4 v = *reinterpret_cast<const float4*>(sv + q * C + c);
          t[q] = fmaf(w.x, v.x, fmaf(w.y, v.y, fmaf(w.z, v.z, fmaf(w.w, v.w, t[q]))));
        }
      }
      const float bias = p.b1 ? __ldg(p.b1 + j) : 0.f;
#pragma unroll
      for (int q = 0; q < kSeSpb; ++q) {
        const float r = fmaxf(warp_sum(t[q]) + bias, 0.f);
        if (lane == 0) {
          float* dst = hid + q * Cr + j;
          if (K == 1) *dst = r;
          else
            for (int k = 0; k < K; ++k) st_cluster_f32(dst, uint32_t(k), r);
        }
      }
    }
    if (K > 1) cluster_sync_all(); else __syncthreads();   // (last remote access: a CTA may leave whenever it is done)
  }

  {
    // gate[c] = sigmoid(W2[c, :] . hidden + b2[c]) for this CTA's C / K channels: a thread per channel (coalesced
    // rows of w2t); with fewer channels than threads the ranges of hidden units go to separate thread groups
    const int nout = C / K, c_lo = rank * nout;
    auto finish = [&](int c, const float (&t)[kSeSpb]) {
      float a = 0.f, b = 0.f;
      if constexpr (kLin) {
        a = __ldg(p.a3 + c);
        b = __ldg(p.b3 + c);
      }
#pragma unroll
      for (int q = 0; q < kSeSpb; ++q)
        if (n0 + q < p.N) {
          const float g = 1.f / (1.f + __expf(-t[q]));
          if constexpr (kLin) {
            p.out0[size_t(n0 + q) * C + c] = g * a;
            p.out1[size_t(n0 + q) * C + c] = g * b;
          } else {
            p.out0[size_t(n0 + q) * C + c] = g;
          }
        }
    };
    // The sum over the hidden units is DEFINED as b2 + p_0 + p_1 + ... over `pc` equal ranges of units (each summed in
    // order from zero), pc a function of (C, Cr) only: whatever the cluster size -- it follows the batch -- a face gets
    // bit-identical gates.  pc = min(Cr / 8, 4096 / C) rounded down to a power of two: every range keeps a batch of 8
    // loads and the smallest slice (C / 4 channels, 1024 threads) still has a thread group per range.
    const int pc = se_gate_ranges(C, Cr);
    const int jper = (Cr + pc - 1) / pc;
    auto range_sum = [&](int c, int part, float (&t)[kSeSpb]) {
#pragma unroll
      for (int q = 0; q < kSeSpb; ++q) t[q] = 0.f;
      const int j_hi = min(Cr, (part + 1) * jper);
#pragma unroll 8
      for (int j = part * jper; j < j_hi; ++j) {
        const float w = __ldg(p.w2t + size_t(j) * C + c);
#pragma unroll
        for (int q = 0; q < kSeSpb; ++q) t[q] = fmaf(w, hid[q * Cr + j], t[q]);
      }
    };
    const int groups = se_gate_groups(pc, nout);
    if (groups == 1) {
      for (int o = tid; o < nout; o += kSeGateThreads) {
        const int c = c_lo + o;
        float t[kSeSpb], r[kSeSpb];
#pragma unroll
        for (int q = 0; q < kSeSpb; ++q) t[q] = p.b2 ? __ldg(p.b2 + c) : 0.f;
        for (int part = 0; part < pc; ++part) {
          range_sum(c, part, r);
#pragma unroll
          for (int q = 0; q < kSeSpb; ++q) t[q] += r[q];
        }
        finish(c, t);
      }
    } else {
      const int grp = tid / nout, o = tid - grp * nout;
      const int c = c_lo + o;
      const int per_group = pc / groups;
      if (grp < groups) {
        for (int part = grp * per_group; part < (grp + 1) * per_group; ++part) {
          float r[kSeSpb];
          range_sum(c, part, r);
#pragma unroll
          for (int q = 0; q < kSeSpb; ++q) red[(part * kSeSpb + q) * nout + o] = r[q];
        }
      }
      __syncthreads();
      if (grp == 0) {
        float t[kSeSpb];
#pragma unroll
        for (int q = 0; q < kSeSpb; ++q) {
          t[q] = p.b2 ? __ldg(p.b2 + c) : 0.f;
          for (int k = 0; k < pc; ++k) t[q] += red[(k * kSeSpb + q) * nout + o];
        }
        finish(c, t);
      }
    }
  }
}

// cluster size for `groups` sample groups: the largest power of two <= 4 that keeps groups * K within the SMs and
// leaves every CTA at least 32 channels and 2 hidden units
static inline int se_gate_cluster_size(int groups, int C, int Cr, int num_sms) {
  int K = 1;
  while (K < 4 && groups * (K * 2) <= num_sms && C % (K * 2) == 0 && C / (K * 2) >= 32 && Cr % (K * 2) == 0 && Cr / (K * 2) >= 2) K *= 2;
  return K;
}

template <bool kLin>
static cudaError_t se_gate_cluster_launch(const SeGateParams& p, int num_sms, cudaStream_t stream, int force_k = 0) {
  const int groups = (p.N + kSeSpb - 1) / kSeSpb;
  static const int env_k = [] { const char* e = getenv("XEMO_SE_GATE_K"); return e ? atoi(e) : 0; }();   // A/B knob: cap on K
  int K = force_k > 0 ? force_k : se_gate_cluster_size(groups, p.C, p.Cr, num_sms);
  if (env_k > 0 && K > env_k) K = env_k;
  if (p.C % K || p.Cr % K) return cudaErrorInvalidValue;
  const int pc = se_gate_ranges(p.C, p.Cr), nout = p.C / K;
  const size_t red = se_gate_groups(pc, nout) > 1 ? size_t(pc) * kSeSpb * nout : 0;
  const size_t smem = (size_t(kSeSpb) * ((kLin ? p.Cm : 0) + p.C + p.Cr) + red) * sizeof(float);
  if (smem > 48 * 1024) return cudaErrorInvalidValue;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(unsigned(groups * K));
  cfg.blockDim = dim3(kSeGateThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = unsigned(K);
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  if constexpr (kLin) {
    switch (p.Cm / 64) {
      case 1: return cudaLaunchKernelEx(&cfg, se_gate_cluster_kernel<true, 1>, p);
      case 2: return cudaLaunchKernelEx(&cfg, se_gate_cluster_kernel<true, 2>, p);
      case 4: return cudaLaunchKernelEx(&cfg, se_gate_cluster_kernel<true, 4>, p);
      case 8: return cudaLaunchKernelEx(&cfg, se_gate_cluster_kernel<true, 8>, p);
      default: return cudaErrorInvalidValue;
    }
  } else {
    return cudaLaunchKernelEx(&cfg, se_gate_cluster_kernel<false, 1>, p);
  }
}

}  // namespace xemo
