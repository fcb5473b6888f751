// conv_launch.cuh -- host-side planning (tile shape, pipeline depth, tensor maps) and launch of the
// tcgen05 implicit-GEMM convolution kernels.
#pragma once
#include "conv_fprop.cuh"
#include "tma_host.h"

namespace xemo {

struct ConvGeom {
  int N, H, W, Cin;   // input NHWC
  int Kout, R, S;     // filters [Kout][R][S][Cin]
  int sh, sw;         // stride
  int pt, pb, pl, pr; // MatConvNet pad = [top bottom left right]
  // vl_nnconv output size: floor((H + pt + pb - R) / sh) + 1
  int OH() const { return (H + pt + pb - R) / sh + 1; }
  int OW() const { return (W + pl + pr - S) / sw + 1; }
};

struct ConvEpilogue {
  const float* scale = nullptr;
  const float* shift = nullptr;
  const __half* residual = nullptr;
  int relu = 0;
  __half* out = nullptr;
  float* out_f32 = nullptr;
};

struct ConvPlan {
  CUtensorMap tmA, tmB;
  ConvFpropParams p;
  int bk = 0;
  int grid = 0;
  int smem = 0;
  double flops = 0;
};

constexpr int kSmemBudget = 227 * 1024;

inline int conv_pick_bk(int cin) { return (cin % 64 == 0) ? 64 : (cin % 32 == 0) ? 32 : (cin % 16 == 0) ? 16 : 0; }

// Pick the N tile: a multiple of 16 dividing Kout, <= 256, minimising (waves x per-tile cost).
inline int conv_pick_block_n(int M, int Kout, int k_steps16, int num_sms) {
  int best = 0;
  double best_cost = 1e300;
  const int m_tiles = (M + kConvBlockM - 1) / kConvBlockM;
  for (int bn = 16; bn <= 256; bn += 16) {
    if (Kout % bn) continue;
    const long tiles = long(m_tiles) * (Kout / bn);
    const long waves = (tiles + num_sms - 1) / num_sms;
    // MMA issue cost per K=16 step is ~max(bn,64)/2 cycles at M=128; operand fetch adds a floor; the
    // epilogue is ~bn*6 cycles per tile but overlaps the next tile's main loop.
    const double mainloop = double(k_steps16) * (bn > 96 ? bn * 0.5 : 48.0);
    const double epi = bn * 6.0 + 300.0;
    const double tile_cost = (mainloop > epi ? mainloop : epi) + 200.0;
    const double cost = waves * tile_cost;
    if (cost < best_cost * 0.999) { best_cost = cost; best = bn; }
  }
  return best;
}

inline bool conv_fprop_plan(ConvPlan* plan, const ConvGeom& g, const __half* x, const __half* w,
                            const ConvEpilogue& e, int num_sms, int force_block_n = 0) {
  const int bk = conv_pick_bk(g.Cin);
  if (!bk) { fprintf(stderr, "[xemo] conv: Cin=%d must be a multiple of 16\n", g.Cin); return false; }
  if (g.Kout % 16) { fprintf(stderr, "[xemo] conv: Kout=%d must be a multiple of 16\n", g.Kout); return false; }
  const int OH = g.OH(), OW = g.OW();
  if (OH <= 0 || OW <= 0) return false;
  ConvFpropParams& p = plan->p;
  p.M = g.N * OH * OW;
  p.Kout = g.Kout;
  p.Cin = g.Cin; p.R = g.R; p.S = g.S;
  p.OH = OH; p.OW = OW;
  p.stride_h = g.sh; p.stride_w = g.sw; p.pad_t = g.pt; p.pad_l = g.pl;
  p.kc_blocks = g.Cin / bk;
  const int k_steps16 = g.R * g.S * g.Cin / 16;
  p.block_n = force_block_n ? force_block_n : conv_pick_block_n(p.M, g.Kout, k_steps16, num_sms);
  if (p.block_n <= 0 || g.Kout % p.block_n) return false;
  p.num_m_tiles = (p.M + kConvBlockM - 1) / kConvBlockM;
  p.num_n_tiles = g.Kout / p.block_n;
  const int stage_bytes = conv_stage_bytes(bk, p.block_n);
  int stages = (kSmemBudget - 1024 - 256) / stage_bytes;
  if (stages > 12) stages = 12;
  if (stages < 2) return false;
  p.num_stages = stages;
  p.scale = e.scale; p.shift = e.shift; p.residual = e.residual; p.relu = e.relu;
  p.out = e.out; p.out_f32 = e.out_f32;
  plan->bk = bk;
  plan->smem = stages * stage_bytes + 1024 + (2 * stages + 4) * 8 + 16;
  const int tiles = p.num_m_tiles * p.num_n_tiles;
  plan->grid = tiles < num_sms ? tiles : num_sms;
  plan->flops = 2.0 * double(p.M) * g.Kout * g.R * g.S * g.Cin;

  const CUtensorMapSwizzle swz = swizzle_for_bytes(bk * 2);
  // upper corner: pad_upper - (filter - 1); the effective pad_upper only matters through the number of
  // output pixels per row/column, which must equal OW/OH: use the exact remainder-free value.
  const int upper_w = (OW - 1) * g.sw - g.pl - (g.W - 1);   // == pr' - (S-1) with pr' trimmed to the floor
  const int upper_h = (OH - 1) * g.sh - g.pt - (g.H - 1);
  if (!make_tmap_im2col_nhwc_f16(&plan->tmA, x, g.N, g.H, g.W, g.Cin, -g.pl, -g.pt, upper_w, upper_h, g.sw, g.sh,
                                 uint32_t(bk), uint32_t(kConvBlockM), swz))
    return false;
  if (!make_tmap_2d_f16(&plan->tmB, w, uint64_t(g.Kout), uint64_t(g.R) * g.S * g.Cin, uint64_t(g.R) * g.S * g.Cin,
                        uint32_t(bk), uint32_t(p.block_n), swz))
    return false;
  return true;
}

inline cudaError_t conv_fprop_run(const ConvPlan& plan, cudaStream_t stream) {
  cudaError_t err = cudaSuccess;
  switch (plan.bk) {
    case 64: {
      static bool attr = false;
      if (!attr) { err = cudaFuncSetAttribute(conv_fprop_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBudget); attr = true; }
      if (err != cudaSuccess) return err;
      conv_fprop_kernel<64><<<plan.grid, kConvThreads, plan.smem, stream>>>(plan.tmA, plan.tmB, plan.p);
      break;
    }
    case 32: {
      static bool attr = false;
      if (!attr) { err = cudaFuncSetAttribute(conv_fprop_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBudget); attr = true; }
      if (err != cudaSuccess) return err;
      conv_fprop_kernel<32><<<plan.grid, kConvThreads, plan.smem, stream>>>(plan.tmA, plan.tmB, plan.p);
      break;
    }
    case 16: {
      static bool attr = false;
      if (!attr) { err = cudaFuncSetAttribute(conv_fprop_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBudget); attr = true; }
      if (err != cudaSuccess) return err;
      conv_fprop_kernel<16><<<plan.grid, kConvThreads, plan.smem, stream>>>(plan.tmA, plan.tmB, plan.p);
      break;
    }
    default: return cudaErrorInvalidValue;
  }
  return cudaGetLastError();
}

}  // namespace xemo
