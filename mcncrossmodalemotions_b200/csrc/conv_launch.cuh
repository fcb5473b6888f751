// conv_launch.cuh -- host-side planning (tile shape, pipeline depth, tensor maps) and launch of the
// tcgen05 implicit-GEMM convolution kernels.
#pragma once
#include "conv_fprop.cuh"
#include "conv_wgrad.cuh"
#include <stdlib.h>
#include <string.h>

#include "tma_host.h"

namespace xemo {

struct ConvGeom {
  int N, H, W, Cin;   // input NHWC
  int Kout, R, S;     // filters [Kout][R][S][Cin]
  int sh, sw;         // stride
  int pt, pb, pl, pr; // MatConvNet pad = [top bottom left right]
  // explicit output size (0 = derive).  Used by the parity-decomposed data gradient of strided
  // convolutions, whose sub-problems produce ceil((H - ph) / sh) rows regardless of the padding.
  int oh_override = 0, ow_override = 0;
  // vl_nnconv output size: floor((H + pt + pb - R) / sh) + 1
  int OH() const { return oh_override ? oh_override : (H + pt + pb - R) / sh + 1; }
  int OW() const { return ow_override ? ow_override : (W + pl + pr - S) / sw + 1; }
};

struct ConvEpilogue {
  const float* scale = nullptr;
  const float* shift = nullptr;
  const __half* residual = nullptr;
  int relu = 0;
  __half* out = nullptr;
  float* out_f32 = nullptr;
  int ldc = 0;                 // row pitch of out/out_f32/residual (0 -> Kout)
  // strided-output mode: row address = n*out_sn + oh*out_sh + ow*out_sw (elements)
  int strided_out = 0;
  long long out_sn = 0, out_sh = 0, out_sw = 0;
  long long out_stile = 0;     // 0 -> block_n (N tiles are plain column ranges); else the element offset between N tiles
  int allow_tma_epilogue = 1;  // 0 forces the direct-store epilogue
  // per-(image, channel) scale / shift ([N][Kout] fp32; replaces scale / shift), fp16 TMA-store outputs only
  const float* nc_scale = nullptr;
  const float* nc_shift = nullptr;
};

struct ConvPlan {
  CUtensorMap tmA, tmB, tmOut, tmRes;
  ConvFpropParams p;
  int bk = 0;
  int grid = 0;
  int smem = 0;
  int ctas = 1;     // 2: CTA pairs (cluster of two, tcgen05.mma.cta_group::2, M = 256)
  double flops = 0;
};

// CTA pairs pay when the tile's operand traffic is dominated by the filter: long reductions and wide N tiles.  Measured on
// B200 with the bring-up harness (profiles/r02_selftest_2cta.log): 3x3 c128 28x28 +12 %, 3x3 c256 14x14 +7 %, 3x3 c512 7x7
// +18 %, student conv2 +18 %, conv3 +13 %, a 4096^2 x 16384 GEMM +11 % (1469 TFLOP/s); a short reduction with a wide output
// (1x1 c512 -> k2048, 8 k-iterations) loses 20 %.
// XEMO_CONV_2CTA: 0 = never, 1 (default) = the rule in conv_fprop_plan, 2 = whenever legal (bring-up / A-B measurements);
// xemo_debug_set_conv_pair_mode overrides the environment at run time (tests).
inline int& conv_pair_mode_override() { static int v = -1; return v; }
inline int conv_pair_mode() {
  static const int mode = [] { const char* e = getenv("XEMO_CONV_2CTA"); return e ? (e[0] - '0') : 1; }();
  return conv_pair_mode_override() >= 0 ? conv_pair_mode_override() : mode;
}

constexpr int kSmemBudget = 227 * 1024;

inline int conv_pick_bk(int cin) { return (cin % 64 == 0) ? 64 : (cin % 32 == 0) ? 32 : (cin % 16 == 0) ? 16 : 0; }

// Pick the N tile: a multiple of 16 dividing Kout, <= 256, minimising (waves x per-tile cost).
// Per-tile cost = max(MMA issue, operand fetch, epilogue): the fetch term is what the profiler showed to bind the
// K-heavy layers (profiles/ncu/r01_epilogue_and_issue_studies.txt): TMA fills shared memory at ~60 B/clk/SM with
// 128-byte rows (BK = 64) and proportionally less with 64- / 32-byte rows, and every tile re-fetches its
// block_n x K filter slice unless the filter is resident (one N tile, <= 112 KB).  The term only enters when there are
// at least four full waves of M tiles (wave quantisation does not blur the comparison there; smaller problems keep the
// measured fetch-blind choice).  XEMO_CONV_COSTMODEL=0 restores the fetch-blind model everywhere.  (Applying the fetch term
// to every problem size -- fc6's data gradient and the 7 x 7 teacher layers on 256-wide tiles -- was measured in round 2:
// 5.57 vs 5.54 ms teacher forward, 9.00 vs 9.02 ms student step at batch 256, i.e. neutral; removed.)
inline int conv_pick_block_n(int M, int Kout, int k_steps16, int num_sms, int bk = 64, int k_iters = 0) {
  static const int fetch_mode = [] { const char* e = getenv("XEMO_CONV_COSTMODEL"); return e ? (e[0] - '0') : 1; }();
  const bool fetch_aware = fetch_mode != 0;
  const int min_m_tiles = 4 * num_sms;
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
    double fetch = 0.0;
    if (fetch_aware && k_iters > 0 && m_tiles >= min_m_tiles) {
      const bool resident = (bn == Kout) && (long(k_iters) * conv_b_slot_bytes(bk, bn) <= 112 * 1024) && tiles >= num_sms;
      const double bytes = double(k_iters) * (kConvBlockM * bk * 2 + (resident ? 0 : bn * bk * 2));
      fetch = bytes / (60.0 * bk / 64.0);
    }
    double tile_cost = mainloop > epi ? mainloop : epi;
    if (fetch > tile_cost) tile_cost = fetch;
    tile_cost += 200.0;
    const double cost = waves * tile_cost;
    if (cost < best_cost * 0.999) { best_cost = cost; best = bn; }
  }
  return best;
}

inline bool conv_fprop_plan(ConvPlan* plan, const ConvGeom& g, const __half* x, const __half* w,
                            const ConvEpilogue& e, int num_sms, int force_block_n = 0, bool encode_maps = true) {
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
  p.block_n = force_block_n ? force_block_n : conv_pick_block_n(p.M, g.Kout, k_steps16, num_sms, bk, g.R * g.S * p.kc_blocks);
  if (p.block_n <= 0 || g.Kout % p.block_n) return false;
  p.num_m_tiles = (p.M + kConvBlockM - 1) / kConvBlockM;
  p.num_n_tiles = g.Kout / p.block_n;
  p.scale = e.scale; p.shift = e.shift; p.residual = e.residual; p.relu = e.relu;
  p.out = e.out; p.out_f32 = e.out_f32;
  p.ldc = e.ldc ? e.ldc : g.Kout;
  p.strided_out = e.strided_out; p.out_sn = e.out_sn; p.out_sh = e.out_sh; p.out_sw = e.out_sw;
  p.out_stile = e.out_stile ? e.out_stile : p.block_n;
  p.use_tma_store = (e.allow_tma_epilogue && e.out && !e.strided_out && (p.ldc % 8 == 0)) ? 1 : 0;
  p.use_tma_residual = (p.use_tma_store && e.residual) ? 1 : 0;
  p.epi_cw = (p.block_n % 64 == 0) ? 64 : (p.block_n % 32 == 0) ? 32 : 16;
  p.nc_scale = e.nc_scale; p.nc_shift = e.nc_shift; p.nc_hw = OH * OW;
  const int nc_bytes = e.nc_scale ? kEpiGroups * 2 * kNcSlots * 256 * 4 : 0;
  if (e.nc_scale) {
    // the fast epilogue only; a 128-row tile may touch at most kNcSlots images
    if (!e.nc_shift || !p.use_tma_store || e.out_f32 || bk != 64 || (kConvBlockM + p.nc_hw - 2) / p.nc_hw + 1 > kNcSlots) return false;
  }
  // staging buffers per epilogue group: two when the pipeline still keeps enough stages for the K loop
  const int k_iters = g.R * g.S * p.kc_blocks;
  const int want_stages = k_iters + 1 < 3 ? k_iters + 1 : 3;
  const int tiles = p.num_m_tiles * p.num_n_tiles;
  // resident filter (one load per CTA instead of one per tile): a single N tile, at most 112 KB (the teacher stem in
  // pixel-pair form: 7 x 16 KB), and at least one tile per CTA.  XEMO_CONV_BRES=0 disables (A/B measurements).
  static const bool bres_enabled = [] { const char* e = getenv("XEMO_CONV_BRES"); return !(e && e[0] == '0'); }();
  const int bres_bytes = k_iters * conv_b_slot_bytes(bk, p.block_n);
  p.b_resident = (bres_enabled && p.num_n_tiles == 1 && bres_bytes <= 112 * 1024 && tiles >= num_sms) ? 1 : 0;
  // CTA pairs: legal when the N tile splits into two halves of whole 16-row groups, the filter is not resident and the
  // per-(image, channel) epilogue is off (its image slots assume in-range M tiles)
  const bool pair_legal = !p.b_resident && !e.nc_scale && p.block_n % 32 == 0 && p.num_m_tiles >= 2 && num_sms % 2 == 0;
  const bool pair_wanted = conv_pair_mode() == 2 || (conv_pair_mode() == 1 && k_iters >= 16 && p.block_n >= 128 && tiles >= num_sms);
  plan->ctas = (pair_legal && pair_wanted) ? 2 : 1;
  p.num_pair_tiles = ((p.num_m_tiles + 1) / 2) * p.num_n_tiles;
  const int stage_bytes = conv_stage_bytes(bk, p.block_n / plan->ctas, p.b_resident);
  const int fixed_bytes = p.b_resident ? bres_bytes : 0;
  int stages = 0, epi_bytes = 0;
  for (int bufs = 2; bufs >= 1; --bufs) {
    epi_bytes = (p.use_tma_store ? kEpiGroups * bufs * kEpiStageBytes : 0) + (p.use_tma_residual ? kEpiGroups * bufs * kEpiStageBytes : 0);
    stages = (kSmemBudget - 1024 - 256 - 4096 - nc_bytes - epi_bytes - fixed_bytes) / stage_bytes;
    p.epi_bufs = bufs;
    if (stages >= want_stages) break;
  }
  if (stages > 12) stages = 12;
  if (stages < 2) return false;
  p.num_stages = stages;
  plan->bk = bk;
  plan->smem = stages * stage_bytes + fixed_bytes + epi_bytes + 1024 + kEpiGroups * 2 * 256 * 4 + nc_bytes + (2 * stages + 9) * 8 + 16;
  plan->grid = tiles < num_sms ? tiles : num_sms;
  if (plan->ctas == 2) plan->grid = 2 * (p.num_pair_tiles < num_sms / 2 ? p.num_pair_tiles : num_sms / 2);
  plan->flops = 2.0 * double(p.M) * g.Kout * g.R * g.S * g.Cin;
  if (!encode_maps) return true;

  const CUtensorMapSwizzle swz = swizzle_for_bytes(bk * 2);
  // upper corner: pad_upper - (filter - 1); the effective pad_upper only matters through the number of
  // output pixels per row/column, which must equal OW/OH: use the exact remainder-free value.
  const int upper_w = (OW - 1) * g.sw - g.pl - (g.W - 1);   // == pr' - (S-1) with pr' trimmed to the floor
  const int upper_h = (OH - 1) * g.sh - g.pt - (g.H - 1);
  if (!make_tmap_im2col_nhwc_f16(&plan->tmA, x, g.N, g.H, g.W, g.Cin, -g.pl, -g.pt, upper_w, upper_h, g.sw, g.sh,
                                 uint32_t(bk), uint32_t(kConvBlockM), swz))
    return false;
  if (!make_tmap_2d_f16(&plan->tmB, w, uint64_t(g.Kout), uint64_t(g.R) * g.S * g.Cin, uint64_t(g.R) * g.S * g.Cin,
                        uint32_t(bk), uint32_t(p.block_n / plan->ctas), swz))
    return false;
  memset(&plan->tmOut, 0, sizeof(CUtensorMap));
  memset(&plan->tmRes, 0, sizeof(CUtensorMap));
  if (p.use_tma_store &&
      !make_tmap_2d_f16(&plan->tmOut, e.out, uint64_t(p.M), uint64_t(g.Kout), uint64_t(p.ldc), uint32_t(p.epi_cw),
                        uint32_t(kConvBlockM), swizzle_for_bytes(p.epi_cw * 2)))
    return false;
  if (p.use_tma_residual &&
      !make_tmap_2d_f16(&plan->tmRes, e.residual, uint64_t(p.M), uint64_t(g.Kout), uint64_t(p.ldc),
                        uint32_t(p.epi_cw), uint32_t(kConvBlockM), swizzle_for_bytes(p.epi_cw * 2)))
    return false;
  return true;
}

template <int BK>
inline cudaError_t conv_fprop_run_pair(const ConvPlan& plan, cudaStream_t stream) {
  static bool attr = false;
  cudaError_t err = cudaSuccess;
  if (!attr) { err = cudaFuncSetAttribute(conv_fprop_kernel<BK, false, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBudget); attr = true; }
  if (err != cudaSuccess) return err;
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(plan.grid);
  cfg.blockDim = dim3(kConvThreads);
  cfg.dynamicSmemBytes = plan.smem;
  cfg.stream = stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, conv_fprop_kernel<BK, false, 2>, plan.tmA, plan.tmB, plan.tmOut, plan.tmRes, plan.p);
}

inline cudaError_t conv_fprop_run(const ConvPlan& plan, cudaStream_t stream) {
  cudaError_t err = cudaSuccess;
  if (plan.ctas == 2) {
    switch (plan.bk) {
      case 64: return conv_fprop_run_pair<64>(plan, stream);
      case 32: return conv_fprop_run_pair<32>(plan, stream);
      case 16: return conv_fprop_run_pair<16>(plan, stream);
      default: return cudaErrorInvalidValue;
    }
  }
  if (plan.p.nc_scale) {   // per-(image, channel) epilogue: BK = 64 only (checked by the plan)
    static bool attr = false;
    if (!attr) { err = cudaFuncSetAttribute(conv_fprop_kernel<64, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBudget); attr = true; }
    if (err != cudaSuccess) return err;
    conv_fprop_kernel<64, true><<<plan.grid, kConvThreads, plan.smem, stream>>>(plan.tmA, plan.tmB, plan.tmOut, plan.tmRes, plan.p);
    return cudaGetLastError();
  }
  switch (plan.bk) {
    case 64: {
      static bool attr = false;
      if (!attr) { err = cudaFuncSetAttribute(conv_fprop_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBudget); attr = true; }
      if (err != cudaSuccess) return err;
      conv_fprop_kernel<64><<<plan.grid, kConvThreads, plan.smem, stream>>>(plan.tmA, plan.tmB, plan.tmOut, plan.tmRes, plan.p);
      break;
    }
    case 32: {
      static bool attr = false;
      if (!attr) { err = cudaFuncSetAttribute(conv_fprop_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBudget); attr = true; }
      if (err != cudaSuccess) return err;
      conv_fprop_kernel<32><<<plan.grid, kConvThreads, plan.smem, stream>>>(plan.tmA, plan.tmB, plan.tmOut, plan.tmRes, plan.p);
      break;
    }
    case 16: {
      static bool attr = false;
      if (!attr) { err = cudaFuncSetAttribute(conv_fprop_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBudget); attr = true; }
      if (err != cudaSuccess) return err;
      conv_fprop_kernel<16><<<plan.grid, kConvThreads, plan.smem, stream>>>(plan.tmA, plan.tmB, plan.tmOut, plan.tmRes, plan.p);
      break;
    }
    default: return cudaErrorInvalidValue;
  }
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// filter gradient
struct WgradPlan {
  CUtensorMap tmY, tmX;
  ConvWgradParams p;
  int grid = 0;
  int smem = 0;
  double flops = 0;
};

inline int pick_chunk(int n) { return (n % 64 == 0) ? 64 : (n % 32 == 0) ? 32 : 16; }

// x: NHWC fp16 [N][H][W][Cin] (Cin % 16 == 0); dy: [P][ldy] fp16 (ldy % 16 == 0, Kout <= ldy);
// dF: [Kout][R][S][Cin] fp32, accumulated into.
// deterministic: no split of the pixel reduction -- every dF element is produced by exactly one work item, so its value
// does not depend on the order in which `red.global.add` operations land (at the price of idle SMs on layers with few
// output tiles)
inline bool conv_wgrad_plan(WgradPlan* plan, const ConvGeom& g, const __half* x, const __half* dy, int ldy, float* dF,
                            float scale, int num_sms, bool encode_maps = true, bool deterministic = false) {
  if (g.Cin % 16 || ldy % 16 || g.Kout > ldy) {
    fprintf(stderr, "[xemo] wgrad: Cin=%d / ldy=%d must be multiples of 16 and Kout=%d <= ldy\n", g.Cin, ldy, g.Kout);
    return false;
  }
  const int OH = g.OH(), OW = g.OW();
  if (OH <= 0 || OW <= 0) return false;
  ConvWgradParams& p = plan->p;
  p.P = g.N * OH * OW;
  p.Kout = g.Kout;
  p.ldy = ldy;
  p.Cin = g.Cin; p.R = g.R; p.S = g.S;
  p.OH = OH; p.OW = OW;
  p.stride_h = g.sh; p.stride_w = g.sw; p.pad_t = g.pt; p.pad_l = g.pl;
  p.chunk_a = ldy >= 64 ? 64 : pick_chunk(ldy);
  p.chunk_b = pick_chunk(g.Cin);
  // channels per sub-tile: the largest divisor of Cin that is a multiple of chunk_b and <= 256
  int block_c = p.chunk_b;
  for (int c = p.chunk_b; c <= 256 && c <= g.Cin; c += p.chunk_b)
    if (g.Cin % c == 0) block_c = c;
  p.block_c = block_c;
  p.c_tiles = g.Cin / block_c;
  const int total_sub = g.R * g.S * p.c_tiles;
  p.m_tiles = (g.Kout + kWgBlockM - 1) / kWgBlockM;
  // Stage shape.  The kernel is bound by the shared-memory fill rate (TMA rows), so among the shapes that fit --
  // mt 128-kout dY tiles and T sub-tiles of block_c channels per stage, mt * T * block_c <= 512 TMEM columns, at
  // least 3 stages (4 for stages under 48 KB) -- take the one with the most FLOP per staged byte,
  //   mt * T * 128 * block_c / (mt * 128 + T * block_c)  per pixel,
  // preferring the larger TMA boxes (128 > 64 pixels); short reductions (fc layers) stay at 32 pixels, mt = 1.
  const int P_total = g.N * OH * OW;
  const int budget = kSmemBudget - 2048;
  // (measured on B200: two-tile items win when one sub-tile spans all input channels -- conv2 / conv3 / conv5 of the
  // student, 5-15 % -- and lose when the channels are split over sub-tiles -- conv4, -25 %; XEMO_WGRAD_MT2=0 disables)
  static const bool mt2_enabled = [] { const char* e = getenv("XEMO_WGRAD_MT2"); return !(e && e[0] == '0'); }();
  int pix = 0, T = 1, mt = 1;
  double best = -1.0;
  for (int cm = (mt2_enabled && p.m_tiles >= 2 && p.c_tiles == 1) ? 2 : 1; cm >= 1; --cm) {
    const int t_cap = 512 / (cm * block_c);
    for (int ct = t_cap < total_sub ? t_cap : total_sub; ct >= 1; --ct) {
      for (int cand = 128; cand >= 64; cand >>= 1) {
        if (P_total < cand * 8) continue;
        const int sb = wgrad_stage_bytes(ct, block_c, cand, cm);
        const int need = sb >= 48 * 1024 ? 3 : 4;
        if (sb * need > budget) continue;
        double score = double(cm) * ct * 128.0 * block_c / (cm * 128.0 + double(ct) * block_c);
        if (cand == 128) score *= 1.05;
        if (score > best) { best = score; pix = cand; T = ct; mt = cm; }
      }
    }
  }
  if (pix == 0) {  // short reduction: 32-pixel stages
    const int t_max = (512 / block_c) < total_sub ? (512 / block_c) : total_sub;
    pix = 32; T = t_max; mt = 1;
    while (T > 1 && wgrad_stage_bytes(T, block_c, 32) * 3 > budget) --T;
  }
  p.T = T;
  p.pix = pix;
  p.mt = mt;
  p.m_items = (p.m_tiles + mt - 1) / mt;
  p.groups = (total_sub + T - 1) / T;
  const int pix_blocks = (p.P + p.pix - 1) / p.pix;
  const int base_items = p.m_items * p.groups;
  // split the pixel reduction so that there are ~2 items per SM, each at least 8 pixel blocks long
  // (rounded DOWN so that the item count stays within two full waves of the persistent grid)
  int splits = (2 * num_sms) / base_items;
  const int max_splits = (pix_blocks + 7) / 8;
  if (splits > max_splits) splits = max_splits;
  if (splits < 1 || deterministic) splits = 1;
  p.pix_blocks_per_split = (pix_blocks + splits - 1) / splits;
  p.splits = (pix_blocks + p.pix_blocks_per_split - 1) / p.pix_blocks_per_split;
  const int stage_bytes = wgrad_stage_bytes(T, block_c, p.pix, mt);
  int stages = (kSmemBudget - 2048) / stage_bytes;
  if (stages > 8) stages = 8;
  if (stages < 2) return false;
  p.num_stages = stages;
  p.dF = dF;
  p.scale = scale;
  p.red_vec = p.splits <= 64 ? 1 : 0;
  plan->smem = stages * stage_bytes + 1024 + (2 * stages + 2) * 8 + 16;
  const int items = p.m_items * p.groups * p.splits;
  plan->grid = items < num_sms ? items : num_sms;
  plan->flops = 2.0 * double(p.P) * g.Kout * g.R * g.S * g.Cin;
  if (!encode_maps) return true;
  const int upper_w = (OW - 1) * g.sw - g.pl - (g.W - 1);
  const int upper_h = (OH - 1) * g.sh - g.pt - (g.H - 1);
  if (!make_tmap_im2col_nhwc_f16(&plan->tmX, x, g.N, g.H, g.W, g.Cin, -g.pl, -g.pt, upper_w, upper_h, g.sw, g.sh,
                                 uint32_t(p.chunk_b), uint32_t(p.pix), swizzle_for_bytes(p.chunk_b * 2)))
    return false;
  if (!make_tmap_2d_f16(&plan->tmY, dy, uint64_t(p.P), uint64_t(ldy), uint64_t(ldy), uint32_t(p.chunk_a),
                        uint32_t(p.pix), swizzle_for_bytes(p.chunk_a * 2)))
    return false;
  return true;
}

inline cudaError_t conv_wgrad_run(const WgradPlan& plan, cudaStream_t stream) {
  static bool attr = false;
  if (!attr) {
    cudaError_t err = cudaFuncSetAttribute(conv_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBudget);
    if (err != cudaSuccess) return err;
    attr = true;
  }
  conv_wgrad_kernel<<<plan.grid, kWgThreads, plan.smem, stream>>>(plan.tmY, plan.tmX, plan.p);
  return cudaGetLastError();
}

}  // namespace xemo
