// conv_fprop.cuh -- implicit-GEMM convolution forward on tcgen05 tensor cores (sm_100a).
//
// Replaces (behind the C ABI in include/xemo.h) the arithmetic of MatConvNet's vl_nnconv forward,
// the operator that dagnn.Conv.forward calls from dag.eval at
//   /root/reference/emoVoxCeleb/fetch_emovoxceleb_imdb.m:129, external/compute_visual_feats.m:90,
//   external/compute_audio_feats.m:126 and inside cnn_train_dag (emoVoxCeleb/run_distillation.m:170).
// The same kernel computes the data gradient (vl_nnconv backward, DX) of stride-1 convolutions by
// running on dY with flipped/transposed filters, and of strided ones as a sum over output parities
// written through the strided-output mode.
//
// GEMM view:  D[M = N*OH*OW, Kout] = sum_{r,s,c} X[n, oh*sy+r-pt, ow*sx+s-pl, c] * F[kout, r, s, c]
//   A (activations) : NHWC fp16, fetched by TMA in im2col mode -- 128 consecutive output pixels x BK
//                     channels per (r,s) filter tap, zero-filled outside the image, 128/64/32B-swizzled.
//   B (filters)     : [Kout][R][S][C] fp16 (K-major), fetched by tiled TMA, same swizzle.
//   D (accumulator) : fp32 in TMEM, 128 lanes x block_n columns, double-buffered (2 x 256 columns) so
//                     the epilogue of tile i overlaps the main loop of tile i+1.
// Warp roles (320 threads, 1 CTA / SM, persistent over tiles):
//   warp 0   : TMA producer (one elected lane)
//   warp 1   : tcgen05.mma issuer (one elected lane)
//   warps 2-9: epilogue, as two independent groups of four warps (one warp per TMEM lane quadrant) that
//              take alternate column chunks of the accumulator tiles: tcgen05.ld, per-channel
//              scale/shift (folded test-mode BN or bias, staged in smem), optional residual add,
//              optional ReLU; fp16 chunk staged in swizzled smem and written with a TMA store (residual
//              chunks prefetched with TMA loads), or stored directly (fp32 / strided outputs).  While
//              one group waits for its residual / drains its store the other one computes, which is
//              what the HBM-bound 1x1 layers (K = 64: one MMA k-block per 128 x 256 tile) need.
//              Warp 2 also owns TMEM alloc/free.
#pragma once
#include <type_traits>

#include "xemo_ptx.cuh"

namespace xemo {

constexpr int kConvBlockM = 128;
constexpr int kConvThreads = 320;
constexpr int kEpiGroups = 2;
constexpr int kConvTmemCols = 512;
constexpr int kEpiStageBytes = kConvBlockM * 64 * 2;  // one [128 x 64] fp16 staging tile

struct ConvFpropParams {
  int M;       // N*OH*OW output pixels
  int Kout;    // output channels
  int Cin, R, S;
  int OH, OW;
  int stride_h, stride_w, pad_t, pad_l;
  int block_n;      // N tile: multiple of 16, <= 256, divides Kout
  int num_m_tiles;  // ceil(M / 128)
  int num_n_tiles;  // Kout / block_n
  int kc_blocks;    // Cin / BK
  int num_stages;
  // epilogue: y = acc * scale[k] + shift[k] (+ residual) ; relu ; store
  const float* scale;       // [Kout] or nullptr (=1)
  const float* shift;       // [Kout] or nullptr (=0)
  const __half* residual;   // [M, ldc] or nullptr
  int relu;
  __half* out;              // fp16 output or nullptr
  float* out_f32;           // [M, ldc] fp32 or nullptr
  int ldc;                  // row pitch (elements) of out / out_f32 / residual in dense mode
  // strided-output mode (dgrad of strided convs): row address = n*out_sn + oh*out_sh + ow*out_sw
  int strided_out;
  long long out_sn, out_sh, out_sw;
  long long out_stile;      // strided-output mode: element offset of N tile t is t * out_stile (block_n: plain columns)
  // TMA epilogue
  int use_tma_store;        // fp16 `out` written through tmOut
  int use_tma_residual;     // residual read through tmRes
  int epi_cw;               // staging chunk width in columns: 64 / 32 / 16
  int epi_bufs;             // staging (and residual) buffers per epilogue group: 1 or 2
  // resident filter: when the whole [Kout x R*S*Cin] filter is one N tile and fits beside the pipeline, it is loaded
  // ONCE per CTA (k_iters slots) instead of once per tile -- for the K = 64..576 layers at 56 x 56 and the stems the
  // per-tile filter re-fetch is 40-65 % of the L2->smem traffic and of the TMA row requests
  int b_resident;
  // per-(image, channel) epilogue vectors (kNC instantiation only, see conv_fprop_kernel): the SE block's
  // excite fused into its expand convolution,  y = relu(gate[n,c] * (a[c]*acc + b[c]) + shortcut)
  //   = relu(nc_scale[n,c]*acc + nc_shift[n,c] + residual),  nc_scale = gate*a, nc_shift = gate*b  ([N][Kout] fp32)
  const float* nc_scale;
  const float* nc_shift;
  int nc_hw;                // pixels per image (OH*OW): image of a GEMM row = row / nc_hw
  // CTA pairs (kCtas = 2 instantiation): tiles are taken by clusters of two CTAs, which own two vertically adjacent M tiles
  // of one N tile and execute them as ONE M = 256 tcgen05.mma.cta_group::2 -- each CTA stages its own 128 rows of A and
  // half of the filter tile, so the filter traffic per output row halves
  int num_pair_tiles;       // ceil(num_m_tiles / 2) * num_n_tiles
};

constexpr int kNcSlots = 4;   // images one 128-row tile can touch (7 x 7 maps: 128 / 49 -> up to 4)

template <int BK>
struct ConvSwizzle;
template <>
struct ConvSwizzle<64> { static constexpr uint32_t mode = 2; };  // 128B
template <>
struct ConvSwizzle<32> { static constexpr uint32_t mode = 4; };  // 64B
template <>
struct ConvSwizzle<16> { static constexpr uint32_t mode = 6; };  // 32B

__host__ __device__ inline int conv_b_slot_bytes(int bk, int block_n) { return ((block_n * bk * 2) + 1023) & ~1023; }
__host__ __device__ inline int conv_stage_bytes(int bk, int block_n, int b_resident = 0) {
  const int a_bytes = kConvBlockM * bk * 2;
  return a_bytes + (b_resident ? 0 : conv_b_slot_bytes(bk, block_n));
}

// byte offset of 16-byte chunk `chunk` of row `row` inside a TMA-swizzled staging tile with row pitch
// `pitch` bytes (128 -> SW128, 64 -> SW64, 32 -> SW32): address bits [4,7) ^= bits [7,10) & mask.
__device__ __forceinline__ uint32_t swz_off(int row, int chunk, int pitch, uint32_t mask) {
  uint32_t off = uint32_t(row) * uint32_t(pitch) + uint32_t(chunk) * 16u;
  return off ^ (((off >> 7) & mask) << 4);
}

__device__ __forceinline__ void epi_bar_sync(int group) { asm volatile("bar.sync %0, 128;" ::"r"(group + 1) : "memory"); }

// kNC = true (SE blocks by linearity): scale / shift are per (image, channel) instead of per channel.  A separate
// instantiation so that the code of the default kernels is untouched.
template <int BK, bool kNC = false, int kCtas = 1>
__global__ void __launch_bounds__(kConvThreads, 1)
conv_fprop_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                  const __grid_constant__ CUtensorMap tmOut, const __grid_constant__ CUtensorMap tmRes,
                  const ConvFpropParams p) {
  extern __shared__ uint8_t smem_raw[];
  // swizzled TMA/UMMA tiles want 1024-byte alignment
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));

  constexpr int kABytes = kConvBlockM * BK * 2;
  const int b_rows = p.block_n / kCtas;      // filter rows this CTA stages (a CTA pair splits the N tile)
  const int b_bytes = b_rows * BK * 2;
  const int b_res = p.b_resident;
  const int stage_bytes = conv_stage_bytes(BK, b_rows, b_res);
  const int b_slot = conv_b_slot_bytes(BK, b_rows);
  const uint32_t cta_rank = kCtas == 2 ? cluster_ctarank() : 0u;
  // persistent loop: 1 CTA -> tile = m_tile * num_n_tiles + n_tile; CTA pair -> the cluster takes pair tile t, this CTA the
  // M tile 2 * (t / num_n_tiles) + rank (possibly one past the last: it then computes and stores nothing but zeros, clipped)
  const int tile0 = kCtas == 2 ? int(blockIdx.x >> 1) : int(blockIdx.x);
  const int tile_step = kCtas == 2 ? int(gridDim.x >> 1) : int(gridDim.x);
  auto tile_mn = [&](int t, int& m_tile, int& n_tile) {
    const int q = t / p.num_n_tiles;
    n_tile = t - q * p.num_n_tiles;
    m_tile = kCtas == 2 ? 2 * q + int(cta_rank) : q;
  };
  const int num_stages = p.num_stages;
  const int k_iters = p.R * p.S * p.kc_blocks;

  const int epi_bufs = p.epi_bufs;
  uint8_t* b_res_buf = smem + size_t(num_stages) * stage_bytes;                      // k_iters filter slots (if resident)
  uint8_t* epi_store_buf = b_res_buf + (b_res ? size_t(k_iters) * b_slot : 0);       // epi_bufs x 16 KB per group (if used)
  uint8_t* epi_res_buf = epi_store_buf + (p.use_tma_store ? kEpiGroups * epi_bufs * kEpiStageBytes : 0);
  uint8_t* after = epi_res_buf + (p.use_tma_residual ? kEpiGroups * epi_bufs * kEpiStageBytes : 0);
  float* epi_ss = reinterpret_cast<float*>(after);                                   // [group][scale | shift][256]
  after += kEpiGroups * 2 * 256 * sizeof(float);
  float* epi_nc = reinterpret_cast<float*>(after);                                   // [group][scale | shift][slot][256] (kNC)
  if (kNC) after += kEpiGroups * 2 * kNcSlots * 256 * sizeof(float);
  uint64_t* bars = reinterpret_cast<uint64_t*>(after);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + num_stages;
  uint64_t* tmem_full_bar = bars + 2 * num_stages;
  uint64_t* tmem_empty_bar = bars + 2 * num_stages + 2;
  uint64_t* res_bar = bars + 2 * num_stages + 4;  // [group][buf]
  uint64_t* b_res_bar = bars + 2 * num_stages + 8;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 2 * num_stages + 9);

  const int warp = threadIdx.x >> 5;  // warp-uniform
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    prefetch_tensormap(&tmA);
    prefetch_tensormap(&tmB);
    if (p.use_tma_store) prefetch_tensormap(&tmOut);
    if (p.use_tma_residual) prefetch_tensormap(&tmRes);
    for (int s = 0; s < num_stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tmem_full_bar[a], 1);
      mbar_init(&tmem_empty_bar[a], 4 * kEpiGroups * kCtas);  // one arrive per epilogue warp (of both CTAs of a pair)
    }
    for (int a = 0; a < 4; ++a) mbar_init(&res_bar[a], 1);
    mbar_init(b_res_bar, 1);
    fence_barrier_init();
  }
  if (warp == 2) {
    if (kCtas == 2) { tmem_alloc_pair(tmem_ptr_smem, kConvTmemCols); tmem_relinquish_pair(); }
    else { tmem_alloc(tmem_ptr_smem, kConvTmemCols); tmem_relinquish(); }
  }
  tc_fence_before();
  if (kCtas == 2) cluster_sync_all(); else __syncthreads();   // (pair: the peer's barriers must exist before anything signals them)
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  const int num_tiles = kCtas == 2 ? p.num_pair_tiles : p.num_m_tiles * p.num_n_tiles;

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      const int ohw = p.OH * p.OW;
      if (b_res && tile0 < num_tiles) {  // the whole filter, once (num_n_tiles == 1)
        mbar_arrive_expect_tx(b_res_bar, uint32_t(k_iters) * uint32_t(b_bytes));
        for (int it = 0; it < k_iters; ++it) tma_load_2d(&tmB, b_res_bar, b_res_buf + size_t(it) * b_slot, it * BK, 0);
      }
      for (int tile = tile0; tile < num_tiles; tile += tile_step) {
        int m_tile, n_tile;
        tile_mn(tile, m_tile, n_tile);
        const int m0 = m_tile * kConvBlockM;
        const int n_img = m0 / ohw;
        const int rem = m0 - n_img * ohw;
        const int oh = rem / p.OW;
        const int ow = rem - oh * p.OW;
        const int w_base = ow * p.stride_w - p.pad_l;
        const int h_base = oh * p.stride_h - p.pad_t;
        const int n0 = n_tile * p.block_n + int(cta_rank) * b_rows;
        for (int r = 0; r < p.R; ++r) {
          for (int s = 0; s < p.S; ++s) {
            const int kbase = (r * p.S + s) * p.Cin;
            for (int kc = 0; kc < p.kc_blocks; ++kc) {
              mbar_wait(&empty_bar[stage], phase ^ 1);
              uint8_t* sa = smem + size_t(stage) * stage_bytes;
              uint8_t* sb = sa + kABytes;
              if constexpr (kCtas == 2) {
                // both CTAs' bytes are counted on the leader's barrier; only the leader arrives on it
                if (cta_rank == 0) mbar_arrive_expect_tx(&full_bar[stage], uint32_t(2 * (kABytes + b_bytes)));
                tma_load_im2col_4d_pair(&tmA, &full_bar[stage], sa, kc * BK, w_base, h_base, n_img, uint16_t(s), uint16_t(r));
                tma_load_2d_pair(&tmB, &full_bar[stage], sb, kbase + kc * BK, n0);
              } else {
                mbar_arrive_expect_tx(&full_bar[stage], uint32_t(kABytes + (b_res ? 0 : b_bytes)));
                tma_load_im2col_4d(&tmA, &full_bar[stage], sa, kc * BK, w_base, h_base, n_img, uint16_t(s),
                                   uint16_t(r));
                if (!b_res) tma_load_2d(&tmB, &full_bar[stage], sb, kbase + kc * BK, n0);
              }
              if (++stage == num_stages) { stage = 0; phase ^= 1; }
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer (CTA pair: the leader only)
    if (cta_rank == 0 && elect_one()) {
      const uint32_t idesc = make_idesc_f16(kConvBlockM * kCtas, p.block_n, 0, 0);
      constexpr uint32_t kSbo = 8 * BK * 2;  // 8 rows of BK fp16
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      if (b_res && tile0 < num_tiles) mbar_wait(b_res_bar, 0);
      const uint64_t desc0 = make_smem_desc(smem_u32(smem), 16, kSbo, ConvSwizzle<BK>::mode);
      const uint32_t desc_hi = uint32_t(desc0 >> 32), desc_lo0 = uint32_t(desc0);
      const uint32_t stage_inc = uint32_t(stage_bytes) >> 4, bslot_inc = uint32_t(b_slot) >> 4;
      const uint32_t bres_off = (smem_u32(b_res_buf) - smem_u32(smem)) >> 4;
      for (int tile = tile0; tile < num_tiles; tile += tile_step) {
        mbar_wait(&tmem_empty_bar[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + uint32_t(acc) * 256u;
        for (int it = 0; it < k_iters; ++it) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          // descriptors: the (smem base) descriptor is built once; only the 14-bit address field moves (no carry:
          // smem < 256 KB) -- the single issuing thread has 32..128 cycles per MMA
          const uint32_t a_lo = desc_lo0 + uint32_t(stage) * stage_inc;
          const uint32_t b_lo = b_res ? desc_lo0 + bres_off + uint32_t(it) * bslot_inc : a_lo + (uint32_t(kABytes) >> 4);
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            // advance 16 fp16 = 32 bytes along K inside the swizzle atom: +2 in the (addr>>4) field
            if constexpr (kCtas == 2)
              umma_f16_ss_pair(d_tmem, (uint64_t(desc_hi) << 32) | (a_lo + 2u * k), (uint64_t(desc_hi) << 32) | (b_lo + 2u * k), idesc,
                               (it > 0 || k > 0) ? 1u : 0u);
            else
              umma_f16_ss(d_tmem, (uint64_t(desc_hi) << 32) | (a_lo + 2u * k), (uint64_t(desc_hi) << 32) | (b_lo + 2u * k), idesc,
                          (it > 0 || k > 0) ? 1u : 0u);
          }
          // smem slot free once these MMAs retire (pair: the slot of BOTH CTAs -- the commit arrives on both empty barriers)
          if constexpr (kCtas == 2) umma_commit_pair(&empty_bar[stage]); else umma_commit(&empty_bar[stage]);
          if (++stage == num_stages) { stage = 0; phase ^= 1; }
        }
        if constexpr (kCtas == 2) umma_commit_pair(&tmem_full_bar[acc]); else umma_commit(&tmem_full_bar[acc]);  // accumulator complete -> epilogue
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
      }
    }
  } else {
    // ------------------------------------------------------------ epilogue warps 2..9 (two groups)
    const int quad = warp & 3;            // TMEM lane quadrant this warp may access
    const int group = (warp - 2) >> 2;    // 0: warps 2-5, 1: warps 6-9
    const int tg = int(threadIdx.x) - 64 - group * 128;  // thread index inside the group
    const int row_in_tile = quad * 32 + lane;
    const bool leader = (tg == 0);
    const int cw = p.epi_cw;
    const int chunks_per_tile = p.block_n / cw;
    const int pitch = cw * 2;
    const uint32_t swz_mask = (cw == 64) ? 7u : (cw == 32) ? 3u : 1u;
    const uint32_t chunk_bytes = uint32_t(kConvBlockM) * uint32_t(pitch);
    const int ohw = p.OH * p.OW;
    uint8_t* sbuf0 = epi_store_buf + group * epi_bufs * kEpiStageBytes;
    uint8_t* rbuf0 = epi_res_buf + group * epi_bufs * kEpiStageBytes;
    float* ss_scale = epi_ss + group * 512;   // this group's copy of the N tile's per-channel scale / shift
    float* ss_shift = ss_scale + 256;
    int ss_n_tile = -1;
    // fast path (fp16 output through the TMA store, the case of every layer inside the fused graphs): explicit
    // LDS / STS, no per-element control flow.  A thread's row inside a swizzled staging tile: byte offset
    // row*pitch + chunk16*16 with address bits [4,7) ^= bits [7,10) & mask -- the XOR term depends on the row only.
    const bool fast = p.use_tma_store && !p.out_f32;
    const uint32_t ss_scale_u32 = smem_u32(ss_scale), ss_shift_u32 = smem_u32(ss_shift);
    const uint32_t row_base = uint32_t(row_in_tile) * uint32_t(pitch);
    const uint32_t row_xor = ((row_base >> 7) & swz_mask) << 4;
    const float relu_floor = p.relu ? 0.f : -INFINITY;
    uint64_t* rbar0 = &res_bar[group * 2];
    int acc = 0;
    uint32_t acc_phase = 0;
    uint32_t mine = 0;  // chunks this group has processed (phase of its residual barrier)

    // residual chunk `G` (global chunk index of this CTA: tile_iter * chunks_per_tile + q) -> TMA prefetch into
    // buffer `buf` of this group
    auto prefetch_residual = [&](long long G, int buf) {
      uint64_t* rbar = rbar0 + buf;
      uint8_t* rbuf = rbuf0 + buf * kEpiStageBytes;
      const long long titer = G / chunks_per_tile;
      const int q = int(G - titer * chunks_per_tile);
      const long long tile = (long long)tile0 + titer * tile_step;
      if (tile >= num_tiles) return;
      int nm, nn;
      tile_mn(int(tile), nm, nn);
      mbar_arrive_expect_tx(rbar, chunk_bytes);
      tma_load_2d(&tmRes, rbar, rbuf, nn * p.block_n + q * cw, nm * kConvBlockM);
    };
    if (p.use_tma_residual && leader)
      for (int bq = 0; bq < epi_bufs; ++bq) prefetch_residual(group + 2 * bq, bq);

    long long G0 = 0;  // global chunk index of chunk 0 of the current tile
    for (int tile = tile0; tile < num_tiles; tile += tile_step, G0 += chunks_per_tile) {
      int m_tile, n_tile;
      tile_mn(tile, m_tile, n_tile);
      const int m0 = m_tile * kConvBlockM;
      const int row = m0 + row_in_tile;
      const int n0 = n_tile * p.block_n;
      if (n_tile != ss_n_tile) {
        // per-channel epilogue vectors of this N tile: fetched once per N-tile change (once per kernel when the
        // filter is a single N tile) and before the accumulator wait, so the global-load latency is off the
        // per-chunk critical path.  The previous chunk ended with a group barrier: nobody still reads the old values.
        for (int i = tg; i < p.block_n; i += 128) {
          ss_scale[i] = p.scale ? __ldg(p.scale + n0 + i) : 1.f;
          ss_shift[i] = p.shift ? __ldg(p.shift + n0 + i) : 0.f;
        }
        ss_n_tile = n_tile;
        epi_bar_sync(group);
      }
      int nc_slot = 0;
      if constexpr (kNC) {
        // the images this tile's rows belong to: their [block_n] scale / shift rows go to this group's slots (the previous
        // chunk ended with a group barrier, so nobody still reads the old ones); a thread's slot = its row's image - first
        const int img0 = m0 / p.nc_hw;
        const int last_row = min(m0 + kConvBlockM, p.M) - 1;
        const int nimg = last_row / p.nc_hw - img0 + 1;   // <= kNcSlots (host check: nc_hw >= 43)
        float* ncs = epi_nc + group * (2 * kNcSlots * 256);
        // 16-byte loads, both vectors in flight before the stores: with two images and a 256-wide tile that is ONE
        // round trip to L2 per tile (the scalar loop walked four dependent load -> store pairs per thread: 17 % of the
        // kernel's stall samples, plus the barrier waits of everybody else; profiles/ncu/r02_full_fprop_nc56.txt)
        const int q4 = p.block_n >> 2;
        for (int i = tg; i < nimg * q4; i += 128) {
          const int sl = i / q4, c = (i - sl * q4) * 4;
          const size_t off = size_t(img0 + sl) * p.Kout + n0 + c;
          const float4 vs = __ldg(reinterpret_cast<const float4*>(p.nc_scale + off));
          const float4 vh = __ldg(reinterpret_cast<const float4*>(p.nc_shift + off));
          *reinterpret_cast<float4*>(ncs + sl * 256 + c) = vs;
          *reinterpret_cast<float4*>(ncs + (kNcSlots + sl) * 256 + c) = vh;
        }
        nc_slot = min(row, p.M - 1) / p.nc_hw - img0;
        epi_bar_sync(group);
      }
      mbar_wait(&tmem_full_bar[acc], acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (uint32_t(quad * 32) << 16) + uint32_t(acc) * 256u;
      const bool row_ok = row < p.M;
      size_t row_off;
      if (p.strided_out) {
        const int n_img = row / ohw;
        const int rem = row - n_img * ohw;
        const int oh = rem / p.OW;
        const int ow = rem - oh * p.OW;
        row_off = size_t(n_img) * p.out_sn + size_t(oh) * p.out_sh + size_t(ow) * p.out_sw + size_t(n_tile) * p.out_stile;
      } else {
        row_off = size_t(row) * p.ldc + n0;
      }

      for (int q = int((group + 2 - (G0 & 1)) & 1); q < chunks_per_tile; q += 2, ++mine) {
        const int buf = epi_bufs == 2 ? int(mine & 1u) : 0;
        uint8_t* sbuf = sbuf0 + buf * kEpiStageBytes;
        uint8_t* rbuf = rbuf0 + buf * kEpiStageBytes;
        // the TMA store that last read this staging buffer (epi_bufs chunks ago) must have drained it
        if (p.use_tma_store && leader) {
          if (epi_bufs == 2) tma_store_wait_read<1>(); else tma_store_wait_read<0>();
        }
        uint32_t va[16], vb[16];
        tmem_ld16(taddr + uint32_t(q * cw), va);  // in flight across the waits below
        epi_bar_sync(group);  // staging buffer free
        if (p.use_tma_residual) mbar_wait(rbar0 + buf, (epi_bufs == 2 ? (mine >> 1) : mine) & 1u);

        if (fast) {
          const uint32_t sbuf_u32 = smem_u32(sbuf), rbuf_u32 = smem_u32(rbuf);
          // (kNC) this thread's image slot of the per-(image, channel) vectors
          const uint32_t nc_scale_u32 = kNC ? smem_u32(epi_nc + group * (2 * kNcSlots * 256) + nc_slot * 256) : 0u;
          const uint32_t nc_shift_u32 = kNC ? smem_u32(epi_nc + group * (2 * kNcSlots * 256) + (kNcSlots + nc_slot) * 256) : 0u;
          // the 16 columns' scale / shift vectors: fetched BEFORE the wait on the accumulator load they go with, so that the
          // shared-memory latency hides behind it (ncu: the first FFMA after the eight LDS held ~10 % of the epilogue's
          // stall samples on the HBM-bound layers, two epilogue warps per scheduler having nothing else to issue)
          auto load_ss = [&](int jj, float4 (&sc)[4], float4 (&sh)[4]) {
            const uint32_t jb = uint32_t(q * cw + jj) * 4u;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              sc[i] = lds_f4((kNC ? nc_scale_u32 : ss_scale_u32) + jb + i * 16);
              sh[i] = lds_f4((kNC ? nc_shift_u32 : ss_shift_u32) + jb + i * 16);
            }
          };
          auto process_fast = [&](const uint32_t (&v)[16], int jj, const float4 (&sc)[4], const float4 (&sh)[4], auto res_tag) {
            constexpr bool kRes = decltype(res_tag)::value;
            float x[16];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              x[4 * i] = fmaf(__uint_as_float(v[4 * i]), sc[i].x, sh[i].x);
              x[4 * i + 1] = fmaf(__uint_as_float(v[4 * i + 1]), sc[i].y, sh[i].y);
              x[4 * i + 2] = fmaf(__uint_as_float(v[4 * i + 2]), sc[i].z, sh[i].z);
              x[4 * i + 3] = fmaf(__uint_as_float(v[4 * i + 3]), sc[i].w, sh[i].w);
            }
            const uint32_t c0 = (row_base + uint32_t(jj) * 2u) ^ row_xor, c1 = (row_base + uint32_t(jj) * 2u + 16u) ^ row_xor;
            if constexpr (kRes) {
              const uint4 r0 = lds_u4(rbuf_u32 + c0), r1 = lds_u4(rbuf_u32 + c1);
              const __half2* ra = reinterpret_cast<const __half2*>(&r0);
              const __half2* rb = reinterpret_cast<const __half2*>(&r1);
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const float2 fa = __half22float2(ra[i]), fb = __half22float2(rb[i]);
                x[2 * i] += fa.x; x[2 * i + 1] += fa.y;
                x[8 + 2 * i] += fb.x; x[8 + 2 * i + 1] += fb.y;
              }
            }
            uint4 o[2];
            __half2* o2 = reinterpret_cast<__half2*>(o);
#pragma unroll
            for (int i = 0; i < 8; ++i) o2[i] = __floats2half2_rn(fmaxf(x[2 * i], relu_floor), fmaxf(x[2 * i + 1], relu_floor));
            sts_u4(sbuf_u32 + c0, o[0]);
            sts_u4(sbuf_u32 + c1, o[1]);
          };
          auto run_fast = [&](auto res_tag) {
            float4 sc[4], sh[4];
            for (int jj = 0; jj < cw; jj += 32) {
              load_ss(jj, sc, sh);
              tmem_ld_wait();
              const bool second = jj + 16 < cw;
              if (second) tmem_ld16(taddr + uint32_t(q * cw + jj + 16), vb);
              process_fast(va, jj, sc, sh, res_tag);
              if (second) {
                load_ss(jj + 16, sc, sh);
                tmem_ld_wait();
                if (jj + 32 < cw) tmem_ld16(taddr + uint32_t(q * cw + jj + 32), va);
                process_fast(vb, jj + 16, sc, sh, res_tag);
              }
            }
          };
          if (p.use_tma_residual) run_fast(std::true_type{}); else run_fast(std::false_type{});
          fence_proxy_async_smem();
          epi_bar_sync(group);
          if (leader) {
            tma_store_2d(&tmOut, sbuf, n0 + q * cw, m0);
            tma_store_commit();
            if (p.use_tma_residual) prefetch_residual(G0 + q + 2 * epi_bufs, buf);
          }
          continue;
        }

        // 16 accumulator columns -> scale/shift (+residual, ReLU) -> fp16 staging / direct stores
        auto process = [&](const uint32_t (&v)[16], int jj) {
          const int j = q * cw + jj;  // column inside the tile
          float x[16];
#pragma unroll
          for (int i = 0; i < 16; i += 4) {
            const float4 sc = *reinterpret_cast<const float4*>(ss_scale + j + i);
            const float4 sh = *reinterpret_cast<const float4*>(ss_shift + j + i);
            x[i] = fmaf(__uint_as_float(v[i]), sc.x, sh.x);
            x[i + 1] = fmaf(__uint_as_float(v[i + 1]), sc.y, sh.y);
            x[i + 2] = fmaf(__uint_as_float(v[i + 2]), sc.z, sh.z);
            x[i + 3] = fmaf(__uint_as_float(v[i + 3]), sc.w, sh.w);
          }
          if (p.residual) {
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              uint4 rv;
              if (p.use_tma_residual) {
                rv = *reinterpret_cast<const uint4*>(rbuf + swz_off(row_in_tile, jj / 8 + h, pitch, swz_mask));
              } else if (row_ok) {
                rv = __ldg(reinterpret_cast<const uint4*>(p.residual + row_off + j) + h);
              } else {
                rv = make_uint4(0, 0, 0, 0);
              }
              const __half2* r2 = reinterpret_cast<const __half2*>(&rv);
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const float2 f = __half22float2(r2[i]);
                x[h * 8 + 2 * i] += f.x;
                x[h * 8 + 2 * i + 1] += f.y;
              }
            }
          }
          if (p.relu) {
#pragma unroll
            for (int i = 0; i < 16; ++i) x[i] = fmaxf(x[i], 0.f);
          }
          if (p.out) {
            uint4 o[2];
            __half2* o2 = reinterpret_cast<__half2*>(o);
#pragma unroll
            for (int i = 0; i < 8; ++i) o2[i] = __floats2half2_rn(x[2 * i], x[2 * i + 1]);
            if (p.use_tma_store) {
              *reinterpret_cast<uint4*>(sbuf + swz_off(row_in_tile, jj / 8, pitch, swz_mask)) = o[0];
              *reinterpret_cast<uint4*>(sbuf + swz_off(row_in_tile, jj / 8 + 1, pitch, swz_mask)) = o[1];
            } else if (row_ok) {
              uint4* op = reinterpret_cast<uint4*>(p.out + row_off + j);
              op[0] = o[0];
              op[1] = o[1];
            }
          }
          if (p.out_f32 && row_ok) {
            float4* fp = reinterpret_cast<float4*>(p.out_f32 + row_off + j);
#pragma unroll
            for (int i = 0; i < 4; ++i) fp[i] = make_float4(x[4 * i], x[4 * i + 1], x[4 * i + 2], x[4 * i + 3]);
          }
        };
        // software pipeline over the chunk: the tcgen05.ld of the next 16 columns is in flight while the current
        // 16 are processed (the first load was issued before the barrier / residual waits above)
        for (int jj = 0; jj < cw; jj += 32) {
          tmem_ld_wait();
          const bool second = jj + 16 < cw;
          if (second) tmem_ld16(taddr + uint32_t(q * cw + jj + 16), vb);
          process(va, jj);
          if (second) {
            tmem_ld_wait();
            if (jj + 32 < cw) tmem_ld16(taddr + uint32_t(q * cw + jj + 32), va);
            process(vb, jj + 16);
          }
        }
        if (p.use_tma_store) {
          fence_proxy_async_smem();  // generic-proxy smem writes -> visible to the TMA (async proxy)
          epi_bar_sync(group);       // chunk complete in sbuf; everyone is done with rbuf / scale / shift
          if (leader) {
            tma_store_2d(&tmOut, sbuf, n0 + q * cw, m0);  // rows beyond M are clipped by the tensor map
            tma_store_commit();
            if (p.use_tma_residual) prefetch_residual(G0 + q + 2 * epi_bufs, buf);
          }
        } else {
          epi_bar_sync(group);       // keeps the group in step (the next chunk's barrier count assumes it)
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {   // the accumulator buffer is free: tell the MMA issuer (pair: the one in the leader CTA)
        if (kCtas == 2 && cta_rank != 0) mbar_arrive_remote(&tmem_empty_bar[acc], 0); else mbar_arrive(&tmem_empty_bar[acc]);
      }
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
    if (p.use_tma_store && leader) tma_store_wait<0>();
  }

  tc_fence_before();
  if (kCtas == 2) cluster_sync_all(); else __syncthreads();   // (pair: nobody leaves while the other CTA may still touch it)
  if (warp == 2) {
    tc_fence_after();
    if (kCtas == 2) tmem_dealloc_pair(tmem_base, kConvTmemCols); else tmem_dealloc(tmem_base, kConvTmemCols);
  }
}

}  // namespace xemo
