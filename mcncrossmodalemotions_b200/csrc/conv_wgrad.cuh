// conv_wgrad.cuh -- filter gradient of vl_nnconv (DF) as an implicit GEMM on tcgen05 (sm_100a).
//
//   dF[kout, r, s, c] = sum_{pixels p=(n,oh,ow)} dY[p, kout] * X[n, oh*sy+r-pt, ow*sx+s-pl, c]
//
// The reduction runs over output pixels, so both operands are "MN-major" for the tensor core: a tile
// of dY is [32 pixels][128 kout] with kout contiguous, a tile of im2col(X) is [32 pixels][block_c
// channels] with channels contiguous (the same TMA im2col loads as the forward pass, one per filter
// tap).  One work item owns a 128-kout slice, up to T (tap, channel-tile) sub-tiles whose fp32
// accumulators fill the 512 TMEM columns (the dY tile is fetched once per stage and reused by all T),
// and one split of the pixel range; partial sums are combined with fp32 red.global.add.
//
// Reference call site: dagnn.Conv.backward under cnn_train_dag, emoVoxCeleb/run_distillation.m:170.
#pragma once
#include "xemo_ptx.cuh"

namespace xemo {

constexpr int kWgPixMax = 128;    // pixels (GEMM-K) per pipeline stage: 32 / 64 / 128, chosen by the plan
constexpr int kWgBlockM = 128;    // kout per item
constexpr int kWgThreads = 192;

struct ConvWgradParams {
  int P;        // N*OH*OW pixels
  int Kout;     // logical output channels (rows of dF)
  int ldy;      // row pitch of dY (elements)
  int Cin, R, S;
  int OH, OW;
  int stride_h, stride_w, pad_t, pad_l;
  int chunk_a;  // contiguous kout elements per smem chunk: 64 / 32 / 16
  int chunk_b;  // contiguous channel elements per smem chunk
  int block_c;  // channels per sub-tile (multiple of 16, <= 256)
  int c_tiles;  // ceil(Cin / block_c)
  int T;        // sub-tiles per item, mt * T * block_c <= 512
  int groups;   // ceil(R*S*c_tiles / T)
  int m_tiles;  // ceil(Kout / 128)
  int mt;       // 128-kout slices per item: 1 or 2
  int m_items;  // ceil(m_tiles / mt)
  int splits;
  int pix_blocks_per_split;
  int num_stages;
  int pix;      // pixels per stage (multiple of 16)
  float* dF;    // [Kout][R][S][Cin] fp32, accumulated into (caller zeroes)
  float scale;  // applied to the accumulators before the atomic add (1/grad_scale)
  int red_vec;  // fp32x4 reductions (large dF) or scalar ones (a small dF that hundreds of splits hit at once:
                // measured 0.88 vs 0.74 ms on the stem, where 296 items reduce into 24 KB)
};

__host__ __device__ inline int wgrad_stage_bytes(int T, int block_c, int pix, int mt = 1) {
  return mt * pix * kWgBlockM * 2 + T * pix * block_c * 2;
}

__global__ void __launch_bounds__(kWgThreads, 1)
conv_wgrad_kernel(const __grid_constant__ CUtensorMap tmY, const __grid_constant__ CUtensorMap tmX,
                  const ConvWgradParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int kWgPix = p.pix;
  const int a_bytes = kWgPix * kWgBlockM * 2;
  const int b_sub_bytes = kWgPix * p.block_c * 2;
  const int stage_bytes = wgrad_stage_bytes(p.T, p.block_c, p.pix, p.mt);
  const int num_stages = p.num_stages;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + size_t(num_stages) * stage_bytes);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + num_stages;
  uint64_t* tmem_full_bar = bars + 2 * num_stages;
  uint64_t* tmem_empty_bar = bars + 2 * num_stages + 1;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 2 * num_stages + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    prefetch_tensormap(&tmY);
    prefetch_tensormap(&tmX);
    for (int s = 0; s < num_stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(tmem_full_bar, 1);
    mbar_init(tmem_empty_bar, 4);
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_ptr_smem, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  const int num_items = p.m_items * p.groups * p.splits;
  // item order: all (kout slice, tap group) items of one pixel range are neighbours, so the CTAs of a wave stream the
  // SAME ranges of dY / x together and DRAM serves each once (with the pixel range fastest, a wave covered every range
  // for half of the tap groups and the next wave fetched everything again: 1.79 GB read for 0.88 GB of operands on conv2)
  const int per_split = p.m_items * p.groups;
  const int total_sub = p.R * p.S * p.c_tiles;
  const int pix_blocks = (p.P + kWgPix - 1) / kWgPix;
  const int ohw = p.OH * p.OW;

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      const int a_chunks = kWgBlockM / p.chunk_a;
      const int a_chunk_bytes = kWgPix * p.chunk_a * 2;
      const int b_chunks = p.block_c / p.chunk_b;
      const int b_chunk_bytes = kWgPix * p.chunk_b * 2;
      for (int item = blockIdx.x; item < num_items; item += gridDim.x) {
        const int split = item / per_split;
        const int grp = (item % per_split) % p.groups;
        const int m0 = ((item % per_split) / p.groups) * p.mt;
        const int nm = min(p.mt, p.m_tiles - m0);
        const int sub0 = grp * p.T;
        const int nsub = min(p.T, total_sub - sub0);
        const int pb0 = split * p.pix_blocks_per_split;
        const int pb1 = min(pb0 + p.pix_blocks_per_split, pix_blocks);
        // pixel coordinates of the block, advanced incrementally (two integer divisions per stage are a visible
        // part of this single thread's per-stage budget)
        const int tap0 = sub0 / p.c_tiles, ct0 = sub0 - tap0 * p.c_tiles;
        const int r0 = tap0 / p.S, s0 = tap0 - r0 * p.S;
        int p0 = pb0 * kWgPix;
        int n_img = p0 / ohw;
        int rem0 = p0 - n_img * ohw;
        int oh = rem0 / p.OW;
        int ow = rem0 - oh * p.OW;
        for (int pb = pb0; pb < pb1; ++pb) {
          const int w_base = ow * p.stride_w - p.pad_l;
          const int h_base = oh * p.stride_h - p.pad_t;
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + size_t(stage) * stage_bytes;
          uint8_t* sb = sa + p.mt * a_bytes;
          mbar_arrive_expect_tx(&full_bar[stage], uint32_t(nm * a_bytes + nsub * b_sub_bytes));
          for (int mi = 0; mi < nm; ++mi)
            for (int ca = 0; ca < a_chunks; ++ca)
              tma_load_2d(&tmY, &full_bar[stage], sa + mi * a_bytes + ca * a_chunk_bytes,
                          (m0 + mi) * kWgBlockM + ca * p.chunk_a, p0);
          int ct = ct0, r = r0, s = s0;   // (channel tile, filter tap) of sub-tile sub0 + t, advanced without divisions
          for (int t = 0; t < nsub; ++t) {
            const int c0 = ct * p.block_c;
            for (int cb = 0; cb < b_chunks; ++cb)
              tma_load_im2col_4d(&tmX, &full_bar[stage], sb + t * b_sub_bytes + cb * b_chunk_bytes,
                                 c0 + cb * p.chunk_b, w_base, h_base, n_img, uint16_t(s), uint16_t(r));
            if (++ct == p.c_tiles) {
              ct = 0;
              if (++s == p.S) { s = 0; ++r; }
            }
          }
          if (++stage == num_stages) { stage = 0; phase ^= 1; }
          p0 += kWgPix;
          ow += kWgPix;
          while (ow >= p.OW) { ow -= p.OW; ++oh; }
          while (oh >= p.OH) { oh -= p.OH; ++n_img; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer
    if (elect_one()) {
      const uint32_t idesc0 = make_idesc_f16(kWgBlockM, 0, 1, 1);  // A and B MN-major; the N field is added per MMA
      const int max_merge = p.block_c >= 256 ? 1 : 256 / p.block_c;
      const uint32_t swz_a = (p.chunk_a == 64) ? 2u : (p.chunk_a == 32) ? 4u : 6u;
      const uint32_t swz_b = (p.chunk_b == 64) ? 2u : (p.chunk_b == 32) ? 4u : 6u;
      const uint32_t sbo_a = 8 * p.chunk_a * 2, lbo_a = kWgPix * p.chunk_a * 2;
      const uint32_t sbo_b = 8 * p.chunk_b * 2, lbo_b = kWgPix * p.chunk_b * 2;
      // One thread issues every MMA of the CTA and an N = 96..128 MMA retires in 48-64 cycles, so the issue loop must
      // stay well under that per MMA: the descriptors of (stage 0, tile 0, k 0) are built once and only their 14-bit
      // address field (bytes >> 4; smem < 256 KB, so no carry out of the field) is advanced with 32-bit adds.
      const uint64_t a_desc0 = make_smem_desc(smem_u32(smem), lbo_a, sbo_a, swz_a);
      const uint64_t b_desc0 = make_smem_desc(smem_u32(smem) + uint32_t(p.mt * a_bytes), lbo_b, sbo_b, swz_b);
      const uint32_t a_hi = uint32_t(a_desc0 >> 32), b_hi = uint32_t(b_desc0 >> 32);
      const uint32_t a_lo0 = uint32_t(a_desc0), b_lo0 = uint32_t(b_desc0);
      const uint32_t stage_inc = uint32_t(stage_bytes) >> 4, a_mi_inc = uint32_t(a_bytes) >> 4, a_k_inc = (2 * sbo_a) >> 4;
      const uint32_t b_t_inc = uint32_t(b_sub_bytes) >> 4, b_k_inc = (2 * sbo_b) >> 4;
      const uint32_t d_mi_inc = uint32_t(p.T * p.block_c), d_t_inc = uint32_t(p.block_c);
      const int k_steps = kWgPix / 16;
      int stage = 0;
      uint32_t phase = 0;
      uint32_t item_phase = 0;
      for (int item = blockIdx.x; item < num_items; item += gridDim.x) {
        const int split = item / per_split;
        const int grp = (item % per_split) % p.groups;
        const int m0 = ((item % per_split) / p.groups) * p.mt;
        const int nm = min(p.mt, p.m_tiles - m0);
        const int sub0 = grp * p.T;
        const int nsub = min(p.T, total_sub - sub0);
        const int pb0 = split * p.pix_blocks_per_split;
        const int pb1 = min(pb0 + p.pix_blocks_per_split, pix_blocks);
        mbar_wait(tmem_empty_bar, item_phase ^ 1);
        tc_fence_after();
        uint32_t accumulate = 0;
        for (int pb = pb0; pb < pb1; ++pb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          uint32_t a_k = a_lo0 + uint32_t(stage) * stage_inc, b_k = b_lo0 + uint32_t(stage) * stage_inc;
          if (nm == 1 && nsub <= max_merge) {
            // one MMA per k step (the stem: N = 64, 32 tensor cycles each): the tightest possible issue loop
            const uint32_t idesc1 = idesc0 + (uint32_t(nsub * p.block_c) >> 3 << 17);
#pragma unroll 4
            for (int k = 0; k < k_steps; ++k) {
              umma_f16_ss(tmem_base, (uint64_t(a_hi) << 32) | a_k, (uint64_t(b_hi) << 32) | b_k, idesc1, accumulate);
              accumulate = 1;
              a_k += a_k_inc;
              b_k += b_k_inc;
            }
            umma_commit(&empty_bar[stage]);
            if (++stage == num_stages) { stage = 0; phase ^= 1; }
            continue;
          }
          for (int k = 0; k < k_steps; ++k) {   // 16 pixels (GEMM-K) = two 8-row groups: advance by 2*SBO
            uint32_t a_d = a_k, d_mi = tmem_base;
            for (int mi = 0; mi < nm; ++mi) {
              const uint64_t a_desc = (uint64_t(a_hi) << 32) | a_d;
              uint32_t b_d = b_k, d = d_mi;
              // consecutive sub-tiles are consecutive chunk sequences in smem (b_sub_bytes = b_chunks * LBO) and
              // consecutive TMEM column ranges: up to 256 / block_c of them go out as ONE MMA with N = tm * block_c
              for (int t = 0; t < nsub;) {
                const int tm = min(nsub - t, max_merge);
                umma_f16_ss(d, a_desc, (uint64_t(b_hi) << 32) | b_d, idesc0 + (uint32_t(tm * p.block_c) >> 3 << 17), accumulate);
                b_d += uint32_t(tm) * b_t_inc;
                d += uint32_t(tm) * d_t_inc;
                t += tm;
              }
              a_d += a_mi_inc;
              d_mi += d_mi_inc;
            }
            accumulate = 1;
            a_k += a_k_inc;
            b_k += b_k_inc;
          }
          umma_commit(&empty_bar[stage]);
          if (++stage == num_stages) { stage = 0; phase ^= 1; }
        }
        umma_commit(tmem_full_bar);
        item_phase ^= 1;
      }
    }
  } else {
    // ------------------------------------------------------------ epilogue warps 2..5
    const int quad = warp & 3;
    const int row_in_tile = quad * 32 + lane;
    uint32_t item_phase = 0;
    for (int item = blockIdx.x; item < num_items; item += gridDim.x) {
      const int grp = (item % per_split) % p.groups;
      const int m0 = ((item % per_split) / p.groups) * p.mt;
      const int nm = min(p.mt, p.m_tiles - m0);
      const int split = item / per_split;
      const int sub0 = grp * p.T;
      const int nsub = min(p.T, total_sub - sub0);
      const int pb0 = split * p.pix_blocks_per_split;
      const bool has_work = pb0 < pix_blocks;
      mbar_wait(tmem_full_bar, item_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (uint32_t(quad * 32) << 16);
      for (int mi = 0; mi < nm; ++mi) {
        const int kout = (m0 + mi) * kWgBlockM + row_in_tile;
        for (int t = 0; t < nsub; ++t) {
          const int sub = sub0 + t;
          const int tap = sub / p.c_tiles;
          const int c0 = (sub - tap * p.c_tiles) * p.block_c;
          float* dst = p.dF + (size_t(kout) * p.R * p.S + tap) * p.Cin + c0;
          for (int j = 0; j < p.block_c; j += 16) {
            uint32_t v[16];
            tmem_ld16(taddr + uint32_t((mi * p.T + t) * p.block_c + j), v);
            tmem_ld_wait();
            // Cin and block_c are multiples of 16: a 16-column group is inside the filter or not at all
            if (has_work && kout < p.Kout && c0 + j < p.Cin) {
              if (p.red_vec) {
#pragma unroll
                for (int i = 0; i < 16; i += 4)
                  red_add_v4(dst + j + i, __uint_as_float(v[i]) * p.scale, __uint_as_float(v[i + 1]) * p.scale,
                             __uint_as_float(v[i + 2]) * p.scale, __uint_as_float(v[i + 3]) * p.scale);
              } else {
#pragma unroll
                for (int i = 0; i < 16; ++i) atomicAdd(dst + j + i, __uint_as_float(v[i]) * p.scale);
              }
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tmem_empty_bar);
      item_phase ^= 1;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace xemo
