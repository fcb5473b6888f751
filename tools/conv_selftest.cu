// conv_selftest.cu -- standalone GPU bring-up harness for the tcgen05 implicit-GEMM convolution.
// Not part of the product: it checks conv_fprop_kernel against a naive CUDA-core convolution on the
// same fp16 inputs, prints timing, and can dump what an im2col TMA load actually puts in shared memory.
//
//   conv_selftest list            -> number of cases
//   conv_selftest case <i>        -> run one correctness case (exit 0 = pass)
//   conv_selftest perf <i>        -> run one timing case (sampled check)
//   conv_selftest diag            -> im2col TMA smem dump
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "../mcncrossmodalemotions_b200/csrc/conv_launch.cuh"

using namespace xemo;

#define CK(x)                                                                          \
  do {                                                                                 \
    cudaError_t e_ = (x);                                                              \
    if (e_ != cudaSuccess) {                                                           \
      fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
      exit(2);                                                                         \
    }                                                                                  \
  } while (0)

struct Case {
  const char* name;
  ConvGeom g;
  int epi;  // 0 none, 1 scale+shift+relu, 2 scale+shift+residual+relu
  int force_bn;
};

static const Case kCases[] = {
    {"1x1 c64 k64", {2, 16, 16, 64, 64, 1, 1, 1, 1, 0, 0, 0, 0}, 0, 0},
    {"1x1 c256 k128", {2, 16, 16, 256, 128, 1, 1, 1, 1, 0, 0, 0, 0}, 0, 0},
    {"3x3 p1 c64 k64 14x14 n3", {3, 14, 14, 64, 64, 3, 3, 1, 1, 1, 1, 1, 1}, 0, 0},
    {"3x3 p1 c128 k256 7x7 n5", {5, 7, 7, 128, 256, 3, 3, 1, 1, 1, 1, 1, 1}, 1, 0},
    {"1x1 s2 c256 k512 28->14", {2, 28, 28, 256, 512, 1, 1, 2, 2, 0, 0, 0, 0}, 0, 0},
    {"5x5 s2 p1 c96 k256 30x21 (bk32)", {2, 30, 21, 96, 256, 5, 5, 2, 2, 1, 1, 1, 1}, 0, 0},
    {"3x3 p1 c384 k256 30x17", {2, 30, 17, 384, 256, 3, 3, 1, 1, 1, 1, 1, 1}, 2, 0},
    {"9x1 c256 k512 9x8 (fc6-like)", {3, 9, 8, 256, 512, 9, 1, 1, 1, 0, 0, 0, 0}, 0, 0},
    {"7x1 s(2,1) p(3,3,0,0) c32 k64 (conv1 trick)", {2, 40, 20, 32, 64, 7, 1, 2, 1, 3, 3, 0, 0}, 1, 0},
    {"3x3 c16 k32 (bk16)", {2, 12, 12, 16, 32, 3, 3, 1, 1, 1, 1, 1, 1}, 0, 0},
    {"3x3 asym pad [0 1 0 1] s2 c64 k96", {2, 15, 15, 64, 96, 3, 3, 2, 2, 0, 1, 0, 1}, 0, 0},
    {"1x1 c64 k256 bn256", {4, 16, 16, 64, 256, 1, 1, 1, 1, 0, 0, 0, 0}, 2, 256},
    {"3x3 p1 c64 k64 56x56 n4 (multi-tile persistent)", {4, 56, 56, 64, 64, 3, 3, 1, 1, 1, 1, 1, 1}, 1, 0},
    {"1x1 c2048 k512 7x7 n8 (long K)", {8, 7, 7, 2048, 512, 1, 1, 1, 1, 0, 0, 0, 0}, 0, 0},
    {"4x1 c16 k96 (student conv1 s2d)", {2, 33, 20, 16, 96, 4, 1, 1, 1, 0, 0, 0, 0}, 0, 0},
};
static const int kNumCases = sizeof(kCases) / sizeof(kCases[0]);

static const Case kPerf[] = {
    {"res2 3x3 c64 k64 56x56 n256", {256, 56, 56, 64, 64, 3, 3, 1, 1, 1, 1, 1, 1}, 1, 0},
    {"res2 1x1 c64 k256 56x56 n256", {256, 56, 56, 64, 256, 1, 1, 1, 1, 0, 0, 0, 0}, 2, 0},
    {"res2 1x1 c256 k64 56x56 n256", {256, 56, 56, 256, 64, 1, 1, 1, 1, 0, 0, 0, 0}, 1, 0},
    {"res3 3x3 c128 k128 28x28 n256", {256, 28, 28, 128, 128, 3, 3, 1, 1, 1, 1, 1, 1}, 1, 0},
    {"res4 3x3 c256 k256 14x14 n256", {256, 14, 14, 256, 256, 3, 3, 1, 1, 1, 1, 1, 1}, 1, 0},
    {"res4 1x1 c1024 k256 14x14 n256", {256, 14, 14, 1024, 256, 1, 1, 1, 1, 0, 0, 0, 0}, 1, 0},
    {"res5 3x3 c512 k512 7x7 n256", {256, 7, 7, 512, 512, 3, 3, 1, 1, 1, 1, 1, 1}, 1, 0},
    {"res5 1x1 c512 k2048 7x7 n256", {256, 7, 7, 512, 2048, 1, 1, 1, 1, 0, 0, 0, 0}, 2, 0},
    {"vox conv2 5x5 s2 c96 k256 126x73 n128", {128, 126, 73, 96, 256, 5, 5, 2, 2, 1, 1, 1, 1}, 0, 0},
    {"vox conv3 3x3 c256 k384 30x17 n128", {128, 30, 17, 256, 384, 3, 3, 1, 1, 1, 1, 1, 1}, 0, 0},
    {"big gemm 1x1 c4096 k4096 m16384", {16, 32, 32, 4096, 4096, 1, 1, 1, 1, 0, 0, 0, 0}, 0, 0},
    {"big gemm bn128", {16, 32, 32, 4096, 4096, 1, 1, 1, 1, 0, 0, 0, 0}, 0, 128},
};
static const int kNumPerf = sizeof(kPerf) / sizeof(kPerf[0]);

__global__ void naive_conv(const __half* x, const __half* w, const float* scale, const float* shift,
                           const __half* residual, int relu, float* y, ConvGeom g, int OH, int OW, int row_step) {
  const long total_rows = (long(g.N) * OH * OW + row_step - 1) / row_step;
  const long idx = blockIdx.x * long(blockDim.x) + threadIdx.x;
  if (idx >= total_rows * g.Kout) return;
  const int k = int(idx % g.Kout);
  const long m = (idx / g.Kout) * row_step;
  const int ow = int(m % OW);
  const int oh = int((m / OW) % OH);
  const int n = int(m / (long(OW) * OH));
  float acc = 0.f;
  for (int r = 0; r < g.R; ++r) {
    const int h = oh * g.sh + r - g.pt;
    if (h < 0 || h >= g.H) continue;
    for (int s = 0; s < g.S; ++s) {
      const int wv = ow * g.sw + s - g.pl;
      if (wv < 0 || wv >= g.W) continue;
      const __half* xp = x + ((size_t(n) * g.H + h) * g.W + wv) * g.Cin;
      const __half* wp = w + ((size_t(k) * g.R + r) * g.S + s) * g.Cin;
      for (int c = 0; c < g.Cin; ++c) acc += __half2float(xp[c]) * __half2float(wp[c]);
    }
  }
  if (scale) acc *= scale[k];
  if (shift) acc += shift[k];
  if (residual) acc += __half2float(residual[m * g.Kout + k]);
  if (relu) acc = fmaxf(acc, 0.f);
  y[(idx / g.Kout) * g.Kout + k] = acc;
}

static float frand(uint32_t& s) {
  s = s * 1664525u + 1013904223u;
  return float((s >> 8) & 0xFFFF) / 65536.f - 0.5f;
}

static int run_case(const Case& c, bool perf) {
  const ConvGeom& g = c.g;
  const int OH = g.OH(), OW = g.OW();
  const size_t nx = size_t(g.N) * g.H * g.W * g.Cin;
  const size_t nw = size_t(g.Kout) * g.R * g.S * g.Cin;
  const size_t M = size_t(g.N) * OH * OW;
  const size_t ny = M * g.Kout;
  printf("[%s] N=%d HxW=%dx%d Cin=%d Kout=%d RxS=%dx%d stride=%d,%d pad=%d,%d,%d,%d -> OHxOW=%dx%d M=%zu\n", c.name,
         g.N, g.H, g.W, g.Cin, g.Kout, g.R, g.S, g.sh, g.sw, g.pt, g.pb, g.pl, g.pr, OH, OW, M);
  std::vector<__half> hx(nx), hw(nw), hres(c.epi == 2 ? ny : 0);
  std::vector<float> hscale(g.Kout), hshift(g.Kout);
  uint32_t seed = 1234u + uint32_t(g.Cin * 7 + g.Kout);
  for (auto& v : hx) v = __float2half(frand(seed) * 2.f);
  const float wscale = 2.f / sqrtf(float(g.R * g.S * g.Cin));
  for (auto& v : hw) v = __float2half(frand(seed) * wscale * 2.f);
  for (auto& v : hres) v = __float2half(frand(seed));
  for (int k = 0; k < g.Kout; ++k) { hscale[k] = 0.5f + frand(seed) * 0.5f + 0.5f; hshift[k] = frand(seed) * 0.2f; }
  __half *dx, *dw, *dres = nullptr, *dy16;
  float *dscale, *dshift, *dy32, *dref;
  CK(cudaMalloc(&dx, nx * 2));
  CK(cudaMalloc(&dw, nw * 2));
  CK(cudaMalloc(&dy16, ny * 2));
  CK(cudaMalloc(&dy32, ny * 4));
  CK(cudaMalloc(&dscale, g.Kout * 4));
  CK(cudaMalloc(&dshift, g.Kout * 4));
  CK(cudaMemcpy(dx, hx.data(), nx * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dw, hw.data(), nw * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dscale, hscale.data(), g.Kout * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dshift, hshift.data(), g.Kout * 4, cudaMemcpyHostToDevice));
  if (c.epi == 2) {
    CK(cudaMalloc(&dres, ny * 2));
    CK(cudaMemcpy(dres, hres.data(), ny * 2, cudaMemcpyHostToDevice));
  }
  CK(cudaMemset(dy16, 0xFF, ny * 2));
  CK(cudaMemset(dy32, 0xFF, ny * 4));

  int dev = 0, sms = 0;
  CK(cudaGetDevice(&dev));
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));

  ConvEpilogue e;
  if (c.epi >= 1) { e.scale = dscale; e.shift = dshift; e.relu = 1; }
  if (c.epi == 2) e.residual = dres;
  e.out = dy16;
  e.out_f32 = perf ? nullptr : dy32;
  ConvPlan plan;
  if (!conv_fprop_plan(&plan, g, dx, dw, e, sms, c.force_bn)) { printf("  PLAN FAILED\n"); return 3; }
  printf("  plan: bk=%d block_n=%d stages=%d grid=%d smem=%d tiles=%dx%d\n", plan.bk, plan.p.block_n,
         plan.p.num_stages, plan.grid, plan.smem, plan.p.num_m_tiles, plan.p.num_n_tiles);
  CK(conv_fprop_run(plan, 0));
  CK(cudaDeviceSynchronize());

  if (perf) {
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    for (int i = 0; i < 3; ++i) CK(conv_fprop_run(plan, 0));
    const int iters = 10;
    CK(cudaEventRecord(e0));
    for (int i = 0; i < iters; ++i) CK(conv_fprop_run(plan, 0));
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    ms /= iters;
    const double bytes = double(nx) * 2 + double(nw) * 2 + double(ny) * 2 * (c.epi == 2 ? 2 : 1);
    printf("  time %.3f ms  %.1f TFLOP/s  (min-traffic %.1f MB -> %.0f GB/s)\n", ms, plan.flops / ms * 1e-9,
           bytes * 1e-6, bytes / ms * 1e-6);
  }

  // reference (sampled rows for perf cases)
  const int row_step = perf ? 61 : 1;
  const size_t ref_rows = (M + row_step - 1) / row_step;
  CK(cudaMalloc(&dref, ref_rows * g.Kout * 4));
  const long total = long(ref_rows) * g.Kout;
  naive_conv<<<unsigned((total + 255) / 256), 256>>>(dx, dw, e.scale, e.shift, e.residual, e.relu, dref, g, OH, OW,
                                                     row_step);
  CK(cudaDeviceSynchronize());
  std::vector<float> href(ref_rows * g.Kout);
  CK(cudaMemcpy(href.data(), dref, href.size() * 4, cudaMemcpyDeviceToHost));
  std::vector<__half> hy16(ny);
  CK(cudaMemcpy(hy16.data(), dy16, ny * 2, cudaMemcpyDeviceToHost));
  std::vector<float> hy32;
  if (!perf) { hy32.resize(ny); CK(cudaMemcpy(hy32.data(), dy32, ny * 4, cudaMemcpyDeviceToHost)); }

  double max_ref = 0, max_err32 = 0, max_err16 = 0;
  size_t bad = 0, shown = 0;
  for (size_t rr = 0; rr < ref_rows; ++rr) {
    const size_t m = rr * row_step;
    for (int k = 0; k < g.Kout; ++k) {
      const float ref = href[rr * g.Kout + k];
      const float got16 = __half2float(hy16[m * g.Kout + k]);
      const float got32 = perf ? got16 : hy32[m * g.Kout + k];
      max_ref = fmax(max_ref, fabs(ref));
      const double e32 = fabs(double(got32) - ref), e16 = fabs(double(got16) - ref);
      if (!(e32 == e32)) { max_err32 = INFINITY; }
      max_err32 = fmax(max_err32, e32);
      max_err16 = fmax(max_err16, e16);
      const double tol = perf ? 4e-3 : 2e-4;
      if (!(e32 <= tol * (1.0 + fabs(ref)))) {
        ++bad;
        if (shown < 12) {
          const int ow = int(m % OW), oh = int((m / OW) % OH), n = int(m / (size_t(OW) * OH));
          printf("  MISMATCH m=%zu (n=%d oh=%d ow=%d) k=%d got32=%g got16=%g ref=%g\n", m, n, oh, ow, k, got32, got16, ref);
          ++shown;
        }
      }
    }
  }
  printf("  max|ref|=%.4g  max_err(fp32 out)=%.3g  max_err(fp16 out)=%.3g  bad=%zu/%zu  => %s\n", max_ref, max_err32,
         max_err16, bad, ref_rows * g.Kout, bad ? "FAIL" : "PASS");
  return bad ? 1 : 0;
}

// ------------------------------------------------------------------ filter gradient
struct WCase {
  const char* name;
  ConvGeom g;
  int kpad;  // dY row pitch (0 -> Kout)
};
static const WCase kWCases[] = {
    {"wgrad 1x1 c64 k128", {2, 16, 16, 64, 128, 1, 1, 1, 1, 0, 0, 0, 0}, 0},
    {"wgrad 3x3 p1 c64 k64 14x14 n3", {3, 14, 14, 64, 64, 3, 3, 1, 1, 1, 1, 1, 1}, 0},
    {"wgrad 3x3 p1 c256 k384 30x17 n2", {2, 30, 17, 256, 384, 3, 3, 1, 1, 1, 1, 1, 1}, 0},
    {"wgrad 5x5 s2 p1 c96 k256 30x21", {2, 30, 21, 96, 256, 5, 5, 2, 2, 1, 1, 1, 1}, 0},
    {"wgrad 9x1 c256 k512 9x8 (fc6-like)", {3, 9, 8, 256, 512, 9, 1, 1, 1, 0, 0, 0, 0}, 0},
    {"wgrad 4x1 c16 k96 (student conv1 s2d)", {2, 33, 20, 16, 96, 4, 1, 1, 1, 0, 0, 0, 0}, 0},
    {"wgrad 1x1 c1024 k8 pitch16 (fc8)", {37, 1, 1, 1024, 8, 1, 1, 1, 1, 0, 0, 0, 0}, 16},
    {"wgrad 1x1 c4096 k1024 n50 (fc7)", {50, 1, 1, 4096, 1024, 1, 1, 1, 1, 0, 0, 0, 0}, 0},
    {"wgrad 3x3 c384 k256 (block_c 192)", {2, 10, 9, 384, 256, 3, 3, 1, 1, 1, 1, 1, 1}, 0},
    {"wgrad 3x3 c32 k48 pitch48", {2, 10, 9, 32, 48, 3, 3, 1, 1, 1, 1, 1, 1}, 0},
};
static const int kNumWCases = sizeof(kWCases) / sizeof(kWCases[0]);
static const WCase kWPerf[] = {
    {"vox conv1 s2d 4x1 c16 k96 257x148 n128", {128, 257, 148, 16, 96, 4, 1, 1, 1, 0, 0, 0, 0}, 0},
    {"vox conv2 5x5 s2 c96 k256 126x73 n128", {128, 126, 73, 96, 256, 5, 5, 2, 2, 1, 1, 1, 1}, 0},
    {"vox conv3 3x3 c256 k384 30x17 n128", {128, 30, 17, 256, 384, 3, 3, 1, 1, 1, 1, 1, 1}, 0},
    {"vox conv4 3x3 c384 k256 30x17 n128", {128, 30, 17, 384, 256, 3, 3, 1, 1, 1, 1, 1, 1}, 0},
    {"vox fc6 9x1 c256 k4096 9x8 n128", {128, 9, 8, 256, 4096, 9, 1, 1, 1, 0, 0, 0, 0}, 0},
    {"vox fc7 1x1 c4096 k1024 n128", {128, 1, 1, 4096, 1024, 1, 1, 1, 1, 0, 0, 0, 0}, 0},
};
static const int kNumWPerf = sizeof(kWPerf) / sizeof(kWPerf[0]);

// one thread per sampled filter element (k, r, s, c)
__global__ void naive_wgrad(const __half* x, const __half* dy, int ldy, float* ref, ConvGeom g, int OH, int OW,
                            long total, int step) {
  const long t = blockIdx.x * long(blockDim.x) + threadIdx.x;
  if (t * step >= total) return;
  const long i = t * step;
  const int c = int(i % g.Cin);
  const int s = int((i / g.Cin) % g.S);
  const int r = int((i / (long(g.Cin) * g.S)) % g.R);
  const int k = int(i / (long(g.Cin) * g.S * g.R));
  double acc = 0;
  for (int n = 0; n < g.N; ++n)
    for (int oh = 0; oh < OH; ++oh) {
      const int h = oh * g.sh + r - g.pt;
      if (h < 0 || h >= g.H) continue;
      for (int ow = 0; ow < OW; ++ow) {
        const int w = ow * g.sw + s - g.pl;
        if (w < 0 || w >= g.W) continue;
        acc += double(__half2float(x[((size_t(n) * g.H + h) * g.W + w) * g.Cin + c])) *
               double(__half2float(dy[((size_t(n) * OH + oh) * OW + ow) * ldy + k]));
      }
    }
  ref[t] = float(acc);
}

static int run_wcase(const WCase& c, bool perf) {
  const ConvGeom& g = c.g;
  const int OH = g.OH(), OW = g.OW();
  const int ldy = c.kpad ? c.kpad : g.Kout;
  const size_t nx = size_t(g.N) * g.H * g.W * g.Cin;
  const size_t P = size_t(g.N) * OH * OW;
  const size_t ny = P * ldy;
  const size_t nw = size_t(g.Kout) * g.R * g.S * g.Cin;
  printf("[%s] N=%d HxW=%dx%d Cin=%d Kout=%d (pitch %d) RxS=%dx%d stride=%d,%d pad=%d,%d,%d,%d P=%zu\n", c.name, g.N, g.H,
         g.W, g.Cin, g.Kout, ldy, g.R, g.S, g.sh, g.sw, g.pt, g.pb, g.pl, g.pr, P);
  std::vector<__half> hx(nx), hy(ny);
  uint32_t seed = 4321u + uint32_t(g.Cin * 3 + g.Kout);
  for (auto& v : hx) v = __float2half(frand(seed) * 2.f);
  for (size_t i = 0; i < ny; ++i) hy[i] = __float2half((int(i % ldy) < g.Kout) ? frand(seed) * 2.f : 0.f);
  __half *dx, *dy;
  float *ddf, *dref;
  CK(cudaMalloc(&dx, nx * 2));
  CK(cudaMalloc(&dy, ny * 2));
  CK(cudaMalloc(&ddf, nw * 4));
  CK(cudaMemcpy(dx, hx.data(), nx * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dy, hy.data(), ny * 2, cudaMemcpyHostToDevice));
  CK(cudaMemset(ddf, 0, nw * 4));
  int dev = 0, sms = 0;
  CK(cudaGetDevice(&dev));
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  WgradPlan plan;
  if (!conv_wgrad_plan(&plan, g, dx, dy, ldy, ddf, 1.f, sms)) { printf("  PLAN FAILED\n"); return 3; }
  printf("  plan: chunk_a=%d chunk_b=%d block_c=%d c_tiles=%d T=%d groups=%d m_tiles=%d splits=%d (x%d blocks) stages=%d grid=%d smem=%d\n",
         plan.p.chunk_a, plan.p.chunk_b, plan.p.block_c, plan.p.c_tiles, plan.p.T, plan.p.groups, plan.p.m_tiles,
         plan.p.splits, plan.p.pix_blocks_per_split, plan.p.num_stages, plan.grid, plan.smem);
  CK(conv_wgrad_run(plan, 0));
  CK(cudaDeviceSynchronize());
  std::vector<float> hdf(nw);
  CK(cudaMemcpy(hdf.data(), ddf, nw * 4, cudaMemcpyDeviceToHost));
  if (perf) {
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    for (int i = 0; i < 3; ++i) CK(conv_wgrad_run(plan, 0));
    const int iters = 10;
    CK(cudaEventRecord(e0));
    for (int i = 0; i < iters; ++i) CK(conv_wgrad_run(plan, 0));
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    ms /= iters;
    const double bytes = double(nx) * 2 + double(ny) * 2 + double(nw) * 4;
    printf("  time %.3f ms  %.1f TFLOP/s  (min-traffic %.1f MB -> %.0f GB/s)\n", ms, plan.flops / ms * 1e-9, bytes * 1e-6,
           bytes / ms * 1e-6);
  }
  const int step = perf ? 997 : 1;
  const long total = long(nw);
  const long nref = (total + step - 1) / step;
  CK(cudaMalloc(&dref, nref * 4));
  naive_wgrad<<<unsigned((nref + 127) / 128), 128>>>(dx, dy, ldy, dref, g, OH, OW, total, step);
  CK(cudaDeviceSynchronize());
  std::vector<float> href(nref);
  CK(cudaMemcpy(href.data(), dref, nref * 4, cudaMemcpyDeviceToHost));
  double max_ref = 0, max_err = 0;
  for (long t = 0; t < nref; ++t) max_ref = fmax(max_ref, fabs(href[t]));
  size_t bad = 0, shown = 0;
  for (long t = 0; t < nref; ++t) {
    const double e = fabs(double(hdf[t * step]) - href[t]);
    if (!(e == e)) max_err = INFINITY;
    max_err = fmax(max_err, e);
    if (!(e <= 1e-4 * max_ref + 1e-5)) {
      ++bad;
      if (shown < 12) {
        const long i = t * step;
        printf("  MISMATCH k=%ld r=%ld s=%ld c=%ld got=%g ref=%g\n", i / (long(g.Cin) * g.S * g.R),
               (i / (long(g.Cin) * g.S)) % g.R, (i / g.Cin) % g.S, i % g.Cin, hdf[i], href[t]);
        ++shown;
      }
    }
  }
  printf("  max|ref|=%.4g  max_err=%.3g  bad=%zu/%ld  => %s\n", max_ref, max_err, bad, nref, bad ? "FAIL" : "PASS");
  return bad ? 1 : 0;
}

// ------------------------------------------------------------------ im2col diagnostics
__global__ void im2col_dump_kernel(const __grid_constant__ CUtensorMap tm, __half* out, int bytes, int c, int w, int h,
                                   int n, int off_w, int off_h) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    fence_barrier_init();
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    mbar_arrive_expect_tx(&bar, bytes);
    tma_load_im2col_4d(&tm, &bar, smem, c, w, h, n, uint16_t(off_w), uint16_t(off_h));
  }
  mbar_wait(&bar, 0);
  for (int i = threadIdx.x; i < bytes / 2; i += blockDim.x) out[i] = reinterpret_cast<__half*>(smem)[i];
}

static int run_diag() {
  // x[n,h,w,c] encodes one coordinate at a time so the fp16 values stay exact.
  const int N = 3, H = 6, W = 5, C = 64;
  const int R = 3, S = 3, pad = 1, stride = 1;
  const int OH = (H + 2 * pad - R) / stride + 1, OW = (W + 2 * pad - S) / stride + 1;
  const size_t nx = size_t(N) * H * W * C;
  __half* dx;
  __half* dout;
  const int bytes = 128 * 64 * 2;
  CK(cudaMalloc(&dx, nx * 2));
  CK(cudaMalloc(&dout, bytes));
  CUtensorMap tm;
  if (!make_tmap_im2col_nhwc_f16(&tm, dx, N, H, W, C, -pad, -pad, pad - (S - 1), pad - (R - 1), stride, stride, 64, 128,
                                 CU_TENSOR_MAP_SWIZZLE_128B)) {
    printf("diag: tensor map failed\n");
    return 3;
  }
  CK(cudaFuncSetAttribute(im2col_dump_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes + 1024));
  const char* names[4] = {"n", "h", "w", "c"};
  std::vector<std::vector<float>> dumps(4);
  const int start_m = 7;  // first output pixel of the box: (n=0, oh=1, ow=2) with OW=5
  const int oh0 = (start_m / OW) % OH, ow0 = start_m % OW, n0 = start_m / (OW * OH);
  const int off_w = 2, off_h = 1;  // filter tap s=2, r=1
  for (int which = 0; which < 4; ++which) {
    std::vector<__half> hx(nx);
    for (int n = 0; n < N; ++n)
      for (int h = 0; h < H; ++h)
        for (int w = 0; w < W; ++w)
          for (int c = 0; c < C; ++c) {
            const int v[4] = {n + 1, h + 1, w + 1, c + 1};
            hx[((size_t(n) * H + h) * W + w) * C + c] = __float2half(float(v[which]));
          }
    CK(cudaMemcpy(dx, hx.data(), nx * 2, cudaMemcpyHostToDevice));
    CK(cudaMemset(dout, 0xFF, bytes));
    im2col_dump_kernel<<<1, 128, bytes + 1024>>>(tm, dout, bytes, 0, ow0 * stride - pad, oh0 * stride - pad, n0, off_w,
                                                 off_h);
    CK(cudaDeviceSynchronize());
    std::vector<__half> ho(bytes / 2);
    CK(cudaMemcpy(ho.data(), dout, bytes, cudaMemcpyDeviceToHost));
    dumps[which].resize(bytes / 2);
    for (int i = 0; i < bytes / 2; ++i) dumps[which][i] = __half2float(ho[i]);
  }
  // Expected: smem row i (128 B, SW128: 16-byte chunk j stored at chunk j ^ (i & 7)) holds output pixel
  // start_m + i, tap (r=1,s=2): input (n, oh+1-1, ow+2-1), zeros when outside the image or n >= N.
  int bad = 0;
  for (int i = 0; i < 128; ++i) {
    const int m = start_m + i;
    const int ow = m % OW, oh = (m / OW) % OH, n = m / (OW * OH);
    const int h = oh * stride - pad + off_h, w = ow * stride - pad + off_w;
    const bool inside = n < N && h >= 0 && h < H && w >= 0 && w < W;
    for (int ch = 0; ch < 64; ++ch) {
      const int chunk = ch / 8, phys_chunk = chunk ^ (i & 7);
      const int idx = i * 64 + phys_chunk * 8 + (ch % 8);
      const float exp_v[4] = {inside ? float(n + 1) : 0.f, inside ? float(h + 1) : 0.f, inside ? float(w + 1) : 0.f,
                              inside ? float(ch + 1) : 0.f};
      for (int which = 0; which < 4; ++which)
        if (dumps[which][idx] != exp_v[which]) {
          if (bad < 24)
            printf("  diag mismatch row %d ch %d field %s: got %g expected %g (pixel n=%d oh=%d ow=%d -> h=%d w=%d)\n", i,
                   ch, names[which], dumps[which][idx], exp_v[which], n, oh, ow, h, w);
          ++bad;
        }
    }
  }
  printf("diag im2col: %d mismatches => %s\n", bad, bad ? "FAIL" : "PASS");
  if (bad) {
    printf("  raw rows (channel-0 slot and first element of each row): n h w c\n");
    for (int i = 0; i < 40; ++i)
      printf("   row %3d: first elem n=%g h=%g w=%g c=%g\n", i, dumps[0][i * 64], dumps[1][i * 64], dumps[2][i * 64],
             dumps[3][i * 64]);
  }
  return bad ? 1 : 0;
}

int main(int argc, char** argv) {
  if (argc < 2) { printf("usage: conv_selftest list|case i|perf i|diag\n"); return 64; }
  if (!strcmp(argv[1], "list")) { printf("%d %d %d %d\n", kNumCases, kNumPerf, kNumWCases, kNumWPerf); return 0; }
  if (!tma_api().ok) { printf("TMA driver entry points unavailable\n"); return 4; }
  if (!strcmp(argv[1], "diag")) return run_diag();
  if (argc < 3) return 64;
  const int i = atoi(argv[2]);
  if (!strcmp(argv[1], "case")) { if (i < 0 || i >= kNumCases) return 64; return run_case(kCases[i], false); }
  if (!strcmp(argv[1], "perf")) { if (i < 0 || i >= kNumPerf) return 64; return run_case(kPerf[i], true); }
  if (!strcmp(argv[1], "wcase")) { if (i < 0 || i >= kNumWCases) return 64; return run_wcase(kWCases[i], false); }
  if (!strcmp(argv[1], "wperf")) { if (i < 0 || i >= kNumWPerf) return 64; return run_wcase(kWPerf[i], true); }
  return 64;
}
