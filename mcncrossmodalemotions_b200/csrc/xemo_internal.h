// xemo_internal.h -- context object and helpers shared by the translation units of libxemo.so.
#pragma once
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>

#include <mutex>
#include <string>
#include <utility>
#include <vector>

#include "../../include/xemo.h"

struct xemo_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;   // current launch stream (xemo_set_stream)
  cudaStream_t primary = nullptr;  // the stream the context was created on: capture, graph launches, sync
  bool own_stream = false;
  int num_sms = 0;
  std::string err;
  uint64_t launches = 0;        // kernels launched (eager) + kernel nodes replayed
  bool capturing = false;
  uint64_t capture_mark = 0;    // value of `launches` when the capture began
  int deterministic = 0;        // xemo_set_deterministic: no order-dependent floating-point reductions in the training step
  int conv_precision = 0;       // xemo_set_conv_precision: 0 = fp16 operands, 1 = split fp16 x 3 (fp32-equivalent)
  cudaMemPool_t pool = nullptr; // scratch of the boundary operators (xemo_vl.cu): stream-ordered, kept until xemo_destroy
};

struct xemo_graph {
  cudaGraph_t graph = nullptr;
  cudaGraphExec_t exec = nullptr;
  int num_kernels = 0;
};

namespace xemo {

inline int fail(xemo_ctx* ctx, int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  if (ctx) ctx->err = buf;
  return code;
}

#define XEMO_CUDA(ctx, call)                                                                              \
  do {                                                                                                    \
    cudaError_t e_ = (call);                                                                              \
    if (e_ != cudaSuccess)                                                                                \
      return ::xemo::fail(ctx, XEMO_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
  } while (0)

// after a kernel launch: count it and surface launch-configuration errors
#define XEMO_LAUNCHED(ctx, n)                                                                             \
  do {                                                                                                    \
    cudaError_t e_ = cudaGetLastError();                                                                  \
    if (e_ != cudaSuccess)                                                                                \
      return ::xemo::fail(ctx, XEMO_ERR_CUDA, "kernel launch failed: %s (%s:%d)", cudaGetErrorString(e_), __FILE__, __LINE__); \
    (ctx)->launches += (n);                                                                               \
  } while (0)

#define XEMO_REQUIRE(ctx, cond, ...)                                          \
  do {                                                                        \
    if (!(cond)) return ::xemo::fail(ctx, XEMO_ERR_INVALID, __VA_ARGS__);     \
  } while (0)

inline int grid_for(size_t work_items, int threads, int num_sms, int per_sm = 8) {
  size_t blocks = (work_items + threads - 1) / threads;
  const size_t cap = size_t(num_sms) * per_sm;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return int(blocks);
}

// Blocks of `kernel` (at `threads` threads, `smem` dynamic bytes) that one SM holds at once, cached per kernel.  The
// grid-stride kernels size their grids to WHOLE waves of this number: a cap of 8 blocks per SM for a kernel of which 3 or
// 6 fit ran 2.67 / 1.33 waves whose last one left most SMs idle (ncu: 34 % / 52 % of the warp slots active on
// bn_bwd_apply / maxpool_bwd_3x3s2).
template <typename K>
inline int resident_blocks(K kernel, int threads, size_t smem = 0) {
  static std::mutex mu;
  static std::vector<std::pair<const void*, int>> cache;
  const void* key = reinterpret_cast<const void*>(kernel);
  std::lock_guard<std::mutex> lock(mu);
  for (const auto& e : cache)
    if (e.first == key) return e.second;
  int n = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kernel, threads, smem) != cudaSuccess || n < 1) {
    cudaGetLastError();
    n = 1;
  }
  cache.emplace_back(key, n);
  return n;
}

inline int pad_to(int v, int m) { return (v + m - 1) / m * m; }

}  // namespace xemo
