// xemo_internal.h -- context object and helpers shared by the translation units of libxemo.so.
#pragma once
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>

#include <string>
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

inline int pad_to(int v, int m) { return (v + m - 1) / m * m; }

}  // namespace xemo
