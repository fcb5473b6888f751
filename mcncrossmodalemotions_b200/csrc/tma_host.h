// tma_host.h -- host-side CUtensorMap construction (tiled + im2col) without linking libcuda:
// the driver entry points are resolved through the runtime (cudaGetDriverEntryPoint).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

namespace xemo {

typedef CUresult (*PFN_tmEncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                      const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                      CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                      CUtensorMapFloatOOBfill);
typedef CUresult (*PFN_tmEncodeIm2col)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                       const cuuint64_t*, const int*, const int*, cuuint32_t, cuuint32_t,
                                       const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                       CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

struct TmaApi {
  PFN_tmEncodeTiled tiled = nullptr;
  PFN_tmEncodeIm2col im2col = nullptr;
  int driver_version = 0;
  bool ok = false;
};

inline const TmaApi& tma_api() {
  static TmaApi api = [] {
    TmaApi a;
    cudaDriverEntryPointQueryResult q1, q2;
    void* f1 = nullptr;
    void* f2 = nullptr;
    cudaError_t e1 = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f1, cudaEnableDefault, &q1);
    cudaError_t e2 = cudaGetDriverEntryPoint("cuTensorMapEncodeIm2col", &f2, cudaEnableDefault, &q2);
    cudaDriverGetVersion(&a.driver_version);
    a.tiled = reinterpret_cast<PFN_tmEncodeTiled>(f1);
    a.im2col = reinterpret_cast<PFN_tmEncodeIm2col>(f2);
    a.ok = (e1 == cudaSuccess && e2 == cudaSuccess && f1 && f2 && q1 == cudaDriverEntryPointSuccess &&
            q2 == cudaDriverEntryPointSuccess);
    return a;
  }();
  return api;
}

inline CUtensorMapSwizzle swizzle_for_bytes(int inner_bytes) {
  switch (inner_bytes) {
    case 128: return CU_TENSOR_MAP_SWIZZLE_128B;
    case 64: return CU_TENSOR_MAP_SWIZZLE_64B;
    case 32: return CU_TENSOR_MAP_SWIZZLE_32B;
    default: return CU_TENSOR_MAP_SWIZZLE_NONE;
  }
}

// 2-D fp16 matrix [rows][cols] (cols contiguous, row pitch ld elements); box = box_cols x box_rows.
inline bool make_tmap_2d_f16(CUtensorMap* tm, const void* base, uint64_t rows, uint64_t cols, uint64_t ld,
                             uint32_t box_cols, uint32_t box_rows, CUtensorMapSwizzle swz) {
  const TmaApi& api = tma_api();
  if (!api.ok) return false;
  cuuint64_t gdim[2] = {cols, rows};
  cuuint64_t gstride[1] = {ld * 2};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = api.tiled(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(base), gdim, gstride, box,
                         estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) fprintf(stderr, "[xemo] cuTensorMapEncodeTiled failed: %d\n", int(r));
  return r == CUDA_SUCCESS;
}

// im2col-mode map over an NHWC fp16 activation tensor.  lower/upper are the bounding-box corners
// in (W,H) order: lower = -pad_lower, upper = pad_upper - (filter-1)*dilation (fprop convention).
inline bool make_tmap_im2col_nhwc_f16(CUtensorMap* tm, const void* base, int N, int H, int W, int C,
                                      int lower_w, int lower_h, int upper_w, int upper_h, int stride_w,
                                      int stride_h, uint32_t channels_per_pixel, uint32_t pixels_per_column,
                                      CUtensorMapSwizzle swz) {
  const TmaApi& api = tma_api();
  if (!api.ok) return false;
  cuuint64_t gdim[4] = {cuuint64_t(C), cuuint64_t(W), cuuint64_t(H), cuuint64_t(N)};
  cuuint64_t gstride[3] = {cuuint64_t(C) * 2, cuuint64_t(W) * C * 2, cuuint64_t(H) * W * C * 2};
  int lower[2] = {lower_w, lower_h};
  int upper[2] = {upper_w, upper_h};
  cuuint32_t estr[4] = {1, cuuint32_t(stride_w), cuuint32_t(stride_h), 1};
  CUresult r = api.im2col(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void*>(base), gdim, gstride, lower,
                          upper, channels_per_pixel, pixels_per_column, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swz,
                          CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    fprintf(stderr, "[xemo] cuTensorMapEncodeIm2col failed: %d\n", int(r));
    return false;
  }
  // Drivers up to CUDA 13.1 set a descriptor bit that breaks im2col loads from tensors smaller than
  // 128 KiB; clear it (same workaround CUTLASS applies in make_im2col_tma_copy_desc).
  if (api.driver_version <= 13010) {
    const uint64_t bytes = uint64_t(N) * H * W * C * 2;
    if (bytes < 131072) reinterpret_cast<uint64_t*>(tm)[1] &= ~(1ull << 21);
  }
  return true;
}

}  // namespace xemo
