// Host side of the TMA descriptors: cuTensorMapEncodeTiled through the runtime's driver entry
// point (no -lcuda at link time) and the NHWC activation map every streaming kernel uses.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstring>

namespace yunet {

typedef CUresult (*TmaEncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                CUtensorMapFloatOOBfill);

inline TmaEncodeFn tma_encode_fn() {
  static TmaEncodeFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    cudaDriverEntryPointQueryResult q;
    void* p = nullptr;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess)
      fn = reinterpret_cast<TmaEncodeFn>(p);
  }
  return fn;
}

// NHWC fp32 activation (B, H, W, C) as the 4-D tensor (C, W, H, B); box = (box_c channels,
// box_w columns, box_h rows, 1 image).  Out-of-image elements of a box are filled with zeros.
// `pixel_stride` / `image_stride` in floats (C and H*W*C for a dense tensor).
inline cudaError_t make_nhwc_map(CUtensorMap* tm, const float* base, int C, int W, int H, int B,
                                 long long pixel_stride, long long image_stride, int box_c, int box_w,
                                 int box_h, CUtensorMapSwizzle swz) {
  TmaEncodeFn enc = tma_encode_fn();
  if (!enc) return cudaErrorNotSupported;
  memset(tm, 0, sizeof *tm);
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
  cuuint64_t strides[3] = {(cuuint64_t)pixel_stride * 4, (cuuint64_t)W * pixel_stride * 4,
                           (cuuint64_t)image_stride * 4};
  cuuint32_t box[4] = {(cuuint32_t)box_c, (cuuint32_t)box_w, (cuuint32_t)box_h, 1};
  cuuint32_t es[4] = {1, 1, 1, 1};
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(base), dims, strides, box,
                   es, CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? cudaSuccess : cudaErrorInvalidValue;
}

}  // namespace yunet
