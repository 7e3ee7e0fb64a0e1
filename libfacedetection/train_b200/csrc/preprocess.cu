// GPU input pipeline: the pixel work of the reference's training transforms on a
// ragged batch of decoded uint8 images, one launch per batch:
//   RandomSquareCrop (mmdet/datasets/pipelines/transforms.py:1126-1146: square patch at (left, top)
//     of side `size`, area outside the image filled with 128) ->
//   Resize(keep_ratio=False) to S x S (transforms.py:258-263 -> mmcv.imresize -> cv2.resize,
//     INTER_LINEAR on the float32 image) ->
//   RandomFlip horizontal (transforms.py:527-530 -> mmcv.imflip) ->
//   Normalize(mean 0, std 1, to_rgb=False) + DefaultFormatBundle: float32 NCHW, BGR, 0..255.
// The random decisions (patch, flip) are drawn on the host exactly like the reference draws them
// (host mirror: pipeline.py); this kernel is the deterministic part.  One thread per output pixel,
// 4 x 3 byte gathers from the source, coalesced fp32 stores per plane.
#include "kernels.h"

namespace yunet {

namespace {

// cv2's INTER_LINEAR source coordinate (imgproc/resize.cpp, the x/y tables; probed against
// OpenCV 4.13 with impulse images): f = (d + 0.5) * scale - 0.5 in double, s = floor(f), the
// fraction f - s cast to float; s < 0 -> (0, 0); s >= n - 1 -> (n - 1, 0)
__device__ __forceinline__ void lin_coord(int d, double scale, int n, int& s0, int& s1, float& f) {
  const double fd = ((double)d + 0.5) * scale - 0.5;
  int s = (int)floor(fd);
  float fx = (float)(fd - (double)s);
  if (s < 0) { s = 0; fx = 0.f; }
  if (s >= n - 1) { s = n - 1; fx = 0.f; }
  s0 = s;
  s1 = (s + 1 < n) ? s + 1 : n - 1;
  f = fx;
}

__global__ void preprocess_u8_kernel(const unsigned char* __restrict__ pixels,
                                     const long long* __restrict__ offsets, const int* __restrict__ hw,
                                     const int* __restrict__ crop, int S, float pad,
                                     float* __restrict__ out) {
  const int b = blockIdx.y;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= S * S) return;
  const int y = i / S, x = i % S;
  const int H = hw[b * 2], W = hw[b * 2 + 1];
  const int left = crop[b * 4], top = crop[b * 4 + 1], size = crop[b * 4 + 2], flip = crop[b * 4 + 3];
  const unsigned char* img = pixels + offsets[b];
  const double scale = 1.0 / ((double)S / (double)size);   // cv2: scale_x = 1. / inv_scale_x
  const int xs = flip ? (S - 1 - x) : x;        // column of the un-flipped resized patch
  int cx0, cx1, cy0, cy1;
  float fx, fy;
  lin_coord(xs, scale, size, cx0, cx1, fx);
  lin_coord(y, scale, size, cy0, cy1, fy);
  const int ix0 = cx0 + left, ix1 = cx1 + left, iy0 = cy0 + top, iy1 = cy1 + top;
  const bool vx0 = ix0 >= 0 && ix0 < W, vx1 = ix1 >= 0 && ix1 < W;
  const bool vy0 = iy0 >= 0 && iy0 < H, vy1 = iy1 >= 0 && iy1 < H;
  const unsigned char* p00 = img + ((long long)iy0 * W + ix0) * 3;
  const unsigned char* p01 = img + ((long long)iy0 * W + ix1) * 3;
  const unsigned char* p10 = img + ((long long)iy1 * W + ix0) * 3;
  const unsigned char* p11 = img + ((long long)iy1 * W + ix1) * 3;
  const float a0 = 1.f - fx, a1 = fx, b0 = 1.f - fy, b1 = fy;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const float v00 = (vy0 && vx0) ? (float)p00[c] : pad;
    const float v01 = (vy0 && vx1) ? (float)p01[c] : pad;
    const float v10 = (vy1 && vx0) ? (float)p10[c] : pad;
    const float v11 = (vy1 && vx1) ? (float)p11[c] : pad;
    // horizontal pass, then vertical pass (cv2's order); plain multiplies and adds
    const float r0 = __fadd_rn(__fmul_rn(v00, a0), __fmul_rn(v01, a1));
    const float r1 = __fadd_rn(__fmul_rn(v10, a0), __fmul_rn(v11, a1));
    out[(((long long)b * 3 + c) * S + y) * S + x] = __fadd_rn(__fmul_rn(r0, b0), __fmul_rn(r1, b1));
  }
}

}  // namespace

cudaError_t launch_preprocess_u8(const unsigned char* pixels, const long long* offsets, const int* hw,
                                 const int* crop, int B, int S, float pad, float* out, cudaStream_t s) {
  dim3 grid((S * S + 255) / 256, B);
  preprocess_u8_kernel<<<grid, 256, 0, s>>>(pixels, offsets, hw, crop, S, pad, out);
  return cudaGetLastError();
}

}  // namespace yunet
