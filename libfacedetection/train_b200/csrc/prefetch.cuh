// L2 prefetch of the NEXT tile's global inputs by the persistent CUDA-core kernels: one
// cp.async.bulk.prefetch.L2 per contiguous tile row (NHWC rows of a tile are contiguous), issued
// by a handful of threads at the start of the current tile, so that the next tile's loads hit L2
// instead of paying the DRAM latency behind a barrier.
#pragma once
#include <cstdint>

namespace yunet {

__device__ __forceinline__ void l2_prefetch_bulk(const void* p, uint32_t bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}
__device__ __forceinline__ void l2_prefetch_line(const void* p) {
  asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
}

// rows [y_lo, y_hi) x columns [x_lo, x_hi) of an NHWC image with C channels (clamped to the image);
// thread `t` (0-based, any subset of the CTA) takes row y_lo + t
template <int C>
__device__ __forceinline__ void l2_prefetch_tile(const float* img, int H, int W, int y_lo, int y_hi,
                                                 int x_lo, int x_hi, int t) {
  if (y_lo < 0) y_lo = 0;
  if (x_lo < 0) x_lo = 0;
  if (y_hi > H) y_hi = H;
  if (x_hi > W) x_hi = W;
  const int y = y_lo + t;
  if (t >= 0 && y < y_hi && x_hi > x_lo)
    l2_prefetch_bulk(img + ((long long)y * W + x_lo) * C, (uint32_t)((x_hi - x_lo) * C * 4));
}

// 16-byte asynchronous global -> shared copy (LDGSTS); `valid == false` writes 16 zero bytes (the
// source address must still be a mapped one).  Double-buffered staging of the next tile by the
// persistent CUDA-core kernels: commit one group per tile, wait_group<1> before consuming.
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc, bool valid) {
  const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
  const int n = valid ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(gsrc), "r"(n) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

}  // namespace yunet
