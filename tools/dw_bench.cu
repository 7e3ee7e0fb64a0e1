// Stand-alone micro-benchmark of the depthwise stage of the warp-specialised forward kernel: 10 warps
// per CTA (warp = 4 output columns, lane = channel pair), one CTA per SM, NB blocks of 3 rows each:
// 6 LDS.64 + 36 FFMA2 (+ statistics) + 4 STG.64 per row and thread, no barriers.  Prints cycles per
// block for every combination of {loads, FMAs, stores} so the stage's floor can be read directly.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/dw_bench tools/dw_bench.cu
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>
#include "../libfacedetection/train_b200/csrc/f32x2.cuh"
using namespace yunet;

__device__ __forceinline__ float2 lds64(uint32_t addr) {
  float2 v;
  asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(addr));
  return v;
}

template <int LOADS, int FMAS, int STORES, int STATS>
__global__ void __launch_bounds__(320, 1) dw_kernel(float* out, int nb, long long row_floats, long long* cycles) {
  extern __shared__ __align__(1024) unsigned char smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int q2 = lane, cg = warp;
  for (int i = tid; i < 65536 / 4; i += 320) reinterpret_cast<float*>(smem)[i] = (float)(i & 255) * 1e-3f;
  __syncthreads();
  float2 w2r[9];
#pragma unroll
  for (int k = 0; k < 9; ++k) w2r[k] = make_float2(0.01f * k + 0.001f * q2, 0.02f * k);
  const float2 bias2 = make_float2(0.1f, 0.2f);
  float2 wa[6], wb[6];
#pragma unroll
  for (int d = 0; d < 6; ++d) { wa[d] = make_float2(0.f, 0.f); wb[d] = wa[d]; }
  uint32_t coff[6];
#pragma unroll
  for (int d = 0; d < 6; ++d) {
    const int cc = cg * 4 + d;
    coff[d] = (uint32_t)(cc * 256 + (((q2 >> 1) ^ (cc & 7)) << 4) + (q2 & 1) * 8);
  }
  const uint32_t ybase = (uint32_t)__cvta_generic_to_shared(smem);
  const uint32_t rowstride = 42 * 256;
  double st[4] = {0, 0, 0, 0};
  float* dst0 = out + (long long)blockIdx.x * 3 * row_floats * 0 + cg * 4 * 64 + q2 * 2;
  const long long t0 = clock64();
  for (int j = 0; j < nb; ++j) {
    const uint32_t ys = ybase + (j & 1) * 32768;
    float* dst = dst0 + ((long long)(blockIdx.x * nb + j) % 4096) * 3 * row_floats;
    float s1x = 0.f, s1y = 0.f, s2x = 0.f, s2y = 0.f;
#pragma unroll
    for (int ii = 0; ii < 3; ++ii) {
      float2 nc[6];
      const uint32_t rbase = ys + (uint32_t)ii * rowstride;
#pragma unroll
      for (int d = 0; d < 6; ++d) nc[d] = LOADS ? lds64(rbase + coff[d]) : make_float2(wb[d].y + 1.f, wb[d].x);
      float* drow = dst + (long long)ii * row_floats;
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        float ox = bias2.x, oy = bias2.y;
        if (FMAS) {
#pragma unroll
          for (int d = 0; d < 3; ++d) {
            fma2(ox, oy, w2r[d].x, w2r[d].y, wa[e + d].x, wa[e + d].y);
            fma2(ox, oy, w2r[3 + d].x, w2r[3 + d].y, wb[e + d].x, wb[e + d].y);
            fma2(ox, oy, w2r[6 + d].x, w2r[6 + d].y, nc[e + d].x, nc[e + d].y);
          }
        } else { ox += nc[e].x + nc[e + 2].y; oy += nc[e + 1].y; }
        if (STORES) *reinterpret_cast<float2*>(drow + e * 64) = make_float2(ox, oy);
        if (STATS) { add2(s1x, s1y, s1x, s1y, ox, oy); fma2(s2x, s2y, ox, oy, ox, oy); }
        else { s1x += ox; }
      }
#pragma unroll
      for (int d = 0; d < 6; ++d) { wa[d] = wb[d]; wb[d] = nc[d]; }
    }
    st[0] += s1x; st[1] += s1y; st[2] += s2x; st[3] += s2y;
  }
  const long long t1 = clock64();
  if (tid == 0) cycles[blockIdx.x] = t1 - t0;
  if (st[0] + st[1] + st[2] + st[3] == 12345.678) out[0] = 1.f;
}

template <int L, int F, int S, int T>
void run(const char* name, float* out, long long* cyc, int nb) {
  cudaFuncSetAttribute(dw_kernel<L, F, S, T>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  dw_kernel<L, F, S, T><<<148, 320, 200 * 1024>>>(out, nb, 80 * 64, cyc);
  cudaEventRecord(e0);
  dw_kernel<L, F, S, T><<<148, 320, 200 * 1024>>>(out, nb, 80 * 64, cyc);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  long long h[148];
  cudaMemcpy(h, cyc, sizeof h, cudaMemcpyDeviceToHost);
  double avg = 0;
  for (int i = 0; i < 148; ++i) avg += h[i];
  printf("%-34s %8.1f cycles/block  (%.3f ms, err %s)\n", name, avg / 148 / nb, ms, cudaGetErrorString(cudaGetLastError()));
}

int main() {
  float* out; long long* cyc;
  cudaMalloc(&out, (size_t)4096 * 3 * 80 * 64 * 4 + (1 << 20));
  cudaMalloc(&cyc, 148 * 8);
  const int nb = 200;
  run<1, 1, 1, 1>("loads + fmas + stores + stats", out, cyc, nb);
  run<1, 1, 0, 1>("loads + fmas + stats", out, cyc, nb);
  run<1, 1, 1, 0>("loads + fmas + stores", out, cyc, nb);
  run<0, 1, 0, 0>("fmas only", out, cyc, nb);
  run<1, 0, 0, 0>("loads only", out, cyc, nb);
  run<0, 0, 1, 0>("stores only", out, cyc, nb);
  run<1, 0, 1, 0>("loads + stores", out, cyc, nb);
  run<0, 1, 1, 1>("fmas + stores + stats", out, cyc, nb);
  return 0;
}
