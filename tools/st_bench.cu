// Store-path micro-benchmark for the depthwise stage: 10 warps per CTA, one CTA per SM, every block
// iteration writes 3 rows x 40 pixels x 256 B = 30 KB.  Variants: STG.64 / STG.128 / 256-bit stores
// straight from registers, and shared-memory staging + one cp.async.bulk (10 KB) per row.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/st_bench tools/st_bench.cu
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// MODE 0: STG.64 (lane = channel pair, warp = 4 columns); 1: STG.128 (lane = quad, 2 px per instr);
// 2: st.global.v8 (lane = 8 channels, 4 px per instr); 3: STS.64 staging + bulk store per row;
// 4: like 0 with st.global.cs; 5: like 3 but STS.128
template <int MODE>
__global__ void __launch_bounds__(320, 1) st_kernel(float* out, int nb, long long row_floats, long long* cycles) {
  extern __shared__ __align__(1024) unsigned char smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  float vx = tid * 1e-3f, vy = lane * 2e-3f;
  const long long t0 = clock64();
  int ring = 0;
  for (int j = 0; j < nb; ++j) {
    float* dst = out + ((long long)(blockIdx.x * nb + j) % 4096) * 3 * row_floats;
#pragma unroll
    for (int ii = 0; ii < 3; ++ii) {
      float* drow = dst + (long long)ii * row_floats;
      vx += 1.f; vy += 2.f;
      if (MODE == 0 || MODE == 4) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          float* p = drow + (warp * 4 + e) * 64 + lane * 2;
          if (MODE == 0) *reinterpret_cast<float2*>(p) = make_float2(vx + e, vy);
          else asm volatile("st.global.cs.v2.f32 [%0], {%1, %2};" ::"l"(p), "f"(vx + e), "f"(vy) : "memory");
        }
      } else if (MODE == 1) {
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          float* p = drow + (warp * 4 + e * 2 + (lane >> 4)) * 64 + (lane & 15) * 4;
          *reinterpret_cast<float4*>(p) = make_float4(vx + e, vy, vx, vy);
        }
      } else if (MODE == 2) {
        float* p = drow + (warp * 4 + (lane >> 3)) * 64 + (lane & 7) * 8;
        asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "f"(vx), "f"(vy), "f"(vx),
                     "f"(vy), "f"(vx), "f"(vy), "f"(vx), "f"(vy)
                     : "memory");
      } else {
        unsigned char* buf = smem + ring * 10240;
        // the bulk store that last read this buffer must be done (at most 3 groups stay in flight)
        if (tid == 0) asm volatile("cp.async.bulk.wait_group.read 3;" ::: "memory");
        asm volatile("bar.sync 1, 320;" ::: "memory");
        if (MODE == 3) {
#pragma unroll
          for (int e = 0; e < 4; ++e)
            *reinterpret_cast<float2*>(buf + (warp * 4 + e) * 256 + lane * 8) = make_float2(vx + e, vy);
        } else {
#pragma unroll
          for (int e = 0; e < 2; ++e)
            *reinterpret_cast<float4*>(buf + (warp * 4 + e * 2 + (lane >> 4)) * 256 + (lane & 15) * 16) =
                make_float4(vx + e, vy, vx, vy);
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("bar.sync 1, 320;" ::: "memory");
        if (tid == 0) {
          asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(drow), "r"(smem_u32(buf)),
                       "r"(10240)
                       : "memory");
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
        ring = (ring + 1) & 3;
      }
    }
  }
  if (MODE == 3 || MODE == 5) { if (tid == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
  const long long t1 = clock64();
  if (tid == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int MODE>
void run(const char* name, float* out, long long* cyc, int nb) {
  cudaFuncSetAttribute(st_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  st_kernel<MODE><<<148, 320, 200 * 1024>>>(out, nb, 80 * 64, cyc);
  cudaEventRecord(e0);
  st_kernel<MODE><<<148, 320, 200 * 1024>>>(out, nb, 80 * 64, cyc);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  long long h[148];
  cudaMemcpy(h, cyc, sizeof h, cudaMemcpyDeviceToHost);
  double avg = 0;
  for (int i = 0; i < 148; ++i) avg += h[i];
  printf("%-44s %8.1f cycles/block  %.3f ms  %.2f TB/s  (%s)\n", name, avg / 148 / nb, ms,
         148.0 * nb * 30720 / (ms * 1e-3) / 1e12, cudaGetErrorString(cudaGetLastError()));
}

int main() {
  float* out; long long* cyc;
  cudaMalloc(&out, (size_t)4096 * 3 * 80 * 64 * 4 + (1 << 20));
  cudaMalloc(&cyc, 148 * 8);
  const int nb = 400;
  run<0>("STG.64   (256 B / warp instr)", out, cyc, nb);
  run<4>("STG.64.cs", out, cyc, nb);
  run<1>("STG.128  (512 B / warp instr)", out, cyc, nb);
  run<2>("STG.256  (1 KB / warp instr)", out, cyc, nb);
  run<3>("STS.64 + cp.async.bulk 10 KB per row", out, cyc, nb);
  run<5>("STS.128 + cp.async.bulk 10 KB per row", out, cyc, nb);
  return 0;
}
