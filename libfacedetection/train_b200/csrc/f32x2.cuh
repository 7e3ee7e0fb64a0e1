// Packed fp32 FMA (sm_100: FFMA2, `fma.rn.f32x2`): two IEEE round-to-nearest fused multiply-adds
// per instruction -- bit-identical to two fmaf() calls, half the issue slots and fma-pipe cycles
// (plain FFMA issues every other cycle per scheduler on Blackwell).  ptxas folds the {s, s} packing
// of a scalar into a broadcast operand and keeps accumulator pairs in aligned register pairs.
#pragma once
#include <cuda_runtime.h>

namespace yunet {

// (c0, c1) = (a0 * b0 + c0, a1 * b1 + c1)
__device__ __forceinline__ void fma2(float& c0, float& c1, float a0, float a1, float b0, float b1) {
  unsigned long long A, B, C;
  asm("mov.b64 %0, {%1, %2};" : "=l"(A) : "f"(a0), "f"(a1));
  asm("mov.b64 %0, {%1, %2};" : "=l"(B) : "f"(b0), "f"(b1));
  asm("mov.b64 %0, {%1, %2};" : "=l"(C) : "f"(c0), "f"(c1));
  asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(C) : "l"(A), "l"(B));
  asm("mov.b64 {%0, %1}, %2;" : "=f"(c0), "=f"(c1) : "l"(C));
}
// packed add / subtract (FADD2): (a0 + b0, a1 + b1), (a0 - b0, a1 - b1), IEEE round-to-nearest
__device__ __forceinline__ void add2(float& c0, float& c1, float a0, float a1, float b0, float b1) {
  unsigned long long A, B, C;
  asm("mov.b64 %0, {%1, %2};" : "=l"(A) : "f"(a0), "f"(a1));
  asm("mov.b64 %0, {%1, %2};" : "=l"(B) : "f"(b0), "f"(b1));
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(C) : "l"(A), "l"(B));
  asm("mov.b64 {%0, %1}, %2;" : "=f"(c0), "=f"(c1) : "l"(C));
}
__device__ __forceinline__ void sub2(float& c0, float& c1, float a0, float a1, float b0, float b1) {
  unsigned long long A, B, C;
  asm("mov.b64 %0, {%1, %2};" : "=l"(A) : "f"(a0), "f"(a1));
  asm("mov.b64 %0, {%1, %2};" : "=l"(B) : "f"(b0), "f"(b1));
  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(C) : "l"(A), "l"(B));
  asm("mov.b64 {%0, %1}, %2;" : "=f"(c0), "=f"(c1) : "l"(C));
}
// acc = a * b + acc, component-wise
__device__ __forceinline__ void fma4p(float4& acc, const float4& a, const float4& b) {
  fma2(acc.x, acc.y, a.x, a.y, b.x, b.y);
  fma2(acc.z, acc.w, a.z, a.w, b.z, b.w);
}
// acc = s * b + acc with a scalar s
__device__ __forceinline__ void fma4s(float4& acc, float s, const float4& b) {
  fma2(acc.x, acc.y, s, s, b.x, b.y);
  fma2(acc.z, acc.w, s, s, b.z, b.w);
}

}  // namespace yunet
