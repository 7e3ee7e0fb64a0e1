// Standalone bring-up test of the tcgen05 building blocks in csrc/tc_common.cuh (sm_100a):
//   D[128 x 64] = A[128 x 64] * W[64 x 64]^T  with 3xTF32 error compensation,
//   A loaded by TMA (SWIZZLE_128B) -> registers -> hi/lo split -> TMEM (mode TS) or smem (mode SS),
//   W hi/lo in shared memory (K-major SW128), accumulator in TMEM, read back with tcgen05.ld.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tc_gemm_test tools/tc_gemm_test.cu
// Run under `timeout`; every device-side wait is bounded.
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include "../libfacedetection/train_b200/csrc/tc_common.cuh"

using namespace yunet::tc;

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(2); } } while (0)

struct Smem {
  alignas(1024) float a_raw[2][128 * 32];   // TMA landing zone, 2 k-blocks
  alignas(1024) float a_hi[2][128 * 32];    // SS mode only
  alignas(1024) float a_lo[2][128 * 32];
  alignas(1024) float b_hi[2][64 * 32];
  alignas(1024) float b_lo[2][64 * 32];
  alignas(8) uint64_t tma_bar;
  alignas(8) uint64_t mma_bar;
  uint32_t tmem_base;
};

// mode 0: TS (A in TMEM) 3xTF32; 1: SS (A in smem) 3xTF32; 2: TS single pass (hi*hi only)
__global__ void __launch_bounds__(128) tc_gemm_kernel(const __grid_constant__ CUtensorMap tmapA,
                                                      const float* __restrict__ W, float* out,
                                                      int mode, int* status) {
  extern __shared__ unsigned char smem_bytes[];
  Smem& s = *reinterpret_cast<Smem*>((reinterpret_cast<uintptr_t>(smem_bytes) + 1023) & ~uintptr_t(1023));
  const int tid = threadIdx.x, warp = tid >> 5;
  if (warp == 0) tmem_alloc<256>(&s.tmem_base);
  if (tid == 0) {
    mbar_init(&s.tma_bar, 1);
    mbar_init(&s.mma_bar, 1);
    mbar_fence_init();
    tma_prefetch_desc(&tmapA);
  }
  // W hi / lo -> smem, K-major SW128
  for (int i = tid; i < 64 * 64; i += 128) {
    const int n = i / 64, k = i % 64;
    const float w = W[i];
    const uint32_t off = sw128_offset(64, n, k);
    *reinterpret_cast<uint32_t*>(reinterpret_cast<unsigned char*>(s.b_hi) + off) = tf32_hi(w);
    *reinterpret_cast<uint32_t*>(reinterpret_cast<unsigned char*>(s.b_lo) + off) = tf32_lo(w);
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tbase = s.tmem_base;
  if (tid == 0) {
    mbar_arrive_expect_tx(&s.tma_bar, 2 * 128 * 32 * 4);
    tma_load_2d(s.a_raw[0], &tmapA, &s.tma_bar, 0, 0);
    tma_load_2d(s.a_raw[1], &tmapA, &s.tma_bar, 32, 0);
  }
  bool ok = mbar_wait(&s.tma_bar, 0);
  if (!ok) { if (tid == 0) status[0] = 1; }
  // thread t owns row t: read swizzled, split, and stage
  const uint32_t lane_addr = tbase + ((uint32_t)(warp * 32) << 16);
  const uint32_t D_COL = 0, AHI_COL = 64, ALO_COL = 128;
  if (ok) {
    for (int g = 0; g < 4; ++g) {          // 4 groups of 16 columns
      uint32_t hi[16], lo[16];
      for (int c4 = 0; c4 < 4; ++c4) {
        const int c = g * 4 + c4;          // 16-byte chunk 0..15 of the 256-byte row
        const int kb = c >> 3, cc = c & 7;
        const float4 v = *reinterpret_cast<const float4*>(
            reinterpret_cast<const unsigned char*>(s.a_raw[kb]) + tid * 128 + ((cc ^ (tid & 7)) << 4));
        const float f[4] = {v.x, v.y, v.z, v.w};
        for (int j = 0; j < 4; ++j) { hi[c4 * 4 + j] = tf32_hi(f[j]); lo[c4 * 4 + j] = tf32_lo(f[j]); }
        if (mode == 1) {
          uint4 h4 = make_uint4(hi[c4 * 4], hi[c4 * 4 + 1], hi[c4 * 4 + 2], hi[c4 * 4 + 3]);
          uint4 l4 = make_uint4(lo[c4 * 4], lo[c4 * 4 + 1], lo[c4 * 4 + 2], lo[c4 * 4 + 3]);
          const uint32_t off = tid * 128 + ((cc ^ (tid & 7)) << 4);
          *reinterpret_cast<uint4*>(reinterpret_cast<unsigned char*>(s.a_hi[kb]) + off) = h4;
          *reinterpret_cast<uint4*>(reinterpret_cast<unsigned char*>(s.a_lo[kb]) + off) = l4;
        }
      }
      if (mode != 1) {
        tmem_st16(lane_addr + AHI_COL + g * 16, hi);
        tmem_st16(lane_addr + ALO_COL + g * 16, lo);
      }
    }
    if (mode != 1) tmem_wait_st();
    else fence_proxy_async_smem();
  }
  tc_fence_before();
  __syncthreads();
  if (tid == 0 && ok) {
    tc_fence_after();
    constexpr uint32_t idesc = make_idesc_tf32(128, 64);
    const uint32_t bhi = smem_u32(s.b_hi), blo = smem_u32(s.b_lo);
    const uint32_t ahi = smem_u32(s.a_hi), alo = smem_u32(s.a_lo);
    uint32_t acc = 0;
    for (int pass = 0; pass < 3; ++pass) {
      if (mode == 2 && pass != 2) continue;
      for (int k = 0; k < 8; ++k) {
        const uint32_t koff = (k >> 2) * 64 * 128 + (k & 3) * 32;     // B: [kblock][64 rows][128 B]
        const uint32_t aoff = (k >> 2) * 128 * 128 + (k & 3) * 32;    // A (SS): [kblock][128 rows][128 B]
        const uint64_t bd = make_desc_sw128_kmajor((pass == 1 ? blo : bhi) + koff);
        if (mode == 1) {
          const uint64_t ad = make_desc_sw128_kmajor((pass == 0 ? alo : ahi) + aoff);
          mma_tf32_ss(tbase + D_COL, ad, bd, idesc, acc);
        } else {
          const uint32_t at = tbase + (pass == 0 ? ALO_COL : AHI_COL) + k * 8;
          mma_tf32_ts(tbase + D_COL, at, bd, idesc, acc);
        }
        acc = 1;
      }
    }
    mma_commit(&s.mma_bar);
  }
  bool ok2 = ok && mbar_wait(&s.mma_bar, 0);
  if (ok && !ok2 && tid == 0) status[0] = 2;
  tc_fence_after();
  if (ok2) {
    for (int g = 0; g < 4; ++g) {
      uint32_t v[16];
      tmem_ld16(lane_addr + D_COL + g * 16, v);
      tmem_wait_ld();
      for (int j = 0; j < 16; ++j) out[tid * 64 + g * 16 + j] = __uint_as_float(v[j]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<256>(tbase);
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                             const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                             CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                             CUtensorMapFloatOOBfill);

int main() {
  const int M = 128, K = 64, N = 64;
  std::vector<float> A(M * K), W(N * K), D(M * N);
  srand(1);
  for (auto& v : A) v = (float)rand() / RAND_MAX * 4.f - 1.f;
  for (auto& v : W) v = ((float)rand() / RAND_MAX - 0.5f) * 0.5f;
  float *dA, *dW, *dD; int* dS;
  CK(cudaMalloc(&dA, A.size() * 4)); CK(cudaMalloc(&dW, W.size() * 4)); CK(cudaMalloc(&dD, D.size() * 4));
  CK(cudaMalloc(&dS, 4));
  CK(cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dW, W.data(), W.size() * 4, cudaMemcpyHostToDevice));
  EncodeFn encode = nullptr;
  cudaDriverEntryPointQueryResult qres;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&encode, cudaEnableDefault, &qres));
  if (!encode) { printf("no cuTensorMapEncodeTiled\n"); return 2; }
  CUtensorMap tm;
  cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)M};
  cuuint64_t strides[1] = {(cuuint64_t)K * 4};
  cuuint32_t box[2] = {32, 128};
  cuuint32_t es[2] = {1, 1};
  CUresult r = encode(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, dA, dims, strides, box, es,
                      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                      CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); return 2; }
  const size_t smem = sizeof(Smem) + 1024;
  CK(cudaFuncSetAttribute(tc_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  std::vector<double> ref(M * N);
  for (int m = 0; m < M; ++m) for (int n = 0; n < N; ++n) {
    double s = 0; for (int k = 0; k < K; ++k) s += (double)A[m * K + k] * W[n * K + k];
    ref[m * N + n] = s;
  }
  int rc = 0;
  const char* names[3] = {"TS 3xTF32 (A in TMEM)", "SS 3xTF32 (A in smem)", "TS 1xTF32"};
  for (int mode = 0; mode < 3; ++mode) {
    CK(cudaMemset(dD, 0, D.size() * 4)); CK(cudaMemset(dS, 0, 4));
    tc_gemm_kernel<<<1, 128, smem>>>(tm, dW, dD, mode, dS);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("mode %d (%s): CUDA error %s\n", mode, names[mode], cudaGetErrorString(e)); return 3; }
    int st = 0;
    CK(cudaMemcpy(&st, dS, 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost));
    double maxerr = 0, maxref = 0;
    for (int i = 0; i < M * N; ++i) { maxerr = fmax(maxerr, fabs(D[i] - ref[i])); maxref = fmax(maxref, fabs(ref[i])); }
    printf("mode %d (%s): status %d, max abs err %.3e, rel %.3e  (D[0]=%f ref %f, D[last]=%f ref %f)\n", mode,
           names[mode], st, maxerr, maxerr / maxref, D[0], ref[0], D[M * N - 1], ref[M * N - 1]);
    if (mode < 2 && (st != 0 || maxerr / maxref > 2e-6)) rc = 1;
  }
  printf(rc == 0 ? "TC_GEMM_TEST PASS\n" : "TC_GEMM_TEST FAIL\n");
  return rc;
}
