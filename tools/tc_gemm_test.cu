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
  alignas(1024) float a_raw32[2][128 * 32];  // TMA landing zone, SWIZZLE_128B_ATOM_32B
  alignas(1024) float b_hi[2][64 * 32];
  alignas(1024) float b_lo[2][64 * 32];
  alignas(8) uint64_t tma_bar;
  alignas(8) uint64_t mma_bar;
  uint32_t tmem_base;
};

// mode 0: TS (A in TMEM) 3xTF32; 1: SS (A in smem) 3xTF32; 2: TS single pass (hi*hi only)
// mode 3 / 4 (not in the run list): MN-major operands in the plain SWIZZLE_128B layout -- the
//         hardware returns zeros: 32-bit MN-major operands exist only in SWIZZLE_128B_BASE32B
// mode 5: TS, D = A * W (NOT transposed): B = W[k][n] as an MN-major SWIZZLE_128B_BASE32B operand
//         (variants probe the LBO / SBO semantics: blocks of 32 n `lbo` apart, 4-row groups 512 B)
// mode 6: TS, D[128 x 64] = [X_hi ; X_lo]^T (X_hi + X_lo): A = X^T staged in TMEM (lane = column of
//         X, hi on lanes 0..63, lo on 64..127, TMEM column = row of X), B = X as MN-major BASE32B;
//         rows m and m+64 add up to (X^T X)[m][n].  Also checks the TMA SWIZZLE_128B_ATOM_32B
//         landing layout (32-byte chunk index XOR (row & 3)).
__global__ void __launch_bounds__(128) tc_gemm_kernel(const __grid_constant__ CUtensorMap tmapA, const __grid_constant__ CUtensorMap tmapA32,
                                                      const float* __restrict__ W, float* out,
                                                      int mode, int* status, int variant) {
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
  if (mode == 5) {
    // W[k = co][n = ci] as an MN-major operand, SWIZZLE_128B_BASE32B: 2 n-blocks of [64 rows][128 B]
    for (int i = tid; i < 64 * 64; i += 128) {
      const int k = i / 64, n = i % 64;
      const float w = W[i];
      const uint32_t off = (n >> 5) * 8192 + k * 128 + ((((n & 31) >> 3) ^ (k & 3)) << 5) + (n & 7) * 4;
      *reinterpret_cast<uint32_t*>(reinterpret_cast<unsigned char*>(s.b_hi) + off) = tf32_hi(w);
      *reinterpret_cast<uint32_t*>(reinterpret_cast<unsigned char*>(s.b_lo) + off) = tf32_lo(w);
    }
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tbase = s.tmem_base;
  if (tid == 0) {
    mbar_arrive_expect_tx(&s.tma_bar, 4 * 128 * 32 * 4);
    tma_load_2d(s.a_raw[0], &tmapA, &s.tma_bar, 0, 0);
    tma_load_2d(s.a_raw[1], &tmapA, &s.tma_bar, 32, 0);
    tma_load_2d(s.a_raw32[0], &tmapA32, &s.tma_bar, 0, 0);
    tma_load_2d(s.a_raw32[1], &tmapA32, &s.tma_bar, 32, 0);
  }
  bool ok = mbar_wait(&s.tma_bar, 0);
  if (!ok) { if (tid == 0) status[0] = 1; }
  if (ok && mode == 6) {
    // which row bits does the TMA 32B-atom swizzle use?  H1: chunk32 ^= row & 3, H2: chunk32 ^= (row >> 1) & 3
    int bad1 = 0, bad2 = 0;
    for (int c = 0; c < 64; ++c) {
      const float v = *(reinterpret_cast<const float*>(reinterpret_cast<const unsigned char*>(s.a_raw[c >> 5]) +
                        tid * 128 + ((((c & 31) >> 2) ^ (tid & 7)) << 4)) + (c & 3));
      const unsigned char* b32 = reinterpret_cast<const unsigned char*>(s.a_raw32[c >> 5]) + tid * 128;
      const float h1 = *reinterpret_cast<const float*>(b32 + ((((c & 31) >> 3) ^ (tid & 3)) << 5) + (c & 7) * 4);
      const float h2 = *reinterpret_cast<const float*>(b32 + ((((c & 31) >> 3) ^ ((tid >> 1) & 3)) << 5) + (c & 7) * 4);
      bad1 += (h1 != v); bad2 += (h2 != v);
    }
    atomicAdd(&status[1], bad1); atomicAdd(&status[2], bad2);
  }
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
        if (mode == 6) {
          // B of the weight-gradient GEMM: [pixel][32 ch] rows, SWIZZLE_128B_BASE32B (chunk32 ^= row & 3)
          uint4 h4 = make_uint4(hi[c4 * 4], hi[c4 * 4 + 1], hi[c4 * 4 + 2], hi[c4 * 4 + 3]);
          uint4 l4 = make_uint4(lo[c4 * 4], lo[c4 * 4 + 1], lo[c4 * 4 + 2], lo[c4 * 4 + 3]);
          const uint32_t off = tid * 128 + ((((cc >> 1) ^ (tid & 3)) << 5) | ((cc & 1) << 4));
          *reinterpret_cast<uint4*>(reinterpret_cast<unsigned char*>(s.a_hi[kb]) + off) = h4;
          *reinterpret_cast<uint4*>(reinterpret_cast<unsigned char*>(s.a_lo[kb]) + off) = l4;
        } else if (mode == 1 || mode >= 3) {
          uint4 h4 = make_uint4(hi[c4 * 4], hi[c4 * 4 + 1], hi[c4 * 4 + 2], hi[c4 * 4 + 3]);
          uint4 l4 = make_uint4(lo[c4 * 4], lo[c4 * 4 + 1], lo[c4 * 4 + 2], lo[c4 * 4 + 3]);
          const uint32_t off = tid * 128 + ((cc ^ (tid & 7)) << 4);
          *reinterpret_cast<uint4*>(reinterpret_cast<unsigned char*>(s.a_hi[kb]) + off) = h4;
          *reinterpret_cast<uint4*>(reinterpret_cast<unsigned char*>(s.a_lo[kb]) + off) = l4;
        }
      }
      if ((mode != 1 && mode < 3) || mode == 5) {
        tmem_st16(lane_addr + AHI_COL + g * 16, hi);
        tmem_st16(lane_addr + ALO_COL + g * 16, lo);
      }
    }
    if (mode == 6) {
      // A of the weight-gradient GEMM in TMEM, transposed: lane m < 64 = hi(X[p][m]), lane 64+m = lo(X[p][m]),
      // column = pixel p (128 columns at AHI_COL)
      const int m = tid & 63, islo = tid >> 6;
      for (int g = 0; g < 8; ++g) {
        uint32_t v[16];
        for (int j = 0; j < 16; ++j) {
          const int p = g * 16 + j;
          const float x = *(reinterpret_cast<const float*>(reinterpret_cast<const unsigned char*>(s.a_raw[m >> 5]) +
                            p * 128 + ((((m & 31) >> 2) ^ (p & 7)) << 4)) + (m & 3));
          v[j] = islo ? tf32_lo(x) : tf32_hi(x);
        }
        tmem_st16(lane_addr + AHI_COL + g * 16, v);
      }
    }
    if ((mode != 1 && mode < 3) || mode >= 5) tmem_wait_st();
    fence_proxy_async_smem();
  }
  tc_fence_before();
  __syncthreads();
  if (tid == 0 && ok) {
    tc_fence_after();
    constexpr uint32_t idesc = make_idesc_tf32(128, 64);
    const uint32_t bhi = smem_u32(s.b_hi), blo = smem_u32(s.b_lo);
    const uint32_t ahi = smem_u32(s.a_hi), alo = smem_u32(s.a_lo);
    uint32_t acc = 0;
    if (mode == 3) {
      constexpr uint32_t idesc3 = make_idesc_tf32(128, 64, 0, 1);
      for (int pass = 0; pass < 3; ++pass)
        for (int k = 0; k < 8; ++k) {       // K = co: 8 rows of the W buffer per step
          const uint32_t aoff = (k >> 2) * 128 * 128 + (k & 3) * 32;
          const uint64_t ad = make_desc_sw128_kmajor((pass == 0 ? alo : ahi) + aoff);
          const uint64_t bd = variant ? make_desc_sw128_mnmajor((pass == 1 ? blo : bhi) + k * 1024, 1024, 64 * 128) : make_desc_sw128_mnmajor((pass == 1 ? blo : bhi) + k * 1024, 64 * 128);
          mma_tf32_ss(tbase + D_COL, ad, bd, idesc3, acc);
          acc = 1;
        }
    } else if (mode == 5) {
      constexpr uint32_t idesc5 = make_idesc_tf32(128, 64, 0, 1);
      const uint32_t lbo = variant == 1 ? 512 : 8192, sbo = variant == 1 ? 8192 : (variant == 2 ? 1024 : 512);
      for (int pass = 0; pass < 3; ++pass)
        for (int k = 0; k < 8; ++k) {       // K = co: 8 rows (1024 B) of the W buffer per step
          const uint32_t at = tbase + (pass == 0 ? ALO_COL : AHI_COL) + k * 8;
          const uint64_t bd = make_desc_sw128_mnmajor((pass == 1 ? blo : bhi) + k * 1024, lbo, sbo, 1);
          mma_tf32_ts(tbase + D_COL, at, bd, idesc5, acc);
          acc = 1;
        }
    } else if (mode == 6) {
      constexpr uint32_t idesc6 = make_idesc_tf32(128, 64, 0, 1);
      const uint32_t lbo = variant == 1 ? 512 : 16384, sbo = variant == 1 ? 16384 : (variant == 2 ? 1024 : 512);
      for (int pass = 0; pass < 2; ++pass)
        for (int k = 0; k < 16; ++k) {      // K = pixels: 8 per step
          const uint64_t bd = make_desc_sw128_mnmajor((pass == 0 ? ahi : alo) + k * 1024, lbo, sbo, 1);
          mma_tf32_ts(tbase + D_COL, tbase + AHI_COL + k * 8, bd, idesc6, acc);
          acc = 1;
        }
    } else if (mode == 4) {
      constexpr uint32_t idesc4 = make_idesc_tf32(128, 64, 1, 1);
      for (int pass = 0; pass < 2; ++pass)
        for (int k = 0; k < 16; ++k) {      // K = rows (pixels): 8 per step, 1024 B apart
          const uint64_t ad = variant ? make_desc_sw128_mnmajor(ahi + k * 1024, 1024, 128 * 128) : make_desc_sw128_mnmajor(ahi + k * 1024, 128 * 128);   // hi0 hi1 lo0 lo1
          const uint64_t bd = variant ? make_desc_sw128_mnmajor((pass == 0 ? ahi : alo) + k * 1024, 1024, 128 * 128) : make_desc_sw128_mnmajor((pass == 0 ? ahi : alo) + k * 1024, 128 * 128);
          mma_tf32_ss(tbase + D_COL, ad, bd, idesc4, acc);
          acc = 1;
        }
    } else
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
  CK(cudaMalloc(&dS, 16));
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
  CUtensorMap tm32;
  r = encode(&tm32, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, dA, dims, strides, box, es,
             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B,
             CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { printf("encode (32B atom) failed %d\n", (int)r); return 2; }
  const size_t smem = sizeof(Smem) + 1024;
  CK(cudaFuncSetAttribute(tc_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  std::vector<double> ref(M * N);
  for (int m = 0; m < M; ++m) for (int n = 0; n < N; ++n) {
    double s = 0; for (int k = 0; k < K; ++k) s += (double)A[m * K + k] * W[n * K + k];
    ref[m * N + n] = s;
  }
  std::vector<double> ref3(M * N), ref4(64 * 64);
  for (int m = 0; m < M; ++m) for (int n = 0; n < N; ++n) {
    double s = 0; for (int k = 0; k < K; ++k) s += (double)A[m * K + k] * W[k * 64 + n];
    ref3[m * N + n] = s;
  }
  for (int m = 0; m < 64; ++m) for (int n = 0; n < 64; ++n) {
    double s = 0; for (int p = 0; p < M; ++p) s += (double)A[p * K + m] * A[p * K + n];
    ref4[m * 64 + n] = s;
  }
  int rc = 0;
  const char* names[7] = {"TS 3xTF32 (A in TMEM)", "SS 3xTF32 (A in smem)", "TS 1xTF32",
                          "SS, B MN-major SW128 (D = A W)", "SS, A and B MN-major SW128, stacked hi/lo (A^T A)",
                          "TS, B MN-major SW128_BASE32B (D = A W)",
                          "TS, A = X^T stacked hi/lo in TMEM, B MN-major SW128_BASE32B (X^T X)"};
  const int modes[9] = {0, 1, 2, 5, 5, 5, 6, 6, 6}, variants[9] = {0, 0, 0, 0, 1, 2, 0, 1, 2};
  for (int mi = 0; mi < 9; ++mi) {
    const int mode = modes[mi], variant = variants[mi];
    CK(cudaMemset(dD, 0, D.size() * 4)); CK(cudaMemset(dS, 0, 16));
    tc_gemm_kernel<<<1, 128, smem>>>(tm, tm32, dW, dD, mode, dS, variant);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("mode %d (%s): CUDA error %s\n", mode, names[mode], cudaGetErrorString(e)); return 3; }
    int st4[4] = {0, 0, 0, 0};
    CK(cudaMemcpy(st4, dS, 16, cudaMemcpyDeviceToHost));
    const int st = st4[0];
    if (mode == 6 && variant == 0)
      printf("   TMA SWIZZLE_128B_ATOM_32B landing layout: mismatches vs chunk32^=(row&3): %d, vs chunk32^=((row>>1)&3): %d\n", st4[1], st4[2]);
    CK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost));
    double maxerr = 0, maxref = 0;
    if (mode == 4 || mode == 6) {
      for (int i = 0; i < 64 * 64; ++i) {
        const double v = (double)D[i] + (double)D[64 * 64 + i];
        maxerr = fmax(maxerr, fabs(v - ref4[i])); maxref = fmax(maxref, fabs(ref4[i]));
      }
      printf("   hi rows D[0]=%f lo rows D[64*64]=%f ref %f\n", D[0], D[64 * 64], ref4[0]);
    } else {
      const std::vector<double>& rr = (mode == 3 || mode == 5) ? ref3 : ref;
      for (int i = 0; i < M * N; ++i) { maxerr = fmax(maxerr, fabs(D[i] - rr[i])); maxref = fmax(maxref, fabs(rr[i])); }
    }
    printf("mode %d%s (%s): status %d, max abs err %.3e, rel %.3e  (D[0]=%f ref %f, D[last]=%f ref %f)\n", mode,
           variant == 1 ? " [LBO/SBO swapped]" : (variant == 2 ? " [SBO 1024]" : ""), names[mode], st, maxerr, maxerr / maxref, D[0], ref[0], D[M * N - 1], ref[M * N - 1]);
    if (mode != 2 && !variant && (st != 0 || maxerr / maxref > 2e-6)) rc = 1;
    if (mode == 6 && variant == 0 && st4[1] != 0) rc = 1;
  }
  printf(rc == 0 ? "TC_GEMM_TEST PASS\n" : "TC_GEMM_TEST FAIL\n");
  return rc;
}
