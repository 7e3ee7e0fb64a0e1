// Fused ConvDPUnit forward as a warp-specialised, persistent streaming pipeline (sm_100a).
// Reference semantics: mmdet/models/utils/yunet_layer.py:30-36  (relu(bn(dw3x3(pw1x1(x))))).
//
// The image is cut into vertical STRIPS of SW (<= 40) interior columns; a strip is streamed top to
// bottom in BLOCKS of RB rows x (SW + 2) halo columns (<= 128 pixels = the M dimension of one
// tcgen05.mma = the 128 TMEM lanes).  Only the two halo COLUMNS of a strip are recomputed; rows are
// never recomputed because the depthwise stage keeps its 3-row window in registers while the
// strip streams by (the halo ROWS exist only at the top / bottom of the image).  The global block
// sequence (image, strip, block) is split evenly over the CTAs (one per SM); a CTA that starts in
// the middle of a strip first replays the preceding block to prime the window (<= 1 extra block
// per CTA), so the load balance is exact for every layer shape.
//
// Roles (one CTA per SM, every role loops over the CTA's blocks; all hand-offs are mbarriers):
//   producer  MODE 0: one thread issues the TMA boxes of block j+2 (cp.async.bulk.tensor.4d,
//             SWIZZLE_128B / 64B, zero fill outside the image) into a 3-stage shared-memory ring.
//             MODE 1/2: four loader warps read the 2x2 max-pool window / the up-add pair with
//             128-bit coalesced loads, apply BN+ReLU and write the same swizzled layout.
//   convert   thread = pixel = TMEM lane: BN + ReLU (FFMA2), tf32 hi/lo split (3xTF32: single
//             TF32 misses the 1e-3 parity bar), tcgen05.st into the A buffer of the block's parity.
//             MODE 0 runs TWO groups of 4 warps that ping-pong on even / odd blocks, so one
//             group converts while the other waits for its MMAs and drains the accumulator.
//   mma       one elected thread: 3 x CIN/8 tcgen05.mma.kind::tf32 (A from TMEM, W1 hi/lo K-major
//             SW128 in shared memory), accumulators double-buffered in TMEM, tcgen05.commit.
//   epilogue  the convert warps: tcgen05.ld, + bias (FADD2), zero outside the image -> y ring (2 slots).
//   depthwise warp = 4 output columns, lane = channel pair; per block row 6 LDS.64, 3x3 stencil on
//             the register window (FFMA2), z stored once (coalesced 256 B per pixel), BN statistics
//             (fp32 per block, fp64 across blocks).  Measured alternatives (tools/dw_bench.cu,
//             tools/st_bench.cu, profiles/r2_*): the stage computes in 515 cycles per block but the
//             30 KB of stores need ~1050 cycles of the SM's store path whatever the instruction
//             (STG.64/128/256, cp.async.bulk); staging the row in shared memory + one bulk store per
//             row behind a DW-wide barrier was slower (0.28 ms vs 0.20 ms for the 80x80 unit).
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "kernels.h"
#include "tc_common.cuh"
#include "f32x2.cuh"
#include "tma_host.h"

namespace yunet {

namespace {

using namespace tc;

struct StripGeom {
  int SW;    // interior columns of a strip
  int SWH;   // SW + 2
  int RB;    // rows per block
  int NB;    // blocks per strip = ceil((H + 2) / RB)
  int nsx;   // strips per image
  int G;     // blocks in total = B * nsx * NB
  int dbg;   // YUNET_WS_DBG bit mask (perf experiments only): 1 skip depthwise work, 2 skip convert work, 4 skip epilogue work
};

__device__ __forceinline__ void bn_coeffs_ws(const BnRef& r, int c, float& scale, float& shift) {
  float m, v;
  if (r.train) {
    double dm = r.sum[c] * r.inv_count;
    double dv = r.sumsq[c] * r.inv_count - dm * dm;
    if (dv < 0.0) dv = 0.0;
    m = (float)dm; v = (float)dv;
  } else {
    m = r.rmean[c]; v = r.rvar[c];
  }
  const float rstd = 1.0f / sqrtf(v + kBnEps);
  scale = r.gamma[c] * rstd;
  shift = r.beta[c] - m * scale;
}

constexpr int NS = 3;                        // input ring stages
constexpr int NY = 2;                        // y ring slots
constexpr uint32_t TMEM_COLS = 512;
constexpr uint32_t COL_A = 0;                // A buffer b: hi at b*128, lo at b*128 + CIN
constexpr uint32_t COL_D = 256;              // D buffer b at 256 + b*64

template <int CIN, int COUT, int MODE>
struct WsCfg {
  static constexpr int ROWB = (CIN >= 32 ? 32 : CIN) * 4;   // bytes of one pixel row of a k-block (128 / 64)
  static constexpr int NKB = CIN >= 32 ? CIN / 32 : 1;      // k-blocks of 32 channels
  static constexpr int CPR = ROWB / 16;                     // 16-byte chunks per k-block row
  static constexpr int NCH = CIN / 4;                       // 16-byte chunks per pixel
  static constexpr uint32_t KB_BYTES = 128 * ROWB;
  static constexpr uint32_t STAGE_BYTES = NKB * KB_BYTES;
  static constexpr int KS = CIN / 8;                        // k-steps (tf32: K = 8 per MMA)
  static constexpr int NP = COUT / 2;                       // channel pairs (depthwise lanes)
  static constexpr int CGW = 32 / NP;                       // column groups per depthwise warp
  static constexpr int DW_WARPS = (10 + CGW - 1) / CGW;
  static constexpr int NG = (MODE == 0) ? 2 : 1;            // convert / epilogue groups
  static constexpr int LAG = (NG == 1) ? 1 : 0;             // NG == 1: convert(j+1) before epilogue(j)
  static constexpr int LD_WARPS = (MODE == 0) ? 0 : 4;
  static constexpr int W_TMA = 0, W_MMA = 1, W_CV = 2;
  static constexpr int W_DW = W_CV + 4 * NG;
  static constexpr int W_LD = W_DW + DW_WARPS;
  static constexpr int NWARPS = W_LD + LD_WARPS;
  static constexpr int NT = NWARPS * 32;
  static constexpr int YROW = COUT * 4;                     // bytes of one y pixel
  static constexpr uint32_t YSLOT = 128 * YROW;
  static constexpr uint32_t B_BLOCK = COUT * 128;           // one k-block (32 channels) of W1 hi (or lo)
  static constexpr uint32_t OFF_IN = 0;
  static constexpr uint32_t OFF_Y = (NS * STAGE_BYTES + 1023) / 1024 * 1024;
  static constexpr uint32_t OFF_BHI = OFF_Y + (NY * YSLOT + 1023) / 1024 * 1024;
  static constexpr uint32_t OFF_BLO = OFF_BHI + (NKB * B_BLOCK + 1023) / 1024 * 1024;
  static constexpr uint32_t OFF_W2 = OFF_BLO + (NKB * B_BLOCK + 1023) / 1024 * 1024;   // [9][COUT]
  static constexpr uint32_t OFF_B1 = OFF_W2 + 9 * COUT * 4;
  static constexpr uint32_t OFF_B2 = OFF_B1 + COUT * 4;
  static constexpr uint32_t OFF_SC = OFF_B2 + COUT * 4;               // scale / shift of operand a
  static constexpr uint32_t OFF_SH = OFF_SC + CIN * 4;
  static constexpr uint32_t OFF_SCB = OFF_SH + CIN * 4;               // operand b (up-add)
  static constexpr uint32_t OFF_SHB = OFF_SCB + CIN * 4;
  static constexpr uint32_t OFF_RED = OFF_SHB + CIN * 4;              // double [10 column groups][2][COUT]
  static constexpr uint32_t OFF_BAR = OFF_RED + 10 * 2 * COUT * 8;
  static constexpr uint32_t SMEM = OFF_BAR + 256;
  static_assert(OFF_RED % 8 == 0 && OFF_BAR % 8 == 0, "alignment");
  static_assert(CIN == 16 || CIN == 32 || CIN == 64, "CIN");
  static_assert(COUT == 16 || COUT == 32 || COUT == 64, "COUT");
};

// barrier indices
enum : int {
  BAR_IN_FULL = 0,                 // [NS]
  BAR_IN_EMPTY = BAR_IN_FULL + NS, // [NS]
  BAR_A_FULL = BAR_IN_EMPTY + NS,  // [2]
  BAR_MMA_DONE = BAR_A_FULL + 2,   // [2]
  BAR_D_EMPTY = BAR_MMA_DONE + 2,  // [2]
  BAR_Y_FULL = BAR_D_EMPTY + 2,    // [NY]
  BAR_Y_EMPTY = BAR_Y_FULL + NY,   // [NY]
  BAR_COUNT = BAR_Y_EMPTY + NY
};

struct BlkIter {
  int b, sx, blk;
  __device__ __forceinline__ void init(int g, const StripGeom& geo) {
    const int sid = g / geo.NB;
    blk = g - sid * geo.NB;
    b = sid / geo.nsx;
    sx = sid - b * geo.nsx;
  }
  __device__ __forceinline__ void next(const StripGeom& geo) {
    if (++blk == geo.NB) {
      blk = 0;
      if (++sx == geo.nsx) { sx = 0; ++b; }
    }
  }
};

// swizzle term of a y-ring pixel: depends on the pixel's COLUMN inside the block row only, so the
// depthwise stage can keep fixed per-thread column offsets and add a uniform row offset
template <int COUT>
__device__ __forceinline__ int y_swz(int col) { return COUT == 16 ? ((col >> 1) & 3) : (col & 7); }

__device__ __forceinline__ bool elect_one() {
  uint32_t p;
  asm volatile("{\n\t.reg .pred pe;\n\telect.sync _|pe, 0xffffffff;\n\tselp.u32 %0, 1, 0, pe;\n\t}" : "=r"(p));
  return p != 0;
}
__device__ __forceinline__ float2 lds64(uint32_t addr) {
  float2 v;
  asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(addr));
  return v;
}

// Swizzled row accesses as  (A ^ imm) + imm  computed next to the access (volatile): the compiler
// otherwise hoists the 16 loop-invariant swizzled addresses of a row out of the block loop and
// spills them (measured: 13 LDL per epilogue).  A = row base | (swizzle << 4), row base 16*CPR aligned.
__device__ __forceinline__ float4 lds128_xor(uint32_t A, uint32_t x) {
  float4 v;
  asm volatile("{\n\t.reg .u32 t;\n\txor.b32 t, %4, %5;\n\tld.shared.v4.f32 {%0, %1, %2, %3}, [t];\n\t}"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "r"(A), "r"(x));
  return v;
}
__device__ __forceinline__ void sts128_xor(uint32_t A, uint32_t xr, uint32_t x, uint32_t y, uint32_t z, uint32_t w) {
  asm volatile("{\n\t.reg .u32 t;\n\txor.b32 t, %0, %1;\n\tst.shared.v4.b32 [t], {%2, %3, %4, %5};\n\t}" ::"r"(A),
               "r"(xr), "r"(x), "r"(y), "r"(z), "r"(w)
               : "memory");
}

#ifdef YUNET_WS_TIMING
#define WT_DECL long long wt_t = clock64();
#define WT(k) do { if (blockIdx.x == 0 && (lane == 0 || warp == 1) && CIN == 64 && COUT == 64 && MODE == 0 && a.H >= 80) { const long long t_ = clock64(); atomicAdd(status + 32 + (k), (int)(t_ - wt_t)); wt_t = t_; } } while (0)
#else
#define WT_DECL
#define WT(k)
#endif

template <int CIN, int COUT, int MODE, int RBT>
__global__ void __launch_bounds__(WsCfg<CIN, COUT, MODE>::NT, 1)
unit_fwd_ws_kernel(const __grid_constant__ CUtensorMap tmap, const UnitFwdArgs a, const StripGeom geo,
                   int* status) {
  using C = WsCfg<CIN, COUT, MODE>;
  extern __shared__ __align__(1024) unsigned char smem_dyn[];
  // round the base up to 1024 B with an OFFSET (not through an integer cast): the pointer stays in
  // the shared address space for the compiler, so every access below is LDS / STS, not generic LD / ST
  unsigned char* smem = smem_dyn + ((1024u - (smem_u32(smem_dyn) & 1023u)) & 1023u);
  unsigned char* sIn = smem + C::OFF_IN;
  unsigned char* sY = smem + C::OFF_Y;
  unsigned char* sBhi = smem + C::OFF_BHI;
  unsigned char* sBlo = smem + C::OFF_BLO;
  float* sW2 = reinterpret_cast<float*>(smem + C::OFF_W2);
  float* sB1 = reinterpret_cast<float*>(smem + C::OFF_B1);
  float* sB2 = reinterpret_cast<float*>(smem + C::OFF_B2);
  float* sSc = reinterpret_cast<float*>(smem + C::OFF_SC);
  float* sSh = reinterpret_cast<float*>(smem + C::OFF_SH);
  float* sScB = reinterpret_cast<float*>(smem + C::OFF_SCB);
  float* sShB = reinterpret_cast<float*>(smem + C::OFF_SHB);
  double* sRed = reinterpret_cast<double*>(smem + C::OFF_RED);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::OFF_BAR);
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + BAR_COUNT);
  volatile int* abort_flag = reinterpret_cast<volatile int*>(tmem_ptr + 1);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int RB = RBT > 0 ? RBT : geo.RB;
  const int SWH = geo.SWH;

  // ---- one-time setup
  if (warp == C::W_MMA) tmem_alloc<TMEM_COLS>(tmem_ptr);
  if (tid == 0) {
    for (int i = 0; i < NS; ++i) {
      mbar_init(&bars[BAR_IN_FULL + i], MODE == 0 ? 1 : C::LD_WARPS);
      mbar_init(&bars[BAR_IN_EMPTY + i], 4);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&bars[BAR_A_FULL + i], 4);
      mbar_init(&bars[BAR_MMA_DONE + i], 1);
      mbar_init(&bars[BAR_D_EMPTY + i], 4);
    }
    for (int i = 0; i < NY; ++i) {
      mbar_init(&bars[BAR_Y_FULL + i], 4);
      mbar_init(&bars[BAR_Y_EMPTY + i], C::DW_WARPS);
    }
    *abort_flag = 0;
    mbar_fence_init();
    if (MODE == 0) tma_prefetch_desc(&tmap);
  }
  for (int i = tid; i < COUT * CIN; i += C::NT) {       // W1[co][ci] -> hi / lo, K-major SW128
    const int n = i / CIN, k = i % CIN;
    const float w = __ldg(a.w1 + i);
    const uint32_t off = sw128_offset(COUT, n, k);
    *reinterpret_cast<uint32_t*>(sBhi + off) = tf32_hi(w);
    *reinterpret_cast<uint32_t*>(sBlo + off) = tf32_lo(w);
  }
  for (int i = tid; i < 9 * COUT; i += C::NT) {
    const int k = i / COUT, co = i % COUT;
    sW2[i] = __ldg(a.w2 + co * 9 + k);
  }
  if (tid < COUT) { sB1[tid] = __ldg(a.b1 + tid); sB2[tid] = __ldg(a.b2 + tid); }
  if (tid < CIN) {
    float sc, sh;
    bn_coeffs_ws(a.bna, tid, sc, sh);
    sSc[tid] = sc; sSh[tid] = sh;
    if (MODE == 2) {
      bn_coeffs_ws(a.bnb, tid, sc, sh);
      sScB[tid] = sc; sShB[tid] = sh;
    }
  }
  for (int i = tid; i < 10 * 2 * COUT; i += C::NT) sRed[i] = 0.0;
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  // all 512 columns are allocated, so the allocation starts at TMEM address 0 (checked: a different
  // base would only cost the uniform-register issue path, but the kernel relies on it being constant)
  constexpr uint32_t tbase = 0;
  if (tid == 0 && *tmem_ptr != 0) { atomicExch(status, 20); *abort_flag = 1; }

  // ---- this CTA's share of the global block sequence (+ one priming block inside a strip)
  const int g0 = (int)(((long long)blockIdx.x * geo.G) / gridDim.x);
  const int g1 = (int)(((long long)(blockIdx.x + 1) * geo.G) / gridDim.x);
  const int prime = (g0 % geo.NB != 0) ? 1 : 0;
  const int gstart = g0 - prime;
  const int nblk = g1 - gstart;

  if (warp == C::W_TMA) {
    // ================================================================= TMA producer (MODE 0)
    if (MODE == 0 && lane == 0) {
      BlkIter it;
      it.init(gstart, geo);
      const uint32_t box_bytes = (uint32_t)(RB * SWH) * (uint32_t)(C::ROWB * C::NKB);
      for (int j = 0; j < nblk; ++j, it.next(geo)) {
        const int s = j % NS;
        const uint32_t par = (uint32_t)((j / NS) & 1);
        if (!mbar_wait_abort(&bars[BAR_IN_EMPTY + s], par ^ 1u, abort_flag)) { atomicExch(status, 21); break; }
        uint64_t* bar = &bars[BAR_IN_FULL + s];
        unsigned char* dst = sIn + s * C::STAGE_BYTES;
        mbar_arrive_expect_tx(bar, box_bytes);
        const int x = it.sx * geo.SW - 1, y = it.blk * RB - 1;
#pragma unroll
        for (int kb = 0; kb < C::NKB; ++kb) tma_load_4d(dst + kb * C::KB_BYTES, &tmap, bar, kb * 32, x, y, it.b);
      }
    }
  } else if (warp == C::W_MMA) {
    // ================================================================= MMA issuer
    // ONE elected thread runs the whole issue loop with warp-uniform operands only (TMEM base 0,
    // shared-memory descriptors from the uniform dynamic-smem base): the compiler then keeps every
    // descriptor in uniform registers and emits back-to-back UTCHMMA (1 instruction per MMA; with a
    // per-instruction elect it was ~11 instructions and ~48 cycles per 32-cycle MMA)
    constexpr uint32_t idesc = make_idesc_tf32(128, COUT);
    const uint64_t dbhi = make_desc_sw128_kmajor(smem_u32(sBhi));
    const uint64_t dblo = make_desc_sw128_kmajor(smem_u32(sBlo));
    if (elect_one()) {
      for (int j = 0; j < nblk; ++j) {
        const int buf = j & 1;
        const uint32_t par = (uint32_t)((j >> 1) & 1);
        WT_DECL
        if (!mbar_wait_abort(&bars[BAR_A_FULL + buf], par, abort_flag)) { atomicExch(status, 22); break; }
        if (!mbar_wait_abort(&bars[BAR_D_EMPTY + buf], par ^ 1u, abort_flag)) { atomicExch(status, 23); break; }
        WT(0);
        tc_fence_after();
        const uint32_t dcol = COL_D + buf * 64;
        const uint32_t ahi = COL_A + buf * 128, alo = ahi + CIN;
        uint32_t acc = 0;
#pragma unroll
        for (int pass = 0; pass < 3; ++pass) {
#pragma unroll
          for (int k = 0; k < C::KS; ++k) {
            const uint32_t koff = ((k >> 2) * C::B_BLOCK + (k & 3) * 32) >> 4;
            mma_tf32_ts(dcol, (pass == 0 ? alo : ahi) + k * 8, (pass == 1 ? dblo : dbhi) + koff, idesc, acc);
            acc = 1;
          }
        }
        mma_commit(&bars[BAR_MMA_DONE + buf]);
        WT(1);
      }
    }
    __syncwarp();
  } else if (warp >= C::W_CV && warp < C::W_CV + 4 * C::NG) {
    // ================================================================= convert + epilogue
    const int grp = (warp - C::W_CV) >> 2;              // ping-pong group: blocks j == grp (mod NG)
    const int quarter = warp & 3;                       // TMEM lane quarter of this warp
    const int m = quarter * 32 + lane;                  // pixel of the block == TMEM lane
    const uint32_t lane_addr = tbase + ((uint32_t)(quarter * 32) << 16);
    const int mr = m / SWH, mc = m - mr * SWH;
    const int swz_in = (C::ROWB == 128) ? (m & 7) : ((m >> 1) & 3);
    const int yx = y_swz<COUT>(mc);
    bool ok = true;
    for (int j = grp; j < nblk + C::LAG * C::NG && ok; j += C::NG) {
      WT_DECL
      if (j < nblk) {
        // ---- convert block j
        const int s = j % NS, buf = j & 1;
        if (!mbar_wait_abort(&bars[BAR_IN_FULL + s], (uint32_t)((j / NS) & 1), abort_flag)) { if (lane == 0) atomicExch(status, 24); ok = false; }
        ok = __all_sync(0xffffffffu, ok);
        WT(2);
        if (ok && !(geo.dbg & 2)) {
          const uint32_t rowA = smem_u32(sIn) + s * C::STAGE_BYTES + m * C::ROWB + (swz_in << 4);
          const uint32_t acol = lane_addr + COL_A + buf * 128;
#pragma unroll
          for (int g = 0; g < CIN / 16; ++g) {
            uint32_t hi[16], lo[16];
#pragma unroll
            for (int c4 = 0; c4 < 4; ++c4) {
              const int c = g * 4 + c4;
              const int kb = c / C::CPR, cc = c % C::CPR;
              const float4 z = lds128_xor(rowA + kb * C::KB_BYTES, (uint32_t)(cc << 4));
              float v0 = z.x, v1 = z.y, v2 = z.z, v3 = z.w;     // MODE 1/2: already activated
              if (MODE == 0) {
                const float4 sc = *reinterpret_cast<const float4*>(sSc + c * 4);
                const float4 sh = *reinterpret_cast<const float4*>(sSh + c * 4);
                v0 = sh.x; v1 = sh.y; v2 = sh.z; v3 = sh.w;
                fma2(v0, v1, z.x, z.y, sc.x, sc.y);
                fma2(v2, v3, z.z, z.w, sc.z, sc.w);
                v0 = fmaxf(v0, 0.f); v1 = fmaxf(v1, 0.f); v2 = fmaxf(v2, 0.f); v3 = fmaxf(v3, 0.f);
              }
              const float h0 = __uint_as_float(tf32_hi(v0)), h1 = __uint_as_float(tf32_hi(v1));
              const float h2 = __uint_as_float(tf32_hi(v2)), h3 = __uint_as_float(tf32_hi(v3));
              float l0, l1, l2, l3;
              sub2(l0, l1, v0, v1, h0, h1);
              sub2(l2, l3, v2, v3, h2, h3);
              hi[c4 * 4 + 0] = __float_as_uint(h0); hi[c4 * 4 + 1] = __float_as_uint(h1);
              hi[c4 * 4 + 2] = __float_as_uint(h2); hi[c4 * 4 + 3] = __float_as_uint(h3);
              lo[c4 * 4 + 0] = __float_as_uint(l0); lo[c4 * 4 + 1] = __float_as_uint(l1);
              lo[c4 * 4 + 2] = __float_as_uint(l2); lo[c4 * 4 + 3] = __float_as_uint(l3);
            }
            tmem_st16(acol + g * 16, hi);
            tmem_st16(acol + CIN + g * 16, lo);
          }
          tmem_wait_st();
        }
        if (ok) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            mbar_arrive(&bars[BAR_A_FULL + buf]);
            mbar_arrive(&bars[BAR_IN_EMPTY + s]);
          }
        }
        WT(3);
      }
      const int jj = j - C::LAG * C::NG;
      if (jj >= 0 && ok) {
        // ---- epilogue of block jj
        const int buf = jj & 1, slot = jj % NY;
        if (!mbar_wait_abort(&bars[BAR_MMA_DONE + buf], (uint32_t)((jj >> 1) & 1), abort_flag)) { if (lane == 0) atomicExch(status, 25); ok = false; }
        ok = __all_sync(0xffffffffu, ok);
        WT(4);
        if (ok) {
          tc_fence_after();
          // the accumulator row of this pixel, RAW (bias and the zero padding of y are applied by the
          // depthwise stage, see there): tcgen05.ld -> registers -> y ring
          if (!mbar_wait_abort(&bars[BAR_Y_EMPTY + slot], (uint32_t)(((jj / NY) & 1) ^ 1), abort_flag)) { if (lane == 0) atomicExch(status, 26); ok = false; }
          ok = __all_sync(0xffffffffu, ok);
          WT(5);
          if (ok) {
            const uint32_t yA = smem_u32(sY) + slot * C::YSLOT + m * C::YROW + (yx << 4);
            constexpr int PW = COUT < 32 ? COUT : 32;       // accumulator columns per piece (register budget)
#pragma unroll
            for (int h = 0; h < COUT / PW; ++h) {
              if (geo.dbg & 4) {
                if (h == COUT / PW - 1) { tc_fence_before(); __syncwarp(); if (lane == 0) mbar_arrive(&bars[BAR_D_EMPTY + buf]); }
                continue;
              }
              uint32_t v[PW];
              if (PW == 32) tmem_ld32(lane_addr + COL_D + buf * 64 + h * 32, v);
              else tmem_ld16(lane_addr + COL_D + buf * 64, v);
              tmem_wait_ld();
              if (h == COUT / PW - 1) {
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&bars[BAR_D_EMPTY + buf]);   // TMEM buffer back to the MMA warp
              }
#pragma unroll
              for (int c4 = 0; c4 < PW / 4; ++c4) {
                const int c = h * (PW / 4) + c4;
                sts128_xor(yA, (uint32_t)(c << 4), v[c4 * 4], v[c4 * 4 + 1], v[c4 * 4 + 2], v[c4 * 4 + 3]);
              }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&bars[BAR_Y_FULL + slot]);
          }
        }
        WT(6);
      }
    }
  } else if (warp >= C::W_DW && warp < C::W_DW + C::DW_WARPS) {
    // ================================================================= depthwise 3x3 + store + statistics
    const int q2 = lane % C::NP;                               // channel pair
    const int cg = (warp - C::W_DW) * C::CGW + lane / C::NP;   // column group: interior columns 4cg .. 4cg+3
    const bool active = cg < 10 && cg * 4 < geo.SW;
    float2 w2r[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) w2r[k] = *reinterpret_cast<const float2*>(sW2 + k * COUT + q2 * 2);
    // The y ring holds the RAW pointwise result u = W1 a (no bias).  With y = u + b1 inside the image
    // and y = 0 outside (the zero padding of the depthwise conv):
    //   z = b2 + sum_k w2_k y_k = (b2 + b1 sum_k w2_k) + sum_k w2_k u'_k,   u' = u inside, -b1 outside,
    // so interior blocks run on the raw values with one constant, and border blocks substitute -b1.
    float2 bias2 = *reinterpret_cast<const float2*>(sB2 + q2 * 2);
    const float2 b1v = *reinterpret_cast<const float2*>(sB1 + q2 * 2);
    {
      float sx_ = 0.f, sy_ = 0.f;
#pragma unroll
      for (int k = 0; k < 9; ++k) { sx_ += w2r[k].x; sy_ += w2r[k].y; }
      bias2.x = fmaf(b1v.x, sx_, bias2.x);
      bias2.y = fmaf(b1v.y, sy_, bias2.y);
    }
    const float2 nb1 = make_float2(-b1v.x, -b1v.y);
    float2 wa[6], wb[6];
#pragma unroll
    for (int d = 0; d < 6; ++d) { wa[d] = make_float2(0.f, 0.f); wb[d] = wa[d]; }
    // this thread's fp64 statistics live in its private slots of the shared reduction area
    double* myred = sRed + (cg < 10 ? cg : 0) * 2 * COUT + q2 * 2;
    // fixed per-thread offsets of the 6 halo columns inside a block row (clamped to the row)
    uint32_t coff[6];
#pragma unroll
    for (int d = 0; d < 6; ++d) {
      int cc = cg * 4 + d;
      if (cc > SWH - 1) cc = SWH - 1;
      coff[d] = (uint32_t)(cc * C::YROW + (((q2 >> 1) ^ y_swz<COUT>(cc)) << 4) + (q2 & 1) * 8);
    }
    const uint32_t ybase = smem_u32(sY);
    const uint32_t rowstride = (uint32_t)(SWH * C::YROW);
    const long long grow = (long long)a.W * COUT;               // floats per output row
    BlkIter it;
    it.init(gstart, geo);
    bool ok = true;
    for (int j = 0; j < nblk && ok; ++j, it.next(geo)) {
      const int slot = j % NY;
      WT_DECL
      if (!mbar_wait_abort(&bars[BAR_Y_FULL + slot], (uint32_t)((j / NY) & 1), abort_flag)) { if (lane == 0) atomicExch(status, 27); ok = false; }
      ok = __all_sync(0xffffffffu, ok);
      WT(7);
      if (!ok) break;
      if (active && !(geo.dbg & 1)) {
        const uint32_t ys = ybase + slot * C::YSLOT;
        const bool owned = j >= prime;
        const int x0 = it.sx * geo.SW + cg * 4;
        const int orow0 = it.blk * RB - 2;
        float* dst = a.zout + (long long)it.b * a.out_batch_stride + ((long long)orow0 * a.W + x0) * COUT + q2 * 2;
        // fast path (interior blocks: every row and column of this thread is a real output): no
        // predicates, no per-store address selection; the general path handles the image borders,
        // ragged strips and the priming block
        // (fast also needs every LOADED row / column inside the image: rows orow0+1 .. orow0+RB,
        // columns x0-1 .. x0+4)
        const bool fast = owned && orow0 >= 0 && orow0 + RB < a.H && (cg * 4 + 3 < geo.SW) && x0 >= 1 && (x0 + 4 < a.W);
        float s1x = 0.f, s1y = 0.f, s2x = 0.f, s2y = 0.f;
        if (fast) {
          auto row_fast = [&](int ii) {
            float2 nc[6];
            const uint32_t rbase = ys + (uint32_t)ii * rowstride;
#pragma unroll
            for (int d = 0; d < 6; ++d) nc[d] = lds64(rbase + coff[d]);
            float* drow = dst + (long long)ii * grow;
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              float ox = bias2.x, oy = bias2.y;
              if (!(geo.dbg & 16)) {
#pragma unroll
              for (int d = 0; d < 3; ++d) {
                fma2(ox, oy, w2r[d].x, w2r[d].y, wa[e + d].x, wa[e + d].y);
                fma2(ox, oy, w2r[3 + d].x, w2r[3 + d].y, wb[e + d].x, wb[e + d].y);
                fma2(ox, oy, w2r[6 + d].x, w2r[6 + d].y, nc[e + d].x, nc[e + d].y);
              }
              } else { ox += nc[e].x + nc[e + 1].x + nc[e + 2].x; oy += nc[e].y + nc[e+1].y + nc[e+2].y; }
              if (!(geo.dbg & 8)) *reinterpret_cast<float2*>(drow + e * COUT) = make_float2(ox, oy);
              add2(s1x, s1y, s1x, s1y, ox, oy);
              fma2(s2x, s2y, ox, oy, ox, oy);
            }
#pragma unroll
            for (int d = 0; d < 6; ++d) { wa[d] = wb[d]; wb[d] = nc[d]; }
          };
          if (RBT > 0) {
#pragma unroll
            for (int ii = 0; ii < (RBT > 0 ? RBT : 1); ++ii) row_fast(ii);
          } else {
            for (int ii = 0; ii < RB; ++ii) row_fast(ii);
          }
        } else {
          bool cok[4], cin[6];
#pragma unroll
          for (int e = 0; e < 4; ++e) cok[e] = owned && (cg * 4 + e < geo.SW) && (x0 + e < a.W);
#pragma unroll
          for (int d = 0; d < 6; ++d) cin[d] = (x0 - 1 + d >= 0) && (x0 - 1 + d < a.W);
          for (int ii = 0; ii < RB; ++ii) {
            float2 nc[6];
            const uint32_t rbase = ys + (uint32_t)ii * rowstride;
            const int lrow = orow0 + 1 + ii;                 // image row of the loaded y row
            const bool rin = lrow >= 0 && lrow < a.H;
#pragma unroll
            for (int d = 0; d < 6; ++d) {
              const float2 t = lds64(rbase + coff[d]);
              nc[d] = (rin && cin[d]) ? t : nb1;
            }
            const int orow = orow0 + ii;
            const bool rowok = orow >= 0 && orow < a.H;
            float* drow = dst + (long long)ii * grow;
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              float ox = bias2.x, oy = bias2.y;
#pragma unroll
              for (int d = 0; d < 3; ++d) {
                fma2(ox, oy, w2r[d].x, w2r[d].y, wa[e + d].x, wa[e + d].y);
                fma2(ox, oy, w2r[3 + d].x, w2r[3 + d].y, wb[e + d].x, wb[e + d].y);
                fma2(ox, oy, w2r[6 + d].x, w2r[6 + d].y, nc[e + d].x, nc[e + d].y);
              }
              if (rowok && cok[e]) {
                *reinterpret_cast<float2*>(drow + e * COUT) = make_float2(ox, oy);
                add2(s1x, s1y, s1x, s1y, ox, oy);
                fma2(s2x, s2y, ox, oy, ox, oy);
              }
            }
#pragma unroll
            for (int d = 0; d < 6; ++d) { wa[d] = wb[d]; wb[d] = nc[d]; }
          }
        }
        if (a.osum != nullptr && owned) {
          double2 t1 = *reinterpret_cast<double2*>(myred), t2 = *reinterpret_cast<double2*>(myred + COUT);
          t1.x += s1x; t1.y += s1y; t2.x += s2x; t2.y += s2y;
          *reinterpret_cast<double2*>(myred) = t1;
          *reinterpret_cast<double2*>(myred + COUT) = t2;
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars[BAR_Y_EMPTY + slot]);
      WT(8);
    }
  } else if (MODE != 0 && warp >= C::W_LD) {
    // ================================================================= loader warps (pool / up-add)
    constexpr int PPI = 128 / C::NCH;                     // pixels per pass of the 128 loader threads
    const int lt = (warp - C::W_LD) * 32 + lane;          // 0..127
    const int ch = lt % C::NCH;                           // 16-byte chunk of the pixel row
    const int p0 = lt / C::NCH;                           // first pixel; then += PPI
    const float4 sc = *reinterpret_cast<const float4*>(sSc + ch * 4);
    const float4 sh = *reinterpret_cast<const float4*>(sSh + ch * 4);
    float4 scb = sc, shb = sh;
    if (MODE == 2) {
      scb = *reinterpret_cast<const float4*>(sScB + ch * 4);
      shb = *reinterpret_cast<const float4*>(sShB + ch * 4);
    }
    const int npx = RB * SWH;
    const int kbo = (ch / C::CPR) * C::KB_BYTES, cc = ch % C::CPR;
    BlkIter it;
    it.init(gstart, geo);
    bool ok = true;
    for (int j = 0; j < nblk && ok; ++j, it.next(geo)) {
      const int s = j % NS;
      if (!mbar_wait_abort(&bars[BAR_IN_EMPTY + s], (uint32_t)(((j / NS) & 1) ^ 1), abort_flag)) { if (lane == 0) atomicExch(status, 28); ok = false; }
      ok = __all_sync(0xffffffffu, ok);
      if (!ok) break;
      unsigned char* dst = sIn + s * C::STAGE_BYTES + kbo;
      const int gy0 = it.blk * RB - 1, gx0 = it.sx * geo.SW - 1;
      // U pixels per step: all their global loads first (memory-level parallelism: the loaders are a
      // chain of dependent load round trips per block), then the math
      constexpr int U = 4;
      int pr = p0 / SWH, pc = p0 - pr * SWH;
      for (int p = p0; p < npx; p += U * PPI) {
        float4 zz[U][MODE == 1 ? 4 : 2];
        bool inb[U];
        int pp[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
          pp[u] = p + u * PPI;
          const int gy = gy0 + pr, gx = gx0 + pc;
          inb[u] = pp[u] < npx && gy >= 0 && gy < a.H && gx >= 0 && gx < a.W;
          if (inb[u]) {
            if (MODE == 1) {
              const int W2 = a.W * 2;
              const float* src = a.za + (((long long)it.b * (a.H * 2) + gy * 2) * W2 + gx * 2) * CIN + ch * 4;
              zz[u][0] = __ldg(reinterpret_cast<const float4*>(src));
              zz[u][1] = __ldg(reinterpret_cast<const float4*>(src + CIN));
              zz[u][2] = __ldg(reinterpret_cast<const float4*>(src + (long long)W2 * CIN));
              zz[u][3] = __ldg(reinterpret_cast<const float4*>(src + (long long)W2 * CIN + CIN));
            } else {
              zz[u][0] = __ldg(reinterpret_cast<const float4*>(
                  a.za + (((long long)it.b * a.H + gy) * a.W + gx) * CIN + ch * 4));
              const int Hb = a.H >> 1, Wb = a.W >> 1;
              zz[u][1] = __ldg(reinterpret_cast<const float4*>(
                  a.zb + (((long long)it.b * Hb + (gy >> 1)) * Wb + (gx >> 1)) * CIN + ch * 4));
            }
          }
          pc += PPI;
          while (pc >= SWH) { pc -= SWH; ++pr; }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
          if (pp[u] >= npx) continue;
          float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
          if (inb[u]) {
            if (MODE == 1) {
              const float4 z00 = zz[u][0], z01 = zz[u][1], z10 = zz[u][2], z11 = zz[u][3];
              v.x = fmaxf(fmaxf(fmaxf(fmaf(z00.x, sc.x, sh.x), fmaf(z01.x, sc.x, sh.x)), fmaxf(fmaf(z10.x, sc.x, sh.x), fmaf(z11.x, sc.x, sh.x))), 0.f);
              v.y = fmaxf(fmaxf(fmaxf(fmaf(z00.y, sc.y, sh.y), fmaf(z01.y, sc.y, sh.y)), fmaxf(fmaf(z10.y, sc.y, sh.y), fmaf(z11.y, sc.y, sh.y))), 0.f);
              v.z = fmaxf(fmaxf(fmaxf(fmaf(z00.z, sc.z, sh.z), fmaf(z01.z, sc.z, sh.z)), fmaxf(fmaf(z10.z, sc.z, sh.z), fmaf(z11.z, sc.z, sh.z))), 0.f);
              v.w = fmaxf(fmaxf(fmaxf(fmaf(z00.w, sc.w, sh.w), fmaf(z01.w, sc.w, sh.w)), fmaxf(fmaf(z10.w, sc.w, sh.w), fmaf(z11.w, sc.w, sh.w))), 0.f);
            } else {
              const float4 z = zz[u][0], zb = zz[u][1];
              v.x = fmaxf(fmaf(z.x, sc.x, sh.x), 0.f) + fmaxf(fmaf(zb.x, scb.x, shb.x), 0.f);
              v.y = fmaxf(fmaf(z.y, sc.y, sh.y), 0.f) + fmaxf(fmaf(zb.y, scb.y, shb.y), 0.f);
              v.z = fmaxf(fmaf(z.z, sc.z, sh.z), 0.f) + fmaxf(fmaf(zb.z, scb.z, shb.z), 0.f);
              v.w = fmaxf(fmaf(z.w, sc.w, sh.w), 0.f) + fmaxf(fmaf(zb.w, scb.w, shb.w), 0.f);
            }
          }
          const int px = pp[u];
          const int sw = (C::ROWB == 128) ? (px & 7) : ((px >> 1) & 3);
          *reinterpret_cast<float4*>(dst + px * C::ROWB + ((cc ^ sw) << 4)) = v;
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars[BAR_IN_FULL + s]);
    }
  }

  // ---- teardown: statistics of this CTA -> global, TMEM released
  tc_fence_before();
  __syncthreads();
  if (a.osum != nullptr && tid < 2 * COUT) {
    double s = 0.0;
#pragma unroll
    for (int w = 0; w < 10; ++w) s += sRed[w * 2 * COUT + tid];
    if (tid < COUT) atomicAdd(a.osum + tid, s);
    else atomicAdd(a.osumsq + (tid - COUT), s);
  }
  if (warp == C::W_MMA) {
    tc_fence_after();
    tmem_dealloc<TMEM_COLS>(tbase);
  }
}

template <int CIN, int COUT, int MODE, int RBT>
cudaError_t launch_ws_t(const CUtensorMap& tm, const UnitFwdArgs& a, const StripGeom& geo, int num_sms,
                        int* status, cudaStream_t s) {
  using C = WsCfg<CIN, COUT, MODE>;
  const size_t smem = C::SMEM + 1024;
  auto kern = unit_fwd_ws_kernel<CIN, COUT, MODE, RBT>;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    configured = true;
  }
  int grid = num_sms;
  if (grid > geo.G) grid = geo.G;
  kern<<<grid, C::NT, smem, s>>>(tm, a, geo, status);
  return cudaGetLastError();
}

template <int CIN, int COUT, int MODE>
cudaError_t launch_ws_rb(const CUtensorMap& tm, const UnitFwdArgs& a, const StripGeom& geo, int num_sms,
                         int* status, cudaStream_t s) {
  switch (geo.RB) {
    case 3: return launch_ws_t<CIN, COUT, MODE, 3>(tm, a, geo, num_sms, status, s);
    case 5: return launch_ws_t<CIN, COUT, MODE, 5>(tm, a, geo, num_sms, status, s);
    default: return launch_ws_t<CIN, COUT, MODE, 0>(tm, a, geo, num_sms, status, s);
  }
}

}  // namespace

int unit_fwd_ws_supported(int cin, int cout, int mode) {
  if (tma_encode_fn() == nullptr) return 0;
  if (cin == 64 && cout == 64) return mode >= 0 && mode <= 2;
  if (cin == 64 && cout == 16) return mode == 0;
  // 16 -> 16: a 126-pixel block moves only 16 KB, the per-block hand-offs dominate and the CUDA-core
  // kernel is as fast (measured 0.42 vs 0.41 ms, pooled 0.24 vs 0.18 ms): left to kernels_fwd.cu
  if (cin == 16 && (cout == 64 || cout == 32)) return mode == 0;
  if (cin == 32 && (cout == 32 || cout == 64)) return mode == 0;   // yunet_s stage 2
  return 0;
}

// `status`: device int, set non-zero if a bounded wait inside the kernel timed out.
cudaError_t launch_unit_fwd_ws(int cin, int cout, int mode, const UnitFwdArgs& a, int num_sms, int* status,
                               cudaStream_t s) {
  StripGeom geo;
  geo.nsx = (a.W + 39) / 40;
  geo.SW = (a.W + geo.nsx - 1) / geo.nsx;
  geo.SWH = geo.SW + 2;
  geo.RB = 128 / geo.SWH;
  if (geo.RB > a.H + 2) geo.RB = a.H + 2;
  geo.NB = (a.H + 2 + geo.RB - 1) / geo.RB;
  geo.G = a.B * geo.nsx * geo.NB;
  {
    static int dbg = -1;
    if (dbg < 0) { const char* e = getenv("YUNET_WS_DBG"); dbg = e ? atoi(e) : 0; }
    geo.dbg = dbg;
  }
  CUtensorMap tm;
  memset(&tm, 0, sizeof tm);
  if (mode == 0) {
    cudaError_t e = make_nhwc_map(&tm, a.za, cin, a.W, a.H, a.B, cin, (long long)a.H * a.W * cin,
                                  cin >= 32 ? 32 : cin, geo.SWH, geo.RB,
                                  cin >= 32 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B);
    if (e != cudaSuccess) return e;
  }
  if (cin == 64 && cout == 64 && mode == 0) return launch_ws_rb<64, 64, 0>(tm, a, geo, num_sms, status, s);
  if (cin == 64 && cout == 64 && mode == 1) return launch_ws_rb<64, 64, 1>(tm, a, geo, num_sms, status, s);
  if (cin == 64 && cout == 64 && mode == 2) return launch_ws_rb<64, 64, 2>(tm, a, geo, num_sms, status, s);
  if (cin == 64 && cout == 16 && mode == 0) return launch_ws_rb<64, 16, 0>(tm, a, geo, num_sms, status, s);
  if (cin == 16 && cout == 16 && mode == 0) return launch_ws_rb<16, 16, 0>(tm, a, geo, num_sms, status, s);
  if (cin == 16 && cout == 16 && mode == 1) return launch_ws_rb<16, 16, 1>(tm, a, geo, num_sms, status, s);
  if (cin == 16 && cout == 64 && mode == 0) return launch_ws_rb<16, 64, 0>(tm, a, geo, num_sms, status, s);
  if (cin == 16 && cout == 32 && mode == 0) return launch_ws_rb<16, 32, 0>(tm, a, geo, num_sms, status, s);
  if (cin == 32 && cout == 32 && mode == 0) return launch_ws_rb<32, 32, 0>(tm, a, geo, num_sms, status, s);
  if (cin == 32 && cout == 64 && mode == 0) return launch_ws_rb<32, 64, 0>(tm, a, geo, num_sms, status, s);
  return cudaErrorInvalidValue;
}

}  // namespace yunet
