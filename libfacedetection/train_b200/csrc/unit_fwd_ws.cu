// Fused ConvDPUnit forward (CIN = 64) as a warp-specialised, persistent streaming pipeline (sm_100a).
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
//             SWIZZLE_128B, zero fill outside the image) into a 3-stage shared-memory ring.
//             MODE 1/2: four loader warps read the 2x2 max-pool window / the up-add pair with
//             128-bit coalesced loads, apply BN+ReLU and write the same swizzled layout.
//   convert   4 warps, thread = pixel = TMEM lane: BN + ReLU, tf32 hi/lo split (3xTF32: single
//             TF32 misses the 1e-3 parity bar), tcgen05.st into one of two A buffers in TMEM.
//   mma       one elected thread: 24 x tcgen05.mma.kind::tf32 (A from TMEM, W1 hi/lo K-major SW128 in
//             shared memory), accumulators double-buffered in TMEM, tcgen05.commit -> mbarrier.
//   epilogue  the convert warps again (convert(j+1) runs before epilogue(j), so it overlaps the MMAs
//             of block j): tcgen05.ld, + bias, zero outside the image, -> y ring (2 slots).
//   depthwise 5 warps (COUT = 64): thread = (4 output columns, 4 channels); per block row 6 LDS.128,
//             3x3 stencil on the register window, z stored once with 128-bit coalesced stores,
//             BN statistics (fp32 per block, fp64 across blocks).
#include <cstdio>
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

constexpr int CIN = 64;
constexpr int NS = 3;                        // input ring stages
constexpr int NY = 2;                        // y ring slots
constexpr uint32_t STAGE_BYTES = 32768;      // 2 channel blocks x 128 pixels x 128 B
constexpr uint32_t TMEM_COLS = 512;
constexpr uint32_t COL_A = 0;                // A buffer b: hi at b*128, lo at b*128 + 64
constexpr uint32_t COL_D = 256;              // D buffer b at 256 + b*64

template <int COUT, int MODE>
struct WsCfg {
  static constexpr int NQ = COUT / 4;                  // channel quads
  static constexpr int DW_THREADS = 10 * NQ;           // 10 column groups of 4
  static constexpr int DW_WARPS = (DW_THREADS + 31) / 32;
  static constexpr int LD_WARPS = (MODE == 0) ? 0 : 4;
  static constexpr int W_TMA = 0, W_MMA = 1, W_CV = 2, W_DW = 6;
  static constexpr int W_LD = W_DW + DW_WARPS;
  static constexpr int NWARPS = W_LD + LD_WARPS;
  static constexpr int NT = NWARPS * 32;
  static constexpr uint32_t YSLOT = 128 * COUT * 4;
  static constexpr uint32_t B_BLOCK = COUT * 128;      // one k-block (32 channels) of W1 hi (or lo)
  static constexpr uint32_t OFF_IN = 0;
  static constexpr uint32_t OFF_Y = OFF_IN + NS * STAGE_BYTES;
  static constexpr uint32_t OFF_BHI = OFF_Y + ((NY * YSLOT + 1023) / 1024) * 1024;
  static constexpr uint32_t OFF_BLO = OFF_BHI + 2 * B_BLOCK;
  static constexpr uint32_t OFF_W2 = OFF_BLO + 2 * B_BLOCK;           // [9][COUT]
  static constexpr uint32_t OFF_B1 = OFF_W2 + 9 * COUT * 4;
  static constexpr uint32_t OFF_B2 = OFF_B1 + COUT * 4;
  static constexpr uint32_t OFF_SC = OFF_B2 + COUT * 4;               // [64] scale / shift of operand a
  static constexpr uint32_t OFF_SH = OFF_SC + CIN * 4;
  static constexpr uint32_t OFF_SCB = OFF_SH + CIN * 4;               // operand b (up-add)
  static constexpr uint32_t OFF_SHB = OFF_SCB + CIN * 4;
  static constexpr uint32_t OFF_RED = OFF_SHB + CIN * 4;              // double [DW_WARPS][2][COUT]
  static constexpr uint32_t OFF_BAR = OFF_RED + DW_WARPS * 2 * COUT * 8;
  static constexpr uint32_t SMEM = OFF_BAR + 256;
  static_assert((2 * B_BLOCK) % 1024 == 0, "operand alignment");
  static_assert(OFF_RED % 8 == 0 && OFF_BAR % 8 == 0, "alignment");
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

#ifdef YUNET_WS_TIMING
#define WT_DECL long long wt_t = clock64();
#define WT(k) do { if (blockIdx.x == 0 && lane == 0 && COUT == 64 && MODE == 0 && a.H >= 80) { const long long t_ = clock64(); atomicAdd(status + 32 + (k), (int)(t_ - wt_t)); wt_t = t_; } } while (0)
#else
#define WT_DECL
#define WT(k)
#endif

template <int COUT, int MODE, int RBT>
__global__ void __launch_bounds__(WsCfg<COUT, MODE>::NT, 1)
unit_fwd_ws_kernel(const __grid_constant__ CUtensorMap tmap, const UnitFwdArgs a, const StripGeom geo,
                   int* status) {
  using C = WsCfg<COUT, MODE>;
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
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tbase = warp_uniform(*tmem_ptr);

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
      const uint32_t box_bytes = (uint32_t)(RB * SWH) * 128u * 2u;
      for (int j = 0; j < nblk; ++j, it.next(geo)) {
        const int s = j % NS;
        const uint32_t par = (uint32_t)((j / NS) & 1);
        if (!mbar_wait_abort(&bars[BAR_IN_EMPTY + s], par ^ 1u, abort_flag)) { atomicExch(status, 21); break; }
        uint64_t* bar = &bars[BAR_IN_FULL + s];
        unsigned char* dst = sIn + s * STAGE_BYTES;
        mbar_arrive_expect_tx(bar, box_bytes);
        const int x = it.sx * geo.SW - 1, y = it.blk * RB - 1;
        tma_load_4d(dst, &tmap, bar, 0, x, y, it.b);
        tma_load_4d(dst + 16384, &tmap, bar, 32, x, y, it.b);
      }
    }
  } else if (warp == C::W_MMA) {
    // ================================================================= MMA issuer
    constexpr uint32_t idesc = make_idesc_tf32(128, COUT);
    const uint64_t dbhi = make_desc_sw128_kmajor(smem_u32(sBhi));
    const uint64_t dblo = make_desc_sw128_kmajor(smem_u32(sBlo));
    bool ok = true;
    for (int j = 0; j < nblk && ok; ++j) {
      const int buf = j & 1;
      const uint32_t par = (uint32_t)((j >> 1) & 1);
      WT_DECL
      if (!mbar_wait_abort(&bars[BAR_A_FULL + buf], par, abort_flag)) { if (lane == 0) atomicExch(status, 22); ok = false; }
      if (ok && !mbar_wait_abort(&bars[BAR_D_EMPTY + buf], par ^ 1u, abort_flag)) { if (lane == 0) atomicExch(status, 23); ok = false; }
      ok = __all_sync(0xffffffffu, ok);
      WT(0);
      if (ok) {
        tc_fence_after();
        const uint32_t dcol = tbase + COL_D + buf * 64;
        const uint32_t ahi = tbase + COL_A + buf * 128, alo = ahi + 64;
        uint32_t acc = 0;
#pragma unroll
        for (int pass = 0; pass < 3; ++pass) {
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const uint32_t koff = ((k >> 2) * C::B_BLOCK + (k & 3) * 32) >> 4;
            mma_tf32_ts_elect(dcol, (pass == 0 ? alo : ahi) + k * 8, (pass == 1 ? dblo : dbhi) + koff, idesc, acc);
            acc = 1;
          }
        }
        mma_commit_elect(&bars[BAR_MMA_DONE + buf]);
      }
      WT(1);
    }
  } else if (warp >= C::W_CV && warp < C::W_CV + 4) {
    // ================================================================= convert + epilogue
    const int quarter = warp & 3;                       // TMEM lane quarter of this warp
    const int m = quarter * 32 + lane;                  // pixel of the block == TMEM lane
    const uint32_t lane_addr = tbase + ((uint32_t)(quarter * 32) << 16);
    const int mr = m / SWH, mc = m - mr * SWH;
    const bool mvalid = m < RB * SWH;
    BlkIter itc, ite;
    itc.init(gstart, geo);
    ite = itc;
    bool ok = true;
    for (int j = 0; j <= nblk && ok; ++j) {
      WT_DECL
      if (j < nblk) {
        // ---- convert block j
        const int s = j % NS, buf = j & 1;
        if (!mbar_wait_abort(&bars[BAR_IN_FULL + s], (uint32_t)((j / NS) & 1), abort_flag)) { if (lane == 0) atomicExch(status, 24); ok = false; }
        ok = __all_sync(0xffffffffu, ok);
        WT(2);
        if (ok) {
          const unsigned char* rowp = sIn + s * STAGE_BYTES + m * 128;
          const uint32_t acol = lane_addr + COL_A + buf * 128;
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            uint32_t hi[16], lo[16];
#pragma unroll
            for (int c4 = 0; c4 < 4; ++c4) {
              const int c = g * 4 + c4;
              const int kb = c >> 3, cc = c & 7;
              const float4 z = *reinterpret_cast<const float4*>(rowp + kb * 16384 + ((cc ^ (m & 7)) << 4));
              float v0 = z.x, v1 = z.y, v2 = z.z, v3 = z.w;     // MODE 1/2: already activated
              if (MODE == 0) {
                const float4 sc = *reinterpret_cast<const float4*>(sSc + c * 4);
                const float4 sh = *reinterpret_cast<const float4*>(sSh + c * 4);
                v0 = fmaxf(fmaf(z.x, sc.x, sh.x), 0.f); v1 = fmaxf(fmaf(z.y, sc.y, sh.y), 0.f);
                v2 = fmaxf(fmaf(z.z, sc.z, sh.z), 0.f); v3 = fmaxf(fmaf(z.w, sc.w, sh.w), 0.f);
              }
              hi[c4 * 4 + 0] = tf32_hi(v0); lo[c4 * 4 + 0] = tf32_lo(v0);
              hi[c4 * 4 + 1] = tf32_hi(v1); lo[c4 * 4 + 1] = tf32_lo(v1);
              hi[c4 * 4 + 2] = tf32_hi(v2); lo[c4 * 4 + 2] = tf32_lo(v2);
              hi[c4 * 4 + 3] = tf32_hi(v3); lo[c4 * 4 + 3] = tf32_lo(v3);
            }
            tmem_st16(acol + g * 16, hi);
            tmem_st16(acol + 64 + g * 16, lo);
          }
          tmem_wait_st();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            mbar_arrive(&bars[BAR_A_FULL + buf]);
            mbar_arrive(&bars[BAR_IN_EMPTY + s]);
          }
        }
        itc.next(geo);
        WT(3);
      }
      if (j >= 1 && ok) {
        // ---- epilogue of block j-1
        const int jj = j - 1, buf = jj & 1, slot = jj % NY;
        if (!mbar_wait_abort(&bars[BAR_MMA_DONE + buf], (uint32_t)((jj >> 1) & 1), abort_flag)) { if (lane == 0) atomicExch(status, 25); ok = false; }
        ok = __all_sync(0xffffffffu, ok);
        WT(4);
        if (ok) {
          tc_fence_after();
          uint32_t v[COUT];
#pragma unroll
          for (int g = 0; g < COUT / 16; ++g) tmem_ld16(lane_addr + COL_D + buf * 64 + g * 16, v + g * 16);
          tmem_wait_ld();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&bars[BAR_D_EMPTY + buf]);
          if (!mbar_wait_abort(&bars[BAR_Y_EMPTY + slot], (uint32_t)(((jj / NY) & 1) ^ 1), abort_flag)) { if (lane == 0) atomicExch(status, 26); ok = false; }
          ok = __all_sync(0xffffffffu, ok);
          WT(5);
          if (ok) {
            const int gy = ite.blk * RB - 1 + mr, gx = ite.sx * geo.SW - 1 + mc;
            const bool in = mvalid && gy >= 0 && gy < a.H && gx >= 0 && gx < a.W;
            unsigned char* yrow = sY + slot * C::YSLOT + m * (COUT * 4);
            const int yx = (COUT == 64) ? (m & 7) : ((m >> 1) & 3);
#pragma unroll
            for (int c = 0; c < COUT / 4; ++c) {
              const float4 bb = *reinterpret_cast<const float4*>(sB1 + c * 4);
              float4 o;
              o.x = in ? __uint_as_float(v[c * 4 + 0]) + bb.x : 0.f;
              o.y = in ? __uint_as_float(v[c * 4 + 1]) + bb.y : 0.f;
              o.z = in ? __uint_as_float(v[c * 4 + 2]) + bb.z : 0.f;
              o.w = in ? __uint_as_float(v[c * 4 + 3]) + bb.w : 0.f;
              *reinterpret_cast<float4*>(yrow + ((c ^ yx) << 4)) = o;
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&bars[BAR_Y_FULL + slot]);
          }
        }
        ite.next(geo);
        WT(6);
      }
    }
  } else if (warp >= C::W_DW && warp < C::W_DW + C::DW_WARPS) {
    // ================================================================= depthwise 3x3 + store + statistics
    const int t = (warp - C::W_DW) * 32 + lane;
    const int q = t % C::NQ, cg = t / C::NQ;
    const bool active = t < C::DW_THREADS && cg * 4 < geo.SW;
    float4 w2r[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) w2r[k] = *reinterpret_cast<const float4*>(sW2 + k * COUT + q * 4);
    const float4 bias2 = *reinterpret_cast<const float4*>(sB2 + q * 4);
    float4 wa[6], wb[6];
#pragma unroll
    for (int d = 0; d < 6; ++d) { wa[d] = make_float4(0.f, 0.f, 0.f, 0.f); wb[d] = wa[d]; }
    double st1[4] = {0, 0, 0, 0}, st2[4] = {0, 0, 0, 0};
    // halo columns this thread reads (clamped to the block row) and their swizzled byte offsets
    int colp[6];
#pragma unroll
    for (int d = 0; d < 6; ++d) { const int cc = cg * 4 + d; colp[d] = cc < SWH ? cc : SWH - 1; }
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
      if (active) {
        const unsigned char* ys = sY + slot * C::YSLOT;
        const bool owned = j >= prime;
        const int x0 = it.sx * geo.SW + cg * 4;
        float* dst_img = a.zout + (long long)it.b * a.out_batch_stride + q * 4;
        float4 s1 = make_float4(0.f, 0.f, 0.f, 0.f), s2 = s1;
        auto row_step = [&](int ii) {
          float4 nc[6];
#pragma unroll
          for (int d = 0; d < 6; ++d) {
            const int p = ii * SWH + colp[d];
            const int sw = (COUT == 64) ? (q ^ (p & 7)) : (q ^ ((p >> 1) & 3));
            nc[d] = *reinterpret_cast<const float4*>(ys + p * (COUT * 4) + (sw << 4));
          }
          const int orow = it.blk * RB + ii - 2;
          const bool rowok = owned && orow >= 0 && orow < a.H;
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            float4 o = bias2;
#pragma unroll
            for (int d = 0; d < 3; ++d) {
              fma4p(o, w2r[d], wa[e + d]); fma4p(o, w2r[3 + d], wb[e + d]); fma4p(o, w2r[6 + d], nc[e + d]);
            }
            const int x = x0 + e;
            if (rowok && cg * 4 + e < geo.SW && x < a.W) {
              *reinterpret_cast<float4*>(dst_img + ((long long)orow * a.W + x) * COUT) = o;
              s1.x += o.x; s1.y += o.y; s1.z += o.z; s1.w += o.w;
              fma4p(s2, o, o);
            }
          }
#pragma unroll
          for (int d = 0; d < 6; ++d) { wa[d] = wb[d]; wb[d] = nc[d]; }
        };
        if (RBT > 0) {
#pragma unroll
          for (int ii = 0; ii < (RBT > 0 ? RBT : 1); ++ii) row_step(ii);
        } else {
          for (int ii = 0; ii < RB; ++ii) row_step(ii);
        }
        st1[0] += s1.x; st1[1] += s1.y; st1[2] += s1.z; st1[3] += s1.w;
        st2[0] += s2.x; st2[1] += s2.y; st2[2] += s2.z; st2[3] += s2.w;
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars[BAR_Y_EMPTY + slot]);
      WT(8);
    }
    // statistics: lanes sharing a channel quad reduce in the warp, one row per warp in shared memory
    if (a.osum != nullptr) {
      if (!(t < C::DW_THREADS)) {
#pragma unroll
        for (int c = 0; c < 4; ++c) { st1[c] = 0.0; st2[c] = 0.0; }
      }
#pragma unroll
      for (int c = 0; c < 4; ++c) {
#pragma unroll
        for (int o = 16; o >= C::NQ; o >>= 1) {
          st1[c] += __shfl_xor_sync(0xffffffffu, st1[c], o);
          st2[c] += __shfl_xor_sync(0xffffffffu, st2[c], o);
        }
      }
      if (lane < C::NQ) {
        double* r = sRed + (warp - C::W_DW) * 2 * COUT;
#pragma unroll
        for (int c = 0; c < 4; ++c) { r[lane * 4 + c] = st1[c]; r[COUT + lane * 4 + c] = st2[c]; }
      }
    }
  } else if (MODE != 0 && warp >= C::W_LD) {
    // ================================================================= loader warps (pool / up-add)
    const int lt = (warp - C::W_LD) * 32 + lane;          // 0..127
    const int ch = lt & 15;                               // 16-byte chunk of the pixel row
    const int p0 = lt >> 4;                               // first pixel; then += 8
    const float4 sc = *reinterpret_cast<const float4*>(sSc + ch * 4);
    const float4 sh = *reinterpret_cast<const float4*>(sSh + ch * 4);
    float4 scb = sc, shb = sh;
    if (MODE == 2) {
      scb = *reinterpret_cast<const float4*>(sScB + ch * 4);
      shb = *reinterpret_cast<const float4*>(sShB + ch * 4);
    }
    const int npx = RB * SWH;
    BlkIter it;
    it.init(gstart, geo);
    bool ok = true;
    for (int j = 0; j < nblk && ok; ++j, it.next(geo)) {
      const int s = j % NS;
      if (!mbar_wait_abort(&bars[BAR_IN_EMPTY + s], (uint32_t)(((j / NS) & 1) ^ 1), abort_flag)) { if (lane == 0) atomicExch(status, 28); ok = false; }
      ok = __all_sync(0xffffffffu, ok);
      if (!ok) break;
      unsigned char* dst = sIn + s * STAGE_BYTES + (ch >> 3) * 16384;
      const int gy0 = it.blk * RB - 1, gx0 = it.sx * geo.SW - 1;
      // (row, col) of pixel p0, advanced by 8 pixels per step without divisions
      int pr = p0 / SWH, pc = p0 - pr * SWH;
#pragma unroll 4
      for (int p = p0; p < npx; p += 8) {
        const int gy = gy0 + pr, gx = gx0 + pc;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (gy >= 0 && gy < a.H && gx >= 0 && gx < a.W) {
          if (MODE == 1) {
            const int W2 = a.W * 2;
            const float* src = a.za + (((long long)it.b * (a.H * 2) + gy * 2) * W2 + gx * 2) * CIN + ch * 4;
            const float4 z00 = __ldg(reinterpret_cast<const float4*>(src));
            const float4 z01 = __ldg(reinterpret_cast<const float4*>(src + CIN));
            const float4 z10 = __ldg(reinterpret_cast<const float4*>(src + (long long)W2 * CIN));
            const float4 z11 = __ldg(reinterpret_cast<const float4*>(src + (long long)W2 * CIN + CIN));
            v.x = fmaxf(fmaxf(fmaxf(fmaf(z00.x, sc.x, sh.x), fmaf(z01.x, sc.x, sh.x)), fmaxf(fmaf(z10.x, sc.x, sh.x), fmaf(z11.x, sc.x, sh.x))), 0.f);
            v.y = fmaxf(fmaxf(fmaxf(fmaf(z00.y, sc.y, sh.y), fmaf(z01.y, sc.y, sh.y)), fmaxf(fmaf(z10.y, sc.y, sh.y), fmaf(z11.y, sc.y, sh.y))), 0.f);
            v.z = fmaxf(fmaxf(fmaxf(fmaf(z00.z, sc.z, sh.z), fmaf(z01.z, sc.z, sh.z)), fmaxf(fmaf(z10.z, sc.z, sh.z), fmaf(z11.z, sc.z, sh.z))), 0.f);
            v.w = fmaxf(fmaxf(fmaxf(fmaf(z00.w, sc.w, sh.w), fmaf(z01.w, sc.w, sh.w)), fmaxf(fmaf(z10.w, sc.w, sh.w), fmaf(z11.w, sc.w, sh.w))), 0.f);
          } else {
            const float4 z = __ldg(reinterpret_cast<const float4*>(
                a.za + (((long long)it.b * a.H + gy) * a.W + gx) * CIN + ch * 4));
            const int Hb = a.H >> 1, Wb = a.W >> 1;
            const float4 zb = __ldg(reinterpret_cast<const float4*>(
                a.zb + (((long long)it.b * Hb + (gy >> 1)) * Wb + (gx >> 1)) * CIN + ch * 4));
            v.x = fmaxf(fmaf(z.x, sc.x, sh.x), 0.f) + fmaxf(fmaf(zb.x, scb.x, shb.x), 0.f);
            v.y = fmaxf(fmaf(z.y, sc.y, sh.y), 0.f) + fmaxf(fmaf(zb.y, scb.y, shb.y), 0.f);
            v.z = fmaxf(fmaf(z.z, sc.z, sh.z), 0.f) + fmaxf(fmaf(zb.z, scb.z, shb.z), 0.f);
            v.w = fmaxf(fmaf(z.w, sc.w, sh.w), 0.f) + fmaxf(fmaf(zb.w, scb.w, shb.w), 0.f);
          }
        }
        *reinterpret_cast<float4*>(dst + p * 128 + (((ch & 7) ^ (p & 7)) << 4)) = v;
        pc += 8;
        while (pc >= SWH) { pc -= SWH; ++pr; }
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
    for (int w = 0; w < C::DW_WARPS; ++w) s += sRed[w * 2 * COUT + tid];
    if (tid < COUT) atomicAdd(a.osum + tid, s);
    else atomicAdd(a.osumsq + (tid - COUT), s);
  }
  if (warp == C::W_MMA) {
    tc_fence_after();
    tmem_dealloc<TMEM_COLS>(tbase);
  }
}

template <int COUT, int MODE, int RBT>
cudaError_t launch_ws_t(const CUtensorMap& tm, const UnitFwdArgs& a, const StripGeom& geo, int num_sms,
                        int* status, cudaStream_t s) {
  using C = WsCfg<COUT, MODE>;
  const size_t smem = C::SMEM + 1024;
  auto kern = unit_fwd_ws_kernel<COUT, MODE, RBT>;
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

template <int COUT, int MODE>
cudaError_t launch_ws_rb(const CUtensorMap& tm, const UnitFwdArgs& a, const StripGeom& geo, int num_sms,
                         int* status, cudaStream_t s) {
  switch (geo.RB) {
    case 3: return launch_ws_t<COUT, MODE, 3>(tm, a, geo, num_sms, status, s);
    case 5: return launch_ws_t<COUT, MODE, 5>(tm, a, geo, num_sms, status, s);
    default: return launch_ws_t<COUT, MODE, 0>(tm, a, geo, num_sms, status, s);
  }
}

}  // namespace

int unit_fwd_ws_supported(int cin, int cout, int mode) {
  if (cin != 64 || tma_encode_fn() == nullptr) return 0;
  if (cout == 64) return mode >= 0 && mode <= 2;
  return cout == 16 && mode == 0;
}

// `status`: device int, set non-zero if a bounded wait inside the kernel timed out.
cudaError_t launch_unit_fwd_ws(int cout, int mode, const UnitFwdArgs& a, int num_sms, int* status,
                               cudaStream_t s) {
  StripGeom geo;
  geo.nsx = (a.W + 39) / 40;
  geo.SW = (a.W + geo.nsx - 1) / geo.nsx;
  geo.SWH = geo.SW + 2;
  geo.RB = 128 / geo.SWH;
  if (geo.RB > a.H + 2) geo.RB = a.H + 2;
  geo.NB = (a.H + 2 + geo.RB - 1) / geo.RB;
  geo.G = a.B * geo.nsx * geo.NB;
  CUtensorMap tm;
  memset(&tm, 0, sizeof tm);
  if (mode == 0) {
    cudaError_t e = make_nhwc_map(&tm, a.za, CIN, a.W, a.H, a.B, CIN, (long long)a.H * a.W * CIN, 32,
                                  geo.SWH, geo.RB, CU_TENSOR_MAP_SWIZZLE_128B);
    if (e != cudaSuccess) return e;
  }
  if (cout == 64 && mode == 0) return launch_ws_rb<64, 0>(tm, a, geo, num_sms, status, s);
  if (cout == 64 && mode == 1) return launch_ws_rb<64, 1>(tm, a, geo, num_sms, status, s);
  if (cout == 64 && mode == 2) return launch_ws_rb<64, 2>(tm, a, geo, num_sms, status, s);
  if (cout == 16 && mode == 0) return launch_ws_rb<16, 0>(tm, a, geo, num_sms, status, s);
  return cudaErrorInvalidValue;
}

}  // namespace yunet
