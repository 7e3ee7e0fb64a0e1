// Fused ConvDPUnit forward with the pointwise 1x1 conv on the 5th-gen tensor cores (sm_100a):
//
//   TMA (cp.async.bulk.tensor.4d, SWIZZLE_128B, zero fill at the image border) brings the 16x16
//   halo tile of the NHWC pre-BN input into shared memory -> each thread reads ONE pixel row
//   (conflict-free thanks to the hardware swizzle), applies BN+ReLU, splits every value into
//   tf32 hi + lo (3xTF32 error compensation: single-pass TF32 misses the 1e-3 parity bar,
//   SURVEY §0) and stages them as the A operand in TENSOR MEMORY (tcgen05.st) -> one elected
//   thread issues tcgen05.mma.kind::tf32 (A from TMEM, W1 hi/lo from shared memory, K-major
//   SW128 descriptors, fp32 accumulators in TMEM): D = Alo*Bhi + Ahi*Blo + Ahi*Bhi ->
//   tcgen05.ld brings the accumulator row of each pixel back, + bias, zero outside the image ->
//   shared memory -> depthwise 3x3 stencil -> pre-BN output stored once + BN statistics.
//
// Persistent CTAs (2 per SM, 256 TMEM columns each), 14x14 output pixels per tile in two M=128
// row blocks that pipeline against each other (block 1 converts while block 0's MMAs run, block 0
// does its epilogue while block 1's MMAs run).  CIN = 64, plain load mode (no pool / up-add).
// Reference semantics: mmdet/models/utils/yunet_layer.py:30-36.
#include <cstdio>
#include <cstring>

#include "kernels.h"
#include "tc_common.cuh"
#include "f32x2.cuh"
#include "prefetch.cuh"

namespace yunet {

namespace {

using namespace tc;

constexpr int NT = 256;
constexpr int CIN = 64;
constexpr int HT = 16;            // halo tile edge
constexpr int OT = HT - 2;        // 14 output pixels per edge
constexpr int HPIX = HT * HT;     // 256 halo pixels = 2 x M128
constexpr uint32_t RAW_BYTES = HPIX * CIN * 4;   // 65536: TMA landing zone, later the y tile
constexpr uint32_t TMEM_COLS = 256;
constexpr uint32_t COL_D0 = 0, COL_D1 = 64, COL_AHI = 128, COL_ALO = 192;

__device__ __forceinline__ void bn_coeffs_tc(const BnRef& r, int c, float& scale, float& shift) {
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

template <int COUT, int MODE>
struct TcCfg {
  // NBUF = 2 (plain mode only) double-buffers the TMA landing zone so the next tile's load
  // overlaps the current tile's math, at 1 CTA / SM.  Measured on B200 (profiles/): two
  // co-resident CTAs with a single buffer each hide more latency (0.386 ms vs 0.464 ms on the
  // 80x80 unit) because the per-tile instruction stream, not the load, is the long pole.
  static constexpr int NBUF = 1;
  static constexpr int CTAS_PER_SM = (NBUF == 2) ? 1 : 2;
  static constexpr uint32_t B_BLOCK = COUT * 128;          // bytes of one k-block of W1 hi (or lo)
  static constexpr uint32_t OFF_BHI = NBUF * RAW_BYTES;
  static constexpr uint32_t OFF_BLO = OFF_BHI + 2 * B_BLOCK;
  static constexpr uint32_t OFF_W2 = OFF_BLO + 2 * B_BLOCK;           // [9][COUT]
  static constexpr uint32_t OFF_B1 = OFF_W2 + 9 * COUT * 4;
  static constexpr uint32_t OFF_B2 = OFF_B1 + COUT * 4;
  static constexpr uint32_t OFF_SC = OFF_B2 + COUT * 4;               // [64]
  static constexpr uint32_t OFF_SH = OFF_SC + CIN * 4;
  static constexpr uint32_t OFF_SCB = OFF_SH + CIN * 4;               // up-add operand b
  static constexpr uint32_t OFF_SHB = OFF_SCB + CIN * 4;
  static constexpr uint32_t OFF_BAR = OFF_SHB + CIN * 4;              // 4 mbarriers + tmem ptr
  static constexpr uint32_t SMEM = OFF_BAR + 64;
  static constexpr int NQ = COUT / 4;
  static constexpr int RGN = NT / (NQ * 16);
  static constexpr int RPT = (OT + RGN - 1) / RGN;
  static_assert((2 * B_BLOCK) % 1024 == 0, "operand alignment");
};

// y tile in shared memory: [256 pixels][COUT] fp32, 16-byte chunks XOR-swizzled with pixel & 7
template <int COUT>
__device__ __forceinline__ float* y_chunk(unsigned char* base, int pix, int chunk) {
  constexpr int ROWB = COUT * 4;
  constexpr int MASK = (COUT == 64) ? 7 : 3;     // 16 chunks (64 ch) / 4 chunks (16 ch) per pixel
  return reinterpret_cast<float*>(base + pix * ROWB + ((chunk ^ (pix & MASK)) << 4));
}

template <int COUT, int MODE>
__global__ void __launch_bounds__(NT, 2)
unit_fwd_tc_kernel(const __grid_constant__ CUtensorMap tmap, const UnitFwdArgs a, int* status) {
  using C = TcCfg<COUT, MODE>;
  extern __shared__ __align__(1024) unsigned char smem_dyn[];
  // round the base up to 1024 B with an OFFSET (not through an integer cast): the pointer stays in
  // the shared address space for the compiler, so every access below is LDS / STS, not generic LD / ST
  unsigned char* smem = smem_dyn + ((1024u - (smem_u32(smem_dyn) & 1023u)) & 1023u);
  unsigned char* raw0 = smem;                                 // NBUF x [2 m][2 kb][128 px][128 B]
  unsigned char* sBhi = smem + C::OFF_BHI;
  unsigned char* sBlo = smem + C::OFF_BLO;
  float* sW2 = reinterpret_cast<float*>(smem + C::OFF_W2);
  float* sB1 = reinterpret_cast<float*>(smem + C::OFF_B1);
  float* sB2 = reinterpret_cast<float*>(smem + C::OFF_B2);
  float* sSc = reinterpret_cast<float*>(smem + C::OFF_SC);
  float* sSh = reinterpret_cast<float*>(smem + C::OFF_SH);
  float* sScB = reinterpret_cast<float*>(smem + C::OFF_SCB);
  float* sShB = reinterpret_cast<float*>(smem + C::OFF_SHB);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::OFF_BAR);   // [0] tma buf0, [1] mma0, [2] mma1, [3] tma buf1
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 4);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int mblk = warp >> 2;            // which M=128 block this warp converts / reads back
  const int quarter = warp & 3;          // TMEM lane quarter this warp may touch

  // ---- one-time setup
  if (warp == 0) tmem_alloc<TMEM_COLS>(tmem_ptr);
  if (tid == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    mbar_init(&bars[2], 1);
    mbar_init(&bars[3], 1);
    mbar_fence_init();
    tma_prefetch_desc(&tmap);
  }
  for (int i = tid; i < COUT * CIN; i += NT) {        // W1[co][ci] -> hi / lo, K-major SW128
    const int n = i / CIN, k = i % CIN;
    const float w = __ldg(a.w1 + i);
    const uint32_t off = sw128_offset(COUT, n, k);
    *reinterpret_cast<uint32_t*>(sBhi + off) = tf32_hi(w);
    *reinterpret_cast<uint32_t*>(sBlo + off) = tf32_lo(w);
  }
  for (int i = tid; i < 9 * COUT; i += NT) {
    const int k = i / COUT, co = i % COUT;
    sW2[i] = __ldg(a.w2 + co * 9 + k);
  }
  if (tid < COUT) { sB1[tid] = __ldg(a.b1 + tid); sB2[tid] = __ldg(a.b2 + tid); }
  if (tid < CIN) {
    float sc, sh;
    bn_coeffs_tc(a.bna, tid, sc, sh);
    sSc[tid] = sc; sSh[tid] = sh;
    if (MODE == 2) {
      bn_coeffs_tc(a.bnb, tid, sc, sh);
      sScB[tid] = sc; sShB[tid] = sh;
    }
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tbase = warp_uniform(*tmem_ptr);
  const uint32_t lane_addr = tbase + ((uint32_t)(quarter * 32) << 16);
  constexpr uint32_t idesc = make_idesc_tf32(128, COUT);

  // depthwise-stage mapping + persistent statistics
  const int dq = tid % C::NQ;
  const int dx = (tid / C::NQ) % 16;
  const int drg = tid / (C::NQ * 16);
  const int dr0 = drg * C::RPT;
  const int dr1 = (dr0 + C::RPT < OT) ? dr0 + C::RPT : OT;
  double st1[4] = {0, 0, 0, 0}, st2[4] = {0, 0, 0, 0};
  float4 w2r[9];
#pragma unroll
  for (int k = 0; k < 9; ++k) w2r[k] = *reinterpret_cast<const float4*>(sW2 + k * COUT + dq * 4);
  const float4 bias2 = *reinterpret_cast<const float4*>(sB2 + dq * 4);

  const int tiles_x = (a.W + OT - 1) / OT, tiles_y = (a.H + OT - 1) / OT;
  const int ntiles = tiles_x * tiles_y * a.B;
  bool alive = true;
  uint32_t it = 0;
  // TMA: 4 boxes of (32 ch, 16 cols, 8 rows, 1 image) = 16 KB each into landing buffer `buf`
  auto issue_tma = [&](int tile_, int buf) {
    int t_ = tile_;
    const int tx_ = t_ % tiles_x; t_ /= tiles_x;
    const int ty_ = t_ % tiles_y;
    const int b_ = t_ / tiles_y;
    uint64_t* bar = &bars[buf == 0 ? 0 : 3];
    unsigned char* dst = raw0 + buf * RAW_BYTES;
    mbar_arrive_expect_tx(bar, RAW_BYTES);
#pragma unroll
    for (int m = 0; m < 2; ++m)
#pragma unroll
      for (int kb = 0; kb < 2; ++kb)
        tma_load_4d(dst + (m * 2 + kb) * 16384, &tmap, bar, kb * 32, tx_ * OT - 1, ty_ * OT - 1 + m * 8, b_);
  };
  if (MODE == 0 && tid == 0 && (int)blockIdx.x < ntiles) issue_tma(blockIdx.x, 0);
  for (int tile = blockIdx.x; tile < ntiles && alive; tile += gridDim.x, ++it) {
    const uint32_t ph = it & 1;
    int t = tile;
    const int tx = t % tiles_x; t /= tiles_x;
    const int ty = t % tiles_y;
    const int b = t / tiles_y;
    const int x0 = tx * OT, y0 = ty * OT;
    const int buf = (MODE == 0 && C::NBUF == 2) ? (int)(it & 1) : 0;
    unsigned char* raw = raw0 + buf * RAW_BYTES;

    if (MODE == 0) {
      // prefetch the next tile into the other buffer (its previous user finished at the barrier
      // that ended the last iteration), then wait for this tile's data
      if (C::NBUF == 2) {
        if (tid == 0 && tile + (int)gridDim.x < ntiles) issue_tma(tile + gridDim.x, buf ^ 1);
        if (!mbar_wait(&bars[buf == 0 ? 0 : 3], (it >> 1) & 1)) { alive = false; if (lane == 0) atomicExch(status, 1); }
      } else {
        if (tid == 0 && it > 0) issue_tma(tile, 0);       // tile 0 was issued before the loop
        // the landing buffer is busy until the end of the tile (the y tile aliases it): pull the
        // next tile's boxes into L2 now so that its TMA, issued one tile later, is an L2 hit
        if (tid == 32 && tile + (int)gridDim.x < ntiles) {
          int t_ = tile + gridDim.x;
          const int tx_ = t_ % tiles_x; t_ /= tiles_x;
          const int ty_ = t_ % tiles_y;
          const int b_ = t_ / tiles_y;
#pragma unroll
          for (int m = 0; m < 2; ++m)
#pragma unroll
            for (int kb = 0; kb < 2; ++kb)
              tma_prefetch_l2_4d(&tmap, kb * 32, tx_ * OT - 1, ty_ * OT - 1 + m * 8, b_);
        }
        if (!mbar_wait(&bars[0], ph)) { alive = false; if (lane == 0) atomicExch(status, 1); }
      }
    } else {
      // next tile's operand rows -> L2 (one bulk request per image row)
      if (tile + (int)gridDim.x < ntiles && warp == 7) {
        int t_ = tile + gridDim.x;
        const int tx_ = t_ % tiles_x; t_ /= tiles_x;
        const int ty_ = t_ % tiles_y;
        const int b_ = t_ / tiles_y;
        const int nx0 = tx_ * OT, ny0 = ty_ * OT;
        if (MODE == 1) {
          // pooled operand (4x the bytes): measured slower with the prefetch, left to the loads
        } else {
          l2_prefetch_tile<64>(a.za + (long long)b_ * a.H * a.W * 64, a.H, a.W, ny0 - 1, ny0 - 1 + HT,
                               nx0 - 1, nx0 - 1 + HT, lane);
          l2_prefetch_tile<64>(a.zb + (long long)b_ * (a.H >> 1) * (a.W >> 1) * 64, a.H >> 1, a.W >> 1,
                               (ny0 - 1) >> 1, ((ny0 - 1 + HT) >> 1) + 1, (nx0 - 1) >> 1,
                               ((nx0 - 1 + HT) >> 1) + 1, lane - 16);
        }
      }
      // ---- pooled / up-added operand: cooperative vector loads, activation applied on the way,
      // written in the layout the TMA would have produced (row = pixel, 16 B chunks ^ (row & 7))
#pragma unroll 4
      for (int it = 0; it < HPIX * 16 / NT; ++it) {
        const int i = tid + it * NT;
        const int pix = i >> 4, ch = i & 15;
        const int gy = y0 - 1 + pix / HT, gx = x0 - 1 + pix % HT;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (gy >= 0 && gy < a.H && gx >= 0 && gx < a.W) {
          const float4 sc = *reinterpret_cast<const float4*>(sSc + ch * 4);
          const float4 sh = *reinterpret_cast<const float4*>(sSh + ch * 4);
          if (MODE == 1) {
            const int W2 = a.W * 2;
            const float* p = a.za + (((long long)b * (a.H * 2) + gy * 2) * W2 + gx * 2) * CIN + ch * 4;
            const float4 z00 = __ldg(reinterpret_cast<const float4*>(p));
            const float4 z01 = __ldg(reinterpret_cast<const float4*>(p + CIN));
            const float4 z10 = __ldg(reinterpret_cast<const float4*>(p + (long long)W2 * CIN));
            const float4 z11 = __ldg(reinterpret_cast<const float4*>(p + (long long)W2 * CIN + CIN));
            v.x = fmaxf(fmaxf(fmaxf(fmaf(z00.x, sc.x, sh.x), fmaf(z01.x, sc.x, sh.x)), fmaxf(fmaf(z10.x, sc.x, sh.x), fmaf(z11.x, sc.x, sh.x))), 0.f);
            v.y = fmaxf(fmaxf(fmaxf(fmaf(z00.y, sc.y, sh.y), fmaf(z01.y, sc.y, sh.y)), fmaxf(fmaf(z10.y, sc.y, sh.y), fmaf(z11.y, sc.y, sh.y))), 0.f);
            v.z = fmaxf(fmaxf(fmaxf(fmaf(z00.z, sc.z, sh.z), fmaf(z01.z, sc.z, sh.z)), fmaxf(fmaf(z10.z, sc.z, sh.z), fmaf(z11.z, sc.z, sh.z))), 0.f);
            v.w = fmaxf(fmaxf(fmaxf(fmaf(z00.w, sc.w, sh.w), fmaf(z01.w, sc.w, sh.w)), fmaxf(fmaf(z10.w, sc.w, sh.w), fmaf(z11.w, sc.w, sh.w))), 0.f);
          } else {
            const float* p = a.za + (((long long)b * a.H + gy) * a.W + gx) * CIN + ch * 4;
            const float4 z = __ldg(reinterpret_cast<const float4*>(p));
            const int Hb = a.H >> 1, Wb = a.W >> 1;
            const float* pb = a.zb + (((long long)b * Hb + (gy >> 1)) * Wb + (gx >> 1)) * CIN + ch * 4;
            const float4 zb = __ldg(reinterpret_cast<const float4*>(pb));
            const float4 scb = *reinterpret_cast<const float4*>(sScB + ch * 4);
            const float4 shb = *reinterpret_cast<const float4*>(sShB + ch * 4);
            v.x = fmaxf(fmaf(z.x, sc.x, sh.x), 0.f) + fmaxf(fmaf(zb.x, scb.x, shb.x), 0.f);
            v.y = fmaxf(fmaf(z.y, sc.y, sh.y), 0.f) + fmaxf(fmaf(zb.y, scb.y, shb.y), 0.f);
            v.z = fmaxf(fmaf(z.z, sc.z, sh.z), 0.f) + fmaxf(fmaf(zb.z, scb.z, shb.z), 0.f);
            v.w = fmaxf(fmaf(z.w, sc.w, sh.w), 0.f) + fmaxf(fmaf(zb.w, scb.w, shb.w), 0.f);
          }
        }
        const int r = pix & 127;
        *reinterpret_cast<float4*>(raw + ((pix >> 7) * 2 + (ch >> 3)) * 16384 + r * 128 +
                                   (((ch & 7) ^ (r & 7)) << 4)) = v;
      }
      __syncthreads();
    }

    // ---- conversion of M block `mblk`: block 1 first waits until block 0's MMAs released the
    // shared A columns of TMEM
    const int r = quarter * 32 + lane;            // row inside the M block == TMEM lane
    if (mblk == 1 && alive) {
      if (!mbar_wait(&bars[1], ph)) { alive = false; if (lane == 0) atomicExch(status, 2); }
      tc_fence_after();
    }
    if (alive) {
      const unsigned char* rowp = raw + mblk * 32768 + r * 128;
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        uint32_t hi[16], lo[16];
#pragma unroll
        for (int c4 = 0; c4 < 4; ++c4) {
          const int c = g * 4 + c4;
          const int kb = c >> 3, cc = c & 7;
          const float4 z = *reinterpret_cast<const float4*>(rowp + kb * 16384 + ((cc ^ (r & 7)) << 4));
          const float4 sc = *reinterpret_cast<const float4*>(sSc + c * 4);
          const float4 sh = *reinterpret_cast<const float4*>(sSh + c * 4);
          float v0 = z.x, v1 = z.y, v2 = z.z, v3 = z.w;     // MODE 1/2: already activated
          if (MODE == 0) {
            v0 = fmaxf(fmaf(z.x, sc.x, sh.x), 0.f); v1 = fmaxf(fmaf(z.y, sc.y, sh.y), 0.f);
            v2 = fmaxf(fmaf(z.z, sc.z, sh.z), 0.f); v3 = fmaxf(fmaf(z.w, sc.w, sh.w), 0.f);
          }
          hi[c4 * 4 + 0] = tf32_hi(v0); lo[c4 * 4 + 0] = tf32_lo(v0);
          hi[c4 * 4 + 1] = tf32_hi(v1); lo[c4 * 4 + 1] = tf32_lo(v1);
          hi[c4 * 4 + 2] = tf32_hi(v2); lo[c4 * 4 + 2] = tf32_lo(v2);
          hi[c4 * 4 + 3] = tf32_hi(v3); lo[c4 * 4 + 3] = tf32_lo(v3);
        }
        tmem_st16(lane_addr + COL_AHI + g * 16, hi);
        tmem_st16(lane_addr + COL_ALO + g * 16, lo);
      }
      tmem_wait_st();
    }
    tc_fence_before();
    // the four warps of this M block meet, then one thread issues the block's 24 MMAs
    asm volatile("bar.sync %0, 128;" ::"r"(1 + mblk) : "memory");
    // whole first warp of the M block, one elected lane issues (see tc_common.cuh)
    if (warp_uniform((uint32_t)(warp & 3)) == 0 && __all_sync(0xffffffffu, alive)) {
      tc_fence_after();
      const uint32_t dcol = tbase + (mblk == 0 ? COL_D0 : COL_D1);
      // descriptors of the buffer bases once; an in-buffer byte offset adds (offset >> 4)
      const uint64_t dbhi = make_desc_sw128_kmajor(smem_u32(sBhi));
      const uint64_t dblo = make_desc_sw128_kmajor(smem_u32(sBlo));
      uint32_t acc = 0;
#pragma unroll
      for (int pass = 0; pass < 3; ++pass) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const uint32_t koff = ((k >> 2) * C::B_BLOCK + (k & 3) * 32) >> 4;
          const uint32_t at = tbase + (pass == 0 ? COL_ALO : COL_AHI) + k * 8;
          mma_tf32_ts_elect(dcol, at, (pass == 1 ? dblo : dbhi) + koff, idesc, acc);
          acc = 1;
        }
      }
      mma_commit_elect(&bars[1 + mblk]);
    }
    // ---- epilogue of M block `mblk`: accumulators -> y tile (over the consumed raw block)
    if (alive) {
      if (!mbar_wait(&bars[1 + mblk], ph)) { alive = false; if (lane == 0) atomicExch(status, 3); }
      tc_fence_after();
    }
    if (alive) {
      const int pix = mblk * 128 + r;
      const int gy = y0 - 1 + pix / HT, gx = x0 - 1 + pix % HT;
      const bool in = gy >= 0 && gy < a.H && gx >= 0 && gx < a.W;
      const uint32_t dcol = lane_addr + (mblk == 0 ? COL_D0 : COL_D1);
      unsigned char* ybase = raw + pix * (COUT * 4);
      const int yx = (pix & ((COUT == 64) ? 7 : 3)) << 4;
#pragma unroll
      for (int g = 0; g < COUT / 16; ++g) {
        uint32_t v[16];
        tmem_ld16(dcol + g * 16, v);
        tmem_wait_ld();
#pragma unroll
        for (int c4 = 0; c4 < 4; ++c4) {
          const float4 bb = *reinterpret_cast<const float4*>(sB1 + g * 16 + c4 * 4);
          float4 o;
          o.x = in ? __uint_as_float(v[c4 * 4 + 0]) + bb.x : 0.f;
          o.y = in ? __uint_as_float(v[c4 * 4 + 1]) + bb.y : 0.f;
          o.z = in ? __uint_as_float(v[c4 * 4 + 2]) + bb.z : 0.f;
          o.w = in ? __uint_as_float(v[c4 * 4 + 3]) + bb.w : 0.f;
          *reinterpret_cast<float4*>(ybase + (((g * 4 + c4) << 4) ^ yx)) = o;
        }
      }
    }
    tc_fence_before();
    __syncthreads();

    // ---- depthwise 3x3 on the 16x16 y tile -> 14x14 outputs, store z, statistics.
    // The swizzle term of a y-tile address depends only on the pixel COLUMN (the tile is 16 wide
    // and 16 % 8 == 0), so three column base pointers are formed once and the fully unrolled row
    // loop uses immediate offsets.
    if (alive && dx < OT && dr0 < OT) {
      constexpr int ROWB = COUT * 4;
      constexpr int MASK = (COUT == 64) ? 7 : 3;
      constexpr int RS = HT * ROWB;            // bytes between tile rows
      const unsigned char* col[3];
#pragma unroll
      for (int d = 0; d < 3; ++d) {
        const int px = dx + d;
        col[d] = raw + (dr0 * HT + px) * ROWB + ((dq ^ (px & MASK)) << 4);
      }
      float4 ra[3], rb[3], rc[3];
#pragma unroll
      for (int d = 0; d < 3; ++d) {
        ra[d] = *reinterpret_cast<const float4*>(col[d]);
        rb[d] = *reinterpret_cast<const float4*>(col[d] + RS);
      }
      float4 s1 = make_float4(0.f, 0.f, 0.f, 0.f), s2 = s1;
      const int gx = x0 + dx;
      float* dst0 = a.zout + (long long)b * a.out_batch_stride + ((long long)(y0 + dr0) * a.W + gx) * COUT + dq * 4;
      const long long dst_rs = (long long)a.W * COUT;
#pragma unroll
      for (int i = 0; i < C::RPT; ++i) {
        if (dr0 + i < dr1) {
#pragma unroll
          for (int d = 0; d < 3; ++d) rc[d] = *reinterpret_cast<const float4*>(col[d] + (i + 2) * RS);
          float4 o = bias2;
#pragma unroll
          for (int d = 0; d < 3; ++d) {
            fma4p(o, w2r[d], ra[d]); fma4p(o, w2r[3 + d], rb[d]); fma4p(o, w2r[6 + d], rc[d]);
          }
          if (y0 + dr0 + i < a.H && gx < a.W) {
            *reinterpret_cast<float4*>(dst0 + i * dst_rs) = o;
            s1.x += o.x; s1.y += o.y; s1.z += o.z; s1.w += o.w;
            fma4p(s2, o, o);
          }
#pragma unroll
          for (int d = 0; d < 3; ++d) { ra[d] = rb[d]; rb[d] = rc[d]; }
        }
      }
      st1[0] += s1.x; st1[1] += s1.y; st1[2] += s1.z; st1[3] += s1.w;
      st2[0] += s2.x; st2[1] += s2.y; st2[2] += s2.z; st2[3] += s2.w;
    }
    // the y tile (generic-proxy writes) is overwritten by the next TMA (async proxy); the
    // barrier also makes the (error-path) exit decision uniform across the CTA
    fence_proxy_async_smem();
    alive = __syncthreads_and(alive ? 1 : 0) != 0;
  }

  // ---- statistics: lanes sharing a channel quad reduce in the warp, then fp64 atomics
  if (a.osum != nullptr) {
#pragma unroll
    for (int c = 0; c < 4; ++c) {
#pragma unroll
      for (int o = 16; o >= C::NQ; o >>= 1) {
        st1[c] += __shfl_xor_sync(0xffffffffu, st1[c], o);
        st2[c] += __shfl_xor_sync(0xffffffffu, st2[c], o);
      }
    }
    if (lane < C::NQ) {
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        atomicAdd(a.osum + dq * 4 + c, st1[c]);
        atomicAdd(a.osumsq + dq * 4 + c, st2[c]);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<TMEM_COLS>(tbase);
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                             const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                             CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                             CUtensorMapFloatOOBfill);

EncodeFn get_encode() {
  static EncodeFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    cudaDriverEntryPointQueryResult q;
    void* p = nullptr;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess)
      fn = reinterpret_cast<EncodeFn>(p);
  }
  return fn;
}

template <int COUT, int MODE>
cudaError_t launch_tc_t(const CUtensorMap& tm, const UnitFwdArgs& a, int num_sms, int* status,
                        cudaStream_t s) {
  using C = TcCfg<COUT, MODE>;
  const size_t smem = C::SMEM + 1024;
  auto kern = unit_fwd_tc_kernel<COUT, MODE>;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    configured = true;
  }
  const int ntiles = ((a.W + OT - 1) / OT) * ((a.H + OT - 1) / OT) * a.B;
  int grid = C::CTAS_PER_SM * num_sms;
  if (grid > ntiles) grid = ntiles;
  kern<<<grid, NT, smem, s>>>(tm, a, status);
  return cudaGetLastError();
}

}  // namespace

int unit_fwd_tc_supported(int cin, int cout, int mode) {
  if (cin != 64 || get_encode() == nullptr) return 0;
  if (cout == 64) return mode >= 0 && mode <= 2;
  return cout == 16 && mode == 0;
}

// `status`: device int, set non-zero if a bounded wait inside the kernel timed out.
cudaError_t launch_unit_fwd_tc(int cout, int mode, const UnitFwdArgs& a, int num_sms, int* status,
                               cudaStream_t s) {
  EncodeFn enc = get_encode();
  if (!enc) return cudaErrorNotSupported;
  CUtensorMap tm;
  memset(&tm, 0, sizeof tm);
  if (mode == 0) {
    // NHWC input as a 4-D tensor (C, W, H, B); box = 32 channels x 16 cols x 8 rows x 1 image
    cuuint64_t dims[4] = {(cuuint64_t)CIN, (cuuint64_t)a.W, (cuuint64_t)a.H, (cuuint64_t)a.B};
    cuuint64_t strides[3] = {(cuuint64_t)CIN * 4, (cuuint64_t)a.W * CIN * 4,
                             (cuuint64_t)a.H * a.W * CIN * 4};
    cuuint32_t box[4] = {32, HT, 8, 1};
    cuuint32_t es[4] = {1, 1, 1, 1};
    CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(a.za), dims, strides,
                     box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return cudaErrorInvalidValue;
  }
  if (cout == 64 && mode == 0) return launch_tc_t<64, 0>(tm, a, num_sms, status, s);
  if (cout == 64 && mode == 1) return launch_tc_t<64, 1>(tm, a, num_sms, status, s);
  if (cout == 64 && mode == 2) return launch_tc_t<64, 2>(tm, a, num_sms, status, s);
  if (cout == 16 && mode == 0) return launch_tc_t<16, 0>(tm, a, num_sms, status, s);
  return cudaErrorInvalidValue;
}

}  // namespace yunet
