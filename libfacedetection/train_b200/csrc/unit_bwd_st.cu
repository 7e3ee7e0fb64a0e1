// Fused ConvDPUnit backward (64 -> 64 channels, BatchNorm on the output) on STRIPS (sm_100a): the
// tcgen05 kernel of unit_bwd_tc.cu with the tile geometry of the streaming forward kernel.
//
//   strip   = SW (<= 40) interior columns + 2 halo columns, streamed top to bottom in blocks of RB
//             rows (RB x (SW+2) <= 128 pixels = one M=128 MMA block = the 128 TMEM lanes); the
//             interior is RB x SW of the 128 lanes (120 of 128 for 3 x 40, vs 84 of 128 for the
//             6 x 14 interior of the 8 x 16 halo tiles of unit_bwd_tc.cu)
//   rows    are never recomputed: the depthwise-backward stage keeps a 2-row window of g in registers
//             while the strip streams by.  The du / z_out boxes of a block are shifted one row DOWN
//             against its z_in box (rows 3j+1 .. 3j+3 vs 3j .. 3j+2): with the window rows 3j-1, 3j
//             the new g rows give dy for exactly the block's own rows, so dy, a and z_in of a pixel
//             live at the same tile index for the three GEMMs and the epilogue
//   prime   a block sequence that starts inside a strip (CTA range boundary) or at its top first
//             runs a g-only step on the preceding block (TMA du / z_out, g pass, window roll)
//   balance the global (image, strip, block) sequence is split evenly over the CTAs
//
//   per block: TMA z_in (SWIZZLE_128B_ATOM_32B), du, z_out (SWIZZLE_128B), du / z_out prefetched one
//   step ahead -> g pass in place -> a = relu(bn(z_in)) rows -> tf32 hi/lo -> TMEM + shared ->
//   MMA 1 (y recomputed) -> depthwise backward on the register window (dy, dW2, db2, db1) ->
//   dy rows -> TMEM -> MMA 2 (h = dy W1) and dy^T -> TMEM -> MMA 3 (dW1 tile, K = the 128 pixels,
//   B = the pixel-major a tile as an MN-major operand) -> epilogue (ReLU mask / pool winner /
//   up-add children, statistics, du_in stored once).  All MMAs are issued by one elected thread with
//   warp-uniform operands (TMEM base 0), i.e. back-to-back UTCHMMA.
// Same math as unit_bwd_kernel (kernels_bwd.cu); semantics: autograd of yunet_layer.py:30-36.
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "kernels.h"
#include "tc_common.cuh"
#include "f32x2.cuh"
#include "prefetch.cuh"
#include "tma_host.h"

namespace yunet {

namespace {

using namespace tc;

constexpr int NT = 256;
constexpr int C64 = 64;
constexpr uint32_t TILE_BYTES = 128 * C64 * 4;   // 32 KB
constexpr uint32_t TMEM_COLS = 512;
constexpr uint32_t COL_D1 = 0, COL_D2 = 64, COL_AHI = 128, COL_ALO = 192;
constexpr uint32_t COL_Z = 256;      // z_in rows (MODE 0), read back by the epilogue
constexpr uint32_t COL_DYT = 320;    // dy^T, 128 pixel columns
constexpr uint32_t COL_DW = 448;     // per-tile dW1: lanes 0..63 dy_hi^T a, lanes 64..127 dy_lo^T a

struct StripB {
  int SW;    // interior columns of a strip
  int SWH;   // SW + 2
  int RB;    // rows per block
  int NB;    // blocks per strip = ceil(H / RB)
  int nsx;   // strips per image
  int G;     // blocks in total = B * nsx * NB
};

// one step of a CTA's sequence: a real block or the g-only priming step in front of it
struct Step {
  int b, sx, blk;    // image, strip, block (for a priming step: the block BEFORE the one it primes)
  int prime;         // 1: g-only
  int valid;
};

__device__ __forceinline__ bool elect_one_b() {
  uint32_t p;
  asm volatile("{\n\t.reg .pred pe;\n\telect.sync _|pe, 0xffffffff;\n\tselp.u32 %0, 1, 0, pe;\n\t}" : "=r"(p));
  return p != 0;
}

struct Off {
  static constexpr uint32_t RAW = 0;                       // z_in tile (TMA, 2 ch-blocks of 16 KB) -> a_hi
  static constexpr uint32_t AL = RAW + TILE_BYTES;         // a_lo (same layout); h tile in MODE 1/2
  static constexpr uint32_t G = AL + TILE_BYTES;           // du (TMA) -> g halo tile, TMA layout
  static constexpr uint32_t Y = G + TILE_BYTES;            // z_out (TMA) -> y -> dy tile [128][64] swizzled
  static constexpr uint32_t B1HI = Y + TILE_BYTES;         // W1 hi  [co][ci] K-major SW128
  static constexpr uint32_t B1LO = B1HI + 16384;
  static constexpr uint32_t B2HI = B1LO + 16384;           // W1^T hi [ci][co] K-major SW128
  static constexpr uint32_t B2LO = B2HI + 16384;
  static constexpr uint32_t W2 = B2LO + 16384;             // [9][64]
  static constexpr uint32_t B1 = W2 + 9 * 64 * 4;          // bias1 [64]
  static constexpr uint32_t CA = B1 + 256;                 // scale, shift, mean, rstd of the input [4][64]
  static constexpr uint32_t CB = CA + 1024;                // same for the up-sampled operand b [4][64]
  static constexpr uint32_t CO = CB + 1024;                // gscale, m1, m2, mean, rstd of the output [5][64]
  static constexpr uint32_t BAR = CO + 1280;               // 6 mbarriers + tmem ptr
  static constexpr uint32_t TOTAL = BAR + 64;
};

// [128 pixels][64 ch] fp32 tile, 16-byte chunks XOR-swizzled with (pixel & 7)
__device__ __forceinline__ float* tchunk(unsigned char* base, int pix, int chunk) {
  return reinterpret_cast<float*>(base + pix * 256 + ((chunk ^ (pix & 7)) << 4));
}
// chunk of the TMA-written z_in tile: [kblock][128 pixels][128 B], chunks ^ (pixel & 7)
__device__ __forceinline__ const float* rchunk(const unsigned char* raw, int pix, int chunk) {
  return reinterpret_cast<const float*>(raw + (chunk >> 3) * 16384 + pix * 128 +
                                        (((chunk & 7) ^ (pix & 7)) << 4));
}

// byte offset of a 16-byte chunk inside the z_in / a tiles: [ch block][128 pixels][128 B] with the
// 32-byte halves of a row XOR-swizzled by (pixel & 3)  (TMA SWIZZLE_128B_ATOM_32B == UMMA
// SWIZZLE_128B_BASE32B, so the tile doubles as the MN-major B operand of the dW1 GEMM)
__device__ __forceinline__ uint32_t zoff(int pix, int chunk) {
  return (uint32_t)((chunk >> 3) * 16384 + pix * 128 +
                    (((((chunk & 7) >> 1) ^ (pix & 3)) << 5) | ((chunk & 1) << 4)));
}

// Optional per-phase cycle counters (make TIMING=1): thread 0 of CTA 0 adds the clock64() deltas of
// the 80x80 plain units to status[32 + phase]; read with tools/phase_timing.py.
#ifdef YUNET_PHASE_TIMING
#define PT_DECL long long pt_t0 = clock64(); const bool pt_on = (blockIdx.x == 0 && threadIdx.x == 0 && MODE == 0 && a.H >= 80);
#define PT(k) do { if (pt_on) { const long long t_ = clock64(); atomicAdd(status + 32 + (k), (int)(t_ - pt_t0)); pt_t0 = t_; } } while (0)
#else
#define PT_DECL
#define PT(k)
#endif

// 256-bit global accesses (sm_100: LDG/STG.E.ENL2.256), 32-byte aligned
__device__ __forceinline__ void stg_v8(float* p, const float (&v)[8]) {
  asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "f"(v[0]), "f"(v[1]),
               "f"(v[2]), "f"(v[3]), "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7])
               : "memory");
}
__device__ __forceinline__ void ldg_v8(const float* p, float (&v)[8]) {
  asm volatile("ld.global.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]),
                 "=f"(v[7])
               : "l"(p)
               : "memory");
}

struct Coef4 { float scale, shift, mean, rstd; };
__device__ __forceinline__ Coef4 bn_coef_tc(const BnRef& r, int c) {
  Coef4 k;
  double dm = r.sum[c] * r.inv_count;
  double dv = r.sumsq[c] * r.inv_count - dm * dm;
  if (dv < 0.0) dv = 0.0;
  k.mean = (float)dm;
  k.rstd = 1.0f / sqrtf((float)dv + kBnEps);
  k.scale = r.gamma[c] * k.rstd;
  k.shift = r.beta[c] - k.mean * k.scale;
  return k;
}

template <int MODE, int RBT>
__global__ void __launch_bounds__(NT, 1)
unit_bwd_st_kernel(const __grid_constant__ CUtensorMap tmap, const __grid_constant__ CUtensorMap tmap_du,
                   const __grid_constant__ CUtensorMap tmap_zo, const UnitBwdArgs a, const StripB geo, int* status) {
  extern __shared__ __align__(1024) unsigned char smem_dyn[];
  unsigned char* smem = smem_dyn + ((1024u - (smem_u32(smem_dyn) & 1023u)) & 1023u);
  unsigned char* raw = smem + Off::RAW;               // MODE 0: z_in -> a_hi; else activated a -> a_hi
  unsigned char* sAL = smem + Off::AL;               // a_lo
  unsigned char* sH = smem + Off::AL;                // MODE 1/2: h tile for the routing pass
  unsigned char* sG = smem + Off::G;
  unsigned char* sY = smem + Off::Y;
  float* sW2 = reinterpret_cast<float*>(smem + Off::W2);
  float* sB1 = reinterpret_cast<float*>(smem + Off::B1);
  float* sCa = reinterpret_cast<float*>(smem + Off::CA);
  float* sCb = reinterpret_cast<float*>(smem + Off::CB);
  float* sCo = reinterpret_cast<float*>(smem + Off::CO);
  // [0] z_in, [1] mma1, [2] mma2, [3] mma3 (dW1): one completion per REAL block; [4] du, [5] z_out:
  // one completion per step (real or priming)
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Off::BAR);
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 6);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int warp_u = (int)warp_uniform((uint32_t)warp);      // provably warp-uniform copy
  const int quarter = warp & 3;      // TMEM lane quarter
  const int half = warp >> 2;        // which 32 of the 64 channels this warp converts / reads back
  const int row = quarter * 32 + lane;   // pixel of the tile == TMEM lane
  const int RB = RBT;
  const int SWH = geo.SWH, SW = geo.SW;

  if (warp == 0) tmem_alloc<TMEM_COLS>(tmem_ptr);
  if (tid == 0) {
#pragma unroll
    for (int i = 0; i < 6; ++i) mbar_init(&bars[i], 1);
    mbar_fence_init();
    tma_prefetch_desc(&tmap); tma_prefetch_desc(&tmap_du); tma_prefetch_desc(&tmap_zo);
  }
  for (int i = tid; i < 64 * 64; i += NT) {
    const int co = i / 64, ci = i % 64;
    const float w = __ldg(a.w1 + i);
    const uint32_t o1 = sw128_offset(64, co, ci);     // GEMM1: B[n=co][k=ci]
    const uint32_t o2 = sw128_offset(64, ci, co);     // GEMM2: B[n=ci][k=co]
    *reinterpret_cast<uint32_t*>(smem + Off::B1HI + o1) = tf32_hi(w);
    *reinterpret_cast<uint32_t*>(smem + Off::B1LO + o1) = tf32_lo(w);
    *reinterpret_cast<uint32_t*>(smem + Off::B2HI + o2) = tf32_hi(w);
    *reinterpret_cast<uint32_t*>(smem + Off::B2LO + o2) = tf32_lo(w);
  }
  for (int i = tid; i < 9 * 64; i += NT) sW2[i] = __ldg(a.w2 + (i % 64) * 9 + i / 64);
  if (tid < 64) {
    sB1[tid] = __ldg(a.b1 + tid);
    const Coef4 ki = bn_coef_tc(a.bna, tid);
    sCa[tid] = ki.scale; sCa[64 + tid] = ki.shift; sCa[128 + tid] = ki.mean; sCa[192 + tid] = ki.rstd;
    if (MODE == 2) {
      const Coef4 kb = bn_coef_tc(a.bnb, tid);
      sCb[tid] = kb.scale; sCb[64 + tid] = kb.shift; sCb[128 + tid] = kb.mean; sCb[192 + tid] = kb.rstd;
    }
    const Coef4 ko = bn_coef_tc(a.bno, tid);
    sCo[tid] = a.bno.gamma[tid] * ko.rstd;
    sCo[64 + tid] = (float)(a.dsum[tid] * a.bno.inv_count);
    sCo[128 + tid] = (float)(a.dsumzh[tid] * a.bno.inv_count);
    sCo[192 + tid] = ko.mean;
    sCo[256 + tid] = ko.rstd;
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  // the whole TMEM (512 columns) is allocated: its base is 0 (checked), which keeps every MMA
  // operand warp-uniform for the single-thread issue loops below
  constexpr uint32_t tbase = 0;
  bool alive = (*tmem_ptr == 0);
  if (!alive && tid == 0) atomicExch(status, 10);
  const uint32_t lane_addr = tbase + ((uint32_t)(quarter * 32) << 16);
  constexpr uint32_t idesc = make_idesc_tf32(128, 64);
  constexpr uint32_t idesc_dw = make_idesc_tf32(128, 64, 0, 1);     // B MN-major
  const uint64_t dB1hi = make_desc_sw128_kmajor(smem_u32(smem + Off::B1HI));
  const uint64_t dB1lo = make_desc_sw128_kmajor(smem_u32(smem + Off::B1LO));
  const uint64_t dB2hi = make_desc_sw128_kmajor(smem_u32(smem + Off::B2HI));
  const uint64_t dB2lo = make_desc_sw128_kmajor(smem_u32(smem + Off::B2LO));
  const uint64_t dAhi = make_desc_sw128_mnmajor(smem_u32(smem + Off::RAW), 16384, 512, 1);
  const uint64_t dAlo = make_desc_sw128_mnmajor(smem_u32(smem + Off::AL), 16384, 512, 1);

  // ---- persistent accumulators
  // depthwise-backward stage: thread -> (channel pair q2, column group cg of 5 interior columns);
  // the g rows r-1 and r of its 7 columns stay in registers while the strip streams by
  const int q2 = tid & 31, cg = tid >> 5;
  float2 w2r[9], gw2[9];
#pragma unroll
  for (int k = 0; k < 9; ++k) {
    w2r[k] = *reinterpret_cast<const float2*>(sW2 + k * 64 + q2 * 2);
    gw2[k] = make_float2(0.f, 0.f);
  }
  float2 gb2 = make_float2(0.f, 0.f), gb1 = gb2;
  float2 wa[7], wb[7];
#pragma unroll
  for (int d = 0; d < 7; ++d) { wa[d] = make_float2(0.f, 0.f); wb[d] = wa[d]; }
  // byte offsets of the thread's 7 g columns inside a tile row of the TMA layout
  // ([ch block][pixel][128 B], 16-byte chunks ^ (pixel & 7)) without the row term
  int gcolx[7];
#pragma unroll
  for (int d = 0; d < 7; ++d) { const int c = cg * 5 + d; gcolx[d] = c < SWH ? c : SWH - 1; }
  // dW1: thread (TMEM lane `row`, column half) accumulates row (row & 63) of dW1, input channels
  // half*32 .. +31; lanes 64..127 carry the dy_lo part of the same rows
  float gw1[32];
#pragma unroll
  for (int j = 0; j < 32; ++j) gw1[j] = 0.f;
  // g pass: the channel quad of a thread is fixed (256 % 16 == 0): coefficients in registers
  const float4 cgs = *reinterpret_cast<const float4*>(sCo + (tid & 15) * 4);
  const float4 cm1 = *reinterpret_cast<const float4*>(sCo + 64 + (tid & 15) * 4);
  const float4 cmu = *reinterpret_cast<const float4*>(sCo + 192 + (tid & 15) * 4);
  float4 ck;                                          // rstd * mean(du * zhat)
  {
    const float4 m2 = *reinterpret_cast<const float4*>(sCo + 128 + (tid & 15) * 4);
    const float4 rs = *reinterpret_cast<const float4*>(sCo + 256 + (tid & 15) * 4);
    ck = make_float4(rs.x * m2.x, rs.y * m2.y, rs.z * m2.z, rs.w * m2.w);
  }
  // pooled / up-add routing: thread -> (channel quad eq, pixels tid/16 + 16k)
  const int eq = tid & 15;
  float4 sa1 = make_float4(0.f, 0.f, 0.f, 0.f), sa2 = sa1, sb1 = sa1, sb2 = sa1;
  // statistics of du_in: lane L of a warp owns channel half*32 + L
  double s1 = 0.0, s2 = 0.0;

  // ---- this CTA's share of the global block sequence
  const int g0 = (int)(((long long)blockIdx.x * geo.G) / gridDim.x);
  const int g1 = (int)(((long long)(blockIdx.x + 1) * geo.G) / gridDim.x);
  auto decode = [&](int g, int& b_, int& sx_, int& blk_) {
    const int sid = g / geo.NB;
    blk_ = g - sid * geo.NB;
    b_ = sid / geo.nsx;
    sx_ = sid - b_ * geo.nsx;
  };
  // step sequence: [prime(g0 - 1)] g0 ... ; a priming step in front of every block that starts a
  // strip or the CTA's range
  auto first_step = [&]() {
    Step s;
    decode(g0, s.b, s.sx, s.blk);
    s.blk -= 1; s.prime = 1; s.valid = g0 < g1;
    return s;
  };
  auto next_step = [&](const Step& c, int gcur) {     // gcur: real block index the step belongs to
    Step n;
    if (c.prime) { n = c; n.blk += 1; n.prime = 0; return n; }
    const int gn = gcur + 1;
    n.valid = gn < g1;
    decode(gn < geo.G ? gn : 0, n.b, n.sx, n.blk);
    n.prime = (n.blk == 0) ? 1 : 0;
    if (n.prime) n.blk = -1;
    return n;
  };
  auto issue = [&](const CUtensorMap* m, unsigned char* dst, uint64_t* bar, int sx_, int y_, int b_) {
    mbar_arrive_expect_tx(bar, (uint32_t)(RB * SWH) * 256u);
    tma_load_4d(dst, m, bar, 0, sx_ * SW - 1, y_, b_);
    tma_load_4d(dst + 16384, m, bar, 32, sx_ * SW - 1, y_, b_);
  };
  Step st = first_step();
  int gcur = g0;
  if (tid == 0 && st.valid && alive) {
    issue(&tmap_du, sG, &bars[4], st.sx, st.blk * RB + 1, st.b);
    issue(&tmap_zo, sY, &bars[5], st.sx, st.blk * RB + 1, st.b);
  }
  uint32_t it_s = 0, it_r = 0;        // completed steps / real blocks (barrier phases)
  const int hy = row / SWH, hx = row - hy * SWH;       // tile pixel of this thread (TMEM lane)
  while (st.valid && alive) {
    const uint32_t phs = it_s & 1, phr = it_r & 1;
    const int b = st.b, sx = st.sx, blk = st.blk;
    const int y0 = blk * RB, x0 = sx * SW;             // image row / column of the block's first centre pixel
    const Step nx = next_step(st, gcur);
    const long long img_off = (long long)b * a.H * a.W * C64;
    PT_DECL

    // ---- T0: stage the operand (MODE 1/2, real blocks); turn the prefetched du / z_out tiles into g
    const float* za_img = a.za + (long long)b * a.H * a.W * C64 * (MODE == 1 ? 4 : 1);
    if (MODE != 0 && nx.valid && warp == 2) {
      // the operand of the NEXT real block is read with plain loads at its T0: pull its rows into L2 now
      const int nblk = nx.prime ? nx.blk + 1 : nx.blk;
      const int ny0 = nblk * RB, nx0 = nx.sx * SW;
      if (MODE == 1) {
        l2_prefetch_tile<C64>(a.za + (long long)nx.b * a.H * a.W * C64 * 4, a.H * 2, a.W * 2, ny0 * 2,
                              (ny0 + RB) * 2, (nx0 - 1) * 2, (nx0 - 1 + SWH) * 2, lane);
      } else {
        l2_prefetch_tile<C64>(a.za + (long long)nx.b * a.H * a.W * C64, a.H, a.W, ny0, ny0 + RB, nx0 - 1,
                              nx0 - 1 + SWH, lane);
        l2_prefetch_tile<C64>(a.zb + (long long)nx.b * (a.H >> 1) * (a.W >> 1) * C64, a.H >> 1, a.W >> 1,
                              ny0 >> 1, ((ny0 + RB - 1) >> 1) + 1, (nx0 - 1) >> 1, ((nx0 + SWH - 2) >> 1) + 1,
                              lane - 8);
      }
    }
    if (MODE != 0 && !st.prime) {
      // pooled / up-added operand a: vector loads (latency overlaps the waits on du / z_out)
      int py = 0, px = tid >> 4;          // tile pixel (tid >> 4) + 16 k without a division
#pragma unroll 4
      for (int k = 0; k < 128 * 16 / NT; ++k, px += 16) {
        while (px >= SWH) { px -= SWH; ++py; }
        const int pix = (tid >> 4) + 16 * k, ch = tid & 15;
        const int gy = y0 + py, gx = x0 - 1 + px;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (py < RB && gy < a.H && gx >= 0 && gx < a.W) {
          const float4 sc = *reinterpret_cast<const float4*>(sCa + ch * 4);
          const float4 sh = *reinterpret_cast<const float4*>(sCa + 64 + ch * 4);
          if (MODE == 1) {
            const int W2 = a.W * 2;
            const float* p = za_img + ((long long)(gy * 2) * W2 + gx * 2) * C64 + ch * 4;
            const float4 z00 = __ldg(reinterpret_cast<const float4*>(p));
            const float4 z01 = __ldg(reinterpret_cast<const float4*>(p + C64));
            const float4 z10 = __ldg(reinterpret_cast<const float4*>(p + (long long)W2 * C64));
            const float4 z11 = __ldg(reinterpret_cast<const float4*>(p + (long long)W2 * C64 + C64));
            v.x = fmaxf(fmaxf(fmaxf(fmaf(z00.x, sc.x, sh.x), fmaf(z01.x, sc.x, sh.x)), fmaxf(fmaf(z10.x, sc.x, sh.x), fmaf(z11.x, sc.x, sh.x))), 0.f);
            v.y = fmaxf(fmaxf(fmaxf(fmaf(z00.y, sc.y, sh.y), fmaf(z01.y, sc.y, sh.y)), fmaxf(fmaf(z10.y, sc.y, sh.y), fmaf(z11.y, sc.y, sh.y))), 0.f);
            v.z = fmaxf(fmaxf(fmaxf(fmaf(z00.z, sc.z, sh.z), fmaf(z01.z, sc.z, sh.z)), fmaxf(fmaf(z10.z, sc.z, sh.z), fmaf(z11.z, sc.z, sh.z))), 0.f);
            v.w = fmaxf(fmaxf(fmaxf(fmaf(z00.w, sc.w, sh.w), fmaf(z01.w, sc.w, sh.w)), fmaxf(fmaf(z10.w, sc.w, sh.w), fmaf(z11.w, sc.w, sh.w))), 0.f);
          } else {
            const float4 z = __ldg(reinterpret_cast<const float4*>(za_img + ((long long)gy * a.W + gx) * C64 + ch * 4));
            const int Hb = a.H >> 1, Wb = a.W >> 1;
            const float4 zb = __ldg(reinterpret_cast<const float4*>(
                a.zb + (((long long)b * Hb + (gy >> 1)) * Wb + (gx >> 1)) * C64 + ch * 4));
            const float4 scb = *reinterpret_cast<const float4*>(sCb + ch * 4);
            const float4 shb = *reinterpret_cast<const float4*>(sCb + 64 + ch * 4);
            v.x = fmaxf(fmaf(z.x, sc.x, sh.x), 0.f) + fmaxf(fmaf(zb.x, scb.x, shb.x), 0.f);
            v.y = fmaxf(fmaf(z.y, sc.y, sh.y), 0.f) + fmaxf(fmaf(zb.y, scb.y, shb.y), 0.f);
            v.z = fmaxf(fmaf(z.z, sc.z, sh.z), 0.f) + fmaxf(fmaf(zb.z, scb.z, shb.z), 0.f);
            v.w = fmaxf(fmaf(z.w, sc.w, sh.w), 0.f) + fmaxf(fmaf(zb.w, scb.w, shb.w), 0.f);
          }
        }
        *reinterpret_cast<float4*>(raw + zoff(pix, ch)) = v;
      }
    }
    if (!mbar_wait(&bars[4], phs)) { alive = false; if (lane == 0) atomicExch(status, 14); }
    if (alive && !mbar_wait(&bars[5], phs)) { alive = false; if (lane == 0) atomicExch(status, 15); }
    PT(0);
    if (alive) {
      // g = gamma*rstd*(du - mean(du) - zhat*mean(du*zhat)) in place; the g tile holds image rows
      // y0+1 .. y0+RB (one row below the block's own rows), exact 0 outside the image
      int py = 0, px = tid >> 4;          // tile pixel (tid >> 4) + 16 k without a division
#pragma unroll 4
      for (int k = 0; k < 128 * 16 / NT; ++k, px += 16) {
        while (px >= SWH) { px -= SWH; ++py; }
        const int pix = (tid >> 4) + 16 * k, ch = tid & 15;
        const int gy = y0 + 1 + py, gx = x0 - 1 + px;
        float* gp = const_cast<float*>(rchunk(sG, pix, ch));
        float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
        if (py < RB && gy >= 0 && gy < a.H && gx >= 0 && gx < a.W) {
          const float4 d = *reinterpret_cast<const float4*>(gp);
          const float4 z = *reinterpret_cast<const float4*>(rchunk(sY, pix, ch));
          g.x = cgs.x * (d.x - cm1.x - (z.x - cmu.x) * ck.x);
          g.y = cgs.y * (d.y - cm1.y - (z.y - cmu.y) * ck.y);
          g.z = cgs.z * (d.z - cm1.z - (z.z - cmu.z) * ck.z);
          g.w = cgs.w * (d.w - cm1.w - (z.w - cmu.w) * ck.w);
        }
        *reinterpret_cast<float4*>(gp) = g;
      }
    }
    PT(1);
    if (st.prime) {
      // ---- priming step: only roll the g window over the rows of this block
      __syncthreads();
      if (alive) {
#pragma unroll
        for (int ii = 0; ii < RBT; ++ii) {
#pragma unroll
          for (int d = 0; d < 7; ++d) {
            const int p = ii * SWH + gcolx[d];
            wa[d] = wb[d];
            wb[d] = *reinterpret_cast<const float2*>(sG + (q2 >> 4) * 16384 + p * 128 + ((((q2 >> 1) & 7) ^ (p & 7)) << 4) + (q2 & 1) * 8);
          }
        }
      }
      fence_proxy_async_smem();
      __syncthreads();
      if (tid == 0 && nx.valid && alive) {
        issue(&tmap_du, sG, &bars[4], nx.sx, nx.blk * RB + 1, nx.b);
        issue(&tmap_zo, sY, &bars[5], nx.sx, nx.blk * RB + 1, nx.b);
        if (MODE == 0) issue(&tmap, raw, &bars[0], nx.sx, nx.blk * RB, nx.b);   // the real block that follows
      }
      alive = __syncthreads_and(alive ? 1 : 0) != 0;
      ++it_s;
      st = nx;
      continue;
    }
    if (MODE == 0) {
      if (alive && !mbar_wait(&bars[0], phr)) { alive = false; if (lane == 0) atomicExch(status, 11); }
    } else {
      __syncthreads();      // operand a staged by all threads
    }
    PT(2);

    // ---- T1: a = relu(bn(z_in)) row per thread (32 channels per warp half) -> hi/lo -> TMEM (A of
    // MMA 1) and -> shared (hi in place, lo beside it: B of MMA 3); z_in itself -> TMEM
    const int gy_r = y0 + hy, gx_r = x0 - 1 + hx;
    const bool interior = hy < RB && hx >= 1 && hx <= SW && gy_r < a.H && gx_r < a.W;
    auto issue_mma1 = [&](int part) {
      if (warp_u == 0) {
        if (elect_one_b()) {
          tc_fence_after();
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) {
            const int k = (kk >> 1) * 4 + part * 2 + (kk & 1);
            const uint32_t koff = ((k >> 2) * 8192 + (k & 3) * 32) >> 4;
            mma_tf32_ts(tbase + COL_D1, tbase + COL_ALO + k * 8, dB1hi + koff, idesc, (part | kk) != 0);
            mma_tf32_ts(tbase + COL_D1, tbase + COL_AHI + k * 8, dB1lo + koff, idesc, 1);
            mma_tf32_ts(tbase + COL_D1, tbase + COL_AHI + k * 8, dB1hi + koff, idesc, 1);
          }
          if (part == 1) mma_commit(&bars[1]);
        }
        __syncwarp();
      }
    };
#pragma unroll
    for (int g16 = 0; g16 < 2; ++g16) {
      if (alive) {
        uint32_t hi[16], lo[16], zr[16];
        float4 zin[4];       // all four loads before the in-place stores (which may alias for the compiler)
#pragma unroll
        for (int c4 = 0; c4 < 4; ++c4)
          zin[c4] = *reinterpret_cast<const float4*>(raw + zoff(row, half * 8 + g16 * 4 + c4));
#pragma unroll
        for (int c4 = 0; c4 < 4; ++c4) {
          const int ch = half * 8 + g16 * 4 + c4;
          const uint32_t zo = zoff(row, ch);
          float4 z = zin[c4];
          // tile pixels beyond the block (lanes >= RB*SWH) hold whatever was in shared memory: they
          // must enter the pixel-contraction of MMA 3 as exact zeros (0 * Inf = NaN otherwise)
          if (hy >= RB) z = make_float4(0.f, 0.f, 0.f, 0.f);
          float v[4] = {z.x, z.y, z.z, z.w};
          if (hy >= RB) { v[0] = v[1] = v[2] = v[3] = 0.f; }
          else if (MODE == 0) {
            const float4 sc = *reinterpret_cast<const float4*>(sCa + ch * 4);
            const float4 sh = *reinterpret_cast<const float4*>(sCa + 64 + ch * 4);
            zr[c4 * 4 + 0] = __float_as_uint(z.x); zr[c4 * 4 + 1] = __float_as_uint(z.y);
            zr[c4 * 4 + 2] = __float_as_uint(z.z); zr[c4 * 4 + 3] = __float_as_uint(z.w);
            v[0] = fmaxf(fmaf(z.x, sc.x, sh.x), 0.f); v[1] = fmaxf(fmaf(z.y, sc.y, sh.y), 0.f);
            v[2] = fmaxf(fmaf(z.z, sc.z, sh.z), 0.f); v[3] = fmaxf(fmaf(z.w, sc.w, sh.w), 0.f);
          }
#pragma unroll
          for (int j = 0; j < 4; ++j) { hi[c4 * 4 + j] = tf32_hi(v[j]); lo[c4 * 4 + j] = tf32_lo(v[j]); }
          *reinterpret_cast<uint4*>(raw + zo) = make_uint4(hi[c4 * 4], hi[c4 * 4 + 1], hi[c4 * 4 + 2], hi[c4 * 4 + 3]);
          *reinterpret_cast<uint4*>(sAL + zo) = make_uint4(lo[c4 * 4], lo[c4 * 4 + 1], lo[c4 * 4 + 2], lo[c4 * 4 + 3]);
        }
        tmem_st16(lane_addr + COL_AHI + half * 32 + g16 * 16, hi);
        tmem_st16(lane_addr + COL_ALO + half * 32 + g16 * 16, lo);
        if (MODE == 0) tmem_st16(lane_addr + COL_Z + half * 32 + g16 * 16, zr);
        tmem_wait_st();
      }
      if (g16 == 1) fence_proxy_async_smem();   // a_hi / a_lo (generic writes) are read by MMA 3
      tc_fence_before();
      alive = __syncthreads_and(alive ? 1 : 0) != 0;
      if (alive) issue_mma1(g16);       // ---- T2: MMA 1   D1 = a W1^T
    }
    PT(3);
    if (alive && !mbar_wait(&bars[1], phr)) { alive = false; if (lane == 0) atomicExch(status, 12); }
    tc_fence_after();
    PT(4);
    // ---- T3: y (+bias) for the block's own in-image pixels, exact 0 elsewhere -> sY (z_out was
    // consumed by the g pass)
    if (alive) {
#pragma unroll
      for (int g16 = 0; g16 < 2; ++g16) {
        uint32_t v[16];
        tmem_ld16(lane_addr + COL_D1 + half * 32 + g16 * 16, v);
        tmem_wait_ld();
#pragma unroll
        for (int c4 = 0; c4 < 4; ++c4) {
          const int ch = half * 8 + g16 * 4 + c4;
          const float4 bb = *reinterpret_cast<const float4*>(sB1 + ch * 4);
          float4 o;
          o.x = interior ? __uint_as_float(v[c4 * 4 + 0]) + bb.x : 0.f;
          o.y = interior ? __uint_as_float(v[c4 * 4 + 1]) + bb.y : 0.f;
          o.z = interior ? __uint_as_float(v[c4 * 4 + 2]) + bb.z : 0.f;
          o.w = interior ? __uint_as_float(v[c4 * 4 + 3]) + bb.w : 0.f;
          *reinterpret_cast<float4*>(tchunk(sY, row, ch)) = o;
        }
      }
    }
    tc_fence_before();
    __syncthreads();
    PT(5);

    // ---- T4: depthwise backward: dy in place over y, dW2, db2, db1.  New g row r+1 (tile row ii)
    // with the window rows r-1, r gives dy of the block's own row r = y0 + ii.
    if (alive) {
      const bool cgact = cg * 5 < SW;
#pragma unroll
      for (int ii = 0; ii < RBT; ++ii) {
        float2 nc[7];
#pragma unroll
        for (int d = 0; d < 7; ++d) {
          const int p = ii * SWH + gcolx[d];
          nc[d] = *reinterpret_cast<const float2*>(sG + (q2 >> 4) * 16384 + p * 128 + ((((q2 >> 1) & 7) ^ (p & 7)) << 4) + (q2 & 1) * 8);
        }
        if (cgact) {
          const bool rin = (y0 + ii) < a.H;
#pragma unroll
          for (int e = 0; e < 5; ++e) {
            const int col = cg * 5 + 1 + e;                       // tile column of the centre pixel
            const int pc = ii * SWH + col;
            float2* yp = reinterpret_cast<float2*>(reinterpret_cast<unsigned char*>(tchunk(sY, pc, q2 >> 1)) + (q2 & 1) * 8);
            if (col <= SW) {
              const float2 y = *yp;
              float dyx = 0.f, dyy = 0.f;
#pragma unroll
              for (int kx = 0; kx < 3; ++kx) {
                const float2 g0_ = nc[e + 2 - kx], g1_ = wb[e + 2 - kx], g2_ = wa[e + 2 - kx];
                fma2(dyx, dyy, w2r[kx].x, w2r[kx].y, g0_.x, g0_.y);
                fma2(dyx, dyy, w2r[3 + kx].x, w2r[3 + kx].y, g1_.x, g1_.y);
                fma2(dyx, dyy, w2r[6 + kx].x, w2r[6 + kx].y, g2_.x, g2_.y);
                fma2(gw2[kx].x, gw2[kx].y, y.x, y.y, g0_.x, g0_.y);
                fma2(gw2[3 + kx].x, gw2[3 + kx].y, y.x, y.y, g1_.x, g1_.y);
                fma2(gw2[6 + kx].x, gw2[6 + kx].y, y.x, y.y, g2_.x, g2_.y);
              }
              gb2.x += wb[e + 1].x; gb2.y += wb[e + 1].y;
              const bool in = rin && (x0 - 1 + col) < a.W;
              if (!in) { dyx = 0.f; dyy = 0.f; }
              gb1.x += dyx; gb1.y += dyy;
              *yp = make_float2(dyx, dyy);
            }
          }
        }
#pragma unroll
        for (int d = 0; d < 7; ++d) { wa[d] = wb[d]; wb[d] = nc[d]; }
      }
    }
    fence_proxy_async_smem();      // generic accesses to the g buffer precede its TMA refill
    __syncthreads();
    if (tid == 0 && nx.valid && alive) issue(&tmap_du, sG, &bars[4], nx.sx, nx.blk * RB + 1, nx.b);
    PT(6);

    // ---- T5 / T6: dy rows -> hi/lo -> TMEM (A columns are free: MMA 1 completed), in two halves with
    // the MMAs of  D2 = dy W1  issued behind each; the dy^T staging for MMA 3 runs under the first batch
    auto issue_mma2 = [&](int part, bool commit) {
      if (warp_u == 0) {
        if (elect_one_b()) {
          tc_fence_after();
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) {
            const int k = (kk >> 1) * 4 + part * 2 + (kk & 1);
            const uint32_t koff = ((k >> 2) * 8192 + (k & 3) * 32) >> 4;
            mma_tf32_ts(tbase + COL_D2, tbase + COL_ALO + k * 8, dB2hi + koff, idesc, (part | kk) != 0);
            mma_tf32_ts(tbase + COL_D2, tbase + COL_AHI + k * 8, dB2lo + koff, idesc, 1);
            mma_tf32_ts(tbase + COL_D2, tbase + COL_AHI + k * 8, dB2hi + koff, idesc, 1);
          }
          if (commit) mma_commit(&bars[2]);
        }
        __syncwarp();
      }
    };
    auto convert_dy = [&](int g16) {
      uint32_t hi[16], lo[16];
#pragma unroll
      for (int c4 = 0; c4 < 4; ++c4) {
        const int ch = half * 8 + g16 * 4 + c4;
        const float4 v = *reinterpret_cast<const float4*>(tchunk(sY, row, ch));
        hi[c4 * 4 + 0] = tf32_hi(v.x); lo[c4 * 4 + 0] = tf32_lo(v.x);
        hi[c4 * 4 + 1] = tf32_hi(v.y); lo[c4 * 4 + 1] = tf32_lo(v.y);
        hi[c4 * 4 + 2] = tf32_hi(v.z); lo[c4 * 4 + 2] = tf32_lo(v.z);
        hi[c4 * 4 + 3] = tf32_hi(v.w); lo[c4 * 4 + 3] = tf32_lo(v.w);
      }
      tmem_st16(lane_addr + COL_AHI + half * 32 + g16 * 16, hi);
      tmem_st16(lane_addr + COL_ALO + half * 32 + g16 * 16, lo);
    };
    if (alive) { convert_dy(0); tmem_wait_st(); }
    tc_fence_before();
    alive = __syncthreads_and(alive ? 1 : 0) != 0;
    if (alive) issue_mma2(0, false);
    if (alive) { convert_dy(1); tmem_wait_st(); }
    tc_fence_before();
    alive = __syncthreads_and(alive ? 1 : 0) != 0;
    if (alive) issue_mma2(1, true);
    PT(7);
    // dy^T for MMA 3 (staged while the second MMA 2 batch runs): this thread's TMEM lane is output
    // channel (row & 63), hi part on lanes 0..63 and lo part on lanes 64..127; its warp half covers
    // 64 of the 128 pixel columns.  A warp reads 32 consecutive channels of one pixel per load:
    // conflict-free.
    if (alive) {
      const int m = row & 63;
      const bool islo = warp_uniform((uint32_t)(quarter >= 2)) != 0;
      const unsigned char* ybase = sY + half * 64 * 256 + (m & 3) * 4;
      const int c16 = m >> 2;
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        uint32_t v[16];
#pragma unroll
        for (int j = 0; j < 16; ++j)
          v[j] = __float_as_uint(*reinterpret_cast<const float*>(ybase + (g * 16 + j) * 256 + ((c16 ^ (j & 7)) << 4)));
        if (islo) {            // warp-uniform (lane quarters 2, 3)
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] = tf32_lo(__uint_as_float(v[j]));
        } else {
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] &= 0xFFFFE000u;
        }
        tmem_st16(lane_addr + COL_DYT + half * 64 + g * 16, v);
      }
      tmem_wait_st();
    }
    fence_proxy_async_smem();      // y / dy (generic accesses) precede the TMA refill of that buffer
    tc_fence_before();
    alive = __syncthreads_and(alive ? 1 : 0) != 0;
    if (alive && warp_u == 0) {
      if (elect_one_b()) {
        tc_fence_after();
        // ---- MMA 3: D_dw[128 x 64] = dy^T(stacked hi | lo, TMEM) x a (MN-major smem: a_hi, then a_lo);
        // K = 128 pixels in 16 steps of 8 rows (1024 B); fresh accumulator every block (the running
        // sum is kept in registers with round-to-nearest adds)
#pragma unroll
        for (int pass = 0; pass < 2; ++pass)
#pragma unroll
          for (int k = 0; k < 16; ++k)
            mma_tf32_ts(tbase + COL_DW, tbase + COL_DYT + k * 8,
                        (pass == 0 ? dAhi : dAlo) + (uint32_t)(k * 1024 >> 4), idesc_dw, (pass | k) != 0);
        mma_commit(&bars[3]);
      }
      __syncwarp();
    }
    // every warp is done reading y / dy: refill that buffer with the next step's z_out
    if (tid == 32 && nx.valid && alive) issue(&tmap_zo, sY, &bars[5], nx.sx, nx.blk * RB + 1, nx.b);
    if (alive && !mbar_wait(&bars[2], phr)) { alive = false; if (lane == 0) atomicExch(status, 13); }
    tc_fence_after();
    PT(8);
    // dW1 of this block -> registers (MODE 0 does it after the epilogue, under which MMA 3 runs)
    auto collect_dw1 = [&]() {
      if (alive && !mbar_wait(&bars[3], phr)) { alive = false; if (lane == 0) atomicExch(status, 16); }
      tc_fence_after();
      if (alive) {
#pragma unroll
        for (int g16 = 0; g16 < 2; ++g16) {
          uint32_t dv[16];
          tmem_ld16(lane_addr + COL_DW + half * 32 + g16 * 16, dv);
          tmem_wait_ld();
#pragma unroll
          for (int j = 0; j < 16; ++j) gw1[g16 * 16 + j] += __uint_as_float(dv[j]);
        }
      }
    };
    if (MODE != 0) collect_dw1();      // the routing pass stages h in the a_lo buffer
    if (MODE != 0) {
      // ---- T7': h rows -> shared, then route through the max-pool winner / the up-add children
      if (alive) {
#pragma unroll
        for (int g16 = 0; g16 < 2; ++g16) {
          uint32_t hv[16];
          tmem_ld16(lane_addr + COL_D2 + half * 32 + g16 * 16, hv);
          tmem_wait_ld();
#pragma unroll
          for (int c4 = 0; c4 < 4; ++c4)
            *reinterpret_cast<float4*>(tchunk(sH, row, half * 8 + g16 * 4 + c4)) =
                make_float4(__uint_as_float(hv[c4 * 4]), __uint_as_float(hv[c4 * 4 + 1]),
                            __uint_as_float(hv[c4 * 4 + 2]), __uint_as_float(hv[c4 * 4 + 3]));
        }
      }
      __syncthreads();
      if (alive) {
        const float4 sc = *reinterpret_cast<const float4*>(sCa + eq * 4);
        const float4 sh = *reinterpret_cast<const float4*>(sCa + 64 + eq * 4);
        const float4 mu = *reinterpret_cast<const float4*>(sCa + 128 + eq * 4);
        const float4 rs = *reinterpret_cast<const float4*>(sCa + 192 + eq * 4);
        for (int idx = tid >> 4; idx < RB * SW; idx += NT / 16) {
          const int r = idx / SW, x = idx - r * SW;
          const int gy = y0 + r, gx = x0 + x;
          if (gy >= a.H || gx >= a.W) continue;
          const float4 h = *reinterpret_cast<const float4*>(tchunk(sH, r * SWH + x + 1, eq));
          const float hh[4] = {h.x, h.y, h.z, h.w};
          if (MODE == 2) {
            const long long off = ((long long)gy * a.W + gx) * C64 + eq * 4;
            const float4 z = __ldg(reinterpret_cast<const float4*>(za_img + off));
            float4 d;
            d.x = fmaf(z.x, sc.x, sh.x) > 0.f ? h.x : 0.f; d.y = fmaf(z.y, sc.y, sh.y) > 0.f ? h.y : 0.f;
            d.z = fmaf(z.z, sc.z, sh.z) > 0.f ? h.z : 0.f; d.w = fmaf(z.w, sc.w, sh.w) > 0.f ? h.w : 0.f;
            sa1.x += d.x; sa1.y += d.y; sa1.z += d.z; sa1.w += d.w;
            sa2.x = fmaf(d.x, (z.x - mu.x) * rs.x, sa2.x); sa2.y = fmaf(d.y, (z.y - mu.y) * rs.y, sa2.y);
            sa2.z = fmaf(d.z, (z.z - mu.z) * rs.z, sa2.z); sa2.w = fmaf(d.w, (z.w - mu.w) * rs.w, sa2.w);
            float4* p = reinterpret_cast<float4*>(a.dua + img_off + off);
            if (a.acc_a) { const float4 o = *p; d.x += o.x; d.y += o.y; d.z += o.z; d.w += o.w; }
            *p = d;
          } else {
            // first maximum of the 2x2 window (ATen order), then the ReLU mask
            const int W2 = a.W * 2;
            const long long o00 = ((long long)(gy * 2) * W2 + gx * 2) * C64 + eq * 4;
            const long long offs[4] = {o00, o00 + C64, o00 + (long long)W2 * C64, o00 + (long long)W2 * C64 + C64};
            float zz[4][4], vv[4][4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const float4 z = __ldg(reinterpret_cast<const float4*>(za_img + offs[j]));
              zz[j][0] = z.x; zz[j][1] = z.y; zz[j][2] = z.z; zz[j][3] = z.w;
              vv[j][0] = fmaxf(fmaf(z.x, sc.x, sh.x), 0.f); vv[j][1] = fmaxf(fmaf(z.y, sc.y, sh.y), 0.f);
              vv[j][2] = fmaxf(fmaf(z.z, sc.z, sh.z), 0.f); vv[j][3] = fmaxf(fmaf(z.w, sc.w, sh.w), 0.f);
            }
            const float muv[4] = {mu.x, mu.y, mu.z, mu.w}, rsv[4] = {rs.x, rs.y, rs.z, rs.w};
            float dd[4][4], s1v[4], s2v[4];
#pragma unroll
            for (int c = 0; c < 4; ++c) {
              float best = vv[0][c]; int bj = 0;
#pragma unroll
              for (int j = 1; j < 4; ++j) if (vv[j][c] > best) { best = vv[j][c]; bj = j; }
              const float hv = best > 0.f ? hh[c] : 0.f;
              float zb = zz[0][c];
#pragma unroll
              for (int j = 1; j < 4; ++j) if (j == bj) zb = zz[j][c];
              s1v[c] = hv;
              s2v[c] = hv * ((zb - muv[c]) * rsv[c]);
#pragma unroll
              for (int j = 0; j < 4; ++j) dd[j][c] = (j == bj) ? hv : 0.f;
            }
            sa1.x += s1v[0]; sa1.y += s1v[1]; sa1.z += s1v[2]; sa1.w += s1v[3];
            sa2.x += s2v[0]; sa2.y += s2v[1]; sa2.z += s2v[2]; sa2.w += s2v[3];
            float* dbase = a.dua + (long long)b * a.H * a.W * C64 * 4;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              float4 o = make_float4(dd[j][0], dd[j][1], dd[j][2], dd[j][3]);
              float4* p = reinterpret_cast<float4*>(dbase + offs[j]);
              if (a.acc_a) { const float4 q = *p; o.x += q.x; o.y += q.y; o.z += q.z; o.w += q.w; }
              *p = o;
            }
          }
        }
        if (MODE == 2) {
          // up-sampled operand: a low-res pixel gathers its 2x2 children (RB, SW and the strip origin
          // are even in this mode)
          const float4 scb = *reinterpret_cast<const float4*>(sCb + eq * 4);
          const float4 shb = *reinterpret_cast<const float4*>(sCb + 64 + eq * 4);
          const float4 mub = *reinterpret_cast<const float4*>(sCb + 128 + eq * 4);
          const float4 rsb = *reinterpret_cast<const float4*>(sCb + 192 + eq * 4);
          const int Hb = a.H >> 1, Wb = a.W >> 1;
          for (int idx = tid >> 4; idx < (RB / 2) * (SW / 2); idx += NT / 16) {
            const int ly = idx / (SW / 2), lx = idx - ly * (SW / 2);
            const int gy = (y0 >> 1) + ly, gx = (x0 >> 1) + lx;
            if (gy >= Hb || gx >= Wb) continue;
            const int p00 = (2 * ly) * SWH + 2 * lx + 1;
            const float4 h0 = *reinterpret_cast<const float4*>(tchunk(sH, p00, eq));
            const float4 h1 = *reinterpret_cast<const float4*>(tchunk(sH, p00 + 1, eq));
            const float4 h2 = *reinterpret_cast<const float4*>(tchunk(sH, p00 + SWH, eq));
            const float4 h3 = *reinterpret_cast<const float4*>(tchunk(sH, p00 + SWH + 1, eq));
            const float4 hs = make_float4(h0.x + h1.x + h2.x + h3.x, h0.y + h1.y + h2.y + h3.y,
                                          h0.z + h1.z + h2.z + h3.z, h0.w + h1.w + h2.w + h3.w);
            const long long off = (((long long)b * Hb + gy) * Wb + gx) * C64 + eq * 4;
            const float4 z = __ldg(reinterpret_cast<const float4*>(a.zb + off));
            float4 d;
            d.x = fmaf(z.x, scb.x, shb.x) > 0.f ? hs.x : 0.f; d.y = fmaf(z.y, scb.y, shb.y) > 0.f ? hs.y : 0.f;
            d.z = fmaf(z.z, scb.z, shb.z) > 0.f ? hs.z : 0.f; d.w = fmaf(z.w, scb.w, shb.w) > 0.f ? hs.w : 0.f;
            sb1.x += d.x; sb1.y += d.y; sb1.z += d.z; sb1.w += d.w;
            sb2.x = fmaf(d.x, (z.x - mub.x) * rsb.x, sb2.x); sb2.y = fmaf(d.y, (z.y - mub.y) * rsb.y, sb2.y);
            sb2.z = fmaf(d.z, (z.z - mub.z) * rsb.z, sb2.z); sb2.w = fmaf(d.w, (z.w - mub.w) * rsb.w, sb2.w);
            float4* p = reinterpret_cast<float4*>(a.dub + off);
            if (a.acc_b) { const float4 o = *p; d.x += o.x; d.y += o.y; d.z += o.z; d.w += o.w; }
            *p = d;
          }
        }
      }
    } else
    // ---- T7: epilogue: du_in = h * [u_in > 0], statistics (dy of out-of-image pixels was zeroed,
    // so h is zero there; only the block's own in-image pixels write)
    {
      float v1[32], v2[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) { v1[j] = 0.f; v2[j] = 0.f; }
      if (alive) {
#pragma unroll
        for (int g16 = 0; g16 < 2; ++g16) {
          uint32_t hv[16], zv[16];
          tmem_ld16(lane_addr + COL_D2 + half * 32 + g16 * 16, hv);
          tmem_ld16(lane_addr + COL_Z + half * 32 + g16 * 16, zv);
          tmem_wait_ld();
          if (interior) {
            float* dst = a.dua + img_off + ((long long)gy_r * a.W + gx_r) * C64;
#pragma unroll
            for (int c4 = 0; c4 < 4; ++c4) {
              const int ch = half * 8 + g16 * 4 + c4;
              const float4 z = make_float4(__uint_as_float(zv[c4 * 4]), __uint_as_float(zv[c4 * 4 + 1]),
                                           __uint_as_float(zv[c4 * 4 + 2]), __uint_as_float(zv[c4 * 4 + 3]));
              const float4 sc = *reinterpret_cast<const float4*>(sCa + ch * 4);
              const float4 sh = *reinterpret_cast<const float4*>(sCa + 64 + ch * 4);
              const float4 mu = *reinterpret_cast<const float4*>(sCa + 128 + ch * 4);
              const float4 rs = *reinterpret_cast<const float4*>(sCa + 192 + ch * 4);
              float4 d;
              d.x = fmaf(z.x, sc.x, sh.x) > 0.f ? __uint_as_float(hv[c4 * 4 + 0]) : 0.f;
              d.y = fmaf(z.y, sc.y, sh.y) > 0.f ? __uint_as_float(hv[c4 * 4 + 1]) : 0.f;
              d.z = fmaf(z.z, sc.z, sh.z) > 0.f ? __uint_as_float(hv[c4 * 4 + 2]) : 0.f;
              d.w = fmaf(z.w, sc.w, sh.w) > 0.f ? __uint_as_float(hv[c4 * 4 + 3]) : 0.f;
              const int j = g16 * 16 + c4 * 4;
              v1[j] = d.x; v1[j + 1] = d.y; v1[j + 2] = d.z; v1[j + 3] = d.w;
              v2[j] = d.x * ((z.x - mu.x) * rs.x); v2[j + 1] = d.y * ((z.y - mu.y) * rs.y);
              v2[j + 2] = d.z * ((z.z - mu.z) * rs.z); v2[j + 3] = d.w * ((z.w - mu.w) * rs.w);
            }
            // 256-bit accesses: a lane owns a whole pixel row, so every access touches 32 distinct
            // 128-byte lines; two chunks (one full 32-byte sector) per access halve the line visits
#pragma unroll
            for (int c8 = 0; c8 < 2; ++c8) {
              float* p = dst + (half * 8 + g16 * 4 + c8 * 2) * 4;
              const int j = g16 * 16 + c8 * 8;
              float o[8];
#pragma unroll
              for (int q = 0; q < 8; ++q) o[q] = v1[j + q];
              if (a.acc_a) {
                float r[8];
                ldg_v8(p, r);
#pragma unroll
                for (int q = 0; q < 8; ++q) o[q] += r[q];
              }
              stg_v8(p, o);
            }
          }
        }
      }
      // butterfly transpose-reduce over the warp's 32 pixels: lane L ends with channel half*32+L
#pragma unroll
      for (int s = 16; s >= 1; s >>= 1) {
        const bool upper = (lane & s) != 0;
#pragma unroll
        for (int i = 0; i < s; ++i) {
          const float snd1 = upper ? v1[i] : v1[i + s];
          const float kp1 = upper ? v1[i + s] : v1[i];
          v1[i] = kp1 + __shfl_xor_sync(0xffffffffu, snd1, s);
          const float snd2 = upper ? v2[i] : v2[i + s];
          const float kp2 = upper ? v2[i + s] : v2[i];
          v2[i] = kp2 + __shfl_xor_sync(0xffffffffu, snd2, s);
        }
      }
      s1 += (double)v1[0];
      s2 += (double)v2[0];
      PT(9);
      collect_dw1();
      PT(10);
    }
    // MMA 3 has completed (every thread waited on it): the a_hi / a_lo buffers are free; z_in of the
    // next real block lands behind its g pass (a priming step in between issues it itself)
    fence_proxy_async_smem();
    tc_fence_before();
    alive = __syncthreads_and(alive ? 1 : 0) != 0;
    if (MODE == 0 && tid == 0 && nx.valid && !nx.prime && alive) issue(&tmap, raw, &bars[0], nx.sx, nx.blk * RB, nx.b);
    ++it_s; ++it_r;
    st = nx;
    ++gcur;
    PT(11);
#ifdef YUNET_PHASE_TIMING
    if (pt_on) atomicAdd(status + 32 + 15, 1);
#endif
  }

  // ---- flush: parameter gradients -> shared-memory reduction -> one partial vector per CTA
  {
    constexpr int NW1 = 64 * 64, NP = NW1 + 11 * 64;
    float* sRed = reinterpret_cast<float*>(raw);       // tile buffers are free (no TMA in flight)
    __syncthreads();
    for (int i = tid; i < NP; i += NT) sRed[i] = 0.f;
    __syncthreads();
    // TMEM lane `row` holds row (row & 63) of dW1 (hi part on lanes < 64, lo part above)
    if (row < 64) {
#pragma unroll
      for (int j = 0; j < 32; ++j) sRed[row * 64 + half * 32 + j] = gw1[j];
    }
    __syncthreads();
    if (row >= 64) {
#pragma unroll
      for (int j = 0; j < 32; ++j) sRed[(row - 64) * 64 + half * 32 + j] += gw1[j];
    }
    // depthwise weight / bias gradients: thread (channel pair q2, column group cg): staged per column
    // group and summed in a fixed order (bit-reproducible, no shared-memory atomics)
    float* sStage = reinterpret_cast<float*>(sAL);      // [8 column groups][11][64 channels]
#pragma unroll
    for (int k = 0; k < 9; ++k) {
      sStage[(cg * 11 + k) * 64 + q2 * 2] = gw2[k].x;
      sStage[(cg * 11 + k) * 64 + q2 * 2 + 1] = gw2[k].y;
    }
    sStage[(cg * 11 + 9) * 64 + q2 * 2] = gb2.x; sStage[(cg * 11 + 9) * 64 + q2 * 2 + 1] = gb2.y;
    sStage[(cg * 11 + 10) * 64 + q2 * 2] = gb1.x; sStage[(cg * 11 + 10) * 64 + q2 * 2 + 1] = gb1.y;
    __syncthreads();
    for (int i = tid; i < 11 * 64; i += NT) {
      const int v = i / 64, ch = i % 64;
      float t = 0.f;
#pragma unroll
      for (int c = 0; c < 8; ++c) t += sStage[(c * 11 + v) * 64 + ch];
      if (v < 9) sRed[NW1 + 64 + ch * 9 + v] = t;
      else if (v == 9) sRed[NW1 + 640 + ch] = t;
      else sRed[NW1 + ch] = t;
    }
    __syncthreads();
    float* dst = a.partial + (long long)blockIdx.x * kPartialStride;
    for (int i = tid; i < NP; i += NT) dst[i] = sRed[i];
  }
  if (MODE == 0) {
    atomicAdd(a.dsum_a + half * 32 + lane, s1);
    atomicAdd(a.dsumzh_a + half * 32 + lane, s2);
  } else {
    float vals[16] = {sa1.x, sa1.y, sa1.z, sa1.w, sa2.x, sa2.y, sa2.z, sa2.w,
                      sb1.x, sb1.y, sb1.z, sb1.w, sb2.x, sb2.y, sb2.z, sb2.w};
#pragma unroll
    for (int v = 0; v < 16; ++v) vals[v] += __shfl_xor_sync(0xffffffffu, vals[v], 16);
    if (lane < 16) {
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        atomicAdd(a.dsum_a + eq * 4 + c, (double)vals[c]);
        atomicAdd(a.dsumzh_a + eq * 4 + c, (double)vals[4 + c]);
        if (MODE == 2) {
          atomicAdd(a.dsum_b + eq * 4 + c, (double)vals[8 + c]);
          atomicAdd(a.dsumzh_b + eq * 4 + c, (double)vals[12 + c]);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<TMEM_COLS>(tbase);
}

}  // namespace

int unit_bwd_st_supported(int cin, int cout, int mode, int has_bn, int H, int W) {
  if (!(cin == 64 && cout == 64 && mode >= 0 && mode <= 2 && has_bn && tma_encode_fn() != nullptr)) return 0;
  const int nsx = (W + 39) / 40, SW = (W + nsx - 1) / nsx;
  if (mode == 2 && ((SW & 1) || (H & 1))) return 0;     // up-add children pair up inside a block
  return 1;
}

// Measured on B200 (bs 256, profiles/r2_*): the strips win where a strip is long (H >= 40: 0.87 vs
// 1.07 ms at 80x80, 0.255 vs 0.296 ms at 40x40); short strips pay one priming step per 4..7 blocks and
// the up-add units run 2-row blocks (84 of 128 lanes), where the halo tiles of unit_bwd_tc.cu are as good.
int unit_bwd_st_preferred(int mode, int H, int W) {
  (void)W;
  if (mode == 2) return 0;
  if (mode == 1) return H >= 20;
  return H >= 40;
}

cudaError_t launch_unit_bwd_st(int mode, const UnitBwdArgs& a, int num_sms, int* status, cudaStream_t s) {
  if (a.dout_batch_stride != (long long)a.H * a.W * C64) return cudaErrorInvalidValue;
  StripB geo;
  geo.nsx = (a.W + 39) / 40;
  geo.SW = (a.W + geo.nsx - 1) / geo.nsx;
  if (mode == 2 && (geo.SW & 1)) return cudaErrorInvalidValue;      // up-add: even strips (W is even)
  geo.SWH = geo.SW + 2;
  geo.RB = (mode == 2) ? 2 : 3;                                     // up-add children pair up inside a block
  if (geo.RB * geo.SWH > 128) return cudaErrorInvalidValue;
  geo.NB = (a.H + geo.RB - 1) / geo.RB;
  geo.G = a.B * geo.nsx * geo.NB;
  CUtensorMap tm[3];
  memset(tm, 0, sizeof tm);
  const float* base[3] = {a.za, a.dout, a.zout};
  for (int i = (mode == 0 ? 0 : 1); i < 3; ++i) {
    // z_in lands in the layout MMA 3 reads as an MN-major operand (32-byte swizzle atoms)
    cudaError_t e = make_nhwc_map(&tm[i], base[i], 64, a.W, a.H, a.B, 64, (long long)a.H * a.W * 64, 32, geo.SWH,
                                  geo.RB, i == 0 ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B);
    if (e != cudaSuccess) return e;
  }
  const size_t smem = Off::TOTAL + 1024;
  int grid = num_sms < geo.G ? num_sms : geo.G;
  if (grid > kMaxPartialCtas) grid = kMaxPartialCtas;
  auto go = [&](auto kern) -> cudaError_t {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    kern<<<grid, NT, smem, s>>>(tm[0], tm[1], tm[2], a, geo, status);
    e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    return launch_reduce_partials(a.partial, grid, 64 * 64 + 11 * 64, a.gw1, s);
  };
  if (mode == 0) return go(unit_bwd_st_kernel<0, 3>);
  if (mode == 1) return go(unit_bwd_st_kernel<1, 3>);
  if (mode == 2) return go(unit_bwd_st_kernel<2, 2>);
  return cudaErrorInvalidValue;
}

}  // namespace yunet
