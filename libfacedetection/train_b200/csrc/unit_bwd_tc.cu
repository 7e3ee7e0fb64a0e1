// Fused ConvDPUnit backward (64 -> 64 channels, plain load mode, BatchNorm on the output) with
// the two pixel-major GEMMs on the 5th-gen tensor cores (sm_100a), 3xTF32 error-compensated:
//
//   per tile (8 x 16 halo pixels = one M=128 block, 6 x 14 interior):
//     TMA      z_in halo tile (SWIZZLE_128B)                              -> shared
//     LDG      du, z_out -> g = gamma*rstd*(du - mean(du) - zhat*mean(du*zhat)) -> shared (halo)
//     convert  a = relu(bn(z_in)) row per thread -> tf32 hi/lo -> TMEM (tcgen05.st)
//     MMA 1    y = a W1^T            (recomputed pointwise output; never stored in the forward)
//     dw-bwd   dy = sum_k W2[k] g[q-d_k], dW2 += y g[q-d_k], db2 += g, db1 += dy   (CUDA cores)
//     convert  dy rows -> hi/lo -> TMEM
//     MMA 2    h = dy W1             || overlapped with ||  dW1 += dy^T a  (CUDA cores, fp32)
//     epilogue du_in = h * [u_in > 0] written once (or accumulated), sum(du_in), sum(du_in*zhat)
//
// Persistent CTAs (1 per SM); parameter gradients and statistics live in registers across all
// tiles and are flushed once.  Same math as unit_bwd_kernel (kernels_bwd.cu), which stays the
// exact-fp32 reference path and serves the other channel configurations / load modes.
#include <cstdio>
#include <cstring>

#include "kernels.h"
#include "tc_common.cuh"

namespace yunet {

namespace {

using namespace tc;

constexpr int NT = 256;
constexpr int C64 = 64;
constexpr int HR = 8, HC = 16;          // halo tile rows / cols  (128 pixels = TMEM lanes)
constexpr int IR = HR - 2, IC = HC - 2; // 6 x 14 interior
constexpr uint32_t TILE_BYTES = 128 * C64 * 4;   // 32 KB
constexpr uint32_t TMEM_COLS = 256;
constexpr uint32_t COL_D1 = 0, COL_D2 = 64, COL_AHI = 128, COL_ALO = 192;

struct Off {
  static constexpr uint32_t RAW = 0;                       // z_in tile (TMA, 2 k-blocks of 16 KB)
  static constexpr uint32_t G = RAW + TILE_BYTES;          // g halo tile   [128][64] swizzled
  static constexpr uint32_t Y = G + TILE_BYTES;            // y -> dy tile  [128][64] swizzled
  static constexpr uint32_t B1HI = Y + TILE_BYTES;         // W1 hi  [co][ci] K-major SW128
  static constexpr uint32_t B1LO = B1HI + 16384;
  static constexpr uint32_t B2HI = B1LO + 16384;           // W1^T hi [ci][co] K-major SW128
  static constexpr uint32_t B2LO = B2HI + 16384;
  static constexpr uint32_t W2 = B2LO + 16384;             // [9][64]
  static constexpr uint32_t B1 = W2 + 9 * 64 * 4;          // bias1 [64]
  static constexpr uint32_t CA = B1 + 256;                 // scale, shift, mean, rstd of the input [4][64]
  static constexpr uint32_t CO = CA + 1024;                // gscale, m1, m2, mean, rstd of the output [5][64]
  static constexpr uint32_t BAR = CO + 1280;               // 3 mbarriers + tmem ptr
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

__global__ void __launch_bounds__(NT, 1)
unit_bwd_tc_kernel(const __grid_constant__ CUtensorMap tmap, const UnitBwdArgs a, int* status) {
  extern __shared__ unsigned char smem_dyn[];
  unsigned char* smem = reinterpret_cast<unsigned char*>(
      (reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~uintptr_t(1023));
  unsigned char* raw = smem + Off::RAW;
  unsigned char* sG = smem + Off::G;
  unsigned char* sY = smem + Off::Y;
  float* sW2 = reinterpret_cast<float*>(smem + Off::W2);
  float* sB1 = reinterpret_cast<float*>(smem + Off::B1);
  float* sCa = reinterpret_cast<float*>(smem + Off::CA);
  float* sCo = reinterpret_cast<float*>(smem + Off::CO);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Off::BAR);   // tma, mma1, mma2
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 3);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int quarter = warp & 3;      // TMEM lane quarter
  const int half = warp >> 2;        // which 32 of the 64 channels this warp converts / reads back
  const int row = quarter * 32 + lane;   // pixel of the tile == TMEM lane

  if (warp == 0) tmem_alloc<TMEM_COLS>(tmem_ptr);
  if (tid == 0) {
    mbar_init(&bars[0], 1); mbar_init(&bars[1], 1); mbar_init(&bars[2], 1);
    mbar_fence_init();
    tma_prefetch_desc(&tmap);
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
  const uint32_t tbase = *tmem_ptr;
  const uint32_t lane_addr = tbase + ((uint32_t)(quarter * 32) << 16);
  constexpr uint32_t idesc = make_idesc_tf32(128, 64);

  // ---- persistent accumulators
  // depthwise stage: thread -> (channel quad, interior column), marches the 6 interior rows
  const int dq = tid & 15, dx = tid >> 4;            // dx < 14 active
  float4 gw2[9];
#pragma unroll
  for (int k = 0; k < 9; ++k) gw2[k] = make_float4(0.f, 0.f, 0.f, 0.f);
  float4 gb2 = make_float4(0.f, 0.f, 0.f, 0.f), gb1 = gb2;
  // dW1: 4x4 block per thread
  const int co3 = (tid >> 4) * 4, ci3 = (tid & 15) * 4;
  float gw1[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) gw1[i][j] = 0.f;
  const float4 sc3 = *reinterpret_cast<const float4*>(sCa + ci3);
  const float4 sh3 = *reinterpret_cast<const float4*>(sCa + 64 + ci3);
  // statistics of du_in: lane L of a warp owns channel half*32 + L
  double s1 = 0.0, s2 = 0.0;

  const int tiles_x = (a.W + IC - 1) / IC, tiles_y = (a.H + IR - 1) / IR;
  const int ntiles = tiles_x * tiles_y * a.B;
  bool alive = true;
  uint32_t it = 0;
  for (int tile = blockIdx.x; tile < ntiles && alive; tile += gridDim.x, ++it) {
    const uint32_t ph = it & 1;
    int t = tile;
    const int tx = t % tiles_x; t /= tiles_x;
    const int ty = t % tiles_y;
    const int b = t / tiles_y;
    const int x0 = tx * IC, y0 = ty * IR;          // interior origin; halo origin = (x0-1, y0-1)
    const long long img_off = (long long)b * a.H * a.W * C64;

    // ---- T0: TMA for z_in, vector loads for du / z_out -> g
    if (tid == 0) {
      mbar_arrive_expect_tx(&bars[0], TILE_BYTES);
      tma_load_4d(raw, &tmap, &bars[0], 0, x0 - 1, y0 - 1, b);
      tma_load_4d(raw + 16384, &tmap, &bars[0], 32, x0 - 1, y0 - 1, b);
    }
    {
      const float* dimg = a.dout + (long long)b * a.dout_batch_stride;
      const float* zimg = a.zout + img_off;
#pragma unroll 4
      for (int k = 0; k < 128 * 16 / NT; ++k) {
        const int i = tid + k * NT;
        const int pix = i >> 4, ch = i & 15;
        const int gy = y0 - 1 + pix / HC, gx = x0 - 1 + pix % HC;
        float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
        if (gy >= 0 && gy < a.H && gx >= 0 && gx < a.W) {
          const long long off = ((long long)gy * a.W + gx) * C64 + ch * 4;
          const float4 d = __ldg(reinterpret_cast<const float4*>(dimg + off));
          const float4 z = __ldg(reinterpret_cast<const float4*>(zimg + off));
          const float4 gs = *reinterpret_cast<const float4*>(sCo + ch * 4);
          const float4 m1 = *reinterpret_cast<const float4*>(sCo + 64 + ch * 4);
          const float4 m2 = *reinterpret_cast<const float4*>(sCo + 128 + ch * 4);
          const float4 mu = *reinterpret_cast<const float4*>(sCo + 192 + ch * 4);
          const float4 rs = *reinterpret_cast<const float4*>(sCo + 256 + ch * 4);
          g.x = gs.x * (d.x - m1.x - (z.x - mu.x) * rs.x * m2.x);
          g.y = gs.y * (d.y - m1.y - (z.y - mu.y) * rs.y * m2.y);
          g.z = gs.z * (d.z - m1.z - (z.z - mu.z) * rs.z * m2.z);
          g.w = gs.w * (d.w - m1.w - (z.w - mu.w) * rs.w * m2.w);
        }
        *reinterpret_cast<float4*>(tchunk(sG, pix, ch)) = g;
      }
    }
    if (!mbar_wait(&bars[0], ph)) { alive = false; if (lane == 0) atomicExch(status, 11); }

    // ---- T1: a = relu(bn(z_in)) row per thread (32 channels per warp half) -> hi/lo -> TMEM
    const int hy = row / HC, hx = row % HC;
    const int gy_r = y0 - 1 + hy, gx_r = x0 - 1 + hx;
    const bool interior = hy >= 1 && hy <= IR && hx >= 1 && hx <= IC && gy_r < a.H && gx_r < a.W;
    if (alive) {
#pragma unroll
      for (int g16 = 0; g16 < 2; ++g16) {
        uint32_t hi[16], lo[16];
#pragma unroll
        for (int c4 = 0; c4 < 4; ++c4) {
          const int ch = half * 8 + g16 * 4 + c4;
          const float4 z = *reinterpret_cast<const float4*>(rchunk(raw, row, ch));
          const float4 sc = *reinterpret_cast<const float4*>(sCa + ch * 4);
          const float4 sh = *reinterpret_cast<const float4*>(sCa + 64 + ch * 4);
          const float v[4] = {fmaxf(fmaf(z.x, sc.x, sh.x), 0.f), fmaxf(fmaf(z.y, sc.y, sh.y), 0.f),
                              fmaxf(fmaf(z.z, sc.z, sh.z), 0.f), fmaxf(fmaf(z.w, sc.w, sh.w), 0.f)};
#pragma unroll
          for (int j = 0; j < 4; ++j) { hi[c4 * 4 + j] = tf32_hi(v[j]); lo[c4 * 4 + j] = tf32_lo(v[j]); }
        }
        tmem_st16(lane_addr + COL_AHI + half * 32 + g16 * 16, hi);
        tmem_st16(lane_addr + COL_ALO + half * 32 + g16 * 16, lo);
      }
      tmem_wait_st();
    }
    tc_fence_before();
    __syncthreads();
    // ---- T2: MMA 1   D1 = a W1^T
    if (tid == 0 && alive) {
      tc_fence_after();
      const uint32_t bhi = smem_u32(smem + Off::B1HI), blo = smem_u32(smem + Off::B1LO);
      uint32_t acc = 0;
#pragma unroll
      for (int pass = 0; pass < 3; ++pass)
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const uint32_t koff = (k >> 2) * 8192 + (k & 3) * 32;
          mma_tf32_ts(tbase + COL_D1, tbase + (pass == 0 ? COL_ALO : COL_AHI) + k * 8,
                      make_desc_sw128_kmajor((pass == 1 ? blo : bhi) + koff), idesc, acc);
          acc = 1;
        }
      mma_commit(&bars[1]);
    }
    if (alive && !mbar_wait(&bars[1], ph)) { alive = false; if (lane == 0) atomicExch(status, 12); }
    tc_fence_after();
    // ---- T3: y (+bias) for interior in-image pixels, exact 0 elsewhere -> sY
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

    // ---- T4: depthwise backward on the interior: dy in place over y, dW2, db2, db1
    if (alive && dx < IC) {
      float4 w2r[9];
#pragma unroll
      for (int k = 0; k < 9; ++k) w2r[k] = *reinterpret_cast<const float4*>(sW2 + k * 64 + dq * 4);
      float4 ra[3], rb[3], rc[3];
#pragma unroll
      for (int d = 0; d < 3; ++d) {
        ra[d] = *reinterpret_cast<const float4*>(tchunk(sG, 0 * HC + dx + d, dq));
        rb[d] = *reinterpret_cast<const float4*>(tchunk(sG, 1 * HC + dx + d, dq));
      }
#pragma unroll
      for (int r = 0; r < IR; ++r) {          // interior row r <-> halo row r+1
#pragma unroll
        for (int d = 0; d < 3; ++d)
          rc[d] = *reinterpret_cast<const float4*>(tchunk(sG, (r + 2) * HC + dx + d, dq));
        const int pix = (r + 1) * HC + dx + 1;
        const bool in = (y0 + r) < a.H && (x0 + dx) < a.W;
        const float4 y = *reinterpret_cast<const float4*>(tchunk(sY, pix, dq));
        float4 dy = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
          const float4 g0 = rc[2 - kx], g1 = rb[2 - kx], g2 = ra[2 - kx];
          const float4 w0 = w2r[kx], w1 = w2r[3 + kx], w2 = w2r[6 + kx];
          dy.x = fmaf(w0.x, g0.x, dy.x); dy.y = fmaf(w0.y, g0.y, dy.y); dy.z = fmaf(w0.z, g0.z, dy.z); dy.w = fmaf(w0.w, g0.w, dy.w);
          dy.x = fmaf(w1.x, g1.x, dy.x); dy.y = fmaf(w1.y, g1.y, dy.y); dy.z = fmaf(w1.z, g1.z, dy.z); dy.w = fmaf(w1.w, g1.w, dy.w);
          dy.x = fmaf(w2.x, g2.x, dy.x); dy.y = fmaf(w2.y, g2.y, dy.y); dy.z = fmaf(w2.z, g2.z, dy.z); dy.w = fmaf(w2.w, g2.w, dy.w);
          gw2[kx].x = fmaf(y.x, g0.x, gw2[kx].x); gw2[kx].y = fmaf(y.y, g0.y, gw2[kx].y); gw2[kx].z = fmaf(y.z, g0.z, gw2[kx].z); gw2[kx].w = fmaf(y.w, g0.w, gw2[kx].w);
          gw2[3 + kx].x = fmaf(y.x, g1.x, gw2[3 + kx].x); gw2[3 + kx].y = fmaf(y.y, g1.y, gw2[3 + kx].y); gw2[3 + kx].z = fmaf(y.z, g1.z, gw2[3 + kx].z); gw2[3 + kx].w = fmaf(y.w, g1.w, gw2[3 + kx].w);
          gw2[6 + kx].x = fmaf(y.x, g2.x, gw2[6 + kx].x); gw2[6 + kx].y = fmaf(y.y, g2.y, gw2[6 + kx].y); gw2[6 + kx].z = fmaf(y.z, g2.z, gw2[6 + kx].z); gw2[6 + kx].w = fmaf(y.w, g2.w, gw2[6 + kx].w);
        }
        gb2.x += rb[1].x; gb2.y += rb[1].y; gb2.z += rb[1].z; gb2.w += rb[1].w;
        if (!in) dy = make_float4(0.f, 0.f, 0.f, 0.f);
        gb1.x += dy.x; gb1.y += dy.y; gb1.z += dy.z; gb1.w += dy.w;
        *reinterpret_cast<float4*>(tchunk(sY, pix, dq)) = dy;
#pragma unroll
        for (int d = 0; d < 3; ++d) { ra[d] = rb[d]; rb[d] = rc[d]; }
      }
    }
    __syncthreads();

    // ---- T5: dy rows -> hi/lo -> TMEM (A columns are free: MMA 1 completed)
    if (alive) {
#pragma unroll
      for (int g16 = 0; g16 < 2; ++g16) {
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
      }
      tmem_wait_st();
    }
    tc_fence_before();
    __syncthreads();
    // ---- T6: MMA 2   D2 = dy W1   (async) ...
    if (tid == 0 && alive) {
      tc_fence_after();
      const uint32_t bhi = smem_u32(smem + Off::B2HI), blo = smem_u32(smem + Off::B2LO);
      uint32_t acc = 0;
#pragma unroll
      for (int pass = 0; pass < 3; ++pass)
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const uint32_t koff = (k >> 2) * 8192 + (k & 3) * 32;
          mma_tf32_ts(tbase + COL_D2, tbase + (pass == 0 ? COL_ALO : COL_AHI) + k * 8,
                      make_desc_sw128_kmajor((pass == 1 ? blo : bhi) + koff), idesc, acc);
          acc = 1;
        }
      mma_commit(&bars[2]);
    }
    // ---- ... while the CUDA cores do  dW1 += dy^T a  over the interior pixels
    if (alive) {
      const int cch = co3 >> 2, ach = ci3 >> 2;
      for (int r = 1; r <= IR; ++r) {
#pragma unroll 2
        for (int x = 1; x <= IC; ++x) {
          const int pix = r * HC + x;
          const float4 d4 = *reinterpret_cast<const float4*>(tchunk(sY, pix, cch));
          const float4 z4 = *reinterpret_cast<const float4*>(rchunk(raw, pix, ach));
          float4 a4;
          a4.x = fmaxf(fmaf(z4.x, sc3.x, sh3.x), 0.f); a4.y = fmaxf(fmaf(z4.y, sc3.y, sh3.y), 0.f);
          a4.z = fmaxf(fmaf(z4.z, sc3.z, sh3.z), 0.f); a4.w = fmaxf(fmaf(z4.w, sc3.w, sh3.w), 0.f);
          const float dd[4] = {d4.x, d4.y, d4.z, d4.w};
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            gw1[i][0] = fmaf(dd[i], a4.x, gw1[i][0]); gw1[i][1] = fmaf(dd[i], a4.y, gw1[i][1]);
            gw1[i][2] = fmaf(dd[i], a4.z, gw1[i][2]); gw1[i][3] = fmaf(dd[i], a4.w, gw1[i][3]);
          }
        }
      }
    }
    if (alive && !mbar_wait(&bars[2], ph)) { alive = false; if (lane == 0) atomicExch(status, 13); }
    tc_fence_after();
    // ---- T7: epilogue: du_in = h * [u_in > 0], statistics (dy of out-of-image pixels was zeroed,
    // so h is zero there; only interior in-image rows write)
    {
      float v1[32], v2[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) { v1[j] = 0.f; v2[j] = 0.f; }
      if (alive) {
#pragma unroll
        for (int g16 = 0; g16 < 2; ++g16) {
          uint32_t hv[16];
          tmem_ld16(lane_addr + COL_D2 + half * 32 + g16 * 16, hv);
          tmem_wait_ld();
          if (interior) {
            float* dst = a.dua + img_off + ((long long)gy_r * a.W + gx_r) * C64;
#pragma unroll
            for (int c4 = 0; c4 < 4; ++c4) {
              const int ch = half * 8 + g16 * 4 + c4;
              const float4 z = *reinterpret_cast<const float4*>(rchunk(raw, row, ch));
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
              float4* p = reinterpret_cast<float4*>(dst + ch * 4);
              if (a.acc_a) { const float4 o = *p; d.x += o.x; d.y += o.y; d.z += o.z; d.w += o.w; }
              *p = d;
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
    }
    fence_proxy_async_smem();
    tc_fence_before();
    alive = __syncthreads_and(alive ? 1 : 0) != 0;
  }

  // ---- flush
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) atomicAdd(a.gw1 + (co3 + i) * 64 + ci3 + j, gw1[i][j]);
  {
    float vals[44];
#pragma unroll
    for (int k = 0; k < 9; ++k) { vals[k * 4] = gw2[k].x; vals[k * 4 + 1] = gw2[k].y; vals[k * 4 + 2] = gw2[k].z; vals[k * 4 + 3] = gw2[k].w; }
    vals[36] = gb2.x; vals[37] = gb2.y; vals[38] = gb2.z; vals[39] = gb2.w;
    vals[40] = gb1.x; vals[41] = gb1.y; vals[42] = gb1.z; vals[43] = gb1.w;
#pragma unroll
    for (int v = 0; v < 44; ++v) vals[v] += __shfl_xor_sync(0xffffffffu, vals[v], 16);
    if (lane < 16) {
#pragma unroll
      for (int k = 0; k < 9; ++k)
#pragma unroll
        for (int c = 0; c < 4; ++c) atomicAdd(a.gw2 + (dq * 4 + c) * 9 + k, vals[k * 4 + c]);
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        atomicAdd(a.gb2 + dq * 4 + c, vals[36 + c]);
        atomicAdd(a.gb1 + dq * 4 + c, vals[40 + c]);
      }
    }
  }
  atomicAdd(a.dsum_a + half * 32 + lane, s1);
  atomicAdd(a.dsumzh_a + half * 32 + lane, s2);
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<TMEM_COLS>(tbase);
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                             const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                             CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                             CUtensorMapFloatOOBfill);
EncodeFn get_encode_bwd() {
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

}  // namespace

int unit_bwd_tc_supported(int cin, int cout, int mode, int has_bn) {
  return cin == 64 && cout == 64 && mode == 0 && has_bn && get_encode_bwd() != nullptr;
}

cudaError_t launch_unit_bwd_tc(const UnitBwdArgs& a, int num_sms, int* status, cudaStream_t s) {
  EncodeFn enc = get_encode_bwd();
  if (!enc) return cudaErrorNotSupported;
  CUtensorMap tm;
  cuuint64_t dims[4] = {64, (cuuint64_t)a.W, (cuuint64_t)a.H, (cuuint64_t)a.B};
  cuuint64_t strides[3] = {256, (cuuint64_t)a.W * 256, (cuuint64_t)a.H * a.W * 256};
  cuuint32_t box[4] = {32, HC, HR, 1};
  cuuint32_t es[4] = {1, 1, 1, 1};
  CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(a.za), dims, strides, box,
                   es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return cudaErrorInvalidValue;
  const size_t smem = Off::TOTAL + 1024;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(unit_bwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    configured = true;
  }
  const int ntiles = ((a.W + IC - 1) / IC) * ((a.H + IR - 1) / IR) * a.B;
  int grid = num_sms < ntiles ? num_sms : ntiles;
  unit_bwd_tc_kernel<<<grid, NT, smem, s>>>(tm, a, status);
  return cudaGetLastError();
}

}  // namespace yunet
