// Forward kernels of the YuNet hot path, fp32 NHWC, sm_100a.
//
//   unit_fwd_kernel  — one fused ConvDPUnit (mmdet/models/utils/yunet_layer.py:30-36):
//       load prologue: a = relu(bn(z_in)) [+ 2x2 max-pool (yunet_backbone.py:40) | + nearest-up2
//       add (tfpn.py:39-40)] -> pointwise 1x1 GEMM (+bias) on the halo tile -> depthwise 3x3
//       stencil (+bias) from shared memory -> store pre-BN z once + per-channel sum / sum^2
//       (train-mode BatchNorm statistics, applied by the *consumer's* prologue).
//   stem_fwd_kernel  — Conv_head.conv1: dense 3x3 stride-2 conv 3->16 (yunet_layer.py:51,58),
//       reads NCHW input, writes NHWC pre-BN + statistics.
//   bn_update_running_kernel — running_mean / running_var update (torch BatchNorm2d, momentum,
//       unbiased variance) for all BatchNorm layers in one launch.
//
// The pointwise GEMM here is the fp32 CUDA-core version (exact fp32, the parity baseline);
// the tcgen05 3xTF32 version for the 64-channel units lives in unit_fwd_tc.cu.
#include <cstdio>

#include "kernels.h"
#include "prefetch.cuh"
#include "f32x2.cuh"

namespace yunet {

namespace {

constexpr int NT = 256;

__device__ __forceinline__ void bn_coeffs(const BnRef& r, int c, float& scale, float& shift) {
  float m, v;
  if (r.train) {
    double dm = r.sum[c] * r.inv_count;
    double dv = r.sumsq[c] * r.inv_count - dm * dm;
    if (dv < 0.0) dv = 0.0;
    m = (float)dm;
    v = (float)dv;
  } else {
    m = r.rmean[c];
    v = r.rvar[c];
  }
  float rstd = 1.0f / sqrtf(v + kBnEps);
  scale = r.gamma[c] * rstd;
  shift = r.beta[c] - m * scale;
}

__device__ __forceinline__ float4 bn_relu4(float4 z, float4 sc, float4 sh) {
  float4 r;
  r.x = fmaxf(fmaf(z.x, sc.x, sh.x), 0.f);
  r.y = fmaxf(fmaf(z.y, sc.y, sh.y), 0.f);
  r.z = fmaxf(fmaf(z.z, sc.z, sh.z), 0.f);
  r.w = fmaxf(fmaf(z.w, sc.w, sh.w), 0.f);
  return r;
}
__device__ __forceinline__ float4 max4(float4 a, float4 b) {
  return make_float4(fmaxf(a.x, b.x), fmaxf(a.y, b.y), fmaxf(a.z, b.z), fmaxf(a.w, b.w));
}
__device__ __forceinline__ float4 add4(float4 a, float4 b) {
  return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w);
}
__device__ __forceinline__ void fma4(float4& acc, float4 w, float4 v) {
  fma4p(acc, w, v);
}
__device__ __forceinline__ float4 ldg4(const float* p) {
  return __ldg(reinterpret_cast<const float4*>(p));
}

template <int CIN, int COUT>
struct FwdCfg {
  // output tile: 8 x 16 pixels; the 16-channel units (the 160^2 / 80^2 layers: few FLOPs, many
  // pixels) take 16 x 32 so the per-tile barriers and the exposed load latency are amortised
  // over 4x the pixels
  static constexpr bool BIG = (CIN <= 16 && COUT <= 16);
  static constexpr int TH = BIG ? 16 : 8, TW = BIG ? 32 : 16;
  static constexpr int HH = TH + 2, HW = TW + 2;     // halo tile
  static constexpr int HP = HH * HW;                 // 180 (612) halo pixels
  static constexpr int HPP = BIG ? 640 : 192;        // padded to the GEMM's pixel blocking
  static constexpr int CPT = (COUT == 16) ? 4 : 8;   // output channels per thread in the GEMM
  static constexpr int NCG = COUT / CPT;             // channel groups
  static constexpr int NPG = NT / NCG;               // pixel groups
  static constexpr int PPT = HPP / NPG;              // pixels per thread
  static constexpr int AS = CIN + 4;                 // sA row stride (floats)
  static constexpr int R0 = (HPP * AS > HP * COUT) ? HPP * AS : HP * COUT;  // sA / sY union
  static constexpr int NQ = COUT / 4;                // channel quads (dw stage)
  static constexpr int RG = NT / (NQ * TW);          // row groups
  static constexpr int RPT = TH / RG;                // rows per thread
  static constexpr int SMEM_FLOATS = R0 + CIN * COUT + 9 * COUT + 2 * COUT + 4 * CIN;
  static_assert(HPP % NPG == 0, "pixel blocking");
  static_assert(NQ * TW * RG == NT && RG * RPT == TH, "dw mapping");
};

template <int CIN, int COUT, int MODE>
__global__ void __launch_bounds__(NT, 2) unit_fwd_kernel(const UnitFwdArgs a) {
  using C = FwdCfg<CIN, COUT>;
  extern __shared__ float4 smem_raw[];
  float* smem = reinterpret_cast<float*>(smem_raw);
  float* sA = smem;                       // [C::HPP][AS]
  float* sY = smem;                       // [C::HP][COUT]   (aliases sA after the GEMM)
  float* sW1t = smem + C::R0;             // [CIN][COUT]
  float* sW2 = sW1t + CIN * COUT;         // [9][COUT]
  float* sB1 = sW2 + 9 * COUT;            // [COUT]
  float* sB2 = sB1 + COUT;                // [COUT]
  float* sScA = sB2 + COUT;               // [CIN]
  float* sShA = sScA + CIN;
  float* sScB = sShA + CIN;
  float* sShB = sScB + CIN;
  // PF (plain input, where it fits beside a second CTA): the next tile's raw halo rows are copied in
  // with cp.async while this tile runs its GEMM / depthwise stages (the staging loads were ~40 % of
  // the stall samples of the 160x160 16 -> 16 unit)
  constexpr bool PF = MODE == 0 && (C::SMEM_FLOATS + C::HP * CIN) * 4 <= 112 * 1024;
  float* sRaw = sShB + CIN;               // [C::HP][CIN] raw z of the next / current tile (PF)

  const int tid = threadIdx.x;
  const int tiles_x = (a.W + C::TW - 1) / C::TW;
  const int tiles_y = (a.H + C::TH - 1) / C::TH;
  const int ntiles = tiles_x * tiles_y * a.B;

  // ---- stage 0 (once per persistent CTA): weights + BN coefficients into shared memory
#pragma unroll
  for (int it = 0; it < CIN * COUT / NT; ++it) {
    const int i = tid + it * NT;
    int ci = i / COUT, co = i % COUT;
    sW1t[i] = __ldg(a.w1 + co * CIN + ci);
  }
  for (int i = tid; i < 9 * COUT; i += NT) {
    int k = i / COUT, co = i % COUT;
    sW2[i] = __ldg(a.w2 + co * 9 + k);
  }
  if (tid < COUT) { sB1[tid] = __ldg(a.b1 + tid); sB2[tid] = __ldg(a.b2 + tid); }
  if (tid < CIN) {
    float sc, sh;
    bn_coeffs(a.bna, tid, sc, sh);
    sScA[tid] = sc; sShA[tid] = sh;
    if (MODE == 2) {
      bn_coeffs(a.bnb, tid, sc, sh);
      sScB[tid] = sc; sShB[tid] = sh;
    }
  }
  __syncthreads();

  // depthwise-stage mapping, weights in registers, BatchNorm statistics accumulated over all tiles
  const int dq = tid % C::NQ;
  const int dx = (tid / C::NQ) % C::TW;
  const int dr0 = (tid / (C::NQ * C::TW)) * C::RPT;
  double st1[4] = {0.0, 0.0, 0.0, 0.0}, st2[4] = {0.0, 0.0, 0.0, 0.0};

  auto stage_in = [&](int tile_) {
    int t = tile_;
    const int tx = t % tiles_x; t /= tiles_x;
    const int ty = t % tiles_y;
    const int b = t / tiles_y;
    const int x0 = tx * C::TW, y0 = ty * C::TH;
    constexpr int Q = CIN / 4;
    const float* img = a.za + (long long)b * a.H * a.W * CIN;
#pragma unroll 4
    for (int it = 0; it < C::HPP * Q / NT; ++it) {
      const int i = tid + it * NT;
      const int pix = i / Q, q = i % Q;
      if (pix < C::HP) {
        const int hy = pix / C::HW, hx = pix % C::HW;
        const int gy = y0 + hy - 1, gx = x0 + hx - 1;
        const bool in = gy >= 0 && gy < a.H && gx >= 0 && gx < a.W;
        cp_async16(sRaw + pix * CIN + q * 4, in ? img + ((long long)gy * a.W + gx) * CIN + q * 4 : a.za, in);
      }
    }
    cp_async_commit();
  };
  if (PF && (int)blockIdx.x < ntiles) stage_in(blockIdx.x);

  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
  int t = tile;
  const int tx = t % tiles_x; t /= tiles_x;
  const int ty = t % tiles_y;
  const int b = t / tiles_y;
  const int x0 = tx * C::TW, y0 = ty * C::TH;

  // ---- L2 prefetch of the next tile's input rows (one bulk request per row)
  if (!PF) {
    const int nt = tile + gridDim.x;
    if (nt < ntiles && tid < 96) {
      int t2 = nt;
      const int ntx = t2 % tiles_x; t2 /= tiles_x;
      const int nty = t2 % tiles_y;
      const int nb = t2 / tiles_y;
      const int nx0 = ntx * C::TW, ny0 = nty * C::TH;
      if (MODE == 1) {
        l2_prefetch_tile<CIN>(a.za + (long long)nb * a.H * a.W * 4 * CIN, a.H * 2, a.W * 2,
                              (ny0 - 1) * 2, (ny0 + C::TH + 1) * 2, (nx0 - 1) * 2,
                              (nx0 + C::TW + 1) * 2, tid);
      } else {
        if (tid < 64)
          l2_prefetch_tile<CIN>(a.za + (long long)nb * a.H * a.W * CIN, a.H, a.W, ny0 - 1,
                                ny0 + C::TH + 1, nx0 - 1, nx0 + C::TW + 1, tid);
        else if (MODE == 2)
          l2_prefetch_tile<CIN>(a.zb + (long long)nb * (a.H >> 1) * (a.W >> 1) * CIN, a.H >> 1,
                                a.W >> 1, (ny0 - 1) >> 1, ((ny0 + C::TH) >> 1) + 1, (nx0 - 1) >> 1,
                                ((nx0 + C::TW) >> 1) + 1, tid - 64);
      }
    }
  }

  // ---- stage 1: halo tile of activated inputs -> sA
  {
    constexpr int Q = CIN / 4;
    static_assert((C::HPP * Q) % NT == 0, "prologue trip count");
    // fixed trip count + partial unroll: the global loads of several iterations are in flight
    // together instead of one dependent load per iteration
    if (PF) cp_async_wait<0>();         // a thread reads back exactly the chunks it copied
#pragma unroll 4
    for (int it = 0; it < C::HPP * Q / NT; ++it) {
      const int i = tid + it * NT;
      const int pix = i / Q, q = i % Q;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (pix < C::HP) {
        const int hy = pix / C::HW, hx = pix % C::HW;
        const int gy = y0 + hy - 1, gx = x0 + hx - 1;
        if (gy >= 0 && gy < a.H && gx >= 0 && gx < a.W) {
          const float4 sc = *reinterpret_cast<const float4*>(sScA + q * 4);
          const float4 sh = *reinterpret_cast<const float4*>(sShA + q * 4);
          if (MODE == 0) {
            const float* p = a.za + (((long long)b * a.H + gy) * a.W + gx) * CIN + q * 4;
            v = bn_relu4(PF ? *reinterpret_cast<const float4*>(sRaw + pix * CIN + q * 4) : ldg4(p), sc, sh);
          } else if (MODE == 1) {
            const int W2 = a.W * 2;
            const float* p = a.za + (((long long)b * (a.H * 2) + gy * 2) * W2 + gx * 2) * CIN + q * 4;
            float4 v00 = bn_relu4(ldg4(p), sc, sh);
            float4 v01 = bn_relu4(ldg4(p + CIN), sc, sh);
            float4 v10 = bn_relu4(ldg4(p + (long long)W2 * CIN), sc, sh);
            float4 v11 = bn_relu4(ldg4(p + (long long)W2 * CIN + CIN), sc, sh);
            v = max4(max4(v00, v01), max4(v10, v11));
          } else {
            const float* p = a.za + (((long long)b * a.H + gy) * a.W + gx) * CIN + q * 4;
            v = bn_relu4(ldg4(p), sc, sh);
            const int Hb = a.H >> 1, Wb = a.W >> 1;
            const float* pb = a.zb + (((long long)b * Hb + (gy >> 1)) * Wb + (gx >> 1)) * CIN + q * 4;
            const float4 scb = *reinterpret_cast<const float4*>(sScB + q * 4);
            const float4 shb = *reinterpret_cast<const float4*>(sShB + q * 4);
            v = add4(v, bn_relu4(ldg4(pb), scb, shb));
          }
        }
      }
      *reinterpret_cast<float4*>(sA + pix * C::AS + q * 4) = v;
    }
  }
  __syncthreads();
  if (PF && tile + (int)gridDim.x < ntiles) stage_in(tile + gridDim.x);

  // ---- stage 2: pointwise GEMM  y[p][co] = sum_ci a[p][ci] * W1[co][ci]
  const int cg = tid % C::NCG;
  const int pg = tid / C::NCG;
  float acc[C::PPT][C::CPT];
#pragma unroll
  for (int i = 0; i < C::PPT; ++i)
#pragma unroll
    for (int j = 0; j < C::CPT; ++j) acc[i][j] = 0.f;

#pragma unroll 2
  for (int k4 = 0; k4 < CIN / 4; ++k4) {
    float4 av[C::PPT];
#pragma unroll
    for (int i = 0; i < C::PPT; ++i)
      av[i] = *reinterpret_cast<const float4*>(sA + (pg + i * C::NPG) * C::AS + k4 * 4);
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      float wv[C::CPT];
#pragma unroll
      for (int j4 = 0; j4 < C::CPT / 4; ++j4) {
        float4 w = *reinterpret_cast<const float4*>(sW1t + (k4 * 4 + kk) * COUT + cg * C::CPT + j4 * 4);
        wv[j4 * 4 + 0] = w.x; wv[j4 * 4 + 1] = w.y; wv[j4 * 4 + 2] = w.z; wv[j4 * 4 + 3] = w.w;
      }
#pragma unroll
      for (int i = 0; i < C::PPT; ++i) {
        const float ak = kk == 0 ? av[i].x : kk == 1 ? av[i].y : kk == 2 ? av[i].z : av[i].w;
#pragma unroll
        for (int j = 0; j < C::CPT; j += 2) fma2(acc[i][j], acc[i][j + 1], ak, ak, wv[j], wv[j + 1]);
      }
    }
  }
  __syncthreads();   // every read of sA is done; the region becomes sY

  // y (+bias) for in-image pixels, exact zero outside (the depthwise conv zero-pads y)
#pragma unroll
  for (int i = 0; i < C::PPT; ++i) {
    const int pix = pg + i * C::NPG;
    if (pix < C::HP) {
      const int hy = pix / C::HW, hx = pix % C::HW;
      const int gy = y0 + hy - 1, gx = x0 + hx - 1;
      const bool in = gy >= 0 && gy < a.H && gx >= 0 && gx < a.W;
#pragma unroll
      for (int j4 = 0; j4 < C::CPT / 4; ++j4) {
        float4 o;
        const float* bb = sB1 + cg * C::CPT + j4 * 4;
        o.x = in ? acc[i][j4 * 4 + 0] + bb[0] : 0.f;
        o.y = in ? acc[i][j4 * 4 + 1] + bb[1] : 0.f;
        o.z = in ? acc[i][j4 * 4 + 2] + bb[2] : 0.f;
        o.w = in ? acc[i][j4 * 4 + 3] + bb[3] : 0.f;
        *reinterpret_cast<float4*>(sY + pix * COUT + cg * C::CPT + j4 * 4) = o;
      }
    }
  }
  __syncthreads();

  // ---- stage 3: depthwise 3x3 from shared memory, store z, statistics
  {
    float4 w[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) w[k] = *reinterpret_cast<const float4*>(sW2 + k * COUT + dq * 4);
    const float4 bias = *reinterpret_cast<const float4*>(sB2 + dq * 4);
    float4 ra[3], rb[3], rc[3];
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      ra[d] = *reinterpret_cast<const float4*>(sY + ((dr0 + 0) * C::HW + dx + d) * COUT + dq * 4);
      rb[d] = *reinterpret_cast<const float4*>(sY + ((dr0 + 1) * C::HW + dx + d) * COUT + dq * 4);
    }
    float4 s1 = make_float4(0.f, 0.f, 0.f, 0.f), s2 = s1;
    const int gx = x0 + dx;
#pragma unroll
    for (int i = 0; i < C::RPT; ++i) {
#pragma unroll
      for (int d = 0; d < 3; ++d)
        rc[d] = *reinterpret_cast<const float4*>(sY + ((dr0 + i + 2) * C::HW + dx + d) * COUT + dq * 4);
      float4 o = bias;
#pragma unroll
      for (int d = 0; d < 3; ++d) {
        fma4(o, w[d], ra[d]);
        fma4(o, w[3 + d], rb[d]);
        fma4(o, w[6 + d], rc[d]);
      }
      const int gy = y0 + dr0 + i;
      if (gy < a.H && gx < a.W) {
        float* dst = a.zout + (long long)b * a.out_batch_stride + ((long long)gy * a.W + gx) * COUT + dq * 4;
        *reinterpret_cast<float4*>(dst) = o;
        s1 = add4(s1, o);
        fma4p(s2, o, o);
      }
#pragma unroll
      for (int d = 0; d < 3; ++d) { ra[d] = rb[d]; rb[d] = rc[d]; }
    }
    st1[0] += s1.x; st1[1] += s1.y; st1[2] += s1.z; st1[3] += s1.w;
    st2[0] += s2.x; st2[1] += s2.y; st2[2] += s2.z; st2[3] += s2.w;
  }
  __syncthreads();   // sY is overwritten (as sA) by the next tile
  }  // tile loop

  // ---- statistics: fp32 within a tile row-strip, fp64 across tiles / lanes / CTAs
  if (a.osum != nullptr) {
#pragma unroll
    for (int c = 0; c < 4; ++c) {
#pragma unroll
      for (int o = 16; o >= C::NQ; o >>= 1) {
        st1[c] += __shfl_xor_sync(0xffffffffu, st1[c], o);
        st2[c] += __shfl_xor_sync(0xffffffffu, st2[c], o);
      }
    }
    if ((tid & 31) < C::NQ) {
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        atomicAdd(a.osum + dq * 4 + c, st1[c]);
        atomicAdd(a.osumsq + dq * 4 + c, st2[c]);
      }
    }
  }
}

// ------------------------------------------------------------------------------- stem
// Persistent CTAs; tile = 16 rows x 32 cols of output pixels, 2 pixels (rows ly, ly+8) per thread
// so every broadcast weight load feeds two pixels; per-lane fp64 statistics across tiles.
constexpr int ST_TH = 16, ST_TW = 32;
constexpr int ST_IH = 2 * ST_TH + 1, ST_IW = 2 * ST_TW + 1;   // 33 x 65 input patch
// the patch is staged as aligned float4 columns -4 .. 2*ST_TW-1 around the tile's first input column
// (17 per row; input column x of the patch sits at float index x + 3)
constexpr int ST_C4 = 2 * ST_TW / 4 + 1;
constexpr int ST_IWP = ST_C4 * 4;                             // row stride 68

__global__ void __launch_bounds__(256, 2) stem_fwd_kernel(const StemArgs a) {
  // two patch buffers: the next tile is copied in (cp.async, zero-filled outside the image) while
  // the current one is convolved -- the staging loads were 25 % of the stall samples (long scoreboard)
  extern __shared__ __align__(16) float stem_smem[];
  typedef float Patch[3][ST_IH][ST_IWP];
  Patch* sInB = reinterpret_cast<Patch*>(stem_smem);
  __shared__ __align__(16) float sW[27][16];
  __shared__ float sB[16];
  __shared__ double sRedD[8][32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int Ho = a.Hin / 2, Wo = a.Win / 2;
  const int tiles_x = (Wo + ST_TW - 1) / ST_TW, tiles_y = (Ho + ST_TH - 1) / ST_TH;
  const int ntiles = tiles_x * tiles_y * a.B;
  for (int i = tid; i < 27 * 16; i += 256) {
    int k = i / 16, co = i % 16;
    sW[k][co] = __ldg(a.w + co * 27 + k);
  }
  if (tid < 16) sB[tid] = __ldg(a.b + tid);
  const int lx = tid % ST_TW, ly = tid / ST_TW;    // ly in 0..7; second pixel at ly + 8
  double stat = 0.0;    // lane L: sum of channel L (L < 16) / sum of squares of channel L-16

  auto stage = [&](int tile, int buf) {
    int t = tile;
    const int tx = t % tiles_x; t /= tiles_x;
    const int ty = t % tiles_y;
    const int b = t / tiles_y;
    const int gx0 = 2 * tx * ST_TW - 4, iy0 = 2 * ty * ST_TH - 1;
    float* dst = &sInB[buf][0][0][0];
    for (int i = tid; i < 3 * ST_IH * ST_C4; i += 256) {
      const int rowid = i / ST_C4, j = i - rowid * ST_C4;
      const int c = rowid / ST_IH, r = rowid - c * ST_IH;
      const int gy = iy0 + r, gx = gx0 + 4 * j;            // Win % 4 == 0: a chunk is all in or all out
      const bool in = gy >= 0 && gy < a.Hin && gx >= 0 && gx < a.Win;
      const float* src = in ? a.img + (((long long)b * 3 + c) * a.Hin + gy) * a.Win + gx : a.img;
      cp_async16(dst + rowid * ST_IWP + 4 * j, src, in);
    }
  };
  if ((int)blockIdx.x < ntiles) stage(blockIdx.x, 0);
  cp_async_commit();

  int buf = 0;
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, buf ^= 1) {
    int t = tile;
    const int tx = t % tiles_x; t /= tiles_x;
    const int ty = t % tiles_y;
    const int b = t / tiles_y;
    const int ox0 = tx * ST_TW, oy0 = ty * ST_TH;
    __syncthreads();      // the readers of the other buffer (previous tile) are done; covers the weight setup
    if (tile + (int)gridDim.x < ntiles) stage(tile + gridDim.x, buf ^ 1);
    cp_async_commit();
    cp_async_wait<1>();   // everything but the group just committed: this tile's patch has landed
    __syncthreads();
    const Patch& sIn = sInB[buf];

    float acc[2][16];
#pragma unroll
    for (int j = 0; j < 16; ++j) { acc[0][j] = sB[j]; acc[1][j] = sB[j]; }
#pragma unroll
    for (int c = 0; c < 3; ++c)
#pragma unroll
      for (int ky = 0; ky < 3; ++ky)
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
          const float v0 = sIn[c][2 * ly + ky][2 * lx + kx + 3];
          const float v1 = sIn[c][2 * (ly + 8) + ky][2 * lx + kx + 3];
          const int k = c * 9 + ky * 3 + kx;
#pragma unroll
          for (int j4 = 0; j4 < 4; ++j4) {
            const float4 w = *reinterpret_cast<const float4*>(&sW[k][j4 * 4]);
            fma2(acc[0][j4 * 4 + 0], acc[0][j4 * 4 + 1], v0, v0, w.x, w.y);
            fma2(acc[0][j4 * 4 + 2], acc[0][j4 * 4 + 3], v0, v0, w.z, w.w);
            fma2(acc[1][j4 * 4 + 0], acc[1][j4 * 4 + 1], v1, v1, w.x, w.y);
            fma2(acc[1][j4 * 4 + 2], acc[1][j4 * 4 + 3], v1, v1, w.z, w.w);
          }
        }
    float vals[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) vals[j] = 0.f;
    const int ox = ox0 + lx;
#pragma unroll
    for (int p = 0; p < 2; ++p) {
      const int oy = oy0 + ly + p * 8;
      if (oy < Ho && ox < Wo) {
        float* dst = a.zout + (((long long)b * Ho + oy) * Wo + ox) * 16;
#pragma unroll
        for (int j4 = 0; j4 < 4; ++j4)
          *reinterpret_cast<float4*>(dst + j4 * 4) =
              make_float4(acc[p][j4 * 4], acc[p][j4 * 4 + 1], acc[p][j4 * 4 + 2], acc[p][j4 * 4 + 3]);
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          vals[j] += acc[p][j];
          vals[16 + j] = fmaf(acc[p][j], acc[p][j], vals[16 + j]);
        }
      }
    }
    if (a.osum != nullptr) {
      // butterfly transpose-reduce: 31 shuffles leave the warp total of value index L on lane L
#pragma unroll
      for (int s = 16; s >= 1; s >>= 1) {
        const bool upper = (lane & s) != 0;
#pragma unroll
        for (int i = 0; i < s; ++i) {
          const float send = upper ? vals[i] : vals[i + s];
          const float keep = upper ? vals[i + s] : vals[i];
          vals[i] = keep + __shfl_xor_sync(0xffffffffu, send, s);
        }
      }
      stat += (double)vals[0];
    }
  }
  if (a.osum != nullptr) {
    sRedD[warp][lane] = stat;
    __syncthreads();
    if (tid < 32) {
      double s = 0.0;
      for (int w = 0; w < 8; ++w) s += sRedD[w][tid];
      atomicAdd((tid < 16 ? a.osum : a.osumsq) + (tid & 15), s);
    }
  }
}

// ------------------------------------------------------------------------------- BN running stats
__global__ void bn_update_running_kernel(const BnFinalizeArgs a, const double* sum,
                                         const double* sumsq, float* rmean, float* rvar,
                                         float momentum) {
  const int i = blockIdx.x;
  const int c = threadIdx.x;
  if (i >= a.n || c >= a.C[i]) return;
  const long long o = a.ch_off[i] + c;
  const double n = a.count[i];
  const double m = sum[o] / n;
  double v = sumsq[o] / n - m * m;
  if (v < 0.0) v = 0.0;
  const double unbiased = n > 1.0 ? v * n / (n - 1.0) : v;
  rmean[o] = (1.f - momentum) * rmean[o] + momentum * (float)m;
  rvar[o] = (1.f - momentum) * rvar[o] + momentum * (float)unbiased;
}

// ------------------------------------------------------------------------------- read activation
__global__ void read_activation_kernel(const float* z, const BnRef bn, int has_bn, int B, int H,
                                       int W, int C, long long batch_stride, float* out) {
  const long long total = (long long)B * C * H * W;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int x = i % W;
    const int y = (i / W) % H;
    const int c = (i / ((long long)W * H)) % C;
    const int b = i / ((long long)W * H * C);
    float v = z[(long long)b * batch_stride + ((long long)y * W + x) * C + c];
    if (has_bn) {
      float sc, sh;
      bn_coeffs(bn, c, sc, sh);
      v = fmaxf(fmaf(v, sc, sh), 0.f);
    }
    out[i] = v;
  }
}

__global__ void grid_priors_kernel(float* priors, int h0, int w0, int s0, int h1, int w1, int s1,
                                   int h2, int w2, int s2) {
  const int n0 = h0 * w0, n1 = h1 * w1, n2 = h2 * w2;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n0 + n1 + n2) return;
  int j = i, w = w0, s = s0;
  if (i >= n0 + n1) { j = i - n0 - n1; w = w2; s = s2; }
  else if (i >= n0) { j = i - n0; w = w1; s = s1; }
  priors[i * 4 + 0] = (float)((j % w) * s);
  priors[i * 4 + 1] = (float)((j / w) * s);
  priors[i * 4 + 2] = (float)s;
  priors[i * 4 + 3] = (float)s;
}

template <int CIN, int COUT, int MODE>
cudaError_t launch_unit_fwd_t(const UnitFwdArgs& a, int num_sms, cudaStream_t s) {
  using C = FwdCfg<CIN, COUT>;
  constexpr bool PF = MODE == 0 && (C::SMEM_FLOATS + C::HP * CIN) * 4 <= 112 * 1024;      // as in the kernel
  const size_t smem = sizeof(float) * (C::SMEM_FLOATS + (PF ? C::HP * CIN : 0));
  auto kern = unit_fwd_kernel<CIN, COUT, MODE>;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    configured = true;
  }
  const int tiles = ((a.W + C::TW - 1) / C::TW) * ((a.H + C::TH - 1) / C::TH) * a.B;
  int grid = 2 * num_sms;            // persistent: __launch_bounds__(NT, 2)
  if (grid > tiles) grid = tiles;
  kern<<<grid, NT, smem, s>>>(a);
  return cudaGetLastError();
}

template <int CIN, int COUT>
cudaError_t launch_unit_fwd_m(int mode, const UnitFwdArgs& a, int num_sms, cudaStream_t s) {
  switch (mode) {
    case 0: return launch_unit_fwd_t<CIN, COUT, 0>(a, num_sms, s);
    case 1: return launch_unit_fwd_t<CIN, COUT, 1>(a, num_sms, s);
    case 2: return launch_unit_fwd_t<CIN, COUT, 2>(a, num_sms, s);
  }
  return cudaErrorInvalidValue;
}

}  // namespace

int unit_fwd_supported(int cin, int cout) {
  return (cin == 16 && (cout == 16 || cout == 32 || cout == 64)) ||
         (cin == 32 && (cout == 32 || cout == 64)) || (cin == 64 && (cout == 64 || cout == 16));
}

cudaError_t launch_unit_fwd(int cin, int cout, int mode, const UnitFwdArgs& a, int num_sms,
                            cudaStream_t s) {
  if (cin == 16 && cout == 16) return launch_unit_fwd_m<16, 16>(mode, a, num_sms, s);
  if (cin == 16 && cout == 32) return launch_unit_fwd_m<16, 32>(mode, a, num_sms, s);
  if (cin == 16 && cout == 64) return launch_unit_fwd_m<16, 64>(mode, a, num_sms, s);
  if (cin == 32 && cout == 32) return launch_unit_fwd_m<32, 32>(mode, a, num_sms, s);
  if (cin == 32 && cout == 64) return launch_unit_fwd_m<32, 64>(mode, a, num_sms, s);
  if (cin == 64 && cout == 64) return launch_unit_fwd_m<64, 64>(mode, a, num_sms, s);
  if (cin == 64 && cout == 16) return launch_unit_fwd_m<64, 16>(mode, a, num_sms, s);
  return cudaErrorInvalidValue;
}

cudaError_t launch_stem_fwd(const StemArgs& a, int num_sms, cudaStream_t s) {
  if (a.Win & 3) return cudaErrorInvalidValue;     // float4 staging of the image rows
  const int Ho = a.Hin / 2, Wo = a.Win / 2;
  const int tiles = ((Wo + ST_TW - 1) / ST_TW) * ((Ho + ST_TH - 1) / ST_TH) * a.B;
  int grid = 2 * num_sms;
  if (grid > tiles) grid = tiles;
  const int smem = 2 * 3 * ST_IH * ST_IWP * (int)sizeof(float);
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(stem_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return e;
    configured = true;
  }
  stem_fwd_kernel<<<grid, 256, smem, s>>>(a);
  return cudaGetLastError();
}

cudaError_t launch_bn_update_running(const BnFinalizeArgs& a, const double* sum,
                                     const double* sumsq, float* rmean, float* rvar,
                                     float momentum, cudaStream_t s) {
  bn_update_running_kernel<<<a.n, 64, 0, s>>>(a, sum, sumsq, rmean, rvar, momentum);
  return cudaGetLastError();
}

cudaError_t launch_read_activation(const float* z, const BnRef& bn, int has_bn, int B, int H,
                                   int W, int C, long long batch_stride, float* out_nchw,
                                   cudaStream_t s) {
  const long long total = (long long)B * C * H * W;
  int blocks = (int)((total + 255) / 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  if (blocks < 1) blocks = 1;
  read_activation_kernel<<<blocks, 256, 0, s>>>(z, bn, has_bn, B, H, W, C, batch_stride, out_nchw);
  return cudaGetLastError();
}

cudaError_t launch_grid_priors(float* priors, const int* lh, const int* lw, const int* strides,
                               cudaStream_t s) {
  const int n = lh[0] * lw[0] + lh[1] * lw[1] + lh[2] * lw[2];
  grid_priors_kernel<<<(n + 255) / 256, 256, 0, s>>>(priors, lh[0], lw[0], strides[0], lh[1], lw[1],
                                                     strides[1], lh[2], lw[2], strides[2]);
  return cudaGetLastError();
}

}  // namespace yunet
