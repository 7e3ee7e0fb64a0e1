// Backward kernels of the YuNet hot path, fp32 NHWC, sm_100a.
//
// One fused kernel per ConvDPUnit (autograd of mmdet/models/utils/yunet_layer.py:30-36 plus the
// BatchNorm/ReLU/max-pool/upsample-add that the forward folded into the load prologues).
// Persistent CTAs walk the pixel tiles; per tile:
//   load   g   = BN-backward of the output gradient on the halo tile
//              = gamma*rstd*(du - mean(du) - zhat*mean(du*zhat))            (has_bn)
//          a   = prologue(z_in) on the interior tile (recomputed, never stored)
//   GEMM1  y   = a W1^T + b1                     (recomputed pointwise output)
//   dw-bwd dy  = sum_k W2[k] * g[q - d_k];  dW2 += y * g[q - d_k];  db2 += g;  db1 += dy
//   GEMM2  h   = dy W1                           (gradient wrt a)
//   GEMM3  dW1 += dy^T a
//   epilogue   du_in = h * relu-mask, routed through max-pool argmax / summed over the 2x2
//              children for the upsampled operand; written once (or accumulated when the input
//              has an earlier consumer in backward order) + sum(du_in), sum(du_in*zhat_in) for the
//              producer's own BN backward.
// Parameter gradients are accumulated in registers across all tiles of a CTA and flushed with one
// round of atomics at the end.
#include <cstdio>
#include <cstdlib>

#include "kernels.h"
#include "prefetch.cuh"
#include "f32x2.cuh"

namespace yunet {

namespace {

constexpr int NT = 256;

struct Coef {  // per-channel BN constants of an input tensor
  float scale, shift, mean, rstd;
};

__device__ __forceinline__ Coef bn_coef(const BnRef& r, int c) {
  Coef k;
  double dm = r.sum[c] * r.inv_count;
  double dv = r.sumsq[c] * r.inv_count - dm * dm;
  if (dv < 0.0) dv = 0.0;
  k.mean = (float)dm;
  k.rstd = 1.0f / sqrtf((float)dv + kBnEps);
  k.scale = r.gamma[c] * k.rstd;
  k.shift = r.beta[c] - k.mean * k.scale;
  return k;
}

__device__ __forceinline__ float4 ldg4(const float* p) {
  return __ldg(reinterpret_cast<const float4*>(p));
}
__device__ __forceinline__ float4 lds4(const float* p) {
  return *reinterpret_cast<const float4*>(p);
}
__device__ __forceinline__ float2 lds2(const float* p) {
  return *reinterpret_cast<const float2*>(p);
}
__device__ __forceinline__ void sts4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
__device__ __forceinline__ float4 f4(float v) { return make_float4(v, v, v, v); }
__device__ __forceinline__ float4 add4(float4 a, float4 b) {
  return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w);
}
__device__ __forceinline__ float4 mul4(float4 a, float4 b) {
  return make_float4(a.x * b.x, a.y * b.y, a.z * b.z, a.w * b.w);
}
__device__ __forceinline__ void fma4(float4& acc, float4 a, float4 b) { fma4p(acc, a, b); }
__device__ __forceinline__ float4 bnu4(float4 z, float4 sc, float4 sh) {  // u = z*scale+shift
  return make_float4(fmaf(z.x, sc.x, sh.x), fmaf(z.y, sc.y, sh.y), fmaf(z.z, sc.z, sh.z),
                     fmaf(z.w, sc.w, sh.w));
}
__device__ __forceinline__ float4 relu4(float4 u) {
  return make_float4(fmaxf(u.x, 0.f), fmaxf(u.y, 0.f), fmaxf(u.z, 0.f), fmaxf(u.w, 0.f));
}
__device__ __forceinline__ float comp(const float4& v, int i) {
  return i == 0 ? v.x : i == 1 ? v.y : i == 2 ? v.z : v.w;
}

template <int CIN, int COUT, int PAIRM = 1>
struct BwdCfg {
  // interior tile 8 x 16; the 16 -> 16 units (160^2 / 80^2 layers) take 16 x 16 so that the
  // per-tile barriers and load latency are amortised over twice the pixels at equal occupancy
  static constexpr bool BIG = (CIN <= 16 && COUT <= 16);
  static constexpr int TH = BIG ? 16 : 8, TW = 16;
  static constexpr int HH = TH + 2, HW = TW + 2;
  static constexpr int HP = HH * HW;   // 180 (324) halo pixels
  static constexpr int TP = TH * TW;   // 128 (256) interior pixels
  static constexpr int AS = CIN + 4;
  static constexpr int YS = COUT + 4;
  // GEMM1: N = COUT
  static constexpr int CPT1 = (COUT == 16) ? 4 : 8;
  static constexpr int NCG1 = COUT / CPT1;
  static constexpr int NPG1 = NT / NCG1;
  static constexpr int PPT1 = TP / NPG1;
  // GEMM2: N = CIN
  static constexpr int CPT2 = (CIN == 16) ? 4 : 8;
  static constexpr int NCG2 = CIN / CPT2;
  static constexpr int NPG2 = NT / NCG2;
  static constexpr int PPT2 = TP / NPG2;
  // GEMM3: 4x4 output blocks, pixel-split groups
  static constexpr int NOUT = (COUT / 4) * (CIN / 4);
  static constexpr int G3 = NT / NOUT;
  // depthwise stage: thread = (channel quad, column) x RPT rows; for COUT <= 32 a thread takes a channel
  // PAIR instead (twice the rows): 22 persistent + 40 transient registers instead of 44 + 76, which is what
  // lets the two-CTAs-per-SM build (128 registers) run without spilling the gradient accumulators
  // (the builds with CIN * COUT <= 1024; for COUT = 64 the pair threads walk the 16 columns in two passes)
  static constexpr bool PAIR = PAIRM != 0 && CIN * COUT <= 1024;
  static constexpr int NQ = PAIR ? COUT / 2 : COUT / 4;
  static constexpr int TWP = (NQ * TW <= NT) ? TW : NT / NQ;   // columns per pass
  static constexpr int CPASS = TW / TWP;
  static constexpr int RG = NT / (NQ * TWP);
  static constexpr int RPT = TH / RG;
  static_assert(RG >= 1 && TH % RG == 0 && TW % TWP == 0 && (PAIR || CPASS == 1), "depthwise mapping");
  // epilogue
  static constexpr int QI = CIN / 4;
  static constexpr int EPT = TP * QI / NT;          // items per thread (hi-res)
  static constexpr int EPTB = (TP / 4) * QI / NT;   // items per thread (lo-res, UPADD), may be 0
  static constexpr int SMEM_FLOATS = HP * COUT + TP * AS + TP * YS + 2 * CIN * COUT + 9 * COUT +
                                     COUT + 8 * CIN + 5 * COUT;
  static_assert(NOUT <= NT && NT % NOUT == 0, "gemm3 mapping");
  static_assert(TP % NPG1 == 0 && TP % NPG2 == 0, "gemm mapping");
};

template <int CIN, int COUT, int MODE, int HAS_BN, int OCC, int PAIRM>
__global__ void __launch_bounds__(NT, OCC) unit_bwd_kernel(const UnitBwdArgs a) {
  using C = BwdCfg<CIN, COUT, PAIRM>;
  extern __shared__ float4 smem_raw[];
  float* smem = reinterpret_cast<float*>(smem_raw);
  float* sG = smem;                        // [C::HP][COUT]
  float* sA = sG + C::HP * COUT;              // [C::TP][AS]     (later h)
  float* sY = sA + C::TP * C::AS;             // [C::TP][YS]     (y, then dy)
  float* sW1 = sY + C::TP * C::YS;            // [COUT][CIN]
  float* sW1t = sW1 + CIN * COUT;          // [CIN][COUT]
  float* sW2 = sW1t + CIN * COUT;          // [9][COUT]
  float* sB1 = sW2 + 9 * COUT;             // [COUT]
  float* sCa = sB1 + COUT;                 // [4][CIN]  scale, shift, mean, rstd of input a
  float* sCb = sCa + 4 * CIN;              // [4][CIN]  ... of input b
  float* sCo = sCb + 4 * CIN;              // [5][COUT] gscale, m1, m2, mean, rstd of the output BN
  // PF: the next tile's du / z_out halo rows are copied in with cp.async while this tile is in its
  // GEMM / depthwise / epilogue phases (the staging loads were ~25 % of the stall samples); only where
  // the extra buffers leave room for the CTAs per SM the kernel is built for
  constexpr int RAWF = (HAS_BN ? 2 : 1) * C::HP * COUT;
  constexpr bool PF = OCC == 2 && (C::SMEM_FLOATS + RAWF) * 4 <= 112 * 1024;
  float* sRawD = sCo + 5 * COUT;           // [C::HP][COUT] raw du      (PF)
  float* sRawZ = sRawD + C::HP * COUT;     // [C::HP][COUT] raw z_out   (PF, HAS_BN)

  const int tid = threadIdx.x;
  // ---- one-time setup
  for (int i = tid; i < CIN * COUT; i += NT) {
    sW1[i] = __ldg(a.w1 + i);
    int ci = i / COUT, co = i % COUT;
    sW1t[i] = __ldg(a.w1 + co * CIN + ci);
  }
  for (int i = tid; i < 9 * COUT; i += NT) {
    int k = i / COUT, co = i % COUT;
    sW2[i] = __ldg(a.w2 + co * 9 + k);
  }
  if (tid < COUT) {
    sB1[tid] = __ldg(a.b1 + tid);
    if (HAS_BN) {
      Coef k = bn_coef(a.bno, tid);
      sCo[0 * COUT + tid] = a.bno.gamma[tid] * k.rstd;
      sCo[1 * COUT + tid] = (float)(a.dsum[tid] * a.bno.inv_count);
      sCo[2 * COUT + tid] = (float)(a.dsumzh[tid] * a.bno.inv_count);
      sCo[3 * COUT + tid] = k.mean;
      sCo[4 * COUT + tid] = k.rstd;
    }
  }
  if (tid < CIN) {
    Coef k = bn_coef(a.bna, tid);
    sCa[0 * CIN + tid] = k.scale; sCa[1 * CIN + tid] = k.shift;
    sCa[2 * CIN + tid] = k.mean;  sCa[3 * CIN + tid] = k.rstd;
    if (MODE == 2) {
      Coef kb = bn_coef(a.bnb, tid);
      sCb[0 * CIN + tid] = kb.scale; sCb[1 * CIN + tid] = kb.shift;
      sCb[2 * CIN + tid] = kb.mean;  sCb[3 * CIN + tid] = kb.rstd;
    }
  }
  __syncthreads();

  // ---- persistent accumulators
  // depthwise-stage mapping
  const int dq = tid % C::NQ;
  const int dx0 = (tid / C::NQ) % C::TWP;
  const int dr0 = (tid / (C::NQ * C::TWP)) * C::RPT;
  float4 gw2[C::PAIR ? 1 : 9];
  float2 pw2[C::PAIR ? 9 : 1];      // PAIR: the channel pair's nine tap gradients
#pragma unroll
  for (int k = 0; k < (C::PAIR ? 1 : 9); ++k) gw2[k] = f4(0.f);
#pragma unroll
  for (int k = 0; k < (C::PAIR ? 9 : 1); ++k) pw2[k] = make_float2(0.f, 0.f);
  float4 gb2 = f4(0.f), gb1 = f4(0.f);   // PAIR: .x/.y used
  // GEMM3 mapping
  const int o3 = tid % C::NOUT;
  const int g3 = tid / C::NOUT;
  const int co3 = (o3 / (CIN / 4)) * 4;
  const int ci3 = (o3 % (CIN / 4)) * 4;
  float gw1[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) gw1[i][j] = 0.f;
  // epilogue mapping / statistics of the input gradients
  const int eq = tid % C::QI;
  float4 sa1 = f4(0.f), sa2 = f4(0.f), sb1 = f4(0.f), sb2 = f4(0.f);

  const int tiles_x = (a.W + C::TW - 1) / C::TW;
  const int tiles_y = (a.H + C::TH - 1) / C::TH;
  const int ntiles = tiles_x * tiles_y * a.B;
  const long long in_img_stride = (MODE == 1) ? (long long)a.H * a.W * 4 * CIN : (long long)a.H * a.W * CIN;

  auto stage_g = [&](int tile_) {
    int t = tile_;
    const int tx = t % tiles_x; t /= tiles_x;
    const int ty = t % tiles_y;
    const int b = t / tiles_y;
    const int x0 = tx * C::TW, y0 = ty * C::TH;
    constexpr int Q = COUT / 4;
    const float* dimg = a.dout + (long long)b * a.dout_batch_stride;
    const float* zimg = HAS_BN ? a.zout + (long long)b * a.H * a.W * COUT : nullptr;
#pragma unroll 4
    for (int it = 0; it < (C::HP * Q + NT - 1) / NT; ++it) {
      const int i = tid + it * NT;
      if (i >= C::HP * Q) break;
      const int pix = i / Q, q = i % Q;
      const int gy = y0 + pix / C::HW - 1, gx = x0 + pix % C::HW - 1;
      const bool in = gy >= 0 && gy < a.H && gx >= 0 && gx < a.W;
      const long long off = in ? ((long long)gy * a.W + gx) * COUT + q * 4 : 0;
      cp_async16(sRawD + pix * COUT + q * 4, dimg + off, in);
      if (HAS_BN) cp_async16(sRawZ + pix * COUT + q * 4, zimg + off, in);
    }
    cp_async_commit();
  };
  if (PF && (int)blockIdx.x < ntiles) stage_g(blockIdx.x);

  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    int t = tile;
    const int tx = t % tiles_x; t /= tiles_x;
    const int ty = t % tiles_y;
    const int b = t / tiles_y;
    const int x0 = tx * C::TW, y0 = ty * C::TH;
    const float* za_img = a.za + (long long)b * in_img_stride;

    // ---- L2 prefetch of the next tile's inputs (one bulk request per tile row)
    {
      const int nt = tile + gridDim.x;
      if (nt < ntiles && tid < 3 * 64) {
        int t2 = nt;
        const int ntx = t2 % tiles_x; t2 /= tiles_x;
        const int nty = t2 % tiles_y;
        const int nb = t2 / tiles_y;
        const int nx0 = ntx * C::TW, ny0 = nty * C::TH;
        const int which = tid >> 6, r = tid & 63;
        if (which == 0) {
          if (!PF)
            l2_prefetch_tile<COUT>(a.dout + (long long)nb * a.dout_batch_stride, a.H, a.W, ny0 - 1,
                                   ny0 + C::TH + 1, nx0 - 1, nx0 + C::TW + 1, r);
        } else if (which == 1) {
          if (HAS_BN && !PF)
            l2_prefetch_tile<COUT>(a.zout + (long long)nb * a.H * a.W * COUT, a.H, a.W, ny0 - 1,
                                   ny0 + C::TH + 1, nx0 - 1, nx0 + C::TW + 1, r);
        } else if (MODE == 1) {
          // pooled operand: 4x the rows at twice the width -- measured slower with the prefetch
        } else {
          l2_prefetch_tile<CIN>(a.za + (long long)nb * in_img_stride, a.H, a.W, ny0, ny0 + C::TH,
                                nx0, nx0 + C::TW, r);
          if (MODE == 2)
            l2_prefetch_tile<CIN>(a.zb + (long long)nb * (a.H >> 1) * (a.W >> 1) * CIN, a.H >> 1,
                                  a.W >> 1, ny0 >> 1, (ny0 + C::TH) >> 1, nx0 >> 1,
                                  (nx0 + C::TW) >> 1, r - 32);
        }
      }
    }

    // ---- S0a: g on the halo tile
    {
      constexpr int Q = COUT / 4;
      const float* dimg = a.dout + (long long)b * a.dout_batch_stride;
      const float* zimg = HAS_BN ? a.zout + (long long)b * a.H * a.W * COUT : nullptr;
      if (PF) cp_async_wait<0>();       // a thread reads back exactly the chunks it copied
      // NT % Q == 0: a thread's items all belong to channel quad tid % Q -- its BN-backward constants
      // are read once per tile instead of once per item
      static_assert(NT % Q == 0, "channel quad of a thread's items");
      const int q = tid % Q;
      float4 gs, m1, m2, mu, rs;
      if (HAS_BN) {
        gs = lds4(sCo + 0 * COUT + q * 4); m1 = lds4(sCo + 1 * COUT + q * 4);
        m2 = lds4(sCo + 2 * COUT + q * 4); mu = lds4(sCo + 3 * COUT + q * 4);
        rs = lds4(sCo + 4 * COUT + q * 4);
      }
#pragma unroll 4
      for (int it = 0; it < (C::HP * Q + NT - 1) / NT; ++it) {
        const int i = tid + it * NT;
        if (i >= C::HP * Q) break;
        const int pix = i / Q;
        const int gy = y0 + pix / C::HW - 1, gx = x0 + pix % C::HW - 1;
        float4 g = f4(0.f);
        if (gy >= 0 && gy < a.H && gx >= 0 && gx < a.W) {
          const long long off = ((long long)gy * a.W + gx) * COUT + q * 4;
          g = PF ? lds4(sRawD + pix * COUT + q * 4) : ldg4(dimg + off);
          if (HAS_BN) {
            const float4 z = PF ? lds4(sRawZ + pix * COUT + q * 4) : ldg4(zimg + off);
            g.x = gs.x * (g.x - m1.x - (z.x - mu.x) * rs.x * m2.x);
            g.y = gs.y * (g.y - m1.y - (z.y - mu.y) * rs.y * m2.y);
            g.z = gs.z * (g.z - m1.z - (z.z - mu.z) * rs.z * m2.z);
            g.w = gs.w * (g.w - m1.w - (z.w - mu.w) * rs.w * m2.w);
          }
        }
        sts4(sG + pix * COUT + q * 4, g);
      }
    }
    // ---- S0b: activated input on the interior tile
    {
      constexpr int Q = CIN / 4;
#pragma unroll 4
      for (int it = 0; it < C::TP * Q / NT; ++it) {
        const int i = tid + it * NT;
        const int pix = i / Q, q = i % Q;
        const int gy = y0 + pix / C::TW, gx = x0 + pix % C::TW;
        float4 v = f4(0.f);
        if (gy < a.H && gx < a.W) {
          const float4 sc = lds4(sCa + q * 4), sh = lds4(sCa + CIN + q * 4);
          if (MODE == 0) {
            v = relu4(bnu4(ldg4(za_img + ((long long)gy * a.W + gx) * CIN + q * 4), sc, sh));
          } else if (MODE == 1) {
            const int W2 = a.W * 2;
            const float* p = za_img + ((long long)(gy * 2) * W2 + gx * 2) * CIN + q * 4;
            float4 v00 = relu4(bnu4(ldg4(p), sc, sh));
            float4 v01 = relu4(bnu4(ldg4(p + CIN), sc, sh));
            float4 v10 = relu4(bnu4(ldg4(p + (long long)W2 * CIN), sc, sh));
            float4 v11 = relu4(bnu4(ldg4(p + (long long)W2 * CIN + CIN), sc, sh));
            v.x = fmaxf(fmaxf(v00.x, v01.x), fmaxf(v10.x, v11.x));
            v.y = fmaxf(fmaxf(v00.y, v01.y), fmaxf(v10.y, v11.y));
            v.z = fmaxf(fmaxf(v00.z, v01.z), fmaxf(v10.z, v11.z));
            v.w = fmaxf(fmaxf(v00.w, v01.w), fmaxf(v10.w, v11.w));
          } else {
            v = relu4(bnu4(ldg4(za_img + ((long long)gy * a.W + gx) * CIN + q * 4), sc, sh));
            const int Hb = a.H >> 1, Wb = a.W >> 1;
            const float* pb = a.zb + (((long long)b * Hb + (gy >> 1)) * Wb + (gx >> 1)) * CIN + q * 4;
            const float4 scb = lds4(sCb + q * 4), shb = lds4(sCb + CIN + q * 4);
            v = add4(v, relu4(bnu4(ldg4(pb), scb, shb)));
          }
        }
        sts4(sA + pix * C::AS + q * 4, v);
      }
    }
    __syncthreads();
    if (PF && tile + (int)gridDim.x < ntiles) stage_g(tile + gridDim.x);

    // ---- S1: GEMM1  y = a W1^T + b1  (zero for out-of-image pixels)
    {
      const int cg = tid % C::NCG1, pg = tid / C::NCG1;
      float acc[C::PPT1][C::CPT1];
#pragma unroll
      for (int i = 0; i < C::PPT1; ++i)
#pragma unroll
        for (int j = 0; j < C::CPT1; ++j) acc[i][j] = 0.f;
#pragma unroll 2
      for (int k4 = 0; k4 < CIN / 4; ++k4) {
        float4 av[C::PPT1];
#pragma unroll
        for (int i = 0; i < C::PPT1; ++i) av[i] = lds4(sA + (pg + i * C::NPG1) * C::AS + k4 * 4);
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
          float wv[C::CPT1];
#pragma unroll
          for (int j4 = 0; j4 < C::CPT1 / 4; ++j4) {
            const float4 w = lds4(sW1t + (k4 * 4 + kk) * COUT + cg * C::CPT1 + j4 * 4);
            wv[j4 * 4] = w.x; wv[j4 * 4 + 1] = w.y; wv[j4 * 4 + 2] = w.z; wv[j4 * 4 + 3] = w.w;
          }
#pragma unroll
          for (int i = 0; i < C::PPT1; ++i) {
            const float ak = comp(av[i], kk);
#pragma unroll
            for (int j = 0; j < C::CPT1; j += 2) fma2(acc[i][j], acc[i][j + 1], ak, ak, wv[j], wv[j + 1]);
          }
        }
      }
#pragma unroll
      for (int i = 0; i < C::PPT1; ++i) {
        const int pix = pg + i * C::NPG1;
        const int gy = y0 + pix / C::TW, gx = x0 + pix % C::TW;
        const bool in = gy < a.H && gx < a.W;
#pragma unroll
        for (int j4 = 0; j4 < C::CPT1 / 4; ++j4) {
          const float* bb = sB1 + cg * C::CPT1 + j4 * 4;
          float4 o;
          o.x = in ? acc[i][j4 * 4 + 0] + bb[0] : 0.f;
          o.y = in ? acc[i][j4 * 4 + 1] + bb[1] : 0.f;
          o.z = in ? acc[i][j4 * 4 + 2] + bb[2] : 0.f;
          o.w = in ? acc[i][j4 * 4 + 3] + bb[3] : 0.f;
          sts4(sY + pix * C::YS + cg * C::CPT1 + j4 * 4, o);
        }
      }
    }
    __syncthreads();

    // ---- S2: depthwise backward: dy (in place over y), dW2, db2, db1
    if constexpr (C::PAIR) {
      float2 w2r[9];
#pragma unroll
      for (int k = 0; k < 9; ++k) w2r[k] = lds2(sW2 + k * COUT + dq * 2);
#pragma unroll 1
      for (int cp = 0; cp < C::CPASS; ++cp) {
      const int dx = dx0 + cp * C::TWP;
      float2 ra[3], rb[3], rc[3];
#pragma unroll
      for (int d = 0; d < 3; ++d) {
        ra[d] = lds2(sG + ((dr0 + 0) * C::HW + dx + d) * COUT + dq * 2);
        rb[d] = lds2(sG + ((dr0 + 1) * C::HW + dx + d) * COUT + dq * 2);
      }
      const int gx = x0 + dx;
#pragma unroll
      for (int i = 0; i < C::RPT; ++i) {
#pragma unroll
        for (int d = 0; d < 3; ++d)
          rc[d] = lds2(sG + ((dr0 + i + 2) * C::HW + dx + d) * COUT + dq * 2);
        const int r = dr0 + i;
        const int pix = r * C::TW + dx;
        const bool in = (y0 + r) < a.H && gx < a.W;
        const float2 y = lds2(sY + pix * C::YS + dq * 2);
        float2 dy = make_float2(0.f, 0.f);
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
          fma2(dy.x, dy.y, w2r[0 * 3 + kx].x, w2r[0 * 3 + kx].y, rc[2 - kx].x, rc[2 - kx].y);
          fma2(dy.x, dy.y, w2r[1 * 3 + kx].x, w2r[1 * 3 + kx].y, rb[2 - kx].x, rb[2 - kx].y);
          fma2(dy.x, dy.y, w2r[2 * 3 + kx].x, w2r[2 * 3 + kx].y, ra[2 - kx].x, ra[2 - kx].y);
          fma2(pw2[0 * 3 + kx].x, pw2[0 * 3 + kx].y, y.x, y.y, rc[2 - kx].x, rc[2 - kx].y);
          fma2(pw2[1 * 3 + kx].x, pw2[1 * 3 + kx].y, y.x, y.y, rb[2 - kx].x, rb[2 - kx].y);
          fma2(pw2[2 * 3 + kx].x, pw2[2 * 3 + kx].y, y.x, y.y, ra[2 - kx].x, ra[2 - kx].y);
        }
        gb2.x += rb[1].x; gb2.y += rb[1].y;
        if (!in) dy = make_float2(0.f, 0.f);
        gb1.x += dy.x; gb1.y += dy.y;
        *reinterpret_cast<float2*>(sY + pix * C::YS + dq * 2) = dy;
#pragma unroll
        for (int d = 0; d < 3; ++d) { ra[d] = rb[d]; rb[d] = rc[d]; }
      }
      }
    } else {
      const int dx = dx0;
      float4 w2r[9];
#pragma unroll
      for (int k = 0; k < 9; ++k) w2r[k] = lds4(sW2 + k * COUT + dq * 4);
      float4 ra[3], rb[3], rc[3];
#pragma unroll
      for (int d = 0; d < 3; ++d) {
        ra[d] = lds4(sG + ((dr0 + 0) * C::HW + dx + d) * COUT + dq * 4);
        rb[d] = lds4(sG + ((dr0 + 1) * C::HW + dx + d) * COUT + dq * 4);
      }
      const int gx = x0 + dx;
#pragma unroll
      for (int i = 0; i < C::RPT; ++i) {
#pragma unroll
        for (int d = 0; d < 3; ++d)
          rc[d] = lds4(sG + ((dr0 + i + 2) * C::HW + dx + d) * COUT + dq * 4);
        const int r = dr0 + i;
        const int pix = r * C::TW + dx;
        const bool in = (y0 + r) < a.H && gx < a.W;
        const float4 y = lds4(sY + pix * C::YS + dq * 4);
        float4 dy = f4(0.f);
        // tap (ky,kx) pairs with window element [2-ky][2-kx]; window rows: ra (0), rb (1), rc (2)
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
          fma4(dy, w2r[0 * 3 + kx], rc[2 - kx]);
          fma4(dy, w2r[1 * 3 + kx], rb[2 - kx]);
          fma4(dy, w2r[2 * 3 + kx], ra[2 - kx]);
          fma4(gw2[0 * 3 + kx], y, rc[2 - kx]);
          fma4(gw2[1 * 3 + kx], y, rb[2 - kx]);
          fma4(gw2[2 * 3 + kx], y, ra[2 - kx]);
        }
        gb2 = add4(gb2, rb[1]);
        if (!in) dy = f4(0.f);
        gb1 = add4(gb1, dy);
        sts4(sY + pix * C::YS + dq * 4, dy);
#pragma unroll
        for (int d = 0; d < 3; ++d) { ra[d] = rb[d]; rb[d] = rc[d]; }
      }
    }
    __syncthreads();

    // ---- S3: GEMM2  h = dy W1 ;  GEMM3  dW1 += dy^T a
    float hacc[C::PPT2][C::CPT2];
    const int cg2 = tid % C::NCG2, pg2 = tid / C::NCG2;
    {
#pragma unroll
      for (int i = 0; i < C::PPT2; ++i)
#pragma unroll
        for (int j = 0; j < C::CPT2; ++j) hacc[i][j] = 0.f;
#pragma unroll 2
      for (int k4 = 0; k4 < COUT / 4; ++k4) {
        float4 dv[C::PPT2];
#pragma unroll
        for (int i = 0; i < C::PPT2; ++i) dv[i] = lds4(sY + (pg2 + i * C::NPG2) * C::YS + k4 * 4);
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
          float wv[C::CPT2];
#pragma unroll
          for (int j4 = 0; j4 < C::CPT2 / 4; ++j4) {
            const float4 w = lds4(sW1 + (k4 * 4 + kk) * CIN + cg2 * C::CPT2 + j4 * 4);
            wv[j4 * 4] = w.x; wv[j4 * 4 + 1] = w.y; wv[j4 * 4 + 2] = w.z; wv[j4 * 4 + 3] = w.w;
          }
#pragma unroll
          for (int i = 0; i < C::PPT2; ++i) {
            const float dk = comp(dv[i], kk);
#pragma unroll
            for (int j = 0; j < C::CPT2; j += 2) fma2(hacc[i][j], hacc[i][j + 1], dk, dk, wv[j], wv[j + 1]);
          }
        }
      }
      // GEMM3: pixels g3, g3+G3, ...
#pragma unroll 4
      for (int p = g3; p < C::TP; p += C::G3) {
        const float4 d4 = lds4(sY + p * C::YS + co3);
        const float4 a4 = lds4(sA + p * C::AS + ci3);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float di = comp(d4, i);
          fma2(gw1[i][0], gw1[i][1], di, di, a4.x, a4.y);
          fma2(gw1[i][2], gw1[i][3], di, di, a4.z, a4.w);
        }
      }
    }
    __syncthreads();
    // ---- S4: h -> shared (over sA)
#pragma unroll
    for (int i = 0; i < C::PPT2; ++i) {
      const int pix = pg2 + i * C::NPG2;
#pragma unroll
      for (int j4 = 0; j4 < C::CPT2 / 4; ++j4)
        sts4(sA + pix * C::AS + cg2 * C::CPT2 + j4 * 4,
             make_float4(hacc[i][j4 * 4], hacc[i][j4 * 4 + 1], hacc[i][j4 * 4 + 2], hacc[i][j4 * 4 + 3]));
    }
    __syncthreads();

    // ---- S5: epilogue: route h to the input gradients
    {
      const float4 sc = lds4(sCa + eq * 4), sh = lds4(sCa + CIN + eq * 4);
      const float4 mu = lds4(sCa + 2 * CIN + eq * 4), rs = lds4(sCa + 3 * CIN + eq * 4);
#pragma unroll
      for (int it = 0; it < C::EPT; ++it) {
        const int pix = (tid + it * NT) / C::QI;
        const int gy = y0 + pix / C::TW, gx = x0 + pix % C::TW;
        if (gy >= a.H || gx >= a.W) continue;
        const float4 h = lds4(sA + pix * C::AS + eq * 4);
        if (MODE == 0 || MODE == 2) {
          const long long off = ((long long)gy * a.W + gx) * CIN + eq * 4;
          const float4 z = ldg4(za_img + off);
          const float4 u = bnu4(z, sc, sh);
          float4 d;
          d.x = u.x > 0.f ? h.x : 0.f; d.y = u.y > 0.f ? h.y : 0.f;
          d.z = u.z > 0.f ? h.z : 0.f; d.w = u.w > 0.f ? h.w : 0.f;
          sa1 = add4(sa1, d);
          sa2.x = fmaf(d.x, (z.x - mu.x) * rs.x, sa2.x); sa2.y = fmaf(d.y, (z.y - mu.y) * rs.y, sa2.y);
          sa2.z = fmaf(d.z, (z.z - mu.z) * rs.z, sa2.z); sa2.w = fmaf(d.w, (z.w - mu.w) * rs.w, sa2.w);
          float* dst = a.dua + (long long)b * in_img_stride + off;
          if (a.acc_a) d = add4(d, *reinterpret_cast<const float4*>(dst));
          sts4(dst, d);
        } else {
          // max-pool: gradient goes to the first maximum of the 2x2 window (ATen order), then
          // through the ReLU mask
          const int W2 = a.W * 2;
          const long long o00 = ((long long)(gy * 2) * W2 + gx * 2) * CIN + eq * 4;
          const long long offs[4] = {o00, o00 + CIN, o00 + (long long)W2 * CIN, o00 + (long long)W2 * CIN + CIN};
          float4 z[4], v[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) { z[j] = ldg4(za_img + offs[j]); v[j] = relu4(bnu4(z[j], sc, sh)); }
          float4 d[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) d[j] = f4(0.f);
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            float best = comp(v[0], c); int bj = 0;
#pragma unroll
            for (int j = 1; j < 4; ++j) { const float vj = comp(v[j], c); if (vj > best) { best = vj; bj = j; } }
            const float hv = best > 0.f ? comp(h, c) : 0.f;
            float zb = comp(z[0], c);       // z of the winner by selects (a run-time index would put z[] in local memory)
#pragma unroll
            for (int j = 1; j < 4; ++j) if (j == bj) zb = comp(z[j], c);
            const float zh = (zb - comp(mu, c)) * comp(rs, c);
            if (c == 0) { sa1.x += hv; sa2.x = fmaf(hv, zh, sa2.x); }
            if (c == 1) { sa1.y += hv; sa2.y = fmaf(hv, zh, sa2.y); }
            if (c == 2) { sa1.z += hv; sa2.z = fmaf(hv, zh, sa2.z); }
            if (c == 3) { sa1.w += hv; sa2.w = fmaf(hv, zh, sa2.w); }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const float val = (j == bj) ? hv : 0.f;
              if (c == 0) d[j].x = val;
              if (c == 1) d[j].y = val;
              if (c == 2) d[j].z = val;
              if (c == 3) d[j].w = val;
            }
          }
          float* dbase = a.dua + (long long)b * in_img_stride;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            float4 o = d[j];
            if (a.acc_a) o = add4(o, *reinterpret_cast<const float4*>(dbase + offs[j]));
            sts4(dbase + offs[j], o);
          }
        }
      }
      if (MODE == 2) {
        // upsampled operand: one low-res pixel gathers its 2x2 children
        const float4 scb = lds4(sCb + eq * 4), shb = lds4(sCb + CIN + eq * 4);
        const float4 mub = lds4(sCb + 2 * CIN + eq * 4), rsb = lds4(sCb + 3 * CIN + eq * 4);
        const int Hb = a.H >> 1, Wb = a.W >> 1;
        for (int idx = tid; idx < (C::TP / 4) * C::QI; idx += NT) {
          const int lp = idx / C::QI;                 // 0..31 : (C::TH/2) x (C::TW/2)
          const int ly = lp / (C::TW / 2), lx = lp % (C::TW / 2);
          const int gy = (y0 >> 1) + ly, gx = (x0 >> 1) + lx;
          if (gy >= Hb || gx >= Wb) continue;
          const int p00 = (ly * 2) * C::TW + lx * 2;
          float4 hs = lds4(sA + p00 * C::AS + eq * 4);
          hs = add4(hs, lds4(sA + (p00 + 1) * C::AS + eq * 4));
          hs = add4(hs, lds4(sA + (p00 + C::TW) * C::AS + eq * 4));
          hs = add4(hs, lds4(sA + (p00 + C::TW + 1) * C::AS + eq * 4));
          const long long off = (((long long)b * Hb + gy) * Wb + gx) * CIN + eq * 4;
          const float4 z = ldg4(a.zb + off);
          const float4 u = bnu4(z, scb, shb);
          float4 d;
          d.x = u.x > 0.f ? hs.x : 0.f; d.y = u.y > 0.f ? hs.y : 0.f;
          d.z = u.z > 0.f ? hs.z : 0.f; d.w = u.w > 0.f ? hs.w : 0.f;
          sb1 = add4(sb1, d);
          sb2.x = fmaf(d.x, (z.x - mub.x) * rsb.x, sb2.x); sb2.y = fmaf(d.y, (z.y - mub.y) * rsb.y, sb2.y);
          sb2.z = fmaf(d.z, (z.z - mub.z) * rsb.z, sb2.z); sb2.w = fmaf(d.w, (z.w - mub.w) * rsb.w, sb2.w);
          float* dst = a.dub + off;
          if (a.acc_b) d = add4(d, *reinterpret_cast<const float4*>(dst));
          sts4(dst, d);
        }
      }
    }
    __syncthreads();   // sA (h) and sG are rewritten by the next tile
  }

  // ---- flush: parameter gradients -> shared-memory reduction -> one partial vector per CTA
  {
    constexpr int NW1 = COUT * CIN, NP = NW1 + 11 * COUT;
    float* sRed = smem;                       // the tile buffers are free now
    __syncthreads();
    for (int i = tid; i < NP; i += NT) sRed[i] = 0.f;
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) atomicAdd(sRed + (co3 + i) * CIN + ci3 + j, gw1[i][j]);
    constexpr int NQ = C::NQ;   // lanes l, l+NQ, ... of a warp share a channel quad (pair): reduce first
    constexpr int CPL = C::PAIR ? 2 : 4;          // channels per lane
    float vals[11 * CPL];
#pragma unroll
    for (int k = 0; k < 9; ++k) {
      if constexpr (C::PAIR) {
        vals[k * 2] = pw2[k].x; vals[k * 2 + 1] = pw2[k].y;
      } else {
        vals[k * 4] = gw2[k].x; vals[k * 4 + 1] = gw2[k].y; vals[k * 4 + 2] = gw2[k].z; vals[k * 4 + 3] = gw2[k].w;
      }
    }
    vals[9 * CPL] = gb2.x; vals[9 * CPL + 1] = gb2.y;
    vals[10 * CPL] = gb1.x; vals[10 * CPL + 1] = gb1.y;
    if constexpr (!C::PAIR) {
      vals[38] = gb2.z; vals[39] = gb2.w;
      vals[42] = gb1.z; vals[43] = gb1.w;
    }
#pragma unroll
    for (int v = 0; v < 11 * CPL; ++v) {
#pragma unroll
      for (int o = 16; o >= NQ; o >>= 1) vals[v] += __shfl_xor_sync(0xffffffffu, vals[v], o);
    }
    if ((tid & 31) < NQ) {
#pragma unroll
      for (int k = 0; k < 9; ++k)
#pragma unroll
        for (int c = 0; c < CPL; ++c) atomicAdd(sRed + NW1 + COUT + (dq * CPL + c) * 9 + k, vals[k * CPL + c]);
#pragma unroll
      for (int c = 0; c < CPL; ++c) {
        atomicAdd(sRed + NW1 + 10 * COUT + dq * CPL + c, vals[9 * CPL + c]);   // gb2
        atomicAdd(sRed + NW1 + dq * CPL + c, vals[10 * CPL + c]);              // gb1
      }
    }
    __syncthreads();
    float* dst = a.partial + (long long)blockIdx.x * kPartialStride;
    for (int i = tid; i < NP; i += NT) dst[i] = sRed[i];
  }
  {
    constexpr int QI = C::QI;
    float vals[16] = {sa1.x, sa1.y, sa1.z, sa1.w, sa2.x, sa2.y, sa2.z, sa2.w,
                      sb1.x, sb1.y, sb1.z, sb1.w, sb2.x, sb2.y, sb2.z, sb2.w};
#pragma unroll
    for (int v = 0; v < 16; ++v) {
#pragma unroll
      for (int o = 16; o >= QI; o >>= 1) vals[v] += __shfl_xor_sync(0xffffffffu, vals[v], o);
    }
    if ((tid & 31) < QI) {
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
}

// ------------------------------------------------------------------------------- stem backward
// d(conv1.weight)[co][c][ky][kx] = sum_p g[p][co] * img[c][2y+ky-1][2x+kx-1],  d(bias) = sum_p g
constexpr int ST_TH = 8, ST_TW = 32;
constexpr int ST_IH = 2 * ST_TH + 1, ST_IW = 2 * ST_TW + 1;
// the patch is staged as aligned float4 columns -4 .. 2*ST_TW-1 around the tile's first input column
// (17 per row; input column x of the patch sits at float index x + 3)
constexpr int ST_C4 = 2 * ST_TW / 4 + 1;
constexpr int ST_IWP = ST_C4 * 4;
constexpr int ST_NPG = 28;       // pixel groups of the weight-gradient GEMM (9 threads each)

// Thread mapping of the GEMM  dW[co][c,ky,kx] += g[p][co] * in[c][2y+ky-1][2x+kx-1]:
//   thread = (c,ky) x pixel group, 16 co x 3 kx accumulators in registers; per pixel 4 broadcast
//   LDS.128 of g + 3 LDS of the input row feed 24 packed FMAs (the previous 4 x 3 tile issued an
//   index computation and 4 loads per 12 FMAs and was issue-bound at 73 warp instructions / pixel).
// Double buffered: the next tile's image patch and raw du / z_out rows are copied in with cp.async
// while the current tile is reduced (the staging loads were ~28 % of the stall samples).
constexpr int ST_PATCH = 3 * ST_IH * ST_IWP;                 // floats
constexpr int ST_BUF = ST_PATCH + 2 * ST_TH * ST_TW * 16;    // + du tile (-> g in place) + z_out tile

__global__ void __launch_bounds__(256, 2) stem_bwd_kernel(const StemBwdArgs a) {
  extern __shared__ __align__(16) float stem_smem[];
  __shared__ float sCo[5][16];
  const int tid = threadIdx.x;
  const int Ho = a.Hin / 2, Wo = a.Win / 2;
  if (tid < 16) {
    Coef k = bn_coef(a.bno, tid);
    sCo[0][tid] = a.bno.gamma[tid] * k.rstd;
    sCo[1][tid] = (float)(a.dsum[tid] * a.bno.inv_count);
    sCo[2][tid] = (float)(a.dsumzh[tid] * a.bno.inv_count);
    sCo[3][tid] = k.mean;
    sCo[4][tid] = k.rstd;
  }
  __syncthreads();
  const int ck = tid % 9, pg = tid / 9;         // ck = c*3 + ky; pg 0..27 (tid >= 252 idle in the GEMM)
  const int c = ck / 3, ky = ck - c * 3;
  float acc[16][3];
#pragma unroll
  for (int i = 0; i < 16; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) acc[i][j] = 0.f;
  // BN-backward constants of this thread's channel quad (tid & 3 is the quad of every g item it stages)
  const int q = tid & 3;
  float4 k_gs, k_m1, k_m2, k_mu, k_rs;
  k_gs = lds4(&sCo[0][q * 4]); k_m1 = lds4(&sCo[1][q * 4]); k_m2 = lds4(&sCo[2][q * 4]);
  k_mu = lds4(&sCo[3][q * 4]); k_rs = lds4(&sCo[4][q * 4]);
  float4 bsum = f4(0.f);                        // bias gradient partial of channels 4q .. 4q+3

  const int tiles_x = (Wo + ST_TW - 1) / ST_TW, tiles_y = (Ho + ST_TH - 1) / ST_TH;
  const int ntiles = tiles_x * tiles_y * a.B;
  auto stage = [&](int tile, int buf) {
    int t = tile;
    const int tx = t % tiles_x; t /= tiles_x;
    const int ty = t % tiles_y;
    const int b = t / tiles_y;
    const int ox0 = tx * ST_TW, oy0 = ty * ST_TH;
    const int gx0 = 2 * ox0 - 4, iy0 = 2 * oy0 - 1;
    float* dst = stem_smem + buf * ST_BUF;
    for (int i = tid; i < 3 * ST_IH * ST_C4; i += 256) {
      const int rowid = i / ST_C4, j = i - rowid * ST_C4;
      const int cc = rowid / ST_IH, r = rowid - cc * ST_IH;
      const int gy = iy0 + r, gx = gx0 + 4 * j;            // Win % 4 == 0: a chunk is all in or all out
      const bool in = gy >= 0 && gy < a.Hin && gx >= 0 && gx < a.Win;
      const float* src = in ? a.img + (((long long)b * 3 + cc) * a.Hin + gy) * a.Win + gx : a.img;
      cp_async16(dst + rowid * ST_IWP + 4 * j, src, in);
    }
    // du / z_out: item (pixel, quad) -> the thread that later turns it into g (no barrier in between)
#pragma unroll
    for (int it = 0; it < ST_TH * ST_TW * 4 / 256; ++it) {
      const int pix = (tid >> 2) + it * 64;
      const int oy = oy0 + pix / ST_TW, ox = ox0 + pix % ST_TW;
      const bool in = oy < Ho && ox < Wo;
      const long long off = in ? (((long long)b * Ho + oy) * Wo + ox) * 16 + q * 4 : 0;
      cp_async16(dst + ST_PATCH + pix * 16 + q * 4, a.du + off, in);
      cp_async16(dst + ST_PATCH + ST_TH * ST_TW * 16 + pix * 16 + q * 4, a.zout + off, in);
    }
  };
  if ((int)blockIdx.x < ntiles) stage(blockIdx.x, 0);
  cp_async_commit();

  int buf = 0;
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, buf ^= 1) {
    int t = tile;
    const int tx = t % tiles_x; t /= tiles_x;
    const int ty = t % tiles_y;
    const int ox0 = tx * ST_TW, oy0 = ty * ST_TH;
    __syncthreads();      // the GEMM readers of the other buffer (previous tile) are done
    if (tile + (int)gridDim.x < ntiles) stage(tile + gridDim.x, buf ^ 1);
    cp_async_commit();
    cp_async_wait<1>();   // this tile's copies (all but the group just committed) have landed
    float* sIn = stem_smem + buf * ST_BUF;                     // [3][ST_IH][ST_IWP]
    float* sGs = sIn + ST_PATCH;                               // [pixels][16]: du -> g in place
    const float* sZ = sGs + ST_TH * ST_TW * 16;
#pragma unroll
    for (int it = 0; it < ST_TH * ST_TW * 4 / 256; ++it) {
      const int pix = (tid >> 2) + it * 64;
      const int oy = oy0 + pix / ST_TW, ox = ox0 + pix % ST_TW;
      float4 g = f4(0.f);
      if (oy < Ho && ox < Wo) {
        const float4 d = lds4(sGs + pix * 16 + q * 4), z = lds4(sZ + pix * 16 + q * 4);
        g.x = k_gs.x * (d.x - k_m1.x - (z.x - k_mu.x) * k_rs.x * k_m2.x);
        g.y = k_gs.y * (d.y - k_m1.y - (z.y - k_mu.y) * k_rs.y * k_m2.y);
        g.z = k_gs.z * (d.z - k_m1.z - (z.z - k_mu.z) * k_rs.z * k_m2.z);
        g.w = k_gs.w * (d.w - k_m1.w - (z.w - k_mu.w) * k_rs.w * k_m2.w);
      }
      bsum = add4(bsum, g);
      sts4(sGs + pix * 16 + q * 4, g);
    }
    __syncthreads();
    if (tid < 9 * ST_NPG) {
#pragma unroll 2
      for (int p = pg; p < ST_TH * ST_TW; p += ST_NPG) {
        const int ly = p / ST_TW, lx = p - ly * ST_TW;
        const float* row = sIn + (c * ST_IH + 2 * ly + ky) * ST_IWP + 2 * lx + 3;
        const float v0 = row[0], v1 = row[1], v2 = row[2];
#pragma unroll
        for (int j4 = 0; j4 < 4; ++j4) {
          const float4 g = lds4(sGs + p * 16 + j4 * 4);
          fma2(acc[j4 * 4 + 0][0], acc[j4 * 4 + 1][0], g.x, g.y, v0, v0);
          fma2(acc[j4 * 4 + 2][0], acc[j4 * 4 + 3][0], g.z, g.w, v0, v0);
          fma2(acc[j4 * 4 + 0][1], acc[j4 * 4 + 1][1], g.x, g.y, v1, v1);
          fma2(acc[j4 * 4 + 2][1], acc[j4 * 4 + 3][1], g.z, g.w, v1, v1);
          fma2(acc[j4 * 4 + 0][2], acc[j4 * 4 + 1][2], g.x, g.y, v2, v2);
          fma2(acc[j4 * 4 + 2][2], acc[j4 * 4 + 3][2], g.z, g.w, v2, v2);
        }
      }
    }
  }
  cp_async_wait<0>();
  // flush: shared-memory reduction, one partial vector [gw (432) | gb (16)] per CTA
  float* sRed = stem_smem;
  __syncthreads();
  for (int i = tid; i < 448; i += 256) sRed[i] = 0.f;
  __syncthreads();
  if (tid < 9 * ST_NPG) {
#pragma unroll
    for (int i = 0; i < 16; ++i)
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) atomicAdd(sRed + i * 27 + c * 9 + ky * 3 + kx, acc[i][kx]);
  }
  atomicAdd(sRed + 432 + q * 4 + 0, bsum.x);
  atomicAdd(sRed + 432 + q * 4 + 1, bsum.y);
  atomicAdd(sRed + 432 + q * 4 + 2, bsum.z);
  atomicAdd(sRed + 432 + q * 4 + 3, bsum.w);
  __syncthreads();
  float* dst = a.partial + (long long)blockIdx.x * kPartialStride;
  for (int i = tid; i < 448; i += 256) dst[i] = sRed[i];
}

__global__ void reduce_partials_kernel(const float* __restrict__ partial, int ncta, int n,
                                       float* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
  int c = 0;
  for (; c + 3 < ncta; c += 4) {
    s0 += partial[(long long)(c + 0) * kPartialStride + i];
    s1 += partial[(long long)(c + 1) * kPartialStride + i];
    s2 += partial[(long long)(c + 2) * kPartialStride + i];
    s3 += partial[(long long)(c + 3) * kPartialStride + i];
  }
  for (; c < ncta; ++c) s0 += partial[(long long)c * kPartialStride + i];
  out[i] = (s0 + s1) + (s2 + s3);
}

__global__ void bn_param_grads_kernel(const BnFinalizeArgs a, const double* dsum,
                                      const double* dsumzh, float* grad) {
  const int i = blockIdx.x, c = threadIdx.x;
  if (i >= a.n || c >= a.C[i]) return;
  grad[a.gamma_off[i] + c] = (float)dsumzh[a.ch_off[i] + c];
  grad[a.beta_off[i] + c] = (float)dsum[a.ch_off[i] + c];
}

template <int CIN, int COUT, int MODE, int HAS_BN, int OCC, int PAIRM = 1>
cudaError_t launch_unit_bwd_o(const UnitBwdArgs& a, int num_sms, cudaStream_t s) {
  using C = BwdCfg<CIN, COUT, PAIRM>;
  constexpr int RAWF = (HAS_BN ? 2 : 1) * C::HP * COUT;
  constexpr bool PF = OCC == 2 && (C::SMEM_FLOATS + RAWF) * 4 <= 112 * 1024;      // as in the kernel
  const size_t smem = sizeof(float) * (C::SMEM_FLOATS + (PF ? RAWF : 0));
  auto kern = unit_bwd_kernel<CIN, COUT, MODE, HAS_BN, OCC, PAIRM>;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    configured = true;
  }
  const int ntiles = ((a.W + C::TW - 1) / C::TW) * ((a.H + C::TH - 1) / C::TH) * a.B;
  const int per_sm = OCC;
  int grid = per_sm * num_sms < ntiles ? per_sm * num_sms : ntiles;
  if (grid > kMaxPartialCtas) grid = kMaxPartialCtas;
  kern<<<grid, NT, smem, s>>>(a);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  // gw1, gb1, gw2, gb2 are adjacent in the bucket in this order (plan.cpp)
  return launch_reduce_partials(a.partial, grid, COUT * CIN + 11 * COUT, a.gw1, s);
}

// CTAs per SM: two for the small-weight units (128 registers; with the channel-pair depthwise mapping
// the persistent gradient accumulators no longer spill).  Development knobs, read once:
// YUNET_BWD_OCC=1 selects the single-CTA build of the 16-channel units, YUNET_BWD_QUAD=1 the round-1
// quad mapping of the depthwise stage (spills ~500 B per thread at 128 registers) for A/B timing.
template <int CIN, int COUT, int MODE, int HAS_BN>
cudaError_t launch_unit_bwd_t(const UnitBwdArgs& a, int num_sms, cudaStream_t s) {
  constexpr int DEF = (CIN * COUT <= 1024) ? 2 : 1;
  if (CIN == 16 && DEF == 2) {
    static const int occ = [] { const char* e = getenv("YUNET_BWD_OCC"); return e ? atoi(e) : 0; }();
    if (occ == 1) return launch_unit_bwd_o<CIN, COUT, MODE, HAS_BN, (CIN == 16 ? 1 : DEF)>(a, num_sms, s);
  }
  if (DEF == 2) {
    static const int quad = [] { const char* e = getenv("YUNET_BWD_QUAD"); return e ? atoi(e) : 0; }();
    if (quad == 1) return launch_unit_bwd_o<CIN, COUT, MODE, HAS_BN, DEF, (DEF == 2 ? 0 : 1)>(a, num_sms, s);
  }
  return launch_unit_bwd_o<CIN, COUT, MODE, HAS_BN, DEF>(a, num_sms, s);
}

template <int CIN, int COUT>
cudaError_t launch_unit_bwd_m(int mode, const UnitBwdArgs& a, int num_sms, cudaStream_t s) {
  if (a.has_bn) {
    switch (mode) {
      case 0: return launch_unit_bwd_t<CIN, COUT, 0, 1>(a, num_sms, s);
      case 1: return launch_unit_bwd_t<CIN, COUT, 1, 1>(a, num_sms, s);
      case 2: return launch_unit_bwd_t<CIN, COUT, 2, 1>(a, num_sms, s);
    }
  } else if (mode == 0) {
    return launch_unit_bwd_t<CIN, COUT, 0, 0>(a, num_sms, s);
  }
  return cudaErrorInvalidValue;
}

}  // namespace

cudaError_t launch_unit_bwd(int cin, int cout, int mode, const UnitBwdArgs& a, int num_sms,
                            cudaStream_t s) {
  if (cin == 16 && cout == 16) return launch_unit_bwd_m<16, 16>(mode, a, num_sms, s);
  if (cin == 16 && cout == 32) return launch_unit_bwd_m<16, 32>(mode, a, num_sms, s);
  if (cin == 16 && cout == 64) return launch_unit_bwd_m<16, 64>(mode, a, num_sms, s);
  if (cin == 32 && cout == 32) return launch_unit_bwd_m<32, 32>(mode, a, num_sms, s);
  if (cin == 32 && cout == 64) return launch_unit_bwd_m<32, 64>(mode, a, num_sms, s);
  if (cin == 64 && cout == 64) return launch_unit_bwd_m<64, 64>(mode, a, num_sms, s);
  if (cin == 64 && cout == 16) return launch_unit_bwd_m<64, 16>(mode, a, num_sms, s);
  return cudaErrorInvalidValue;
}

cudaError_t launch_stem_bwd(const StemBwdArgs& a, int num_sms, cudaStream_t s) {
  if (a.Win & 3) return cudaErrorInvalidValue;     // float4 staging of the image rows
  const int Ho = a.Hin / 2, Wo = a.Win / 2;
  const int ntiles = ((Wo + ST_TW - 1) / ST_TW) * ((Ho + ST_TH - 1) / ST_TH) * a.B;
  int grid = 2 * num_sms < ntiles ? 2 * num_sms : ntiles;
  if (grid > kMaxPartialCtas) grid = kMaxPartialCtas;
  const int smem = 2 * ST_BUF * (int)sizeof(float);
  static bool configured = false;
  if (!configured) {
    cudaError_t e0 = cudaFuncSetAttribute(stem_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e0 != cudaSuccess) return e0;
    configured = true;
  }
  stem_bwd_kernel<<<grid, 256, smem, s>>>(a);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  return launch_reduce_partials(a.partial, grid, 448, a.gw, s);   // [weight (432) | bias (16)]
}

cudaError_t launch_reduce_partials(const float* partial, int ncta, int n, float* out, cudaStream_t s) {
  reduce_partials_kernel<<<(n + 127) / 128, 128, 0, s>>>(partial, ncta, n, out);
  return cudaGetLastError();
}

cudaError_t launch_bn_param_grads(const BnFinalizeArgs& a, const double* dsum,
                                  const double* dsumzh, float* grad_bucket, cudaStream_t s) {
  bn_param_grads_kernel<<<a.n, 64, 0, s>>>(a, dsum, dsumzh, grad_bucket);
  return cudaGetLastError();
}

}  // namespace yunet
