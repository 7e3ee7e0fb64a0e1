// Batched SimOTA assignment and the multi-task head loss with its gradient, sm_100a.
//
// simota_assign_kernel — one CTA per image; replaces the per-image Python loop
//   multi_apply(_get_target_single) (mmdet/models/dense_heads/yunet_head.py:483-489,536-604) and
//   SimOTAAssigner._assign / get_in_gt_and_in_center_info / dynamic_k_matching
//   (mmdet/core/bbox/assigners/sim_ota_assigner.py:95-257), with bbox_overlaps of
//   mmdet/core/bbox/iou_calculators/iou2d_calculator.py:213-253.  The fp32 operation order of
//   the reference is kept (explicit round-to-nearest intrinsics, no FMA contraction) because the
//   additive INF=1e5 quantises costs and `int()` truncates the top-k IoU sum.
//   Tie-break where torch.topk leaves the order unspecified: lowest prior index first.
// loss_grad_kernel — YuNet_Head.loss (yunet_head.py:493-534): sigmoid-BCE objectness over all
//   priors, sigmoid-BCE classification with IoU-soft targets, smooth-EIoU
//   (mmdet/models/losses/iou_loss.py:194-227) through the box decode (yunet_head.py:376-386) and
//   SmoothL1 landmarks (losses/smooth_l1_loss.py:24-32, losses/utils.py:42-59), plus
//   d(loss)/d(preds) in the same pass.
#include <cfloat>
#include <cstring>
#include <cstdio>

#include "kernels.h"

namespace yunet {

namespace {

constexpr int NT = 256;
constexpr int NW = NT / 32;
// the assignment runs one CTA per image and its time is that of the image with the most faces (one warp
// per gt in phase 2): 16 warps halve the rounds of the crowded images, two CTAs still fit an SM
constexpr int NTA = 512;
constexpr int NWA = NTA / 32;
constexpr int KTOP = 10;   // per-lane candidate list length (candidate_topk <= 10)
constexpr int GT_ROW = 19;
constexpr int PC = 16;     // prediction channels

struct __align__(16) Cand {   // one valid prior of the image (32 bytes)
  float x1, y1, x2, y2;       // decoded box
  float cls_cost;
  int idx;                    // prior index
  int cnt;                    // number of gts that selected this prior
  int gsel;                   // (min) gt index that selected it
};

__device__ __forceinline__ float sigmoid_ref(float x) {
  return __fdiv_rn(1.0f, __fadd_rn(1.0f, expf(-x)));
}

__device__ __forceinline__ void prior_of(const LevelGeom& g, int p, float& px, float& py, float& s) {
  int l = 0, j = p;
  if (p >= g.off[2]) { l = 2; j = p - g.off[2]; }
  else if (p >= g.off[1]) { l = 1; j = p - g.off[1]; }
  const int w = g.w[l];
  s = (float)g.stride[l];
  px = (float)((j % w) * g.stride[l]);
  py = (float)((j / w) * g.stride[l]);
}

// is the (offset) prior centre inside the gt box / inside the gt centre region?
// sim_ota_assigner.py:186-228
__device__ __forceinline__ void in_flags(float x, float y, float s, float radius, float4 gb,
                                         bool& in_gt, bool& in_ct) {
  const float l_ = __fsub_rn(x, gb.x), t_ = __fsub_rn(y, gb.y);
  const float r_ = __fsub_rn(gb.z, x), b_ = __fsub_rn(gb.w, y);
  in_gt = fminf(fminf(l_, t_), fminf(r_, b_)) > 0.f;
  const float cx = __fdiv_rn(__fadd_rn(gb.x, gb.z), 2.0f);
  const float cy = __fdiv_rn(__fadd_rn(gb.y, gb.w), 2.0f);
  const float rs = __fmul_rn(radius, s);
  const float cl_ = __fsub_rn(x, __fsub_rn(cx, rs));
  const float ct_ = __fsub_rn(y, __fsub_rn(cy, rs));
  const float cr_ = __fsub_rn(__fadd_rn(cx, rs), x);
  const float cb_ = __fsub_rn(__fadd_rn(cy, rs), y);
  in_ct = fminf(fminf(cl_, ct_), fminf(cr_, cb_)) > 0.f;
}

// bbox_overlaps(mode='iou'), iou2d_calculator.py:213-253
__device__ __forceinline__ float pair_iou(float x1, float y1, float x2, float y2, float4 gb) {
  const float area1 = __fmul_rn(__fsub_rn(x2, x1), __fsub_rn(y2, y1));
  const float area2 = __fmul_rn(__fsub_rn(gb.z, gb.x), __fsub_rn(gb.w, gb.y));
  const float w = fmaxf(__fsub_rn(fminf(x2, gb.z), fmaxf(x1, gb.x)), 0.f);
  const float h = fmaxf(__fsub_rn(fminf(y2, gb.w), fmaxf(y1, gb.y)), 0.f);
  const float overlap = __fmul_rn(w, h);
  float uni = __fsub_rn(__fadd_rn(area1, area2), overlap);
  uni = fmaxf(uni, 1e-6f);
  return __fdiv_rn(overlap, uni);
}

// sim_ota_assigner.py:154-169
__device__ __forceinline__ float pair_cost(const yunet_loss_cfg_dev& lc, float cls_cost, float iou,
                                           bool in_both) {
  const float iou_cost = -logf(__fadd_rn(iou, 1e-7f));
  const float t = __fadd_rn(__fmul_rn(cls_cost, lc.cls_weight), __fmul_rn(iou_cost, lc.iou_weight));
  return __fadd_rn(t, in_both ? 0.0f : 100000.0f);
}

__device__ __forceinline__ unsigned ordered_bits(float f) {
  const unsigned b = __float_as_uint(f);
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}

__device__ __forceinline__ float4 load_gt_box(const float* gt, int g) {
  const float* r = gt + (long long)g * GT_ROW;
  return make_float4(__ldg(r), __ldg(r + 1), __ldg(r + 2), __ldg(r + 3));
}

// decode (yunet_head.py:376-386) with the non-offset prior
__device__ __forceinline__ void decode_box(const float* pr, float px, float py, float s, float& x1,
                                           float& y1, float& x2, float& y2) {
  const float cx = __fadd_rn(__fmul_rn(pr[1], s), px);
  const float cy = __fadd_rn(__fmul_rn(pr[2], s), py);
  const float w = __fmul_rn(expf(pr[3]), s);
  const float h = __fmul_rn(expf(pr[4]), s);
  const float hw = __fdiv_rn(w, 2.0f), hh = __fdiv_rn(h, 2.0f);
  x1 = __fsub_rn(cx, hw); y1 = __fsub_rn(cy, hh);
  x2 = __fadd_rn(cx, hw); y2 = __fadd_rn(cy, hh);
}

// Explicit per-prior inputs of SimOTAAssigner.assign (sim_ota_assigner.py:38-93): scores (P) =
// sigmoid(cls)*sigmoid(obj) as the head passes them, priors (P,4) = [cx, cy, stride_w, stride_h]
// already offset by half a stride (yunet_head.py:570-573; stride_w is used for both axes, they are
// equal in every YuNet level), decoded boxes (P,4).  EXT = false derives all three from `preds`.
struct AssignExt { const float* scores; const float* priors; const float* boxes; };

template <bool EXT>
__global__ void __launch_bounds__(NTA) simota_assign_kernel(
    const yunet_loss_cfg_dev lc, const LevelGeom geo, const float* __restrict__ preds,
    const float* __restrict__ gt, const int* __restrict__ gt_offsets, int* __restrict__ assigned,
    float* __restrict__ matched_iou, float* counters, Cand* gscratch, int vcap, const AssignExt ext) {
  // centre of prior p as the assigner sees it (offset prior) and its stride
  auto offset_prior = [&](int p, float& ox, float& oy, float& s) {
    if (EXT) {
      ox = __ldg(ext.priors + 4 * p); oy = __ldg(ext.priors + 4 * p + 1); s = __ldg(ext.priors + 4 * p + 2);
    } else {
      float px, py;
      prior_of(geo, p, px, py, s);
      ox = __fadd_rn(px, __fmul_rn(s, 0.5f));   // yunet_head.py:572-573
      oy = __fadd_rn(py, __fmul_rn(s, 0.5f));
    }
  };
  extern __shared__ float4 smem_raw[];
  Cand* scand = reinterpret_cast<Cand*>(smem_raw);                       // [vcap]
  unsigned char* sflag = reinterpret_cast<unsigned char*>(scand + vcap); // [P]
  __shared__ int s_warp[NWA];
  __shared__ int s_base;
  __shared__ int s_V;
  __shared__ float s_red[2][NWA];

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int b = blockIdx.x;
  const int P = geo.P;
  const int g0 = gt_offsets[b];
  const int G = gt_offsets[b + 1] - g0;
  const float* gtb = gt + (long long)g0 * GT_ROW;
  const float* pb = preds + (long long)b * P * PC;
  int* asg = assigned + (long long)b * P;
  float* miou = matched_iou + (long long)b * P;

  // ---- pass A: validity flag per prior, zero the outputs
  int nvalid_local = 0;
  for (int p = tid; p < P; p += NTA) {
    float ox, oy, s;
    offset_prior(p, ox, oy, s);
    bool any_gt = false, any_ct = false;
    for (int g = 0; g < G; ++g) {
      bool ig, ic;
      in_flags(ox, oy, s, lc.center_radius, load_gt_box(gtb, g), ig, ic);
      any_gt |= ig; any_ct |= ic;
    }
    const bool v = any_gt || any_ct;
    sflag[p] = v ? 1 : 0;
    nvalid_local += v ? 1 : 0;
    asg[p] = 0;
    miou[p] = 0.f;
  }
  // total V
  {
    int v = nvalid_local;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) s_warp[warp] = v;
    __syncthreads();
    if (tid == 0) {
      int t = 0;
      for (int w = 0; w < NWA; ++w) t += s_warp[w];
      s_V = t;
      s_base = 0;
    }
    __syncthreads();
  }
  const int V = s_V;
  if (V == 0 || G == 0) return;
  Cand* cand = (V <= vcap) ? scand : (gscratch + (long long)b * P);

  // ---- pass B: ordered compaction + per-candidate decode and classification cost
  for (int base = 0; base < P; base += NTA) {
    const int p = base + tid;
    const bool v = p < P && sflag[p] != 0;
    const unsigned m = __ballot_sync(0xffffffffu, v);
    if (lane == 0) s_warp[warp] = __popc(m);
    __syncthreads();
    int off = s_base;
    for (int w = 0; w < warp; ++w) off += s_warp[w];
    if (v) {
      const int slot = off + __popc(m & ((1u << lane) - 1u));
      Cand c;
      float sc;
      if (EXT) {
        c.x1 = __ldg(ext.boxes + 4 * p); c.y1 = __ldg(ext.boxes + 4 * p + 1);
        c.x2 = __ldg(ext.boxes + 4 * p + 2); c.y2 = __ldg(ext.boxes + 4 * p + 3);
        sc = __ldg(ext.scores + p);
      } else {
        float px, py, s;
        prior_of(geo, p, px, py, s);
        const float* pr = pb + (long long)p * PC;
        decode_box(pr, px, py, s, c.x1, c.y1, c.x2, c.y2);
        // yunet_head.py:576: cls.sigmoid() * obj.sigmoid()
        sc = __fmul_rn(sigmoid_ref(pr[0]), sigmoid_ref(pr[5]));
      }
      // sim_ota_assigner.py:160-165: BCE of sqrt(score) against the one-hot label (= 1), torch
      // clamps log at -100
      c.cls_cost = -fmaxf(logf(sqrtf(sc)), -100.0f);
      c.idx = p;
      c.cnt = 0;
      c.gsel = 0x7fffffff;
      cand[slot] = c;
    }
    __syncthreads();
    if (tid == 0) {
      int t = 0;
      for (int w = 0; w < NWA; ++w) t += s_warp[w];
      s_base += t;
    }
    __syncthreads();
  }

  // ---- phase 2: one warp per gt: dynamic k from the top-k IoUs, then the k cheapest candidates
  const int K = lc.candidate_topk < KTOP ? lc.candidate_topk : KTOP;
  for (int g = warp; g < G; g += NWA) {
    const float4 gb = load_gt_box(gtb, g);
    float ti[KTOP];
    unsigned long long kc[KTOP];
#pragma unroll
    for (int i = 0; i < KTOP; ++i) { ti[i] = -1.0f; kc[i] = ~0ull; }
    for (int v = lane; v < V; v += 32) {
      const Cand c = cand[v];
      const float iou = pair_iou(c.x1, c.y1, c.x2, c.y2, gb);
      float ox, oy, s;
      offset_prior(c.idx, ox, oy, s);
      bool ig, ic;
      in_flags(ox, oy, s, lc.center_radius, gb, ig, ic);
      const float cost = pair_cost(lc, c.cls_cost, iou, ig && ic);
      if (iou > ti[KTOP - 1]) {
        ti[KTOP - 1] = iou;
#pragma unroll
        for (int i = KTOP - 1; i > 0; --i)
          if (ti[i] > ti[i - 1]) { const float t = ti[i]; ti[i] = ti[i - 1]; ti[i - 1] = t; }
      }
      const unsigned long long key = ((unsigned long long)ordered_bits(cost) << 32) | (unsigned)v;
      if (key < kc[KTOP - 1]) {
        kc[KTOP - 1] = key;
#pragma unroll
        for (int i = KTOP - 1; i > 0; --i)
          if (kc[i] < kc[i - 1]) { const unsigned long long t = kc[i]; kc[i] = kc[i - 1]; kc[i - 1] = t; }
      }
    }
    // merge the per-lane IoU lists: sum of the top-K (descending order, like topk(...).sum(0))
    float sum = 0.f;
    for (int r = 0; r < K; ++r) {
      float m = ti[0];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
      if (m < 0.f) break;
      const unsigned who = __ballot_sync(0xffffffffu, ti[0] == m);
      if (lane == __ffs(who) - 1) {
#pragma unroll
        for (int i = 0; i < KTOP - 1; ++i) ti[i] = ti[i + 1];
        ti[KTOP - 1] = -1.0f;
      }
      sum = __fadd_rn(sum, m);
    }
    int dyn_k = (int)sum;   // .int() truncation, sim_ota_assigner.py:236
    if (dyn_k < 1) dyn_k = 1;
    for (int r = 0; r < dyn_k; ++r) {
      unsigned long long m = kc[0];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long other = __shfl_xor_sync(0xffffffffu, m, o);
        m = other < m ? other : m;
      }
      if (m == ~0ull) break;
      if (kc[0] == m) {
#pragma unroll
        for (int i = 0; i < KTOP - 1; ++i) kc[i] = kc[i + 1];
        kc[KTOP - 1] = ~0ull;
      }
      if (lane == 0) {
        const int slot = (int)(m & 0xffffffffu);
        atomicAdd(&cand[slot].cnt, 1);
        atomicMin(&cand[slot].gsel, g);
      }
    }
  }
  __syncthreads();

  // ---- phase 3: resolve multi-matched priors (argmin over ALL gts, sim_ota_assigner.py:244-249),
  // matched IoU, outputs, counters
  float npos = 0.f, wsum = 0.f;
  for (int v = tid; v < V; v += NTA) {
    const Cand c = cand[v];
    if (c.cnt == 0) continue;
    int gsel = c.gsel;
    if (c.cnt > 1) {
      float ox, oy, s;
      offset_prior(c.idx, ox, oy, s);
      float best = FLT_MAX;
      for (int g = 0; g < G; ++g) {
        const float4 gb = load_gt_box(gtb, g);
        bool ig, ic;
        in_flags(ox, oy, s, lc.center_radius, gb, ig, ic);
        const float cost = pair_cost(lc, c.cls_cost, pair_iou(c.x1, c.y1, c.x2, c.y2, gb), ig && ic);
        if (cost < best) { best = cost; gsel = g; }
      }
    }
    const float iou = pair_iou(c.x1, c.y1, c.x2, c.y2, load_gt_box(gtb, gsel));
    asg[c.idx] = gsel + 1;
    miou[c.idx] = iou;
    npos += 1.f;
    const float* w = gtb + (long long)gsel * GT_ROW + 14;
    // torch.mean over the 5 landmark weights (yunet_head.py:598-599)
    wsum += __fdiv_rn(__fadd_rn(__fadd_rn(__fadd_rn(__fadd_rn(__ldg(w), __ldg(w + 1)), __ldg(w + 2)), __ldg(w + 3)), __ldg(w + 4)), 5.0f);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    npos += __shfl_xor_sync(0xffffffffu, npos, o);
    wsum += __shfl_xor_sync(0xffffffffu, wsum, o);
  }
  if (lane == 0) { s_red[0][warp] = npos; s_red[1][warp] = wsum; }
  __syncthreads();
  if (tid == 0) {
    float a = 0.f, c = 0.f;
    for (int w = 0; w < NWA; ++w) { a += s_red[0][w]; c += s_red[1][w]; }
    atomicAdd(counters + 0, a);
    atomicAdd(counters + 1, c);
  }
}

// ------------------------------------------------------------------------------- loss + gradient
__device__ __forceinline__ float bce_logits(float x, float t) {
  // binary_cross_entropy_with_logits: max(x,0) - x*t + log1p(exp(-|x|))
  return fmaxf(x, 0.f) - x * t + log1pf(expf(-fabsf(x)));
}

__device__ __forceinline__ void min_bwd(float a, float b, float g, float& ga, float& gb) {
  // torch.min(a,b) backward: ties split evenly
  if (a < b) ga += g; else if (b < a) gb += g; else { ga += 0.5f * g; gb += 0.5f * g; }
}
__device__ __forceinline__ void max_bwd(float a, float b, float g, float& ga, float& gb) {
  if (a > b) ga += g; else if (b > a) gb += g; else { ga += 0.5f * g; gb += 0.5f * g; }
}

__global__ void __launch_bounds__(NT) loss_grad_kernel(
    const yunet_loss_cfg_dev lc, const LevelGeom geo, const float* __restrict__ preds,
    const float* __restrict__ gt, const int* __restrict__ gt_offsets,
    const int* __restrict__ assigned, const float* __restrict__ matched_iou,
    const float* __restrict__ counters, const float* __restrict__ num_total, float s_cls,
    float s_bbox, float s_obj, float s_kps, int B, float* losses, float* d_preds) {
  __shared__ float s_red[4][NW];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int P = geo.P;
  const long long n = (long long)B * P;
  const long long i = (long long)blockIdx.x * NT + tid;
  // yunet_head.py:497: max(reduce_mean(num_pos), 1.0)
  const float N = fmaxf(__ldg(num_total), 1.0f);
  const float invN = 1.0f / N;
  // losses/utils.py:51-55: loss.sum() / (avg_factor + eps), avg_factor = sum(kps_weight)
  const float inv_kw = 1.0f / (__ldg(counters + 1) + 1.1920928955078125e-07f);
  float l_cls = 0.f, l_bbox = 0.f, l_obj = 0.f, l_kps = 0.f;
  if (i < n) {
    const int b = (int)(i / P), p = (int)(i % P);
    const float* pr = preds + i * PC;
    float pv[PC];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(pr) + q);
      pv[q * 4] = v.x; pv[q * 4 + 1] = v.y; pv[q * 4 + 2] = v.z; pv[q * 4 + 3] = v.w;
    }
    float dv[PC];
#pragma unroll
    for (int k = 0; k < PC; ++k) dv[k] = 0.f;
    const int a = assigned[i];
    const float tobj = a > 0 ? 1.f : 0.f;
    l_obj = bce_logits(pv[5], tobj);
    dv[5] = (1.0f / (1.0f + expf(-pv[5])) - tobj) * (lc.w_obj * s_obj * invN);
    if (a > 0) {
      const float* gr = gt + (long long)(gt_offsets[b] + a - 1) * GT_ROW;
      const float tiou = matched_iou[i];
      l_cls = bce_logits(pv[0], tiou);
      dv[0] = (1.0f / (1.0f + expf(-pv[0])) - tiou) * (lc.w_cls * s_cls * invN);
      float px, py, s;
      prior_of(geo, p, px, py, s);
      // ---- smooth EIoU on the decoded box
      const float cx = pv[1] * s + px, cy = pv[2] * s + py;
      const float w = expf(pv[3]) * s, h = expf(pv[4]) * s;
      const float px1 = cx - w / 2, py1 = cy - h / 2, px2 = cx + w / 2, py2 = cy + h / 2;
      const float tx1 = __ldg(gr), ty1 = __ldg(gr + 1), tx2 = __ldg(gr + 2), ty2 = __ldg(gr + 3);
      const float ex1 = fminf(px1, tx1), ey1 = fminf(py1, ty1);
      const float ix1 = fmaxf(px1, tx1), iy1 = fmaxf(py1, ty1);
      const float ix2 = fminf(px2, tx2), iy2 = fminf(py2, ty2);
      const float xmin = fminf(ix1, ix2), ymin = fminf(iy1, iy2);
      const float xmax = fmaxf(ix1, ix2), ymax = fmaxf(iy1, iy2);
      const float A_ = ix2 - ex1, B_ = iy2 - ey1, C_ = xmin - ex1, D_ = ymin - ey1;
      const float E_ = ix1 - ex1, F_ = ymax - ey1, G_ = xmax - ex1, H_ = iy1 - ey1;
      const float I = A_ * B_ + C_ * D_ - E_ * F_ - G_ * H_;
      const float U = (px2 - px1) * (py2 - py1) + (tx2 - tx1) * (ty2 - ty1) - I + lc.eiou_eps;
      const float q = 1.0f - I / U;
      const bool sm = q < lc.smooth_point;
      l_bbox = sm ? 0.5f * q * q / lc.smooth_point : q - 0.5f * lc.smooth_point;
      const float gq = (sm ? q / lc.smooth_point : 1.0f) * (lc.w_bbox * s_bbox * invN);
      // q = 1 - I/U, U = Ap + At - I + eps
      const float gI = gq * (-1.0f / U - I / (U * U));
      const float gAp = gq * (I / (U * U));
      float d_ix2 = gI * B_, d_iy2 = gI * A_, d_xmin = gI * D_, d_ymin = gI * C_;
      float d_ix1 = -gI * F_, d_ymax = -gI * E_, d_xmax = -gI * H_, d_iy1 = -gI * G_;
      float d_ex1 = gI * (-B_ - D_ + F_ + H_), d_ey1 = gI * (-A_ - C_ + E_ + G_);
      min_bwd(ix1, ix2, d_xmin, d_ix1, d_ix2);
      min_bwd(iy1, iy2, d_ymin, d_iy1, d_iy2);
      max_bwd(ix1, ix2, d_xmax, d_ix1, d_ix2);
      max_bwd(iy1, iy2, d_ymax, d_iy1, d_iy2);
      float d_px1 = 0.f, d_py1 = 0.f, d_px2 = 0.f, d_py2 = 0.f, dump = 0.f;
      max_bwd(px1, tx1, d_ix1, d_px1, dump);
      max_bwd(py1, ty1, d_iy1, d_py1, dump);
      min_bwd(px2, tx2, d_ix2, d_px2, dump);
      min_bwd(py2, ty2, d_iy2, d_py2, dump);
      min_bwd(px1, tx1, d_ex1, d_px1, dump);
      min_bwd(py1, ty1, d_ey1, d_py1, dump);
      d_px2 += gAp * (py2 - py1); d_px1 -= gAp * (py2 - py1);
      d_py2 += gAp * (px2 - px1); d_py1 -= gAp * (px2 - px1);
      dv[1] = (d_px1 + d_px2) * s;
      dv[2] = (d_py1 + d_py2) * s;
      dv[3] = 0.5f * (d_px2 - d_px1) * w;
      dv[4] = 0.5f * (d_py2 - d_py1) * h;
      // ---- landmarks: SmoothL1 on (kps - prior_xy)/stride, weight = mean visibility
      const float kw = (__ldg(gr + 14) + __ldg(gr + 15) + __ldg(gr + 16) + __ldg(gr + 17) + __ldg(gr + 18)) / 5.0f;
      const float gk = lc.w_kps * s_kps * kw * inv_kw;
#pragma unroll
      for (int k = 0; k < 10; ++k) {
        const float tgt = (__ldg(gr + 4 + k) - ((k & 1) ? py : px)) / s;
        const float d = pv[6 + k] - tgt;
        const float ad = fabsf(d);
        if (ad < lc.beta) {
          l_kps += 0.5f * ad * ad / lc.beta * kw;
          dv[6 + k] = d / lc.beta * gk;
        } else {
          l_kps += (ad - 0.5f * lc.beta) * kw;
          dv[6 + k] = (d > 0.f ? 1.f : -1.f) * gk;
        }
      }
    }
    if (d_preds != nullptr) {
      float4* dst = reinterpret_cast<float4*>(d_preds + i * PC);
#pragma unroll
      for (int q = 0; q < 4; ++q) dst[q] = make_float4(dv[q * 4], dv[q * 4 + 1], dv[q * 4 + 2], dv[q * 4 + 3]);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    l_cls += __shfl_xor_sync(0xffffffffu, l_cls, o);
    l_bbox += __shfl_xor_sync(0xffffffffu, l_bbox, o);
    l_obj += __shfl_xor_sync(0xffffffffu, l_obj, o);
    l_kps += __shfl_xor_sync(0xffffffffu, l_kps, o);
  }
  if (lane == 0) { s_red[0][warp] = l_cls; s_red[1][warp] = l_bbox; s_red[2][warp] = l_obj; s_red[3][warp] = l_kps; }
  __syncthreads();
  if (tid < 4) {
    float t = 0.f;
    for (int w = 0; w < NW; ++w) t += s_red[tid][w];
    const float scale = tid == 0 ? lc.w_cls * invN : tid == 1 ? lc.w_bbox * invN : tid == 2 ? lc.w_obj * invN : lc.w_kps * inv_kw;
    if (t != 0.f) atomicAdd(losses + tid, t * scale);
  }
}

}  // namespace

size_t simota_workspace_bytes(int B, int P) {
  return P <= 2112 ? 0 : (size_t)B * P * sizeof(Cand);
}

cudaError_t launch_simota_assign(const yunet_loss_cfg_dev& lc, const LevelGeom& g,
                                 const float* preds, const float* gt, const int* gt_offsets, int B,
                                 int* assigned, float* matched_iou, float* counters, void* ws,
                                 cudaStream_t s) {
  cudaError_t e = cudaMemsetAsync(counters, 0, 4 * sizeof(float), s);
  if (e != cudaSuccess) return e;
  const int vcap = g.P <= 2112 ? g.P : 6400;
  const size_t smem = (size_t)vcap * sizeof(Cand) + ((g.P + 15) & ~15);
  static size_t configured = 0;
  if (smem > configured) {
    e = cudaFuncSetAttribute(simota_assign_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    configured = smem;
  }
  simota_assign_kernel<false><<<B, NTA, smem, s>>>(lc, g, preds, gt, gt_offsets, assigned, matched_iou,
                                                  counters, reinterpret_cast<Cand*>(ws), vcap, AssignExt{});
  return cudaGetLastError();
}

// One image, explicit scores / offset priors / decoded boxes (the SimOTAAssigner.assign surface).
cudaError_t launch_simota_assign_ext(const yunet_loss_cfg_dev& lc, int P, const float* scores,
                                     const float* priors, const float* boxes, const float* gt,
                                     const int* gt_offsets, int* assigned, float* matched_iou,
                                     float* counters, void* ws, cudaStream_t s) {
  cudaError_t e = cudaMemsetAsync(counters, 0, 4 * sizeof(float), s);
  if (e != cudaSuccess) return e;
  LevelGeom g;
  memset(&g, 0, sizeof g);
  g.P = P;
  const int vcap = P <= 2112 ? P : 6400;
  const size_t smem = (size_t)vcap * sizeof(Cand) + ((P + 15) & ~15);
  if (smem > 227 * 1024) return cudaErrorInvalidValue;
  static size_t configured = 0;
  if (smem > configured) {
    e = cudaFuncSetAttribute(simota_assign_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    configured = smem;
  }
  AssignExt ext{scores, priors, boxes};
  simota_assign_kernel<true><<<1, NTA, smem, s>>>(lc, g, nullptr, gt, gt_offsets, assigned, matched_iou,
                                                 counters, reinterpret_cast<Cand*>(ws), vcap, ext);
  return cudaGetLastError();
}

cudaError_t launch_loss_grad(const yunet_loss_cfg_dev& lc, const LevelGeom& g, const float* preds,
                             const float* gt, const int* gt_offsets, const int* assigned,
                             const float* matched_iou, const float* counters,
                             const float* num_total, float s_cls, float s_bbox, float s_obj,
                             float s_kps, int B, float* losses, float* d_preds, cudaStream_t s) {
  cudaError_t e = cudaMemsetAsync(losses, 0, 4 * sizeof(float), s);
  if (e != cudaSuccess) return e;
  const long long n = (long long)B * g.P;
  const int blocks = (int)((n + NT - 1) / NT);
  loss_grad_kernel<<<blocks, NT, 0, s>>>(lc, g, preds, gt, gt_offsets, assigned, matched_iou,
                                         counters, num_total, s_cls, s_bbox, s_obj, s_kps, B,
                                         losses, d_preds);
  return cudaGetLastError();
}

}  // namespace yunet
