// Decode + score filter + greedy NMS, one CTA per image, sm_100a.
//
// Replaces YuNet_Head.get_bboxes / _bboxes_nms (mmdet/models/dense_heads/yunet_head.py:290-374,
// 404-416) and the third-party mmcv.ops.nms.batched_nms it calls (mmcv-full 1.3.17..1.6.0; single
// class so the class offset is a no-op): keep priors with sigmoid(obj)*sigmoid(cls) >= score_thr,
// score = sigmoid(cls)*sigmoid(obj), sort by score descending (stable: lower prior index first),
// greedily suppress IoU > iou_thr with IoU = inter / (a + b - inter) (no +1 offset), emit
// [x1,y1,x2,y2,score] in score order.  fp32 operation order follows the CPU kernel the reference
// ends up in, so results are bit-comparable.
#include <cstdint>
#include <cstdio>

#include "kernels.h"

namespace yunet {

namespace {

constexpr int NT = 512;
constexpr int NW = NT / 32;
constexpr int PC = 16;

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

__device__ __forceinline__ unsigned ordered_bits(float f) {
  const unsigned b = __float_as_uint(f);
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}

// GLOBAL = false: keys / boxes live in shared memory (P <= 8704: every training / 640x640 shape).
// GLOBAL = true : the same algorithm on a per-image scratch area of the caller's workspace (keys
// [npow2] u64, then boxes / scores / indices / flags [P]) for origin-size evaluation inputs
// (WIDER test modes 1 / 2, tools/test_widerface.py:80-92: tens of thousands of priors per image).
template <bool GLOBAL>
__global__ void __launch_bounds__(NT) decode_nms_kernel(
    const LevelGeom geo, const float* __restrict__ preds, float score_thr, float iou_thr,
    const float* __restrict__ scale_factors, int max_det, float* __restrict__ dets,
    float* __restrict__ det_kps, int* __restrict__ det_count, int npow2_cap,
    unsigned char* __restrict__ scratch, size_t scratch_stride) {
  extern __shared__ float4 smem_raw[];
  // region A: sort keys (u64) [npow2]  — shared memory: later reused as boxes float4[n] + scores[n] + idx[n]
  unsigned char* region = GLOBAL ? scratch + (size_t)blockIdx.x * scratch_stride
                                 : reinterpret_cast<unsigned char*>(smem_raw);
  unsigned long long* skey = reinterpret_cast<unsigned long long*>(region);
  __shared__ int s_warp[NW];
  __shared__ int s_base;
  __shared__ int s_n;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int b = blockIdx.x;
  const int P = geo.P;
  const float* pb = preds + (long long)b * P * PC;

  // ---- phase 1: score filter + ordered compaction of (score, prior) keys
  if (tid == 0) s_base = 0;
  __syncthreads();
  for (int base = 0; base < P; base += NT) {
    const int p = base + tid;
    bool v = false;
    float score = 0.f;
    if (p < P) {
      const float cls = sigmoid_ref(__ldg(pb + (long long)p * PC + 0));
      const float obj = sigmoid_ref(__ldg(pb + (long long)p * PC + 5));
      score = __fmul_rn(cls, obj);             // max_scores * score_factor (yunet_head.py:409)
      v = __fmul_rn(obj, cls) >= score_thr;    // score_factor * max_scores >= thr (:406)
    }
    const unsigned m = __ballot_sync(0xffffffffu, v);
    if (lane == 0) s_warp[warp] = __popc(m);
    __syncthreads();
    int off = s_base;
    for (int w = 0; w < warp; ++w) off += s_warp[w];
    if (v) {
      const int slot = off + __popc(m & ((1u << lane) - 1u));
      // ascending key order == descending score, ascending prior index
      skey[slot] = ((unsigned long long)(~ordered_bits(score)) << 32) | (unsigned)p;
    }
    __syncthreads();
    if (tid == 0) {
      int t = 0;
      for (int w = 0; w < NW; ++w) t += s_warp[w];
      s_base += t;
    }
    __syncthreads();
  }
  const int n = s_base;
  if (n == 0) {
    if (tid == 0) det_count[b] = 0;
    return;
  }
  int np2 = 1;
  while (np2 < n) np2 <<= 1;
  for (int i = n + tid; i < np2; i += NT) skey[i] = ~0ull;
  __syncthreads();
  // ---- phase 2: bitonic sort of the keys
  for (int k = 2; k <= np2; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = tid; i < np2; i += NT) {
        const int ixj = i ^ j;
        if (ixj > i) {
          const unsigned long long a = skey[i], c = skey[ixj];
          const bool up = (i & k) == 0;
          if ((a > c) == up) { skey[i] = c; skey[ixj] = a; }
        }
      }
      __syncthreads();
    }
  }
  // ---- phase 3: gather decoded boxes in sorted order (keys -> registers -> boxes over the keys)
  // each thread owns candidates tid, tid+NT, ...; read all keys first, then overwrite the region
  constexpr int MAXPT = 17;   // ceil(8704 / 512)
  unsigned long long mykeys[MAXPT];
  if (!GLOBAL) {
#pragma unroll
    for (int c = 0; c < MAXPT; ++c) {
      const int i = tid + c * NT;
      mykeys[c] = i < n ? skey[i] : 0ull;
    }
    __syncthreads();
  }
  float4* sbox = reinterpret_cast<float4*>(GLOBAL ? region + (size_t)npow2_cap * 8 : region);   // [n]
  float* sscore = reinterpret_cast<float*>(sbox + n);            // [n]
  int* sidx = reinterpret_cast<int*>(sscore + n);                // [n]
  unsigned char* ssup = reinterpret_cast<unsigned char*>(sidx + n);  // [n]
  auto gather = [&](int i, unsigned long long key) {
    const int p = (int)(key & 0xffffffffu);
    const float* pr = pb + (long long)p * PC;
    float px, py, s;
    prior_of(geo, p, px, py, s);
    const float cx = __fadd_rn(__fmul_rn(__ldg(pr + 1), s), px);
    const float cy = __fadd_rn(__fmul_rn(__ldg(pr + 2), s), py);
    const float w = __fmul_rn(expf(__ldg(pr + 3)), s);
    const float h = __fmul_rn(expf(__ldg(pr + 4)), s);
    float4 bx;
    bx.x = __fsub_rn(cx, __fdiv_rn(w, 2.0f)); bx.y = __fsub_rn(cy, __fdiv_rn(h, 2.0f));
    bx.z = __fadd_rn(cx, __fdiv_rn(w, 2.0f)); bx.w = __fadd_rn(cy, __fdiv_rn(h, 2.0f));
    if (scale_factors != nullptr) {            // yunet_head.py:359-363
      const float* sf = scale_factors + b * 4;
      bx.x = __fdiv_rn(bx.x, __ldg(sf)); bx.y = __fdiv_rn(bx.y, __ldg(sf + 1));
      bx.z = __fdiv_rn(bx.z, __ldg(sf + 2)); bx.w = __fdiv_rn(bx.w, __ldg(sf + 3));
    }
    sbox[i] = bx;
    sscore[i] = __fmul_rn(sigmoid_ref(__ldg(pr + 0)), sigmoid_ref(__ldg(pr + 5)));
    sidx[i] = p;
    ssup[i] = 0;
  };
  if (GLOBAL) {
    for (int i = tid; i < n; i += NT) gather(i, skey[i]);
  } else {
#pragma unroll
    for (int c = 0; c < MAXPT; ++c) {
      const int i = tid + c * NT;
      if (i < n) gather(i, mykeys[c]);
    }
  }
  __syncthreads();
  // ---- phase 4: greedy suppression in score order
  for (int i = 0; i < n; ++i) {
    if (ssup[i]) continue;            // uniform: written before the last barrier
    const float4 bi = sbox[i];
    const float iarea = __fmul_rn(__fsub_rn(bi.z, bi.x), __fsub_rn(bi.w, bi.y));
    for (int j = i + 1 + tid; j < n; j += NT) {
      if (ssup[j]) continue;
      const float4 bj = sbox[j];
      const float xx1 = fmaxf(bi.x, bj.x), yy1 = fmaxf(bi.y, bj.y);
      const float xx2 = fminf(bi.z, bj.z), yy2 = fminf(bi.w, bj.w);
      const float w = fmaxf(0.f, __fsub_rn(xx2, xx1)), h = fmaxf(0.f, __fsub_rn(yy2, yy1));
      const float inter = __fmul_rn(w, h);
      const float jarea = __fmul_rn(__fsub_rn(bj.z, bj.x), __fsub_rn(bj.w, bj.y));
      const float ovr = __fdiv_rn(inter, __fsub_rn(__fadd_rn(iarea, jarea), inter));
      if (ovr > iou_thr) ssup[j] = 1;
    }
    __syncthreads();
  }
  // ---- phase 5: ordered output of the survivors
  if (tid == 0) s_base = 0;
  __syncthreads();
  for (int base = 0; base < n; base += NT) {
    const int i = base + tid;
    const bool keep = i < n && ssup[i] == 0;
    const unsigned m = __ballot_sync(0xffffffffu, keep);
    if (lane == 0) s_warp[warp] = __popc(m);
    __syncthreads();
    int off = s_base;
    for (int w = 0; w < warp; ++w) off += s_warp[w];
    if (keep) {
      const int slot = off + __popc(m & ((1u << lane) - 1u));
      if (slot < max_det) {
        const float4 bx = sbox[i];
        float* d = dets + ((long long)b * max_det + slot) * 5;
        d[0] = bx.x; d[1] = bx.y; d[2] = bx.z; d[3] = bx.w; d[4] = sscore[i];
        if (det_kps != nullptr) {
          const int p = sidx[i];
          const float* pr = pb + (long long)p * PC + 6;
          float px, py, s;
          prior_of(geo, p, px, py, s);
          float* dk = det_kps + ((long long)b * max_det + slot) * 10;
#pragma unroll
          for (int k = 0; k < 10; ++k)   // _kps_decode, yunet_head.py:388-393
            dk[k] = __fadd_rn(__fmul_rn(__ldg(pr + k), s), (k & 1) ? py : px);
        }
      }
    }
    __syncthreads();
    if (tid == 0) {
      int t = 0;
      for (int w = 0; w < NW; ++w) t += s_warp[w];
      s_base += t;
    }
    __syncthreads();
  }
  // rows beyond max_det were not written: the count a caller may index with is clamped
  if (tid == 0) det_count[b] = s_base < max_det ? s_base : max_det;
}

}  // namespace

constexpr int kNmsSharedMaxP = 512 * 17;   // MAXPT bound of the shared-memory variant

static size_t nms_scratch_stride(int P) {
  int np2 = 1;
  while (np2 < P) np2 <<= 1;
  size_t b = (size_t)np2 * 8 + (size_t)P * (16 + 4 + 4 + 1) + 64;
  return (b + 255) / 256 * 256;
}

size_t nms_workspace_bytes(int B, int P) {
  if (P <= kNmsSharedMaxP) return 256;   // everything lives in shared memory; non-zero size for the ABI
  return (size_t)B * nms_scratch_stride(P) + 256;
}

cudaError_t launch_decode_nms(const LevelGeom& g, const float* preds, int B, float score_thr,
                              float iou_thr, const float* scale_factors, int max_det, float* dets,
                              float* det_kps, int* det_count, void* ws, cudaStream_t s) {
  int np2 = 1;
  while (np2 < g.P) np2 <<= 1;
  if (g.P > kNmsSharedMaxP) {
    unsigned char* scratch = reinterpret_cast<unsigned char*>(
        (reinterpret_cast<uintptr_t>(ws) + 255) & ~uintptr_t(255));
    decode_nms_kernel<true><<<B, NT, 0, s>>>(g, preds, score_thr, iou_thr, scale_factors, max_det, dets,
                                             det_kps, det_count, np2, scratch, nms_scratch_stride(g.P));
    return cudaGetLastError();
  }
  size_t a = (size_t)np2 * sizeof(unsigned long long);
  size_t bbytes = (size_t)g.P * (16 + 4 + 4 + 1) + 64;
  size_t smem = a > bbytes ? a : bbytes;
  static size_t configured = 0;
  if (smem > configured) {
    cudaError_t e = cudaFuncSetAttribute(decode_nms_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    configured = smem;
  }
  decode_nms_kernel<false><<<B, NT, smem, s>>>(g, preds, score_thr, iou_thr, scale_factors, max_det, dets,
                                               det_kps, det_count, np2, nullptr, 0);
  return cudaGetLastError();
}

}  // namespace yunet
