// Device-side argument structs and host launchers of the YuNet hot-path kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>

#include <cstddef>
#include <cstdint>

namespace yunet {

constexpr float kBnEps = 1e-5f;  // torch.nn.BatchNorm2d default (yunet_layer.py:27,54)

// How to normalise a stored pre-BN tensor while loading it:  a = relu(z*scale + shift).
struct BnRef {
  const double* sum;    // train: batch sum / sum of squares per channel
  const double* sumsq;
  const float* rmean;   // eval: running statistics
  const float* rvar;
  const float* gamma;
  const float* beta;
  double inv_count;     // 1 / (B*H*W) of the normalised tensor
  int train;
};

struct UnitFwdArgs {
  const float* za;      // NHWC input (PLAIN: HxW, POOL: 2Hx2W, UPADD: HxW)
  const float* zb;      // UPADD only: NHWC at (H/2)x(W/2)
  BnRef bna, bnb;
  const float* w1;      // [COUT][CIN]
  const float* b1;      // [COUT]
  const float* w2;      // [COUT][3][3]
  const float* b2;      // [COUT]
  float* zout;          // NHWC-like: pixel stride COUT, image stride out_batch_stride
  long long out_batch_stride;
  double* osum;         // statistics of zout (nullptr: none)
  double* osumsq;
  int B, H, W;          // output resolution
};

struct UnitBwdArgs {
  // forward operands (recomputation)
  const float* za;
  const float* zb;
  BnRef bna, bnb;
  const float* w1; const float* b1; const float* w2;
  // this unit's output side
  const float* zout;            // pre-BN output saved by forward (has_bn only)
  const float* dout;            // has_bn: du (ReLU-masked grad wrt BN output); else d(loss)/d(z)
  long long dout_batch_stride;  // floats (image stride of dout)
  BnRef bno;                    // BN of the output (train statistics)
  const double* dsum;           // sum(du), sum(du*zhat) per output channel (has_bn only)
  const double* dsumzh;
  // gradients of the inputs
  float* dua; float* dub;
  int acc_a, acc_b;             // accumulate (1) or overwrite (0)
  double* dsum_a; double* dsumzh_a;   // statistics of dua for the producer's BN backward
  double* dsum_b; double* dsumzh_b;
  // parameter gradients: every CTA reduces its contributions in shared memory and writes ONE
  // partial vector [gw1 | gb1 | gw2 | gb2] to partial[blockIdx.x * kPartialStride]; a follow-up
  // reduce kernel sums the partials into the gradient bucket (no contended global atomics,
  // bit-reproducible sums)
  float* gw1; float* gb1; float* gw2; float* gb2;
  float* partial;
  int B, H, W;
  int has_bn;
};

struct StemArgs {
  const float* img;     // (B,3,Hin,Win) NCHW
  const float* w;       // [16][3][3][3]
  const float* b;       // [16]
  float* zout;          // (B,Hin/2,Win/2,16) NHWC
  double* osum; double* osumsq;
  int B, Hin, Win;
};

constexpr int kPartialStride = 5120;   // floats per CTA partial (>= 64*64 + 11*64)
constexpr int kMaxPartialCtas = 512;

struct StemBwdArgs {
  const float* img;
  const float* zout;     // pre-BN stem output
  const float* du;       // masked grad wrt BN output
  BnRef bno;
  const double* dsum; const double* dsumzh;
  float* gw; float* gb;
  float* partial;
  int B, Hin, Win;
};

constexpr int kMaxBn = 40;
struct BnFinalizeArgs {
  int n;
  int C[kMaxBn];
  long long ch_off[kMaxBn];
  double count[kMaxBn];
  long long gamma_off[kMaxBn];
  long long beta_off[kMaxBn];
};

// ---- launchers (kernels_fwd.cu) ----
cudaError_t launch_unit_fwd(int cin, int cout, int mode, const UnitFwdArgs& a, int num_sms,
                            cudaStream_t s);
cudaError_t launch_stem_fwd(const StemArgs& a, int num_sms, cudaStream_t s);
cudaError_t launch_bn_update_running(const BnFinalizeArgs& a, const double* sum,
                                     const double* sumsq, float* rmean, float* rvar,
                                     float momentum, cudaStream_t s);
cudaError_t launch_read_activation(const float* z, const BnRef& bn, int has_bn, int B, int H,
                                   int W, int C, long long batch_stride, float* out_nchw,
                                   cudaStream_t s);
cudaError_t launch_grid_priors(float* priors, const int* lh, const int* lw, const int* strides,
                               cudaStream_t s);
int unit_fwd_supported(int cin, int cout);

// ---- unit_fwd_tc.cu: tcgen05 / TMEM / TMA version of the fused unit (CIN = 64, plain load) ----
int unit_fwd_tc_supported(int cin, int cout, int mode);
cudaError_t launch_unit_fwd_tc(int cout, int mode, const UnitFwdArgs& a, int num_sms, int* status,
                               cudaStream_t s);

// ---- unit_fwd_ws.cu: warp-specialised streaming version (strips, TMA ring, TMEM double buffers) ----
int unit_fwd_ws_supported(int cin, int cout, int mode);
cudaError_t launch_unit_fwd_ws(int cin, int cout, int mode, const UnitFwdArgs& a, int num_sms, int* status,
                               cudaStream_t s);

// ---- unit_bwd_tc.cu: tcgen05 version of the fused unit backward (64 -> 64, plain load, BN) ----
int unit_bwd_tc_supported(int cin, int cout, int mode, int has_bn);
cudaError_t launch_unit_bwd_tc(int mode, const UnitBwdArgs& a, int num_sms, int* status,
                               cudaStream_t s);

// ---- unit_bwd_st.cu: the tcgen05 backward on strips (register-window depthwise stage, 94 % tile use) ----
int unit_bwd_st_supported(int cin, int cout, int mode, int has_bn, int H, int W);
int unit_bwd_st_preferred(int mode, int H, int W);   // shapes where it beats the per-tile kernel
cudaError_t launch_unit_bwd_st(int mode, const UnitBwdArgs& a, int num_sms, int* status, cudaStream_t s);

// ---- launchers (kernels_bwd.cu) ----
cudaError_t launch_unit_bwd(int cin, int cout, int mode, const UnitBwdArgs& a, int num_sms,
                            cudaStream_t s);
cudaError_t launch_stem_bwd(const StemBwdArgs& a, int num_sms, cudaStream_t s);
// out[i] = sum_c partial[c * kPartialStride + i], i < n   (deterministic parameter-gradient flush)
cudaError_t launch_reduce_partials(const float* partial, int ncta, int n, float* out, cudaStream_t s);
// d(gamma) = sum(du*zhat), d(beta) = sum(du) for every BN, from the statistics area
cudaError_t launch_bn_param_grads(const BnFinalizeArgs& a, const double* dsum,
                                  const double* dsumzh, float* grad_bucket, cudaStream_t s);

// ---- simota.cu ----
struct yunet_loss_cfg_dev {
  float center_radius; int candidate_topk; float iou_weight, cls_weight;
  float w_cls, w_bbox, w_obj, w_kps; float smooth_point, eiou_eps, beta;
};
struct LevelGeom {
  int off[3]; int h[3]; int w[3]; int stride[3]; int P;
};
cudaError_t launch_simota_assign(const yunet_loss_cfg_dev& lc, const LevelGeom& g,
                                 const float* preds, const float* gt, const int* gt_offsets, int B,
                                 int* assigned, float* matched_iou, float* counters, void* ws,
                                 cudaStream_t s);
cudaError_t launch_simota_assign_ext(const yunet_loss_cfg_dev& lc, int P, const float* scores,
                                     const float* priors, const float* boxes, const float* gt,
                                     const int* gt_offsets, int* assigned, float* matched_iou,
                                     float* counters, void* ws, cudaStream_t s);
size_t simota_workspace_bytes(int B, int P);
cudaError_t launch_loss_grad(const yunet_loss_cfg_dev& lc, const LevelGeom& g, const float* preds,
                             const float* gt, const int* gt_offsets, const int* assigned,
                             const float* matched_iou, const float* counters,
                             const float* num_total, float s_cls, float s_bbox, float s_obj,
                             float s_kps, int B, float* losses, float* d_preds, cudaStream_t s);

// ---- nms.cu ----
size_t nms_workspace_bytes(int B, int P);
cudaError_t launch_decode_nms(const LevelGeom& g, const float* preds, int B, float score_thr,
                              float iou_thr, const float* scale_factors, int max_det, float* dets,
                              float* det_kps, int* det_count, void* ws, cudaStream_t s);

// ---- sgd.cu ----
cudaError_t launch_preprocess_u8(const unsigned char* pixels, const long long* offsets, const int* hw,
                                 const int* crop, int B, int S, float pad, float* out, cudaStream_t s);
cudaError_t launch_sgd_dev(float* params, const float* grad, float* mom, long long n, const float* lr_dev,
                           float momentum, float wd, float grad_scale, cudaStream_t s);
cudaError_t launch_sgd(float* params, const float* grad, float* mom, long long n, float lr,
                       float momentum, float wd, float grad_scale, cudaStream_t s);

}  // namespace yunet
