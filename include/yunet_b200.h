/*
 * yunet_b200.h — C ABI of libyunet_b200.so: the B200-native (sm_100a) YuNet hot path.
 *
 * Drop-in boundary for ShiqiYu/libfacedetection.train (reference @ 0047ac24).  The reference has
 * no native code (setup.py:215 `ext_modules=[]`); everything below replaces work the reference
 * delegates to ATen/cuDNN, mmcv._ext and NCCL through these Python call sites:
 *
 *   yunet_forward          mmdet/models/backbones/yunet_backbone.py:33-41 (YuNetBackbone.forward)
 *                          mmdet/models/necks/tfpn.py:33-45               (TFPN.forward)
 *                          mmdet/models/dense_heads/yunet_head.py:175-247 (YuNet_Head.forward)
 *                          mmdet/models/utils/yunet_layer.py:30-36,57-62  (ConvDPUnit, Conv_head)
 *                          + flatten/cat of yunet_head.py:456-477
 *   yunet_simota_assign    yunet_head.py:536-604 (_get_target_single, batched over images)
 *                          mmdet/core/bbox/assigners/sim_ota_assigner.py:95-257
 *                          mmdet/core/bbox/iou_calculators/iou2d_calculator.py:213-253
 *                          mmdet/core/anchor/point_generator.py:80-175, yunet_head.py:376-386
 *   yunet_loss_grad        yunet_head.py:493-534 with losses/iou_loss.py:194-227 (EIoU),
 *                          losses/cross_entropy_loss.py:117-145 (sigmoid BCE),
 *                          losses/smooth_l1_loss.py:24-32, losses/utils.py:42-59, and their autograd
 *   yunet_backward         autograd of everything yunet_forward computes (loss.backward() in
 *                          mmcv OptimizerHook, registered at mmdet/apis/train.py:182-198)
 *   yunet_sgd_step         torch.optim.SGD as configured in configs/yunet_n.py:1 + the DDP
 *                          gradient mean of mmdet/apis/train.py:156-161
 *   yunet_decode_nms       yunet_head.py:290-374 (get_bboxes) + :404-416 (_bboxes_nms ->
 *                          mmcv.ops.nms.batched_nms, mmcv-full 1.3.17..1.6.0)
 *   yunet_grid_priors      point_generator.py:80-175 (MlvlPointGenerator.grid_priors)
 *
 * Conventions: extern "C", plain pointers and sizes.  Every device buffer is caller-owned (the
 * library never allocates or frees device memory, and holds no mutable global state outside the
 * opaque ctx).  Every launch goes to the cudaStream_t passed in (as void*), no host
 * synchronisation happens inside.  Return value: 0 = ok, negative = argument/shape error,
 * positive = cudaError_t.  yunet_last_error(ctx) gives the message of the last failure.
 *
 * Layouts: images NCHW fp32 0..255 (configs/yunet_n.py:27); predictions (B, P, 16) fp32 with
 * channels [cls, bbox dx dy dw dh, obj, kps x0 y0 .. x4 y4] and priors ordered level-major
 * (stride 8,16,32), row-major inside a level; parameters one flat fp32 bucket whose sub-ranges
 * are exactly the reference state_dict tensors (OIHW), see yunet_param_info; activations NHWC
 * fp32 inside the caller-provided workspace.
 */
#ifndef YUNET_B200_H_
#define YUNET_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define YUNET_MAX_STAGES 8
#define YUNET_PRED_CH 16
#define YUNET_GT_ROW 19 /* x1 y1 x2 y2, 5x(kx ky), 5x weight */

typedef struct yunet_ctx yunet_ctx;

/* Architecture, mirroring the `model` dict of configs/yunet_{n,s}.py:104-145. */
typedef struct {
  int num_stages;                          /* len(stage_channels) */
  int stage_channels[YUNET_MAX_STAGES][3]; /* stage 0: {in, mid, out}; stage i>0: {in, out, 0} */
  int downsample_mask;                     /* bit i set: F.max_pool2d(x, 2) after stage i */
  int out_idx[3];                          /* backbone stages feeding the neck */
  int shared_stacked_convs;                /* bbox_head.shared_stacked_convs (0 or 1..) */
  int feat_channels;                       /* bbox_head.feat_channels */
  int num_classes;                         /* must be 1 */
  int kps_num;                             /* must be 5 */
  int strides[3];                          /* prior_generator.strides */
} yunet_arch_cfg;

/* SimOTA + loss hyper-parameters (configs/yunet_n.py:122-138, sim_ota_assigner.py:28-36). */
typedef struct {
  float center_radius;   /* 2.5 */
  int candidate_topk;    /* 10 */
  float iou_weight;      /* 3.0 */
  float cls_weight;      /* 1.0 */
  float loss_cls_weight; /* 1.0 */
  float loss_bbox_weight;/* 5.0 */
  float loss_obj_weight; /* 1.0 */
  float loss_kps_weight; /* 0.1 */
  float eiou_smooth_point; /* 0.1 */
  float eiou_eps;          /* 1e-6 */
  float smooth_l1_beta;    /* 0.1111111111111111 */
} yunet_loss_cfg;

int yunet_ctx_create(const yunet_arch_cfg* cfg, yunet_ctx** out);
void yunet_ctx_destroy(yunet_ctx* ctx);
const char* yunet_last_error(const yunet_ctx* ctx);
const char* yunet_version(void);

/* ---- parameter bucket description (host only) ------------------------------------------- */
long long yunet_num_params(const yunet_ctx* ctx);      /* floats in the flat parameter bucket */
int yunet_param_count(const yunet_ctx* ctx);           /* number of state_dict parameter tensors */
/* i-th tensor: reference state_dict key, offset (floats) in the bucket, shape (OIHW / (C,)) */
int yunet_param_info(const yunet_ctx* ctx, int i, char* name, int name_cap, long long* offset,
                     int* ndim, int shape[4]);
long long yunet_num_bn_channels(const yunet_ctx* ctx); /* running buffer = 2x this many floats:
                                                          [all running_mean | all running_var] */
int yunet_bn_count(const yunet_ctx* ctx);
/* i-th BatchNorm: state_dict prefix (e.g. "backbone.model0.bn1"), channel offset, channels */
int yunet_bn_info(const yunet_ctx* ctx, int i, char* name, int name_cap, long long* ch_offset,
                  int* channels);

/* ---- geometry ---------------------------------------------------------------------------- */
int yunet_num_priors(const yunet_ctx* ctx, int H, int W);
/* bytes of workspace yunet_forward/backward need for a (B,3,H,W) batch; train!=0 adds the
 * activation-gradient and statistics areas */
size_t yunet_workspace_bytes(const yunet_ctx* ctx, int B, int H, int W, int train);
/* priors (P,4) = [x, y, stride, stride], device pointer */
int yunet_grid_priors(yunet_ctx* ctx, int H, int W, float* priors, void* stream);

/* ---- forward ------------------------------------------------------------------------------
 * img (B,3,H,W) fp32 device; params: flat bucket; bn_running: 2*num_bn_channels floats.
 * train != 0: batch statistics (saved in ws for backward) and running-stat update with
 * `momentum` (torch BatchNorm2d semantics, unbiased running_var); train == 0: running stats.
 * preds: (B, P, 16).  ws: workspace of yunet_workspace_bytes(...). */
int yunet_forward(yunet_ctx* ctx, const float* img, const float* params, float* bn_running,
                  int B, int H, int W, int train, float momentum, float* preds, void* ws,
                  size_t ws_bytes, void* stream);

/* ---- SimOTA assignment (no grad) -----------------------------------------------------------
 * gt: (sumG, 19) rows, gt_offsets: (B+1) int32 CSR (every image needs >= 1 gt, as in the
 * reference).  Outputs per prior: assigned_gt (B,P) int32 = 1-based gt index inside the image or
 * 0 (AssignResult.gt_inds), matched_iou (B,P) fp32 (AssignResult.max_overlaps for positives,
 * 0 elsewhere).  counters (device, 4 floats, zeroed by this call): [num_pos, sum_kps_weight, 0, 0].
 * ws: scratch of yunet_assign_workspace_bytes (only used for images whose candidate set does not
 * fit in shared memory; NULL allowed when that size is 0). */
size_t yunet_assign_workspace_bytes(const yunet_ctx* ctx, int B, int H, int W); /* may be 0 */
int yunet_simota_assign(yunet_ctx* ctx, const yunet_loss_cfg* lc, const float* preds,
                        const float* gt, const int* gt_offsets, int B, int H, int W,
                        int* assigned_gt, float* matched_iou, float* counters, void* ws,
                        size_t ws_bytes, void* stream);

/* SimOTAAssigner.assign for ONE image with explicit inputs (sim_ota_assigner.py:38-93, as called by
 * yunet_head.py:575-577): scores (P) = sigmoid(cls)*sigmoid(obj), priors (P,4) = [cx, cy, stride_w,
 * stride_h] already offset by half a stride, decoded_boxes (P,4) xyxy, gt (G,19) rows (only the
 * box columns 0..3 are read), gt_offsets = {0, G}.  Outputs as yunet_simota_assign with B = 1.
 * ws: P * 32 bytes of scratch when P > 2112 (NULL otherwise). */
int yunet_simota_assign_ext(yunet_ctx* ctx, const yunet_loss_cfg* lc, int P, const float* scores,
                            const float* priors, const float* decoded_boxes, const float* gt,
                            const int* gt_offsets, int* assigned_gt, float* matched_iou,
                            float* counters, void* ws, size_t ws_bytes, void* stream);

/* ---- losses + d(loss)/d(preds) --------------------------------------------------------------
 * num_total_samples: device pointer to 1 float = reduce_mean(num_pos) over ranks BEFORE the
 * max(.,1) clamp (the caller all-reduces counters[0] and divides by world size; single process:
 * pass counters).  loss_scale[4]: upstream gradient of (loss_cls, loss_bbox, loss_obj, loss_kps)
 * (all 1 for `loss = sum`), host floats.  Outputs: losses (device, 4 floats,
 * [cls, bbox, obj, kps]) and d_preds (B,P,16) (may be NULL for loss only). */
int yunet_loss_grad(yunet_ctx* ctx, const yunet_loss_cfg* lc, const float* preds, const float* gt,
                    const int* gt_offsets, const int* assigned_gt, const float* matched_iou,
                    const float* counters, const float* num_total_samples, const float* loss_scale,
                    int B, int H, int W, float* losses, float* d_preds, void* stream);

/* ---- backward -------------------------------------------------------------------------------
 * Uses the activations saved in ws by the preceding yunet_forward(train=1) on the same img.
 * grad_bucket: flat fp32, same layout as params; overwritten (not accumulated). */
int yunet_backward(yunet_ctx* ctx, const float* img, const float* params, const float* d_preds,
                   int B, int H, int W, float* grad_bucket, void* ws, size_t ws_bytes,
                   void* stream);

/* ---- optimiser: g = grad*grad_scale + wd*w ; v = mom*v + g ; w -= lr*v  (torch.optim.SGD) --- */
int yunet_sgd_step(yunet_ctx* ctx, float* params, const float* grad_bucket, float* momentum_buf,
                   long long n, float lr, float momentum, float weight_decay, float grad_scale,
                   void* stream);

/* the same step with the learning rate read from device memory (1 float): every launch parameter of
 * a training iteration is then constant, so the iteration can be captured once and replayed as a
 * CUDA graph while the LR schedule (configs/yunet_n.py:4-11) changes lr per iteration */
int yunet_sgd_step_dev(yunet_ctx* ctx, float* params, const float* grad_bucket, float* momentum_buf,
                       long long n, const float* lr_dev, float momentum, float weight_decay,
                       float grad_scale, void* stream);

/* ---- decode + score filter + NMS -------------------------------------------------------------
 * scale_factors: NULL or (B,4) device floats dividing the boxes (rescale=True).  dets:
 * (B, max_det, 5) [x1,y1,x2,y2,score] in descending score order; det_kps: NULL or
 * (B, max_det, 10) decoded landmarks; det_count (B) int32 (number kept, may exceed max_det
 * in which case only max_det rows are written).  ws: yunet_nms_workspace_bytes(B,H,W). */
size_t yunet_nms_workspace_bytes(const yunet_ctx* ctx, int B, int H, int W);
int yunet_decode_nms(yunet_ctx* ctx, const float* preds, int B, int H, int W, float score_thr,
                     float iou_thr, const float* scale_factors, int max_det, float* dets,
                     float* det_kps, int* det_count, void* ws, size_t ws_bytes, void* stream);

/* ---- input pipeline: RandomSquareCrop + Resize + RandomFlip + Normalize ----------
 * The pixel work of mmdet/datasets/pipelines/transforms.py:1126-1146 (square crop, outside = 128),
 * :258-263 (mmcv.imresize, cv2 INTER_LINEAR on float32, keep_ratio=False) and :527-530
 * (horizontal flip) followed by Normalize(mean 0, std 1) + DefaultFormatBundle, for a ragged batch
 * of decoded images in ONE launch.  pixels: concatenated uint8 HWC (BGR) images; offsets (B) int64
 * byte offset of each image; hw (B,2) int32 height, width; crop (B,4) int32 [left, top, size,
 * flip] in source-image coordinates (left / top may be negative, size may exceed the image);
 * out: (B,3,S,S) fp32.  The random decisions are made by the caller (host mirror: pipeline.py). */
int yunet_preprocess_u8(yunet_ctx* ctx, const unsigned char* pixels, const long long* offsets,
                        const int* hw, const int* crop, int B, int S, float pad_value, float* out,
                        void* stream);

/* ---- debugging / parity helpers ----------------------------------------------------------------
 * yunet_unit_get describes the fused execution plan (used by the formulation tests and the host
 * mirror).  yunet_read_activation copies an internal NHWC activation (pre-BN output `z` of unit
 * `unit_index`, in execution order, -1 = stem) out of the workspace as NCHW with BatchNorm+ReLU
 * applied — what the reference module would have returned.  Used by the per-unit parity tests and
 * by the plugin classes' standalone forward. */
typedef struct {
  char name[96];          /* reference module path ("backbone.model2.conv1", ...) */
  int cin, cout;
  int mode;               /* 0 plain, 1 2x2 max-pool on load, 2 a + nearest_up2(b) on load */
  int in_a, in_b, out;    /* tensor ids (in_b = -1 unless mode 2) */
  int div;                /* output resolution = (H/div, W/div) */
  int has_bn;
  int acc_a, acc_b;       /* backward: accumulate into (1) or overwrite (0) the input gradient */
  int bn_out, bn_a, bn_b; /* BatchNorm indices of the output / inputs (-1: none) */
  int pred_level;         /* >= 0: output is level `pred_level` of the prediction tensor */
  long long w1, b1, w2, b2, gamma, beta; /* offsets (floats) into the parameter bucket */
} yunet_unit_desc;
int yunet_unit_count(const yunet_ctx* ctx);
/* unit_index in execution order; -1 = the stem conv (cin 3, cout 16, w1/b1 = conv weight/bias,
 * gamma/beta = bn1) */
int yunet_unit_get(const yunet_ctx* ctx, int unit_index, yunet_unit_desc* out);
int yunet_read_activation(yunet_ctx* ctx, int unit_index, const float* params,
                          const float* bn_running, int B, int H, int W, int train, const void* ws,
                          float* out_nchw, void* stream);

/* ---- measurement helpers (bench.py) ---------------------------------------------------------------
 * yunet_launch_count: kernels launched through this ctx so far.  yunet_profile_begin/end bracket
 * a region in which every kernel launch of the entry points above is timed with its own pair of
 * CUDA events on the launching stream; yunet_profile_end synchronises on them and returns the
 * number of records, yunet_profile_get returns name ("fwd:<unit>", "bwd:<unit>", ...), duration
 * and the algorithmic bytes of record i. */
long long yunet_launch_count(const yunet_ctx* ctx);
/* options (default 1): "tc_forward" runs the 64-input-channel units, "tc_backward" the 64->64
 * plain-load unit backward through the tcgen05/TMEM/TMA kernels (3xTF32, fp32-accurate); 0 selects
 * the exact-fp32 CUDA-core kernels.  Device-side bounded waits report into the status block
 * (yunet_ws_offset kind 3: int[64], [0] forward, [1] backward; all zero = ok). */
int yunet_set_option(yunet_ctx* ctx, const char* name, int value);
/* byte offset inside the workspace of tensor `tensor_id`'s pre-BN activation (kind 0), its
 * activation gradient (kind 1, train only) or the BatchNorm statistics block (kind 2: double
 * [4][num_bn_channels] = sum z, sum z^2, sum du, sum du*zhat) or the status block (kind 3);
 * -1 if absent.  Debug / tests. */
long long yunet_ws_offset(const yunet_ctx* ctx, int B, int H, int W, int train, int tensor_id,
                          int kind);
int yunet_profile_begin(yunet_ctx* ctx);
int yunet_profile_end(yunet_ctx* ctx);
int yunet_profile_get(const yunet_ctx* ctx, int i, char* name, int name_cap, float* ms,
                      double* algorithmic_bytes);

#ifdef __cplusplus
}
#endif
#endif /* YUNET_B200_H_ */
