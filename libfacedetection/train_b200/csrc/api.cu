// C-ABI of libyunet_b200.so (declared in include/yunet_b200.h): ctx, plan queries, and the
// orchestration of the fused-unit kernels for forward / backward.  No device memory is allocated
// here; every launch goes to the caller's stream.
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <new>
#include <string>
#include <vector>

#include "kernels.h"
#include "plan.h"

using namespace yunet;

struct ProfEvent {
  std::string name;
  cudaEvent_t start = nullptr, stop = nullptr;
  double bytes = 0.0;   // algorithmic bytes of the launch, 0 if not a streaming kernel
};

struct yunet_ctx {
  Plan plan;
  std::string err;
  int num_sms = 148;
  bool sms_known = false;
  long long launches = 0;       // kernels launched by this ctx (bench.py's gpu_launches)
  int opt_tc_forward = 1;       // use the tcgen05 unit kernel where it applies (default on)
  int opt_tc_backward = 1;      // same for the unit backward (64->64 plain units)
  int opt_st_backward = 1;      // strip-streaming tcgen05 unit backward (unit_bwd_st.cu) where it applies
  int opt_ws_forward = 1;       // warp-specialised streaming unit kernel (unit_fwd_ws.cu) where it applies
  bool profiling = false;
  std::vector<ProfEvent> prof;
  std::vector<float> prof_ms;
};

namespace {

int fail(yunet_ctx* c, int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  if (c) c->err = buf;
  return code;
}

int cuda_fail(yunet_ctx* c, cudaError_t e, const char* what) {
  if (e == cudaSuccess) return 0;
  return fail(c, (int)e, "%s: %s", what, cudaGetErrorString(e));
}

bool shape_ok(const yunet_ctx* c, int B, int H, int W) {
  const int m = c->plan.cfg.strides[2];
  return B > 0 && H > 0 && W > 0 && H % m == 0 && W % m == 0;
}

void ensure_sms(yunet_ctx* c) {
  if (c->sms_known) return;
  int dev = 0, n = 0;
  if (cudaGetDevice(&dev) == cudaSuccess &&
      cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0)
    c->num_sms = n;
  c->sms_known = true;
}

struct Views {
  const Plan& p;
  const WsLayout& L;
  char* ws;
  const float* params;
  const float* bn_running;
  float* preds;
  const float* d_preds;
  int B, H, W, train;

  float* z(int t) const { return reinterpret_cast<float*>(ws + L.z_off[t]); }
  float* du(int t) const { return reinterpret_cast<float*>(ws + L.du_off[t]); }
  int* status() const { return reinterpret_cast<int*>(ws + L.status_off); }
  float* partial() const { return reinterpret_cast<float*>(ws + L.partial_off); }
  double* stat(int which) const {
    return reinterpret_cast<double*>(ws + L.stats_off) + (size_t)which * p.num_bn_ch;
  }
  BnRef bnref(int t) const {
    BnRef r;
    memset(&r, 0, sizeof r);
    const TensorDesc& td = p.tensors[t];
    if (td.bn < 0) return r;
    const BnDesc& bn = p.bns[td.bn];
    r.sum = stat(0) + bn.ch_off;
    r.sumsq = stat(1) + bn.ch_off;
    r.rmean = bn_running ? bn_running + bn.ch_off : nullptr;
    r.rvar = bn_running ? bn_running + p.num_bn_ch + bn.ch_off : nullptr;
    r.gamma = params + bn.gamma;
    r.beta = params + bn.beta;
    r.inv_count = 1.0 / ((double)B * (H / td.div) * (W / td.div));
    r.train = train;
    return r;
  }
};

BnFinalizeArgs make_bn_args(const Plan& p, int B, int H, int W) {
  BnFinalizeArgs a;
  memset(&a, 0, sizeof a);
  a.n = (int)p.bns.size();
  for (int i = 0; i < a.n; ++i) {
    const BnDesc& bn = p.bns[i];
    const TensorDesc& td = p.tensors[bn.tensor];
    a.C[i] = bn.C;
    a.ch_off[i] = bn.ch_off;
    a.count[i] = (double)B * (H / td.div) * (W / td.div);
    a.gamma_off[i] = bn.gamma;
    a.beta_off[i] = bn.beta;
  }
  return a;
}

yunet_loss_cfg_dev to_dev(const yunet_loss_cfg* lc) {
  yunet_loss_cfg_dev d;
  d.center_radius = lc->center_radius;
  d.candidate_topk = lc->candidate_topk;
  d.iou_weight = lc->iou_weight;
  d.cls_weight = lc->cls_weight;
  d.w_cls = lc->loss_cls_weight;
  d.w_bbox = lc->loss_bbox_weight;
  d.w_obj = lc->loss_obj_weight;
  d.w_kps = lc->loss_kps_weight;
  d.smooth_point = lc->eiou_smooth_point;
  d.eiou_eps = lc->eiou_eps;
  d.beta = lc->smooth_l1_beta;
  return d;
}

LevelGeom make_geom(const Plan& p, int H, int W) {
  LevelGeom g;
  int off = 0;
  for (int l = 0; l < 3; ++l) {
    g.stride[l] = p.cfg.strides[l];
    g.h[l] = H / g.stride[l];
    g.w[l] = W / g.stride[l];
    g.off[l] = off;
    off += g.h[l] * g.w[l];
  }
  g.P = off;
  return g;
}

// RAII: counts the launch and, when profiling, brackets it with CUDA events on the stream.
struct Scope {
  yunet_ctx* c;
  cudaStream_t s;
  int idx = -1;
  Scope(yunet_ctx* c_, cudaStream_t s_, const std::string& name, double bytes = 0.0) : c(c_), s(s_) {
    c->launches++;
    if (!c->profiling) return;
    ProfEvent e;
    e.name = name;
    e.bytes = bytes;
    if (cudaEventCreate(&e.start) != cudaSuccess || cudaEventCreate(&e.stop) != cudaSuccess) return;
    cudaEventRecord(e.start, s);
    c->prof.push_back(e);
    idx = (int)c->prof.size() - 1;
  }
  ~Scope() {
    if (idx >= 0) cudaEventRecord(c->prof[idx].stop, s);
  }
};

void copy_name(char* dst, int cap, const std::string& s) {
  if (!dst || cap <= 0) return;
  snprintf(dst, (size_t)cap, "%s", s.c_str());
}

}  // namespace

extern "C" {

const char* yunet_version(void) { return "yunet_b200 0.1 (sm_100a)"; }

int yunet_ctx_create(const yunet_arch_cfg* cfg, yunet_ctx** out) {
  if (!cfg || !out) return -1;
  yunet_ctx* c = new (std::nothrow) yunet_ctx();
  if (!c) return -2;
  if (!c->plan.build(*cfg)) {
    // keep the ctx alive so the caller can read the message
    c->err = c->plan.error;
    *out = c;
    return -3;
  }
  if (c->plan.bns.size() > (size_t)kMaxBn) {
    c->err = "too many BatchNorm layers";
    *out = c;
    return -3;
  }
  *out = c;
  return 0;
}

void yunet_ctx_destroy(yunet_ctx* ctx) {
  if (!ctx) return;
  for (ProfEvent& e : ctx->prof) { cudaEventDestroy(e.start); cudaEventDestroy(e.stop); }
  delete ctx;
}

const char* yunet_last_error(const yunet_ctx* ctx) { return ctx ? ctx->err.c_str() : "null ctx"; }

long long yunet_num_params(const yunet_ctx* ctx) { return ctx->plan.num_params; }
int yunet_param_count(const yunet_ctx* ctx) { return (int)ctx->plan.params.size(); }

int yunet_param_info(const yunet_ctx* ctx, int i, char* name, int name_cap, long long* offset,
                     int* ndim, int shape[4]) {
  if (i < 0 || i >= (int)ctx->plan.params.size()) return -1;
  const ParamInfo& pi = ctx->plan.params[i];
  copy_name(name, name_cap, pi.name);
  if (offset) *offset = pi.offset;
  if (ndim) *ndim = pi.ndim;
  if (shape) for (int k = 0; k < 4; ++k) shape[k] = pi.shape[k];
  return 0;
}

long long yunet_num_bn_channels(const yunet_ctx* ctx) { return ctx->plan.num_bn_ch; }
int yunet_bn_count(const yunet_ctx* ctx) { return (int)ctx->plan.bns.size(); }

int yunet_bn_info(const yunet_ctx* ctx, int i, char* name, int name_cap, long long* ch_offset,
                  int* channels) {
  if (i < 0 || i >= (int)ctx->plan.bns.size()) return -1;
  const BnDesc& b = ctx->plan.bns[i];
  copy_name(name, name_cap, b.name);
  if (ch_offset) *ch_offset = b.ch_off;
  if (channels) *channels = b.C;
  return 0;
}

int yunet_num_priors(const yunet_ctx* ctx, int H, int W) {
  if (!shape_ok(ctx, 1, H, W)) return -1;
  return make_geom(ctx->plan, H, W).P;
}

size_t yunet_workspace_bytes(const yunet_ctx* ctx, int B, int H, int W, int train) {
  if (!shape_ok(ctx, B, H, W)) return 0;
  return make_layout(ctx->plan, B, H, W, train != 0).total;
}

int yunet_grid_priors(yunet_ctx* ctx, int H, int W, float* priors, void* stream) {
  if (!shape_ok(ctx, 1, H, W) || !priors) return fail(ctx, -1, "grid_priors: bad arguments");
  LevelGeom g = make_geom(ctx->plan, H, W);
  return cuda_fail(ctx, launch_grid_priors(priors, g.h, g.w, g.stride, (cudaStream_t)stream),
                   "grid_priors");
}

int yunet_unit_count(const yunet_ctx* ctx) { return (int)ctx->plan.units.size(); }

int yunet_unit_get(const yunet_ctx* ctx, int i, yunet_unit_desc* d) {
  const Plan& p = ctx->plan;
  if (!d || i < -1 || i >= (int)p.units.size()) return -1;
  memset(d, 0, sizeof *d);
  if (i == -1) {
    snprintf(d->name, sizeof d->name, "backbone.model0.conv1");
    d->cin = p.stem_cin; d->cout = p.stem_cout; d->mode = 0; d->in_a = -1; d->in_b = -1;
    d->out = p.stem_out; d->div = 2; d->has_bn = 1;
    d->bn_out = p.tensors[p.stem_out].bn; d->bn_a = -1; d->bn_b = -1; d->pred_level = -1;
    d->w1 = p.stem_w; d->b1 = p.stem_b;
    d->gamma = p.bns[d->bn_out].gamma; d->beta = p.bns[d->bn_out].beta;
    return 0;
  }
  const UnitDesc& u = p.units[i];
  snprintf(d->name, sizeof d->name, "%s", u.name.c_str());
  d->cin = u.cin; d->cout = u.cout; d->mode = u.mode; d->in_a = u.in_a; d->in_b = u.in_b;
  d->out = u.out; d->div = u.div; d->has_bn = u.has_bn ? 1 : 0;
  d->acc_a = u.acc_a ? 1 : 0; d->acc_b = u.acc_b ? 1 : 0;
  d->bn_out = p.tensors[u.out].bn;
  d->bn_a = p.tensors[u.in_a].bn;
  d->bn_b = u.in_b >= 0 ? p.tensors[u.in_b].bn : -1;
  d->pred_level = p.tensors[u.out].pred_level;
  d->w1 = u.w1; d->b1 = u.b1; d->w2 = u.w2; d->b2 = u.b2; d->gamma = u.gamma; d->beta = u.beta;
  return 0;
}

int yunet_forward(yunet_ctx* ctx, const float* img, const float* params, float* bn_running,
                  int B, int H, int W, int train, float momentum, float* preds, void* ws,
                  size_t ws_bytes, void* stream) {
  if (!ctx) return -1;
  if (!img || !params || !preds || !ws) return fail(ctx, -1, "forward: null pointer");
  if (!train && !bn_running) return fail(ctx, -1, "forward: eval mode needs bn_running");
  if (!shape_ok(ctx, B, H, W)) return fail(ctx, -2, "forward: H and W must be multiples of %d", ctx->plan.cfg.strides[2]);
  const Plan& p = ctx->plan;
  WsLayout L = make_layout(p, B, H, W, train != 0);
  if (ws_bytes < L.total) return fail(ctx, -3, "forward: workspace too small (%zu < %zu)", ws_bytes, L.total);
  cudaStream_t s = (cudaStream_t)stream;
  Views v{p, L, (char*)ws, params, bn_running, preds, nullptr, B, H, W, train};
  cudaError_t e;
  ensure_sms(ctx);
  if (train) {
    e = cudaMemsetAsync((char*)ws + L.stats_off, 0, L.stats_bytes, s);
    if (e != cudaSuccess) return cuda_fail(ctx, e, "forward: memset statistics");
  }
  if (ctx->opt_tc_forward || ctx->opt_ws_forward) {
    e = cudaMemsetAsync((char*)ws + L.status_off, 0, 256, s);
    if (e != cudaSuccess) return cuda_fail(ctx, e, "forward: memset status");
  }
  {
    StemArgs a;
    a.img = img; a.w = params + p.stem_w; a.b = params + p.stem_b; a.zout = v.z(p.stem_out);
    const BnDesc& bn = p.bns[p.tensors[p.stem_out].bn];
    a.osum = train ? v.stat(0) + bn.ch_off : nullptr;
    a.osumsq = train ? v.stat(1) + bn.ch_off : nullptr;
    a.B = B; a.Hin = H; a.Win = W;
    {
      Scope sc(ctx, s, "fwd:stem", 4.0 * B * (3.0 * H * W + 16.0 * (H / 2) * (W / 2)));
      e = launch_stem_fwd(a, ctx->num_sms, s);
    }
    if (e != cudaSuccess) return cuda_fail(ctx, e, "forward: stem");
  }
  for (const UnitDesc& u : p.units) {
    UnitFwdArgs a;
    memset(&a, 0, sizeof a);
    a.za = v.z(u.in_a);
    a.bna = v.bnref(u.in_a);
    if (u.in_b >= 0) { a.zb = v.z(u.in_b); a.bnb = v.bnref(u.in_b); }
    a.w1 = params + u.w1; a.b1 = params + u.b1; a.w2 = params + u.w2; a.b2 = params + u.b2;
    a.B = B; a.H = H / u.div; a.W = W / u.div;
    const TensorDesc& to = p.tensors[u.out];
    if (to.pred_level >= 0) {
      a.zout = preds + (size_t)L.level_off[to.pred_level] * YUNET_PRED_CH;
      a.out_batch_stride = (long long)L.P * YUNET_PRED_CH;
    } else {
      a.zout = v.z(u.out);
      a.out_batch_stride = (long long)a.H * a.W * u.cout;
    }
    if (train && u.has_bn) {
      const BnDesc& bn = p.bns[to.bn];
      a.osum = v.stat(0) + bn.ch_off;
      a.osumsq = v.stat(1) + bn.ch_off;
    }
    {
      const double hw = (double)a.H * a.W;
      double bytes = 4.0 * B * (u.cin * hw * (u.mode == LOAD_POOL ? 4.0 : 1.0) + u.cout * hw);
      if (u.mode == LOAD_UPADD) bytes += 4.0 * B * u.cin * hw / 4.0;
      const bool ws = ctx->opt_ws_forward && unit_fwd_ws_supported(u.cin, u.cout, u.mode);
      const bool tc = !ws && ctx->opt_tc_forward && unit_fwd_tc_supported(u.cin, u.cout, u.mode);
      Scope sc(ctx, s, (ws ? "fwd_ws:" : tc ? "fwd_tc:" : "fwd:") + u.name, bytes);
      e = ws ? launch_unit_fwd_ws(u.cin, u.cout, u.mode, a, ctx->num_sms, v.status(), s)
          : tc ? launch_unit_fwd_tc(u.cout, u.mode, a, ctx->num_sms, v.status(), s)
               : launch_unit_fwd(u.cin, u.cout, u.mode, a, ctx->num_sms, s);
    }
    if (e != cudaSuccess) return fail(ctx, (int)e, "forward: unit %s: %s", u.name.c_str(), cudaGetErrorString(e));
  }
  if (train && bn_running) {
    BnFinalizeArgs fa = make_bn_args(p, B, H, W);
    Scope sc(ctx, s, "fwd:bn_running");
    e = launch_bn_update_running(fa, v.stat(0), v.stat(1), bn_running, bn_running + p.num_bn_ch,
                                 momentum, s);
    if (e != cudaSuccess) return cuda_fail(ctx, e, "forward: running statistics");
  }
  return 0;
}

int yunet_read_activation(yunet_ctx* ctx, int unit_index, const float* params,
                          const float* bn_running, int B, int H, int W, int train, const void* ws,
                          float* out_nchw, void* stream) {
  if (!ctx || !ws || !out_nchw || !params) return fail(ctx, -1, "read_activation: null pointer");
  const Plan& p = ctx->plan;
  if (unit_index < -1 || unit_index >= (int)p.units.size()) return fail(ctx, -1, "read_activation: bad unit index");
  if (!shape_ok(ctx, B, H, W)) return fail(ctx, -2, "read_activation: bad shape");
  const int t = unit_index < 0 ? p.stem_out : p.units[unit_index].out;
  const TensorDesc& td = p.tensors[t];
  if (td.pred_level >= 0) return fail(ctx, -1, "read_activation: prediction tensors live in `preds`");
  WsLayout L = make_layout(p, B, H, W, train != 0);
  Views v{p, L, (char*)const_cast<void*>(ws), params, bn_running, nullptr, nullptr, B, H, W, train};
  const int h = H / td.div, w = W / td.div;
  return cuda_fail(ctx, launch_read_activation(v.z(t), v.bnref(t), td.bn >= 0, B, h, w, td.C,
                                               (long long)h * w * td.C, out_nchw,
                                               (cudaStream_t)stream), "read_activation");
}

size_t yunet_assign_workspace_bytes(const yunet_ctx* ctx, int B, int H, int W) {
  if (!shape_ok(ctx, B, H, W)) return 0;
  return simota_workspace_bytes(B, make_geom(ctx->plan, H, W).P);
}

int yunet_simota_assign(yunet_ctx* ctx, const yunet_loss_cfg* lc, const float* preds,
                        const float* gt, const int* gt_offsets, int B, int H, int W,
                        int* assigned_gt, float* matched_iou, float* counters, void* ws,
                        size_t ws_bytes, void* stream) {
  if (!ctx || !lc || !preds || !gt || !gt_offsets || !assigned_gt || !matched_iou || !counters)
    return fail(ctx, -1, "simota_assign: null pointer");
  if (!shape_ok(ctx, B, H, W)) return fail(ctx, -2, "simota_assign: bad shape");
  if (lc->candidate_topk < 1 || lc->candidate_topk > 10) return fail(ctx, -2, "simota_assign: candidate_topk must be in 1..10");
  LevelGeom g = make_geom(ctx->plan, H, W);
  const size_t need = simota_workspace_bytes(B, g.P);
  if (need > 0 && (!ws || ws_bytes < need)) return fail(ctx, -3, "simota_assign: workspace too small");
  Scope sc(ctx, (cudaStream_t)stream, "simota_assign");
  return cuda_fail(ctx, launch_simota_assign(to_dev(lc), g, preds, gt, gt_offsets, B, assigned_gt,
                                             matched_iou, counters, ws, (cudaStream_t)stream),
                   "simota_assign");
}

int yunet_simota_assign_ext(yunet_ctx* ctx, const yunet_loss_cfg* lc, int P, const float* scores,
                            const float* priors, const float* decoded_boxes, const float* gt,
                            const int* gt_offsets, int* assigned_gt, float* matched_iou,
                            float* counters, void* ws, size_t ws_bytes, void* stream) {
  if (!ctx || !lc || !scores || !priors || !decoded_boxes || !gt || !gt_offsets || !assigned_gt ||
      !matched_iou || !counters)
    return fail(ctx, -1, "simota_assign_ext: null pointer");
  if (P <= 0) return fail(ctx, -2, "simota_assign_ext: bad prior count");
  if (lc->candidate_topk < 1 || lc->candidate_topk > 10) return fail(ctx, -2, "simota_assign_ext: candidate_topk must be in 1..10");
  const size_t need = simota_workspace_bytes(1, P);
  if (need > 0 && (!ws || ws_bytes < need)) return fail(ctx, -3, "simota_assign_ext: workspace too small");
  Scope sc(ctx, (cudaStream_t)stream, "simota_assign_ext");
  return cuda_fail(ctx, launch_simota_assign_ext(to_dev(lc), P, scores, priors, decoded_boxes, gt, gt_offsets,
                                                 assigned_gt, matched_iou, counters, ws, (cudaStream_t)stream),
                   "simota_assign_ext");
}

int yunet_loss_grad(yunet_ctx* ctx, const yunet_loss_cfg* lc, const float* preds, const float* gt,
                    const int* gt_offsets, const int* assigned_gt, const float* matched_iou,
                    const float* counters, const float* num_total_samples, const float* loss_scale,
                    int B, int H, int W, float* losses, float* d_preds, void* stream) {
  if (!ctx || !lc || !preds || !gt || !gt_offsets || !assigned_gt || !matched_iou || !counters ||
      !num_total_samples || !losses)
    return fail(ctx, -1, "loss_grad: null pointer");
  if (!shape_ok(ctx, B, H, W)) return fail(ctx, -2, "loss_grad: bad shape");
  LevelGeom g = make_geom(ctx->plan, H, W);
  float sc[4] = {1.f, 1.f, 1.f, 1.f};
  if (loss_scale) for (int i = 0; i < 4; ++i) sc[i] = loss_scale[i];
  Scope scope(ctx, (cudaStream_t)stream, "loss_grad");
  return cuda_fail(ctx, launch_loss_grad(to_dev(lc), g, preds, gt, gt_offsets, assigned_gt,
                                         matched_iou, counters, num_total_samples, sc[0], sc[1],
                                         sc[2], sc[3], B, losses, d_preds, (cudaStream_t)stream),
                   "loss_grad");
}

int yunet_backward(yunet_ctx* ctx, const float* img, const float* params, const float* d_preds,
                   int B, int H, int W, float* grad_bucket, void* ws, size_t ws_bytes,
                   void* stream) {
  if (!ctx) return -1;
  if (!img || !params || !d_preds || !grad_bucket || !ws) return fail(ctx, -1, "backward: null pointer");
  if (!shape_ok(ctx, B, H, W)) return fail(ctx, -2, "backward: bad shape");
  const Plan& p = ctx->plan;
  WsLayout L = make_layout(p, B, H, W, true);
  if (ws_bytes < L.total) return fail(ctx, -3, "backward: workspace too small (%zu < %zu)", ws_bytes, L.total);
  ensure_sms(ctx);
  cudaStream_t s = (cudaStream_t)stream;
  Views v{p, L, (char*)ws, params, nullptr, nullptr, d_preds, B, H, W, 1};
  cudaError_t e;
  e = cudaMemsetAsync(grad_bucket, 0, sizeof(float) * (size_t)p.num_params, s);
  if (e != cudaSuccess) return cuda_fail(ctx, e, "backward: memset grads");
  e = cudaMemsetAsync(v.stat(2), 0, sizeof(double) * 2 * (size_t)p.num_bn_ch, s);
  if (e != cudaSuccess) return cuda_fail(ctx, e, "backward: memset statistics");
  e = cudaMemsetAsync(v.status() + 1, 0, sizeof(int), s);
  if (e != cudaSuccess) return cuda_fail(ctx, e, "backward: memset status");
  for (int i = (int)p.units.size() - 1; i >= 0; --i) {
    const UnitDesc& u = p.units[i];
    UnitBwdArgs a;
    memset(&a, 0, sizeof a);
    a.za = v.z(u.in_a);
    a.bna = v.bnref(u.in_a);
    if (u.in_b >= 0) { a.zb = v.z(u.in_b); a.bnb = v.bnref(u.in_b); }
    a.w1 = params + u.w1; a.b1 = params + u.b1; a.w2 = params + u.w2;
    a.B = B; a.H = H / u.div; a.W = W / u.div;
    a.has_bn = u.has_bn ? 1 : 0;
    const TensorDesc& to = p.tensors[u.out];
    if (to.pred_level >= 0) {
      a.dout = d_preds + (size_t)L.level_off[to.pred_level] * YUNET_PRED_CH;
      a.dout_batch_stride = (long long)L.P * YUNET_PRED_CH;
    } else {
      a.dout = v.du(u.out);
      a.dout_batch_stride = (long long)a.H * a.W * u.cout;
      a.zout = v.z(u.out);
      a.bno = v.bnref(u.out);
      const BnDesc& bn = p.bns[to.bn];
      a.dsum = v.stat(2) + bn.ch_off;
      a.dsumzh = v.stat(3) + bn.ch_off;
    }
    a.dua = v.du(u.in_a);
    a.acc_a = u.acc_a ? 1 : 0;
    {
      const BnDesc& bn = p.bns[p.tensors[u.in_a].bn];
      a.dsum_a = v.stat(2) + bn.ch_off;
      a.dsumzh_a = v.stat(3) + bn.ch_off;
    }
    if (u.in_b >= 0) {
      a.dub = v.du(u.in_b);
      a.acc_b = u.acc_b ? 1 : 0;
      const BnDesc& bn = p.bns[p.tensors[u.in_b].bn];
      a.dsum_b = v.stat(2) + bn.ch_off;
      a.dsumzh_b = v.stat(3) + bn.ch_off;
    }
    a.gw1 = grad_bucket + u.w1; a.gb1 = grad_bucket + u.b1;
    a.gw2 = grad_bucket + u.w2; a.gb2 = grad_bucket + u.b2;
    a.partial = v.partial();
    if (u.b1 != u.w1 + (long long)u.cout * u.cin || u.w2 != u.b1 + u.cout || u.b2 != u.w2 + 9LL * u.cout)
      return fail(ctx, -4, "backward: unit %s: parameter group is not contiguous", u.name.c_str());
    {
      const double hw = (double)a.H * a.W;
      double bytes = 4.0 * B * (2.0 * u.cin * hw * (u.mode == LOAD_POOL ? 4.0 : 1.0) +
                                (u.has_bn ? 2.0 : 1.0) * u.cout * hw);
      if (u.mode == LOAD_UPADD) bytes += 4.0 * B * 2.0 * u.cin * hw / 4.0;
      // st_backward: 0 off, 1 where it is the faster kernel (default), 2 wherever it applies (tests)
      const bool st = ctx->opt_st_backward && unit_bwd_st_supported(u.cin, u.cout, u.mode, a.has_bn, a.H, a.W) &&
                      (ctx->opt_st_backward >= 2 || unit_bwd_st_preferred(u.mode, a.H, a.W));
      const bool tc = !st && ctx->opt_tc_backward && unit_bwd_tc_supported(u.cin, u.cout, u.mode, a.has_bn);
      Scope sc(ctx, s, (st ? "bwd_st:" : tc ? "bwd_tc:" : "bwd:") + u.name, bytes);
      e = st ? launch_unit_bwd_st(u.mode, a, ctx->num_sms, v.status() + 1, s)
          : tc ? launch_unit_bwd_tc(u.mode, a, ctx->num_sms, v.status() + 1, s)
               : launch_unit_bwd(u.cin, u.cout, u.mode, a, ctx->num_sms, s);
      ctx->launches++;      // every backward launcher is followed by its reduce_partials_kernel
    }
    if (e != cudaSuccess) return fail(ctx, (int)e, "backward: unit %s: %s", u.name.c_str(), cudaGetErrorString(e));
  }
  {
    StemBwdArgs a;
    memset(&a, 0, sizeof a);
    a.img = img; a.zout = v.z(p.stem_out); a.du = v.du(p.stem_out); a.bno = v.bnref(p.stem_out);
    const BnDesc& bn = p.bns[p.tensors[p.stem_out].bn];
    a.dsum = v.stat(2) + bn.ch_off; a.dsumzh = v.stat(3) + bn.ch_off;
    a.gw = grad_bucket + p.stem_w; a.gb = grad_bucket + p.stem_b;
    a.partial = v.partial();
    if (p.stem_b != p.stem_w + 432) return fail(ctx, -4, "backward: stem parameters are not contiguous");
    a.B = B; a.Hin = H; a.Win = W;
    Scope sc(ctx, s, "bwd:stem", 4.0 * B * (3.0 * H * W + 2.0 * 16.0 * (H / 2) * (W / 2)));
    e = launch_stem_bwd(a, ctx->num_sms, s);
    ctx->launches++;        // + reduce_partials_kernel
    if (e != cudaSuccess) return cuda_fail(ctx, e, "backward: stem");
  }
  BnFinalizeArgs fa = make_bn_args(p, B, H, W);
  {
    Scope sc(ctx, s, "bwd:bn_param_grads");
    e = launch_bn_param_grads(fa, v.stat(2), v.stat(3), grad_bucket, s);
  }
  return cuda_fail(ctx, e, "backward: bn parameter grads");
}

int yunet_sgd_step(yunet_ctx* ctx, float* params, const float* grad_bucket, float* momentum_buf,
                   long long n, float lr, float momentum, float weight_decay, float grad_scale,
                   void* stream) {
  if (!params || !grad_bucket || !momentum_buf || n <= 0) return fail(ctx, -1, "sgd_step: bad arguments");
  Scope sc(ctx, (cudaStream_t)stream, "sgd_step");
  return cuda_fail(ctx, launch_sgd(params, grad_bucket, momentum_buf, n, lr, momentum, weight_decay,
                                   grad_scale, (cudaStream_t)stream), "sgd_step");
}

int yunet_sgd_step_dev(yunet_ctx* ctx, float* params, const float* grad_bucket, float* momentum_buf,
                       long long n, const float* lr_dev, float momentum, float weight_decay,
                       float grad_scale, void* stream) {
  if (!params || !grad_bucket || !momentum_buf || !lr_dev || n <= 0) return fail(ctx, -1, "sgd_step_dev: bad arguments");
  Scope sc(ctx, (cudaStream_t)stream, "sgd_step");
  return cuda_fail(ctx, launch_sgd_dev(params, grad_bucket, momentum_buf, n, lr_dev, momentum, weight_decay,
                                       grad_scale, (cudaStream_t)stream), "sgd_step_dev");
}

size_t yunet_nms_workspace_bytes(const yunet_ctx* ctx, int B, int H, int W) {
  if (!shape_ok(ctx, B, H, W)) return 0;
  return nms_workspace_bytes(B, make_geom(ctx->plan, H, W).P);
}

int yunet_decode_nms(yunet_ctx* ctx, const float* preds, int B, int H, int W, float score_thr,
                     float iou_thr, const float* scale_factors, int max_det, float* dets,
                     float* det_kps, int* det_count, void* ws, size_t ws_bytes, void* stream) {
  if (!ctx || !preds || !dets || !det_count || !ws) return fail(ctx, -1, "decode_nms: null pointer");
  if (!shape_ok(ctx, B, H, W) || max_det <= 0) return fail(ctx, -2, "decode_nms: bad shape");
  LevelGeom g = make_geom(ctx->plan, H, W);
  if (ws_bytes < nms_workspace_bytes(B, g.P)) return fail(ctx, -3, "decode_nms: workspace too small");
  Scope sc(ctx, (cudaStream_t)stream, "decode_nms");
  return cuda_fail(ctx, launch_decode_nms(g, preds, B, score_thr, iou_thr, scale_factors, max_det,
                                          dets, det_kps, det_count, ws, (cudaStream_t)stream),
                   "decode_nms");
}

int yunet_preprocess_u8(yunet_ctx* ctx, const unsigned char* pixels, const long long* offsets,
                        const int* hw, const int* crop, int B, int S, float pad_value, float* out,
                        void* stream) {
  if (!ctx || !pixels || !offsets || !hw || !crop || !out) return fail(ctx, -1, "preprocess_u8: null pointer");
  if (B <= 0 || S <= 0 || S > 4096) return fail(ctx, -2, "preprocess_u8: bad shape");
  Scope sc(ctx, (cudaStream_t)stream, "preprocess_u8");
  return cuda_fail(ctx, launch_preprocess_u8(pixels, offsets, hw, crop, B, S, pad_value, out,
                                             (cudaStream_t)stream),
                   "preprocess_u8");
}

long long yunet_launch_count(const yunet_ctx* ctx) { return ctx ? ctx->launches : 0; }

int yunet_set_option(yunet_ctx* ctx, const char* name, int value) {
  if (!ctx || !name) return -1;
  if (strcmp(name, "tc_forward") == 0) { ctx->opt_tc_forward = value ? 1 : 0; return 0; }
  if (strcmp(name, "tc_backward") == 0) { ctx->opt_tc_backward = value ? 1 : 0; return 0; }
  if (strcmp(name, "ws_forward") == 0) { ctx->opt_ws_forward = value ? 1 : 0; return 0; }
  if (strcmp(name, "st_backward") == 0) { ctx->opt_st_backward = value < 0 ? 0 : (value > 2 ? 2 : value); return 0; }
  return fail(ctx, -1, "unknown option %s", name);
}

long long yunet_ws_offset(const yunet_ctx* ctx, int B, int H, int W, int train, int tensor_id,
                          int kind) {
  if (!ctx || !shape_ok(ctx, B, H, W)) return -1;
  const Plan& p = ctx->plan;
  WsLayout L = make_layout(p, B, H, W, train != 0);
  if (kind == 2) return (long long)L.stats_off;
  if (kind == 3) return (long long)L.status_off;
  if (tensor_id < 0 || tensor_id >= (int)p.tensors.size() || p.tensors[tensor_id].pred_level >= 0) return -1;
  if (kind == 0) return (long long)L.z_off[tensor_id];
  if (kind == 1 && train) return (long long)L.du_off[tensor_id];
  return -1;
}

int yunet_profile_begin(yunet_ctx* ctx) {
  if (!ctx) return -1;
  for (ProfEvent& e : ctx->prof) { cudaEventDestroy(e.start); cudaEventDestroy(e.stop); }
  ctx->prof.clear();
  ctx->prof_ms.clear();
  ctx->profiling = true;
  return 0;
}

int yunet_profile_end(yunet_ctx* ctx) {
  if (!ctx) return -1;
  ctx->profiling = false;
  ctx->prof_ms.assign(ctx->prof.size(), 0.f);
  for (size_t i = 0; i < ctx->prof.size(); ++i) {
    cudaError_t e = cudaEventSynchronize(ctx->prof[i].stop);
    if (e != cudaSuccess) return cuda_fail(ctx, e, "profile_end");
    cudaEventElapsedTime(&ctx->prof_ms[i], ctx->prof[i].start, ctx->prof[i].stop);
  }
  return (int)ctx->prof.size();
}

int yunet_profile_get(const yunet_ctx* ctx, int i, char* name, int name_cap, float* ms,
                      double* algorithmic_bytes) {
  if (!ctx || i < 0 || i >= (int)ctx->prof_ms.size()) return -1;
  copy_name(name, name_cap, ctx->prof[i].name);
  if (ms) *ms = ctx->prof_ms[i];
  if (algorithmic_bytes) *algorithmic_bytes = ctx->prof[i].bytes;
  return 0;
}

}  // extern "C"
