// Graph builder: arch cfg -> fused-unit execution plan + flat parameter bucket layout.
// See plan.h for the reference structures this mirrors.
#include "plan.h"

#include <cstdio>

namespace yunet {

namespace {

struct Builder {
  Plan& p;
  long long cursor = 0;

  long long take(const std::string& name, int ndim, int s0, int s1, int s2, int s3) {
    ParamInfo pi;
    pi.name = name;
    pi.offset = cursor;
    pi.ndim = ndim;
    pi.shape[0] = s0; pi.shape[1] = s1; pi.shape[2] = s2; pi.shape[3] = s3;
    long long n = 1;
    for (int i = 0; i < ndim; ++i) n *= pi.shape[i];
    cursor += n;
    p.params.push_back(pi);
    return pi.offset;
  }

  int add_tensor(int C, int div) {
    TensorDesc t;
    t.C = C;
    t.div = div;
    p.tensors.push_back(t);
    return (int)p.tensors.size() - 1;
  }

  int add_bn(const std::string& name, int C, long long gamma, long long beta, int tensor) {
    BnDesc b;
    b.name = name;
    b.C = C;
    b.ch_off = p.num_bn_ch;
    b.gamma = gamma;
    b.beta = beta;
    b.tensor = tensor;
    p.num_bn_ch += C;
    p.bns.push_back(b);
    p.tensors[tensor].bn = (int)p.bns.size() - 1;
    return (int)p.bns.size() - 1;
  }

  // A ConvDPUnit with BatchNorm+ReLU (mmdet/models/utils/yunet_layer.py:4-36).
  int add_dp_unit(const std::string& name, int cin, int cout, int mode, int in_a, int in_b,
                  int div) {
    UnitDesc u;
    u.name = name;
    u.cin = cin;
    u.cout = cout;
    u.mode = mode;
    u.in_a = in_a;
    u.in_b = in_b;
    u.div = div;
    u.has_bn = true;
    u.w1 = take(name + ".conv1.weight", 4, cout, cin, 1, 1);
    u.b1 = take(name + ".conv1.bias", 1, cout, 0, 0, 0);
    u.w2 = take(name + ".conv2.weight", 4, cout, 1, 3, 3);
    u.b2 = take(name + ".conv2.bias", 1, cout, 0, 0, 0);
    u.gamma = take(name + ".bn.weight", 1, cout, 0, 0, 0);
    u.beta = take(name + ".bn.bias", 1, cout, 0, 0, 0);
    u.out = add_tensor(cout, div);
    add_bn(name + ".bn", cout, u.gamma, u.beta, u.out);
    p.tensors[in_a].n_consumers++;
    if (in_b >= 0) p.tensors[in_b].n_consumers++;
    p.units.push_back(u);
    return u.out;
  }
};

}  // namespace

bool Plan::build(const yunet_arch_cfg& c) {
  cfg = c;
  tensors.clear(); units.clear(); bns.clear(); params.clear();
  num_params = 0; num_bn_ch = 0;
  char buf[256];
  if (c.num_stages < 2 || c.num_stages > YUNET_MAX_STAGES) { error = "num_stages out of range"; return false; }
  if (c.num_classes != 1 || c.kps_num != 5) {
    error = "only num_classes=1, kps_num=5 (16 prediction channels) is supported";
    return false;
  }
  auto ok_ch = [](int ch) { return ch == 16 || ch == 32 || ch == 64; };
  Builder b{*this};

  // ---- backbone stage 0: Conv_head (yunet_layer.py:39-62)
  stem_cin = c.stage_channels[0][0];
  stem_cout = c.stage_channels[0][1];
  if (stem_cin != 3 || stem_cout != 16) { error = "stem must be 3->16"; return false; }
  stem_w = b.take("backbone.model0.conv1.weight", 4, stem_cout, stem_cin, 3, 3);
  stem_b = b.take("backbone.model0.conv1.bias", 1, stem_cout, 0, 0, 0);
  int div = 2;
  stem_out = b.add_tensor(stem_cout, div);
  int cur;
  {
    // Conv_head registers conv1, conv2 (ConvDPUnit), bn1 in this order (yunet_layer.py:51-55)
    if (!ok_ch(c.stage_channels[0][2])) { error = "unsupported channel count"; return false; }
    // bn1 parameters come after conv2.* in the reference state_dict; the bucket order is free,
    // keep them adjacent to the stem conv for locality.
    long long g = b.take("backbone.model0.bn1.weight", 1, stem_cout, 0, 0, 0);
    long long be = b.take("backbone.model0.bn1.bias", 1, stem_cout, 0, 0, 0);
    b.add_bn("backbone.model0.bn1", stem_cout, g, be, stem_out);
    cur = b.add_dp_unit("backbone.model0.conv2", stem_cout, c.stage_channels[0][2], LOAD_PLAIN,
                        stem_out, -1, div);
  }
  int feat[3] = {-1, -1, -1};
  int feat_div[3] = {0, 0, 0};
  bool pool_next = (c.downsample_mask & 1) != 0;
  for (int k = 0; k < 3; ++k) if (c.out_idx[k] == 0) { feat[k] = cur; feat_div[k] = div; }
  int cur_c = c.stage_channels[0][2];
  for (int i = 1; i < c.num_stages; ++i) {
    int cin = c.stage_channels[i][0], cout = c.stage_channels[i][1];
    if (cin != cur_c || !ok_ch(cin) || !ok_ch(cout)) {
      snprintf(buf, sizeof buf, "stage %d: unsupported channels %d->%d (prev %d)", i, cin, cout, cur_c);
      error = buf;
      return false;
    }
    int mode = LOAD_PLAIN;
    if (pool_next) { mode = LOAD_POOL; div *= 2; }
    snprintf(buf, sizeof buf, "backbone.model%d", i);
    // Conv4layerBlock (yunet_layer.py:65-82)
    cur = b.add_dp_unit(std::string(buf) + ".conv1", cin, cin, mode, cur, -1, div);
    cur = b.add_dp_unit(std::string(buf) + ".conv2", cin, cout, LOAD_PLAIN, cur, -1, div);
    cur_c = cout;
    for (int k = 0; k < 3; ++k) if (c.out_idx[k] == i) { feat[k] = cur; feat_div[k] = div; }
    pool_next = ((c.downsample_mask >> i) & 1) != 0;
  }
  for (int k = 0; k < 3; ++k) {
    if (feat[k] < 0) { error = "out_idx does not name three backbone stages"; return false; }
    if (feat_div[k] != c.strides[k]) {
      snprintf(buf, sizeof buf, "level %d: feature stride %d != prior stride %d", k, feat_div[k], c.strides[k]);
      error = buf;
      return false;
    }
    if (k > 0 && feat_div[k] != 2 * feat_div[k - 1]) { error = "levels must differ by x2"; return false; }
    if (tensors[feat[k]].C != 64 || c.feat_channels != 64) { error = "neck/head channels must be 64"; return false; }
  }

  // ---- neck: TFPN (tfpn.py:33-45), executed top-down (2,1,0); bucket order follows execution.
  int lat[3];
  for (int i = 2; i >= 0; --i) {
    snprintf(buf, sizeof buf, "neck.lateral_convs.%d", i);
    if (i == 2) lat[i] = b.add_dp_unit(buf, 64, 64, LOAD_PLAIN, feat[2], -1, feat_div[2]);
    else lat[i] = b.add_dp_unit(buf, 64, 64, LOAD_UPADD, feat[i], lat[i + 1], feat_div[i]);
  }

  // ---- head (yunet_head.py:112-156,175-247), stacked_convs == 0
  for (int l = 0; l < 3; ++l) {
    int x = lat[l];
    for (int j = 0; j < c.shared_stacked_convs; ++j) {
      snprintf(buf, sizeof buf, "bbox_head.multi_level_share_convs.%d.%d", l, j);
      x = b.add_dp_unit(buf, 64, 64, LOAD_PLAIN, x, -1, feat_div[l]);
    }
    // the four branch ConvDPUnits (cls 1, bbox 4, obj 1, kps 10; no BN/ReLU) fused into one
    // 64->16 unit whose parameter block is the concatenation of the four reference tensors.
    UnitDesc u;
    snprintf(buf, sizeof buf, "bbox_head.level%d.branches", l);
    u.name = buf;
    u.cin = 64; u.cout = YUNET_PRED_CH; u.mode = LOAD_PLAIN; u.in_a = x; u.in_b = -1;
    u.div = feat_div[l]; u.has_bn = false;
    const char* br[4] = {"cls", "bbox", "obj", "kps"};
    const int bc[4] = {1, 4, 1, 10};
    for (int part = 0; part < 4; ++part) {
      for (int k = 0; k < 4; ++k) {
        snprintf(buf, sizeof buf, "bbox_head.multi_level_%s.%d", br[k], l);
        std::string n = buf;
        long long off = 0;
        if (part == 0) off = b.take(n + ".conv1.weight", 4, bc[k], 64, 1, 1);
        if (part == 1) off = b.take(n + ".conv1.bias", 1, bc[k], 0, 0, 0);
        if (part == 2) off = b.take(n + ".conv2.weight", 4, bc[k], 1, 3, 3);
        if (part == 3) off = b.take(n + ".conv2.bias", 1, bc[k], 0, 0, 0);
        if (k == 0) {
          if (part == 0) u.w1 = off;
          if (part == 1) u.b1 = off;
          if (part == 2) u.w2 = off;
          if (part == 3) u.b2 = off;
        }
      }
    }
    u.out = b.add_tensor(YUNET_PRED_CH, feat_div[l]);
    tensors[u.out].pred_level = l;
    tensors[x].n_consumers++;
    level_tensor[l] = u.out;
    units.push_back(u);
  }
  num_params = b.cursor;

  // every group base must be 16-byte aligned for the vectorised weight loads
  for (const UnitDesc& u : units) {
    if ((u.w1 | u.b1 | u.w2 | u.b2 | u.gamma | u.beta) & 3) { error = "internal: unaligned parameter group"; return false; }
  }

  // ---- backward bookkeeping: first writer (in reverse execution order) overwrites, later add
  std::vector<char> written(tensors.size(), 0);
  for (int i = (int)units.size() - 1; i >= 0; --i) {
    UnitDesc& u = units[i];
    u.acc_a = written[u.in_a] != 0;
    written[u.in_a] = 1;
    if (u.in_b >= 0) {
      u.acc_b = written[u.in_b] != 0;
      written[u.in_b] = 1;
    }
  }
  return true;
}

WsLayout make_layout(const Plan& p, int B, int H, int W, bool train) {
  WsLayout L;
  auto align = [](size_t x) { return (x + 255) & ~size_t(255); };
  size_t cur = 0;
  L.z_off.assign(p.tensors.size(), 0);
  L.du_off.assign(p.tensors.size(), 0);
  for (size_t i = 0; i < p.tensors.size(); ++i) {
    const TensorDesc& t = p.tensors[i];
    if (t.pred_level >= 0) continue;
    size_t bytes = (size_t)B * (H / t.div) * (W / t.div) * t.C * sizeof(float);
    L.z_off[i] = cur;
    cur = align(cur + bytes);
  }
  if (train) {
    for (size_t i = 0; i < p.tensors.size(); ++i) {
      const TensorDesc& t = p.tensors[i];
      if (t.pred_level >= 0) continue;
      size_t bytes = (size_t)B * (H / t.div) * (W / t.div) * t.C * sizeof(float);
      L.du_off[i] = cur;
      cur = align(cur + bytes);
    }
  }
  L.stats_off = cur;
  L.stats_bytes = sizeof(double) * 4 * (size_t)p.num_bn_ch;
  cur = align(cur + L.stats_bytes);
  L.status_off = cur;
  cur = align(cur + 256);
  L.partial_off = cur;
  if (train) cur = align(cur + sizeof(float) * 512 * 5120);   // kMaxPartialCtas x kPartialStride
  L.total = cur;
  int off = 0;
  for (int l = 0; l < 3; ++l) {
    L.level_h[l] = H / p.cfg.strides[l];
    L.level_w[l] = W / p.cfg.strides[l];
    L.level_off[l] = off;
    off += L.level_h[l] * L.level_w[l];
  }
  L.P = off;
  return L;
}

}  // namespace yunet
