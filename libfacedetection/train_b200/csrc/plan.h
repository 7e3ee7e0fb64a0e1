// Execution plan of the YuNet hot path: the graph of fused ConvDPUnit kernels, the flat parameter
// bucket layout and the workspace layout.  Built once per ctx from yunet_arch_cfg.
//
// Reference structure being planned (not code): mmdet/models/backbones/yunet_backbone.py:11-41,
// mmdet/models/necks/tfpn.py:11-45, mmdet/models/dense_heads/yunet_head.py:112-156,175-247.
#pragma once
#include <cstddef>
#include <cstdint>
#include <string>
#include <vector>

#include "../../../include/yunet_b200.h"

namespace yunet {

enum LoadMode : int { LOAD_PLAIN = 0, LOAD_POOL = 1, LOAD_UPADD = 2 };

// One NHWC activation tensor held in the workspace as its *pre-BatchNorm* value z; the consumer
// applies BN+ReLU (and max-pool / upsample-add) while loading it.
struct TensorDesc {
  int C = 0;
  int div = 1;       // spatial size = (H/div, W/div)
  int bn = -1;       // BatchNorm index normalising this tensor (-1: none, e.g. head outputs)
  int pred_level = -1;  // >=0: lives in the (B,P,16) prediction tensor, not in the workspace
  int n_consumers = 0;
};

struct UnitDesc {
  std::string name;  // reference module path, e.g. "backbone.model2.conv1"
  int cin = 0, cout = 0;
  int mode = LOAD_PLAIN;
  int in_a = -1, in_b = -1;  // tensor ids (in_b only for LOAD_UPADD: a + nearest_up2(b))
  int out = -1;
  int div = 1;               // output resolution divisor
  bool has_bn = true;
  // offsets (floats) into the flat parameter bucket
  long long w1 = 0, b1 = 0, w2 = 0, b2 = 0, gamma = 0, beta = 0;
  // backward bookkeeping: does this unit overwrite (first writer in backward order) or
  // accumulate into the gradient of its inputs?
  bool acc_a = false, acc_b = false;
};

struct BnDesc {
  std::string name;   // state_dict prefix, e.g. "backbone.model0.bn1"
  int C = 0;
  long long ch_off = 0;     // channel offset into the running-stat / statistics arrays
  long long gamma = 0, beta = 0;  // parameter bucket offsets
  int tensor = -1;
};

struct ParamInfo {
  std::string name;
  long long offset;
  int ndim;
  int shape[4];
};

struct Plan {
  yunet_arch_cfg cfg;
  std::vector<TensorDesc> tensors;
  std::vector<UnitDesc> units;  // execution order (forward)
  std::vector<BnDesc> bns;
  std::vector<ParamInfo> params;  // reference state_dict order
  long long num_params = 0;
  long long num_bn_ch = 0;
  // stem (Conv_head.conv1 + bn1): 3x3 stride 2 dense conv
  int stem_cin = 3, stem_cout = 16;
  long long stem_w = 0, stem_b = 0;
  int stem_out = 0;  // tensor id
  int level_tensor[3] = {-1, -1, -1};
  std::string error;

  bool build(const yunet_arch_cfg& c);
};

// Workspace layout for a given (B,H,W).  All offsets in bytes from the workspace base.
struct WsLayout {
  std::vector<size_t> z_off;    // per tensor (pred tensors: unused)
  std::vector<size_t> du_off;   // per tensor, train only
  size_t stats_off = 0;         // double [4][num_bn_ch]: sum, sumsq, dsum, dsum_zh
  size_t stats_bytes = 0;
  size_t status_off = 0;        // int[64]: device-side error flags (bounded waits of the TC kernels)
  size_t partial_off = 0;       // per-CTA parameter-gradient partials (train only)
  size_t total = 0;
  int P = 0;
  int level_off[3] = {0, 0, 0};  // prior offset of each level
  int level_h[3] = {0, 0, 0}, level_w[3] = {0, 0, 0};
};

WsLayout make_layout(const Plan& p, int B, int H, int W, bool train);

}  // namespace yunet
