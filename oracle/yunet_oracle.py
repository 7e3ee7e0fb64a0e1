"""TEST INFRASTRUCTURE — portable CPU oracle for the YuNet hot path (plain PyTorch fp32).

This is a *restatement* of the reference algorithm (ShiqiYu/libfacedetection.train @ 0047ac24),
written from its behaviour, with every function citing the reference ``file:line`` it follows
(paths relative to the reference root).  It exists because the reference is Python + mmcv and
cannot travel to the GPU box.  It is pinned (``tests/test_oracle_pinned.py``,
``oracle/gen_golden.py``) against

  * the unmodified reference run through ``oracle/ref_loader.py`` in the development container
    (forward maps, SimOTA assignment indices, four losses, every parameter gradient, SGD step,
    decode+NMS detections), and
  * the committed fixtures in ``tests/golden/`` generated from the reference by
    ``oracle/gen_golden.py`` (these travel to the GPU box).

Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs
of ``bench.py`` may import this module; the product path (``libfacedetection/train_b200``) never
does and fails loudly when its CUDA library is missing.

Floating-point work: parity bar is 1e-3 relative fp32 on values, exact on prior/assignment
indices (``BASELINE.json: north_star``).
"""
import math

import torch
import torch.nn.functional as F

# ----------------------------------------------------------------------------- architecture
ARCH = {
    # configs/yunet_n.py:104-145
    'yunet_n': dict(stage_channels=[[3, 16, 16], [16, 64], [64, 64], [64, 64], [64, 64], [64, 64]],
                    downsample_idx=[0, 2, 3, 4], out_idx=[3, 4, 5],
                    neck_channels=[64, 64, 64], neck_out_idx=[0, 1, 2],
                    in_channels=64, feat_channels=64, shared_stacked_convs=1, stacked_convs=0,
                    num_classes=1, kps_num=5, strides=[8, 16, 32]),
    # configs/yunet_s.py:104-145
    'yunet_s': dict(stage_channels=[[3, 16, 16], [16, 32], [32, 64], [64, 64], [64, 64], [64, 64]],
                    downsample_idx=[0, 2, 3, 4], out_idx=[3, 4, 5],
                    neck_channels=[64, 64, 64], neck_out_idx=[0, 1, 2],
                    in_channels=64, feat_channels=64, shared_stacked_convs=0, stacked_convs=0,
                    num_classes=1, kps_num=5, strides=[8, 16, 32]),
}

BN_EPS = 1e-5        # torch.nn.BatchNorm2d default, mmdet/models/utils/yunet_layer.py:27
BN_MOMENTUM = 0.1


def _dp_keys(prefix, bn=True):
    ks = [f'{prefix}.conv1.weight', f'{prefix}.conv1.bias', f'{prefix}.conv2.weight',
          f'{prefix}.conv2.bias']
    if bn:
        ks += [f'{prefix}.bn.weight', f'{prefix}.bn.bias']
    return ks


def dp_unit_prefixes(arch):
    """All ConvDPUnit prefixes in execution order, with (cin, cout, has_bn)."""
    a = ARCH[arch] if isinstance(arch, str) else arch
    units = []
    sc = a['stage_channels']
    units.append(('backbone.model0.conv2', sc[0][1], sc[0][2], True))
    for i in range(1, len(sc)):
        cin, cout = sc[i]
        units.append((f'backbone.model{i}.conv1', cin, cin, True))
        units.append((f'backbone.model{i}.conv2', cin, cout, True))
    for i, c in enumerate(a['neck_channels']):
        units.append((f'neck.lateral_convs.{i}', c, c, True))
    for lvl in range(len(a['strides'])):
        for j in range(a['shared_stacked_convs']):
            cin = a['in_channels'] if j == 0 else a['feat_channels']
            units.append((f'bbox_head.multi_level_share_convs.{lvl}.{j}', cin, a['feat_channels'],
                          True))
    chn = a['feat_channels'] if a['shared_stacked_convs'] > 0 else a['in_channels']
    for lvl in range(len(a['strides'])):
        units.append((f'bbox_head.multi_level_cls.{lvl}', chn, a['num_classes'], False))
        units.append((f'bbox_head.multi_level_bbox.{lvl}', chn, 4, False))
        units.append((f'bbox_head.multi_level_obj.{lvl}', chn, 1, False))
        units.append((f'bbox_head.multi_level_kps.{lvl}', chn, a['kps_num'] * 2, False))
    return units


def init_params(arch, seed=0):
    """Reference init: Xavier-normal conv weights, conv bias 0.02, BN gamma 1 / beta 0
    (mmdet/models/backbones/yunet_backbone.py:21-31, necks/tfpn.py:21-31,
    dense_heads/yunet_head.py:158-168).  Returns (params, buffers) keyed like the reference
    state_dict."""
    a = ARCH[arch]
    g = torch.Generator().manual_seed(seed)
    P, Bf = {}, {}

    def conv(key, cout, cin_per_group, k):
        w = torch.empty(cout, cin_per_group, k, k)
        fan_in = cin_per_group * k * k
        fan_out = cout * k * k
        std = math.sqrt(2.0 / float(fan_in + fan_out))
        w.normal_(0, std, generator=g)
        P[key + '.weight'] = w
        P[key + '.bias'] = torch.full((cout,), 0.02)

    def bn(key, c):
        P[key + '.weight'] = torch.ones(c)
        P[key + '.bias'] = torch.zeros(c)
        Bf[key + '.running_mean'] = torch.zeros(c)
        Bf[key + '.running_var'] = torch.ones(c)
        Bf[key + '.num_batches_tracked'] = torch.zeros((), dtype=torch.long)

    sc = a['stage_channels']
    conv('backbone.model0.conv1', sc[0][1], sc[0][0], 3)
    bn('backbone.model0.bn1', sc[0][1])
    for prefix, cin, cout, has_bn in dp_unit_prefixes(arch):
        conv(prefix + '.conv1', cout, cin, 1)
        conv(prefix + '.conv2', cout, 1, 3)
        if has_bn:
            bn(prefix + '.bn', cout)
    return P, Bf


def split_state_dict(sd):
    """Reference state_dict → (params, buffers), float32 / long, detached clones."""
    P, Bf = {}, {}
    for k, v in sd.items():
        if 'running_' in k or 'num_batches_tracked' in k:
            Bf[k] = v.detach().clone()
        else:
            P[k] = v.detach().clone()
    return P, Bf


# ----------------------------------------------------------------------------- model forward
def _bn(x, P, Bf, prefix, training):
    # torch.nn.BatchNorm2d semantics: batch statistics + running-stat update when training
    # (mmdet/models/utils/yunet_layer.py:27,34,54,59)
    rm, rv = Bf[prefix + '.running_mean'], Bf[prefix + '.running_var']
    if training:
        Bf[prefix + '.num_batches_tracked'] += 1
    return F.batch_norm(x, rm, rv, P[prefix + '.weight'], P[prefix + '.bias'], training,
                        BN_MOMENTUM, BN_EPS)


def conv_dp_unit(x, P, Bf, prefix, with_bn_relu, training):
    """ConvDPUnit.forward — mmdet/models/utils/yunet_layer.py:30-36:
    pointwise 1x1 (bias) -> depthwise 3x3 pad 1 (bias) -> [BN -> ReLU]."""
    w1 = P[prefix + '.conv1.weight']
    x = F.conv2d(x, w1, P[prefix + '.conv1.bias'])
    w2 = P[prefix + '.conv2.weight']
    x = F.conv2d(x, w2, P[prefix + '.conv2.bias'], padding=1, groups=w2.shape[0])
    if with_bn_relu:
        x = F.relu(_bn(x, P, Bf, prefix + '.bn', training))
    return x


def backbone_forward(img, P, Bf, arch, training):
    """YuNetBackbone.forward — mmdet/models/backbones/yunet_backbone.py:33-41, with Conv_head
    (yunet_layer.py:57-62) as stage 0 and Conv4layerBlock (yunet_layer.py:79-82) for the rest."""
    a = ARCH[arch]
    x = F.conv2d(img, P['backbone.model0.conv1.weight'], P['backbone.model0.conv1.bias'],
                 stride=2, padding=1)
    x = F.relu(_bn(x, P, Bf, 'backbone.model0.bn1', training))
    x = conv_dp_unit(x, P, Bf, 'backbone.model0.conv2', True, training)
    out = []
    n = len(a['stage_channels'])
    for i in range(n):
        if i > 0:
            x = conv_dp_unit(x, P, Bf, f'backbone.model{i}.conv1', True, training)
            x = conv_dp_unit(x, P, Bf, f'backbone.model{i}.conv2', True, training)
        if i in a['out_idx']:
            out.append(x)
        if i in a['downsample_idx']:
            x = F.max_pool2d(x, 2)
    return out


def neck_forward(feats, P, Bf, arch, training):
    """TFPN.forward — mmdet/models/necks/tfpn.py:33-45 (top-down, nearest x2, add)."""
    a = ARCH[arch]
    feats = list(feats)
    for i in range(len(feats) - 1, 0, -1):
        feats[i] = conv_dp_unit(feats[i], P, Bf, f'neck.lateral_convs.{i}', True, training)
        feats[i - 1] = feats[i - 1] + F.interpolate(feats[i], scale_factor=2., mode='nearest')
    feats[0] = conv_dp_unit(feats[0], P, Bf, 'neck.lateral_convs.0', True, training)
    return [feats[i] for i in a['neck_out_idx']]


def head_forward(feats, P, Bf, arch, training):
    """YuNet_Head.forward — mmdet/models/dense_heads/yunet_head.py:175-247 (stacked_convs == 0
    branch): optional shared ConvDPUnits then cls/bbox/obj/kps ConvDPUnits without BN/ReLU."""
    a = ARCH[arch]
    feats = list(feats)
    for lvl in range(len(feats)):
        for j in range(a['shared_stacked_convs']):
            feats[lvl] = conv_dp_unit(feats[lvl], P, Bf,
                                      f'bbox_head.multi_level_share_convs.{lvl}.{j}', True,
                                      training)
    outs = {k: [] for k in ('cls', 'bbox', 'obj', 'kps')}
    for lvl, f in enumerate(feats):
        for k in outs:
            outs[k].append(conv_dp_unit(f, P, Bf, f'bbox_head.multi_level_{k}.{lvl}', False,
                                        training))
    return outs['cls'], outs['bbox'], outs['obj'], outs['kps']


def model_forward(img, P, Bf, arch, training=False):
    """SingleStageDetector.extract_feat + bbox_head (detectors/single_stage.py:52-57,
    detectors/yunet.py:83-86) → (cls_preds, bbox_preds, obj_preds, kps_preds), NCHW per level."""
    feats = backbone_forward(img, P, Bf, arch, training)
    feats = neck_forward(feats, P, Bf, arch, training)
    return head_forward(feats, P, Bf, arch, training)


# ----------------------------------------------------------------------------- priors / decode
def grid_priors(featmap_sizes, strides, dtype=torch.float32):
    """MlvlPointGenerator.grid_priors(with_stride=True), offset 0 —
    mmdet/core/anchor/point_generator.py:80-175: rows (x=j*s, y=i*s, s, s), x fastest."""
    out = []
    for (h, w), s in zip(featmap_sizes, strides):
        xs = (torch.arange(0, w) * s).to(dtype)
        ys = (torch.arange(0, h) * s).to(dtype)
        yy, xx = torch.meshgrid(ys, xs, indexing='ij')
        st = torch.full((h * w,), float(s), dtype=dtype)
        out.append(torch.stack([xx.reshape(-1), yy.reshape(-1), st, st], -1))
    return out


def flatten_preds(cls_preds, bbox_preds, obj_preds, kps_preds):
    """yunet_head.py:456-477: permute(0,2,3,1).reshape(B,-1,C) per level, cat over levels."""
    B = cls_preds[0].shape[0]

    def fl(lst, c):
        return torch.cat([t.permute(0, 2, 3, 1).reshape(B, -1, c) for t in lst], 1)

    return (fl(cls_preds, cls_preds[0].shape[1]), fl(bbox_preds, 4),
            fl(obj_preds, 1).squeeze(-1), fl(kps_preds, kps_preds[0].shape[1]))


def bbox_decode(priors, bbox_preds):
    """YuNet_Head._bbox_decode — yunet_head.py:376-386."""
    xys = (bbox_preds[..., :2] * priors[..., 2:]) + priors[..., :2]
    whs = bbox_preds[..., 2:].exp() * priors[..., 2:]
    return torch.stack([xys[..., 0] - whs[..., 0] / 2, xys[..., 1] - whs[..., 1] / 2,
                        xys[..., 0] + whs[..., 0] / 2, xys[..., 1] + whs[..., 1] / 2], -1)


def kps_encode(priors, kps):
    """YuNet_Head._kps_encode — yunet_head.py:395-402."""
    n = kps.shape[-1] // 2
    return torch.cat([(kps[..., [2 * i, 2 * i + 1]] - priors[..., :2]) / priors[..., 2:]
                      for i in range(n)], -1)


def kps_decode(priors, kps_preds):
    """YuNet_Head._kps_decode — yunet_head.py:388-393."""
    n = kps_preds.shape[-1] // 2
    return torch.cat([(kps_preds[..., [2 * i, 2 * i + 1]] * priors[..., 2:]) + priors[..., :2]
                      for i in range(n)], -1)


# ----------------------------------------------------------------------------- SimOTA
def bbox_overlaps(b1, b2, eps=1e-6):
    """bbox_overlaps(mode='iou', is_aligned=False) —
    mmdet/core/bbox/iou_calculators/iou2d_calculator.py:213-253."""
    area1 = (b1[:, 2] - b1[:, 0]) * (b1[:, 3] - b1[:, 1])
    area2 = (b2[:, 2] - b2[:, 0]) * (b2[:, 3] - b2[:, 1])
    lt = torch.max(b1[:, None, :2], b2[None, :, :2])
    rb = torch.min(b1[:, None, 2:], b2[None, :, 2:])
    wh = (rb - lt).clamp(min=0)
    overlap = wh[..., 0] * wh[..., 1]
    union = area1[:, None] + area2[None, :] - overlap
    union = torch.max(union, union.new_tensor([eps]))
    return overlap / union


def in_gt_and_in_center(priors, gt_bboxes, center_radius=2.5):
    """SimOTAAssigner.get_in_gt_and_in_center_info — sim_ota_assigner.py:186-228."""
    x, y, sx, sy = (priors[:, i:i + 1] for i in range(4))
    l_, t_ = x - gt_bboxes[:, 0], y - gt_bboxes[:, 1]
    r_, b_ = gt_bboxes[:, 2] - x, gt_bboxes[:, 3] - y
    is_in_gts = torch.stack([l_, t_, r_, b_], 1).min(1).values > 0
    cx = (gt_bboxes[:, 0] + gt_bboxes[:, 2]) / 2.0
    cy = (gt_bboxes[:, 1] + gt_bboxes[:, 3]) / 2.0
    cl_ = x - (cx - center_radius * sx)
    ct_ = y - (cy - center_radius * sy)
    cr_ = (cx + center_radius * sx) - x
    cb_ = (cy + center_radius * sy) - y
    is_in_cts = torch.stack([cl_, ct_, cr_, cb_], 1).min(1).values > 0
    valid = (is_in_gts.sum(1) > 0) | (is_in_cts.sum(1) > 0)
    return valid, is_in_gts[valid] & is_in_cts[valid]


def simota_assign(pred_scores, priors, decoded_bboxes, gt_bboxes, gt_labels,
                  center_radius=2.5, candidate_topk=10, iou_weight=3.0, cls_weight=1.0,
                  eps=1e-7, return_debug=False):
    """SimOTAAssigner._assign + dynamic_k_matching — sim_ota_assigner.py:95-257.

    Returns (assigned_gt_inds (P,) long, 1-based / 0, max_overlaps (P,) float).  Uses the same
    torch ops as the reference (``topk``, ``min``) so tie behaviour is that of the installed
    torch build."""
    INF = 100000.0
    num_gt = gt_bboxes.size(0)
    num_bboxes = decoded_bboxes.size(0)
    assigned = decoded_bboxes.new_full((num_bboxes,), 0, dtype=torch.long)
    if num_gt == 0 or num_bboxes == 0:
        return assigned, decoded_bboxes.new_zeros((num_bboxes,))
    valid_mask, in_both = in_gt_and_in_center(priors, gt_bboxes, center_radius)
    vb = decoded_bboxes[valid_mask]
    vs = pred_scores[valid_mask]
    num_valid = vb.size(0)
    if num_valid == 0:
        return assigned, decoded_bboxes.new_zeros((num_bboxes,))
    ious = bbox_overlaps(vb, gt_bboxes)
    iou_cost = -torch.log(ious + eps)
    onehot = F.one_hot(gt_labels.to(torch.int64), pred_scores.shape[-1]).float() \
        .unsqueeze(0).repeat(num_valid, 1, 1)
    vs = vs.unsqueeze(1).repeat(1, num_gt, 1)
    cls_cost = F.binary_cross_entropy(vs.to(torch.float32).sqrt_(), onehot,
                                      reduction='none').sum(-1)
    cost = cls_cost * cls_weight + iou_cost * iou_weight + (~in_both) * INF
    # dynamic_k_matching (sim_ota_assigner.py:230-257)
    matching = torch.zeros_like(cost, dtype=torch.uint8)
    topk_ious, _ = torch.topk(ious, min(candidate_topk, num_valid), dim=0)
    dynamic_ks = torch.clamp(topk_ious.sum(0).int(), min=1)
    for g in range(num_gt):
        _, pos_idx = torch.topk(cost[:, g], k=int(dynamic_ks[g]), largest=False)
        matching[:, g][pos_idx] = 1
    multi = matching.sum(1) > 1
    if multi.sum() > 0:
        _, cost_argmin = torch.min(cost[multi, :], dim=1)
        matching[multi, :] *= 0
        matching[multi, cost_argmin] = 1
    fg = matching.sum(1) > 0
    valid_mask = valid_mask.clone()
    valid_mask[valid_mask.clone()] = fg
    matched_gt = matching[fg, :].argmax(1)
    matched_iou = (matching * ious).sum(1)[fg]
    assigned[valid_mask] = matched_gt + 1
    max_overlaps = assigned.new_full((num_bboxes,), -INF, dtype=torch.float32)
    max_overlaps[valid_mask] = matched_iou
    if return_debug:
        return assigned, max_overlaps, dict(cost=cost, ious=ious, dynamic_ks=dynamic_ks,
                                            valid=in_gt_and_in_center(priors, gt_bboxes,
                                                                      center_radius)[0])
    return assigned, max_overlaps


def get_target_single(cls_preds, objectness, priors, decoded_bboxes, gt_bboxes, gt_labels,
                      gt_kpss, num_classes=1, nk=5, **assigner_kw):
    """YuNet_Head._get_target_single — yunet_head.py:536-604 (requires num_gts >= 1, as the
    reference does in practice)."""
    gt_bboxes = gt_bboxes.to(decoded_bboxes.dtype)
    gt_kpss = gt_kpss.to(decoded_bboxes.dtype)
    offset_priors = torch.cat([priors[:, :2] + priors[:, 2:] * 0.5, priors[:, 2:]], -1)
    assigned, max_ov = simota_assign(cls_preds.sigmoid() * objectness.unsqueeze(1).sigmoid(),
                                     offset_priors, decoded_bboxes, gt_bboxes, gt_labels,
                                     **assigner_kw)
    # PseudoSampler.sample — pseudo_sampler.py:35-41, sampling_result.py:23-47
    pos_inds = torch.nonzero(assigned > 0, as_tuple=False).squeeze(-1).unique()
    pos_gt = assigned[pos_inds] - 1
    pos_ious = max_ov[pos_inds]
    cls_target = F.one_hot(gt_labels[pos_gt].long(), num_classes) * pos_ious.unsqueeze(-1)
    obj_target = torch.zeros_like(objectness).unsqueeze(-1)
    obj_target[pos_inds] = 1
    bbox_target = gt_bboxes[pos_gt]
    kps_target = gt_kpss[pos_gt, :, :2].reshape((-1, nk * 2))
    kps_weight = torch.mean(gt_kpss[pos_gt, :, 2], dim=1, keepdim=True)
    fg = torch.zeros_like(objectness).to(torch.bool)
    fg[pos_inds] = 1
    return fg, cls_target, obj_target, bbox_target, kps_target, kps_weight, pos_inds.size(0), \
        assigned, max_ov


# ----------------------------------------------------------------------------- losses
def eiou_loss_elem(pred, target, smooth_point=0.1, eps=1e-6):
    """eiou_loss — mmdet/models/losses/iou_loss.py:194-227 (EIoULoss default eps 1e-6, :536)."""
    px1, py1, px2, py2 = pred[:, 0], pred[:, 1], pred[:, 2], pred[:, 3]
    tx1, ty1, tx2, ty2 = target[:, 0], target[:, 1], target[:, 2], target[:, 3]
    ex1, ey1 = torch.min(px1, tx1), torch.min(py1, ty1)
    ix1, iy1 = torch.max(px1, tx1), torch.max(py1, ty1)
    ix2, iy2 = torch.min(px2, tx2), torch.min(py2, ty2)
    xmin, ymin = torch.min(ix1, ix2), torch.min(iy1, iy2)
    xmax, ymax = torch.max(ix1, ix2), torch.max(iy1, iy2)
    inter = (ix2 - ex1) * (iy2 - ey1) + (xmin - ex1) * (ymin - ey1) - (ix1 - ex1) * (
        ymax - ey1) - (xmax - ex1) * (iy1 - ey1)
    union = (px2 - px1) * (py2 - py1) + (tx2 - tx1) * (ty2 - ty1) - inter + eps
    ious = 1 - (inter / union)
    sm = (ious < smooth_point).detach().float()
    return 0.5 * sm * (ious ** 2) / smooth_point + (1 - sm) * (ious - 0.5 * smooth_point)


def smooth_l1_elem(pred, target, beta):
    """smooth_l1_loss — mmdet/models/losses/smooth_l1_loss.py:24-32."""
    diff = torch.abs(pred - target)
    return torch.where(diff < beta, 0.5 * diff * diff / beta, diff - 0.5 * beta)


def head_loss(cls_preds, bbox_preds, obj_preds, kps_preds, gt_bboxes, gt_labels, gt_kpss,
              strides=(8, 16, 32), world_mean_num_pos=None, return_assign=False):
    """YuNet_Head.loss — yunet_head.py:418-534, with the config's loss settings
    (configs/yunet_n.py:122-136): sigmoid-BCE cls (pos only, IoU-soft targets, sum),
    EIoU bbox x5 (sum), sigmoid-BCE obj (all priors, sum), SmoothL1(beta=1/9) kps x0.1 with
    avg_factor = sum(kps_weight).  cls/bbox/obj are divided by
    max(reduce_mean(num_pos), 1) (yunet_head.py:493-497, dist_utils.py:68-74)."""
    B = cls_preds[0].shape[0]
    nk = kps_preds[0].shape[1] // 2
    sizes = [t.shape[2:] for t in cls_preds]
    priors = torch.cat(grid_priors(sizes, strides, cls_preds[0].dtype))
    f_cls, f_bbox, f_obj, f_kps = flatten_preds(cls_preds, bbox_preds, obj_preds, kps_preds)
    f_priors = priors.unsqueeze(0).repeat(B, 1, 1)
    f_boxes = bbox_decode(f_priors, f_bbox)
    res = []
    with torch.no_grad():
        for b in range(B):  # multi_apply(_get_target_single, ...) — yunet_head.py:483-489
            res.append(get_target_single(f_cls[b].detach(), f_obj[b].detach(), f_priors[b],
                                         f_boxes[b].detach(), gt_bboxes[b], gt_labels[b],
                                         gt_kpss[b], num_classes=f_cls.shape[-1], nk=nk))
    num_pos = torch.tensor(sum(r[6] for r in res), dtype=torch.float)
    if world_mean_num_pos is not None:
        num_pos = torch.tensor(float(world_mean_num_pos))
    num_total = max(num_pos, torch.tensor(1.0))
    pos_masks = torch.cat([r[0] for r in res], 0)
    cls_t = torch.cat([r[1] for r in res], 0)
    obj_t = torch.cat([r[2] for r in res], 0)
    bbox_t = torch.cat([r[3] for r in res], 0)
    kps_t = torch.cat([r[4] for r in res], 0)
    kps_w = torch.cat([r[5] for r in res], 0)

    loss_bbox = 5.0 * eiou_loss_elem(f_boxes.view(-1, 4)[pos_masks], bbox_t).sum() / num_total
    loss_obj = F.binary_cross_entropy_with_logits(f_obj.reshape(-1, 1), obj_t,
                                                  reduction='none').sum() / num_total
    loss_cls = F.binary_cross_entropy_with_logits(f_cls.view(-1, f_cls.shape[-1])[pos_masks],
                                                  cls_t.float(),
                                                  reduction='none').sum() / num_total
    enc = kps_encode(f_priors.view(-1, 4)[pos_masks], kps_t)
    l = smooth_l1_elem(f_kps.view(-1, nk * 2)[pos_masks], enc, 0.1111111111111111)
    l = l * kps_w.view(-1, 1)
    loss_kps = 0.1 * (l.sum() / (torch.sum(kps_w) + torch.finfo(torch.float32).eps))
    losses = dict(loss_cls=loss_cls, loss_bbox=loss_bbox, loss_obj=loss_obj, loss_kps=loss_kps)
    if return_assign:
        assigned = torch.stack([r[7] for r in res])
        max_ov = torch.stack([r[8] for r in res])
        return losses, dict(assigned_gt_inds=assigned, max_overlaps=max_ov,
                            num_pos=float(sum(r[6] for r in res)), kps_weight_sum=float(kps_w.sum()))
    return losses


def train_forward_backward(img, P, Bf, arch, gt_bboxes, gt_labels, gt_kpss):
    """BaseDetector.train_step core (detectors/base.py:219-252, detectors/yunet.py:21-51):
    forward_train -> _parse_losses (sum of four losses) -> backward.  Returns
    (losses dict of floats, grads dict keyed like P, assign dict)."""
    Pg = {k: v.detach().clone().requires_grad_(True) for k, v in P.items()}
    outs = model_forward(img, Pg, Bf, arch, training=True)
    losses, assign = head_loss(*outs, gt_bboxes, gt_labels, gt_kpss,
                               strides=ARCH[arch]['strides'], return_assign=True)
    total = sum(losses.values())
    total.backward()
    grads = {k: (v.grad if v.grad is not None else torch.zeros_like(v)) for k, v in Pg.items()}
    return {k: float(v) for k, v in losses.items()}, grads, assign, outs


def sgd_step(P, grads, momentum_buf, lr=0.01, momentum=0.9, weight_decay=0.0005):
    """torch.optim.SGD step as configured in configs/yunet_n.py:1 (all params decayed)."""
    for k in P:
        g = grads[k] + weight_decay * P[k]
        if k not in momentum_buf:
            momentum_buf[k] = g.clone()
        else:
            momentum_buf[k].mul_(momentum).add_(g)
        P[k] = P[k] - lr * momentum_buf[k]


# ----------------------------------------------------------------------------- decode + NMS
def nms_greedy(boxes, scores, iou_thr):
    """Greedy IoU-NMS as mmcv.ops.nms / torchvision.ops.nms do it (mmcv-full 1.3.17..1.6.0,
    call site yunet_head.py:415): sort by score descending (stable: lower index first on ties),
    suppress IoU > thr, IoU = inter / (a + b - inter) without +1 offset.  Returns keep indices
    in score order.  Pure-loop restatement; pinned against torchvision in the CPU tests."""
    n = boxes.shape[0]
    if n == 0:
        return torch.zeros((0,), dtype=torch.long)
    order = torch.sort(scores, descending=True, stable=True).indices
    b = boxes[order].numpy()
    x1, y1, x2, y2 = b[:, 0], b[:, 1], b[:, 2], b[:, 3]
    areas = (x2 - x1) * (y2 - y1)
    suppressed = [False] * n
    keep = []
    import numpy as np
    thr = np.float32(iou_thr)
    for i in range(n):
        if suppressed[i]:
            continue
        keep.append(i)
        xx1 = np.maximum(x1[i], x1[i + 1:])
        yy1 = np.maximum(y1[i], y1[i + 1:])
        xx2 = np.minimum(x2[i], x2[i + 1:])
        yy2 = np.minimum(y2[i], y2[i + 1:])
        w = np.maximum(np.float32(0), xx2 - xx1)
        h = np.maximum(np.float32(0), yy2 - yy1)
        inter = w * h
        ovr = inter / (areas[i] + areas[i + 1:] - inter)
        for j in np.nonzero(ovr > thr)[0]:
            suppressed[i + 1 + int(j)] = True
    return order[torch.tensor(keep, dtype=torch.long)]


def get_bboxes(cls_preds, bbox_preds, obj_preds, kps_preds, strides=(8, 16, 32), score_thr=0.02,
               iou_thr=0.45, scale_factors=None, with_kps=False, use_torchvision=False):
    """YuNet_Head.get_bboxes + _bboxes_nms — yunet_head.py:290-374,404-416 with
    test_cfg of configs/yunet_n.py:139-145 (no top-k caps).  Returns per image (dets (n,5), labels)
    [+ decoded keypoints (n,10) when ``with_kps``, the export-path extra of
    tools/compare_inference.py:381-386]."""
    B = cls_preds[0].shape[0]
    sizes = [t.shape[2:] for t in cls_preds]
    priors = torch.cat(grid_priors(sizes, strides, cls_preds[0].dtype))
    f_cls, f_bbox, f_obj, f_kps = flatten_preds(cls_preds, bbox_preds, obj_preds, kps_preds)
    f_cls = f_cls.sigmoid()
    f_obj = f_obj.sigmoid()
    boxes = bbox_decode(priors, f_bbox)
    kps = kps_decode(priors, f_kps)
    if scale_factors is not None:
        boxes = boxes / torch.as_tensor(scale_factors, dtype=boxes.dtype).unsqueeze(1)
    out = []
    for b in range(B):
        max_scores, labels = torch.max(f_cls[b], 1)
        valid = f_obj[b] * max_scores >= score_thr
        bb = boxes[b][valid]
        sc = max_scores[valid] * f_obj[b][valid]
        lb = labels[valid]
        kk = kps[b][valid]
        if lb.numel() == 0:  # yunet_head.py:412-413 returns the empty (0,4) boxes as-is
            out.append((bb, lb, kk) if with_kps else (bb, lb))
            continue
        if use_torchvision:
            import torchvision
            keep = torchvision.ops.nms(bb, sc, iou_thr)
        else:
            keep = nms_greedy(bb, sc, iou_thr)
        dets = torch.cat([bb[keep], sc[keep][:, None]], -1)
        out.append((dets, lb[keep], kk[keep]) if with_kps else (dets, lb[keep]))
    return out
