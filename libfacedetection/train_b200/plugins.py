"""Host-side mirror of the reference's mmdet plugin surface for the YuNet path.

Classes are registered under the reference's own names with identical constructor kwargs,
``forward`` signatures and ``state_dict`` keys, so ``weights/yunet_{n,s}.pth`` load strict and a
config written for the reference builds them unchanged:

  ``YuNetBackbone``   mmdet/models/backbones/yunet_backbone.py:8-41
  ``TFPN``            mmdet/models/necks/tfpn.py:8-45
  ``YuNet_Head``      mmdet/models/dense_heads/yunet_head.py:16-604
  ``SimOTAAssigner``  mmdet/core/bbox/assigners/sim_ota_assigner.py:12-36
  ``YuNet``           mmdet/models/detectors/yunet.py:7-86 (+ single_stage.py:17-57, base.py:184-252)
  ``SGD``             ``optimizer = dict(type='SGD', ...)`` (configs/yunet_n.py:1) as mmcv's
                      ``build_optimizer`` resolves it: ``torch.optim.SGD`` with a one-launch fused step

Execution is fused across the three modules: ``YuNetBackbone.forward`` returns a light
``FusedFeatures`` handle that ``TFPN.forward`` passes through and ``YuNet_Head`` consumes — the
reference's own ``SingleStageDetector.extract_feat`` -> ``bbox_head.forward_train`` call sequence
(single_stage.py:52-57, yunet.py:46-51) drives the hand-written kernels without modification.
The torch modules only *hold* the parameters (as views into the engine's flat bucket); no torch
operator runs on the hot path.
"""
import weakref

import numpy as np
import torch
import torch.nn as nn

from . import _capi
from .engine import YuNetEngine
from . import synthetic


# --------------------------------------------------------------------------------- registries
class Registry:
    """Minimal stand-in for ``mmcv.utils.Registry`` (same decorator / build protocol)."""

    def __init__(self, name):
        self.name = name
        self.module_dict = {}

    def register_module(self, name=None, force=False, module=None):
        def deco(cls):
            self.module_dict[name or cls.__name__] = cls
            return cls
        if module is not None:
            return deco(module)
        return deco

    def get(self, key):
        return self.module_dict.get(key)

    def build(self, cfg, default_args=None):
        args = dict(cfg)
        if default_args:
            for k, v in default_args.items():
                args.setdefault(k, v)
        t = args.pop('type')
        cls = self.get(t) if isinstance(t, str) else t
        if cls is None:
            raise KeyError(f'{t} is not in the {self.name} registry')
        return cls(**args)


BACKBONES = Registry('backbone')
NECKS = Registry('neck')
HEADS = Registry('head')
DETECTORS = Registry('detector')
BBOX_ASSIGNERS = Registry('bbox_assigner')
OPTIMIZERS = Registry('optimizer')


def register_into_mmdet(force=True):
    """Swap the reference's implementations for these in mmdet's own registries (call after
    ``import mmdet`` in an environment that has mmcv)."""
    from mmdet.models.builder import MODELS
    from mmdet.core.bbox.builder import BBOX_ASSIGNERS as MM_ASSIGNERS
    for cls in (YuNetBackbone, TFPN, YuNet_Head, YuNet):
        MODELS.register_module(name=cls.__name__, force=force, module=cls)
    MM_ASSIGNERS.register_module(name='SimOTAAssigner', force=force, module=SimOTAAssigner)
    try:    # mmcv.runner.optimizer.builder registers the torch optimizers under their class names
        from mmcv.runner.optimizer.builder import OPTIMIZERS as MM_OPTIMIZERS
        MM_OPTIMIZERS.register_module(name='SGD', force=force, module=SGD)
    except (ImportError, AttributeError, TypeError):     # an mmcv without that module / a partial stub
        pass


# --------------------------------------------------------------------------------- containers
def _init_like_reference(module):
    # yunet_backbone.py:21-31 / tfpn.py:21-31 / yunet_head.py:158-168
    for m in module.modules():
        if isinstance(m, nn.Conv2d):
            if m.bias is not None:
                nn.init.xavier_normal_(m.weight.data)
                m.bias.data.fill_(0.02)
            else:
                m.weight.data.normal_(0, 0.01)
        elif isinstance(m, nn.BatchNorm2d):
            m.weight.data.fill_(1)
            m.bias.data.zero_()


class ConvDPUnit(nn.Module):
    """Parameter container with the reference layout (yunet_layer.py:4-36)."""

    def __init__(self, in_channels, out_channels, withBNRelu=True):
        super().__init__()
        self.in_channels, self.out_channels, self.withBNRelu = in_channels, out_channels, withBNRelu
        self.conv1 = nn.Conv2d(in_channels, out_channels, 1, 1, 0, bias=True)
        self.conv2 = nn.Conv2d(out_channels, out_channels, 3, 1, 1, bias=True, groups=out_channels)
        if withBNRelu:
            self.bn = nn.BatchNorm2d(out_channels)

    def forward(self, x):
        raise RuntimeError('ConvDPUnit runs inside the fused sm_100a kernels; call the detector / '
                           'head that owns it')


class Conv_head(nn.Module):  # yunet_layer.py:39-62

    def __init__(self, in_channels, mid_channels, out_channels):
        super().__init__()
        self.conv1 = nn.Conv2d(in_channels, mid_channels, 3, 2, 1, bias=True)
        self.conv2 = ConvDPUnit(mid_channels, out_channels, True)
        self.bn1 = nn.BatchNorm2d(mid_channels)


class Conv4layerBlock(nn.Module):  # yunet_layer.py:65-82

    def __init__(self, in_channels, out_channels, withBNRelu=True):
        super().__init__()
        self.conv1 = ConvDPUnit(in_channels, in_channels, True)
        self.conv2 = ConvDPUnit(in_channels, out_channels, withBNRelu)


class FusedFeatures(list):
    """What flows between backbone, neck and head: the input batch plus the modules that have
    claimed it.  Behaves as a list of three per-level entries so code that only forwards or counts
    the features (single_stage.py:52-57) keeps working; ``materialize()`` gives real NCHW tensors
    of the current stage (backbone or neck outputs)."""

    def __init__(self, img, backbone):
        super().__init__([None, None, None])
        self.img = img
        self.backbone = backbone
        self.neck = None

    def materialize(self, head=None, train=False):
        eng = _engine_for(self.backbone, self.neck, head)
        eng.sync_from_modules()
        B, _, H, W = self.img.shape
        eng.core.forward(self.img, train=train)
        units = eng.core.ctx.units()
        want = ['neck.lateral_convs.%d' % i for i in range(3)] if self.neck is not None else None
        out = []
        if want is None:
            a = eng.core.arch
            names = [f'backbone.model{i}.conv2' for i in a['out_idx']]
        else:
            names = want
        for n in names:
            idx = [i for i, u in enumerate(units) if u.name.decode() == n][0]
            out.append(eng.core.read_activation(idx, B, H, W, train=train))
        return out


# --------------------------------------------------------------------------------- plugins
@BACKBONES.register_module()
class YuNetBackbone(nn.Module):

    def __init__(self, stage_channels, downsample_idx, out_idx):
        super().__init__()
        self.layer_num = len(stage_channels)
        self.stage_channels = [list(s) for s in stage_channels]
        self.downsample_idx = list(downsample_idx)
        self.out_idx = list(out_idx)
        self.model0 = Conv_head(*stage_channels[0])
        for i in range(1, self.layer_num):
            self.add_module(f'model{i}', Conv4layerBlock(*stage_channels[i]))
        self.init_weights()

    def init_weights(self, pretrained=None):
        _init_like_reference(self)

    def forward(self, x):
        return FusedFeatures(x, self)


@NECKS.register_module()
class TFPN(nn.Module):

    def __init__(self, in_channels, out_idx):
        super().__init__()
        self.num_layers = len(in_channels)
        self.out_idx = list(out_idx)
        self.lateral_convs = nn.ModuleList(
            [ConvDPUnit(in_channels[i], in_channels[i], True) for i in range(self.num_layers)])
        self.init_weights()

    def init_weights(self):
        _init_like_reference(self)

    def forward(self, feats):
        if not isinstance(feats, FusedFeatures):
            raise TypeError('TFPN (B200) consumes the FusedFeatures handle produced by YuNetBackbone')
        feats.neck = self
        return feats


class AssignResult:
    """The fields of mmdet's ``AssignResult`` the YuNet head reads (assign_result.py): ``num_gts``,
    ``gt_inds`` (P,) long, 1-based gt index or 0, ``max_overlaps`` (P,) matched IoU (-1e5 for
    non-positives, all zero when nothing could be assigned), ``labels`` (P,) long or None."""

    def __init__(self, num_gts, gt_inds, max_overlaps, labels=None):
        self.num_gts = num_gts
        self.gt_inds = gt_inds
        self.max_overlaps = max_overlaps
        self.labels = labels

    @property
    def num_preds(self):
        return len(self.gt_inds)


@BBOX_ASSIGNERS.register_module()
class SimOTAAssigner:
    """``mmdet/core/bbox/assigners/sim_ota_assigner.py:10-93``.  Inside the fused head the batched
    ``yunet_simota_assign`` kernel assigns all images at once; ``assign`` is the reference's
    per-image entry point (``yunet_head.py:575-577``) on the same kernel through
    ``yunet_simota_assign_ext``, so a reference head that calls it keeps working."""

    def __init__(self, center_radius=2.5, candidate_topk=10, iou_weight=3.0, cls_weight=1.0):
        self.center_radius = center_radius
        self.candidate_topk = candidate_topk
        self.iou_weight = iou_weight
        self.cls_weight = cls_weight
        self._engine = None

    def _core(self, device):
        if self._engine is None or self._engine.device != device:
            lc = _capi.default_loss_cfg()
            lc.center_radius = float(self.center_radius)
            lc.candidate_topk = int(self.candidate_topk)
            lc.iou_weight = float(self.iou_weight)
            lc.cls_weight = float(self.cls_weight)
            self._engine = YuNetEngine('yunet_n', device=device, loss_cfg=lc)
        return self._engine

    def assign(self, pred_scores, priors, decoded_bboxes, gt_bboxes, gt_labels,
               gt_bboxes_ignore=None, eps=1e-7):
        """pred_scores (P,1), priors (P,4) ``[cx, cy, stride_w, stride_h]`` (offset by half a
        stride), decoded_bboxes (P,4), gt_bboxes (G,4), gt_labels (G,) -> ``AssignResult``."""
        if pred_scores.dim() == 2 and pred_scores.shape[1] != 1:
            raise NotImplementedError('the B200 SimOTA kernel is single-class (YuNet: num_classes=1)')
        dev = decoded_bboxes.device
        if dev.type != 'cuda':
            raise RuntimeError('SimOTAAssigner.assign needs CUDA tensors (there is no CPU fallback)')
        num_gt, P_ = int(gt_bboxes.shape[0]), int(decoded_bboxes.shape[0])
        gt_inds = torch.zeros(P_, dtype=torch.long, device=dev)
        labels = None if gt_labels is None else torch.full((P_,), -1, dtype=torch.long, device=dev)
        if num_gt == 0 or P_ == 0:
            return AssignResult(num_gt, gt_inds, torch.zeros(P_, device=dev), labels)
        assigned, miou, counters = self._core(dev).assign_ext(pred_scores.reshape(-1), priors,
                                                               decoded_bboxes, gt_bboxes)
        gt_inds = assigned.long()
        pos = gt_inds > 0
        if not bool(pos.any()):      # no prior inside any gt box / centre region
            return AssignResult(num_gt, gt_inds, torch.zeros(P_, device=dev), labels)
        max_overlaps = torch.where(pos, miou, torch.full_like(miou, -100000.0))
        if labels is not None:
            labels[pos] = gt_labels[gt_inds[pos] - 1].long()
        return AssignResult(num_gt, gt_inds, max_overlaps, labels)


def _loss_weight(cfg, default):
    return float(cfg.get('loss_weight', default)) if cfg else default


@HEADS.register_module()
class YuNet_Head(nn.Module):

    def __init__(self, num_classes, in_channels, feat_channels=256, shared_stacked_convs=2,
                 stacked_convs=2, loss_cls=None, loss_bbox=None, use_kps=False, kps_num=5,
                 loss_kps=None, prior_generator=None, train_cfg=None, test_cfg=None, loss_obj=None):
        super().__init__()
        if stacked_convs != 0:
            raise NotImplementedError('the YuNet configs use stacked_convs=0 (configs/yunet_n.py:118)')
        if not use_kps or kps_num != 5 or num_classes != 1:
            raise NotImplementedError('fused head supports num_classes=1, use_kps=True, kps_num=5')
        if loss_bbox is not None and loss_bbox.get('type', 'EIoULoss') != 'EIoULoss':
            raise NotImplementedError('fused loss implements EIoULoss (configs/yunet_n.py:127)')
        self.num_classes, self.NK = num_classes, kps_num
        self.in_channels, self.feat_channels = in_channels, feat_channels
        self.shared_stack_convs, self.stacked_convs = shared_stacked_convs, stacked_convs
        self.use_kps = use_kps
        pg = prior_generator or dict(strides=[8, 16, 32], offset=0)
        self.strides = [s if isinstance(s, int) else s[0] for s in pg['strides']]
        self.train_cfg, self.test_cfg = train_cfg, test_cfg
        self.assigner = None
        if train_cfg:
            acfg = dict(train_cfg['assigner'])
            acfg.pop('type', None)
            self.assigner = SimOTAAssigner(**acfg)
        lc = _capi.default_loss_cfg()
        lc.loss_cls_weight = _loss_weight(loss_cls, 1.0)
        lc.loss_bbox_weight = _loss_weight(loss_bbox, 5.0)
        lc.loss_obj_weight = _loss_weight(loss_obj, 1.0)
        lc.loss_kps_weight = _loss_weight(loss_kps, 0.1)
        if loss_kps:
            lc.smooth_l1_beta = float(loss_kps.get('beta', 1.0))
        if loss_bbox:
            lc.eiou_eps = float(loss_bbox.get('eps', 1e-6))
            lc.eiou_smooth_point = float(loss_bbox.get('smooth_point', 0.1))
        if self.assigner is not None:
            lc.center_radius = self.assigner.center_radius
            lc.candidate_topk = self.assigner.candidate_topk
            lc.iou_weight = self.assigner.iou_weight
            lc.cls_weight = self.assigner.cls_weight
        self.loss_cfg = lc
        if shared_stacked_convs > 0:
            self.multi_level_share_convs = nn.ModuleList()
        self.multi_level_cls = nn.ModuleList()
        self.multi_level_bbox = nn.ModuleList()
        self.multi_level_obj = nn.ModuleList()
        self.multi_level_kps = nn.ModuleList()
        for _ in self.strides:
            if shared_stacked_convs > 0:
                convs = [ConvDPUnit(in_channels if i == 0 else feat_channels, feat_channels)
                         for i in range(shared_stacked_convs)]
                self.multi_level_share_convs.append(nn.Sequential(*convs))
            chn = in_channels if shared_stacked_convs == 0 else feat_channels
            self.multi_level_cls.append(ConvDPUnit(chn, num_classes, False))
            self.multi_level_bbox.append(ConvDPUnit(chn, 4, False))
            self.multi_level_kps.append(ConvDPUnit(chn, kps_num * 2, False))
            self.multi_level_obj.append(ConvDPUnit(chn, 1, False))
        self.init_weights()

    def init_weights(self):
        _init_like_reference(self)

    # ---- forward (yunet_head.py:175-247): raw NCHW maps per level
    def forward(self, feats):
        if not isinstance(feats, FusedFeatures):
            raise TypeError('YuNet_Head (B200) consumes the FusedFeatures handle of YuNetBackbone/TFPN')
        eng = _engine_for(feats.backbone, feats.neck, self)
        eng.sync_from_modules()
        preds = eng.core.forward(feats.img, train=self.training)
        return _split_levels(preds, feats.img.shape[2], feats.img.shape[3], self.strides)

    # ---- forward_train (yunet_head.py:249-288) -> dict of four losses with autograd history
    def forward_train(self, x, img_metas, gt_bboxes, gt_labels=None, gt_keypointss=None,
                      gt_bboxes_ignore=None, proposal_cfg=None, **kwargs):
        if not isinstance(x, FusedFeatures):
            raise TypeError('YuNet_Head (B200) consumes the FusedFeatures handle of YuNetBackbone/TFPN')
        eng = _engine_for(x.backbone, x.neck, self)
        return eng.losses(x.img, gt_bboxes, gt_keypointss)

    # ---- loss (yunet_head.py:418-534) on caller-provided head outputs: what a reference-style
    # ``forward_train`` (``outs = self(x); self.loss(*outs, gt_bboxes, gt_labels, gt_keypointss,
    # img_metas)``, yunet_head.py:276-283) or a foreign backbone calls.  SimOTA + the four losses run
    # in the same two kernels as the fused path; the gradient flows back into the given tensors.
    def loss(self, cls_scores, bbox_preds, objectnesses, kps_preds, gt_bboxes, gt_labels, gt_kpss,
             img_metas, gt_bboxes_ignore=None):
        if self.strides != [8, 16, 32]:
            raise NotImplementedError('the fused loss is built for strides [8, 16, 32] (configs/yunet_n.py:122)')
        if gt_kpss is None:
            raise ValueError('YuNet_Head.loss needs gt_kpss (use_kps=True, configs/yunet_n.py:129)')
        B = cls_scores[0].shape[0]
        H, W = cls_scores[0].shape[2] * self.strides[0], cls_scores[0].shape[3] * self.strides[0]
        dev = cls_scores[0].device
        if dev.type != 'cuda':
            raise RuntimeError('YuNet_Head.loss needs CUDA tensors (there is no CPU fallback)')

        def fl(lst, c):     # the reference's permute(0, 2, 3, 1).reshape(B, -1, c) per level, then cat
            return torch.cat([t.permute(0, 2, 3, 1).reshape(B, -1, c) for t in lst], 1)

        preds = torch.cat([fl(cls_scores, 1), fl(bbox_preds, 4), fl(objectnesses, 1),
                           fl(kps_preds, 10)], -1).float().contiguous()
        core = self.__dict__.get('_b200_loss_engine')
        if core is None or core.device != dev:
            core = YuNetEngine('yunet_n', device=dev, loss_cfg=self.loss_cfg)
            self.__dict__['_b200_loss_engine'] = core
        gt, offs = pack_gt_csr_device(gt_bboxes, gt_kpss, dev)
        l = _HeadLoss.apply(core, gt, offs, H, W, preds)
        return dict(loss_cls=l[0], loss_bbox=l[1], loss_obj=l[2], loss_kps=l[3])

    # ---- get_bboxes (yunet_head.py:290-374): decode + score filter + NMS on the GPU
    def get_bboxes(self, cls_scores, bbox_preds, objectnesses, kps_preds, img_metas=None, cfg=None,
                   rescale=False, with_nms=True):
        cfg = self.test_cfg if cfg is None else cfg
        B = cls_scores[0].shape[0]
        H, W = cls_scores[0].shape[2] * self.strides[0], cls_scores[0].shape[3] * self.strides[0]

        def fl(lst, c):
            return torch.cat([t.permute(0, 2, 3, 1).reshape(B, -1, c) for t in lst], 1)

        preds = torch.cat([fl(cls_scores, 1), fl(bbox_preds, 4), fl(objectnesses, 1),
                           fl(kps_preds, 10)], -1).contiguous()
        sf = None
        if rescale:
            sf = torch.as_tensor(np.array([m['scale_factor'] for m in img_metas], np.float32),
                                 device=preds.device).reshape(B, 4).contiguous()
        core = _any_engine(preds.device)
        nms = cfg['nms'] if isinstance(cfg, dict) else cfg.nms
        thr = cfg['score_thr'] if isinstance(cfg, dict) else cfg.score_thr
        iou = nms.get('iou_threshold', nms.get('iou_thr', 0.45))
        dets, counts, _ = core.decode_nms(preds, H, W, float(thr), float(iou), scale_factors=sf)
        counts = counts.cpu().tolist()
        out = []
        for b in range(B):
            d = dets[b, :counts[b]].clone()
            out.append((d, torch.zeros(counts[b], dtype=torch.long, device=d.device)))
        return out


def _split_levels(preds, H, W, strides):
    B = preds.shape[0]
    outs = ([], [], [], [])
    off = 0
    for s in strides:
        h, w = H // s, W // s
        sl = preds[:, off:off + h * w].reshape(B, h, w, 16).permute(0, 3, 1, 2)
        off += h * w
        outs[0].append(sl[:, 0:1]); outs[1].append(sl[:, 1:5])
        outs[2].append(sl[:, 5:6]); outs[3].append(sl[:, 6:16])
    return outs


# --------------------------------------------------------------------------------- fused engine glue
class _FusedLoss(torch.autograd.Function):
    """loss dict with autograd history: backward runs ``yunet_backward`` and hands every module
    parameter its slice of the flat gradient bucket."""

    @staticmethod
    def forward(ctx, glue, img, gt, offs, *params):
        core = glue.core
        B, _, H, W = img.shape
        core.train_step_forward = True
        preds = core.forward(img, train=True)
        assigned, miou, counters = core.assign(preds, gt, offs, H, W)
        num_total = counters
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            num_total = counters[:1].clone() / dist.get_world_size()   # reduce_mean
            dist.all_reduce(num_total)
        losses, d_preds = core.loss_grad(preds, gt, offs, assigned, miou, counters, num_total, H, W)
        ctx.glue, ctx.img = glue, img
        ctx.saved = (preds, gt, offs, assigned, miou, counters, num_total, d_preds)
        out = losses.clone()
        return out[0], out[1], out[2], out[3]

    @staticmethod
    def backward(ctx, g_cls, g_bbox, g_obj, g_kps):
        glue, img = ctx.glue, ctx.img
        core = glue.core
        preds, gt, offs, assigned, miou, counters, num_total, d_preds = ctx.saved
        scale = [float(g) for g in (g_cls, g_bbox, g_obj, g_kps)]
        if scale != [1.0, 1.0, 1.0, 1.0]:
            B, _, H, W = img.shape
            _, d_preds = core.loss_grad(preds, gt, offs, assigned, miou, counters, num_total, H, W,
                                        loss_scale=scale)
        core.backward(img, d_preds)
        # every parameter gets a VIEW of the flat gradient bucket (no copies): autograd's
        # AccumulateGrad adopts it as ``.grad`` when the gradient was None (zero_grad(set_to_none=True),
        # torch's default) and adds it in place otherwise; the bucket is rewritten by the next backward.
        # The views must be FRESH objects: AccumulateGrad only steals a gradient nobody else references
        # (a cached tuple of views makes it clone every one of them).
        return (None, None, None, None) + glue.fresh_grad_views()


class _HeadLoss(torch.autograd.Function):
    """``YuNet_Head.loss`` on explicit predictions (B, P, 16): ``yunet_simota_assign`` +
    ``yunet_loss_grad``; backward returns d loss / d preds (recomputed with the upstream scales when
    the four losses are not simply summed)."""

    @staticmethod
    def forward(ctx, core, gt, offs, H, W, preds):
        assigned, miou, counters = core.assign(preds, gt, offs, H, W)
        num_total = counters
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            num_total = counters[:1].clone() / dist.get_world_size()   # reduce_mean, yunet_head.py:493-497
            dist.all_reduce(num_total)
        losses, d_preds = core.loss_grad(preds, gt, offs, assigned, miou, counters, num_total, H, W)
        # the engine's buffers are reused by its next call: keep private copies for backward
        ctx.core, ctx.hw = core, (H, W)
        ctx.saved = (preds.detach(), gt, offs, assigned.clone(), miou.clone(), counters.clone(),
                     num_total.clone(), d_preds.clone())
        out = losses.clone()
        return out[0], out[1], out[2], out[3]

    @staticmethod
    def backward(ctx, g_cls, g_bbox, g_obj, g_kps):
        preds, gt, offs, assigned, miou, counters, num_total, d_preds = ctx.saved
        scale = [float(g) for g in (g_cls, g_bbox, g_obj, g_kps)]
        if scale != [1.0, 1.0, 1.0, 1.0]:
            H, W = ctx.hw
            _, d = ctx.core.loss_grad(preds, gt, offs, assigned, miou, counters, num_total, H, W,
                                      loss_scale=scale)
            d_preds = d.clone()
        return None, None, None, None, None, d_preds


def pack_gt_csr_device(gt_bboxes, gt_keypointss, device):
    """Ragged per-image GT lists -> the CSR arrays of the C ABI, built ON THE DEVICE (the reference
    hands the head device tensors; no ``.cpu()`` per step): gt (sumG, 19) fp32 rows
    ``[x1,y1,x2,y2, kx0,ky0,..,kx4,ky4, w0..w4]`` and offsets (B+1,) int32.  The row counts come from
    the tensor shapes, so nothing synchronises."""
    counts = [int(b.shape[0]) for b in gt_bboxes]
    offs = torch.tensor(np.concatenate([[0], np.cumsum(counts)]).astype(np.int32))
    offs = offs.to(device, non_blocking=True)
    total = int(sum(counts))
    gt = torch.empty(max(total, 0), 19, device=device, dtype=torch.float32)
    if total:
        # (a metadata call per image costs more than the copy at 256 images: reshape only odd shapes)
        bb = torch.cat([b if b.dim() == 2 else b.reshape(-1, 4) for b in gt_bboxes])
        kp = torch.cat([k if k.dim() == 3 else k.reshape(-1, 5, 3) for k in gt_keypointss])
        bb = bb.to(device=device, dtype=torch.float32)
        kp = kp.to(device=device, dtype=torch.float32)
        gt[:, :4] = bb
        gt[:, 4:14] = kp[:, :, :2].reshape(-1, 10)
        gt[:, 14:19] = kp[:, :, 2]
    return gt, offs


class _Glue:
    """One engine per (backbone, neck, head) triple; module parameters become views of the
    engine's flat bucket so optimiser updates land where the kernels read."""

    def __init__(self, backbone, neck, head):
        dev = next(head.parameters()).device
        if dev.type != 'cuda':
            raise RuntimeError('the B200 YuNet plugins need their parameters on a CUDA device')
        arch = dict(stage_channels=backbone.stage_channels, downsample_idx=backbone.downsample_idx,
                    out_idx=backbone.out_idx, shared_stacked_convs=head.shared_stack_convs,
                    feat_channels=head.feat_channels)
        self.core = YuNetEngine(arch, device=dev, loss_cfg=head.loss_cfg)
        self.modules = {'backbone': backbone, 'neck': neck, 'bbox_head': head}
        self.names = [n for n, _, _ in self.core.param_table]
        mv = self.core.param_views(self.core.momentum_buf)
        self.momentum_views = tuple(mv[k] for k in self.names)    # ``SGD.state[p]['momentum_buffer']``
        self.momentum_ptrs = [v.data_ptr() for v in self.momentum_views]
        # the gradient views handed out by every backward: addresses are fixed, the view objects are not
        table = self.core.param_table
        self._shapes = [tuple(shape) for _, _, shape in table]
        self._sizes = [int(np.prod(shape)) for _, _, shape in table]
        self._dense = all(table[i + 1][1] == table[i][1] + self._sizes[i] for i in range(len(table) - 1)) \
            and table[0][1] == 0 and table[-1][1] + self._sizes[-1] == self.core.grads.numel()
        self.grad_ptrs = [self.core.grads.data_ptr() + 4 * off for _, off, _ in table]
        self._adopt()
        _LIVE_GLUES.add(self)

    def fresh_grad_views(self):
        g = self.core.grads
        if self._dense:      # one split call instead of a slice per parameter
            return tuple(v.view(sh) for v, sh in zip(g.split(self._sizes), self._shapes))
        views = self.core.param_views(g)
        return tuple(views[k] for k in self.names)

    def _named(self):
        p, b = {}, {}
        for prefix, m in self.modules.items():
            for k, v in m.named_parameters():
                p[f'{prefix}.{k}'] = v
            for k, v in m.named_buffers():
                b[f'{prefix}.{k}'] = v
        return p, b

    def _adopt(self):
        p, b = self._named()
        views = self.core.param_views()
        assert set(views) == set(p), sorted(set(views) ^ set(p))
        with torch.no_grad():
            for k, v in views.items():
                v.copy_(p[k].detach().reshape(v.shape))
                p[k].data = v                     # parameter now aliases the bucket
            nbn = self.core.ctx.num_bn_channels
            for name, off, ch in self.core.bn_table:
                rm = self.core.bn_running[off:off + ch]
                rv = self.core.bn_running[nbn + off:nbn + off + ch]
                rm.copy_(b[name + '.running_mean'])
                rv.copy_(b[name + '.running_var'])
                owner, leaf = self._resolve(name)
                owner.running_mean = rm
                owner.running_var = rv
        self.params = [p[k] for k in self.names]
        self.param_ptrs = [q.data_ptr() for q in self.params]
        # per-step bookkeeping without walking the module tree by name: the parameters in module
        # order with the addresses they must keep, and the BatchNorm step counters
        self._mod_params = [q for m in self.modules.values() for q in m.parameters()]
        self._mod_ptrs = [q.data_ptr() for q in self._mod_params]
        self._nbt = [mod.num_batches_tracked for m in self.modules.values() for mod in m.modules()
                     if isinstance(mod, nn.BatchNorm2d)]

    def _resolve(self, dotted):
        prefix, rest = dotted.split('.', 1)
        m = self.modules[prefix]
        for part in rest.split('.'):
            m = getattr(m, part) if not part.isdigit() else m[int(part)]
        return m, rest

    def sync_from_modules(self):
        """Re-adopt if something (``.to()``, ``load_state_dict`` on a fresh tensor) re-pointed a
        parameter away from the bucket."""
        cur = [q for m in self.modules.values() for q in m.parameters()]
        same = len(cur) == len(self._mod_params)
        if same:
            for q, q0, ptr in zip(cur, self._mod_params, self._mod_ptrs):
                if q is not q0 or q.data_ptr() != ptr:
                    same = False
                    break
        if not same:
            self._adopt()

    def losses(self, img, gt_bboxes, gt_keypointss):
        self.sync_from_modules()
        gt, offs = pack_gt_csr_device(gt_bboxes, gt_keypointss, img.device)
        l = _FusedLoss.apply(self, img.contiguous(), gt, offs, *self.params)
        torch._foreach_add_(self._nbt, 1)       # every BatchNorm2d.num_batches_tracked, one launch
        return dict(loss_cls=l[0], loss_bbox=l[1], loss_obj=l[2], loss_kps=l[3])


_ENGINES = {}
_LIVE_GLUES = weakref.WeakSet()     # looked up by ``SGD`` to find the bucket its parameters alias


def _engine_for(backbone, neck, head):
    """The engine glue of a (backbone, neck, head) triple lives ON the head module (plain attribute,
    not a sub-module), so it is freed with the model instead of accumulating in a global table."""
    if neck is None or head is None:
        raise RuntimeError('the fused path needs backbone, neck and head')
    g = head.__dict__.get('_b200_glue')
    if g is None or g.modules['backbone'] is not backbone or g.modules['neck'] is not neck:
        g = _Glue(backbone, neck, head)
        head.__dict__['_b200_glue'] = g
    return g


def _any_engine(device):
    e = _ENGINES.get(str(device))
    if e is None:
        e = YuNetEngine('yunet_n', device=device)
        _ENGINES[str(device)] = e
    return e


class SGD(torch.optim.SGD):
    """``torch.optim.SGD`` — what mmcv's ``build_optimizer`` makes of the reference's
    ``optimizer = dict(type='SGD', lr=0.01, momentum=0.9, weight_decay=0.0005)`` (configs/yunet_n.py:1,
    mmdet/apis/train.py:117) — whose ``step()`` is ONE ``yunet_sgd_step`` launch over the flat bucket when
    that is exactly what the stock step would compute: a single param group without dampening /
    nesterov / maximize whose parameters are all the views of one engine's bucket and whose ``.grad``
    tensors are all the matching views of its gradient bucket (what ``_FusedLoss.backward`` hands out;
    in-place edits such as ``clip_grad_norm_`` or DDP's averaging land in the bucket and are honoured).
    Anything else (a ``None`` gradient, accumulated gradients in foreign storage, several groups)
    takes ``torch.optim.SGD.step`` unchanged.  ``state[p]['momentum_buffer']`` are views of the engine's
    flat momentum bucket, so ``state_dict()`` / ``load_state_dict()`` keep the reference checkpoint
    layout and both paths share one state (a zero buffer reproduces torch's first step ``buf = grad``)."""

    def _fused_glue(self):
        if len(self.param_groups) != 1:
            return None
        g = self.param_groups[0]
        if g.get('dampening', 0) != 0 or g.get('nesterov', False) or g.get('maximize', False) or \
                g.get('differentiable', False) or isinstance(g['lr'], torch.Tensor):
            return None
        ps = g['params']
        glue = getattr(self, '_b200_glue_ref', None)
        glue = glue() if glue is not None else None
        if glue is None or len(glue.params) != len(ps):
            glue = None
            ids = {id(q) for q in ps}
            for cand in list(_LIVE_GLUES):
                if len(cand.params) == len(ps) and all(id(q) in ids for q in cand.params):
                    glue = cand
                    break
            if glue is None:
                return None
            self._b200_glue_ref = weakref.ref(glue)
        for q, pptr, gptr in zip(glue.params, glue.param_ptrs, glue.grad_ptrs):
            gr = q.grad
            if gr is None or gr.data_ptr() != gptr or q.data_ptr() != pptr:
                return None
        return glue

    @torch.no_grad()
    def step(self, closure=None):
        glue = self._fused_glue()
        if glue is None:
            return super().step(closure)
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        g = self.param_groups[0]
        if g['momentum'] != 0:
            # one shared momentum state: whatever a stock step or load_state_dict left in ``state`` is
            # moved into the flat bucket, and ``state`` then holds the bucket's views
            for q, view, mptr in zip(glue.params, glue.momentum_views, glue.momentum_ptrs):
                st = self.state[q]
                buf = st.get('momentum_buffer')
                if buf is None or buf.data_ptr() != mptr:
                    if buf is not None:
                        view.copy_(buf.reshape(view.shape))
                    st['momentum_buffer'] = view
        glue.core.sgd_step(lr=float(g['lr']), momentum=float(g['momentum']),
                           weight_decay=float(g['weight_decay']), grad_scale=1.0)
        return loss


OPTIMIZERS.register_module(module=SGD)


@DETECTORS.register_module()
class YuNet(nn.Module):
    """mmdet/models/detectors/yunet.py:7-86 on the fused engine."""

    def __init__(self, backbone, neck, bbox_head, train_cfg=None, test_cfg=None, pretrained=None,
                 init_cfg=None):
        super().__init__()
        b = dict(backbone); b.pop('type', None)
        n = dict(neck); n.pop('type', None)
        h = dict(bbox_head); h.pop('type', None)
        h.update(train_cfg=train_cfg, test_cfg=test_cfg)     # single_stage.py:28-29
        self.backbone = YuNetBackbone(**b)
        self.neck = TFPN(**n)
        self.bbox_head = YuNet_Head(**h)
        self.train_cfg, self.test_cfg = train_cfg, test_cfg

    def extract_feat(self, img):                              # single_stage.py:52-57
        return self.neck(self.backbone(img))

    def forward_train(self, img, img_metas, gt_bboxes, gt_labels, gt_keypointss=None,
                      gt_bboxes_ignore=None):                 # yunet.py:21-51
        x = self.extract_feat(img)
        return self.bbox_head.forward_train(x, img_metas, gt_bboxes, gt_labels, gt_keypointss,
                                            gt_bboxes_ignore)

    def feature_test(self, img):                              # yunet.py:83-86
        return self.bbox_head(self.extract_feat(img))

    def simple_test(self, img, img_metas, rescale=False):     # yunet.py:53-81
        outs = self.bbox_head(self.extract_feat(img))
        results = self.bbox_head.get_bboxes(*outs, img_metas, rescale=rescale)
        # bbox2result (mmdet/core/bbox/transforms.py:116-133): per image, list over classes
        return [[d.detach().cpu().numpy()] for d, _ in results]

    def forward(self, img, img_metas=None, return_loss=True, **kwargs):   # base.py:168-182
        if return_loss:
            return self.forward_train(img, img_metas, **kwargs)
        return self.simple_test(img, img_metas, **kwargs)

    @staticmethod
    def _parse_losses(losses):                                # base.py:184-217
        """``(loss, log_vars)``: ``loss`` = sum of the entries whose key contains 'loss' (with autograd
        history), ``log_vars`` = python floats, averaged over the ranks when a process group is up
        (base.py:210-214).  The reference all-reduces and ``.item()``s entry by entry; here the entries
        travel as ONE stacked vector: one collective and one device-to-host read per step."""
        from collections import OrderedDict
        import torch.distributed as dist
        log_vars = OrderedDict()
        for name, value in losses.items():
            if isinstance(value, torch.Tensor):
                log_vars[name] = value.mean() if value.dim() else value
            elif isinstance(value, list):
                log_vars[name] = sum(_l.mean() for _l in value)
            else:
                raise TypeError(f'{name} is not a tensor or list of tensors')
        loss = sum(v for k, v in log_vars.items() if 'loss' in k)
        log_vars['loss'] = loss
        vec = torch.stack([v.detach().float() for v in log_vars.values()])
        if dist.is_available() and dist.is_initialized():
            vec = vec / dist.get_world_size()
            dist.all_reduce(vec)
        for k, x in zip(list(log_vars), vec.tolist()):
            log_vars[k] = x
        return loss, log_vars

    def train_step(self, data, optimizer=None):               # base.py:219-252
        losses = self(**data)
        loss, log_vars = self._parse_losses(losses)
        return dict(loss=loss, log_vars=log_vars, num_samples=len(data['img_metas']))
