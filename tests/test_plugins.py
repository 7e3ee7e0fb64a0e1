"""CPU tests of the plugin surface: same names, ctor kwargs and state_dict keys as the reference
(mmdet/models/{backbones/yunet_backbone,necks/tfpn,dense_heads/yunet_head}.py), so the shipped
checkpoints load strict."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN
from libfacedetection.train_b200 import plugins

CFG = {
    'yunet_n': dict(stage=[[3, 16, 16], [16, 64], [64, 64], [64, 64], [64, 64], [64, 64]], shared=1),
    'yunet_s': dict(stage=[[3, 16, 16], [16, 32], [32, 64], [64, 64], [64, 64], [64, 64]], shared=0),
}


def model_cfg(arch):
    c = CFG[arch]
    # the `model` dict of configs/yunet_{n,s}.py:104-145, verbatim structure
    return dict(
        type='YuNet',
        backbone=dict(type='YuNetBackbone', stage_channels=c['stage'], downsample_idx=[0, 2, 3, 4],
                      out_idx=[3, 4, 5]),
        neck=dict(type='TFPN', in_channels=[64, 64, 64], out_idx=[0, 1, 2]),
        bbox_head=dict(
            type='YuNet_Head', num_classes=1, in_channels=64, shared_stacked_convs=c['shared'],
            stacked_convs=0, feat_channels=64,
            prior_generator=dict(type='MlvlPointGenerator', offset=0, strides=[8, 16, 32]),
            loss_cls=dict(type='CrossEntropyLoss', use_sigmoid=True, reduction='sum', loss_weight=1.0),
            loss_bbox=dict(type='EIoULoss', loss_weight=5.0, reduction='sum'),
            use_kps=True, kps_num=5,
            loss_kps=dict(type='SmoothL1Loss', beta=0.1111111111111111, loss_weight=0.1),
            loss_obj=dict(type='CrossEntropyLoss', use_sigmoid=True, reduction='sum', loss_weight=1.0)),
        train_cfg=dict(assigner=dict(type='SimOTAAssigner', center_radius=2.5)),
        test_cfg=dict(nms_pre=-1, min_bbox_size=0, score_thr=0.02,
                      nms=dict(type='nms', iou_threshold=0.45), max_per_img=-1))


@pytest.mark.parametrize('arch', ['yunet_n', 'yunet_s'])
def test_checkpoint_loads_strict(arch):
    model = plugins.DETECTORS.build(model_cfg(arch))
    d = np.load(os.path.join(GOLDEN, f'weights_{arch}.npz'))
    sd = {k: torch.from_numpy(d[k]) for k in d.files}
    mine = model.state_dict()
    assert set(mine) == set(sd)
    for k in sd:
        assert tuple(mine[k].shape) == tuple(sd[k].shape), k
    model.load_state_dict(sd, strict=True)
    assert sum(p.numel() for p in model.parameters()) == {'yunet_n': 75856, 'yunet_s': 54608}[arch]
    lc = model.bbox_head.loss_cfg
    assert abs(lc.loss_bbox_weight - 5.0) < 1e-6 and abs(lc.smooth_l1_beta - 1 / 9) < 1e-6
    assert abs(lc.center_radius - 2.5) < 1e-6 and lc.candidate_topk == 10


def test_registries_hold_reference_names():
    assert plugins.BACKBONES.get('YuNetBackbone') is plugins.YuNetBackbone
    assert plugins.NECKS.get('TFPN') is plugins.TFPN
    assert plugins.HEADS.get('YuNet_Head') is plugins.YuNet_Head
    assert plugins.BBOX_ASSIGNERS.get('SimOTAAssigner') is plugins.SimOTAAssigner
    a = plugins.BBOX_ASSIGNERS.build(dict(type='SimOTAAssigner', center_radius=2.5))
    assert (a.center_radius, a.candidate_topk, a.iou_weight, a.cls_weight) == (2.5, 10, 3.0, 1.0)


def test_reference_init_statistics():
    torch.manual_seed(0)
    m = plugins.DETECTORS.build(model_cfg('yunet_n'))
    sd = m.state_dict()
    assert float(sd['backbone.model2.conv1.conv1.bias'].mean()) == pytest.approx(0.02)
    assert float(sd['backbone.model2.conv1.bn.weight'].mean()) == 1.0
    w = sd['backbone.model2.conv1.conv1.weight']     # xavier normal: std = sqrt(2/(64+64))
    assert abs(float(w.std()) - (2.0 / 128) ** 0.5) < 0.01


def test_cpu_parameters_are_refused():
    m = plugins.DETECTORS.build(model_cfg('yunet_s'))
    with pytest.raises(RuntimeError):
        m.feature_test(torch.zeros(1, 3, 64, 64))


@pytest.mark.reference
def test_register_into_mmdet_builds_the_plugins_from_the_reference_config():
    """With the reference's own registries (imported from /root/reference through the mmcv stub of
    oracle/ref_loader.py): after ``register_into_mmdet()`` the real ``configs/yunet_n.py`` model
    dict builds the plugin classes — the reference's ``build_detector`` call site needs no change —
    and the assigner the head config names resolves to the plugin with an ``assign`` method."""
    from oracle import ref_loader
    if not ref_loader.reference_available():
        pytest.skip('/root/reference not present')
    ref_loader.install()
    from mmdet.models.builder import MODELS, build_detector
    from mmdet.core.bbox.builder import BBOX_ASSIGNERS as MM_ASSIGNERS, build_assigner
    saved = {k: MODELS.get(k) for k in ('YuNetBackbone', 'TFPN', 'YuNet_Head', 'YuNet')}
    saved_a = MM_ASSIGNERS.get('SimOTAAssigner')
    try:
        plugins.register_into_mmdet(force=True)
        cfg = ref_loader.load_config('yunet_n')
        det = build_detector(cfg.model)
        assert type(det) is plugins.YuNet
        assert type(det.backbone) is plugins.YuNetBackbone and type(det.neck) is plugins.TFPN
        assert type(det.bbox_head) is plugins.YuNet_Head
        import torch
        sd = torch.load('/root/reference/weights/yunet_n.pth', map_location='cpu', weights_only=False)['state_dict']
        det.load_state_dict(sd, strict=True)          # "All keys matched"
        asg = build_assigner(cfg.model.train_cfg.assigner)
        assert type(asg) is plugins.SimOTAAssigner and callable(asg.assign)
        assert asg.center_radius == 2.5 and asg.candidate_topk == 10
        # optimizer = dict(type='SGD', ...) (configs/yunet_n.py:1) resolves to the fused subclass
        from mmcv.runner.optimizer.builder import OPTIMIZERS as MM_OPT
        assert MM_OPT.get('SGD') is plugins.SGD and issubclass(MM_OPT.get('SGD'), torch.optim.SGD)
        assert cfg.optimizer['type'] == 'SGD'
    finally:
        for k, v in saved.items():
            MODELS.register_module(name=k, force=True, module=v)
        MM_ASSIGNERS.register_module(name='SimOTAAssigner', force=True, module=saved_a)


# ------------------------------------------------------------------------------ plugins.SGD host logic
class _StubCore:
    """The flat buckets of an engine on the CPU with ``sgd_step`` restated in torch (csrc/sgd.cu)."""

    def __init__(self, shapes):
        self.shapes = shapes
        n = sum(int(np.prod(s)) for s in shapes)
        self.params, self.grads, self.momentum_buf = torch.zeros(n), torch.zeros(n), torch.zeros(n)
        self.calls = 0

    def views(self, bucket):
        out, off = [], 0
        for s in self.shapes:
            n = int(np.prod(s))
            out.append(bucket[off:off + n].view(s))
            off += n
        return tuple(out)

    def sgd_step(self, lr, momentum, weight_decay, grad_scale):
        self.calls += 1
        g = self.grads * grad_scale + weight_decay * self.params
        self.momentum_buf.mul_(momentum).add_(g)
        self.params.sub_(lr * self.momentum_buf)


class _StubGlue:
    def __init__(self, shapes, seed=0):
        self.core = _StubCore(shapes)
        g = torch.Generator().manual_seed(seed)
        self.core.params.copy_(torch.randn(self.core.params.numel(), generator=g))
        self.params = [torch.nn.Parameter(v) for v in self.core.views(self.core.params)]
        for q, v in zip(self.params, self.core.views(self.core.params)):
            q.data = v
        self.grad_views = self.core.views(self.core.grads)
        self.momentum_views = self.core.views(self.core.momentum_buf)
        self.param_ptrs = [q.data_ptr() for q in self.params]
        self.grad_ptrs = [v.data_ptr() for v in self.grad_views]
        self.momentum_ptrs = [v.data_ptr() for v in self.momentum_views]
        plugins._LIVE_GLUES.add(self)

    def backward(self, seed):
        g = torch.Generator().manual_seed(100 + seed)
        self.core.grads.copy_(torch.randn(self.core.grads.numel(), generator=g))
        for q, v in zip(self.params, self.grad_views):
            q.grad = v


SHAPES = [(4, 3, 1, 1), (4,), (4, 1, 3, 3), (4,), (7,)]
KW = dict(lr=0.01, momentum=0.9, weight_decay=0.0005)


def test_fused_sgd_equals_torch_sgd_and_shares_the_momentum_state():
    ga, gb = _StubGlue(SHAPES), _StubGlue(SHAPES)
    fused = plugins.SGD(ga.params, **KW)
    stock = torch.optim.SGD([torch.nn.Parameter(q.detach().clone()) for q in gb.params], **KW)
    for it in range(4):
        ga.backward(it)
        gb.backward(it)
        for q, v in zip(stock.param_groups[0]['params'], gb.grad_views):
            q.grad = v.clone()
        fused.step()
        stock.step()
        for a, b in zip(ga.params, stock.param_groups[0]['params']):
            assert torch.allclose(a, b, rtol=1e-6, atol=1e-7)
    assert ga.core.calls == 4                               # every step took the one-launch path
    for q, view in zip(ga.params, ga.momentum_views):       # state holds the bucket's views
        assert fused.state[q]['momentum_buffer'].data_ptr() == view.data_ptr()
    for a, b in zip(ga.params, stock.param_groups[0]['params']):
        assert torch.allclose(fused.state[a]['momentum_buffer'], stock.state[b]['momentum_buffer'],
                              rtol=1e-6, atol=1e-7)
    # state_dict round trip into a fresh optimizer (resume): buffers are re-homed into the bucket
    import copy
    sd = copy.deepcopy(fused.state_dict())      # what torch.save / torch.load hand back
    gc = _StubGlue(SHAPES)
    gc.core.params.copy_(ga.core.params)
    resumed = plugins.SGD(gc.params, **KW)
    resumed.load_state_dict(sd)
    ga.backward(9); gc.backward(9)
    fused.step(); resumed.step()
    assert gc.core.calls == 1
    assert torch.allclose(gc.core.params, ga.core.params, rtol=1e-6, atol=1e-7)
    assert torch.allclose(gc.core.momentum_buf, ga.core.momentum_buf, rtol=1e-6, atol=1e-7)


def test_fused_sgd_falls_back_to_the_stock_step():
    # a None gradient, a gradient in foreign storage, several groups, nesterov: torch.optim.SGD.step
    for case in ('none_grad', 'foreign_grad', 'two_groups', 'nesterov', 'tensor_lr'):
        ga, gb = _StubGlue(SHAPES), _StubGlue(SHAPES)
        kw = dict(KW, nesterov=True) if case == 'nesterov' else dict(KW, lr=torch.tensor(0.01)) if case == 'tensor_lr' else KW
        if case == 'two_groups':
            mk = lambda ps: [dict(params=ps[:2]), dict(params=ps[2:], weight_decay=0.0)]
        else:
            mk = lambda ps: ps
        stock_params = [torch.nn.Parameter(q.detach().clone()) for q in gb.params]
        fused, stock = plugins.SGD(mk(ga.params), **kw), torch.optim.SGD(mk(stock_params), **kw)
        for it in range(2):
            ga.backward(it)
            for q, v in zip(stock_params, ga.grad_views):
                q.grad = v.clone()
            if case == 'none_grad':
                ga.params[1].grad = None
                stock_params[1].grad = None
            if case == 'foreign_grad':
                ga.params[2].grad = ga.params[2].grad.clone()
            fused.step()
            stock.step()
        assert ga.core.calls == 0, case
        for a, b in zip(ga.params, stock_params):
            assert torch.equal(a.detach(), b.detach()), case
    # and a fused step after stock steps picks their momentum buffers up
    ga, gb = _StubGlue(SHAPES), _StubGlue(SHAPES)
    stock_params = [torch.nn.Parameter(q.detach().clone()) for q in gb.params]
    fused, stock = plugins.SGD(ga.params, **KW), torch.optim.SGD(stock_params, **KW)
    for it in range(3):
        ga.backward(it)
        for q, v in zip(stock_params, ga.grad_views):
            q.grad = v.clone()
        if it == 0:
            ga.params[0].grad = ga.params[0].grad.clone()       # forces the stock path once
        fused.step()
        stock.step()
    assert ga.core.calls == 2
    for a, b in zip(ga.params, stock_params):
        assert torch.allclose(a, b, rtol=1e-6, atol=1e-7)


def test_optimizer_registry_builds_the_fused_sgd_from_the_reference_config():
    # configs/yunet_n.py:1  optimizer = dict(type='SGD', lr=0.01, momentum=0.9, weight_decay=0.0005)
    ga = _StubGlue(SHAPES)
    opt = plugins.OPTIMIZERS.build(dict(type='SGD', lr=0.01, momentum=0.9, weight_decay=0.0005),
                                   default_args=dict(params=ga.params))
    assert type(opt) is plugins.SGD and isinstance(opt, torch.optim.SGD)


def test_device_side_gt_packing_equals_the_host_packer_incl_empty_images():
    """``pack_gt_csr_device`` (torch ops, any device) == ``synthetic.pack_gt_csr`` (numpy), with images
    without faces given as (0, 4) / (0, 5, 3) or as bare empty tensors."""
    from libfacedetection.train_b200 import synthetic
    gb, gl, gk = synthetic.make_gt(6, 320, 3)
    gb[2], gk[2] = gb[2][:0], gk[2][:0]                       # an image without faces, shaped (0, 4) / (0, 5, 3)
    ref_gt, ref_offs = synthetic.pack_gt_csr(gb, gk)
    tb = [torch.from_numpy(x) for x in gb]
    tk = [torch.from_numpy(x) for x in gk]
    gt, offs = plugins.pack_gt_csr_device(tb, tk, torch.device('cpu'))
    assert torch.equal(offs, torch.from_numpy(ref_offs)) and offs.dtype == torch.int32
    assert torch.equal(gt, torch.from_numpy(ref_gt)) and gt.dtype == torch.float32
    tb[2], tk[2] = torch.zeros(0), torch.zeros(0)             # ... or as bare empty tensors
    gt2, offs2 = plugins.pack_gt_csr_device(tb, tk, torch.device('cpu'))
    assert torch.equal(gt2, gt) and torch.equal(offs2, offs)
    gt0, offs0 = plugins.pack_gt_csr_device([torch.zeros(0, 4)] * 3, [torch.zeros(0, 5, 3)] * 3, torch.device('cpu'))
    assert gt0.shape == (0, 19) and offs0.tolist() == [0, 0, 0, 0]
