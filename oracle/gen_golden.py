"""TEST INFRASTRUCTURE — generate the committed fixtures in ``tests/golden`` from the UNMODIFIED
reference (``/root/reference`` through ``oracle/ref_loader.py``) and, in the same run, check the
portable restatement ``oracle/yunet_oracle.py`` against it.

    python oracle/gen_golden.py            # writes tests/golden/*.npz, asserts oracle == reference

Only runs where ``/root/reference`` exists (development container).  Inputs are seeded numpy
generators (``libfacedetection/train_b200/synthetic.py``) so the fixtures carry only the outputs
plus the trained weights the parity cases need (``weights/yunet_{n,s}.pth`` state_dicts).
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import ref_loader, yunet_oracle as orc  # noqa: E402
from libfacedetection.train_b200 import synthetic  # noqa: E402

GOLD = os.path.join(ROOT, 'tests', 'golden')


def rel_err(a, b):
    a, b = a.detach().double(), b.detach().double()
    return float((a - b).abs().max() / (b.abs().max() + 1e-30))


def to_np(d):
    return {k: v.detach().cpu().numpy() for k, v in d.items()}


def train_case(arch, B, seed, size=320):
    """One reference train step (forward_train -> _parse_losses -> backward -> SGD)."""
    model, cfg = ref_loader.build_reference_model(arch, pretrained=True)
    model.train()
    img = torch.from_numpy(synthetic.make_images(B, size, seed))
    gb, gl, gk = synthetic.make_gt(B, size, seed)
    gt_b = [torch.from_numpy(x) for x in gb]
    gt_l = [torch.from_numpy(x) for x in gl]
    gt_k = [torch.from_numpy(x) for x in gk]
    P0, Bf0 = orc.split_state_dict(model.state_dict())

    # ---- reference: hook the assigner to record per-image results
    rec = []
    assigner = model.bbox_head.assigner
    orig_assign = assigner.assign

    def spy(*a, **k):
        r = orig_assign(*a, **k)
        rec.append((r.gt_inds.clone(), r.max_overlaps.clone()))
        return r

    assigner.assign = spy
    opt = torch.optim.SGD(model.parameters(), lr=0.01, momentum=0.9, weight_decay=0.0005)
    opt.zero_grad()
    losses = model.forward_train(img, [{}] * B, gt_b, gt_l, gt_k, None)
    loss, log_vars = model._parse_losses(losses)
    loss.backward()
    ref_grads = {k: p.grad.detach().clone() for k, p in model.named_parameters()}
    with torch.no_grad():
        model.eval()
        # raw head maps of the SAME (train-mode BN) forward are not retrievable after the fact;
        # recompute them in train mode below through hooks instead
        model.train()
    opt.step()
    ref_after = {k: v.detach().clone() for k, v in model.state_dict().items()}
    ref_assigned = torch.stack([r[0] for r in rec])
    ref_maxov = torch.stack([r[1] for r in rec])

    # ---- oracle on the same inputs
    P, Bf = {k: v.clone() for k, v in P0.items()}, {k: v.clone() for k, v in Bf0.items()}
    o_losses, o_grads, o_assign, o_outs = orc.train_forward_backward(img, P, Bf, arch, gt_b, gt_l,
                                                                     gt_k)
    for k in ('loss_cls', 'loss_bbox', 'loss_obj', 'loss_kps'):
        r = float(losses[k])
        assert abs(o_losses[k] - r) <= 1e-5 * max(1.0, abs(r)), (k, o_losses[k], r)
    assert torch.equal(o_assign['assigned_gt_inds'], ref_assigned), 'assignment mismatch'
    assert torch.equal(o_assign['max_overlaps'], ref_maxov)
    worst = max(rel_err(o_grads[k], ref_grads[k]) for k in ref_grads)
    assert worst < 1e-4, worst
    mom = {}
    orc.sgd_step(P, o_grads, mom)
    worst_p = max(rel_err(P[k], ref_after[k]) for k in P)
    assert worst_p < 1e-5, worst_p
    for k in Bf:
        if Bf[k].dtype.is_floating_point:
            assert rel_err(Bf[k], ref_after[k]) < 1e-5, k
        else:
            assert int(Bf[k]) == int(ref_after[k]), k
    print(f'[train {arch} B={B}] oracle==reference: losses ok, assignment exact, '
          f'grad rel {worst:.2e}, param-after rel {worst_p:.2e}')

    f_cls, f_bbox, f_obj, f_kps = orc.flatten_preds(*o_outs)
    preds = torch.cat([f_cls, f_bbox, f_obj.unsqueeze(-1), f_kps], -1)  # (B,P,16)
    out = dict(arch=arch, B=B, seed=seed, size=size,
               losses=np.array([float(losses[k]) for k in
                                ('loss_cls', 'loss_bbox', 'loss_obj', 'loss_kps')], np.float64),
               assigned_gt_inds=ref_assigned.numpy().astype(np.int32),
               max_overlaps=ref_maxov.numpy(),
               preds=preds.detach().numpy())
    for k, v in ref_grads.items():
        out['grad/' + k] = v.numpy()
    for k, v in ref_after.items():
        out['after/' + k] = v.numpy()
    np.savez_compressed(os.path.join(GOLD, f'train_{arch}_b{B}.npz'), **out)


def forward_case(arch, size=320):
    """Config 1: single image eval forward (feature_test) + OpenCV-DNN ONNX cross-check."""
    model, cfg = ref_loader.build_reference_model(arch, pretrained=True)
    model.eval()
    torch.manual_seed(0)
    img = torch.rand(1, 3, size, size) * 255
    with torch.no_grad():
        outs = model.feature_test(img)
    P, Bf = orc.split_state_dict(model.state_dict())
    with torch.no_grad():
        o = orc.model_forward(img, P, Bf, arch, training=False)
    worst = max(rel_err(a, b) for la, lb in zip(o, outs) for a, b in zip(la, lb))
    assert worst < 1e-5, worst
    f_cls, f_bbox, f_obj, f_kps = orc.flatten_preds(*outs)
    preds = torch.cat([f_cls, f_bbox, f_obj.unsqueeze(-1), f_kps], -1)
    msg = f'[forward {arch} {size}] oracle==reference rel {worst:.2e}'
    onnx_path = os.path.join(ref_loader.REFERENCE_ROOT, 'onnx', f'{arch}_{size}_{size}.onnx')
    if os.path.exists(onnx_path):
        import cv2
        net = cv2.dnn.readNetFromONNX(onnx_path)
        net.setInput(img.numpy())
        names = net.getUnconnectedOutLayersNames()
        res = dict(zip(names, net.forward(names)))
        w = 0.0
        off = 0
        for lvl, s in enumerate((8, 16, 32)):
            n = (size // s) ** 2
            sl = preds[0, off:off + n]
            off += n
            pairs = [(torch.sigmoid(sl[:, 0:1]), res[f'cls_{s}']), (sl[:, 1:5], res[f'bbox_{s}']),
                     (torch.sigmoid(sl[:, 5:6]), res[f'obj_{s}']), (sl[:, 6:16], res[f'kps_{s}'])]
            for a, b in pairs:
                w = max(w, rel_err(a, torch.from_numpy(b).reshape(a.shape)))
        assert w < 1e-4, w
        msg += f'; reference==OpenCV-DNN({os.path.basename(onnx_path)}) rel {w:.2e}'
    print(msg)
    # detections of simple_test (decode + NMS) on the same image
    with torch.no_grad():
        dets = model.bbox_head.get_bboxes(*outs, img_metas=[{}], rescale=False)
    np.savez_compressed(os.path.join(GOLD, f'forward_{arch}_{size}.npz'), arch=arch, size=size,
                        preds=preds.numpy(), dets=dets[0][0].numpy())


def nms_case(seed=0, B=2, size=640):
    """Decode + NMS on synthetic clustered logits (SURVEY §8d config-4 variant): reference
    get_bboxes (batched_nms stub -> torchvision.ops.nms) vs the oracle's greedy restatement."""
    model, cfg = ref_loader.build_reference_model('yunet_n', pretrained=False)
    rng = np.random.default_rng(seed)
    lv = [size // s for s in (8, 16, 32)]
    cls, bbox, obj, kps = [], [], [], []
    centers = rng.uniform(0.1, 0.9, (B, 20, 2)) * size
    for h, s in zip(lv, (8, 16, 32)):
        yy, xx = np.meshgrid(np.arange(h) * s, np.arange(h) * s, indexing='ij')
        d = np.min(np.hypot(xx[None, None] - centers[:, :, 0, None, None],
                            yy[None, None] - centers[:, :, 1, None, None]), 1)   # (B,h,h)
        logit = 3.0 - d / (0.45 * s) + rng.normal(0, 0.7, d.shape)
        cls.append(torch.from_numpy(logit[:, None].astype(np.float32)))
        obj.append(torch.from_numpy((logit[:, None] + rng.normal(0, 0.5, d.shape)[:, None])
                                    .astype(np.float32)))
        bbox.append(torch.from_numpy(rng.normal(0, 0.6, (B, 4, h, h)).astype(np.float32)))
        kps.append(torch.from_numpy(rng.normal(0, 1.0, (B, 10, h, h)).astype(np.float32)))
    with torch.no_grad():
        ref = model.bbox_head.get_bboxes(cls, bbox, obj, kps, img_metas=[{}] * B, rescale=False)
    mine = orc.get_bboxes(cls, bbox, obj, kps)
    out = dict(seed=seed, B=B, size=size)
    for b in range(B):
        assert ref[b][0].shape == mine[b][0].shape, (ref[b][0].shape, mine[b][0].shape)
        assert torch.equal(ref[b][0], mine[b][0]), 'NMS dets differ'
        out[f'dets{b}'] = ref[b][0].numpy()
    f_cls, f_bbox, f_obj, f_kps = orc.flatten_preds(cls, bbox, obj, kps)
    out['preds'] = torch.cat([f_cls, f_bbox, f_obj.unsqueeze(-1), f_kps], -1).numpy()
    print(f'[nms B={B} {size}] oracle greedy NMS == reference (torchvision) bit-exact; '
          f'dets/img {[int(r[0].shape[0]) for r in ref]}')
    np.savez_compressed(os.path.join(GOLD, f'nms_synth_{size}.npz'), **out)


def weights():
    for arch in ('yunet_n', 'yunet_s'):
        ck = torch.load(os.path.join(ref_loader.REFERENCE_ROOT, 'weights', f'{arch}.pth'),
                        map_location='cpu', weights_only=False)
        np.savez_compressed(os.path.join(GOLD, f'weights_{arch}.npz'),
                            **{k: v.numpy() for k, v in ck['state_dict'].items()})


def docstring_vectors():
    """Known-answer vectors the reference carries in docstrings:
    losses/utils.py:72-90 and iou2d_calculator.py:168-189."""
    ref_loader.install()
    from mmdet.core.bbox.iou_calculators import bbox_overlaps
    b1 = torch.FloatTensor([[0, 0, 10, 10], [10, 10, 20, 20], [32, 32, 38, 42]])
    b2 = torch.FloatTensor([[0, 0, 10, 20], [0, 10, 10, 19], [10, 10, 20, 20]])
    ref = bbox_overlaps(b1, b2)
    assert torch.equal(ref, orc.bbox_overlaps(b1, b2))
    np.savez(os.path.join(GOLD, 'iou_docstring.npz'), b1=b1.numpy(), b2=b2.numpy(),
             iou=ref.numpy())
    print('[docstring] bbox_overlaps vector ok')


if __name__ == '__main__':
    os.makedirs(GOLD, exist_ok=True)
    torch.set_num_threads(8)
    weights()
    docstring_vectors()
    for arch in ('yunet_n', 'yunet_s'):
        forward_case(arch, 320)
    forward_case('yunet_n', 640)
    train_case('yunet_n', 4, seed=0)
    train_case('yunet_s', 4, seed=1)
    nms_case()
    print('golden fixtures written to', GOLD)
