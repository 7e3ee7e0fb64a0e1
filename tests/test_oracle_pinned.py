"""CPU tests: the portable oracle (oracle/yunet_oracle.py) against (a) the committed fixtures
generated from the unmodified reference (tests/golden, oracle/gen_golden.py) and (b) the reference
itself when /root/reference is present (development container only)."""
import os

import numpy as np
import pytest
import torch

from oracle import yunet_oracle as orc
from oracle import ref_loader
from libfacedetection.train_b200 import synthetic
from conftest import GOLDEN


def _weights(arch):
    d = np.load(os.path.join(GOLDEN, f'weights_{arch}.npz'))
    return orc.split_state_dict({k: torch.from_numpy(d[k]) for k in d.files})


def _rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / (np.abs(b).max() + 1e-30))


def _split_preds(preds, size, arch='yunet_n'):
    """(B,P,16) flattened -> per-level NCHW lists, inverse of flatten_preds."""
    B = preds.shape[0]
    outs = ([], [], [], [])
    off = 0
    for s in (8, 16, 32):
        h = size // s
        sl = preds[:, off:off + h * h].reshape(B, h, h, 16).permute(0, 3, 1, 2)
        off += h * h
        outs[0].append(sl[:, 0:1]); outs[1].append(sl[:, 1:5])
        outs[2].append(sl[:, 5:6]); outs[3].append(sl[:, 6:16])
    return outs


@pytest.mark.parametrize('arch,size', [('yunet_n', 320), ('yunet_s', 320), ('yunet_n', 640)])
def test_forward_matches_golden(arch, size):
    g = np.load(os.path.join(GOLDEN, f'forward_{arch}_{size}.npz'))
    P, Bf = _weights(arch)
    torch.manual_seed(0)
    img = torch.rand(1, 3, size, size) * 255
    with torch.no_grad():
        outs = orc.model_forward(img, P, Bf, arch, training=False)
    f = orc.flatten_preds(*outs)
    preds = torch.cat([f[0], f[1], f[2].unsqueeze(-1), f[3]], -1).numpy()
    assert _rel(preds, g['preds']) < 1e-5
    dets = orc.get_bboxes(*outs)[0][0].numpy()
    assert dets.shape == g['dets'].shape
    if dets.size:
        assert _rel(dets, g['dets']) < 1e-5


@pytest.mark.parametrize('arch,seed', [('yunet_n', 0), ('yunet_s', 1)])
def test_train_step_matches_golden(arch, seed):
    g = np.load(os.path.join(GOLDEN, f'train_{arch}_b4.npz'))
    B, size = int(g['B']), int(g['size'])
    P, Bf = _weights(arch)
    img = torch.from_numpy(synthetic.make_images(B, size, seed))
    gb, gl, gk = synthetic.make_gt(B, size, seed)
    losses, grads, assign, outs = orc.train_forward_backward(
        img, P, Bf, arch, [torch.from_numpy(x) for x in gb], [torch.from_numpy(x) for x in gl],
        [torch.from_numpy(x) for x in gk])
    ref_l = g['losses']
    for i, k in enumerate(('loss_cls', 'loss_bbox', 'loss_obj', 'loss_kps')):
        assert abs(losses[k] - ref_l[i]) <= 1e-4 * max(1.0, abs(ref_l[i])), k
    assert np.array_equal(assign['assigned_gt_inds'].numpy(), g['assigned_gt_inds'])
    np.testing.assert_allclose(assign['max_overlaps'].numpy(), g['max_overlaps'], rtol=1e-5)
    for k in grads:
        assert _rel(grads[k].numpy(), g['grad/' + k]) < 1e-4, k
    mom = {}
    orc.sgd_step(P, grads, mom)
    for k in P:
        assert _rel(P[k].numpy(), g['after/' + k]) < 1e-5, k
    for k in Bf:
        if Bf[k].dtype.is_floating_point:
            assert _rel(Bf[k].numpy(), g['after/' + k]) < 1e-5, k


def test_nms_matches_golden_and_torchvision():
    g = np.load(os.path.join(GOLDEN, 'nms_synth_640.npz'))
    preds = torch.from_numpy(g['preds'])
    outs = _split_preds(preds, int(g['size']))
    mine = orc.get_bboxes(*outs)
    tv = orc.get_bboxes(*outs, use_torchvision=True)
    for b in range(preds.shape[0]):
        assert np.array_equal(mine[b][0].numpy(), g[f'dets{b}'])
        assert torch.equal(mine[b][0], tv[b][0])


def test_docstring_vectors():
    # mmdet/core/bbox/iou_calculators/iou2d_calculator.py:168-189, losses/utils.py:72-90
    g = np.load(os.path.join(GOLDEN, 'iou_docstring.npz'))
    iou = orc.bbox_overlaps(torch.from_numpy(g['b1']), torch.from_numpy(g['b2']))
    assert np.array_equal(iou.numpy(), g['iou'])
    assert tuple(orc.bbox_overlaps(torch.empty(0, 4), torch.FloatTensor([[0, 0, 10, 9]])).shape) == (0, 1)


def test_priors_known_values():
    # SURVEY a7 probe values of MlvlPointGenerator.grid_priors
    p = torch.cat(orc.grid_priors([(40, 40), (20, 20), (10, 10)], (8, 16, 32)))
    assert p.shape == (2100, 4)
    assert p[1].tolist() == [8, 0, 8, 8] and p[40].tolist() == [0, 8, 8, 8]
    assert p[1600].tolist() == [0, 0, 16, 16] and p[2099].tolist() == [288, 288, 32, 32]


def test_param_count():
    for arch, n in (('yunet_n', 75856), ('yunet_s', 54608)):
        P, _ = orc.init_params(arch)
        assert sum(v.numel() for v in P.values()) == n
        W, _ = _weights(arch)
        assert {k: tuple(v.shape) for k, v in P.items()} == {k: tuple(v.shape) for k, v in W.items()}


@pytest.mark.reference
@pytest.mark.skipif(not ref_loader.reference_available(), reason='needs /root/reference')
def test_live_reference_forward_and_init():
    model, cfg = ref_loader.build_reference_model('yunet_s', pretrained=True)
    model.eval()
    P, Bf = orc.split_state_dict(model.state_dict())
    img = torch.rand(2, 3, 320, 320) * 255
    with torch.no_grad():
        ref = model.feature_test(img)
        mine = orc.model_forward(img, P, Bf, 'yunet_s', False)
    for la, lb in zip(mine, ref):
        for a, b in zip(la, lb):
            assert torch.equal(a, b)
