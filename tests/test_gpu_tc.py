"""GPU: the tcgen05 / TMEM / TMA unit kernel (3xTF32) against the exact-fp32 CUDA-core path of the
same library and against the oracle."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from conftest import GOLDEN  # noqa: E402
from oracle import yunet_oracle as orc  # noqa: E402


def _engine(arch):
    from libfacedetection.train_b200 import YuNetEngine
    eng = YuNetEngine(arch)
    d = np.load(os.path.join(GOLDEN, f'weights_{arch}.npz'))
    eng.load_state_dict({k: torch.from_numpy(d[k]) for k in d.files})
    return eng


def _rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).abs().max() / (b.abs().max() + 1e-30))


@pytest.mark.parametrize('arch', ['yunet_n', 'yunet_s'])
@pytest.mark.parametrize('train', [False, True])
@pytest.mark.parametrize('shape', [(3, 96, 160), (2, 320, 320)])
@pytest.mark.parametrize('path', ['tc', 'ws'])
def test_tc_forward_equals_fp32_path(arch, train, shape, path):
    """`tc`: the per-tile tcgen05 kernel (unit_fwd_tc.cu); `ws`: the warp-specialised streaming
    kernel (unit_fwd_ws.cu).  Both against the exact-fp32 CUDA-core path."""
    B, H, W = shape
    eng = _engine(arch)
    rng = np.random.default_rng(21)
    img = torch.from_numpy(rng.random((B, 3, H, W), dtype=np.float32) * 255).cuda()
    eng.set_option('tc_forward', 0)
    eng.set_option('ws_forward', 0)
    ref = eng.forward(img, train=train).clone()
    ref_units = [eng.read_activation(i, B, H, W, train=train).clone()
                 for i, u in enumerate(eng.ctx.units()) if u.pred_level < 0]
    eng.set_option('tc_forward', 1)
    eng.set_option('ws_forward', 1 if path == 'ws' else 0)
    out = eng.forward(img, train=train)
    torch.cuda.synchronize()
    flags = eng.status_flags(B, H, W, train)
    assert int(flags.abs().sum()) == 0, f'tensor-core kernel reported {flags[:4].tolist()}'
    k = 0
    worst = 0.0
    for i, u in enumerate(eng.ctx.units()):
        if u.pred_level >= 0:
            continue
        e = _rel(eng.read_activation(i, B, H, W, train=train), ref_units[k])
        k += 1
        worst = max(worst, e)
        assert e < 1e-4, f'unit {i} {u.name.decode()}: {e:.3e}'
    e = _rel(out, ref)
    print(f'{arch} train={train} {shape} {path}: tc vs fp32 preds {e:.3e}, worst unit {worst:.3e}')
    assert e < 1e-4


def test_tc_forward_matches_reference_golden():
    g = np.load(os.path.join(GOLDEN, 'forward_yunet_n_640.npz'))
    eng = _engine('yunet_n')
    eng.set_option('tc_forward', 1)
    eng.set_option('ws_forward', 0)
    torch.manual_seed(0)
    img = (torch.rand(1, 3, 640, 640) * 255).cuda()
    preds = eng.forward(img, train=False)
    e = _rel(preds, torch.from_numpy(g['preds']))
    print(f'tc forward vs reference golden (640): {e:.3e}')
    assert e < 1e-3


@pytest.mark.parametrize('arch', ['yunet_n', 'yunet_s'])
@pytest.mark.parametrize('shape', [(3, 96, 160), (2, 320, 320), (2, 640, 320)])
@pytest.mark.parametrize('path', ['tc', 'st'])
def test_tc_backward_equals_fp32_path(arch, shape, path):
    """Same forward, same upstream gradient: parameter gradients of the tcgen05 backward kernels
    (`tc`: per-tile unit_bwd_tc.cu, `st`: strip-streaming unit_bwd_st.cu) vs the exact-fp32
    CUDA-core backward (identical ReLU/pool decisions by construction)."""
    B, H, W = shape
    eng = _engine(arch)
    rng = np.random.default_rng(33)
    img = torch.from_numpy(rng.random((B, 3, H, W), dtype=np.float32) * 255).cuda()
    preds = eng.forward(img, train=True)
    d_preds = torch.from_numpy(rng.standard_normal(tuple(preds.shape)).astype(np.float32)).cuda()
    eng.set_option('tc_backward', 0)
    eng.set_option('st_backward', 0)
    g0 = eng.backward(img, d_preds).clone()
    eng.set_option('tc_backward', 1)
    eng.set_option('st_backward', 2 if path == 'st' else 0)      # 2: the strip kernel wherever it applies
    g1 = eng.backward(img, d_preds).clone()
    torch.cuda.synchronize()
    flags = eng.status_flags(B, H, W, True)
    assert int(flags.abs().sum()) == 0, f'tensor-core kernel reported {flags[:4].tolist()}'
    gmax = float(g0.abs().max())
    worst = 0.0
    for name, off, shp in eng.ctx.params():
        n = int(np.prod(shp))
        a, b = g1[off:off + n], g0[off:off + n]
        scale = float(b.abs().max())
        err = float((a - b).abs().max())
        worst = max(worst, err / (scale + 1e-4 * gmax))
        assert err <= 1e-4 * scale + 1e-5 * gmax, (name, err, scale)
    print(f'{arch} {shape} {path}: tc backward vs fp32 backward, worst normalised grad err {worst:.3e}')
