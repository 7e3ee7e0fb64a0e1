"""GPU: every intermediate of the fused backward (activation gradients, BatchNorm-backward
statistics, per-unit parameter gradients) against the float64 emulation of the same plan
(tests/test_formulation.py::Emu, itself pinned to the oracle's autograd on CPU).  Ragged 96x160
input so partially filled tiles occur at every level; random upstream gradient."""
import ctypes as C

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from test_formulation import Emu, _flat_params, _weights  # noqa: E402
from libfacedetection.train_b200 import _capi  # noqa: E402


def _ws_tensor(eng, B, H, W, tid, kind, shape, dtype=torch.float32):
    off = _capi.lib.yunet_ws_offset(eng.h, B, H, W, 1, tid, kind)
    assert off >= 0
    ws = eng.workspace(B, H, W, True)
    n = int(np.prod(shape)) * (8 if dtype == torch.float64 else 4)
    return ws[off:off + n].view(dtype).view(shape)


@pytest.mark.parametrize('arch', ['yunet_n', 'yunet_s'])
def test_backward_intermediates(arch):
    from libfacedetection.train_b200 import YuNetEngine
    eng = YuNetEngine(arch)
    P, Bf = _weights(arch)
    eng.load_state_dict({**P, **Bf})
    B, H, W = 2, 96, 160
    rng = np.random.default_rng(9)
    img = torch.from_numpy(rng.random((B, 3, H, W), dtype=np.float32) * 255)
    flat = _flat_params(eng.ctx, P)
    emu = Emu(eng.ctx, flat, B, H, W)
    preds = eng.forward(img.cuda(), train=True)
    # ReLU decisions of the fp32 kernels, used by the float64 emulation only where |u| < 1e-4
    # (there rounding decides the branch; a single flip moves all upstream gradients by ~1e-2)
    units_all = eng.ctx.units()
    ov = {eng.ctx.units(include_stem=True)[0].out:
          eng.read_activation(-1, B, H, W, train=True).permute(0, 2, 3, 1).cpu() > 0}
    for i, u in enumerate(units_all):
        if u.has_bn:
            ov[u.out] = eng.read_activation(i, B, H, W, train=True).permute(0, 2, 3, 1).cpu() > 0
    emu.mask_override = ov
    ref_preds = emu.forward(img)
    assert float((preds.cpu().double() - ref_preds).abs().max() / ref_preds.abs().max()) < 1e-4
    d_preds = torch.from_numpy(rng.standard_normal(tuple(ref_preds.shape)).astype(np.float32))
    ref_grad = emu.backward(img, d_preds)
    eng.backward(img.cuda(), d_preds.cuda())
    torch.cuda.synchronize()
    units = eng.ctx.units()
    stem = eng.ctx.units(include_stem=True)[0]
    nbn = eng.ctx.num_bn_channels
    stats = _ws_tensor(eng, B, H, W, 0, 2, (4, nbn), torch.float64).cpu()
    bn_off = {name: off for name, off, ch in eng.ctx.bns()}
    rows = []
    bad_dump = []
    # tensors in backward order: what each unit's backward WROTE (du of its inputs)
    seen = set()
    order = []
    for u in reversed(units):
        for t in (u.in_a, u.in_b):
            if t >= 0 and t not in seen:
                seen.add(t)
                order.append((t, u.name.decode()))
    bn_of = {stem.out: 'backbone.model0.bn1'}
    for u in units:
        if u.has_bn:
            bn_of[u.out] = u.name.decode() + '.bn'
    for t, writer in order:
        ref = emu.du[t]
        mine = _ws_tensor(eng, B, H, W, t, 1, tuple(ref.shape)).cpu().double()
        e = float((mine - ref).abs().max() / (ref.abs().max() + 1e-30))
        o = bn_off[bn_of[t]]
        Cc = ref.shape[-1]
        s1 = float((stats[2, o:o + Cc] - emu.S1[t]).abs().max() / (emu.S1[t].abs().max() + 1e-30))
        s2 = float((stats[3, o:o + Cc] - emu.S2[t]).abs().max() / (emu.S2[t].abs().max() + 1e-30))
        rows.append((f'du[t{t}] first written by {writer}', e, s1, s2))
        if e > 1e-4 and not bad_dump:
            bad_dump.append(t)
            d = (mine - ref).abs()
            scale = float(ref.abs().max())
            print(f'--- first bad tensor t{t} ({writer}): shape {tuple(ref.shape)}')
            pc = d.amax((0, 1, 2)) / scale
            print('    err by channel:', ' '.join(f'{float(x):.1e}' for x in pc))
            pm = d.amax(-1) / scale          # (B,H,W)
            for b in range(pm.shape[0]):
                print(f'    image {b}: err by pixel (rows), x->')
                for y in range(pm.shape[1]):
                    print('      ' + ' '.join('#' if float(v) > 1e-3 else ('+' if float(v) > 1e-5 else '.') for v in pm[b, y]))
            # second launch: determinism
            eng.backward(img.cuda(), d_preds.cuda())
            torch.cuda.synchronize()
            again = _ws_tensor(eng, B, H, W, t, 1, tuple(ref.shape)).cpu().double()
            print('    identical on a second backward:', bool(torch.equal(again, mine)))
    print('\nactivation gradients (rel err), sum(du), sum(du*zhat):')
    for r in rows:
        print(f'   {r[0]:62s} {r[1]:.2e} {r[2]:.2e} {r[3]:.2e}')
    g = eng.grads.cpu().double()
    gmax = float(ref_grad.abs().max())
    prow = []
    for name, off, shape in eng.ctx.params():
        n = int(np.prod(shape))
        r = ref_grad[off:off + n]
        e = float((g[off:off + n] - r).abs().max() / (r.abs().max() + 1e-5 * gmax))
        prow.append((e, name))
    prow.sort(reverse=True)
    print('parameter gradients, worst 10 (normalised err):')
    for e, n in prow[:10]:
        print(f'   {e:.2e} {n}')
    assert max(r[1] for r in rows) < 1e-4 and max(r[2] for r in rows) < 1e-3
    # tensors with an identically-zero true gradient (conv biases in front of a train-mode BN) hold
    # fp32 rounding residue only: bounded relative to the whole gradient instead
    for name, off, shape in eng.ctx.params():
        n = int(np.prod(shape))
        r = ref_grad[off:off + n]
        err = float((g[off:off + n] - r).abs().max())
        scale = float(r.abs().max())
        if scale < 1e-6 * gmax:
            assert err < 1e-4 * gmax, (name, err)
        else:
            assert err < 1e-3 * scale + 1e-5 * gmax, (name, err, scale)
