"""CPU tests of the *fused formulation* the CUDA kernels implement, driven by the real execution
plan exported by libyunet_b200.so (host-only C-ABI calls, no GPU):

  forward : every unit stores its pre-BN output z + per-channel sum / sum^2; consumers apply
            BN+ReLU (+2x2 max-pool | + nearest-up2 add) while loading.
  backward: per unit, g = BN-backward(du) from the (sum du, sum du*zhat) statistics, recomputed
            pointwise output y, depthwise transposed stencil, dW/db, h = dy W1, routed through the
            ReLU mask / pool argmax / upsample children to the producers (overwrite or accumulate
            exactly as the plan's acc flags say).

Both are checked against the oracle (autograd of the reference restatement).  This pins the math,
the graph wiring, the parameter-bucket offsets and the accumulate flags before any kernel runs.
"""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import yunet_oracle as orc
from libfacedetection.train_b200 import _capi, synthetic
from conftest import GOLDEN

EPS = 1e-5


def _ctx(arch):
    a = orc.ARCH[arch]
    return _capi.Ctx(_capi.make_arch_cfg(a['stage_channels'], a['downsample_idx'], a['out_idx'],
                                         a['shared_stacked_convs'], a['feat_channels']))


def _flat_params(ctx, P):
    flat = torch.zeros(ctx.num_params, dtype=torch.float64)
    seen = 0
    for name, off, shape in ctx.params():
        t = P[name].double().reshape(-1)
        flat[off:off + t.numel()] = t
        seen += t.numel()
    assert seen == ctx.num_params == sum(v.numel() for v in P.values())
    return flat


def _nhwc(x):
    return x.permute(0, 2, 3, 1).contiguous()


class Emu:
    """float64 emulation of the fused plan (NHWC tensors)."""

    def __init__(self, ctx, flat, B, H, W):
        self.ctx, self.w, self.B, self.H, self.W = ctx, flat, B, H, W
        self.units = ctx.units()
        self.stem = ctx.units(include_stem=True)[0]
        self.z, self.stat = {}, {}

    def coef(self, t):
        z = self.z[t]
        mean = z.mean((0, 1, 2))
        var = (z * z).mean((0, 1, 2)) - mean * mean
        rstd = 1.0 / torch.sqrt(var + EPS)
        return mean, rstd

    def act(self, t, gamma_off, beta_off):
        C = self.z[t].shape[-1]
        mean, rstd = self.coef(t)
        g, b = self.w[gamma_off:gamma_off + C], self.w[beta_off:beta_off + C]
        u = (self.z[t] - mean) * rstd * g + b
        m = u > 0
        ov = getattr(self, 'mask_override', None)
        if ov is not None and t in ov:
            # rounding-fragile elements (|u| ~ 0): follow the fp32 kernels' ReLU decision
            m = torch.where(u.abs() < 1e-4, ov[t], m)
        self.mask = getattr(self, 'mask', {})
        self.mask[t] = m
        return u * m

    def bn_of_tensor(self):
        m = {self.stem.out: (self.stem.gamma, self.stem.beta)}
        for u in self.units:
            if u.has_bn:
                m[u.out] = (u.gamma, u.beta)
        return m

    def load(self, u):
        bn = self.bn_of_tensor()
        a = self.act(u.in_a, *bn[u.in_a])
        if u.mode == 1:
            B, H2, W2, C = a.shape
            a = a.reshape(B, H2 // 2, 2, W2 // 2, 2, C).amax((2, 4))
        elif u.mode == 2:
            b = self.act(u.in_b, *bn[u.in_b])
            a = a + b.repeat_interleave(2, 1).repeat_interleave(2, 2)
        return a

    def unit_params(self, u):
        w = self.w
        W1 = w[u.w1:u.w1 + u.cout * u.cin].reshape(u.cout, u.cin)
        b1 = w[u.b1:u.b1 + u.cout]
        W2 = w[u.w2:u.w2 + u.cout * 9].reshape(u.cout, 3, 3)
        b2 = w[u.b2:u.b2 + u.cout]
        return W1, b1, W2, b2

    @staticmethod
    def dw(y, W2, b2):
        # zero-pads y (NOT pw(0)+b1) — yunet_layer.py:32 conv2 has padding=1
        yp = F.pad(y, (0, 0, 1, 1, 1, 1))
        H, W = y.shape[1], y.shape[2]
        out = torch.zeros_like(y) + b2
        for ky in range(3):
            for kx in range(3):
                out = out + yp[:, ky:ky + H, kx:kx + W, :] * W2[:, ky, kx]
        return out

    def forward(self, img):
        s = self.stem
        wst = self.w[s.w1:s.w1 + 16 * 27].reshape(16, 3, 3, 3)
        bst = self.w[s.b1:s.b1 + 16]
        self.z[s.out] = _nhwc(F.conv2d(img.double(), wst, bst, stride=2, padding=1))
        preds = {}
        for u in self.units:
            a = self.load(u)
            W1, b1, W2, b2 = self.unit_params(u)
            y = a @ W1.t() + b1
            z = self.dw(y, W2, b2)
            self.z[u.out] = z
            if u.pred_level >= 0:
                preds[u.pred_level] = z
        B = img.shape[0]
        return torch.cat([preds[l].reshape(B, -1, 16) for l in range(3)], 1)

    def backward(self, img, d_preds):
        """d_preds (B,P,16) -> flat parameter gradient, following the kernel decomposition."""
        grad = torch.zeros_like(self.w)
        bn = self.bn_of_tensor()
        du, S1, S2 = {}, {}, {}
        self.du, self.S1, self.S2 = du, S1, S2
        B = img.shape[0]
        off = 0
        dlevel = {}
        for l, s in enumerate((8, 16, 32)):
            h, w = self.H // s, self.W // s
            dlevel[l] = d_preds[:, off:off + h * w].reshape(B, h, w, 16).double()
            off += h * w

        def g_of(t, D):
            C = D.shape[-1]
            mean, rstd = self.coef(t)
            gam = self.w[bn[t][0]:bn[t][0] + C]
            zh = (self.z[t] - mean) * rstd
            N = D.shape[0] * D.shape[1] * D.shape[2]
            return gam * rstd * (D - S1[t] / N - zh * S2[t] / N), zh

        def add_stats(t, contrib):
            mean, rstd = self.coef(t)
            zh = (self.z[t] - mean) * rstd
            S1[t] = S1.get(t, 0) + contrib.sum((0, 1, 2))
            S2[t] = S2.get(t, 0) + (contrib * zh).sum((0, 1, 2))

        for u in reversed(self.units):
            W1, b1, W2, b2 = self.unit_params(u)
            if u.has_bn:
                g, _ = g_of(u.out, du[u.out])
                C = u.cout
                grad[u.gamma:u.gamma + C] = S2[u.out]
                grad[u.beta:u.beta + C] = S1[u.out]
            else:
                g = dlevel[u.pred_level]
            a = self.load(u)
            y = a @ W1.t() + b1
            H, W = y.shape[1], y.shape[2]
            gp = F.pad(g, (0, 0, 1, 1, 1, 1))
            dy = torch.zeros_like(y)
            gW2 = torch.zeros_like(W2)
            for ky in range(3):
                for kx in range(3):
                    # window element [2-ky][2-kx] of the halo tile == g[q - (ky-1, kx-1)]
                    gs = gp[:, 2 - ky:2 - ky + H, 2 - kx:2 - kx + W, :]
                    dy = dy + gs * W2[:, ky, kx]
                    gW2[:, ky, kx] = (y * gs).sum((0, 1, 2))
            grad[u.w2:u.w2 + u.cout * 9] = gW2.reshape(-1)
            grad[u.b2:u.b2 + u.cout] = g.sum((0, 1, 2))
            grad[u.b1:u.b1 + u.cout] = dy.sum((0, 1, 2))
            grad[u.w1:u.w1 + u.cout * u.cin] = torch.einsum('bhwo,bhwi->oi', dy, a).reshape(-1)
            h = dy @ W1
            # ---- prologue backward
            ua = self.act(u.in_a, *bn[u.in_a])
            ma = self.mask[u.in_a]
            if u.mode in (0, 2):
                contrib = h * ma
            else:
                Bn, H2, W2_, C = ua.shape
                win = ua.reshape(Bn, H2 // 2, 2, W2_ // 2, 2, C).permute(0, 1, 3, 5, 2, 4).reshape(
                    Bn, H2 // 2, W2_ // 2, C, 4)
                best = win.argmax(-1)    # first maximum (row-major window order)
                # torch.argmax returns the first max on CPU for ties? make it explicit:
                mx = win.amax(-1, keepdim=True)
                first = (win == mx).double()
                first = first * (first.cumsum(-1) == 1)
                mwin = ma.reshape(Bn, H2 // 2, 2, W2_ // 2, 2, C).permute(0, 1, 3, 5, 2, 4).reshape(
                    Bn, H2 // 2, W2_ // 2, C, 4).double()
                hv = (h * ((mwin * first).sum(-1) > 0)).unsqueeze(-1) * first
                contrib = hv.reshape(Bn, H2 // 2, W2_ // 2, C, 2, 2).permute(0, 1, 4, 2, 5, 3).reshape(
                    Bn, H2, W2_, C)
            if u.acc_a:
                du[u.in_a] = du[u.in_a] + contrib
            else:
                assert u.in_a not in du, 'overwrite of an already written gradient'
                du[u.in_a] = contrib
            add_stats(u.in_a, contrib)
            if u.mode == 2:
                self.act(u.in_b, *bn[u.in_b])
                Bn, Hh, Wh, C = h.shape
                hs = h.reshape(Bn, Hh // 2, 2, Wh // 2, 2, C).sum((2, 4))
                cb = hs * self.mask[u.in_b]
                if u.acc_b:
                    du[u.in_b] = du[u.in_b] + cb
                else:
                    assert u.in_b not in du
                    du[u.in_b] = cb
                add_stats(u.in_b, cb)
        # ---- stem
        s = self.stem
        g, _ = g_of(s.out, du[s.out])
        grad[s.gamma:s.gamma + 16] = S2[s.out]
        grad[s.beta:s.beta + 16] = S1[s.out]
        gn = g.permute(0, 3, 1, 2)
        imgp = F.pad(img.double(), (1, 1, 1, 1))
        Ho, Wo = gn.shape[2], gn.shape[3]
        gw = torch.zeros(16, 3, 3, 3, dtype=torch.float64)
        for ky in range(3):
            for kx in range(3):
                patch = imgp[:, :, ky:ky + 2 * Ho:2, kx:kx + 2 * Wo:2]
                gw[:, :, ky, kx] = torch.einsum('bohw,bchw->oc', gn, patch)
        grad[s.w1:s.w1 + 432] = gw.reshape(-1)
        grad[s.b1:s.b1 + 16] = g.sum((0, 1, 2))
        return grad


def _weights(arch):
    d = np.load(os.path.join(GOLDEN, f'weights_{arch}.npz'))
    return orc.split_state_dict({k: torch.from_numpy(d[k]) for k in d.files})


def _rel(a, b):
    return float((a - b).abs().max() / (b.abs().max() + 1e-30))


@pytest.mark.parametrize('arch', ['yunet_n', 'yunet_s'])
def test_plan_bucket_covers_state_dict(arch):
    ctx = _ctx(arch)
    P, _ = orc.init_params(arch)
    names = {n: (o, s) for n, o, s in ctx.params()}
    assert set(names) == set(P)
    for k, v in P.items():
        assert tuple(v.shape) == names[k][1], k
    # ranges tile the bucket exactly once
    spans = sorted((o, o + int(np.prod(s))) for o, s in names.values())
    assert spans[0][0] == 0 and spans[-1][1] == ctx.num_params
    assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
    assert ctx.num_params == {'yunet_n': 75856, 'yunet_s': 54608}[arch]
    assert ctx.num_priors(320, 320) == 2100 and ctx.num_priors(640, 640) == 8400
    bn = ctx.bns()
    _, Bf = orc.init_params(arch)
    assert {n + '.running_mean' for n, _, _ in bn} == {k for k in Bf if k.endswith('running_mean')}


@pytest.mark.parametrize('arch', ['yunet_n', 'yunet_s'])
def test_fused_forward_and_backward_match_oracle(arch):
    torch.manual_seed(0)
    ctx = _ctx(arch)
    P, Bf = _weights(arch)
    flat = _flat_params(ctx, P)
    B, S = 2, 64
    img = torch.from_numpy(synthetic.make_images(B, S, 3))
    emu = Emu(ctx, flat, B, S, S)
    preds = emu.forward(img)
    # oracle, float64, train-mode BN
    P64 = {k: v.double().requires_grad_(True) for k, v in P.items()}
    Bf64 = {k: (v.double() if v.dtype.is_floating_point else v.clone()) for k, v in Bf.items()}
    outs = orc.model_forward(img.double(), P64, Bf64, arch, training=True)
    f = orc.flatten_preds(*outs)
    ref = torch.cat([f[0], f[1], f[2].unsqueeze(-1), f[3]], -1)
    assert _rel(preds, ref.detach()) < 1e-9
    d_preds = torch.randn(ref.shape, dtype=torch.float64)
    (ref * d_preds).sum().backward()
    grad = emu.backward(img, d_preds)
    for name, off, shape in ctx.params():
        n = int(np.prod(shape))
        mine = grad[off:off + n].reshape(shape)
        r = P64[name].grad
        scale = max(float(r.abs().max()), 1e-6 * float(grad.abs().max()))
        assert float((mine - r).abs().max()) <= 1e-7 * scale + 1e-9 * float(grad.abs().max()), name
