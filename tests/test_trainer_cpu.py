"""CPU tests of the trainer's host logic: LR schedule of configs/yunet_n.py:4-11 and the parameter
order used for optimizer-state interchange with the reference checkpoints."""
import os

import numpy as np
import pytest

from conftest import GOLDEN
from libfacedetection.train_b200 import trainer, _capi
from libfacedetection.train_b200.engine import ARCHS


def test_lr_schedule_matches_reference_config():
    # warm-up: linear from 0.001*lr over 1500 iterations (mmcv StepLrUpdaterHook semantics)
    assert trainer.lr_at(0, 0) == pytest.approx(0.01 * 0.001)
    assert trainer.lr_at(750, 0) == pytest.approx(0.01 * (1 - 0.5 * 0.999))
    assert trainer.lr_at(1500, 0) == pytest.approx(0.01)
    assert trainer.lr_at(10 ** 5, 399) == pytest.approx(0.01)
    assert trainer.lr_at(10 ** 5, 400) == pytest.approx(0.001)
    assert trainer.lr_at(10 ** 5, 544) == pytest.approx(0.0001)


@pytest.mark.parametrize('arch', ['yunet_n', 'yunet_s'])
def test_param_order_is_the_reference_state_dict_order(arch):
    a = ARCHS[arch]
    ctx = _capi.Ctx(_capi.make_arch_cfg(a['stage_channels'], a['downsample_idx'], a['out_idx'],
                                        a['shared_stacked_convs']))
    names = [n for n, _, _ in ctx.params()]
    d = np.load(os.path.join(GOLDEN, f'weights_{arch}.npz'))
    ref = [k for k in d.files if 'running_' not in k and 'num_batches' not in k]
    mine = trainer.reference_param_order(names)
    assert sorted(mine) == sorted(ref)
    if arch == 'yunet_n':
        # the yunet_n checkpoint was written by the current reference code: identical order
        assert mine == ref
    else:
        # weights/yunet_s.pth predates a reordering (kps before obj); loading follows the file
        swap = lambda n: n.replace('multi_level_obj', '#').replace('multi_level_kps', 'multi_level_obj').replace('#', 'multi_level_kps')
        assert [k for k in mine if 'multi_level_obj' not in k and 'multi_level_kps' not in k] == \
            [k for k in ref if 'multi_level_obj' not in k and 'multi_level_kps' not in k]


class _StubEngine:
    """CPU stand-in with the real parameter table / bucket layout of the plan (no GPU)."""

    def __init__(self, arch, seed):
        import torch
        a = ARCHS[arch]
        self.ctx = _capi.Ctx(_capi.make_arch_cfg(a['stage_channels'], a['downsample_idx'], a['out_idx'],
                                                 a['shared_stacked_convs']))
        self.param_table = self.ctx.params()
        self.bn_table = self.ctx.bns()
        g = torch.Generator().manual_seed(seed)
        n, nbn = self.ctx.num_params, self.ctx.num_bn_channels
        self.params = torch.randn(n, generator=g)
        self.momentum_buf = torch.randn(n, generator=g)
        self.bn_running = torch.rand(2 * nbn, generator=g)
        self.num_batches_tracked = 7

    def param_views(self, flat=None):
        flat = self.params if flat is None else flat
        return {n: flat[o:o + int(np.prod(s))].view(*s) for n, o, s in self.param_table}

    def state_dict(self):
        import torch
        sd = {k: v.clone() for k, v in self.param_views().items()}
        nbn = self.ctx.num_bn_channels
        for name, off, ch in self.bn_table:
            sd[name + '.running_mean'] = self.bn_running[off:off + ch].clone()
            sd[name + '.running_var'] = self.bn_running[nbn + off:nbn + off + ch].clone()
            sd[name + '.num_batches_tracked'] = torch.tensor(self.num_batches_tracked)
        return sd

    def load_state_dict(self, sd, strict=True):
        views = self.param_views()
        nbn = self.ctx.num_bn_channels
        for k, v in sd.items():
            if k in views:
                views[k].copy_(v)
        for name, off, ch in self.bn_table:
            self.bn_running[off:off + ch] = sd[name + '.running_mean']
            self.bn_running[nbn + off:nbn + off + ch] = sd[name + '.running_var']


@pytest.mark.parametrize('arch', ['yunet_n', 'yunet_s'])
def test_checkpoint_round_trip_restores_parameters_and_momentum(arch, tmp_path):
    """optimizer state must survive save -> load, including parameters whose
    plan (bucket) order differs from the reference order and share a shape (neck laterals)."""
    import torch
    a, b = _StubEngine(arch, 1), _StubEngine(arch, 2)
    path = str(tmp_path / 'ck.pth')
    trainer.save_checkpoint(a, path, epoch=3, iteration=17)
    meta = trainer.load_checkpoint(b, path)
    assert meta['epoch'] == 3 and meta['iter'] == 17
    assert torch.equal(a.params, b.params)
    assert torch.equal(a.momentum_buf, b.momentum_buf)
    assert torch.equal(a.bn_running, b.bn_running)
    ck = torch.load(path, weights_only=False)
    # file layout: reference key order; optimizer index i <-> i-th parameter key of the file
    pkeys = [k for k in ck['state_dict'] if 'running_' not in k and 'num_batches' not in k]
    assert pkeys == trainer.reference_param_order(a)
    mom = a.param_views(a.momentum_buf)
    for i, k in enumerate(pkeys):
        assert torch.equal(ck['optimizer']['state'][i]['momentum_buffer'], mom[k])
    # a foreign file (no names, like the reference's own checkpoints) resolves by key order
    del ck['optimizer']['param_names']
    torch.save(ck, path)
    c = _StubEngine(arch, 3)
    trainer.load_checkpoint(c, path)
    assert torch.equal(a.momentum_buf, c.momentum_buf)


def test_state_dict_order_is_the_reference_file_order():
    d = np.load(os.path.join(GOLDEN, 'weights_yunet_n.npz'))
    assert trainer.reference_state_dict_order(sorted(d.files)) == list(d.files)
