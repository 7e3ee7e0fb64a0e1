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
