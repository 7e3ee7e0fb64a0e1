"""SURVEY §8f N2 — the ``yunet_preprocess_u8`` kernel (through the C ABI) against the output images
of the unmodified reference transforms (fixtures: oracle/gen_golden_pipeline.py), and the batcher
feeding a training step."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN

pytestmark = pytest.mark.gpu

from libfacedetection.train_b200 import YuNetEngine, pipeline as P  # noqa: E402

TOL = 1e-3      # absolute, on 0..255 pixel values (the numpy emulation of the formula is bit-exact)


@pytest.mark.parametrize('S', [64, 320])
def test_preprocess_kernel_matches_reference_transforms(S):
    g = np.load(os.path.join(GOLDEN, f'pipeline_S{S}.npz'))
    n = int(g['n'])
    eng = YuNetEngine('yunet_n')
    aug = P.GpuAugmenter(eng, size=S)
    images = [g[f'{i}/img'] for i in range(n)]
    params, gb, gk = [], [], []
    for i in range(n):
        np.random.seed(1000 + i)
        p, b, k, _ = P.augment_sample(images[i].shape[0], images[i].shape[1], g[f'{i}/boxes'],
                                      g[f'{i}/kps'], g[f'{i}/labels'], S)
        params.append(p); gb.append(b); gk.append(k)
    out = aug.pixels(images, params).cpu().numpy()            # ragged batch, one launch
    assert out.shape == (n, 3, S, S)
    worst = 0.0
    for i in range(n):
        ref = g[f'{i}/out_img'].transpose(2, 0, 1)             # HWC -> CHW (DefaultFormatBundle)
        worst = max(worst, float(np.abs(out[i] - ref).max()))
        assert np.array_equal(gb[i], g[f'{i}/out_boxes']) and np.array_equal(gk[i], g[f'{i}/out_kps'])
    print(f'preprocess_u8 S={S}: worst abs pixel error {worst:.3e}')
    assert worst <= TOL


def test_augmenter_feeds_a_training_step():
    g = np.load(os.path.join(GOLDEN, 'pipeline_S64.npz'))
    n = int(g['n'])
    eng = YuNetEngine('yunet_n')
    eng.init_weights(0)
    aug = P.GpuAugmenter(eng, size=320)
    np.random.seed(7)
    img, gt, offs = aug([g[f'{i}/img'] for i in range(n)], [g[f'{i}/boxes'] for i in range(n)],
                        [g[f'{i}/kps'] for i in range(n)], [g[f'{i}/labels'] for i in range(n)])
    assert img.shape == (n, 3, 320, 320) and img.dtype == torch.float32
    assert float(img.min()) >= 0.0 and float(img.max()) <= 255.0
    assert gt.shape[1] == 19 and int(offs[-1]) == gt.shape[0] >= n
    losses = eng.train_step(img, gt, offs, lr=1e-5)
    assert torch.isfinite(losses).all()
