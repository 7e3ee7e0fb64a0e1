"""SURVEY §8f N2 — input pipeline, host side (CPU): the random decisions and ground-truth arithmetic
of ``pipeline.augment_sample`` against the unmodified reference transforms (fixtures written by
``oracle/gen_golden_pipeline.py``), and a float32 numpy emulation of the ``yunet_preprocess_u8``
kernel formula against the reference's output image (the kernel itself: tests/test_gpu_pipeline.py)."""
import os

import numpy as np
import pytest

from libfacedetection.train_b200 import pipeline as P

GOLD = os.path.join(os.path.dirname(__file__), 'golden')


def emulate_kernel(img, left, top, side, flip, S, pad=128.0):
    """The arithmetic of csrc/preprocess.cu in numpy float32 (HWC output)."""
    H, W = img.shape[:2]
    scale = 1.0 / (S / side)          # cv2: scale_x = 1. / inv_scale_x, inv_scale_x = dst / src

    def coord(d, n):
        f = (d.astype(np.float64) + 0.5) * scale - 0.5            # cv2 keeps the coordinate in double,
        s = np.floor(f).astype(np.int64)
        f = (f - s).astype(np.float32)                            # only the fraction becomes float
        lo = s < 0
        s[lo], f[lo] = 0, 0
        hi = s >= n - 1
        s[hi], f[hi] = n - 1, 0
        return s, np.minimum(s + 1, n - 1), f.astype(np.float32)

    xs = np.arange(S)
    xs = (S - 1 - xs) if flip else xs
    cx0, cx1, fx = coord(xs, side)
    cy0, cy1, fy = coord(np.arange(S), side)

    def sample(cy, cx):
        iy, ix = cy[:, None] + top, cx[None, :] + left
        ok = (iy >= 0) & (iy < H) & (ix >= 0) & (ix < W)
        v = img[np.clip(iy, 0, H - 1), np.clip(ix, 0, W - 1)].astype(np.float32)
        return np.where(ok[..., None], v, np.float32(pad))

    a0, a1 = (1 - fx)[None, :, None], fx[None, :, None]
    b0, b1 = (1 - fy)[:, None, None], fy[:, None, None]
    r0 = sample(cy0, cx0) * a0 + sample(cy0, cx1) * a1
    r1 = sample(cy1, cx0) * a0 + sample(cy1, cx1) * a1
    return (r0 * b0 + r1 * b1).astype(np.float32)


@pytest.mark.parametrize('S', [64, 320])
def test_host_decisions_and_ground_truth_match_reference(S):
    g = np.load(os.path.join(GOLD, f'pipeline_S{S}.npz'))
    for i in range(int(g['n'])):
        img = g[f'{i}/img']
        np.random.seed(1000 + i)          # the reference transforms draw from the numpy global RNG
        (left, top, side, flip), b, k, l = P.augment_sample(img.shape[0], img.shape[1], g[f'{i}/boxes'],
                                                            g[f'{i}/kps'], g[f'{i}/labels'], S)
        assert flip == int(g[f'{i}/flip'])
        assert np.array_equal(b, g[f'{i}/out_boxes'])
        assert np.array_equal(k, g[f'{i}/out_kps'])
        assert np.array_equal(l, g[f'{i}/out_labels'])
        assert b.dtype == np.float32 and k.dtype == np.float32


@pytest.mark.parametrize('S', [64, 320])
def test_kernel_formula_matches_reference_pixels(S):
    g = np.load(os.path.join(GOLD, f'pipeline_S{S}.npz'))
    worst = 0.0
    seen_pad = seen_flip = False
    for i in range(int(g['n'])):
        img = g[f'{i}/img']
        np.random.seed(1000 + i)
        (left, top, side, flip), _, _, _ = P.augment_sample(img.shape[0], img.shape[1], g[f'{i}/boxes'],
                                                            g[f'{i}/kps'], g[f'{i}/labels'], S)
        seen_pad |= left < 0 or top < 0 or left + side > img.shape[1] or top + side > img.shape[0]
        seen_flip |= bool(flip)
        out = emulate_kernel(img, left, top, side, flip, S)
        worst = max(worst, float(np.abs(out - g[f'{i}/out_img']).max()))
    assert worst < 2e-3, worst             # pixel values 0..255: < 1e-5 relative
    if S == 64:
        assert seen_pad and seen_flip      # the fixture exercises the padded and the mirrored paths


def test_crop_retry_and_flip_order_edge_cases():
    # a box outside every small patch forces the scale retries (max_scale > 1 keeps drawing)
    boxes = np.array([[2., 2., 6., 6.]], np.float32)
    np.random.seed(3)
    patch = P.sample_square_crop(100, 140, boxes)
    c = (boxes[0, :2] + boxes[0, 2:]) / 2
    assert patch[0] < c[0] < patch[2] and patch[1] < c[1] < patch[3]
    assert patch[2] - patch[0] == patch[3] - patch[1]
    k = np.arange(15, dtype=np.float32).reshape(1, 5, 3)
    _, kf = P.flip_gt(np.array([[10., 0., 30., 5.]], np.float32), k, 64)
    assert np.array_equal(kf[0, :, 2], k[0, [1, 0, 2, 4, 3], 2])
    assert np.array_equal(kf[0, :, 0], 64 - k[0, [1, 0, 2, 4, 3], 0])
