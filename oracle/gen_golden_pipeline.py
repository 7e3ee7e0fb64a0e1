"""TEST INFRASTRUCTURE (development container only: reads /root/reference).

Pins the host mirror ``libfacedetection.train_b200.pipeline`` and the ``yunet_preprocess_u8`` kernel
against the reference's *unmodified* training transforms (mmdet/datasets/pipelines/transforms.py:
RandomSquareCrop, Resize, RandomFlip, Normalize, run through the stubbed mmcv of ``ref_loader`` with
the three image functions they call implemented the way mmcv 1.x implements them on the cv2
backend).  Writes ``tests/golden/pipeline_S{S}.npz``: seeded uint8 source images + annotations,
and per sample the reference's output image (float32 HWC), boxes, landmarks, labels and the
(left, top, side, flip) decisions recovered from it.
"""
import os
import sys

import cv2
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_loader  # noqa: E402

GOLD = os.path.join(ROOT, 'tests', 'golden')


def load_reference_transforms():
    ref_loader.install()
    if not hasattr(np, 'int'):
        np.int = int                      # numpy >= 1.24 dropped the alias the reference still uses
    import pycocotools
    pycocotools.__version__ = '12.0.2'    # the stub must pass mmdet/datasets/retinaface.py:9
    import mmcv

    def imresize(img, size, return_scale=False, interpolation='bilinear', out=None, backend=None):
        # mmcv/image/geometric.py imresize, cv2 backend: cv2.resize(img, size, interpolation=INTER_LINEAR)
        h, w = img.shape[:2]
        assert interpolation == 'bilinear'
        resized = cv2.resize(img, size, dst=out, interpolation=cv2.INTER_LINEAR)
        if not return_scale:
            return resized
        return resized, size[0] / w, size[1] / h

    def imflip(img, direction='horizontal'):
        assert direction == 'horizontal'
        return np.flip(img, axis=1)

    def imnormalize(img, mean, std, to_rgb=True):
        img = img.copy().astype(np.float32)
        mean = np.float64(mean.reshape(1, -1))
        stdinv = 1 / np.float64(std.reshape(1, -1))
        if to_rgb:
            cv2.cvtColor(img, cv2.COLOR_BGR2RGB, img)
        cv2.subtract(img, mean, img)
        cv2.multiply(img, stdinv, img)
        return img

    mmcv.imresize, mmcv.imflip, mmcv.imnormalize = imresize, imflip, imnormalize
    from mmdet.datasets.pipelines import transforms as T
    T.mmcv.imresize, T.mmcv.imflip, T.mmcv.imnormalize = imresize, imflip, imnormalize
    return T


def make_sample(rng, i):
    """A decoded uint8 BGR image of odd size with smooth content + a few faces."""
    h, w = int(rng.integers(90, 260)), int(rng.integers(90, 260))
    yy, xx = np.mgrid[0:h, 0:w]
    img = np.stack([(xx * 3 + yy * 2 + 17 * c) % 256 for c in range(3)], -1).astype(np.uint8)
    img = (img.astype(np.int32) + rng.integers(0, 3, (h // 8 + 1, w // 8 + 1, 3)).repeat(8, 0).repeat(8, 1)[:h, :w] * 20).clip(0, 255).astype(np.uint8)   # blocky, compressible
    n = int(rng.integers(1, 5))
    cx, cy = rng.uniform(10, w - 10, n), rng.uniform(10, h - 10, n)
    s = rng.uniform(8, 40, n)
    boxes = np.stack([cx - s, cy - s, cx + s, cy + s], 1).astype(np.float32)
    kps = np.zeros((n, 5, 3), np.float32)
    kps[:, :, 0] = cx[:, None] + rng.uniform(-0.6, 0.6, (n, 5)) * s[:, None]
    kps[:, :, 1] = cy[:, None] + rng.uniform(-0.6, 0.6, (n, 5)) * s[:, None]
    kps[:, :, 2] = (rng.uniform(0, 1, (n, 5)) > 0.2).astype(np.float32)
    return img, boxes, kps, np.zeros(n, np.int64)


def run_case(S, nsamples=12, seed=5):
    T = load_reference_transforms()
    crop_choice = [0.5, 0.7, 0.9, 1.1, 1.3, 1.5]
    tr = [T.RandomSquareCrop(crop_choice=crop_choice), T.Resize(img_scale=(S, S), keep_ratio=False),
          T.RandomFlip(flip_ratio=0.5), T.Normalize(mean=[0., 0., 0.], std=[1., 1., 1.], to_rgb=False)]
    rng = np.random.default_rng(seed)
    out = {'S': S, 'n': nsamples, 'seed': seed}
    for i in range(nsamples):
        img, boxes, kps, labels = make_sample(rng, i)
        results = dict(img=img.astype(np.float32), img_shape=img.shape, ori_shape=img.shape,
                       img_fields=['img'], bbox_fields=['gt_bboxes'], keypoints_fields=['gt_keypointss'],
                       gt_bboxes=boxes.copy(), gt_labels=labels.copy(), gt_keypointss=kps.copy())
        np.random.seed(1000 + i)                     # the transforms draw from the numpy global RNG
        for t in tr:
            results = t(results)
        out[f'{i}/img'] = img
        out[f'{i}/boxes'], out[f'{i}/kps'], out[f'{i}/labels'] = boxes, kps, labels
        out[f'{i}/out_img'] = np.ascontiguousarray(results['img']).astype(np.float32)
        out[f'{i}/out_boxes'] = results['gt_bboxes'].astype(np.float32)
        out[f'{i}/out_kps'] = results['gt_keypointss'].astype(np.float32)
        out[f'{i}/out_labels'] = results['gt_labels']
        out[f'{i}/flip'] = bool(results['flip'])
        out[f'{i}/scale_factor'] = results['scale_factor']
    np.savez_compressed(os.path.join(GOLD, f'pipeline_S{S}.npz'), **out)
    return out


def check_host_mirror(out):
    from libfacedetection.train_b200 import pipeline as P
    S = int(out['S'])
    for i in range(int(out['n'])):
        img = out[f'{i}/img']
        np.random.seed(1000 + i)
        (left, top, side, flip), b, k, l = P.augment_sample(img.shape[0], img.shape[1], out[f'{i}/boxes'],
                                                            out[f'{i}/kps'], out[f'{i}/labels'], S)
        assert flip == int(out[f'{i}/flip']), i
        assert np.array_equal(b, out[f'{i}/out_boxes']), (i, b, out[f'{i}/out_boxes'])
        assert np.array_equal(k, out[f'{i}/out_kps']), i
        assert np.array_equal(l, out[f'{i}/out_labels']), i
        assert np.float32(S / side) == out[f'{i}/scale_factor'][0]
    print(f'[pipeline S={S}] host mirror reproduces the reference decisions and ground truth exactly '
          f'({int(out["n"])} samples)')


if __name__ == '__main__':
    for S in (64, 320):
        check_host_mirror(run_case(S, nsamples=10 if S == 64 else 2))
