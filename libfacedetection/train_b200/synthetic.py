"""Synthetic WIDER-shaped training batches.

The reference's real input pipeline (``mmdet/datasets/retinaface.py:18-150`` + the crop/resize/
flip transforms) is out of scope; the benchmark and the parity tests use seeded synthetic batches
whose ground-truth statistics follow ``data/widerface/labelv2/train/labelv2.txt``:
faces per image ~ clip(round(lognormal(1.2, 1.2)), 1, 64); width ~ clip(exp(N(3.0, 0.8)), 4, 200),
height = 1.25 width; 5 landmarks uniform inside the box; a face carries landmarks (weight 1) with
probability 0.48, otherwise weight 0 (``retinaface.py:32-49``).  Every image has >= 1 face, as
``RandomSquareCrop`` guarantees (``mmdet/datasets/pipelines/transforms.py:1096-1098``).
"""
import numpy as np


def make_gt(batch, size=320, seed=0, max_faces=64):
    """Returns lists (len ``batch``) of gt_bboxes (G,4) f32 xyxy, gt_labels (G,) i64 zeros,
    gt_keypointss (G,5,3) f32 (x, y, weight)."""
    rng = np.random.default_rng(seed)
    boxes, labels, kpss = [], [], []
    for _ in range(batch):
        g = int(np.clip(np.round(rng.lognormal(1.2, 1.2)), 1, max_faces))
        w = np.clip(np.exp(rng.normal(3.0, 0.8, g)), 4, min(200, size * 0.6))
        h = np.minimum(1.25 * w, size * 0.75)
        x1 = rng.uniform(0, size - w)
        y1 = rng.uniform(0, size - h)
        bb = np.stack([x1, y1, x1 + w, y1 + h], 1).astype(np.float32)
        kp = np.empty((g, 5, 3), np.float32)
        kp[:, :, 0] = (x1[:, None] + rng.uniform(0, 1, (g, 5)) * w[:, None])
        kp[:, :, 1] = (y1[:, None] + rng.uniform(0, 1, (g, 5)) * h[:, None])
        has = (rng.uniform(0, 1, g) < 0.48).astype(np.float32)
        kp[:, :, 2] = has[:, None]
        boxes.append(bb)
        labels.append(np.zeros((g,), np.int64))
        kpss.append(kp)
    return boxes, labels, kpss


def make_images(batch, size=320, seed=0):
    """float32 NCHW BGR 0..255 un-normalised (``configs/yunet_n.py:27``), uniform noise."""
    rng = np.random.default_rng(seed + 1000003)
    return (rng.random((batch, 3, size, size), dtype=np.float32) * np.float32(255.0))


def pack_gt_csr(boxes, kpss):
    """Ragged GT lists → CSR arrays for the C-ABI: gt (sumG, 19) f32 rows
    [x1,y1,x2,y2, kx0,ky0,...,kx4,ky4, w0..w4] and offsets (B+1,) int32."""
    offs = np.zeros(len(boxes) + 1, np.int32)
    rows = []
    for i, (b, k) in enumerate(zip(boxes, kpss)):
        b = np.asarray(b, np.float32).reshape(-1, 4)
        k = np.asarray(k, np.float32).reshape(-1, 5, 3)
        offs[i + 1] = offs[i] + b.shape[0]
        rows.append(np.concatenate([b, k[:, :, :2].reshape(-1, 10), k[:, :, 2]], 1))
    gt = np.concatenate(rows, 0).astype(np.float32) if rows else np.zeros((0, 19), np.float32)
    return np.ascontiguousarray(gt), offs
