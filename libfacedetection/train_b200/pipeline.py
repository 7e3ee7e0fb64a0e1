"""GPU input pipeline: the reference's training transforms for one batch,

    RandomSquareCrop(crop_choice) -> Resize(S, S, keep_ratio=False) -> RandomFlip(0.5)
    -> Normalize(mean 0, std 1, to_rgb=False) -> DefaultFormatBundle          (configs/yunet_n.py:37-57)

split the way the hardware wants it: the *random decisions* and the handful of box / landmark
coordinates stay on the host and follow the reference statement by statement (same numpy global
RNG calls in the same order, same dtypes — ``tests/test_pipeline.py`` pins them against the
unmodified transforms), the *pixels* — decoded uint8 images of arbitrary sizes — go to the device
once (1 B per sample instead of 4) and are cropped, padded, resized and flipped by ONE kernel launch
per batch (``csrc/preprocess.cu`` via ``yunet_preprocess_u8``) straight into the ``(B,3,S,S)``
fp32 tensor the detector reads.

Reference behaviour mirrored (mmdet/datasets/pipelines/transforms.py):
  * ``RandomSquareCrop.__call__`` :1014-1155 — ``np.random.choice(crop_choice)`` per scale retry,
    ``np.random.randint`` for left / top (``from numpy import random``, :10), at most 250 placements
    per scale, a placement is kept when at least one box centre lies strictly inside; boxes and
    landmarks of the kept faces are clipped to the patch and shifted; pixels outside the image
    are 128.
  * ``Resize._resize_img/_resize_bboxes/_resize_keypoints`` :241-298 — scale factors ``S / side``
    as float32, coordinates clipped to ``[0, S]``.
  * ``RandomFlip.__call__`` :488-535 — ``np.random.choice([horizontal, None], p=[r, 1-r])``, boxes
    ``x -> S - x`` (swapped), landmarks re-ordered ``[1,0,2,4,3]`` and mirrored (:439-486).
"""
import numpy as np
import torch

from . import _capi

CROP_CHOICE_N = (0.5, 0.7, 0.9, 1.1, 1.3, 1.5)       # configs/yunet_n.py:41-43
CROP_CHOICE_S = (0.3, 0.45, 0.6, 0.8, 1.0)           # configs/yunet_s.py:41
FLIP_ORDER = (1, 0, 2, 4, 3)
PAD_VALUE = 128.0


def _centers_in_patch(boxes, patch):
    center = (boxes[:, :2] + boxes[:, 2:]) / 2
    return ((center[:, 0] > patch[0]) * (center[:, 1] > patch[1]) *
            (center[:, 0] < patch[2]) * (center[:, 1] < patch[3]))


def sample_square_crop(h, w, boxes, crop_choice=CROP_CHOICE_N, rng=None):
    """Patch ``(left, top, right, bottom)`` (int64) of RandomSquareCrop for an ``h x w`` image with
    ground-truth ``boxes`` (N,4).  ``rng``: object with ``choice`` / ``randint`` (default: the numpy
    global generator the reference uses)."""
    rng = np.random if rng is None else rng
    max_scale = np.amax(crop_choice)
    scale_retry = 0
    scale = None
    while True:
        scale_retry += 1
        if scale_retry == 1 or max_scale > 1.0:
            scale = rng.choice(crop_choice)
        else:
            scale = scale * 1.2
        for _ in range(250):
            short_side = min(w, h)
            cw = int(scale * short_side)
            ch = cw
            if w == cw:
                left = 0
            elif w > cw:
                left = rng.randint(0, w - cw)
            else:
                left = rng.randint(w - cw, 0)
            if h == ch:
                top = 0
            elif h > ch:
                top = rng.randint(0, h - ch)
            else:
                top = rng.randint(h - ch, 0)
            patch = np.array((int(left), int(top), int(left + cw), int(top + ch)), dtype=np.int64)
            if _centers_in_patch(boxes, patch).any():
                return patch


def crop_gt(boxes, kps, patch, clip=True):
    """Boxes (N,4) / landmarks (N,5,3) of the faces whose centre is inside ``patch``, clipped to it
    and shifted to patch coordinates; also returns the boolean keep mask."""
    mask = _centers_in_patch(boxes, patch)
    b = boxes.copy()[mask]
    if clip:
        b[:, 2:] = b[:, 2:].clip(max=patch[2:])
        b[:, :2] = b[:, :2].clip(min=patch[:2])
    b -= np.tile(patch[:2], 2)
    k = kps.copy()[mask, :, :]
    if clip:
        k[:, :, :2] = k[:, :, :2].clip(max=patch[2:])
        k[:, :, :2] = k[:, :, :2].clip(min=patch[:2])
    k[:, :, 0] -= patch[0]
    k[:, :, 1] -= patch[1]
    return b, k, mask


def resize_gt(boxes, kps, side, S, clip=True):
    w_scale = S / side
    h_scale = S / side
    factor = np.array([w_scale, h_scale, w_scale, h_scale], dtype=np.float32)
    b = boxes * factor
    if clip:
        b[:, 0::2] = np.clip(b[:, 0::2], 0, S)
        b[:, 1::2] = np.clip(b[:, 1::2], 0, S)
    k = kps.copy()
    k[:, :, 0] *= factor[0]
    k[:, :, 1] *= factor[1]
    if clip:
        k[:, :, 0] = np.clip(k[:, :, 0], 0, S)
        k[:, :, 1] = np.clip(k[:, :, 1], 0, S)
    return b, k


def flip_gt(boxes, kps, S):
    b = boxes.copy()
    b[..., 0::4] = S - boxes[..., 2::4]
    b[..., 2::4] = S - boxes[..., 0::4]
    k = kps.copy()
    for idx, a in enumerate(FLIP_ORDER):
        k[:, idx, :] = kps[:, a, :]
    k[..., 0] = S - k[..., 0]
    return b, k


def sample_flip(flip_ratio=0.5, rng=None):
    rng = np.random if rng is None else rng
    return int(rng.choice(2, p=[flip_ratio, 1 - flip_ratio])) == 0      # index 0 = 'horizontal'


def augment_sample(h, w, boxes, kps, labels, S, crop_choice=CROP_CHOICE_N, flip_ratio=0.5, rng=None):
    """All host-side decisions and ground-truth arithmetic of one sample.  Returns
    ``(left, top, side, flip)`` for the kernel and the transformed ``boxes, kps, labels``."""
    patch = sample_square_crop(h, w, boxes, crop_choice, rng)
    b, k, mask = crop_gt(boxes, kps, patch)
    lab = labels[mask]
    side = int(patch[2] - patch[0])
    b, k = resize_gt(b, k, side, S)
    flip = sample_flip(flip_ratio, rng)
    if flip:
        b, k = flip_gt(b, k, S)
    return (int(patch[0]), int(patch[1]), side, int(flip)), b, k, lab


class GpuAugmenter:
    """Batches decoded uint8 BGR images (any sizes) + annotations into the detector's inputs.

    >>> aug = GpuAugmenter(engine, size=320)
    >>> img, gt, offsets = aug(images, boxes, kps, labels)     # img (B,3,320,320) fp32 on the GPU
    >>> engine.train_step(img, gt, offsets, lr=...)
    """

    def __init__(self, engine, size=320, crop_choice=CROP_CHOICE_N, flip_ratio=0.5, rng=None):
        self.engine = engine
        self.size = int(size)
        self.crop_choice = tuple(crop_choice)
        self.flip_ratio = flip_ratio
        self.rng = rng

    def decide(self, images, boxes, kps, labels):
        """Host part: per-sample kernel parameters and transformed ground truth."""
        params, gb, gk, gl = [], [], [], []
        for img, b, k, l in zip(images, boxes, kps, labels):
            p, b2, k2, l2 = augment_sample(img.shape[0], img.shape[1], b, k, l, self.size,
                                           self.crop_choice, self.flip_ratio, self.rng)
            params.append(p)
            gb.append(b2); gk.append(k2); gl.append(l2)
        return params, gb, gk, gl

    def pixels(self, images, params, out=None):
        """Device part: one H2D copy of the raw bytes, one kernel launch."""
        eng, S, B = self.engine, self.size, len(images)
        sizes = [int(im.shape[0]) * int(im.shape[1]) * 3 for im in images]
        offs = np.zeros(B, np.int64)
        offs[1:] = np.cumsum(sizes)[:-1]
        flat = torch.empty(int(sum(sizes)), dtype=torch.uint8, pin_memory=True)
        fl = flat.numpy()
        for im, o, n in zip(images, offs, sizes):
            assert im.dtype == np.uint8 and im.ndim == 3 and im.shape[2] == 3
            fl[o:o + n] = np.ascontiguousarray(im).reshape(-1)
        dev = eng.device
        d_pix = flat.to(dev, non_blocking=True)
        d_off = torch.from_numpy(offs).to(dev)
        d_hw = torch.tensor([[im.shape[0], im.shape[1]] for im in images], dtype=torch.int32, device=dev)
        d_crop = torch.tensor(params, dtype=torch.int32, device=dev)
        if out is None:
            out = torch.empty(B, 3, S, S, device=dev)
        _capi.check(eng.h, _capi.lib.yunet_preprocess_u8(
            eng.h, d_pix.data_ptr(), d_off.data_ptr(), d_hw.data_ptr(), d_crop.data_ptr(), B, S,
            PAD_VALUE, out.data_ptr(), torch.cuda.current_stream(dev).cuda_stream), 'yunet_preprocess_u8')
        return out

    def __call__(self, images, boxes, kps, labels):
        from . import synthetic
        params, gb, gk, gl = self.decide(images, boxes, kps, labels)
        img = self.pixels(images, params)
        gt, offs = synthetic.pack_gt_csr(gb, gk)
        dev = self.engine.device
        return img, torch.from_numpy(gt).to(dev), torch.from_numpy(offs).to(dev)
