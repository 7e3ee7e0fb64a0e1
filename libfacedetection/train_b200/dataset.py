"""RetinaFace-format annotations (input side): the label file reader of the
reference's ``RetinaFaceDataset`` (``mmdet/datasets/retinaface.py:28-150``) without mmdet — what a
dataloader worker needs besides ``cv2.imread`` to feed ``pipeline.GpuAugmenter``.

File format (``labelv2.txt``): ``# <relative path> <width> <height>`` starts an image, every
following line is one face: ``x1 y1 x2 y2`` then either 5 x ``(x y flag)`` landmarks (a triple of
-1 = missing -> weight 0, otherwise weight 1) or a single ignore flag.  Faces smaller than
``min_size`` or flagged go to the ``*_ignore`` lists; training images without any face are dropped
(``retinaface.py:93-94``).
"""
import os

import numpy as np

NK = 5


def parse_ann_line(line, min_size=None, test_mode=False):
    """One face line -> ``dict(bbox (4,) f32, kps (5,3) f32 [x, y, weight], ignore, cat)``."""
    v = np.array(line.split(), dtype=np.float64)
    box = v[:4].astype(np.float32)
    too_small = False
    if min_size is not None:
        if test_mode:
            raise AssertionError('min_size is a training-only filter')
        too_small = bool((box[2] - box[0]) < min_size or (box[3] - box[1]) < min_size)
    kps = np.zeros((NK, 3), dtype=np.float32)
    flagged = False
    if v.size > 5:                                   # x y flag per landmark (a trailing score may follow)
        kps = v[4:4 + 3 * NK].astype(np.float32).reshape(NK, 3)
        missing = (kps == -1).all(axis=1)
        if (kps[~missing, 2] < 0).any():
            raise AssertionError('negative landmark flag')
        kps[:, 2] = np.where(missing, 0.0, 1.0)      # the third column becomes the loss weight
    elif v.size == 5:                                # box + ignore flag
        flagged = v[4] == 1
    elif not test_mode:
        raise AssertionError('a box without landmarks or flag is only valid in test mode')
    return dict(bbox=box, kps=kps, ignore=bool(too_small or flagged), cat='FG')


def load_annotations(ann_file, min_size=None, test_mode=False):
    """List of ``dict(filename, width, height, objs)`` in file order; training images whose face list
    is empty are dropped."""
    images, current = [], None
    with open(ann_file, 'r') as f:
        for raw in f:
            raw = raw.strip()
            if raw.startswith('#'):
                path, width, height = raw[1:].split()[:3]
                current = dict(filename=path, width=int(width), height=int(height), objs=[])
                # a repeated header replaces the earlier entry but keeps its position (dict semantics)
                for k, im in enumerate(images):
                    if im['filename'] == path:
                        images[k] = current
                        break
                else:
                    images.append(current)
            elif current is None:
                raise AssertionError('face line before the first "# <image>" header')
            else:
                current['objs'].append(parse_ann_line(raw, min_size, test_mode))
    return [im for im in images if im['objs'] or test_mode]


def get_ann_info(data_info):
    """``bboxes (N,4) f32, labels (N,) i64, keypointss (N,5,3) f32, bboxes_ignore, labels_ignore``."""
    used = [o for o in data_info['objs'] if not o['ignore']]
    skipped = [o for o in data_info['objs'] if o['ignore']]

    def stack(objs, key, shape):
        if not objs:
            return np.zeros((0,) + shape, np.float32)
        return np.stack([o[key] for o in objs]).astype(np.float32)

    return dict(bboxes=stack(used, 'bbox', (4,)), labels=np.zeros(len(used), np.int64),
                keypointss=stack(used, 'kps', (NK, 3)),
                bboxes_ignore=stack(skipped, 'bbox', (4,)), labels_ignore=np.zeros(len(skipped), np.int64))


class RetinaFaceSamples:
    """Indexable training samples: ``(uint8 BGR HWC image, bboxes, keypointss, labels)`` — the inputs
    of ``pipeline.GpuAugmenter``.  Images are decoded with ``cv2.imread`` like the reference's
    ``LoadImageFromFile`` (mmcv ``imread``, cv2 backend, colour)."""

    def __init__(self, ann_file, img_prefix='', min_size=None, test_mode=False, imread=None):
        self.infos = load_annotations(ann_file, min_size, test_mode)
        self.img_prefix = img_prefix
        if imread is None:
            import cv2
            imread = lambda p: cv2.imread(p, cv2.IMREAD_COLOR)     # noqa: E731
        self.imread = imread

    def __len__(self):
        return len(self.infos)

    def __getitem__(self, i):
        info = self.infos[i]
        ann = get_ann_info(info)
        img = self.imread(os.path.join(self.img_prefix, info['filename']))
        return img, ann['bboxes'], ann['keypointss'], ann['labels']

    def num_faces(self, i):
        """Usable (non-ignored) faces of sample ``i`` from the annotations alone (no image decode)."""
        return int(get_ann_info(self.infos[i])['bboxes'].shape[0])

    def batches(self, batch_size, shuffle=True, rng=None, drop_last=True):
        """Lists ``(images, bboxes, keypointss, labels)`` of ``batch_size`` samples; images without a
        usable face (all ignored) are skipped, as RandomSquareCrop needs at least one centre."""
        order = np.arange(len(self))
        if shuffle:
            (np.random if rng is None else rng).shuffle(order)
        cur = [[], [], [], []]
        for i in order:
            s = self[i]
            if s[1].shape[0] == 0:
                continue
            for c, v in zip(cur, s):
                c.append(v)
            if len(cur[0]) == batch_size:
                yield tuple(cur)
                cur = [[], [], [], []]
        if cur[0] and not drop_last:
            yield tuple(cur)
