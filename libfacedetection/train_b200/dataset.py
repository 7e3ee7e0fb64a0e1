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
    values = [float(x) for x in line.strip().split()]
    bbox = np.array(values[0:4], dtype=np.float32)
    kps = np.zeros((NK, 3), dtype=np.float32)
    ignore = False
    if min_size is not None:
        assert not test_mode
        w = bbox[2] - bbox[0]
        h = bbox[3] - bbox[1]
        if w < min_size or h < min_size:
            ignore = True
    if len(values) > 4:
        if len(values) > 5:
            kps = np.array(values[4:19], dtype=np.float32).reshape((NK, 3))
            for li in range(kps.shape[0]):
                if (kps[li, :] == -1).all():
                    kps[li][2] = 0.0          # weight 0: landmark not annotated
                else:
                    assert kps[li][2] >= 0
                    kps[li][2] = 1.0
        elif not ignore:
            ignore = (values[4] == 1)
    else:
        assert test_mode
    return dict(bbox=bbox, kps=kps, ignore=ignore, cat='FG')


def load_annotations(ann_file, min_size=None, test_mode=False):
    """List of ``dict(filename, width, height, objs)`` in file order."""
    name = None
    bbox_map = {}
    with open(ann_file, 'r') as f:
        for line in f:
            line = line.strip()
            if line.startswith('#'):
                value = line[1:].strip().split()
                name = value[0]
                bbox_map[name] = dict(width=int(value[1]), height=int(value[2]), objs=[])
                continue
            assert name is not None
            bbox_map[name]['objs'].append(line)
    data_infos = []
    for name, item in bbox_map.items():
        objs = [parse_ann_line(l, min_size, test_mode) for l in item['objs']]
        if len(objs) == 0 and not test_mode:
            continue
        data_infos.append(dict(filename=name, width=item['width'], height=item['height'], objs=objs))
    return data_infos


def get_ann_info(data_info):
    """``bboxes (N,4) f32, labels (N,) i64, keypointss (N,5,3) f32, bboxes_ignore, labels_ignore``."""
    bboxes, keypointss, labels, bboxes_ignore, labels_ignore = [], [], [], [], []
    for obj in data_info['objs']:
        if obj['ignore']:
            bboxes_ignore.append(obj['bbox'])
            labels_ignore.append(0)
        else:
            bboxes.append(obj['bbox'])
            labels.append(0)
            keypointss.append(obj['kps'])
    if not bboxes:
        bboxes, labels, keypointss = np.zeros((0, 4)), np.zeros((0,)), np.zeros((0, NK, 3))
    else:
        bboxes, labels, keypointss = np.array(bboxes, ndmin=2), np.array(labels), np.array(keypointss, ndmin=3)
    if not bboxes_ignore:
        bboxes_ignore, labels_ignore = np.zeros((0, 4)), np.zeros((0,))
    else:
        bboxes_ignore, labels_ignore = np.array(bboxes_ignore, ndmin=2), np.array(labels_ignore)
    return dict(bboxes=bboxes.astype(np.float32), labels=labels.astype(np.int64),
                keypointss=keypointss.astype(np.float32),
                bboxes_ignore=bboxes_ignore.astype(np.float32),
                labels_ignore=labels_ignore.astype(np.int64))


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
