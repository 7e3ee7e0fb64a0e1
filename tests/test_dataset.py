"""RetinaFace label reader (host code) against the reference's ``RetinaFaceDataset`` parser
(fixtures: oracle/gen_golden_dataset.py) and its hand-off to the augmentation host mirror."""
import os

import numpy as np
import pytest

from libfacedetection.train_b200 import dataset as D, pipeline as P

GOLD = os.path.join(os.path.dirname(__file__), 'golden')
LABEL = os.path.join(GOLD, 'labelv2_synth.txt')


@pytest.mark.parametrize('tag,min_size,test_mode', [('train', None, False), ('train_min8', 8, False),
                                                    ('test', None, True)])
def test_parser_identical_to_reference(tag, min_size, test_mode):
    g = np.load(os.path.join(GOLD, 'dataset_synth.npz'))
    infos = D.load_annotations(LABEL, min_size, test_mode)
    assert [d['filename'] for d in infos] == list(g[f'{tag}/files'])
    assert np.array_equal(np.array([[d['width'], d['height']] for d in infos]), g[f'{tag}/sizes'])
    for i, info in enumerate(infos):
        ann = D.get_ann_info(info)
        for k, v in ann.items():
            ref = g[f'{tag}/{i}/{k}']
            assert v.dtype == ref.dtype and v.shape == ref.shape and np.array_equal(v, ref), (i, k)


def test_parser_semantics():
    infos = D.load_annotations(LABEL)
    assert len(infos) == 3                                    # the image without faces is dropped in training
    a0 = D.get_ann_info(infos[0])
    assert a0['bboxes'].shape == (3, 4) and a0['keypointss'].shape == (3, 5, 3)
    assert np.array_equal(a0['keypointss'][1, :, 2], np.zeros(5))          # all landmarks missing
    assert np.array_equal(a0['keypointss'][2, :, 2], [1, 0, 1, 0, 1])
    a1 = D.get_ann_info(infos[1])
    assert a1['bboxes'].shape == (1, 4) and a1['bboxes_ignore'].shape == (1, 4)   # flag 1 = ignore
    assert len(D.load_annotations(LABEL, test_mode=True)) == 4


def test_samples_feed_the_augmenter_host_side():
    sizes = {i['filename']: (i['height'], i['width']) for i in D.load_annotations(LABEL)}
    fake = lambda p: np.full(sizes[os.path.relpath(p, 'root')] + (3,), 7, np.uint8)   # noqa: E731
    ds = D.RetinaFaceSamples(LABEL, img_prefix='root', imread=fake)
    assert len(ds) == 3
    np.random.seed(0)
    batches = list(ds.batches(2, shuffle=True, drop_last=False))
    assert sum(len(b[0]) for b in batches) == 3
    for images, boxes, kps, labels in batches:
        for img, b, k, l in zip(images, boxes, kps, labels):
            assert img.dtype == np.uint8 and b.dtype == np.float32 and k.shape[1:] == (5, 3)
            (left, top, side, flip), b2, k2, l2 = P.augment_sample(img.shape[0], img.shape[1], b, k, l, 320)
            assert side > 0 and b2.shape[0] >= 1 and b2.shape[0] == k2.shape[0] == l2.shape[0]
            assert (b2 >= 0).all() and (b2 <= 320).all()
