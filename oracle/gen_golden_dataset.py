"""TEST INFRASTRUCTURE (development container only: reads /root/reference).

Pins ``libfacedetection.train_b200.dataset`` against the reference's ``RetinaFaceDataset`` label
parser (``mmdet/datasets/retinaface.py``): a synthetic ``labelv2.txt`` covering every line shape is
written to ``tests/golden/labelv2_synth.txt``, parsed by the unmodified class methods (called on a
bare instance, the mmcv dataset machinery is not needed), and the resulting annotation arrays are
stored in ``tests/golden/dataset_synth.npz``.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_loader  # noqa: E402

GOLD = os.path.join(ROOT, 'tests', 'golden')

LABEL = """# 0--Parade/0_Parade_a.jpg 1024 768
10.5 20 110.25 140 30 40 0 50 41 0 40 60 1 32 80 0 52 81 0 0.9
200 210 205 216 -1 -1 -1 -1 -1 -1 -1 -1 -1 -1 -1 -1 -1 -1 -1 0.3
300 300 380 390 310 320 0 -1 -1 -1 340 350 1 -1 -1 -1 360 370 0 0.7
# 1--Handshaking/1_Handshaking_b.jpg 640 480
5 6 7 8 1
50 60 150 170 0
# 2--Empty/2_Empty_c.jpg 320 240
# 3--Mixed/3_Mixed_d.jpg 800 600
100 100 104 104 101 101 0 102 101 0 102 102 0 101 103 0 103 103 0 0.5
400 100 500 220 0
"""


def main():
    ref_loader.install()
    import pycocotools
    pycocotools.__version__ = '12.0.2'
    from mmdet.datasets.retinaface import RetinaFaceDataset
    path = os.path.join(GOLD, 'labelv2_synth.txt')
    with open(path, 'w') as f:
        f.write(LABEL)
    from libfacedetection.train_b200 import dataset as D
    out = {}
    for tag, min_size, test_mode in (('train', None, False), ('train_min8', 8, False), ('test', None, True)):
        ds = object.__new__(RetinaFaceDataset)
        ds.NK, ds.min_size, ds.test_mode = 5, min_size, test_mode
        ds.cat2label = {'FG': 0}
        ds.data_infos = ds.load_annotations(path)
        ours = D.load_annotations(path, min_size, test_mode)
        assert [d['filename'] for d in ds.data_infos] == [d['filename'] for d in ours]
        out[f'{tag}/files'] = np.array([d['filename'] for d in ds.data_infos])
        out[f'{tag}/sizes'] = np.array([[d['width'], d['height']] for d in ds.data_infos])
        for i in range(len(ds.data_infos)):
            ann = ds.get_ann_info(i)
            mine = D.get_ann_info(ours[i])
            for k, v in ann.items():
                assert v.dtype == mine[k].dtype and np.array_equal(v, mine[k]), (tag, i, k)
                out[f'{tag}/{i}/{k}'] = v
    np.savez_compressed(os.path.join(GOLD, 'dataset_synth.npz'), **out)
    print('[dataset] label parser identical to RetinaFaceDataset for train / min_size / test modes')


if __name__ == '__main__':
    main()
