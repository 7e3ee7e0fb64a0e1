"""SURVEY §8f N4 — WIDER Face AP protocol (host code): the array restatement in
``evaluation.py`` against the unmodified reference evaluator (``core/evaluation/widerface.py``;
fixture with its outputs on a synthetic dataset: oracle/gen_golden_evaluation.py)."""
import copy
import os

import numpy as np

from libfacedetection.train_b200 import evaluation as E

GOLD = os.path.join(os.path.dirname(__file__), 'golden')


def _load():
    g = np.load(os.path.join(GOLD, 'evaluation_synth.npz'))
    gt, pred, match = {}, {}, {}
    for key in g.files:
        parts = key.split('/')
        if parts[0] == 'gt':
            gt.setdefault(parts[1], {}).setdefault(parts[2], {})[parts[3]] = g[key]
        elif parts[0] == 'pred':
            pred.setdefault(parts[1], {})[parts[2]] = g[key]
        elif parts[0] == 'match':
            match.setdefault((parts[1], parts[2]), {})[parts[3]] = g[key]
    return g['aps'], gt, pred, match


def test_aps_identical_to_reference_evaluator():
    aps_ref, gt, pred, _ = _load()
    aps, curves = E.wider_evaluation(copy.deepcopy(pred), gt, 0.5, return_curves=True)
    assert np.array_equal(np.asarray(aps), aps_ref)           # bit-identical float64
    assert aps[0] <= aps[1] <= aps[2] or True                 # no ordering guarantee on synthetic data
    for c in curves:
        assert c.shape == (1000, 2) and np.nanmax(c[:, 1]) <= 1.0


def test_per_image_matching_identical_to_reference():
    _, gt, pred, match = _load()
    normed = E.norm_score(copy.deepcopy(pred))
    assert len(match) >= 10
    for (ev, name), m in match.items():
        g = gt[ev][name]
        ignore = np.zeros(len(g['boxes']), dtype=int)
        ignore[g['hard'] - 1] = 1
        pr, pl = E.image_eval(normed[ev][name], g['boxes'], ignore, 0.5)
        assert np.array_equal(pr, m['pred_recall']) and np.array_equal(pl, m['proposal'])


def test_edge_cases_and_result_packing():
    # image without faces / without detections contribute nothing but do not crash
    gt = {'e': {'a': dict(boxes=np.zeros((0, 4)), easy=np.zeros(0, int), medium=np.zeros(0, int), hard=np.zeros(0, int)),
                'b': dict(boxes=np.array([[10., 10., 20., 20.]]), easy=np.array([1]), medium=np.array([1]), hard=np.array([1]))}}
    res = {}
    E.detections_to_results(res, 'e', 'a', np.zeros((0, 5), np.float32))
    E.detections_to_results(res, 'e', 'b', np.array([[10., 10., 30., 30., 0.9], [200., 200., 220., 220., 0.5]], np.float32))
    assert np.allclose(res['e']['b'][0], [10, 10, 20, 20, 0.9]) and res['e']['b'].dtype == np.float32
    aps = E.wider_evaluation(res, gt)
    assert all(abs(a - 1.0) < 1e-12 for a in aps)             # the one face is found first: AP = 1
    txt = E.prediction_file_text('e', 'b.jpg', np.array([[1., 2., 4., 6., 0.25]]))
    assert txt == 'e/b.jpg\n1\n1.00000 2.00000 3.00000 4.00000 0.25\n'
    # IoU helper: +1 convention of the reference
    o = E.pairwise_overlap(np.array([[0., 0., 9., 9.]]), np.array([[0., 0., 9., 9.], [5., 5., 14., 14.], [20., 20., 30., 30.]]))
    assert np.allclose(o[0], [1.0, 25.0 / 175.0, 0.0])


def test_test_time_preparation_and_driver():
    rng = np.random.default_rng(0)
    img = rng.integers(0, 255, (300, 500, 3), dtype=np.uint8)
    chw, f = E.prepare_test_image(img, mode=0)
    assert chw.shape == (3, 640, 640) and chw.dtype == np.float32
    assert np.allclose(f, [640 / 500, 384 / 300, 640 / 500, 384 / 300])       # 500x300 -> 640x384
    assert chw[:, 384:, :].max() == 0 and chw[:, :384, :].max() > 0            # zero pad below
    chw2, f2 = E.prepare_test_image(img, mode=2)
    assert chw2.shape == (3, 320, 512) and np.array_equal(f2, np.ones(4, np.float32))
    assert np.array_equal(chw2[:, :300, :500], img.transpose(2, 0, 1).astype(np.float32))
    # a detector that returns the ground truth (in network coordinates) scores AP = 1 on every subset
    gt = {'e': {'a': dict(boxes=np.array([[50., 60., 40., 40.], [200., 100., 80., 90.]]),
                          easy=np.array([1]), medium=np.array([1, 2]), hard=np.array([1, 2]))}}

    def detect(chw, factor):
        b = gt['e']['a']['boxes']
        xyxy = np.concatenate([b[:, :2], b[:, :2] + b[:, 2:]], 1) * factor / factor    # rescale=True output
        return np.concatenate([xyxy, [[0.9], [0.8]]], 1).astype(np.float32)
    aps, results = E.evaluate_wider(detect, [('e', 'a', img)], gt, mode=0)
    assert all(abs(a - 1.0) < 1e-12 for a in aps) and results['e']['a'].shape == (2, 5)
