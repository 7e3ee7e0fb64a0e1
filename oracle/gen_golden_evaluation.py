"""TEST INFRASTRUCTURE (development container only: reads /root/reference).

Pins ``libfacedetection.train_b200.evaluation`` against the unmodified WIDER evaluation of the
reference (``mmdet/core/evaluation/widerface.py``): a synthetic dataset in the nested object-array
structure ``scipy.io.loadmat`` gives for the WIDER ``.mat`` files is pushed through the reference's
``wider_evaluation`` (its ``get_gt_boxes`` replaced by the synthetic structure, its process pool by a
serial ``starmap``), and the three APs plus a few per-image match lists are stored in
``tests/golden/evaluation_synth.npz`` together with the inputs.
"""
import importlib.util
import itertools
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
GOLD = os.path.join(ROOT, 'tests', 'golden')
REF = os.environ.get('YUNET_REFERENCE_ROOT', '/root/reference')


def load_reference_module():
    for alias, typ in (('float', float), ('int', int)):
        if not hasattr(np, alias):
            setattr(np, alias, typ)          # numpy >= 1.24 dropped the aliases the reference uses
    spec = importlib.util.spec_from_file_location(
        'ref_widerface', os.path.join(REF, 'mmdet', 'core', 'evaluation', 'widerface.py'))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


class SerialPool:
    def __init__(self, *a, **k):
        pass

    def starmap(self, fn, it):
        return list(itertools.starmap(fn, it))


def synth(seed=3, events=3, images=7):
    rng = np.random.default_rng(seed)
    gt, pred = {}, {}
    for e in range(events):
        ev = f'{e}--Event{e}'
        gt[ev], pred[ev] = {}, {}
        for j in range(images):
            name = f'{e}_Event{e}_{j}'
            g = int(rng.integers(0, 9)) if (e, j) != (0, 0) else 0          # one image without faces
            xy = rng.uniform(0, 500, (g, 2))
            wh = rng.uniform(6, 120, (g, 2))
            boxes = np.floor(np.concatenate([xy, wh], 1))
            keep = {}
            idx = np.arange(1, g + 1)
            keep['hard'] = idx[rng.uniform(size=g) < 0.9]
            keep['medium'] = keep['hard'][rng.uniform(size=len(keep['hard'])) < 0.8]
            keep['easy'] = keep['medium'][rng.uniform(size=len(keep['medium'])) < 0.6]
            gt[ev][name] = dict(boxes=boxes.astype('float'), **keep)
            # detections: jittered copies of some faces + false positives, float32, descending score
            sel = rng.uniform(size=g) < 0.8
            det = boxes[sel] + rng.normal(0, 3, (int(sel.sum()), 4))
            fp = np.concatenate([rng.uniform(0, 500, (int(rng.integers(0, 6)), 2)),
                                 rng.uniform(6, 100, (0, 2))], 1) if False else None
            nfp = int(rng.integers(0, 6)) if (e, j) != (1, 1) else 0
            fpb = np.concatenate([rng.uniform(0, 500, (nfp, 2)), rng.uniform(6, 100, (nfp, 2))], 1)
            det = np.concatenate([det, fpb], 0)
            if (e, j) == (2, 2):
                det = det[:0]                                               # one image without detections
            score = np.sort(rng.uniform(0.02, 1.0, det.shape[0]))[::-1]
            pred[ev][name] = np.concatenate([det, score[:, None]], 1).astype(np.float32)
    return gt, pred


def to_mat_structure(gt):
    """The nested (N,1) object arrays of loadmat: X[i][0][j][0] addressing."""
    def col(items):
        a = np.empty((len(items), 1), dtype=object)
        for i, it in enumerate(items):
            a[i, 0] = it
        return a
    events = list(gt)
    event_list = col([np.array([ev]) for ev in events])
    file_list = col([col([np.array([name]) for name in gt[ev]]) for ev in events])
    facebox = col([col([gt[ev][n]['boxes'] for n in gt[ev]]) for ev in events])
    lists = {s: col([col([gt[ev][n][s].reshape(-1, 1) for n in gt[ev]]) for ev in events])
             for s in ('easy', 'medium', 'hard')}
    return facebox, event_list, file_list, lists['hard'], lists['medium'], lists['easy']


def main():
    import copy
    import multiprocessing
    ref = load_reference_module()
    gt, pred = synth()
    structure = to_mat_structure(gt)
    ref.get_gt_boxes = lambda gt_dir: structure
    multiprocessing.Pool = SerialPool
    aps_ref = ref.wider_evaluation(copy.deepcopy(pred), 'unused', 0.5)
    from libfacedetection.train_b200 import evaluation as E
    aps, curves = E.wider_evaluation(copy.deepcopy(pred), gt, 0.5, return_curves=True)
    print('reference APs', aps_ref, '\nours        ', aps)
    assert np.array_equal(np.asarray(aps_ref, np.float64), np.asarray(aps, np.float64)), 'AP mismatch'
    # per-image match lists for a few images (reference image_eval with the serial pool)
    out = {'aps': np.asarray(aps_ref, np.float64)}
    normed = ref.norm_score(copy.deepcopy(pred))
    k = 0
    for ev in gt:
        for name, g in gt[ev].items():
            p = normed[ev][name]
            out[f'gt/{ev}/{name}/boxes'] = g['boxes']
            for s in ('easy', 'medium', 'hard'):
                out[f'gt/{ev}/{name}/{s}'] = g[s]
            out[f'pred/{ev}/{name}'] = pred[ev][name]
            if len(g['boxes']) == 0 or len(p) == 0:
                continue
            ignore = np.zeros(len(g['boxes']), dtype=int)
            ignore[g['hard'] - 1] = 1
            pr, pl = ref.image_eval(p, g['boxes'], ignore, 0.5, SerialPool())
            info, _ = ref.img_pr_info(1000, p, pl, pr)
            pr2, pl2 = E.image_eval(p, g['boxes'], ignore, 0.5)
            assert np.array_equal(pr, pr2) and np.array_equal(pl, pl2), (ev, name)
            assert np.array_equal(info, E.img_pr_info(1000, p, pl2, pr2)), (ev, name)
            out[f'match/{ev}/{name}/pred_recall'], out[f'match/{ev}/{name}/proposal'] = pr, pl
            k += 1
    np.savez_compressed(os.path.join(GOLD, 'evaluation_synth.npz'), **out)
    print(f'[evaluation] APs and {k} per-image match lists identical to the reference; fixture written')


if __name__ == '__main__':
    main()
