"""WIDER Face evaluation: the AP easy / medium / hard protocol of
``mmdet/core/evaluation/widerface.py`` (``wider_evaluation`` :274-346 and its helpers ``norm_score``
:159-180, ``image_eval`` :183-220, ``img_pr_info`` :223-243, ``dataset_pr_info`` :246-251, ``voc_ap``
:254-271), restated with array operations instead of the reference's per-prediction process pool and
per-threshold Python loops (the reference spends minutes on the 3 226 validation images; this runs in
seconds), and the result packing of ``tools/test_widerface.py:141-173``.

Same arithmetic, same outputs (``tests/test_evaluation.py`` pins the three APs, the PR curves and the
per-image match lists against the unmodified reference functions on a synthetic dataset): boxes are
``[x, y, w, h]`` with the +1 pixel convention, a prediction is matched to the ground-truth box of
largest IoU, matches to boxes outside the difficulty's keep list are ignored, 1 000 score
thresholds, VOC-style area under the precision envelope.

Ground truth is passed in memory: ``{event: {image: dict(boxes=(G,4) xywh, easy=idx, medium=idx,
hard=idx)}}`` with the 1-based keep indices of the WIDER ``.mat`` files; ``load_wider_gt`` builds it
from ``wider_face_val.mat`` / ``wider_{easy,medium,hard}_val.mat`` when they are available.
"""
import os

import numpy as np

SETTINGS = ('easy', 'medium', 'hard')
THRESH_NUM = 1000


def norm_score(pred):
    """Min-max normalise all scores of ``{event: {image: (N,5)}}`` in place (widerface.py:159-180;
    the bounds start from -1 / 2 like the reference's accumulators)."""
    arrays = [v for images in pred.values() for v in images.values() if len(v)]
    hi = max([-1] + [np.max(v[:, -1]) for v in arrays])
    lo = min([2] + [np.min(v[:, -1]) for v in arrays])
    span = hi - lo
    for v in arrays:
        v[:, -1] = (v[:, -1] - lo).astype(np.float64) / span
    return pred


def pairwise_overlap(pred_xyxy, gt_xyxy):
    """(N,G) IoU with the +1 convention of ``bbox_overlap`` (widerface.py:39-52)."""
    p, g = pred_xyxy[:, None, :], gt_xyxy[None, :, :]
    w = np.minimum(p[..., 2], g[..., 2]) - np.maximum(p[..., 0], g[..., 0]) + 1
    h = np.minimum(p[..., 3], g[..., 3]) - np.maximum(p[..., 1], g[..., 1]) + 1
    inter = w * h
    parea = (p[..., 2] - p[..., 0] + 1) * (p[..., 3] - p[..., 1] + 1)
    garea = (g[..., 2] - g[..., 0] + 1) * (g[..., 3] - g[..., 1] + 1)
    o = inter / (garea + parea - inter)
    o[w <= 0] = 0
    o[h <= 0] = 0
    return o


def image_eval(pred, gt, ignore, iou_thresh):
    """Single-image matching (widerface.py:183-220) without the sequential loop.

    ``ignore[g] == 1`` marks a box of the evaluated difficulty (the reference's naming); predictions
    whose best box is not in it get ``proposal = -1``.  ``pred_recall[h]`` = number of distinct kept
    boxes matched by predictions ``0..h``."""
    n = pred.shape[0]
    _pred = pred.copy()                      # native dtype, like the reference (float32 detections
    _gt = gt.copy()                          # stay float32 until they meet the float64 boxes)
    _pred[:, 2] = _pred[:, 2] + _pred[:, 0]
    _pred[:, 3] = _pred[:, 3] + _pred[:, 1]
    _gt[:, 2] = _gt[:, 2] + _gt[:, 0]
    _gt[:, 3] = _gt[:, 3] + _gt[:, 1]
    ov = pairwise_overlap(_pred[:, :4], _gt)
    max_idx = ov.argmax(1)
    hit = ov[np.arange(n), max_idx] >= iou_thresh
    kept = np.asarray(ignore)[max_idx] != 0
    proposal_list = np.ones(n)
    proposal_list[hit & ~kept] = -1
    # first prediction that reaches each kept box switches it on
    first = np.zeros(n)
    cand = np.where(hit & kept)[0]
    if cand.size:
        _, first_pos = np.unique(max_idx[cand], return_index=True)
        first[cand[first_pos]] = 1
    pred_recall = np.cumsum(first)
    return pred_recall, proposal_list


def img_pr_info(thresh_num, pred_info, proposal_list, pred_recall):
    """Per-image (valid predictions, recalled boxes) at each score threshold (widerface.py:223-243)."""
    t = np.arange(thresh_num)
    thresh = 1 - (t + 1) / thresh_num
    scores = pred_info[:, 4]
    above = scores[None, :] >= thresh[:, None]                # (T, N)
    any_above = above.any(1)
    n = scores.shape[0]
    r_index = n - 1 - np.argmax(above[:, ::-1], axis=1)       # last index with score >= thresh
    valid_cum = np.cumsum(proposal_list == 1)
    pr_info = np.zeros((thresh_num, 2))
    pr_info[any_above, 0] = valid_cum[r_index[any_above]]
    pr_info[any_above, 1] = pred_recall[r_index[any_above]]
    return pr_info


def dataset_pr_info(thresh_num, pr_curve, count_face):
    out = np.zeros((thresh_num, 2))
    with np.errstate(divide='ignore', invalid='ignore'):
        out[:, 0] = pr_curve[:, 1] / pr_curve[:, 0]
    out[:, 1] = pr_curve[:, 1] / count_face
    return out


def voc_ap(rec, prec):
    mrec = np.concatenate(([0.], rec, [1.]))
    mpre = np.concatenate(([0.], prec, [0.]))
    mpre = np.maximum.accumulate(mpre[::-1])[::-1]            # precision envelope
    i = np.where(mrec[1:] != mrec[:-1])[0]
    return np.sum((mrec[i + 1] - mrec[i]) * mpre[i + 1])


def wider_evaluation(pred, gt, iou_thresh=0.5, thresh_num=THRESH_NUM, return_curves=False):
    """``[AP_easy, AP_medium, AP_hard]`` for predictions ``{event: {image: (N,5) [x,y,w,h,score]}}``
    (descending score, as the detector emits them) against the in-memory ground truth ``gt``."""
    pred = norm_score(pred)
    aps, curves = [], []
    for setting in SETTINGS:
        count_face = 0
        pr_curve = np.zeros((thresh_num, 2))
        for event, images in gt.items():
            pred_list = pred[event]
            for name, g in images.items():
                pred_info = pred_list[name]
                gt_boxes = np.asarray(g['boxes']).astype('float').reshape(-1, 4)
                keep_index = np.asarray(g[setting]).reshape(-1).astype(np.int64)
                count_face += len(keep_index)
                if len(gt_boxes) == 0 or len(pred_info) == 0:
                    continue
                ignore = np.zeros(gt_boxes.shape[0], dtype=np.int64)
                if len(keep_index) != 0:
                    ignore[keep_index - 1] = 1
                pred_recall, proposal_list = image_eval(pred_info, gt_boxes, ignore, iou_thresh)
                pr_curve += img_pr_info(thresh_num, pred_info, proposal_list, pred_recall)
        curve = dataset_pr_info(thresh_num, pr_curve, count_face)
        aps.append(voc_ap(curve[:, 1], curve[:, 0]))
        curves.append(curve)
    return (aps, curves) if return_curves else aps


def detections_to_results(results, event, image, dets_xyxy_score):
    """Store one image's detector output the way ``tools/test_widerface.py:149-160`` does
    (``[x1,y1,x2,y2,s]`` -> ``[x,y,w,h,s]``)."""
    xywh = np.array(dets_xyxy_score, copy=True).reshape(-1, 5)
    xywh[:, 2] = xywh[:, 2] - xywh[:, 0]
    xywh[:, 3] = xywh[:, 3] - xywh[:, 1]
    results.setdefault(event, {})[image] = xywh
    return results


def prediction_file_text(event, image_file, dets_xyxy_score):
    """The per-image ``.txt`` of ``tools/test_widerface.py:161-173`` (official submission format)."""
    d = np.asarray(dets_xyxy_score).reshape(-1, 5)
    lines = ['%s' % '/'.join([event, image_file]), '%d' % d.shape[0]]
    lines += ['%.5f %.5f %.5f %.5f %g' % (b[0], b[1], b[2] - b[0], b[3] - b[1], b[4]) for b in d]
    return '\n'.join(lines) + '\n'


def load_wider_gt(gt_dir):
    """In-memory ground truth from the four WIDER ``.mat`` files (``get_gt_boxes``, widerface.py:63-81)."""
    from scipy.io import loadmat
    gt_mat = loadmat(os.path.join(gt_dir, 'wider_face_val.mat'))
    keep = {s: loadmat(os.path.join(gt_dir, f'wider_{s}_val.mat'))['gt_list'] for s in SETTINGS}
    out = {}
    for i in range(len(gt_mat['event_list'])):
        event = str(gt_mat['event_list'][i][0][0])
        out[event] = {}
        for j in range(len(gt_mat['file_list'][i][0])):
            name = str(gt_mat['file_list'][i][0][j][0][0])
            out[event][name] = dict(boxes=gt_mat['face_bbx_list'][i][0][j][0].astype('float'),
                                    **{s: keep[s][i][0][j][0] for s in SETTINGS})
    return out


# --------------------------------------------------------------------------- test-time drivers
def prepare_test_image(img_u8, mode=0, divisor=32):
    """Host side of the reference's test pipeline (``configs/yunet_n.py:59-78`` as patched by
    ``tools/test_widerface.py:76-96``): mode 0 = keep-ratio resize into 640x640, mode > 30 = into
    (mode, mode), each followed by a zero pad to that size; mode 2 = original size padded to a
    multiple of ``divisor``; mode 1 = keep-ratio resize into 1100 x 1650 (short x long side), padded
    to the next multiple of ``divisor`` (the reference pads to exactly 1100 x 1650, a shape its own
    TFPN cannot add: 1100 / 8 = 137 vs 2 * (1100 // 16) = 136).  Returns the float32 CHW image (BGR,
    0..255) and the ``[w_scale, h_scale, w_scale, h_scale]`` float32 factor detections are divided by.

    Resizing uses ``cv2.resize`` on the uint8 image exactly like ``mmcv.imrescale`` (cv2 backend)."""
    h, w = img_u8.shape[:2]
    if mode == 1:
        scale = min(1650 / max(h, w), 1100 / min(h, w))                 # mmcv.rescale_size((1100, 1650))
        new_w, new_h = int(w * float(scale) + 0.5), int(h * float(scale) + 0.5)
        import cv2
        img = cv2.resize(img_u8, (new_w, new_h), interpolation=cv2.INTER_LINEAR)
        factor = np.array([new_w / w, new_h / h, new_w / w, new_h / h], dtype=np.float32)
        out_h, out_w = -(-new_h // divisor) * divisor, -(-new_w // divisor) * divisor
    elif mode == 2:
        out_h, out_w = -(-h // divisor) * divisor, -(-w // divisor) * divisor
        img, factor = img_u8, np.ones(4, np.float32)
    else:
        side = 640 if mode == 0 else int(mode)
        scale = min(side / max(h, w), side / min(h, w))                 # mmcv.rescale_size
        new_w, new_h = int(w * float(scale) + 0.5), int(h * float(scale) + 0.5)
        import cv2
        img = cv2.resize(img_u8, (new_w, new_h), interpolation=cv2.INTER_LINEAR)
        w_scale, h_scale = new_w / w, new_h / h
        factor = np.array([w_scale, h_scale, w_scale, h_scale], dtype=np.float32)
        out_h = out_w = -(-side // divisor) * divisor                   # Pad(size=(side, side))
    chw = np.zeros((3, out_h, out_w), np.float32)
    chw[:, :img.shape[0], :img.shape[1]] = img.transpose(2, 0, 1)
    return chw, factor


def engine_detector(engine, score_thr=0.02, iou_thr=0.45):
    """``detect(chw_float32, factor) -> (N,5)`` on the GPU engine (boxes rescaled by ``factor``)."""
    import torch

    def detect(chw, factor):
        img = torch.from_numpy(chw).to(engine.device)[None]
        sf = torch.from_numpy(np.asarray(factor, np.float32)).to(engine.device)[None]
        dets, counts, _ = engine.detect(img, score_thr, iou_thr, scale_factors=sf)
        n = int(counts[0])
        return dets[0, :n].cpu().numpy()
    return detect


def evaluate_wider(detect, samples, gt, mode=0, iou_thresh=0.5):
    """``samples``: iterable of ``(event, image_name, uint8 BGR image)``; ``detect`` as returned by
    ``engine_detector``.  Returns ``([AP_easy, AP_medium, AP_hard], results)``."""
    results = {}
    for event, name, img in samples:
        chw, factor = prepare_test_image(img, mode)
        detections_to_results(results, event, name, detect(chw, factor))
    for event, images in gt.items():                    # images the detector never saw count as empty
        for name in images:
            results.setdefault(event, {}).setdefault(name, np.zeros((0, 5), np.float32))
    import copy
    return wider_evaluation(copy.deepcopy(results), gt, iou_thresh), results
