"""CPU oracle for the VOC detection evaluation -- TEST INFRASTRUCTURE ONLY.

Numpy restatement of eval_dets.voc_ap (eval_dets.py:8-35) and of the matching loop of eval_dets.voc_eval
(eval_dets.py:66-125) on in-memory arrays.  Pinned: live against the unmodified reference function run on
its own VOC_test annotations (tests/test_oracle_vs_reference.py) and against golden vectors it produced
(tests/golden/voc_eval.npz, tests/golden/make_golden.py)."""
import numpy as np


def voc_ap(rec, prec, use_07_metric=False):
    """eval_dets.py:8-35."""
    if use_07_metric:
        ap = 0.
        for t in np.arange(0., 1.1, 0.1):
            p = 0 if np.sum(rec >= t) == 0 else np.max(prec[rec >= t])
            ap = ap + p / 11
        return ap
    mrec = np.concatenate(([0.], rec, [1.]))
    mpre = np.concatenate(([0.], prec, [0.]))
    for i in range(mpre.size - 1, 0, -1):
        mpre[i - 1] = np.maximum(mpre[i - 1], mpre[i])
    i = np.where(mrec[1:] != mrec[:-1])[0]
    return np.sum((mrec[i + 1] - mrec[i]) * mpre[i + 1])


def voc_match(image_ids, confidence, boxes, gt_by_image, ovthresh=0.5, stable=True):
    """eval_dets.py:66-125.  image_ids: list of image names per detection; confidence (nd,) f64; boxes (nd,4) f64;
    gt_by_image: name -> (bbox (g,4) array, difficult (g,) bool array) for ONE class.  Returns rec, prec, ap (07)."""
    recs = {k: {'bbox': np.asarray(v[0], dtype=float).reshape(-1, 4), 'difficult': np.asarray(v[1], dtype=bool),
                'det': [False] * len(v[1])} for k, v in gt_by_image.items()}
    npos = sum(int(np.sum(~r['difficult'])) for r in recs.values())
    order = np.argsort(-confidence, kind='stable') if stable else np.argsort(-confidence)
    bb_sorted = boxes[order, :]
    ids = [image_ids[i] for i in order]
    nd = len(ids)
    tp, fp = np.zeros(nd), np.zeros(nd)
    for d in range(nd):
        r = recs[ids[d]]
        bb = bb_sorted[d, :].astype(float)
        ovmax, jmax = -np.inf, -1
        gt = r['bbox']
        if gt.size > 0:
            iw = np.maximum(np.minimum(gt[:, 2], bb[2]) - np.maximum(gt[:, 0], bb[0]) + 1., 0.)
            ih = np.maximum(np.minimum(gt[:, 3], bb[3]) - np.maximum(gt[:, 1], bb[1]) + 1., 0.)
            inters = iw * ih
            uni = ((bb[2] - bb[0] + 1.) * (bb[3] - bb[1] + 1.) + (gt[:, 2] - gt[:, 0] + 1.) * (gt[:, 3] - gt[:, 1] + 1.) - inters)
            overlaps = inters / uni
            ovmax, jmax = np.max(overlaps), np.argmax(overlaps)
        if ovmax > ovthresh:
            if not r['difficult'][jmax]:
                if not r['det'][jmax]:
                    tp[d] = 1.
                    r['det'][jmax] = 1
                else:
                    fp[d] = 1.
        else:
            fp[d] = 1.
    fp, tp = np.cumsum(fp), np.cumsum(tp)
    rec = tp / float(npos)
    prec = tp / np.maximum(tp + fp, np.finfo(np.float64).eps)
    return rec, prec, voc_ap(rec, prec, use_07_metric=True)
