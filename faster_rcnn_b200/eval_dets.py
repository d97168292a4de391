"""Drop-in for the reference's `eval_dets.py` (VOC07 detection evaluation) and the detection-file writer of
`voc_dets.py:114-129` -- a widening row of the hot path.

The det<->GT matching (float64 IoU with the devkit's +1 convention, first-index arg-max, greedy "already
detected" marking) and the precision / recall / 11-point AP arithmetic run on the GPU (csrc/evalmatch.cu);
reading the annotation XML and the `comp3_det_test_<cls>.txt` files, and ordering the parsed detections by
confidence, stay on the host like any file IO.  Equal confidences are ordered stably (the reference's default
argsort leaves their order implementation-defined).
"""
import os
from xml.etree import ElementTree

import numpy as np

from . import ops
from .runtime import get_context

_THRESHOLDS_07 = np.arange(0., 1.1, 0.1)          # eval_dets.py:12 (the float values matter: 0.30000000000000004 ...)


def voc_ap(rec, prec, use_07_metric=False):
    """eval_dets.py:8-35 on host arrays (the GPU path of `voc_eval` computes the 07 metric itself)."""
    if use_07_metric:
        ap = 0.
        for t in _THRESHOLDS_07:
            p = 0 if np.sum(rec >= t) == 0 else np.max(prec[rec >= t])
            ap = ap + p / 11
        return ap
    mrec = np.concatenate(([0.], rec, [1.]))
    mpre = np.concatenate(([0.], prec, [0.]))
    for i in range(mpre.size - 1, 0, -1):
        mpre[i - 1] = np.maximum(mpre[i - 1], mpre[i])
    i = np.where(mrec[1:] != mrec[:-1])[0]
    return np.sum((mrec[i + 1] - mrec[i]) * mpre[i + 1])


def read_voc_objects(voc_path, imagename):
    """[(class name, [x1,y1,x2,y2] 0-based, difficult)] of one annotation file (data/voc_data_helpers.py:99-118)."""
    root = ElementTree.parse(os.path.join(voc_path, 'Annotations', imagename + '.xml')).getroot()
    out = []
    for obj in root.findall('object'):
        bb = obj.find('bndbox')
        box = [int(float(bb.find(k).text)) - 1 for k in ('xmin', 'ymin', 'xmax', 'ymax')]
        out.append((obj.find('name').text, box, int(obj.find('difficult').text) == 1))
    return out


def voc_eval_arrays(image_ids, confidence, boxes, gt_by_image, imagenames, ovthresh=0.5):
    """Core of voc_eval on in-memory data for ONE class.  image_ids: image name per detection; confidence (nd,);
    boxes (nd,4); gt_by_image: name -> (bbox (g,4), difficult (g,)); imagenames: the evaluated image set.
    Returns (rec, prec, ap) like the reference (eval_dets.py:66-125)."""
    ctx = get_context()
    index = {name: i for i, name in enumerate(imagenames)}
    gt_counts = np.zeros(len(imagenames) + 1, dtype=np.int64)
    gt_rows, gt_diff = [], []
    for i, name in enumerate(imagenames):
        bbox, difficult = gt_by_image.get(name, (np.zeros((0, 4)), np.zeros(0, bool)))
        bbox = np.asarray(bbox, dtype=np.float64).reshape(-1, 4)
        gt_counts[i + 1] = len(bbox)
        gt_rows.append(bbox)
        gt_diff.append(np.asarray(difficult, dtype=bool).reshape(-1))
    gt_offsets = np.cumsum(gt_counts).astype(np.int32)
    gt_boxes = np.concatenate(gt_rows) if gt_rows else np.zeros((0, 4))
    gt_difficult = np.concatenate(gt_diff) if gt_diff else np.zeros(0, bool)
    npos = float(np.sum(~gt_difficult))

    confidence = np.asarray(confidence, dtype=np.float64).reshape(-1)
    nd = len(confidence)
    if nd == 0:
        return np.zeros(0), np.zeros(0), 0.0
    order = np.argsort(-confidence, kind='stable')
    bb_sorted = np.asarray(boxes, dtype=np.float64).reshape(-1, 4)[order]
    img_of_rank = np.array([index[image_ids[i]] for i in order], dtype=np.int64)     # KeyError like the reference
    by_img = np.argsort(img_of_rank, kind='stable').astype(np.int32)                 # ranks grouped by image, ascending
    det_offsets = np.zeros(len(imagenames) + 1, dtype=np.int32)
    det_offsets[1:] = np.cumsum(np.bincount(img_of_rank, minlength=len(imagenames)))

    dummy_box = np.zeros((1, 4))
    tp, fp = ops.voc_match(ctx.to_device(bb_sorted), ctx.to_device(det_offsets), ctx.to_device(by_img),
                           ctx.to_device(gt_boxes if len(gt_boxes) else dummy_box),
                           ctx.to_device((gt_difficult if len(gt_difficult) else np.zeros(1, bool)).view(np.uint8)),
                           ctx.to_device(gt_offsets), ovthresh)
    rec, prec, ap = ops.voc_pr_ap(tp, fp, npos, ctx.to_device(_THRESHOLDS_07))
    return ctx.to_host(rec), ctx.to_host(prec), float(ctx.to_host(ap)[0])


def voc_eval(voc_path, det_file, imageset_path, cls_name, ovthresh=0.5):
    """Same signature and return value as the reference (eval_dets.py:38-125)."""
    with open(imageset_path, 'r') as f:
        imagenames = [line.strip() for line in f.readlines()]
    gt_by_image = {}
    for name in imagenames:
        objs = [o for o in read_voc_objects(voc_path, name) if o[0] == cls_name]
        gt_by_image[name] = (np.array([o[1] for o in objs], dtype=np.float64).reshape(-1, 4),
                             np.array([o[2] for o in objs], dtype=bool))
    with open(det_file, 'r') as f:
        split = [x.strip().split(' ') for x in f.readlines()]
    image_ids = [x[0] for x in split]
    confidence = np.array([float(x[1]) for x in split])
    boxes = np.array([[float(z) for z in x[2:]] for x in split]).reshape(-1, 4)
    return voc_eval_arrays(image_ids, confidence, boxes, gt_by_image, imagenames, ovthresh)


def get_voc_results_filename(dets_path, cls_name):
    return os.path.join(dets_path, 'comp3_det_test_{}.txt'.format(cls_name))


def write_dets(dets, out_dir):
    """voc_dets.py:114-129: one `comp3_det_test_<cls>.txt` per class, `image prob x1 y1 x2 y2` with +1 coordinates.
    `dets`: {cls_name: {image_name: [{'bbox', 'prob', ...}]}} as built by voc_dets.get_dets_by_cls."""
    os.makedirs(out_dir, exist_ok=True)
    for cls_name, cls_dets in dets.items():
        with open(get_voc_results_filename(out_dir, cls_name), 'w') as out:
            for image_name, image_dets in cls_dets.items():
                for det in image_dets:
                    x1, y1, x2, y2 = det['bbox'] + 1
                    out.write("{} {} {} {} {} {}\n".format(image_name, det['prob'], x1, y1, x2, y2))


def eval_all(dets_path, voc_path, class_mapping, img_set='val'):
    """eval_dets.py:134-152 without the printing: {cls_name: ap} and the mean AP."""
    aps = {}
    imageset_file = os.path.join(voc_path, 'ImageSets', 'Main', img_set + '.txt')
    for cls_name, _ in sorted(class_mapping.items()):
        if cls_name == 'bg':
            continue
        aps[cls_name] = voc_eval(voc_path, get_voc_results_filename(dets_path, cls_name), imageset_file, cls_name, 0.5)[2]
    return aps, float(np.mean(list(aps.values())))
