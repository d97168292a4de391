"""CPU oracle for the region-proposal / target-assignment / NMS hot path.

TEST INFRASTRUCTURE ONLY.  This module is a numpy restatement of the arithmetic
of the reference (Kelicious/faster_rcnn, pure Python + numpy).  It is imported
only by ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py``; the product package
``faster_rcnn_b200`` never imports it and has no CPU fallback.

Parity pin: the numpy half of the reference imports and runs in the authoring
container, so every function here is checked (a) live against the imported
reference (``tests/test_oracle_vs_reference.py``, skipped where
``/root/reference`` is absent) and (b) against fixtures generated FROM the
reference by ``tests/golden/make_golden.py`` and committed under
``tests/golden/``.  The RoI layer restatement lives in ``roi_oracle.py`` and is
"parity unpinned" (TensorFlow 1.3 is not available), see its header.

All citations are ``file:line`` under ``/root/reference/faster_rcnn``.

Dtype notes (numpy >= 2 promotion, verified by probe): ``int16/int16 -> f64``,
``int16*int16 -> int16``, ``np.maximum(int16_array, f32_scalar) -> f32``,
``f32_scalar * int16_scalar -> f32``, ``(int16+int16)/2 -> f64``.
"""
import math
import random as _py_random

import numpy as np

# shared_constants.py:5  (float32 on purpose: it decides the dtype of `regr / MULT`)
BBREG_MULT = np.array([10, 10, 5, 5], dtype=np.float32)

# rpn_util.py:11-15
RPN_POS_IOU = 0.7
RPN_NEG_IOU = 0.3
RPN_BATCH = 256
RPN_MAX_POS = 128
# det_util.py:7-8
DET_MIN_IOU = 0.1
DET_POS_IOU = 0.5


# --------------------------------------------------------------------------
# anchors and geometry
# --------------------------------------------------------------------------
def anchor_table(scales=(16, 32, 64, 128, 256, 512), ratios=((1, 1), (1, 2), (2, 1))):
    """[height, width] per anchor, scale-major / ratio-minor.  util.py:242-253."""
    rows, norm = [], []
    for s in scales:
        for rh, rw in ratios:
            rows.append([s * rh, s * rw])
            norm.append(math.sqrt(s * rh * s * rw) / s)
    return (np.array(rows) // np.array(norm)[:, None]).astype(int)


def conv_dims_resnet(height, width):
    """conv4 output size of the ResNet base.  resnet.py:78-93."""
    out = []
    for d in (height, width):
        d += 6
        for f in (7, 3, 1, 1):
            d = (d - f) // 2 + 1
        out.append(d)
    return out


def conv_dims_vgg(height, width):
    """vgg.py:60-61 (STRIDE = 16)."""
    return height // 16, width // 16


def feature_anchors(rows, cols, dims_hw):
    """Feature-space anchors, centre = cell index.  det_util.py:162-175.

    Returns (rows*cols*A, 4) float32 [x1,y1,x2,y2], flat index (y*cols+x)*A+a.
    """
    dims_hw = np.asarray(dims_hw)
    out = np.zeros((rows, cols, len(dims_hw), 4), dtype=np.float32)
    xs, ys = np.meshgrid(np.arange(cols), np.arange(rows))
    for a, (ah, aw) in enumerate(dims_hw):
        out[:, :, a, 0] = xs - aw // 2
        out[:, :, a, 1] = ys - ah // 2
        out[:, :, a, 2] = out[:, :, a, 0] + aw
        out[:, :, a, 3] = out[:, :, a, 1] + ah
    return out.reshape(-1, 4)


def decode_boxes(boxes, deltas):
    """float32 delta decode with the reference's op order.  util.py:111-142.

    `boxes` (N,4) f32 is consumed (mutated) like the reference does.
    """
    b = boxes
    b[:, 2] -= b[:, 0]                       # w
    b[:, 3] -= b[:, 1]                       # h
    b[:, 0] += b[:, 2] / 2                   # cx
    b[:, 1] += b[:, 3] / 2                   # cy
    b[:, 0] += deltas[:, 0] * b[:, 2]
    b[:, 1] += deltas[:, 1] * b[:, 3]
    b[:, 2] *= np.exp(deltas[:, 2])
    b[:, 3] *= np.exp(deltas[:, 3])
    b[:, 0] -= b[:, 2] / 2
    b[:, 1] -= b[:, 3] / 2
    np.round(b, out=b)                       # half-to-even on x, y, w, h
    b[:, 2] += b[:, 0]
    b[:, 3] += b[:, 1]
    return b


def sanitize_boxes(cols, rows, b):
    """min 1-cell size, then clip to the map.  det_util.py:179-192 (this order)."""
    b[:, 2] = np.maximum(b[:, 0] + 1, b[:, 2])
    b[:, 3] = np.maximum(b[:, 1] + 1, b[:, 3])
    b[:, 0] = np.maximum(0, b[:, 0])
    b[:, 1] = np.maximum(0, b[:, 1])
    b[:, 2] = np.minimum(cols - 1, b[:, 2])
    b[:, 3] = np.minimum(rows - 1, b[:, 3])
    return b


def valid_box_indices(b):
    """det_util.py:196-205."""
    return np.where((b[:, 2] > b[:, 0]) & (b[:, 3] > b[:, 1]))[0]


def proposals_from_rpn(regr_out, anchor_dims, stride):
    """regr_out (1,R,C,4A) f32 -> (R*C*A,4) f32 boxes.  det_util.py:370-380."""
    rows, cols = regr_out.shape[1:3]
    anc = feature_anchors(rows, cols, np.asarray(anchor_dims) // stride)
    deltas = regr_out[0].reshape(-1, 4) / BBREG_MULT
    return sanitize_boxes(cols, rows, decode_boxes(anc, deltas))


def topk_proposals(boxes, probs, k, stable=True):
    """valid filter + descending score sort + truncate + int16 cast.

    det_util.py:68-76 (k=12000) and :147-155 (k=8000).  The reference uses the
    default (unstable) argsort; `stable=True` is the total order this repo
    defines for ties: descending score, ties by DESCENDING original index
    (what `argsort(kind='stable')[::-1]` yields).  Returns (boxes i16, probs,
    original flat indices).
    """
    v = valid_box_indices(boxes)
    b, p = boxes[v], probs[v]
    order = (p.argsort(kind='stable') if stable else p.argsort())[::-1][:k]
    return b[order].astype('int16'), p[order], v[order]


def greedy_nms(boxes, probs, overlap_thresh=0.7, max_boxes=300, stable=True):
    """Greedy NMS with the +1 area convention.  det_util.py:209-256.

    Returns the pick list (indices into `boxes`, pick order = descending score).
    Arithmetic follows the input dtype exactly as numpy does in the reference:
    int16 boxes -> int16 areas, int16/int16 -> float64 ratio; float64 boxes ->
    all float64.  `stable=True`: candidates are visited by (score desc,
    position desc), i.e. `argsort(kind='stable')` consumed from the back.
    """
    if len(boxes) == 0:
        return np.zeros(0, dtype=np.int64)
    x1, y1, x2, y2 = boxes[:, 0], boxes[:, 1], boxes[:, 2], boxes[:, 3]
    area = (x2 - x1 + 1) * (y2 - y1 + 1)
    queue = np.argsort(probs, kind='stable') if stable else np.argsort(probs)
    picked = []
    while queue.size:
        i, rest = queue[-1], queue[:-1]
        picked.append(i)
        iw = np.maximum(0, np.minimum(x2[i], x2[rest]) - np.maximum(x1[i], x1[rest]) + 1)
        ih = np.maximum(0, np.minimum(y2[i], y2[rest]) - np.maximum(y1[i], y1[rest]) + 1)
        inter = iw * ih
        ratio = inter / (area[i] + area[rest] - inter)
        queue = rest[np.where(ratio <= overlap_thresh)[0]]
        if len(picked) >= max_boxes:
            break
    return np.asarray(picked, dtype=np.int64)


def nms(boxes, probs, overlap_thresh=0.7, max_boxes=300, stable=True):
    """Reference-shaped wrapper: (boxes[pick], probs[pick]); [] on empty input
    (det_util.py:220-221)."""
    if len(boxes) == 0:
        return []
    pick = greedy_nms(boxes, probs, overlap_thresh, max_boxes, stable)
    return boxes[pick], probs[pick]


# --------------------------------------------------------------------------
# IoU matrix and regression parameters
# --------------------------------------------------------------------------
def iou_matrix(boxes1, boxes2):
    """(N,G) float32 IoU, NO +1 convention.  util.py:146-177.

    Op order matters for bit parity: union = (area1 + area2[g]) - inter.
    """
    out = np.zeros((len(boxes1), len(boxes2)), dtype=np.float32)
    a1 = (boxes1[:, 2] - boxes1[:, 0]) * (boxes1[:, 3] - boxes1[:, 1])
    a2 = (boxes2[:, 2] - boxes2[:, 0]) * (boxes2[:, 3] - boxes2[:, 1])
    for g, gt in enumerate(boxes2):
        w = np.maximum(0, np.minimum(boxes1[:, 2], gt[2]) - np.maximum(boxes1[:, 0], gt[0]))
        h = np.maximum(0, np.minimum(boxes1[:, 3], gt[3]) - np.maximum(boxes1[:, 1], gt[1]))
        inter = w * h
        out[:, g] = inter / (a1 + a2[g] - inter)
    return out


def regression_params(anchor, target):
    """(tx,ty,tw,th) that map `anchor` onto `target`.  util.py:180-206.

    Scalars keep their numpy types, so promotion is whatever numpy does for the
    caller's dtypes (see module docstring).
    """
    gx1, gy1, gx2, gy2 = target
    ax1, ay1, ax2, ay2 = anchor
    gcx, gcy = (gx2 + gx1) / 2.0, (gy2 + gy1) / 2.0
    gw, gh = gx2 - gx1, gy2 - gy1
    acx, acy = (ax2 + ax1) / 2.0, (ay2 + ay1) / 2.0
    aw, ah = ax2 - ax1, ay2 - ay1
    return (gcx - acx) / aw, (gcy - acy) / ah, np.log(gw / aw), np.log(gh / ah)


# --------------------------------------------------------------------------
# RPN training targets
# --------------------------------------------------------------------------
def pixel_anchors(rows, cols, dims_hw, stride):
    """Pixel-space anchors, centre = int(stride*(cell+0.5)).  rpn_util.py:276-298
    (+ :160-166, :184-189).  (N,4) float32."""
    dims_hw = np.asarray(dims_hw)
    n_a = len(dims_hw)
    flat = np.arange(rows * cols * n_a)
    per_row = cols * n_a
    cy, rem = flat // per_row, flat % per_row
    cx, a = rem // n_a, rem % n_a
    px = (stride * (cx + 0.5)).astype('int32')
    py = (stride * (cy + 0.5)).astype('int32')
    ah, aw = dims_hw[a, 0], dims_hw[a, 1]
    out = np.zeros((len(flat), 4), dtype=np.float32)
    out[:, 0] = px - aw // 2
    out[:, 1] = py - ah // 2
    out[:, 2] = out[:, 0] + aw
    out[:, 3] = out[:, 1] + ah
    return out


def out_of_bounds_indices(anchors, img_w, img_h):
    """rpn_util.py:302-310."""
    bad = (anchors[:, 0] < 0) | (anchors[:, 1] < 0) | (anchors[:, 2] >= img_w) | (anchors[:, 3] >= img_h)
    return np.where(bad)[0]


def label_anchors(img_w, img_h, gt_f32, rows, cols, dims_hw, stride):
    """Anchor labels before sampling.  rpn_util.py:54-103.

    gt_f32: (G,4) float32 pixel-space GT corners (util.get_bbox_coords output).
    Returns can_use (N,) bool, is_pos (N,) bool, bbreg (N,4) float32.
    """
    dims_hw = np.asarray(dims_hw)
    n_a = len(dims_hw)
    n = rows * cols * n_a
    bbreg = np.zeros((n, 4), dtype=np.float32)
    can_use = np.zeros(n, dtype=bool)
    is_pos = np.zeros(n, dtype=bool)

    anc = pixel_anchors(rows, cols, dims_hw, stride)
    oob = out_of_bounds_indices(anc, img_w, img_h)
    iou = iou_matrix(anc, gt_f32)
    best_iou_a = np.amax(iou, axis=1)
    best_gt_a = np.argmax(iou, axis=1)
    best_iou_g = np.amax(iou, axis=0)
    best_anchor_g = np.argmax(iou, axis=0)

    pos = np.where(best_iou_a > RPN_POS_IOU)[0]
    extra = best_anchor_g[np.where(best_iou_g > 0.0)]
    pos = np.unique(np.concatenate((pos, extra)))
    can_use[pos] = 1
    is_pos[pos] = 1

    for i in pos:
        y, r = divmod(int(i), cols * n_a)
        x, a = divmod(r, n_a)
        cx, cy = int(stride * (x + 0.5)), int(stride * (y + 0.5))     # rpn_util.py:169-180
        ah, aw = dims_hw[a]
        ax1, ay1 = cx - aw // 2, cy - ah // 2                          # shapes.py:309-323
        corners = np.array([ax1, ay1, ax1 + aw, ay1 + ah])
        bbreg[i, :] = BBREG_MULT * regression_params(corners, gt_f32[best_gt_a[i]])

    neg = np.where(np.logical_and(is_pos == 0, best_iou_a < RPN_NEG_IOU))[0]
    can_use[neg] = 1
    can_use[oob] = 0
    return can_use, is_pos, bbreg


def sample_rpn(is_pos, can_use, rng=_py_random):
    """256-anchor mini-batch balancing with Python's `random`.  rpn_util.py:324-350.
    Mutates and returns `can_use`."""
    pos = np.where(np.logical_and(is_pos == 1, can_use == 1))[0]
    neg = np.where(np.logical_and(is_pos == 0, can_use == 1))[0]
    n_pos, n_neg = len(pos), len(neg)
    if n_pos > RPN_MAX_POS:
        off = rng.sample(range(n_pos), n_pos - RPN_MAX_POS)
        can_use[pos[off]] = 0
        n_pos = RPN_MAX_POS
    if n_neg + n_pos > RPN_BATCH:
        off = rng.sample(range(n_neg), n_neg + n_pos - RPN_BATCH)
        can_use[neg[off]] = 0
    return can_use


def pack_rpn_targets(can_use, is_pos, bbreg, rows, cols, n_a):
    """Keras y_true layouts.  rpn_util.py:125-140.
    y_class (1,R,C,2A) bool = [can_use | is_pos]; y_bbreg (1,R,C,8A) f32 =
    [repeat(is_pos & can_use, 4) | targets]."""
    ip = is_pos.reshape(rows, cols, n_a)
    cu = can_use.reshape(rows, cols, n_a)
    y_class = np.concatenate([cu, ip], axis=2)[None]
    sel = np.repeat(np.logical_and(ip, cu), 4, axis=2)
    y_bbreg = np.concatenate([sel, bbreg.reshape(rows, cols, 4 * n_a)], axis=2)[None]
    return y_class, y_bbreg


# --------------------------------------------------------------------------
# detector training targets
# --------------------------------------------------------------------------
def label_rois(rois_i16, gt_f64, gt_cls_idx, n_classes):
    """RoI x GT labelling before sampling.  det_util.py:310-366.

    rois_i16: (n,4) int16 (feature units); gt_f64: (G,4) float64 GT corners
    already divided by the stride (python-float precision, det_util.py:312);
    gt_cls_idx: (G,) class index of each GT; n_classes counts 'bg' (last).
    Returns eligible_rois (m,4) i16, y_class (m,K) int32, y_transform (m,8(K-1)) f32.
    """
    gt_f32 = np.zeros((len(gt_f64), 4), dtype=np.float32)
    gt_f32[:] = gt_f64                                           # util.py:229-239
    iou = iou_matrix(rois_i16, gt_f32)
    best = np.amax(iou, axis=1)
    best_gt = np.argmax(iou, axis=1)
    elig = np.where(best >= DET_MIN_IOU)[0]
    is_pos = best[elig] >= DET_POS_IOU
    k_fg = n_classes - 1
    y_class = np.zeros((len(elig), n_classes), dtype=np.int32)
    labels = np.zeros((len(elig), 4 * k_fg), dtype=np.float32)
    targs = np.zeros((len(elig), 4 * k_fg), dtype=np.float32)
    for row, (i, pos) in enumerate(zip(elig, is_pos)):
        if not pos:
            y_class[row, n_classes - 1] = 1
            continue
        c = int(gt_cls_idx[best_gt[i]])
        y_class[row, c] = 1
        labels[row, 4 * c:4 * c + 4] = 1
        targs[row, 4 * c:4 * c + 4] = regression_params(rois_i16[i], gt_f64[best_gt[i]])
        targs[row, 4 * c:4 * c + 4] *= BBREG_MULT               # f32 * f32 (det_util.py:351-352)
    return rois_i16[elig], y_class, np.concatenate([labels, targs], axis=1)


def sample_det(is_pos, num_rois, rng=np.random):
    """64-RoI mini-batch (<=25% positives) with numpy's legacy global RNG.
    det_util.py:260-306.  Returns a python list, positives first."""
    want_pos = num_rois // 4
    pos = np.where(is_pos)[0]
    neg = np.where(np.logical_not(is_pos))[0]
    if len(pos) == 0:
        sel_pos = []
    elif len(pos) < want_pos:
        sel_pos = pos.tolist()
    else:
        sel_pos = rng.choice(pos, want_pos, replace=False).tolist()
    want_neg = num_rois - len(sel_pos)
    if len(neg) == 0:
        sel_neg = []
    elif len(neg) < want_neg:
        sel_neg = rng.choice(neg, want_neg, replace=True).tolist()
    else:
        sel_neg = rng.choice(neg, want_neg, replace=False).tolist()
    if len(sel_neg) == 0 and len(pos) > 0:
        sel_neg = np.tile(pos, want_neg // len(pos) + 1)[:want_neg].tolist()
    return sel_pos + sel_neg


# --------------------------------------------------------------------------
# detector post-processing (inference)
# --------------------------------------------------------------------------
def decode_scalar(box, deltas):
    """Unrounded, unclipped scalar decode.  util.py:55-74.  numpy scalar types
    of the inputs decide the precision of each op (tx*wa is f32 for f32 x i16)."""
    x1, y1, x2, y2 = box
    cxa, cya = (x1 + x2) / 2, (y1 + y2) / 2
    wa, ha = x2 - x1, y2 - y1
    tx, ty, tw, th = deltas
    cx, cy = tx * wa + cxa, ty * ha + cya
    w, h = math.exp(tw) * wa, math.exp(th) * ha
    x, y = cx - w / 2, cy - h / 2
    return x, y, x + w, y + h


def det_postprocess(rois_i16, out_cls, out_reg, bg_idx, stride, resize_ratio,
                    det_threshold=0.0, nms_thresh=0.5, max_boxes=2000, stable=True):
    """Per-class post-processing of detector outputs.  voc_dets.py:51-86.

    rois_i16 (M,4), out_cls (M,K) f32, out_reg (M,4(K-1)) f32 are the rows the
    detector saw (M = padded multiple of 64, duplicates included,
    voc_dets.py:42-46).  Returns a list of (class_idx, int bbox[4], prob) in the
    reference's order: classes by first appearance, rows in NMS pick order.
    """
    by_cls = {}
    for r in range(len(rois_i16)):
        c = int(np.argmax(out_cls[r]))
        conf = out_cls[r, c]
        if c == bg_idx or conf < det_threshold:
            continue
        t = out_reg[r, 4 * c:4 * c + 4] / BBREG_MULT
        x1, y1, x2, y2 = rois_i16[r]
        p = decode_scalar([x1, y1, x2, y2], t)
        by_cls.setdefault(c, ([], []))
        by_cls[c][0].append([stride * v for v in p])
        by_cls[c][1].append(conf)
    dets = []
    for c, (bb, pp) in by_cls.items():
        bb, pp = np.array(bb), np.array(pp)
        pick = greedy_nms(bb, pp, nms_thresh, max_boxes, stable)
        for i in pick:
            dets.append((c, np.array([int(round(v / resize_ratio)) for v in bb[i]]), pp[i]))
    return dets
