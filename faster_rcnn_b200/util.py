"""Drop-in for the box maths of the reference's `util.py` (numpy in, numpy out).

Array functions run on the GPU through libfrcnn_b200.so; the scalar helpers (`transform`,
`get_reg_params`, `calc_iou`, `get_anchors`, `get_bbox_coords`) are host value logic exactly as
in the reference -- inside the fused kernels the same formulas run on the device (label.cu,
postproc.cu).  Citations: file:line under /root/reference/faster_rcnn.
"""
import math

import numpy as np
import torch

from . import ops
from .runtime import get_context
from .shared_constants import DEFAULT_ANCHOR_RATIOS, DEFAULT_ANCHOR_SCALES, _anchor_table


def calc_iou(coords1, coords2):
    """IoU of two [x1,y1,x2,y2] boxes, no +1 convention (util.py:8-38)."""
    ax1, ay1, ax2, ay2 = coords1
    bx1, by1, bx2, by2 = coords2
    iw = min(ax2, bx2) - max(ax1, bx1)
    ih = min(ay2, by2) - max(ay1, by1)
    if iw <= 0 or ih <= 0:
        return 0.0
    inter = iw * ih
    union = (ax2 - ax1) * (ay2 - ay1) + (bx2 - bx1) * (by2 - by1) - inter
    return inter / union


def transform(anchor_coords, reg_targets):
    """Scalar delta decode, unrounded and unclipped (util.py:55-74)."""
    x1, y1, x2, y2 = anchor_coords
    tx, ty, tw, th = reg_targets
    cxa, cya = (x1 + x2) / 2, (y1 + y2) / 2
    wa, ha = x2 - x1, y2 - y1
    cx, cy = tx * wa + cxa, ty * ha + cya
    w, h = math.exp(tw) * wa, math.exp(th) * ha
    x, y = cx - w / 2, cy - h / 2
    return x, y, x + w, y + h


def transform_np_inplace(coords, reg_targets):
    """Decode (N,4) f32 boxes with (N,4) deltas IN PLACE on the GPU and return `coords`
    (util.py:111-142: float32, separate roundings, half-to-even round of x, y, w, h)."""
    if not (isinstance(coords, np.ndarray) and coords.dtype == np.float32 and coords.ndim == 2
            and coords.shape[1] == 4):
        raise TypeError("coords must be a float32 (N,4) numpy array")
    if len(coords) == 0:
        return coords
    ctx = get_context()
    dev = ctx.to_device(coords)
    deltas = ctx.to_device(np.asarray(reg_targets, dtype=np.float32).reshape(-1, 4))
    ops.box_transform_(dev, deltas)
    coords[...] = ctx.to_host(dev)
    return coords


def cross_ious(boxes1, boxes2):
    """(N,G) float32 IoU matrix on the GPU (util.py:146-177).  boxes1 int16 or float32."""
    boxes1, boxes2 = np.asarray(boxes1), np.asarray(boxes2, dtype=np.float32)
    if len(boxes1) == 0 or len(boxes2) == 0:
        return np.zeros((len(boxes1), len(boxes2)), dtype=np.float32)
    if boxes1.dtype != np.int16:
        boxes1 = boxes1.astype(np.float32, copy=False)
    ctx = get_context()
    return ctx.to_host(ops.cross_ious(ctx.to_device(boxes1), ctx.to_device(boxes2)))


def get_reg_params(anchor_coords, bbox_coords):
    """(tx,ty,tw,th) mapping an anchor onto a box (util.py:180-206).  Scalars keep their numpy
    types, so precision follows numpy's promotion exactly like the reference."""
    bx1, by1, bx2, by2 = bbox_coords
    ax1, ay1, ax2, ay2 = anchor_coords
    bcx, bcy = (bx2 + bx1) / 2.0, (by2 + by1) / 2.0
    bw, bh = bx2 - bx1, by2 - by1
    acx, acy = (ax2 + ax1) / 2.0, (ay2 + ay1) / 2.0
    aw, ah = ax2 - ax1, ay2 - ay1
    return (bcx - acx) / aw, (bcy - acy) / ah, np.log(bw / aw), np.log(bh / ah)


def resize_imgs(imgs, min_size=600, max_size=1000):
    """util.py:209-226 (host-side, not on the hot path)."""
    pairs = [img.resize_within_bounds(min_size=min_size, max_size=max_size) for img in imgs]
    return [p[0] for p in pairs], [p[1] for p in pairs]


def get_bbox_coords(gt_boxes):
    """(G,4) float32 array of the boxes' corners (util.py:229-239)."""
    out = np.zeros((len(gt_boxes), 4), dtype=np.float32)
    for i, box in enumerate(gt_boxes):
        out[i] = box.corners
    return out


def get_anchors(anchor_scales=DEFAULT_ANCHOR_SCALES, anchor_ratios=DEFAULT_ANCHOR_RATIOS):
    """[height, width] per anchor (util.py:242-253)."""
    return _anchor_table(anchor_scales, anchor_ratios)
