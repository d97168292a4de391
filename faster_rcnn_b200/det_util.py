"""Drop-in for the reference's `det_util.py`: proposals, NMS and detector training targets on the GPU.

Same call signatures, array layouts and RNG stream as the reference (citations: file:line under
/root/reference/faster_rcnn).  Decode + sanitize + validity + top-k (proposals.cu), greedy NMS
(nms.cu) and RoI x GT labelling (label.cu) run in libfrcnn_b200.so; only `np.random.choice` of
the mini-batch sampling stays on the host (the reference's legacy global RNG stream).

Tie order: the reference sorts with numpy's default (unstable) argsort, so its order among equal
scores is implementation-defined.  This implementation defines it: descending score, ties by
descending position (what `argsort(kind='stable')` consumed from the back yields).
"""
import numpy as np
import torch

from . import _lib, ops
from .custom_decorators import profile
from .runtime import get_context
from .shared_constants import BBREG_MULTIPLIERS, DEFAULT_ANCHORS  # noqa: F401

CLASSIFIER_MIN_OVERLAP = 0.1   # det_util.py:7-10; compiled into label.cu as float32 constants
CLASSIFIER_POS_OVERLAP = 0.5
PROBABLE_THRESHOLD = 0.05

TRAIN_PRE_NMS_TOPK, TRAIN_POST_NMS = 12000, 2000     # det_util.py:73,77
TEST_PRE_NMS_TOPK, TEST_POST_NMS = 8000, 300         # det_util.py:153,156
RPN_NMS_THRESH = 0.7


class DetTrainingManager:
    """Generates detector (Fast R-CNN head) inputs for an image (reference: det_util.py:13-158).

    `rpn_model` is any object with `.output` (list; 3 entries = conv features are returned too)
    and `.predict_on_batch(batch) -> [cls_out (1,R,C,A), regr_out (1,R,C,4A)[, conv_out]]`
    yielding numpy arrays or CUDA torch tensors (tensors stay on the device)."""

    def __init__(self, rpn_model, class_mapping, preprocess_func, num_rois=64, stride=16, anchor_dims=DEFAULT_ANCHORS):
        self.rpn_model = rpn_model
        self.class_mapping = class_mapping
        self.preprocess_func = preprocess_func
        self.num_rois = num_rois
        self.stride = stride
        self.anchor_dims = anchor_dims
        self._cache = {}
        self.conv_only = True if len(rpn_model.output) == 3 else False

    @profile
    def batched_image(self, image):
        return np.expand_dims(self.preprocess_func(image.data), axis=0)

    @profile
    def _out_from_image(self, batched_img):
        return self.rpn_model.predict_on_batch(batched_img)

    def _rpn_outputs(self, image):
        outs = self._out_from_image(self.batched_image(image))
        conv_out = outs[2] if self.conv_only else None
        return outs[0], outs[1], conv_out

    @profile
    def _rois_from_image(self, image):
        """(roi_coords (N,4) f32, roi_probs (N,) f32, conv_out) -- det_util.py:44-55."""
        cls_out, regr_out, conv_out = self._rpn_outputs(image)
        return self._get_rois(_as_numpy(regr_out), self.anchor_dims), _as_numpy(cls_out).reshape((-1)), conv_out

    @profile
    def _get_rois(self, regr_out, anchor_dims):
        return _get_rois(regr_out, anchor_dims, self.stride)

    def _proposals(self, image, k, max_boxes):
        """RPN outputs -> NMS-ed proposals, fused on the device (decode never leaves the GPU)."""
        ctx = get_context()
        cls_out, regr_out, conv_out = self._rpn_outputs(image)
        rois, _, count = ops.proposals(_as_device(ctx, regr_out), _as_device(ctx, cls_out), self.anchor_dims,
                                       self.stride, k, RPN_NMS_THRESH, max_boxes)
        return rois, count, conv_out

    @profile
    def _process(self, image):
        """det_util.py:63-87: proposals (k=12000, NMS 0.7 -> 2000) then RoI x GT labelling; cached."""
        ctx = get_context()
        rois, count, conv_out = self._proposals(image, TRAIN_PRE_NMS_TOPK, TRAIN_POST_NMS)
        gt64, gt_cls = _gt_feature_boxes(image, self.class_mapping, self.stride)
        n_cls = len(self.class_mapping)
        out_rois, y_cls, y_tr, _, m = ops.label_rois(rois, ctx.to_device(gt64[None]), ctx.to_device(gt_cls[None]),
                                                     ctx.to_device(np.array([len(gt64)], dtype=np.int32)), n_cls,
                                                     n_roi=count)
        m = int(ctx.to_host(m)[0])
        cache_obj = {'rois': ctx.to_host(out_rois[0, :m]), 'y_class_num': ctx.to_host(y_cls[0, :m]),
                     'y_transform': ctx.to_host(y_tr[0, :m])}
        if conv_out is not None:
            cache_obj['conv_out'] = conv_out
        self._cache[image.cache_key] = cache_obj

    @profile
    def get_training_input(self, image):
        """(first_input, rois (1,64,4) i16, y_class_num (1,64,K) i32, y_transform (1,64,8(K-1)) f32) or
        4 x None when no RoI is eligible (det_util.py:90-133)."""
        if image.cache_key not in self._cache:
            self._process(image)
        results = self._cache[image.cache_key]
        if len(results['rois']) == 0:
            return None, None, None, None
        rois, y_class_num, y_transform = results['rois'], results['y_class_num'], results['y_transform']
        found_object = y_class_num[:, -1] == 0            # 'bg' is the last class
        sampled_idxs = _get_det_samples(found_object, self.num_rois)
        rois, y_class_num, y_transform = rois[sampled_idxs], y_class_num[sampled_idxs], y_transform[sampled_idxs]
        first_input = results['conv_out'] if self.conv_only else self.batched_image(image)
        if self.conv_only:
            del self._cache[image.cache_key]
        return first_input, np.expand_dims(rois, axis=0), np.expand_dims(y_class_num, axis=0), \
            np.expand_dims(y_transform, axis=0)

    @profile
    def get_det_inputs(self, image):
        """(conv_out, rois (<=300,4) i16) for inference: k=8000, NMS 0.7 -> 300 (det_util.py:136-158)."""
        ctx = get_context()
        rois, count, conv_out = self._proposals(image, TEST_PRE_NMS_TOPK, TEST_POST_NMS)
        n = int(ctx.to_host(count)[0])
        if n == 0:
            raise ValueError("no valid proposal (the reference fails unpacking nms()'s [] here too)")
        return conv_out, ctx.to_host(rois[0, :n])


def _as_numpy(x):
    return x.detach().cpu().numpy() if isinstance(x, torch.Tensor) else np.asarray(x)


def _as_device(ctx, x):
    if isinstance(x, torch.Tensor):
        return x.to(ctx.device, torch.float32)
    return ctx.to_device(x, np.float32)


def _gt_feature_boxes(image, class_mapping, stride):
    """GT corners in feature units as float64 (python-float precision of `gt_box.resize(1/stride)`,
    det_util.py:312) + class indices."""
    if class_mapping['bg'] != len(class_mapping) - 1:
        raise NotImplementedError("'bg' must be the last class index (the reference assumes it too: det_util.py:120)")
    boxes = [gt_box.resize(1 / stride) for gt_box in image.gt_boxes]
    if not boxes:
        raise ValueError("image without ground-truth boxes (the reference fails in np.amax here too)")
    gt64 = np.array([np.asarray(b.corners, dtype=np.float64) for b in boxes], dtype=np.float64).reshape(-1, 4)
    gt_cls = np.array([class_mapping[b.obj_cls] for b in boxes], dtype=np.int32)
    return gt64, gt_cls


# ---- module-level functions with the reference's names ------------------------------------------
@profile
def _get_anchor_coords(conv_rows, conv_cols, anchor_dims, multiplier=1):
    """(R,C,A,4) f32 feature-space anchors, generated on the GPU (det_util.py:162-175)."""
    ctx = get_context()
    dims = np.asarray(anchor_dims) * multiplier
    return ctx.to_host(ops.anchor_grid(dims, conv_rows, conv_cols)).reshape(conv_rows, conv_cols, len(dims), 4)


@profile
def _sanitize_boxes_inplace(conv_cols, conv_rows, coords):
    """min 1-cell size then clip to the map, IN PLACE on (N,4) f32 (det_util.py:179-192)."""
    if len(coords) == 0:
        return coords
    ctx = get_context()
    dev = ctx.to_device(coords, np.float32)
    ops.box_transform_(dev, None, sanitize=(conv_cols, conv_rows))
    coords[...] = ctx.to_host(dev)
    return coords


@profile
def _get_valid_box_idxs(boxes):
    """ascending indices of boxes with positive width and height (det_util.py:196-205)."""
    if len(boxes) == 0:
        return np.zeros(0, dtype=np.int64)
    ctx = get_context()
    index, count = ops.valid_boxes(ctx.to_device(boxes, np.float32))
    n = int(ctx.to_host(count)[0])
    return ctx.to_host(index[:n]).astype(np.int64)


@profile
def nms(boxes, probs, overlap_thresh=0.7, max_boxes=300):
    """Greedy NMS, +1 area convention (det_util.py:209-256).  Returns (boxes[pick], probs[pick]) in
    pick order (descending score); `[]` for empty input like the reference.  int16 boxes take the
    exact-integer RPN kernel, everything else the float64 kernel (identical to numpy for float64 and
    integer inputs)."""
    if len(boxes) == 0:
        return []
    ctx = get_context()
    boxes, probs = np.asarray(boxes), np.asarray(probs)
    n = len(boxes)
    scores = ctx.to_device(probs.reshape(1, n), np.float32)
    if probs.dtype != np.float32 and not np.array_equal(probs.astype(np.float32).astype(probs.dtype), probs):
        raise TypeError("probs must be exactly representable in float32")
    if boxes.dtype == np.int16:
        if n > _lib.NMS_MAX_UNSORTED and not bool(np.all(probs[:-1] > probs[1:])):
            raise ValueError("nms: more than %d int16 boxes need strictly descending scores" % _lib.NMS_MAX_UNSORTED)
        keep, count, _, _ = ops.nms_i16(ctx.to_device(boxes.reshape(1, n, 4)), scores, None, overlap_thresh, max_boxes)
        pick = ctx.to_host(keep[0, :int(ctx.to_host(count)[0])])
    else:
        if n > _lib.NMS_F64_MAX:
            raise ValueError("nms: at most %d non-int16 boxes per call" % _lib.NMS_F64_MAX)
        offs = ctx.to_device(np.array([0, n], dtype=np.int32))
        keep, count = ops.nms_f64(ctx.to_device(boxes.reshape(n, 4), np.float64), scores.reshape(n), offs, n,
                                  overlap_thresh, max_boxes)
        pick = ctx.to_host(keep[0, :int(ctx.to_host(count)[0])])
    return boxes[pick], probs[pick]


@profile
def _get_det_samples(is_pos, num_desired_rois):
    """64-RoI mini-batch, <= 25 % positives, numpy's legacy global RNG in the reference's call order
    (det_util.py:260-306).  Returns a python list, positives first."""
    want_pos = num_desired_rois // 4
    pos = np.where(is_pos)[0]
    neg = np.where(np.logical_not(is_pos))[0]
    if len(pos) == 0:
        chosen_pos = []
    elif len(pos) < want_pos:
        chosen_pos = pos.tolist()
    else:
        chosen_pos = np.random.choice(pos, want_pos, replace=False).tolist()
    want_neg = num_desired_rois - len(chosen_pos)
    if len(neg) == 0:
        chosen_neg = []
    else:
        chosen_neg = np.random.choice(neg, want_neg, replace=len(neg) < want_neg).tolist()
    if len(chosen_neg) == 0 and len(pos) > 0:
        chosen_neg = np.tile(pos, want_neg // len(pos) + 1)[:want_neg].tolist()
    return chosen_pos + chosen_neg


@profile
def _rois_to_truth(rois, image, class_mapping, stride=16):
    """RoI x GT labelling before sampling (det_util.py:310-334): (eligible_rois (m,4) i16,
    y_class_num (m,K) i32 one-hot, y_transform (m,8(K-1)) f32)."""
    ctx = get_context()
    rois = np.ascontiguousarray(rois, dtype=np.int16)
    gt64, gt_cls = _gt_feature_boxes(image, class_mapping, stride)
    k = len(class_mapping)
    if len(rois) == 0:
        return rois, np.zeros((0, k), np.int32), np.zeros((0, 8 * (k - 1)), np.float32)
    out_rois, y_cls, y_tr, _, m = ops.label_rois(ctx.to_device(rois[None]), ctx.to_device(gt64[None]),
                                                 ctx.to_device(gt_cls[None]),
                                                 ctx.to_device(np.array([len(gt64)], dtype=np.int32)), k)
    m = int(ctx.to_host(m)[0])
    return ctx.to_host(out_rois[0, :m]), ctx.to_host(y_cls[0, :m]), ctx.to_host(y_tr[0, :m])


@profile
def _one_hot_encode_cls(obj_classes, class_to_num):
    """(n,K) int32 one-hot (det_util.py:358-366); host helper kept for API compatibility -- the
    fused label_rois kernel writes this layout directly."""
    out = np.zeros((len(obj_classes), len(class_to_num)), dtype=np.int32)
    out[np.arange(len(obj_classes)), [class_to_num[c] for c in obj_classes]] = 1
    return out


@profile
def _one_hot_encode_bbreg(rois, gt_boxes, is_pos, class_mapping):
    """(n, 8(K-1)) f32 = [labels | targets] (det_util.py:338-354); host helper kept for API
    compatibility -- the fused label_rois kernel writes this layout directly."""
    from .util import get_reg_params
    k_fg = len(class_mapping) - 1
    labels = np.zeros((len(rois), 4 * k_fg), dtype=np.float32)
    targs = np.zeros((len(rois), 4 * k_fg), dtype=np.float32)
    for i, (roi, gt_box, pos) in enumerate(zip(rois, gt_boxes, is_pos)):
        if pos:
            c = class_mapping[gt_box.obj_cls]
            labels[i, 4 * c:4 * c + 4] = 1
            targs[i, 4 * c:4 * c + 4] = get_reg_params(roi, gt_box.corners)
            targs[i, 4 * c:4 * c + 4] *= BBREG_MULTIPLIERS
    return np.concatenate([labels, targs], axis=1)


@profile
def _get_rois(regr_out, anchor_dims, stride):
    """regr_out (1,R,C,4A) f32 -> every decoded + sanitized box (R*C*A,4) f32 (det_util.py:370-380)."""
    ctx = get_context()
    regr = _as_device(ctx, regr_out)
    _, rows, cols, a4 = regr.shape
    cls = ctx.empty((1, rows, cols, a4 // 4), torch.float32).zero_()
    dense = ops.decode_topk(regr[:1], cls, anchor_dims, stride, 1, want_dense=True)[4]
    return ctx.to_host(dense[0])
