"""Drop-in for the detection post-processing of the reference's `voc_dets.get_dets`
(voc_dets.py:20-88): RoI batching with the reference's padding rule, per-RoI arg-max class,
float64 box decode, per-class NMS and rescaling -- steps 2-4 on the GPU (postproc.cu).

`detector` is any object with `.predict([conv_out, batch_rois]) -> (out_cls (1,64,K),
out_reg (1,64,4(K-1)))`; the detector itself (dense layers) is not part of this package.
"""
import numpy as np
import torch

from . import ops
from .runtime import get_context

DEFAULT_DET_THRESHOLD = 0.0   # voc_dets.py:17


def pad_roi_batches(rois, num_rois=64):
    """Splits (n,4) rois into ceil(n/num_rois) batches; the last one is padded with copies of ITS
    first RoI (voc_dets.py:37-46).  Returns (n_batches*num_rois, 4)."""
    n = rois.shape[0]
    n_batches = -(-n // num_rois)
    pad = n_batches * num_rois - n
    if pad == 0:
        return rois
    first_of_last = rois[(n_batches - 1) * num_rois]
    return np.concatenate([rois, np.tile(first_of_last, (pad, 1))])


def postprocess(rois, out_cls, out_reg, class_mapping, resize_ratio, stride=16,
                det_threshold=DEFAULT_DET_THRESHOLD, nms_thresh=0.5, max_boxes=2000):
    """voc_dets.py:51-86 for one image.  rois (M,4) int16, out_cls (M,K) f32, out_reg (M,4(K-1)) f32
    are the rows the detector saw (padding duplicates included).  Returns the reference's list of
    {'bbox': int array[4], 'cls_name': str, 'prob': float32}."""
    return postprocess_batch(rois[None], out_cls[None], out_reg[None], class_mapping, [resize_ratio], stride,
                             det_threshold, nms_thresh, max_boxes)[0]


def postprocess_batch(rois, out_cls, out_reg, class_mapping, resize_ratios, stride=16,
                      det_threshold=DEFAULT_DET_THRESHOLD, nms_thresh=0.5, max_boxes=2000):
    """Batched post-processing: rois (B,M,4), out_cls (B,M,K), out_reg (B,M,4(K-1)), one launch."""
    ctx = get_context()
    boxes, probs, cls, count = _postprocess_device(ctx, rois, out_cls, out_reg, class_mapping, resize_ratios, stride,
                                                   det_threshold, nms_thresh, max_boxes)
    boxes, probs, cls, count = (ctx.to_host(t) for t in (boxes, probs, cls, count))
    names = {v: k for k, v in class_mapping.items()}
    out = []
    for b in range(len(count)):
        out.append([{'bbox': boxes[b, i].astype(np.int64), 'cls_name': names[int(cls[b, i])], 'prob': probs[b, i]}
                    for i in range(int(count[b]))])
    return out


def _postprocess_device(ctx, rois, out_cls, out_reg, class_mapping, resize_ratios, stride, det_threshold,
                        nms_thresh, max_boxes):
    def dev(x, dtype):
        return x.to(ctx.device) if isinstance(x, torch.Tensor) else ctx.to_device(x, dtype)
    return ops.det_postprocess(dev(rois, np.int16), dev(out_cls, np.float32), dev(out_reg, np.float32),
                               dev(np.asarray(resize_ratios, dtype=np.float64), np.float64), class_mapping['bg'],
                               stride, det_threshold, nms_thresh, max_boxes)


def get_dets(training_manager, detector, image, resize_ratio, num_rois=64, stride=16,
             det_threshold=DEFAULT_DET_THRESHOLD):
    """Same signature and return value as the reference's get_dets (voc_dets.py:20-88)."""
    conv_out, rois = training_manager.get_det_inputs(image)
    padded = pad_roi_batches(rois, num_rois)
    cls_parts, reg_parts = [], []
    for start in range(0, len(padded), num_rois):
        out_cls, out_reg = detector.predict([conv_out, np.expand_dims(padded[start:start + num_rois], axis=0)])
        cls_parts.append(out_cls[0])
        reg_parts.append(out_reg[0])
    cat = torch.cat if isinstance(cls_parts[0], torch.Tensor) else np.concatenate
    return postprocess(padded, cat(cls_parts), cat(reg_parts), training_manager.class_mapping, resize_ratio, stride,
                       det_threshold)
