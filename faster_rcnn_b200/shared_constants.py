"""Constants of the hot path; the values are part of the contract (reference:
shared_constants.py:1-18)."""
import math

import numpy as np

# float32 on purpose: it fixes the dtype of `regr_out / BBREG_MULTIPLIERS` (det_util.py:376)
BBREG_MULTIPLIERS = np.array([10, 10, 5, 5], dtype=np.float32)

DEFAULT_ANCHOR_SCALES = np.array([16, 32, 64, 128, 256, 512])
DEFAULT_ANCHOR_RATIOS = np.array([[1, 1], [1, 2], [2, 1]])


def _anchor_table(scales, ratios):
    # [height, width] per anchor, scale-major; side lengths floor-divided by sqrt(ratio area)
    dims = np.array([[s * rh, s * rw] for s in scales for rh, rw in ratios])
    norm = np.array([math.sqrt(s * rh * s * rw) / s for s in scales for rh, rw in ratios])
    return (dims // norm[:, None]).astype(int)


DEFAULT_ANCHORS = _anchor_table(DEFAULT_ANCHOR_SCALES, DEFAULT_ANCHOR_RATIOS)
DEFAULT_ANCHORS_PER_LOC = len(DEFAULT_ANCHORS)
DEFAULT_NUM_ITERATIONS = 10
DEFAULT_LEARN_RATE = 1e-3
DEFAULT_MOMENTUM = 0.9
RESIZE_MIN_SIZE = 600
RESIZE_MAX_SIZE = 1000
NUM_ROIS = 64
