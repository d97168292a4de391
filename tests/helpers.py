"""Shared helpers of the test-suite."""
import os

import numpy as np

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden(name):
    return np.load(os.path.join(GOLDEN_DIR, name + ".npz"))


def dev(x, dtype=None):
    """numpy -> CUDA torch tensor."""
    import torch
    return torch.from_numpy(np.ascontiguousarray(x, dtype=dtype)).cuda()


def host(t):
    return t.cpu().numpy()


class FakeImage:
    """Duck-typed image (the managers read .width .height .gt_boxes .cache_key .data)."""

    def __init__(self, name, width, height, gts, data=None):
        from faster_rcnn_b200.shapes import Box, GroundTruthBox
        self.name, self.width, self.height = name, width, height
        self.gt_boxes = [GroundTruthBox(c, False, Box(x1, y1, x2, y2)) for c, x1, y1, x2, y2 in gts]
        self.cache_key = name + "False"
        self.data = data


class FakeRpn:
    """Stands in for the Keras RPN model: returns canned head outputs."""

    def __init__(self, cls, regr, conv=None):
        self.cls, self.regr, self.conv = cls, regr, conv
        self.output = [0, 1, 2] if conv is not None else [0, 1]

    def predict_on_batch(self, batch):
        return [self.cls, self.regr] + ([self.conv] if self.conv is not None else [])


def flipped_rows(got, want):
    """rows where two integer-valued box arrays differ (decode flips caused by expf ulps)."""
    return np.where(np.any(got != want, axis=1))[0]


def voc_eval_golden(cls):
    """(image_ids, confidence, boxes, gt_by_image, imagenames, rec, prec, ap) of tests/golden/voc_eval.npz."""
    g = golden("voc_eval")
    gt, start = {}, 0
    for name, cnt in zip(g[cls + "_gt_names"].tolist(), g[cls + "_gt_counts"].tolist()):
        gt[name] = (g[cls + "_gt_boxes"][start:start + cnt], g[cls + "_gt_difficult"][start:start + cnt])
        start += cnt
    names = g["names"].tolist()
    for name in names:
        gt.setdefault(name, (np.zeros((0, 4)), np.zeros(0, bool)))
    return (g[cls + "_ids"].tolist(), g[cls + "_conf"], g[cls + "_boxes"], gt, names, g[cls + "_rec"], g[cls + "_prec"],
            float(g[cls + "_ap"]))
