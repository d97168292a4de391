"""Minimal host-side value objects accepted by the drop-in managers.

The managers only read ``.width .height .gt_boxes[*].corners/.obj_cls/.resize()
.cache_key .data`` (reference: shapes.py:5-132,187-305), so any duck-typed
object works, including the reference's own ``shapes.Image``.  These light
classes exist so tests, the bench and users without the reference checkout can
build inputs; there is no image decoding here (pixels are out of scope for the
proposal / target path).
"""
import numpy as np


class Box:
    """Axis-aligned box, corners [x1, y1, x2, y2] (reference: shapes.py:307-408)."""
    __slots__ = ("x1", "y1", "x2", "y2")

    def __init__(self, x1, y1, x2, y2):
        self.x1, self.y1, self.x2, self.y2 = x1, y1, x2, y2

    @property
    def corners(self):
        return np.array([self.x1, self.y1, self.x2, self.y2])

    def resize(self, ratio):
        return Box(self.x1 * ratio, self.y1 * ratio, self.x2 * ratio, self.y2 * ratio)

    def __repr__(self):
        return "Box(%r, %r, %r, %r)" % (self.x1, self.y1, self.x2, self.y2)


class GroundTruthBox:
    """Labelled object (reference: shapes.py:187-305)."""
    __slots__ = ("obj_cls", "difficult", "box")

    def __init__(self, obj_cls, difficult, box):
        self.obj_cls, self.difficult, self.box = obj_cls, difficult, box

    @property
    def corners(self):
        return self.box.corners

    def resize(self, ratio):
        return GroundTruthBox(self.obj_cls, self.difficult, self.box.resize(ratio))


class Image:
    """Image metadata + optional in-memory pixels (reference: shapes.py:5-132)."""

    def __init__(self, name, width, height, gt_boxes=(), flipped=False, data=None):
        self.name, self.width, self.height = name, width, height
        self.gt_boxes, self.flipped, self._data = list(gt_boxes), flipped, data

    @property
    def cache_key(self):
        return self.name + str(self.flipped)

    @property
    def data(self):
        if self._data is None:
            raise ValueError("image %s carries no pixels" % self.name)
        return self._data

    def resize(self, ratio):
        w, h = int(round(ratio * self.width)), int(round(ratio * self.height))
        return Image(self.name, w, h, [g.resize(ratio) for g in self.gt_boxes], self.flipped, self._data)

    def resize_within_bounds(self, min_size, max_size):
        short, long_ = min(self.width, self.height), max(self.width, self.height)
        r_min = min_size / short
        ratio = max_size / long_ if r_min * long_ > max_size else r_min
        return self.resize(ratio), ratio
