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

    @staticmethod
    def from_center_dims_int(x_center, y_center, width, height):
        """integer anchor box around a centre: x1 = cx - w // 2, x2 = x1 + w (reference: shapes.py:309-323)."""
        left, top = x_center - width // 2, y_center - height // 2
        return Box(left, top, left + width, top + height)

    @staticmethod
    def from_corners(coords):
        return Box(*coords)

    @property
    def width(self):
        return self.x2 - self.x1

    @property
    def height(self):
        return self.y2 - self.y1

    @property
    def x_center(self):
        return (self.x2 + self.x1) / 2

    @property
    def y_center(self):
        return (self.y1 + self.y2) / 2

    @property
    def corners(self):
        return np.array([self.x1, self.y1, self.x2, self.y2])

    @property
    def corner_dims(self):
        return np.array([self.x1, self.y1, self.width, self.height])

    @property
    def center_dims(self):
        return np.array([self.x_center, self.y_center, self.width, self.height])

    def resize(self, ratio):
        return Box(self.x1 * ratio, self.y1 * ratio, self.x2 * ratio, self.y2 * ratio)

    def horizontal_flip(self, image_width):
        """the same box in the horizontally mirrored image (reference: shapes.py:292-300)."""
        return Box(image_width - self.x2, self.y1, image_width - self.x1, self.y2)

    def __repr__(self):
        return "Box(%r, %r, %r, %r)" % (self.x1, self.y1, self.x2, self.y2)


class GroundTruthBox:
    """Labelled object (reference: shapes.py:187-305)."""
    __slots__ = ("obj_cls", "difficult", "box")

    def __init__(self, obj_cls, difficult, box):
        self.obj_cls, self.difficult, self.box = obj_cls, difficult, box

    x1 = property(lambda self: self.box.x1)
    y1 = property(lambda self: self.box.y1)
    x2 = property(lambda self: self.box.x2)
    y2 = property(lambda self: self.box.y2)
    width = property(lambda self: self.box.width)
    height = property(lambda self: self.box.height)
    x_center = property(lambda self: self.box.x_center)
    y_center = property(lambda self: self.box.y_center)
    corner_dims = property(lambda self: self.box.corner_dims)
    center_dims = property(lambda self: self.box.center_dims)

    @property
    def corners(self):
        return self.box.corners

    def resize(self, ratio):
        return GroundTruthBox(self.obj_cls, self.difficult, self.box.resize(ratio))

    def horizontal_flip(self, image_width):
        return GroundTruthBox(self.obj_cls, self.difficult, self.box.horizontal_flip(image_width))


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

    def horizontal_flip(self):
        """mirrored copy: boxes flipped about the image width, `flipped` toggled, so `cache_key` differs
        (reference: shapes.py:126-132,227-234); pixels are mirrored by the caller's loader, not here."""
        flipped = None if self._data is None else self._data[:, ::-1]
        return Image(self.name, self.width, self.height, [g.horizontal_flip(self.width) for g in self.gt_boxes],
                     not self.flipped, flipped)

    @property
    def num_gt_boxes(self):
        return len(self.gt_boxes)
