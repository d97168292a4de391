"""Seeded synthetic workloads for the hot path.

There is no network for datasets or trained weights, so RPN / detector head
outputs are synthetic.  Every generator is a pure function of its seed
(`np.random.default_rng`), produces the reference's array layouts and is used
identically by tests, `__graft_entry__.smoke()` and `bench.py`.
"""
import numpy as np

VOC_CLASSES = ['aeroplane', 'bicycle', 'bird', 'boat', 'bottle', 'bus', 'car', 'cat', 'chair', 'cow',
               'diningtable', 'dog', 'horse', 'motorbike', 'person', 'pottedplant', 'sheep', 'sofa',
               'train', 'tvmonitor']
VOC_CLASS_MAPPING = dict({c: i for i, c in enumerate(VOC_CLASSES)}, bg=20)   # data/voc_data_helpers.py:10-32


def rpn_outputs(rows, cols, n_anchors, seed, clustered=False, n_objects=12):
    """cls (1,R,C,A) f32 with unique (tie-free) scores and regr (1,R,C,4A) f32.

    clustered=False: scores are a random permutation of (i+0.5)/N and deltas
    are N(0,1)*[1,1,1.5,1.5] (before the /[10,10,5,5] of det_util.py:376).
    clustered=True: scores peak around `n_objects` random centres so the top-k
    proposals overlap heavily and NMS has to scan deep, as with a trained RPN.
    """
    rng = np.random.default_rng(seed)
    n = rows * cols * n_anchors
    base = (rng.permutation(n).astype(np.float64) + 0.5) / n
    regr = rng.standard_normal((n, 4)) * np.array([1.0, 1.0, 1.5, 1.5])
    if clustered:
        ys, xs = np.meshgrid(np.arange(rows), np.arange(cols), indexing='ij')
        heat = np.zeros((rows, cols))
        for _ in range(n_objects):
            cy, cx = rng.uniform(0, rows), rng.uniform(0, cols)
            sy, sx = rng.uniform(1.5, 6.0), rng.uniform(1.5, 6.0)
            heat = np.maximum(heat, np.exp(-0.5 * (((ys - cy) / sy) ** 2 + ((xs - cx) / sx) ** 2)))
        heat = np.repeat(heat.reshape(-1), n_anchors)
        # rank-transform keeps scores unique: order by heat + small noise
        order = np.argsort(heat + 0.15 * base, kind='stable')
        base = np.empty(n)
        base[order] = (np.arange(n) + 0.5) / n
        regr *= 0.35
    cls = base.astype(np.float32).reshape(1, rows, cols, n_anchors)
    assert len(np.unique(cls)) == n, "scores must be tie-free"
    return cls, regr.astype(np.float32).reshape(1, rows, cols, 4 * n_anchors)


def feature_map(rows, cols, channels, seed):
    """conv feature map (1,R,C,Cf) f32 ~ N(0,1)."""
    return np.random.default_rng(seed).standard_normal((1, rows, cols, channels), dtype=np.float32)


def gt_boxes(n_gt, img_w, img_h, seed, classes=VOC_CLASSES):
    """`n_gt` pixel-space GT boxes with w,h ~ U{20..399}, classes cycling.
    Returns list of (cls_name, x1, y1, x2, y2) with integer corners inside the image."""
    rng = np.random.default_rng(seed)
    out = []
    for i in range(n_gt):
        w = int(min(rng.integers(20, 400), img_w - 2))
        h = int(min(rng.integers(20, 400), img_h - 2))
        x1 = int(rng.integers(0, img_w - w))
        y1 = int(rng.integers(0, img_h - h))
        out.append((classes[i % len(classes)], x1, y1, x1 + w, y1 + h))
    return out


def detector_outputs(n_rows, n_classes, seed):
    """out_cls (n_rows,K) f32 softmax of N(0,2) logits, out_reg (n_rows,4(K-1)) f32 ~ N(0,1)."""
    rng = np.random.default_rng(seed)
    logits = 2.0 * rng.standard_normal((n_rows, n_classes))
    e = np.exp(logits - logits.max(axis=1, keepdims=True))
    cls = (e / e.sum(axis=1, keepdims=True)).astype(np.float32)
    reg = rng.standard_normal((n_rows, 4 * (n_classes - 1))).astype(np.float32)
    return cls, reg


def random_rois(n, rows, cols, seed):
    """(n,4) int16 valid feature-space RoIs (x2>x1, y2>y1, inside the map)."""
    rng = np.random.default_rng(seed)
    x1 = rng.integers(0, cols - 1, n)
    y1 = rng.integers(0, rows - 1, n)
    x2 = np.minimum(cols - 1, x1 + 1 + rng.integers(0, 24, n))
    y2 = np.minimum(rows - 1, y1 + 1 + rng.integers(0, 24, n))
    return np.stack([x1, y1, x2, y2], axis=1).astype(np.int16)
