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


def np_exp_f32_simd(x):
    """numpy's own float32 `exp` kernel (numpy >= 1.17 on x86 with AVX2 / AVX512F: simd_exp_FLOAT in
    loops_exponent_log.dispatch.c.src), restated with FMAs emulated in float64.  The device's np_expf() (common.cuh)
    is this sequence; numpy_exp_is_simd_kernel() tells whether the numpy of this machine dispatches to it."""
    f = np.float32
    x = np.asarray(x, f)

    def fma(a, b, c):
        return (np.asarray(a, np.float64) * np.asarray(b, np.float64) + np.asarray(c, np.float64)).astype(f)
    q = (x * f(1.44269504088896340736)).astype(f)
    q = ((q + f(12582912.0)).astype(f) - f(12582912.0)).astype(f)
    r = fma(q, f(-6.93145752e-1), x)
    r = fma(q, f(-1.42860677e-6), r)
    num = fma(f(5.082762527590693718096e-04), r, f(6.757896990527504603057e-03))
    for c in (5.114512081637298353406e-02, 2.473615434895520810817e-01, 7.257664613233124478488e-01, 9.999999999980870924916e-01):
        num = fma(num, r, f(c))
    den = fma(fma(f(2.159509375685829852307e-02), r, f(-2.742335390411667452936e-01)), r, f(1.0))
    return np.ldexp((num / den).astype(f), q.astype(np.int32)).astype(f)


def numpy_exp_is_simd_kernel():
    """True when np.exp on a STRIDED float32 column (what util.py:131 passes) equals np_exp_f32_simd bit for bit."""
    x = (np.random.default_rng(7).standard_normal((200000, 4)) * 0.6).astype(np.float32)
    return bool(np.array_equal(np.exp(x[:, 2]), np_exp_f32_simd(x[:, 2])))


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
