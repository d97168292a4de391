"""Pins the RoI-layer oracle.  Resize mode: TensorFlow 1.3 is unavailable, so the TF `ResizeBilinear` op is executed
by an independent implementation, OpenCV's cv2.dnn TensorFlow importer (fixtures + live): the oracle's taps reproduce
it bit for bit, its TF-form output to a few ulp.  Also torch.nn.functional.grid_sample on explicit legacy-coordinate
grids (tolerance only) with torch autograd for the backward.  Max mode: torchvision.ops.roi_pool.  CPU only."""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import roi_oracle as R


GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "resize_bilinear_cv2dnn.npz")


def _check_against_tf_resize_bilinear(x, y_cv2, pool):
    """x (h,w,c) crop, y_cv2 = cv2.dnn's TF ResizeBilinear of it.  Returns the worst deviation of the oracle's TF-form
    output in units of one float32 ulp of the largest tap magnitude."""
    # (1) same taps, OpenCV's association of the four-tap sum: bit for bit
    assert np.array_equal(R.resize_bilinear_opencv_form(x, pool), y_cv2)
    # (2) TF's association (the oracle proper): float32 rounding only
    h, w = x.shape[:2]
    got = R.roi_resize_fwd(x, np.array([[0, 0, w, h]]), pool)[0]
    ulp = np.spacing(np.float32(np.abs(x).max()))
    return float(np.abs(got - y_cv2).max() / ulp)


def test_resize_taps_pinned_by_cv2dnn_tensorflow_importer_golden():
    """Fixtures from tests/golden/make_golden_resize.py: a hand-encoded TensorFlow GraphDef with the ResizeBilinear op
    (align_corners=false) run by cv2.dnn 4.13 on crops from 1x1 to 38x63, pool 7 and 3."""
    z = np.load(GOLDEN)
    keys = sorted(k for k in z.files if k.startswith("x_"))
    assert len(keys) == 32
    worst, exact = 0.0, 0
    for k in keys:
        pool = int(k.rsplit("_", 1)[1])
        dev = _check_against_tf_resize_bilinear(z[k], z["y" + k[1:]], pool)
        worst, exact = max(worst, dev), exact + (dev == 0.0)
    assert worst <= 4.0, worst            # tolerance: 4 ulp of the largest tap (observed 2.0); 17 of 32 cases are exact
    assert exact >= 4                     # 1x1 / 7x7 / integer-ratio crops have no rounding at all


def test_resize_taps_pinned_by_cv2dnn_tensorflow_importer_live():
    """The same check against cv2.dnn run here, on every crop size the C1 / C5 proposals produce (1..24 cells per
    side) plus the full map."""
    cv2 = pytest.importorskip("cv2")
    if not hasattr(cv2, "dnn"):
        pytest.skip("cv2 without dnn")
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    from make_golden_resize import cv2_resize_bilinear
    rng = np.random.default_rng(5)
    sizes = [(h, w) for h in range(1, 25) for w in (1, 2, 3, 5, 7, 8, 11, 14, 17, 24)] + [(38, 63), (37, 62), (38, 94)]
    worst = 0.0
    for h, w in sizes:
        x = rng.standard_normal((h, w, 3), dtype=np.float32)
        worst = max(worst, _check_against_tf_resize_bilinear(x, cv2_resize_bilinear(x, 7), 7))
    assert worst <= 4.0, worst


def _torch_resize(feat, rois, pool):
    """crop + legacy bilinear (src = i * in/out, clamp at the border) via grid_sample(align_corners=True)."""
    outs = []
    for x1, y1, x2, y2 in rois.tolist():
        crop = feat[y1:y2, x1:x2].permute(2, 0, 1)[None]                     # (1,C,h,w)
        h, w = crop.shape[2:]
        sy = torch.arange(pool, dtype=torch.float64) * (h / pool)
        sx = torch.arange(pool, dtype=torch.float64) * (w / pool)
        gy = (2 * sy / (h - 1) - 1) if h > 1 else torch.zeros(pool, dtype=torch.float64)
        gx = (2 * sx / (w - 1) - 1) if w > 1 else torch.zeros(pool, dtype=torch.float64)
        grid = torch.stack(torch.meshgrid(gy, gx, indexing='ij')[::-1], dim=-1)[None]
        outs.append(F.grid_sample(crop.double(), grid, mode='bilinear', padding_mode='border', align_corners=True)[0]
                    .permute(1, 2, 0))
    return torch.stack(outs)


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_resize_forward_and_backward_match_grid_sample(seed):
    rng = np.random.default_rng(seed)
    h, w, c, n = 13, 16, 5, 12
    feat = rng.standard_normal((h, w, c)).astype(np.float32)
    x1, y1 = rng.integers(0, w - 1, n), rng.integers(0, h - 1, n)
    rois = np.stack([x1, y1, np.minimum(w, x1 + 1 + rng.integers(0, w, n)), np.minimum(h, y1 + 1 + rng.integers(0, h, n))], 1)
    got = R.roi_resize_fwd(feat, rois, 7)
    ft = torch.from_numpy(feat).double().requires_grad_(True)
    want = _torch_resize(ft, torch.from_numpy(rois), 7)
    assert np.abs(got - want.detach().numpy()).max() < 1e-5
    gout = rng.standard_normal(got.shape).astype(np.float32)
    (want * torch.from_numpy(gout).double()).sum().backward()
    assert np.abs(R.roi_resize_bwd(gout, rois, feat.shape) - ft.grad.numpy()).max() < 1e-4


def test_resize_identity_and_upsample():
    feat = np.arange(7 * 7 * 2, dtype=np.float32).reshape(7, 7, 2)
    assert np.array_equal(R.roi_resize_fwd(feat, np.array([[0, 0, 7, 7]]), 7)[0], feat)
    one = R.roi_resize_fwd(feat, np.array([[3, 2, 4, 3]]), 7)[0]             # 1x1 crop -> constant
    assert np.array_equal(one, np.broadcast_to(feat[2, 3], (7, 7, 2)))


def test_max_pool_spec():
    rng = np.random.default_rng(3)
    feat = rng.standard_normal((9, 11, 4)).astype(np.float32)
    rois = np.array([[0, 0, 11, 9], [2, 1, 5, 3], [4, 4, 5, 5], [1, 0, 10, 8]])
    out, arg = R.roi_max_fwd(feat, rois, 7)
    flat = feat.reshape(-1, 4)
    assert np.array_equal(np.take_along_axis(flat, arg.reshape(-1, 4), 0).reshape(out.shape), out)
    # every cell of the crop belongs to at least one bin: the global max of the crop is an output
    for r, (x1, y1, x2, y2) in enumerate(rois):
        assert np.array_equal(out[r].max(axis=(0, 1)), feat[y1:y2, x1:x2].max(axis=(0, 1)))
    gout = rng.standard_normal(out.shape).astype(np.float32)
    dx = R.roi_max_bwd(gout, arg, feat.shape)
    assert abs(float(dx.sum()) - float(gout.sum())) < 1e-3


@pytest.mark.parametrize("seed", [0, 1, 2, 3])
def test_max_mode_is_torchvision_roi_pool(seed):
    """Pins the max-pool variant (which the reference does not have) to an independent implementation: the oracle's
    bins, outputs and arg-max rule are those of torchvision.ops.roi_pool (the Fast R-CNN RoIPool: hstart =
    floor(ph*h/P), hend = ceil((ph+1)*h/P), first maximum in row-major order) once the exclusive x2/y2 of this
    path are passed as inclusive ends.  Outputs bit for bit; the backward against torchvision's autograd, with
    tie-rich features so that the arg-max rule is exercised."""
    tv = pytest.importorskip("torchvision")
    rng = np.random.default_rng(seed)
    for _ in range(25):
        h, w, c, n = int(rng.integers(1, 40)), int(rng.integers(1, 64)), 3, 16
        pool = int(rng.choice([1, 2, 3, 7, 9]))
        feat = rng.standard_normal((h, w, c), dtype=np.float32)
        if rng.random() < 0.5:
            feat = np.round(feat * 2) / 2                                    # ties
        x1, y1 = rng.integers(0, w, n), rng.integers(0, h, n)
        rois = np.stack([x1, y1, np.minimum(w, x1 + 1 + rng.integers(0, w, n)), np.minimum(h, y1 + 1 + rng.integers(0, h, n))], 1).astype(np.int16)
        out, arg = R.roi_max_fwd(feat, rois, pool)
        t = torch.from_numpy(feat).permute(2, 0, 1)[None].clone().requires_grad_(True)
        boxes = torch.tensor([[0, a, b, c2 - 1, d - 1] for a, b, c2, d in rois.tolist()], dtype=torch.float32)
        want = tv.ops.roi_pool(t, boxes, output_size=pool, spatial_scale=1.0)                 # (K, C, P, P)
        assert np.array_equal(out, want.detach().permute(0, 2, 3, 1).numpy())
        g = rng.standard_normal(out.shape, dtype=np.float32)
        want.backward(torch.from_numpy(g).permute(0, 3, 1, 2))
        got = R.roi_max_bwd(g, arg, feat.shape)
        ref = t.grad[0].permute(1, 2, 0).numpy()
        assert np.abs(got - ref).max() <= 1e-5 * max(1.0, float(np.abs(ref).max()))
