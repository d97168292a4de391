"""Property tests (hypothesis) of the path's invariants -- the reference has no unit tests of its own for these
functions, so the invariants are stated here: on the oracle (CPU) and, marked `gpu`, on the
CUDA kernels through the C ABI with hypothesis-drawn inputs compared bit for bit against the oracle."""
import numpy as np
import pytest
from hypothesis import HealthCheck, given, settings
from hypothesis import strategies as st

from oracle import frcnn_oracle as O
from oracle import roi_oracle as R

COMMON = dict(deadline=None, suppress_health_check=[HealthCheck.too_slow, HealthCheck.function_scoped_fixture])


@st.composite
def int16_boxes(draw, max_n=60, span=80):
    n = draw(st.integers(1, max_n))
    rows = draw(st.lists(st.tuples(st.integers(0, span), st.integers(0, span), st.integers(0, 30), st.integers(0, 30)),
                         min_size=n, max_size=n))
    boxes = np.array([[x, y, x + w, y + h] for x, y, w, h in rows], np.int16)
    tie_free = draw(st.booleans())
    if tie_free:
        probs = (np.array(draw(st.permutations(list(range(n)))), np.float32) + 0.5) / n
    else:
        probs = np.array(draw(st.lists(st.integers(0, 3), min_size=n, max_size=n)), np.float32) / 4
    return boxes, probs.astype(np.float32)


def _iou_plus1(a, b):
    iw = max(0, min(a[2], b[2]) - max(a[0], b[0]) + 1)
    ih = max(0, min(a[3], b[3]) - max(a[1], b[1]) + 1)
    inter = iw * ih
    return inter / ((a[2] - a[0] + 1) * (a[3] - a[1] + 1) + (b[2] - b[0] + 1) * (b[3] - b[1] + 1) - inter)


@settings(max_examples=120, **COMMON)
@given(int16_boxes(), st.sampled_from([0.3, 0.5, 0.7, 0.9]), st.integers(1, 80))
def test_nms_invariants_on_the_oracle(data, thresh, max_boxes):
    boxes, probs = data
    b = boxes.astype(np.int64)
    pick = O.greedy_nms(boxes, probs, thresh, max_boxes)
    assert 1 <= len(pick) <= min(max_boxes, len(boxes)) and len(set(pick.tolist())) == len(pick)
    assert np.all(np.diff(probs[pick]) <= 0)                                   # pick order = descending score
    for i in range(len(pick)):                                                  # kept boxes do not suppress each other
        for j in range(i):
            assert _iou_plus1(b[pick[j]], b[pick[i]]) <= thresh
    if len(pick) < max_boxes:                                                   # every dropped box is covered by a kept one
        for d in set(range(len(boxes))) - set(pick.tolist()):
            assert any(_iou_plus1(b[k], b[d]) > thresh and probs[k] >= probs[d] for k in pick)
        again = O.greedy_nms(boxes[pick], probs[pick], thresh, max_boxes)       # idempotent
        assert np.array_equal(np.sort(again), np.arange(len(pick)))             # (ties are revisited in reverse position order)
        if len(np.unique(probs)) == len(probs):
            assert np.array_equal(again, np.arange(len(pick)))


@settings(max_examples=100, **COMMON)
@given(st.integers(1, 12), st.integers(1, 12), st.integers(0, 2 ** 31 - 1))
def test_iou_matrix_and_labels_invariants(n, g, seed):
    rng = np.random.default_rng(seed)

    def boxes(m):
        xy = rng.integers(0, 100, (m, 2))
        return np.concatenate([xy, xy + rng.integers(1, 60, (m, 2))], axis=1).astype(np.float32)
    a, b = boxes(n), boxes(g)
    iou = O.iou_matrix(a, b)
    assert iou.shape == (n, g) and iou.dtype == np.float32 and np.all(iou >= 0) and np.all(iou <= 1)
    assert np.array_equal(iou, O.iou_matrix(b, a).T)                            # symmetric, bit for bit
    assert np.all(np.diag(O.iou_matrix(a, a)) == 1.0)
    shift = np.float32(17.0)                                                    # translation invariance (exact for integers)
    assert np.array_equal(iou, O.iou_matrix(a + shift, b + shift))


@settings(max_examples=60, **COMMON)
@given(st.integers(1, 9), st.integers(1, 9), st.integers(1, 6), st.integers(1, 7), st.integers(0, 2 ** 31 - 1))
def test_roi_layer_invariants_on_the_oracle(h, w, n, pool, seed):
    rng = np.random.default_rng(seed)
    x1, y1 = rng.integers(0, w, n), rng.integers(0, h, n)
    rois = np.stack([x1, y1, np.minimum(w, x1 + 1 + rng.integers(0, w, n)), np.minimum(h, y1 + 1 + rng.integers(0, h, n))], 1).astype(np.int16)
    const = np.full((h, w, 3), 2.5, np.float32)
    assert np.all(R.roi_resize_fwd(const, rois, pool) == 2.5)                   # interpolation of a constant is the constant
    feat = rng.standard_normal((h, w, 3), dtype=np.float32)
    out = R.roi_resize_fwd(feat, rois, pool)
    mo, ma = R.roi_max_fwd(feat, rois, pool)
    for r, (a, b, c, d) in enumerate(rois):
        crop = feat[b:d, a:c]
        assert out[r].min() >= crop.min() - 1e-6 and out[r].max() <= crop.max() + 1e-6      # convex combination of the crop
        assert np.array_equal(mo[r].max(axis=(0, 1)), crop.max(axis=(0, 1)))               # bins cover the crop
        assert np.array_equal(feat.reshape(-1, 3)[ma[r].reshape(-1, 3), np.arange(3)], mo[r].reshape(-1, 3))
    g = rng.standard_normal(out.shape, dtype=np.float32)
    lhs = float((out.astype(np.float64) * g).sum())                             # backward is the adjoint of the forward
    rhs = float((R.roi_resize_bwd(g, rois, feat.shape).astype(np.float64) * feat).sum())
    assert abs(lhs - rhs) <= 1e-4 * max(1.0, abs(lhs))
    assert abs(float(R.roi_max_bwd(g, ma, feat.shape).sum()) - float(g.sum())) <= 1e-3 * max(1.0, float(np.abs(g).sum()))


@pytest.mark.gpu
@settings(max_examples=60, **COMMON)
@given(int16_boxes(max_n=200, span=120), st.sampled_from([0.0, 0.3, 0.5, 0.7, 0.75, 1.0]), st.integers(1, 300))
def test_gpu_nms_equals_oracle_on_drawn_inputs(data, thresh, max_boxes):
    from faster_rcnn_b200 import ops
    from helpers import dev, host
    boxes, probs = data
    pick = O.greedy_nms(boxes, probs, thresh, max_boxes)
    ki, kc, kb, _ = ops.nms_i16(dev(boxes[None]), dev(probs[None]), None, thresh, max_boxes)
    m = int(host(kc)[0])
    assert m == len(pick) and np.array_equal(host(ki)[0, :m], pick) and np.array_equal(host(kb)[0, :m], boxes[pick])


@pytest.mark.gpu
@settings(max_examples=40, **COMMON)
@given(st.integers(1, 12), st.integers(1, 14), st.sampled_from([4, 12, 64, 132]), st.integers(1, 40), st.sampled_from([1, 2, 7, 9]),
       st.integers(0, 2 ** 31 - 1))
def test_gpu_roi_layer_equals_oracle_on_drawn_inputs(h, w, c, n, pool, seed):
    from faster_rcnn_b200 import ops
    from helpers import dev, host
    rng = np.random.default_rng(seed)
    feat = np.round(rng.standard_normal((h, w, c), dtype=np.float32) * 4) / 4     # plenty of ties for the max mode
    x1, y1 = rng.integers(0, w, n), rng.integers(0, h, n)
    rois = np.stack([x1, y1, np.minimum(w, x1 + 1 + rng.integers(0, w, n)), np.minimum(h, y1 + 1 + rng.integers(0, h, n))], 1).astype(np.int16)
    assert np.array_equal(host(ops.roi_forward(dev(feat[None]), dev(rois[None]), pool, "resize"))[0], R.roi_resize_fwd(feat, rois, pool))
    mo, ma = ops.roi_forward(dev(feat[None]), dev(rois[None]), pool, "max")
    wo, wa = R.roi_max_fwd(feat, rois, pool)
    assert np.array_equal(host(mo)[0], wo) and np.array_equal(host(ma)[0], wa)
    g = rng.standard_normal((n, pool, pool, c), dtype=np.float32)
    gm = host(ops.roi_backward(dev(g[None]), dev(rois[None]), (1, h, w, c), "max", argmax=ma))[0]
    assert np.array_equal(gm, R.roi_max_bwd(g, wa, (h, w, c)))
