"""GPU parity: proposal stage (K-a decode/top-k, K-b NMS) through the C ABI vs the numpy oracle
and the reference-generated golden vectors.  Bit-exact for boxes/indices/keep lists.

The decode evaluates numpy's own float32 exp kernel (np_expf in common.cuh), so on machines whose numpy dispatches to
that kernel (x86 with AVX2 / AVX512F, numpy >= 1.17) decoded boxes are bit-exact too
(test_decode_is_bit_exact_with_numpy_simd_exp).  Elsewhere numpy falls back to libm's expf, which differs by an ulp
and can flip the half-to-even rounding of a decoded coordinate by one cell in ~1e-5 of the anchors: the generic tests
therefore tolerate MAX_FLIPS rows per image and compare downstream stages against the oracle run on the SAME
decoded boxes."""
import numpy as np
import pytest

from helpers import dev, flipped_rows, golden, host
from oracle import frcnn_oracle as O

pytestmark = pytest.mark.gpu

MAX_FLIPS = 3


@pytest.fixture(scope="module")
def ops():
    from faster_rcnn_b200 import ops as _ops
    return _ops


def _synth(rows, cols, scales, seed, clustered):
    from faster_rcnn_b200 import synth
    dims = O.anchor_table(scales) if scales else O.anchor_table()
    cls, regr = synth.rpn_outputs(rows, cols, len(dims), seed, clustered=clustered)
    return dims, cls, regr


def test_decode_is_bit_exact_with_numpy_simd_exp(ops):
    """1.4 M anchors (64 images of the C1 shape) with wide size deltas: every decoded coordinate equals the oracle's.
    With CUDA's expf this sweep shows ~1e-5 flipped rows per anchor; with np_expf none."""
    from helpers import numpy_exp_is_simd_kernel
    if not numpy_exp_is_simd_kernel():
        pytest.skip("this machine's numpy does not use its SIMD float32 exp kernel")
    from faster_rcnn_b200 import synth
    dims = O.anchor_table([128, 256, 512])
    pairs = [synth.rpn_outputs(38, 63, 9, 5000 + i) for i in range(64)]
    cls, regr = np.concatenate([p[0] for p in pairs]), np.concatenate([p[1] for p in pairs])
    regr[32:] *= 3.0                                            # sizes up to e^(+-3): far beyond what a trained head emits
    dense = host(ops.decode_topk(dev(regr), dev(cls), dims, 16, 8000, want_dense=True)[4])
    flips = sum(len(flipped_rows(dense[i], O.proposals_from_rpn(regr[i:i + 1].copy(), dims, 16))) for i in range(64))
    assert flips == 0, flips


def _check_decode_topk(ops, dims, cls, regr, k, want_dense=None):
    boxes, scores, index, count, dense = ops.decode_topk(dev(regr), dev(cls), dims, 16, k, want_dense=True)
    dense = host(dense)[0]
    if want_dense is None:
        want_dense = O.proposals_from_rpn(regr.copy(), dims, 16)
    flips = flipped_rows(dense, want_dense)
    assert len(flips) <= MAX_FLIPS, "decode: %d rows differ from the oracle" % len(flips)
    if len(flips):
        assert np.abs(dense[flips] - want_dense[flips]).max() <= 1.0
    # top-k is exact given the decoded boxes
    wb, wp, widx = O.topk_proposals(dense.copy(), cls.reshape(-1), k)
    n = int(host(count)[0])
    assert n == len(wb)
    assert np.array_equal(host(boxes)[0, :n], wb) and np.array_equal(host(scores)[0, :n], wp)
    assert np.array_equal(host(index)[0, :n], widx)
    assert np.all(host(index)[0, n:] == -1)
    return dense, len(flips)


@pytest.mark.parametrize("rows,cols,scales,seed,k,clustered", [
    (38, 63, [128, 256, 512], 1, 8000, False),      # C1 VOC ResNet-50
    (38, 63, [128, 256, 512], 2, 12000, True),      # training k
    (38, 94, None, 3, 12000, True),                 # C3 KITTI, 18 anchors incl. zero-sized feature anchors
    (38, 94, None, 4, 8000, False),
    (37, 62, [128, 256, 512], 5, 8000, False),      # VGG16 map
    (5, 7, [128, 256, 512], 6, 8000, False),        # k > number of anchors
    (1, 1, [128, 256, 512], 7, 4, False),
])
def test_decode_topk_vs_oracle(ops, rows, cols, scales, seed, k, clustered):
    dims, cls, regr = _synth(rows, cols, scales, seed, clustered)
    _check_decode_topk(ops, dims, cls, regr, k)


@pytest.mark.parametrize("tag", ["voc_small", "voc_clustered", "kitti_small"])
def test_proposal_stage_vs_golden(ops, tag):
    g = golden("proposals_" + tag)
    k, max_boxes = int(g["k"]), int(g["max_boxes"])
    dense, n_flips = _check_decode_topk(ops, g["anchor_dims"], g["cls"], g["regr"], k, want_dense=g["dense"])
    # NMS kernel on the reference's own top-k boxes: bit-exact keep list
    ki, kc, kb, ks = ops.nms_i16(dev(g["topk_boxes"][None]), dev(g["topk_probs"][None]), None, 0.7, max_boxes)
    n = int(host(kc)[0])
    assert n == len(g["nms_boxes"])
    assert np.array_equal(host(kb)[0, :n], g["nms_boxes"]) and np.array_equal(host(ks)[0, :n], g["nms_probs"])
    # fused decode -> top-k -> NMS on the device
    rois, scores, count = ops.proposals(dev(g["regr"]), dev(g["cls"]), g["anchor_dims"], 16, k, 0.7, max_boxes)
    if n_flips == 0:
        n = int(host(count)[0])
        assert np.array_equal(host(rois)[0, :n], g["nms_boxes"]) and np.array_equal(host(scores)[0, :n], g["nms_probs"])


@pytest.mark.parametrize("rows,cols,scales,seed,k,max_boxes,clustered", [
    (38, 63, [128, 256, 512], 11, 8000, 300, False),
    (38, 63, [128, 256, 512], 12, 8000, 300, True),
    (38, 63, [128, 256, 512], 13, 12000, 2000, True),
    (38, 94, None, 14, 12000, 2000, True),
    (38, 94, None, 15, 8000, 300, False),
])
def test_fused_proposals_vs_oracle(ops, rows, cols, scales, seed, k, max_boxes, clustered):
    dims, cls, regr = _synth(rows, cols, scales, seed, clustered)
    dense = host(ops.decode_topk(dev(regr), dev(cls), dims, 16, k, want_dense=True)[4])[0]
    wb, wp, _ = O.topk_proposals(dense.copy(), cls.reshape(-1), k)
    pick = O.greedy_nms(wb, wp, 0.7, max_boxes)
    rois, scores, count = ops.proposals(dev(regr), dev(cls), dims, 16, k, 0.7, max_boxes)
    n = int(host(count)[0])
    assert n == len(pick)
    assert np.array_equal(host(rois)[0, :n], wb[pick]) and np.array_equal(host(scores)[0, :n], wp[pick])
    assert np.all(host(rois)[0, n:] == 0)


def test_fused_proposals_batch(ops):
    """A batch of images = independent problems in one launch sequence."""
    dims = O.anchor_table([128, 256, 512])
    from faster_rcnn_b200 import synth
    pairs = [synth.rpn_outputs(38, 63, 9, 20 + i, clustered=bool(i % 2)) for i in range(5)]
    cls = np.concatenate([p[0] for p in pairs])
    regr = np.concatenate([p[1] for p in pairs])
    rois, scores, count = ops.proposals(dev(regr), dev(cls), dims, 16, 8000, 0.7, 300)
    for b in range(5):
        one_r, one_s, one_c = ops.proposals(dev(regr[b:b + 1]), dev(cls[b:b + 1]), dims, 16, 8000, 0.7, 300)
        n = int(host(one_c)[0])
        assert int(host(count)[b]) == n
        assert np.array_equal(host(rois)[b], host(one_r)[0]) and np.array_equal(host(scores)[b], host(one_s)[0])


# ------------------------------------------------------------------------------------------------
# NMS kernel in isolation
# ------------------------------------------------------------------------------------------------
def _rand_boxes_i16(rng, n, size=64, max_wh=24):
    x1 = rng.integers(0, size - 1, n)
    y1 = rng.integers(0, size - 1, n)
    x2 = np.minimum(size - 1, x1 + 1 + rng.integers(0, max_wh, n))
    y2 = np.minimum(size - 1, y1 + 1 + rng.integers(0, max_wh, n))
    return np.stack([x1, y1, x2, y2], axis=1).astype(np.int16)


def _nms_dev(ops, boxes, probs, thresh, max_boxes):
    ki, kc, kb, ks = ops.nms_i16(dev(boxes[None]), dev(probs[None]), None, thresh, max_boxes)
    n = int(host(kc)[0])
    assert np.all(host(ki)[0, n:] == -1)
    return host(ki)[0, :n], host(kb)[0, :n], host(ks)[0, :n]


@pytest.mark.parametrize("n,thresh,max_boxes", [
    (1, 0.7, 300), (2, 0.7, 300), (63, 0.7, 300), (64, 0.7, 300), (65, 0.7, 300), (129, 0.5, 10),
    (1000, 0.7, 300), (1001, 0.3, 2000), (4097, 0.7, 300), (8000, 0.7, 300), (12000, 0.7, 2000),
    (16384, 0.7, 300), (500, 0.0, 300), (500, 1.0, 300), (500, 0.7, 1),
])
def test_nms_i16_unsorted_vs_oracle(ops, n, thresh, max_boxes):
    rng = np.random.default_rng(n)
    boxes = _rand_boxes_i16(rng, n)
    probs = (rng.permutation(n).astype(np.float32) + 0.5) / n          # unique, arbitrary order
    pick = O.greedy_nms(boxes, probs, thresh, max_boxes)
    ki, kb, ks = _nms_dev(ops, boxes, probs, thresh, max_boxes)
    assert np.array_equal(ki, pick) and np.array_equal(kb, boxes[pick]) and np.array_equal(ks, probs[pick])


@pytest.mark.parametrize("n", [2, 100, 3001, 20000])
def test_nms_i16_sorted_input_tma_path(ops, n):
    """strictly descending scores (the layout K-a emits): identity order, bulk-copy load path;
    odd n exercises the non-16-byte-multiple fallback; n = 20000 > FRCNN_NMS_MAX_UNSORTED."""
    rng = np.random.default_rng(n + 7)
    boxes = _rand_boxes_i16(rng, n, size=96)
    probs = np.sort((rng.permutation(n).astype(np.float32) + 0.5) / n)[::-1].copy()
    pick = O.greedy_nms(boxes, probs, 0.7, 300)
    ki, kb, ks = _nms_dev(ops, boxes, probs, 0.7, 300)
    assert np.array_equal(ki, pick) and np.array_equal(kb, boxes[pick])


def test_nms_i16_ties_follow_the_stable_order(ops):
    """equal scores: visit order = (score desc, position desc), i.e. argsort(kind='stable') from the back."""
    rng = np.random.default_rng(3)
    boxes = _rand_boxes_i16(rng, 2000)
    probs = (rng.integers(0, 40, 2000) / 40).astype(np.float32)
    pick = O.greedy_nms(boxes, probs, 0.7, 300, stable=True)
    ki, _, _ = _nms_dev(ops, boxes, probs, 0.7, 300)
    assert np.array_equal(ki, pick)
    same = np.tile(np.array([[3, 4, 20, 30]], dtype=np.int16), (70, 1))           # all identical boxes
    ki, _, _ = _nms_dev(ops, same, np.ones(70, dtype=np.float32), 0.7, 300)
    assert ki.tolist() == [69]


def test_nms_i16_exact_threshold_boundary(ops):
    """inter/union == 0.7 exactly must survive (<=), one cell more must not; the reference divides in f64."""
    a = [0, 0, 9, 9]                       # area 100
    b = [0, 3, 9, 9]                       # area 70, inter 70, union 100 -> 0.7 exactly
    c = [0, 2, 9, 9]                       # area 80 -> 0.8
    boxes = np.array([a, b, c], dtype=np.int16)
    probs = np.array([0.9, 0.8, 0.7], dtype=np.float32)
    pick = O.greedy_nms(boxes, probs, 0.7, 300)
    ki, _, _ = _nms_dev(ops, boxes, probs, 0.7, 300)
    assert pick.tolist() == [0, 1] and ki.tolist() == [0, 1]


def test_nms_i16_ragged_batch(ops):
    rng = np.random.default_rng(9)
    ns = [0, 1, 777, 4000, 64]
    n_max = max(ns)
    boxes = np.zeros((len(ns), n_max, 4), dtype=np.int16)
    probs = np.zeros((len(ns), n_max), dtype=np.float32)
    for b, n in enumerate(ns):
        boxes[b, :n] = _rand_boxes_i16(rng, n)
        probs[b, :n] = (rng.permutation(n).astype(np.float32) + 0.5) / max(n, 1)
    ki, kc, kb, ks = ops.nms_i16(dev(boxes), dev(probs), dev(np.array(ns, dtype=np.int32)), 0.7, 300)
    for b, n in enumerate(ns):
        pick = O.greedy_nms(boxes[b, :n], probs[b, :n], 0.7, 300)
        assert int(host(kc)[b]) == len(pick)
        assert np.array_equal(host(ki)[b, :len(pick)], pick)


def test_nms_properties_full_size(ops):
    """size-independent properties at BASELINE's full size (12000 -> 2000): kept boxes are pairwise
    below the threshold, every dropped box ahead of the cut is covered by an earlier kept one, and
    NMS of the kept set is the identity (idempotence)."""
    from faster_rcnn_b200 import synth
    dims = O.anchor_table()
    cls, regr = synth.rpn_outputs(38, 94, 18, 77, clustered=True)
    tb, ts, _, tc = ops.decode_topk(dev(regr), dev(cls), dims, 16, 12000)
    n = int(host(tc)[0])
    ki, kc, kb, ks = ops.nms_i16(tb, ts, tc, 0.7, 2000)
    m = int(host(kc)[0])
    boxes, keep = host(tb)[0, :n].astype(np.int64), host(ki)[0, :m]
    assert np.all(np.diff(keep) > 0)                       # pick order = descending score = ascending position
    area = (boxes[:, 2] - boxes[:, 0] + 1) * (boxes[:, 3] - boxes[:, 1] + 1)

    def over(i_idx, j_idx):                                 # exact integer test of inter/union > 0.7
        a, b = boxes[i_idx][:, None, :], boxes[j_idx][None, :, :]
        iw = np.maximum(0, np.minimum(a[..., 2], b[..., 2]) - np.maximum(a[..., 0], b[..., 0]) + 1)
        ih = np.maximum(0, np.minimum(a[..., 3], b[..., 3]) - np.maximum(a[..., 1], b[..., 1]) + 1)
        inter = iw * ih
        union = area[i_idx][:, None] + area[j_idx][None, :] - inter
        return 10 * inter > 7 * union
    kk = over(keep, keep)
    np.fill_diagonal(kk, False)
    assert not kk.any()
    last = keep[-1]
    dropped = np.setdiff1d(np.arange(last), keep)
    cover = over(dropped, keep) & (keep[None, :] < dropped[:, None])
    assert cover.any(axis=1).all()
    ki2, kc2, _, _ = ops.nms_i16(kb[:, :m].contiguous(), ks[:, :m].contiguous(), None, 0.7, 2000)
    assert int(host(kc2)[0]) == m and np.array_equal(host(ki2)[0, :m], np.arange(m))


# ------------------------------------------------------------------------------------------------
# float64 segmented NMS
# ------------------------------------------------------------------------------------------------
def test_nms_f64_vs_golden_and_oracle(ops):
    g = golden("nms_f64_iou")
    offs = dev(np.array([0, 400], dtype=np.int32))
    ki, kc = ops.nms_f64(dev(g["boxes"]), dev(g["probs"]), offs, 400, 0.5, 2000)
    n = int(host(kc)[0])
    pick = host(ki)[0, :n]
    assert np.array_equal(g["boxes"][pick], g["nms_boxes"]) and np.array_equal(g["probs"][pick], g["nms_probs"])
    # several segments of different length incl. an empty one, ties inside a segment
    rng = np.random.default_rng(21)
    lens = [0, 1, 37, 320, 5]
    total = sum(lens)
    xy, wh = rng.uniform(0, 300, (total, 2)), rng.uniform(5, 150, (total, 2))
    boxes = np.concatenate([xy, xy + wh], axis=1)
    probs = (rng.integers(0, 50, total) / 50).astype(np.float32)
    so = np.concatenate([[0], np.cumsum(lens)]).astype(np.int32)
    ki, kc = ops.nms_f64(dev(boxes), dev(probs), dev(so), max(lens), 0.5, 2000)
    for s, ln in enumerate(lens):
        pick = O.greedy_nms(boxes[so[s]:so[s + 1]], probs[so[s]:so[s + 1]], 0.5, 2000)
        assert int(host(kc)[s]) == len(pick) and np.array_equal(host(ki)[s, :len(pick)], pick)


# ------------------------------------------------------------------------------------------------
# drop-in module surface (numpy in / numpy out)
# ------------------------------------------------------------------------------------------------
def test_dropin_det_util_functions():
    from faster_rcnn_b200 import det_util, synth
    g = golden("proposals_voc_clustered")
    rois = det_util._get_rois(g["regr"], g["anchor_dims"], 16)
    flips = flipped_rows(rois, g["dense"])
    assert rois.dtype == np.float32 and len(flips) <= MAX_FLIPS
    assert np.array_equal(det_util._get_valid_box_idxs(g["dense"]), g["valid"])
    nb, npb = det_util.nms(g["topk_boxes"], g["topk_probs"], max_boxes=int(g["max_boxes"]), overlap_thresh=0.7)
    assert nb.dtype == np.int16 and np.array_equal(nb, g["nms_boxes"]) and np.array_equal(npb, g["nms_probs"])
    assert det_util.nms(g["topk_boxes"][:0], g["topk_probs"][:0]) == []
    f = golden("nms_f64_iou")
    nb, npb = det_util.nms(f["boxes"], f["probs"], overlap_thresh=0.5, max_boxes=2000)
    assert nb.dtype == np.float64 and np.array_equal(nb, f["nms_boxes"]) and np.array_equal(npb, f["nms_probs"])
    # anchors + sanitize helpers
    dims = O.anchor_table() // 16
    anc = det_util._get_anchor_coords(7, 9, dims)
    assert anc.shape == (7, 9, 18, 4) and np.array_equal(anc.reshape(-1, 4), O.feature_anchors(7, 9, dims))
    rng = np.random.default_rng(1)
    raw = np.round(rng.normal(20, 30, (500, 4))).astype(np.float32)
    want = O.sanitize_boxes(63, 38, raw.copy())
    got = det_util._sanitize_boxes_inplace(63, 38, raw)
    assert got is raw and np.array_equal(raw, want)
    cls, regr = synth.rpn_outputs(38, 63, 9, 31)
    from helpers import FakeImage, FakeRpn
    mgr = det_util.DetTrainingManager(FakeRpn(cls, regr, conv=np.zeros((1, 38, 63, 8), np.float32)),
                                      synth.VOC_CLASS_MAPPING, lambda d: d, anchor_dims=O.anchor_table([128, 256, 512]))
    conv, rois = mgr.get_det_inputs(FakeImage("a", 1000, 600, [], data=np.zeros((600, 1000, 3), np.float32)))
    dense = det_util._get_rois(regr, O.anchor_table([128, 256, 512]), 16)
    wb, wp, _ = O.topk_proposals(dense.copy(), cls.reshape(-1), 8000)
    pick = O.greedy_nms(wb, wp, 0.7, 300)
    assert mgr.conv_only and conv.shape == (1, 38, 63, 8) and rois.dtype == np.int16 and np.array_equal(rois, wb[pick])


def test_pipeline_host_staged_equals_device_path(ops):
    """ProposalRoiPipeline.__call__ (pinned host inputs, chunked upload overlapping compute) must equal the
    device-resident path and the per-stage ops, for batch sizes that are not a multiple of the chunk."""
    import torch
    from faster_rcnn_b200 import synth
    from faster_rcnn_b200.pipeline import ProposalRoiPipeline
    dims = O.anchor_table([128, 256, 512])
    b = 5
    pairs = [synth.rpn_outputs(19, 25, 9, 40 + i, clustered=bool(i % 2)) for i in range(b)]
    cls, regr = np.concatenate([p[0] for p in pairs]), np.concatenate([p[1] for p in pairs])
    feat = np.concatenate([synth.feature_map(19, 25, 32, 50 + i) for i in range(b)])
    pipe = ProposalRoiPipeline(dims, 16, 2000, 0.7, 300, 64, 7, h2d_chunk=2)
    rois, scores, count, pooled = pipe(cls, regr, feat)
    d_rois, d_scores, d_count, d_padded, d_pooled = pipe.run_device(dev(cls), dev(regr), dev(feat))
    assert np.array_equal(rois, host(d_rois)) and np.array_equal(scores, host(d_scores)) and np.array_equal(count, host(d_count))
    assert torch.equal(pooled, d_pooled)
    keep = rois.copy()
    again = pipe(cls[::-1].copy(), regr[::-1].copy(), feat[::-1].copy())     # staging buffers are reused across calls ...
    assert not np.shares_memory(again[0], rois) and np.array_equal(rois, keep)   # ... but results are the caller's own arrays
    assert np.array_equal(again[0], rois[::-1]) and torch.equal(again[3], d_pooled.flip(0))
    from oracle import roi_oracle as R
    n0 = int(count[0])
    assert np.array_equal(host(pooled)[0, :n0], R.roi_resize_fwd(feat[0], rois[0, :n0], 7))
    # max mode through the same host-staged call (the north-star's pooling variant)
    mpipe = ProposalRoiPipeline(dims, 16, 2000, 0.7, 300, 64, 7, "max", h2d_chunk=2)
    m_rois, _, m_count, (m_pooled, m_arg) = mpipe(cls, regr, feat)
    assert np.array_equal(m_rois, rois) and np.array_equal(m_count, count)
    dm_pooled, dm_arg = mpipe.run_device(dev(cls), dev(regr), dev(feat))[4]
    assert torch.equal(m_pooled, dm_pooled) and torch.equal(m_arg, dm_arg)
    wout, warg = R.roi_max_fwd(feat[0], rois[0, :n0], 7)
    assert np.array_equal(host(m_pooled)[0, :n0], wout) and np.array_equal(host(m_arg)[0, :n0], warg)


def test_pipeline_cuda_graph_replay_equals_eager(ops):
    """ProposalRoiPipeline.capture: one graph launch per step, same rois / counts / pooled features as the eager path,
    also after the inputs change and after other calls have used (and grown) the shared handle's arena."""
    import torch
    from faster_rcnn_b200 import synth
    from faster_rcnn_b200.pipeline import ProposalRoiPipeline
    from faster_rcnn_b200.util import get_anchors
    dims = get_anchors([128, 256, 512])
    rows, cols, ch, b = 19, 25, 64, 3
    pipe = ProposalRoiPipeline(dims, 16, 2000, 0.7, 300, 64, 7, "resize")

    def inputs(seed):
        pairs = [synth.rpn_outputs(rows, cols, len(dims), seed + i, clustered=True) for i in range(b)]
        return (dev(np.concatenate([p[0] for p in pairs])), dev(np.concatenate([p[1] for p in pairs])),
                torch.randn((b, rows, cols, ch), device="cuda", generator=torch.Generator("cuda").manual_seed(seed)))

    a = inputs(10)
    run = pipe.capture(*a)
    for seed in (10, 20):
        x = inputs(seed)
        want = [t.clone() for t in pipe.run_device(*x)]
        ops.proposals(*inputs(99)[1::-1], dims, 16, 8000, 0.7, 300)          # unrelated call on the shared handle
        got = run(*x)
        torch.cuda.synchronize()
        for w, g in zip(want, got):
            assert torch.equal(w, g)


@pytest.mark.parametrize("levels,k,max_boxes", [(20, 8000, 300), (3, 3000, 300), (1, 500, 64)])
def test_fused_proposals_with_tied_scores(ops, levels, k, max_boxes):
    """Saturated / quantised RPN scores (a trained sigmoid head outputs exact 1.0s): the top-k boundary and the NMS visit
    order among equal scores follow the documented total order (two stable sorts, like the reference's argsort calls
    with kind='stable'), also on the fused decode -> top-k -> NMS path."""
    dims, cls, regr = _synth(38, 63, [128, 256, 512], 31, True)
    cls = (np.ceil(cls * levels) / levels).astype(np.float32)
    dense = host(ops.decode_topk(dev(regr), dev(cls), dims, 16, k, want_dense=True)[4])[0]
    wb, wp, widx = O.topk_proposals(dense.copy(), cls.reshape(-1), k, stable=True)
    boxes, scores, index, count = ops.decode_topk(dev(regr), dev(cls), dims, 16, k)
    n = int(host(count)[0])
    assert n == len(wb) and np.array_equal(host(index)[0, :n], widx) and np.array_equal(host(boxes)[0, :n], wb)
    pick = O.greedy_nms(wb, wp, 0.7, max_boxes, stable=True)
    rois, sc, cnt = ops.proposals(dev(regr), dev(cls), dims, 16, k, 0.7, max_boxes)
    m = int(host(cnt)[0])
    assert m == len(pick) and np.array_equal(host(rois)[0, :m], wb[pick]) and np.array_equal(host(sc)[0, :m], wp[pick])


def test_one_handle_two_streams_share_scratch_safely(ops):
    """The per-handle scratch arena is reused by every call; a call on another stream must first wait for the previous
    stream's work (capi.cu stream_handover).  Back-to-back fused proposal calls on two streams without any host
    synchronisation give the single-stream results."""
    import torch
    dims = O.anchor_table([128, 256, 512])
    a = _synth(38, 63, [128, 256, 512], 71, True)
    b = _synth(38, 63, [128, 256, 512], 72, True)
    big = [torch.cat([dev(x[i])] * 32) for x in (a, b) for i in (1, 2)]       # cls_a, regr_a, cls_b, regr_b, 32 images each
    want_a = [t.clone() for t in ops.proposals(big[1], big[0], dims, 16, 12000, 0.7, 2000)]
    want_b = [t.clone() for t in ops.proposals(big[3], big[2], dims, 16, 12000, 0.7, 2000)]
    torch.cuda.synchronize()
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    for _ in range(3):
        with torch.cuda.stream(s1):
            got_a = ops.proposals(big[1], big[0], dims, 16, 12000, 0.7, 2000)
        with torch.cuda.stream(s2):
            got_b = ops.proposals(big[3], big[2], dims, 16, 12000, 0.7, 2000)
        torch.cuda.synchronize()
        for w, g in zip(want_a + want_b, list(got_a) + list(got_b)):
            assert torch.equal(w, g)


def _score_family(name, n, rng):
    """Objectness arrays that exercise every slice-membership path of proposals_kernel: the bucket window is
    [2^-32, 1) with 64 buckets per octave, everything else lands in the two clamped end buckets."""
    u = (rng.permutation(n).astype(np.float64) + 0.5) / n
    if name == "logits":                       # negative and > 1: both clamped buckets, fall-back selects
        return ((u - 0.5) * 30.0).astype(np.float32)
    if name == "sigmoid_of_logits":            # a trained head: mass near 0 and a saturated top
        return (1.0 / (1.0 + np.exp(-(u - 0.7) * 40.0))).astype(np.float32)
    if name == "tiny":                         # everything below the window
        return (u * 1e-12).astype(np.float32)
    if name == "above_one":
        return (1.0 + u * 5.0).astype(np.float32)
    if name == "saturated_top":                # thousands of exact 1.0 (ties inside the clamped top bucket)
        return np.minimum(u * 1.6, 1.0).astype(np.float32)
    if name == "one_bucket":                   # all scores inside one 1/64-octave bucket: stash overflow -> fall-back
        return (0.75 + u * 1e-3).astype(np.float32)
    if name == "two_values":
        return np.where(u < 0.5, 0.25, 0.5).astype(np.float32)
    raise KeyError(name)


@pytest.mark.parametrize("family", ["logits", "sigmoid_of_logits", "tiny", "above_one", "saturated_top", "one_bucket",
                                    "two_values"])
@pytest.mark.parametrize("k", [8000, 12000, 700])
def test_decode_topk_score_distributions(ops, family, k):
    dims = O.anchor_table([128, 256, 512])
    from faster_rcnn_b200 import synth
    _, regr = synth.rpn_outputs(38, 63, 9, 300 + k, clustered=True)
    rng = np.random.default_rng(k)
    cls = _score_family(family, 38 * 63 * 9, rng).reshape(1, 38, 63, 9)
    _check_decode_topk(ops, dims, cls, regr, k)


def test_decode_topk_large_k(ops):
    """k = 20000 of 36864 anchors: sixteen CTAs per image, 2048-key sorts."""
    dims = O.anchor_table([128, 256, 512])
    from faster_rcnn_b200 import synth
    cls, regr = synth.rpn_outputs(64, 64, 9, 901, clustered=True)
    _check_decode_topk(ops, dims, cls, regr, 20000)
    _check_decode_topk(ops, dims, cls, regr, 32768)


def test_decode_topk_batch_uses_narrow_ctas(ops):
    """24 images x 8 slices exceeds the SM count: the 256-thread variant of the kernel, several CTAs per SM."""
    dims = O.anchor_table([128, 256, 512])
    from faster_rcnn_b200 import synth
    b = 24
    pairs = [synth.rpn_outputs(38, 63, 9, 700 + i, clustered=bool(i % 2)) for i in range(b)]
    cls, regr = np.concatenate([p[0] for p in pairs]), np.concatenate([p[1] for p in pairs])
    cls[3] = _score_family("logits", cls[3].size, np.random.default_rng(3)).reshape(cls[3].shape)
    cls[4] = _score_family("one_bucket", cls[4].size, np.random.default_rng(4)).reshape(cls[4].shape)
    boxes, scores, index, count, dense = ops.decode_topk(dev(regr), dev(cls), dims, 16, 8000, want_dense=True)
    boxes, scores, index, count, dense = host(boxes), host(scores), host(index), host(count), host(dense)
    for i in range(b):
        wb, wp, widx = O.topk_proposals(dense[i].copy(), cls[i].reshape(-1), 8000)
        n = int(count[i])
        assert n == len(wb), i
        assert np.array_equal(index[i, :n], widx) and np.array_equal(boxes[i, :n], wb) and np.array_equal(scores[i, :n], wp), i
        assert np.all(index[i, n:] == -1)


def test_decode_topk_repeated_launches_are_identical(ops):
    """proposals_kernel exchanges histograms between the CTAs of a cluster; every launch must give the same answer
    whatever the timing (L2 flushed or hot, wide or narrow CTAs, back-to-back launches)."""
    import torch
    dims = O.anchor_table([128, 256, 512])
    from faster_rcnn_b200 import synth
    junk = torch.empty(160 << 20, dtype=torch.uint8, device="cuda")
    for rows, cols, k, b in [(10, 12, 8000, 1), (38, 63, 8000, 1), (20, 20, 1500, 3), (38, 63, 8000, 24)]:
        pairs = [synth.rpn_outputs(rows, cols, 9, 800 + i, clustered=bool(i % 2)) for i in range(b)]
        cls, regr = dev(np.concatenate([p[0] for p in pairs])), dev(np.concatenate([p[1] for p in pairs]))
        ref = [t.clone() for t in ops.decode_topk(regr, cls, dims, 16, k)]
        for rep in range(60):
            if rep % 3 == 1:
                junk.fill_(rep)
            got = ops.decode_topk(regr, cls, dims, 16, k)
            assert all(torch.equal(a, g) for a, g in zip(ref, got)), (rows, cols, k, b, rep)
