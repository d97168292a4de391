"""GPU parity: training-target assignment (K-c anchor labelling, sampling + packing, K-c' RoI
labelling) and detector post-processing (K-e) through the C ABI vs the numpy oracle and the
reference-generated golden vectors.  Labels, indices and int boxes bit-exact; float32 regression
targets bit-exact except for the documented `log` exception (<= 1 ulp, asserted)."""
import random

import numpy as np
import pytest

from helpers import FakeImage, FakeRpn, dev, golden, host
from oracle import frcnn_oracle as O
from oracle import roi_oracle as R

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops():
    from faster_rcnn_b200 import ops as _ops
    return _ops


def assert_f32_ulp(got, want, max_ulp=1):
    """float32 arrays equal up to `max_ulp` units in the last place (device log/exp vs libm)."""
    got, want = np.ascontiguousarray(got, np.float32), np.ascontiguousarray(want, np.float32)
    assert got.shape == want.shape
    gi, wi = got.view(np.int32).astype(np.int64), want.view(np.int32).astype(np.int64)
    gi = np.where(gi < 0, -(gi & 0x7fffffff), gi)
    wi = np.where(wi < 0, -(wi & 0x7fffffff), wi)
    worst = int(np.abs(gi - wi).max()) if got.size else 0
    assert worst <= max_ulp, "max difference %d ulp" % worst
    return worst


def _label(ops, gt_list, wh_list, rows, cols, dims):
    g_max = max(len(g) for g in gt_list)
    gt = np.zeros((len(gt_list), g_max, 4), np.float32)
    for b, g in enumerate(gt_list):
        gt[b, :len(g)] = g
    n_gt = np.array([len(g) for g in gt_list], np.int32)
    out = ops.label_anchors(dev(gt), dev(n_gt), dev(np.array(wh_list, np.int32)), rows, cols, dims, 16)
    return [host(t) for t in out]


@pytest.mark.parametrize("tag", ["000005_resnet", "000005_vgg", "synth50"])
def test_label_anchors_vs_golden(ops, tag):
    g = golden("rpn_labels_" + tag)
    rows, cols = (int(v) for v in g["conv"])
    cu, ip, bb, counts = _label(ops, [g["gt"]], [g["img_wh"]], rows, cols, g["anchor_dims"])
    assert np.array_equal(cu[0].view(np.bool_), g["can_use"]) and np.array_equal(ip[0].view(np.bool_), g["is_pos"])
    assert_f32_ulp(bb[0], g["bbreg"])
    assert counts[0].tolist() == [int((g["can_use"] & g["is_pos"]).sum()), int((g["can_use"] & ~g["is_pos"]).sum())]


@pytest.mark.parametrize("rows_cols,wh,scales,n_gts", [
    ((38, 63), (1000, 600), [128, 256, 512], [50, 3, 1, 17]),       # C4 VOC shape, ragged GT counts
    ((38, 94), (1500, 600), None, [50, 50]),                         # KITTI 18 anchors / location
])
def test_label_anchors_batch_vs_oracle(ops, rows_cols, wh, scales, n_gts):
    from faster_rcnn_b200 import synth
    rows, cols = rows_cols
    dims = O.anchor_table(scales) if scales else O.anchor_table()
    gts = [np.array([g[1:] for g in synth.gt_boxes(n, wh[0], wh[1], 40 + i)], np.float32) for i, n in enumerate(n_gts)]
    cu, ip, bb, counts = _label(ops, gts, [wh] * len(gts), rows, cols, dims)
    for b, gt in enumerate(gts):
        wcu, wip, wbb = O.label_anchors(wh[0], wh[1], gt, rows, cols, dims, 16)
        assert np.array_equal(cu[b].view(np.bool_), wcu) and np.array_equal(ip[b].view(np.bool_), wip)
        assert_f32_ulp(bb[b], wbb)
        assert counts[b].tolist() == [int((wcu & wip).sum()), int((wcu & ~wip).sum())]


def test_label_anchors_on_real_voc_ground_truth(ops):
    """200 real VOC2007 annotation sets: labels must hash to the digests recorded from the reference;
    images of equal conv shape go through the kernel as one batch."""
    import hashlib
    g = golden("voc_gt_200")
    dims = O.anchor_table([128, 256, 512])
    groups = {}
    for i in range(len(g["names"])):
        w, h = (int(v) for v in g["img_wh"][i])
        groups.setdefault(tuple(O.conv_dims_resnet(h, w)), []).append(i)
    checked = 0
    for (rows, cols), idxs in groups.items():
        gts = [g["gt"][g["offsets"][i]:g["offsets"][i + 1]] for i in idxs]
        cu, ip, bb, counts = _label(ops, gts, [g["img_wh"][i] for i in idxs], rows, cols, dims)
        for b, i in enumerate(idxs):
            cub, ipb = cu[b].view(np.bool_), ip[b].view(np.bool_)
            assert hashlib.sha1(cub.tobytes() + ipb.tobytes()).hexdigest()[:16] == str(g["label_sha1"][i]), g["names"][i]
            assert int(ipb.sum()) == int(g["n_pos"][i]) and int(cub.sum()) == int(g["n_use"][i])
            checked += 1
    assert checked == 200


def test_label_anchors_degenerate_gt(ops):
    """GT outside every anchor (max IoU 0 -> no forced positive), duplicated GTs (shared arg-max anchor)."""
    dims = O.anchor_table([128, 256, 512])
    gt = np.array([[5000, 5000, 5100, 5100], [100, 100, 300, 260], [100, 100, 300, 260]], np.float32)
    cu, ip, bb, _ = _label(ops, [gt], [(800, 600)], 38, 50, dims)
    wcu, wip, wbb = O.label_anchors(800, 600, gt, 38, 50, dims, 16)
    assert np.array_equal(cu[0].view(np.bool_), wcu) and np.array_equal(ip[0].view(np.bool_), wip)
    assert_f32_ulp(bb[0], wbb)


@pytest.mark.parametrize("tag", ["000005_resnet", "synth50"])
def test_rpn_manager_dropin_vs_golden(tag):
    """rpn_util.RpnTrainingManager with the reference's signature; sampling replays `random` seeded as
    in the reference's own test (train_rpn_test.py:17-18)."""
    from faster_rcnn_b200 import rpn_util
    g = golden("rpn_labels_" + tag)
    w, h = (int(v) for v in g["img_wh"])
    img = FakeImage(tag, w, h, [("chair", *row) for row in g["gt"].tolist()])
    conv = tuple(int(v) for v in g["conv"])
    mgr = rpn_util.RpnTrainingManager(lambda hh, ww: conv, 16, preprocess_func=None, anchor_dims=g["anchor_dims"])
    mgr._process(img)
    c = mgr._cache[img.cache_key]
    assert c['can_use'].dtype == np.bool_ and np.array_equal(c['can_use'], g["can_use"])
    assert np.array_equal(c['is_pos'], g["is_pos"])
    assert_f32_ulp(c['bbreg_targets'], g["bbreg"])
    random.seed(int(g["py_random_seed"]))
    y_class, y_bbreg = mgr.rpn_y_true(img)
    assert img.cache_key not in mgr._cache                              # consumed like the reference
    assert y_class.dtype == np.bool_ and y_class.shape == g["y_class"].shape and np.array_equal(y_class, g["y_class"])
    assert y_bbreg.dtype == np.float32 and y_bbreg.shape == g["y_bbreg"].shape
    assert_f32_ulp(y_bbreg, g["y_bbreg"])
    # uncached path and the batched variant give the same answer with the same RNG state
    random.seed(int(g["py_random_seed"]))
    y2c, y2b = mgr.rpn_y_true(img)
    assert np.array_equal(y2c, y_class) and np.array_equal(y2b, y_bbreg)
    random.seed(int(g["py_random_seed"]))
    y3c, y3b = mgr.rpn_y_true_batch([img, img])
    assert np.array_equal(y3c[0], y_class[0]) and np.array_equal(y3b[0], y_bbreg[0])
    assert y3c[1].sum() > 0


def test_rpn_sampling_many_positives(ops):
    """> 128 positives and > 256 usable anchors: both random.sample branches (rpn_util.py:338-348)."""
    from faster_rcnn_b200 import rpn_util
    rng = np.random.default_rng(5)
    n = 9 * 40 * 30
    is_pos = rng.random(n) < 0.05
    can_use = rng.random(n) < 0.6
    random.seed(3)
    want = O.sample_rpn(is_pos.copy(), can_use.copy())
    random.seed(3)
    got = rpn_util._apply_sampling(is_pos, can_use)
    assert got is can_use and np.array_equal(got, want)
    assert int((got & is_pos).sum()) == 128 and int(got.sum()) == 256


def test_pack_rpn_targets_vs_oracle(ops):
    rng = np.random.default_rng(8)
    rows, cols, a, b = 7, 9, 9, 3
    n = rows * cols * a
    cu = (rng.random((b, n)) < 0.5)
    ip = (rng.random((b, n)) < 0.2)
    bb = rng.standard_normal((b, n, 4)).astype(np.float32)
    y_class, y_bbreg = ops.pack_rpn_targets(dev(cu.view(np.uint8)), dev(ip.view(np.uint8)), dev(bb), rows, cols, a)
    for i in range(b):
        wc, wb = O.pack_rpn_targets(cu[i], ip[i], bb[i], rows, cols, a)
        assert np.array_equal(host(y_class)[i].view(np.bool_), wc[0]) and np.array_equal(host(y_bbreg)[i], wb[0])


# ------------------------------------------------------------------------------------------------
# detector targets
# ------------------------------------------------------------------------------------------------
def test_label_rois_vs_golden(ops):
    g = golden("det_labels")
    gt64 = np.array([[v * (1 / 16) for v in row] for row in g["gt_pixels"].tolist()], np.float64)
    out = ops.label_rois(dev(g["rois"][None]), dev(gt64[None]), dev(g["gt_cls"][None], np.int32),
                         dev(np.array([len(gt64)], np.int32)), 21)
    rois, y_cls, y_tr, src, m = [host(t) for t in out]
    m = int(m[0])
    assert m == len(g["eligible_rois"])
    assert np.array_equal(rois[0, :m], g["eligible_rois"]) and np.array_equal(y_cls[0, :m], g["y_class_num"])
    assert_f32_ulp(y_tr[0, :m], g["y_transform"])
    assert np.array_equal(g["rois"][src[0, :m]], g["eligible_rois"])


def test_det_manager_dropin_vs_oracle():
    """DetTrainingManager.get_training_input: proposals (12000 -> NMS 2000) -> labelling -> 64-RoI sampling with
    numpy's global RNG seeded like the reference's test (train_det_test.py:3-6)."""
    from faster_rcnn_b200 import det_util, synth
    dims = O.anchor_table([128, 256, 512])
    cls, regr = synth.rpn_outputs(38, 63, 9, 51, clustered=True)
    gts = synth.gt_boxes(12, 1000, 600, 52)
    img = FakeImage("train", 1000, 600, gts, data=np.zeros((4, 4, 3), np.float32))
    mapping = synth.VOC_CLASS_MAPPING
    mgr = det_util.DetTrainingManager(FakeRpn(cls, regr), mapping, lambda d: d, anchor_dims=dims)
    np.random.seed(1337)
    first, rois, y_cls, y_tr = mgr.get_training_input(img)
    # oracle pipeline on the device's decoded boxes
    dense = det_util._get_rois(regr, dims, 16)
    wb, wp, _ = O.topk_proposals(dense.copy(), cls.reshape(-1), 12000)
    nms_rois = wb[O.greedy_nms(wb, wp, 0.7, 2000)]
    gt64 = np.array([[v * (1 / 16) for v in g[1:]] for g in gts], np.float64)
    gidx = np.array([mapping[g[0]] for g in gts])
    e_rois, w_cls, w_tr = O.label_rois(nms_rois, gt64, gidx, 21)
    np.random.seed(1337)
    sel = O.sample_det(w_cls[:, -1] == 0, 64)
    assert first.shape == (1, 4, 4, 3) and not mgr.conv_only
    assert rois.shape == (1, 64, 4) and rois.dtype == np.int16 and np.array_equal(rois[0], e_rois[sel])
    assert y_cls.shape == (1, 64, 21) and y_cls.dtype == np.int32 and np.array_equal(y_cls[0], w_cls[sel])
    assert y_tr.shape == (1, 64, 160) and y_tr.dtype == np.float32
    assert_f32_ulp(y_tr[0], w_tr[sel])
    # module-level function with the reference's signature
    r2, c2, t2 = det_util._rois_to_truth(nms_rois, img, mapping, stride=16)
    assert np.array_equal(r2, e_rois) and np.array_equal(c2, w_cls)
    assert_f32_ulp(t2, w_tr)


# ------------------------------------------------------------------------------------------------
# detector post-processing
# ------------------------------------------------------------------------------------------------
def test_det_postprocess_vs_golden(ops):
    from faster_rcnn_b200 import synth, voc_dets
    g = golden("det_postprocess")
    dets = voc_dets.postprocess(g["rois"], g["out_cls"], g["out_reg"], synth.VOC_CLASS_MAPPING,
                                float(g["resize_ratio"]), int(g["stride"]))
    assert len(dets) == len(g["det_cls"])
    assert [synth.VOC_CLASS_MAPPING[d['cls_name']] for d in dets] == g["det_cls"].tolist()
    assert np.array_equal(np.array([d['bbox'] for d in dets]), g["det_boxes"])
    assert np.array_equal(np.array([d['prob'] for d in dets], np.float32), g["det_probs"])


def test_det_postprocess_batch_vs_oracle(ops):
    """C2: batch of images, 320 rows (300 RoIs + the reference's padding duplicates), 21 classes."""
    from faster_rcnn_b200 import synth, voc_dets
    n_img = 6
    rois = np.stack([voc_dets.pad_roi_batches(synth.random_rois(300, 37, 62, 60 + i)) for i in range(n_img)])
    outs = [synth.detector_outputs(320, 21, 70 + i) for i in range(n_img)]
    out_cls, out_reg = np.stack([o[0] for o in outs]), np.stack([o[1] for o in outs])
    ratios = [1.6, 1.0, 0.75, 2.2, 1.6, 1.3]
    assert rois.shape == (n_img, 320, 4) and np.array_equal(rois[0, 300:], np.tile(rois[0, 256], (20, 1)))
    for thr in (0.0, 0.3):
        got = voc_dets.postprocess_batch(rois, out_cls, out_reg, synth.VOC_CLASS_MAPPING, ratios, 16, det_threshold=thr)
        for b in range(n_img):
            want = O.det_postprocess(rois[b], out_cls[b], out_reg[b], 20, 16, ratios[b], det_threshold=thr)
            assert len(got[b]) == len(want)
            for d, (wc, wbox, wp) in zip(got[b], want):
                assert synth.VOC_CLASS_MAPPING[d['cls_name']] == wc and d['bbox'].tolist() == wbox.tolist() and d['prob'] == wp


def test_get_dets_dropin():
    """voc_dets.get_dets signature with a fake detector: batching/padding + post-processing."""
    from faster_rcnn_b200 import det_util, synth, voc_dets
    dims = O.anchor_table([128, 256, 512])
    cls, regr = synth.rpn_outputs(37, 62, 9, 81)
    mapping = synth.VOC_CLASS_MAPPING
    mgr = det_util.DetTrainingManager(FakeRpn(cls, regr, conv=np.zeros((1, 37, 62, 4), np.float32)), mapping,
                                      lambda d: d, anchor_dims=dims)
    calls = []

    class Detector:
        def predict(self, x):
            conv, batch = x
            assert batch.shape == (1, 64, 4)
            oc, orr = synth.detector_outputs(64, 21, 90 + len(calls))
            calls.append(batch[0].copy())
            return oc[None], orr[None]
    img = FakeImage("x", 992, 592, [], data=np.zeros((4, 4, 3), np.float32))
    dets = voc_dets.get_dets(mgr, Detector(), img, 1.6)
    _, rois = mgr.get_det_inputs(img)
    padded = voc_dets.pad_roi_batches(rois)
    assert np.array_equal(np.concatenate(calls), padded)
    oc = np.concatenate([synth.detector_outputs(64, 21, 90 + i)[0] for i in range(len(calls))])
    orr = np.concatenate([synth.detector_outputs(64, 21, 90 + i)[1] for i in range(len(calls))])
    want = O.det_postprocess(padded, oc, orr, 20, 16, 1.6)
    assert len(dets) == len(want) > 0
    for d, (wc, wbox, wp) in zip(dets, want):
        assert mapping[d['cls_name']] == wc and d['bbox'].tolist() == wbox.tolist() and d['prob'] == wp


def test_detection_pipeline_equals_per_image_get_dets():
    """DetectionPipeline (batched, device-resident) == voc_dets.get_dets image by image, incl. an image with fewer
    than 300 proposals whose unused fixed-shape rows must be ignored."""
    import torch
    from faster_rcnn_b200 import det_util, synth, voc_dets
    from faster_rcnn_b200.pipeline import DetectionPipeline
    dims = O.anchor_table([128, 256, 512])
    mapping = synth.VOC_CLASS_MAPPING
    shapes = [(37, 62, False), (37, 62, True), (37, 62, False)]
    pairs = [synth.rpn_outputs(r, c, 9, 700 + i, clustered=cl) for i, (r, c, cl) in enumerate(shapes)]
    pairs[2][1][..., 2::4] = 20.0            # image 2: every box blows up to the whole map -> identical boxes,
    pairs[2][1][..., 3::4] = 20.0            # NMS keeps a handful, most fixed-shape rows are unused
    cls, regr = np.concatenate([p[0] for p in pairs]), np.concatenate([p[1] for p in pairs])
    feat = np.concatenate([synth.feature_map(37, 62, 16, 710 + i) for i in range(3)])
    gen = torch.Generator(device="cuda").manual_seed(5)
    w_cls = torch.randn((16, 21), generator=gen, device="cuda")
    w_reg = torch.randn((16, 80), generator=gen, device="cuda")

    def head(pooled, rois):                                  # tiny deterministic stand-in for the dense detector head
        x = pooled.mean(dim=(2, 3)) + rois.to(torch.float32).sum(dim=2, keepdim=True) * 0.01
        return torch.softmax(x @ w_cls, dim=2), x @ w_reg

    ratios = [1.6, 1.0, 2.0]
    pipe = DetectionPipeline(head, mapping, anchor_dims=dims)
    got = pipe.detect(cls, regr, feat, ratios)

    class Detector:                                          # the same head behind the reference's Keras-style interface
        def predict(self, x):
            conv, batch = x
            pooled = ops_mod.roi_forward(dev(conv), dev(batch), 7)
            oc, orr = head(pooled, dev(batch))
            return host(oc), host(orr)
    from faster_rcnn_b200 import ops as ops_mod
    n_rois = []
    for b in range(3):
        mgr = det_util.DetTrainingManager(FakeRpn(cls[b:b + 1], regr[b:b + 1], conv=feat[b:b + 1]), mapping, lambda d: d,
                                          anchor_dims=dims)
        img = FakeImage("i%d" % b, 992, 592, [], data=np.zeros((4, 4, 3), np.float32))
        want = voc_dets.get_dets(mgr, Detector(), img, ratios[b])
        n_rois.append(len(mgr.get_det_inputs(img)[1]))
        assert len(got[b]) == len(want)
        for d, w in zip(got[b], want):
            assert d['cls_name'] == w['cls_name'] and d['bbox'].tolist() == w['bbox'].tolist() and d['prob'] == w['prob']
    assert n_rois[0] == 300 and n_rois[2] < 256, n_rois     # image 2 really exercises the ignored rows


# ------------------------------------------------------------------------------------------------
# stand-alone helpers (util.py surface)
# ------------------------------------------------------------------------------------------------
def test_util_dropin_vs_golden():
    from faster_rcnn_b200 import rpn_util, util
    g = golden("nms_f64_iou")
    assert np.array_equal(util.cross_ious(g["anchors"], g["gt"]), g["iou_f32"])
    assert np.array_equal(util.cross_ious(g["rois_i16"], g["gt_feat"]), g["iou_i16"])
    dims = O.anchor_table([128, 256, 512])
    anc = rpn_util._get_all_anchor_coords(6, 9, dims, 16)
    assert anc.dtype == np.float32 and np.array_equal(anc, g["anchors"])
    assert np.array_equal(rpn_util._get_out_of_bounds_idxs(anc, 144, 96), O.out_of_bounds_indices(anc, 144, 96))
    rng = np.random.default_rng(2)
    boxes = O.feature_anchors(9, 11, dims // 16)
    deltas = (rng.standard_normal((len(boxes), 4)) * [0.1, 0.1, 0.3, 0.3]).astype(np.float32)
    want = O.decode_boxes(boxes.copy(), deltas)
    got = util.transform_np_inplace(boxes, deltas)
    assert got is boxes and np.sum(np.any(got != want, axis=1)) <= 1


def test_det_training_pipeline_equals_per_image_manager():
    """pipeline.DetTrainingPipeline (batched, device-resident) == DetTrainingManager.get_training_input called on each
    image in turn with the same numpy RNG stream, incl. an image whose GT overlaps no proposal (4 x None)."""
    import torch
    from faster_rcnn_b200 import det_util, synth
    from faster_rcnn_b200.pipeline import DetTrainingPipeline
    dims = O.anchor_table([128, 256, 512])
    mapping = synth.VOC_CLASS_MAPPING
    rows, cols, ch, b = 19, 25, 32, 4
    heads = [synth.rpn_outputs(rows, cols, 9, 70 + i, clustered=True) for i in range(b)]
    gts = [synth.gt_boxes(6, 400, 300, 80 + i) for i in range(b)]
    gts[2] = [("cat", 1, 1, 3, 3)]                                   # 2x2 px object: IoU < 0.1 with every proposal
    images = [FakeImage("im%d" % i, 400, 300, g, data=np.zeros((2, 2, 3), np.float32)) for i, g in enumerate(gts)]
    feat = np.random.default_rng(5).standard_normal((b, rows, cols, ch), dtype=np.float32)

    np.random.seed(1337)
    want = []
    for (cls, regr), img in zip(heads, images):
        mgr = det_util.DetTrainingManager(FakeRpn(cls, regr), mapping, lambda d: d, anchor_dims=dims)
        want.append(mgr.get_training_input(img))
    assert want[2] == (None, None, None, None) and want[0][1] is not None

    pipe = DetTrainingPipeline(mapping, dims)
    np.random.seed(1337)
    rois, y_cls, y_tr, pooled, has = pipe(np.concatenate([h[0] for h in heads]), np.concatenate([h[1] for h in heads]), feat, images)
    assert has.tolist() == [True, True, False, True]
    rois, y_cls, y_tr = host(rois), host(y_cls), host(y_tr)
    for i in range(b):
        if want[i][1] is None:
            assert not rois[i].any() and not y_cls[i].any() and not y_tr[i].any()
            continue
        assert np.array_equal(rois[i], want[i][1][0]) and np.array_equal(y_cls[i], want[i][2][0])
        assert np.array_equal(y_tr[i], want[i][3][0])
        assert np.array_equal(host(pooled)[i], R.roi_resize_fwd(feat[i], rois[i], 7))
