"""Pins the numpy oracle against the committed golden vectors (tests/golden/*.npz, produced from the
unmodified reference by tests/golden/make_golden.py).  Runs anywhere, no reference checkout, no GPU."""
import os
import random

import numpy as np
import pytest

from oracle import frcnn_oracle as O

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


def test_anchor_tables():
    g = golden("anchors")
    assert np.array_equal(O.anchor_table([128, 256, 512]), g["voc"])
    assert np.array_equal(O.anchor_table(), g["default"])
    assert g["voc"].tolist()[:3] == [[128, 128], [90, 181], [181, 90]]
    assert (g["default"] // 16).tolist()[:6] == [[1, 1], [0, 1], [1, 0], [2, 2], [1, 2], [2, 1]]


@pytest.mark.parametrize("tag", ["voc_small", "voc_clustered", "kitti_small"])
def test_proposal_stage(tag):
    g = golden("proposals_" + tag)
    dense = O.proposals_from_rpn(g["regr"].copy(), g["anchor_dims"], 16)
    assert dense.dtype == np.float32 and np.array_equal(dense, g["dense"])
    assert np.array_equal(O.valid_box_indices(dense), g["valid"])
    b, p, idx = O.topk_proposals(dense, g["cls"].reshape(-1), int(g["k"]))
    assert b.dtype == np.int16 and np.array_equal(b, g["topk_boxes"]) and np.array_equal(p, g["topk_probs"])
    assert np.array_equal(idx, g["topk_index"])
    nb, npb = O.nms(b, p, 0.7, int(g["max_boxes"]))
    assert np.array_equal(nb, g["nms_boxes"]) and np.array_equal(npb, g["nms_probs"])


@pytest.mark.parametrize("tag", ["000005_resnet", "000005_vgg", "synth50"])
def test_rpn_labels(tag):
    g = golden("rpn_labels_" + tag)
    rows, cols = (int(v) for v in g["conv"])
    w, h = (int(v) for v in g["img_wh"])
    cu, ip, bb = O.label_anchors(w, h, g["gt"], rows, cols, g["anchor_dims"], 16)
    assert np.array_equal(cu, g["can_use"]) and np.array_equal(ip, g["is_pos"]) and np.array_equal(bb, g["bbreg"])
    random.seed(int(g["py_random_seed"]))
    cu2 = O.sample_rpn(ip, cu.copy())
    y_class, y_bbreg = O.pack_rpn_targets(cu2, ip, bb, rows, cols, len(g["anchor_dims"]))
    assert y_class.dtype == np.bool_ and np.array_equal(y_class, g["y_class"])
    assert y_bbreg.dtype == np.float32 and np.array_equal(y_bbreg, g["y_bbreg"])


def test_known_answer_000005():
    """the reference's labels on the one VOC image it ships."""
    g = golden("rpn_labels_000005_resnet")
    assert tuple(g["conv"]) == (38, 50) and tuple(g["img_wh"]) == (800, 600)
    assert np.where(g["is_pos"])[0].tolist() == [7454, 10586, 11036, 11486, 11963, 12413, 12863, 13079, 13529, 13680, 13979]
    assert int(g["can_use"].sum()) == 5287
    assert abs(float(np.abs(g["bbreg"]).sum()) - 33.408202) < 1e-4


def test_det_labels_and_sampling():
    g = golden("det_labels")
    gt64 = np.array([[v * (1 / 16) for v in row] for row in g["gt_pixels"].tolist()], dtype=np.float64)
    rois, y_cls, y_tr = O.label_rois(g["rois"], gt64, g["gt_cls"], 21)
    assert np.array_equal(rois, g["eligible_rois"]) and np.array_equal(y_cls, g["y_class_num"])
    assert y_tr.dtype == np.float32 and np.array_equal(y_tr, g["y_transform"])
    np.random.seed(int(g["np_random_seed"]))
    assert O.sample_det(y_cls[:, -1] == 0, 64) == g["sampled"].tolist()


def test_det_postprocess():
    g = golden("det_postprocess")
    dets = O.det_postprocess(g["rois"], g["out_cls"], g["out_reg"], 20, int(g["stride"]), float(g["resize_ratio"]))
    assert len(dets) == len(g["det_cls"]) > 50
    assert [d[0] for d in dets] == g["det_cls"].tolist()
    assert np.array_equal(np.array([d[1] for d in dets]), g["det_boxes"])
    assert np.array_equal(np.array([d[2] for d in dets], dtype=np.float32), g["det_probs"])


def test_nms_f64_and_iou():
    g = golden("nms_f64_iou")
    nb, npb = O.nms(g["boxes"], g["probs"], 0.5, 2000)
    assert np.array_equal(nb, g["nms_boxes"]) and np.array_equal(npb, g["nms_probs"])
    assert np.array_equal(O.iou_matrix(g["anchors"], g["gt"]), g["iou_f32"])
    assert np.array_equal(O.iou_matrix(g["rois_i16"], g["gt_feat"]), g["iou_i16"])
    assert np.array_equal(O.pixel_anchors(6, 9, O.anchor_table([128, 256, 512]), 16), g["anchors"])
    assert O.nms(g["boxes"][:0], g["probs"][:0]) == []


def test_labels_on_real_voc_ground_truth():
    """200 real VOC2007 annotation sets (SURVEY 8c iv): the oracle's labels hash to the reference's."""
    import hashlib
    g = golden("voc_gt_200")
    dims = O.anchor_table([128, 256, 512])
    for i in range(0, len(g["names"]), 4):                       # every 4th image keeps the CPU suite fast
        w, h = (int(v) for v in g["img_wh"][i])
        gt = g["gt"][g["offsets"][i]:g["offsets"][i + 1]]
        rows, cols = O.conv_dims_resnet(h, w)
        cu, ip, _ = O.label_anchors(w, h, gt, rows, cols, dims, 16)
        assert hashlib.sha1(cu.tobytes() + ip.tobytes()).hexdigest()[:16] == str(g["label_sha1"][i])
        assert int(ip.sum()) == int(g["n_pos"][i]) and int(cu.sum()) == int(g["n_use"][i])


def test_numpy_float32_exp_kernel_restatement():
    """np_expf (csrc/common.cuh) restates numpy's SIMD float32 exp; its Python twin in tests/helpers.py must reproduce
    np.exp bit for bit wherever numpy dispatches to that kernel (this container: AVX512F), incl. the range ends."""
    from helpers import np_exp_f32_simd, numpy_exp_is_simd_kernel
    if not numpy_exp_is_simd_kernel():
        pytest.skip("this machine's numpy does not use its SIMD float32 exp kernel")
    rng = np.random.default_rng(11)
    x = np.concatenate([(rng.standard_normal(2_000_000) * s).astype(np.float32) for s in (0.05, 0.5, 3.0, 20.0)])
    assert np.array_equal(np.exp(x), np_exp_f32_simd(x))
    # not the correctly rounded exp: the reason a double-precision exp on the device would NOT reproduce the reference
    exact = np.exp(x.astype(np.float64)).astype(np.float32)
    assert 0.2 < np.mean(np.exp(x) != exact) < 0.6
