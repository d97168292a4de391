"""Host-side logic of the drop-in package that needs no GPU: RNG replay, scalar box maths, padding,
sharding.  Checked against the oracle (itself pinned to the reference)."""
import random

import numpy as np
import pytest

from faster_rcnn_b200 import det_util, parallel, rpn_util, shared_constants, synth, util, voc_dets
from faster_rcnn_b200.custom_layers import RoiResizeConv
from oracle import frcnn_oracle as O


def test_constants_and_anchor_tables():
    assert np.array_equal(shared_constants.DEFAULT_ANCHORS, O.anchor_table())
    assert shared_constants.BBREG_MULTIPLIERS.dtype == np.float32
    assert np.array_equal(util.get_anchors([128, 256, 512]), O.anchor_table([128, 256, 512]))
    assert (rpn_util.POS_OVERLAP, rpn_util.NEG_OVERLAP, rpn_util.SAMPLE_SIZE, rpn_util.MAX_POS_SAMPLES) == (0.7, 0.3, 256, 128)
    assert (det_util.CLASSIFIER_MIN_OVERLAP, det_util.CLASSIFIER_POS_OVERLAP) == (0.1, 0.5)


def test_rng_replay_matches_reference_call_order():
    """_draw_switch_offs must consume Python's `random` exactly like rpn_util._apply_sampling."""
    rng = np.random.default_rng(0)
    for n_pos_frac, n_use_frac in ((0.05, 0.6), (0.001, 0.5), (0.3, 0.01)):
        is_pos = rng.random(5000) < n_pos_frac
        can_use = rng.random(5000) < n_use_frac
        random.seed(11)
        want = O.sample_rpn(is_pos, can_use.copy())
        random.seed(11)
        off_pos, off_neg = rpn_util._draw_switch_offs(rpn_util._host_counts(can_use, is_pos))
        got = can_use.copy()
        got[np.where(is_pos & can_use)[0][off_pos[0]]] = False
        got[np.where(~is_pos & can_use)[0][off_neg[0]]] = False
        assert np.array_equal(got, want)
        assert random.random() == (random.seed(11), O.sample_rpn(is_pos, can_use.copy()), random.random())[2]


def test_det_sampling_matches_oracle():
    flags = np.random.default_rng(1).random(300) < 0.2
    for f in (flags, np.zeros_like(flags), np.ones_like(flags), flags & (np.arange(300) < 40), flags[:50]):
        np.random.seed(1337)
        want = O.sample_det(f, 64)
        np.random.seed(1337)
        assert det_util._get_det_samples(f, 64) == want


def test_scalar_box_maths():
    rng = np.random.default_rng(2)
    for _ in range(20):
        a = np.array([10, 20, 138, 148]) + rng.integers(0, 5, 4)
        g = rng.uniform(0, 300, 4).astype(np.float32)
        g[2:] += g[:2] + 5
        assert util.get_reg_params(a, g) == O.regression_params(a, g)
        box = np.array([3, 4, 20, 30], np.int16)
        t = rng.standard_normal(4).astype(np.float32) / shared_constants.BBREG_MULTIPLIERS
        assert util.transform(list(box), t) == O.decode_scalar(list(box), t)
    assert util.calc_iou([0, 0, 10, 10], [5, 5, 15, 15]) == pytest.approx(25 / 175)
    assert util.calc_iou([0, 0, 10, 10], [10, 10, 20, 20]) == 0.0
    y, x, a = rpn_util._idx_to_conv(7454, 50, 9)
    assert (y, x, a) == (16, 28, 2) and rpn_util._get_conv_center(x, y, 16) == (456, 264)


def test_one_hot_helpers_match_oracle_layout():
    mapping = synth.VOC_CLASS_MAPPING
    assert det_util._one_hot_encode_cls(['bg', 'cat', 'aeroplane'], mapping).tolist()[1][7] == 1
    from faster_rcnn_b200.shapes import Box, GroundTruthBox
    rois = np.array([[1, 2, 9, 12], [0, 0, 5, 5]], np.int16)
    gts = [GroundTruthBox('cat', False, Box(1.5, 2.25, 8.0, 11.0)), None]
    out = det_util._one_hot_encode_bbreg(rois, gts, [True, False], mapping)
    assert out.shape == (2, 160) and out[0, 28:32].tolist() == [1, 1, 1, 1] and not out[1].any()
    want = np.float32(O.regression_params(rois[0], gts[0].corners)) * shared_constants.BBREG_MULTIPLIERS
    assert np.array_equal(out[0, 80 + 28:80 + 32], want)


def test_padding_rule_and_sharding():
    rois = synth.random_rois(300, 37, 62, 1)
    padded = voc_dets.pad_roi_batches(rois, 64)
    assert padded.shape == (320, 4) and np.array_equal(padded[:300], rois)
    assert np.array_equal(padded[300:], np.tile(rois[256], (20, 1)))            # first RoI of the LAST batch
    assert voc_dets.pad_roi_batches(rois[:128], 64) is rois[:128] or len(voc_dets.pad_roi_batches(rois[:128], 64)) == 128
    for n, world in ((64, 8), (65, 8), (3, 8), (128, 2), (0, 4)):
        blocks = [parallel.shard_range(n, r, world) for r in range(world)]
        assert blocks[0][0] == 0 and blocks[-1][1] == n
        assert all(blocks[i][1] == blocks[i + 1][0] for i in range(world - 1))
        assert max(b - a for a, b in blocks) - min(b - a for a, b in blocks) <= 1


def test_roi_layer_surface_without_gpu():
    layer = RoiResizeConv(7, 64)
    layer.build([(None, 38, 63, 1024), (None, 64, 4)])
    assert layer.compute_output_shape(None) == (None, 64, 7, 7, 1024)
    assert layer.get_config() == {'pool_size': 7, 'num_rois': 64}
    assert RoiResizeConv.from_config(layer.get_config()).num_rois == 64
    with pytest.raises(ValueError):
        RoiResizeConv(7, 64, mode="avg")


def test_synthetic_inputs_are_deterministic_and_tie_free():
    a = synth.rpn_outputs(5, 7, 9, 3, clustered=True)
    b = synth.rpn_outputs(5, 7, 9, 3, clustered=True)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]) and len(np.unique(a[0])) == 5 * 7 * 9
    r = synth.random_rois(100, 38, 63, 0)
    assert np.all(r[:, 2] > r[:, 0]) and np.all(r[:, 3] > r[:, 1]) and r[:, 2].max() <= 62 and r[:, 3].max() <= 37
