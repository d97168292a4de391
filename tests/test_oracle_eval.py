"""VOC evaluation oracle vs golden vectors produced by the unmodified reference `eval_dets.voc_eval` on its own
VOC_test annotations (tests/golden/make_golden.py: voc_eval_case).  CPU only."""
import numpy as np
import pytest

from helpers import voc_eval_golden
from oracle import eval_oracle as E


@pytest.mark.parametrize("cls", ["person", "chair", "car"])
def test_voc_match_vs_reference_golden(cls):
    ids, conf, boxes, gt, names, rec, prec, ap = voc_eval_golden(cls)
    r, p, a = E.voc_match(ids, conf, boxes, gt)
    assert np.array_equal(r, rec) and np.array_equal(p, prec) and a == ap
    assert 0.05 < ap < 0.9 and len(rec) > 300


def test_voc_ap_both_metrics():
    rec = np.array([0.1, 0.2, 0.2, 0.4, 0.4, 0.5])
    prec = np.array([1.0, 1.0, 0.67, 0.75, 0.6, 0.5])
    assert abs(E.voc_ap(rec, prec, True) - (1 + 1 + 1 + 0.75 + 0.75 + 0.5) / 11) < 1e-12
    assert abs(E.voc_ap(rec, prec, False) - (0.2 * 1.0 + 0.2 * 0.75 + 0.1 * 0.5)) < 1e-12
