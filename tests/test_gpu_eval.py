"""GPU parity of the VOC evaluation kernels (widening row) through the C ABI vs golden vectors of the
unmodified reference `eval_dets.voc_eval` and vs the oracle: float64, bit-exact rec / prec / ap."""
import os

import numpy as np
import pytest

from helpers import voc_eval_golden
from oracle import eval_oracle as E

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("cls", ["person", "chair", "car"])
def test_voc_eval_arrays_vs_reference_golden(cls):
    from faster_rcnn_b200 import eval_dets
    ids, conf, boxes, gt, names, rec, prec, ap = voc_eval_golden(cls)
    r, p, a = eval_dets.voc_eval_arrays(ids, conf, boxes, gt, names)
    assert np.array_equal(r, rec) and np.array_equal(p, prec) and a == ap
    assert eval_dets.voc_ap(r, p, use_07_metric=True) == ap and eval_dets.voc_ap(r, p) == E.voc_ap(rec, prec)


def test_voc_eval_edge_cases():
    from faster_rcnn_b200 import eval_dets
    rng = np.random.default_rng(3)
    names = ["img%03d" % i for i in range(70)]
    gt = {}
    for i, n in enumerate(names):
        g = i % 5 * 9                                           # 0, 9, 18, 27, 36 boxes: more than one warp round
        xy = rng.uniform(0, 400, (g, 2))
        gt[n] = (np.concatenate([xy, xy + rng.uniform(10, 120, (g, 2))], axis=1).round(), rng.random(g) < 0.2)
    ids, boxes = [], []
    for n in names:
        for b in gt[n][0]:
            for _ in range(rng.integers(0, 4)):
                ids.append(n)
                boxes.append(b + rng.normal(0, 4, 4))
        ids.append(n)
        boxes.append(np.array([500., 500., 520., 530.]))
    boxes = np.array(boxes)
    conf = (rng.integers(0, 50, len(ids)) / 50.0)               # many ties: the stable order is the definition
    want = E.voc_match(ids, conf, boxes, gt, stable=True)
    got = eval_dets.voc_eval_arrays(ids, conf, boxes, gt, names)
    assert np.array_equal(got[0], want[0]) and np.array_equal(got[1], want[1]) and got[2] == want[2]
    # identical duplicates of one GT: only the first is a true positive
    one = {"a": (np.array([[10., 10., 50., 50.]]), np.array([False]))}
    r, p, a = eval_dets.voc_eval_arrays(["a"] * 3, np.array([0.9, 0.8, 0.7]), np.tile([[10., 10., 50., 50.]], (3, 1)), one, ["a"])
    assert r.tolist() == [1.0, 1.0, 1.0] and p.tolist() == [1.0, 0.5, 1.0 / 3.0] and a == E.voc_ap(r, p, True)   # 11 x (1/11)
    assert eval_dets.voc_eval_arrays([], np.zeros(0), np.zeros((0, 4)), one, ["a"])[2] == 0.0


def test_voc_eval_file_interface(tmp_path):
    """voc_eval / write_dets / eval_all with the reference's file formats on a synthetic VOC directory."""
    from faster_rcnn_b200 import eval_dets
    root = str(tmp_path / "VOC")
    os.makedirs(os.path.join(root, "Annotations"))
    os.makedirs(os.path.join(root, "ImageSets", "Main"))
    rng = np.random.default_rng(5)
    names = ["%06d" % i for i in range(1, 31)]
    mapping = {"cat": 0, "dog": 1, "bg": 2}
    gt = {c: {} for c in ("cat", "dog")}
    dets = {c: {} for c in ("cat", "dog")}
    for n in names:
        objs = []
        for _ in range(rng.integers(0, 4)):
            c = ("cat", "dog")[rng.integers(0, 2)]
            x1, y1 = rng.integers(1, 200, 2)
            w, h = rng.integers(20, 150, 2)
            diff = int(rng.random() < 0.2)
            objs.append((c, x1, y1, x1 + w, y1 + h, diff))
        xml = "<annotation><filename>%s.jpg</filename><size><width>500</width><height>375</height><depth>3</depth></size>%s</annotation>" % (
            n, "".join("<object><name>%s</name><difficult>%d</difficult><bndbox><xmin>%d</xmin><ymin>%d</ymin><xmax>%d</xmax>"
                       "<ymax>%d</ymax></bndbox></object>" % (c, d, a, b, e, f) for c, a, b, e, f, d in objs))
        open(os.path.join(root, "Annotations", n + ".xml"), "w").write(xml)
        for c in ("cat", "dog"):
            mine = [o for o in objs if o[0] == c]
            gt[c][n] = (np.array([[o[1] - 1, o[2] - 1, o[3] - 1, o[4] - 1] for o in mine], float).reshape(-1, 4),
                        np.array([bool(o[5]) for o in mine]))
            dets[c][n] = [{'bbox': np.array(o[1:5]) + rng.integers(-6, 7, 4), 'prob': np.float32(rng.random())} for o in mine] + \
                         [{'bbox': np.array([300, 300, 340, 350]), 'prob': np.float32(rng.random())}]
    open(os.path.join(root, "ImageSets", "Main", "val.txt"), "w").write("\n".join(names) + "\n")
    out_dir = str(tmp_path / "dets")
    eval_dets.write_dets(dets, out_dir)
    first = open(eval_dets.get_voc_results_filename(out_dir, "cat")).readline().split(" ")
    assert first[0] == names[0] and len(first) == 6
    aps, mean_ap = eval_dets.eval_all(out_dir, root, mapping)
    for c in ("cat", "dog"):
        ids = [n for n in names for _ in dets[c][n]]
        conf = np.array([float(str(d['prob'])) for n in names for d in dets[c][n]])          # what write_dets printed
        boxes = np.array([(d['bbox'] + 1).astype(float) for n in names for d in dets[c][n]])
        want = E.voc_match(ids, conf, boxes, gt[c])
        assert aps[c] == want[2]
    assert abs(mean_ap - np.mean([aps["cat"], aps["dog"]])) < 1e-15 and 0 < mean_ap <= 1
