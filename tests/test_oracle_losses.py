"""Cross-checks the loss oracle (parity unpinned: Keras 2.0.8 / TF 1.3 are unavailable) against an independent
torch restatement of the published formulas, and pins the reference's masking quirk.  CPU only."""
import numpy as np
import torch

from oracle import frcnn_oracle as O
from oracle import loss_oracle as L


def test_rpn_losses_match_torch_restatement():
    rng = np.random.default_rng(0)
    r, c, a = 6, 7, 9
    cu = rng.random((r, c, a)) < 0.4
    ip = rng.random((r, c, a)) < 0.1
    bb = rng.standard_normal((r * c * a, 4)).astype(np.float32)
    y_class, y_bbreg = O.pack_rpn_targets(cu.reshape(-1), ip.reshape(-1), bb, r, c, a)
    p = rng.uniform(0, 1, (r, c, a)).astype(np.float32)
    q = rng.standard_normal((r, c, 4 * a)).astype(np.float32) * 2
    z, pt = torch.from_numpy(ip.astype(np.float64)), torch.from_numpy(p).double().clamp(1e-7, 1 - 1e-7)
    want_cls = (torch.from_numpy(cu.astype(np.float64)) * torch.nn.functional.binary_cross_entropy(pt, z, reduction='none')).sum() / 256
    assert abs(L.rpn_cls_loss(y_class[0], p, a) - want_cls.item()) < 1e-5 * want_cls.item()
    d = np.abs(y_bbreg[0][..., 4 * a:].astype(np.float64) - q)
    s = np.where(d <= 1, 0.5 * d * d, d - 0.5).sum()
    want_reg = y_bbreg[0][..., :4 * a].mean() * 10 * s / 2400           # the mask multiplies the SUM (loss_functions.py:44)
    assert abs(L.rpn_bbreg_loss(y_bbreg[0], q, a) - want_reg) < 1e-5 * want_reg


def test_det_losses_match_torch_restatement():
    rng = np.random.default_rng(1)
    m, k = 64, 21
    cls = rng.integers(0, k, m)
    y_cls = np.eye(k, dtype=np.int32)[cls]
    y_tr = np.zeros((m, 160), np.float32)
    for i, c in enumerate(cls):
        if c < 20:
            y_tr[i, 4 * c:4 * c + 4] = 1
            y_tr[i, 80 + 4 * c:84 + 4 * c] = rng.standard_normal(4)
    logits = rng.standard_normal((m, k))
    p = (np.exp(logits) / np.exp(logits).sum(1, keepdims=True)).astype(np.float32)
    q = rng.standard_normal((m, 80)).astype(np.float32)
    want_cls = torch.nn.functional.nll_loss(torch.log(torch.from_numpy(p).double()), torch.from_numpy(cls)).item()
    assert abs(L.det_cls_loss(y_cls, p) - want_cls) < 1e-5 * want_cls
    x = torch.from_numpy(y_tr[:, 80:]).double() - torch.from_numpy(q).double()
    sl1 = torch.nn.functional.smooth_l1_loss(torch.from_numpy(q).double(), torch.from_numpy(y_tr[:, 80:]).double(), reduction='none')
    mask = torch.from_numpy(y_tr[:, :80]).double()
    want_reg = ((mask * sl1).sum() / (1e-4 + mask).sum()).item()
    assert x.shape == sl1.shape and abs(L.det_bbreg_loss(y_tr, q, 20) - want_reg) < 1e-5 * want_reg
