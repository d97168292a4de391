"""GPU parity of the masked RPN / detector loss kernels (widening row) through the C ABI vs
oracle/loss_oracle.py (float32, tolerance 2e-6 relative: reductions run in a different but fixed order), and of
their gradients vs torch autograd of an independent torch restatement of the Keras formulas."""
import numpy as np
import pytest

from helpers import dev, golden, host
from oracle import frcnn_oracle as O
from oracle import loss_oracle as L

pytestmark = pytest.mark.gpu
RTOL = 2e-6


def _rpn_case(seed):
    from faster_rcnn_b200 import synth
    dims = O.anchor_table([128, 256, 512])
    gt = np.array([g[1:] for g in synth.gt_boxes(12, 1000, 600, seed)], np.float32)
    cu, ip, bb = O.label_anchors(1000, 600, gt, 38, 63, dims, 16)
    rng = np.random.default_rng(seed)
    cls_pred = rng.uniform(0, 1, 38 * 63 * 9).astype(np.float32)
    cls_pred[:50] = [0.0, 1.0, 1e-9, 1 - 1e-9, 0.5] * 10                      # clipping range
    reg_pred = (rng.standard_normal((38 * 63 * 9, 4)) * 1.5).astype(np.float32)
    return cu, ip, bb, cls_pred, reg_pred


@pytest.mark.parametrize("seed", [1, 2])
def test_rpn_losses_vs_oracle(seed):
    from faster_rcnn_b200 import loss_functions, ops
    cu, ip, bb, cls_pred, reg_pred = _rpn_case(seed)
    cu2 = O.sample_rpn(ip, cu.copy())
    y_class, y_bbreg = O.pack_rpn_targets(cu2, ip, bb, 38, 63, 9)
    want_cls = L.rpn_cls_loss(y_class[0], cls_pred.reshape(38, 63, 9), 9)
    want_reg = L.rpn_bbreg_loss(y_bbreg[0], reg_pred.reshape(38, 63, 36), 9)
    loss = host(ops.rpn_losses(dev(cu2.view(np.uint8)[None]), dev(ip.view(np.uint8)[None]), dev(bb[None]),
                               dev(cls_pred[None]), dev(reg_pred[None])))
    assert abs(loss[0, 0] - want_cls) <= RTOL * abs(want_cls) and abs(loss[0, 1] - want_reg) <= RTOL * abs(want_reg)
    assert want_cls > 0 and want_reg > 0
    # the reference's factories on the Keras layouts
    got_cls = loss_functions.cls_loss_rpn(9)(y_class, cls_pred.reshape(1, 38, 63, 9))
    got_reg = loss_functions.bbreg_loss_rpn(9)(y_bbreg, reg_pred.reshape(1, 38, 63, 36))
    assert abs(got_cls - want_cls) <= RTOL * abs(want_cls) and abs(got_reg - want_reg) <= RTOL * abs(want_reg)


def test_det_losses_vs_oracle():
    from faster_rcnn_b200 import loss_functions, ops, synth
    g = golden("det_labels")
    sel = g["sampled"]
    y_cls, y_tr = g["y_class_num"][sel], g["y_transform"][sel]
    cls_pred, reg_pred = synth.detector_outputs(64, 21, 9)
    cls_pred[0] = 0.0
    cls_pred[0, 3] = 1.0                                                      # exercises the clip of the normalised output
    want_cls, want_reg = L.det_cls_loss(y_cls, cls_pred), L.det_bbreg_loss(y_tr, reg_pred, 20)
    loss = host(ops.det_losses(dev(y_cls[None]), dev(y_tr[None]), dev(cls_pred[None]), dev(reg_pred[None])))
    assert abs(loss[0, 0] - want_cls) <= RTOL * abs(want_cls) and abs(loss[0, 1] - want_reg) <= RTOL * abs(want_reg)
    assert abs(loss_functions.cls_loss_det(y_cls[None], cls_pred[None]) - want_cls) <= RTOL * abs(want_cls)
    assert abs(loss_functions.bbreg_loss_det(20)(y_tr[None], reg_pred[None]) - want_reg) <= RTOL * abs(want_reg)
    # batch of images = independent problems
    both = host(ops.det_losses(dev(np.stack([y_cls, y_cls[::-1]])), dev(np.stack([y_tr, y_tr[::-1]])),
                               dev(np.stack([cls_pred, cls_pred[::-1]])), dev(np.stack([reg_pred, reg_pred[::-1]]))))
    assert np.allclose(both[0], loss[0], rtol=1e-6) and np.allclose(both[1], loss[0], rtol=1e-5)


def test_loss_gradients_vs_torch_autograd():
    import torch
    from faster_rcnn_b200 import loss_functions, synth
    cu, ip, bb, cls_pred, reg_pred = _rpn_case(3)
    cls_pred = np.clip(cls_pred, 1e-4, 1 - 1e-4)
    cu_t, ip_t, bb_t = dev(cu.view(np.uint8)[None]), dev(ip.view(np.uint8)[None]), dev(bb[None])
    p = dev(cls_pred[None]).requires_grad_(True)
    q = dev(reg_pred[None]).requires_grad_(True)
    loss = loss_functions.rpn_losses(p, q, cu_t, ip_t, bb_t)
    (loss[0, 0] * 2.0 + loss[0, 1] * 3.0).backward()
    # independent torch restatement (double precision) of loss_functions.py:15-48
    pd = dev(cls_pred[None]).double().requires_grad_(True)
    qd = dev(reg_pred[None]).double().requires_grad_(True)
    sel, z = cu_t.double(), ip_t.double()
    pc = pd.clamp(1e-7, 1 - 1e-7)
    x = torch.log(pc / (1 - pc))
    bce = torch.clamp(x, min=0) - x * z + torch.log1p(torch.exp(-x.abs()))
    l_cls = (sel * bce).sum() / 256
    d = (bb_t.double() - qd).abs()
    s = torch.where(d <= 1, 0.5 * d * d, d - 0.5).sum()
    mask4 = (sel * z).unsqueeze(-1).expand(-1, -1, 4)
    l_reg = (mask4 * (10 * s / 2400)).mean()
    (l_cls * 2.0 + l_reg * 3.0).backward()
    assert torch.allclose(p.grad.double(), pd.grad, rtol=1e-4, atol=1e-9)
    assert torch.allclose(q.grad.double(), qd.grad, rtol=1e-4, atol=1e-12)
    assert abs(loss[0, 0].item() - l_cls.item()) < 1e-5 * l_cls.item() and abs(loss[0, 1].item() - l_reg.item()) < 1e-5 * l_reg.item()

    g = golden("det_labels")
    y_cls, y_tr = g["y_class_num"][g["sampled"]], g["y_transform"][g["sampled"]]
    cp, rp = synth.detector_outputs(64, 21, 11)
    a = dev(cp[None]).requires_grad_(True)
    b = dev(rp[None]).requires_grad_(True)
    dl = loss_functions.det_losses(a, b, dev(y_cls[None]), dev(y_tr[None]))
    (dl[0, 0] + dl[0, 1] * 0.5).backward()
    ad = dev(cp[None]).double().requires_grad_(True)
    bd = dev(rp[None]).double().requires_grad_(True)
    yn = (ad / ad.sum(dim=-1, keepdim=True)).clamp(1e-7, 1 - 1e-7)
    l_c = (-(dev(y_cls[None]).double() * torch.log(yn)).sum(dim=-1)).mean()
    mask, tgt = dev(y_tr[None])[..., :80].double(), dev(y_tr[None])[..., 80:].double()
    xx = tgt - bd
    l_b = (mask * torch.where(xx.abs() <= 1, 0.5 * xx * xx, xx.abs() - 0.5)).sum() / (1e-4 + mask).sum()
    (l_c + l_b * 0.5).backward()
    assert torch.allclose(a.grad.double(), ad.grad, rtol=1e-4, atol=1e-9)
    assert torch.allclose(b.grad.double(), bd.grad, rtol=1e-4, atol=1e-12)
