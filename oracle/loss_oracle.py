"""CPU oracle for the masked RPN / detector losses -- TEST INFRASTRUCTURE ONLY.

PARITY UNPINNED.  The reference's losses (loss_functions.py:15-76) are Keras-backend expressions evaluated
by tensorflow==1.3.0 through keras==2.0.8, neither of which is available here, and the reference has no
test for them.  This file restates the published formulas of that stack in numpy float32:

  K.binary_crossentropy(target, output)     keras/backend/tensorflow_backend.py (2.0.8):
        output = clip(output, 1e-7, 1 - 1e-7); x = log(output / (1 - output))
        tf.nn.sigmoid_cross_entropy_with_logits(labels=z, logits=x) = max(x, 0) - x*z + log1p(exp(-|x|))
  categorical_crossentropy(target, output):
        output /= sum(output, axis=-1); output = clip(output, 1e-7, 1 - 1e-7); -sum(target * log(output), axis=-1)
  Keras reports K.mean(loss_tensor) for each output.

Quirk kept on purpose (loss_functions.py:40-46): the RPN box loss multiplies the mask OUTSIDE the sum, so the
loss tensor is `mask * (10 * S / 2400)` with S the smooth-L1 sum over ALL anchors, and Keras reports its mean.

Inputs are the hot path's own targets: y_class (R,C,2A) = [can_use | is_pos], y_bbreg (R,C,8A) =
[repeat(is_pos & can_use, 4) | targets] (rpn_util.py:125-140), y_class_num (M,K) one-hot, y_transform
(M, 8(K-1)) = [labels | targets] (det_util.py:338-366).
"""
import numpy as np

F = np.float32
EPS = F(1e-7)
N_CLS, N_REG, LAMBDA_REG, LAMBDA_REG_DET = F(256), F(2400), F(10.0), F(1.0)     # loss_functions.py:8-12


def _bce(z, p):
    p = np.clip(p.astype(F), EPS, F(1) - EPS)
    x = np.log(p / (F(1) - p)).astype(F)
    return (np.maximum(x, F(0)) - x * z + np.log1p(np.exp(-np.abs(x)))).astype(F)


def _smooth_l1(d):
    a = np.abs(d).astype(F)
    small = (a <= F(1.0)).astype(F)
    return (small * (F(0.5) * a * a) + (F(1) - small) * (a - F(0.5))).astype(F)


def rpn_cls_loss(y_class, cls_pred, n_anchors):
    """loss_functions.py:15-29.  y_class (R,C,2A) bool/float, cls_pred (R,C,A) f32 -> scalar f32."""
    y = y_class.astype(F)
    sel, z = y[..., :n_anchors], y[..., n_anchors:]
    return F(np.sum(sel * _bce(z, cls_pred), dtype=np.float64)) / N_CLS


def rpn_bbreg_loss(y_bbreg, reg_pred, n_anchors):
    """loss_functions.py:32-48 as Keras reports it (mean of the loss tensor).  -> scalar f32."""
    sel = y_bbreg[..., :4 * n_anchors].astype(F)
    s = F(np.sum(_smooth_l1(y_bbreg[..., 4 * n_anchors:].astype(F) - reg_pred.astype(F)), dtype=np.float64))
    return F(np.mean(sel, dtype=np.float64)) * (LAMBDA_REG * s / N_REG)


def det_bbreg_loss(y_transform, reg_pred, n_fg):
    """loss_functions.py:51-66.  y_transform (M,8K') f32, reg_pred (M,4K') f32 -> scalar f32."""
    mask = y_transform[..., :4 * n_fg].astype(F)
    x = y_transform[..., 4 * n_fg:].astype(F) - reg_pred.astype(F)
    num = F(np.sum(mask * _smooth_l1(x), dtype=np.float64))
    den = F(np.sum(F(1e-4) + mask, dtype=np.float64))
    return LAMBDA_REG_DET * num / den


def det_cls_loss(y_class_num, cls_pred):
    """loss_functions.py:69-76.  y_class_num (M,K) one-hot, cls_pred (M,K) f32 -> scalar f32."""
    p = cls_pred.astype(F)
    p = p / np.sum(p, axis=-1, keepdims=True, dtype=F)
    p = np.clip(p, EPS, F(1) - EPS)
    rows = -np.sum(y_class_num.astype(F) * np.log(p), axis=-1, dtype=np.float64)
    return F(np.mean(rows))
