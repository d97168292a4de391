"""CPU oracle for the host-side input pipeline -- TEST INFRASTRUCTURE ONLY.

Reference: `shapes.Image.data` (shapes.py:19-29) = cv2.imread -> cv2.resize(INTER_CUBIC) -> optional cv2.flip(img, 1);
`resnet.preprocess` / `vgg.preprocess` (resnet.py:64-75, vgg.py:52-57) = BGR->RGB, float64, Keras 2.0.8
`preprocess_input` (RGB->BGR again, minus the ImageNet means [103.939, 116.779, 123.68]); GT boxes are scaled with
`Box.resize` and mirrored with `Box.horizontal_flip` (shapes.py:93-132, 292-300, 400-408).

`resize_cubic_u8` restates OpenCV's generic (non-SIMD) uint8 bicubic path (modules/imgproc/src/resize.cpp):
    scale = 1 / (dst / src)  (double);  f = (float)((d + 0.5) * scale - 0.5);  s = floor(f);  f -= s
    coefficients of interpolateCubic with A = -0.75 in float32, each stored as cvRound(c * 2048) in a short
    horizontal pass in int32 over taps s-1 .. s+2 (indices clamped to the image), vertical pass likewise,
    result = (sum + 2^21) >> 22, saturated to uint8.
PINNED against the installed cv2 (tests/test_oracle_image.py): with `cv2.setUseOptimized(False)` cv2.resize takes that
generic path and agrees on all but <= 2e-4 of the pixels (never by more than one grey level); the SIMD-dispatched
default path of the same binary differs from its own generic path -- and therefore from this restatement -- by one
grey level on 0.4 % of the pixels of a natural image (<= 6 % on white noise).  The reference's pixels are thus defined
by the OpenCV build only up to +-1 level; the tolerance of the parity tests is exactly that.
"""
import numpy as np

_F = np.float32
IMAGENET_MEAN_BGR = (103.939, 116.779, 123.68)


def _cubic_coeffs(x):
    a = _F(-0.75)
    x = x.astype(np.float32)
    one = _F(1)
    c0 = ((a * (x + one) - _F(5) * a) * (x + one) + _F(8) * a) * (x + one) - _F(4) * a
    c1 = ((a + _F(2)) * x - (a + _F(3))) * x * x + one
    c2 = ((a + _F(2)) * (one - x) - (a + _F(3))) * (one - x) * (one - x) + one
    c3 = one - c0 - c1 - c2
    return np.stack([c0, c1, c2, c3], axis=1).astype(np.float32)


def cubic_axis_tables(src_size, dst_size):
    """(dst,4) clamped source indices and (dst,4) fixed-point coefficients (x2048) of one axis."""
    scale = np.float64(1.0) / (np.float64(dst_size) / np.float64(src_size))
    d = np.arange(dst_size, dtype=np.float64)
    f = ((d + 0.5) * scale - 0.5).astype(np.float32)
    s = np.floor(f).astype(np.int64)
    f = (f - s.astype(np.float32)).astype(np.float32)
    coef = np.clip(np.rint(_cubic_coeffs(f) * _F(2048)).astype(np.int64), -32768, 32767)     # cvRound: half to even
    idx = np.clip(s[:, None] - 1 + np.arange(4)[None, :], 0, src_size - 1)
    return idx, coef


def resize_cubic_u8(img, dst_w, dst_h, flip=False):
    """img (H,W,C) uint8 -> (dst_h,dst_w,C) uint8, cv2.resize(..., INTER_CUBIC) [+ cv2.flip(.., 1)]."""
    img = np.asarray(img, np.uint8)
    xi, xa = cubic_axis_tables(img.shape[1], dst_w)
    yi, yb = cubic_axis_tables(img.shape[0], dst_h)
    src = img.astype(np.int64)
    hor = np.zeros((img.shape[0], dst_w, img.shape[2]), np.int64)
    for k in range(4):
        hor += src[:, xi[:, k], :] * xa[None, :, k, None]
    ver = np.zeros((dst_h, dst_w, img.shape[2]), np.int64)
    for k in range(4):
        ver += hor[yi[:, k]] * yb[:, k, None, None]
    out = np.clip((ver + (1 << 21)) >> 22, 0, 255).astype(np.uint8)
    return out[:, ::-1].copy() if flip else out


def preprocess_bgr(img_u8, mean_bgr=IMAGENET_MEAN_BGR):
    """resnet.py:64-75: the BGR pixels minus the per-channel means, computed in float64 like the reference; the device
    stores float32 (the backbone's input type), so the comparison value is this array cast to float32."""
    return np.asarray(img_u8, np.float64) - np.asarray(mean_bgr, np.float64)[None, None, :]


def transform_gt(boxes, ratio, flip_width=None):
    """shapes.py:93-101 + 292-300: corners * ratio (Python floats = float64), then, for a mirrored image of width
    `flip_width` (the RESIZED width), x1' = W - x2, x2' = W - x1."""
    b = np.asarray(boxes, np.float64) * np.float64(ratio)
    if flip_width is not None:
        w = np.float64(flip_width)
        b = np.stack([w - b[:, 2], b[:, 1], w - b[:, 0], b[:, 3]], axis=1)
    return b
