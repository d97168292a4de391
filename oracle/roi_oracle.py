"""CPU oracle for the RoI layer (`RoiResizeConv`) -- TEST INFRASTRUCTURE ONLY.

Resize mode -- pinned through an independent executable implementation.  The reference layer
(custom_layers.py:35-56) crops ``img[:, y1:y2, x1:x2, :]`` per RoI and calls
``tf.image.resize_images(crop, (P, P))`` of tensorflow==1.3.0 (requirements.txt:53), i.e. the legacy bilinear
kernel with ``align_corners=False`` and no half-pixel offset.  TensorFlow is a third-party dependency that is
neither vendored under /root/reference nor installable here, and the reference's only tests for this layer are
golden ``.h5`` files that are missing (.MISSING_LARGE_BLOBS).  This file restates the *published* TF-1.x
algorithm (core/kernels/resize_bilinear_op.cc):

    scale = in_size / float(out_size)            (float32)
    src   = out_index * scale                    (float32)
    lo    = (int) src ;  hi = min(lo + 1, in_size - 1) ;  lerp = src - lo
    top    = tl + (tr - tl) * x_lerp
    bottom = bl + (br - bl) * x_lerp
    out    = top + (bottom - top) * y_lerp

and ResizeBilinearGrad (taps scattered in the order TL, TR, BL, BR with weights
(1-ly)(1-lx), (1-ly)lx, ly(1-lx), ly*lx), followed by the zero-padded slice
gradient and an AddN over RoIs in index order.

THE PIN (tests/test_oracle_roi.py, fixtures tests/golden/resize_bilinear_cv2dnn.npz made by
tests/golden/make_golden_resize.py): OpenCV 4.13's cv2.dnn runs a real TensorFlow GraphDef holding the
``ResizeBilinear`` op (align_corners=false) -- an implementation of exactly this op written by other people.
`_axis_taps` (coordinates, border clamp, lerp weights) put through OpenCV's tap expression
(`resize_bilinear_opencv_form`) reproduces cv2.dnn BIT FOR BIT on every fixture and on live random crops, so the
sampling geometry is pinned exactly.  What remains restated is the association of the final four-tap
combination (TF: top/bottom lerps as above; OpenCV: a + ly(b-a) + lx((c-a) + ly(d-c-b+a))), which moves results
by float32 rounding only: `roi_resize_fwd` differs from cv2.dnn by <= 4 ulp-of-the-largest-tap, counted in the
test.  The backward is the exact adjoint of the forward (tests/test_properties.py), so the same taps pin it up
to summation order.  A second, tolerance-only cross-check is `torch.nn.functional.grid_sample` on explicit
legacy-coordinate grids.

`mode="max"` (roi_max_*) is the north-star max-pool variant; the reference has
no such layer.  Its spec is the Fast R-CNN RoIPool:  bin (ph,pw) of an h x w crop
covers rows  y1 + floor(ph*h/P) .. y1 + ceil((ph+1)*h/P) - 1  (cols likewise);
output is the max over the bin, argmax the flat ``y*W + x`` of the FIRST maximum
in row-major scan; backward adds dY to dX[argmax].  PINNED against an independent
implementation: outputs equal ``torchvision.ops.roi_pool`` bit for bit (exclusive
x2/y2 passed as inclusive ends, spatial_scale 1) and the backward equals its
autograd on tie-rich inputs (tests/test_oracle_roi.py).

Layouts: feat (H,W,C) f32 channels-last (batch index 0 of the Keras tensor),
rois (N,4) integer [x1,y1,x2,y2] in feature cells with x2/y2 EXCLUDED from the
crop, out (N,P,P,C).
"""
import numpy as np


def _axis_taps(in_size, out_size):
    """lo, hi, lerp per output index (float32 arithmetic as in TF)."""
    scale = np.float32(in_size) / np.float32(out_size)
    src = np.arange(out_size, dtype=np.float32) * scale
    lo = src.astype(np.int64)
    hi = np.minimum(lo + 1, in_size - 1)
    lerp = (src - lo.astype(np.float32)).astype(np.float32)
    return lo, hi, lerp


def resize_bilinear_opencv_form(crop, pool):
    """The taps of `_axis_taps` combined with OpenCV's expression (modules/dnn/src/layers/resize_layer.cpp, bilinear):
    out = a + ly*(b - a) + lx*((c - a) + ly*(((d - c) - b) + a)), a/b = rows lo/hi at column lo, c/d at column hi.
    Equals cv2.dnn's TF-ResizeBilinear bit for bit -- this is how the taps are pinned (see the header)."""
    crop = np.asarray(crop, dtype=np.float32)
    h, w = crop.shape[:2]
    ylo, yhi, ly = _axis_taps(h, pool)
    xlo, xhi, lx = _axis_taps(w, pool)
    a, b = crop[ylo][:, xlo], crop[yhi][:, xlo]
    c, d = crop[ylo][:, xhi], crop[yhi][:, xhi]
    lyb, lxb = ly[:, None, None], lx[None, :, None]
    return (a + lyb * (b - a)) + lxb * ((c - a) + lyb * (((d - c) - b) + a))


def roi_resize_fwd(feat, rois, pool):
    """custom_layers.py:41-54 with TF-1.3 legacy bilinear (see header)."""
    feat = np.asarray(feat, dtype=np.float32)
    n, c = len(rois), feat.shape[2]
    out = np.zeros((n, pool, pool, c), dtype=np.float32)
    for r in range(n):
        x1, y1, x2, y2 = (int(v) for v in rois[r])       # K.cast(.., 'int32'): truncation
        crop = feat[y1:y2, x1:x2, :]
        h, w = crop.shape[:2]
        if h <= 0 or w <= 0:
            raise ValueError("empty crop for roi %d: %r" % (r, rois[r]))
        ylo, yhi, ly = _axis_taps(h, pool)
        xlo, xhi, lx = _axis_taps(w, pool)
        lxb = lx[None, :, None]
        tl, tr = crop[ylo][:, xlo], crop[ylo][:, xhi]
        bl, br = crop[yhi][:, xlo], crop[yhi][:, xhi]
        top = tl + (tr - tl) * lxb
        bot = bl + (br - bl) * lxb
        out[r] = top + (bot - top) * ly[:, None, None]
    return out


def roi_resize_bwd(grad_out, rois, feat_shape):
    """dX (H,W,C) f32.  Accumulation order: RoIs ascending (AddN), inside a RoI
    (ph, pw) ascending and taps TL, TR, BL, BR (ResizeBilinearGrad), each RoI
    first summed into its own zero crop then added to dX (slice gradient)."""
    hh, ww, c = feat_shape
    n, pool = grad_out.shape[0], grad_out.shape[1]
    dx = np.zeros((hh, ww, c), dtype=np.float32)
    one = np.float32(1.0)
    for r in range(n):
        x1, y1, x2, y2 = (int(v) for v in rois[r])
        y2c, x2c = min(y2, hh), min(x2, ww)
        h, w = y2c - y1, x2c - x1
        ylo, yhi, ly = _axis_taps(h, pool)
        xlo, xhi, lx = _axis_taps(w, pool)
        local = np.zeros((h, w, c), dtype=np.float32)
        for ph in range(pool):
            for pw in range(pool):
                g = grad_out[r, ph, pw]
                local[ylo[ph], xlo[pw]] += g * (one - ly[ph]) * (one - lx[pw])
                local[ylo[ph], xhi[pw]] += g * (one - ly[ph]) * lx[pw]
                local[yhi[ph], xlo[pw]] += g * ly[ph] * (one - lx[pw])
                local[yhi[ph], xhi[pw]] += g * ly[ph] * lx[pw]
        dx[y1:y2c, x1:x2c] += local
    return dx


def _bin_edges(size, pool):
    lo = [(p * size) // pool for p in range(pool)]
    hi = [((p + 1) * size + pool - 1) // pool for p in range(pool)]     # exclusive
    return lo, hi


def roi_max_fwd(feat, rois, pool):
    """Max-pool variant (spec in the header).  Returns out (N,P,P,C) f32 and
    argmax (N,P,P,C) int32 (flat y*W+x into the feature map)."""
    feat = np.asarray(feat, dtype=np.float32)
    hh, ww, c = feat.shape
    n = len(rois)
    out = np.zeros((n, pool, pool, c), dtype=np.float32)
    arg = np.zeros((n, pool, pool, c), dtype=np.int32)
    for r in range(n):
        x1, y1, x2, y2 = (int(v) for v in rois[r])
        h, w = y2 - y1, x2 - x1
        if h <= 0 or w <= 0:
            raise ValueError("empty crop for roi %d: %r" % (r, rois[r]))
        ylo, yhi = _bin_edges(h, pool)
        xlo, xhi = _bin_edges(w, pool)
        for ph in range(pool):
            for pw in range(pool):
                ys = np.arange(y1 + ylo[ph], y1 + yhi[ph])
                xs = np.arange(x1 + xlo[pw], x1 + xhi[pw])
                cells = feat[ys][:, xs].reshape(-1, c)                  # row-major scan
                flat = (ys[:, None] * ww + xs[None, :]).reshape(-1)
                best = cells[0].copy()
                best_i = np.full(c, flat[0], dtype=np.int32)
                for k in range(1, len(flat)):
                    upd = cells[k] > best                               # strict: first max wins
                    best = np.where(upd, cells[k], best)
                    best_i = np.where(upd, flat[k], best_i).astype(np.int32)
                out[r, ph, pw], arg[r, ph, pw] = best, best_i
    return out, arg


def roi_max_bwd(grad_out, argmax, feat_shape):
    """dX[argmax] += dY, RoIs then bins in ascending order (float32 adds)."""
    hh, ww, c = feat_shape
    dx = np.zeros((hh * ww, c), dtype=np.float32)
    n, pool = grad_out.shape[0], grad_out.shape[1]
    ch = np.arange(c)
    for r in range(n):
        for ph in range(pool):
            for pw in range(pool):
                dx[argmax[r, ph, pw], ch] += grad_out[r, ph, pw]
    return dx.reshape(hh, ww, c)
