"""GPU parity of the device-side input pipeline through the C ABI: bicubic resize + mirror + mean
subtraction (shapes.py:19-29, resnet.py:64-75) and GT-box scale / mirror (shapes.py:93-132, 292-300).
Bit-exact against oracle/image_oracle.py; one grey level against the installed cv2 (its SIMD and generic code paths
differ from each other by as much)."""
import os

import numpy as np
import pytest

from helpers import dev, host
from oracle import image_oracle as IO
from test_oracle_image import _natural, _write_voc

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops():
    from faster_rcnn_b200 import ops as _ops
    return _ops


@pytest.mark.parametrize("src,dst,cn,batch,flip", [
    ((375, 500), (600, 800), 3, 1, False),      # VOC 000005 -> resize_within_bounds(600, 1000)
    ((375, 500), (600, 800), 3, 2, True),
    ((480, 640), (300, 400), 3, 1, False),      # downscale (no antialiasing in INTER_CUBIC)
    ((37, 53), (111, 97), 1, 3, True),
    ((5, 4), (9, 13), 4, 1, False),             # taps clamped on every side
    ((600, 800), (600, 800), 3, 1, False),      # identity
    ((640, 480), (90, 70), 3, 2, True),         # 7x reduction: eight output rows span > 40 source rows, per-pixel kernel
    ((333, 47), (61, 301), 2, 1, False),        # reduced 5.5x in y (tile limit), enlarged 6.4x in x, ragged tiles
])
def test_resize_kernel_equals_oracle(ops, src, dst, cn, batch, flip):
    rng = np.random.default_rng(src[0] + dst[1] + cn)
    imgs = np.stack([_natural(src[0], src[1], 20 + i)[..., :1].repeat(cn, 2) if cn != 3 else _natural(src[0], src[1], 20 + i)
                     for i in range(batch)])
    imgs[..., -1] = rng.integers(0, 256, imgs.shape[:-1], dtype=np.uint8)            # one white-noise channel
    mean = [103.939, 116.779, 123.68, 7.25][:cn]
    u8, f32 = ops.image_resize_cubic(dev(imgs), dst[0], dst[1], flip=flip, mean=mean)
    u8, f32 = host(u8), host(f32)
    for b in range(batch):
        want = IO.resize_cubic_u8(imgs[b], dst[1], dst[0], flip=flip)
        assert np.array_equal(u8[b], want)
        assert np.array_equal(f32[b], IO.preprocess_bgr(want, mean).astype(np.float32))
    only_f = ops.image_resize_cubic(dev(imgs), dst[0], dst[1], flip=flip, mean=mean, want_u8=False)
    assert np.array_equal(host(only_f), f32)


def test_resize_kernel_vs_installed_cv2(ops):
    cv2 = pytest.importorskip("cv2")
    img = _natural(375, 500, 5)
    got = host(ops.image_resize_cubic(dev(img[None]), 600, 800))[0].astype(int)
    want = cv2.resize(img, (800, 600), interpolation=cv2.INTER_CUBIC).astype(int)
    assert np.abs(got - want).max() <= 1                              # tolerance: one grey level ...
    assert np.mean(got != want) <= 0.08                               # ... on at most 8 % of the pixels (observed 0.4 - 6 %)


def test_gt_transform_equals_oracle(ops):
    rng = np.random.default_rng(2)
    b, g = 5, 50
    boxes = np.sort(rng.uniform(0, 500, (b, g, 2, 2)), axis=2).transpose(0, 1, 3, 2).reshape(b, g, 4).copy()
    boxes = boxes[..., [0, 2, 1, 3]]                                   # x1, y1, x2, y2
    n_box = np.array([50, 3, 0, 17, 50], np.int32)
    ratio = np.array([1.6, 0.75, 1.0, 1.2345678, 2.0])
    width = np.array([800.0, -1.0, 500.0, -1.0, 1000.0])
    got = host(ops.gt_transform(dev(boxes), dev(ratio), dev(width), dev(n_box)))
    for i in range(b):
        want = IO.transform_gt(boxes[i], ratio[i], width[i] if width[i] >= 0 else None)
        assert np.array_equal(got[i, :n_box[i]], want[:n_box[i]]) and not got[i, n_box[i]:].any()
    assert np.array_equal(host(ops.gt_transform(dev(boxes), dev(ratio))), boxes * ratio[:, None, None])


def test_image_object_device_pixels(tmp_path):
    """shapes.Image from a VOC-layout directory: .data (host, the reference's cv2 calls) vs .data_device() /
    .preprocessed_device() (GPU), for the plain and the mirrored copy."""
    cv2 = pytest.importorskip("cv2")
    from faster_rcnn_b200 import args_util
    root = str(tmp_path)
    _write_voc(root, "a1", 500, 375, [("cat", 0, 10, 20, 300, 200)])
    open(os.path.join(root, "ImageSets", "Main", "trainval.txt"), "w").write("a1\n")
    for img in args_util.base_paths_to_imgs(root, "trainval"):
        big, ratio = img.resize_within_bounds(600, 1000)
        raw = cv2.imread(big.image_path)
        want = IO.resize_cubic_u8(raw, big.width, big.height, flip=big.flipped)
        got = host(big.data_device())
        assert (big.width, big.height, ratio) == (800, 600, 1.6) and np.array_equal(got, want)
        assert np.abs(got.astype(int) - big.data.astype(int)).max() <= 1
        pre = host(big.preprocessed_device())
        assert pre.shape == (1, 600, 800, 3) and np.array_equal(pre[0], IO.preprocess_bgr(want).astype(np.float32))
