"""Pins the TF-1.x legacy bilinear semantics behind RoiResizeConv (custom_layers.py:50, tensorflow==1.3.0,
requirements.txt:53) with an independent EXECUTABLE implementation: OpenCV's cv2.dnn.

TensorFlow cannot be installed here, but cv2 4.13 ships a TensorFlow importer whose Resize layer implements the
TF `ResizeBilinear` op (align_corners=false, no half-pixel centres: src = i * in/out, lo = (int)src,
hi = min(lo + 1, in - 1), lerp = src - lo).  This script hand-encodes a three-node TensorFlow GraphDef
(Placeholder -> ResizeBilinear(size=Const[P, P])) in protobuf wire format, runs it with
cv2.dnn.readNetFromTensorflow and stores inputs and outputs in tests/golden/resize_bilinear_cv2dnn.npz.

    python tests/golden/make_golden_resize.py

OpenCV combines the four taps as  a + ly*(b-a) + lx*((c-a) + ly*(((d-c)-b)+a))  while TF's kernel computes
top + (bottom-top)*ly; tests/test_oracle_roi.py therefore checks (1) that the oracle's taps and lerp weights put
through OpenCV's expression reproduce cv2.dnn BIT FOR BIT (this pins coordinates, clamping and weights), and
(2) that the oracle's TF-form output differs from it by float32 rounding only (a few ulp, counted)."""
import os
import struct

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


def _varint(n):
    out = b''
    while True:
        b, n = n & 0x7f, n >> 7
        out += bytes([b | 0x80]) if n else bytes([b])
        if not n:
            return out


def _ld(field, payload):                       # length-delimited field
    return _varint((field << 3) | 2) + _varint(len(payload)) + payload


def _vi(field, n):                             # varint field
    return _varint(field << 3) + _varint(n)


def _node(name, op, inputs=(), attrs=()):      # tensorflow.NodeDef: name=1 op=2 input=3 attr=5 (map<string, AttrValue>)
    b = _ld(1, name.encode()) + _ld(2, op.encode())
    for i in inputs:
        b += _ld(3, i.encode())
    for k, v in attrs:
        b += _ld(5, _ld(1, k.encode()) + _ld(2, v))
    return _ld(1, b)                           # GraphDef.node = 1


def resize_bilinear_graphdef(h, w, c, pool):
    """serialized tensorflow.GraphDef: input (1,h,w,c) f32 -> ResizeBilinear(size=[pool,pool], align_corners=False)."""
    dt_float, dt_int32 = _vi(6, 1), _vi(6, 3)                                    # AttrValue.type = 6
    shape = _ld(7, b''.join(_ld(2, _vi(1, d)) for d in (1, h, w, c)))            # AttrValue.shape = 7, Dim.size = 1
    size = _ld(8, _vi(1, 3) + _ld(2, _ld(2, _vi(1, 2))) + _ld(4, struct.pack('<2i', pool, pool)))   # AttrValue.tensor = 8
    g = _node('input', 'Placeholder', attrs=[('dtype', dt_float), ('shape', shape)])
    g += _node('size', 'Const', attrs=[('dtype', dt_int32), ('value', size)])
    g += _node('resize', 'ResizeBilinear', ['input', 'size'], attrs=[('T', dt_float), ('align_corners', _vi(5, 0))])
    return g


def cv2_resize_bilinear(feat_hwc, pool):
    """feat (h,w,c) f32 -> (pool,pool,c) f32 through cv2.dnn's TensorFlow importer."""
    import cv2
    h, w, c = feat_hwc.shape
    net = cv2.dnn.readNetFromTensorflow(np.frombuffer(resize_bilinear_graphdef(h, w, c, pool), np.uint8))
    net.setInput(np.ascontiguousarray(feat_hwc.transpose(2, 0, 1))[None])
    return np.ascontiguousarray(net.forward()[0].transpose(1, 2, 0))


SHAPES = [(1, 1), (1, 5), (4, 1), (2, 3), (3, 2), (5, 9), (6, 6), (7, 7), (8, 8), (13, 6), (14, 21), (15, 15), (3, 40),
          (20, 2), (23, 24), (38, 63)]


def main():
    import cv2
    rng = np.random.default_rng(20261017)
    arrays = {"cv2_version": np.array(cv2.__version__)}
    for i, (h, w) in enumerate(SHAPES):
        for pool in (7, 3):
            x = rng.standard_normal((h, w, 4), dtype=np.float32)
            arrays["x_%d_%d" % (i, pool)] = x
            arrays["y_%d_%d" % (i, pool)] = cv2_resize_bilinear(x, pool)
    path = os.path.join(HERE, "resize_bilinear_cv2dnn.npz")
    np.savez_compressed(path, **arrays)
    print("%s  %d bytes, cv2 %s" % (path, os.path.getsize(path), cv2.__version__))


if __name__ == "__main__":
    main()
