"""Error behaviour of the C ABI on a real device: bad arguments and unsupported sizes come back as negative status
codes with a message (raised as FrcnnError by the binding) -- nothing aborts, nothing falls back to the CPU."""
import numpy as np
import pytest

from helpers import dev

pytestmark = pytest.mark.gpu


def test_invalid_arguments_raise_with_message():
    import torch
    from faster_rcnn_b200 import _lib, ops
    from faster_rcnn_b200.runtime import get_context, ptr
    ctx = get_context()
    boxes = dev(np.zeros((1, 8, 4), np.int16))
    scores = dev(np.zeros((1, 8), np.float32))
    with pytest.raises(_lib.FrcnnError) as e:
        ops.nms_i16(boxes, scores, None, 0.7, 0)                       # max_boxes must be positive
    assert e.value.code == _lib.ERR_INVALID and "nms_i16" in str(e.value)
    with pytest.raises(_lib.FrcnnError) as e:                          # more anchors per cell than the table supports
        ops.decode_topk(dev(np.zeros((1, 2, 2, 4 * 65), np.float32)), dev(np.zeros((1, 2, 2, 65), np.float32)),
                        np.ones((65, 2), np.int64) * 16, 16, 10)
    assert e.value.code == _lib.ERR_INVALID
    with pytest.raises(_lib.FrcnnError) as e:                          # candidate list larger than one SM's shared memory
        ops.nms_i16(dev(np.zeros((1, 30000, 4), np.int16)), dev(np.zeros((1, 30000), np.float32)), None, 0.7, 300)
    assert e.value.code == _lib.ERR_UNSUPPORTED
    with pytest.raises(_lib.FrcnnError) as e:                          # top-k above the sort capacity
        ops.decode_topk(dev(np.zeros((1, 64, 64, 36), np.float32)), dev(np.zeros((1, 64, 64, 9), np.float32)),
                        np.ones((9, 2), np.int64) * 64, 16, 20000)
    assert e.value.code == _lib.ERR_UNSUPPORTED
    with pytest.raises(_lib.FrcnnError):                               # MAX mode without an arg-max buffer
        ctx.call("frcnn_roi_fwd", 1, ptr(torch.zeros(16, device="cuda")), 2, 2, 4, ptr(torch.zeros(4, dtype=torch.int16, device="cuda")),
                 0, 1, 7, 1, ptr(torch.zeros(49 * 4, device="cuda")), None)
    # the handle stays usable after errors
    ki, kc, _, _ = ops.nms_i16(boxes, scores + dev(np.arange(8, dtype=np.float32)[None]), None, 0.7, 4)
    assert int(kc[0]) >= 1


def test_python_layer_type_checks():
    import torch
    from faster_rcnn_b200 import det_util, ops
    with pytest.raises(TypeError):
        ops.nms_i16(torch.zeros((1, 4, 4), dtype=torch.int16), torch.zeros((1, 4)), None)       # CPU tensors
    with pytest.raises(TypeError):
        ops.nms_i16(dev(np.zeros((1, 4, 4), np.int32)), dev(np.zeros((1, 4), np.float32)), None)  # wrong dtype
    with pytest.raises(ValueError):
        ops.proposals(dev(np.zeros((1, 3, 3, 8), np.float32)), dev(np.zeros((1, 3, 3, 9), np.float32)), np.ones((9, 2), int), 16, 10)
    with pytest.raises(TypeError):
        det_util.nms(np.zeros((3, 4), np.int16), np.array([0.1, 0.2, 1 / 3], dtype=np.float64))  # not float32-representable
    with pytest.raises(ValueError):
        det_util.nms(np.zeros((5000, 4), np.float64), np.arange(5000, dtype=np.float32))         # float path capacity
