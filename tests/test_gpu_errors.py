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
                        np.ones((9, 2), np.int64) * 64, 16, 36000)
    assert e.value.code == _lib.ERR_UNSUPPORTED
    with pytest.raises(_lib.FrcnnError):                               # MAX mode without an arg-max buffer
        ctx.call("frcnn_roi_fwd", 1, ptr(torch.zeros(16, device="cuda")), 2, 2, 4, ptr(torch.zeros(4, dtype=torch.int16, device="cuda")),
                 0, 1, 7, 1, ptr(torch.zeros(49 * 4, device="cuda")), None)
    # the handle stays usable after errors
    ki, kc, _, _ = ops.nms_i16(boxes, scores + dev(np.arange(8, dtype=np.float32)[None]), None, 0.7, 4)
    assert int(kc[0]) >= 1


def test_nms_i16_oversized_unsorted_input_fails_safe():
    """More than FRCNN_NMS_MAX_UNSORTED candidates with tied scores (ties are 'not strictly descending'): the kernel
    must not sort past its shared-memory buffer; the image is flagged with keep_count = -1 and the other image of the
    same launch (strictly descending scores, same size) is processed normally."""
    from faster_rcnn_b200 import ops
    from oracle import frcnn_oracle as O
    n = 20000
    rng = np.random.default_rng(0)
    x1, y1 = rng.integers(0, 50, (2, n)), rng.integers(0, 30, (2, n))
    boxes = np.stack([x1, y1, x1 + rng.integers(1, 12, (2, n)), y1 + rng.integers(1, 12, (2, n))], axis=2).astype(np.int16)
    scores = np.stack([np.full(n, 0.5, np.float32), np.linspace(1.0, 0.01, n, dtype=np.float32)])
    assert np.all(np.diff(scores[1]) < 0)
    ki, kc, kb, ks = ops.nms_i16(dev(boxes), dev(scores), None, 0.7, 300)
    kc, ki = kc.cpu().numpy(), ki.cpu().numpy()
    assert kc[0] == -1 and np.all(ki[0] == -1)
    pick = O.greedy_nms(boxes[1], scores[1], 0.7, 300)
    assert kc[1] == len(pick) and np.array_equal(ki[1, :kc[1]], pick)


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


def test_new_entry_points_reject_bad_arguments():
    import torch
    from faster_rcnn_b200 import _lib, ops, synth
    from faster_rcnn_b200.pipeline import DetTrainingPipeline, ProposalRoiPipeline
    from faster_rcnn_b200.runtime import get_context, ptr
    ctx = get_context()
    rois, y_cls = dev(np.zeros((2, 5, 4), np.int16)), dev(np.zeros((2, 5, 21), np.int32))
    y_tr, index = dev(np.zeros((2, 5, 160), np.float32)), dev(np.zeros((2, 3), np.int32))
    with pytest.raises(ValueError):
        ops.gather_det_samples(rois, y_cls, y_tr[:, :, :100].contiguous(), index)               # 8(K-1) columns expected
    with pytest.raises(_lib.FrcnnError) as e:                                                      # NULL output through the raw ABI
        ctx.call("frcnn_gather_det_samples", ptr(rois), ptr(y_cls), ptr(y_tr), ptr(index), 5, 21, 3, 2, None, None, None)
    assert e.value.code == _lib.ERR_INVALID and "gather_det_samples" in str(e.value)
    # out-of-range and -1 sample rows give zero rows, never an out-of-bounds read
    out = ops.gather_det_samples(rois + 7, y_cls + 1, y_tr + 1.0, dev(np.array([[0, -1, 99], [4, 5, 2]], np.int32)))
    got = [t.cpu().numpy() for t in out]
    assert got[0][0, 0].tolist() == [7, 7, 7, 7] and not got[0][0, 1:].any() and not got[1][1, 1].any() and got[2][1, 2, 0] == 1.0
    with pytest.raises(NotImplementedError):
        DetTrainingPipeline({'bg': 0, 'cat': 1})                                                  # 'bg' must be the last class
    # a captured graph keeps its shapes: other shapes must be captured separately
    dims = np.array([[8, 8], [16, 16]]) * 16
    pipe = ProposalRoiPipeline(dims, 16, 50, 0.7, 10, 8, 7, "resize")
    cls, regr = synth.rpn_outputs(6, 7, 2, 1)
    run = pipe.capture(dev(cls), dev(regr), torch.randn((1, 6, 7, 16), device="cuda"))
    with pytest.raises(RuntimeError):
        run(feat=torch.randn((1, 6, 7, 32), device="cuda"))
