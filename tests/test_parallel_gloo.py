"""N > 1 path on CPU: two gloo ranks shard a batch of images, post-process their block with the ORACLE standing in
for the GPU kernels (the kernels need a device; the sharding / packing / all-gather plumbing does not), and the
gathered buffer must equal the single-process result."""
import os
import socket
import sys

import numpy as np
import torch
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _dets_for(image_ids, m=96):
    """oracle post-processing of the given images -> fixed-size tensors like frcnn_det_postprocess emits."""
    from faster_rcnn_b200 import synth
    from oracle import frcnn_oracle as O
    boxes = torch.zeros((len(image_ids), m, 4), dtype=torch.int32)
    probs = torch.zeros((len(image_ids), m), dtype=torch.float32)
    cls = torch.full((len(image_ids), m), -1, dtype=torch.int32)
    count = torch.zeros(len(image_ids), dtype=torch.int32)
    for b, i in enumerate(image_ids):
        oc, orr = synth.detector_outputs(m, 21, 500 + i)
        dets = O.det_postprocess(synth.random_rois(m, 37, 62, 600 + i), oc, orr, 20, 16, 1.6)
        count[b] = len(dets)
        for r, (c, box, p) in enumerate(dets):
            boxes[b, r], probs[b, r], cls[b, r] = torch.from_numpy(box.astype(np.int32)), float(p), c
    return boxes, probs, cls, count


def _worker(rank, world, port, n_images, out_dir):
    sys.path.insert(0, ROOT)
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    from faster_rcnn_b200 import parallel
    r, w, _ = parallel.init_from_env(backend="gloo")
    start, stop = parallel.shard_range(n_images, r, w)
    dets, counts = parallel.pack_detections(*_dets_for(range(start, stop)))
    all_d, all_c = parallel.all_gather_detections(dets, counts)
    # the proposal stage's exchange: int16 RoI rows travel as int32 views
    rois = (torch.arange((stop - start) * 5 * 4, dtype=torch.int32).reshape(stop - start, 5, 4) + 1000 * rank).to(torch.int16)
    g_rois, g_cnt = parallel.all_gather_rois(rois, torch.full((stop - start,), rank, dtype=torch.int32))
    assert g_rois.dtype == torch.int16 and g_rois.shape == (n_images, 5, 4)
    assert torch.equal(g_rois[start:stop], rois) and g_cnt.tolist() == [0] * (n_images // 2) + [1] * (n_images // 2)
    # asynchronous form (overlaps the RoI layer in bench.py): same buffers after pending.wait()
    a_rois, a_cnt, pending = parallel.all_gather_rois(rois, torch.full((stop - start,), rank, dtype=torch.int32), async_op=True)
    pending.wait()
    assert torch.equal(a_rois, g_rois) and torch.equal(a_cnt, g_cnt)
    torch.save((all_d, all_c), os.path.join(out_dir, "rank%d.pt" % rank))
    torch.distributed.barrier()
    torch.distributed.destroy_process_group()


def test_two_rank_gloo_all_gather_of_detections(tmp_path):
    n_images, world = 6, 2
    mp.start_processes(_worker, args=(world, _free_port(), n_images, str(tmp_path)), nprocs=world, join=True,
                       start_method="spawn")
    sys.path.insert(0, ROOT)
    from faster_rcnn_b200 import parallel
    want_d, want_c = parallel.pack_detections(*_dets_for(range(n_images)))
    for rank in range(world):
        got_d, got_c = torch.load(os.path.join(str(tmp_path), "rank%d.pt" % rank))
        assert torch.equal(got_d, want_d) and torch.equal(got_c, want_c)
    assert int(want_c.sum()) > 0 and want_d.shape == (n_images, 96, 6)
    # rows beyond the count are zero, live rows carry [x1,y1,x2,y2,prob,class]
    assert float(want_d[0, int(want_c[0]):].abs().sum()) == 0.0
    # single process: identity
    d1, c1 = parallel.all_gather_detections(want_d, want_c)
    assert d1 is want_d and c1 is want_c
