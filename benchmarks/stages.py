#!/usr/bin/env python
"""Per-stage device timings at the BASELINE.json configs (C1..C5) with CUDA events.

    python benchmarks/stages.py [--iters 20] [--only roi_fwd,nms] [--json gpurun_out/stages.json]

Every stage is timed alone on device-resident inputs (warm-up, then the mean of `iters` launches bracketed by
events on the launching stream); HBM-bound stages also report algorithmic GB/s against the measured copy peak.
Inputs are larger than L2 where the stage's working set allows it; the small latency-bound stages (NMS, top-k,
labelling) are L2-resident by nature and are reported as latency per image.  This script is the source of the
per-kernel tables in DESIGN.md / profiles/ and the target of the ncu captures; bench.py is the contract line."""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from faster_rcnn_b200 import ops, synth          # noqa: E402
from faster_rcnn_b200.util import get_anchors    # noqa: E402

PEAK = 6551.0
try:
    PEAK = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass


def dev(x, dtype=None):
    return torch.from_numpy(np.ascontiguousarray(x, dtype=dtype)).cuda()


def timeit(fn, iters, warmup=5, min_ms=10.0):
    """mean ms per call.  Warm-up and measurement each cover at least `min_ms` of GPU time, so that a 0.1 ms kernel is
    not timed on clocks that are still ramping.  This is a BURST protocol like MEASURED_PEAKS.json's copy figure
    (best of 10 x 0.66 ms): kept running for hundreds of milliseconds the RoI kernels reach the 1000 W power cap
    (bench.py --steps 300 reports sw_power_cap) and run 6-8 % slower."""
    fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(warmup):
        fn()
    b.record()
    torch.cuda.synchronize()
    per_call = max(a.elapsed_time(b) / warmup, 1e-3)
    for _ in range(max(0, int(min_ms / per_call) - warmup)):
        fn()
    iters = max(iters, int(min_ms / per_call))
    torch.cuda.synchronize()
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


def graph_timeit(fn, calls=20, replays=20):
    """mean ms per call with `calls` back-to-back calls captured in one CUDA graph: the GPU time of a stage whose eager
    call is bound by the host (ctypes + torch.empty + launch, about 30 us per call for the proposal stage)."""
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(3):
            fn()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(calls):
            fn()
    for _ in range(3):
        g.replay()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    a.record()
    for _ in range(replays):
        g.replay()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / (calls * replays)


def rpn_batch(rows, cols, dims, batch, seed, clustered):
    pairs = [synth.rpn_outputs(rows, cols, len(dims), seed + i, clustered=clustered) for i in range(batch)]
    return dev(np.concatenate([p[0] for p in pairs])), dev(np.concatenate([p[1] for p in pairs]))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--only", default="")
    ap.add_argument("--json", default="")
    ap.add_argument("--modes", default="resize,max", help="RoI layer modes to time")
    args = ap.parse_args()
    only = set(filter(None, args.only.split(",")))
    want = lambda name: not only or any(name.startswith(o) for o in only)   # noqa: E731
    res = []

    def rec(name, config, ms, images=1, algo_bytes=None, **extra):
        row = {"stage": name, "config": config, "ms": round(ms, 5), "images": images, "ms_per_image": round(ms / images, 5)}
        if algo_bytes is not None:
            row.update(algo_MB=round(algo_bytes / 1e6, 2), GBps=round(algo_bytes / ms / 1e6, 1),
                       frac_of_measured_peak=round(algo_bytes / ms / 1e6 / PEAK, 4))
        row.update(extra)
        res.append(row)
        print(json.dumps(row), flush=True)

    voc, kitti = get_anchors([128, 256, 512]), get_anchors()
    # ---- proposal stage: C1 (VOC test), C3 (KITTI train), batch 1 (latency) and batch 64 (throughput) ------------
    for tag, rows, cols, dims, k, post, clustered in (("C1 voc 8000->300", 38, 63, voc, 8000, 300, False),
                                                      ("C1 voc clustered 8000->300", 38, 63, voc, 8000, 300, True),
                                                      ("C1 voc train 12000->2000 clustered", 38, 63, voc, 12000, 2000, True),
                                                      ("C3 kitti 12000->2000 clustered", 38, 94, kitti, 12000, 2000, True),
                                                      ("C3 kitti 8000->300", 38, 94, kitti, 8000, 300, False)):
        for batch in (1, 64):
            if not (want("decode") or want("nms") or want("proposals")):
                continue
            cls, regr = rpn_batch(rows, cols, dims, batch, 100, clustered)
            n = rows * cols * len(dims)
            if want("decode"):
                ms = timeit(lambda: ops.decode_topk(regr, cls, dims, 16, k), args.iters)
                rec("decode_topk", "%s b%d" % (tag, batch), ms, batch, batch * (n * 20 + min(k, n) * 16),
                    graph_ms=round(graph_timeit(lambda: ops.decode_topk(regr, cls, dims, 16, k)), 5))
            tb, ts, _, tc = ops.decode_topk(regr, cls, dims, 16, k)
            if want("nms"):
                ms = timeit(lambda: ops.nms_i16(tb, ts, tc, 0.7, post), args.iters)
                kept = int(ops.nms_i16(tb, ts, tc, 0.7, post)[1].float().mean().item())
                rec("nms_i16", "%s b%d" % (tag, batch), ms, batch, kept_mean=kept,
                    graph_ms=round(graph_timeit(lambda: ops.nms_i16(tb, ts, tc, 0.7, post)), 5))
            if want("proposals"):
                ms = timeit(lambda: ops.proposals(regr, cls, dims, 16, k, 0.7, post), args.iters)
                rec("proposals_fused", "%s b%d" % (tag, batch), ms, batch,
                    graph_ms=round(graph_timeit(lambda: ops.proposals(regr, cls, dims, 16, k, 0.7, post)), 5))

    # ---- RoI layer: C1 (320 RoIs) x 64 images, C5 (2000 RoIs, 1 image and 8 images), forward + backward, both modes --
    for tag, n_rois, batch in (("C1 320 rois", 320, 64), ("C5 2000 rois", 2000, 1), ("C5 2000 rois", 2000, 8)):
        if not (want("roi_fwd") or want("roi_bwd")):
            continue
        h, w, c, p = 38, 63, 1024, 7
        feat = torch.randn((batch, h, w, c), device="cuda")
        rois = dev(np.stack([synth.random_rois(n_rois, h, w, 7 + i) for i in range(batch)]))
        out_b = 4 * batch * n_rois * p * p * c
        in_b = 4 * batch * h * w * c + 8 * batch * n_rois
        for mode in args.modes.split(","):
            if want("roi_fwd"):
                ms = timeit(lambda: ops.roi_forward(feat, rois, p, mode), args.iters)
                rec("roi_fwd_" + mode, "%s b%d" % (tag, batch), ms, batch, in_b + out_b * (2 if mode == "max" else 1))
            if want("roi_bwd"):
                gout = torch.randn((batch, n_rois, p, p, c), device="cuda")
                arg = ops.roi_forward(feat, rois, p, "max")[1] if mode == "max" else None
                ms = timeit(lambda: ops.roi_backward(gout, rois, (batch, h, w, c), mode, arg), max(3, args.iters // 4), 2)
                rec("roi_bwd_" + mode, "%s b%d" % (tag, batch), ms, batch, in_b + out_b * (2 if mode == "max" else 1))
                del gout, arg
        del feat

    # ---- library comparison for the max mode: torchvision.ops.roi_pool (NCHW, atomics in the backward) on the same RoIs
    if want("tv_roi_pool"):
        try:
            import torchvision
        except ImportError:
            torchvision = None
        for tag, n_rois, batch in (("C1 320 rois", 320, 64), ("C5 2000 rois", 2000, 1), ("C5 2000 rois", 2000, 8)):
            if torchvision is None:
                break
            h, w, c, p = 38, 63, 1024, 7
            feat = torch.randn((batch, c, h, w), device="cuda", requires_grad=True)
            r = np.stack([synth.random_rois(n_rois, h, w, 7 + i) for i in range(batch)]).astype(np.float32)
            idx = np.repeat(np.arange(batch, dtype=np.float32), n_rois)[:, None]
            boxes = dev(np.concatenate([idx, r.reshape(-1, 4)[:, :2], r.reshape(-1, 4)[:, 2:] - 1], axis=1))
            nbytes = 4 * batch * h * w * c + 8 * batch * n_rois + 2 * 4 * batch * n_rois * p * p * c
            ms = timeit(lambda: torchvision.ops.roi_pool(feat, boxes, p, 1.0), max(3, args.iters // 4), 2)
            rec("tv_roi_pool_fwd", "%s b%d (torchvision %s, library kernel)" % (tag, batch, torchvision.__version__), ms, batch, nbytes)
            out = torchvision.ops.roi_pool(feat, boxes, p, 1.0)
            g = torch.randn_like(out)

            def bwd():
                feat.grad = None
                out.backward(g, retain_graph=True)
            ms = timeit(bwd, max(3, args.iters // 4), 2)
            rec("tv_roi_pool_bwd", "%s b%d (torchvision, atomics)" % (tag, batch), ms, batch, nbytes)
            del feat, out, g

    # ---- device-resident pipeline (decode -> top-k -> NMS -> pad -> RoI layer), eager launches vs one CUDA graph ------
    if want("pipeline"):
        from faster_rcnn_b200.pipeline import ProposalRoiPipeline
        pipe = ProposalRoiPipeline(voc, 16, 8000, 0.7, 300, 64, 7, "resize")
        for batch in (1, 8, 64):
            cls, regr = rpn_batch(38, 63, voc, batch, 100, False)
            feat = torch.randn((batch, 38, 63, 1024), device="cuda")
            ms = timeit(lambda: pipe.run_device(cls, regr, feat), args.iters)
            rec("pipeline_eager", "C1 b%d" % batch, ms, batch)
            run = pipe.capture(cls, regr, feat)
            ms = timeit(lambda: run(), args.iters)
            rec("pipeline_graph", "C1 b%d" % batch, ms, batch)
            del run, feat

    # ---- C4 training targets: batch 128, 50 GT ---------------------------------------------------------------------
    if want("label"):
        batch, rows, cols = 128, 38, 63
        gts = np.stack([np.array([g[1:] for g in synth.gt_boxes(50, 1000, 600, 300 + i)], np.float32) for i in range(batch)])
        gt, n_gt = dev(gts), dev(np.full(batch, 50, np.int32))
        wh = dev(np.tile(np.array([[1000, 600]], np.int32), (batch, 1)))
        ms = timeit(lambda: ops.label_anchors(gt, n_gt, wh, rows, cols, voc, 16), args.iters)
        n = rows * cols * 9
        rec("label_anchors", "C4 voc 50gt b128", ms, batch, batch * (16 * 50 + n * 18))
        cu, ip, bb, _ = ops.label_anchors(gt, n_gt, wh, rows, cols, voc, 16)
        ms = timeit(lambda: ops.pack_rpn_targets(cu, ip, bb, rows, cols, 9), args.iters)
        rec("pack_rpn_targets", "C4 voc b128", ms, batch, batch * (n * 18 + rows * cols * (18 + 72 * 4)))
        rois = dev(np.stack([synth.random_rois(2000, rows, cols, 400 + i) for i in range(batch)]))
        gt64 = dev(gts.astype(np.float64) / 16)
        gcls = dev(np.tile(np.arange(50, dtype=np.int32) % 20, (batch, 1)))
        ms = timeit(lambda: ops.label_rois(rois, gt64, gcls, n_gt, 21), args.iters)
        rec("label_rois", "C4 2000 rois 50gt b128", ms, batch, batch * 2000 * (8 + 8 + 84 + 640 + 4))

    # ---- C4 detector-training inputs, batch 128: proposals 12000 -> 2000, RoI x GT labelling, host draw, gather, RoI layer
    if want("det_training"):
        from faster_rcnn_b200.pipeline import DetTrainingPipeline
        batch, rows, cols = 128, 38, 63
        cls, regr = rpn_batch(rows, cols, voc, batch, 700, True)
        feat = torch.randn((batch, rows, cols, 1024), device="cuda")
        gts = np.stack([np.array([g[1:] for g in synth.gt_boxes(50, 1000, 600, 300 + i)], np.float64) for i in range(batch)]) / 16
        gt, n_gt = dev(gts), dev(np.full(batch, 50, np.int32))
        gcls = dev(np.tile(np.arange(50, dtype=np.int32) % 20, (batch, 1)))
        pipe = DetTrainingPipeline(synth.VOC_CLASS_MAPPING, voc)
        np.random.seed(0)

        def run():
            r, yc, yt, _ = pipe.targets(cls, regr, gt, gcls, n_gt)
            return ops.roi_forward(feat, r, 7, "resize")
        ms = timeit(run, max(3, args.iters // 4), 2)
        rec("det_training_inputs+roi_fwd", "C4 voc 12000->2000, 50 gt, 64 samples b128 (incl. host RNG draw)", ms, batch)
        del feat

    # ---- C2 detector post-processing: 64 images x 320 rows x 21 classes ------------------------------------------------
    if want("postprocess"):
        batch = 64
        rois = dev(np.stack([synth.random_rois(320, 37, 62, 500 + i) for i in range(batch)]))
        outs = [synth.detector_outputs(320, 21, 600 + i) for i in range(batch)]
        oc, orr = dev(np.stack([o[0] for o in outs])), dev(np.stack([o[1] for o in outs]))
        ratio = dev(np.full(batch, 1.6))
        ms = timeit(lambda: ops.det_postprocess(rois, oc, orr, ratio, 20), args.iters)
        rec("det_postprocess", "C2 vgg16 320 rows 21 cls b64", ms, batch)

    # ---- widening rows: masked losses (8f-2) and VOC evaluation (8f-3) ----------------------------------------------
    if want("losses"):
        batch, rows, cols = 128, 38, 63
        n = rows * cols * 9
        gts = np.stack([np.array([g[1:] for g in synth.gt_boxes(50, 1000, 600, 300 + i)], np.float32) for i in range(batch)])
        cu, ip, bb, _ = ops.label_anchors(dev(gts), dev(np.full(batch, 50, np.int32)),
                                          dev(np.tile(np.array([[1000, 600]], np.int32), (batch, 1))), rows, cols, voc, 16)
        cp, rp = torch.rand((batch, n), device="cuda"), torch.randn((batch, n, 4), device="cuda")
        ms = timeit(lambda: ops.rpn_losses(cu, ip, bb, cp, rp, want_grad=True), args.iters)
        rec("rpn_losses+grad", "C4 voc b128", ms, batch, batch * n * (2 + 16 + 4 + 16 + 4 + 16))
        yc = torch.zeros((batch, 64, 21), dtype=torch.int32, device="cuda")
        yc[..., -1] = 1
        yt = torch.zeros((batch, 64, 160), device="cuda")
        pc, pr = torch.softmax(torch.randn((batch, 64, 21), device="cuda"), dim=2), torch.randn((batch, 64, 80), device="cuda")
        ms = timeit(lambda: ops.det_losses(yc, yt, pc, pr, want_grad=True), args.iters)
        rec("det_losses+grad", "64 rois 21 cls b128", ms, batch)
    if want("voc_eval"):
        rng = np.random.default_rng(0)
        n_img, per_img_gt, per_img_det = 4952, 3, 12                      # VOC2007 test size, one class
        g_xy = rng.uniform(0, 400, (n_img * per_img_gt, 2))
        gt = np.concatenate([g_xy, g_xy + rng.uniform(20, 200, (n_img * per_img_gt, 2))], axis=1)
        d_xy = rng.uniform(0, 400, (n_img * per_img_det, 2))
        det = np.concatenate([d_xy, d_xy + rng.uniform(20, 200, (n_img * per_img_det, 2))], axis=1)
        rank_img = rng.integers(0, n_img, n_img * per_img_det)
        by_img = np.argsort(rank_img, kind='stable').astype(np.int32)
        d_off = np.concatenate([[0], np.cumsum(np.bincount(rank_img, minlength=n_img))]).astype(np.int32)
        g_off = (np.arange(n_img + 1) * per_img_gt).astype(np.int32)
        a = [dev(det), dev(d_off), dev(by_img), dev(gt), dev(np.zeros(len(gt), np.uint8)), dev(g_off)]
        thr = dev(np.arange(0., 1.1, 0.1))

        def run():
            tp, fp = ops.voc_match(*a, 0.5)
            return ops.voc_pr_ap(tp, fp, float(len(gt)), thr)
        ms = timeit(run, args.iters)
        rec("voc_match+pr_ap", "4952 images x 12 dets x 3 gt, one class", ms, n_img)

    if args.json:
        os.makedirs(os.path.dirname(os.path.abspath(args.json)), exist_ok=True)
        json.dump(res, open(args.json, "w"), indent=1)


if __name__ == "__main__":
    main()
