#!/usr/bin/env python
"""A/B of the RoI backward variants (env knobs of roi_bwd.cu) at C1 x 64, C5 x 1, C5 x 8, both modes.

    python benchmarks/bwd_ab.py [--json gpurun_out/bwd_ab.json] [--modes resize,max]

Every variant is checked against the round-1 cell-stationary kernels (FRCNN_BWD_IMPL=cell) before it is timed."""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from faster_rcnn_b200 import ops, synth          # noqa: E402
from benchmarks.stages import timeit, PEAK       # noqa: E402

KNOBS = ("FRCNN_BWD_IMPL", "FRCNN_BWD_CPB", "FRCNN_BWD_PARTS")


def setenv(cfg):
    for k in KNOBS:
        os.environ.pop(k, None)
    for k, v in cfg.items():
        os.environ[k] = str(v)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--json", default="")
    ap.add_argument("--modes", default="resize,max")
    ap.add_argument("--iters", type=int, default=10)
    args = ap.parse_args()
    h, w, c, p = 38, 63, 1024, 7
    variants = {
        "resize": [{"FRCNN_BWD_IMPL": "cell"}, {}, {"FRCNN_BWD_CPB": 1},
                   {"FRCNN_BWD_PARTS": 1}, {"FRCNN_BWD_PARTS": 2}, {"FRCNN_BWD_PARTS": 4}, {"FRCNN_BWD_PARTS": 8}],
        "max": [{"FRCNN_BWD_IMPL": "cell"}, {},
                {"FRCNN_BWD_PARTS": 1}, {"FRCNN_BWD_PARTS": 2}, {"FRCNN_BWD_PARTS": 4}, {"FRCNN_BWD_PARTS": 8}],
    }
    res = []
    for tag, n_rois, batch in (("C5x1", 2000, 1), ("C5x8", 2000, 8), ("C1x64", 320, 64)):
        torch.manual_seed(0)
        feat = torch.randn((batch, h, w, c), device="cuda")
        rois = torch.from_numpy(np.stack([synth.random_rois(n_rois, h, w, 7 + i) for i in range(batch)])).cuda()
        gout = torch.randn((batch, n_rois, p, p, c), device="cuda")
        for mode in args.modes.split(","):
            if mode == "maxc":                                    # max mode with the one-byte arg-max, forward + backward
                nb = 4 * batch * h * w * c + 8 * batch * n_rois + 5 * batch * n_rois * p * p * c
                code = ops.roi_forward(feat, rois, p, "max", compact=True)[1]
                want = ops.roi_backward(gout, rois, (batch, h, w, c), "max", ops.roi_forward(feat, rois, p, "max")[1])
                ms = timeit(lambda: ops.roi_forward(feat, rois, p, "max", compact=True), args.iters, 3)
                row = {"case": tag, "mode": "maxc_fwd", "ms": round(ms, 4), "frac_5B": round(nb / ms / 1e6 / PEAK, 3)}
                res.append(row)
                print(json.dumps(row), flush=True)
                for cfg in ({}, {"FRCNN_BWD_CPB": 1}):
                    setenv(cfg)
                    got = ops.roi_backward(gout, rois, (batch, h, w, c), "max", code)
                    err = (got - want).abs().max().item() / want.abs().max().item()
                    ms = timeit(lambda: ops.roi_backward(gout, rois, (batch, h, w, c), "max", code), args.iters, 3)
                    row = {"case": tag, "mode": "maxc_bwd", "cfg": cfg, "ms": round(ms, 4), "frac_5B": round(nb / ms / 1e6 / PEAK, 3),
                           "rel_err_vs_int32": err}
                    res.append(row)
                    print(json.dumps(row), flush=True)
                setenv({})
                del code, want
                continue
            arg = ops.roi_forward(feat, rois, p, "max")[1] if mode == "max" else None
            nbytes = 4 * batch * h * w * c + 8 * batch * n_rois + 4 * batch * n_rois * p * p * c * (2 if mode == "max" else 1)
            setenv({"FRCNN_BWD_IMPL": "cell"})
            want = ops.roi_backward(gout, rois, (batch, h, w, c), mode, arg)
            scale = want.abs().max().item()
            for cfg in variants[mode]:
                setenv(cfg)
                try:
                    got = ops.roi_backward(gout, rois, (batch, h, w, c), mode, arg)
                    torch.cuda.synchronize()
                    err = (got - want).abs().max().item() / scale
                    again = ops.roi_backward(gout, rois, (batch, h, w, c), mode, arg)
                    same = bool(torch.equal(got, again))
                    ms = timeit(lambda: ops.roi_backward(gout, rois, (batch, h, w, c), mode, arg), args.iters, 3)
                    row = {"case": tag, "mode": mode, "cfg": cfg, "ms": round(ms, 4), "frac": round(nbytes / ms / 1e6 / PEAK, 3),
                           "rel_err_vs_cell": err, "bit_reproducible": same}
                except Exception as e:        # noqa: BLE001
                    row = {"case": tag, "mode": mode, "cfg": cfg, "error": str(e)[:200]}
                res.append(row)
                print(json.dumps(row), flush=True)
            del arg, want
        del feat, gout
    setenv({})
    if args.json:
        os.makedirs(os.path.dirname(os.path.abspath(args.json)), exist_ok=True)
        json.dump(res, open(args.json, "w"), indent=1)


if __name__ == "__main__":
    main()
