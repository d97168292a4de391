#!/usr/bin/env python
"""A/B of the resize-mode RoI forward variants (FRCNN_FWD_ASYNC = cp.async ring depth, 0 = register loads) at C1 x 64,
C5 x 1 and C5 x 8; every variant must reproduce the default kernel's output bit for bit."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from faster_rcnn_b200 import ops, synth          # noqa: E402
from benchmarks.stages import timeit, PEAK       # noqa: E402

h, w, c, p = 38, 63, 1024, 7
variants = sys.argv[1].split(",") if len(sys.argv) > 1 else ["0", "2", "3", "4"]
for tag, n_rois, batch in (("C1x64", 320, 64), ("C5x1", 2000, 1), ("C5x8", 2000, 8)):
    torch.manual_seed(0)
    feat = torch.randn((batch, h, w, c), device="cuda")
    if tag == "C1x64":      # the bench's RoIs: NMS output of the synthetic RPN heads, padded to 320
        from faster_rcnn_b200.util import get_anchors
        dims = get_anchors([128, 256, 512])
        pairs = [synth.rpn_outputs(h, w, 9, 1000 + i) for i in range(batch)]
        cls = torch.from_numpy(np.concatenate([q[0] for q in pairs])).cuda()
        regr = torch.from_numpy(np.concatenate([q[1] for q in pairs])).cuda()
        r, _, cnt = ops.proposals(regr, cls, dims, 16, 8000, 0.7, 300)
        rois = ops.pad_rois(r, cnt, 64)[0]
    else:
        rois = torch.from_numpy(np.stack([synth.random_rois(n_rois, h, w, 7 + i) for i in range(batch)])).cuda()
    nbytes = 4 * batch * h * w * c + 8 * batch * rois.shape[1] + 4 * batch * rois.shape[1] * p * p * c
    os.environ["FRCNN_FWD_ASYNC"] = "0"
    want = ops.roi_forward(feat, rois, p, "resize")
    for v in variants:
        os.environ["FRCNN_FWD_ASYNC"] = v
        got = ops.roi_forward(feat, rois, p, "resize")
        same = bool(torch.equal(got, want))
        ms = timeit(lambda: ops.roi_forward(feat, rois, p, "resize"), 20)
        print(json.dumps({"case": tag, "async_depth": int(v), "ms": round(ms, 4), "frac": round(nbytes / ms / 1e6 / PEAK, 3),
                          "bit_identical": same}), flush=True)
os.environ.pop("FRCNN_FWD_ASYNC", None)
