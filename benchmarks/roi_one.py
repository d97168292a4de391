#!/usr/bin/env python
"""One RoI-layer configuration, a few launches (the target of ncu captures):
    python benchmarks/roi_one.py --dir bwd --mode resize --rois 2000 --batch 8 [--reps 3]"""
import argparse
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from faster_rcnn_b200 import ops, synth          # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--dir", default="bwd")
ap.add_argument("--mode", default="resize")
ap.add_argument("--rois", type=int, default=2000)
ap.add_argument("--batch", type=int, default=8)
ap.add_argument("--reps", type=int, default=3)
a = ap.parse_args()
h, w, c, p = 38, 63, 1024, 7
feat = torch.randn((a.batch, h, w, c), device="cuda")
rois = torch.from_numpy(np.stack([synth.random_rois(a.rois, h, w, 7 + i) for i in range(a.batch)])).cuda()
gout = torch.randn((a.batch, a.rois, p, p, c), device="cuda")
arg = ops.roi_forward(feat, rois, p, "max")[1] if a.mode == "max" else None
for _ in range(a.reps):
    if a.dir == "bwd":
        ops.roi_backward(gout, rois, (a.batch, h, w, c), a.mode, arg)
    else:
        ops.roi_forward(feat, rois, p, a.mode)
torch.cuda.synchronize()
