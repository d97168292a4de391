import sys, os
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from faster_rcnn_b200 import ops, synth
from faster_rcnn_b200.util import get_anchors
voc = get_anchors([128, 256, 512])
batch = int(sys.argv[1]) if len(sys.argv) > 1 else 1
pairs = [synth.rpn_outputs(38, 63, 9, 100 + i, clustered=True) for i in range(batch)]
cls = torch.from_numpy(np.concatenate([p[0] for p in pairs])).cuda(); regr = torch.from_numpy(np.concatenate([p[1] for p in pairs])).cuda()
tb, ts, _, tc = ops.decode_topk(regr, cls, voc, 16, 12000)
for _ in range(4):
    out = ops.nms_i16(tb, ts, tc, 0.7, 2000)
torch.cuda.synchronize()
print(int(out[1].float().mean().item()))
