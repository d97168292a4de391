import sys, numpy as np, torch
sys.path.insert(0, '/root/repo')
from faster_rcnn_b200 import ops, synth
from faster_rcnn_b200.util import get_anchors
b=int(sys.argv[1]); k=int(sys.argv[2]); post=int(sys.argv[3])
voc=get_anchors([128,256,512])
pairs=[synth.rpn_outputs(38,63,9,100+i,clustered=True) for i in range(b)]
cls=torch.from_numpy(np.concatenate([p[0] for p in pairs])).cuda(); regr=torch.from_numpy(np.concatenate([p[1] for p in pairs])).cuda()
for _ in range(4): ops.proposals(regr,cls,voc,16,k,0.7,post)
torch.cuda.synchronize()
# quick A/B timing (hot L2; benchmarks/stages.py is the reported measurement): decode + top-k alone, then with NMS
def _time(fn, iters=200):
    for _ in range(20): fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); a.record()
    for _ in range(iters): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / iters * 1e3
print("PROP b=%d k=%d  decode_topk %.1f us   proposals(+nms %d) %.1f us" % (
    b, k, _time(lambda: ops.decode_topk(regr, cls, voc, 16, k)), post, _time(lambda: ops.proposals(regr, cls, voc, 16, k, 0.7, post))))
# GPU time without the Python / launch overhead: 20 calls captured in one CUDA graph
def _graph_time(fn, calls=20, replays=20):
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(3): fn()
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(calls): fn()
    for _ in range(3): g.replay()
    a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); a.record()
    for _ in range(replays): g.replay()
    e.record(); torch.cuda.synchronize()
    return a.elapsed_time(e) / (calls * replays) * 1e3
print("PROPGRAPH b=%d k=%d  decode_topk %.1f us   proposals(+nms %d) %.1f us" % (
    b, k, _graph_time(lambda: ops.decode_topk(regr, cls, voc, 16, k)), post, _graph_time(lambda: ops.proposals(regr, cls, voc, 16, k, 0.7, post))))
