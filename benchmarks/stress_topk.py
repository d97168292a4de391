"""Repeated decode + top-k calls over a mix of map sizes and k against the numpy oracle (bit-exact indices, boxes,
counts): python benchmarks/stress_topk.py [repetitions]."""
import sys, numpy as np, torch
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
from faster_rcnn_b200 import ops, synth
from oracle import frcnn_oracle as O
dims = O.anchor_table([128,256,512])
bad = 0; total = 0
cases = [(5,7,8000),(10,12,8000),(10,12,500),(20,20,8000),(20,20,2500),(38,63,8000),(38,63,12000),(38,63,300),(30,40,6000),(38,94,12000)]
for rep in range(int(sys.argv[1]) if len(sys.argv) > 1 else 5):
    for (rows, cols, k) in cases:
        cls, regr = synth.rpn_outputs(rows, cols, len(dims), 100 + rep, clustered=bool(rep % 2))
        b,s,i,c,d = ops.decode_topk(torch.from_numpy(regr).cuda(), torch.from_numpy(cls).cuda(), dims, 16, k, want_dense=True)
        dense = d.cpu().numpy()[0]
        wb, wp, widx = O.topk_proposals(dense.copy(), cls.reshape(-1), k)
        n = int(c.cpu().numpy()[0]); idx = i.cpu().numpy()[0]
        ok = n == len(wb) and np.array_equal(idx[:n], widx) and np.all(idx[n:] == -1) and np.array_equal(b.cpu().numpy()[0,:n], wb)
        total += 1
        if not ok:
            bad += 1
            print('MISMATCH rep', rep, rows, cols, k, 'count', n, 'want', len(wb), 'first diff', (np.nonzero(idx[:min(n,len(widx))] != widx[:min(n,len(widx))])[0][:3]))
print('STRESS done: %d bad of %d' % (bad, total))
