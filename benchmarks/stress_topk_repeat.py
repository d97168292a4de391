"""Race hunt for proposals_kernel: the same small problems launched thousands of times back to back (several cluster
sizes, wide and narrow CTAs, with and without other work in between); every result must equal the first one.
python benchmarks/stress_topk_repeat.py [repetitions]"""
import sys
import numpy as np
import torch
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
from faster_rcnn_b200 import ops, synth
from oracle import frcnn_oracle as O

dims = O.anchor_table([128, 256, 512])
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 300
bad = 0
total = 0
for (rows, cols, k, batch) in [(10, 12, 8000, 1), (5, 7, 8000, 1), (20, 20, 8000, 1), (20, 20, 1500, 3), (38, 63, 8000, 1),
                               (38, 63, 12000, 2), (16, 16, 2000, 40), (38, 63, 8000, 24), (30, 30, 5000, 1)]:
    pairs = [synth.rpn_outputs(rows, cols, len(dims), 11 + i, clustered=bool(i % 2)) for i in range(batch)]
    cls = torch.from_numpy(np.concatenate([p[0] for p in pairs])).cuda()
    regr = torch.from_numpy(np.concatenate([p[1] for p in pairs])).cuda()
    junk = torch.empty(64 << 20, dtype=torch.uint8, device='cuda')
    ref = [t.clone() for t in ops.decode_topk(regr, cls, dims, 16, k)]
    wb, wp, widx = O.topk_proposals(O.proposals_from_rpn(pairs[0][1].copy(), dims, 16), pairs[0][0].reshape(-1), k)
    n0 = int(ref[3][0])
    ok0 = n0 == len(wb) and np.array_equal(ref[2][0, :n0].cpu().numpy(), widx)
    if not ok0:
        print("FIRST RESULT differs from the oracle", rows, cols, k, batch)
        bad += 1
    for r in range(reps):
        if r % 3 == 1:
            junk.fill_(r & 255)                       # evict L2 / change timing
        if r % 7 == 3:
            torch.cuda.synchronize()
        got = ops.decode_topk(regr, cls, dims, 16, k)
        same = all(torch.equal(a, b) for a, b in zip(ref, got))
        total += 1
        if not same:
            bad += 1
            print("MISMATCH", rows, cols, k, batch, "rep", r, "count", got[3][:4].tolist(), "want", ref[3][:4].tolist())
print("REPEAT-STRESS done: %d bad of %d" % (bad, total))
