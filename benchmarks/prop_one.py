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
