"""Debugging aid: run one configuration of the split forward-backward with a synchronise + name after every launch."""
import sys, os, faulthandler
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from vbx_b200 import synth
from vbx_b200.batch import VbxBatch
faulthandler.dump_traceback_later(40, exit=True)
T, S, iters, fb = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
lens = [T]
d = synth.make_batch(lens, R=128, S=S, seed=3, dtype=np.float32)
dev = torch.device('cuda:0')
vb = VbxBatch(lens, 128, S, device=dev, fb_split=fb)
vb.workspace.fill_(0xFF)
vb.set_option('debug_sync', 1)
g = torch.zeros((T, vb.S), device=dev); g[:, :S] = torch.from_numpy(d['gamma0']).to(dev)
p = torch.zeros((1, vb.S), device=dev); p[0, :S] = 1.0 / S
vb.prepare_scale(torch.from_numpy(d['fea']).to(dev), torch.from_numpy(d['Phi']).to(dev))
out = vb.run(g, p, Fa=0.3, Fb=17.0, loopProb=0.99, maxIters=iters, epsilon=-float('inf'))
torch.cuda.synchronize()
print('done', T, S, out['Li'][0].cpu().numpy()[:3], float(g.sum()))
