"""Representative runs for compute-sanitizer (memcheck / racecheck / synccheck): tcgen05 projection, short and long
(chunked-scan) recordings, state counts 6..64, per-recording state masks.
    compute-sanitizer --tool memcheck --error-exitcode 3 python tools/sanitizer_cases.py
Round 1: 0 errors / 0 hazards with all three tools on a B200."""
import numpy as np, torch, sys
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vbx_b200 import synth
from vbx_b200.batch import VbxBatch
dev=torch.device('cuda:0')
def run(lens,S,iters,D=None,ns=None):
    d=synth.make_batch(lens,R=128,S=S,seed=3,D=D,dtype=np.float32)
    nsa=np.full(len(lens),S,dtype=np.int32) if ns is None else np.asarray(ns,dtype=np.int32)
    vb=VbxBatch(lens,128,nsa,device=dev)
    Sp=vb.S
    g=torch.zeros((sum(lens),Sp),device=dev); g[:,:S]=torch.from_numpy(d['gamma0']).to(dev)
    p=torch.zeros((len(lens),Sp),device=dev)
    for b in range(len(lens)): p[b,:nsa[b]]=1.0/nsa[b]
    if D: vb.prepare_project(torch.from_numpy(d['X']).to(dev), torch.from_numpy(d['V']).to(dev), torch.from_numpy(d['Phi']).to(dev))
    else: vb.prepare_scale(torch.from_numpy(d['fea']).to(dev), torch.from_numpy(d['Phi']).to(dev))
    out=vb.run(g,p,Fa=0.3,Fb=17.0,loopProb=0.99,maxIters=iters,epsilon=-float('inf'),return_model=True)
    torch.cuda.synchronize(); vb.close()
    print(lens[:4],S,'ok',float(out['Li'][0,-1]))
run([300,45,1,129,600],16,2,D=256)
run([37,700,2],6,2,ns=[6,3,5])
run([4100,300],8,2)
run([513,512,511],64,2)
run([100,200],31,2)
