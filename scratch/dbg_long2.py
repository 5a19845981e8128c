import numpy as np, torch, sys
sys.path.insert(0,'.')
from vbx_b200 import synth
from vbx_b200.batch import VbxBatch
from oracle import c_oracle as co
dev=torch.device('cuda:0')
def run(lens,S,iters,ns=None,seed=3):
    d=synth.make_batch(lens,R=128,S=S,seed=seed,dtype=np.float32)
    vb=VbxBatch(lens,128,S if ns is None else ns,device=dev)
    Sp=vb.S
    g=torch.zeros((sum(lens),Sp),device=dev); g[:,:S]=torch.from_numpy(d['gamma0']).to(dev)
    p=torch.zeros((len(lens),Sp),device=dev); p[:,:S]=1.0/S
    vb.prepare_scale(torch.from_numpy(d['fea']).to(dev), torch.from_numpy(d['Phi']).to(dev))
    out=vb.run(g,p,Fa=0.3,Fb=17.0,loopProb=0.99,maxIters=iters,epsilon=-float('inf'))
    torch.cuda.synchronize()
    ref=co.vbx_oracle_batch(d['fea'],d['Phi'],d['offsets'],d['gamma0'],np.full(S,1.0/S),0.3,17.0,0.99,iters,-np.inf)
    gg=g[:,:S].double().cpu().numpy()
    Li=out['Li'].cpu().numpy()
    res=[]
    for b,(lo,hi) in enumerate(zip(d['offsets'][:-1],d['offsets'][1:])):
        e=np.abs(gg[lo:hi]-ref['gamma'][lo:hi]); bad=np.nonzero(~np.isfinite(gg[lo:hi]).all(1))[0]
        res.append((lens[b], float(np.nanmax(e)) if e.size else 0, len(bad), int(bad[0]) if len(bad) else -1, np.isfinite(Li[b]).tolist()))
    print(lens,S,iters,res)
run([4096],4,3)
run([4097],4,1)
run([5000],4,1)
run([4096,300],4,1)
run([4096],16,2)
run([4096],6,1)
run([4096],30,1)
