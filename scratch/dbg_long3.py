import numpy as np, torch, sys
sys.path.insert(0,'.')
from vbx_b200 import synth
from vbx_b200.batch import VbxBatch
dev=torch.device('cuda:0')
S=4; lens=[4096]
d=synth.make_batch(lens,R=128,S=S,seed=3,dtype=np.float32)
vb=VbxBatch(lens,128,S,device=dev)
g=torch.from_numpy(d['gamma0']).to(dev).contiguous(); p=torch.full((1,S),1.0/S,device=dev)
vb.prepare_scale(torch.from_numpy(d['fea']).to(dev), torch.from_numpy(d['Phi']).to(dev))
for it in range(4):
    out=vb.run(g,p,Fa=0.3,Fb=17.0,loopProb=0.99,maxIters=1,epsilon=-float('inf'))
    torch.cuda.synchronize()
    gg=g.cpu().numpy(); bad=np.nonzero(~np.isfinite(gg).all(1))[0]
    ws=vb.workspace
    print('iter',it,'Li',out['Li'].cpu().numpy().ravel(),'pi',p.cpu().numpy().ravel(),'nan rows',len(bad), (bad[0],bad[-1]) if len(bad) else None)
    if len(bad):
        # find chunks with nan
        ch=sorted(set((bad//256).tolist())); print('chunks with nan', ch[:20])
        break
