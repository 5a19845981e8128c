import numpy as np, torch, sys
sys.path.insert(0,'.')
from vbx_b200 import synth
from vbx_b200.batch import VbxBatch
from oracle import c_oracle as co
S=4; lens=[4096]
d=synth.make_batch(lens,R=128,S=S,seed=3,dtype=np.float32)
dev=torch.device('cuda:0')
vb=VbxBatch(lens,128,S,device=dev)
g=torch.from_numpy(d['gamma0']).to(dev).contiguous(); p=torch.full((1,S),1.0/S,device=dev)
vb.prepare_scale(torch.from_numpy(d['fea']).to(dev), torch.from_numpy(d['Phi']).to(dev))
out=vb.run(g,p,Fa=0.3,Fb=17.0,loopProb=0.99,maxIters=1,epsilon=-float('inf'))
torch.cuda.synchronize()
ref=co.vbx_oracle_batch(d['fea'],d['Phi'],d['offsets'],d['gamma0'],np.full(S,1.0/S),0.3,17.0,0.99,1,-np.inf)
gg=g.double().cpu().numpy()
bad=np.nonzero(~np.isfinite(gg).all(1))[0]
print('nan rows:', len(bad), bad[:10], bad[-5:] if len(bad) else None)
err=np.abs(gg-ref['gamma']).max(1)
print('max err per chunk:', [float(np.nanmax(err[c*256:(c+1)*256])) for c in range(16)])
print('Li', out['Li'].cpu().numpy(), ref['Li'])
print('pi', p.cpu().numpy(), ref['pi'])
# peek into workspace? rows sums
print('rowsum first rows', gg[:3].sum(1), gg[255:258].sum(1))
