#!/usr/bin/env python
"""Drive the real-data front end once on a synthetic batch (for ncu launch lists): x-vector transform + PLDA projection
(vbx_prepare_xvectors), AHC initialisation (vbx_ahc), 10 EM iterations, labels (vbx_hard_labels)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from vbx_b200.batch import VbxBatch          # noqa: E402
from vbx_b200 import ahc, pipeline           # noqa: E402


def main():
    B, T = int(os.environ.get('B', 128)), int(os.environ.get('T', 1000))
    dev = torch.device('cuda:0')
    gen = torch.Generator(device='cpu').manual_seed(1)
    rnd = lambda *s: torch.randn(*s, generator=gen)
    spk = rnd(B, 6, 256)
    who = torch.randint(0, 6, (B, T // 5 + 1), generator=gen).repeat_interleave(5, dim=1)[:, :T]
    x_raw = (torch.gather(spk, 1, who[:, :, None].expand(B, T, 256)) + 0.7 * rnd(B, T, 256)).reshape(B * T, 256).to(dev).contiguous()
    q, _ = torch.linalg.qr(rnd(128, 128))
    model = [t.to(dev).contiguous() for t in (rnd(256) * 0.1, rnd(256, 128) / 16, rnd(128) * 0.05, rnd(128) * 0.02,
                                              q * (2.0 + 18.0 * torch.rand(128, generator=gen))[:, None],
                                              torch.linspace(8.0, 0.05, 128))]
    front = VbxBatch([T] * B, 128, 1, device=dev)
    rho, x_norm = front.prepare_xvectors(x_raw, *model)
    labels, thr, _ = ahc.ahc_batch(front, x_norm)
    front.close()
    ns = [int(l.max()) + 1 for l in labels]
    vb = VbxBatch([T] * B, 128, ns, device=dev)
    fea = rho / torch.sqrt(model[5])[None, :]
    vb.prepare_scale(fea.contiguous(), model[5])
    g = torch.zeros((B * T, vb.S), device=dev)
    lab = torch.from_numpy(np.concatenate(labels)).to(dev)
    for b in range(B):
        g[b * T:(b + 1) * T, :ns[b]] = pipeline.soft_init(lab[b * T:(b + 1) * T], ns[b], 7.0)
    p = torch.zeros((B, vb.S), device=dev)
    for b in range(B):
        p[b, :ns[b]] = 1.0 / ns[b]
    out = vb.run(g, p, Fa=0.3, Fb=17.0, loopProb=0.99, maxIters=10, epsilon=1e-6)
    first = vb.hard_labels(g)
    torch.cuda.synchronize()
    print('speakers per recording (first 8):', ns[:8], 'iterations:', out['n_iters'][:8].tolist(), 'labels:', first[:10].tolist())


if __name__ == '__main__':
    main()
