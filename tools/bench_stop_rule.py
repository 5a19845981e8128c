#!/usr/bin/env python
"""The reference's way of calling the loop (maxIters=40, epsilon=1e-6, VBx/vbhmm.py:154-158) on a headline-shaped batch:
how long the float32 + float64-finish schedule takes, where the recordings stop, and the float64 share of the time."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench                                   # noqa: E402  (synthetic batch generator)
from vbx_b200.batch import VbxBatch            # noqa: E402


def main():
    B = int(os.environ.get('B', 4096))
    T, S = 1000, 16
    dev = torch.device('cuda:0')
    lens = np.full(B, T, dtype=np.int64)
    data = bench.make_device_batch(lens, S, seed=17, device=dev)
    vb = VbxBatch(lens, 128, S, device=dev)
    vb.set_option('timing', 1)
    rho = vb.prepare_project(data['X'], data['V'], data['Phi'])
    g = torch.empty((B * T, vb.S), device=dev)
    p = torch.empty((B, vb.S), device=dev)
    out = None
    res = {}
    for eps, iters in ((-float('inf'), 10), (1e-6, 40), (1e-4, 40)):
        for rep in range(2):
            g.copy_(data['gamma0'])
            p.fill_(1.0 / S)
            torch.cuda.synchronize()
            vb.timings(reset=True)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            out = vb.run(g, p, Fa=0.3, Fb=17.0, loopProb=0.99, maxIters=iters, epsilon=eps)
            e1.record()
            torch.cuda.synchronize()
        tm = vb.timings(reset=True)
        n = out['n_iters'].cpu().numpy()
        res[f'epsilon={eps}, maxIters={iters}'] = dict(
            ms=e0.elapsed_time(e1), iterations_mean=float(n.mean()), iterations_hist=np.bincount(n).tolist(),
            float64_ms=tm['exact64'][0], float32_ms=sum(v[0] for k, v in tm.items() if k not in ('exact64', 'project', 'prepare')),
            x_vectors_per_s=B * T / (e0.elapsed_time(e1) * 1e-3))
    print(json.dumps(dict(batch=f'{B} recordings x {T} x-vectors, S={S}', runs=res)))


if __name__ == '__main__':
    main()
