#!/usr/bin/env python
"""Timing of the device AHC initialisation (vbx_ahc) against the CPU restatement (oracle/ahc_oracle.py, numpy + scipy
average linkage) on synthetic recordings.  Prints one JSON line per configuration.  Measurement tool, not a test."""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from vbx_b200.batch import VbxBatch          # noqa: E402
from vbx_b200 import ahc                     # noqa: E402
from oracle import ahc_oracle                # noqa: E402  (CPU baseline leg only)


def synth(T, spk, rng):
    centres = rng.standard_normal((spk, 128))
    who = np.repeat(rng.integers(0, spk, T // 5 + 1), 5)[:T]
    x = centres[who] + 0.8 * rng.standard_normal((T, 128))
    return x / np.linalg.norm(x, axis=1, keepdims=True)


def main():
    dev = torch.device('cuda:0')
    rng = np.random.default_rng(0)
    for n_rec, T in ((1, 1025), (64, 1000), (256, 1000), (1, 8000), (16, 4000)):
        xs = [synth(T, 6, rng) for _ in range(n_rec)]
        vb = VbxBatch([T] * n_rec, 128, 2, device=dev, allocate=False)
        x = torch.from_numpy(np.concatenate(xs)).to(dev).float().contiguous()
        ahc.ahc_batch(vb, x)
        torch.cuda.synchronize()
        t0 = time.time()
        labels, thr, Zs = ahc.ahc_batch(vb, x)
        torch.cuda.synchronize()
        t_gpu = time.time() - t0
        t0 = time.time()
        ref = ahc_oracle.ahc_labels(xs[0])
        t_cpu = time.time() - t0
        same = bool(np.array_equal(ref[0], labels[0]))
        print(json.dumps({'recordings': n_rec, 'T': T, 'gpu_s_batch': round(t_gpu, 4), 'gpu_ms_per_recording': round(1e3 * t_gpu / n_rec, 3),
                          'cpu_s_one_recording': round(t_cpu, 4), 'labels_equal_recording0': same}), flush=True)
        vb.close()


if __name__ == '__main__':
    main()
