"""End-to-end batched VB-HMM from HOST buffers: pinned host x-vectors in, host responsibilities out.

This is the host-buffer call a batch user makes (bench.py `e2e`): the batch is cut into chunks of whole
recordings; chunk k+1 is copied host->device on one stream while chunk k runs the projection + EM loop on
another and chunk k-1 drains device->host, so PCIe transfers overlap the kernels.  Per recording the result is
the same as the reference's per-recording call VBx/vbhmm.py:153-158 (projection + VBx()).
"""
import numpy as np
import torch

from .batch import VbxBatch


class HostPipeline:
    def __init__(self, lengths, D, R, n_states, device=None, n_chunks=None, n_buffers=3):
        self.device = torch.device('cuda', torch.cuda.current_device()) if device is None else torch.device(device)
        lengths = np.asarray(lengths, dtype=np.int64)
        self.lengths, self.D, self.R, self.S_user = lengths, int(D), int(R), int(n_states)
        B, N = len(lengths), int(lengths.sum())
        if n_chunks is None:
            n_chunks = int(min(B, max(1, min(16, N // 200_000))))
        # contiguous chunks of whole recordings with ~equal frame counts
        bounds = [0]
        csum = np.cumsum(lengths)
        for c in range(1, n_chunks):
            bounds.append(int(np.searchsorted(csum, N * c / n_chunks, side='left')) + 1)
        bounds.append(B)
        bounds = sorted(set(min(max(b, 0), B) for b in bounds))
        self.rec_bounds = bounds
        self.n_chunks = len(bounds) - 1
        offs = np.concatenate([[0], csum])
        self.frame_bounds = [int(offs[b]) for b in bounds]
        max_frames = max(self.frame_bounds[i + 1] - self.frame_bounds[i] for i in range(self.n_chunks))
        max_recs = max(bounds[i + 1] - bounds[i] for i in range(self.n_chunks))
        self.plans = [VbxBatch(lengths[bounds[i]:bounds[i + 1]], R, n_states, device=self.device, allocate=False)
                      for i in range(self.n_chunks)]
        self.S = self.plans[0].S
        ws_bytes = max(p.workspace_bytes for p in self.plans)
        nb = min(n_buffers, self.n_chunks)
        dev = self.device
        self.bufs = [dict(X=torch.empty((max_frames, D), dtype=torch.float32, device=dev),
                          rho=torch.empty((max_frames, R), dtype=torch.float32, device=dev),
                          gamma=torch.zeros((max_frames, self.S), dtype=torch.float32, device=dev),
                          pi=torch.empty((max_recs, self.S), dtype=torch.float32, device=dev),
                          ws=torch.empty(ws_bytes, dtype=torch.uint8, device=dev),
                          done=torch.cuda.Event()) for _ in range(nb)]
        self.copy_in = torch.cuda.Stream(device=dev)
        self.copy_out = torch.cuda.Stream(device=dev)
        self.pi0 = torch.zeros(self.S, dtype=torch.float32, device=dev)
        self.pi0[:self.S_user] = 1.0 / self.S_user
        self.n_states_dev = None if self.S_user == self.S else torch.full((max_recs,), self.S_user, dtype=torch.int32, device=dev)
        # pinned host outputs
        self.gamma_h = torch.empty((N, self.S_user), dtype=torch.float32).pin_memory()
        self.pi_h = torch.empty((B, self.S_user), dtype=torch.float32).pin_memory()
        self.Li_h = None
        self.N, self.B = N, B
        self.h2d_bytes = N * D * 4 + N * self.S_user * 4
        self.d2h_bytes = N * self.S_user * 4 + B * self.S_user * 4

    def run(self, X_host, V, Phi, gamma0_host, Fa=1.0, Fb=1.0, loopProb=0.9, maxIters=10, epsilon=1e-4):
        """X_host [N,D], gamma0_host [N,S] pinned float32 host tensors; V [D,R], Phi [R] on the device.
        Returns dict(gamma [N,S], pi [B,S], Li [B,maxIters], n_iters [B]) as (pinned) host tensors."""
        dev = self.device
        main = torch.cuda.current_stream(dev)
        if self.Li_h is None or self.Li_h.shape[1] != maxIters:
            self.Li_h = torch.empty((self.B, maxIters), dtype=torch.float64).pin_memory()
            self.ni_h = torch.empty(self.B, dtype=torch.int32).pin_memory()
        S, Su = self.S, self.S_user
        outs = []
        for c in range(self.n_chunks):
            buf = self.bufs[c % len(self.bufs)]
            f0, f1 = self.frame_bounds[c], self.frame_bounds[c + 1]
            r0, r1 = self.rec_bounds[c], self.rec_bounds[c + 1]
            n, nb = f1 - f0, r1 - r0
            # H2D on the copy-in stream, after the previous user of this buffer set has drained
            self.copy_in.wait_event(buf['done'])
            with torch.cuda.stream(self.copy_in):
                buf['X'][:n].copy_(X_host[f0:f1], non_blocking=True)
                if Su == S:
                    buf['gamma'][:n].copy_(gamma0_host[f0:f1], non_blocking=True)
                else:
                    buf['gamma'][:n, :Su].copy_(gamma0_host[f0:f1], non_blocking=True)
                ready = torch.cuda.Event()
                ready.record(self.copy_in)
            main.wait_event(ready)
            vb = self.plans[c]
            vb.bind(buf['ws'])
            vb.n_states = None if self.n_states_dev is None else self.n_states_dev[:nb]
            vb.prepare_project(buf['X'][:n], V, Phi, out=buf['rho'][:n])
            pi = buf['pi'][:nb]
            pi.copy_(self.pi0.expand_as(pi))
            out = vb.run(buf['gamma'][:n], pi, Fa=Fa, Fb=Fb, loopProb=loopProb, maxIters=maxIters, epsilon=epsilon)
            computed = torch.cuda.Event()
            computed.record(main)
            self.copy_out.wait_event(computed)
            with torch.cuda.stream(self.copy_out):
                self.gamma_h[f0:f1].copy_(buf['gamma'][:n, :Su] if Su != S else buf['gamma'][:n], non_blocking=True)
                self.pi_h[r0:r1].copy_(pi[:, :Su], non_blocking=True)
                self.Li_h[r0:r1].copy_(out['Li'], non_blocking=True)
                self.ni_h[r0:r1].copy_(out['n_iters'], non_blocking=True)
                buf['done'].record(self.copy_out)
            outs.append(out)   # keep Li / n_iters tensors alive until the copies ran
        main.wait_stream(self.copy_out)
        self.copy_out.synchronize()
        return dict(gamma=self.gamma_h, pi=self.pi_h, Li=self.Li_h, n_iters=self.ni_h)
