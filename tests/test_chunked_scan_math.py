"""The exact chunked scan for long recordings inside large batches (vbx_long_kernels.cu), restated
phase by phase in numpy float32 and checked against the reference's own forward_backward() outputs
(tests/golden/forward_backward_cases.npz, VBx/VBx.py:146-175) and the speaker-prior update of VBx/VBx.py:101-104:

    forward   A  per (chunk, basis e_i): the chunk's transfer operator as columns scale_i * u_i (renormalised every frame)
              B  per recording: entry vector of every chunk from the operators, sequentially over the chunks
              C  per chunk: re-run from the true entry vector, writing the normalised forward variables and 1/sigma_t
    backward  A  per (chunk, basis e_i): operator columns mu_i * v_i;  B  beta at the last frame of every chunk (exact scale)
              C  per chunk: re-run, gamma, per-chunk N_s and re-entry statistics;  tail: eq. (24)

Host-side guard of the algebra: cutting a recording into chunks changes float32 rounding only, for any chunk length."""
import os

import numpy as np
import pytest
from scipy.special import logsumexp

GOLD = os.path.join(os.path.dirname(__file__), 'golden', 'forward_backward_cases.npz')
f32, f64 = np.float32, np.float64
EPS = f32(1e-8)


def chunked_forward_backward(lls, ip, loopP, chunk):
    T, S = lls.shape
    P, Q = f32(loopP), f32(1.0 - loopP)
    m = lls.max(axis=1)
    p = np.exp(lls - m[:, None]).astype(f32)
    pi = ip.astype(f32)
    w = (Q * pi + EPS).astype(f32)                                    # VBx/VBx.py:98,159
    K = (T + chunk - 1) // chunk
    bounds = [(c * chunk, min(T, (c + 1) * chunk)) for c in range(K)]

    # ---------------- forward, phase A: operators of the chunks that have a successor ----------------
    def run_forward(t0, t1, base):
        lam, a = 0.0, None                                            # log of the chunk's scale in float64 (the kernel keeps mantissa * 2^exponent)
        for t in range(t0, t1):
            v = p[t] * base
            sig = v.sum(dtype=f32)
            a = v * (f32(1.0) / sig)
            lam += np.log(f64(sig))
            base = P * a + w
        return a, lam
    fa_u, fa_lam = {}, {}
    for c in range(K - 1):
        t0, t1 = bounds[c]
        if c == 0:
            fa_u[c], fa_lam[c] = [run_forward(t0, t1, pi + EPS)[0]], None      # VBx/VBx.py:164
        else:
            cols = [run_forward(t0, t1, P * np.eye(S, dtype=f32)[i] + w) for i in range(S)]
            fa_u[c], fa_lam[c] = [u for u, _ in cols], np.array([l for _, l in cols])
    # ---------------- forward, phase B: entry vector of every chunk ----------------
    astart = {}
    if K > 1:
        a = fa_u[0][0]
        astart[1] = a
        for c in range(1, K - 1):
            on = a > 0
            lmax = fa_lam[c][on].max()
            wgt = np.where(on, a.astype(f64) * np.exp(fa_lam[c] - lmax), 0.0).astype(f32)   # relative to the largest scale
            acc = np.zeros(S, f32)
            for i in range(S):
                if wgt[i] != 0:
                    acc = acc + wgt[i] * fa_u[c][i]
            a = acc * (f32(1.0) / acc.sum(dtype=f32))
            astart[c + 1] = a
    # ---------------- forward, phase C: re-run, per-frame outputs ----------------
    A, rsig = np.zeros((T, S), f32), np.zeros(T, f32)
    for c, (t0, t1) in enumerate(bounds):
        base = pi + EPS if c == 0 else P * astart[c] + w
        for t in range(t0, t1):
            v = p[t] * base
            r = f32(1.0) / v.sum(dtype=f32)
            A[t], rsig[t] = v * r, r
            base = P * A[t] + w

    # ---------------- backward, phase A: operator of chunk c maps b(t1-1) to b(t0-1) ----------------
    def run_backward(t0, t1, b):
        mu = 0.0                                                      # log of the scale, float64
        for fr in range(t1 - 1, t0 - 1, -1):                           # fr = frame t+1 of the step producing b(t)
            u = (p[fr] * rsig[fr]) * b
            bn = P * u + (w * u).sum(dtype=f32)
            tot = bn.sum(dtype=f32)
            if tot > 1e-30:
                b, mu = bn * (f32(1.0) / tot), mu + np.log(f64(tot))
            else:
                b, mu = bn * f32(0.0), -np.inf
        return b, mu
    bb_v, bb_mu = {}, {}
    for c in range(1, K):
        t0, t1 = bounds[c]
        if c == K - 1:
            v, mu = run_backward(t0, t1, np.ones(S, f32))
            bb_v[c], bb_mu[c] = [v], np.array([mu])
        else:
            cols = [run_backward(t0, t1, np.eye(S, dtype=f32)[i]) for i in range(S)]
            bb_v[c], bb_mu[c] = [v for v, _ in cols], np.array([mu_ for _, mu_ in cols])
    # ---------------- backward, phase B: beta at the last frame of every chunk (exact scale) ----------------
    beta = {K - 1: np.ones(S, f32)}
    if K > 1:
        b = (f32(np.exp(bb_mu[K - 1][0])) * bb_v[K - 1][0]).astype(f32)
        beta[K - 2] = b
        for c in range(K - 2, 0, -1):
            wgt = np.where(b > 0, b.astype(f64) * np.exp(bb_mu[c]), 0.0).astype(f32)   # bounded: beta stays in [1e-8, 1e8]
            acc = np.zeros(S, f32)
            for i in range(S):
                if wgt[i] != 0:
                    acc = acc + wgt[i] * bb_v[c][i]
            b = acc
            beta[c - 1] = b
    # ---------------- backward, phase C: gamma, per-chunk N_s and re-entry statistics ----------------
    gamma = np.zeros((T, S), f32)
    occ, ent = np.zeros(S, f64), np.zeros(S, f64)
    for c, (t0, t1) in enumerate(bounds):
        b = beta[c].copy()
        occ_c, ent_c = np.zeros(S, f32), np.zeros(S, f32)
        for t in range(t1 - 1, t0 - 1, -1):
            if t < t1 - 1:
                u = (p[t + 1] * rsig[t + 1]) * b
                ent_c = ent_c + u
                b = P * u + (w * u).sum(dtype=f32)
            g = A[t] * b
            g = g * (f32(1.0) / g.sum(dtype=f32))
            occ_c = occ_c + g
            gamma[t] = g
        if c > 0:
            ent_c = ent_c + (p[t0] * rsig[t0]) * b                     # the step across the chunk boundary
        occ, ent = occ + occ_c.astype(f64), ent + ent_c.astype(f64)     # tail kernel: chunk order, float64
    pn = gamma[0].astype(f64) + (1.0 - loopP) * pi.astype(f64) * ent   # VBx/VBx.py:101-104
    tll = float(-np.log(rsig.astype(f64)).sum() + m.sum())
    return gamma, tll, occ, pn / pn.sum()


def reference_pi_update(z, i):
    lls, ip, loopP = z[f'fb{i}/lls'], z[f'fb{i}/ip'], float(z[f'fb{i}/loopProb'])
    lfw, lbw, tll, gamma = z[f'fb{i}/lfw'], z[f'fb{i}/lbw'], float(z[f'fb{i}/tll']), z[f'fb{i}/gamma']
    pi = gamma[0] + (1.0 - loopP) * ip * np.sum(np.exp(logsumexp(lfw[:-1], axis=1, keepdims=True) + lls[1:] + lbw[1:] - tll), axis=0)
    return pi / pi.sum()


@pytest.mark.parametrize('i', range(4))
@pytest.mark.parametrize('chunk', [8, 16, 10 ** 9])
def test_chunked_scan_equals_reference_forward_backward(i, chunk):
    z = np.load(GOLD)
    lls, ip, loopP = z[f'fb{i}/lls'], z[f'fb{i}/ip'], float(z[f'fb{i}/loopProb'])
    gamma, tll, occ, pi_new = chunked_forward_backward(lls, ip, loopP, chunk)
    assert np.abs(gamma - z[f'fb{i}/gamma']).max() <= 3e-6
    assert abs(tll - float(z[f'fb{i}/tll'])) <= 2e-6 * max(1.0, abs(float(z[f'fb{i}/tll'])))
    assert np.abs(occ - z[f'fb{i}/gamma'].sum(axis=0)).max() <= 1e-5 * lls.shape[0]
    assert np.abs(pi_new - reference_pi_update(z, i)).max() <= 5e-6


def test_long_recording_with_the_kernels_chunk_length():
    """4 500 frames in chunks of 256 (kChunk) with sharp likelihoods and a dead speaker, against a float64 log-domain
    restatement of VBx/VBx.py:146-175 with the O(S) transition."""
    rng = np.random.default_rng(5)
    T, S, loopP = 4500, 6, 0.99
    lls = rng.standard_normal((T, S)) * 12.0
    ip = np.full(S, 1.0 / (S - 1))
    ip[2] = 0.0
    gamma, tll, occ, pi_new = chunked_forward_backward(lls, ip, loopP, 256)
    # float64 truth in the scaled linear domain (exactly the recursion the log-domain reference computes)
    mm = lls.max(axis=1)
    p = np.exp(lls - mm[:, None])
    w = (1.0 - loopP) * ip + 1e-8
    a = np.zeros((T, S)); sig = np.zeros(T)
    cur = p[0] * (ip + 1e-8)
    sig[0] = cur.sum(); a[0] = cur / sig[0]
    for t in range(1, T):
        cur = p[t] * (loopP * a[t - 1] + w)
        sig[t] = cur.sum(); a[t] = cur / sig[t]
    b = np.ones(S); ref = np.zeros((T, S)); ref[T - 1] = a[T - 1]; ent = np.zeros(S)
    for t in range(T - 2, -1, -1):
        u = p[t + 1] * b / sig[t + 1]
        ent += u
        b = loopP * u + (w * u).sum()
        ref[t] = a[t] * b
    ref_tll = np.log(sig).sum() + mm.sum()
    ref_pi = ref[0] + (1.0 - loopP) * ip * ent
    ref_pi /= ref_pi.sum()
    assert np.abs(gamma - ref).max() <= 2e-5
    assert abs(tll - ref_tll) <= 2e-6 * abs(ref_tll)
    assert np.abs(pi_new - ref_pi).max() <= 1e-5
    assert np.abs(gamma.sum(axis=1) - 1.0).max() <= 1e-5
