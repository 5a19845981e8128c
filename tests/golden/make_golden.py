#!/usr/bin/env python
"""Generate the golden vectors under tests/golden/ by running the UNMODIFIED reference.

Run in the build container only (needs /root/reference):  python tests/golden/make_golden.py
The reference `VBx.VBx` / `forward_backward` (VBx/VBx.py) and the helpers of
VBx/diarization_lib.py are imported, never copied.  fastcluster is missing in this image, so the
AHC step of VBx/vbhmm.py:135-146 uses scipy's average linkage (same algorithm, same dendrogram).
"""
import hashlib
import os
import sys

import numpy as np
from scipy.cluster.hierarchy import fcluster, linkage
from scipy.linalg import eigh
from scipy.spatial.distance import squareform
from scipy.special import softmax

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get('VBX_REF', '/root/reference')
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(REF, 'VBx'))

from VBx import VBx as ref_VBx, forward_backward as ref_fb, DER as ref_DER          # noqa: E402  (the reference)
import diarization_lib as ref_dl                                     # noqa: E402  (the reference)
from vbx_b200 import formats, synth                                  # noqa: E402


def sha(path):
    return hashlib.sha256(open(path, 'rb').read()).hexdigest()[:16]


def es2005a():
    """Follows VBx/vbhmm.py:105-162 with run_example.sh:30-34 hyper-parameters."""
    ark = os.path.join(REF, 'exp/ES2005a.ark')
    seg = os.path.join(REF, 'exp/ES2005a.seg')
    plda_f = os.path.join(REF, 'VBx/models/ResNet101_16kHz/plda')
    tr_f = os.path.join(REF, 'VBx/models/ResNet101_16kHz/transform.h5')
    Fa, Fb, loopP, thr_bias, lda_dim, smoothing = 0.3, 17.0, 0.99, -0.015, 128, 5.0

    plda_mu, plda_tr, plda_psi = formats.read_kaldi_plda(plda_f)
    W = np.linalg.inv(plda_tr.T.dot(plda_tr))
    B = np.linalg.inv((plda_tr.T / plda_psi).dot(plda_tr))
    acvar, wccn = eigh(B, W)
    plda_psi = acvar[::-1]
    plda_tr = wccn.T[::-1]

    recs = formats.read_xvectors_by_recording(ark)
    (name, (keys, x_raw)), = recs.items()
    mean1, mean2, lda = formats.read_xvec_transform(tr_f)
    x = ref_dl.l2_norm(lda.T.dot(ref_dl.l2_norm(x_raw - mean1).T).T - mean2)

    scr = ref_dl.cos_similarity(x)
    thr, _ = ref_dl.twoGMMcalib_lin(scr.ravel())
    lin = linkage(squareform(-scr, checks=False), method='average')
    adjust = abs(lin[:, 2].min())
    lin[:, 2] += adjust
    labels_ahc = fcluster(lin, -(thr + thr_bias) + adjust, criterion='distance') - 1

    qinit = np.zeros((len(labels_ahc), labels_ahc.max() + 1))
    qinit[range(len(labels_ahc)), labels_ahc] = 1.0
    qinit = softmax(qinit * smoothing, axis=1)
    fea = (x - plda_mu).dot(plda_tr.T)[:, :lda_dim]
    Phi = plda_psi[:lda_dim]
    q, sp, L = ref_VBx(fea, Phi, pi=qinit.shape[1], gamma=qinit, maxIters=40, epsilon=1e-6,
                       loopProb=loopP, Fa=Fa, Fb=Fb)
    labels = np.argsort(-q, axis=1)[:, 0]
    segs = formats.read_segments(seg)[name]
    assert list(segs[0]) == keys
    starts, ends, out_labels = ref_dl.merge_adjacent_labels(segs[1][:, 0].copy(), segs[1][:, 1].copy(), labels)

    # the committed system output of run_example.sh must be reproduced (up to speaker renaming)
    rttm = formats.read_rttm(os.path.join(REF, 'exp/ES2005a.rttm'))
    assert len(rttm) == len(starts), (len(rttm), len(starts))
    mapping = {}
    for (rec, s, d, lab), s2, e2, l2 in zip(rttm, starts, ends, out_labels):
        assert abs(s - s2) < 1e-5 and abs(d - (e2 - s2)) < 1e-5
        assert mapping.setdefault(int(l2), lab) == lab
    assert len(set(mapping.values())) == len(mapping)
    print('ES2005a: T=%d S=%d iters=%d final ELBO=%.8f; RTTM reproduced (%d segments, map %s)'
          % (fea.shape[0], qinit.shape[1], len(L), L[-1][0], len(starts), mapping))

    np.savez_compressed(
        os.path.join(HERE, 'es2005a.npz'),
        # inputs of the VBx() call at VBx/vbhmm.py:154-158
        fea=fea, Phi=Phi, labels_ahc=labels_ahc.astype(np.int16), smoothing=smoothing,
        Fa=Fa, Fb=Fb, loopProb=loopP, maxIters=40, epsilon=1e-6,
        # outputs
        gamma=q, pi=sp, Li=np.array([l[0] for l in L]),
        # what follows the call: hard labels and merged segments
        labels=labels.astype(np.int16), seg_times=segs[1], rttm_starts=starts, rttm_ends=ends,
        rttm_labels=out_labels.astype(np.int16),
        rttm_ref_labels=np.array([int(r[3]) for r in rttm], dtype=np.int16),
        # the raw inputs of the steps before the call (for the section-8f "next" rows)
        x_raw=x_raw.astype(np.float32), x_lda=x,
        sha_ark=sha(ark), sha_plda=sha(plda_f), sha_transform=sha(tr_f))


def synthetic_cases():
    """Seeded synthetic cases through the reference, covering the recipes' hyper-parameters
    (AMI_run.sh:44-49, DIHARD2_run.sh:42-47, CALLHOME_run.sh:42-47, run_example.sh:30-34) and the
    edge cases of SURVEY.md App. A (T=1, S=1, dead speaker, warm start, vector pi, early stop)."""
    cases = {}

    def run(tag, T, R, S, seed, Fa, Fb, loopP, maxIters=10, epsilon=-np.inf, pi=None,
            warm=False, kill=None):
        rng = np.random.default_rng(seed)
        Phi = synth.plda_phi(R)
        fea, _ = synth.make_recording(T, R, Phi, rng, n_spk=min(4, max(1, S)))
        g0 = synth.dirichlet_rows(T, S, rng)
        pi_in = S if pi is None else np.asarray(pi, dtype=np.float64)
        if kill is not None:                    # a speaker with zero prior and zero mass
            g0[:, kill] = 0.0
            g0 /= g0.sum(1, keepdims=True)
            pi_in = np.full(S, 1.0 / (S - 1)); pi_in[kill] = 0.0
        kw = dict(loopProb=loopP, Fa=Fa, Fb=Fb, pi=pi_in, gamma=g0.copy(), maxIters=maxIters,
                  epsilon=epsilon)
        extra = {}
        if warm:
            _, _, _, a0, l0 = ref_VBx(fea, Phi, return_model=True, **{**kw, 'maxIters': 2})
            kw.update(alpha=a0, invL=l0)
            extra = {'alpha0': a0, 'invL0': l0}
        g, p, L, a, iL = ref_VBx(fea, Phi, return_model=True, **kw)
        cases[tag] = dict(fea=fea, Phi=Phi, gamma0=g0, pi0=(np.full(S, 1.0 / S) if isinstance(pi_in, int) else np.asarray(pi_in, dtype=np.float64)),
                          pi_is_int=pi is None and kill is None,
                          Fa=Fa, Fb=Fb, loopProb=loopP, maxIters=maxIters, epsilon=epsilon,
                          gamma=g, pi=p, Li=np.array([l[0] for l in L]), alpha=a, invL=iL, **extra)
        print('%-22s T=%-5d R=%-4d S=%-3d iters=%-3d ELBO=%.6f' % (tag, T, R, S, len(L), L[-1][0]))

    run('example_hp', 300, 128, 16, 1, 0.3, 17.0, 0.99)
    run('ami_hp', 257, 128, 12, 2, 0.4, 64.0, 0.65)
    run('dihard_hp', 400, 128, 30, 3, 0.2, 6.0, 0.35, maxIters=12)
    run('callhome_hp', 150, 128, 7, 4, 0.4, 17.0, 0.40)
    run('small_r16', 50, 16, 4, 5, 1.0, 1.0, 0.9)
    run('s64', 200, 128, 64, 6, 0.3, 17.0, 0.99, maxIters=6)
    run('t1', 1, 128, 5, 7, 0.3, 17.0, 0.99, maxIters=3)
    run('s1', 40, 128, 1, 8, 0.3, 17.0, 0.99, maxIters=3)
    run('t2', 2, 32, 3, 9, 0.3, 17.0, 0.9, maxIters=3)
    run('dead_speaker', 120, 128, 6, 10, 0.3, 17.0, 0.99, kill=2)
    run('vector_pi', 90, 64, 5, 11, 0.4, 17.0, 0.8, pi=[0.5, 0.2, 0.15, 0.1, 0.05])
    run('warm_start', 160, 128, 8, 12, 0.3, 17.0, 0.99, warm=True, maxIters=5)
    run('early_stop', 300, 128, 16, 13, 0.3, 17.0, 0.99, maxIters=40, epsilon=1e-3)
    run('loop0', 80, 32, 4, 14, 0.5, 5.0, 0.0, maxIters=5)
    run('loop1', 80, 32, 4, 15, 0.5, 5.0, 1.0, maxIters=5)
    flat = {}
    for tag, d in cases.items():
        for k, v in d.items():
            flat[f'{tag}/{k}'] = v
    np.savez_compressed(os.path.join(HERE, 'synthetic_cases.npz'), **flat)

    # forward_backward() alone (VBx/VBx.py:146-175) on random log-likelihoods with a wide dynamic range
    fb = {}
    for i, (T, S, scale) in enumerate([(64, 4, 1.0), (200, 16, 30.0), (77, 31, 100.0), (1, 3, 5.0)]):
        rng = np.random.default_rng(100 + i)
        lls = rng.standard_normal((T, S)) * scale
        ip = rng.dirichlet(np.ones(S))
        if S > 3:
            ip[1] = 0.0
            ip /= ip.sum()
        loopP = [0.9, 0.99, 0.35, 0.5][i]
        tr = np.eye(S) * loopP + (1 - loopP) * ip
        g, tll, lfw, lbw = ref_fb(lls, tr, ip)
        fb.update({f'fb{i}/lls': lls, f'fb{i}/ip': ip, f'fb{i}/loopProb': loopP, f'fb{i}/gamma': g,
                   f'fb{i}/tll': tll, f'fb{i}/lfw': lfw, f'fb{i}/lbw': lbw})
    np.savez_compressed(os.path.join(HERE, 'forward_backward_cases.npz'), **fb)


def es2005a_model():
    """The x-vector transform and the (diagonalised) PLDA of the shipped model, as the arrays VBx/vbhmm.py:125-143
    works with: the inputs of the real-data chain test (x_raw of es2005a.npz -> fea of es2005a.npz)."""
    plda_f = os.path.join(REF, 'VBx/models/ResNet101_16kHz/plda')
    tr_f = os.path.join(REF, 'VBx/models/ResNet101_16kHz/transform.h5')
    plda_mu, plda_tr, plda_psi = formats.read_kaldi_plda(plda_f)
    W = np.linalg.inv(plda_tr.T.dot(plda_tr))
    B = np.linalg.inv((plda_tr.T / plda_psi).dot(plda_tr))
    acvar, wccn = eigh(B, W)
    mean1, mean2, lda = formats.read_xvec_transform(tr_f)
    np.savez_compressed(os.path.join(HERE, 'es2005a_model.npz'), mean1=mean1, mean2=mean2, lda=lda, plda_mu=plda_mu,
                        plda_tr=wccn.T[::-1], plda_psi=acvar[::-1], sha_plda=sha(plda_f), sha_transform=sha(tr_f))


def ahc_cases():
    """AHC initialisation (VBx/vbhmm.py:131-146) with the reference's cos_similarity and twoGMMcalib_lin; scipy's
    average linkage stands in for fastcluster (same algorithm).  Case 0 is ES2005a, the others are synthetic."""
    es = np.load(os.path.join(HERE, 'es2005a.npz'))
    cases = {'es2005a': es['x_lda']}
    rng = np.random.default_rng(77)
    for name, T, spk, noise in (('syn_a', 150, 3, 0.6), ('syn_b', 333, 6, 0.9), ('syn_c', 40, 2, 0.4)):
        centres = rng.standard_normal((spk, 128))
        who = np.repeat(rng.integers(0, spk, T // 5 + 1), 5)[:T]
        cases[name] = ref_dl.l2_norm(centres[who] + noise * rng.standard_normal((T, 128)))
    out = {}
    for name, x in cases.items():
        scr = ref_dl.cos_similarity(x)
        thr, _ = ref_dl.twoGMMcalib_lin(scr.ravel())
        lin = linkage(squareform(-scr, checks=False), method='average')
        Z = lin.copy()
        adjust = abs(lin[:, 2].min())
        lin[:, 2] += adjust
        labels = fcluster(lin, -(thr - 0.015) + adjust, criterion='distance') - 1
        if name != 'es2005a':
            out[name + '/x'] = x
        else:
            assert np.array_equal(labels, es['labels_ahc'])
        out[name + '/thr'] = thr
        out[name + '/Z'] = Z
        out[name + '/labels'] = labels.astype(np.int32)
    np.savez_compressed(os.path.join(HERE, 'ahc_cases.npz'), **out)


def diagnostics_cases():
    """The optional parts of the reference module: DER() (VBx/VBx.py:129-143), the `ref=` trace columns of VBx()
    (VBx/VBx.py:107-109) and forward_backward() (VBx/VBx.py:146-175) with a GENERAL transition matrix."""
    out = {}
    rng = np.random.default_rng(2024)
    for i, (T, S, n_ref) in enumerate([(60, 4, 4), (200, 7, 5), (33, 3, 6)]):
        q = rng.dirichlet(np.ones(S) * 0.3, size=T)
        ref = rng.integers(0, n_ref, size=T)
        ref[0] = n_ref - 1                        # every reference speaker id up to the maximum exists
        out[f'der{i}/q'] = q
        out[f'der{i}/ref'] = ref.astype(np.int32)
        out[f'der{i}/values'] = np.array([ref_DER(q, ref), ref_DER(q, ref, xentropy=True),
                                          ref_DER(q, ref, expected=False), ref_DER(q, ref, expected=False, xentropy=True)])
    # VBx(ref=...) appends [ELBO, DER, cross-entropy] per iteration
    T, R, S = 180, 128, 6
    Phi = synth.plda_phi(R)
    fea, z = synth.make_recording(T, R, Phi, rng, stay=0.9, n_spk=4)
    fea = 0.12 * fea + rng.standard_normal(fea.shape)     # weakly separated speakers: DER and cross-entropy stay away from 0
    g0 = synth.dirichlet_rows(T, S, rng)
    g, p, L = ref_VBx(fea, Phi, loopProb=0.9, Fa=0.3, Fb=17.0, pi=S, gamma=g0.copy(), maxIters=8, epsilon=1e-3, ref=z)
    out.update({'trace/fea': fea, 'trace/Phi': Phi, 'trace/gamma0': g0, 'trace/ref': np.asarray(z, dtype=np.int32),
                'trace/Li': np.array(L), 'trace/gamma': g, 'trace/pi': p})
    print('ref= trace: %d iterations, last row %s' % (len(L), L[-1]))
    # forward_backward with dense random transition matrices
    for i, (T, S, scale) in enumerate([(50, 5, 3.0), (120, 30, 40.0), (1, 4, 2.0), (40, 70, 10.0)]):
        lls = rng.standard_normal((T, S)) * scale
        tr = rng.dirichlet(np.ones(S) * 0.5, size=S)
        ip = rng.dirichlet(np.ones(S))
        post, tll, lfw, lbw = ref_fb(lls, tr, ip)
        out.update({f'fbg{i}/lls': lls, f'fbg{i}/tr': tr, f'fbg{i}/ip': ip, f'fbg{i}/post': post, f'fbg{i}/tll': tll,
                    f'fbg{i}/lfw': lfw, f'fbg{i}/lbw': lbw})
    np.savez_compressed(os.path.join(HERE, 'diagnostics_cases.npz'), **out)


if __name__ == '__main__':
    np.random.seed(0)
    if sys.argv[1:] == ['diag']:
        diagnostics_cases()
        sys.exit(0)
    if sys.argv[1:] == ['model']:
        es2005a_model()
        sys.exit(0)
    if sys.argv[1:] == ['ahc']:
        ahc_cases()
        sys.exit(0)
    es2005a()
    es2005a_model()
    ahc_cases()
    synthetic_cases()
    diagnostics_cases()
