"""Pins both oracles to the golden vectors produced by the unmodified reference
(tests/golden/make_golden.py).  CPU only."""
import os

import numpy as np
import pytest

from oracle import vbx_oracle as po
from oracle import c_oracle as co

GOLD = os.path.join(os.path.dirname(__file__), 'golden')


def load_cases():
    z = np.load(os.path.join(GOLD, 'synthetic_cases.npz'), allow_pickle=False)
    cases = {}
    for k in z.files:
        tag, name = k.split('/')
        cases.setdefault(tag, {})[name] = z[k]
    return cases


CASES = load_cases()


def pi_arg(c):
    return int(len(c['pi0'])) if bool(c['pi_is_int']) else c['pi0']


@pytest.mark.parametrize('tag', sorted(CASES))
def test_numpy_oracle_matches_reference(tag):
    c = CASES[tag]
    kw = {}
    if 'alpha0' in c:
        kw = dict(alpha=c['alpha0'], invL=c['invL0'])
    g, p, L, a, iL = po.vbx_oracle(c['fea'], c['Phi'], loopProb=float(c['loopProb']), Fa=float(c['Fa']),
                                   Fb=float(c['Fb']), pi=pi_arg(c), gamma=c['gamma0'],
                                   maxIters=int(c['maxIters']), epsilon=float(c['epsilon']),
                                   return_model=True, **kw)
    assert len(L) == len(c['Li'])
    np.testing.assert_allclose(g, c['gamma'], rtol=0, atol=1e-9)
    np.testing.assert_allclose(p, c['pi'], rtol=0, atol=1e-10)
    np.testing.assert_allclose([l[0] for l in L], c['Li'], rtol=1e-9)
    np.testing.assert_allclose(a, c['alpha'], rtol=0, atol=1e-9)
    np.testing.assert_allclose(iL, c['invL'], rtol=0, atol=1e-10)


@pytest.mark.parametrize('tag', sorted(CASES))
def test_c_oracle_matches_reference(tag):
    c = CASES[tag]
    T = c['fea'].shape[0]
    kw = {}
    if 'alpha0' in c:
        kw = dict(alpha0=c['alpha0'][None], invL0=c['invL0'][None])
    out = co.vbx_oracle_batch(c['fea'], c['Phi'], np.array([0, T]), c['gamma0'], c['pi0'][None, :],
                              float(c['Fa']), float(c['Fb']), float(c['loopProb']), int(c['maxIters']),
                              float(c['epsilon']), **kw)
    n = int(out['n_iters'][0])
    assert n == len(c['Li'])
    np.testing.assert_allclose(out['gamma'], c['gamma'], rtol=0, atol=2e-9)
    np.testing.assert_allclose(out['pi'][0], c['pi'], rtol=0, atol=1e-9)
    np.testing.assert_allclose(out['Li'][0, :n], c['Li'], rtol=1e-9)
    assert np.all(np.isnan(out['Li'][0, n:]))
    np.testing.assert_allclose(out['alpha'][0], c['alpha'], rtol=0, atol=1e-9)
    np.testing.assert_allclose(out['invL'][0], c['invL'], rtol=0, atol=1e-10)


def test_forward_backward_cases():
    z = np.load(os.path.join(GOLD, 'forward_backward_cases.npz'))
    n = len([k for k in z.files if k.endswith('/lls')])
    for i in range(n):
        lls, ip, P = z[f'fb{i}/lls'], z[f'fb{i}/ip'], float(z[f'fb{i}/loopProb'])
        S = len(ip)
        tr = np.eye(S) * P + (1 - P) * ip[None, :]
        g, tll, lf, lb = po.hmm_forward_backward(lls, tr, ip)
        np.testing.assert_allclose(g, z[f'fb{i}/gamma'], rtol=0, atol=1e-9)
        np.testing.assert_allclose(tll, float(z[f'fb{i}/tll']), rtol=1e-13)
        np.testing.assert_allclose(lf, z[f'fb{i}/lfw'], rtol=1e-12, atol=1e-10)
        np.testing.assert_allclose(lb, z[f'fb{i}/lbw'], rtol=1e-12, atol=1e-10)


def es2005a_inputs():
    z = np.load(os.path.join(GOLD, 'es2005a.npz'))
    lab = z['labels_ahc'].astype(int)
    q = np.zeros((len(lab), lab.max() + 1))
    q[np.arange(len(lab)), lab] = 1.0
    q = np.exp(q * float(z['smoothing']))
    q /= q.sum(1, keepdims=True)                       # softmax(onehot*smoothing), VBx/vbhmm.py:150-152
    return z, q


def test_es2005a_c_oracle():
    z, q = es2005a_inputs()
    T, S = q.shape
    out = co.vbx_oracle_batch(z['fea'], z['Phi'], np.array([0, T]), q, np.full((1, S), 1.0 / S),
                              float(z['Fa']), float(z['Fb']), float(z['loopProb']), int(z['maxIters']),
                              float(z['epsilon']))
    n = int(out['n_iters'][0])
    assert n == len(z['Li']) == 13
    np.testing.assert_allclose(out['Li'][0, :n], z['Li'], rtol=1e-12, atol=1e-7)
    np.testing.assert_allclose(out['gamma'], z['gamma'], rtol=0, atol=1e-8)
    np.testing.assert_allclose(out['pi'][0], z['pi'], rtol=0, atol=1e-9)
    assert np.array_equal(np.argmax(out['gamma'], axis=1), z['labels'])
    # known answers recorded in SURVEY.md section 8(c)
    assert abs(z['Li'][0] - (-76108.82806215)) < 1e-6 and abs(z['Li'][-1] - (-73432.62626820)) < 1e-6


def test_es2005a_numpy_oracle_three_iters():
    # the numpy oracle needs ~0.2 ms/frame/iter; 3 iterations keep the CPU suite fast
    z, q = es2005a_inputs()
    g, p, L = po.vbx_oracle(z['fea'], z['Phi'], loopProb=float(z['loopProb']), Fa=float(z['Fa']),
                            Fb=float(z['Fb']), pi=q.shape[1], gamma=q, maxIters=3, epsilon=-np.inf)
    np.testing.assert_allclose([l[0] for l in L], z['Li'][:3], rtol=1e-10)


def test_c_oracle_ragged_batch_and_state_mask():
    """Batch = independent recordings: each must equal its own single run; n_states masks columns."""
    rng = np.random.default_rng(5)
    from vbx_b200 import synth
    lens = [37, 1, 120, 64]
    d = synth.make_batch(lens, R=32, S=6, seed=3, dtype=np.float64)
    ns = np.array([6, 3, 4, 6], dtype=np.int32)
    g0 = d['gamma0'].copy()
    for b, (lo, hi) in enumerate(zip(d['offsets'][:-1], d['offsets'][1:])):
        g0[lo:hi, ns[b]:] = 0
        g0[lo:hi] /= g0[lo:hi].sum(1, keepdims=True)
    pi0 = np.zeros((4, 6))
    for b in range(4):
        pi0[b, :ns[b]] = 1.0 / ns[b]
    out = co.vbx_oracle_batch(d['fea'], d['Phi'], d['offsets'], g0, pi0, 0.3, 17.0, 0.9, 5, -np.inf, n_states=ns)
    for b, (lo, hi) in enumerate(zip(d['offsets'][:-1], d['offsets'][1:])):
        g, p, L = po.vbx_oracle(d['fea'][lo:hi], d['Phi'], loopProb=0.9, Fa=0.3, Fb=17.0, pi=pi0[b, :ns[b]],
                                gamma=g0[lo:hi, :ns[b]], maxIters=5, epsilon=-np.inf)
        np.testing.assert_allclose(out['gamma'][lo:hi, :ns[b]], g, atol=1e-9)
        assert np.all(out['gamma'][lo:hi, ns[b]:] == 0)
        np.testing.assert_allclose(out['pi'][b, :ns[b]], p, atol=1e-10)
        np.testing.assert_allclose(out['Li'][b], [l[0] for l in L], rtol=1e-12)
