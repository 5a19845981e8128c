"""The optional members of the reference module next to VBx(): DER() (VBx/VBx.py:129-143), the `ref=` trace of VBx()
(VBx/VBx.py:107-109) and forward_backward() (VBx/VBx.py:146-175).  Goldens come from the unmodified reference
(tests/golden/make_golden.py diag)."""
import os

import numpy as np
import pytest

GOLD = os.path.join(os.path.dirname(__file__), 'golden')
Z = np.load(os.path.join(GOLD, 'diagnostics_cases.npz'))


@pytest.mark.parametrize('i', [0, 1, 2])
def test_der_matches_the_reference(i):
    from vbx_b200.api import DER
    q, ref, want = Z[f'der{i}/q'], Z[f'der{i}/ref'], Z[f'der{i}/values']
    got = [DER(q, ref), DER(q, ref, xentropy=True), DER(q, ref, expected=False), DER(q, ref, expected=False, xentropy=True)]
    np.testing.assert_allclose(got, want, rtol=1e-12, atol=1e-12)
    assert DER(q, list(ref)) == got[0]                       # ref may be any integer sequence
    with pytest.raises(ValueError):
        DER(q[:-1], ref)


@pytest.mark.gpu
@pytest.mark.parametrize('precision', ['float64', 'float32'])
def test_vbx_ref_trace_matches_the_reference(precision):
    """VBx(ref=...) reports [ELBO, DER, cross-entropy] per iteration (VBx/VBx.py:107-109)."""
    import vbx_b200.api as api
    api.set_precision(precision)
    try:
        g, p, L = api.VBx(Z['trace/fea'], Z['trace/Phi'], loopProb=0.9, Fa=0.3, Fb=17.0, pi=6, gamma=Z['trace/gamma0'],
                          maxIters=8, epsilon=1e-3, ref=Z['trace/ref'])
    finally:
        api.set_precision('float64')
    want = Z['trace/Li']
    assert len(L) == len(want) and all(len(row) == 3 for row in L)
    got = np.array(L)
    tol = 1e-9 if precision == 'float64' else 1e-4
    np.testing.assert_allclose(got[:, 0], want[:, 0], rtol=tol)
    np.testing.assert_allclose(got[:, 1:], want[:, 1:], rtol=max(tol, 1e-7) * 10, atol=tol)
    assert np.abs(g - Z['trace/gamma']).max() <= (1e-7 if precision == 'float64' else 1e-4)


@pytest.mark.gpu
@pytest.mark.parametrize('i', [0, 1, 2, 3])
def test_forward_backward_general_transition_matrix(i):
    from vbx_b200.dropin.VBx import forward_backward
    lls, tr, ip = Z[f'fbg{i}/lls'], Z[f'fbg{i}/tr'], Z[f'fbg{i}/ip']
    post, tll, lfw, lbw = forward_backward(lls, tr, ip)
    assert post.dtype == np.float64 and post.shape == lls.shape and isinstance(tll, float)
    np.testing.assert_allclose(tll, Z[f'fbg{i}/tll'], rtol=1e-12)
    np.testing.assert_allclose(lfw, Z[f'fbg{i}/lfw'], rtol=1e-11, atol=1e-9)
    np.testing.assert_allclose(lbw, Z[f'fbg{i}/lbw'], rtol=1e-11, atol=1e-9)
    np.testing.assert_allclose(post, Z[f'fbg{i}/post'], rtol=1e-8, atol=1e-12)


@pytest.mark.gpu
@pytest.mark.parametrize('i', [0, 1, 2, 3])
def test_forward_backward_on_the_em_loops_own_matrices(i):
    """The structured matrices of VBx/VBx.py:98 (the cases that pin the oracle)."""
    from vbx_b200.api import forward_backward
    z = np.load(os.path.join(GOLD, 'forward_backward_cases.npz'))
    lls, ip, lp = z[f'fb{i}/lls'], z[f'fb{i}/ip'], float(z[f'fb{i}/loopProb'])
    S = lls.shape[1]
    post, tll, lfw, lbw = forward_backward(lls, np.eye(S) * lp + (1 - lp) * ip, ip)
    np.testing.assert_allclose(tll, z[f'fb{i}/tll'], rtol=1e-12)
    np.testing.assert_allclose(lfw, z[f'fb{i}/lfw'], rtol=1e-11, atol=1e-9)
    np.testing.assert_allclose(lbw, z[f'fb{i}/lbw'], rtol=1e-11, atol=1e-9)
    np.testing.assert_allclose(post, z[f'fb{i}/gamma'], rtol=1e-8, atol=1e-12)
