"""CPU oracle for the VB-HMM EM loop (TEST INFRASTRUCTURE - never on the product path).

This is a float64 numpy restatement of the algorithm implemented by the reference's
`VBx()` (/root/reference VBx/VBx.py:27-126) and `forward_backward()` (VBx/VBx.py:146-175).
It deliberately keeps the reference's *algorithmic structure* -- log-domain recursions with a
dense S x S log-sum-exp per frame executed from a Python loop over frames -- so that timing
it on host cores is representative of the reference's CPU cost (bench.py `cpu_baseline`,
kind "port").  Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s CPU-baseline /
`--impl reference` legs may import this module.

Parity pin: `tests/golden/*.npz` were produced by importing the *unmodified* reference in
the build container (tests/golden/make_golden.py); tests/test_oracle.py checks this module
against every one of them (ES2005a real recording + seeded synthetic cases).

Equation numbers refer to Landini et al., "Bayesian HMM clustering of x-vector sequences
(VBx) in speaker diarization" (cited at VBx/VBx.py:64-67).
"""
import math

import numpy as np

EPS_TR = 1e-8  # additive floor inside the logs of the HMM, VBx/VBx.py:158


def lse(a, axis):
    """log(sum(exp(a))) along `axis` with the usual max shift (stands in for
    scipy.special.logsumexp used at VBx/VBx.py:101,168,171,173; finite inputs only)."""
    top = np.max(a, axis=axis, keepdims=True)
    out = np.log(np.sum(np.exp(a - top), axis=axis, keepdims=True)) + top
    return np.squeeze(out, axis=axis)


def frame_constant(X):
    """Per-frame state-independent term of eq. (23): VBx/VBx.py:87."""
    R = X.shape[1]
    return -0.5 * (np.einsum('tr,tr->t', X, X) + R * math.log(2.0 * math.pi))


def speaker_model(gamma, rho, Phi, FaFb):
    """M-step, eqs (17) and (16): VBx/VBx.py:95-96.  Returns invL, alpha (S x R)."""
    occupancy = gamma.sum(axis=0)                       # N_s
    invL = 1.0 / (1.0 + FaFb * np.outer(occupancy, Phi))
    alpha = FaFb * invL * (gamma.T @ rho)
    return invL, alpha


def plda_loglik(rho, alpha, invL, Phi, G, Fa):
    """E-step observation log-likelihoods, eq. (23): VBx/VBx.py:97 (T x S)."""
    bias = 0.5 * ((invL + alpha * alpha) @ Phi)         # S
    return Fa * (rho @ alpha.T - bias[None, :] + G[:, None])


def hmm_forward_backward(ll, trans, init):
    """Log-domain forward-backward of an ergodic HMM: VBx/VBx.py:146-175.

    ll    T x S per-frame state log-likelihoods
    trans S x S transition probabilities (row = from, column = to)
    init  S     initial state probabilities
    Returns (posteriors T x S, total log-likelihood, log-forward, log-backward).
    """
    T, S = ll.shape
    ltr = np.log(trans + EPS_TR)                        # :159
    lf = np.full((T, S), -np.inf)
    lb = np.full((T, S), -np.inf)
    lf[0] = ll[0] + np.log(init + EPS_TR)               # :164
    lb[T - 1] = 0.0                                     # :165
    ltr_t = ltr.T.copy()
    for t in range(1, T):                               # :167-168
        lf[t] = ll[t] + lse(ltr_t + lf[t - 1][None, :], axis=1)
    for t in range(T - 2, -1, -1):                      # :170-171
        lb[t] = lse(ltr + (ll[t + 1] + lb[t + 1])[None, :], axis=1)
    total = lse(lf[T - 1], axis=0)                      # :173
    post = np.exp(lf + lb - total)                      # :174
    return post, float(total), lf, lb


def vbx_oracle(X, Phi, loopProb=0.9, Fa=1.0, Fb=1.0, pi=10, gamma=None, maxIters=10,
               epsilon=1e-4, alphaQInit=1.0, return_model=False, alpha=None, invL=None,
               rng=None):
    """Float64 restatement of the reference EM loop VBx/VBx.py:74-126 (same argument
    meaning, same `(gamma, pi, Li)` return; `ref`/`plot` diagnostics are not part of the
    oracle).  `rng` replaces the reference's global np.random draw (VBx/VBx.py:82) when
    gamma is None; pass nothing to use np.random exactly like the reference does."""
    X = np.asarray(X, dtype=np.float64)
    Phi = np.asarray(Phi, dtype=np.float64)
    T, R = X.shape
    if type(pi) is int:                                 # :76-77
        pi = np.full(pi, 1.0 / pi)
    pi = np.asarray(pi, dtype=np.float64)
    S = len(pi)
    if gamma is None:                                   # :79-83
        draw = (rng if rng is not None else np.random).gamma(alphaQInit, size=(T, S))
        gamma = draw / draw.sum(axis=1, keepdims=True)
    gamma = np.asarray(gamma, dtype=np.float64)
    assert gamma.shape == (T, S)                        # :85

    G = frame_constant(X)                               # :87
    rho = X * np.sqrt(Phi)[None, :]                     # :88-89, eq. (18)
    FaFb = Fa / Fb
    Li = []
    for it in range(maxIters):
        if it > 0 or alpha is None or invL is None:     # :94 warm start
            invL, alpha = speaker_model(gamma, rho, Phi, FaFb)
        ll = plda_loglik(rho, alpha, invL, Phi, G, Fa)
        trans = loopProb * np.eye(S) + (1.0 - loopProb) * pi[None, :]     # :98, eq. (1)
        gamma, tll, lf, lb = hmm_forward_backward(ll, trans, pi)          # :99
        elbo = tll + 0.5 * Fb * np.sum(np.log(invL) - invL - alpha * alpha + 1.0)   # :100, eq. (25)
        # eq. (24), :101-104 -- expected number of (re)entries into each speaker
        enter = np.exp(lse(lf[:-1], axis=1)[:, None] + ll[1:] + lb[1:] - tll).sum(axis=0)
        pi = gamma[0] + (1.0 - loopProb) * pi * enter
        pi = pi / pi.sum()
        Li.append([float(elbo)])                        # :105
        if it > 0 and elbo - Li[-2][0] < epsilon:       # :122-125
            break
    out = (gamma, pi, Li)
    if return_model:
        out = out + (alpha, invL)
    return out
