"""The stop rule at float64 resolution (elbo_kernel in vbx_kernels.cu, vbx_exact64.cu), emulated
on the CPU: an EM iteration written once in numpy and run in float32 (standing in for the float32 kernels: same
quantities, float64 accumulation of the ELBO like the kernels) or float64 (standing in for the finishing kernels, gamma
stored in float32 between iterations like they do), driven by exactly the decision logic of the kernels:

    float32 iteration k:  d = ELBO_k - ELBO_{k-1},  nb = 2 * 2^-24 * |ELBO_k|
        d >= epsilon + 16 nb  -> keep going in float32
        d <  epsilon -  4 nb  -> stop
        otherwise             -> restore the state that ENTERED iteration k-1 (two-deep snapshot) and redo k-1, k, ...
                                 in float64; iteration k-1 is not tested again, iteration k and later ones are tested exactly

The emulation must stop at the iteration the float64 oracle stops at, for every epsilon - and a pure float32 loop must
not (otherwise the test would prove nothing).  Truth = oracle/vbx_oracle_c.c (float64, pinned to the reference)."""
import numpy as np
import pytest

from oracle import c_oracle
from vbx_b200 import synth

f32, f64 = np.float32, np.float64
NOISE_C, GUARD, SAFE_STOP = 2.0, 16.0, 4.0      # vbx_capi.cu: opt_noise_c, opt_guard_mult; elbo_kernel: 4 nb


def em_iteration(rho, gsum, Phi, gamma, pi, Fa, Fb, P, dt):
    """One EM iteration (VBx/VBx.py:95-104) with every per-frame quantity in `dt`; ELBO sums in float64.
    Scaled linear-domain forward-backward with the O(S) transition, as on the GPU.  -> gamma, pi, ELBO."""
    rho, Phi, gamma, pi = rho.astype(dt), Phi.astype(dt), gamma.astype(dt), pi.astype(dt)
    T, S = gamma.shape
    FaFb, Pd, Q = dt(Fa / Fb), dt(P), dt(1.0 - P)
    Ns = gamma.sum(axis=0, dtype=dt)
    invL = dt(1.0) / (dt(1.0) + FaFb * Ns[:, None] * Phi[None, :])
    alpha = FaFb * invL * (gamma.T @ rho)
    bias = dt(0.5) * ((invL + alpha * alpha) * Phi[None, :]).sum(axis=1, dtype=dt)
    ll = dt(Fa) * (rho @ alpha.T - bias[None, :])                      # without the state-independent G_t
    m = ll.max(axis=1)
    p = np.exp(ll - m[:, None]).astype(dt)
    w = (Q * pi + dt(1e-8)).astype(dt)                                # VBx/VBx.py:98,159
    a = np.empty((T, S), dt)
    sig = np.empty(T, dt)
    cur = p[0] * (pi + dt(1e-8))                                      # VBx/VBx.py:164
    sig[0] = cur.sum(dtype=dt)
    a[0] = cur / sig[0]
    for t in range(1, T):
        cur = p[t] * (Pd * a[t - 1] + w)                              # sum(a[t-1]) = 1
        sig[t] = cur.sum(dtype=dt)
        a[t] = cur / sig[t]
    b = np.ones(S, dt)
    g = np.empty((T, S), dt)
    g[T - 1] = a[T - 1]
    enter = np.zeros(S, f64)
    for t in range(T - 2, -1, -1):
        u = p[t + 1] * b / sig[t + 1]
        enter += u.astype(f64)                                        # sum_{t>=1} p_t b_t / sigma_t   (eq. 24)
        b = Pd * u + (w * u).sum(dtype=dt)
        gt = a[t] * b
        g[t] = gt / gt.sum(dtype=dt)
    tll = float(np.log(sig.astype(f64)).sum() + m.astype(f64).sum()) + Fa * gsum
    reg = (np.log(invL) - invL - alpha * alpha + dt(1.0)).astype(f64).sum()
    elbo = tll + 0.5 * Fb * float(reg)
    pn = g[0].astype(f64) + (1.0 - P) * pi.astype(f64) * enter
    return g, (pn / pn.sum()).astype(dt), elbo


def run_hybrid(rho, gsum, Phi, gamma0, pi0, Fa, Fb, P, max_iters, eps, exact_stop=True):
    """-> n_iters, Li, gamma, switched_at (None if the float32 phase decided everything)."""
    gamma, pi = gamma0.astype(f32), pi0.astype(f32)
    snaps = [None, None]
    Li, prev = [], None
    for k in range(max_iters):
        snaps[k & 1] = (gamma.copy(), pi.copy())                      # snapshot_kernel: state entering iteration k
        gamma, pi, elbo = em_iteration(rho, gsum, Phi, gamma, pi, Fa, Fb, P, f32)
        if k > 0:
            d = elbo - prev
            nb = NOISE_C * 2.0 ** -24 * abs(elbo)
            if exact_stop and not (d >= eps + GUARD * nb) and not (d < eps - SAFE_STOP * nb):
                # ---- float64 finish: redo iterations k-1 and k from the snapshot that entered k-1 ----
                gamma, pi = snaps[(k - 1) & 1]
                pi = pi.astype(f64)
                Li = Li[:k - 1]
                fresh = True
                for j in range(k - 1, max_iters):
                    g64, pi, e64 = em_iteration(rho, gsum, Phi, gamma, pi, Fa, Fb, P, f64)
                    gamma = g64.astype(f32)                           # gamma is stored in float32 between iterations
                    Li.append(e64)
                    if j > 0 and not fresh and e64 - Li[-2] < eps:
                        return j + 1, Li, gamma, k
                    fresh = False
                return max_iters, Li, gamma, k
            Li.append(elbo)
            if d < eps:
                return k + 1, Li, gamma, None
        else:
            Li.append(elbo)
        prev = elbo
    return max_iters, Li, gamma, None


def case(seed, T, R, S):
    rng = np.random.default_rng(seed)
    Phi = synth.plda_phi(R)
    fea, _ = synth.make_recording(T, R, Phi, rng, stay=0.98, n_spk=int(rng.integers(2, 5)))
    gamma0 = synth.dirichlet_rows(T, S, rng)
    G = -0.5 * ((fea * fea).sum(axis=1) + R * np.log(2.0 * np.pi))   # VBx/VBx.py:87
    rho = (fea * np.sqrt(Phi)[None, :]).astype(f32)                   # the kernels keep rho in float32
    return fea, Phi, rho, float(G.sum()), gamma0


@pytest.mark.parametrize('seed,eps', [(s, e) for s in range(6) for e in (1e-3, 1e-5, 1e-6)])
def test_hybrid_schedule_stops_where_the_float64_reference_stops(seed, eps):
    T, R, S, Fa, Fb, P, max_iters = 260 + 40 * seed, 32, 6, 0.3, 17.0, 0.99, 40
    fea, Phi, rho, gsum, gamma0 = case(seed, T, R, S)
    pi0 = np.full(S, 1.0 / S)
    # truth: float64 oracle on the float32-rounded inputs the GPU path sees
    fea32 = (rho.astype(f64) / np.sqrt(Phi)[None, :])
    ref = c_oracle.vbx_oracle_batch(fea32, Phi, np.array([0, T]), gamma0.astype(f32).astype(f64), pi0, Fa, Fb, P, max_iters, eps)
    n_ref = int(ref['n_iters'][0])
    # G of the float32-rounded features (the oracle derives it from its input)
    gsum32 = float((-0.5 * ((fea32 * fea32).sum(axis=1) + R * np.log(2.0 * np.pi))).sum())
    n, Li, gamma, switched = run_hybrid(rho, gsum32, Phi, gamma0, pi0, Fa, Fb, P, max_iters, eps)
    assert n == n_ref, (n, n_ref, switched)
    assert np.abs(gamma.astype(f64) - ref['gamma']).max() <= 1e-4
    li_ref = ref['Li'][0, :n_ref]
    assert np.abs(np.array(Li) - li_ref).max() <= 1e-4 * np.abs(li_ref).max()
    if switched is not None:
        # the float64 tail reproduces the reference's ELBO STEPS (1e-7 ... 1e-2 on |ELBO| ~ 5e3) to a relative 1e-3,
        # i.e. far below the float32 resolution of an ELBO value (~6e-4 here); measured: <= 1.1e-4
        steps, steps_ref = np.diff(np.array(Li)[switched - 1:]), np.diff(li_ref[switched - 1:])
        assert np.all(np.abs(steps - steps_ref) <= 1e-3 * np.abs(steps_ref) + 1e-13 * np.abs(li_ref).max())


def test_a_pure_float32_loop_does_not_follow_the_reference():
    """Control: without the float64 finish the same float32 iterations stop early for tight epsilons on at least some
    recordings - the failure the schedule exists to remove (round 1: iteration 6-10 instead of 13 on ES2005a)."""
    wrong = 0
    for seed in range(6):
        T, R, S, Fa, Fb, P, max_iters, eps = 260 + 40 * seed, 32, 6, 0.3, 17.0, 0.99, 40, 1e-6
        fea, Phi, rho, gsum, gamma0 = case(seed, T, R, S)
        pi0 = np.full(S, 1.0 / S)
        fea32 = (rho.astype(f64) / np.sqrt(Phi)[None, :])
        ref = c_oracle.vbx_oracle_batch(fea32, Phi, np.array([0, T]), gamma0.astype(f32).astype(f64), pi0, Fa, Fb, P, max_iters, eps)
        gsum32 = float((-0.5 * ((fea32 * fea32).sum(axis=1) + R * np.log(2.0 * np.pi))).sum())
        n, _, _, _ = run_hybrid(rho, gsum32, Phi, gamma0, pi0, Fa, Fb, P, max_iters, eps, exact_stop=False)
        wrong += int(n != int(ref['n_iters'][0]))
    assert wrong >= 3          # measured: 6 of 6 at epsilon = 1e-6 (10 of 18 over epsilon = 1e-3, 1e-5, 1e-6)
