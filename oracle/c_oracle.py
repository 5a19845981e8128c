"""ctypes binding of the C oracle (oracle/vbx_oracle_c.c) -- TEST INFRASTRUCTURE only."""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, 'libvbx_oracle.so')


def build(force=False):
    src = os.path.join(_HERE, 'vbx_oracle_c.c')
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(['make', '-C', _HERE, '-B' if force else '-s'], stdout=subprocess.DEVNULL)
    return _SO


_lib = None


def _load():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_SO)
        dp = ctypes.POINTER(ctypes.c_double)
        _lib.vbx_oracle_batch.restype = ctypes.c_int
        _lib.vbx_oracle_batch.argtypes = [
            dp, dp, ctypes.POINTER(ctypes.c_int64), ctypes.c_int, ctypes.c_int, ctypes.c_int,
            ctypes.POINTER(ctypes.c_int32), dp, dp, ctypes.c_double, ctypes.c_double, ctypes.c_double,
            ctypes.c_int, ctypes.c_double, dp, dp, ctypes.c_int, dp, ctypes.POINTER(ctypes.c_int32)]
    return _lib


def _dp(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))


def vbx_oracle_batch(fea, Phi, offsets, gamma0, pi0, Fa, Fb, loopProb, maxIters, epsilon,
                     n_states=None, alpha0=None, invL0=None):
    """Packed ragged batch through the C oracle.
    fea [N,R], gamma0 [N,S], pi0 [B,S] (or [S] broadcast) -> dict(gamma, pi, Li [B,maxIters] NaN padded,
    n_iters [B], alpha [B,S,R], invL [B,S,R])."""
    lib = _load()
    fea = np.ascontiguousarray(fea, dtype=np.float64)
    Phi = np.ascontiguousarray(Phi, dtype=np.float64)
    offsets = np.ascontiguousarray(offsets, dtype=np.int64)
    B = len(offsets) - 1
    N, R = fea.shape
    gamma = np.array(gamma0, dtype=np.float64, order='C', copy=True)
    S = gamma.shape[1]
    pi = np.array(np.broadcast_to(np.asarray(pi0, dtype=np.float64), (B, S)), order='C', copy=True)
    warm = alpha0 is not None and invL0 is not None
    alpha = np.array(alpha0, dtype=np.float64, order='C', copy=True).reshape(B, S, R) if warm \
        else np.zeros((B, S, R))
    invL = np.array(invL0, dtype=np.float64, order='C', copy=True).reshape(B, S, R) if warm \
        else np.zeros((B, S, R))
    Li = np.empty((B, maxIters), dtype=np.float64)
    n_iters = np.zeros(B, dtype=np.int32)
    ns = None
    if n_states is not None:
        ns = np.ascontiguousarray(n_states, dtype=np.int32)
    rc = lib.vbx_oracle_batch(
        _dp(fea), _dp(Phi), offsets.ctypes.data_as(ctypes.POINTER(ctypes.c_int64)), B, R, S,
        ns.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)) if ns is not None else None,
        _dp(gamma), _dp(pi), Fa, Fb, loopProb, maxIters, float(epsilon), _dp(alpha), _dp(invL),
        int(warm), _dp(Li), n_iters.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)))
    assert rc == 0
    return dict(gamma=gamma, pi=pi, Li=Li, n_iters=n_iters, alpha=alpha, invL=invL)
