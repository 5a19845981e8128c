"""Index-level emulation of the mma.sync fragment layouts used by the in-loop contractions (vbx_mma_kernels.cu;
DESIGN.md section 5.3).  The kernels permute the k / n indices of the m16n8k8 fragments so that every thread reads
contiguous floats of a rho row (coalesced, no shared-memory staging); this test rebuilds the per-lane fragments with
the kernels' own index formulas, applies the fragment semantics of `mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32`
(PTX ISA: A a0..a3 = (g, q), (g+8, q), (g, q+4), (g+8, q+4); B b0, b1 = (k=q, n=g), (k=q+4, n=g); C c0..c3 = (g, 2q),
(g, 2q+1), (g+8, 2q), (g+8, 2q+1) with g = lane / 4, q = lane % 4) and checks that the result is the plain matrix
product, every output written exactly once.  Host-side guard of the layout algebra, no arithmetic subtleties."""
import numpy as np
import pytest

LANES = np.arange(32)
G, Q = LANES >> 2, LANES & 3


def mma_m16n8k8(a_frag, b_frag):
    """a_frag [32,4], b_frag [32,2] per-lane registers -> the 16x8 product A.B as a matrix."""
    A = np.zeros((16, 8))
    B = np.zeros((8, 8))
    for lane in LANES:
        g, q = G[lane], Q[lane]
        A[g, q], A[g + 8, q], A[g, q + 4], A[g + 8, q + 4] = a_frag[lane]
        B[q, g], B[q + 4, g] = b_frag[lane]
    return A @ B


def c_fragment_positions():
    """(lane, e) -> (row, col) of the accumulator registers."""
    pos = {}
    for lane in LANES:
        g, q = G[lane], Q[lane]
        pos[(lane, 0)], pos[(lane, 1)], pos[(lane, 2)], pos[(lane, 3)] = (g, 2 * q), (g, 2 * q + 1), (g + 8, 2 * q), (g + 8, 2 * q + 1)
    return pos


def test_alpha_fragment_column_permutation_is_a_bijection():
    """speaker_model_kernel writes Fa*alpha fragment-major: element q = ((i*KS + j)*32 + lane)*2 + e  <->  state
    8i + lane/4, column 16*(j/2) + 4*(lane%4) + 2*(j%2) + e (R = 128) or KQ*(lane%4) + 2j + e (other R)."""
    for R in (128, 64, 32, 100):
        KS = 16 if R == 128 else (R + 7) >> 3
        KQ = 2 * KS
        cols = [16 * (j >> 1) + 4 * fq + 2 * (j & 1) + e if R == 128 else KQ * fq + 2 * j + e
                for j in range(KS) for fq in range(4) for e in range(2)]
        assert sorted(cols) == list(range(8 * KS))                  # every column of the (padded) row exactly once
        assert 8 * KS >= R


@pytest.mark.parametrize('S', [8, 16, 32, 64])
def test_loglik_fragments_give_rho_times_alpha_transposed(S):
    """loglik_mma_kernel, R = 128: D[frame][state] for one 16-frame m-tile, all n-tiles."""
    rng = np.random.default_rng(S)
    R, KS, NT = 128, 16, S // 8
    rho = rng.standard_normal((16, R))
    alpha = rng.standard_normal((S, R))
    # fragment-major alpha as written by the speaker-model kernel
    frag = np.zeros(NT * KS * 64)
    for qi in range(NT * KS * 64):
        e, ln, ij = qi & 1, (qi >> 1) & 31, qi >> 6
        j, i = ij & 15, ij >> 4
        st, fq = 8 * i + (ln >> 2), ln & 3
        frag[qi] = alpha[st, 16 * (j >> 1) + 4 * fq + 2 * (j & 1) + e]
    frag = frag.reshape(NT, KS, 32, 2)
    D = np.zeros((NT, 16, 8))
    for j in range(KS):
        k = j >> 1
        a = np.zeros((32, 4))
        for lane in LANES:
            g, q = G[lane], Q[lane]
            xa, xb = rho[g, 16 * k + 4 * q:16 * k + 4 * q + 4], rho[g + 8, 16 * k + 4 * q:16 * k + 4 * q + 4]   # float4 number k of the thread
            a[lane] = (xa[0], xb[0], xa[1], xb[1]) if j % 2 == 0 else (xa[2], xb[2], xa[3], xb[3])
        for i in range(NT):
            D[i] += mma_m16n8k8(a, frag[i, j])
    full = np.concatenate([D[i] for i in range(NT)], axis=1)           # state 8i + n
    np.testing.assert_allclose(full, rho @ alpha.T, rtol=1e-12, atol=1e-12)
    # finish(): lane (g, q) holds states 8i + 2q, 8i + 2q + 1 of rows g and g + 8
    pos = c_fragment_positions()
    for lane in LANES:
        assert pos[(lane, 0)] == (G[lane], 2 * Q[lane]) and pos[(lane, 3)] == (G[lane] + 8, 2 * Q[lane] + 1)


@pytest.mark.parametrize('S', [16, 32, 64])
def test_mstep_fragments_give_gamma_transposed_times_rho(S):
    """mstep_mma_kernel: one 8-frame chunk, one warp's column range of RW = 32*NQ columns; A = gamma^T (states x frames),
    B = rho with n-tile j = 4k + e holding columns col0 + 32k + e, col0 = 4g; outputs parked at
    r0 = 32*(j/4) + 8q + (j%4), r1 = r0 + 4."""
    rng = np.random.default_rng(S)
    MT = S // 16
    NQ = 4 // MT                                                      # float4 loads per thread and frame (kernel: NTW / 4, NTW = 16 / MT)
    RW = 32 * NQ
    gamma = rng.random((8, S))
    rho = rng.standard_normal((8, RW))
    NTW = 4 * NQ
    out = np.full((S, RW), np.nan)
    written = np.zeros((S, RW), int)
    pos = c_fragment_positions()
    for m in range(MT):
        a = np.zeros((32, 4))
        for lane in LANES:
            g, q = G[lane], Q[lane]
            a[lane] = (gamma[q, 16 * m + g], gamma[q, 16 * m + g + 8], gamma[q + 4, 16 * m + g], gamma[q + 4, 16 * m + g + 8])
        for j in range(NTW):
            k, e = j >> 2, j & 3
            b = np.zeros((32, 2))
            for lane in LANES:
                g, q = G[lane], Q[lane]
                b[lane] = (rho[q, 4 * g + 32 * k + e], rho[q + 4, 4 * g + 32 * k + e])
            C = mma_m16n8k8(a, b)
            for lane in LANES:
                g, q = G[lane], Q[lane]
                r0 = 32 * (j >> 2) + 8 * q + (j & 3)
                r1 = r0 + 4
                for reg, (srow, col) in enumerate([(16 * m + g, r0), (16 * m + g, r1), (16 * m + g + 8, r0), (16 * m + g + 8, r1)]):
                    out[srow, col] = C[pos[(lane, reg)]]
                    written[srow, col] += 1
    assert np.all(written == 1)                                        # every (state, column) exactly once
    np.testing.assert_allclose(out, gamma.T @ rho, rtol=1e-12, atol=1e-12)
