/*
 * CPU oracle #2 for the VB-HMM EM loop -- TEST INFRASTRUCTURE, never on the product path.
 *
 * Plain C, float64, one recording at a time (the Python binding threads over sub-batches).  It restates the algorithm of
 * the reference's VBx() (VBx/VBx.py:74-126) and forward_backward() (VBx/VBx.py:146-175), but in the
 * *scaled linear domain* and using the structure of the transition matrix
 *      A[i][j] = loopProb * (i==j) + (1-loopProb) * pi[j]              (VBx/VBx.py:98)
 * (diagonal + rank one, plus the reference's +1e-8 inside every log, VBx/VBx.py:158-159,164), so each
 * frame costs O(S) instead of an S x S log-sum-exp.  It exists because the numpy oracle
 * (oracle/vbx_oracle.py, pinned to the reference goldens) needs ~0.2 ms per frame per iteration and
 * cannot check full-size batches; tests/test_oracle.py pins THIS file to the numpy oracle and to the
 * reference goldens (agreement ~1e-10), so it inherits the pin.
 *
 * Build: see oracle/Makefile  (gcc -O2 -shared -fPIC).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define VBX_EPS 1e-8 /* VBx/VBx.py:158 */

/* One recording.  fea: T x R (row major).  gamma: T x S in/out.  pi: S in/out.
 * alpha/invL: S x R; if warm != 0 they are used as given for iteration 0 (VBx/VBx.py:94).
 * Li: maxIters doubles out.  Returns the number of iterations executed. */
static int vbx_one(const double *fea, const double *Phi, int64_t T, int R, int S, double *gamma, double *pi,
                   double Fa, double Fb, double loopP, int maxIters, double epsilon, double *alpha,
                   double *invL, int warm, double *Li) {
    const double FaFb = Fa / Fb;
    double *rho = malloc(sizeof(double) * (size_t)T * R);
    double *p = malloc(sizeof(double) * (size_t)T * S);    /* exp(ll - rowmax) */
    double *ah = malloc(sizeof(double) * (size_t)T * S);   /* normalised forward */
    double *sig = malloc(sizeof(double) * (size_t)T);      /* forward scales */
    double *bh = malloc(sizeof(double) * 2 * (size_t)S);   /* backward vector, ping-pong */
    double *w = malloc(sizeof(double) * (size_t)S);
    double *bias = malloc(sizeof(double) * (size_t)S);
    double *enter = malloc(sizeof(double) * (size_t)S);
    double *sqphi = malloc(sizeof(double) * (size_t)R);
    double Gsum = 0.0; /* sum_t G_t, VBx/VBx.py:87 -- G_t is state independent, so it only shifts tll */
    int iters = 0;

    for (int r = 0; r < R; ++r) sqphi[r] = sqrt(Phi[r]);      /* VBx/VBx.py:88 */
    for (int64_t t = 0; t < T; ++t) {
        double n2 = 0.0;
        for (int r = 0; r < R; ++r) {
            double x = fea[t * R + r];
            n2 += x * x;
            rho[t * R + r] = x * sqphi[r];                     /* VBx/VBx.py:89, eq. (18) */
        }
        Gsum += -0.5 * (n2 + R * log(2.0 * M_PI));
    }

    for (int it = 0; it < maxIters; ++it) {
        /* ---- M-step, VBx/VBx.py:95-96, eqs (17),(16) ---- */
        if (it > 0 || !warm) {
            for (int s = 0; s < S; ++s) {
                double Ns = 0.0;
                for (int64_t t = 0; t < T; ++t) Ns += gamma[t * S + s];
                for (int r = 0; r < R; ++r) invL[s * R + r] = 1.0 / (1.0 + FaFb * Ns * Phi[r]);
                for (int r = 0; r < R; ++r) alpha[s * R + r] = 0.0;
            }
            for (int64_t t = 0; t < T; ++t)
                for (int s = 0; s < S; ++s) {
                    double g = gamma[t * S + s];
                    if (g != 0.0)
                        for (int r = 0; r < R; ++r) alpha[s * R + r] += g * rho[t * R + r];
                }
            for (int s = 0; s < S; ++s)
                for (int r = 0; r < R; ++r) alpha[s * R + r] *= FaFb * invL[s * R + r];
        }
        /* ---- eq. (25) regulariser, VBx/VBx.py:100 ---- */
        double reg = 0.0;
        for (int s = 0; s < S; ++s) {
            double c = 0.0;
            for (int r = 0; r < R; ++r) {
                double iL = invL[s * R + r], a = alpha[s * R + r];
                c += (iL + a * a) * Phi[r];
                reg += log(iL) - iL - a * a + 1.0;
            }
            bias[s] = 0.5 * c;
        }
        /* ---- observation likelihoods, VBx/VBx.py:97 eq. (23), without the common G_t ---- */
        double msum = 0.0;
        for (int64_t t = 0; t < T; ++t) {
            double m = -INFINITY;
            for (int s = 0; s < S; ++s) {
                double d = 0.0;
                for (int r = 0; r < R; ++r) d += rho[t * R + r] * alpha[s * R + r];
                d = Fa * (d - bias[s]);
                p[t * S + s] = d;
                if (d > m) m = d;
            }
            for (int s = 0; s < S; ++s) p[t * S + s] = exp(p[t * S + s] - m);
            msum += m;
        }
        /* ---- forward, VBx/VBx.py:164,167-168 with A+eps = loopP*I + 1*w^T ---- */
        for (int s = 0; s < S; ++s) w[s] = (1.0 - loopP) * pi[s] + VBX_EPS;
        double lsig = 0.0;
        {
            double z = 0.0;
            for (int s = 0; s < S; ++s) { ah[s] = p[s] * (pi[s] + VBX_EPS); z += ah[s]; }
            for (int s = 0; s < S; ++s) ah[s] /= z;
            sig[0] = z; lsig += log(z);
        }
        for (int64_t t = 1; t < T; ++t) {
            const double *prev = ah + (t - 1) * S;
            double tot = 0.0, z = 0.0;
            for (int s = 0; s < S; ++s) tot += prev[s];   /* == 1 up to rounding; kept for exactness */
            for (int s = 0; s < S; ++s) {
                double v = p[t * S + s] * (loopP * prev[s] + w[s] * tot);
                ah[t * S + s] = v; z += v;
            }
            for (int s = 0; s < S; ++s) ah[t * S + s] /= z;
            sig[t] = z; lsig += log(z);
        }
        /* total log-likelihood, VBx/VBx.py:173 (+ the G_t we left out) */
        double tll = lsig + msum + Fa * Gsum;
        /* ---- backward + posteriors + eq. (24) statistics, VBx/VBx.py:165,170-171,174,101-103 ---- */
        double *b = bh, *bn = bh + S;
        for (int s = 0; s < S; ++s) { b[s] = 1.0; enter[s] = 0.0; }
        for (int s = 0; s < S; ++s) gamma[(T - 1) * S + s] = ah[(T - 1) * S + s];
        for (int64_t t = T - 2; t >= 0; --t) {
            double dotw = 0.0;
            for (int s = 0; s < S; ++s) {
                double u = p[(t + 1) * S + s] * b[s] / sig[t + 1];
                enter[s] += u;                             /* p*b/sigma at frame t+1 >= 1 */
                bn[s] = u; dotw += w[s] * u;
            }
            for (int s = 0; s < S; ++s) {
                bn[s] = loopP * bn[s] + dotw;
                gamma[t * S + s] = ah[t * S + s] * bn[s];
            }
            double *tmp = b; b = bn; bn = tmp;
        }
        /* ---- speaker priors, VBx/VBx.py:101-104 (no eps in this expression) ---- */
        double pz = 0.0;
        for (int s = 0; s < S; ++s) {
            pi[s] = gamma[s] + (1.0 - loopP) * pi[s] * enter[s];
            pz += pi[s];
        }
        for (int s = 0; s < S; ++s) pi[s] /= pz;
        Li[it] = tll + 0.5 * Fb * reg;                      /* VBx/VBx.py:100,105 */
        iters = it + 1;
        if (it > 0 && Li[it] - Li[it - 1] < epsilon) break; /* VBx/VBx.py:122-125 */
    }
    free(rho); free(p); free(ah); free(sig); free(bh); free(w); free(bias); free(enter); free(sqphi);
    return iters;
}

/* Packed ragged batch: recording b owns rows offsets[b] .. offsets[b+1]-1 of fea / gamma.
 * n_states[b] (or S for all when n_states == NULL) columns of the S-wide gamma/pi rows are live;
 * columns >= n_states[b] are ignored on input and zeroed on output.
 * alpha_io / invL_io: [B,S,R]; read for iteration 0 when warm != 0, always written. */
int vbx_oracle_batch(const double *fea, const double *Phi, const int64_t *offsets, int B, int R, int S,
                     const int32_t *n_states, double *gamma_io, double *pi_io, double Fa, double Fb,
                     double loopP, int maxIters, double epsilon, double *alpha_io, double *invL_io, int warm,
                     double *Li_out, int32_t *n_iters_out) {
    for (int b = 0; b < B; ++b) {
        const int64_t lo = offsets[b], T = offsets[b + 1] - offsets[b];
        const int Sb = n_states ? n_states[b] : S;
        double *Li = Li_out + (size_t)b * maxIters;
        for (int i = 0; i < maxIters; ++i) Li[i] = NAN;
        if (T <= 0 || Sb <= 0) { n_iters_out[b] = 0; continue; }
        double *g = malloc(sizeof(double) * (size_t)T * Sb);
        double *pi = malloc(sizeof(double) * (size_t)Sb);
        double *al = malloc(sizeof(double) * (size_t)Sb * R);
        double *il = malloc(sizeof(double) * (size_t)Sb * R);
        for (int64_t t = 0; t < T; ++t)
            for (int s = 0; s < Sb; ++s) g[t * Sb + s] = gamma_io[(lo + t) * S + s];
        for (int s = 0; s < Sb; ++s) pi[s] = pi_io[(size_t)b * S + s];
        if (warm)
            for (int s = 0; s < Sb; ++s)
                for (int r = 0; r < R; ++r) {
                    al[s * R + r] = alpha_io[((size_t)b * S + s) * R + r];
                    il[s * R + r] = invL_io[((size_t)b * S + s) * R + r];
                }
        n_iters_out[b] = vbx_one(fea + lo * R, Phi, T, R, Sb, g, pi, Fa, Fb, loopP, maxIters, epsilon, al, il,
                                 warm, Li);
        for (int64_t t = 0; t < T; ++t)
            for (int s = 0; s < S; ++s) gamma_io[(lo + t) * S + s] = s < Sb ? g[t * Sb + s] : 0.0;
        for (int s = 0; s < S; ++s) pi_io[(size_t)b * S + s] = s < Sb ? pi[s] : 0.0;
        if (alpha_io && invL_io)
            for (int s = 0; s < S; ++s)
                for (int r = 0; r < R; ++r) {
                    alpha_io[((size_t)b * S + s) * R + r] = s < Sb ? al[s * R + r] : 0.0;
                    invL_io[((size_t)b * S + s) * R + r] = s < Sb ? il[s * R + r] : 0.0;
                }
        free(g); free(pi); free(al); free(il);
    }
    return 0;
}
