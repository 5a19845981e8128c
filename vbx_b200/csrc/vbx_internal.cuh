// Internal declarations shared by the kernel translation units and the C ABI (not installed).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>

#define VBX_EPS_TR 1e-8f  // additive floor inside the HMM logs of the reference, VBx/VBx.py:158

namespace vbx {

constexpr int kLTile = 64;    // frames per CTA tile of the log-likelihood kernel
constexpr int kMTile = 512;   // frames per CTA tile of the M-step accumulation / mma log-likelihood kernels
constexpr int kMaxR = 128;
constexpr int kMaxS = 64;

// Device-resident description of a planned batch (arrays owned by the handle).
struct Plan {
    int32_t n_rec = 0, R = 0, S = 0;
    int64_t n_frames = 0;
    int32_t n_ltiles = 0, n_mtiles = 0;
    int64_t max_T = 0;
    const int64_t *offsets = nullptr;   // [n_rec+1]
    const int32_t *order = nullptr;     // [n_rec] recordings sorted by length, longest first
    const int32_t *ltile_rec = nullptr; // [n_ltiles]
    const int64_t *ltile_f0 = nullptr;  // [n_ltiles] first (global) frame of the tile
    const int32_t *mtile_rec = nullptr; // [n_mtiles]
    const int64_t *mtile_f0 = nullptr;  // [n_mtiles]
    const int32_t *mtile_begin = nullptr; // [n_rec+1] first M-tile of each recording
};

// Caller-provided workspace, carved by the handle.
struct Workspace {
    float *p = nullptr;        // [N,S]  exp(ll - rowmax)
    float *rowmax = nullptr;   // [N]
    float *rsigma = nullptr;   // [N]    1 / forward scale
    float *partial = nullptr;  // [n_mtiles,S,R] per-tile gamma^T rho
    float *A = nullptr;        // [n_rec,S,R]  Fa * alpha
    float *Afrag_hi = nullptr; // [n_rec,NT,KS,32] float2: Fa*alpha split to TF32 hi/lo, mma fragment-major
    float *Afrag_lo = nullptr; //   (NT = max(1,S/8) n-tiles, KS = ceil(R/8) k-steps; see vbx_mma_kernels.cu)
    float *bias = nullptr;     // [n_rec,S]    Fa * 0.5 * sum_r (invL + alpha^2) Phi_r ; +inf for dead columns
    float *occ = nullptr;      // [n_rec,S]    N_s = sum_t gamma
    double *reg = nullptr;     // [n_rec]      0.5 Fb sum (log invL - invL - alpha^2 + 1)
    double *gsum = nullptr;    // [n_rec]      sum_t G_t
    double *gpart = nullptr;   // [n_mtiles]
    double *prev_elbo = nullptr; // [n_rec]
    int32_t *active = nullptr; // [n_rec]
    float *scratch = nullptr;  // [2*kMaxS] write sink for warp lanes that own no recording
};

struct RunParams {
    float Fa, Fb, FaFb, loopP;
    double dFa, dFb, dFaFb, epsilon;
    int32_t max_iters;
};

// launchers (vbx_kernels.cu); each returns the number of kernels launched or -1 on launch error
int launch_prepare_scale(const Plan &pl, const Workspace &ws, const float *fea, const float *Phi, float *rho,
                         cudaStream_t st);
int launch_project_ffma(const Plan &pl, const float *X, int D, const float *V, float *rho, cudaStream_t st);
int launch_g_from_rho(const Plan &pl, const Workspace &ws, const float *rho, const float *Phi, cudaStream_t st);
int launch_run_init(const Plan &pl, const Workspace &ws, const float *gamma, const int32_t *n_states, double *Li,
                    int32_t *n_iters, int32_t *flags, int max_iters, cudaStream_t st);
int launch_mstep_partial(const Plan &pl, const Workspace &ws, const float *rho, const float *gamma, cudaStream_t st);
int launch_speaker_model(const Plan &pl, const Workspace &ws, const RunParams &rp, const float *Phi,
                         const int32_t *n_states, float *alpha_io, float *invL_io, bool from_given,
                         cudaStream_t st);
int launch_loglik(const Plan &pl, const Workspace &ws, const float *rho, cudaStream_t st);
int launch_forward_backward(const Plan &pl, const Workspace &ws, const RunParams &rp, float *gamma, float *pi,
                            const int32_t *n_states, double *Li, int32_t *n_iters, int32_t *flags, int iter,
                            int spl, cudaStream_t st);
// tensor-core (mma.sync 3xTF32) versions of the two in-loop contractions (vbx_mma_kernels.cu)
int launch_mstep_mma(const Plan &pl, const Workspace &ws, const float *rho, const float *gamma, cudaStream_t st);
int launch_loglik_mma(const Plan &pl, const Workspace &ws, const float *rho, cudaStream_t st);
// float64 "exact" path (vbx_f64.cu)
size_t f64_workspace_bytes(const Plan &pl);
int launch_run_f64(const Plan &pl, void *workspace, const double *fea, const double *Phi, double *gamma, double *pi,
                   const int32_t *n_states, double Fa, double Fb, double loopP, int max_iters, double epsilon,
                   double *alpha_io, double *invL_io, int warm, double *Li, int32_t *n_iters, int32_t *flags,
                   cudaStream_t st);
// tcgen05 projection (vbx_project_tc.cu)
int launch_project_tcgen05(const Plan &pl, const float *X, int D, const float *V, const float *Phi, float *rho,
                           float *gframe, cudaStream_t st, std::string *err);
int launch_gsum_from_frames(const Plan &pl, const Workspace &ws, const float *gframe, cudaStream_t st);

}  // namespace vbx
