// Background-track kernels: launch interface (see background_kernels.cu).
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

namespace cb200 {

constexpr int BG_MAX_LEVELS = 32;
constexpr int BG_SMALL_ROWS = 1024;  // from this many block rows on, one CTA finishes the recursion in shared memory
constexpr int BG_SUM_BLOCKS = 1024;
constexpr int BG_MID_ROWS = 65536;    // levels up to this many block rows share one cooperative launch
constexpr int BG_MID_THREADS = 256;

// "no index" in the status records the kernels fill by atomicMin: the value cudaMemsetAsync(0x7f) leaves,
// so that arming a record needs no host-to-device copy
constexpr int64_t STATUS_NONE = 0x7f7f7f7f7f7f7f7fLL;

// outcome of a solve: first unknown whose pivot fell below the reference's floor (1e-12), or
// STATUS_NONE; the reference fails such a solve (cconsenrich.pyx:1090-1095)
struct BackgroundStatus {
    int64_t bad_index;
    double bad_value;
};

size_t background_workspace_bytes(int64_t n);

// weight[i] = sum_j inv[j][i]; rhs[i] = sum_j inv[j][i] * resid[j][i]  (float32 [m][ld] in, float64 out);
// support (or nullptr): number of intervals with weight > 0
cudaError_t launch_background_stats(const float *resid, const float *inv, int64_t m, int64_t n, int64_t ld,
                                    double *weight, double *rhs, unsigned long long *support, cudaStream_t st);

// out[k] = state[k] - sum_j w_jk (data[j][k] - background[k]) / sum_j w_jk, w_jk = 1 / max(munc[j][k] + pad, 1e-12)
// over the valid cells of interval k (core.py:2670-2696); NaN where no cell is valid.  background may be null.
cudaError_t launch_weighted_mean_residual(const float *data, const float *munc, int64_t m, int64_t n, int64_t ld,
                                          const double *state, const double *background, double pad, double *out,
                                          cudaStream_t st);

// core._perIntervalOutputDiagnosticTracks (core.py:7734-7880), device parts.
// munc_trace[k] = sum_j max(munc[j][k] + pad, 1e-12) / obs_prec[k], sum_inv_r[k] = sum_j obs_prec[k] / max(...),
// finite terms only.
cudaError_t launch_diag_obs_sums(const float *munc, int64_t m, int64_t n, int64_t ld, const double *obs_prec, double pad,
                                 double *munc_trace, double *sum_inv_r, cudaStream_t st);
// summed Kalman gain of every interval from the filtered covariance of the one before it
struct DiagGainArgs {
    const float *covar;      // [n][cov_dim][cov_dim] filtered covariances (float32)
    const float *p_noise;    // [n][cov_dim][cov_dim] stored process noise (Q_k at row k - 1) or nullptr
    const double *q_scale;   // [n]
    const double *proc_prec; // [n] clipped kappa, or nullptr
    const double *sum_inv_r; // [n]
    double *sum_gain0, *sum_gain1;  // [n]
    int64_t n;
    int32_t dim, cov_dim;    // state dimension used (1 or 2) and the stored matrices' dimension
    double base_q[4], f[4], cov_init;
};
cudaError_t launch_diag_gain(const DiagGainArgs &a, cudaStream_t st);

// (diag(w) + lam_first D1'D1 + lam D2'D2) x = rhs, optionally with sum(x) = 0; n >= 2.
// workspace: background_workspace_bytes(n) bytes of device memory; status: device BackgroundStatus.
cudaError_t launch_background_solve(const double *w, const double *rhs, int64_t n, double lam, double lam_first,
                                    int zero_center, double *out, void *workspace, BackgroundStatus *status,
                                    cudaStream_t st, int *launches);

}  // namespace cb200
