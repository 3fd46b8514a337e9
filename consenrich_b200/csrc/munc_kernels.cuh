// Dense kernels of the observation-noise (MUNC) stage: launch interface (see munc_kernels.cu).
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

#include "background_kernels.cuh"

namespace cb200 {

constexpr int64_t MUNC_MAX_WINDOW = 8192;  // tile + window cells must fit one CTA's shared memory

int munc_rolling_groups(int64_t window);
size_t munc_rolling_smem(int groups);

// out[j][i] = max(eps, mean of the unmasked local[j][k] over the centred window of i) (float32),
// the cell itself where the whole window is masked.  mask_mode: 0 none, 1 per interval [n],
// 2 per cell [m][mask_ld]; nonzero excludes.  *invalid (device int) is set when an unmasked cell is
// not positive and finite.  1 <= window <= MUNC_MAX_WINDOW.
cudaError_t launch_munc_rolling_mean(const float *local, const uint8_t *mask, int mask_mode, int64_t m, int64_t n,
                                     int64_t ld, int64_t mask_ld, int64_t window, double eps, float *out, int64_t out_ld,
                                     int *invalid, cudaStream_t st);

// outcome of launch_munc_finalize_eb (device memory): counters of cconsenrich.pyx:5355-5362; the invalid_*
// fields hold the first offending interval of each kind or STATUS_NONE (background_kernels.cuh)
struct MuncFinalizeStatus {
    int64_t support, cfloor_finite, cfloor_added, cfloor_missing;
    int64_t invalid_local, invalid_prior, invalid_cfloor;
};

// out[i] = clip(clip((nu_local clip(local[i]) + nu_prior clip(prior[i])) / (nu_local + nu_prior)) + cfloor[i])
// (prior == nullptr / use_eb == 0: no shrinkage; cfloor == nullptr: no count floor; NaN in cfloor: none for
// that interval).  Counters are valid only when no invalid_* index was reported.
cudaError_t launch_munc_finalize_eb(const float *local, const float *prior, const float *cfloor, int64_t n,
                                    double nu_local, double nu_prior, double vfloor, double vcap, int use_eb, float *out,
                                    MuncFinalizeStatus *status, cudaStream_t st);

// cEMA (cconsenrich.pyx:5744-5759): out = the forward-then-backward exponential filter of x (float32 or
// float64 track, n elements); tmp: n elements of scratch; workspace: munc_ema_workspace_bytes(n).
size_t munc_ema_workspace_bytes(int64_t n);
cudaError_t launch_munc_ema(const void *x, void *tmp, void *out, int64_t n, int is_double, double alpha, void *workspace,
                            cudaStream_t st);

// cMuncObservationMomentSeedPass (cconsenrich.pyx:4843-5040): arguments of one launch.  Matrices are
// float32 [m][ld]; per-interval vectors float32 [n]; nullptr = absent.
struct MuncSeedArgs {
    const float *data, *munc;          // [m][ld]
    const float *state_mean, *state_var;  // [n]
    const float *background, *g_var;   // [n] or nullptr
    const float *count_floor;          // [m][ld] or nullptr
    const float *omega_in;             // [n] or nullptr
    const float *rho_in;               // [m][ld] or nullptr (all ones)
    const uint8_t *active;             // nullptr, [n] (active_mode 1) or [m][active_ld] (2); nonzero = active
    float *moment, *rho_out, *local, *variance;  // [m][ld]
    float *omega_raw, *omega_out;      // [n]
    int64_t m, n, ld, active_ld;
    int32_t active_mode, use_weights, student_t, update_weights;
    double pad, d_s, d_omega, omega_min, omega_max, var_floor, var_cap;
};

// *invalid (device int) is set when an active cell fails the reference's input check (pyx:4767-4840)
cudaError_t launch_munc_seed_pass(const MuncSeedArgs &a, int *invalid, cudaStream_t st);

}  // namespace cb200
