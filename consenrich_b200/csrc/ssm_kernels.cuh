// consenrich_b200/csrc/ssm_kernels.cuh -- launch interface between the C ABI (cabi.cu) and the
// sm_100a kernels (ssm_kernels.cu).  Device pointers only.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "ssm_math.cuh"

namespace cb200 {

// Geometry of the look-back scans: every thread owns a run of CHUNK * nsub consecutive scan
// positions, processed in nsub sub-steps of CHUNK; a tile is SCAN_THREADS runs.
constexpr int SCAN_THREADS = 128;
constexpr int CHUNK = 4;
constexpr int TILE_BINS = SCAN_THREADS * CHUNK;  // positions per tile per sub-step
constexpr int MAX_NSUB = 16;
// The 2-state forward filter replays the first HEAD_BINS bins of a chromosome as ONE sequential
// run (the first thread of the first tile continues past its own run if that is shorter): under
// the diffuse prior (stateCovarInit = 1000 carried in float32) the reference's own cross
// covariances are rounding-noise dominated for about ten bins, and only an uninterrupted replay
// of its rounding sequence reproduces them.
constexpr int HEAD_BINS = 16;
constexpr int AGG_PITCH = 16;   // doubles per published tile aggregate (14 used)
constexpr int PREF_PITCH = 8;   // doubles per published tile prefix state (5 used)

struct ScanWorkspace {    // sized by scan_workspace_bytes(n); zeroed once when allocated
    double *tile_agg;     // [ntiles][AGG_PITCH]
    double *tile_pref;    // [ntiles][PREF_PITCH]
    double *partials;     // [ntiles][2]
    int32_t *flags;       // [ntiles]  epoch4 + 1 = aggregate published, epoch4 + 2 = inclusive prefix published
    int32_t *counters;    // [0] dynamic tile ticket, [1] tiles finished (reset by the last tile)
    int32_t epoch4;       // 4 * launch epoch: flags of earlier launches read as "nothing"
    long long *dbg;       // diagnostics: [dbg_tiles][8] globaltimer stamps (cb200_debug_scan_times) or nullptr
    int32_t dbg_tiles;    // tiles the diagnostics buffer holds: launches with more tiles do not stamp
};

struct FwdArgs {
    const double2 *SA, *SB;    // fold statistics per bin: SA = {S0, S1}, SB = {S2, SL}
    const float *lam, *kap, *qs;
    const double *init_state;  // device, or nullptr -> model prior
    float *xf, *Pf, *Qf, *D;
    // 2-state, whole chromosome, device pipelines: the replay also composes, run by run, the smoothing
    // elements the backward scan would otherwise build in a pass of its own (it has every filtered
    // state and process noise in registers at that moment).  smo_run: 9 arrays of smo_pitch doubles
    // (structure of arrays over the forward runs), or nullptr.
    double *smo_run;
    int64_t smo_pitch;
    int32_t nsub;              // sub-steps per run, fixed by the caller (0 = chosen per launch)
    float *q_head;             // device float[d*d] or nullptr: Q of this shard's first bin (row n-1 of
                               // the preceding shard's pNoiseForward)
    double *sums;              // device double[2] or nullptr
    double *agg_out;           // aggregate-only mode: shard aggregate destination
    int64_t n;
    double m, inv_m, mlog2pi;
    Model2 M;
    double state_init, cov_init;
    double lam_min, lam_max, kap_min, kap_max;
    int32_t use_lambda, use_kappa, use_qscale, want_nll, nll_in_d, do_store;
    int32_t head_from;         // 2-state: bins [head_from, HEAD_BINS) belong to the head replay (set at launch)
};

struct BwdArgs {
    const float *xf, *Pf, *Qf;
    const double *tail_state;  // device, or nullptr -> this shard ends the chromosome
    float *xs, *Ps, *lag;
    double *agg_out;
    int64_t n, lag_rows;
    Model2 M;
    int32_t is_last_shard;
    // run elements composed by the forward scan (FwdArgs::smo_run): the backward scan then uses the
    // forward scan's partition into runs (positions counted from npad_fixed = forward tiles x tile
    // length, same nsub) and skips its own first pass.  nullptr: off.
    const double *smo_run;
    int64_t smo_pitch, npad_fixed;
    int32_t nsub;
    // Student-t process precision update fused into the replay (cconsenrich.pyx:8252-8298, 7499-7521):
    // kap_out[k+1] from the smoothed bins k, k+1 and their lag-one covariance, as soon as the replay
    // has them in registers.  kap_out == nullptr: off.
    float *kap_out;
    int32_t no_store;          // with kap_out: do not write xs / Ps / lag (the ECM's inner sweeps need only kappa)
    const float *qs;           // processQScale or nullptr
    double nu, kap_lo, kap_hi;
    double qi00, qi01, qi10, qi11;  // Q0^-1 (state_dim 1: qi00 = 1 / Q0[0])
};

// forward filter with adaptive process noise (apn_kernels.cu): a sequential recursion over the linear
// fold statistics
struct ApnArgs {
    const double2 *SA, *SB;
    const float *lam;
    float *xf, *Pf, *Qf, *D;   // xf/Pf/Qf may be nullptr (do_store == 0); D may be nullptr
    double *sums;              // device double[2] or nullptr: {sum D, sum NLL}
    int64_t n;
    double m, inv_m, mlog2pi;
    Model2 M;
    double state_init, cov_init, lam_min, lam_max;
    double apn_min_q, apn_max_q, apn_thresh, apn_scale, apn_pc;
    double q_diag;             // 0.5 (Q0[0,0] + Q0[1,1]), or Q0[0,0] for the level model
    int32_t use_lambda, want_nll, nll_in_d, do_store;
};
cudaError_t launch_apn_forward(int dim, const ApnArgs &a, cudaStream_t st);

size_t scan_workspace_bytes(int64_t n);
ScanWorkspace scan_workspace_carve(void *base, int64_t n);
int64_t scan_num_tiles(int64_t positions, int nsub);
int scan_pick_nsub(int64_t positions, int which);
void scan_set_nsub_override(int nsub);

// every launcher returns the cudaError_t of the launch (cudaGetLastError)
// rm_logL >= 0: the statistics are written run-major for the lean sweeps (lean_kernels.cuh)
cudaError_t launch_fold(const float *data, const float *munc, int64_t m, int64_t n, int64_t ld, double pad,
                        double2 *SA, double2 *SB, cudaStream_t st, int rm_logL = -1);
cudaError_t launch_forward(int dim, const FwdArgs &a, const ScanWorkspace &ws, bool aggregate_only,
                           cudaStream_t st, int *launches);
cudaError_t launch_backward(int dim, const BwdArgs &a, const ScanWorkspace &ws, bool aggregate_only,
                            cudaStream_t st, int *launches);
cudaError_t launch_residuals(const float *data, int64_t m, int64_t n, int64_t ld, const float *xs, int dim,
                             float *resid, cudaStream_t st);
// v[0..n) = value (multipliers that start at 1 without a warm start)
cudaError_t launch_fill(float *v, int64_t n, float value, cudaStream_t st);
cudaError_t launch_update_lambda(const double2 *SA, const double2 *SB, int64_t n, double m,
                                 const float *xs, const float *Ps, int dim, double nu, double lo, double hi,
                                 float *lam, cudaStream_t st);
cudaError_t launch_update_kappa(int dim, const Model2 &M, int64_t n, const float *xs, const float *Ps,
                                const float *lag, const float *qs, double nu, double lo, double hi, float *kap,
                                cudaStream_t st);
cudaError_t launch_forward_shard_prefix(int dim, const double *aggs, int rank, double state_init,
                                        double cov_init, double *init_state, cudaStream_t st);
cudaError_t launch_backward_shard_prefix(int dim, const double *aggs, int rank, int n_shards,
                                         double *tail_state, cudaStream_t st);
cudaError_t configure_kernels();

}  // namespace cb200
