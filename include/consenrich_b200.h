/*
 * consenrich_b200.h -- C ABI of libconsenrich_b200.so
 *
 * B200 (sm_100a) implementation of Consenrich's state-space hot path: the multi-sample
 * Kalman forward filter, the RTS smoother and the fixed-background ECM (Student-t precision
 * re-weighting) loop.  These entry points are what a binding for the six native functions
 * of the reference's `consenrich.cconsenrich` module would call:
 *
 *   reference (src/consenrich/cconsenrich.pyx)          this library
 *   -------------------------------------------------   ---------------------------------
 *   cforwardPass               :6393-6632  (loop 291)   cb200_host_forward_pass (state_dim 2)
 *   cforwardPassLevel          :6853-7049  (loop 538)   cb200_host_forward_pass (state_dim 1)
 *   cbackwardPass              :6635-6850               cb200_host_backward_pass (state_dim 2)
 *   cbackwardPassLevel         :7052-7150               cb200_host_backward_pass (state_dim 1)
 *   cfixedBackgroundECM        :7660-8442               cb200_host_ecm (state_dim 2)
 *   cfixedBackgroundECMLevel   :7153-7657               cb200_host_ecm (state_dim 1)
 *   _accumulateObservationValue :259-283                cb200_fold_tracks (device fold kernel)
 *   cbackgroundWeightedStats[WithSupport] :9675-9724    cb200_host_background_stats
 *   csolveZeroCenteredBackground :944-1096              cb200_host_background_solve
 *   cMuncSmoothDenseLocalEvidence :5547-5740            cb200_host_munc_smooth_local_evidence
 *   cFinalizeMuncEBTrack :5372-5545                     cb200_host_munc_finalize_eb
 *   cMuncObservationMomentSeedPass :4843-5345           cb200_host_munc_seed_pass
 *   cEMA :5744-5759, 5897-5915                          cb200_host_ema
 *
 * Conventions
 *   - plain C: pointers, sizes, POD structs; no torch / numpy types.
 *   - "tracks" = the m samples (rows), "intervals" = the n genomic bins (columns); matrices are
 *     row-major float32 [m x n] with a row stride `ld` (elements) >= n.
 *   - every function returns CB200_OK or an error code; cb200_last_error() gives the message
 *     (for CB200_ERR_INVALID the text is the reference's ValueError text, so a binding can
 *     re-raise it verbatim).
 *   - cb200_host_* take HOST pointers and run H2D copies, kernels and D2H copies on the
 *     context's stream, returning after the results are in the caller's buffers.
 *   - all other entry points take DEVICE pointers, enqueue work on the context's stream and
 *     return without synchronising (inputs already resident in HBM).
 *   - there is no CPU implementation behind any entry point: without a CUDA device
 *     cb200_ctx_create fails with CB200_ERR_CUDA.
 */
#ifndef CONSENRICH_B200_H
#define CONSENRICH_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CB200_OK 0
#define CB200_ERR_INVALID 1     /* bad argument: maps to the reference's ValueError */
#define CB200_ERR_CUDA 2        /* CUDA runtime / driver failure */
#define CB200_ERR_UNSUPPORTED 3 /* valid in the reference, not expressible as a scan (APN) */

#define CB200_ABI_VERSION 3

#if defined(__GNUC__)
#define CB200_API __attribute__((visibility("default")))
#else
#define CB200_API
#endif

typedef struct cb200_ctx cb200_ctx;

/* State-space model and per-call switches.  Scalars that the reference receives as C `float`
 * (stateInit, stateCovarInit, pad, the multiplier bounds; cconsenrich.pyx:6400-6426) must be
 * rounded to float by the caller before being widened into these doubles. */
typedef struct cb200_model {
    int32_t state_dim;      /* 2 = level + trend (cforwardPass), 1 = level (cforwardPassLevel) */
    int32_t use_lambda;     /* per-interval observation precision multipliers are live */
    int32_t use_kappa;      /* per-interval process precision multipliers are live */
    int32_t use_qscale;     /* per-interval processQScale is live */
    int32_t return_nll;     /* accumulate the Gaussian negative log-likelihood */
    int32_t store_nll_in_d; /* vectorD holds the per-interval NLL instead of NIS */
    int32_t use_apn;        /* ECM_useAPN: adaptive process noise (cconsenrich.pyx:510-527, 688-703).  Live only
                             * without processQScale and with 0.5 (Q0[0,0] + Q0[1,1]) > 1e-12 (pyx:6574-6576), as in
                             * the reference; the forward pass is then a sequential device recursion, not a scan */
    int32_t reserved1;
    double F[4];            /* row-major transition matrix (ignored when state_dim == 1) */
    double Q0[4];           /* row-major base process noise; state_dim 1 uses Q0[0] */
    double state_init, cov_init, pad;
    double lam_min, lam_max; /* observation precision multiplier clamp */
    double kap_min, kap_max; /* process precision multiplier clamp */
    double apn_min_q, apn_max_q, apn_thresh, apn_scale, apn_pc; /* APN_minQ, APN_maxQ, APN_dStatThresh,
                             * APN_dStatScale, APN_dStatPC (rounded to float by the caller) */
} cb200_model;

/* ECM controls (cconsenrich.pyx:7660-7693). */
typedef struct cb200_ecm_opts {
    int32_t max_iters;       /* ECM_fixedBackgroundIters */
    int32_t inner_iters;     /* t_innerIters */
    int32_t update_lambda;   /* ECM_useObsPrecisionReweighting */
    int32_t update_kappa;    /* ECM_useProcessPrecisionReweighting (and not disabled by APN) */
    int32_t want_outputs;    /* returnIntermediates: smoothed tracks + residuals are produced */
    int32_t init_ones;       /* cb200_host_ecm: bit 0 lambda, bit 1 kappa start at 1 (no warm start): the host
                              * array is an output only and is not read */
    double rtol;             /* ECM_fixedBackgroundRtol (rounded to float by the caller) */
    double nu;               /* ECM_robustTNu (rounded to float by the caller) */
} cb200_ecm_opts;

/* ECM outcome; mirrors the diagnostics dict of cconsenrich.pyx:8409-8425. */
typedef struct cb200_ecm_result {
    int32_t iters_done;
    int32_t converged;
    int32_t skipped;          /* n <= 5: filter + smoother only (pyx:7998-8129) */
    int32_t stable_iters;
    int32_t nll_increase_count;
    int32_t has_initial;
    double initial_nll, final_nll, final_abs_rel_change, final_rel_improvement;
} cb200_ecm_result;

/* ---- context ------------------------------------------------------------------------ */
CB200_API int cb200_abi_version(void);
/* The CUDA device the calling thread has selected (cudaGetDevice): the drop-in functions create
 * their context there, so that one process per GPU (torchrun) needs no extra argument. */
CB200_API int cb200_current_device(int *device);
/* stream: a cudaStream_t to enqueue on, or NULL for a stream owned by the context. */
CB200_API int cb200_ctx_create(int device, void *stream, cb200_ctx **out);
CB200_API void cb200_ctx_destroy(cb200_ctx *ctx);
CB200_API int cb200_ctx_set_stream(cb200_ctx *ctx, void *stream);
CB200_API int cb200_ctx_sync(cb200_ctx *ctx);
CB200_API const char *cb200_last_error(void); /* thread-local; valid until the next failing call */
/* number of kernel launches this context has enqueued (bench.py's gpu_launches). */
CB200_API int64_t cb200_ctx_launch_count(const cb200_ctx *ctx);
/* cumulative device time (ms) of the named kernel family since the last reset, measured
 * with CUDA events on the context's stream when timing is enabled: 0 fold, 1 forward scan (lean
 * sweeps: the replay kernel), 2 backward scan (lean: the kappa-carrying replay), 3 residuals,
 * 4 precision updates and run-major copies, 5 background, 6 variance stage, 7 lean run-element
 * composition, 8 lean segment scans, 9 lean publishing backward replay. */
CB200_API int cb200_ctx_enable_timing(cb200_ctx *ctx, int on);
CB200_API int cb200_ctx_kernel_ms(cb200_ctx *ctx, int family, double *ms, int64_t *launches);
CB200_API int cb200_ctx_reset_timing(cb200_ctx *ctx);
/* Event pairs keep neighbouring kernels from overlapping their ramps (about 5 us per launch): with a stride s
 * only every s-th launch of each family is bracketed (cb200_ctx_kernel_ms then reports the sampled launches),
 * while cb200_ctx_kernel_launches counts every launch of the family made with timing enabled. */
CB200_API int cb200_ctx_set_timing_stride(cb200_ctx *ctx, int stride);
CB200_API int cb200_ctx_kernel_launches(cb200_ctx *ctx, int family, int64_t *launches);
/* Tuning knob (process-wide): sub-steps of 4 bins each thread of the scan kernels runs through
 * (1..16), 0 = choose per launch from the track length and the resident tile slots.  Results do not
 * depend on it beyond float64 re-association.  Also settable with CB200_SCAN_NSUB. */
CB200_API int cb200_set_scan_substeps(int nsub);
/* Tuning / diagnostics knob (process-wide).  The inner sweeps of cb200_ecm_device -- 2-state model,
 * F = [[1, f], [0, 1]], process precision the only multiplier fitted (the CLI default of
 * cfixedBackgroundECM, cconsenrich.pyx:7660) -- run on run-major private tracks (csrc/lean_kernels.cuh)
 * when the track has at least 4096 intervals.  mode 0: always the look-back scan kernels; 1: lean sweeps
 * where eligible (default).  log2_run: 5 or 6 fixes the run length at 32 / 64 intervals, 0 chooses per
 * call.  Results agree between the two paths to float64 re-association of each run's start state.
 * Also settable with CB200_NO_LEAN / CB200_LEAN_LOGL. */
CB200_API int cb200_set_lean_sweeps(int mode, int log2_run);
/* Diagnostics.  (tiles > 0, host_out NULL) arms phase stamping: every tile of the following scan
 * launches writes eight %globaltimer values (0 start, 1 run elements composed, 2 prefix known, 3 end,
 * 4 look-back flags ready, 5 window loaded, 6 look-back entered, 7 aggregate published).
 * (host_out non-NULL) copies the stamps of the last launch out ([tiles][8] int64 ns) and disarms.
 * With CB200_DEBUG_STAMP_LAUNCH=k in the environment when arming, only the k-th (0-based) scan
 * launch after arming stamps, so that one launch inside cb200_ecm_device can be looked at. */
CB200_API int cb200_debug_scan_times(cb200_ctx *ctx, int64_t tiles, long long *host_out);

/* ---- memory helpers (so that hosts without torch can stage tracks) -------------------- */
CB200_API int cb200_device_alloc(cb200_ctx *ctx, size_t bytes, void **dptr);
CB200_API int cb200_device_free(cb200_ctx *ctx, void *dptr);
CB200_API int cb200_pinned_alloc(size_t bytes, void **hptr);
CB200_API int cb200_pinned_free(void *hptr);
/* rows x row_bytes, pitched on either side; asynchronous on the context's stream. */
CB200_API int cb200_copy_h2d(cb200_ctx *ctx, void *dst, size_t dst_pitch, const void *src, size_t src_pitch,
                   size_t row_bytes, size_t rows);
CB200_API int cb200_copy_d2h(cb200_ctx *ctx, void *dst, size_t dst_pitch, const void *src, size_t src_pitch,
                   size_t row_bytes, size_t rows);

/* ---- device-resident path ------------------------------------------------------------- */
/* Fold the m observations of every interval into information form, one coalesced pass over
 * data and munc (replaces the per-(interval, sample) calls of _accumulateObservationValue,
 * cconsenrich.pyx:259-283).  stats = 4 * stat_stride doubles: first stat_stride pairs {S0, S1}, then
 * stat_stride pairs {S2, SL}; S0 = sum 1/r, S1 = sum z/r, S2 = sum z^2/r, SL = sum log r,
 * r = max(munc + pad, 1e-12).  Opaque to callers: produced here, consumed by the scans. */
CB200_API int cb200_fold_tracks(cb200_ctx *ctx, const float *data, const float *munc, int64_t m, int64_t n,
                      int64_t ld, double pad, double *stats, int64_t stat_stride);

/* Forward filter as a single-pass decoupled look-back scan (replaces the loops at
 * cconsenrich.pyx:291-529 and 538-707).  lam/kap/qscale: float32 [n] or NULL according to the
 * model flags.  xf [n][d], Pf [n][d][d], Qf [n][d][d] (Q_k stored at k-1) may all be NULL (no
 * store).  D float32 [n] may be NULL.  sums: device double[2] = {sum of float-rounded D, sum
 * NLL}, or NULL.  init_state: device double[5] {x0, x1, P00, P01, P11} (d = 1: {x, P}) that
 * overrides the model's prior -- the carry of the preceding contiguous shard -- or NULL. */
CB200_API int cb200_forward_scan(cb200_ctx *ctx, const cb200_model *model, const double *stats,
                       int64_t stat_stride, int64_t m, int64_t n, const float *lam, const float *kap,
                       const float *qscale, const double *init_state, float *xf, float *Pf, float *Qf,
                       float *D, double *sums);

/* Same as cb200_forward_scan for a shard of a split chromosome: q_head (device float[d*d], or NULL)
 * receives Q of the shard's FIRST interval, which the reference stores in row n-1 of the PRECEDING
 * shard's pNoiseForward (Q_k lives at row k-1) and which that shard's smoother needs. */
CB200_API int cb200_forward_scan_shard(cb200_ctx *ctx, const cb200_model *model, const double *stats,
                             int64_t stat_stride, int64_t m, int64_t n, const float *lam, const float *kap,
                             const float *qscale, const double *init_state, float *xf, float *Pf,
                             float *Qf, float *D, double *sums, float *q_head);

/* Aggregate filtering element of a whole shard (14 doubles for d = 2: A, b, C, eta, J; 5 for
 * d = 1), the only thing ranks exchange when a chromosome is split into contiguous ranges. */
CB200_API int cb200_forward_shard_aggregate(cb200_ctx *ctx, const cb200_model *model, const double *stats,
                                  int64_t stat_stride, int64_t n, const float *lam, const float *kap,
                                  const float *qscale, double *agg);
/* Combine the prior with the gathered aggregates of shards 0..rank-1 (aggs: [n_shards][16]
 * doubles, 16-double pitch) into the init_state of shard `rank`. */
CB200_API int cb200_forward_shard_prefix(cb200_ctx *ctx, const cb200_model *model, const double *aggs,
                               int32_t rank, double *init_state);

/* RTS smoother as a reverse decoupled look-back scan (replaces cconsenrich.pyx:6740-6848 and
 * 7116-7148, residuals excluded).  lag has lag_rows rows.  tail_state: device double[5]
 * smoothed {x, P} of the first interval of the FOLLOWING shard, or NULL when this shard ends
 * the chromosome. */
CB200_API int cb200_backward_scan(cb200_ctx *ctx, const cb200_model *model, int64_t n, const float *xf,
                        const float *Pf, const float *Qf, const double *tail_state, float *xs,
                        float *Ps, float *lag, int64_t lag_rows);
/* cb200_backward_scan of a whole chromosome with the Student-t process precision update
 * (cb200_update_kappa) carried out inside the replay: kap[k+1] is written as soon as the smoothed
 * intervals k and k+1 and their lag-one covariance are in registers; kap[0] = 1.  kap may be the
 * array the preceding forward scan read. */
CB200_API int cb200_backward_scan_kappa(cb200_ctx *ctx, const cb200_model *model, int64_t n, const float *xf,
                              const float *Pf, const float *Qf, float *xs, float *Ps, float *lag,
                              int64_t lag_rows, const float *qscale, double nu, float *kap);
CB200_API int cb200_backward_shard_aggregate(cb200_ctx *ctx, const cb200_model *model, int64_t n, const float *xf,
                                   const float *Pf, const float *Qf, int32_t is_last_shard,
                                   double *agg);
CB200_API int cb200_backward_shard_prefix(cb200_ctx *ctx, const cb200_model *model, const double *aggs,
                                int32_t rank, int32_t n_shards, double *tail_state);

/* postFitResiduals[k][j] = data[j][k] - xs[k][0], float32 [n][m] (transposed with respect to
 * data; cconsenrich.pyx:6846-6848). */
CB200_API int cb200_residuals(cb200_ctx *ctx, const float *data, int64_t m, int64_t n, int64_t ld,
                    const float *xs, int32_t state_dim, float *resid);

/* Student-t precision multipliers (cconsenrich.pyx:8210-8298, 7474-7521). */
CB200_API int cb200_update_lambda(cb200_ctx *ctx, const cb200_model *model, const double *stats,
                        int64_t stat_stride, int64_t m, int64_t n, const float *xs, const float *Ps,
                        double nu, float *lam);
CB200_API int cb200_update_kappa(cb200_ctx *ctx, const cb200_model *model, int64_t n, const float *xs,
                       const float *Ps, const float *lag, const float *qscale, double nu, float *kap);

/* Whole ECM loop on device-resident tracks.  lam / kap are in-out float32 [n] (warm start,
 * already clipped) or NULL when the corresponding update is off.  xs, Ps, lag are required
 * work/output tracks; resid may be NULL.  Blocks until the loop has finished (the stopping
 * rule reads one double per iteration). */
CB200_API int cb200_ecm_device(cb200_ctx *ctx, const cb200_model *model, const cb200_ecm_opts *opts,
                     const float *data, const float *munc, int64_t m, int64_t n, int64_t ld,
                     const float *qscale, float *lam, float *kap, float *xs, float *Ps, float *lag,
                     float *resid, cb200_ecm_result *result, double *nll_path /* [max_iters] or NULL */);

/* ---- reference-facing path (host buffers) ---------------------------------------------- */
/* cforwardPass / cforwardPassLevel.  block_map is range-checked only (pyx:389-392). */
CB200_API int cb200_host_forward_pass(cb200_ctx *ctx, const cb200_model *model, const float *data,
                            const float *munc, int64_t m, int64_t n, const int32_t *block_map,
                            int64_t block_count, const float *lam, const float *kap,
                            const float *qscale, float *xf, float *Pf, float *Qf, float *D,
                            double *sum_d, double *sum_nll);
/* cbackwardPass / cbackwardPassLevel. */
CB200_API int cb200_host_backward_pass(cb200_ctx *ctx, const cb200_model *model, const float *data, int64_t m,
                             int64_t n, const float *xf, const float *Pf, const float *Qf, float *xs,
                             float *Ps, float *lag, int64_t lag_rows, float *resid);
/* Fused forward + backward sweep on one upload of the tracks (what core._runForwardBackward,
 * core.py:4207, does with two native calls).  Any output pointer may be NULL. */
CB200_API int cb200_host_sweep(cb200_ctx *ctx, const cb200_model *model, const float *data, const float *munc,
                     int64_t m, int64_t n, const float *lam, const float *kap, const float *qscale,
                     float *xf, float *Pf, float *Qf, float *D, double *sum_d, double *sum_nll,
                     float *xs, float *Ps, float *lag, int64_t lag_rows, float *resid);
/* cfixedBackgroundECM / cfixedBackgroundECMLevel. */
CB200_API int cb200_host_ecm(cb200_ctx *ctx, const cb200_model *model, const cb200_ecm_opts *opts,
                   const float *data, const float *munc, int64_t m, int64_t n,
                   const int32_t *block_map, int64_t block_count, const float *qscale, float *lam,
                   float *kap, float *xs, float *Ps, float *lag, float *resid,
                   cb200_ecm_result *result, double *nll_path);

/* ---- background track (the caller on the other side of the ECM: core.py:8085-8378) ------ */
/* cbackgroundWeightedStats[WithSupport] (cconsenrich.pyx:9675-9724): weight[i] = sum_j inv[j][i],
 * rhs[i] = sum_j inv[j][i] * resid[j][i], float64, accumulated over j in order (bit-identical to the
 * reference).  resid, inv: float32 [m][ld] on the device.  support: device counter of the intervals
 * with weight > 0, or NULL. */
CB200_API int cb200_background_stats(cb200_ctx *ctx, const float *resid, const float *inv, int64_t m, int64_t n,
                           int64_t ld, double *weight, double *rhs, unsigned long long *support);
/* csolveZeroCenteredBackground (cconsenrich.pyx:944-1096): solves
 * (diag(weight) + lam_first D1'D1 + lam D2'D2) x = rhs, with sum(x) = 0 imposed by a Lagrange
 * multiplier when zero_center is set, by block cyclic reduction (float64; agrees with the reference's
 * sequential LDL' to rounding times the conditioning of the system).  weight, rhs, out: device
 * float64 [n].  *bad_index (host) receives the first unknown whose pivot fell below the reference's
 * 1e-12 floor, or -1, and *bad_value that pivot: the reference raises RuntimeError for such a system. */
CB200_API int cb200_background_solve(cb200_ctx *ctx, const double *weight, const double *rhs, int64_t n, double lam,
                           double lam_first, int32_t zero_center, double *out, int64_t *bad_index,
                           double *bad_value);
/* The same two with HOST arrays (uploads, kernels, downloads; return when the results are in place).  The
 * degenerate shapes the reference special-cases are answered without a launch: a single interval is one
 * division (cconsenrich.pyx:995-1006), no tracks give zero sums. */
CB200_API int cb200_host_background_stats(cb200_ctx *ctx, const float *resid, const float *inv, int64_t m, int64_t n,
                                double *weight, double *rhs, int64_t *support);
CB200_API int cb200_host_background_solve(cb200_ctx *ctx, const double *weight, const double *rhs, int64_t n,
                                double lam, double lam_first, int32_t zero_center, double *out,
                                int64_t *bad_index, double *bad_value);

/* ---- driver-side [tracks x intervals] reductions --------------------------- */
/* The matrix half of core._relativeSignChangePerKB (core.py:2647-2700): out[k] = state[k] minus the
 * inverse-variance weighted mean over the tracks of data[j][k] - background[k], weights
 * 1 / max(munc[j][k] + pad, 1e-12), over the cells that are finite with a positive denominator; NaN where an
 * interval has none.  float32 matrices, float64 vectors; numpy's float64 arithmetic in its order
 * (bit-identical).  The sign-change density the reference derives from `out` is a per-interval vector
 * operation and stays the reference's own function (core.py:2614-2644). */
CB200_API int cb200_weighted_mean_residual(cb200_ctx *ctx, const float *data, const float *munc, int64_t m, int64_t n,
                                 int64_t ld, const double *state, const double *background, double pad,
                                 double *out);
CB200_API int cb200_host_weighted_mean_residual(cb200_ctx *ctx, const float *data, const float *munc, int64_t m,
                                      int64_t n, const double *state, const double *background, double pad,
                                      double *out);

/* The two costly parts of core._perIntervalOutputDiagnosticTracks (core.py:7734-7880; run once per fit when
 * precision diagnostics are requested):
 *   munc_trace[k] = sum_j v_jk / obs_prec[k],  sum_inv_r[k] = sum_j obs_prec[k] / v_jk,  v_jk = max(munc[j][k] + pad,
 *   1e-12), finite terms only (core.py:7786-7800);
 *   sum_gain0/1[k]: the level / trend entries of the summed Kalman gain from the one-step prediction of the
 *   filtered covariance of interval k - 1 (the prior cov_init I for k = 0), core.py:7840-7866 -- the reference's
 *   Python loop over all intervals, here one thread per interval.
 * float32 matrices, float64 vectors in and out; float64 arithmetic, separately rounded operations. */
typedef struct cb200_diag_gain_args {
    const float *covar;       /* [n][cov_dim][cov_dim] stateCovarForward */
    const float *p_noise;     /* [n][cov_dim][cov_dim] pNoiseForward (Q_k at row k-1) or NULL */
    const double *q_scale;    /* [n] processQScale (ones when absent) */
    const double *proc_prec;  /* [n] clipped processPrecExp or NULL (then p_noise, when given, is used) */
    const double *sum_inv_r;  /* [n] */
    double *sum_gain0, *sum_gain1; /* [n] */
    int64_t n;
    int32_t dim, cov_dim;     /* state dimension (1 or 2), dimension of the stored matrices (>= dim) */
    double base_q[4], f[4], cov_init;
} cb200_diag_gain_args;
CB200_API int cb200_diag_obs_sums(cb200_ctx *ctx, const float *munc, int64_t m, int64_t n, int64_t ld,
                        const double *obs_prec, double pad, double *munc_trace, double *sum_inv_r);
CB200_API int cb200_diag_gain(cb200_ctx *ctx, const cb200_diag_gain_args *args);
/* both, with HOST arrays (munc contiguous [m][n]); args->sum_inv_r is ignored (it is computed here) */
CB200_API int cb200_host_interval_diagnostics(cb200_ctx *ctx, const float *munc, int64_t m, const double *obs_prec,
                                    double pad, const cb200_diag_gain_args *args, double *munc_trace,
                                    double *sum_inv_r);

/* ---- observation-noise (MUNC) stage: dense [tracks x intervals] kernels --------------------- */
#define CB200_MUNC_MAX_WINDOW 8192
/* cMuncSmoothDenseLocalEvidence (cconsenrich.pyx:5547-5740): out[j][i] = max(eps, mean of the unmasked
 * local[j][k] over the window of `window` intervals centred on i (clipped at the ends)), the cell itself
 * where the whole window is masked; float32 in and out, float64 sums.  mask_mode 0: no mask; 1: mask is
 * uint8 [n]; 2: mask is uint8 [m][mask_ld]; nonzero excludes a cell.  *invalid (device int32) is set to 1
 * when an unmasked cell is not positive and finite (the reference raises ValueError).
 * 1 <= window <= CB200_MUNC_MAX_WINDOW, otherwise CB200_ERR_UNSUPPORTED. */
CB200_API int cb200_munc_smooth_local_evidence(cb200_ctx *ctx, const float *local, const unsigned char *mask,
                                     int32_t mask_mode, int64_t m, int64_t n, int64_t ld, int64_t mask_ld,
                                     int64_t window, double eps, float *out, int64_t out_ld, int32_t *invalid);
/* The same with HOST arrays (local, out: [m][n] contiguous; mask: [n] or [m][n]); *invalid is a host int32. */
CB200_API int cb200_host_munc_smooth_local_evidence(cb200_ctx *ctx, const float *local, const unsigned char *mask,
                                          int32_t mask_mode, int64_t m, int64_t n, int64_t window, double eps,
                                          float *out, int32_t *invalid);

/* cFinalizeMuncEBTrack (cconsenrich.pyx:5372-5545): per interval, the local variance clipped to
 * [variance_floor, variance_cap], shrunk towards the clipped prior with weights nu_local : nu_prior when
 * use_eb is set, clipped, plus the count floor where that is not NaN, clipped again; float32 in and out,
 * float64 arithmetic with the reference's rounding sequence (bit-identical).  prior / count_floor may be
 * NULL (no shrinkage / no count floor).  The outcome mirrors the reference's diagnostics; an invalid_*
 * field >= 0 names the first interval for which the reference raises ValueError (the one with the
 * smallest index wins; local before prior before count floor at equal index), and the counters and `out`
 * are then meaningless. */
typedef struct cb200_munc_finalize_result {
    int64_t support_count;        /* intervals with local variance above the floor */
    int64_t count_floor_finite;
    int64_t count_floor_added;
    int64_t count_floor_missing;
    int64_t invalid_local;        /* -1 or the first interval whose local variance is not positive and finite */
    int64_t invalid_prior;
    int64_t invalid_count_floor;  /* finite-or-inf but negative / infinite */
} cb200_munc_finalize_result;
/* device arrays; `result` is HOST memory (the call synchronises the stream to fill it) */
CB200_API int cb200_munc_finalize_eb(cb200_ctx *ctx, const float *local, const float *prior, const float *count_floor,
                           int64_t n, double nu_local, double nu_prior, double variance_floor, double variance_cap,
                           int32_t use_eb, float *out, cb200_munc_finalize_result *result);
/* host arrays */
CB200_API int cb200_host_munc_finalize_eb(cb200_ctx *ctx, const float *local, const float *prior,
                                const float *count_floor, int64_t n, double nu_local, double nu_prior,
                                double variance_floor, double variance_cap, int32_t use_eb, float *out,
                                cb200_munc_finalize_result *result);

/* cMuncObservationMomentSeedPass (cconsenrich.pyx:4843-5345): per cell, the squared residual against the
 * seed smoother's state plus its variance (`moment`), the Student-t cell weight (`rho_out`), per interval
 * the track-averaged weight (`omega_raw`, clamped: `omega_out`), and from them the local variance and the
 * total variance (local + count floor), both clipped to [variance_floor, variance_cap].  float32 in and
 * out, float64 arithmetic in the reference's order with separately rounded operations: bit-identical.
 * Matrices are [m][ld] (cb200_munc_seed_pass: device, pitch ld; cb200_host_munc_seed_pass: host,
 * contiguous, ld ignored); vectors are [n]; NULL = absent (background, g_var, count_floor, omega_in,
 * rho_in (treated as ones), active).  active_mode 0: every cell active; 1: active is uint8 [n];
 * 2: uint8 [m][active_ld]; nonzero = active.  *invalid is set to 1 when an active cell fails the
 * reference's input check (pyx:4767-4840; it raises ValueError). */
typedef struct cb200_munc_seed_args {
    const float *data, *munc, *state_mean, *state_var, *background, *g_var, *count_floor, *omega_in, *rho_in;
    const unsigned char *active;
    float *moment, *rho_out, *omega_raw, *omega_out, *local, *variance;
    int64_t m, n, ld, active_ld;
    int32_t active_mode;
    int32_t use_weights;     /* enabled and useSeedWeights */
    int32_t student_t;
    int32_t update_weights;
    double pad, student_t_df, d_omega, omega_min, omega_max, variance_floor, variance_cap;
} cb200_munc_seed_args;
/* device arrays; *invalid is a device int32 */
CB200_API int cb200_munc_seed_pass(cb200_ctx *ctx, const cb200_munc_seed_args *args, int32_t *invalid);
/* host arrays; *invalid is a host int32 */
CB200_API int cb200_host_munc_seed_pass(cb200_ctx *ctx, const cb200_munc_seed_args *args, int32_t *invalid);

/* cEMA (cconsenrich.pyx:5744-5759, 5897-5915): y[0] = x[0], y[i] = alpha x[i] + (1 - alpha) y[i-1] forward,
 * then the same recurrence backward over y, for a float32 (is_double = 0; alpha rounded to float, products
 * and sums in the reference's mixed float/double sequence) or float64 track.  The recurrence is scanned in
 * parallel and replayed with a warm-up, so the result equals the reference's sequential loop to the
 * tolerance its own test states (1e-6 relative for float32, 1e-14 for float64) and in practice to the last
 * bit once (1 - alpha)^256 is below the type's resolution.  0 <= alpha <= 1 (the reference leaves `out`
 * unwritten otherwise; here that is CB200_ERR_INVALID).  x, out: device arrays of n elements. */
CB200_API int cb200_ema(cb200_ctx *ctx, const void *x, int64_t n, int32_t is_double, double alpha, void *out);
CB200_API int cb200_host_ema(cb200_ctx *ctx, const void *x, int64_t n, int32_t is_double, double alpha, void *out);

/* ---- a chromosome split over several GPUs ---------------------------------------
 * cfixedBackgroundECM (2-state, kappa the only multiplier fitted) on ONE chromosome whose bins are split into
 * contiguous ranges, one shard per rank / context.  The shards run the same lean sweeps as a whole chromosome
 * (csrc/lean_kernels.cuh) and exchange one payload of CB200_SPLIT_PAYLOAD doubles per pass, which the caller
 * all-gathers (NCCL; consenrich_b200/sharding.py: SplitECM) -- nothing else crosses the GPUs:
 *   forward : [0..13] the shard's filtering aggregate, [14] kappa and [15] qScale of its first interval
 *   backward: [0..8] the shard's smoothing aggregate, [9..13] the filtered Gaussian of its last interval
 * A pass = compose (local, fills the payload) -> all-gather -> replay (local).  The kappa of a shard's first
 * interval (cconsenrich.pyx:8244-8298 needs smoothed intervals k and k+1 on either side of the boundary) is
 * formed by that shard from the previous shard's last filtered interval, which rides in the backward payload.
 * All pointers are device pointers; every call is asynchronous on the context's stream except cb200_split_end.
 *   sums (forward_replay, with_nll): device double[2], [1] receives the shard's part of the NLL.
 *   set: which of the two forward-track sets the pass writes / reads (the ECM runs the pass that closes an
 *        iteration ahead into the spare set).  gathered_fwd of backward_replay: the forward payloads gathered
 *        for the forward pass that wrote `set`. */
#define CB200_SPLIT_PAYLOAD 16
CB200_API int cb200_split_begin(cb200_ctx *ctx, const cb200_model *model, double nu, const float *data,
                                const float *munc, int64_t m, int64_t n, int64_t ld, const float *qscale,
                                const float *kap, int32_t is_first, int32_t is_last);
CB200_API int cb200_split_forward_compose(cb200_ctx *ctx, double *payload);
CB200_API int cb200_split_forward_replay(cb200_ctx *ctx, const double *gathered, int32_t rank, int32_t world,
                                         int32_t with_nll, int32_t store, int32_t set, double *sums);
CB200_API int cb200_split_backward_compose(cb200_ctx *ctx, int32_t set, double *payload);
CB200_API int cb200_split_backward_replay(cb200_ctx *ctx, const double *gathered_bwd, const double *gathered_fwd,
                                          int32_t rank, int32_t world, int32_t set, int32_t publish, float *xs,
                                          float *Ps, float *lag);
/* writes the shard's kappa back in interval order (kap may be NULL), waits for the stream, ends the split */
CB200_API int cb200_split_end(cb200_ctx *ctx, float *kap);

/* ---- output: bedGraph text ----------------------------------------------
 * The rows the reference appends per chromosome and track with pandas (consenrich.py:9797-9805:
 * to_csv(sep="\t", header=False, index=False, float_format="%.4f", lineterminator="\n")):
 * "chrom\tstart\tend\tvalue\n", value = C's "%.4f" of the float32 track value (NaN -> empty, +-inf ->
 * "inf" / "-inf"), formatted on the device, byte-identical to the reference writer.
 * starts / ends: [n] int64 or NULL (NULL: start = start0 + k step, end = start + step, clipped to end_clip
 * when end_clip > 0).  values: row k at values[k * value_stride].  chrom: at most 32 bytes.
 * cb200_bedgraph_chunk: device arrays; *text receives a DEVICE pointer owned by the context.
 * cb200_host_bedgraph_chunk: host arrays; *text receives a page-locked HOST pointer owned by the context.
 * Either pointer stays valid until the next bedGraph call on the context; *bytes is the text length. */
CB200_API int cb200_bedgraph_chunk(cb200_ctx *ctx, const char *chrom, int64_t n, const int64_t *starts,
                                   const int64_t *ends, int64_t start0, int64_t step, int64_t end_clip,
                                   const float *values, int64_t value_stride, const char **text, int64_t *bytes);
CB200_API int cb200_host_bedgraph_chunk(cb200_ctx *ctx, const char *chrom, int64_t n, const int64_t *starts,
                                        const int64_t *ends, int64_t start0, int64_t step, int64_t end_clip,
                                        const float *values, int64_t value_stride, const char **text, int64_t *bytes);

#ifdef __cplusplus
}
#endif
#endif /* CONSENRICH_B200_H */
