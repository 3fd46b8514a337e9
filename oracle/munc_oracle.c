/*
 * oracle/munc_oracle.c -- TEST INFRASTRUCTURE ONLY (checker; never measured as product, never
 * linked into libconsenrich_b200.so).
 *
 * Sequential CPU restatement of the dense kernels of the reference's observation-noise (MUNC)
 * stage, written from the algorithm in /root/reference/src/consenrich/cconsenrich.pyx with the
 * same arithmetic order; pinned bit-exact against oracle/_ref and tests/golden/munc_golden.npz by
 * tests/test_munc_oracle.py.
 */
#include <math.h>
#include <stddef.h>
#include <stdint.h>

/* _muncSeedMaskAllowsCell with nonzeroMeansActive = False, cconsenrich.pyx:4746-4766 */
static int allows(const uint8_t *mask, int mode, int64_t n, int64_t j, int64_t k) {
    uint8_t v;
    if (mode == 0) return 1;
    v = mode == 1 ? mask[k] : mask[j * n + k];
    return v == 0;
}

/* _muncSmoothDenseLocalEvidenceInvalidIndex, :5547-5574: first unmasked cell that is not positive
 * and finite, or -1 */
int64_t munc_smooth_invalid_index(const float *local, const uint8_t *mask, int mode, int64_t m, int64_t n) {
    for (int64_t j = 0; j < m; ++j)
        for (int64_t i = 0; i < n; ++i) {
            double v;
            if (!allows(mask, mode, n, j, i)) continue;
            v = (double)local[j * n + i];
            if (!isfinite(v) || v <= 0.0) return j * n + i;
        }
    return -1;
}

/* _muncSmoothDenseLocalEvidenceRow for every row, :5577-5639 */
void munc_smooth_rows(const float *local, const uint8_t *mask, int mode, int64_t m, int64_t n, int64_t window,
                      double eps, float *out) {
    const int64_t half = window / 2;
    for (int64_t j = 0; j < m; ++j) {
        const int64_t row = j * n;
        int64_t left = 0, right = 0, count = 0;
        double rolling = 0.0;
        for (int64_t i = 0; i < n; ++i) {
            int64_t tl = i >= half ? i - half : 0;
            int64_t tr = tl + window;
            double v;
            if (tr > n) {
                tr = n;
                tl = tr >= window ? tr - window : 0;
            }
            while (right < tr) {
                if (allows(mask, mode, n, j, right)) {
                    rolling += (double)local[row + right];
                    count += 1;
                }
                right += 1;
            }
            while (left < tl) {
                if (allows(mask, mode, n, j, left)) {
                    rolling -= (double)local[row + left];
                    count -= 1;
                }
                left += 1;
            }
            v = count > 0 ? rolling / (double)count : (double)local[row + i];
            if (v < eps) v = eps;
            out[row + i] = (float)v;
        }
    }
}

/* _finalizeMuncEBTrackLoop, cconsenrich.pyx:5365-5440.  counters: [support, cfFinite, cfAdded, cfMissing];
 * invalid: [local, prior, countFloor] (first offending index of the kind that stopped the loop, others -1) */
void munc_finalize_eb(const float *local, const float *prior, const float *cfloor, float *out, int64_t n,
                      double nu_local, double nu_prior, double post, double vfloor, double vcap, int use_eb,
                      int64_t *counters, int64_t *invalid) {
    counters[0] = counters[1] = counters[2] = counters[3] = 0;
    invalid[0] = invalid[1] = invalid[2] = -1;
    for (int64_t i = 0; i < n; ++i) {
        double lv = (double)local[i], ov;
        if (!isfinite(lv) || lv <= 0.0) { invalid[0] = i; return; }
        if (lv > vfloor) counters[0] += 1;
        if (lv < vfloor) lv = vfloor;
        else if (lv > vcap) lv = vcap;
        if (use_eb) {
            double pv = (double)prior[i];
            if (!isfinite(pv) || pv <= 0.0) { invalid[1] = i; return; }
            if (pv < vfloor) pv = vfloor;
            else if (pv > vcap) pv = vcap;
            ov = ((nu_local * lv) + (nu_prior * pv)) / post;
        } else {
            ov = lv;
        }
        if (ov < vfloor) ov = vfloor;
        else if (ov > vcap) ov = vcap;
        if (cfloor) {
            const double cv = (double)cfloor[i];
            if (cv == cv) {
                if (!isfinite(cv) || cv < 0.0) { invalid[2] = i; return; }
                counters[1] += 1;
                ov += cv;
                if (cv > 0.0) counters[2] += 1;
                if (ov < vfloor) ov = vfloor;
                else if (ov > vcap) ov = vcap;
            } else {
                counters[3] += 1;
            }
        }
        out[i] = (float)ov;
    }
}

/* ---- cMuncObservationMomentSeedPass, cconsenrich.pyx:4767-5040 ---- */
typedef struct {
    const float *data, *munc, *state_mean, *state_var, *background, *g_var, *count_floor, *omega_in, *rho_in;
    const uint8_t *active;
    float *moment, *rho_out, *omega_raw, *omega_out, *local, *variance;
    int64_t m, n;
    int32_t active_mode, use_weights, student_t, update_weights;
    double pad, d_s, d_omega, omega_min, omega_max, var_floor, var_cap;
} seed_args;

static int seed_active(const seed_args *a, int64_t j, int64_t k) {
    /* _muncSeedMaskAllowsCell with nonzeroMeansActive = True */
    if (a->active_mode == 0) return 1;
    return (a->active_mode == 1 ? a->active[k] : a->active[j * a->n + k]) != 0;
}

static double clamp_mult(double v, double lo, double hi) { /* :135-140 */
    if (v < lo) return lo;
    if (v > hi) return hi;
    return v;
}

/* _muncObservationMomentSeedInvalidIndex, :4767-4840 */
int64_t munc_seed_invalid_index(const seed_args *a) {
    const int weighted = a->use_weights && a->student_t;
    for (int64_t j = 0; j < a->m; ++j)
        for (int64_t k = 0; k < a->n; ++k) {
            int64_t idx = j * a->n + k;
            double v;
            if (!seed_active(a, j, k)) continue;
            if (!isfinite((double)a->state_mean[k])) return idx;
            if (!isfinite((double)a->state_var[k])) return idx;
            if (a->background && !isfinite((double)a->background[k])) return idx;
            if (a->g_var && !isfinite((double)a->g_var[k])) return idx;
            if (!isfinite((double)a->data[idx])) return idx;
            v = (double)a->munc[idx] + a->pad;
            if (!isfinite(v) || v <= 0.0) return idx;
            if (a->count_floor) {
                v = (double)a->count_floor[idx];
                if (!isfinite(v) || v < 0.0) return idx;
            }
            if (weighted) {
                if (a->omega_in) {
                    v = (double)a->omega_in[k];
                    if (!isfinite(v) || v <= 0.0) return idx;
                }
                if (!a->update_weights) {
                    v = (double)a->rho_in[idx];
                    if (!isfinite(v) || v <= 0.0) return idx;
                }
            }
        }
    return -1;
}

/* _muncObservationMomentSeedPassInterval for every interval, :4843-5040 */
void munc_seed_pass(const seed_args *a) {
    const int weighted = a->use_weights && a->student_t;
    for (int64_t k = 0; k < a->n; ++k) {
        int64_t active_count = 0;
        double state = (double)a->state_mean[k], mvb = (double)a->state_var[k];
        double bg = 0.0, cv, base, res, mom, rho = 1.0, omega_in = 1.0, omega_raw, omega = 1.0, dbar, lv, tv;
        if (a->background) bg = (double)a->background[k];
        if (a->g_var) mvb += (double)a->g_var[k];
        if (mvb < 0.0) mvb = 0.0;
        if (weighted) {
            omega_in = a->omega_in ? (double)a->omega_in[k] : 1.0;
            if (a->update_weights) {
                dbar = 0.0;
                for (int64_t j = 0; j < a->m; ++j) {
                    int64_t idx = j * a->n + k;
                    if (!seed_active(a, j, k)) {
                        a->moment[idx] = 0.0f;
                        a->rho_out[idx] = 1.0f;
                        continue;
                    }
                    base = (double)a->munc[idx] + a->pad;
                    if (base < a->var_floor) base = a->var_floor;
                    res = (double)a->data[idx] - bg - state;
                    mom = res * res + mvb;
                    rho = (a->d_s + 1.0) / (a->d_s + omega_in * mom / base);
                    a->moment[idx] = (float)mom;
                    a->rho_out[idx] = (float)rho;
                    dbar += mom / base;
                    active_count += 1;
                }
                if (active_count > 0) {
                    dbar = dbar / (double)active_count;
                    omega_raw = (a->d_omega + 1.0) / (a->d_omega + dbar);
                    omega = clamp_mult(omega_raw, a->omega_min, a->omega_max);
                } else {
                    omega_raw = 1.0;
                    omega = 1.0;
                }
            } else {
                omega_raw = omega_in;
                omega = clamp_mult(omega_raw, a->omega_min, a->omega_max);
                for (int64_t j = 0; j < a->m; ++j) {
                    int64_t idx = j * a->n + k;
                    if (!seed_active(a, j, k)) {
                        a->moment[idx] = 0.0f;
                        a->rho_out[idx] = 1.0f;
                        continue;
                    }
                    res = (double)a->data[idx] - bg - state;
                    mom = res * res + mvb;
                    rho = (double)a->rho_in[idx];
                    a->moment[idx] = (float)mom;
                    a->rho_out[idx] = (float)rho;
                }
            }
            a->omega_raw[k] = (float)omega_raw;
            a->omega_out[k] = (float)omega;
        } else {
            a->omega_raw[k] = 1.0f;
            a->omega_out[k] = 1.0f;
        }
        for (int64_t j = 0; j < a->m; ++j) {
            int64_t idx = j * a->n + k;
            cv = a->count_floor ? (double)a->count_floor[idx] : 0.0;
            if (seed_active(a, j, k)) {
                if (!weighted) {
                    res = (double)a->data[idx] - bg - state;
                    mom = res * res + mvb;
                    a->moment[idx] = (float)mom;
                    a->rho_out[idx] = 1.0f;
                    lv = mom - a->pad - cv;
                } else {
                    mom = (double)a->moment[idx];
                    rho = (double)a->rho_out[idx];
                    lv = omega * rho * mom - a->pad - cv;
                }
                tv = lv + cv;
                if (lv < a->var_floor) {
                    lv = a->var_floor;
                    tv = lv + cv;
                }
                if (tv > a->var_cap) {
                    tv = a->var_cap;
                    lv = tv - cv;
                    if (lv < a->var_floor) {
                        lv = a->var_floor;
                        tv = lv + cv;
                    }
                }
            } else {
                lv = (double)a->munc[idx] - cv;
                if (lv < a->var_floor) lv = a->var_floor;
                tv = lv + cv;
                if (tv > a->var_cap) {
                    tv = a->var_cap;
                    lv = tv - cv;
                    if (lv < a->var_floor) {
                        lv = a->var_floor;
                        tv = lv + cv;
                    }
                }
                a->moment[idx] = 0.0f;
                a->rho_out[idx] = 1.0f;
            }
            a->local[idx] = (float)lv;
            a->variance[idx] = (float)tv;
        }
    }
}

/* _cEMA, cconsenrich.pyx:5744-5759, for real_t = float and real_t = double (the C Cython emits: alpha * x
 * in the track's type, (1.0 - alpha) in double).  Returns 1 without writing when alpha is outside [0, 1]. */
int munc_ema_f32(const float *x, float *out, int64_t n, float alpha) {
    if (alpha > 1.0f || alpha < 0.0f) return 1;
    out[0] = x[0];
    for (int64_t i = 1; i < n; ++i) out[i] = alpha * x[i] + (1.0 - alpha) * out[i - 1];
    for (int64_t i = n - 2; i >= 0; --i) out[i] = alpha * out[i] + (1.0 - alpha) * out[i + 1];
    return 0;
}
int munc_ema_f64(const double *x, double *out, int64_t n, double alpha) {
    if (alpha > 1.0 || alpha < 0.0) return 1;
    out[0] = x[0];
    for (int64_t i = 1; i < n; ++i) out[i] = alpha * x[i] + (1.0 - alpha) * out[i - 1];
    for (int64_t i = n - 2; i >= 0; --i) out[i] = alpha * out[i] + (1.0 - alpha) * out[i + 1];
    return 0;
}
