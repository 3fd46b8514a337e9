/*
 * oracle/background_oracle.c -- TEST INFRASTRUCTURE ONLY (checker; never measured as product,
 * never linked into libconsenrich_b200.so).
 *
 * Sequential CPU restatement of the reference's background-track kernels, written from the
 * algorithm in /root/reference/src/consenrich/cconsenrich.pyx, same arithmetic order, so that it
 * agrees with the reference build (oracle/_ref) to the last bit; tests/test_oracle_pinning.py pins
 * that against oracle/_ref and tests/golden/background_golden.npz.
 */
#include <math.h>
#include <stddef.h>
#include <stdint.h>

/* cconsenrich.pyx:905-943: entries of lamFirst D1'D1 + lam D2'D2 */
static double second_diag(int64_t n, int64_t i, double lam) {
    if (n < 3 || lam <= 0.0) return 0.0;
    if (n == 3) return i == 1 ? 4.0 * lam : lam;
    if (i == 0 || i == n - 1) return lam;
    if (i == 1 || i == n - 2) return 5.0 * lam;
    return 6.0 * lam;
}
static double second_off1(int64_t n, int64_t i, double lam) {
    if (n < 3 || lam <= 0.0) return 0.0;
    if (n == 3) return -2.0 * lam;
    if (i == 0 || i == n - 2) return -2.0 * lam;
    return -4.0 * lam;
}
static double first_diag(int64_t n, int64_t i, double lam) {
    if (n < 2 || lam <= 0.0) return 0.0;
    if (i == 0 || i == n - 1) return lam;
    return 2.0 * lam;
}
static double first_off1(int64_t n, double lam) {
    if (n < 2 || lam <= 0.0) return 0.0;
    return -lam;
}

/* cbackgroundWeightedStats[WithSupport], cconsenrich.pyx:9675-9724.  resid, inv: float32 [m][n]. */
int64_t bg_weighted_stats(const float *resid, const float *inv, int64_t m, int64_t n, double *weight, double *rhs) {
    int64_t support = 0;
    for (int64_t i = 0; i < n; ++i) {
        double wsum = 0.0, rsum = 0.0;
        for (int64_t j = 0; j < m; ++j) {
            const double w = (double)inv[j * n + i];
            wsum += w;
            rsum += w * (double)resid[j * n + i];
        }
        weight[i] = wsum;
        rhs[i] = rsum;
        if (wsum > 0.0) support += 1;
    }
    return support;
}

/* csolveZeroCenteredBackground, cconsenrich.pyx:944-1096 (n >= 2; the wrapper handles n < 2).
 * diag, rhs: in = weightTrack / rhsTrack copies, used as scratch; cons, lower: scratch [n].
 * Returns the first index whose pivot was raised to the floor, or -1; *bad_value = that pivot. */
int64_t bg_solve(double *diag, double *rhs, double *cons, double *lower, double *out, int64_t n, double lam,
                 double lam_first, int zero_center, double *bad_value) {
    const double min_pivot = 1.0e-12;
    int64_t bad = -1;
    double off, l2;
    for (int64_t i = 0; i < n; ++i) {
        cons[i] = 1.0;
        lower[i] = 0.0;
    }
    for (int64_t i = 0; i < n; ++i) { /* :1018-1031 */
        diag[i] = diag[i] + first_diag(n, i, lam_first) + second_diag(n, i, lam);
        if (diag[i] < min_pivot) {
            if (bad < 0) { bad = i; *bad_value = diag[i]; }
            diag[i] = min_pivot;
        }
    }
    /* pentadiagonal LDL', :1036-1060 */
    off = first_off1(n, lam_first) + second_off1(n, 0, lam);
    lower[1] = off / diag[0];
    diag[1] = diag[1] - lower[1] * lower[1] * diag[0];
    if (diag[1] < min_pivot) {
        if (bad < 0) { bad = 1; *bad_value = diag[1]; }
        diag[1] = min_pivot;
    }
    for (int64_t i = 2; i < n; ++i) {
        off = first_off1(n, lam_first) + second_off1(n, i - 1, lam);
        lower[i] = (off - lam * lower[i - 1]) / diag[i - 1];
        diag[i] = diag[i] - lower[i] * lower[i] * diag[i - 1] - (lam * lam) / diag[i - 2];
        if (diag[i] < min_pivot) {
            if (bad < 0) { bad = i; *bad_value = diag[i]; }
            diag[i] = min_pivot;
        }
    }
    /* forward solve, :1064-1069 */
    rhs[1] = rhs[1] - lower[1] * rhs[0];
    cons[1] = cons[1] - lower[1] * cons[0];
    for (int64_t i = 2; i < n; ++i) {
        l2 = lam / diag[i - 2];
        rhs[i] = rhs[i] - lower[i] * rhs[i - 1] - l2 * rhs[i - 2];
        cons[i] = cons[i] - lower[i] * cons[i - 1] - l2 * cons[i - 2];
    }
    for (int64_t i = 0; i < n; ++i) { /* :1071-1073 */
        rhs[i] = rhs[i] / diag[i];
        cons[i] = cons[i] / diag[i];
    }
    /* backward solve, :1076-1081 */
    rhs[n - 2] = rhs[n - 2] - lower[n - 1] * rhs[n - 1];
    cons[n - 2] = cons[n - 2] - lower[n - 1] * cons[n - 1];
    for (int64_t i = n - 3; i >= 0; --i) {
        l2 = lam / diag[i];
        rhs[i] = rhs[i] - lower[i + 1] * rhs[i + 1] - l2 * rhs[i + 2];
        cons[i] = cons[i] - lower[i + 1] * cons[i + 1] - l2 * cons[i + 2];
    }
    if (zero_center) { /* :1083-1093 */
        double sum_rhs = 0.0, sum_cons = 0.0, mu;
        for (int64_t i = 0; i < n; ++i) {
            sum_rhs += rhs[i];
            sum_cons += cons[i];
        }
        mu = (fabs(sum_cons) > min_pivot) ? sum_rhs / sum_cons : sum_rhs / (double)n;
        for (int64_t i = 0; i < n; ++i) out[i] = rhs[i] - mu * cons[i];
    } else {
        for (int64_t i = 0; i < n; ++i) out[i] = rhs[i];
    }
    return bad;
}

/* The matrix half of core._relativeSignChangePerKB, core.py:2670-2696 (numpy float64 operations, one row
 * of the matrices at a time): out[k] = state[k] - weighted mean over the valid cells of interval k. */
void bg_weighted_mean_residual(const float *data, const float *munc, int64_t m, int64_t n, const double *state,
                               const double *background, double pad, double *out) {
    for (int64_t k = 0; k < n; ++k) {
        const double sv = state[k], bg = background ? background[k] : 0.0;
        double total = 0.0, wsum = 0.0, mean;
        for (int64_t j = 0; j < m; ++j) {
            const double d = (double)data[j * n + k], den = (double)munc[j * n + k] + pad;
            if (isfinite(sv) && isfinite(d) && isfinite(den) && den > 0.0) {
                const double w = 1.0 / (den > 1.0e-12 ? den : 1.0e-12);
                total += (d - bg) * w;
                wsum += w;
            }
        }
        mean = wsum > 0.0 ? total / wsum : NAN;
        out[k] = sv - mean;
    }
}

/* core._perIntervalOutputDiagnosticTracks, core.py:7786-7800: per-interval sums over the tracks of the effective
 * observation variance and of its inverse, finite terms only. */
void bg_diag_obs_sums(const float *munc, int64_t m, int64_t n, const double *obs_prec, double pad, double *munc_trace,
                      double *sum_inv_r) {
    for (int64_t k = 0; k < n; ++k) {
        double tr = 0.0, si = 0.0;
        for (int64_t j = 0; j < m; ++j) {
            double v = (double)munc[j * n + k] + pad, eff, inv;
            if (!(v >= 1.0e-12)) v = v != v ? v : 1.0e-12; /* np.maximum keeps NaN */
            eff = v / obs_prec[k];
            inv = obs_prec[k] / v;
            if (isfinite(eff)) tr += eff;
            if (isfinite(inv)) si += inv;
        }
        munc_trace[k] = tr;
        sum_inv_r[k] = si;
    }
}

/* the per-interval loop of core.py:7840-7866.  covar, p_noise: float32 [n][c][c]; p_noise / proc_prec may be NULL */
void bg_diag_gain(const float *covar, const float *p_noise, const double *q_scale, const double *proc_prec,
                  const double *sum_inv_r, int64_t n, int dim, int c, const double *base_q, const double *f,
                  double cov_init, double *gain0, double *gain1) {
    double prev[4] = {cov_init, 0.0, 0.0, cov_init};
    for (int64_t k = 0; k < n; ++k) {
        double q[4] = {0, 0, 0, 0}, pred00, pred10 = 0.0, denom;
        int from_track = 0;
        if (p_noise && k > 0) {
            const float *pn = p_noise + (k - 1) * c * c;
            int fin = isfinite((double)pn[0]);
            if (dim == 2) fin = fin && isfinite((double)pn[1]) && isfinite((double)pn[c]) && isfinite((double)pn[c + 1]);
            if (fin) {
                from_track = 1;
                q[0] = (double)pn[0];
                if (dim == 2) { q[1] = (double)pn[1]; q[2] = (double)pn[c]; q[3] = (double)pn[c + 1]; }
            }
        }
        if (!from_track) {
            for (int i = 0; i < dim * dim; ++i) q[i] = base_q[i] * q_scale[k];
            if (!(p_noise && k > 0) && proc_prec)
                for (int i = 0; i < dim * dim; ++i) q[i] /= proc_prec[k];
        }
        if (dim == 2) {
            const double t00 = f[0] * prev[0] + f[1] * prev[2], t01 = f[0] * prev[1] + f[1] * prev[3];
            const double t10 = f[2] * prev[0] + f[3] * prev[2], t11 = f[2] * prev[1] + f[3] * prev[3];
            pred00 = (t00 * f[0] + t01 * f[1]) + q[0];
            pred10 = (t10 * f[0] + t11 * f[1]) + q[2];
        } else {
            pred00 = prev[0] + q[0];
        }
        if (!(pred00 > 0.0)) pred00 = pred00 != pred00 ? pred00 : 0.0;
        denom = 1.0 + pred00 * sum_inv_r[k];
        gain0[k] = gain1[k] = 0.0;
        if (isfinite(denom) && denom > 0.0) {
            const double gs = sum_inv_r[k] / denom;
            gain0[k] = pred00 * gs;
            gain1[k] = pred10 * gs;
        }
        prev[0] = (double)covar[k * c * c];
        if (dim == 2) {
            prev[1] = (double)covar[k * c * c + 1];
            prev[2] = (double)covar[k * c * c + c];
            prev[3] = (double)covar[k * c * c + c + 1];
        }
    }
}
