/*
 * oracle/ssm_oracle.c -- TEST INFRASTRUCTURE ONLY (checker; never measured as product,
 * never linked into libconsenrich_b200.so).
 *
 * Sequential CPU restatement of the reference's state-space hot path, written from the
 * algorithm in /root/reference/src/consenrich/cconsenrich.pyx.  It keeps the reference's
 * arithmetic ORDER and its float32 rounding points so that it agrees with the reference
 * build (oracle/_ref) to the last bit on the same inputs; tests/test_oracle_pinning.py
 * pins that claim against oracle/_ref and against tests/golden/.
 *
 * Each function cites the reference lines it restates.
 *
 * Layouts (all C-contiguous):
 *   data, munc : float32 [m][n]      (m samples/tracks, n bins/intervals)
 *   xf / xs    : float32 [n][d]      d = 2 (level+trend) or 1 (level)
 *   Pf / Ps    : float32 [n][d][d]
 *   Qf         : float32 [>=n-1][d][d]   (Q_k stored at index k-1)
 *   lagC       : float32 [max(n-1,1)][d][d]
 *   resid      : float32 [n][m]      (transposed w.r.t. data)
 */
#include <math.h>
#include <stddef.h>
#include <stdint.h>

#define R32(x) ((double)(float)(x))

typedef struct {
    double state_init, cov_init, pad; /* already rounded to C float by the caller */
    double F[4];                      /* row-major 2x2 (unused by level model)  */
    double Q0[4];                     /* row-major 2x2; level model uses Q0[0]  */
    double lam_min, lam_max;          /* observation precision multiplier clamp */
    double kap_min, kap_max;          /* process precision multiplier clamp     */
    double apn_min_q, apn_max_q, apn_thresh, apn_scale, apn_pc;
    int32_t use_lambda, use_kappa, use_qscale, use_apn;
    int32_t return_nll, store_nll_in_d, do_store, pad_;
} oracle_params;

static inline double clampd(double v, double lo, double hi) {
    /* cconsenrich.pyx:135-140 */
    if (v < lo) return lo;
    if (v > hi) return hi;
    return v;
}

/* cconsenrich.pyx:259-283 (_accumulateObservationValue), called per (bin, sample). */
static inline void fold_one(double z, double level, double v, double pad, double lam,
                            int want_nll, double *s0, double *s1, double *s2, double *sl) {
    double e = z - level;
    double r = v + pad;
    if (r < 1.0e-12) r = 1.0e-12;
    double w = lam / r;
    if (want_nll) *sl += (log(r) - log(lam));
    *s2 += w * (e * e);
    *s1 += w * e;
    *s0 += w;
}

/*
 * 2-state (level/trend) forward filter.  cconsenrich.pyx:291-529.
 * Returns -1 on success or the index of the first out-of-range block id (pyx:389-392).
 */
int64_t oracle_forward2(const float *data, const float *munc, int64_t m, int64_t n,
                        const int32_t *block_map, int64_t block_count, const float *lam,
                        const float *kap, const float *qscale, const oracle_params *p,
                        float *D, float *xf, float *Pf, float *Qf, double *sum_d,
                        double *sum_nll) {
    const double F00 = p->F[0], F01 = p->F[1], F10 = p->F[2], F11 = p->F[3];
    const double q00 = p->Q0[0], q01 = p->Q0[1], q10 = p->Q0[2], q11 = p->Q0[3];
    const double q_diag = 0.5 * (q00 + q11);
    const double LOG2PI = log(6.2831853071795864769);
    double x0 = R32(p->state_init), x1 = 0.0;
    double P00 = R32(p->cov_init), P01 = 0.0, P10 = 0.0, P11 = R32(p->cov_init);
    double apn = 1.0;
    int use_apn = p->use_apn && (q_diag > 1.0e-12); /* pyx:6574-6576 */
    *sum_d = 0.0;
    *sum_nll = 0.0;
    for (int64_t k = 0; k < n; ++k) {
        int64_t b = (int64_t)block_map[k];
        if (b < 0 || b >= block_count) return k;
        double kappa = p->use_kappa ? clampd((double)kap[k], p->kap_min, p->kap_max) : 1.0;
        /* predict mean, round to float32 (pyx:403-406) */
        double xp0 = F00 * x0 + F01 * x1;
        double xp1 = F10 * x0 + F11 * x1;
        x0 = R32(xp0);
        x1 = R32(xp1);
        double qs = p->use_qscale ? (double)qscale[k] : apn;
        double Q00 = (qs / kappa) * q00, Q01 = (qs / kappa) * q01;
        double Q10 = (qs / kappa) * q10, Q11 = (qs / kappa) * q11;
        /* predict covariance, round to float32 (pyx:417-430) */
        double t00 = F00 * P00 + F01 * P10, t01 = F00 * P01 + F01 * P11;
        double t10 = F10 * P00 + F11 * P10, t11 = F10 * P01 + F11 * P11;
        P00 = R32(t00 * F00 + t01 * F01 + Q00);
        P01 = R32(t00 * F10 + t01 * F11 + Q01);
        P10 = R32(t10 * F00 + t11 * F01 + Q10);
        P11 = R32(t10 * F10 + t11 * F11 + Q11);
        double lambda = p->use_lambda ? clampd((double)lam[k], p->lam_min, p->lam_max) : 1.0;
        double s0 = 0.0, s1 = 0.0, s2 = 0.0, sl = 0.0, nll_k = 0.0;
        for (int64_t j = 0; j < m; ++j) {
            size_t idx = (size_t)j * (size_t)n + (size_t)k;
            fold_one((double)data[idx], x0, (double)munc[idx], p->pad, lambda, p->return_nll,
                     &s0, &s1, &s2, &sl);
        }
        /* scalar-innovation update (pyx:458-495) */
        double innov_scale = 1.0 + P00 * s0;
        double gain_like = P00 / innov_scale;
        double quad = s2 - gain_like * (s1 * s1);
        if (quad < 0.0) quad = 0.0;
        if (p->return_nll) {
            nll_k = 0.5 * (sl + log(innov_scale) + quad + ((double)m) * LOG2PI);
            *sum_nll += nll_k;
        }
        double stat = (p->return_nll && p->store_nll_in_d) ? nll_k : quad / ((double)m);
        D[k] = (float)stat;
        *sum_d += (double)D[k];
        double delta0 = s1 / innov_scale;
        x0 = R32(x0 + P00 * delta0);
        x1 = R32(x1 + P10 * delta0);
        double gG = s0 / innov_scale;
        double gH = s0 / (innov_scale * innov_scale);
        double I00 = 1.0 - (P00 * gG);
        double I10 = -(P10 * gG);
        double n00 = (I00 * I00 * P00) + (gH * (P00 * P00));
        double n01 = (I00 * (I10 * P00 + P01)) + (gH * (P00 * P10));
        double n11 = ((I10 * I10 * P00) + 2.0 * I10 * P10 + P11) + (gH * (P10 * P10));
        P00 = R32(n00);
        P01 = R32(n01);
        P10 = P01;
        P11 = R32(n11);
        if (p->do_store) {
            xf[k * 2] = (float)x0;
            xf[k * 2 + 1] = (float)x1;
            Pf[k * 4] = (float)P00;
            Pf[k * 4 + 1] = (float)P01;
            Pf[k * 4 + 2] = (float)P10;
            Pf[k * 4 + 3] = (float)P11;
            if (k > 0) {
                Qf[(k - 1) * 4] = (float)Q00;
                Qf[(k - 1) * 4 + 1] = (float)Q01;
                Qf[(k - 1) * 4 + 2] = (float)Q10;
                Qf[(k - 1) * 4 + 3] = (float)Q11;
            }
        }
        /* adaptive process noise feedback (pyx:510-527) */
        if (use_apn && !p->use_qscale) {
            double pn = 0.5 * (Q00 + Q11);
            double dk = (double)D[k];
            if (D[k] > p->apn_thresh && pn < p->apn_max_q) {
                apn *= sqrt(p->apn_scale * (dk - p->apn_thresh) + p->apn_pc);
            } else if (D[k] <= p->apn_thresh && pn > p->apn_min_q) {
                apn *= 1.0 / sqrt(p->apn_scale * (p->apn_thresh - dk) + p->apn_pc);
            }
            pn = apn * q_diag;
            if (pn < p->apn_min_q)
                apn = p->apn_min_q / q_diag;
            else if (pn > p->apn_max_q)
                apn = p->apn_max_q / q_diag;
        }
    }
    return -1;
}

/* 1-state (level) forward filter.  cconsenrich.pyx:538-707.  No float32 rounding of the
 * carried state in this variant. */
int64_t oracle_forward1(const float *data, const float *munc, int64_t m, int64_t n,
                        const int32_t *block_map, int64_t block_count, const float *lam,
                        const float *kap, const float *qscale, const oracle_params *p,
                        float *D, float *xf, float *Pf, float *Qf, double *sum_d,
                        double *sum_nll) {
    const double q0 = p->Q0[0];
    const double LOG2PI = log(6.2831853071795864769);
    double x = p->state_init, P = p->cov_init, apn = 1.0;
    int use_apn = p->use_apn && (q0 > 1.0e-12); /* pyx:6997-6998 */
    *sum_d = 0.0;
    *sum_nll = 0.0;
    for (int64_t k = 0; k < n; ++k) {
        int64_t b = (int64_t)block_map[k];
        if (b < 0 || b >= block_count) return k;
        double kappa = p->use_kappa ? clampd((double)kap[k], p->kap_min, p->kap_max) : 1.0;
        double qs = p->use_qscale ? (double)qscale[k] : apn;
        double Q = (qs / kappa) * q0;
        P += Q;
        double lambda = p->use_lambda ? clampd((double)lam[k], p->lam_min, p->lam_max) : 1.0;
        double s0 = 0.0, s1 = 0.0, s2 = 0.0, sl = 0.0, nll_k = 0.0;
        for (int64_t j = 0; j < m; ++j) {
            size_t idx = (size_t)j * (size_t)n + (size_t)k;
            fold_one((double)data[idx], x, (double)munc[idx], p->pad, lambda, p->return_nll,
                     &s0, &s1, &s2, &sl);
        }
        double innov_scale = 1.0 + P * s0;
        double gain_like = P / innov_scale;
        double quad = s2 - gain_like * (s1 * s1);
        if (quad < 0.0) quad = 0.0;
        if (p->return_nll) {
            nll_k = 0.5 * (sl + log(innov_scale) + quad + ((double)m) * LOG2PI);
            *sum_nll += nll_k;
        }
        double stat = (p->return_nll && p->store_nll_in_d) ? nll_k : quad / ((double)m);
        D[k] = (float)stat;
        *sum_d += (double)D[k];
        double delta0 = s1 / innov_scale;
        x += P * delta0;
        double gG = s0 / innov_scale;
        double gH = s0 / (innov_scale * innov_scale);
        double IKH = 1.0 - P * gG;
        P = (IKH * IKH * P) + (gH * (P * P));
        if (p->do_store) {
            xf[k] = (float)x;
            Pf[k] = (float)P;
            if (k > 0) Qf[k - 1] = (float)Q;
        }
        if (use_apn && !p->use_qscale) {
            double pn = apn * q0;
            double dk = (double)D[k];
            if (D[k] > p->apn_thresh && pn < p->apn_max_q) {
                apn *= sqrt(p->apn_scale * (dk - p->apn_thresh) + p->apn_pc);
            } else if (D[k] <= p->apn_thresh && pn > p->apn_min_q) {
                apn *= 1.0 / sqrt(p->apn_scale * (p->apn_thresh - dk) + p->apn_pc);
            }
            pn = apn * q0;
            if (pn < p->apn_min_q)
                apn = p->apn_min_q / q0;
            else if (pn > p->apn_max_q)
                apn = p->apn_max_q / q0;
        }
    }
    return -1;
}

/* 2-state RTS smoother + lag-one covariance + residuals.  cconsenrich.pyx:6740-6848.
 * lag_rows = number of rows available in lagC (pyx:6840 guard). */
void oracle_backward2(const float *data, int64_t m, int64_t n, const double *F, const float *xf,
                      const float *Pf, const float *Qf, float *xs, float *Ps, float *lagC,
                      int64_t lag_rows, float *resid) {
    if (n <= 0) return;
    const double F00 = F[0], F01 = F[1], F10 = F[2], F11 = F[3];
    int64_t last = n - 1;
    xs[last * 2] = xf[last * 2];
    xs[last * 2 + 1] = xf[last * 2 + 1];
    for (int c = 0; c < 4; ++c) Ps[last * 4 + c] = Pf[last * 4 + c];
    for (int64_t j = 0; j < m; ++j)
        resid[last * m + j] = (float)((double)data[j * n + last] - (double)xs[last * 2]);
    for (int64_t k = n - 2; k >= 0; --k) {
        double Pf00 = Pf[k * 4], Pf01 = Pf[k * 4 + 1], Pf10 = Pf[k * 4 + 2], Pf11 = Pf[k * 4 + 3];
        double xk0 = xf[k * 2], xk1 = xf[k * 2 + 1];
        double xp0 = F00 * xk0 + F01 * xk1;
        double xp1 = F10 * xk0 + F11 * xk1;
        double Q00 = Qf[k * 4], Q01 = Qf[k * 4 + 1], Q10 = Qf[k * 4 + 2], Q11 = Qf[k * 4 + 3];
        double c00 = F00 * Pf00 + F01 * Pf10, c01 = F00 * Pf01 + F01 * Pf11;
        double c10 = F10 * Pf00 + F11 * Pf10, c11 = F10 * Pf01 + F11 * Pf11;
        double PP00 = c00 * F00 + c01 * F01 + Q00, PP01 = c00 * F10 + c01 * F11 + Q01;
        double PP10 = c10 * F00 + c11 * F01 + Q10, PP11 = c10 * F10 + c11 * F11 + Q11;
        double det = (PP00 * PP11) - (PP01 * PP10);
        double i00 = PP11 / det, i01 = -PP01 / det, i10 = -PP10 / det, i11 = PP00 / det;
        /* P_f F^T */
        c00 = Pf00 * F00 + Pf01 * F01;
        c01 = Pf00 * F10 + Pf01 * F11;
        c10 = Pf10 * F00 + Pf11 * F01;
        c11 = Pf10 * F10 + Pf11 * F11;
        double J00 = c00 * i00 + c01 * i10, J01 = c00 * i01 + c01 * i11;
        double J10 = c10 * i00 + c11 * i10, J11 = c10 * i01 + c11 * i11;
        double dx0 = (double)xs[(k + 1) * 2] - xp0;
        double dx1 = (double)xs[(k + 1) * 2 + 1] - xp1;
        double s0 = xk0 + (J00 * dx0 + J01 * dx1);
        double s1 = xk1 + (J10 * dx0 + J11 * dx1);
        xs[k * 2] = (float)s0;
        xs[k * 2 + 1] = (float)s1;
        double d00 = (double)Ps[(k + 1) * 4] - PP00, d01 = (double)Ps[(k + 1) * 4 + 1] - PP01;
        double d10 = (double)Ps[(k + 1) * 4 + 2] - PP10, d11 = (double)Ps[(k + 1) * 4 + 3] - PP11;
        double r00 = d00 * J00 + d01 * J01, r01 = d00 * J10 + d01 * J11;
        double r10 = d10 * J00 + d11 * J01, r11 = d10 * J10 + d11 * J11;
        double S00 = Pf00 + (J00 * r00 + J01 * r10);
        double S01 = Pf01 + (J00 * r01 + J01 * r11);
        double S11 = Pf11 + (J10 * r01 + J11 * r11);
        Ps[k * 4] = (float)S00;
        Ps[k * 4 + 1] = (float)S01;
        Ps[k * 4 + 2] = (float)S01;
        Ps[k * 4 + 3] = (float)S11;
        double C00 = c00 + (J00 * d00 + J01 * d10), C01 = c01 + (J00 * d01 + J01 * d11);
        double C10 = c10 + (J10 * d00 + J11 * d10), C11 = c11 + (J10 * d01 + J11 * d11);
        if (k < lag_rows) {
            lagC[k * 4] = (float)C00;
            lagC[k * 4 + 1] = (float)C01;
            lagC[k * 4 + 2] = (float)C10;
            lagC[k * 4 + 3] = (float)C11;
        }
        double lvl = (double)xs[k * 2];
        for (int64_t j = 0; j < m; ++j)
            resid[k * m + j] = (float)((double)data[j * n + k] - lvl);
    }
}

/* 1-state RTS smoother.  cconsenrich.pyx:7116-7148. */
void oracle_backward1(const float *data, int64_t m, int64_t n, const float *xf, const float *Pf,
                      const float *Qf, float *xs, float *Ps, float *lagC, int64_t lag_rows,
                      float *resid) {
    if (n <= 0) return;
    int64_t last = n - 1;
    xs[last] = xf[last];
    Ps[last] = Pf[last];
    for (int64_t j = 0; j < m; ++j)
        resid[last * m + j] = (float)((double)data[j * n + last] - (double)xs[last]);
    for (int64_t k = n - 2; k >= 0; --k) {
        double pf = Pf[k], q = Qf[k];
        double pp = pf + q;
        if (pp < 1.0e-12) pp = 1.0e-12;
        double J = pf / pp;
        double dx = (double)xs[k + 1] - (double)xf[k];
        xs[k] = (float)((double)xf[k] + J * dx);
        double dP = (double)Ps[k + 1] - pp;
        double ps = pf + (J * J * dP);
        if (ps < 0.0) ps = 0.0;
        Ps[k] = (float)ps;
        if (k < lag_rows) lagC[k] = (float)(pf + (J * dP));
        double lvl = (double)xs[k];
        for (int64_t j = 0; j < m; ++j)
            resid[k * m + j] = (float)((double)data[j * n + k] - lvl);
    }
}

/* Student-t observation precision update (lambda).  cconsenrich.pyx:8210-8239 (2-state)
 * and 7474-7497 (level); sdim = 2 or 1 selects the stride of xs / Ps. */
void oracle_update_lambda(const float *data, const float *munc, int64_t m, int64_t n,
                          const int32_t *block_map, int64_t block_count, const float *xs,
                          const float *Ps, int sdim, double pad, double nu, double lam_min,
                          double lam_max, float *lam) {
    for (int64_t k = 0; k < n; ++k) {
        int64_t b = (int64_t)block_map[k];
        if (b < 0 || b >= block_count) {
            lam[k] = 1.0f;
            continue;
        }
        double p00 = (double)Ps[k * sdim * sdim];
        if (p00 < 0.0) p00 = 0.0;
        double lvl = (double)xs[k * sdim];
        double u2 = 0.0;
        for (int64_t j = 0; j < m; ++j) {
            double r = (double)munc[j * n + k] + pad;
            if (r < 1.0e-12) r = 1.0e-12;
            double e = (double)data[j * n + k] - lvl;
            u2 += (e * e + p00) / r;
        }
        double w = (nu + (double)m) / (nu + u2);
        if (w < lam_min)
            w = lam_min;
        else if (w > lam_max)
            w = lam_max;
        lam[k] = (float)w;
    }
}

/* Student-t process precision update (kappa), 2-state.  cconsenrich.pyx:8244-8298 with the
 * MAT2 helpers of pyx:4123-4175. */
void oracle_update_kappa2(int64_t n, const int32_t *block_map, int64_t block_count,
                          const float *xs, const float *Ps, const float *lagC, const double *F,
                          const double *Q0, const float *qscale, double nu, double kap_min,
                          double kap_max, float *kap) {
    if (n <= 0) return;
    double det = Q0[0] * Q0[3] - Q0[1] * Q0[2];
    double qi00 = Q0[3] / det, qi01 = -Q0[1] / det, qi10 = -Q0[2] / det, qi11 = Q0[0] / det;
    const double f00 = F[0], f01 = F[1], f10 = F[2], f11 = F[3];
    kap[0] = 1.0f;
    for (int64_t k = 0; k < n - 1; ++k) {
        int64_t b = (int64_t)block_map[k];
        if (b < 0 || b >= block_count) {
            kap[k + 1] = 1.0f;
            continue;
        }
        double x0 = xs[k * 2], x1 = xs[k * 2 + 1], y0 = xs[(k + 1) * 2], y1 = xs[(k + 1) * 2 + 1];
        /* E[x x^T], E[y y^T], E[x y^T] */
        double xx00 = (double)Ps[k * 4] + x0 * x0, xx01 = (double)Ps[k * 4 + 1] + x0 * x1;
        double xx10 = (double)Ps[k * 4 + 2] + x1 * x0, xx11 = (double)Ps[k * 4 + 3] + x1 * x1;
        double yy00 = (double)Ps[(k + 1) * 4] + y0 * y0, yy01 = (double)Ps[(k + 1) * 4 + 1] + y0 * y1;
        double yy10 = (double)Ps[(k + 1) * 4 + 2] + y1 * y0, yy11 = (double)Ps[(k + 1) * 4 + 3] + y1 * y1;
        double xy00 = (double)lagC[k * 4] + x0 * y0, xy01 = (double)lagC[k * 4 + 1] + x0 * y1;
        double xy10 = (double)lagC[k * 4 + 2] + x1 * y0, xy11 = (double)lagC[k * 4 + 3] + x1 * y1;
        /* yx = xy^T ; yx * F^T */
        double yx00 = xy00, yx01 = xy10, yx10 = xy01, yx11 = xy11;
        double a00 = yx00 * f00 + yx01 * f01, a01 = yx00 * f10 + yx01 * f11;
        double a10 = yx10 * f00 + yx11 * f01, a11 = yx10 * f10 + yx11 * f11;
        /* F * xy */
        double b00 = f00 * xy00 + f01 * xy10, b01 = f00 * xy01 + f01 * xy11;
        double b10 = f10 * xy00 + f11 * xy10, b11 = f10 * xy01 + f11 * xy11;
        /* (F * xx) * F^T */
        double g00 = f00 * xx00 + f01 * xx10, g01 = f00 * xx01 + f01 * xx11;
        double g10 = f10 * xx00 + f11 * xx10, g11 = f10 * xx01 + f11 * xx11;
        double h00 = g00 * f00 + g01 * f01, h01 = g00 * f10 + g01 * f11;
        double h10 = g10 * f00 + g11 * f01, h11 = g10 * f10 + g11 * f11;
        double w00 = ((yy00 - a00) - b00) + h00, w01 = ((yy01 - a01) - b01) + h01;
        double w10 = ((yy10 - a10) - b10) + h10, w11 = ((yy11 - a11) - b11) + h11;
        if (w00 < 0.0) w00 = 0.0;
        if (w11 < 0.0) w11 = 0.0;
        double delta = qi00 * w00 + qi01 * w10 + qi10 * w01 + qi11 * w11;
        if (qscale) delta = delta / (double)qscale[k + 1];
        if (delta < 0.0) delta = 0.0;
        double kv = (nu + 2.0) / (nu + delta);
        if (kv < kap_min)
            kv = kap_min;
        else if (kv > kap_max)
            kv = kap_max;
        kap[k + 1] = (float)kv;
    }
}

/* Student-t process precision update (kappa), level model.  cconsenrich.pyx:7499-7521. */
void oracle_update_kappa1(int64_t n, const int32_t *block_map, int64_t block_count,
                          const float *xs, const float *Ps, const float *lagC, double q0,
                          const float *qscale, double nu, double kap_min, double kap_max,
                          float *kap) {
    if (n <= 0) return;
    double q0inv = 1.0 / q0;
    kap[0] = 1.0f;
    for (int64_t k = 0; k < n - 1; ++k) {
        int64_t b = (int64_t)block_map[k];
        if (b < 0 || b >= block_count) {
            kap[k + 1] = 1.0f;
            continue;
        }
        double x0 = xs[k], y0 = xs[k + 1];
        double Pk = Ps[k], Pk1 = Ps[k + 1], Ck = lagC[k];
        double delta = ((Pk1 + y0 * y0) - (2.0 * (Ck + x0 * y0)) + (Pk + x0 * x0)) * q0inv;
        if (qscale) delta = delta / (double)qscale[k + 1];
        if (delta < 0.0) delta = 0.0;
        double kv = (nu + 1.0) / (nu + delta);
        if (kv < kap_min)
            kv = kap_min;
        else if (kv > kap_max)
            kv = kap_max;
        kap[k + 1] = (float)kv;
    }
}
