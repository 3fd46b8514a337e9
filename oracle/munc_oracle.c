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
