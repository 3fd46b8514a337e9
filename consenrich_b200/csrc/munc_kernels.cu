// Dense [tracks x intervals] kernels of the observation-noise (MUNC) stage that feeds the path
// (SURVEY 8f, next #3).  First of them: the centred rolling mean with an exclusion mask that turns
// per-cell local evidence into the local variance track, cMuncSmoothDenseLocalEvidence
// (cconsenrich.pyx:5547-5740).
//
// The reference slides one running sum along each row (add the entering cell, subtract the leaving
// one).  Here a CTA takes a tile of consecutive outputs of one row, loads the tile + window cells its
// windows cover (two aligned float4 per thread and group), forms tile-local float64 prefix sums of the
// unmasked values next to int32 prefix counts (registers, warp shuffles, one pass through shared
// memory) and reads every window as a difference of two prefixes.  The float64 window sums agree with the
// reference's running sum to ~1e-13 relative, far below the float32 rounding of the output.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "munc_kernels.cuh"
#include "ssm_math.cuh"  // cb_div: the correctly rounded quotient without the special-case path

namespace cb200 {
namespace {

constexpr int RM_THREADS = 256;

// mask semantics: _muncSeedMaskAllowsCell with nonzeroMeansActive = False (pyx:4746-4766), nonzero excludes

// window [left, right) of output i (pyx:5601-5609)
__device__ __forceinline__ void window_of(int64_t i, int64_t n, int64_t W, int64_t &left, int64_t &right) {
    const int64_t half = W / 2;
    left = i >= half ? i - half : 0;
    right = left + W;
    if (right > n) {
        right = n;
        left = right >= W ? right - W : 0;
    }
}

// A CTA covers RM_THREADS * 8 * groups cells of one row, starting at a multiple of 4 so that a thread's
// 8 cells of a round are two aligned float4.  A thread scans its 8 cells in registers, the warp scans
// the thread totals by shuffle, the eight warp totals go through shared memory, and only the finished
// exclusive prefixes (float64 sums, int32 counts) are written to shared memory, once.
__global__ void __launch_bounds__(RM_THREADS)
rolling_mean_kernel(const float *__restrict__ local, const uint8_t *__restrict__ mask, int mask_mode, int64_t n,
                    int64_t ld, int64_t mask_ld, int64_t W, double eps, int tile, int groups, int vec_ok,
                    float *__restrict__ out, int64_t out_ld, int *__restrict__ invalid) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int64_t j = blockIdx.y;
    const int64_t i0 = (int64_t)blockIdx.x * tile;
    if (i0 >= n) return;
    const int64_t i1 = min(i0 + (int64_t)tile, n);  // outputs [i0, i1)
    int64_t lo, hi, t;
    window_of(i0, n, W, lo, t);
    window_of(i1 - 1, n, W, t, hi);  // cells [lo, hi) cover every window of the tile
    lo &= ~(int64_t)3;               // aligned start; the extra cells in front are simply part of the prefix
    const int cells = RM_THREADS * 8 * groups, slots = cells + cells / 8 + 2;
    double *ps = reinterpret_cast<double *>(smem_raw);  // exclusive prefixes of the values (padded slots)
    int *pc = reinterpret_cast<int *>(ps + slots);      // ... of the counts
    __shared__ double warp_sum[RM_THREADS / 32];
    __shared__ int warp_cnt[RM_THREADS / 32];
    const float *row = local + j * ld;
    const uint8_t *mrow = mask_mode == 2 ? mask + j * mask_ld : mask;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    // ---- rounds of RM_THREADS * 8 cells: thread t owns cells [8 t, 8 t + 8) of the round ----
    // slot of boundary b (prefix over the first b cells): one pad slot after every 8 cells keeps the
    // per-thread runs of 8 on different banks
    auto slot = [](int b) { return b == 0 ? 0 : b + ((b - 1) >> 3); };
    double carry_s = 0.0;  // totals of the rounds before this one
    int carry_c = 0, bad = 0;
    for (int g = 0; g < groups; ++g) {
        const int c0 = (g * RM_THREADS + tid) * 8;
        const int64_t k = lo + c0;
        float f[8];
        uint8_t ex[8];
        if (vec_ok && k + 8 <= n) {
            const float4 a = *reinterpret_cast<const float4 *>(row + k), b = *reinterpret_cast<const float4 *>(row + k + 4);
            f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
            if (!mask_mode) {
#pragma unroll
                for (int u = 0; u < 8; ++u) ex[u] = 0;
            } else if ((reinterpret_cast<uintptr_t>(mrow + k) & 3) == 0) {
                const uint32_t m0 = *reinterpret_cast<const uint32_t *>(mrow + k);
                const uint32_t m1 = *reinterpret_cast<const uint32_t *>(mrow + k + 4);
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    ex[u] = (uint8_t)(m0 >> (8 * u));
                    ex[4 + u] = (uint8_t)(m1 >> (8 * u));
                }
            } else {
#pragma unroll
                for (int u = 0; u < 8; ++u) ex[u] = mrow[k + u];
            }
        } else {
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const bool in = k + u < n;
                f[u] = in ? row[k + u] : 1.0f;
                ex[u] = in ? (mask_mode ? mrow[k + u] : 0) : 1;  // cells past the row end count as excluded
            }
        }
        // inclusive scan of the own 8 cells in registers
        double ls[8];
        int lc[8];
        double run = 0.0;
        int cnt = 0;
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const bool on = ex[u] == 0;
            if (on && (!(f[u] > 0.0f) || f[u] == INFINITY)) bad = 1;  // not positive and finite (pyx:5569-5571)
            run += on ? (double)f[u] : 0.0;
            cnt += on ? 1 : 0;
            ls[u] = run;
            lc[u] = cnt;
        }
        // exclusive prefix of the thread totals: across the warp by shuffle, across warps through shared memory
        double inc_s = run;
        int inc_c = cnt;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const double os = __shfl_up_sync(0xffffffffu, inc_s, d);
            const int oc = __shfl_up_sync(0xffffffffu, inc_c, d);
            if (lane >= d) {
                inc_s += os;
                inc_c += oc;
            }
        }
        if (g) __syncthreads();  // the warp totals of the previous round have been read
        if (lane == 31) {
            warp_sum[warp] = inc_s;
            warp_cnt[warp] = inc_c;
        }
        __syncthreads();
        double off_s = carry_s + (inc_s - run);
        int off_c = carry_c + (inc_c - cnt);
#pragma unroll
        for (int w = 0; w < RM_THREADS / 32; ++w) {
            if (w < warp) {
                off_s += warp_sum[w];
                off_c += warp_cnt[w];
            }
            carry_s += warp_sum[w];
            carry_c += warp_cnt[w];
        }
        double *pd = ps + slot(c0 + 1);
        int *pi = pc + slot(c0 + 1);
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            pd[u] = off_s + ls[u];
            pi[u] = off_c + lc[u];
        }
    }
    if (bad) atomicOr(invalid, 1);
    if (tid == 0) {
        ps[0] = 0.0;
        pc[0] = 0;
    }
    __syncthreads();

    // ---- outputs: every window is a difference of two prefixes ----
    float *orow = out + j * out_ld;
    const int64_t half = W / 2;
    const bool interior = i0 >= half && (i1 - 1 - half) + W <= n;  // all windows are [i - half, i - half + W)
    const int base = interior ? (int)((i0 - half) - lo) : 0;
    const int nout = (int)(i1 - i0), iw = (int)W;
    if (interior && vec_ok && (out_ld & 3) == 0 && (nout & 3) == 0 && ((i0 & 3) == 0) &&
        (reinterpret_cast<uintptr_t>(out) & 15) == 0) {
        // four consecutive outputs per thread: their boundaries are consecutive too
        for (int o = 4 * tid; o < nout; o += 4 * RM_THREADS) {
            float r[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int ql = slot(base + o + u), qr = slot(base + o + u + iw);
                const int count = pc[qr] - pc[ql];
                double v = count > 0 ? cb_div(ps[qr] - ps[ql], (double)count) : (double)row[i0 + o + u];
                if (v < eps) v = eps;
                r[u] = (float)v;
            }
            *reinterpret_cast<float4 *>(orow + i0 + o) = make_float4(r[0], r[1], r[2], r[3]);
        }
        return;
    }
    for (int o = tid; o < nout; o += RM_THREADS) {
        int bl, br;
        if (interior) {
            bl = base + o;
            br = bl + iw;
        } else {
            int64_t l, r;
            window_of(i0 + o, n, W, l, r);
            bl = (int)(l - lo);
            br = (int)(r - lo);
        }
        const int ql = slot(bl), qr = slot(br);
        const int count = pc[qr] - pc[ql];
        double v = count > 0 ? cb_div(ps[qr] - ps[ql], (double)count) : (double)row[i0 + o];
        if (v < eps) v = eps;
        orow[i0 + o] = (float)v;
    }
}

// ---- cFinalizeMuncEBTrack (cconsenrich.pyx:5372-5440): per-interval shrinkage of the local variance
// towards the prior, clipping, count floor.  Elementwise; separately rounded float64 products, sums and
// quotient, so the float32 output is bit-identical to the reference's loop.
constexpr int FIN_THREADS = 256;

__device__ __forceinline__ double clip_var(double v, double lo, double hi) { return v < lo ? lo : (v > hi ? hi : v); }

__global__ void __launch_bounds__(FIN_THREADS)
finalize_eb_kernel(const float *__restrict__ local, const float *__restrict__ prior, const float *__restrict__ cfloor,
                   int64_t n, double nu_local, double nu_prior, double post, double vfloor, double vcap, int use_eb,
                   float *__restrict__ out, MuncFinalizeStatus *st) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int support = 0, cf_finite = 0, cf_added = 0, cf_missing = 0;
    if (i < n) {
        double lv = (double)local[i];
        bool ok = true;
        if (!isfinite(lv) || lv <= 0.0) {
            atomicMin(reinterpret_cast<long long *>(&st->invalid_local), (long long)i);
            ok = false;
        }
        if (ok) {
            support = lv > vfloor;
            lv = clip_var(lv, vfloor, vcap);
            double ov = lv;
            if (use_eb) {
                double pv = (double)prior[i];
                if (!isfinite(pv) || pv <= 0.0) {
                    atomicMin(reinterpret_cast<long long *>(&st->invalid_prior), (long long)i);
                    ok = false;
                } else {
                    pv = clip_var(pv, vfloor, vcap);
                    ov = __ddiv_rn(__dadd_rn(__dmul_rn(nu_local, lv), __dmul_rn(nu_prior, pv)), post);
                }
            }
            if (ok) {
                ov = clip_var(ov, vfloor, vcap);
                if (cfloor) {
                    const double cv = (double)cfloor[i];
                    if (cv == cv) {  // NaN = no count floor for this interval
                        if (!isfinite(cv) || cv < 0.0) {
                            atomicMin(reinterpret_cast<long long *>(&st->invalid_cfloor), (long long)i);
                            ok = false;
                        } else {
                            cf_finite = 1;
                            ov = __dadd_rn(ov, cv);
                            cf_added = cv > 0.0;
                            ov = clip_var(ov, vfloor, vcap);
                        }
                    } else {
                        cf_missing = 1;
                    }
                }
                if (ok) out[i] = (float)ov;
            }
            if (!ok) support = cf_finite = cf_added = cf_missing = 0;
        }
    }
    // counters: warp ballots, one atomic per warp and counter that has anything to add
    const unsigned b0 = __ballot_sync(0xffffffffu, support), b1 = __ballot_sync(0xffffffffu, cf_finite);
    const unsigned b2 = __ballot_sync(0xffffffffu, cf_added), b3 = __ballot_sync(0xffffffffu, cf_missing);
    if ((threadIdx.x & 31) == 0) {
        if (b0) atomicAdd(reinterpret_cast<unsigned long long *>(&st->support), (unsigned long long)__popc(b0));
        if (b1) atomicAdd(reinterpret_cast<unsigned long long *>(&st->cfloor_finite), (unsigned long long)__popc(b1));
        if (b2) atomicAdd(reinterpret_cast<unsigned long long *>(&st->cfloor_added), (unsigned long long)__popc(b2));
        if (b3) atomicAdd(reinterpret_cast<unsigned long long *>(&st->cfloor_missing), (unsigned long long)__popc(b3));
    }
}

// ---- cMuncObservationMomentSeedPass (cconsenrich.pyx:4843-5040) ----------------------------------
// One thread per interval, the tracks in order (the Student-t weight of an interval averages over its
// active tracks in that order); every product, sum and quotient rounded separately, as the reference's
// C does, so all six outputs are bit-identical.
constexpr int SEED_THREADS = 256;

__device__ __forceinline__ bool seed_active(const MuncSeedArgs &a, int64_t j, int64_t k) {
    if (a.active_mode == 0) return true;
    return (a.active_mode == 1 ? a.active[k] : a.active[j * a.active_ld + k]) != 0;
}

// clip the local variance and the total (local + count floor) the way pyx:5003-5015 does
__device__ __forceinline__ void seed_clip(double &lv, double &tv, double cv, double vfloor, double vcap) {
    if (lv < vfloor) {
        lv = vfloor;
        tv = __dadd_rn(lv, cv);
    }
    if (tv > vcap) {
        tv = vcap;
        lv = __dsub_rn(tv, cv);
        if (lv < vfloor) {
            lv = vfloor;
            tv = __dadd_rn(lv, cv);
        }
    }
}

__global__ void __launch_bounds__(SEED_THREADS) seed_pass_kernel(const MuncSeedArgs a, int *__restrict__ invalid) {
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= a.n) return;
    const bool weighted = a.use_weights && a.student_t;
    const double state = (double)a.state_mean[k];
    double mvb = (double)a.state_var[k];
    const double bg = a.background ? (double)a.background[k] : 0.0;
    const double gv = a.g_var ? (double)a.g_var[k] : 0.0;
    bool bad_k = !isfinite(state) || !isfinite(mvb) || !isfinite(bg) || !isfinite(gv);
    if (a.g_var) mvb = __dadd_rn(mvb, gv);
    if (mvb < 0.0) mvb = 0.0;
    double omega_in = 1.0;
    if (weighted && a.omega_in) {
        omega_in = (double)a.omega_in[k];
        if (!isfinite(omega_in) || omega_in <= 0.0) bad_k = true;
    }
    bool bad = false;
    double omega_raw = 1.0, omega = 1.0;
    if (weighted) {
        if (a.update_weights) {
            double dbar = 0.0;
            int64_t active = 0;
            for (int64_t j = 0; j < a.m; ++j) {
                const int64_t idx = j * a.ld + k;
                if (!seed_active(a, j, k)) {
                    a.moment[idx] = 0.0f;
                    a.rho_out[idx] = 1.0f;
                    continue;
                }
                const double d = (double)a.data[idx];
                const double mu = __dadd_rn((double)a.munc[idx], a.pad);
                if (bad_k || !isfinite(d) || !isfinite(mu) || mu <= 0.0) bad = true;
                double base = mu;
                if (base < a.var_floor) base = a.var_floor;
                const double res = __dsub_rn(__dsub_rn(d, bg), state);
                const double mom = __dadd_rn(__dmul_rn(res, res), mvb);
                const double rho = __ddiv_rn(__dadd_rn(a.d_s, 1.0),
                                             __dadd_rn(a.d_s, __ddiv_rn(__dmul_rn(omega_in, mom), base)));
                a.moment[idx] = (float)mom;
                a.rho_out[idx] = (float)rho;
                dbar = __dadd_rn(dbar, __ddiv_rn(mom, base));
                active += 1;
            }
            if (active > 0) {
                dbar = __ddiv_rn(dbar, (double)active);
                omega_raw = __ddiv_rn(__dadd_rn(a.d_omega, 1.0), __dadd_rn(a.d_omega, dbar));
                omega = omega_raw < a.omega_min ? a.omega_min : (omega_raw > a.omega_max ? a.omega_max : omega_raw);
            }
        } else {
            omega_raw = omega_in;
            omega = omega_raw < a.omega_min ? a.omega_min : (omega_raw > a.omega_max ? a.omega_max : omega_raw);
            for (int64_t j = 0; j < a.m; ++j) {
                const int64_t idx = j * a.ld + k;
                if (!seed_active(a, j, k)) {
                    a.moment[idx] = 0.0f;
                    a.rho_out[idx] = 1.0f;
                    continue;
                }
                const double d = (double)a.data[idx];
                const double mu = __dadd_rn((double)a.munc[idx], a.pad);
                const double rho = a.rho_in ? (double)a.rho_in[idx] : 1.0;
                if (bad_k || !isfinite(d) || !isfinite(mu) || mu <= 0.0 || !isfinite(rho) || rho <= 0.0) bad = true;
                const double res = __dsub_rn(__dsub_rn(d, bg), state);
                a.moment[idx] = (float)__dadd_rn(__dmul_rn(res, res), mvb);
                a.rho_out[idx] = (float)rho;
            }
        }
        a.omega_raw[k] = (float)omega_raw;
        a.omega_out[k] = (float)omega;
    } else {
        a.omega_raw[k] = 1.0f;
        a.omega_out[k] = 1.0f;
    }
    for (int64_t j = 0; j < a.m; ++j) {
        const int64_t idx = j * a.ld + k;
        const double cv = a.count_floor ? (double)a.count_floor[idx] : 0.0;
        double lv, tv;
        if (seed_active(a, j, k)) {
            if (a.count_floor && (!isfinite(cv) || cv < 0.0)) bad = true;
            double mom, rho = 1.0;
            if (!weighted) {
                const double d = (double)a.data[idx];
                const double mu = __dadd_rn((double)a.munc[idx], a.pad);
                if (bad_k || !isfinite(d) || !isfinite(mu) || mu <= 0.0) bad = true;
                const double res = __dsub_rn(__dsub_rn(d, bg), state);
                mom = __dadd_rn(__dmul_rn(res, res), mvb);
                a.moment[idx] = (float)mom;
                a.rho_out[idx] = 1.0f;
                lv = __dsub_rn(__dsub_rn(mom, a.pad), cv);
            } else {
                mom = (double)a.moment[idx];  // the float32 values just stored, as the reference re-reads them
                rho = (double)a.rho_out[idx];
                lv = __dsub_rn(__dsub_rn(__dmul_rn(__dmul_rn(omega, rho), mom), a.pad), cv);
            }
            tv = __dadd_rn(lv, cv);
            seed_clip(lv, tv, cv, a.var_floor, a.var_cap);
        } else {
            lv = __dsub_rn((double)a.munc[idx], cv);
            if (lv < a.var_floor) lv = a.var_floor;
            tv = __dadd_rn(lv, cv);
            if (tv > a.var_cap) {
                tv = a.var_cap;
                lv = __dsub_rn(tv, cv);
                if (lv < a.var_floor) {
                    lv = a.var_floor;
                    tv = __dadd_rn(lv, cv);
                }
            }
            a.moment[idx] = 0.0f;
            a.rho_out[idx] = 1.0f;
        }
        a.local[idx] = (float)lv;
        a.variance[idx] = (float)tv;
    }
    if (bad) atomicOr(invalid, 1);
}

// ---- cEMA (cconsenrich.pyx:5744-5759, 5897-5915): forward then backward exponential filter of a track ----
// y_p = alpha x_p + (1 - alpha) y_{p-1} is an affine recurrence: each thread composes the map of its run
// of EMA_RUN elements in float64, one CTA scans the run maps, and each thread then REPLAYS the
// reference's own arithmetic (its rounding sequence in the track's type) over the run before its own --
// a warm-up that lets the replay forget the scan's un-rounded start value -- and over its own run.
constexpr int EMA_RUN = 256;
constexpr int EMA_THREADS = 128;
constexpr int EMA_SCAN_THREADS = 1024;

template <class T>
__device__ __forceinline__ T ema_step(T alpha, double c, T x, T y_prev);
template <>
__device__ __forceinline__ float ema_step<float>(float alpha, double c, float x, float y_prev) {
    // alpha * x in float; (1.0 - alpha) * y in double; sum in double; stored as float (the C semantics of
    // pyx:5752 for real_t = float)
    return (float)__dadd_rn((double)__fmul_rn(alpha, x), __dmul_rn(c, (double)y_prev));
}
template <>
__device__ __forceinline__ double ema_step<double>(double alpha, double c, double x, double y_prev) {
    return __dadd_rn(__dmul_rn(alpha, x), __dmul_rn(c, y_prev));
}

// element p of the pass in processing order: index p forward, n - 1 - p backward
__device__ __forceinline__ int64_t ema_index(int64_t p, int64_t n, int reversed) { return reversed ? n - 1 - p : p; }

template <class T>
__global__ void __launch_bounds__(EMA_THREADS) ema_run_maps_kernel(const T *__restrict__ src, int64_t n, int reversed,
                                                                   T alpha, double c, int64_t runs,
                                                                   double2 *__restrict__ maps) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= runs) return;
    double A = 1.0, B = 0.0;
    const int64_t p1 = min((r + 1) * EMA_RUN, n);
    for (int64_t p = r * EMA_RUN; p < p1; ++p) {
        const T x = src[ema_index(p, n, reversed)];
        if (p == 0) {  // the pass starts from the first element itself
            A = 0.0;
            B = (double)x;
        } else {
            B = c * B + (double)ema_step<T>(alpha, 0.0, x, (T)0);  // alpha x in the track's arithmetic
            A = c * A;
        }
    }
    maps[r] = make_double2(A, B);
}

// exclusive scan of the run maps: state[r] = value of the pass just before run r (state[0] unused)
__global__ void __launch_bounds__(EMA_SCAN_THREADS) ema_scan_kernel(const double2 *__restrict__ maps, int64_t runs,
                                                                    double *__restrict__ state) {
    __shared__ double sA[EMA_SCAN_THREADS], sB[EMA_SCAN_THREADS];
    const int t = threadIdx.x;
    const int64_t per = (runs + EMA_SCAN_THREADS - 1) / EMA_SCAN_THREADS;
    const int64_t r0 = min((int64_t)t * per, runs), r1 = min(r0 + per, runs);
    double A = 1.0, B = 0.0;  // composition of this thread's runs
    for (int64_t r = r0; r < r1; ++r) {
        const double2 m = maps[r];
        B = m.x * B + m.y;
        A = m.x * A;
    }
    sA[t] = A;
    sB[t] = B;
    __syncthreads();
    for (int d = 1; d < EMA_SCAN_THREADS; d <<= 1) {  // inclusive scan of the thread maps (later after earlier)
        double a = sA[t], b = sB[t];
        if (t >= d) {
            b = a * sB[t - d] + b;
            a = a * sA[t - d];
        }
        __syncthreads();
        sA[t] = a;
        sB[t] = b;
        __syncthreads();
    }
    double y = t > 0 ? sB[t - 1] : 0.0;  // value before this thread's first run (the chain starts with A = 0)
    for (int64_t r = r0; r < r1; ++r) {
        state[r] = y;
        const double2 m = maps[r];
        y = m.x * y + m.y;
    }
}

template <class T>
__global__ void __launch_bounds__(EMA_THREADS) ema_replay_kernel(const T *__restrict__ src, T *__restrict__ dst, int64_t n,
                                                                 int reversed, T alpha, double c, int64_t runs,
                                                                 const double *__restrict__ state) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= runs) return;
    const int64_t own0 = r * EMA_RUN, p1 = min((r + 1) * EMA_RUN, n);
    const int64_t p0 = r > 0 ? own0 - EMA_RUN : 0;  // warm-up: the run before
    T y = r > 1 ? (T)state[r - 1] : (T)0;           // value before the warm-up run (run 0 starts the pass itself)
    for (int64_t p = p0; p < p1; ++p) {
        const int64_t i = ema_index(p, n, reversed);
        const T x = src[i];
        y = p == 0 ? x : ema_step<T>(alpha, c, x, y);
        if (p >= own0) dst[i] = y;
    }
}

template <class T>
cudaError_t ema_launch(const T *x, T *tmp, T *out, int64_t n, double alpha, double2 *maps, double *state,
                       cudaStream_t st) {
    const int64_t runs = (n + EMA_RUN - 1) / EMA_RUN;
    const unsigned grid = (unsigned)((runs + EMA_THREADS - 1) / EMA_THREADS);
    const T a = (T)alpha;
    const double c = 1.0 - (double)a;  // (1.0 - alpha) with alpha in the track's type (pyx:5752)
    for (int pass = 0; pass < 2; ++pass) {
        const T *src = pass == 0 ? x : tmp;
        T *dst = pass == 0 ? tmp : out;
        ema_run_maps_kernel<T><<<grid, EMA_THREADS, 0, st>>>(src, n, pass, a, c, runs, maps);
        ema_scan_kernel<<<1, EMA_SCAN_THREADS, 0, st>>>(maps, runs, state);
        ema_replay_kernel<T><<<grid, EMA_THREADS, 0, st>>>(src, dst, n, pass, a, c, runs, state);
    }
    return cudaGetLastError();
}

}  // namespace

size_t munc_ema_workspace_bytes(int64_t n) {
    const size_t runs = (size_t)((n + EMA_RUN - 1) / EMA_RUN) + 1;
    return runs * 16 + runs * 8 + 64;
}

cudaError_t launch_munc_ema(const void *x, void *tmp, void *out, int64_t n, int is_double, double alpha, void *workspace,
                            cudaStream_t st) {
    if (n <= 0) return cudaSuccess;
    const size_t runs = (size_t)((n + EMA_RUN - 1) / EMA_RUN) + 1;
    double2 *maps = static_cast<double2 *>(workspace);
    double *state = reinterpret_cast<double *>(maps + runs);
    if (is_double)
        return ema_launch<double>(static_cast<const double *>(x), static_cast<double *>(tmp), static_cast<double *>(out), n,
                                  alpha, maps, state, st);
    return ema_launch<float>(static_cast<const float *>(x), static_cast<float *>(tmp), static_cast<float *>(out), n, alpha,
                             maps, state, st);
}

cudaError_t launch_munc_seed_pass(const MuncSeedArgs &a, int *invalid, cudaStream_t st) {
    cudaError_t e = cudaMemsetAsync(invalid, 0, sizeof(int), st);
    if (e != cudaSuccess || a.n <= 0 || a.m <= 0) return e;
    seed_pass_kernel<<<(unsigned)((a.n + SEED_THREADS - 1) / SEED_THREADS), SEED_THREADS, 0, st>>>(a, invalid);
    return cudaGetLastError();
}

cudaError_t launch_munc_finalize_eb(const float *local, const float *prior, const float *cfloor, int64_t n,
                                    double nu_local, double nu_prior, double vfloor, double vcap, int use_eb, float *out,
                                    MuncFinalizeStatus *status, cudaStream_t st) {
    // counters 0, invalid_* = STATUS_NONE: two memsets, no host-to-device copy
    cudaError_t e = cudaMemsetAsync(status, 0, 4 * sizeof(int64_t), st);
    if (e != cudaSuccess) return e;
    e = cudaMemsetAsync(&status->invalid_local, 0x7f, 3 * sizeof(int64_t), st);
    if (e != cudaSuccess || n <= 0) return e;
    finalize_eb_kernel<<<(unsigned)((n + FIN_THREADS - 1) / FIN_THREADS), FIN_THREADS, 0, st>>>(
        local, prior, cfloor, n, nu_local, nu_prior, nu_local + nu_prior, vfloor, vcap, use_eb, out, status);
    return cudaGetLastError();
}

// rounds of 8 cells per thread: a CTA covers RM_THREADS * 8 * groups cells, at least 1.5 windows
// (+ alignment slack), so that a cell is loaded at most about three times even for the widest window
// shared memory admits
int munc_rolling_groups(int64_t window) {
    return (int)((window + window / 2 + 8 + RM_THREADS * 8 - 1) / (RM_THREADS * 8));
}

size_t munc_rolling_smem(int groups) {
    const size_t cells = (size_t)RM_THREADS * 8 * groups, slots = cells + cells / 8 + 2;
    return slots * 8 + slots * 4;
}

cudaError_t launch_munc_rolling_mean(const float *local, const uint8_t *mask, int mask_mode, int64_t m, int64_t n,
                                     int64_t ld, int64_t mask_ld, int64_t window, double eps, float *out, int64_t out_ld,
                                     int *invalid, cudaStream_t st) {
    if (m <= 0 || n <= 0) return cudaSuccess;
    const int groups = munc_rolling_groups(window);
    // outputs per CTA (3: alignment slack of the first cell), a multiple of 4 for the vector stores
    const int tile = (int)(((int64_t)RM_THREADS * 8 * groups - window - 3) & ~(int64_t)3);
    const size_t smem = munc_rolling_smem(groups);
    const int vec_ok = (ld % 4 == 0) && ((reinterpret_cast<uintptr_t>(local) & 15) == 0);
    cudaError_t e = cudaFuncSetAttribute(rolling_mean_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    e = cudaMemsetAsync(invalid, 0, sizeof(int), st);
    if (e != cudaSuccess) return e;
    const dim3 grid((unsigned)((n + tile - 1) / tile), (unsigned)m);
    rolling_mean_kernel<<<grid, RM_THREADS, smem, st>>>(local, mask, mask_mode, n, ld, mask_ld, window, eps, tile, groups,
                                                        vec_ok, out, out_ld, invalid);
    return cudaGetLastError();
}

}  // namespace cb200
