// Background-track kernels: the per-interval weighted statistics of the
// residual matrix and the roughness-penalised solve
//     (diag(w) + lamFirst D1'D1 + lam D2'D2) x = rhs        [, sum(x) = 0 by a Lagrange multiplier]
// that the reference runs once per outer pass and per IRLS pass of its background update
// (cconsenrich.pyx:9675-9724, 944-1096; core.py:8085-8378).
//
// The reference factorises the pentadiagonal matrix sequentially (LDL').  Here the n unknowns are
// paired into 2x2 blocks, which turns the system into a symmetric block-TRIdiagonal one, and that is
// solved by block cyclic reduction: each level eliminates every other block row in parallel (Schur
// complements of a symmetric positive definite matrix stay symmetric positive definite, so no
// pivoting is needed), log2(n/2) levels down, the same number back up.  Two right-hand sides ride
// along: rhs and the vector of ones the zero-sum constraint needs.
//
// A block row is 12 doubles: D (symmetric: d00 d01 d11), U (coupling to the NEXT row; the coupling to
// the previous row is the previous row's U transposed), B (2 unknowns x 2 right-hand sides), pad.
#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#include "background_kernels.cuh"

namespace cb200 {
namespace {

constexpr int ROW = 12;  // doubles per block row
constexpr double MIN_PIVOT = 1.0e-12;  // cconsenrich.pyx:968

struct Row {
    double d00, d01, d11;
    double u00, u01, u10, u11;
    double b00, b01, b10, b11;  // b[r][c]: unknown r of the block, right-hand side c
};

__device__ __forceinline__ Row load_row(const double *p) {
    const double2 *q = reinterpret_cast<const double2 *>(p);
    const double2 a = q[0], b = q[1], c = q[2], d = q[3], e = q[4], f = q[5];
    return Row{a.x, a.y, b.x, b.y, c.x, c.y, d.x, d.y, e.x, e.y, f.x};
}
__device__ __forceinline__ void store_row(double *p, const Row &r) {
    double2 *q = reinterpret_cast<double2 *>(p);
    q[0] = make_double2(r.d00, r.d01);
    q[1] = make_double2(r.d11, r.u00);
    q[2] = make_double2(r.u01, r.u10);
    q[3] = make_double2(r.u11, r.b00);
    q[4] = make_double2(r.b01, r.b10);
    q[5] = make_double2(r.b11, 0.0);
}

// inverse of the symmetric 2x2 block; reports a pivot below the floor (row index) through *bad
struct Inv2 {
    double i00, i01, i11;
};
__device__ __forceinline__ Inv2 inv_spd2(const Row &r, int64_t unknown0, BackgroundStatus *st) {
    // LDL' pivots of the block: d00 and d11 - d01^2 / d00
    const double p0 = r.d00;
    const double p1 = r.d11 - (r.d01 * r.d01) / r.d00;
    if (!(p0 >= MIN_PIVOT) || !(p1 >= MIN_PIVOT)) {
        const long long idx = unknown0 + ((p0 >= MIN_PIVOT) ? 1 : 0);
        const long long old = atomicMin(reinterpret_cast<long long *>(&st->bad_index), idx);
        if (idx < old) st->bad_value = (p0 >= MIN_PIVOT) ? p1 : p0;  // advisory (racy only between bad pivots)
    }
    const double rdet = 1.0 / (r.d00 * r.d11 - r.d01 * r.d01);
    return Inv2{r.d11 * rdet, -r.d01 * rdet, r.d00 * rdet};
}

// ---- penalty coefficients (cconsenrich.pyx:905-943) ----
__device__ __forceinline__ double second_diag(int64_t n, int64_t i, double lam) {
    if (n < 3 || lam <= 0.0) return 0.0;
    if (n == 3) return i == 1 ? 4.0 * lam : lam;
    if (i == 0 || i == n - 1) return lam;
    if (i == 1 || i == n - 2) return 5.0 * lam;
    return 6.0 * lam;
}
__device__ __forceinline__ double second_off1(int64_t n, int64_t i, double lam) {
    if (n < 3 || lam <= 0.0) return 0.0;
    if (n == 3) return -2.0 * lam;
    if (i == 0 || i == n - 2) return -2.0 * lam;
    return -4.0 * lam;
}
__device__ __forceinline__ double first_diag(int64_t n, int64_t i, double lam) {
    if (n < 2 || lam <= 0.0) return 0.0;
    return (i == 0 || i == n - 1) ? lam : 2.0 * lam;
}
__device__ __forceinline__ double first_off1(int64_t n, double lam) { return (n < 2 || lam <= 0.0) ? 0.0 : -lam; }

// entries of the pentadiagonal matrix; unknowns >= n (the pad of an odd n) are decoupled identities
__device__ __forceinline__ double mat_diag(const double *w, int64_t n, int64_t k, double lam, double lam1,
                                           BackgroundStatus *st) {
    if (k >= n) return 1.0;
    double v = w[k] + first_diag(n, k, lam1) + second_diag(n, k, lam);
    if (v < MIN_PIVOT) {  // the reference floors the entry and fails the solve (pyx:1024-1031)
        const long long old = atomicMin(reinterpret_cast<long long *>(&st->bad_index), (long long)k);
        if ((long long)k < old) st->bad_value = v;
        v = MIN_PIVOT;
    }
    return v;
}
__device__ __forceinline__ double mat_off1(int64_t n, int64_t k, double lam, double lam1) {  // (k, k+1)
    return (k + 1 < n) ? first_off1(n, lam1) + second_off1(n, k, lam) : 0.0;
}
__device__ __forceinline__ double mat_off2(int64_t n, int64_t k, double lam) {  // (k, k+2)
    return (k + 2 < n) ? lam : 0.0;
}

// block row i of level 0, straight from the weight / rhs tracks
struct Level0 {
    const double *w, *rhs;
    int64_t n;
    double lam, lam1;
    const double *rhs2;  // second right-hand side; nullptr: the vector of ones
};
__device__ __forceinline__ Row build_row(const Level0 &a, int64_t i, BackgroundStatus *st) {
    const double *w = a.w, *rhs = a.rhs, *rhs2 = a.rhs2;
    const int64_t n = a.n, k = 2 * i;
    const double lam = a.lam, lam1 = a.lam1;
    Row r;
    r.d00 = mat_diag(w, n, k, lam, lam1, st);
    r.d11 = mat_diag(w, n, k + 1, lam, lam1, st);
    r.d01 = mat_off1(n, k, lam, lam1);
    // U: [x_k, x_{k+1}] against [x_{k+2}, x_{k+3}]
    r.u00 = mat_off2(n, k, lam);
    r.u01 = 0.0;
    r.u10 = mat_off1(n, k + 1, lam, lam1);
    r.u11 = mat_off2(n, k + 1, lam);
    r.b00 = rhs[k];
    r.b01 = rhs2 ? rhs2[k] : 1.0;
    r.b10 = (k + 1 < n) ? rhs[k + 1] : 0.0;
    r.b11 = (k + 1 < n) ? (rhs2 ? rhs2[k + 1] : 1.0) : 0.0;
    return r;
}

// rows of a level: stored (12 doubles each) or, for level 0, formed on the fly
struct Rows {
    const double *stored;  // nullptr: level 0
    Level0 l0;
};
__device__ __forceinline__ Row get_row(const Rows &rw, int64_t i, BackgroundStatus *st) {
    return rw.stored ? load_row(rw.stored + i * ROW) : build_row(rw.l0, i, st);
}

// new row j of the next level from rows 2j-1, 2j, 2j+1 of this one
__device__ __forceinline__ void reduce_row(const Rows &in, int64_t rows_in, int64_t j, int64_t stride_unknowns,
                                           double *out, BackgroundStatus *st) {
    const int64_t i = 2 * j;
    Row c = get_row(in, i, st);
    Row o = c;
    o.u00 = o.u01 = o.u10 = o.u11 = 0.0;
    if (i - 1 >= 0) {
        const Row p = get_row(in, i - 1, st);
        const Inv2 v = inv_spd2(p, (i - 1) * stride_unknowns, st);
        // L_i = U_p'.  alpha = L_i inv(D_p) = U_p' V
        const double a00 = p.u00 * v.i00 + p.u10 * v.i01, a01 = p.u00 * v.i01 + p.u10 * v.i11;
        const double a10 = p.u01 * v.i00 + p.u11 * v.i01, a11 = p.u01 * v.i01 + p.u11 * v.i11;
        // D -= alpha U_p
        o.d00 -= a00 * p.u00 + a01 * p.u10;
        o.d01 -= a00 * p.u01 + a01 * p.u11;
        o.d11 -= a10 * p.u01 + a11 * p.u11;
        o.b00 -= a00 * p.b00 + a01 * p.b10;
        o.b01 -= a00 * p.b01 + a01 * p.b11;
        o.b10 -= a10 * p.b00 + a11 * p.b10;
        o.b11 -= a10 * p.b01 + a11 * p.b11;
        // (the new coupling to the previous kept row is that row's new U transposed)
    }
    if (i + 1 < rows_in) {
        const Row q = get_row(in, i + 1, st);
        const Inv2 v = inv_spd2(q, (i + 1) * stride_unknowns, st);
        // gamma = U_i inv(D_q)
        const double g00 = c.u00 * v.i00 + c.u01 * v.i01, g01 = c.u00 * v.i01 + c.u01 * v.i11;
        const double g10 = c.u10 * v.i00 + c.u11 * v.i01, g11 = c.u10 * v.i01 + c.u11 * v.i11;
        // D -= gamma L_q = gamma U_i'
        o.d00 -= g00 * c.u00 + g01 * c.u01;
        o.d01 -= g00 * c.u10 + g01 * c.u11;
        o.d11 -= g10 * c.u10 + g11 * c.u11;
        o.b00 -= g00 * q.b00 + g01 * q.b10;
        o.b01 -= g00 * q.b01 + g01 * q.b11;
        o.b10 -= g10 * q.b00 + g11 * q.b10;
        o.b11 -= g10 * q.b01 + g11 * q.b11;
        // U' = -gamma U_q
        o.u00 = -(g00 * q.u00 + g01 * q.u10);
        o.u01 = -(g00 * q.u01 + g01 * q.u11);
        o.u10 = -(g10 * q.u00 + g11 * q.u10);
        o.u11 = -(g10 * q.u01 + g11 * q.u11);
    }
    store_row(out + j * ROW, o);
}

// X: [rows][4] doubles = x[r][c] of each block row.  Odd rows of this level from their even neighbours
// (already solved: they are the rows of the next level), even rows copied from the next level.
__device__ __forceinline__ void backsub_row(const Rows &lvl, int64_t rows, const double *x_next, int64_t i,
                                            double *x_out, BackgroundStatus *st, int64_t stride_unknowns) {
    double2 *o = reinterpret_cast<double2 *>(x_out + i * 4);
    if ((i & 1) == 0) {
        const double2 *s = reinterpret_cast<const double2 *>(x_next + (i >> 1) * 4);
        o[0] = s[0];
        o[1] = s[1];
        return;
    }
    const Row r = get_row(lvl, i, st);
    const Row p = get_row(lvl, i - 1, st);  // for L_i = U_p'
    const double2 *xp = reinterpret_cast<const double2 *>(x_next + ((i - 1) >> 1) * 4);
    const double2 xp0 = xp[0], xp1 = xp[1];  // x[0][0], x[0][1]; x[1][0], x[1][1]
    double t00 = r.b00 - (p.u00 * xp0.x + p.u10 * xp1.x);
    double t01 = r.b01 - (p.u00 * xp0.y + p.u10 * xp1.y);
    double t10 = r.b10 - (p.u01 * xp0.x + p.u11 * xp1.x);
    double t11 = r.b11 - (p.u01 * xp0.y + p.u11 * xp1.y);
    if (i + 1 < rows) {
        const double2 *xq = reinterpret_cast<const double2 *>(x_next + ((i + 1) >> 1) * 4);
        const double2 xq0 = xq[0], xq1 = xq[1];
        t00 -= r.u00 * xq0.x + r.u01 * xq1.x;
        t01 -= r.u00 * xq0.y + r.u01 * xq1.y;
        t10 -= r.u10 * xq0.x + r.u11 * xq1.x;
        t11 -= r.u10 * xq0.y + r.u11 * xq1.y;
    }
    (void)stride_unknowns;
    const double rdet = 1.0 / (r.d00 * r.d11 - r.d01 * r.d01);
    const double i00 = r.d11 * rdet, i01 = -r.d01 * rdet, i11 = r.d00 * rdet;
    o[0] = make_double2(i00 * t00 + i01 * t10, i00 * t01 + i01 * t11);
    o[1] = make_double2(i01 * t00 + i11 * t10, i01 * t01 + i11 * t11);
}

__global__ void reduce_kernel(const Rows in, int64_t rows_in, int64_t rows_out, int64_t stride_unknowns,
                              double *__restrict__ out, BackgroundStatus *st) {
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j < rows_out) reduce_row(in, rows_in, j, stride_unknowns, out, st);
}

__global__ void backsub_kernel(const Rows lvl, int64_t rows, const double *__restrict__ x_next,
                               double *__restrict__ x_out, BackgroundStatus *st, int64_t stride_unknowns) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < rows) backsub_row(lvl, rows, x_next, i, x_out, st, stride_unknowns);
}

// The small end of the recursion (<= BG_SMALL_ROWS block rows) in ONE CTA, in place in shared memory.
// At the level with stride s the active rows sit at positions 0, s, 2s, ...; those at even multiples
// of s are replaced by the next level's rows, the odd ones stay for the substitution back, which
// needs their coupling to the LEFT too -- so here a row carries its L explicitly (16 doubles).  The
// solution overwrites B.
constexpr int SROW = 16;
struct RowS {
    double l00, l01, l10, l11, d00, d01, d11, u00, u01, u10, u11, b00, b01, b10, b11;
};
__device__ __forceinline__ RowS load_rows(const double *p) {
    const double2 *q = reinterpret_cast<const double2 *>(p);
    const double2 a = q[0], b = q[1], c = q[2], d = q[3], e = q[4], f = q[5], g = q[6], h = q[7];
    return RowS{a.x, a.y, b.x, b.y, c.x, c.y, d.x, d.y, e.x, e.y, f.x, f.y, g.x, g.y, h.x};
}
__device__ __forceinline__ void store_rows(double *p, const RowS &r) {
    double2 *q = reinterpret_cast<double2 *>(p);
    q[0] = make_double2(r.l00, r.l01);
    q[1] = make_double2(r.l10, r.l11);
    q[2] = make_double2(r.d00, r.d01);
    q[3] = make_double2(r.d11, r.u00);
    q[4] = make_double2(r.u01, r.u10);
    q[5] = make_double2(r.u11, r.b00);
    q[6] = make_double2(r.b01, r.b10);
    q[7] = make_double2(r.b11, 0.0);
}
__device__ __forceinline__ Inv2 inv_spd2s(const RowS &r, int64_t unknown0, BackgroundStatus *st) {
    Row t{};
    t.d00 = r.d00; t.d01 = r.d01; t.d11 = r.d11;
    return inv_spd2(t, unknown0, st);
}

__device__ __forceinline__ void small_system(const Rows &in, int64_t rows, int64_t stride_unknowns, double *x_out,
                                             BackgroundStatus *st, double *sm) {
    const int R = (int)rows;
    for (int i = threadIdx.x; i < R; i += blockDim.x) {
        const Row r = get_row(in, i, st);
        RowS t{};
        if (i > 0) {  // L_i = U_{i-1}'
            const Row p = get_row(in, i - 1, st);
            t.l00 = p.u00; t.l01 = p.u10; t.l10 = p.u01; t.l11 = p.u11;
        }
        t.d00 = r.d00; t.d01 = r.d01; t.d11 = r.d11;
        t.u00 = r.u00; t.u01 = r.u01; t.u10 = r.u10; t.u11 = r.u11;
        t.b00 = r.b00; t.b01 = r.b01; t.b10 = r.b10; t.b11 = r.b11;
        store_rows(sm + i * SROW, t);
    }
    __syncthreads();
    int s = 1;
    for (; s < R; s <<= 1) {  // reduce: rows at multiples of 2s from their neighbours at +-s
        for (int pos = threadIdx.x * 2 * s; pos < R; pos += blockDim.x * 2 * s) {
            const RowS c = load_rows(sm + pos * SROW);
            RowS o = c;
            o.l00 = o.l01 = o.l10 = o.l11 = 0.0;
            o.u00 = o.u01 = o.u10 = o.u11 = 0.0;
            if (pos - s >= 0) {
                const RowS p = load_rows(sm + (pos - s) * SROW);
                const Inv2 v = inv_spd2s(p, (int64_t)(pos - s) * stride_unknowns, st);
                const double a00 = c.l00 * v.i00 + c.l01 * v.i01, a01 = c.l00 * v.i01 + c.l01 * v.i11;
                const double a10 = c.l10 * v.i00 + c.l11 * v.i01, a11 = c.l10 * v.i01 + c.l11 * v.i11;
                o.d00 -= a00 * p.u00 + a01 * p.u10;
                o.d01 -= a00 * p.u01 + a01 * p.u11;
                o.d11 -= a10 * p.u01 + a11 * p.u11;
                o.b00 -= a00 * p.b00 + a01 * p.b10;
                o.b01 -= a00 * p.b01 + a01 * p.b11;
                o.b10 -= a10 * p.b00 + a11 * p.b10;
                o.b11 -= a10 * p.b01 + a11 * p.b11;
                o.l00 = -(a00 * p.l00 + a01 * p.l10);
                o.l01 = -(a00 * p.l01 + a01 * p.l11);
                o.l10 = -(a10 * p.l00 + a11 * p.l10);
                o.l11 = -(a10 * p.l01 + a11 * p.l11);
            }
            if (pos + s < R) {
                const RowS q = load_rows(sm + (pos + s) * SROW);
                const Inv2 v = inv_spd2s(q, (int64_t)(pos + s) * stride_unknowns, st);
                const double g00 = c.u00 * v.i00 + c.u01 * v.i01, g01 = c.u00 * v.i01 + c.u01 * v.i11;
                const double g10 = c.u10 * v.i00 + c.u11 * v.i01, g11 = c.u10 * v.i01 + c.u11 * v.i11;
                o.d00 -= g00 * q.l00 + g01 * q.l10;
                o.d01 -= g00 * q.l01 + g01 * q.l11;
                o.d11 -= g10 * q.l01 + g11 * q.l11;
                o.b00 -= g00 * q.b00 + g01 * q.b10;
                o.b01 -= g00 * q.b01 + g01 * q.b11;
                o.b10 -= g10 * q.b00 + g11 * q.b10;
                o.b11 -= g10 * q.b01 + g11 * q.b11;
                o.u00 = -(g00 * q.u00 + g01 * q.u10);
                o.u01 = -(g00 * q.u01 + g01 * q.u11);
                o.u10 = -(g10 * q.u00 + g11 * q.u10);
                o.u11 = -(g10 * q.u01 + g11 * q.u11);
            }
            store_rows(sm + pos * SROW, o);
        }
        __syncthreads();
    }
    // row 0 now stands alone
    if (threadIdx.x == 0) {
        RowS r = load_rows(sm);
        const Inv2 v = inv_spd2s(r, 0, st);
        const double t00 = r.b00, t01 = r.b01, t10 = r.b10, t11 = r.b11;
        r.b00 = v.i00 * t00 + v.i01 * t10;
        r.b01 = v.i00 * t01 + v.i01 * t11;
        r.b10 = v.i01 * t00 + v.i11 * t10;
        r.b11 = v.i01 * t01 + v.i11 * t11;
        store_rows(sm, r);
    }
    __syncthreads();
    // back up: rows at odd multiples of h from the solved rows at +-h
    for (int h = s >> 1; h >= 1; h >>= 1) {
        for (int pos = h + (int)threadIdx.x * 2 * h; pos < R; pos += blockDim.x * 2 * h) {
            RowS r = load_rows(sm + pos * SROW);
            const RowS xl = load_rows(sm + (pos - h) * SROW);
            double t00 = r.b00 - (r.l00 * xl.b00 + r.l01 * xl.b10);
            double t01 = r.b01 - (r.l00 * xl.b01 + r.l01 * xl.b11);
            double t10 = r.b10 - (r.l10 * xl.b00 + r.l11 * xl.b10);
            double t11 = r.b11 - (r.l10 * xl.b01 + r.l11 * xl.b11);
            if (pos + h < R) {
                const RowS xr = load_rows(sm + (pos + h) * SROW);
                t00 -= r.u00 * xr.b00 + r.u01 * xr.b10;
                t01 -= r.u00 * xr.b01 + r.u01 * xr.b11;
                t10 -= r.u10 * xr.b00 + r.u11 * xr.b10;
                t11 -= r.u10 * xr.b01 + r.u11 * xr.b11;
            }
            const double rdet = 1.0 / (r.d00 * r.d11 - r.d01 * r.d01);
            const double i00 = r.d11 * rdet, i01 = -r.d01 * rdet, i11 = r.d00 * rdet;
            r.b00 = i00 * t00 + i01 * t10;
            r.b01 = i00 * t01 + i01 * t11;
            r.b10 = i01 * t00 + i11 * t10;
            r.b11 = i01 * t01 + i11 * t11;
            store_rows(sm + pos * SROW, r);
        }
        __syncthreads();
    }
    for (int i = threadIdx.x; i < R; i += blockDim.x) {
        const RowS r = load_rows(sm + i * SROW);
        double2 *o = reinterpret_cast<double2 *>(x_out + (int64_t)i * 4);
        o[0] = make_double2(r.b00, r.b01);
        o[1] = make_double2(r.b10, r.b11);
    }
}

__global__ void __launch_bounds__(1024) small_system_kernel(const Rows in, int64_t rows, int64_t stride_unknowns,
                                                            double *x_out, BackgroundStatus *st) {
    extern __shared__ __align__(16) double sm_small[];
    small_system(in, rows, stride_unknowns, x_out, st, sm_small);
}

// The middle of the recursion (levels with at most BG_MID_ROWS block rows) in ONE cooperative launch:
// every level is a grid-stride loop followed by a grid barrier, CTA 0 runs the shared-memory tail in
// between.  Replaces a dozen launches of a few CTAs each, which cost latency, not traffic.
struct MidArgs {
    const double *lvl[BG_MAX_LEVELS];  // stored rows of a level (nullptr: level 0, formed on the fly)
    double *x[BG_MAX_LEVELS];
    int64_t rows[BG_MAX_LEVELS], stride[BG_MAX_LEVELS];
    Level0 l0;
    int first, small;  // levels [first, small): reduced / substituted here; level `small`: the tail
};

__global__ void __launch_bounds__(BG_MID_THREADS) mid_system_kernel(const MidArgs a, BackgroundStatus *st) {
    namespace cg = cooperative_groups;
    extern __shared__ __align__(16) double sm_mid[];
    cg::grid_group grid = cg::this_grid();
    const int64_t me = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, all = (int64_t)gridDim.x * blockDim.x;
    auto level = [&](int l) {
        Rows r{};
        r.stored = a.lvl[l];
        r.l0 = a.l0;
        return r;
    };
    for (int l = a.first; l < a.small; ++l) {
        const Rows in = level(l);
        for (int64_t j = me; j < a.rows[l + 1]; j += all)
            reduce_row(in, a.rows[l], j, a.stride[l], const_cast<double *>(a.lvl[l + 1]), st);
        grid.sync();
    }
    if (blockIdx.x == 0) small_system(level(a.small), a.rows[a.small], a.stride[a.small], a.x[a.small], st, sm_mid);
    grid.sync();
    for (int l = a.small - 1; l >= a.first; --l) {
        const Rows in = level(l);
        for (int64_t i = me; i < a.rows[l]; i += all) backsub_row(in, a.rows[l], a.x[l + 1], i, a.x[l], st, a.stride[l]);
        grid.sync();
    }
}

// ---- one step of iterative refinement ----
// Block cyclic reduction is backward stable, but on systems with long stretches of zero weight (masked
// regions: only the roughness penalty holds the solution there) its error constant is ~50x that of the
// reference's sequential LDL'.  One refinement step -- residual of both right-hand sides in float64, a
// second solve for the correction -- brings it to the LDL' level (measured with an extended-precision
// yardstick; further steps gain nothing, the residual itself is the limit).
__global__ void __launch_bounds__(256) penta_residual_kernel(const double *__restrict__ w, const double *__restrict__ rhs,
                                                             int64_t n, double lam, double lam1,
                                                             const double *__restrict__ x, double *__restrict__ r0,
                                                             double *__restrict__ r1, BackgroundStatus *st) {
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    auto col = [&](int64_t u) { return *reinterpret_cast<const double2 *>(x + (u >> 1) * 4 + (u & 1) * 2); };
    const double2 xk = col(k);
    const double ad = mat_diag(w, n, k, lam, lam1, st);
    double a0 = ad * xk.x, a1 = ad * xk.y;
    if (k >= 1) {
        const double b = mat_off1(n, k - 1, lam, lam1);
        const double2 v = col(k - 1);
        a0 = fma(b, v.x, a0);
        a1 = fma(b, v.y, a1);
    }
    if (k + 1 < n) {
        const double b = mat_off1(n, k, lam, lam1);
        const double2 v = col(k + 1);
        a0 = fma(b, v.x, a0);
        a1 = fma(b, v.y, a1);
    }
    if (k >= 2) {
        const double c = mat_off2(n, k - 2, lam);
        const double2 v = col(k - 2);
        a0 = fma(c, v.x, a0);
        a1 = fma(c, v.y, a1);
    }
    if (k + 2 < n) {
        const double c = mat_off2(n, k, lam);
        const double2 v = col(k + 2);
        a0 = fma(c, v.x, a0);
        a1 = fma(c, v.y, a1);
    }
    r0[k] = rhs[k] - a0;
    r1[k] = 1.0 - a1;
}

__global__ void __launch_bounds__(256) accumulate_kernel(double *__restrict__ acc, const double *__restrict__ add, int64_t count) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < count) acc[i] += add[i];
}

// ---- zero-sum constraint: column sums of X, then out = x0 - mu x1 ----
constexpr int SUM_THREADS = 256;

__global__ void __launch_bounds__(SUM_THREADS) column_sums_kernel(const double *__restrict__ x, int64_t n,
                                                                  double *__restrict__ partial) {
    // unknown k lives at x[(k / 2) * 4 + (k % 2) * 2 + c]
    __shared__ double s0[SUM_THREADS], s1[SUM_THREADS];
    double a0 = 0.0, a1 = 0.0;
    for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (int64_t)gridDim.x * blockDim.x) {
        const double2 v = *reinterpret_cast<const double2 *>(x + (k >> 1) * 4 + (k & 1) * 2);
        a0 += v.x;
        a1 += v.y;
    }
    s0[threadIdx.x] = a0;
    s1[threadIdx.x] = a1;
    __syncthreads();
    for (int d = SUM_THREADS / 2; d > 0; d >>= 1) {
        if ((int)threadIdx.x < d) {
            s0[threadIdx.x] += s0[threadIdx.x + d];
            s1[threadIdx.x] += s1[threadIdx.x + d];
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        partial[2 * blockIdx.x] = s0[0];
        partial[2 * blockIdx.x + 1] = s1[0];
    }
}

// mu of the zero-sum constraint from the per-block partial sums (one CTA, fixed reduction tree)
__global__ void __launch_bounds__(SUM_THREADS) mu_kernel(const double *__restrict__ partial, int nparts, int64_t n,
                                                         double *__restrict__ mu) {
    __shared__ double s0[SUM_THREADS], s1[SUM_THREADS];
    double t0 = 0.0, t1 = 0.0;
    for (int p = threadIdx.x; p < nparts; p += SUM_THREADS) {
        t0 += partial[2 * p];
        t1 += partial[2 * p + 1];
    }
    s0[threadIdx.x] = t0;
    s1[threadIdx.x] = t1;
    __syncthreads();
    for (int d = SUM_THREADS / 2; d > 0; d >>= 1) {
        if ((int)threadIdx.x < d) {
            s0[threadIdx.x] += s0[threadIdx.x + d];
            s1[threadIdx.x] += s1[threadIdx.x + d];
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) *mu = (fabs(s1[0]) > MIN_PIVOT) ? s0[0] / s1[0] : s0[0] / (double)n;  // pyx:1078-1081
}

__global__ void finish_kernel(const double *__restrict__ x, int64_t n, const double *__restrict__ mu_ptr,
                              double *__restrict__ out) {
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const double2 v = *reinterpret_cast<const double2 *>(x + (k >> 1) * 4 + (k & 1) * 2);
    out[k] = mu_ptr ? v.x - (*mu_ptr) * v.y : v.x;
}

// ---- weighted statistics of the residual matrix (cconsenrich.pyx:9675-9724) ----
// weight[i] = sum_j inv[j][i], rhs[i] = sum_j inv[j][i] * resid[j][i], accumulated over j in order in
// float64 with separately rounded products -- the reference's arithmetic, bit for bit.
constexpr int STAT_THREADS = 256;

__global__ void __launch_bounds__(STAT_THREADS) weighted_stats_kernel(const float *__restrict__ resid,
                                                                      const float *__restrict__ inv, int64_t m,
                                                                      int64_t n, int64_t ld, double *__restrict__ weight,
                                                                      double *__restrict__ rhs,
                                                                      unsigned long long *__restrict__ support) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int pos = 0;
    if (i < n) {
        double ws = 0.0, rs = 0.0;
        for (int64_t j = 0; j < m; ++j) {
            const double w = (double)__ldcs(inv + j * ld + i);
            const double r = (double)__ldcs(resid + j * ld + i);
            ws = __dadd_rn(ws, w);
            rs = __dadd_rn(rs, __dmul_rn(w, r));
        }
        weight[i] = ws;
        rhs[i] = rs;
        pos = ws > 0.0;
    }
    if (support) {
        const unsigned b = __ballot_sync(0xffffffffu, pos);
        if ((threadIdx.x & 31) == 0 && b) atomicAdd(support, (unsigned long long)__popc(b));
    }
}

// ---- core._relativeSignChangePerKB (core.py:2647-2700), the [m x n] part: per interval the state minus
// the inverse-variance weighted mean of (data - background) over the tracks whose cell is valid.  numpy's
// float64 arithmetic in its order (quotient, product, sums rounded separately): bit-identical.
__global__ void __launch_bounds__(STAT_THREADS) weighted_mean_residual_kernel(
    const float *__restrict__ data, const float *__restrict__ munc, int64_t m, int64_t n, int64_t ld,
    const double *__restrict__ state, const double *__restrict__ background, double pad, double *__restrict__ out) {
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const double sv = state[k], bg = background ? background[k] : 0.0;
    const bool state_ok = isfinite(sv);
    double total = 0.0, wsum = 0.0;
    for (int64_t j = 0; j < m; ++j) {
        const double d = (double)__ldcs(data + j * ld + k);
        const double den = __dadd_rn((double)__ldcs(munc + j * ld + k), pad);
        if (state_ok && isfinite(d) && isfinite(den) && den > 0.0) {
            const double w = __ddiv_rn(1.0, den > 1.0e-12 ? den : 1.0e-12);
            total = __dadd_rn(total, __dmul_rn(__dsub_rn(d, bg), w));
            wsum = __dadd_rn(wsum, w);
        }
    }
    const double mean = wsum > 0.0 ? __ddiv_rn(total, wsum) : nan("");
    out[k] = __dsub_rn(sv, mean);
}

// ---- core._perIntervalOutputDiagnosticTracks (core.py:7734-7880): the two parts that cost time ----
// (1) muncTrace / sumInvR: per interval, sums over the tracks of the effective observation variance and of
//     its inverse, finite terms only (core.py:7786-7800);
// (2) the per-interval loop over the one-step predicted covariance and the summed Kalman gain
//     (core.py:7840-7866) -- interval k needs only the filtered covariance of interval k - 1, so the
//     reference's Python loop of n iterations with 2x2 matrix products is one thread per interval here.
__global__ void __launch_bounds__(STAT_THREADS) diag_obs_sums_kernel(const float *__restrict__ munc, int64_t m, int64_t n,
                                                                     int64_t ld, const double *__restrict__ obs_prec,
                                                                     double pad, double *__restrict__ munc_trace,
                                                                     double *__restrict__ sum_inv_r) {
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const double lam = obs_prec[k];
    double tr = 0.0, si = 0.0;
    for (int64_t j = 0; j < m; ++j) {
        double v = __dadd_rn((double)__ldcs(munc + j * ld + k), pad);
        if (!(v >= 1.0e-12)) v = v != v ? v : 1.0e-12;  // np.maximum propagates NaN
        const double eff = __ddiv_rn(v, lam), inv = __ddiv_rn(lam, v);
        if (isfinite(eff)) tr = __dadd_rn(tr, eff);
        if (isfinite(inv)) si = __dadd_rn(si, inv);
    }
    munc_trace[k] = tr;
    sum_inv_r[k] = si;
}

__global__ void __launch_bounds__(STAT_THREADS) diag_gain_kernel(const DiagGainArgs a) {
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= a.n) return;
    const int d = a.dim, dd = a.cov_dim * a.cov_dim;
    // process noise of interval k (core.py:7841-7851)
    double q00, q01 = 0.0, q10 = 0.0, q11 = 0.0;
    bool from_track = false;
    if (a.p_noise && k > 0) {
        const float *pn = a.p_noise + (k - 1) * dd;
        const double p00 = (double)pn[0];
        bool fin = isfinite(p00);
        if (d == 2) {
            fin = fin && isfinite((double)pn[1]) && isfinite((double)pn[a.cov_dim]) && isfinite((double)pn[a.cov_dim + 1]);
        }
        if (fin) {
            from_track = true;
            q00 = p00;
            if (d == 2) {
                q01 = (double)pn[1];
                q10 = (double)pn[a.cov_dim];
                q11 = (double)pn[a.cov_dim + 1];
            }
        }
    }
    if (!from_track) {
        const double s = a.q_scale[k];
        q00 = __dmul_rn(a.base_q[0], s);
        if (d == 2) {
            q01 = __dmul_rn(a.base_q[1], s);
            q10 = __dmul_rn(a.base_q[2], s);
            q11 = __dmul_rn(a.base_q[3], s);
        }
        if (!(a.p_noise && k > 0) && a.proc_prec) {  // the division sits in the `else` branch of core.py:7847
            const double kp = a.proc_prec[k];
            q00 = __ddiv_rn(q00, kp);
            if (d == 2) {
                q01 = __ddiv_rn(q01, kp);
                q10 = __ddiv_rn(q10, kp);
                q11 = __ddiv_rn(q11, kp);
            }
        }
    }
    // filtered covariance of the interval before (the prior for k = 0)
    double p00, p01 = 0.0, p10 = 0.0, p11 = 0.0;
    if (k == 0) {
        p00 = a.cov_init;
        p11 = a.cov_init;
    } else {
        const float *pc = a.covar + (k - 1) * dd;
        p00 = (double)pc[0];
        if (d == 2) {
            p01 = (double)pc[1];
            p10 = (double)pc[a.cov_dim];
            p11 = (double)pc[a.cov_dim + 1];
        }
    }
    double pred00, pred10 = 0.0;
    if (d == 2) {
        // (F P) F' + Q, products and sums rounded separately (numpy matmul without fused multiply-add)
        const double f00 = a.f[0], f01 = a.f[1], f10 = a.f[2], f11 = a.f[3];
        const double t00 = __dadd_rn(__dmul_rn(f00, p00), __dmul_rn(f01, p10));
        const double t01 = __dadd_rn(__dmul_rn(f00, p01), __dmul_rn(f01, p11));
        const double t10 = __dadd_rn(__dmul_rn(f10, p00), __dmul_rn(f11, p10));
        const double t11 = __dadd_rn(__dmul_rn(f10, p01), __dmul_rn(f11, p11));
        pred00 = __dadd_rn(__dadd_rn(__dmul_rn(t00, f00), __dmul_rn(t01, f01)), q00);
        pred10 = __dadd_rn(__dadd_rn(__dmul_rn(t10, f00), __dmul_rn(t11, f01)), q10);
        (void)q01;
        (void)q11;
    } else {
        pred00 = __dadd_rn(p00, q00);
    }
    if (!(pred00 > 0.0)) pred00 = pred00 != pred00 ? pred00 : 0.0;  // max(float(x), 0.0) keeps a NaN first argument
    const double sir = a.sum_inv_r[k];
    const double denom = __dadd_rn(1.0, __dmul_rn(pred00, sir));
    double g0 = 0.0, g1 = 0.0;
    if (isfinite(denom) && denom > 0.0) {
        const double gs = __ddiv_rn(sir, denom);
        g0 = __dmul_rn(pred00, gs);
        g1 = __dmul_rn(pred10, gs);
    }
    a.sum_gain0[k] = g0;
    a.sum_gain1[k] = g1;
}

}  // namespace

// =====================================================================================
// pivot parity with the reference on degenerate systems
// =====================================================================================
// The reference fails a solve when a pivot of its SEQUENTIAL pentadiagonal LDL' falls below 1e-12
// (cconsenrich.pyx:1016-1055, 1090-1095).  Block cyclic reduction meets different pivots, so on a singular
// system it may sail through (one positive weight under a second-difference penalty: the null space is
// not pinned down, the reference raises, the parallel solve returns ~1e14) or flag a different index.  The
// sequential recurrence is two divisions per unknown and cannot be parallelised, but it only has to run
// when the outcome is in doubt: the parallel solve flagged a pivot, or the weights cannot pin the penalty's
// null space (fewer than three positive weights).  Then one warp replays the reference's recurrence -- lanes
// stream the weights in, lane 0 runs it -- and ITS verdict (first bad index and value, or none) replaces
// the parallel one.
__global__ void __launch_bounds__(256) positive_weight_count_kernel(const double *__restrict__ w, int64_t n,
                                                                    unsigned long long *count) {
    unsigned long long c = 0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        c += w[i] > 0.0 ? 1ull : 0ull;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) c += __shfl_xor_sync(0xffffffffu, c, d);
    if ((threadIdx.x & 31) == 0 && c) atomicAdd(count, c);
}

__global__ void __launch_bounds__(32) ldl_pivot_check_kernel(const double *__restrict__ w, int64_t n, double lam,
                                                             double lam1, const unsigned long long *positive,
                                                             BackgroundStatus *st) {
    if (st->bad_index == STATUS_NONE && *positive >= 3ull) return;  // not in doubt
    __shared__ double sw[32];
    const int lane = threadIdx.x;
    long long bad = -1;
    double bad_value = 0.0;
    // ---- the reference's first loop (pyx:1016-1031): entries of the assembled diagonal below the floor; the
    //      first of them is its verdict whatever the factorisation meets afterwards ----
    {
        long long first = 0x7fffffffffffffffLL;
        for (int64_t i = lane; i < n; i += 32) {
            const double d = w[i] + first_diag(n, i, lam1) + second_diag(n, i, lam);
            if (d < MIN_PIVOT) {
                first = i;
                break;
            }
        }
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) first = min(first, __shfl_xor_sync(0xffffffffu, first, d));
        if (first != 0x7fffffffffffffffLL) {
            if (lane == 0) {
                st->bad_index = first;
                st->bad_value = w[first] + first_diag(n, first, lam1) + second_diag(n, first, lam);
            }
            return;
        }
    }
    // ---- the factorisation's pivots (pyx:1033-1055) ----
    double d1 = 0.0, d2 = 0.0, l1 = 0.0;  // pivots i-1, i-2; first lower diagonal at i-1
    double r = lane < n ? w[lane] : 0.0;
    for (int64_t c0 = 0; c0 < n; c0 += 32) {
        sw[lane] = r;
        __syncwarp();
        if (c0 + 32 + lane < n) r = w[c0 + 32 + lane];
        if (lane == 0) {
            const int cnt = (int)min((int64_t)32, n - c0);
            for (int u = 0; u < cnt; ++u) {
                const int64_t i = c0 + u;
                double d = __dadd_rn(__dadd_rn(sw[u], first_diag(n, i, lam1)), second_diag(n, i, lam));
                // every operation rounded on its own, in the reference's order (no fused multiply-add: a pivot that
                // is zero in exact arithmetic is pure rounding, and the message quotes it)
                if (i == 1) {
                    l1 = __ddiv_rn(__dadd_rn(first_off1(n, lam1), second_off1(n, 0, lam)), d1);
                    d = __dadd_rn(d, -__dmul_rn(__dmul_rn(l1, l1), d1));
                } else if (i >= 2) {
                    l1 = __ddiv_rn(__dadd_rn(__dadd_rn(first_off1(n, lam1), second_off1(n, i - 1, lam)), -__dmul_rn(lam, l1)), d1);
                    d = __dadd_rn(__dadd_rn(d, -__dmul_rn(__dmul_rn(l1, l1), d1)), -__ddiv_rn(__dmul_rn(lam, lam), d2));
                }
                if (i >= 1 && d < MIN_PIVOT) {
                    if (bad < 0) {
                        bad = i;
                        bad_value = d;
                    }
                    d = MIN_PIVOT;
                }
                d2 = d1;
                d1 = d;
            }
        }
        __syncwarp();
    }
    if (lane == 0) {
        st->bad_index = bad >= 0 ? bad : STATUS_NONE;
        st->bad_value = bad_value;
    }
}

// =====================================================================================
// host side
// =====================================================================================
int64_t background_rows(int64_t n) { return (n + 1) / 2; }

size_t background_workspace_bytes(int64_t n) {
    // stored rows of levels 1.. (level 0 is formed on the fly) + solutions of all levels + partial sums
    int64_t rows = background_rows(n), total_rows = 0, total_x = 0;
    bool first = true;
    while (true) {
        if (!first) total_rows += rows;
        total_x += rows;
        first = false;
        if (rows <= 1) break;
        rows = (rows + 1) / 2;
    }
    // + the accumulated solution and the two residual vectors of the refinement step
    const size_t refine = (size_t)background_rows(n) * 4 * 8 + 2 * (size_t)(n + 2) * 8;
    return (size_t)total_rows * ROW * 8 + (size_t)total_x * 4 * 8 + refine + (size_t)(BG_SUM_BLOCKS * 2 + 2) * 8 + 256;
}

cudaError_t launch_background_stats(const float *resid, const float *inv, int64_t m, int64_t n, int64_t ld,
                                    double *weight, double *rhs, unsigned long long *support, cudaStream_t st) {
    if (n <= 0) return cudaSuccess;
    if (support) {
        cudaError_t e = cudaMemsetAsync(support, 0, sizeof(unsigned long long), st);
        if (e != cudaSuccess) return e;
    }
    weighted_stats_kernel<<<(unsigned)((n + STAT_THREADS - 1) / STAT_THREADS), STAT_THREADS, 0, st>>>(
        resid, inv, m, n, ld, weight, rhs, support);
    return cudaGetLastError();
}

cudaError_t launch_diag_obs_sums(const float *munc, int64_t m, int64_t n, int64_t ld, const double *obs_prec, double pad,
                                 double *munc_trace, double *sum_inv_r, cudaStream_t st) {
    if (n <= 0) return cudaSuccess;
    diag_obs_sums_kernel<<<(unsigned)((n + STAT_THREADS - 1) / STAT_THREADS), STAT_THREADS, 0, st>>>(munc, m, n, ld, obs_prec,
                                                                                                  pad, munc_trace, sum_inv_r);
    return cudaGetLastError();
}

cudaError_t launch_diag_gain(const DiagGainArgs &a, cudaStream_t st) {
    if (a.n <= 0) return cudaSuccess;
    diag_gain_kernel<<<(unsigned)((a.n + STAT_THREADS - 1) / STAT_THREADS), STAT_THREADS, 0, st>>>(a);
    return cudaGetLastError();
}

cudaError_t launch_weighted_mean_residual(const float *data, const float *munc, int64_t m, int64_t n, int64_t ld,
                                          const double *state, const double *background, double pad, double *out,
                                          cudaStream_t st) {
    if (n <= 0) return cudaSuccess;
    weighted_mean_residual_kernel<<<(unsigned)((n + STAT_THREADS - 1) / STAT_THREADS), STAT_THREADS, 0, st>>>(
        data, munc, m, n, ld, state, background, pad, out);
    return cudaGetLastError();
}

cudaError_t launch_background_solve(const double *w, const double *rhs, int64_t n, double lam, double lam_first,
                                    int zero_center, double *out, void *workspace, BackgroundStatus *status,
                                    cudaStream_t st, int *launches) {
    if (n <= 0) return cudaSuccess;
    // levels: rows[0] = ceil(n / 2) block rows formed on the fly from w / rhs, each following level half of it
    int64_t rows[BG_MAX_LEVELS], stride[BG_MAX_LEVELS];
    double *lvl[BG_MAX_LEVELS], *x[BG_MAX_LEVELS];
    int nl = 0;
    {
        int64_t r = background_rows(n), s = 2;
        while (true) {
            rows[nl] = r;
            stride[nl] = s;
            ++nl;
            if (r <= 1) break;
            r = (r + 1) / 2;
            s *= 2;
        }
    }
    // the large levels get one launch each way; from the first level with <= BG_SMALL_ROWS rows on,
    // one CTA finishes the recursion in shared memory
    int first_small = 0;
    while (rows[first_small] > BG_SMALL_ROWS) ++first_small;
    double *p = static_cast<double *>(workspace);
    lvl[0] = nullptr;  // never stored
    for (int l = 1; l <= first_small; ++l) {
        lvl[l] = p;
        p += rows[l] * ROW;
    }
    for (int l = 0; l <= first_small; ++l) {
        x[l] = p;
        p += rows[l] * 4;
    }
    double *x_acc = p;  // refinement: the accumulated solution, then the two residual vectors
    p += rows[0] * 4;
    double *res0 = p;
    p += n + 2;
    double *res1 = p;
    p += n + 2;
    double *partial = p;
    double *mu_slot = partial + 2 * BG_SUM_BLOCKS + 2;  // one word of scratch after the zero-sum multiplier
    Level0 l0{w, rhs, n, lam, lam_first, nullptr};
    auto level = [&](int l) {
        Rows r{};
        r.stored = lvl[l];
        r.l0 = l0;
        return r;
    };
    cudaError_t e = cudaMemsetAsync(status, 0x7f, sizeof(BackgroundStatus), st);  // bad_index = STATUS_NONE
    if (e != cudaSuccess) return e;
    // the levels: [0, first_mid) one launch each way; [first_mid, first_small) + the shared-memory tail in
    // one cooperative launch when the device takes it (co-resident CTAs), else launch by launch
    int first_mid = 0;
    while (first_mid < first_small && rows[first_mid] > BG_MID_ROWS) ++first_mid;
    const size_t tail_smem = (size_t)rows[first_small] * SROW * 8;
    int count = 0;
    const int T = 256;
    int dev = 0, sms = 0, coop = 0, per_sm = 0;
    if (first_mid < first_small && !getenv("CB200_BG_NO_COOP") && cudaGetDevice(&dev) == cudaSuccess &&
        cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev) == cudaSuccess && coop &&
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess &&
        cudaFuncSetAttribute(mid_system_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tail_smem) == cudaSuccess &&
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, mid_system_kernel, BG_MID_THREADS, tail_smem) == cudaSuccess &&
        per_sm > 0) {
        coop = 1;
    } else {
        coop = 0;
        first_mid = first_small;
    }
    // one pass down and up the levels for the right-hand sides in l0; the solution lands in x[0]
    auto run_levels = [&]() -> cudaError_t {
        for (int l = 0; l < first_mid; ++l) {
            reduce_kernel<<<(unsigned)((rows[l + 1] + T - 1) / T), T, 0, st>>>(level(l), rows[l], rows[l + 1], stride[l],
                                                                             lvl[l + 1], status);
            ++count;
        }
        if (coop) {
            MidArgs ma{};
            for (int l = 0; l <= first_small; ++l) {
                ma.lvl[l] = lvl[l];
                ma.x[l] = x[l];
                ma.rows[l] = rows[l];
                ma.stride[l] = stride[l];
            }
            ma.l0 = l0;
            ma.first = first_mid;
            ma.small = first_small;
            int64_t want = (rows[first_mid] + BG_MID_THREADS - 1) / BG_MID_THREADS;
            const int64_t cap = (int64_t)sms * per_sm;
            if (want > cap) want = cap;
            if (want < 1) want = 1;
            BackgroundStatus *stp = status;
            void *kargs[] = {&ma, &stp};
            cudaError_t ec = cudaLaunchCooperativeKernel(reinterpret_cast<void *>(mid_system_kernel), dim3((unsigned)want),
                                                         dim3(BG_MID_THREADS), kargs, tail_smem, st);
            if (ec != cudaSuccess) return ec;
            ++count;
        } else {
            // per device, so set on every call (a few hundred nanoseconds)
            cudaError_t ec = cudaFuncSetAttribute(small_system_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                  BG_SMALL_ROWS * SROW * 8);
            if (ec != cudaSuccess) return ec;
            small_system_kernel<<<1, 1024, tail_smem, st>>>(level(first_small), rows[first_small], stride[first_small],
                                                           x[first_small], status);
            ++count;
        }
        for (int l = first_mid - 1; l >= 0; --l) {
            backsub_kernel<<<(unsigned)((rows[l] + T - 1) / T), T, 0, st>>>(level(l), rows[l], x[l + 1], x[l], status,
                                                                          stride[l]);
            ++count;
        }
        return cudaGetLastError();
    };
    e = run_levels();
    if (e != cudaSuccess) return e;
    const double *solution = x[0];
    if (!getenv("CB200_BG_NO_REFINE")) {
        // one refinement step: X += A^-1 (B - A X) for both right-hand sides
        const int64_t xcount = rows[0] * 4;
        e = cudaMemcpyAsync(x_acc, x[0], (size_t)xcount * 8, cudaMemcpyDeviceToDevice, st);
        if (e != cudaSuccess) return e;
        penta_residual_kernel<<<(unsigned)((n + T - 1) / T), T, 0, st>>>(w, rhs, n, lam, lam_first, x_acc, res0, res1, status);
        ++count;
        l0.rhs = res0;
        l0.rhs2 = res1;
        e = run_levels();
        if (e != cudaSuccess) return e;
        accumulate_kernel<<<(unsigned)((xcount + T - 1) / T), T, 0, st>>>(x_acc, x[0], xcount);
        ++count;
        solution = x_acc;
    }
    double *mu = nullptr;
    if (zero_center) {
        int nparts = (int)((n + SUM_THREADS - 1) / SUM_THREADS);
        if (nparts > BG_SUM_BLOCKS) nparts = BG_SUM_BLOCKS;
        mu = partial + 2 * BG_SUM_BLOCKS;
        column_sums_kernel<<<nparts, SUM_THREADS, 0, st>>>(solution, n, partial);
        mu_kernel<<<1, SUM_THREADS, 0, st>>>(partial, nparts, n, mu);
        count += 2;
    }
    finish_kernel<<<(unsigned)((n + T - 1) / T), T, 0, st>>>(solution, n, mu, out);
    ++count;
    if (n >= 2) {
        // the reference's own pivot verdict where the parallel one is in doubt (see ldl_pivot_check_kernel)
        unsigned long long *positive = reinterpret_cast<unsigned long long *>(mu_slot);
        e = cudaMemsetAsync(positive, 0, sizeof(unsigned long long), st);
        if (e != cudaSuccess) return e;
        int blocks = (int)((n + 255) / 256);
        if (blocks > 592) blocks = 592;
        positive_weight_count_kernel<<<blocks, 256, 0, st>>>(w, n, positive);
        ldl_pivot_check_kernel<<<1, 32, 0, st>>>(w, n, lam, lam_first, positive, status);
        count += 2;
    }
    if (launches) *launches += count;
    return cudaGetLastError();
}

}  // namespace cb200
