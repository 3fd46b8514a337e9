// consenrich_b200/csrc/ssm_kernels.cu -- sm_100a kernels of the Consenrich state-space hot path.
//
//   fold_kernel        one coalesced pass over data/munc [m x n] -> per-bin information-form
//                      statistics (replaces _accumulateObservationValue, cconsenrich.pyx:259-283)
//   scan_kernel<Fwd*>  forward Kalman filter as a single-pass decoupled look-back scan of
//                      filtering elements (replaces the loops at cconsenrich.pyx:291-529, 538-707)
//   scan_kernel<Bwd*>  RTS smoother as the reverse scan of smoothing elements
//                      (replaces cconsenrich.pyx:6740-6848, 7116-7148)
//   residual_kernel    postFitResiduals [n x m] = data^T - level   (cconsenrich.pyx:6846-6848)
//   lambda/kappa       Student-t precision multipliers (cconsenrich.pyx:8210-8298, 7474-7521)
//
// The path is HBM-bound byte/double work: no tensor cores.  Global traffic is coalesced
// (bins are the contiguous axis), per-thread bin chunks are transposed through padded shared
// memory records, and the only inter-CTA dependency is the look-back on published tile
// aggregates.
#include <cuda_runtime.h>

#include <cstdlib>

#include "ssm_kernels.cuh"

namespace cb200 {

namespace {

constexpr unsigned FULL = 0xffffffffu;
constexpr int NWARPS = SCAN_THREADS / 32;
#ifndef FWD2_MIN_CTAS
#define FWD2_MIN_CTAS 3
#endif
#ifndef BWD2_MIN_CTAS
#define BWD2_MIN_CTAS 4
#endif
#ifndef SCAN_MIN_CTAS
#define SCAN_MIN_CTAS 4
#endif

__device__ __forceinline__ int ld_acquire(const int *p) {
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ long long gtimer() {
    long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ void st_release(int *p, int v) {
    asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

template <class E>
__device__ __forceinline__ E shfl_up_elem(const E &e, int d) {
    E r;
    const double *s = reinterpret_cast<const double *>(&e);
    double *t = reinterpret_cast<double *>(&r);
#pragma unroll
    for (int i = 0; i < E::N; ++i) t[i] = __shfl_up_sync(FULL, s[i], d);
    return r;
}
template <class E>
__device__ __forceinline__ E shfl_down_elem(const E &e, int d) {
    E r;
    const double *s = reinterpret_cast<const double *>(&e);
    double *t = reinterpret_cast<double *>(&r);
#pragma unroll
    for (int i = 0; i < E::N; ++i) t[i] = __shfl_down_sync(FULL, s[i], d);
    return r;
}
template <class E>
__device__ __forceinline__ E shfl_bcast_elem(const E &e, int src) {
    E r;
    const double *s = reinterpret_cast<const double *>(&e);
    double *t = reinterpret_cast<double *>(&r);
#pragma unroll
    for (int i = 0; i < E::N; ++i) t[i] = __shfl_sync(FULL, s[i], src);
    return r;
}
template <class E>
__device__ __forceinline__ void store_elem(double *dst, const E &e) {
    const double *s = reinterpret_cast<const double *>(&e);
#pragma unroll
    for (int i = 0; i < E::N; ++i) dst[i] = s[i];
}
template <class E>
__device__ __forceinline__ E load_elem(const double *src) {
    E r;
    double *t = reinterpret_cast<double *>(&r);
#pragma unroll
    for (int i = 0; i < E::N; ++i) t[i] = src[i];
    return r;
}
template <class E>
__device__ __forceinline__ E load_elem_cg(const double *src) {
    E r;
    double *t = reinterpret_cast<double *>(&r);
#pragma unroll
    for (int i = 0; i < E::N; ++i) t[i] = __ldcg(src + i);
    return r;
}

// =====================================================================================
// fold kernel
// =====================================================================================
struct FoldAcc {
    double s0, s1, s2, prod;
    int esum;
};

__device__ __forceinline__ void fold_cell(FoldAcc &a, float zf, float vf, double pad) {
    const double z = (double)zf;
    double r = (double)vf + pad;
    if (r < 1.0e-12) r = 1.0e-12;
    const double w = cb_rcp(r);
    const double wz = w * z;
    a.s0 += w;
    a.s1 += wz;
    a.s2 = fma(wz, z, a.s2);
    // sum of logs as the log of a running product; fold_renorm moves its exponent into an integer
    // counter every FOLD_RENORM cells, before it can leave the double range (1e-12 <= r <= ~1e30 per cell)
    a.prod *= r;
}

__device__ __forceinline__ void fold_renorm(FoldAcc &a) {
    const long long bits = __double_as_longlong(a.prod);
    a.esum += (int)((bits >> 52) & 0x7ff) - 1023;
    a.prod = __longlong_as_double((bits & 0x000fffffffffffffLL) | 0x3ff0000000000000LL);
}

__device__ __forceinline__ double fold_logsum(const FoldAcc &a) {
    return log(a.prod) + (double)a.esum * 0.693147180559945309417232121458;
}

constexpr int FOLD_THREADS = 256;
#ifndef FOLD_MIN_CTAS
#define FOLD_MIN_CTAS 4
#endif
constexpr int FOLD_UNROLL = 4;
constexpr int FOLD_RENORM = 4;  // cells between renormalisations of the running product: at most 7 with the tail rows, and (3.4e38)^7, (1e-12)^7 fit a double

// VEC = 4: each thread owns 4 consecutive bins and streams float4 (needs 16-byte aligned rows);
// VEC = 1: scalar loads, any alignment.
template <int VEC>
__global__ void __launch_bounds__(FOLD_THREADS, FOLD_MIN_CTAS)
fold_kernel(const float *__restrict__ data, const float *__restrict__ munc, int64_t m, int64_t n, int64_t ld,
            double pad, double2 *__restrict__ SA, double2 *__restrict__ SB, int rm_logL) {
    const int64_t k0 = ((int64_t)blockIdx.x * FOLD_THREADS + threadIdx.x) * VEC;
    if (k0 >= n) return;
    FoldAcc acc[VEC];
#pragma unroll
    for (int c = 0; c < VEC; ++c) {
        acc[c].s0 = acc[c].s1 = acc[c].s2 = 0.0;
        acc[c].prod = 1.0;
        acc[c].esum = 0;
    }
    const float *dp = data + k0;
    const float *vp = munc + k0;
    int64_t j = 0;
    int since = 0;
    if (VEC == 4) {
        for (; j + FOLD_UNROLL <= m; j += FOLD_UNROLL) {
            float4 z[FOLD_UNROLL], v[FOLD_UNROLL];
#pragma unroll
            for (int u = 0; u < FOLD_UNROLL; ++u) {
                z[u] = __ldcs(reinterpret_cast<const float4 *>(dp + (j + u) * ld));
                v[u] = __ldcs(reinterpret_cast<const float4 *>(vp + (j + u) * ld));
            }
#pragma unroll
            for (int u = 0; u < FOLD_UNROLL; ++u) {
                fold_cell(acc[0], z[u].x, v[u].x, pad);
                fold_cell(acc[1 % VEC], z[u].y, v[u].y, pad);
                fold_cell(acc[2 % VEC], z[u].z, v[u].z, pad);
                fold_cell(acc[3 % VEC], z[u].w, v[u].w, pad);
            }
            since += FOLD_UNROLL;
            if (since >= FOLD_RENORM) {
#pragma unroll
                for (int c = 0; c < VEC; ++c) fold_renorm(acc[c]);
                since = 0;
            }
        }
        for (; j < m; ++j) {
            const float4 z = __ldcs(reinterpret_cast<const float4 *>(dp + j * ld));
            const float4 v = __ldcs(reinterpret_cast<const float4 *>(vp + j * ld));
            fold_cell(acc[0], z.x, v.x, pad);
            fold_cell(acc[1 % VEC], z.y, v.y, pad);
            fold_cell(acc[2 % VEC], z.z, v.z, pad);
            fold_cell(acc[3 % VEC], z.w, v.w, pad);
        }
    } else {
        for (; j + FOLD_UNROLL <= m; j += FOLD_UNROLL) {
            float z[FOLD_UNROLL], v[FOLD_UNROLL];
#pragma unroll
            for (int u = 0; u < FOLD_UNROLL; ++u) {
                z[u] = __ldcs(dp + (j + u) * ld);
                v[u] = __ldcs(vp + (j + u) * ld);
            }
#pragma unroll
            for (int u = 0; u < FOLD_UNROLL; ++u) fold_cell(acc[0], z[u], v[u], pad);
            since += FOLD_UNROLL;
            if (since >= FOLD_RENORM) {
                fold_renorm(acc[0]);
                since = 0;
            }
        }
        for (; j < m; ++j) fold_cell(acc[0], __ldcs(dp + j * ld), __ldcs(vp + j * ld), pad);
    }
#pragma unroll
    for (int c = 0; c < VEC; ++c) {
        const int64_t k = k0 + c;
        if (k < n) {
            // rm_logL >= 0: run-major output for the lean sweeps (lean_kernels.cuh), runs of 1 << rm_logL bins
            int64_t pos = k;
            if (rm_logL >= 0) {
                const int64_t sb = (int64_t)32 << rm_logL;
                const int64_t r = k & (sb - 1);
                pos = (k - r) + ((r & (((int64_t)1 << rm_logL) - 1)) << 5) + (r >> rm_logL);
            }
            SA[pos] = make_double2(acc[c].s0, acc[c].s1);
            SB[pos] = make_double2(acc[c].s2, fold_logsum(acc[c]));
        }
    }
}

// =====================================================================================
// scan traits
// =====================================================================================
// Geometry.  A tile is SCAN_THREADS runs of L = CHUNK * nsub consecutive scan positions, one
// run per thread.  A run is processed in nsub sub-steps of CHUNK positions.  For each sub-step
// the CTA stages the SCAN_THREADS x CHUNK records of all its threads in shared memory with
// cp.async (LDGSTS: no registers, groups of CHUNK consecutive bins = whole sectors), double
// buffered so that the copies of sub-step s+1 fly while sub-step s is computed.  Long runs
// amortise the warp scan, the cross-warp prefix and the look-back over L positions.
//
// Shared-memory layout of a sub-step buffer: BIN_BYTES / 16 planes of 16-byte cells, one cell per
// (thread, slot).  Two access patterns have to be free of bank conflicts: a thread walking its own
// CHUNK slots (lanes t, t+1, ... read the same slot index: stride of one thread) and the staging
// copies (lanes 4o .. 4o+3 touch the CHUNK slots of owner o: stride of one slot).  With 16-byte
// cells a quarter warp covers all 32 banks exactly once in both patterns when slot i of thread t
// sits at cell  t * CHUNK + ((i + (t >> 1)) mod CHUNK):  the rotation by t >> 1 spreads equal slot
// indices of neighbouring threads over the four bank groups that 4 t leaves open.
struct Cells {  // the cells of one thread
    unsigned char *base;  // plane 0 of the buffer
    int t4, rot;
    __device__ __forceinline__ unsigned char *at(int i, int plane, int plane_bytes) const {
        return base + plane * plane_bytes + ((t4 + ((i + rot) & (CHUNK - 1))) << 4);
    }
};

template <int BIN_BYTES>
struct RecGeom {
    static_assert(BIN_BYTES % 16 == 0, "records are made of 16-byte cells");
    static constexpr int PLANES = BIN_BYTES / 16;
    static constexpr int PLANE_BYTES = SCAN_THREADS * CHUNK * 16;
    static constexpr int BUF_BYTES = PLANES * PLANE_BYTES;
    __device__ static __forceinline__ Cells cells(unsigned char *buf, int thread) {
        return Cells{buf, thread * CHUNK, (thread >> 1) & (CHUNK - 1)};
    }
    // 16-byte cell `plane` of staged record g (thread g / CHUNK, slot g % CHUNK)
    __device__ static __forceinline__ unsigned char *cell(unsigned char *buf, int g, int plane) {
        return cells(buf, g / CHUNK).at(g % CHUNK, plane, PLANE_BYTES);
    }
};

// Staging is warp-private: the 32 threads of a warp copy in and store out the records of their
// own 32 runs (element r of lane l is record l + 32 r of the warp's 32 * CHUNK), so the sub-step
// loops need only __syncwarp() and the warps of a CTA drift freely between the two CTA-wide
// barriers around the scan section.
__device__ __forceinline__ int stage_elem(int tid, int r) { return (tid & ~31) * CHUNK + (tid & 31) + 32 * r; }

// scan position of staged record g (thread g / CHUNK, slot g % CHUNK) in sub-step s
__device__ __forceinline__ int64_t stage_pos(int64_t p0, int L, int s, int g) {
    return p0 + (int64_t)((g / CHUNK) * L + s * CHUNK + (g % CHUNK));
}

// cp.async of BYTES (4, 8 or 16) bytes; !ok zero-fills the destination without touching src
template <int BYTES>
__device__ __forceinline__ void cp_async(void *dst_smem, const void *src, bool ok) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(dst_smem);
    const int sz = ok ? BYTES : 0;
    if (BYTES == 16)
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(src), "r"(sz) : "memory");
    else if (BYTES == 8)
        asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(d), "l"(src), "r"(sz) : "memory");
    else
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(d), "l"(src), "r"(sz) : "memory");
}
// same, source known to be valid (no src-size operand)
template <int BYTES>
__device__ __forceinline__ void cp_async_full(void *dst_smem, const void *src) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(dst_smem);
    if (BYTES == 16)
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(src) : "memory");
    else if (BYTES == 8)
        asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d), "l"(src) : "memory");
    else
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(d), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// ---- forward, 2-state ------------------------------------------------------------------
// record per bin, in : [S0 S1][S2 SL][kappa qscale lambda (f32) pad]      (48 bytes, raw copies)
//                 out: [Pf 4xf32][Qf 4xf32][xf 2xf32, D f32, pad]
// scan position = bin index k.
struct FwdRaw {
    double qk, lam;
};

// qk = qScale / clamp(kappa), lam = clamp(lambda) from the raw float triple of a record
// qk = qScale / clamp(kappa) from the raw float32 multipliers
__device__ __forceinline__ double fwd_qk(const FwdArgs &a, float kap, float qscale) {
    // branch-free (selects): the replay loops hoist these ahead of the serial per-bin chain
    const double qs = a.use_qscale ? (double)qscale : 1.0;
    const double kc = a.use_kappa ? clampd((double)kap, a.kap_min, a.kap_max) : 1.0;
    return cb_div(qs, kc);
}

__device__ __forceinline__ FwdRaw fwd_raw(const FwdArgs &a, const unsigned char *cell2) {
    const float4 w = *reinterpret_cast<const float4 *>(cell2);
    FwdRaw r;
    r.qk = fwd_qk(a, w.x, w.y);
    // the reference clamps in double and uses the double value (pyx:432-435)
    r.lam = a.use_lambda ? clampd((double)w.z, a.lam_min, a.lam_max) : 1.0;
    return r;
}

// process-noise scale of bin k read straight from global memory (run boundaries, head replay)
__device__ __forceinline__ double fwd_qk_global(const FwdArgs &a, int64_t k) {
    return fwd_qk(a, a.use_kappa ? a.kap[k] : 1.0f, a.use_qscale ? a.qs[k] : 1.0f);
}

template <bool ALL>
__device__ __forceinline__ void fwd_issue(const FwdArgs &a, unsigned char *buf, int64_t p0, int L, int s, int tid) {
    using G = RecGeom<48>;
#pragma unroll
    for (int r = 0; r < CHUNK; ++r) {
        const int g = stage_elem(tid, r);
        const int64_t k = stage_pos(p0, L, s, g);
        const bool ok = k < a.n;
        const int64_t kk = ok ? k : 0;
        cp_async<16>(G::cell(buf, g, 0), a.SA + kk, ok);
        if (ALL) cp_async<16>(G::cell(buf, g, 1), a.SB + kk, ok);
        unsigned char *d2 = G::cell(buf, g, 2);
        if (a.use_kappa) cp_async<4>(d2, a.kap + kk, ok);
        if (a.use_qscale) cp_async<4>(d2 + 4, a.qs + kk, ok);
        if (a.use_lambda) cp_async<4>(d2 + 8, a.lam + kk, ok);
    }
}

template <bool CANON>
struct Fwd2 {
    using Elem = Filt2;
    using State = State2;
    using Args = FwdArgs;
    using G = RecGeom<48>;
    struct Carry {
        Kf2 s;
        NllAcc acc;
        Smo2 brun;  // smoothing element of the bins replayed so far (FwdArgs::smo_run)
        int nb;     // bins replayed so far
    };
    static constexpr bool HAS_SUMS = true;
    static constexpr int MIN_CTAS = FWD2_MIN_CTAS;  // 3: 168 registers: the replay also carries the smoother's run element
    __device__ static __forceinline__ double *sums(const Args &a) { return a.sums; }
    __device__ static __forceinline__ bool prebuilt(const Args &) { return false; }
    __device__ static __forceinline__ Elem load_prebuilt(const Args &, int64_t) { return filt2_identity(); }

    __device__ static __forceinline__ Elem identity() { return filt2_identity(); }
    __device__ static __forceinline__ Elem combine(const Elem &a, const Elem &b) { return filt2_combine(a, b); }
    __device__ static __forceinline__ State apply(const Elem &e, const State &s) { return filt2_apply(e, s); }
    __device__ static __forceinline__ Elem from_state(const State &s) { return filt2_from_state(s); }
    __device__ static __forceinline__ State elem_state(const Elem &e) {
        return State2{e.b0, e.b1, e.C00, e.C01, e.C11};
    }
    __device__ static __forceinline__ State initial(const Args &a) {
        if (a.init_state) return load_elem_cg<State2>(a.init_state);
        return State2{a.state_init, 0.0, a.cov_init, 0.0, a.cov_init};
    }
    // valid slots [lo, hi) of the CHUNK positions starting at q0
    __device__ static __forceinline__ void bounds(const Args &a, int64_t q0, int &lo, int &hi) {
        const int64_t rem = a.n - q0;
        lo = 0;
        hi = rem <= 0 ? 0 : (rem >= CHUNK ? CHUNK : (int)rem);
    }
    // ALL = false: only what pass 1 reads (S0, S1 and the raw multipliers)
    template <bool ALL>
    __device__ static __forceinline__ void issue(const Args &a, unsigned char *buf, int64_t p0, int L, int s, int tid) {
        fwd_issue<ALL>(a, buf, p0, L, s, tid);
    }

    template <bool FULLC>
    __device__ static __forceinline__ void pass1(const Args &a, const Cells &rec, int lo, int hi, int64_t,
                                                 Elem &g) {
        // the multipliers of the CHUNK bins first: independent of one another and of the chain below
        FwdRaw wr[CHUNK];
#pragma unroll
        for (int i = 0; i < CHUNK; ++i) wr[i] = fwd_raw(a, rec.at(i, 2, G::PLANE_BYTES));
#pragma unroll
        for (int i = 0; i < CHUNK; ++i) {
            if (FULLC || (i >= lo && i < hi)) {
                const double2 s01 = *reinterpret_cast<const double2 *>(rec.at(i, 0, G::PLANE_BYTES));
                const FwdRaw w = wr[i];
                filt2_step<CANON>(g, a.M, w.qk * a.M.q00, w.qk * a.M.q01, w.qk * a.M.q11, w.lam * s01.x,
                                  w.lam * s01.y);
            }
        }
    }

    __device__ static __forceinline__ Carry begin2(const Args &, const State &st) {
        Carry c;
        c.s = Kf2{r32(st.x0), r32(st.x1), r32(st.P00), r32(st.P01), r32(st.P01), r32(st.P11)};
        nll_acc_init(c.acc);
        c.brun = smo2_identity();
        c.nb = 0;
        return c;
    }
    // smoothing element of a bin from its filtered state (float32 values, as stored) and the float32
    // process noise of the NEXT bin, composed onto the run element (later bins first: they are the
    // ones the reverse scan meets first)
    __device__ static __forceinline__ void compose_smo(const Args &a, Smo2 &brun, const Kf2 &f, double Q00, double Q01,
                                                       double Q10, double Q11) {
        const Rts2 r = rts2_gain<CANON>(a.M, f.x0, f.x1, f.P00, f.P01, f.P10, f.P11, r32(Q00), r32(Q01), r32(Q10), r32(Q11));
        brun = smo2_combine(smo2_from_rts(r, f.x0, f.x1, f.P00, f.P01, f.P11), brun);
    }
    // After the replay of a run: its last bin's element (needs the NEXT run's first process noise, or
    // is the terminal element of the chromosome), then the run element goes to global memory.
    __device__ static __forceinline__ void finish_run(const Args &a, int64_t run, int64_t run0, int L, Carry &c) {
        if (!a.smo_run) return;
        const int64_t next = run0 + L;  // first bin of the next run
        if (c.nb > 0) {
            if (next >= a.n) {
                // the run holds the last bin of the chromosome: x_s = x_f, P_s = P_f there
                c.brun = smo2_combine(smo2_from_state(State2{c.s.x0, c.s.x1, c.s.P00, c.s.P01, c.s.P11}), c.brun);
            } else {
                const double qk = fwd_qk_global(a, next);
                compose_smo(a, c.brun, c.s, qk * a.M.q00, qk * a.M.q01, qk * a.M.q10, qk * a.M.q11);
            }
        }
        const double *e = reinterpret_cast<const double *>(&c.brun);
#pragma unroll
        for (int i = 0; i < Smo2::N; ++i) a.smo_run[i * a.smo_pitch + run] = e[i];
    }
    // sum of the run's NLL pieces (one pair of logs per run)
    __device__ static __forceinline__ double finish2(const Args &a, const Carry &c) {
        return (a.want_nll && !a.nll_in_d) ? nll_acc_finish(c.acc, a.m, a.mlog2pi) : 0.0;
    }

    template <bool FULLC>
    __device__ static __forceinline__ void pass2(const Args &a, const Cells &rec, int lo, int hi, int64_t q0,
                                                 Carry &c, double &acc_d, double &acc_nll) {
        Kf2 &s = c.s;
        NllAcc &acc = c.acc;
        const bool per_bin = a.nll_in_d != 0;
        // bins [head_from, HEAD_BINS) are written and summed by the head replay (epilogue)
        const bool near_head = q0 < HEAD_BINS && q0 + CHUNK > a.head_from;
        FwdRaw wr[CHUNK];
#pragma unroll
        for (int i = 0; i < CHUNK; ++i) wr[i] = fwd_raw(a, rec.at(i, 2, G::PLANE_BYTES));
#pragma unroll
        for (int i = 0; i < CHUNK; ++i) {
            if (FULLC || (i >= lo && i < hi)) {
                unsigned char *b0 = rec.at(i, 0, G::PLANE_BYTES), *b1 = rec.at(i, 1, G::PLANE_BYTES), *b2 = rec.at(i, 2, G::PLANE_BYTES);
                const double2 s01 = *reinterpret_cast<const double2 *>(b0);
                const double2 s2l = *reinterpret_cast<const double2 *>(b1);
                const FwdRaw w = wr[i];
                const bool counted = !(near_head && q0 + i >= a.head_from && q0 + i < HEAD_BINS);
                const Kf2 prev = s;  // filtered state of the previous bin, float32 values
                BinOut o;
                kf2_step<CANON>(s, a.M, w.qk, w.lam, s01.x, s01.y, s2l.x, s2l.y, a.m, a.inv_m, a.mlog2pi,
                                a.want_nll != 0 && counted, per_bin, o, acc);
                if (a.smo_run) {
                    if (c.nb > 0) compose_smo(a, c.brun, prev, o.Q00, o.Q01, o.Q10, o.Q11);
                    c.nb += 1;
                }
                const float d = (float)o.stat;
                if (counted) acc_d += (double)d;
                acc_nll += o.nll;
                *reinterpret_cast<float4 *>(b0) = make_float4((float)s.P00, (float)s.P01, (float)s.P10, (float)s.P11);
                *reinterpret_cast<float4 *>(b1) = make_float4((float)o.Q00, (float)o.Q01, (float)o.Q10, (float)o.Q11);
                *reinterpret_cast<float4 *>(b2) = make_float4((float)s.x0, (float)s.x1, d, 0.0f);
            }
        }
        if (a.want_nll && !per_bin) nll_acc_renorm(acc, a.use_lambda != 0);
    }

    __device__ static __forceinline__ void stage_out(const Args &a, unsigned char *recs, int64_t p0, int L, int s,
                                                     int tid) {
#pragma unroll
        for (int r = 0; r < CHUNK; ++r) {
            const int g = stage_elem(tid, r);
            const int64_t k = stage_pos(p0, L, s, g);
            // bins [L, HEAD_BINS) of an unsharded chromosome are written by the head replay instead
            const bool head = k >= a.head_from && k < HEAD_BINS;
            if (k < a.n && !head) {
                const float4 xd = *reinterpret_cast<const float4 *>(G::cell(recs, g, 2));
                if (a.do_store) {
                    reinterpret_cast<float4 *>(a.Pf)[k] = *reinterpret_cast<const float4 *>(G::cell(recs, g, 0));
                    if (k > 0) reinterpret_cast<float4 *>(a.Qf)[k - 1] = *reinterpret_cast<const float4 *>(G::cell(recs, g, 1));
                    reinterpret_cast<float2 *>(a.xf)[k] = make_float2(xd.x, xd.y);
                }
                // Q of the shard's first bin belongs to row n-1 of the preceding shard
                if (k == 0 && a.q_head) *reinterpret_cast<float4 *>(a.q_head) = *reinterpret_cast<const float4 *>(G::cell(recs, g, 1));
                if (a.D) a.D[k] = xd.z;
            }
        }
    }

    // Head replay (see HEAD_BINS): the first thread of the first tile carries its replay on through
    // bins [L, HEAD_BINS), reading the inputs straight from global memory.
    __device__ static __forceinline__ void epilogue(const Args &a, int tile, int tid, int L, Carry &c, double &acc_d,
                                                    double &acc_nll) {
        if (tile != 0 || tid != 0 || a.head_from >= HEAD_BINS) return;
        const int64_t end = a.n < HEAD_BINS ? a.n : HEAD_BINS;
        Kf2 &s = c.s;
        NllAcc acc;
        nll_acc_init(acc);
        for (int64_t k = L; k < end; ++k) {
            const double2 s01 = a.SA[k], s2l = a.SB[k];
            const double qk = fwd_qk_global(a, k);
            const double lam = a.use_lambda ? clampd((double)a.lam[k], a.lam_min, a.lam_max) : 1.0;
            BinOut o;
            kf2_step<CANON>(s, a.M, qk, lam, s01.x, s01.y, s2l.x, a.want_nll ? s2l.y : 0.0, a.m, a.inv_m, a.mlog2pi,
                            a.want_nll != 0, a.nll_in_d != 0, o, acc);
            const float d = (float)o.stat;
            acc_d += (double)d;
            acc_nll += o.nll;
            if (a.do_store) {
                reinterpret_cast<float4 *>(a.Pf)[k] = make_float4((float)s.P00, (float)s.P01, (float)s.P10, (float)s.P11);
                reinterpret_cast<float4 *>(a.Qf)[k - 1] = make_float4((float)o.Q00, (float)o.Q01, (float)o.Q10, (float)o.Q11);
                reinterpret_cast<float2 *>(a.xf)[k] = make_float2((float)s.x0, (float)s.x1);
            }
            if (a.D) a.D[k] = d;
        }
        if (a.want_nll && !a.nll_in_d) acc_nll += nll_acc_finish(acc, a.m, a.mlog2pi);
    }
};

// ---- forward, level --------------------------------------------------------------------
// same input record; out: [xf Pf Qf D] in the first 16 bytes
struct Fwd1 {
    using Elem = Filt1;
    using State = State1;
    using Args = FwdArgs;
    using G = RecGeom<48>;
    struct Carry {
        State1 s;
        NllAcc acc;
    };
    static constexpr bool HAS_SUMS = true;
    static constexpr int MIN_CTAS = SCAN_MIN_CTAS;
    __device__ static __forceinline__ double *sums(const Args &a) { return a.sums; }

    __device__ static __forceinline__ Elem identity() { return filt1_identity(); }
    __device__ static __forceinline__ Elem combine(const Elem &a, const Elem &b) { return filt1_combine(a, b); }
    __device__ static __forceinline__ State apply(const Elem &e, const State &s) { return filt1_apply(e, s); }
    __device__ static __forceinline__ Elem from_state(const State &s) { return filt1_from_state(s); }
    __device__ static __forceinline__ State elem_state(const Elem &e) { return State1{e.b, e.C}; }
    __device__ static __forceinline__ State initial(const Args &a) {
        if (a.init_state) return load_elem_cg<State1>(a.init_state);
        return State1{a.state_init, a.cov_init};
    }
    __device__ static __forceinline__ void bounds(const Args &a, int64_t q0, int &lo, int &hi) {
        Fwd2<false>::bounds(a, q0, lo, hi);
    }
    template <bool ALL>
    __device__ static __forceinline__ void issue(const Args &a, unsigned char *buf, int64_t p0, int L, int s, int tid) {
        fwd_issue<ALL>(a, buf, p0, L, s, tid);
    }
    template <bool FULLC>
    __device__ static __forceinline__ void pass1(const Args &a, const Cells &rec, int lo, int hi, int64_t,
                                                 Elem &g) {
#pragma unroll
        for (int i = 0; i < CHUNK; ++i) {
            if (FULLC || (i >= lo && i < hi)) {
                const double2 s01 = *reinterpret_cast<const double2 *>(rec.at(i, 0, G::PLANE_BYTES));
                const FwdRaw w = fwd_raw(a, rec.at(i, 2, G::PLANE_BYTES));
                filt1_step(g, w.qk * a.M.q00, w.lam * s01.x, w.lam * s01.y);
            }
        }
    }
    __device__ static __forceinline__ Carry begin2(const Args &, const State &st) {
        Carry c;
        c.s = st;
        nll_acc_init(c.acc);
        return c;
    }
    __device__ static __forceinline__ double finish2(const Args &a, const Carry &c) {
        return (a.want_nll && !a.nll_in_d) ? nll_acc_finish(c.acc, a.m, a.mlog2pi) : 0.0;
    }
    __device__ static __forceinline__ void finish_run(const Args &, int64_t, int64_t, int, Carry &) {}
    __device__ static __forceinline__ bool prebuilt(const Args &) { return false; }
    __device__ static __forceinline__ Elem load_prebuilt(const Args &, int64_t) { return filt1_identity(); }
    __device__ static __forceinline__ void epilogue(const Args &, int, int, int, Carry &, double &, double &) {}
    template <bool FULLC>
    __device__ static __forceinline__ void pass2(const Args &a, const Cells &rec, int lo, int hi, int64_t,
                                                 Carry &c, double &acc_d, double &acc_nll) {
        State1 &s = c.s;
        NllAcc &acc = c.acc;
        const bool per_bin = a.nll_in_d != 0;
#pragma unroll
        for (int i = 0; i < CHUNK; ++i) {
            if (FULLC || (i >= lo && i < hi)) {
                unsigned char *b0 = rec.at(i, 0, G::PLANE_BYTES), *b1 = rec.at(i, 1, G::PLANE_BYTES), *b2 = rec.at(i, 2, G::PLANE_BYTES);
                const double2 s01 = *reinterpret_cast<const double2 *>(b0);
                const double2 s2l = *reinterpret_cast<const double2 *>(b1);
                const FwdRaw w = fwd_raw(a, b2);
                BinOut o;
                kf1_step(s, w.qk * a.M.q00, w.lam, s01.x, s01.y, s2l.x, s2l.y, a.m, a.inv_m, a.mlog2pi,
                         a.want_nll != 0, per_bin, o, acc);
                const float d = (float)o.stat;
                acc_d += (double)d;
                acc_nll += o.nll;
                *reinterpret_cast<float4 *>(b0) = make_float4((float)s.x, (float)s.P, (float)o.Q00, d);
            }
        }
        if (a.want_nll && !per_bin) nll_acc_renorm(acc, a.use_lambda != 0);
    }
    __device__ static __forceinline__ void stage_out(const Args &a, unsigned char *recs, int64_t p0, int L, int s,
                                                     int tid) {
#pragma unroll
        for (int r = 0; r < CHUNK; ++r) {
            const int g = stage_elem(tid, r);
            const int64_t k = stage_pos(p0, L, s, g);
            if (k < a.n) {
                const float4 o = *reinterpret_cast<const float4 *>(G::cell(recs, g, 0));
                if (a.do_store) {
                    a.xf[k] = o.x;
                    a.Pf[k] = o.y;
                    if (k > 0) a.Qf[k - 1] = o.z;
                }
                if (k == 0 && a.q_head) a.q_head[0] = o.z;
                if (a.D) a.D[k] = o.w;
            }
        }
    }
};

// ---- backward, 2-state -----------------------------------------------------------------
// scan position q runs against the bins: k = npad - 1 - q with npad = n rounded up to CHUNK, so
// that every staged group of CHUNK positions is a CHUNK-aligned block of bins (whole sectors);
// the first npad - n positions are empty.  record per bin, in: [Pf][Qf row k][xf, pad];
// out: [Ps][lag][xs]
template <bool CANON>
struct Bwd2 {
    using Elem = Smo2;
    using State = State2;
    using Args = BwdArgs;
    using G = RecGeom<48>;
    using Carry = Rs2;
    static constexpr bool HAS_SUMS = false;
    static constexpr int MIN_CTAS = BWD2_MIN_CTAS;
    __device__ static __forceinline__ double *sums(const Args &) { return nullptr; }
    __device__ static __forceinline__ int64_t npad(const Args &a) {
        return a.npad_fixed > 0 ? a.npad_fixed : (a.n + CHUNK - 1) / CHUNK * CHUNK;
    }
    __device__ static __forceinline__ bool prebuilt(const Args &a) { return a.smo_run != nullptr; }
    // run element composed by the forward scan: backward run `run` is forward run (runs - 1 - run)
    __device__ static __forceinline__ Elem load_prebuilt(const Args &a, int64_t fwd_run) {
        Elem e;
        double *d = reinterpret_cast<double *>(&e);
#pragma unroll
        for (int i = 0; i < Elem::N; ++i) d[i] = __ldcg(a.smo_run + i * a.smo_pitch + fwd_run);
        return e;
    }
    __device__ static __forceinline__ void finish_run(const Args &, int64_t, int64_t, int, Carry &) {}

    __device__ static __forceinline__ Elem identity() { return smo2_identity(); }
    __device__ static __forceinline__ Elem combine(const Elem &a, const Elem &b) { return smo2_combine(a, b); }
    __device__ static __forceinline__ State apply(const Elem &e, const State &s) { return smo2_apply(e, s); }
    __device__ static __forceinline__ Elem from_state(const State &s) { return smo2_from_state(s); }
    __device__ static __forceinline__ State elem_state(const Elem &e) {
        return State2{e.g0, e.g1, e.L00, e.L01, e.L11};
    }
    __device__ static __forceinline__ State initial(const Args &a) {
        if (a.tail_state) return load_elem_cg<State2>(a.tail_state);
        return State2{0.0, 0.0, 0.0, 0.0, 0.0};
    }
    __device__ static __forceinline__ void bounds(const Args &a, int64_t q0, int &lo, int &hi) {
        const int64_t np = npad(a);
        const int64_t first = (np - a.n) - q0;  // positions below npad - n hold no bin
        const int64_t rem = np - q0;
        lo = first <= 0 ? 0 : (first >= CHUNK ? CHUNK : (int)first);
        hi = rem <= 0 ? 0 : (rem >= CHUNK ? CHUNK : (int)rem);
    }
    template <bool ALL>
    __device__ static __forceinline__ void issue(const Args &a, unsigned char *buf, int64_t p0, int L, int s, int tid) {
        const int64_t np = npad(a);
        // the thread's CHUNK records are 8 runs apart: bins k0 - 8 L r
        const int g0 = stage_elem(tid, 0);
        const int64_t k0 = np - 1 - stage_pos(p0, L, s, g0);
        const int64_t dk = (int64_t)(32 / CHUNK) * L;
        const bool want_qs = ALL && a.kap_out && a.qs;
        if (k0 - (CHUNK - 1) * dk >= 0 && k0 < a.n - 1) {  // all inside the track, none the last bin
#pragma unroll
            for (int r = 0; r < CHUNK; ++r) {
                const int g = g0 + 32 * r;
                const int64_t k = k0 - r * dk;
                unsigned char *d2 = G::cell(buf, g, 2);
                cp_async_full<16>(G::cell(buf, g, 0), reinterpret_cast<const float4 *>(a.Pf) + k);
                cp_async_full<16>(G::cell(buf, g, 1), reinterpret_cast<const float4 *>(a.Qf) + k);
                cp_async_full<8>(d2, reinterpret_cast<const float2 *>(a.xf) + k);
                if (want_qs) cp_async_full<4>(d2 + 12, a.qs + k + 1);
            }
            return;
        }
#pragma unroll
        for (int r = 0; r < CHUNK; ++r) {
            const int g = g0 + 32 * r;
            const int64_t k = k0 - r * dk;
            const bool ok = k >= 0 && k < a.n;
            const int64_t kk = ok ? k : 0;
            unsigned char *d2 = G::cell(buf, g, 2);
            cp_async<16>(G::cell(buf, g, 0), reinterpret_cast<const float4 *>(a.Pf) + kk, ok);
            // row n-1 of Qf holds nothing unless a following shard supplied it
            cp_async<16>(G::cell(buf, g, 1), reinterpret_cast<const float4 *>(a.Qf) + kk,
                         ok && (k < a.n - 1 || !a.is_last_shard));
            cp_async<8>(d2, reinterpret_cast<const float2 *>(a.xf) + kk, ok);
            if (want_qs) cp_async<4>(d2 + 12, a.qs + (kk + 1 < a.n ? kk + 1 : kk), ok);
        }
    }
    // slot (0..CHUNK-1) of the sub-step starting at position q0 that holds the chromosome's last bin
    // (x_s = x_f there), or -1
    __device__ static __forceinline__ int last_bin_slot(const Args &a, int64_t q0) {
        const int64_t di = (npad(a) - 1 - q0) - (a.n - 1);
        return (a.is_last_shard && di >= 0 && di < CHUNK) ? (int)di : -1;
    }
    template <bool FULLC>
    __device__ static __forceinline__ void pass1(const Args &a, const Cells &rec, int lo, int hi, int64_t q0,
                                                 Elem &g) {
        const int i_end = last_bin_slot(a, q0);
#pragma unroll
        for (int i = 0; i < CHUNK; ++i) {
            if (FULLC || (i >= lo && i < hi)) {
                const float4 P = *reinterpret_cast<const float4 *>(rec.at(i, 0, G::PLANE_BYTES));
                const float4 Q = *reinterpret_cast<const float4 *>(rec.at(i, 1, G::PLANE_BYTES));
                const float2 x = *reinterpret_cast<const float2 *>(rec.at(i, 2, G::PLANE_BYTES));
                Elem e;
                if (i == i_end) {
                    e = smo2_from_state(State2{(double)x.x, (double)x.y, (double)P.x, (double)P.y, (double)P.w});
                } else {
                    const Rts2 r = rts2_gain<CANON>(a.M, x.x, x.y, P.x, P.y, P.z, P.w, Q.x, Q.y, Q.z, Q.w);
                    e = smo2_from_rts(r, x.x, x.y, P.x, P.y, P.w);
                }
                g = smo2_combine(g, e);
            }
        }
    }
    __device__ static __forceinline__ Carry begin2(const Args &, const State &st) {
        return Rs2{r32(st.x0), r32(st.x1), r32(st.P00), r32(st.P01), r32(st.P01), r32(st.P11)};
    }
    __device__ static __forceinline__ double finish2(const Args &, const Carry &) { return 0.0; }
    __device__ static __forceinline__ void epilogue(const Args &, int, int, int, Carry &, double &, double &) {}
    template <bool FULLC>
    __device__ static __forceinline__ void pass2(const Args &a, const Cells &rec, int lo, int hi, int64_t q0,
                                                 Carry &c, double &, double &) {
        const int i_end = last_bin_slot(a, q0);
#pragma unroll
        for (int i = 0; i < CHUNK; ++i) {
            if (FULLC || (i >= lo && i < hi)) {
                unsigned char *b0 = rec.at(i, 0, G::PLANE_BYTES), *b1 = rec.at(i, 1, G::PLANE_BYTES), *b2 = rec.at(i, 2, G::PLANE_BYTES);
                const float4 P = *reinterpret_cast<const float4 *>(b0);
                const float4 Q = *reinterpret_cast<const float4 *>(b1);
                const float2 x = *reinterpret_cast<const float2 *>(b2);
                if (i == i_end) {
                    c = Rs2{(double)x.x, (double)x.y, (double)P.x, (double)P.y, (double)P.z, (double)P.w};
                    // xs = xf, Ps = Pf already in place; lag row n-1 does not exist
                } else {
                    const Rts2 r = rts2_gain<CANON>(a.M, x.x, x.y, P.x, P.y, P.z, P.w, Q.x, Q.y, Q.z, Q.w);
                    const Rs2 nxt = c;  // smoothed bin k+1 as the reference stores it (float32 values)
                    const float qs_next = *reinterpret_cast<const float *>(b2 + 12);
                    Smo2Out o;
                    rts2_step(c, r, x.x, x.y, P.x, P.y, P.w, o);
                    if (!a.no_store) {
                        *reinterpret_cast<float4 *>(b0) = make_float4((float)o.S00, (float)o.S01, (float)o.S01, (float)o.S11);
                        *reinterpret_cast<float4 *>(b1) = make_float4((float)o.C00, (float)o.C01, (float)o.C10, (float)o.C11);
                    }
                    float kv = 1.0f;
                    if (a.kap_out) {
                        // the reference reads its float32 tracks back: c now holds bin k rounded that way
                        kv = (float)kappa2_update<CANON>(a.M, a.qi00, a.qi01, a.qi10, a.qi11, c.x0, c.x1, c.P00, c.P01, c.P10,
                                                         c.P11, nxt.x0, nxt.x1, nxt.P00, nxt.P01, nxt.P10, nxt.P11, r32(o.C00),
                                                         r32(o.C01), r32(o.C10), r32(o.C11), (double)qs_next, a.qs != nullptr,
                                                         a.nu, a.kap_lo, a.kap_hi);
                    }
                    *reinterpret_cast<float4 *>(b2) = make_float4((float)o.xs0, (float)o.xs1, kv, 0.0f);
                }
            }
        }
    }
    __device__ static __forceinline__ void stage_out(const Args &a, unsigned char *recs, int64_t p0, int L, int s,
                                                     int tid) {
        const int64_t np = npad(a);
        const int g0 = stage_elem(tid, 0);
        const int64_t k0 = np - 1 - stage_pos(p0, L, s, g0);
        const int64_t dk = (int64_t)(32 / CHUNK) * L;
#pragma unroll
        for (int r = 0; r < CHUNK; ++r) {
            const int g = g0 + 32 * r;
            const int64_t k = k0 - r * dk;
            if (k >= 0 && k < a.n) {
                const float4 xk = *reinterpret_cast<const float4 *>(G::cell(recs, g, 2));
                if (!a.no_store) {
                    reinterpret_cast<float4 *>(a.Ps)[k] = *reinterpret_cast<const float4 *>(G::cell(recs, g, 0));
                    if (k < a.lag_rows && (k < a.n - 1 || !a.is_last_shard))
                        reinterpret_cast<float4 *>(a.lag)[k] = *reinterpret_cast<const float4 *>(G::cell(recs, g, 1));
                    reinterpret_cast<float2 *>(a.xs)[k] = make_float2(xk.x, xk.y);
                }
                if (a.kap_out) {
                    if (k + 1 < a.n) a.kap_out[k + 1] = xk.z;  // multiplier of the transition k -> k+1
                    if (k == 0) a.kap_out[0] = 1.0f;
                }
            }
        }
    }
};

// ---- backward, level -------------------------------------------------------------------
// record per bin (16 bytes), in: [xf Pf Qf pad]; out: [xs Ps lag pad]
struct Bwd1 {
    using Elem = Smo1;
    using State = State1;
    using Args = BwdArgs;
    using G = RecGeom<16>;
    struct Carry {
        double x, P;
    };
    static constexpr bool HAS_SUMS = false;
    static constexpr int MIN_CTAS = SCAN_MIN_CTAS;
    __device__ static __forceinline__ double *sums(const Args &) { return nullptr; }
    __device__ static __forceinline__ int64_t npad(const Args &a) { return (a.n + CHUNK - 1) / CHUNK * CHUNK; }
    __device__ static __forceinline__ bool prebuilt(const Args &) { return false; }
    __device__ static __forceinline__ Elem load_prebuilt(const Args &, int64_t) { return smo1_identity(); }
    __device__ static __forceinline__ void finish_run(const Args &, int64_t, int64_t, int, Carry &) {}

    __device__ static __forceinline__ Elem identity() { return smo1_identity(); }
    __device__ static __forceinline__ Elem combine(const Elem &a, const Elem &b) { return smo1_combine(a, b); }
    __device__ static __forceinline__ State apply(const Elem &e, const State &s) { return smo1_apply(e, s); }
    __device__ static __forceinline__ Elem from_state(const State &s) { return smo1_from_state(s); }
    __device__ static __forceinline__ State elem_state(const Elem &e) { return State1{e.g, e.L}; }
    __device__ static __forceinline__ State initial(const Args &a) {
        if (a.tail_state) return load_elem_cg<State1>(a.tail_state);
        return State1{0.0, 0.0};
    }
    __device__ static __forceinline__ void bounds(const Args &a, int64_t q0, int &lo, int &hi) {
        const int64_t np = npad(a);
        const int64_t first = (np - a.n) - q0;
        const int64_t rem = np - q0;
        lo = first <= 0 ? 0 : (first >= CHUNK ? CHUNK : (int)first);
        hi = rem <= 0 ? 0 : (rem >= CHUNK ? CHUNK : (int)rem);
    }
    template <bool ALL>
    __device__ static __forceinline__ void issue(const Args &a, unsigned char *buf, int64_t p0, int L, int s, int tid) {
        const int64_t np = npad(a);
#pragma unroll
        for (int r = 0; r < CHUNK; ++r) {
            const int g = stage_elem(tid, r);
            const int64_t k = np - 1 - stage_pos(p0, L, s, g);
            const bool ok = k >= 0 && k < a.n;
            const int64_t kk = ok ? k : 0;
            unsigned char *d = G::cell(buf, g, 0);
            cp_async<4>(d, a.xf + kk, ok);
            cp_async<4>(d + 4, a.Pf + kk, ok);
            cp_async<4>(d + 8, a.Qf + kk, ok && (k < a.n - 1 || !a.is_last_shard));
            if (ALL && a.kap_out && a.qs) cp_async<4>(d + 12, a.qs + (kk + 1 < a.n ? kk + 1 : kk), ok);
        }
    }
    template <bool FULLC>
    __device__ static __forceinline__ void pass1(const Args &a, const Cells &rec, int lo, int hi, int64_t q0,
                                                 Elem &g) {
        const int64_t klast = npad(a) - 1 - q0;
#pragma unroll
        for (int i = 0; i < CHUNK; ++i) {
            if (FULLC || (i >= lo && i < hi)) {
                const float4 v = *reinterpret_cast<const float4 *>(rec.at(i, 0, G::PLANE_BYTES));
                Elem e;
                if (klast - i == a.n - 1 && a.is_last_shard) {
                    e = smo1_from_state(State1{(double)v.x, (double)v.y});
                } else {
                    double pp, J;
                    rts1_gain((double)v.y, (double)v.z, pp, J);
                    e = smo1_from_rts((double)v.x, (double)v.y, pp, J);
                }
                g = smo1_combine(g, e);
            }
        }
    }
    __device__ static __forceinline__ Carry begin2(const Args &, const State &st) { return Carry{r32(st.x), r32(st.P)}; }
    __device__ static __forceinline__ double finish2(const Args &, const Carry &) { return 0.0; }
    __device__ static __forceinline__ void epilogue(const Args &, int, int, int, Carry &, double &, double &) {}
    template <bool FULLC>
    __device__ static __forceinline__ void pass2(const Args &a, const Cells &rec, int lo, int hi, int64_t q0,
                                                 Carry &c, double &, double &) {
        const int64_t klast = npad(a) - 1 - q0;
#pragma unroll
        for (int i = 0; i < CHUNK; ++i) {
            if (FULLC || (i >= lo && i < hi)) {
                unsigned char *b = rec.at(i, 0, G::PLANE_BYTES);
                const float4 v = *reinterpret_cast<const float4 *>(b);
                if (klast - i == a.n - 1 && a.is_last_shard) {
                    c.x = (double)v.x;
                    c.P = (double)v.y;
                } else {
                    const double xf = (double)v.x, pf = (double)v.y;
                    double pp, J;
                    rts1_gain(pf, (double)v.z, pp, J);
                    const double xsv = xf + J * (c.x - xf);
                    const double dP = c.P - pp;
                    double ps = pf + (J * J * dP);
                    if (ps < 0.0) ps = 0.0;
                    const float xs32 = (float)xsv, ps32 = (float)ps, lag32 = (float)(pf + (J * dP));
                    float kv = 1.0f;
                    if (a.kap_out)
                        kv = (float)kappa1_update(a.qi00, (double)xs32, (double)ps32, c.x, c.P, (double)lag32, (double)v.w,
                                                  a.qs != nullptr, a.nu, a.kap_lo, a.kap_hi);
                    *reinterpret_cast<float4 *>(b) = make_float4(xs32, ps32, lag32, kv);
                    c.x = (double)xs32;
                    c.P = (double)ps32;
                }
            }
        }
    }
    __device__ static __forceinline__ void stage_out(const Args &a, unsigned char *recs, int64_t p0, int L, int s,
                                                     int tid) {
        const int64_t np = npad(a);
#pragma unroll
        for (int r = 0; r < CHUNK; ++r) {
            const int g = stage_elem(tid, r);
            const int64_t k = np - 1 - stage_pos(p0, L, s, g);
            if (k >= 0 && k < a.n) {
                const float4 o = *reinterpret_cast<const float4 *>(G::cell(recs, g, 0));
                if (!a.no_store) {
                    a.xs[k] = o.x;
                    a.Ps[k] = o.y;
                    if (k < a.lag_rows && (k < a.n - 1 || !a.is_last_shard)) a.lag[k] = o.z;
                }
                if (a.kap_out) {
                    if (k + 1 < a.n) a.kap_out[k + 1] = o.w;
                    if (k == 0) a.kap_out[0] = 1.0f;
                }
            }
        }
    }
};

// =====================================================================================
// decoupled look-back (one warp of the CTA)
// =====================================================================================
// Cooperative look-back.  When every tile of a wave finishes composing its run elements at about
// the same time, no tile has a published prefix yet.  A single warp walking back 32 aggregates
// per round then needs one round per 32 tiles, and waiting for intermediate results of other
// tiles costs several L2 round trips per generation.  Here the WHOLE CTA (its other warps would
// only wait at a barrier anyway) reduces a window of SCAN_THREADS * c preceding tiles at once:
// thread j takes the c tiles base - j c ... base - j c - (c - 1), composes them in order, the
// warps reduce by an ordered shuffle tree and the four warp results are combined through shared
// memory.  Nothing but plain aggregates (flag 1) and final prefixes (flag 2) is ever published, a
// tile depends on earlier tiles only, and one window of c = 4 spans 512 tiles.
//   flag - epoch4:  1 = aggregate published, 2 = inclusive prefix published
template <class Tr>
__device__ __forceinline__ int flag_state(const ScanWorkspace &ws, int idx) {
    const int v = ld_acquire(ws.flags + idx) - ws.epoch4;
    return (v == 1 || v == 2) ? v : 0;
}

constexpr int LB_MAX_PER_THREAD = 4;

// Exclusive prefix state of `tile` (> 0); every thread of the CTA calls it and gets the result.
// sd_red: NWARPS * N doubles of shared scratch, sd_int: NWARPS ints.
template <class Tr>
__device__ typename Tr::State cooperative_lookback(const typename Tr::Args &a, const ScanWorkspace &ws, int tile,
                                                    int first_wave, int tid, double *sd_red, int *sd_int) {
    using Elem = typename Tr::Elem;
    using State = typename Tr::State;
    constexpr int N = Elem::N;
    const int lane = tid & 31, warp = tid >> 5;
    Elem running = Tr::identity();
    bool have = false;
    int base = tile - 1;  // latest tile not yet covered
    // tiles per thread in this window.  A tile of the first wave has no published prefix near it (its
    // neighbours started together with it): take everything back to tile 0 in one window if it fits.
    // Later tiles find a prefix a few tiles back: one tile per thread keeps the loads minimal.
    int c = 1;
    if (tile < first_wave) {
        c = (tile + 1 + SCAN_THREADS - 1) / SCAN_THREADS;
        if (c > LB_MAX_PER_THREAD) c = LB_MAX_PER_THREAD;
    }
    while (true) {
        // ---- wait until every tile of the window has at least its aggregate out ----
        int f[LB_MAX_PER_THREAD];
        while (true) {
            bool zero = false;
#pragma unroll
            for (int u = 0; u < LB_MAX_PER_THREAD; ++u) {
                if (u < c) {
                    const int idx = base - (tid * c + u);
                    f[u] = idx >= 0 ? flag_state<Tr>(ws, idx) : 2;  // before tile 0: the model prior, a "prefix"
                    zero |= f[u] == 0;
                }
            }
            // a tile publishes its prefix only after all earlier tiles have published aggregates, so
            // zeros can only sit nearer than the nearest prefix: all of them are needed
            if (!__syncthreads_or(zero)) break;
        }
        if (ws.dbg && tid == 0 && !have) ws.dbg[tile * 8 + 4] = gtimer();
        // ---- nearest published prefix of the window (position = distance from base) ----
        int mypos = 1 << 30;
#pragma unroll
        for (int u = LB_MAX_PER_THREAD - 1; u >= 0; --u)
            if (u < c && f[u] == 2) mypos = tid * c + u;
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) mypos = min(mypos, __shfl_xor_sync(FULL, mypos, d));
        if (lane == 0) sd_int[warp] = mypos;
        __syncthreads();
        int first = sd_int[0];
#pragma unroll
        for (int w = 1; w < NWARPS; ++w) first = min(first, sd_int[w]);
        // ---- this thread's tiles, earliest first ----
        Elem e = Tr::identity();
#pragma unroll
        for (int u = LB_MAX_PER_THREAD - 1; u >= 0; --u) {
            if (u < c) {
                const int pos = tid * c + u;
                const int idx = base - pos;
                if (pos < first) {
                    e = Tr::combine(e, load_elem_cg<Elem>(ws.tile_agg + (int64_t)idx * AGG_PITCH));
                } else if (pos == first) {
                    const State s = idx >= 0 ? load_elem_cg<State>(ws.tile_pref + (int64_t)idx * PREF_PITCH)
                                             : Tr::initial(a);
                    e = Tr::from_state(s);  // everything before it is inside the prefix
                }
            }
        }
        if (ws.dbg && tid == 0 && !have) ws.dbg[tile * 8 + 5] = gtimer();
        // ---- ordered reduction: lane i holds later tiles than lane i + d, warp w later than w + 1 ----
        // only as many levels / warps as the window up to the nearest prefix needs
        const int last_thread = first < (1 << 30) ? first / c : SCAN_THREADS - 1;  // thread holding that prefix
        const int last_warp = last_thread >> 5;
        const int span = warp < last_warp ? 32 : (warp == last_warp ? (last_thread & 31) + 1 : 0);
        for (int d = 1; d < span; d <<= 1) {
            const Elem o = shfl_down_elem(e, d);
            if (lane + d < 32) e = Tr::combine(o, e);
        }
        if (lane == 0) store_elem(sd_red + warp * N, e);
        __syncthreads();
        Elem win = load_elem<Elem>(sd_red + last_warp * N);
        for (int w = last_warp - 1; w >= 0; --w) win = Tr::combine(win, load_elem<Elem>(sd_red + w * N));
        running = have ? Tr::combine(win, running) : win;
        have = true;
        if (first < (1 << 30)) break;
        base -= SCAN_THREADS * c;
        c = LB_MAX_PER_THREAD;
        __syncthreads();  // sd_red / sd_int are rewritten by the next window
    }
    return Tr::elem_state(running);
}

// =====================================================================================
// scan kernel
// =====================================================================================
// A CTA works through tiles it claims from a counter, software-pipelined over two tiles: it
// composes the run elements of the NEXT tile and publishes that tile's aggregate (stage 1) before
// it looks back and replays the tile it claimed before (stage 2).  By the time a tile's look-back
// starts, the aggregates of the tiles in front of it have had a whole stage 1 to arrive, so the
// wait that a single-pass scan otherwise spends between its two passes is filled with the next
// tile's first pass.  A claimed tile's stage 1 follows its claim without any wait in between, so
// every aggregate a look-back waits for is being produced by a running CTA: no deadlock whatever
// the number of resident CTAs.
template <class Tr>
struct ScanSmem {
    static constexpr int N = Tr::Elem::N;
    static constexpr int REC_TOTAL = 2 * Tr::G::BUF_BYTES;  // double-buffered records
    // doubles after the records
    static constexpr int OFF_WAGG = 0;                       // [NWARPS][N] warp aggregates
    static constexpr int OFF_TAGG = OFF_WAGG + NWARPS * N;   // [2][N] tile aggregates (pending tile, next tile)
    static constexpr int OFF_RED = OFF_TAGG + 2 * N;         // [NWARPS][2] partial sums
    static constexpr int OFF_LB = OFF_RED + NWARPS * 2;      // [NWARPS][N] look-back: per-warp window results
    static constexpr int OFF_EX = OFF_LB + NWARPS * N;       // [N][SCAN_THREADS] exclusive elements of the pending tile
    static constexpr int OFF_END = OFF_EX + N * SCAN_THREADS;
    static constexpr int BYTES = REC_TOTAL + OFF_END * 8 + 48;  // + s_tile[4], look-back ints[NWARPS]
};

template <class Tr, bool AGG_ONLY>
__global__ void __launch_bounds__(SCAN_THREADS, Tr::MIN_CTAS)
scan_kernel(const typename Tr::Args a, const ScanWorkspace ws, const int ntiles, const int nsub, const int first_wave) {
    using Elem = typename Tr::Elem;
    using State = typename Tr::State;
    using SM = ScanSmem<Tr>;
    extern __shared__ __align__(16) unsigned char smem[];
    double *sd = reinterpret_cast<double *>(smem + SM::REC_TOTAL);
    int *s_tile = reinterpret_cast<int *>(sd + SM::OFF_END);
    int *s_lb = s_tile + 4;
    double *sd_ex = sd + SM::OFF_EX + threadIdx.x;  // this thread's column

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int L = CHUNK * nsub;
    auto buf = [&](int s) -> unsigned char * { return smem + (s & 1) * Tr::G::BUF_BYTES; };

    int cur = -1;  // tile whose run elements are composed and whose replay is pending
    for (int it = 0;; ++it) {
        // ---- claim the next tile ----
        if (tid == 0) s_tile[0] = AGG_ONLY ? (int)blockIdx.x : atomicAdd(ws.counters, 1);
        __syncthreads();
        const int nxt = s_tile[0];
        const bool have_next = nxt < ntiles;

        // ================= stage 1 of tile nxt: run elements, warp scan, aggregate =================
        Elem ex_next = Tr::identity();  // composition of the runs of this tile in front of this thread's
        if (have_next) {
            const int tile = nxt;
            if (ws.dbg && tid == 0) ws.dbg[tile * 8] = gtimer();
            const int64_t p0 = (int64_t)tile * TILE_BINS * nsub;
            const int64_t run0 = p0 + (int64_t)tid * L;
            Elem mine = Tr::identity();
            const bool prebuilt = !AGG_ONLY && Tr::prebuilt(a);
            if (prebuilt) {
                // the forward scan composed this run's element during its replay
                mine = Tr::load_prebuilt(a, (int64_t)ntiles * SCAN_THREADS - 1 - ((int64_t)tile * SCAN_THREADS + tid));
            } else {
                __syncwarp();  // the warp's stores of the preceding replay have read their records
                Tr::template issue<false>(a, buf(0), p0, L, 0, tid);
                cp_async_commit();
            }
            // the copies of sub-step s+1 are in flight while sub-step s is computed
            for (int s = 0; s < (prebuilt ? 0 : nsub); ++s) {
                cp_async_wait_all();
                __syncwarp();  // sub-step s has landed for the whole warp; its buffer (s+1)&1 is no longer read
                if (s + 1 < nsub) {
                    Tr::template issue<false>(a, buf(s + 1), p0, L, s + 1, tid);
                    cp_async_commit();
                }
                const int64_t q0 = run0 + s * CHUNK;
                int lo, hi;
                Tr::bounds(a, q0, lo, hi);
                if (lo == 0 && hi == CHUNK)
                    Tr::template pass1<true>(a, Tr::G::cells(buf(s), tid), lo, hi, q0, mine);
                else if (hi > lo)
                    Tr::template pass1<false>(a, Tr::G::cells(buf(s), tid), lo, hi, q0, mine);
            }
            // inclusive Kogge-Stone scan across the warp
            Elem inc = mine;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const Elem o = shfl_up_elem(inc, d);
                if (lane >= d) inc = Tr::combine(o, inc);
            }
            if (lane == 31) store_elem(sd + SM::OFF_WAGG + warp * SM::N, inc);
            __syncthreads();
            if (ws.dbg && tid == 0) ws.dbg[tile * 8 + 1] = gtimer();
            // every warp forms its own exclusive prefix over the warps before it (lane-redundant); the
            // last warp also forms and publishes the tile aggregate
            Elem wex = warp > 0 ? load_elem<Elem>(sd + SM::OFF_WAGG) : Tr::identity();
            for (int w = 1; w < warp; ++w) wex = Tr::combine(wex, load_elem<Elem>(sd + SM::OFF_WAGG + w * SM::N));
            if (warp == NWARPS - 1) {
                const Elem tagg = Tr::combine(wex, load_elem<Elem>(sd + SM::OFF_WAGG + warp * SM::N));
                if (lane == 0) {
                    store_elem(sd + SM::OFF_TAGG + (it & 1) * SM::N, tagg);
                    store_elem(ws.tile_agg + (int64_t)tile * AGG_PITCH, tagg);
                    if (!AGG_ONLY) {
                        __threadfence();
                        st_release(ws.flags + tile, ws.epoch4 + 1);
                        if (ws.dbg) ws.dbg[tile * 8 + 7] = gtimer();
                    }
                }
            }
            if (!AGG_ONLY) {
                const Elem lex = shfl_up_elem(inc, 1);
                ex_next = lane > 0 ? (warp > 0 ? Tr::combine(wex, lex) : lex) : wex;
            }
        }
        if (AGG_ONLY) return;

        // hand over: the pending tile's exclusive element comes out of this thread's column of
        // shared memory, the next tile's goes in (thread-private, no barrier)
        Elem ex_cur = Tr::identity();
        if (cur >= 0) {
            double *d = reinterpret_cast<double *>(&ex_cur);
#pragma unroll
            for (int i = 0; i < SM::N; ++i) d[i] = sd_ex[i * SCAN_THREADS];
        }
        if (have_next) {
            const double *d = reinterpret_cast<const double *>(&ex_next);
#pragma unroll
            for (int i = 0; i < SM::N; ++i) sd_ex[i * SCAN_THREADS] = d[i];
        }

        // ================= stage 2 of tile cur: look-back, prefix, replay =================
        if (cur >= 0) {
            const int tile = cur;
            const int64_t p0 = (int64_t)tile * TILE_BINS * nsub;
            const int64_t run0 = p0 + (int64_t)tid * L;
            // the replay's first sub-step is fetched underneath the look-back
            __syncwarp();
            Tr::template issue<true>(a, buf(0), p0, L, 0, tid);
            cp_async_commit();
            if (ws.dbg && tid == 0) ws.dbg[tile * 8 + 6] = gtimer();
            const State tpref = tile == 0 ? Tr::initial(a)
                                          : cooperative_lookback<Tr>(a, ws, tile, first_wave, tid, sd + SM::OFF_LB, s_lb);
            cp_async_wait_all();
            __syncthreads();  // the replay's first records have landed
            if (warp == NWARPS - 1 && lane == 0) {
                // inclusive prefix of the tile for the tiles behind it (off this tile's critical path)
                const State incl = Tr::apply(load_elem<Elem>(sd + SM::OFF_TAGG + ((it + 1) & 1) * SM::N), tpref);
                store_elem(ws.tile_pref + (int64_t)tile * PREF_PITCH, incl);
                __threadfence();
                st_release(ws.flags + tile, ws.epoch4 + 2);
            }
            if (ws.dbg && tid == 0) ws.dbg[tile * 8 + 2] = gtimer();
            const State start = tid == 0 ? tpref : Tr::apply(ex_cur, tpref);

            // replay the reference's recursion over the run from its exact start state
            typename Tr::Carry carry = Tr::begin2(a, start);
            double acc0 = 0.0, acc1 = 0.0;
            for (int s = 0; s < nsub; ++s) {
                if (s) {
                    cp_async_wait_all();
                    __syncwarp();  // sub-step s has landed; the warp's stores of sub-step s-1 have read their buffer
                }
                if (s + 1 < nsub) {
                    Tr::template issue<true>(a, buf(s + 1), p0, L, s + 1, tid);
                    cp_async_commit();
                }
                const int64_t q0 = run0 + s * CHUNK;
                int lo, hi;
                Tr::bounds(a, q0, lo, hi);
                if (lo == 0 && hi == CHUNK)
                    Tr::template pass2<true>(a, Tr::G::cells(buf(s), tid), lo, hi, q0, carry, acc0, acc1);
                else if (hi > lo)
                    Tr::template pass2<false>(a, Tr::G::cells(buf(s), tid), lo, hi, q0, carry, acc0, acc1);
                __syncwarp();
                Tr::stage_out(a, buf(s), p0, L, s, tid);
            }
            acc1 += Tr::finish2(a, carry);
            Tr::finish_run(a, (int64_t)tile * SCAN_THREADS + tid, run0, L, carry);
            Tr::epilogue(a, tile, tid, L, carry, acc0, acc1);
            if (ws.dbg && tid == 0) ws.dbg[tile * 8 + 3] = gtimer();

            if (Tr::HAS_SUMS) {
#pragma unroll
                for (int d = 16; d > 0; d >>= 1) {
                    acc0 += __shfl_xor_sync(FULL, acc0, d);
                    acc1 += __shfl_xor_sync(FULL, acc1, d);
                }
                if (lane == 0) {
                    sd[SM::OFF_RED + warp * 2] = acc0;
                    sd[SM::OFF_RED + warp * 2 + 1] = acc1;
                }
                __syncthreads();
                if (tid == 0) {
                    double t0 = 0.0, t1 = 0.0;
#pragma unroll
                    for (int w = 0; w < NWARPS; ++w) {
                        t0 += sd[SM::OFF_RED + w * 2];
                        t1 += sd[SM::OFF_RED + w * 2 + 1];
                    }
                    ws.partials[(int64_t)tile * 2] = t0;
                    ws.partials[(int64_t)tile * 2 + 1] = t1;
                }
            }
        }
        if (!have_next) break;
        cur = nxt;
        __syncthreads();  // s_tile[0] and the scratch of this iteration are rewritten by the next
    }

    // ---- this CTA is done: the last one to leave resets the counters and adds up the partial sums ----
    if (tid == 0) {
        __threadfence();
        const int done = atomicAdd(ws.counters + 1, 1);
        const int last = (done == (int)gridDim.x - 1) ? 1 : 0;
        if (last) {
            // every CTA has made its last (failing) claim: leave the counters ready for the next launch
            ws.counters[0] = 0;
            ws.counters[1] = 0;
            __threadfence();
        }
        s_tile[1] = last;
    }
    if (Tr::HAS_SUMS) {
        __syncthreads();
        if (s_tile[1] && warp == 0) {
            // per-tile partial sums in tile order
            double t0 = 0.0, t1 = 0.0;
            for (int t = lane; t < ntiles; t += 32) {
                t0 += __ldcg(ws.partials + (int64_t)t * 2);
                t1 += __ldcg(ws.partials + (int64_t)t * 2 + 1);
            }
#pragma unroll
            for (int d = 16; d > 0; d >>= 1) {
                t0 += __shfl_xor_sync(FULL, t0, d);
                t1 += __shfl_xor_sync(FULL, t1, d);
            }
            double *sums = Tr::sums(a);
            if (lane == 0 && sums != nullptr) {
                sums[0] = t0;
                sums[1] = t1;
            }
        }
    }
}

// Ordered reduction of the per-tile aggregates of a shard into one element (aggregate-only mode).
template <class Tr>
__global__ void __launch_bounds__(256) reduce_tile_aggs_kernel(const double *tile_agg, int ntiles, double *out) {
    using Elem = typename Tr::Elem;
    __shared__ double sh[256 * Tr::Elem::N];
    const int tid = threadIdx.x;
    const int per = (ntiles + 255) / 256;
    Elem e = Tr::identity();
    for (int i = 0; i < per; ++i) {
        const int t = tid * per + i;
        if (t < ntiles) e = Tr::combine(e, load_elem<Elem>(tile_agg + (int64_t)t * AGG_PITCH));
    }
    store_elem(sh + tid * Tr::Elem::N, e);
    __syncthreads();
    for (int s = 1; s < 256; s <<= 1) {
        if ((tid % (2 * s)) == 0 && tid + s < 256) {
            const Elem a = load_elem<Elem>(sh + tid * Tr::Elem::N);
            const Elem b = load_elem<Elem>(sh + (tid + s) * Tr::Elem::N);
            store_elem(sh + tid * Tr::Elem::N, Tr::combine(a, b));
        }
        __syncthreads();
    }
    if (tid < Tr::Elem::N) out[tid] = sh[tid];
}

template <class Tr>
__global__ void forward_shard_prefix_kernel(const double *aggs, int rank, double state_init, double cov_init,
                                            double *init_state) {
    using Elem = typename Tr::Elem;
    using State = typename Tr::State;
    FwdArgs a{};
    a.init_state = nullptr;
    a.state_init = state_init;
    a.cov_init = cov_init;
    State s = Tr::initial(a);
    for (int r = 0; r < rank; ++r) s = Tr::apply(load_elem<Elem>(aggs + (int64_t)r * AGG_PITCH), s);
    store_elem(init_state, s);
}

template <class Tr>
__global__ void backward_shard_prefix_kernel(const double *aggs, int rank, int n_shards, double *tail_state) {
    using Elem = typename Tr::Elem;
    using State = typename Tr::State;
    BwdArgs a{};
    a.tail_state = nullptr;
    State s = Tr::initial(a);
    for (int r = n_shards - 1; r > rank; --r) s = Tr::apply(load_elem<Elem>(aggs + (int64_t)r * AGG_PITCH), s);
    store_elem(tail_state, s);
}

// =====================================================================================
// residuals: resid[k][j] = data[j][k] - level[k]      (float32 [n][m], transposed w.r.t. data)
// =====================================================================================
// The reference evaluates (float)((double)z - (double)x_s) (pyx:6846-6848).  The double
// difference of two floats is exact, so rounding it to float is the correctly rounded float
// difference: a single FSUB gives the same bits.
//
// A CTA transposes `bins` consecutive bins x ALL m tracks through shared memory, so that what it
// writes -- rows k0 .. k0 + bins of the [n x m] output -- is ONE contiguous stretch of memory
// whatever m is (m = 50: a 32-track tile would end every output row inside a sector).  `bins` is a
// multiple of 32 chosen by the launcher from the shared-memory budget.  Loads: 4-byte cp.async, a warp
// copies 32 consecutive bins of one track (one 128-byte line) per instruction, the whole tile in
// flight at once; the tile row pitch is odd, so both the writes (lanes along the bins) and the reads
// (lanes along the tracks) are free of bank conflicts.  Stores walk the output in memory order, one coalesced 128-byte line per warp
// instruction; (bin, track) advance incrementally, no per-element division.
constexpr int RES_THREADS = 256;
constexpr int RES_SMEM_BUDGET = 96 * 1024;  // two CTAs per SM

__device__ __forceinline__ void cp_async4(unsigned dst_shared, const float *src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst_shared), "l"(src) : "memory");
}

// WHOLE: the tile spans all m tracks (its output rows are one contiguous stretch).  VEC: 16-byte stores
// (the launcher checks the alignment they need).  Index arithmetic is 32-bit and incremental throughout: the
// first version of this kernel spent 37 instructions per element and was issue-bound (profiles/README.md).
template <bool WHOLE, bool VEC>
__global__ void __launch_bounds__(RES_THREADS)
residual_kernel(const float *__restrict__ data, int64_t m_all, int64_t n, int64_t ld, const float *__restrict__ xs,
                int dim, float *__restrict__ resid, int bins, int m_chunk) {
    extern __shared__ float res_tile[];  // [m][bins + 1] then level[bins]
    // tracks [j0, j0 + m) of the m_all; the chunks of one bin range are neighbours in launch order, so the
    // pieces of an output row are written close together in time
    const int chunks = (int)((m_all + m_chunk - 1) / m_chunk);
    const int64_t bid = blockIdx.x;
    const int64_t j0 = (bid % chunks) * m_chunk;
    const int m = (int)min((int64_t)m_chunk, m_all - j0);
    data += j0 * ld;
    const int pitch = bins + 1;
    float *level = res_tile + (size_t)m_chunk * pitch;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t k0 = (bid / chunks) * bins;
    const int nb = (int)min((int64_t)bins, n - k0);
    // ---- load: the whole tile by 4-byte cp.async (no registers, every copy in flight at once); a warp
    //      copies 32 consecutive bins of one track per instruction: one 128-byte line, conflict-free ----
    const int nfull = nb >> 5, tail = nb & 31;
    for (int j = warp; j < m; j += RES_THREADS / 32) {
        const float *g = data + (int64_t)j * ld + k0 + lane;
        unsigned sa = (unsigned)__cvta_generic_to_shared(res_tile + j * pitch + lane);
        int sgm = 0;
        for (; sgm + 4 <= nfull; sgm += 4) {
            cp_async4(sa, g);
            cp_async4(sa + 128, g + 32);
            cp_async4(sa + 256, g + 64);
            cp_async4(sa + 384, g + 96);
            sa += 512;
            g += 128;
        }
        for (; sgm < nfull; ++sgm) {
            cp_async4(sa, g);
            sa += 128;
            g += 32;
        }
        if (lane < tail) cp_async4(sa, g);
    }
    cp_async_commit();
    for (int kk = tid; kk < nb; kk += RES_THREADS) level[kk] = __ldg(xs + (k0 + kk) * dim);
    cp_async_wait_all();
    __syncthreads();
    // ---- store: the [nb][m] block in memory order ----
    const int total = nb * m;
    float *out = resid + k0 * m_all + j0;
    if (VEC) {
        // four consecutive outputs per thread and store: tracks jj .. jj + 3 of one bin (or the wrap into the next)
        const int quads = total >> 2;
        const int dkk = (4 * RES_THREADS) / m, djj = 4 * RES_THREADS - dkk * m;
        int kk = (4 * tid) / m, jj = 4 * tid - kk * m;
        const int m_all32 = (int)m_all;
        for (int q = tid; q < quads; q += RES_THREADS) {
            float4 v;
            if (jj + 3 < m) {
                const float *t = res_tile + jj * pitch + kk;
                const float lv = level[kk];
                v = make_float4(t[0] - lv, t[pitch] - lv, t[2 * pitch] - lv, t[3 * pitch] - lv);
            } else {  // the quad runs over the end of a bin's tracks (WHOLE only: chunks are multiples of 4)
                float w[4];
                int k2 = kk, j2 = jj;
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    w[u] = res_tile[j2 * pitch + k2] - level[k2];
                    if (++j2 == m) {
                        j2 = 0;
                        ++k2;
                    }
                }
                v = make_float4(w[0], w[1], w[2], w[3]);
            }
            float *dst = WHOLE ? out + 4 * q : out + (int64_t)kk * m_all32 + jj;
            __stcs(reinterpret_cast<float4 *>(dst), v);
            kk += dkk;
            jj += djj;
            if (jj >= m) {
                jj -= m;
                kk += 1;
            }
        }
        if (WHOLE) {  // at most three elements left over (last, ragged tile)
            const int idx = 4 * quads + tid;
            if (idx < total) {
                const int k2 = idx / m, j2 = idx - k2 * m;
                __stcs(out + idx, res_tile[j2 * pitch + k2] - level[k2]);
            }
        }
    } else {
        const int dkk = RES_THREADS / m, djj = RES_THREADS - dkk * m;
        int kk = tid / m, jj = tid - kk * m;
        for (int idx = tid; idx < total; idx += RES_THREADS) {
            const float v = res_tile[jj * pitch + kk] - level[kk];
            if (WHOLE)
                __stcs(out + idx, v);
            else
                __stcs(out + (int64_t)kk * m_all + jj, v);
            kk += dkk;
            jj += djj;
            if (jj >= m) {
                jj -= m;
                kk += 1;
            }
        }
    }
}

// =====================================================================================
// Student-t precision multipliers
// =====================================================================================
__global__ void lambda_kernel(const double2 *__restrict__ SA, const double2 *__restrict__ SB, int64_t n, double m,
                              const float *__restrict__ xs,
                              const float *__restrict__ Ps, int dim, double nu, double lo, double hi,
                              float *__restrict__ lam) {
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const double lvl = (double)xs[k * dim];
    const double p00 = (double)Ps[k * dim * dim];
    const double2 s01 = SA[k];
    lam[k] = (float)lambda_update(s01.x, s01.y, SB[k].x, lvl, p00, m, nu, lo, hi);
}

__global__ void kappa2_kernel(Model2 M, double qi00, double qi01, double qi10, double qi11, int64_t n,
                              const float *__restrict__ xs, const float *__restrict__ Ps,
                              const float *__restrict__ lag, const float *__restrict__ qs, double nu, double lo,
                              double hi, float *__restrict__ kap) {
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k == 0) kap[0] = 1.0f;
    if (k >= n - 1) return;
    const float2 x = reinterpret_cast<const float2 *>(xs)[k], y = reinterpret_cast<const float2 *>(xs)[k + 1];
    const float4 P = reinterpret_cast<const float4 *>(Ps)[k], Py = reinterpret_cast<const float4 *>(Ps)[k + 1];
    const float4 C = reinterpret_cast<const float4 *>(lag)[k];
    const double q = qs ? (double)qs[k + 1] : 1.0;
    kap[k + 1] = (float)kappa2_update(M, qi00, qi01, qi10, qi11, x.x, x.y, P.x, P.y, P.z, P.w, y.x, y.y, Py.x, Py.y,
                                      Py.z, Py.w, C.x, C.y, C.z, C.w, q, qs != nullptr, nu, lo, hi);
}

__global__ void kappa1_kernel(double q0inv, int64_t n, const float *__restrict__ xs, const float *__restrict__ Ps,
                              const float *__restrict__ lag, const float *__restrict__ qs, double nu, double lo,
                              double hi, float *__restrict__ kap) {
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k == 0) kap[0] = 1.0f;
    if (k >= n - 1) return;
    const double q = qs ? (double)qs[k + 1] : 1.0;
    kap[k + 1] = (float)kappa1_update(q0inv, (double)xs[k], (double)Ps[k], (double)xs[k + 1], (double)Ps[k + 1],
                                      (double)lag[k], q, qs != nullptr, nu, lo, hi);
}

__global__ void fill_kernel(float *v, int64_t n, float value) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) v[i] = value;
}

static int g_slots[4] = {0, 0, 0, 0};  // Fwd2, Fwd1, Bwd2, Bwd1 tiles in flight

template <class Tr, bool AGG_ONLY>
cudaError_t launch_scan(const typename Tr::Args &a, const ScanWorkspace &ws, int64_t positions, int nsub,
                        int slots, cudaStream_t st, int *launches) {
    const int ntiles = (int)scan_num_tiles(positions, nsub);
    if (ntiles <= 0) return cudaSuccess;
    // aggregate-only: one CTA per tile; full scan: as many CTAs as are resident, each working through
    // the tiles it claims
    const int grid = AGG_ONLY ? ntiles : (slots > 0 && slots < ntiles ? slots : ntiles);
    ScanWorkspace w = ws;
    if (w.dbg && ntiles > w.dbg_tiles) w.dbg = nullptr;  // the stamps of this launch would not fit the armed buffer
    scan_kernel<Tr, AGG_ONLY><<<grid, SCAN_THREADS, ScanSmem<Tr>::BYTES, st>>>(a, w, ntiles, nsub, ntiles);
    if (launches) *launches += 1;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    if (AGG_ONLY) {
        reduce_tile_aggs_kernel<Tr><<<1, 256, 0, st>>>(ws.tile_agg, ntiles, a.agg_out);
        if (launches) *launches += 1;
        e = cudaGetLastError();
    }
    return e;
}

template <class Tr>
cudaError_t set_smem_attr() {
    cudaError_t e = cudaFuncSetAttribute(scan_kernel<Tr, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         ScanSmem<Tr>::BYTES);
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(scan_kernel<Tr, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                ScanSmem<Tr>::BYTES);
}

// resident CTAs of the scan kernels on the current device (tiles in flight)
template <class Tr>
int scan_slots() {
    int per_sm = 0, dev = 0, sms = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 0;
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, scan_kernel<Tr, false>, SCAN_THREADS,
                                                      ScanSmem<Tr>::BYTES) != cudaSuccess)
        return 0;
    return per_sm * sms;
}

bool canonical_F(const Model2 &M) { return M.F00 == 1.0 && M.F10 == 0.0 && M.F11 == 1.0; }

}  // namespace

// =====================================================================================
// host-side launchers
// =====================================================================================
int64_t scan_num_tiles(int64_t positions, int nsub) {
    const int64_t tb = (int64_t)TILE_BINS * nsub;
    return (positions + tb - 1) / tb;
}

static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// workspace for the finest tiling (nsub = 1) of n + CHUNK positions (the reverse scans pad n)
static size_t ws_tiles(int64_t n) { return (size_t)scan_num_tiles(n + CHUNK, 1) + 1; }

size_t scan_workspace_bytes(int64_t n) {
    const size_t t = ws_tiles(n);
    return align_up(t * AGG_PITCH * 8, 256) + align_up(t * PREF_PITCH * 8, 256) + align_up(t * 2 * 8, 256) +
           align_up((t + 4) * 4, 256);
}

ScanWorkspace scan_workspace_carve(void *base, int64_t n) {
    const size_t t = ws_tiles(n);
    unsigned char *p = static_cast<unsigned char *>(base);
    ScanWorkspace ws{};
    ws.tile_agg = reinterpret_cast<double *>(p);
    p += align_up(t * AGG_PITCH * 8, 256);
    ws.tile_pref = reinterpret_cast<double *>(p);
    p += align_up(t * PREF_PITCH * 8, 256);
    ws.partials = reinterpret_cast<double *>(p);
    p += align_up(t * 2 * 8, 256);
    ws.flags = reinterpret_cast<int32_t *>(p);
    ws.counters = ws.flags + t;  // contiguous with flags: one memset clears both
    return ws;
}


static int g_nsub_override = 0;

cudaError_t configure_kernels() {
    cudaError_t e;
    if ((e = set_smem_attr<Fwd2<true>>()) != cudaSuccess) return e;
    if ((e = set_smem_attr<Fwd2<false>>()) != cudaSuccess) return e;
    if ((e = set_smem_attr<Fwd1>()) != cudaSuccess) return e;
    if ((e = set_smem_attr<Bwd2<true>>()) != cudaSuccess) return e;
    if ((e = set_smem_attr<Bwd2<false>>()) != cudaSuccess) return e;
    if ((e = set_smem_attr<Bwd1>()) != cudaSuccess) return e;
    g_slots[0] = scan_slots<Fwd2<true>>();
    g_slots[1] = scan_slots<Fwd1>();
    g_slots[2] = scan_slots<Bwd2<true>>();
    g_slots[3] = scan_slots<Bwd1>();
    const char *ov = getenv("CB200_SCAN_NSUB");
    if (ov) g_nsub_override = atoi(ov);
    return cudaSuccess;
}

void scan_set_nsub_override(int nsub) { g_nsub_override = nsub < 0 ? 0 : nsub; }

// Sub-steps per run.  Longer runs amortise the warp scan and the look-back, but the tiles must still
// cover the machine, and a launch that spills into a second wave of tiles pays for it: measured on
// B200 (chr19 @ 25 bp) both scans run fastest with the longest runs that keep one wave about 94 %
// full (417 tiles: 444 slots for the forward scan at 3 CTAs/SM, 592 for the others at 4).
// which: 0 Fwd2, 1 Fwd1, 2 Bwd2, 3 Bwd1.
int scan_pick_nsub(int64_t positions, int which) {
    if (g_nsub_override > 0) return g_nsub_override > MAX_NSUB ? MAX_NSUB : g_nsub_override;
    const int slots = g_slots[which] > 0 ? g_slots[which] : 444;
    const int64_t want_tiles = (int64_t)((slots > 444 ? 0.705 : 0.94) * slots);
    int64_t ns = (positions + TILE_BINS * want_tiles - 1) / (TILE_BINS * want_tiles);  // tiles <= want_tiles
    if (ns < 1) ns = 1;
    if (ns > MAX_NSUB) ns = MAX_NSUB;
    return (int)ns;
}

cudaError_t launch_fill(float *v, int64_t n, float value, cudaStream_t st) {
    if (n <= 0) return cudaSuccess;
    fill_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(v, n, value);
    return cudaGetLastError();
}

cudaError_t launch_fold(const float *data, const float *munc, int64_t m, int64_t n, int64_t ld, double pad,
                        double2 *SA, double2 *SB, cudaStream_t st, int rm_logL) {
    if (n <= 0) return cudaSuccess;
    const bool vec = (ld % 4 == 0) && ((reinterpret_cast<uintptr_t>(data) & 15) == 0) &&
                     ((reinterpret_cast<uintptr_t>(munc) & 15) == 0);
    if (vec) {
        const int64_t threads = (n + 3) / 4;
        const unsigned grid = (unsigned)((threads + FOLD_THREADS - 1) / FOLD_THREADS);
        fold_kernel<4><<<grid, FOLD_THREADS, 0, st>>>(data, munc, m, n, ld, pad, SA, SB, rm_logL);
    } else {
        const unsigned grid = (unsigned)((n + FOLD_THREADS - 1) / FOLD_THREADS);
        fold_kernel<1><<<grid, FOLD_THREADS, 0, st>>>(data, munc, m, n, ld, pad, SA, SB, rm_logL);
    }
    return cudaGetLastError();
}

cudaError_t launch_forward(int dim, const FwdArgs &a_in, const ScanWorkspace &ws, bool aggregate_only,
                           cudaStream_t st, int *launches) {
    if (dim == 2) {
        const int ns = a_in.nsub > 0 ? a_in.nsub : scan_pick_nsub(a_in.n, 0);
        FwdArgs a = a_in;
        a.head_from = (a.init_state == nullptr && !aggregate_only) ? CHUNK * ns : HEAD_BINS;
        if (canonical_F(a.M))
            return aggregate_only ? launch_scan<Fwd2<true>, true>(a, ws, a.n, ns, g_slots[0], st, launches)
                                  : launch_scan<Fwd2<true>, false>(a, ws, a.n, ns, g_slots[0], st, launches);
        return aggregate_only ? launch_scan<Fwd2<false>, true>(a, ws, a.n, ns, g_slots[0], st, launches)
                              : launch_scan<Fwd2<false>, false>(a, ws, a.n, ns, g_slots[0], st, launches);
    }
    FwdArgs a = a_in;
    a.head_from = HEAD_BINS;
    const int ns = scan_pick_nsub(a.n, 1);
    return aggregate_only ? launch_scan<Fwd1, true>(a, ws, a.n, ns, g_slots[1], st, launches)
                          : launch_scan<Fwd1, false>(a, ws, a.n, ns, g_slots[1], st, launches);
}

cudaError_t launch_backward(int dim, const BwdArgs &a, const ScanWorkspace &ws, bool aggregate_only,
                            cudaStream_t st, int *launches) {
    const int64_t positions = (dim == 2 && a.npad_fixed > 0) ? a.npad_fixed : (a.n + CHUNK - 1) / CHUNK * CHUNK;
    if (dim == 2) {
        const int ns = a.nsub > 0 ? a.nsub : scan_pick_nsub(positions, 2);
        if (canonical_F(a.M))
            return aggregate_only ? launch_scan<Bwd2<true>, true>(a, ws, positions, ns, g_slots[2], st, launches)
                                  : launch_scan<Bwd2<true>, false>(a, ws, positions, ns, g_slots[2], st, launches);
        return aggregate_only ? launch_scan<Bwd2<false>, true>(a, ws, positions, ns, g_slots[2], st, launches)
                              : launch_scan<Bwd2<false>, false>(a, ws, positions, ns, g_slots[2], st, launches);
    }
    const int ns = scan_pick_nsub(positions, 3);
    return aggregate_only ? launch_scan<Bwd1, true>(a, ws, positions, ns, g_slots[3], st, launches)
                          : launch_scan<Bwd1, false>(a, ws, positions, ns, g_slots[3], st, launches);
}

cudaError_t launch_residuals(const float *data, int64_t m, int64_t n, int64_t ld, const float *xs, int dim,
                             float *resid, cudaStream_t st) {
    if (n <= 0 || m <= 0) return cudaSuccess;
    // bins per CTA: as many 32-bin groups as the shared-memory budget holds for m tracks (at least
    // one; m beyond ~1700 tracks would need more than a CTA's shared memory for a single group)
    const int64_t max_bytes = 200 * 1024;
    // Few tracks: a tile spans all of them, so the output rows it writes are one contiguous stretch whatever m
    // is.  Many tracks: chunks of 128 (output pieces of 512 aligned bytes, input rows visited for 640 bytes):
    // a tile of all 1000 tracks would visit every input row for 128 bytes only.
    const int64_t m_chunk = m <= 96 ? m : 128;
    const int64_t chunks = (m + m_chunk - 1) / m_chunk;
    int64_t bins = (RES_SMEM_BUDGET / 4 - 1) / (m_chunk + 1) / 32 * 32;
    if (bins < 32) bins = 32;
    if (bins > 512) bins = 512;
    // short tracks: enough CTAs to cover the machine
    while (bins > 32 && (n + bins - 1) / bins < 296) bins -= 32;
    const int64_t smem = (m_chunk * (bins + 1) + bins) * 4;
    if (smem > max_bytes) return cudaErrorInvalidValue;
    const bool whole = chunks == 1;
    // 16-byte stores: the output block of a tile starts on a 16-byte boundary (k0 is a multiple of 32) when the
    // array does; a chunked tile also needs every output row piece aligned (m and the chunk multiples of 4)
    const bool vec = (reinterpret_cast<uintptr_t>(resid) & 15) == 0 && (whole || (m % 4 == 0 && m_chunk % 4 == 0)) &&
                     m * 4 * (int64_t)RES_THREADS < (int64_t)1 << 30;
    auto kern = whole ? (vec ? residual_kernel<true, true> : residual_kernel<true, false>)
                      : (vec ? residual_kernel<false, true> : residual_kernel<false, false>);
    static bool configured[4] = {false, false, false, false};
    const int which = (whole ? 2 : 0) + (vec ? 1 : 0);
    if (!configured[which]) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)max_bytes);
        if (e != cudaSuccess) return e;
        configured[which] = true;
    }
    const int64_t blocks = (n + bins - 1) / bins;
    if (blocks * chunks > 0x7fffffffLL) return cudaErrorInvalidValue;
    kern<<<(unsigned)(blocks * chunks), RES_THREADS, (size_t)smem, st>>>(data, m, n, ld, xs, dim, resid, (int)bins, (int)m_chunk);
    return cudaGetLastError();
}

cudaError_t launch_update_lambda(const double2 *SA, const double2 *SB, int64_t n, double m,
                                 const float *xs, const float *Ps, int dim, double nu, double lo, double hi,
                                 float *lam, cudaStream_t st) {
    if (n <= 0) return cudaSuccess;
    lambda_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(SA, SB, n, m, xs, Ps, dim, nu, lo, hi, lam);
    return cudaGetLastError();
}

cudaError_t launch_update_kappa(int dim, const Model2 &M, int64_t n, const float *xs, const float *Ps,
                                const float *lag, const float *qs, double nu, double lo, double hi, float *kap,
                                cudaStream_t st) {
    if (n <= 0) return cudaSuccess;
    const unsigned grid = (unsigned)((n + 255) / 256);
    if (dim == 2) {
        const double det = M.q00 * M.q11 - M.q01 * M.q10;
        kappa2_kernel<<<grid, 256, 0, st>>>(M, M.q11 / det, -M.q01 / det, -M.q10 / det, M.q00 / det, n, xs, Ps, lag,
                                            qs, nu, lo, hi, kap);
    } else {
        kappa1_kernel<<<grid, 256, 0, st>>>(1.0 / M.q00, n, xs, Ps, lag, qs, nu, lo, hi, kap);
    }
    return cudaGetLastError();
}

cudaError_t launch_forward_shard_prefix(int dim, const double *aggs, int rank, double state_init,
                                        double cov_init, double *init_state, cudaStream_t st) {
    if (dim == 2)
        forward_shard_prefix_kernel<Fwd2<false>><<<1, 1, 0, st>>>(aggs, rank, state_init, cov_init, init_state);
    else
        forward_shard_prefix_kernel<Fwd1><<<1, 1, 0, st>>>(aggs, rank, state_init, cov_init, init_state);
    return cudaGetLastError();
}

cudaError_t launch_backward_shard_prefix(int dim, const double *aggs, int rank, int n_shards,
                                         double *tail_state, cudaStream_t st) {
    if (dim == 2)
        backward_shard_prefix_kernel<Bwd2<false>><<<1, 1, 0, st>>>(aggs, rank, n_shards, tail_state);
    else
        backward_shard_prefix_kernel<Bwd1><<<1, 1, 0, st>>>(aggs, rank, n_shards, tail_state);
    return cudaGetLastError();
}

}  // namespace cb200
