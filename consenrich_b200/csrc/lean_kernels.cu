// consenrich_b200/csrc/lean_kernels.cu -- sm_100a kernels of the ECM's inner sweeps on run-major
// private tracks (layout and launch sequence: lean_kernels.cuh).  The arithmetic of every bin is the
// same ssm_math.cuh code the look-back kernels (ssm_kernels.cu) run: filt2_step for the run
// elements, kf2_step / rts2_step / kappa2_update for the replay of the reference's recursions
// (cconsenrich.pyx:388-529, 6758-6848, 8252-8298).  What differs is the data movement: one run per
// lane, element i of a warp's 32 runs in one 512-byte row, registers <-> HBM directly.
#include <cuda_runtime.h>

#include "lean_kernels.cuh"

namespace cb200 {

namespace {

constexpr unsigned FULL = 0xffffffffu;
#ifndef LEAN_COMPOSE_CTAS
#define LEAN_COMPOSE_CTAS 2
#endif
#ifndef LEAN_FWD_CTAS
#define LEAN_FWD_CTAS 2
#endif
#ifndef LEAN_BWD_CTAS
#define LEAN_BWD_CTAS 2
#endif

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
// element j of E at p[j * pitch]
template <class E>
__device__ __forceinline__ void store_strided(double *p, int pitch, const E &e) {
    const double *s = reinterpret_cast<const double *>(&e);
#pragma unroll
    for (int i = 0; i < E::N; ++i) p[i * pitch] = s[i];
}
template <class E>
__device__ __forceinline__ E load_strided(const double *p, int pitch) {
    E r;
    double *t = reinterpret_cast<double *>(&r);
#pragma unroll
    for (int i = 0; i < E::N; ++i) t[i] = p[i * pitch];
    return r;
}

// qScale_k / clamp(kappa_k)   (cconsenrich.pyx:394-401, 408-415)
__device__ __forceinline__ double lean_qk(bool has_qs, float kap, float qs, double kmin, double kmax) {
    const double q = has_qs ? (double)qs : 1.0;
    return cb_div(q, clampd((double)kap, kmin, kmax));
}

struct Seg {  // where a lane's run sits
    int lane, w, L, valid;
    int64_t base, k0;
};
__device__ __forceinline__ Seg seg_of(const LeanGeom &g) {
    Seg s;
    s.lane = threadIdx.x & 31;
    s.w = blockIdx.x * LEAN_WARPS + (threadIdx.x >> 5);
    s.L = 1 << g.logL;
    s.base = (int64_t)s.w * ((int64_t)32 << g.logL);
    s.k0 = s.base + (int64_t)s.lane * s.L;
    const int64_t rem = g.n - s.k0;
    s.valid = s.w < g.W ? (int)(rem <= 0 ? 0 : (rem >= s.L ? s.L : rem)) : 0;
    return s;
}

// =====================================================================================
// forward 1/3: run elements
// =====================================================================================
struct ComposeIn {
    double2 s[4];
    float kp[4], q[4];
};

__global__ void __launch_bounds__(LEAN_THREADS, LEAN_COMPOSE_CTAS) lean_fwd_compose_kernel(const LeanFwdArgs a) {
    __shared__ double sh[LEAN_WARPS * Filt2::N], sh_ex[LEAN_WARPS * Filt2::N];
    const Seg sg = seg_of(a.g);  // segments past the end of the track (last group) hold no valid bin: identity
    const double2 *SA = a.SA + sg.base + sg.lane;
    const float *kap = a.kap + sg.base + sg.lane;
    const bool has_qs = a.qs != nullptr;
    const float *qs = has_qs ? a.qs + sg.base + sg.lane : nullptr;
    Filt2 g = filt2_identity();
    auto load = [&](ComposeIn &r, int i0) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int idx = (i0 + u) << 5;
            r.s[u] = SA[idx];
            r.kp[u] = kap[idx];
            r.q[u] = has_qs ? qs[idx] : 1.0f;
        }
    };
    auto step = [&](const ComposeIn &r, int i0) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            if (i0 + u < sg.valid) {
                const double qk = lean_qk(has_qs, r.kp[u], r.q[u], a.kap_min, a.kap_max);
                filt2_step<true>(g, a.M, qk * a.M.q00, qk * a.M.q01, qk * a.M.q11, r.s[u].x, r.s[u].y);
            }
        }
    };
    // groups of 4 bins; the inputs of the next group are in flight while this one is composed
    ComposeIn c0, c1;
    load(c0, 0);
#pragma unroll 1
    for (int i0 = 0; i0 < sg.L; i0 += 8) {
        load(c1, i0 + 4);
        step(c0, i0);
        if (i0 + 8 < sg.L) load(c0, i0 + 8);
        step(c1, i0 + 4);
    }
    // inclusive Kogge-Stone scan over the warp's 32 runs, then over the CTA's warps
    Filt2 inc = g;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const Filt2 o = shfl_up_elem(inc, d);
        if (sg.lane >= d) inc = filt2_combine(o, inc);
    }
    Filt2 ex = shfl_up_elem(inc, 1);
    if (sg.lane == 0) ex = filt2_identity();
    const int warp = threadIdx.x >> 5;
    if (sg.lane == 31) store_strided(sh + warp * Filt2::N, 1, inc);
    __syncthreads();
    if (warp == 0) {
        Filt2 v = sg.lane < LEAN_WARPS ? load_strided<Filt2>(sh + sg.lane * Filt2::N, 1) : filt2_identity();
#pragma unroll
        for (int d = 1; d < LEAN_WARPS; d <<= 1) {
            const Filt2 o = shfl_up_elem(v, d);
            if (sg.lane >= d) v = filt2_combine(o, v);
        }
        if (sg.lane == LEAN_WARPS - 1) store_strided(a.sc.fagg + blockIdx.x, a.g.Gp, v);  // the group's aggregate
        Filt2 x = shfl_up_elem(v, 1);
        if (sg.lane == 0) x = filt2_identity();
        if (sg.lane < LEAN_WARPS) store_strided(sh_ex + sg.lane * Filt2::N, 1, x);
    }
    __syncthreads();
    if (warp > 0) {
        const Filt2 wex = load_strided<Filt2>(sh_ex + warp * Filt2::N, 1);
        ex = sg.lane > 0 ? filt2_combine(wex, ex) : wex;
    }
    store_strided(a.sc.fex + (int64_t)sg.w * (Filt2::N * 32) + sg.lane, 32, ex);
}

// =====================================================================================
// segment scans (one CTA): forward prefix states, backward suffix states
// =====================================================================================
struct FiltOps {
    using Elem = Filt2;
    __device__ static __forceinline__ Elem identity() { return filt2_identity(); }
    __device__ static __forceinline__ Elem combine(const Elem &a, const Elem &b) { return filt2_combine(a, b); }
    __device__ static __forceinline__ State2 apply(const Elem &e, const State2 &s) { return filt2_apply(e, s); }
};
struct SmoOps {
    using Elem = Smo2;
    __device__ static __forceinline__ Elem identity() { return smo2_identity(); }
    __device__ static __forceinline__ Elem combine(const Elem &a, const Elem &b) { return smo2_combine(a, b); }
    __device__ static __forceinline__ State2 apply(const Elem &e, const State2 &s) { return smo2_apply(e, s); }
};

// Scan position j is group j (REVERSE = false) or group G - 1 - j (REVERSE = true); out[.][grp]
// receives the Gaussian the group starts from: `first` pushed through the aggregates of all scan
// positions before it.  agg: [N][pitch], out: [5][pitch].
template <class Ops, bool REVERSE>
__global__ void __launch_bounds__(LEAN_SCAN_THREADS) lean_group_scan_kernel(const double *agg, int G, int pitch,
                                                                           State2 first, const double *gathered,
                                                                           int rank, int world, double *out) {
    using Elem = typename Ops::Elem;
    if (gathered) {
        // a shard of a split chromosome: `first` pushed through the aggregates of the shards before it in
        // scan order (every thread does the same few applies: no barrier, no extra launch)
        if (!REVERSE) {
            for (int r = 0; r < rank; ++r) first = Ops::apply(load_strided<Elem>(gathered + (int64_t)r * LEAN_PAYLOAD, 1), first);
        } else {
            for (int r = world - 1; r > rank; --r) first = Ops::apply(load_strided<Elem>(gathered + (int64_t)r * LEAN_PAYLOAD, 1), first);
        }
    }
    constexpr int N = Elem::N;
    constexpr int NW = LEAN_SCAN_THREADS / 32;
    __shared__ double sh[NW * N];
    __shared__ double sh_ex[NW * N];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int c = (G + LEAN_SCAN_THREADS - 1) / LEAN_SCAN_THREADS;
    const int j0 = tid * c;
    auto grp = [&](int j) { return REVERSE ? G - 1 - j : j; };
    Elem e = Ops::identity();
    for (int u = 0; u < c; ++u) {
        const int j = j0 + u;
        if (j < G) {
            const Elem x = load_strided<Elem>(agg + grp(j), pitch);
            e = u == 0 ? x : Ops::combine(e, x);
        }
    }
    const int warps_live = min(NW, (G + 32 * c - 1) / (32 * c));  // warps that hold any group
    Elem inc = e;
    if (warp < warps_live) {
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const Elem o = shfl_up_elem(inc, d);
            if (lane >= d) inc = Ops::combine(o, inc);
        }
        if (lane == 31) store_strided(sh + warp * N, 1, inc);
    }
    __syncthreads();
    if (warp == 0 && warps_live > 1) {
        Elem v = lane < warps_live ? load_strided<Elem>(sh + lane * N, 1) : Ops::identity();
        for (int d = 1; d < warps_live; d <<= 1) {
            const Elem o = shfl_up_elem(v, d);
            if (lane >= d) v = Ops::combine(o, v);
        }
        Elem x = shfl_up_elem(v, 1);
        if (lane == 0) x = Ops::identity();
        if (lane < NW) store_strided(sh_ex + lane * N, 1, x);
    }
    __syncthreads();
    if (warp >= warps_live) return;
    Elem tex = shfl_up_elem(inc, 1);
    State2 st = first;
    if (tid > 0) {
        if (warp > 0) {
            const Elem wex = load_strided<Elem>(sh_ex + warp * N, 1);
            tex = lane > 0 ? Ops::combine(wex, tex) : wex;
        }
        st = Ops::apply(tex, first);
    }
    for (int u = 0; u < c; ++u) {
        const int j = j0 + u;
        if (j < G) {
            const int s = grp(j);
            store_strided(out + s, pitch, st);
            if (u + 1 < c) st = Ops::apply(load_strided<Elem>(agg + s, pitch), st);
        }
    }
}

// The whole aggregate of a shard (split chromosomes): ordered reduction of the group aggregates.
template <class Ops, bool REVERSE>
__global__ void __launch_bounds__(LEAN_SCAN_THREADS) lean_reduce_groups_kernel(const double *agg, int G, int pitch,
                                                                              double *out, const float *kap, const float *qs,
                                                                              const float4 *A, const float4 *B,
                                                                              int64_t last_pos) {
    using Elem = typename Ops::Elem;
    constexpr int N = Elem::N;
    constexpr int NW = LEAN_SCAN_THREADS / 32;
    __shared__ double sh[NW * N];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int c = (G + LEAN_SCAN_THREADS - 1) / LEAN_SCAN_THREADS;
    const int j0 = tid * c;
    Elem e = Ops::identity();
    for (int u = 0; u < c; ++u) {
        const int j = j0 + u;
        if (j < G) {
            const Elem x = load_strided<Elem>(agg + (REVERSE ? G - 1 - j : j), pitch);
            e = u == 0 ? x : Ops::combine(e, x);
        }
    }
    // ordered tree: lane l covers scan positions before lane l + d
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const Elem o = shfl_down_elem(e, d);
        if ((lane & (2 * d - 1)) == 0) e = Ops::combine(e, o);
    }
    if (lane == 0) store_strided(sh + warp * N, 1, e);
    __syncthreads();
    if (warp == 0) {
        Elem v = lane < NW ? load_strided<Elem>(sh + lane * N, 1) : Ops::identity();
#pragma unroll
        for (int d = 1; d < NW; d <<= 1) {
            const Elem o = shfl_down_elem(v, d);
            if ((lane & (2 * d - 1)) == 0) v = Ops::combine(v, o);
        }
        if (lane == 0) {
            store_strided(out, 1, v);
            // the rest of the payload: forward [14], [15] = kappa, qScale of the shard's first bin; backward
            // [9..13] = filtered Gaussian of its last bin
            if (!REVERSE) {
                out[14] = (double)kap[0];
                out[15] = qs ? (double)qs[0] : 1.0;
            } else {
                const float4 a = A[last_pos], b = B[last_pos];
                out[9] = (double)a.x; out[10] = (double)a.y; out[11] = (double)a.z; out[12] = (double)a.w;
                out[13] = (double)b.x;
            }
        }
    }
}

// =====================================================================================
// forward 3/3: replay of the reference recursion, compact track out, smoother run elements
// =====================================================================================
template <bool NLL>
struct ReplayIn {
    double2 s01[4], s2l[4];
    float kp[4], q[4];
};

// smoothing element of a bin from its filtered state (float32 values, as stored) and the float32
// process noise of the NEXT bin, composed onto the run element (later bins are the ones the
// reverse scan meets first)
__device__ __forceinline__ void compose_smo(const Model2 &M, Smo2 &brun, const Kf2 &f, double Q00, double Q01,
                                            double Q10, double Q11) {
    (void)Q10;  // the scan requires a symmetric Q0 (cb200: check_model)
    brun = smo2_combine(smo2_from_filtered_canon(M.F01, f.x0, f.x1, f.P00, f.P01, f.P11, r32(Q00), r32(Q01), r32(Q11)),
                        brun);
}

template <bool NLL>
__global__ void __launch_bounds__(LEAN_THREADS, LEAN_FWD_CTAS) lean_fwd_replay_kernel(const LeanFwdArgs a) {
    __shared__ double sh_part[LEAN_WARPS];
    __shared__ double sh[LEAN_WARPS * Smo2::N], sh_ex[LEAN_WARPS * Smo2::N];
    __shared__ int sh_last;
    const Seg sg = seg_of(a.g);
    const int warp = threadIdx.x >> 5;
    double nll = 0.0;
    Smo2 brun = smo2_identity();
    {
        const bool has_qs = a.qs != nullptr;
        // start state of the run: the group's Gaussian pushed through the runs of the group in front of it
        State2 start = load_strided<State2>(a.sc.fpref + blockIdx.x, a.g.Gp);
        if (threadIdx.x > 0)
            start = filt2_apply(load_strided<Filt2>(a.sc.fex + (int64_t)sg.w * (Filt2::N * 32) + sg.lane, 32), start);
        Kf2 s{r32(start.x0), r32(start.x1), r32(start.P00), r32(start.P01), r32(start.P01), r32(start.P11)};
        NllAcc acc;
        nll_acc_init(acc);
        const double2 *SA = a.SA + sg.base + sg.lane;
        const double2 *SB = a.SB + sg.base + sg.lane;
        const float *kap = a.kap + sg.base + sg.lane;
        const float *qs = has_qs ? a.qs + sg.base + sg.lane : nullptr;
        float4 *tA = a.trk.A + sg.base + sg.lane;
        float4 *tB = a.trk.B + sg.base + sg.lane;
        const bool store = a.do_store != 0;
        auto load = [&](ReplayIn<NLL> &r, int i0) {
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int idx = (i0 + u) << 5;
                r.s01[u] = __ldcs(SA + idx);
                if (NLL) r.s2l[u] = __ldcs(SB + idx);
                r.kp[u] = __ldcs(kap + idx);
                r.q[u] = has_qs ? __ldcs(qs + idx) : 1.0f;
            }
        };
        auto step = [&](const ReplayIn<NLL> &r, int i0) {
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int i = i0 + u;
                if (i < sg.valid) {
                    const double qk = lean_qk(has_qs, r.kp[u], r.q[u], a.kap_min, a.kap_max);
                    const Kf2 prev = s;  // filtered state of the previous bin, float32 values
                    BinOut o;
                    kf2_step<true>(s, a.M, qk, 1.0, r.s01[u].x, r.s01[u].y, NLL ? r.s2l[u].x : 0.0,
                                   NLL ? r.s2l[u].y : 0.0, a.m, a.inv_m, a.mlog2pi, NLL, false, o, acc);
                    if (store) {
                        if (i > 0) compose_smo(a.M, brun, prev, o.Q00, o.Q01, o.Q10, o.Q11);
                        tA[i << 5] = make_float4((float)s.x0, (float)s.x1, (float)s.P00, (float)s.P01);
                        tB[i << 5] = make_float4((float)s.P11, (float)o.Q00, (float)o.Q01, (float)o.Q11);
                    }
                }
            }
            if (NLL) nll_acc_renorm(acc, false);
        };
        ReplayIn<NLL> c0, c1;
        load(c0, 0);
#pragma unroll 1
        for (int i0 = 0; i0 < sg.L; i0 += 8) {
            load(c1, i0 + 4);
            step(c0, i0);
            if (i0 + 8 < sg.L) load(c0, i0 + 8);
            step(c1, i0 + 4);
        }
        if (NLL) nll = nll_acc_finish(acc, a.m, a.mlog2pi);
        if (store) {
            // the run's last bin: its element needs the NEXT run's first process noise, or it is the
            // terminal element of the chromosome (x_s = x_f, P_s = P_f there)
            if (sg.valid > 0) {
                const int64_t next = sg.k0 + sg.L;
                if (next >= a.g.n && a.sh.is_last) {
                    brun = smo2_combine(smo2_from_state(State2{s.x0, s.x1, s.P00, s.P01, s.P11}), brun);
                } else if (next >= a.g.n) {
                    // the shard's last bin: the next bin is the following shard's first one
                    const double qk = lean_qk(has_qs, (float)a.sh.fwd_next[14], (float)a.sh.fwd_next[15], a.kap_min, a.kap_max);
                    compose_smo(a.M, brun, s, qk * a.M.q00, qk * a.M.q01, qk * a.M.q10, qk * a.M.q11);
                } else {
                    const int64_t pn = a.g.index(next);
                    const double qk = lean_qk(has_qs, a.kap[pn], has_qs ? a.qs[pn] : 1.0f, a.kap_min, a.kap_max);
                    compose_smo(a.M, brun, s, qk * a.M.q00, qk * a.M.q01, qk * a.M.q10, qk * a.M.q11);
                }
            }
        }
    }
    if (a.do_store) {
        // reverse inclusive scan: lane l ends up with the composition of runs 31 .. l, then the same over
        // the CTA's warps (warp 7 first)
        Smo2 inc = brun;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const Smo2 o = shfl_down_elem(inc, d);
            if (sg.lane + d < 32) inc = smo2_combine(o, inc);
        }
        Smo2 ex = shfl_down_elem(inc, 1);
        if (sg.lane == 31) ex = smo2_identity();
        if (sg.lane == 0) store_strided(sh + warp * Smo2::N, 1, inc);
        __syncthreads();
        if (warp == 0) {
            // lane l holds warp LEAN_WARPS - 1 - l: scan position l of the reverse order
            Smo2 v = sg.lane < LEAN_WARPS ? load_strided<Smo2>(sh + (LEAN_WARPS - 1 - sg.lane) * Smo2::N, 1) : smo2_identity();
#pragma unroll
            for (int d = 1; d < LEAN_WARPS; d <<= 1) {
                const Smo2 o = shfl_up_elem(v, d);
                if (sg.lane >= d) v = smo2_combine(o, v);
            }
            if (sg.lane == LEAN_WARPS - 1) store_strided(a.trk.sagg + blockIdx.x, a.g.Gp, v);  // the group's aggregate
            Smo2 x = shfl_up_elem(v, 1);
            if (sg.lane == 0) x = smo2_identity();
            if (sg.lane < LEAN_WARPS) store_strided(sh_ex + (LEAN_WARPS - 1 - sg.lane) * Smo2::N, 1, x);
        }
        __syncthreads();
        if (warp < LEAN_WARPS - 1) {
            const Smo2 wex = load_strided<Smo2>(sh_ex + warp * Smo2::N, 1);
            ex = sg.lane < 31 ? smo2_combine(wex, ex) : wex;
        }
        store_strided(a.trk.sex + (int64_t)sg.w * (Smo2::N * 32) + sg.lane, 32, ex);
    }
    if (!NLL) return;
    // ---- sum of the NLL pieces: per segment, then (last CTA to finish) over the segments in order ----
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) nll += __shfl_xor_sync(FULL, nll, d);
    if (sg.lane == 0) sh_part[warp] = nll;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
#pragma unroll
        for (int i = 0; i < LEAN_WARPS; ++i) t += sh_part[i];
        a.sc.partials[blockIdx.x] = t;
        __threadfence();
        const int done = atomicAdd(a.sc.counter, 1);
        sh_last = done == (int)gridDim.x - 1;
    }
    __syncthreads();
    if (!sh_last) return;
    __threadfence();
    double t = 0.0;
    for (int i = threadIdx.x; i < a.g.G; i += LEAN_THREADS) t += __ldcg(a.sc.partials + i);
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) t += __shfl_xor_sync(FULL, t, d);
    __syncthreads();  // sh_part is re-used
    if (sg.lane == 0) sh_part[warp] = t;
    __syncthreads();
    if (threadIdx.x == 0) {
        double tot = 0.0;
#pragma unroll
        for (int i = 0; i < LEAN_WARPS; ++i) tot += sh_part[i];
        if (a.sums) {
            a.sums[0] = 0.0;
            a.sums[1] = tot;
        }
        *a.sc.counter = 0;
    }
}

// =====================================================================================
// backward 2/2: replay of the RTS recursion (+ kappa update, or the public tracks)
// =====================================================================================
struct BwdIn {
    float4 a[4], b[4];
    float q[4];
};

template <bool KAPPA, bool PUBLIC>
__global__ void __launch_bounds__(LEAN_THREADS, LEAN_BWD_CTAS) lean_bwd_replay_kernel(const LeanBwdArgs a) {
    const Seg sg = seg_of(a.g);
    if (sg.valid == 0) return;
    const bool has_qs = a.qs != nullptr;
    // smoothed Gaussian of the first bin of the next run: the one just beyond the group pushed through the
    // runs of the group behind this one
    State2 st = load_strided<State2>(a.ssuf + blockIdx.x, a.g.Gp);
    if (threadIdx.x < LEAN_THREADS - 1)
        st = smo2_apply(load_strided<Smo2>(a.trk.sex + (int64_t)sg.w * (Smo2::N * 32) + sg.lane, 32), st);
    Rs2 c{r32(st.x0), r32(st.x1), r32(st.P00), r32(st.P01), r32(st.P01), r32(st.P11)};
    const float4 *tA = a.trk.A + sg.base + sg.lane;
    const float4 *tB = a.trk.B + sg.base + sg.lane;
    const float *qs = has_qs ? a.qs + sg.base + sg.lane : nullptr;
    float *kout = KAPPA ? a.kap_out + sg.base + sg.lane : nullptr;
    // process noise / qScale of the bin after the one being smoothed, and where its kappa goes
    float qn0 = 0.f, qn1 = 0.f, qn2 = 0.f, qsn = 1.f;
    float *knext = nullptr;
    const int64_t next = sg.k0 + sg.L;
    if (next < a.g.n) {
        const int64_t pn = a.g.index(next);
        const float4 b = a.trk.B[pn];
        qn0 = b.y; qn1 = b.z; qn2 = b.w;
        if (has_qs) qsn = a.qs[pn];
        if (KAPPA) knext = a.kap_out + pn;
    } else if (!a.sh.is_last) {
        // the bin after the shard's last one is the next shard's first: its process noise as that shard's
        // forward pass stored it (float32 of qk Q0); its kappa is that shard's to compute
        const float nk = (float)a.sh.fwd_next[14], nq = (float)a.sh.fwd_next[15];
        const double qk = lean_qk(has_qs, nk, nq, a.kap_lo, a.kap_hi);
        qn0 = (float)(qk * a.M.q00); qn1 = (float)(qk * a.M.q01); qn2 = (float)(qk * a.M.q11);
        qsn = has_qs ? nq : 1.0f;
        if (KAPPA) knext = a.kap_discard;
    }
    auto load = [&](BwdIn &r, int i0) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int idx = (i0 + u) << 5;
            r.a[u] = __ldcs(tA + idx);
            r.b[u] = __ldcs(tB + idx);
            r.q[u] = has_qs ? __ldcs(qs + idx) : 1.0f;
        }
    };
    auto step = [&](const BwdIn &r, int i0) {
#pragma unroll
        for (int u = 3; u >= 0; --u) {
            const int i = i0 + u;
            if (i < sg.valid) {
                const float4 fa = r.a[u], fb = r.b[u];
                const int64_t k = sg.k0 + i;
                if (k == a.g.n - 1 && a.sh.is_last) {
                    // the chromosome's last bin: x_s = x_f, P_s = P_f
                    c = Rs2{(double)fa.x, (double)fa.y, (double)fa.z, (double)fa.w, (double)fa.w, (double)fb.x};
                    if (PUBLIC) {
                        reinterpret_cast<float2 *>(a.xs)[k] = make_float2(fa.x, fa.y);
                        reinterpret_cast<float4 *>(a.Ps)[k] = make_float4(fa.z, fa.w, fa.w, fb.x);
                    }
                } else {
                    const Rts2 g = rts2_gain<true>(a.M, fa.x, fa.y, fa.z, fa.w, fa.w, fb.x, qn0, qn1, qn1, qn2);
                    const Rs2 nxt = c;  // smoothed bin k+1 as the reference stores it (float32 values)
                    Smo2Out o;
                    rts2_step(c, g, fa.x, fa.y, fa.z, fa.w, fb.x, o);
                    if (KAPPA) {
                        // the reference reads its float32 tracks back: c now holds bin k rounded that way
                        *knext = (float)kappa2_update<true>(a.M, a.qi00, a.qi01, a.qi10, a.qi11, c.x0, c.x1, c.P00, c.P01,
                                                            c.P10, c.P11, nxt.x0, nxt.x1, nxt.P00, nxt.P01, nxt.P10,
                                                            nxt.P11, r32(o.C00), r32(o.C01), r32(o.C10), r32(o.C11),
                                                            (double)qsn, has_qs, a.nu, a.kap_lo, a.kap_hi);
                    }
                    if (PUBLIC) {
                        reinterpret_cast<float2 *>(a.xs)[k] = make_float2((float)o.xs0, (float)o.xs1);
                        reinterpret_cast<float4 *>(a.Ps)[k] =
                            make_float4((float)o.S00, (float)o.S01, (float)o.S01, (float)o.S11);
                        if (k < a.lag_rows)
                            reinterpret_cast<float4 *>(a.lag)[k] =
                                make_float4((float)o.C00, (float)o.C01, (float)o.C10, (float)o.C11);
                    }
                }
                qn0 = fb.y; qn1 = fb.z; qn2 = fb.w;
                qsn = r.q[u];
                if (KAPPA) knext = kout + (i << 5);
            }
        }
    };
    BwdIn c0, c1;
    load(c0, sg.L - 4);
#pragma unroll 1
    for (int i0 = sg.L - 4; i0 >= 0; i0 -= 8) {
        load(c1, i0 - 4);
        step(c0, i0);
        if (i0 - 8 >= 0) load(c0, i0 - 8);
        step(c1, i0 - 4);
    }
    if (KAPPA && sg.k0 == 0) {
        if (a.sh.is_first) {
            kout[0] = 1.0f;  // kappa_0 is not estimated (cconsenrich.pyx:8252)
        } else {
            // the multiplier of this shard's first bin: one more RTS step into the previous shard's last bin
            // (its filtered Gaussian came with the backward payload); c holds the smoothed first bin
            const double px0 = a.sh.bwd_prev[9], px1 = a.sh.bwd_prev[10], pP00 = a.sh.bwd_prev[11],
                         pP01 = a.sh.bwd_prev[12], pP11 = a.sh.bwd_prev[13];
            const Rts2 g = rts2_gain<true>(a.M, px0, px1, pP00, pP01, pP01, pP11, qn0, qn1, qn1, qn2);
            const Rs2 nxt = c;
            Smo2Out o;
            rts2_step(c, g, px0, px1, pP00, pP01, pP11, o);
            kout[0] = (float)kappa2_update<true>(a.M, a.qi00, a.qi01, a.qi10, a.qi11, c.x0, c.x1, c.P00, c.P01, c.P10,
                                                 c.P11, nxt.x0, nxt.x1, nxt.P00, nxt.P01, nxt.P10, nxt.P11, r32(o.C00),
                                                 r32(o.C01), r32(o.C10), r32(o.C11), (double)qsn, has_qs, a.nu, a.kap_lo,
                                                 a.kap_hi);
        }
    }
}

// =====================================================================================
// run-major <-> linear vectors
// =====================================================================================
__global__ void lean_gather_kernel(const float *__restrict__ linear, float *__restrict__ rm, LeanGeom g, float fill) {
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;  // run-major position
    if (p >= g.npad()) return;
    const int64_t sb = g.seg_bins();
    const int64_t r = p & (sb - 1);
    const int64_t k = (p - r) + ((r & 31) << g.logL) + (r >> 5);
    rm[p] = k < g.n ? linear[k] : fill;
}

__global__ void lean_scatter_kernel(const float *__restrict__ rm, float *__restrict__ linear, LeanGeom g) {
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= g.n) return;
    linear[k] = rm[g.index(k)];
}

__global__ void lean_fill_kernel(float *rm, int64_t count, float value) {
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p < count) rm[p] = value;
}

}  // namespace

// =====================================================================================
// launchers
// =====================================================================================
cudaError_t lean_configure() { return cudaSuccess; }

static unsigned lean_grid(const LeanGeom &g) { return (unsigned)g.G; }

cudaError_t lean_fwd_compose(const LeanFwdArgs &a, cudaStream_t st) {
    lean_fwd_compose_kernel<<<lean_grid(a.g), LEAN_THREADS, 0, st>>>(a);
    return cudaGetLastError();
}

cudaError_t lean_fwd_prefix(const LeanFwdArgs &a, cudaStream_t st) {
    const State2 prior{a.state_init, 0.0, a.cov_init, 0.0, a.cov_init};
    lean_group_scan_kernel<FiltOps, false><<<1, LEAN_SCAN_THREADS, 0, st>>>(a.sc.fagg, a.g.G, a.g.Gp, prior, a.sh.gathered,
                                                                      a.sh.rank, a.sh.world, a.sc.fpref);
    return cudaGetLastError();
}

cudaError_t lean_fwd_replay(const LeanFwdArgs &a, cudaStream_t st) {
    if (a.want_nll)
        lean_fwd_replay_kernel<true><<<lean_grid(a.g), LEAN_THREADS, 0, st>>>(a);
    else
        lean_fwd_replay_kernel<false><<<lean_grid(a.g), LEAN_THREADS, 0, st>>>(a);
    return cudaGetLastError();
}

cudaError_t lean_bwd_suffix(const LeanBwdArgs &a, cudaStream_t st) {
    const State2 beyond{0.0, 0.0, 0.0, 0.0, 0.0};
    lean_group_scan_kernel<SmoOps, true><<<1, LEAN_SCAN_THREADS, 0, st>>>(a.trk.sagg, a.g.G, a.g.Gp, beyond, a.sh.gathered,
                                                                    a.sh.rank, a.sh.world, a.ssuf);
    return cudaGetLastError();
}

cudaError_t lean_bwd_replay(const LeanBwdArgs &a, bool publish, cudaStream_t st) {
    if (publish)
        lean_bwd_replay_kernel<false, true><<<lean_grid(a.g), LEAN_THREADS, 0, st>>>(a);
    else
        lean_bwd_replay_kernel<true, false><<<lean_grid(a.g), LEAN_THREADS, 0, st>>>(a);
    return cudaGetLastError();
}

cudaError_t lean_shard_payload(const LeanFwdArgs &a, const LeanTrack &trk, bool backward, double *payload, cudaStream_t st) {
    const int64_t last = a.g.index(a.g.n - 1);
    if (backward)
        lean_reduce_groups_kernel<SmoOps, true><<<1, LEAN_SCAN_THREADS, 0, st>>>(trk.sagg, a.g.G, a.g.Gp, payload, a.kap, a.qs,
                                                                               trk.A, trk.B, last);
    else
        lean_reduce_groups_kernel<FiltOps, false><<<1, LEAN_SCAN_THREADS, 0, st>>>(a.sc.fagg, a.g.G, a.g.Gp, payload, a.kap,
                                                                                 a.qs, trk.A, trk.B, last);
    return cudaGetLastError();
}

cudaError_t lean_gather_f32(const float *linear, float *run_major, const LeanGeom &g, float fill, cudaStream_t st) {
    const int64_t np = g.npad();
    if (np <= 0) return cudaSuccess;
    lean_gather_kernel<<<(unsigned)((np + 255) / 256), 256, 0, st>>>(linear, run_major, g, fill);
    return cudaGetLastError();
}

cudaError_t lean_scatter_f32(const float *run_major, float *linear, const LeanGeom &g, cudaStream_t st) {
    if (g.n <= 0) return cudaSuccess;
    lean_scatter_kernel<<<(unsigned)((g.n + 255) / 256), 256, 0, st>>>(run_major, linear, g);
    return cudaGetLastError();
}

cudaError_t lean_fill_f32(float *run_major, const LeanGeom &g, float value, cudaStream_t st) {
    const int64_t np = g.npad();
    if (np <= 0) return cudaSuccess;
    lean_fill_kernel<<<(unsigned)((np + 255) / 256), 256, 0, st>>>(run_major, np, value);
    return cudaGetLastError();
}

}  // namespace cb200
