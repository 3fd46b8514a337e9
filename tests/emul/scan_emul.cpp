// tests/emul/scan_emul.cpp -- TEST INFRASTRUCTURE ONLY (never linked into the product).
//
// CPU emulation of the tiling the sm_100a scan kernels use (thread chunk -> Kogge-Stone
// scan inside a tile -> look-back over tile aggregates -> reference-ordered replay), built
// by g++ from the SAME header the kernels compile (consenrich_b200/csrc/ssm_math.cuh).
// It lets `pytest -m "not gpu"` check the scan algebra against the oracle on a box with
// no GPU.  The product never loads this file.
#include <cstdint>
#include <cstdlib>
#include <vector>

#include "../../consenrich_b200/csrc/ssm_math.cuh"

using namespace cb200;

namespace {

template <class E, class Comb>
void kogge_stone(std::vector<E> &v, Comb comb) {
    const size_t n = v.size();
    for (size_t d = 1; d < n; d <<= 1) {
        std::vector<E> nxt(v);
        for (size_t i = d; i < n; ++i) nxt[i] = comb(v[i - d], v[i]);
        v.swap(nxt);
    }
}

// ordered tree reduction of v[0..n) (v[0] earliest), the shape of the warp look-back
template <class E, class Comb>
E tree_reduce(std::vector<E> v, Comb comb) {
    while (v.size() > 1) {
        std::vector<E> nxt;
        for (size_t i = 0; i + 1 < v.size(); i += 2) nxt.push_back(comb(v[i], v[i + 1]));
        if (v.size() & 1) nxt.push_back(v.back());
        v.swap(nxt);
    }
    return v[0];
}

struct Lcg {
    uint64_t s;
    uint32_t next() {
        s = s * 6364136223846793005ULL + 1442695040888963407ULL;
        return (uint32_t)(s >> 33);
    }
};

}  // namespace

struct emul_params {
    double F[4], Q0[4];
    double state_init, cov_init;
    double lam_min, lam_max, kap_min, kap_max;
    int32_t use_lambda, use_kappa, use_qscale, return_nll, store_nll_in_d, do_store;
    int32_t chunk, tile_chunks;  // bins per thread, threads per tile
    int32_t canon, pad_;         // use the F = [[1, f], [0, 1]] specialisations
    uint64_t seed;               // randomises the look-back window lengths
};

namespace {

// 2-state forward filter from fold statistics.  Outputs as the reference (Qf[k-1] = Q_k).
// init5 (x0, x1, P00, P01, P11) replaces the model prior for a shard that continues a chromosome;
// q_head receives Q of the first bin in that case.
void forward2_impl(const double *S0, const double *S1, const double *S2, const double *SL, int64_t m,
                   int64_t n, const float *lam, const float *kap, const float *qscale,
                   const emul_params *p, float *D, float *xf, float *Pf, float *Qf, double *sum_d,
                   double *sum_nll, const double *init5, float *q_head) {
    Model2 M{p->F[0], p->F[1], p->F[2], p->F[3], p->Q0[0], p->Q0[1], p->Q0[2], p->Q0[3]};
    const int64_t L = p->chunk, T = p->tile_chunks;
    const int64_t nchunks = (n + L - 1) / L, ntiles = (nchunks + T - 1) / T;
    auto qk_of = [&](int64_t k) {
        double kappa = p->use_kappa ? clampd((double)kap[k], p->kap_min, p->kap_max) : 1.0;
        double qs = p->use_qscale ? (double)qscale[k] : 1.0;
        return qs / kappa;
    };
    auto lam_of = [&](int64_t k) {
        return p->use_lambda ? clampd((double)lam[k], p->lam_min, p->lam_max) : 1.0;
    };
    // pass 1: chunk aggregates
    std::vector<Filt2> agg(nchunks);
    for (int64_t c = 0; c < nchunks; ++c) {
        Filt2 g = filt2_identity();
        for (int64_t k = c * L; k < n && k < (c + 1) * L; ++k) {
            double qk = qk_of(k), l = lam_of(k);
            if (p->canon)
                filt2_step<true>(g, M, qk * M.q00, qk * M.q01, qk * M.q11, l * S0[k], l * S1[k]);
            else
                filt2_step<false>(g, M, qk * M.q00, qk * M.q01, qk * M.q11, l * S0[k], l * S1[k]);
        }
        agg[c] = g;
    }
    // tile-local inclusive scans + tile aggregates
    std::vector<Filt2> tile_agg(ntiles);
    std::vector<std::vector<Filt2>> incl(ntiles);
    for (int64_t t = 0; t < ntiles; ++t) {
        int64_t c0 = t * T, c1 = std::min(nchunks, c0 + T);
        incl[t].assign(agg.begin() + c0, agg.begin() + c1);
        kogge_stone(incl[t], filt2_combine);
        tile_agg[t] = incl[t].back();
    }
    // look-back: tile prefix states
    std::vector<State2> tile_prefix(ntiles + 1);
    tile_prefix[0] = init5 ? State2{init5[0], init5[1], init5[2], init5[3], init5[4]}
                           : State2{p->state_init, 0.0, p->cov_init, 0.0, p->cov_init};
    Lcg rng{p->seed | 1};
    for (int64_t t = 1; t <= ntiles; ++t) {
        // exclusive prefix of tile t = prefix(t - w) (x) agg[t-w .. t-1], random window w
        int64_t w = 1 + (int64_t)(rng.next() % 40);
        if (w > t) w = t;
        std::vector<Filt2> win;
        win.push_back(filt2_from_state(tile_prefix[t - w]));
        for (int64_t u = t - w; u < t; ++u) win.push_back(tile_agg[u]);
        Filt2 r = tree_reduce(win, filt2_combine);
        tile_prefix[t] = State2{r.b0, r.b1, r.C00, r.C01, r.C11};
    }
    // pass 2: replay every chunk from its exclusive prefix state in the reference's order
    const double mlog2pi = (double)m * log(6.2831853071795864769);
    *sum_d = 0.0;
    *sum_nll = 0.0;
    Kf2 head_carry{};
    for (int64_t c = 0; c < nchunks; ++c) {
        int64_t t = c / T, i = c % T;
        State2 s0 = tile_prefix[t];
        if (i > 0) s0 = filt2_apply(incl[t][i - 1], s0);
        Kf2 s{s0.x0, s0.x1, s0.P00, s0.P01, s0.P01, s0.P11};
        if (c > 0 || init5) {  // the reference carries float32-rounded values between bins
            s.x0 = r32(s.x0); s.x1 = r32(s.x1);
            s.P00 = r32(s.P00); s.P01 = r32(s.P01); s.P10 = s.P01; s.P11 = r32(s.P11);
        }
        // head replay: runs that start inside the first 16 bins continue the first run's replay
        if (c > 0 && c * L < 16 && !init5) s = head_carry;
        NllAcc acc;
        nll_acc_init(acc);
        for (int64_t k = c * L; k < n && k < (c + 1) * L; ++k) {
            BinOut o;
            if (p->canon)
                kf2_step<true>(s, M, qk_of(k), lam_of(k), S0[k], S1[k], S2[k], SL[k], (double)m, 1.0 / (double)m,
                               mlog2pi, p->return_nll != 0, p->store_nll_in_d != 0, o, acc);
            else
                kf2_step<false>(s, M, qk_of(k), lam_of(k), S0[k], S1[k], S2[k], SL[k], (double)m, 1.0 / (double)m,
                                mlog2pi, p->return_nll != 0, p->store_nll_in_d != 0, o, acc);
            D[k] = (float)o.stat;
            *sum_d += (double)D[k];
            *sum_nll += o.nll;
            if (p->do_store) {
                xf[k * 2] = (float)s.x0; xf[k * 2 + 1] = (float)s.x1;
                Pf[k * 4] = (float)s.P00; Pf[k * 4 + 1] = (float)s.P01;
                Pf[k * 4 + 2] = (float)s.P10; Pf[k * 4 + 3] = (float)s.P11;
                if (k > 0) {
                    Qf[(k - 1) * 4] = (float)o.Q00; Qf[(k - 1) * 4 + 1] = (float)o.Q01;
                    Qf[(k - 1) * 4 + 2] = (float)o.Q10; Qf[(k - 1) * 4 + 3] = (float)o.Q11;
                } else if (q_head) {
                    q_head[0] = (float)o.Q00; q_head[1] = (float)o.Q01;
                    q_head[2] = (float)o.Q10; q_head[3] = (float)o.Q11;
                }
            }
        }
        if (p->return_nll && !p->store_nll_in_d) *sum_nll += nll_acc_finish(acc, (double)m, mlog2pi);
        head_carry = s;
    }
}

}  // namespace

extern "C" {

void emul_forward2(const double *S0, const double *S1, const double *S2, const double *SL, int64_t m,
                   int64_t n, const float *lam, const float *kap, const float *qscale,
                   const emul_params *p, float *D, float *xf, float *Pf, float *Qf, double *sum_d,
                   double *sum_nll) {
    forward2_impl(S0, S1, S2, SL, m, n, lam, kap, qscale, p, D, xf, Pf, Qf, sum_d, sum_nll, nullptr, nullptr);
}

// ---- shards of a split chromosome (what the cb200_*_shard_* entry points compute) -------------
void emul_forward2_shard(const double *S0, const double *S1, const double *S2, const double *SL, int64_t m,
                         int64_t n, const float *lam, const float *kap, const float *qscale,
                         const emul_params *p, float *D, float *xf, float *Pf, float *Qf, double *sum_d,
                         double *sum_nll, const double *init5, float *q_head) {
    forward2_impl(S0, S1, S2, SL, m, n, lam, kap, qscale, p, D, xf, Pf, Qf, sum_d, sum_nll, init5, q_head);
}

// filtering element of a whole range: 14 doubles in the field order of Filt2
void emul_forward2_aggregate(const double *S0, const double *S1, int64_t n, const float *lam, const float *kap,
                             const float *qscale, const emul_params *p, double *agg) {
    Model2 M{p->F[0], p->F[1], p->F[2], p->F[3], p->Q0[0], p->Q0[1], p->Q0[2], p->Q0[3]};
    Filt2 g = filt2_identity();
    for (int64_t k = 0; k < n; ++k) {
        const double kappa = p->use_kappa ? clampd((double)kap[k], p->kap_min, p->kap_max) : 1.0;
        const double qk = (p->use_qscale ? (double)qscale[k] : 1.0) / kappa;
        const double l = p->use_lambda ? clampd((double)lam[k], p->lam_min, p->lam_max) : 1.0;
        filt2_step<true>(g, M, qk * M.q00, qk * M.q01, qk * M.q11, l * S0[k], l * S1[k]);
    }
    const double *src = reinterpret_cast<const double *>(&g);
    for (int i = 0; i < 14; ++i) agg[i] = src[i];
}

// prior pushed through the aggregates of shards 0..rank-1 (16-double pitch) -> init5
void emul_forward2_prefix(const double *aggs, int rank, const emul_params *p, double *init5) {
    State2 s{p->state_init, 0.0, p->cov_init, 0.0, p->cov_init};
    for (int r = 0; r < rank; ++r) {
        Filt2 g;
        double *dst = reinterpret_cast<double *>(&g);
        for (int i = 0; i < 14; ++i) dst[i] = aggs[r * 16 + i];
        s = filt2_apply(g, s);
    }
    init5[0] = s.x0; init5[1] = s.x1; init5[2] = s.P00; init5[3] = s.P01; init5[4] = s.P11;
}

void emul_forward1(const double *S0, const double *S1, const double *S2, const double *SL, int64_t m,
                   int64_t n, const float *lam, const float *kap, const float *qscale,
                   const emul_params *p, float *D, float *xf, float *Pf, float *Qf, double *sum_d,
                   double *sum_nll) {
    const double q0 = p->Q0[0];
    const int64_t L = p->chunk, T = p->tile_chunks;
    const int64_t nchunks = (n + L - 1) / L, ntiles = (nchunks + T - 1) / T;
    auto q_of = [&](int64_t k) {
        double kappa = p->use_kappa ? clampd((double)kap[k], p->kap_min, p->kap_max) : 1.0;
        double qs = p->use_qscale ? (double)qscale[k] : 1.0;
        return (qs / kappa) * q0;
    };
    auto lam_of = [&](int64_t k) {
        return p->use_lambda ? clampd((double)lam[k], p->lam_min, p->lam_max) : 1.0;
    };
    std::vector<Filt1> agg(nchunks);
    for (int64_t c = 0; c < nchunks; ++c) {
        Filt1 g = filt1_identity();
        for (int64_t k = c * L; k < n && k < (c + 1) * L; ++k) {
            double l = lam_of(k);
            filt1_step(g, q_of(k), l * S0[k], l * S1[k]);
        }
        agg[c] = g;
    }
    std::vector<Filt1> tile_agg(ntiles);
    std::vector<std::vector<Filt1>> incl(ntiles);
    for (int64_t t = 0; t < ntiles; ++t) {
        int64_t c0 = t * T, c1 = std::min(nchunks, c0 + T);
        incl[t].assign(agg.begin() + c0, agg.begin() + c1);
        kogge_stone(incl[t], filt1_combine);
        tile_agg[t] = incl[t].back();
    }
    std::vector<State1> tile_prefix(ntiles + 1);
    tile_prefix[0] = State1{p->state_init, p->cov_init};
    Lcg rng{p->seed | 1};
    for (int64_t t = 1; t <= ntiles; ++t) {
        int64_t w = 1 + (int64_t)(rng.next() % 40);
        if (w > t) w = t;
        std::vector<Filt1> win;
        win.push_back(filt1_from_state(tile_prefix[t - w]));
        for (int64_t u = t - w; u < t; ++u) win.push_back(tile_agg[u]);
        Filt1 r = tree_reduce(win, filt1_combine);
        tile_prefix[t] = State1{r.b, r.C};
    }
    const double mlog2pi = (double)m * log(6.2831853071795864769);
    *sum_d = 0.0;
    *sum_nll = 0.0;
    for (int64_t c = 0; c < nchunks; ++c) {
        int64_t t = c / T, i = c % T;
        State1 s = tile_prefix[t];
        if (i > 0) s = filt1_apply(incl[t][i - 1], s);
        NllAcc acc;
        nll_acc_init(acc);
        for (int64_t k = c * L; k < n && k < (c + 1) * L; ++k) {
            BinOut o;
            kf1_step(s, q_of(k), lam_of(k), S0[k], S1[k], S2[k], SL[k], (double)m, 1.0 / (double)m, mlog2pi,
                     p->return_nll != 0, p->store_nll_in_d != 0, o, acc);
            D[k] = (float)o.stat;
            *sum_d += (double)D[k];
            *sum_nll += o.nll;
            if (p->do_store) {
                xf[k] = (float)s.x;
                Pf[k] = (float)s.P;
                if (k > 0) Qf[k - 1] = (float)o.Q00;
            }
        }
        if (p->return_nll && !p->store_nll_in_d) *sum_nll += nll_acc_finish(acc, (double)m, mlog2pi);
    }
}

}  // extern "C"

namespace {

// 2-state RTS smoother as a reverse scan.  Position p = n-1-k runs forward.  tail5 = smoothed
// (x, P) of the first bin of the FOLLOWING shard (then row n-1 of Qf is live), or null.
void backward2_impl(int64_t n, const double *F, const float *xf, const float *Pf, const float *Qf,
                    const emul_params *p, float *xs, float *Ps, float *lagC, int64_t lag_rows,
                    const double *tail5) {
    if (n <= 0) return;
    const bool last = tail5 == nullptr;
    Model2 M{F[0], F[1], F[2], F[3], 0, 0, 0, 0};
    const int64_t L = p->chunk, T = p->tile_chunks;
    const int64_t nchunks = (n + L - 1) / L, ntiles = (nchunks + T - 1) / T;
    auto elem_of = [&](int64_t k) {
        if (k == n - 1 && last)
            return smo2_from_state(State2{xf[k * 2], xf[k * 2 + 1], Pf[k * 4], Pf[k * 4 + 1], Pf[k * 4 + 3]});
        Rts2 r = p->canon ? rts2_gain<true>(M, xf[k * 2], xf[k * 2 + 1], Pf[k * 4], Pf[k * 4 + 1], Pf[k * 4 + 2],
                                            Pf[k * 4 + 3], Qf[k * 4], Qf[k * 4 + 1], Qf[k * 4 + 2], Qf[k * 4 + 3])
                          : rts2_gain<false>(M, xf[k * 2], xf[k * 2 + 1], Pf[k * 4], Pf[k * 4 + 1], Pf[k * 4 + 2],
                                             Pf[k * 4 + 3], Qf[k * 4], Qf[k * 4 + 1], Qf[k * 4 + 2], Qf[k * 4 + 3]);
        return smo2_from_rts(r, xf[k * 2], xf[k * 2 + 1], Pf[k * 4], Pf[k * 4 + 1], Pf[k * 4 + 3]);
    };
    std::vector<Smo2> agg(nchunks);
    for (int64_t c = 0; c < nchunks; ++c) {
        Smo2 g = smo2_identity();
        for (int64_t q = c * L; q < n && q < (c + 1) * L; ++q) g = smo2_combine(g, elem_of(n - 1 - q));
        agg[c] = g;
    }
    std::vector<Smo2> tile_agg(ntiles);
    std::vector<std::vector<Smo2>> incl(ntiles);
    for (int64_t t = 0; t < ntiles; ++t) {
        int64_t c0 = t * T, c1 = std::min(nchunks, c0 + T);
        incl[t].assign(agg.begin() + c0, agg.begin() + c1);
        kogge_stone(incl[t], smo2_combine);
        tile_agg[t] = incl[t].back();
    }
    std::vector<State2> tile_prefix(ntiles + 1);
    tile_prefix[0] = last ? State2{0, 0, 0, 0, 0}  // irrelevant: bin n-1's element ignores it
                          : State2{tail5[0], tail5[1], tail5[2], tail5[3], tail5[4]};
    Lcg rng{p->seed | 1};
    for (int64_t t = 1; t <= ntiles; ++t) {
        int64_t w = 1 + (int64_t)(rng.next() % 40);
        if (w > t) w = t;
        std::vector<Smo2> win;
        win.push_back(smo2_from_state(tile_prefix[t - w]));
        for (int64_t u = t - w; u < t; ++u) win.push_back(tile_agg[u]);
        Smo2 r = tree_reduce(win, smo2_combine);
        tile_prefix[t] = State2{r.g0, r.g1, r.L00, r.L01, r.L11};
    }
    for (int64_t c = 0; c < nchunks; ++c) {
        int64_t t = c / T, i = c % T;
        State2 s0 = tile_prefix[t];
        if (i > 0) s0 = smo2_apply(incl[t][i - 1], s0);
        Rs2 cy{r32(s0.x0), r32(s0.x1), r32(s0.P00), r32(s0.P01), r32(s0.P01), r32(s0.P11)};
        for (int64_t q = c * L; q < n && q < (c + 1) * L; ++q) {
            int64_t k = n - 1 - q;
            if (k == n - 1 && last) {
                for (int e = 0; e < 2; ++e) xs[k * 2 + e] = xf[k * 2 + e];
                for (int e = 0; e < 4; ++e) Ps[k * 4 + e] = Pf[k * 4 + e];
                cy = Rs2{xf[k * 2], xf[k * 2 + 1], Pf[k * 4], Pf[k * 4 + 1], Pf[k * 4 + 2], Pf[k * 4 + 3]};
                continue;
            }
            Rts2 r = p->canon ? rts2_gain<true>(M, xf[k * 2], xf[k * 2 + 1], Pf[k * 4], Pf[k * 4 + 1], Pf[k * 4 + 2],
                                                Pf[k * 4 + 3], Qf[k * 4], Qf[k * 4 + 1], Qf[k * 4 + 2], Qf[k * 4 + 3])
                              : rts2_gain<false>(M, xf[k * 2], xf[k * 2 + 1], Pf[k * 4], Pf[k * 4 + 1], Pf[k * 4 + 2],
                                                 Pf[k * 4 + 3], Qf[k * 4], Qf[k * 4 + 1], Qf[k * 4 + 2], Qf[k * 4 + 3]);
            Smo2Out o;
            rts2_step(cy, r, xf[k * 2], xf[k * 2 + 1], Pf[k * 4], Pf[k * 4 + 1], Pf[k * 4 + 3], o);
            xs[k * 2] = (float)o.xs0; xs[k * 2 + 1] = (float)o.xs1;
            Ps[k * 4] = (float)o.S00; Ps[k * 4 + 1] = (float)o.S01;
            Ps[k * 4 + 2] = (float)o.S01; Ps[k * 4 + 3] = (float)o.S11;
            if (k < lag_rows) {
                lagC[k * 4] = (float)o.C00; lagC[k * 4 + 1] = (float)o.C01;
                lagC[k * 4 + 2] = (float)o.C10; lagC[k * 4 + 3] = (float)o.C11;
            }
        }
    }
}

}  // namespace

extern "C" {

void emul_backward2(int64_t n, const double *F, const float *xf, const float *Pf, const float *Qf,
                    const emul_params *p, float *xs, float *Ps, float *lagC, int64_t lag_rows) {
    backward2_impl(n, F, xf, Pf, Qf, p, xs, Ps, lagC, lag_rows, nullptr);
}

void emul_backward2_shard(int64_t n, const double *F, const float *xf, const float *Pf, const float *Qf,
                          const emul_params *p, float *xs, float *Ps, float *lagC, int64_t lag_rows,
                          const double *tail5) {
    backward2_impl(n, F, xf, Pf, Qf, p, xs, Ps, lagC, lag_rows, tail5);
}

// smoothing element of a whole range: 9 doubles in the field order of Smo2
void emul_backward2_aggregate(int64_t n, const double *F, const float *xf, const float *Pf, const float *Qf,
                              int is_last, double *agg) {
    Model2 M{F[0], F[1], F[2], F[3], 0, 0, 0, 0};
    Smo2 g = smo2_identity();
    for (int64_t k = n - 1; k >= 0; --k) {
        Smo2 e;
        if (k == n - 1 && is_last) {
            e = smo2_from_state(State2{xf[k * 2], xf[k * 2 + 1], Pf[k * 4], Pf[k * 4 + 1], Pf[k * 4 + 3]});
        } else {
            Rts2 r = rts2_gain<true>(M, xf[k * 2], xf[k * 2 + 1], Pf[k * 4], Pf[k * 4 + 1], Pf[k * 4 + 2],
                                     Pf[k * 4 + 3], Qf[k * 4], Qf[k * 4 + 1], Qf[k * 4 + 2], Qf[k * 4 + 3]);
            e = smo2_from_rts(r, xf[k * 2], xf[k * 2 + 1], Pf[k * 4], Pf[k * 4 + 1], Pf[k * 4 + 3]);
        }
        g = smo2_combine(g, e);
    }
    const double *src = reinterpret_cast<const double *>(&g);
    for (int i = 0; i < 9; ++i) agg[i] = src[i];
}

// tail state of shard `rank`: the aggregates of the shards after it, applied from the end
void emul_backward2_prefix(const double *aggs, int rank, int n_shards, double *tail5) {
    State2 s{0, 0, 0, 0, 0};
    for (int r = n_shards - 1; r > rank; --r) {
        Smo2 g;
        double *dst = reinterpret_cast<double *>(&g);
        for (int i = 0; i < 9; ++i) dst[i] = aggs[r * 16 + i];
        s = smo2_apply(g, s);
    }
    tail5[0] = s.x0; tail5[1] = s.x1; tail5[2] = s.P00; tail5[3] = s.P01; tail5[4] = s.P11;
}

void emul_backward1(int64_t n, const float *xf, const float *Pf, const float *Qf, const emul_params *p,
                    float *xs, float *Ps, float *lagC, int64_t lag_rows) {
    if (n <= 0) return;
    const int64_t L = p->chunk, T = p->tile_chunks;
    const int64_t nchunks = (n + L - 1) / L, ntiles = (nchunks + T - 1) / T;
    auto elem_of = [&](int64_t k) {
        if (k == n - 1) return smo1_from_state(State1{xf[k], Pf[k]});
        double pp, J;
        rts1_gain(Pf[k], Qf[k], pp, J);
        return smo1_from_rts(xf[k], Pf[k], pp, J);
    };
    std::vector<Smo1> agg(nchunks);
    for (int64_t c = 0; c < nchunks; ++c) {
        Smo1 g = smo1_identity();
        for (int64_t q = c * L; q < n && q < (c + 1) * L; ++q) g = smo1_combine(g, elem_of(n - 1 - q));
        agg[c] = g;
    }
    std::vector<Smo1> tile_agg(ntiles);
    std::vector<std::vector<Smo1>> incl(ntiles);
    for (int64_t t = 0; t < ntiles; ++t) {
        int64_t c0 = t * T, c1 = std::min(nchunks, c0 + T);
        incl[t].assign(agg.begin() + c0, agg.begin() + c1);
        kogge_stone(incl[t], smo1_combine);
        tile_agg[t] = incl[t].back();
    }
    std::vector<State1> tile_prefix(ntiles + 1);
    tile_prefix[0] = State1{0, 0};
    Lcg rng{p->seed | 1};
    for (int64_t t = 1; t <= ntiles; ++t) {
        int64_t w = 1 + (int64_t)(rng.next() % 40);
        if (w > t) w = t;
        std::vector<Smo1> win;
        win.push_back(smo1_from_state(tile_prefix[t - w]));
        for (int64_t u = t - w; u < t; ++u) win.push_back(tile_agg[u]);
        Smo1 r = tree_reduce(win, smo1_combine);
        tile_prefix[t] = State1{r.g, r.L};
    }
    for (int64_t c = 0; c < nchunks; ++c) {
        int64_t t = c / T, i = c % T;
        State1 s0 = tile_prefix[t];
        if (i > 0) s0 = smo1_apply(incl[t][i - 1], s0);
        double cx = r32(s0.x), cP = r32(s0.P);
        for (int64_t q = c * L; q < n && q < (c + 1) * L; ++q) {
            int64_t k = n - 1 - q;
            if (k == n - 1) {
                xs[k] = xf[k];
                Ps[k] = Pf[k];
                cx = xf[k];
                cP = Pf[k];
                continue;
            }
            double pf = Pf[k], pp, J;
            rts1_gain(pf, Qf[k], pp, J);
            double dx = cx - (double)xf[k];
            double x = (double)xf[k] + J * dx;
            double dP = cP - pp;
            double ps = pf + (J * J * dP);
            if (ps < 0.0) ps = 0.0;
            xs[k] = (float)x;
            Ps[k] = (float)ps;
            if (k < lag_rows) lagC[k] = (float)(pf + (J * dP));
            cx = (double)xs[k];
            cP = (double)Ps[k];
        }
    }
}

// fold statistics in the accumulation order of the fold kernel (sample-major per bin)
void emul_fold(const float *data, const float *munc, int64_t m, int64_t n, int64_t ld, double pad,
               double *S0, double *S1, double *S2, double *SL) {
    for (int64_t k = 0; k < n; ++k) {
        double s0 = 0, s1 = 0, s2 = 0, sl = 0;
        for (int64_t j = 0; j < m; ++j) {
            double z = data[j * ld + k], r = (double)munc[j * ld + k] + pad;
            if (r < 1.0e-12) r = 1.0e-12;
            double w = 1.0 / r;
            s0 += w;
            s1 += w * z;
            s2 += w * z * z;
            sl += log(r);
        }
        S0[k] = s0; S1[k] = s1; S2[k] = s2; SL[k] = sl;
    }
}

// Smoothing element of one bin two ways: the general route (rts2_gain + smo2_from_rts) and the
// canonical-F shortcut the lean forward replay uses (smo2_from_filtered_canon).  out: 2 x 9 doubles.
void emul_smoothing_element(double dF, const double *xP, const double *Q, double *out) {
    Model2 M{};
    M.F00 = 1.0; M.F01 = dF; M.F10 = 0.0; M.F11 = 1.0;
    const Rts2 r = rts2_gain<true>(M, xP[0], xP[1], xP[2], xP[3], xP[3], xP[4], Q[0], Q[1], Q[1], Q[2]);
    const Smo2 a = smo2_from_rts(r, xP[0], xP[1], xP[2], xP[3], xP[4]);
    const Smo2 b = smo2_from_filtered_canon(dF, xP[0], xP[1], xP[2], xP[3], xP[4], Q[0], Q[1], Q[2]);
    const double *pa = reinterpret_cast<const double *>(&a), *pb = reinterpret_cast<const double *>(&b);
    for (int i = 0; i < 9; ++i) {
        out[i] = pa[i];
        out[9 + i] = pb[i];
    }
}

}  // extern "C"
