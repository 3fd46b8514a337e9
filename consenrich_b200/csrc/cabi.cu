// consenrich_b200/csrc/cabi.cu -- C ABI of libconsenrich_b200.so (include/consenrich_b200.h).
//
// Host-side runtime around the sm_100a kernels: context (stream, device arena, launch
// accounting, optional per-kernel CUDA-event timing), the device-resident entry points, the
// ECM driver (reference cconsenrich.pyx:7877-8442 / 7188-7657) and the reference-facing
// host-buffer entry points.  There is no CPU implementation behind any of them.
#include <cuda_runtime.h>

#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <utility>
#include <string>
#include <vector>

#include "../../include/consenrich_b200.h"
#include "ssm_kernels.cuh"
#include "lean_kernels.cuh"
#include "background_kernels.cuh"
#include "munc_kernels.cuh"
#include "writer_kernels.cuh"

using namespace cb200;

namespace {

thread_local std::string g_err;

int fail(int code, const char *fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}

#define CU_TRY(expr)                                                                              \
    do {                                                                                          \
        cudaError_t _e = (expr);                                                                  \
        if (_e != cudaSuccess)                                                                    \
            return fail(CB200_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), \
                        __FILE__, __LINE__);                                                      \
    } while (0)

#define CB_TRY(expr)                \
    do {                            \
        int _rc = (expr);           \
        if (_rc != CB200_OK) return _rc; \
    } while (0)

// Entry points run on the context's device whatever device the calling thread had selected
// (one process per GPU under torchrun: rank r's thread sits on device r), and hand the
// thread's selection back on return.
struct DeviceGuard {
    int prev = -1;
    bool switched = false;
    explicit DeviceGuard(int device) {
        if (cudaGetDevice(&prev) == cudaSuccess && prev != device) switched = cudaSetDevice(device) == cudaSuccess;
    }
    ~DeviceGuard() {
        if (switched) cudaSetDevice(prev);
    }
};

struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
};

enum Family { FAM_FOLD = 0, FAM_FWD = 1, FAM_BWD = 2, FAM_RESID = 3, FAM_PREC = 4, FAM_BG = 5, FAM_MUNC = 6,
              FAM_COMPOSE = 7, FAM_SEGSCAN = 8, FAM_PUBLISH = 9, FAM_WRITER = 10, FAM_COUNT = 11 };

struct TimedSpan {
    cudaEvent_t a, b;
    int fam;
};

}  // namespace

static void split_free(void *p);

struct cb200_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    int64_t launches = 0;
    int32_t scan_epoch = 0;  // bumped per look-back scan launch (tags the workspace flags)
    int64_t scan_ws_n = 0;   // track length the scan workspace is laid out for
    long long *scan_dbg = nullptr;  // diagnostics buffer (cb200_debug_scan_times)
    int64_t scan_dbg_tiles = 0;
    int scan_dbg_pick = -1;  // CB200_DEBUG_STAMP_LAUNCH: only this scan launch after arming stamps (-1: all)
    int scan_dbg_seen = 0;
    // arena (device)
    DevBuf scan_ws, stats, sums, data, munc, xf, Pf, Qf, xf2, Pf2, Qf2, smo, smo2, D, xs, Ps, lag, resid, lam, kap, qs, shard;
    DevBuf bg_ws, bg_w, bg_rhs, bg_out, bg_status;  // background solve: workspace, operands, outcome
    // lean sweeps (lean_kernels.cuh): run-major statistics, multipliers, two sets of forward tracks, scratch
    DevBuf ln_SA, ln_SB, ln_kap, ln_qs, ln_A[2], ln_B[2], ln_sagg[2], ln_sex[2], ln_fagg, ln_fex, ln_fpref, ln_ssuf, ln_part;
    void *split = nullptr;  // SplitState of a chromosome shard between cb200_split_begin and cb200_split_end
    DevBuf wr_text, wr_tiles, wr_vals, wr_starts, wr_ends;  // bedGraph writer
    char *wr_host = nullptr;  // pinned text buffer
    size_t wr_host_cap = 0;
    DevBuf mask;  // MUNC stage: exclusion mask
    DevBuf seed_mat[6], seed_vec[7];  // MUNC seed pass: count floor, rho in, 4 outputs; 5 input + 2 output vectors
    double *sums_host = nullptr;  // pinned double[2]
    // timing
    bool timing = false;
    int timing_stride = 1;   // every timing_stride-th launch of a family is bracketed by events (1: all)
    std::vector<TimedSpan> spans;
    std::vector<cudaEvent_t> pool;
    double fam_ms[FAM_COUNT] = {};
    int64_t fam_n[FAM_COUNT] = {};      // launches bracketed by events
    int64_t fam_all[FAM_COUNT] = {};    // launches made while timing was enabled
};

namespace {

int ensure(cb200_ctx *c, DevBuf &b, size_t bytes) {
    if (bytes <= b.cap && b.p) return CB200_OK;
    if (b.p) {
        // buffers may still be in flight on the stream
        CU_TRY(cudaStreamSynchronize(c->stream));
        CU_TRY(cudaFree(b.p));
        b.p = nullptr;
        b.cap = 0;
    }
    size_t want = bytes < 256 ? 256 : bytes;
    CU_TRY(cudaMalloc(&b.p, want));
    b.cap = want;
    return CB200_OK;
}

// workspace view for the next look-back scan launch
int next_scan_ws(cb200_ctx *c, int64_t n, ScanWorkspace *ws) {
    // The layout (where flags and counters live) is fixed per allocation: the flags carry launch
    // epochs and the counters are left at zero by each launch, so they must not move between
    // launches of different lengths.
    bool fresh = false;
    if (n > c->scan_ws_n || !c->scan_ws.p) {
        const int64_t n_alloc = n + n / 4 + 4096;
        CB_TRY(ensure(c, c->scan_ws, scan_workspace_bytes(n_alloc)));
        c->scan_ws_n = n_alloc;
        fresh = true;
    }
    if (fresh || c->scan_epoch >= (1 << 28)) {
        CU_TRY(cudaMemsetAsync(c->scan_ws.p, 0, c->scan_ws.cap, c->stream));
        c->scan_epoch = 0;
    }
    c->scan_epoch += 1;
    *ws = scan_workspace_carve(c->scan_ws.p, c->scan_ws_n);
    ws->epoch4 = c->scan_epoch * 4;
    ws->dbg = nullptr;
    ws->dbg_tiles = 0;
    if (c->scan_dbg && (c->scan_dbg_pick < 0 || c->scan_dbg_seen == c->scan_dbg_pick)) {
        ws->dbg = c->scan_dbg;
        ws->dbg_tiles = (int32_t)(c->scan_dbg_tiles > 0x7fffffff ? 0x7fffffff : c->scan_dbg_tiles);
    }
    if (c->scan_dbg) c->scan_dbg_seen += 1;
    return CB200_OK;
}

struct Span {
    cb200_ctx *c;
    TimedSpan s{};
    bool on = false;
    Span(cb200_ctx *ctx, int fam) : c(ctx) {
        if (!c->timing) return;
        if ((c->fam_all[fam]++ % c->timing_stride) != 0) return;
        auto get = [&](cudaEvent_t &e) {
            if (!c->pool.empty()) {
                e = c->pool.back();
                c->pool.pop_back();
                return true;
            }
            return cudaEventCreate(&e) == cudaSuccess;
        };
        if (!get(s.a)) return;
        if (!get(s.b)) {
            c->pool.push_back(s.a);
            return;
        }
        s.fam = fam;
        on = true;
        cudaEventRecord(s.a, c->stream);
    }
    ~Span() {
        if (!on) return;
        cudaEventRecord(s.b, c->stream);
        c->spans.push_back(s);
    }
};

int resolve_spans(cb200_ctx *c) {
    if (c->spans.empty()) return CB200_OK;
    CU_TRY(cudaStreamSynchronize(c->stream));
    for (auto &s : c->spans) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, s.a, s.b) == cudaSuccess) {
            c->fam_ms[s.fam] += (double)ms;
            c->fam_n[s.fam] += 1;
        }
        c->pool.push_back(s.a);
        c->pool.push_back(s.b);
    }
    c->spans.clear();
    return CB200_OK;
}

int check_model(const cb200_model *mo) {
    if (!mo) return fail(CB200_ERR_INVALID, "model is NULL");
    if (mo->state_dim != 1 && mo->state_dim != 2) return fail(CB200_ERR_INVALID, "state_dim must be 1 or 2");
    if (mo->lam_min <= 0.0 || mo->lam_max <= 0.0 || mo->lam_max < mo->lam_min)
        return fail(CB200_ERR_INVALID, "observation precision multiplier bounds must satisfy 0 < min <= max");
    if (mo->kap_min <= 0.0 || mo->kap_max <= 0.0 || mo->kap_max < mo->kap_min)
        return fail(CB200_ERR_INVALID, "process precision multiplier bounds must satisfy 0 < min <= max");
    if (mo->state_dim == 1) {
        if (!(mo->Q0[0] > 0.0)) return fail(CB200_ERR_INVALID, "matrixQ0[0, 0] must be positive");
    } else {
        const double asym = fabs(mo->Q0[1] - mo->Q0[2]);
        const double scale = fabs(mo->Q0[0]) + fabs(mo->Q0[3]) + fabs(mo->Q0[1]) + 1e-300;
        if (asym > 1e-12 * scale)
            return fail(CB200_ERR_UNSUPPORTED,
                        "matrixQ0 must be symmetric: the filtering scan carries symmetric covariances");
    }
    return CB200_OK;
}

Model2 to_model2(const cb200_model *mo) {
    Model2 M;
    if (mo->state_dim == 2) {
        M.F00 = mo->F[0]; M.F01 = mo->F[1]; M.F10 = mo->F[2]; M.F11 = mo->F[3];
        M.q00 = mo->Q0[0]; M.q01 = mo->Q0[1]; M.q10 = mo->Q0[2]; M.q11 = mo->Q0[3];
    } else {
        M.F00 = 1.0; M.F01 = 0.0; M.F10 = 0.0; M.F11 = 1.0;
        M.q00 = mo->Q0[0]; M.q01 = M.q10 = M.q11 = 0.0;
    }
    return M;
}

bool aligned(const void *p, size_t a) { return (reinterpret_cast<uintptr_t>(p) % a) == 0; }

int g_lean_off = -1;   // CB200_NO_LEAN / cb200_set_lean_sweeps: always the look-back kernels
int g_lean_logL = 0;   // CB200_LEAN_LOGL / cb200_set_lean_sweeps: run length override (5 or 6)

void lean_read_env() {
    if (g_lean_off >= 0) return;
    const char *e = getenv("CB200_NO_LEAN");
    g_lean_off = (e && *e && *e != '0') ? 1 : 0;
    const char *l = getenv("CB200_LEAN_LOGL");
    g_lean_logL = (l && *l) ? atoi(l) : 0;
}

void lean_set_mode(int mode, int log2_run) {
    g_lean_off = mode ? 0 : 1;
    g_lean_logL = log2_run;
}

int64_t round_up(int64_t v, int64_t a) { return (v + a - 1) / a * a; }

// ---- device-level building blocks ------------------------------------------------------
struct SmoFuse {  // smoothing run elements composed by the forward replay for the backward scan
    double *run = nullptr;  // 9 arrays of `pitch` doubles
    int64_t pitch = 0;
    int nsub = 0;           // sub-steps per run both scans use
    int64_t npad = 0;       // forward tiles x tile length: where the backward scan counts positions from
};

int do_fold(cb200_ctx *c, const float *data, const float *munc, int64_t m, int64_t n, int64_t ld, double pad,
            double *stats, int64_t stride) {
    Span sp(c, FAM_FOLD);
    CU_TRY(launch_fold(data, munc, m, n, ld, pad, reinterpret_cast<double2 *>(stats),
                       reinterpret_cast<double2 *>(stats + 2 * stride), c->stream));
    c->launches += 1;
    return CB200_OK;
}

// ECM_useAPN is live exactly when the reference's loop would feed D_k back into the process noise
// (cconsenrich.pyx:510, 6574-6576, 6997-6998)
bool apn_live(const cb200_model *mo) {
    if (!mo->use_apn || mo->use_qscale) return false;
    const double q_diag = mo->state_dim == 2 ? 0.5 * (mo->Q0[0] + mo->Q0[3]) : mo->Q0[0];
    return q_diag > 1.0e-12;
}

int do_forward(cb200_ctx *c, const cb200_model *mo, const double *stats, int64_t stride, int64_t m, int64_t n,
               const float *lam, const float *kap, const float *qs, const double *init_state, float *xf, float *Pf,
               float *Qf, float *D, double *sums, double *agg_out, bool aggregate_only, float *q_head = nullptr,
               const SmoFuse *sf = nullptr) {
    const int d = mo->state_dim;
    const bool store = (xf != nullptr);
    if (store && (!Pf || !Qf)) return fail(CB200_ERR_INVALID, "xf, Pf and Qf must be given together");
    if (store && d == 2 && (!aligned(xf, 8) || !aligned(Pf, 16) || !aligned(Qf, 16)))
        return fail(CB200_ERR_INVALID, "forward outputs must be 8/16-byte aligned device pointers");
    if (mo->use_lambda && !lam) return fail(CB200_ERR_INVALID, "lambdaExp is required when use_lambda is set");
    if (mo->use_kappa && !kap) return fail(CB200_ERR_INVALID, "processPrecExp is required when use_kappa is set");
    if (mo->use_qscale && !qs) return fail(CB200_ERR_INVALID, "processQScale is required when use_qscale is set");
    if (apn_live(mo)) {
        // adaptive process noise: a sequential recursion over the fold statistics (apn_kernels.cu)
        if (aggregate_only || init_state)
            return fail(CB200_ERR_UNSUPPORTED, "adaptive process noise cannot run on a shard of a split chromosome: the "
                                               "feedback is sequential over the whole chromosome");
        if (mo->use_kappa) return fail(CB200_ERR_INVALID, "process precision multipliers are not live under adaptive process noise");
        ApnArgs p{};
        p.SA = reinterpret_cast<const double2 *>(stats);
        p.SB = reinterpret_cast<const double2 *>(stats + 2 * stride);
        p.lam = lam;
        p.xf = xf; p.Pf = Pf; p.Qf = Qf; p.D = D;
        p.sums = sums;
        p.n = n;
        p.m = (double)m;
        p.inv_m = 1.0 / (double)m;
        p.mlog2pi = (double)m * log(6.2831853071795864769);
        p.M = to_model2(mo);
        p.state_init = mo->state_init;
        p.cov_init = mo->cov_init;
        p.lam_min = mo->lam_min; p.lam_max = mo->lam_max;
        p.apn_min_q = mo->apn_min_q; p.apn_max_q = mo->apn_max_q; p.apn_thresh = mo->apn_thresh;
        p.apn_scale = mo->apn_scale; p.apn_pc = mo->apn_pc;
        p.q_diag = d == 2 ? 0.5 * (mo->Q0[0] + mo->Q0[3]) : mo->Q0[0];
        p.use_lambda = mo->use_lambda; p.want_nll = mo->return_nll; p.nll_in_d = mo->store_nll_in_d;
        p.do_store = store ? 1 : 0;
        {
            Span sp(c, FAM_FWD);
            CU_TRY(launch_apn_forward(d, p, c->stream));
        }
        c->launches += 1;
        return CB200_OK;
    }
    ScanWorkspace ws;
    CB_TRY(next_scan_ws(c, n, &ws));
    FwdArgs a{};
    a.SA = reinterpret_cast<const double2 *>(stats);
    a.SB = reinterpret_cast<const double2 *>(stats + 2 * stride);
    a.lam = lam; a.kap = kap; a.qs = qs;
    a.init_state = init_state;
    a.xf = xf; a.Pf = Pf; a.Qf = Qf; a.D = D;
    a.q_head = q_head;
    if (sf && sf->run && d == 2 && !aggregate_only && !init_state) {
        a.smo_run = sf->run;
        a.smo_pitch = sf->pitch;
        a.nsub = sf->nsub;
    }
    a.sums = sums;
    a.agg_out = agg_out;
    a.n = n;
    a.m = (double)m;
    a.inv_m = 1.0 / (double)m;
    a.mlog2pi = (double)m * log(6.2831853071795864769);
    a.M = to_model2(mo);
    a.state_init = mo->state_init;
    a.cov_init = mo->cov_init;
    a.lam_min = mo->lam_min; a.lam_max = mo->lam_max;
    a.kap_min = mo->kap_min; a.kap_max = mo->kap_max;
    a.use_lambda = mo->use_lambda; a.use_kappa = mo->use_kappa; a.use_qscale = mo->use_qscale;
    a.want_nll = mo->return_nll; a.nll_in_d = mo->store_nll_in_d;
    a.do_store = store ? 1 : 0;
    int launches = 0;
    {
        Span sp(c, FAM_FWD);
        CU_TRY(launch_forward(d, a, ws, aggregate_only, c->stream, &launches));
    }
    c->launches += launches;
    return CB200_OK;
}

struct KappaFuse {  // Student-t process precision update carried out inside the backward replay
    float *kap_out = nullptr;
    bool no_store = false;  // do not write the smoothed tracks: only kappa is wanted
    const float *qs = nullptr;
    double nu = 0.0;
};

int do_backward(cb200_ctx *c, const cb200_model *mo, int64_t n, const float *xf, const float *Pf, const float *Qf,
                const double *tail_state, int is_last, float *xs, float *Ps, float *lag, int64_t lag_rows,
                double *agg_out, bool aggregate_only, const KappaFuse *kf = nullptr, const SmoFuse *sf = nullptr) {
    const int d = mo->state_dim;
    if (d == 2 && (!aligned(xf, 8) || !aligned(Pf, 16) || !aligned(Qf, 16) ||
                   (!aggregate_only && (!aligned(xs, 8) || !aligned(Ps, 16) || !aligned(lag, 16)))))
        return fail(CB200_ERR_INVALID, "smoother tracks must be 8/16-byte aligned device pointers");
    ScanWorkspace ws;
    CB_TRY(next_scan_ws(c, n, &ws));
    BwdArgs a{};
    a.xf = xf; a.Pf = Pf; a.Qf = Qf;
    a.tail_state = tail_state;
    a.xs = xs; a.Ps = Ps; a.lag = lag;
    a.agg_out = agg_out;
    a.n = n;
    a.lag_rows = lag_rows;
    a.M = to_model2(mo);
    a.is_last_shard = is_last;
    if (sf && sf->run && d == 2 && !aggregate_only && is_last && !tail_state) {
        a.smo_run = sf->run;
        a.smo_pitch = sf->pitch;
        a.npad_fixed = sf->npad;
        a.nsub = sf->nsub;
    }
    if (kf && kf->kap_out && !aggregate_only) {
        a.kap_out = kf->kap_out;
        a.no_store = kf->no_store ? 1 : 0;
        a.qs = kf->qs;
        a.nu = kf->nu;
        a.kap_lo = mo->kap_min;
        a.kap_hi = mo->kap_max;
        if (d == 2) {
            const double det = mo->Q0[0] * mo->Q0[3] - mo->Q0[1] * mo->Q0[2];
            if (det == 0.0) return fail(CB200_ERR_INVALID, "matrixQ0 is singular");
            a.qi00 = mo->Q0[3] / det; a.qi01 = -mo->Q0[1] / det; a.qi10 = -mo->Q0[2] / det; a.qi11 = mo->Q0[0] / det;
        } else {
            a.qi00 = 1.0 / mo->Q0[0];
        }
    }
    int launches = 0;
    {
        Span sp(c, FAM_BWD);
        CU_TRY(launch_backward(d, a, ws, aggregate_only, c->stream, &launches));
    }
    c->launches += launches;
    return CB200_OK;
}

int do_residuals(cb200_ctx *c, const float *data, int64_t m, int64_t n, int64_t ld, const float *xs, int dim,
                 float *resid) {
    Span sp(c, FAM_RESID);
    CU_TRY(launch_residuals(data, m, n, ld, xs, dim, resid, c->stream));
    c->launches += 1;
    return CB200_OK;
}

int check_block_map(const int32_t *bm, int64_t n, int64_t block_count) {
    if (block_count <= 0) return fail(CB200_ERR_INVALID, "blockCount must be positive");
    if (!bm) return fail(CB200_ERR_INVALID, "intervalToBlockMap length must match intervalCount");
    int32_t lo = 0, hi = 0;
    if (n > 0) lo = hi = bm[0];
    for (int64_t k = 1; k < n; ++k) {
        const int32_t b = bm[k];
        lo = b < lo ? b : lo;
        hi = b > hi ? b : hi;
    }
    if (n > 0 && (lo < 0 || (int64_t)hi >= block_count))
        return fail(CB200_ERR_INVALID, "intervalToBlockMap has out-of-range block id");
    return CB200_OK;
}

int h2d(cb200_ctx *c, void *dst, const void *src, size_t bytes) {
    if (bytes == 0) return CB200_OK;
    CU_TRY(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, c->stream));
    return CB200_OK;
}
int d2h(cb200_ctx *c, void *dst, const void *src, size_t bytes) {
    if (bytes == 0) return CB200_OK;
    CU_TRY(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, c->stream));
    return CB200_OK;
}

// upload [m x n] host tracks into a pitched device buffer; returns ld
int upload_tracks(cb200_ctx *c, DevBuf &buf, const float *host, int64_t m, int64_t n, int64_t *ld_out) {
    const int64_t ld = round_up(n, 32);
    CB_TRY(ensure(c, buf, (size_t)m * (size_t)ld * sizeof(float)));
    CU_TRY(cudaMemcpy2DAsync(buf.p, (size_t)ld * 4, host, (size_t)n * 4, (size_t)n * 4, (size_t)m,
                             cudaMemcpyHostToDevice, c->stream));
    *ld_out = ld;
    return CB200_OK;
}

int upload_vec(cb200_ctx *c, DevBuf &buf, const float *host, int64_t n, bool live, const float **dev_out) {
    *dev_out = nullptr;
    if (!live) return CB200_OK;
    if (!host) return fail(CB200_ERR_INVALID, "a per-interval vector flagged live in the model is NULL");
    CB_TRY(ensure(c, buf, (size_t)n * 4));
    CB_TRY(h2d(c, buf.p, host, (size_t)n * 4));
    *dev_out = static_cast<const float *>(buf.p);
    return CB200_OK;
}

// a multiplier vector that starts at 1: filled on the device, nothing crosses the bus
int ones_vec(cb200_ctx *c, DevBuf &buf, int64_t n, const float **dev_out) {
    CB_TRY(ensure(c, buf, (size_t)n * 4));
    CU_TRY(launch_fill(static_cast<float *>(buf.p), n, 1.0f, c->stream));
    c->launches += 1;
    *dev_out = static_cast<const float *>(buf.p);
    return CB200_OK;
}

int read_sums(cb200_ctx *c, double *out2) {
    CB_TRY(d2h(c, c->sums_host, c->sums.p, 16));
    CU_TRY(cudaStreamSynchronize(c->stream));
    out2[0] = c->sums_host[0];
    out2[1] = c->sums_host[1];
    return CB200_OK;
}

}  // namespace

// =====================================================================================
// context
// =====================================================================================
extern "C" {

int cb200_abi_version(void) { return CB200_ABI_VERSION; }

int cb200_current_device(int *device) {
    if (!device) return fail(CB200_ERR_INVALID, "device is NULL");
    CU_TRY(cudaGetDevice(device));
    return CB200_OK;
}

const char *cb200_last_error(void) { return g_err.c_str(); }

int cb200_ctx_create(int device, void *stream, cb200_ctx **out) {
    if (!out) return fail(CB200_ERR_INVALID, "out is NULL");
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count <= 0)
        return fail(CB200_ERR_CUDA, "no CUDA device available (%s); libconsenrich_b200 has no CPU path",
                    e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
    if (device < 0 || device >= count) return fail(CB200_ERR_INVALID, "device %d out of range [0, %d)", device, count);
    DeviceGuard _dg(device);
    cudaDeviceProp prop;
    CU_TRY(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10)
        return fail(CB200_ERR_CUDA, "device %d is sm_%d%d; this library carries sm_100a code only", device, prop.major,
                    prop.minor);
    cb200_ctx *c = new cb200_ctx();
    c->device = device;
    if (stream) {
        c->stream = static_cast<cudaStream_t>(stream);
    } else {
        if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess) {
            delete c;
            return fail(CB200_ERR_CUDA, "cudaStreamCreate failed");
        }
        c->own_stream = true;
    }
    if (configure_kernels() != cudaSuccess || cudaMalloc(&c->sums.p, 256) != cudaSuccess ||
        cudaMallocHost(reinterpret_cast<void **>(&c->sums_host), 64) != cudaSuccess) {
        cudaError_t le = cudaGetLastError();
        cb200_ctx_destroy(c);
        return fail(CB200_ERR_CUDA, "context setup failed: %s", cudaGetErrorString(le));
    }
    c->sums.cap = 256;
    *out = c;
    return CB200_OK;
}

void cb200_ctx_destroy(cb200_ctx *c) {
    DeviceGuard _dg(c ? c->device : 0);
    if (!c) return;
    split_free(c->split);
    c->split = nullptr;
    if (c->stream) cudaStreamSynchronize(c->stream);
    DevBuf *bufs[] = {&c->scan_ws, &c->stats, &c->sums, &c->data, &c->munc, &c->xf, &c->Pf, &c->Qf, &c->xf2, &c->Pf2,
                      &c->Qf2, &c->smo, &c->smo2, &c->D,
                      &c->xs, &c->Ps, &c->lag, &c->resid, &c->lam, &c->kap, &c->qs, &c->shard,
                      &c->bg_ws, &c->bg_w, &c->bg_rhs, &c->bg_out, &c->bg_status, &c->mask,
                      &c->ln_SA, &c->ln_SB, &c->ln_kap, &c->ln_qs, &c->ln_A[0], &c->ln_A[1], &c->ln_B[0], &c->ln_B[1],
                      &c->ln_sagg[0], &c->ln_sagg[1], &c->ln_sex[0], &c->ln_sex[1], &c->ln_fagg, &c->ln_fex, &c->ln_fpref,
                      &c->ln_ssuf, &c->ln_part, &c->wr_text, &c->wr_tiles, &c->wr_vals, &c->wr_starts, &c->wr_ends,
                      &c->seed_mat[0], &c->seed_mat[1], &c->seed_mat[2], &c->seed_mat[3], &c->seed_mat[4], &c->seed_mat[5],
                      &c->seed_vec[0], &c->seed_vec[1], &c->seed_vec[2], &c->seed_vec[3], &c->seed_vec[4], &c->seed_vec[5],
                      &c->seed_vec[6]};
    for (DevBuf *b : bufs)
        if (b->p) cudaFree(b->p);
    if (c->sums_host) cudaFreeHost(c->sums_host);
    if (c->wr_host) cudaFreeHost(c->wr_host);
    for (auto &s : c->spans) {
        cudaEventDestroy(s.a);
        cudaEventDestroy(s.b);
    }
    for (auto &e : c->pool) cudaEventDestroy(e);
    if (c->own_stream && c->stream) cudaStreamDestroy(c->stream);
    delete c;
}

int cb200_ctx_set_stream(cb200_ctx *c, void *stream) {
    DeviceGuard _dg(c ? c->device : 0);
    if (!c) return fail(CB200_ERR_INVALID, "ctx is NULL");
    CU_TRY(cudaStreamSynchronize(c->stream));
    if (c->own_stream) {
        cudaStreamDestroy(c->stream);
        c->own_stream = false;
    }
    if (stream) {
        c->stream = static_cast<cudaStream_t>(stream);
    } else {
        CU_TRY(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
        c->own_stream = true;
    }
    return CB200_OK;
}

int cb200_ctx_sync(cb200_ctx *c) {
    DeviceGuard _dg(c ? c->device : 0);
    if (!c) return fail(CB200_ERR_INVALID, "ctx is NULL");
    CU_TRY(cudaStreamSynchronize(c->stream));
    return CB200_OK;
}

int64_t cb200_ctx_launch_count(const cb200_ctx *c) {
    DeviceGuard _dg(c ? c->device : 0); return c ? c->launches : 0; }

int cb200_ctx_enable_timing(cb200_ctx *c, int on) {
    DeviceGuard _dg(c ? c->device : 0);
    if (!c) return fail(CB200_ERR_INVALID, "ctx is NULL");
    CB_TRY(resolve_spans(c));
    c->timing = on != 0;
    return CB200_OK;
}

int cb200_ctx_kernel_ms(cb200_ctx *c, int family, double *ms, int64_t *launches) {
    DeviceGuard _dg(c ? c->device : 0);
    if (!c || family < 0 || family >= FAM_COUNT) return fail(CB200_ERR_INVALID, "bad kernel family");
    CB_TRY(resolve_spans(c));
    if (ms) *ms = c->fam_ms[family];
    if (launches) *launches = c->fam_n[family];
    return CB200_OK;
}

int cb200_ctx_kernel_launches(cb200_ctx *c, int family, int64_t *launches) {
    DeviceGuard _dg(c ? c->device : 0);
    if (!c || family < 0 || family >= FAM_COUNT || !launches) return fail(CB200_ERR_INVALID, "bad kernel family");
    *launches = c->fam_all[family];
    return CB200_OK;
}

int cb200_ctx_set_timing_stride(cb200_ctx *c, int stride) {
    DeviceGuard _dg(c ? c->device : 0);
    if (!c || stride < 1) return fail(CB200_ERR_INVALID, "timing stride must be >= 1");
    c->timing_stride = stride;
    return CB200_OK;
}

int cb200_debug_scan_times(cb200_ctx *c, int64_t tiles, long long *host_out) {
    DeviceGuard _dg(c ? c->device : 0);
    if (!c) return fail(CB200_ERR_INVALID, "ctx is NULL");
    CU_TRY(cudaStreamSynchronize(c->stream));
    if (host_out && c->scan_dbg) {
        const int64_t t = tiles < c->scan_dbg_tiles ? tiles : c->scan_dbg_tiles;
        CU_TRY(cudaMemcpy(host_out, c->scan_dbg, (size_t)t * 8 * sizeof(long long), cudaMemcpyDeviceToHost));
    }
    if (c->scan_dbg) {
        CU_TRY(cudaFree(c->scan_dbg));
        c->scan_dbg = nullptr;
        c->scan_dbg_tiles = 0;
    }
    if (tiles > 0 && !host_out) {  // arm: the next scan launches stamp their phases
        CU_TRY(cudaMalloc(reinterpret_cast<void **>(&c->scan_dbg), (size_t)tiles * 8 * sizeof(long long)));
        CU_TRY(cudaMemset(c->scan_dbg, 0, (size_t)tiles * 8 * sizeof(long long)));
        c->scan_dbg_tiles = tiles;
        const char *pick = getenv("CB200_DEBUG_STAMP_LAUNCH");
        c->scan_dbg_pick = pick && *pick ? atoi(pick) : -1;
        c->scan_dbg_seen = 0;
    }
    return CB200_OK;
}

int cb200_set_scan_substeps(int nsub) {
    if (nsub < 0 || nsub > MAX_NSUB) return fail(CB200_ERR_INVALID, "scan sub-steps must be in [0, %d]", MAX_NSUB);
    scan_set_nsub_override(nsub);
    return CB200_OK;
}

int cb200_set_lean_sweeps(int mode, int log2_run) {
    if (mode != 0 && mode != 1) return fail(CB200_ERR_INVALID, "lean sweep mode must be 0 or 1");
    if (log2_run != 0 && log2_run != 5 && log2_run != 6) return fail(CB200_ERR_INVALID, "log2_run must be 0, 5 or 6");
    lean_set_mode(mode, log2_run);
    return CB200_OK;
}

int cb200_ctx_reset_timing(cb200_ctx *c) {
    DeviceGuard _dg(c ? c->device : 0);
    if (!c) return fail(CB200_ERR_INVALID, "ctx is NULL");
    CB_TRY(resolve_spans(c));
    for (int i = 0; i < FAM_COUNT; ++i) {
        c->fam_ms[i] = 0.0;
        c->fam_n[i] = 0;
        c->fam_all[i] = 0;
    }
    return CB200_OK;
}

// =====================================================================================
// memory helpers
// =====================================================================================
int cb200_device_alloc(cb200_ctx *c, size_t bytes, void **dptr) {
    DeviceGuard _dg(c ? c->device : 0);
    if (!c || !dptr) return fail(CB200_ERR_INVALID, "ctx/dptr is NULL");
    CU_TRY(cudaSetDevice(c->device));
    CU_TRY(cudaMalloc(dptr, bytes < 256 ? 256 : bytes));
    return CB200_OK;
}

int cb200_device_free(cb200_ctx *c, void *dptr) {
    DeviceGuard _dg(c ? c->device : 0);
    if (!c) return fail(CB200_ERR_INVALID, "ctx is NULL");
    if (!dptr) return CB200_OK;
    CU_TRY(cudaStreamSynchronize(c->stream));
    CU_TRY(cudaFree(dptr));
    return CB200_OK;
}

int cb200_pinned_alloc(size_t bytes, void **hptr) {
    if (!hptr) return fail(CB200_ERR_INVALID, "hptr is NULL");
    CU_TRY(cudaMallocHost(hptr, bytes < 64 ? 64 : bytes));
    return CB200_OK;
}

int cb200_pinned_free(void *hptr) {
    if (!hptr) return CB200_OK;
    CU_TRY(cudaFreeHost(hptr));
    return CB200_OK;
}

int cb200_copy_h2d(cb200_ctx *c, void *dst, size_t dst_pitch, const void *src, size_t src_pitch, size_t row_bytes,
                   size_t rows) {
    DeviceGuard _dg(c ? c->device : 0);
    if (!c) return fail(CB200_ERR_INVALID, "ctx is NULL");
    if (rows == 0 || row_bytes == 0) return CB200_OK;
    CU_TRY(cudaMemcpy2DAsync(dst, dst_pitch, src, src_pitch, row_bytes, rows, cudaMemcpyHostToDevice, c->stream));
    return CB200_OK;
}

int cb200_copy_d2h(cb200_ctx *c, void *dst, size_t dst_pitch, const void *src, size_t src_pitch, size_t row_bytes,
                   size_t rows) {
    DeviceGuard _dg(c ? c->device : 0);
    if (!c) return fail(CB200_ERR_INVALID, "ctx is NULL");
    if (rows == 0 || row_bytes == 0) return CB200_OK;
    CU_TRY(cudaMemcpy2DAsync(dst, dst_pitch, src, src_pitch, row_bytes, rows, cudaMemcpyDeviceToHost, c->stream));
    return CB200_OK;
}

// =====================================================================================
// device-resident path
// =====================================================================================
int cb200_fold_tracks(cb200_ctx *c, const float *data, const float *munc, int64_t m, int64_t n, int64_t ld,
                      double pad, double *stats, int64_t stat_stride) {
    DeviceGuard _dg(c ? c->device : 0);
    if (!c || !data || !munc || !stats) return fail(CB200_ERR_INVALID, "NULL argument");
    if (m <= 0 || n <= 0) return CB200_OK;
    if (ld < n) return fail(CB200_ERR_INVALID, "row stride ld must be >= n");
    if (stat_stride < n || (stat_stride & 1)) return fail(CB200_ERR_INVALID, "stat_stride must be even and >= n");
    if (!aligned(stats, 16)) return fail(CB200_ERR_INVALID, "stats must be 16-byte aligned");
    return do_fold(c, data, munc, m, n, ld, pad, stats, stat_stride);
}

int cb200_forward_scan(cb200_ctx *c, const cb200_model *mo, const double *stats, int64_t stat_stride, int64_t m,
                       int64_t n, const float *lam, const float *kap, const float *qscale, const double *init_state,
                       float *xf, float *Pf, float *Qf, float *D, double *sums) {
    DeviceGuard _dg(c ? c->device : 0);
    if (!c || !stats) return fail(CB200_ERR_INVALID, "NULL argument");
    CB_TRY(check_model(mo));
    if (n <= 0) return CB200_OK;
    return do_forward(c, mo, stats, stat_stride, m, n, lam, kap, qscale, init_state, xf, Pf, Qf, D, sums, nullptr,
                      false);
}

int cb200_forward_scan_shard(cb200_ctx *c, const cb200_model *mo, const double *stats, int64_t stat_stride, int64_t m,
                             int64_t n, const float *lam, const float *kap, const float *qscale,
                             const double *init_state, float *xf, float *Pf, float *Qf, float *D, double *sums,
                             float *q_head) {
    DeviceGuard _dg(c ? c->device : 0);
    if (!c || !stats) return fail(CB200_ERR_INVALID, "NULL argument");
    CB_TRY(check_model(mo));
    if (n <= 0) return fail(CB200_ERR_INVALID, "a shard must hold at least one interval");
    return do_forward(c, mo, stats, stat_stride, m, n, lam, kap, qscale, init_state, xf, Pf, Qf, D, sums, nullptr,
                      false, q_head);
}

int cb200_forward_shard_aggregate(cb200_ctx *c, const cb200_model *mo, const double *stats, int64_t stat_stride,
                                  int64_t n, const float *lam, const float *kap, const float *qscale, double *agg) {
    DeviceGuard _dg(c ? c->device : 0);
    if (!c || !stats || !agg) return fail(CB200_ERR_INVALID, "NULL argument");
    CB_TRY(check_model(mo));
    if (n <= 0) return fail(CB200_ERR_INVALID, "a shard must hold at least one interval");
    return do_forward(c, mo, stats, stat_stride, 1, n, lam, kap, qscale, nullptr, nullptr, nullptr, nullptr, nullptr,
                      nullptr, agg, true);
}

int cb200_forward_shard_prefix(cb200_ctx *c, const cb200_model *mo, const double *aggs, int32_t rank,
                               double *init_state) {
    DeviceGuard _dg(c ? c->device : 0);
    if (!c || !aggs || !init_state) return fail(CB200_ERR_INVALID, "NULL argument");
    CB_TRY(check_model(mo));
    CU_TRY(launch_forward_shard_prefix(mo->state_dim, aggs, rank, mo->state_init, mo->cov_init, init_state,
                                       c->stream));
    c->launches += 1;
    return CB200_OK;
}

int cb200_backward_scan(cb200_ctx *c, const cb200_model *mo, int64_t n, const float *xf, const float *Pf,
                        const float *Qf, const double *tail_state, float *xs, float *Ps, float *lag,
                        int64_t lag_rows) {
    DeviceGuard _dg(c ? c->device : 0);
    if (!c || !xf || !Pf || !Qf || !xs || !Ps || !lag) return fail(CB200_ERR_INVALID, "NULL argument");
    CB_TRY(check_model(mo));
    if (n <= 0) return CB200_OK;
    return do_backward(c, mo, n, xf, Pf, Qf, tail_state, tail_state == nullptr, xs, Ps, lag, lag_rows, nullptr, false);
}

int cb200_backward_scan_kappa(cb200_ctx *c, const cb200_model *mo, int64_t n, const float *xf, const float *Pf,
                              const float *Qf, float *xs, float *Ps, float *lag, int64_t lag_rows, const float *qscale,
                              double nu, float *kap) {
    DeviceGuard _dg(c ? c->device : 0);
    if (!c || !xf || !Pf || !Qf || !xs || !Ps || !lag || !kap) return fail(CB200_ERR_INVALID, "NULL argument");
    CB_TRY(check_model(mo));
    if (n <= 0) return CB200_OK;
    KappaFuse kf;
    kf.kap_out = kap;
    kf.qs = qscale;
    kf.nu = nu;
    return do_backward(c, mo, n, xf, Pf, Qf, nullptr, 1, xs, Ps, lag, lag_rows, nullptr, false, &kf);
}

int cb200_backward_shard_aggregate(cb200_ctx *c, const cb200_model *mo, int64_t n, const float *xf, const float *Pf,
                                   const float *Qf, int32_t is_last_shard, double *agg) {
    DeviceGuard _dg(c ? c->device : 0);
    if (!c || !xf || !Pf || !Qf || !agg) return fail(CB200_ERR_INVALID, "NULL argument");
    CB_TRY(check_model(mo));
    if (n <= 0) return fail(CB200_ERR_INVALID, "a shard must hold at least one interval");
    return do_backward(c, mo, n, xf, Pf, Qf, nullptr, is_last_shard, nullptr, nullptr, nullptr, 0, agg, true);
}

int cb200_backward_shard_prefix(cb200_ctx *c, const cb200_model *mo, const double *aggs, int32_t rank,
                                int32_t n_shards, double *tail_state) {
    DeviceGuard _dg(c ? c->device : 0);
    if (!c || !aggs || !tail_state) return fail(CB200_ERR_INVALID, "NULL argument");
    CB_TRY(check_model(mo));
    CU_TRY(launch_backward_shard_prefix(mo->state_dim, aggs, rank, n_shards, tail_state, c->stream));
    c->launches += 1;
    return CB200_OK;
}

int cb200_residuals(cb200_ctx *c, const float *data, int64_t m, int64_t n, int64_t ld, const float *xs,
                    int32_t state_dim, float *resid) {
    DeviceGuard _dg(c ? c->device : 0);
    if (!c || !data || !xs || !resid) return fail(CB200_ERR_INVALID, "NULL argument");
    if (state_dim != 1 && state_dim != 2) return fail(CB200_ERR_INVALID, "state_dim must be 1 or 2");
    if (m <= 0 || n <= 0) return CB200_OK;
    return do_residuals(c, data, m, n, ld, xs, state_dim, resid);
}

int cb200_update_lambda(cb200_ctx *c, const cb200_model *mo, const double *stats, int64_t stat_stride, int64_t m,
                        int64_t n, const float *xs, const float *Ps, double nu, float *lam) {
    DeviceGuard _dg(c ? c->device : 0);
    if (!c || !stats || !xs || !Ps || !lam) return fail(CB200_ERR_INVALID, "NULL argument");
    CB_TRY(check_model(mo));
    Span sp(c, FAM_PREC);
    CU_TRY(launch_update_lambda(reinterpret_cast<const double2 *>(stats),
                                reinterpret_cast<const double2 *>(stats + 2 * stat_stride), n, (double)m, xs, Ps,
                                mo->state_dim, nu, mo->lam_min, mo->lam_max, lam, c->stream));
    c->launches += 1;
    return CB200_OK;
}

int cb200_update_kappa(cb200_ctx *c, const cb200_model *mo, int64_t n, const float *xs, const float *Ps,
                       const float *lag, const float *qscale, double nu, float *kap) {
    DeviceGuard _dg(c ? c->device : 0);
    if (!c || !xs || !Ps || !lag || !kap) return fail(CB200_ERR_INVALID, "NULL argument");
    CB_TRY(check_model(mo));
    if (mo->state_dim == 2) {
        const double det = mo->Q0[0] * mo->Q0[3] - mo->Q0[1] * mo->Q0[2];
        if (det == 0.0) return fail(CB200_ERR_INVALID, "matrixQ0 is singular");
    }
    Span sp(c, FAM_PREC);
    CU_TRY(launch_update_kappa(mo->state_dim, to_model2(mo), n, xs, Ps, lag, qscale, nu, mo->kap_min, mo->kap_max,
                               kap, c->stream));
    c->launches += 1;
    return CB200_OK;
}

}  // extern "C"

// =====================================================================================
// ECM on run-major private tracks (lean_kernels.cuh)
// =====================================================================================
namespace {

bool lean_eligible(const cb200_model *mo, const cb200_ecm_opts *op, int64_t n, const float *lam, const float *kap) {
    lean_read_env();
    if (g_lean_off || op->max_iters <= 0 || op->inner_iters <= 0) return false;
    return mo->state_dim == 2 && lam == nullptr && kap != nullptr && n >= LEAN_MIN_BINS && mo->F[0] == 1.0 &&
           mo->F[2] == 0.0 && mo->F[3] == 1.0;
}

struct EcmLoopState {
    double prev = 1.0e16, cur = 0.0, init_nll = 0.0, rel_impr = 0.0, abs_rel = 0.0;
    bool has_init = false, converged = false;
    int iters_done = 0, stable = 0, inc = 0;
};

// convergence bookkeeping of one ECM iteration (cconsenrich.pyx:8337-8407); true = stop
bool ecm_iteration_done(EcmLoopState &L, double rtol, int patience) {
    const bool has_prev = L.has_init;
    if (!has_prev) {
        L.init_nll = L.cur;
        L.has_init = true;
    } else if (L.cur > L.prev + (1.0e-12 * fmax(fabs(L.prev), 1.0))) {
        L.inc += 1;
    }
    double delta, scale;
    if (has_prev) {
        delta = fabs(L.cur - L.prev);
        scale = fabs(L.prev);
    } else {
        delta = 0.0;
        scale = fabs(L.cur);
    }
    scale = fmax(scale, fabs(L.cur));
    scale = fmax(scale, 1.0);
    if (has_prev) {
        L.rel_impr = (L.prev - L.cur) / scale;
        L.abs_rel = delta / scale;
    } else {
        L.rel_impr = L.abs_rel = 0.0;
    }
    const double tol = rtol * scale;
    L.prev = L.cur;
    L.stable = (has_prev && delta <= tol) ? L.stable + 1 : 0;
    if (L.stable >= patience) {
        L.converged = true;
        return true;
    }
    return false;
}

// Everything one lean ECM needs on the device: geometry, the two pass argument blocks, the two sets of
// forward tracks.  lean_setup allocates (context arena), folds the tracks into run-major statistics and
// brings kappa / qScale into run-major order.
struct LeanRun {
    LeanGeom g;
    LeanFwdArgs fa;
    LeanBwdArgs ba;
    LeanTrack trk[2];
    float *kap_rm = nullptr, *qs_rm = nullptr;
};

int lean_setup(cb200_ctx *c, const cb200_model *mo, double nu, const float *data, const float *munc, int64_t m,
               int64_t n, int64_t ld, const float *qscale, const float *kap, LeanRun *R) {
    // run length: 32 bins (measured on B200: faster than 64 at chr19 and at chr1 @ 25 bp alike)
    int logL = g_lean_logL >= 5 && g_lean_logL <= 6 ? g_lean_logL : 5;
    const LeanGeom g = lean_geom(n, logL);
    const size_t np = (size_t)g.npad(), G = (size_t)g.G, Gp = (size_t)g.Gp, segs = G * LEAN_WARPS;
    CB_TRY(ensure(c, c->ln_SA, np * 16));
    CB_TRY(ensure(c, c->ln_SB, np * 16));
    CB_TRY(ensure(c, c->ln_kap, np * 4));
    if (qscale) CB_TRY(ensure(c, c->ln_qs, np * 4));
    for (int s = 0; s < 2; ++s) {
        CB_TRY(ensure(c, c->ln_A[s], np * 16));
        CB_TRY(ensure(c, c->ln_B[s], np * 16));
        CB_TRY(ensure(c, c->ln_sagg[s], Gp * 9 * 8));
        CB_TRY(ensure(c, c->ln_sex[s], segs * 9 * 32 * 8));
    }
    CB_TRY(ensure(c, c->ln_fagg, Gp * 14 * 8));
    CB_TRY(ensure(c, c->ln_fex, segs * 14 * 32 * 8));
    CB_TRY(ensure(c, c->ln_fpref, Gp * 5 * 8));
    CB_TRY(ensure(c, c->ln_ssuf, Gp * 5 * 8));
    if (!c->ln_part.p || c->ln_part.cap < 512 + G * 8) {
        CB_TRY(ensure(c, c->ln_part, 512 + G * 8 + G * 2));  // headroom: the counter must start at zero only once
        CU_TRY(cudaMemsetAsync(c->ln_part.p, 0, 512, c->stream));
    }
    // layout of ln_part: [0] counter (int32, left at zero by every launch); [128] the float a non-last shard
    // drops its boundary kappa into; partial sums from byte 512
    {
        Span sp(c, FAM_FOLD);
        CU_TRY(launch_fold(data, munc, m, n, ld, mo->pad, static_cast<double2 *>(c->ln_SA.p),
                           static_cast<double2 *>(c->ln_SB.p), c->stream, logL));
        c->launches += 1;
    }
    float *kap_rm = static_cast<float *>(c->ln_kap.p);
    float *qs_rm = nullptr;
    {
        Span sp(c, FAM_PREC);
        CU_TRY(lean_gather_f32(kap, kap_rm, g, 1.0f, c->stream));
        c->launches += 1;
        if (qscale) {
            qs_rm = static_cast<float *>(c->ln_qs.p);
            CU_TRY(lean_gather_f32(qscale, qs_rm, g, 1.0f, c->stream));
            c->launches += 1;
        }
    }
    R->g = g;
    R->kap_rm = kap_rm;
    R->qs_rm = qs_rm;
    for (int s = 0; s < 2; ++s) {
        R->trk[s].A = static_cast<float4 *>(c->ln_A[s].p);
        R->trk[s].B = static_cast<float4 *>(c->ln_B[s].p);
        R->trk[s].sagg = static_cast<double *>(c->ln_sagg[s].p);
        R->trk[s].sex = static_cast<double *>(c->ln_sex[s].p);
    }
    unsigned char *part = static_cast<unsigned char *>(c->ln_part.p);
    LeanFwdArgs fa{};
    fa.g = g;
    fa.SA = static_cast<const double2 *>(c->ln_SA.p);
    fa.SB = static_cast<const double2 *>(c->ln_SB.p);
    fa.kap = kap_rm;
    fa.qs = qs_rm;
    fa.sc.fagg = static_cast<double *>(c->ln_fagg.p);
    fa.sc.fex = static_cast<double *>(c->ln_fex.p);
    fa.sc.fpref = static_cast<double *>(c->ln_fpref.p);
    fa.sc.counter = reinterpret_cast<int32_t *>(part);
    fa.sc.partials = reinterpret_cast<double *>(part + 512);
    fa.sums = static_cast<double *>(c->sums.p);
    fa.m = (double)m;
    fa.inv_m = 1.0 / (double)m;
    fa.mlog2pi = (double)m * log(6.2831853071795864769);
    fa.M = to_model2(mo);
    fa.state_init = mo->state_init;
    fa.cov_init = mo->cov_init;
    fa.kap_min = mo->kap_min;
    fa.kap_max = mo->kap_max;
    fa.sh.is_first = fa.sh.is_last = 1;
    LeanBwdArgs ba{};
    ba.g = g;
    ba.ssuf = static_cast<double *>(c->ln_ssuf.p);
    ba.qs = qs_rm;
    ba.kap_out = kap_rm;
    ba.lag_rows = n > 1 ? n - 1 : 1;
    ba.M = fa.M;
    ba.nu = nu;
    ba.kap_lo = mo->kap_min;
    ba.kap_hi = mo->kap_max;
    ba.sh.is_first = ba.sh.is_last = 1;
    ba.kap_discard = reinterpret_cast<float *>(part + 128);
    {
        const double det = mo->Q0[0] * mo->Q0[3] - mo->Q0[1] * mo->Q0[2];
        ba.qi00 = mo->Q0[3] / det; ba.qi01 = -mo->Q0[1] / det; ba.qi10 = -mo->Q0[2] / det; ba.qi11 = mo->Q0[0] / det;
    }
    R->fa = fa;
    R->ba = ba;
    return CB200_OK;
}

int ecm_device_lean(cb200_ctx *c, const cb200_model *mo, const cb200_ecm_opts *op, const float *data,
                    const float *munc, int64_t m, int64_t n, int64_t ld, const float *qscale, float *kap, float *xs,
                    float *Ps, float *lag, float *resid, cb200_ecm_result *res, double *nll_path) {
    LeanRun R;
    CB_TRY(lean_setup(c, mo, op->nu, data, munc, m, n, ld, qscale, kap, &R));
    LeanFwdArgs &fa = R.fa;
    LeanBwdArgs &ba = R.ba;
    LeanTrack *trk = R.trk;
    const LeanGeom g = R.g;
    float *kap_rm = R.kap_rm;
    ba.xs = xs; ba.Ps = Ps; ba.lag = lag;
    int cur_set = 0;
    auto forward = [&](int set, bool with_nll, bool store) -> int {
        fa.trk = trk[set];
        fa.want_nll = with_nll ? 1 : 0;
        fa.do_store = store ? 1 : 0;
        {
            Span sp(c, FAM_COMPOSE);
            CU_TRY(lean_fwd_compose(fa, c->stream));
        }
        {
            Span sp(c, FAM_SEGSCAN);
            CU_TRY(lean_fwd_prefix(fa, c->stream));
        }
        {
            Span sp(c, FAM_FWD);
            CU_TRY(lean_fwd_replay(fa, c->stream));
        }
        c->launches += 3;
        return CB200_OK;
    };
    auto backward = [&](bool publish) -> int {
        ba.trk = trk[cur_set];
        if (!publish) {
            Span sp(c, FAM_SEGSCAN);
            CU_TRY(lean_bwd_suffix(ba, c->stream));
            c->launches += 1;
        }
        {
            Span sp(c, publish ? FAM_PUBLISH : FAM_BWD);
            CU_TRY(lean_bwd_replay(ba, publish, c->stream));
        }
        c->launches += 1;
        return CB200_OK;
    };
    EcmLoopState L;
    bool opened = false;  // the forward pass of the next sweep has already run (into the current set)
    for (int i = 0; i < op->max_iters; ++i) {
        L.iters_done = i + 1;
        for (int t = 0; t < op->inner_iters; ++t) {
            if (!opened) CB_TRY(forward(cur_set, false, true));
            opened = false;
            CB_TRY(backward(false));  // kappa is rewritten in place: only the forward pass reads it
        }
        const bool ahead = i + 1 < op->max_iters && op->inner_iters > 0;
        // NLL of this iteration; with another iteration to come it is the by-product of that iteration's
        // opening forward pass, run ahead into the spare set
        CB_TRY(forward(ahead ? 1 - cur_set : cur_set, true, ahead));
        double s2[2];
        CB_TRY(read_sums(c, s2));
        L.cur = s2[1];
        if (nll_path) nll_path[i] = L.cur;
        if (ecm_iteration_done(L, op->rtol, 2)) break;  // tracks run ahead into the spare set are dropped
        if (ahead) {
            cur_set = 1 - cur_set;
            opened = true;
        }
    }
    {
        Span sp(c, FAM_PREC);
        CU_TRY(lean_scatter_f32(kap_rm, kap, g, c->stream));
        c->launches += 1;
    }
    CB_TRY(backward(true));  // the smoothed tracks the call returns: the last sweep's, in the reference's layouts
    if (resid) CB_TRY(do_residuals(c, data, m, n, ld, xs, 2, resid));
    CU_TRY(cudaStreamSynchronize(c->stream));
    res->iters_done = L.iters_done;
    res->converged = L.converged ? 1 : 0;
    res->skipped = 0;
    res->stable_iters = L.stable;
    res->nll_increase_count = L.inc;
    res->has_initial = L.has_init ? 1 : 0;
    res->initial_nll = L.init_nll;
    res->final_nll = L.prev;
    res->final_abs_rel_change = L.abs_rel;
    res->final_rel_improvement = L.rel_impr;
    return CB200_OK;
}

struct SplitState {
    LeanRun R;
    cb200_model mo;
    int64_t m = 0, n = 0;
    int is_first = 1, is_last = 1;
};

}  // namespace

static void split_free(void *p) { delete static_cast<SplitState *>(p); }

// =====================================================================================
// a chromosome split over several GPUs: one shard per context
// =====================================================================================
extern "C" {

int cb200_split_begin(cb200_ctx *c, const cb200_model *mo_in, double nu, const float *data, const float *munc, int64_t m,
                      int64_t n, int64_t ld, const float *qscale, const float *kap, int32_t is_first, int32_t is_last) {
    DeviceGuard _dg(c ? c->device : 0);
    if (!c || !data || !munc || !kap) return fail(CB200_ERR_INVALID, "NULL argument");
    CB_TRY(check_model(mo_in));
    cb200_model mo = *mo_in;
    mo.use_lambda = 0;
    mo.use_kappa = 1;
    mo.use_qscale = qscale != nullptr;
    if (mo.state_dim != 2 || mo.F[0] != 1.0 || mo.F[2] != 0.0 || mo.F[3] != 1.0)
        return fail(CB200_ERR_UNSUPPORTED, "split chromosomes run the 2-state model with F = [[1, f], [0, 1]]");
    if (apn_live(&mo)) return fail(CB200_ERR_UNSUPPORTED, "adaptive process noise cannot run on a split chromosome");
    if (n < 64 || m <= 0) return fail(CB200_ERR_INVALID, "a shard needs at least 64 intervals");
    {
        const double det = mo.Q0[0] * mo.Q0[3] - mo.Q0[1] * mo.Q0[2];
        if (det == 0.0) return fail(CB200_ERR_INVALID, "matrixQ0 is singular");
    }
    lean_read_env();
    split_free(c->split);
    c->split = nullptr;
    SplitState *S = new SplitState();
    S->mo = mo;
    S->m = m;
    S->n = n;
    S->is_first = is_first != 0;
    S->is_last = is_last != 0;
    const int rc = lean_setup(c, &mo, nu, data, munc, m, n, ld, qscale, kap, &S->R);
    if (rc != CB200_OK) {
        delete S;
        return rc;
    }
    S->R.fa.sh.is_first = S->R.ba.sh.is_first = S->is_first;
    S->R.fa.sh.is_last = S->R.ba.sh.is_last = S->is_last;
    S->R.ba.lag_rows = S->is_last ? (n > 1 ? n - 1 : 1) : n;
    c->split = S;
    return CB200_OK;
}

static SplitState *split_of(cb200_ctx *c) { return c ? static_cast<SplitState *>(c->split) : nullptr; }

int cb200_split_forward_compose(cb200_ctx *c, double *payload) {
    DeviceGuard _dg(c ? c->device : 0);
    SplitState *S = split_of(c);
    if (!S || !payload) return fail(CB200_ERR_INVALID, "no split in progress (cb200_split_begin) or NULL payload");
    {
        Span sp(c, FAM_COMPOSE);
        CU_TRY(lean_fwd_compose(S->R.fa, c->stream));
    }
    {
        Span sp(c, FAM_SEGSCAN);
        CU_TRY(lean_shard_payload(S->R.fa, S->R.trk[0], false, payload, c->stream));
    }
    c->launches += 2;
    return CB200_OK;
}

int cb200_split_forward_replay(cb200_ctx *c, const double *gathered, int32_t rank, int32_t world, int32_t with_nll,
                               int32_t store, int32_t set, double *sums) {
    DeviceGuard _dg(c ? c->device : 0);
    SplitState *S = split_of(c);
    if (!S || !gathered) return fail(CB200_ERR_INVALID, "no split in progress (cb200_split_begin) or NULL payloads");
    if (rank < 0 || rank >= world || set < 0 || set > 1) return fail(CB200_ERR_INVALID, "bad rank / track set");
    if ((rank == 0) != (S->is_first != 0) || (rank == world - 1) != (S->is_last != 0))
        return fail(CB200_ERR_INVALID, "rank does not match the shard's position given to cb200_split_begin");
    LeanFwdArgs fa = S->R.fa;
    fa.trk = S->R.trk[set];
    fa.want_nll = with_nll ? 1 : 0;
    fa.do_store = store ? 1 : 0;
    fa.sums = sums ? sums : static_cast<double *>(c->sums.p);
    fa.sh.gathered = gathered;
    fa.sh.rank = rank;
    fa.sh.world = world;
    fa.sh.fwd_next = S->is_last ? nullptr : gathered + (int64_t)(rank + 1) * LEAN_PAYLOAD;
    {
        Span sp(c, FAM_SEGSCAN);
        CU_TRY(lean_fwd_prefix(fa, c->stream));
    }
    {
        Span sp(c, FAM_FWD);
        CU_TRY(lean_fwd_replay(fa, c->stream));
    }
    c->launches += 2;
    return CB200_OK;
}

int cb200_split_backward_compose(cb200_ctx *c, int32_t set, double *payload) {
    DeviceGuard _dg(c ? c->device : 0);
    SplitState *S = split_of(c);
    if (!S || !payload || set < 0 || set > 1) return fail(CB200_ERR_INVALID, "no split in progress, NULL payload or bad track set");
    {
        Span sp(c, FAM_SEGSCAN);
        CU_TRY(lean_shard_payload(S->R.fa, S->R.trk[set], true, payload, c->stream));
    }
    c->launches += 1;
    return CB200_OK;
}

int cb200_split_backward_replay(cb200_ctx *c, const double *gathered_bwd, const double *gathered_fwd, int32_t rank,
                                int32_t world, int32_t set, int32_t publish, float *xs, float *Ps, float *lag) {
    DeviceGuard _dg(c ? c->device : 0);
    SplitState *S = split_of(c);
    if (!S || !gathered_bwd || !gathered_fwd) return fail(CB200_ERR_INVALID, "no split in progress or NULL payloads");
    if (rank < 0 || rank >= world || set < 0 || set > 1) return fail(CB200_ERR_INVALID, "bad rank / track set");
    if (publish && (!xs || !Ps || !lag)) return fail(CB200_ERR_INVALID, "a publishing pass needs xs, Ps and lag");
    LeanBwdArgs ba = S->R.ba;
    ba.trk = S->R.trk[set];
    ba.xs = xs; ba.Ps = Ps; ba.lag = lag;
    ba.sh.gathered = gathered_bwd;
    ba.sh.rank = rank;
    ba.sh.world = world;
    ba.sh.fwd_next = S->is_last ? nullptr : gathered_fwd + (int64_t)(rank + 1) * LEAN_PAYLOAD;
    ba.sh.bwd_prev = S->is_first ? nullptr : gathered_bwd + (int64_t)(rank - 1) * LEAN_PAYLOAD;
    if (!publish) {
        // a publishing pass re-uses the suffix states of the kappa-carrying pass before it (same tracks)
        Span sp(c, FAM_SEGSCAN);
        CU_TRY(lean_bwd_suffix(ba, c->stream));
        c->launches += 1;
    }
    {
        Span sp(c, publish ? FAM_PUBLISH : FAM_BWD);
        CU_TRY(lean_bwd_replay(ba, publish != 0, c->stream));
    }
    c->launches += 1;
    return CB200_OK;
}

int cb200_split_end(cb200_ctx *c, float *kap) {
    DeviceGuard _dg(c ? c->device : 0);
    SplitState *S = split_of(c);
    if (!S) return fail(CB200_ERR_INVALID, "no split in progress (cb200_split_begin)");
    if (kap) {
        Span sp(c, FAM_PREC);
        CU_TRY(lean_scatter_f32(S->R.kap_rm, kap, S->R.g, c->stream));
        c->launches += 1;
    }
    CU_TRY(cudaStreamSynchronize(c->stream));
    split_free(c->split);
    c->split = nullptr;
    return CB200_OK;
}

}  // extern "C"

namespace {
}  // namespace

// =====================================================================================
// ECM (cconsenrich.pyx:7877-8442, 7188-7657)
// =====================================================================================
extern "C" {
int cb200_ecm_device(cb200_ctx *c, const cb200_model *mo_in, const cb200_ecm_opts *op, const float *data,
                     const float *munc, int64_t m, int64_t n, int64_t ld, const float *qscale, float *lam, float *kap,
                     float *xs, float *Ps, float *lag, float *resid, cb200_ecm_result *res, double *nll_path) {
    DeviceGuard _dg(c ? c->device : 0);
    if (!c || !op || !res || !data || !munc || !xs || !Ps || !lag) return fail(CB200_ERR_INVALID, "NULL argument");
    CB_TRY(check_model(mo_in));
    memset(res, 0, sizeof(*res));
    if (n <= 0 || m <= 0) {
        res->skipped = 1;
        return CB200_OK;
    }
    cb200_model mo = *mo_in;
    const int d = mo.state_dim;
    if (d == 2) {
        const double det = mo.Q0[0] * mo.Q0[3] - mo.Q0[1] * mo.Q0[2];
        if (det == 0.0) return fail(CB200_ERR_INVALID, "matrixQ0 is singular");
    }
    mo.use_lambda = lam != nullptr;
    mo.use_kappa = kap != nullptr;
    mo.use_qscale = qscale != nullptr;
    mo.store_nll_in_d = 0;
    const int64_t stride = round_up(n, 32);
    CB_TRY(ensure(c, c->stats, (size_t)stride * 4 * 8));
    CB_TRY(ensure(c, c->xf, (size_t)n * d * 4));
    CB_TRY(ensure(c, c->Pf, (size_t)n * d * d * 4));
    CB_TRY(ensure(c, c->Qf, (size_t)n * d * d * 4));
    CB_TRY(ensure(c, c->xf2, (size_t)n * d * 4));
    CB_TRY(ensure(c, c->Pf2, (size_t)n * d * d * 4));
    CB_TRY(ensure(c, c->Qf2, (size_t)n * d * d * 4));
    double *stats = static_cast<double *>(c->stats.p);
    // The inner sweeps of the CLI-default configuration (2-state model, F = [[1, f], [0, 1]], kappa the
    // only multiplier being fitted) run on run-major private tracks: lean_kernels.cuh.
    if (lean_eligible(&mo, op, n, lam, kap))
        return ecm_device_lean(c, &mo, op, data, munc, m, n, ld, qscale, kap, xs, Ps, lag, resid, res, nll_path);
    // two sets of forward tracks: the NLL pass that closes an iteration is, when another iteration
    // follows, the storing forward pass that iteration would open with (same multipliers, same
    // arithmetic), run ahead of time into the spare set
    float *xf = static_cast<float *>(c->xf.p), *Pf = static_cast<float *>(c->Pf.p), *Qf = static_cast<float *>(c->Qf.p);
    float *xf2 = static_cast<float *>(c->xf2.p), *Pf2 = static_cast<float *>(c->Pf2.p), *Qf2 = static_cast<float *>(c->Qf2.p);
    double *sums = static_cast<double *>(c->sums.p);
    CB_TRY(do_fold(c, data, munc, m, n, ld, mo.pad, stats, stride));

    // 2-state model: the forward replay composes the smoother's run elements, so the backward scan
    // needs no first pass.  Both scans then share one partition into runs.  Not when the runs are
    // shorter than the head replay (tiny tracks): the head rewrites bins other threads composed from.
    SmoFuse sfa, sfb;  // current / spare set, like the forward tracks
    {
        const int ns = scan_pick_nsub(n, 0);
        if (d == 2 && CHUNK * ns >= HEAD_BINS && !apn_live(&mo)) {
            const int64_t tiles = scan_num_tiles(n, ns);
            const int64_t runs = tiles * SCAN_THREADS;
            const int64_t pitch = round_up(runs, 32);
            CB_TRY(ensure(c, c->smo, (size_t)pitch * 9 * 8));
            CB_TRY(ensure(c, c->smo2, (size_t)pitch * 9 * 8));
            sfa.run = static_cast<double *>(c->smo.p);
            sfb.run = static_cast<double *>(c->smo2.p);
            sfa.pitch = sfb.pitch = pitch;
            sfa.nsub = sfb.nsub = ns;
            sfa.npad = sfb.npad = tiles * TILE_BINS * ns;
        }
    }

    // The smoothed tracks of an inner sweep are read by the multiplier updates only.  When kappa is
    // the only one (the CLI default) and rides on the backward replay, the replay does not store them
    // at all; one plain backward pass after the loop writes the tracks the call returns (it reads the
    // forward tracks of the last sweep, which the kappa update does not touch).
    const bool lean = kap != nullptr && lam == nullptr;
    auto forward_store = [&](float *oxf, float *oPf, float *oQf, bool with_nll) -> int {
        cb200_model f = mo;
        f.return_nll = with_nll ? 1 : 0;
        return do_forward(c, &f, stats, stride, m, n, lam, kap, qscale, nullptr, oxf, oPf, oQf, nullptr,
                          with_nll ? sums : nullptr, nullptr, false, nullptr, oxf == xf ? &sfa : &sfb);
    };
    // with_kappa: the kappa update of this inner iteration rides on the backward replay
    auto backward = [&](bool with_kappa) -> int {
        KappaFuse kf;
        if (with_kappa) {
            kf.kap_out = kap;
            kf.qs = qscale;
            kf.nu = op->nu;
            kf.no_store = lean;
        }
        return do_backward(c, &mo, n, xf, Pf, Qf, nullptr, 1, xs, Ps, lag, n > 1 ? n - 1 : 1, nullptr, false, &kf, &sfa);
    };
    auto sweep = [&](bool with_kappa) -> int {
        CB_TRY(forward_store(xf, Pf, Qf, false));
        return backward(with_kappa);
    };
    auto nll_only = [&](double *out) -> int {
        cb200_model f = mo;
        f.return_nll = 1;
        CB_TRY(do_forward(c, &f, stats, stride, m, n, lam, kap, qscale, nullptr, nullptr, nullptr, nullptr, nullptr,
                          sums, nullptr, false));
        double s[2];
        CB_TRY(read_sums(c, s));
        *out = s[1];
        return CB200_OK;
    };

    const int patience = 2;
    if (n <= 5) {  // pyx:7998-8129: filter + smoother only
        double cur = 0.0;
        CB_TRY(sweep(false));
        CB_TRY(nll_only(&cur));
        res->skipped = 1;
        res->initial_nll = res->final_nll = cur;
        if (resid) CB_TRY(do_residuals(c, data, m, n, ld, xs, d, resid));
        CU_TRY(cudaStreamSynchronize(c->stream));
        return CB200_OK;
    }

    double prev = 1.0e16, cur = 0.0, init_nll = 0.0, rel_impr = 0.0, abs_rel = 0.0;
    bool has_init = false, converged = false;
    int iters_done = 0, stable = 0, inc = 0;
    bool opened = false;  // the forward pass of the next sweep has already run (into xf, Pf, Qf)
    for (int i = 0; i < op->max_iters; ++i) {
        iters_done = i + 1;
        for (int t = 0; t < op->inner_iters; ++t) {
            // kappa is read by the forward pass only, so the backward pass may overwrite it in place; the
            // lambda update (separate kernel) reads the smoothed tracks, not kappa
            if (opened) {
                CB_TRY(backward(kap != nullptr));
                opened = false;
            } else {
                CB_TRY(sweep(kap != nullptr));
            }
            if (lam) {
                Span sp(c, FAM_PREC);
                CU_TRY(launch_update_lambda(reinterpret_cast<const double2 *>(stats),
                                            reinterpret_cast<const double2 *>(stats + 2 * stride), n, (double)m, xs, Ps,
                                            d, op->nu, mo.lam_min, mo.lam_max, lam, c->stream));
                c->launches += 1;
            }
        }
        const bool ahead = i + 1 < op->max_iters && op->inner_iters > 0;
        if (ahead) {  // NLL of this iteration = by-product of the next iteration's opening forward pass
            CB_TRY(forward_store(xf2, Pf2, Qf2, true));
            double s2[2];
            CB_TRY(read_sums(c, s2));
            cur = s2[1];
        } else {
            CB_TRY(nll_only(&cur));
        }
        if (nll_path) nll_path[i] = cur;
        const bool has_prev = has_init;
        if (!has_prev) {
            init_nll = cur;
            has_init = true;
        } else if (cur > prev + (1.0e-12 * fmax(fabs(prev), 1.0))) {
            inc += 1;
        }
        double delta, scale;
        if (has_prev) {
            delta = fabs(cur - prev);
            scale = fabs(prev);
        } else {
            delta = 0.0;
            scale = fabs(cur);
        }
        scale = fmax(scale, fabs(cur));
        scale = fmax(scale, 1.0);
        if (has_prev) {
            rel_impr = (prev - cur) / scale;
            abs_rel = delta / scale;
        } else {
            rel_impr = abs_rel = 0.0;
        }
        const double tol = op->rtol * scale;
        prev = cur;
        stable = (has_prev && delta <= tol) ? stable + 1 : 0;
        if (stable >= patience) {
            converged = true;
            break;  // the tracks run ahead into the spare set are dropped
        }
        if (ahead) {  // the next iteration continues from the tracks just written
            std::swap(xf, xf2);
            std::swap(Pf, Pf2);
            std::swap(Qf, Qf2);
            std::swap(sfa, sfb);
            opened = true;
        }
    }
    if (lean && iters_done > 0 && op->inner_iters > 0)
        CB_TRY(do_backward(c, &mo, n, xf, Pf, Qf, nullptr, 1, xs, Ps, lag, n > 1 ? n - 1 : 1, nullptr, false, nullptr, &sfa));
    if (resid) CB_TRY(do_residuals(c, data, m, n, ld, xs, d, resid));
    CU_TRY(cudaStreamSynchronize(c->stream));
    res->iters_done = iters_done;
    res->converged = converged ? 1 : 0;
    res->skipped = 0;
    res->stable_iters = stable;
    res->nll_increase_count = inc;
    res->has_initial = has_init ? 1 : 0;
    res->initial_nll = init_nll;
    res->final_nll = prev;
    res->final_abs_rel_change = abs_rel;
    res->final_rel_improvement = rel_impr;
    return CB200_OK;
}

// =====================================================================================
// reference-facing path (host buffers)
// =====================================================================================
int cb200_host_sweep(cb200_ctx *c, const cb200_model *mo, const float *data, const float *munc, int64_t m, int64_t n,
                     const float *lam, const float *kap, const float *qscale, float *xf, float *Pf, float *Qf, float *D,
                     double *sum_d, double *sum_nll, float *xs, float *Ps, float *lag, int64_t lag_rows, float *resid) {
    DeviceGuard _dg(c ? c->device : 0);
    if (!c || !data || !munc) return fail(CB200_ERR_INVALID, "NULL argument");
    CB_TRY(check_model(mo));
    if (sum_d) *sum_d = 0.0;
    if (sum_nll) *sum_nll = 0.0;
    if (n <= 0 || m <= 0) return CB200_OK;
    const int d = mo->state_dim;
    const bool want_bwd = xs || Ps || lag || resid;
    const bool store = xf || Pf || Qf || want_bwd;
    int64_t ld = 0, ld2 = 0;
    CB_TRY(upload_tracks(c, c->data, data, m, n, &ld));
    CB_TRY(upload_tracks(c, c->munc, munc, m, n, &ld2));
    const float *dlam, *dkap, *dqs;
    CB_TRY(upload_vec(c, c->lam, lam, n, mo->use_lambda != 0, &dlam));
    CB_TRY(upload_vec(c, c->kap, kap, n, mo->use_kappa != 0, &dkap));
    CB_TRY(upload_vec(c, c->qs, qscale, n, mo->use_qscale != 0, &dqs));
    const int64_t stride = round_up(n, 32);
    CB_TRY(ensure(c, c->stats, (size_t)stride * 4 * 8));
    double *stats = static_cast<double *>(c->stats.p);
    CB_TRY(do_fold(c, static_cast<const float *>(c->data.p), static_cast<const float *>(c->munc.p), m, n, ld, mo->pad,
                   stats, stride));
    float *dxf = nullptr, *dPf = nullptr, *dQf = nullptr;
    if (store) {
        CB_TRY(ensure(c, c->xf, (size_t)n * d * 4));
        CB_TRY(ensure(c, c->Pf, (size_t)n * d * d * 4));
        CB_TRY(ensure(c, c->Qf, (size_t)n * d * d * 4));
        dxf = static_cast<float *>(c->xf.p);
        dPf = static_cast<float *>(c->Pf.p);
        dQf = static_cast<float *>(c->Qf.p);
    }
    CB_TRY(ensure(c, c->D, (size_t)n * 4));
    CB_TRY(do_forward(c, mo, stats, stride, m, n, dlam, dkap, dqs, nullptr, dxf, dPf, dQf, static_cast<float *>(c->D.p),
                      static_cast<double *>(c->sums.p), nullptr, false));
    if (xf) CB_TRY(d2h(c, xf, dxf, (size_t)n * d * 4));
    if (Pf) CB_TRY(d2h(c, Pf, dPf, (size_t)n * d * d * 4));
    if (Qf && n > 1) CB_TRY(d2h(c, Qf, dQf, (size_t)(n - 1) * d * d * 4));
    if (D) CB_TRY(d2h(c, D, c->D.p, (size_t)n * 4));
    if (want_bwd) {
        CB_TRY(ensure(c, c->xs, (size_t)n * d * 4));
        CB_TRY(ensure(c, c->Ps, (size_t)n * d * d * 4));
        CB_TRY(ensure(c, c->lag, (size_t)n * d * d * 4));
        const int64_t rows = n > 1 ? n - 1 : 1;
        CB_TRY(do_backward(c, mo, n, dxf, dPf, dQf, nullptr, 1, static_cast<float *>(c->xs.p),
                           static_cast<float *>(c->Ps.p), static_cast<float *>(c->lag.p), rows, nullptr, false));
        if (xs) CB_TRY(d2h(c, xs, c->xs.p, (size_t)n * d * 4));
        if (Ps) CB_TRY(d2h(c, Ps, c->Ps.p, (size_t)n * d * d * 4));
        if (lag && n > 1) {
            const int64_t r = lag_rows < n - 1 ? lag_rows : n - 1;
            CB_TRY(d2h(c, lag, c->lag.p, (size_t)r * d * d * 4));
        }
        if (resid) {
            CB_TRY(ensure(c, c->resid, (size_t)n * m * 4));
            CB_TRY(do_residuals(c, static_cast<const float *>(c->data.p), m, n, ld, static_cast<float *>(c->xs.p), d,
                                static_cast<float *>(c->resid.p)));
            CB_TRY(d2h(c, resid, c->resid.p, (size_t)n * m * 4));
        }
    }
    double s[2];
    CB_TRY(read_sums(c, s));
    if (sum_d) *sum_d = s[0];
    if (sum_nll) *sum_nll = s[1];
    return CB200_OK;
}

int cb200_host_forward_pass(cb200_ctx *c, const cb200_model *mo, const float *data, const float *munc, int64_t m,
                            int64_t n, const int32_t *block_map, int64_t block_count, const float *lam,
                            const float *kap, const float *qscale, float *xf, float *Pf, float *Qf, float *D,
                            double *sum_d, double *sum_nll) {
    DeviceGuard _dg(c ? c->device : 0);
    if (n > 0 && m > 0) CB_TRY(check_block_map(block_map, n, block_count));
    return cb200_host_sweep(c, mo, data, munc, m, n, lam, kap, qscale, xf, Pf, Qf, D, sum_d, sum_nll, nullptr, nullptr,
                            nullptr, 0, nullptr);
}

int cb200_host_backward_pass(cb200_ctx *c, const cb200_model *mo, const float *data, int64_t m, int64_t n,
                             const float *xf, const float *Pf, const float *Qf, float *xs, float *Ps, float *lag,
                             int64_t lag_rows, float *resid) {
    DeviceGuard _dg(c ? c->device : 0);
    if (!c || !xf || !Pf || !Qf || !xs || !Ps || !lag) return fail(CB200_ERR_INVALID, "NULL argument");
    CB_TRY(check_model(mo));
    if (n <= 0) return CB200_OK;
    const int d = mo->state_dim;
    CB_TRY(ensure(c, c->xf, (size_t)n * d * 4));
    CB_TRY(ensure(c, c->Pf, (size_t)n * d * d * 4));
    CB_TRY(ensure(c, c->Qf, (size_t)n * d * d * 4));
    CB_TRY(ensure(c, c->xs, (size_t)n * d * 4));
    CB_TRY(ensure(c, c->Ps, (size_t)n * d * d * 4));
    CB_TRY(ensure(c, c->lag, (size_t)n * d * d * 4));
    CB_TRY(h2d(c, c->xf.p, xf, (size_t)n * d * 4));
    CB_TRY(h2d(c, c->Pf.p, Pf, (size_t)n * d * d * 4));
    if (n > 1) CB_TRY(h2d(c, c->Qf.p, Qf, (size_t)(n - 1) * d * d * 4));
    const int64_t rows = n > 1 ? n - 1 : 1;
    CB_TRY(do_backward(c, mo, n, static_cast<float *>(c->xf.p), static_cast<float *>(c->Pf.p),
                       static_cast<float *>(c->Qf.p), nullptr, 1, static_cast<float *>(c->xs.p),
                       static_cast<float *>(c->Ps.p), static_cast<float *>(c->lag.p), rows, nullptr, false));
    CB_TRY(d2h(c, xs, c->xs.p, (size_t)n * d * 4));
    CB_TRY(d2h(c, Ps, c->Ps.p, (size_t)n * d * d * 4));
    if (n > 1) {
        const int64_t r = lag_rows < n - 1 ? lag_rows : n - 1;
        CB_TRY(d2h(c, lag, c->lag.p, (size_t)r * d * d * 4));
    }
    if (resid && m > 0) {
        if (!data) return fail(CB200_ERR_INVALID, "matrixData is required for residuals");
        int64_t ld = 0;
        CB_TRY(upload_tracks(c, c->data, data, m, n, &ld));
        CB_TRY(ensure(c, c->resid, (size_t)n * m * 4));
        CB_TRY(do_residuals(c, static_cast<const float *>(c->data.p), m, n, ld, static_cast<float *>(c->xs.p), d,
                            static_cast<float *>(c->resid.p)));
        CB_TRY(d2h(c, resid, c->resid.p, (size_t)n * m * 4));
    }
    CU_TRY(cudaStreamSynchronize(c->stream));
    return CB200_OK;
}

int cb200_host_ecm(cb200_ctx *c, const cb200_model *mo, const cb200_ecm_opts *op, const float *data,
                   const float *munc, int64_t m, int64_t n, const int32_t *block_map, int64_t block_count,
                   const float *qscale, float *lam, float *kap, float *xs, float *Ps, float *lag, float *resid,
                   cb200_ecm_result *res, double *nll_path) {
    DeviceGuard _dg(c ? c->device : 0);
    if (!c || !op || !res || !data || !munc) return fail(CB200_ERR_INVALID, "NULL argument");
    CB_TRY(check_model(mo));
    memset(res, 0, sizeof(*res));
    if (n <= 0 || m <= 0) {
        res->skipped = 1;
        return CB200_OK;
    }
    CB_TRY(check_block_map(block_map, n, block_count));
    const int d = mo->state_dim;
    int64_t ld = 0, ld2 = 0;
    CB_TRY(upload_tracks(c, c->data, data, m, n, &ld));
    CB_TRY(upload_tracks(c, c->munc, munc, m, n, &ld2));
    const float *dqs, *dlam_c, *dkap_c;
    CB_TRY(upload_vec(c, c->qs, qscale, n, qscale != nullptr, &dqs));
    if (lam && (op->init_ones & 1)) CB_TRY(ones_vec(c, c->lam, n, &dlam_c));
    else CB_TRY(upload_vec(c, c->lam, lam, n, lam != nullptr, &dlam_c));
    if (kap && (op->init_ones & 2)) CB_TRY(ones_vec(c, c->kap, n, &dkap_c));
    else CB_TRY(upload_vec(c, c->kap, kap, n, kap != nullptr, &dkap_c));
    CB_TRY(ensure(c, c->xs, (size_t)n * d * 4));
    CB_TRY(ensure(c, c->Ps, (size_t)n * d * d * 4));
    CB_TRY(ensure(c, c->lag, (size_t)n * d * d * 4));
    float *dres = nullptr;
    if (resid) {
        CB_TRY(ensure(c, c->resid, (size_t)n * m * 4));
        dres = static_cast<float *>(c->resid.p);
    }
    CB_TRY(cb200_ecm_device(c, mo, op, static_cast<const float *>(c->data.p), static_cast<const float *>(c->munc.p), m,
                            n, ld, dqs, const_cast<float *>(dlam_c), const_cast<float *>(dkap_c),
                            static_cast<float *>(c->xs.p), static_cast<float *>(c->Ps.p),
                            static_cast<float *>(c->lag.p), dres, res, nll_path));
    if (xs) CB_TRY(d2h(c, xs, c->xs.p, (size_t)n * d * 4));
    if (Ps) CB_TRY(d2h(c, Ps, c->Ps.p, (size_t)n * d * d * 4));
    if (lag && n > 1) CB_TRY(d2h(c, lag, c->lag.p, (size_t)(n - 1) * d * d * 4));
    if (resid) CB_TRY(d2h(c, resid, dres, (size_t)n * m * 4));
    if (lam) CB_TRY(d2h(c, lam, c->lam.p, (size_t)n * 4));
    if (kap) CB_TRY(d2h(c, kap, c->kap.p, (size_t)n * 4));
    CU_TRY(cudaStreamSynchronize(c->stream));
    return CB200_OK;
}

// ---- background track ----------------------------------------------------------------
int cb200_background_stats(cb200_ctx *c, const float *resid, const float *inv, int64_t m, int64_t n, int64_t ld,
                           double *weight, double *rhs, unsigned long long *support) {
    DeviceGuard _dg(c ? c->device : 0);
    if (!c || !resid || !inv || !weight || !rhs) return fail(CB200_ERR_INVALID, "NULL argument");
    if (m < 0 || n < 0 || ld < n) return fail(CB200_ERR_INVALID, "bad matrix shape");
    if (n == 0) return CB200_OK;
    Span sp(c, FAM_BG);
    CU_TRY(launch_background_stats(resid, inv, m, n, ld, weight, rhs, support, c->stream));
    c->launches += 1;
    return CB200_OK;
}

static int check_background_args(int64_t n, double lam, double lam_first) {
    // messages of cconsenrich.pyx:985-990
    if (!std::isfinite(lam_first) || lam_first < 0.0) return fail(CB200_ERR_INVALID, "lamFirst must be finite and nonnegative");
    if (!std::isfinite(lam) || lam < 0.0) return fail(CB200_ERR_INVALID, "lam must be finite and nonnegative");
    if (n < 0) return fail(CB200_ERR_INVALID, "n must be nonnegative");
    return CB200_OK;
}

int cb200_background_solve(cb200_ctx *c, const double *weight, const double *rhs, int64_t n, double lam,
                           double lam_first, int32_t zero_center, double *out, int64_t *bad_index, double *bad_value) {
    DeviceGuard _dg(c ? c->device : 0);
    if (!c || !weight || !rhs || !out) return fail(CB200_ERR_INVALID, "NULL argument");
    CB_TRY(check_background_args(n, lam, lam_first));
    if (bad_index) *bad_index = -1;
    if (bad_value) *bad_value = 0.0;
    if (n == 0) return CB200_OK;
    if (n == 1) return fail(CB200_ERR_INVALID, "a single interval is solved by the host entry point");
    CB_TRY(ensure(c, c->bg_ws, background_workspace_bytes(n)));
    CB_TRY(ensure(c, c->bg_status, sizeof(BackgroundStatus)));
    int launches = 0;
    {
        Span sp(c, FAM_BG);
        CU_TRY(launch_background_solve(weight, rhs, n, lam, lam_first, zero_center, out, c->bg_ws.p,
                                       static_cast<BackgroundStatus *>(c->bg_status.p), c->stream, &launches));
    }
    c->launches += launches;
    if (bad_index || bad_value) {
        BackgroundStatus st;
        CU_TRY(cudaMemcpyAsync(&st, c->bg_status.p, sizeof(st), cudaMemcpyDeviceToHost, c->stream));
        CU_TRY(cudaStreamSynchronize(c->stream));
        if (st.bad_index != STATUS_NONE) {
            if (bad_index) *bad_index = st.bad_index;
            if (bad_value) *bad_value = st.bad_value;
        }
    }
    return CB200_OK;
}

int cb200_host_background_stats(cb200_ctx *c, const float *resid, const float *inv, int64_t m, int64_t n,
                                double *weight, double *rhs, int64_t *support) {
    DeviceGuard _dg(c ? c->device : 0);
    if (!c || !weight || !rhs) return fail(CB200_ERR_INVALID, "NULL argument");
    if (support) *support = 0;
    if (n <= 0) return CB200_OK;
    if (m <= 0) {  // empty sums
        memset(weight, 0, (size_t)n * 8);
        memset(rhs, 0, (size_t)n * 8);
        return CB200_OK;
    }
    if (!resid || !inv) return fail(CB200_ERR_INVALID, "NULL argument");
    int64_t ld = 0, ld2 = 0;
    CB_TRY(upload_tracks(c, c->data, resid, m, n, &ld));
    CB_TRY(upload_tracks(c, c->munc, inv, m, n, &ld2));
    CB_TRY(ensure(c, c->bg_w, (size_t)n * 8));
    CB_TRY(ensure(c, c->bg_rhs, (size_t)n * 8));
    CB_TRY(ensure(c, c->bg_status, sizeof(BackgroundStatus)));
    unsigned long long *dsup = support ? reinterpret_cast<unsigned long long *>(c->bg_status.p) : nullptr;
    CB_TRY(cb200_background_stats(c, static_cast<const float *>(c->data.p), static_cast<const float *>(c->munc.p), m, n, ld,
                                  static_cast<double *>(c->bg_w.p), static_cast<double *>(c->bg_rhs.p), dsup));
    CB_TRY(d2h(c, weight, c->bg_w.p, (size_t)n * 8));
    CB_TRY(d2h(c, rhs, c->bg_rhs.p, (size_t)n * 8));
    unsigned long long sup = 0;
    if (support) CB_TRY(d2h(c, &sup, dsup, sizeof(sup)));
    CU_TRY(cudaStreamSynchronize(c->stream));
    if (support) *support = (int64_t)sup;
    return CB200_OK;
}

int cb200_host_background_solve(cb200_ctx *c, const double *weight, const double *rhs, int64_t n, double lam,
                                double lam_first, int32_t zero_center, double *out, int64_t *bad_index,
                                double *bad_value) {
    DeviceGuard _dg(c ? c->device : 0);
    if (!c) return fail(CB200_ERR_INVALID, "ctx is NULL");
    CB_TRY(check_background_args(n, lam, lam_first));
    if (bad_index) *bad_index = -1;
    if (bad_value) *bad_value = 0.0;
    if (n == 0) return CB200_OK;
    if (!weight || !rhs || !out) return fail(CB200_ERR_INVALID, "NULL argument");
    if (n == 1) {  // cconsenrich.pyx:995-1006
        out[0] = 0.0;
        if (!zero_center) {
            if (weight[0] < 1.0e-12) {
                if (bad_index) *bad_index = 0;
                if (bad_value) *bad_value = weight[0];
            } else {
                out[0] = rhs[0] / weight[0];
            }
        }
        return CB200_OK;
    }
    CB_TRY(ensure(c, c->bg_w, (size_t)n * 8));
    CB_TRY(ensure(c, c->bg_rhs, (size_t)n * 8));
    CB_TRY(ensure(c, c->bg_out, (size_t)n * 8));
    CB_TRY(h2d(c, c->bg_w.p, weight, (size_t)n * 8));
    CB_TRY(h2d(c, c->bg_rhs.p, rhs, (size_t)n * 8));
    int64_t bi = -1;
    double bv = 0.0;
    CB_TRY(cb200_background_solve(c, static_cast<const double *>(c->bg_w.p), static_cast<const double *>(c->bg_rhs.p), n,
                                  lam, lam_first, zero_center, static_cast<double *>(c->bg_out.p), &bi, &bv));
    if (bad_index) *bad_index = bi;
    if (bad_value) *bad_value = bv;
    CB_TRY(d2h(c, out, c->bg_out.p, (size_t)n * 8));
    CU_TRY(cudaStreamSynchronize(c->stream));
    return CB200_OK;
}

// ---- driver-side reductions -------------------------------------------------------------
int cb200_weighted_mean_residual(cb200_ctx *c, const float *data, const float *munc, int64_t m, int64_t n, int64_t ld,
                                 const double *state, const double *background, double pad, double *out) {
    DeviceGuard _dg(c ? c->device : 0);
    if (!c) return fail(CB200_ERR_INVALID, "ctx is NULL");
    if (m < 0 || n < 0 || ld < n) return fail(CB200_ERR_INVALID, "bad matrix shape");
    if (n == 0) return CB200_OK;
    if (!state || !out || (m > 0 && (!data || !munc))) return fail(CB200_ERR_INVALID, "NULL argument");
    Span sp(c, FAM_BG);
    CU_TRY(launch_weighted_mean_residual(data, munc, m, n, ld, state, background, pad, out, c->stream));
    c->launches += 1;
    return CB200_OK;
}

int cb200_host_weighted_mean_residual(cb200_ctx *c, const float *data, const float *munc, int64_t m, int64_t n,
                                      const double *state, const double *background, double pad, double *out) {
    DeviceGuard _dg(c ? c->device : 0);
    if (!c) return fail(CB200_ERR_INVALID, "ctx is NULL");
    if (n <= 0) return n < 0 ? fail(CB200_ERR_INVALID, "bad matrix shape") : CB200_OK;
    if (!state || !out || (m > 0 && (!data || !munc))) return fail(CB200_ERR_INVALID, "NULL argument");
    int64_t ld = round_up(n, 32), t = 0;
    if (m > 0) {
        CB_TRY(upload_tracks(c, c->data, data, m, n, &t));
        CB_TRY(upload_tracks(c, c->munc, munc, m, n, &t));
    }
    CB_TRY(ensure(c, c->bg_w, (size_t)n * 8));
    CB_TRY(ensure(c, c->bg_rhs, (size_t)n * 8));
    CB_TRY(ensure(c, c->bg_out, (size_t)n * 8));
    CB_TRY(h2d(c, c->bg_w.p, state, (size_t)n * 8));
    if (background) CB_TRY(h2d(c, c->bg_rhs.p, background, (size_t)n * 8));
    CB_TRY(cb200_weighted_mean_residual(c, static_cast<const float *>(c->data.p), static_cast<const float *>(c->munc.p), m, n,
                                        ld, static_cast<const double *>(c->bg_w.p),
                                        background ? static_cast<const double *>(c->bg_rhs.p) : nullptr, pad,
                                        static_cast<double *>(c->bg_out.p)));
    CB_TRY(d2h(c, out, c->bg_out.p, (size_t)n * 8));
    CU_TRY(cudaStreamSynchronize(c->stream));
    return CB200_OK;
}

int cb200_diag_obs_sums(cb200_ctx *c, const float *munc, int64_t m, int64_t n, int64_t ld, const double *obs_prec,
                        double pad, double *munc_trace, double *sum_inv_r) {
    DeviceGuard _dg(c ? c->device : 0);
    if (!c) return fail(CB200_ERR_INVALID, "ctx is NULL");
    if (m < 0 || n < 0 || ld < n) return fail(CB200_ERR_INVALID, "bad matrix shape");
    if (n == 0) return CB200_OK;
    if (!obs_prec || !munc_trace || !sum_inv_r || (m > 0 && !munc)) return fail(CB200_ERR_INVALID, "NULL argument");
    Span sp(c, FAM_BG);
    CU_TRY(launch_diag_obs_sums(munc, m, n, ld, obs_prec, pad, munc_trace, sum_inv_r, c->stream));
    c->launches += 1;
    return CB200_OK;
}

static int to_diag_args(const cb200_diag_gain_args *a, DiagGainArgs *k) {
    if (a->n < 0 || (a->dim != 1 && a->dim != 2) || a->cov_dim < a->dim)
        return fail(CB200_ERR_INVALID, "stateCovarForward shape does not match stateModel");
    k->covar = a->covar; k->p_noise = a->p_noise; k->q_scale = a->q_scale; k->proc_prec = a->proc_prec;
    k->sum_inv_r = a->sum_inv_r; k->sum_gain0 = a->sum_gain0; k->sum_gain1 = a->sum_gain1;
    k->n = a->n; k->dim = a->dim; k->cov_dim = a->cov_dim; k->cov_init = a->cov_init;
    for (int i = 0; i < 4; ++i) {
        k->base_q[i] = a->base_q[i];
        k->f[i] = a->f[i];
    }
    return CB200_OK;
}

int cb200_diag_gain(cb200_ctx *c, const cb200_diag_gain_args *a) {
    DeviceGuard _dg(c ? c->device : 0);
    if (!c || !a) return fail(CB200_ERR_INVALID, "NULL argument");
    DiagGainArgs k{};
    CB_TRY(to_diag_args(a, &k));
    if (a->n == 0) return CB200_OK;
    if (!a->covar || !a->q_scale || !a->sum_inv_r || !a->sum_gain0 || !a->sum_gain1)
        return fail(CB200_ERR_INVALID, "NULL argument");
    Span sp(c, FAM_BG);
    CU_TRY(launch_diag_gain(k, c->stream));
    c->launches += 1;
    return CB200_OK;
}

int cb200_host_interval_diagnostics(cb200_ctx *c, const float *munc, int64_t m, const double *obs_prec, double pad,
                                    const cb200_diag_gain_args *a, double *munc_trace, double *sum_inv_r) {
    DeviceGuard _dg(c ? c->device : 0);
    if (!c || !a) return fail(CB200_ERR_INVALID, "NULL argument");
    DiagGainArgs k{};
    CB_TRY(to_diag_args(a, &k));
    const int64_t n = a->n;
    if (n == 0) return CB200_OK;
    if (!obs_prec || !munc_trace || !sum_inv_r || (m > 0 && !munc) || !a->covar || !a->q_scale || !a->sum_gain0 ||
        !a->sum_gain1)
        return fail(CB200_ERR_INVALID, "NULL argument");
    int64_t ld = round_up(n, 32);
    if (m > 0) CB_TRY(upload_tracks(c, c->munc, munc, m, n, &ld));
    const size_t cov_bytes = (size_t)n * a->cov_dim * a->cov_dim * 4;
    CB_TRY(ensure(c, c->Pf, cov_bytes));
    CB_TRY(h2d(c, c->Pf.p, a->covar, cov_bytes));
    k.covar = static_cast<const float *>(c->Pf.p);
    if (a->p_noise) {
        CB_TRY(ensure(c, c->Qf, cov_bytes));
        CB_TRY(h2d(c, c->Qf.p, a->p_noise, cov_bytes));
        k.p_noise = static_cast<const float *>(c->Qf.p);
    }
    // double vectors: obs_prec, q_scale, proc_prec in; munc_trace, sum_inv_r, gains out
    DevBuf *vec[7] = {&c->bg_w, &c->bg_rhs, &c->bg_out, &c->seed_vec[0], &c->seed_vec[1], &c->seed_vec[2], &c->seed_vec[3]};
    for (DevBuf *b : vec) CB_TRY(ensure(c, *b, (size_t)n * 8));
    double *d_obs = static_cast<double *>(vec[0]->p), *d_qs = static_cast<double *>(vec[1]->p);
    double *d_pp = static_cast<double *>(vec[2]->p), *d_tr = static_cast<double *>(vec[3]->p);
    double *d_si = static_cast<double *>(vec[4]->p), *d_g0 = static_cast<double *>(vec[5]->p);
    double *d_g1 = static_cast<double *>(vec[6]->p);
    CB_TRY(h2d(c, d_obs, obs_prec, (size_t)n * 8));
    CB_TRY(h2d(c, d_qs, a->q_scale, (size_t)n * 8));
    if (a->proc_prec) CB_TRY(h2d(c, d_pp, a->proc_prec, (size_t)n * 8));
    CB_TRY(cb200_diag_obs_sums(c, static_cast<const float *>(c->munc.p), m, n, ld, d_obs, pad, d_tr, d_si));
    k.q_scale = d_qs;
    k.proc_prec = a->proc_prec ? d_pp : nullptr;
    k.sum_inv_r = d_si;
    k.sum_gain0 = d_g0;
    k.sum_gain1 = d_g1;
    {
        Span sp(c, FAM_BG);
        CU_TRY(launch_diag_gain(k, c->stream));
    }
    c->launches += 1;
    CB_TRY(d2h(c, munc_trace, d_tr, (size_t)n * 8));
    CB_TRY(d2h(c, sum_inv_r, d_si, (size_t)n * 8));
    CB_TRY(d2h(c, a->sum_gain0, d_g0, (size_t)n * 8));
    CB_TRY(d2h(c, a->sum_gain1, d_g1, (size_t)n * 8));
    CU_TRY(cudaStreamSynchronize(c->stream));
    return CB200_OK;
}

// ---- observation-noise (MUNC) stage -----------------------------------------------------
static int check_munc_smooth_args(int32_t mask_mode, int64_t window, double eps) {
    // messages of cconsenrich.pyx:5665-5668
    if (window < 1) return fail(CB200_ERR_INVALID, "windowIntervals must be positive");
    if (!(eps > 0.0) || !std::isfinite(eps)) return fail(CB200_ERR_INVALID, "eps must be positive and finite");
    if (mask_mode < 0 || mask_mode > 2) return fail(CB200_ERR_INVALID, "excludeMask must be one- or two-dimensional");
    if (window > MUNC_MAX_WINDOW)
        return fail(CB200_ERR_UNSUPPORTED, "windowIntervals above %lld is not supported on the device",
                    (long long)MUNC_MAX_WINDOW);
    return CB200_OK;
}

int cb200_munc_smooth_local_evidence(cb200_ctx *c, const float *local, const unsigned char *mask, int32_t mask_mode,
                                     int64_t m, int64_t n, int64_t ld, int64_t mask_ld, int64_t window, double eps,
                                     float *out, int64_t out_ld, int32_t *invalid) {
    DeviceGuard _dg(c ? c->device : 0);
    if (!c || !invalid) return fail(CB200_ERR_INVALID, "NULL argument");
    CB_TRY(check_munc_smooth_args(mask_mode, window, eps));
    if (m < 0 || n < 0 || ld < n || out_ld < n) return fail(CB200_ERR_INVALID, "bad matrix shape");
    if (mask_mode != 0 && !mask) return fail(CB200_ERR_INVALID, "excludeMask is NULL");
    if (m == 0 || n == 0) {
        CU_TRY(cudaMemsetAsync(invalid, 0, sizeof(int32_t), c->stream));
        return CB200_OK;
    }
    if (!local || !out) return fail(CB200_ERR_INVALID, "NULL argument");
    Span sp(c, FAM_MUNC);
    CU_TRY(launch_munc_rolling_mean(local, mask, mask_mode, m, n, ld, mask_ld, window, eps, out, out_ld,
                                    reinterpret_cast<int *>(invalid), c->stream));
    c->launches += 1;
    return CB200_OK;
}

int cb200_host_munc_smooth_local_evidence(cb200_ctx *c, const float *local, const unsigned char *mask,
                                          int32_t mask_mode, int64_t m, int64_t n, int64_t window, double eps,
                                          float *out, int32_t *invalid) {
    DeviceGuard _dg(c ? c->device : 0);
    if (!c || !invalid) return fail(CB200_ERR_INVALID, "NULL argument");
    CB_TRY(check_munc_smooth_args(mask_mode, window, eps));
    *invalid = 0;
    if (m <= 0 || n <= 0) return CB200_OK;
    if (!local || !out || (mask_mode != 0 && !mask)) return fail(CB200_ERR_INVALID, "NULL argument");
    int64_t ld = 0;
    CB_TRY(upload_tracks(c, c->data, local, m, n, &ld));
    CB_TRY(ensure(c, c->munc, (size_t)m * (size_t)ld * 4));  // output, same pitch
    const size_t mask_bytes = mask_mode == 1 ? (size_t)n : (mask_mode == 2 ? (size_t)m * (size_t)n : 0);
    if (mask_bytes) {
        CB_TRY(ensure(c, c->mask, mask_bytes));
        CB_TRY(h2d(c, c->mask.p, mask, mask_bytes));
    }
    CB_TRY(ensure(c, c->bg_status, sizeof(BackgroundStatus)));
    int32_t *dflag = reinterpret_cast<int32_t *>(c->bg_status.p);
    CB_TRY(cb200_munc_smooth_local_evidence(c, static_cast<const float *>(c->data.p),
                                            static_cast<const unsigned char *>(c->mask.p), mask_mode, m, n, ld, n, window,
                                            eps, static_cast<float *>(c->munc.p), ld, dflag));
    CB_TRY(d2h(c, invalid, dflag, sizeof(int32_t)));
    CU_TRY(cudaMemcpy2DAsync(out, (size_t)n * 4, c->munc.p, (size_t)ld * 4, (size_t)n * 4, (size_t)m,
                             cudaMemcpyDeviceToHost, c->stream));
    CU_TRY(cudaStreamSynchronize(c->stream));
    return CB200_OK;
}

int cb200_munc_finalize_eb(cb200_ctx *c, const float *local, const float *prior, const float *count_floor, int64_t n,
                           double nu_local, double nu_prior, double variance_floor, double variance_cap, int32_t use_eb,
                           float *out, cb200_munc_finalize_result *result) {
    DeviceGuard _dg(c ? c->device : 0);
    if (!c || !result) return fail(CB200_ERR_INVALID, "NULL argument");
    // messages of cconsenrich.pyx:5468-5482
    if (!(variance_floor > 0.0) || !std::isfinite(variance_floor))
        return fail(CB200_ERR_INVALID, "varianceFloor must be positive and finite");
    if (variance_cap < variance_floor || !std::isfinite(variance_cap))
        return fail(CB200_ERR_INVALID, "varianceCap must be finite and at least varianceFloor");
    if (use_eb) {
        if (!prior && n > 0) return fail(CB200_ERR_INVALID, "priorVarianceTrack is required for MUNC EB finalization");
        if (!std::isfinite(nu_local) || nu_local <= 0.0) return fail(CB200_ERR_INVALID, "nuLocal must be positive and finite");
        if (!std::isfinite(nu_prior) || nu_prior <= 0.0) return fail(CB200_ERR_INVALID, "nuPrior must be positive and finite");
        if (!std::isfinite(nu_local + nu_prior) || nu_local + nu_prior <= 0.0)
            return fail(CB200_ERR_INVALID, "posterior sample size must be positive and finite");
    }
    if (n < 0) return fail(CB200_ERR_INVALID, "n must be nonnegative");
    memset(result, 0, sizeof(*result));
    result->invalid_local = result->invalid_prior = result->invalid_count_floor = -1;
    if (n == 0) return CB200_OK;
    if (!local || !out) return fail(CB200_ERR_INVALID, "NULL argument");
    CB_TRY(ensure(c, c->bg_status, sizeof(MuncFinalizeStatus) > sizeof(BackgroundStatus) ? sizeof(MuncFinalizeStatus)
                                                                                       : sizeof(BackgroundStatus)));
    MuncFinalizeStatus *dst = static_cast<MuncFinalizeStatus *>(c->bg_status.p);
    {
        Span sp(c, FAM_MUNC);
        CU_TRY(launch_munc_finalize_eb(local, use_eb ? prior : nullptr, count_floor, n, nu_local, nu_prior, variance_floor,
                                       variance_cap, use_eb ? 1 : 0, out, dst, c->stream));
    }
    c->launches += 1;
    MuncFinalizeStatus h;
    CU_TRY(cudaMemcpyAsync(&h, dst, sizeof(h), cudaMemcpyDeviceToHost, c->stream));
    CU_TRY(cudaStreamSynchronize(c->stream));
    result->support_count = h.support;
    result->count_floor_finite = h.cfloor_finite;
    result->count_floor_added = h.cfloor_added;
    result->count_floor_missing = h.cfloor_missing;
    // the reference stops at the first offending interval; at equal index local is tested before prior
    // before the count floor
    const int64_t il = h.invalid_local, ip = h.invalid_prior, ic = h.invalid_cfloor;
    if (il != STATUS_NONE && il <= ip && il <= ic) result->invalid_local = il;
    else if (ip != STATUS_NONE && ip <= ic) result->invalid_prior = ip;
    else if (ic != STATUS_NONE) result->invalid_count_floor = ic;
    return CB200_OK;
}

int cb200_host_munc_finalize_eb(cb200_ctx *c, const float *local, const float *prior, const float *count_floor, int64_t n,
                                double nu_local, double nu_prior, double variance_floor, double variance_cap,
                                int32_t use_eb, float *out, cb200_munc_finalize_result *result) {
    DeviceGuard _dg(c ? c->device : 0);
    if (!c || !result) return fail(CB200_ERR_INVALID, "NULL argument");
    if (n > 0 && (!local || !out)) return fail(CB200_ERR_INVALID, "NULL argument");
    const float *dl = nullptr, *dp = nullptr, *dc = nullptr;
    if (n > 0) {
        CB_TRY(upload_vec(c, c->lam, local, n, true, &dl));
        CB_TRY(upload_vec(c, c->kap, prior, n, use_eb && prior != nullptr, &dp));
        CB_TRY(upload_vec(c, c->qs, count_floor, n, count_floor != nullptr, &dc));
        CB_TRY(ensure(c, c->D, (size_t)n * 4));
    }
    if (use_eb && !prior && n > 0) return fail(CB200_ERR_INVALID, "priorVarianceTrack is required for MUNC EB finalization");
    CB_TRY(cb200_munc_finalize_eb(c, dl, dp, dc, n, nu_local, nu_prior, variance_floor, variance_cap, use_eb,
                                  static_cast<float *>(c->D.p), result));
    if (n > 0) {
        CB_TRY(d2h(c, out, c->D.p, (size_t)n * 4));
        CU_TRY(cudaStreamSynchronize(c->stream));
    }
    return CB200_OK;
}

static int check_munc_seed_args(const cb200_munc_seed_args *a) {
    // messages of cconsenrich.pyx:5112-5132
    if (a->m < 0 || a->n < 0) return fail(CB200_ERR_INVALID, "bad matrix shape");
    if (a->pad < 0.0 || !std::isfinite(a->pad)) return fail(CB200_ERR_INVALID, "pad must be finite and nonnegative");
    if (!(a->variance_floor > 0.0) || !std::isfinite(a->variance_floor))
        return fail(CB200_ERR_INVALID, "varianceFloor must be positive and finite");
    if (!std::isfinite(a->variance_cap) || a->variance_cap < a->variance_floor)
        return fail(CB200_ERR_INVALID, "varianceCap must be greater than or equal to varianceFloor");
    if (a->use_weights && a->student_t &&
        (a->student_t_df <= 0.0 || a->d_omega <= 0.0 || !std::isfinite(a->student_t_df) || !std::isfinite(a->d_omega) ||
         a->omega_min <= 0.0 || a->omega_max < a->omega_min || !std::isfinite(a->omega_min) ||
         !std::isfinite(a->omega_max)))
        return fail(CB200_ERR_INVALID, "seed weight parameters are invalid");
    if (a->active_mode < 0 || a->active_mode > 2) return fail(CB200_ERR_INVALID, "activeMask must be one- or two-dimensional");
    return CB200_OK;
}

static MuncSeedArgs to_seed_args(const cb200_munc_seed_args *a) {
    MuncSeedArgs k{};
    k.data = a->data; k.munc = a->munc; k.state_mean = a->state_mean; k.state_var = a->state_var;
    k.background = a->background; k.g_var = a->g_var; k.count_floor = a->count_floor; k.omega_in = a->omega_in;
    k.rho_in = a->rho_in; k.active = a->active_mode ? a->active : nullptr;
    k.moment = a->moment; k.rho_out = a->rho_out; k.local = a->local; k.variance = a->variance;
    k.omega_raw = a->omega_raw; k.omega_out = a->omega_out;
    k.m = a->m; k.n = a->n; k.ld = a->ld; k.active_ld = a->active_ld;
    k.active_mode = a->active_mode; k.use_weights = a->use_weights; k.student_t = a->student_t;
    k.update_weights = a->update_weights;
    k.pad = a->pad; k.d_s = a->student_t_df; k.d_omega = a->d_omega; k.omega_min = a->omega_min;
    k.omega_max = a->omega_max; k.var_floor = a->variance_floor; k.var_cap = a->variance_cap;
    return k;
}

int cb200_munc_seed_pass(cb200_ctx *c, const cb200_munc_seed_args *a, int32_t *invalid) {
    DeviceGuard _dg(c ? c->device : 0);
    if (!c || !a || !invalid) return fail(CB200_ERR_INVALID, "NULL argument");
    CB_TRY(check_munc_seed_args(a));
    if (a->m > 0 && a->n > 0) {
        if (!a->data || !a->munc || !a->state_mean || !a->state_var || !a->moment || !a->rho_out || !a->omega_raw ||
            !a->omega_out || !a->local || !a->variance || (a->active_mode && !a->active))
            return fail(CB200_ERR_INVALID, "NULL argument");
        if (a->ld < a->n) return fail(CB200_ERR_INVALID, "bad matrix shape");
    }
    Span sp(c, FAM_MUNC);
    CU_TRY(launch_munc_seed_pass(to_seed_args(a), reinterpret_cast<int *>(invalid), c->stream));
    c->launches += 1;
    return CB200_OK;
}

int cb200_host_munc_seed_pass(cb200_ctx *c, const cb200_munc_seed_args *a, int32_t *invalid) {
    DeviceGuard _dg(c ? c->device : 0);
    if (!c || !a || !invalid) return fail(CB200_ERR_INVALID, "NULL argument");
    CB_TRY(check_munc_seed_args(a));
    *invalid = 0;
    const int64_t m = a->m, n = a->n;
    if (n <= 0) return CB200_OK;
    if (!a->state_mean || !a->state_var || !a->omega_raw || !a->omega_out) return fail(CB200_ERR_INVALID, "NULL argument");
    if (m > 0 && (!a->data || !a->munc || !a->moment || !a->rho_out || !a->local || !a->variance ||
                  (a->active_mode && !a->active)))
        return fail(CB200_ERR_INVALID, "NULL argument");
    cb200_munc_seed_args d = *a;
    int64_t ld = round_up(n, 32), t = 0;
    d.ld = ld;
    if (m > 0) {
        CB_TRY(upload_tracks(c, c->data, a->data, m, n, &t));
        CB_TRY(upload_tracks(c, c->munc, a->munc, m, n, &t));
        d.data = static_cast<const float *>(c->data.p);
        d.munc = static_cast<const float *>(c->munc.p);
        if (a->count_floor) {
            CB_TRY(upload_tracks(c, c->seed_mat[0], a->count_floor, m, n, &t));
            d.count_floor = static_cast<const float *>(c->seed_mat[0].p);
        }
        if (a->rho_in) {
            CB_TRY(upload_tracks(c, c->seed_mat[1], a->rho_in, m, n, &t));
            d.rho_in = static_cast<const float *>(c->seed_mat[1].p);
        }
        float **outs[4] = {&d.moment, &d.rho_out, &d.local, &d.variance};
        for (int i = 0; i < 4; ++i) {
            CB_TRY(ensure(c, c->seed_mat[2 + i], (size_t)m * (size_t)ld * 4));
            *outs[i] = static_cast<float *>(c->seed_mat[2 + i].p);
        }
        if (a->active_mode) {
            const size_t bytes = a->active_mode == 1 ? (size_t)n : (size_t)m * (size_t)n;
            CB_TRY(ensure(c, c->mask, bytes));
            CB_TRY(h2d(c, c->mask.p, a->active, bytes));
            d.active = static_cast<const unsigned char *>(c->mask.p);
            d.active_ld = n;
        }
    }
    const float *vin[5] = {a->state_mean, a->state_var, a->background, a->g_var, a->omega_in};
    const float **vdst[5] = {&d.state_mean, &d.state_var, &d.background, &d.g_var, &d.omega_in};
    for (int i = 0; i < 5; ++i) CB_TRY(upload_vec(c, c->seed_vec[i], vin[i], n, vin[i] != nullptr, vdst[i]));
    CB_TRY(ensure(c, c->seed_vec[5], (size_t)n * 4));
    CB_TRY(ensure(c, c->seed_vec[6], (size_t)n * 4));
    d.omega_raw = static_cast<float *>(c->seed_vec[5].p);
    d.omega_out = static_cast<float *>(c->seed_vec[6].p);
    CB_TRY(ensure(c, c->bg_status, sizeof(MuncFinalizeStatus)));
    int32_t *dflag = reinterpret_cast<int32_t *>(c->bg_status.p);
    if (m > 0) {
        CB_TRY(cb200_munc_seed_pass(c, &d, dflag));
        CB_TRY(d2h(c, invalid, dflag, sizeof(int32_t)));
        float *houts[4] = {a->moment, a->rho_out, a->local, a->variance};
        float *douts[4] = {d.moment, d.rho_out, d.local, d.variance};
        for (int i = 0; i < 4; ++i)
            CU_TRY(cudaMemcpy2DAsync(houts[i], (size_t)n * 4, douts[i], (size_t)ld * 4, (size_t)n * 4, (size_t)m,
                                     cudaMemcpyDeviceToHost, c->stream));
        CB_TRY(d2h(c, a->omega_raw, d.omega_raw, (size_t)n * 4));
        CB_TRY(d2h(c, a->omega_out, d.omega_out, (size_t)n * 4));
        CU_TRY(cudaStreamSynchronize(c->stream));
    } else {
        // no tracks: the reference still writes the per-interval weights (1 without an active cell)
        for (int64_t k = 0; k < n; ++k) {
            const bool weighted = a->use_weights && a->student_t;
            double raw = 1.0, om = 1.0;
            if (weighted && !a->update_weights) {
                raw = a->omega_in ? (double)a->omega_in[k] : 1.0;
                om = raw < a->omega_min ? a->omega_min : (raw > a->omega_max ? a->omega_max : raw);
            }
            a->omega_raw[k] = (float)raw;
            a->omega_out[k] = (float)om;
        }
    }
    return CB200_OK;
}

int cb200_ema(cb200_ctx *c, const void *x, int64_t n, int32_t is_double, double alpha, void *out) {
    DeviceGuard _dg(c ? c->device : 0);
    if (!c) return fail(CB200_ERR_INVALID, "ctx is NULL");
    if (!(alpha >= 0.0 && alpha <= 1.0)) return fail(CB200_ERR_INVALID, "alpha must lie in [0, 1]");
    if (n < 0) return fail(CB200_ERR_INVALID, "n must be nonnegative");
    if (n == 0) return CB200_OK;
    if (!x || !out) return fail(CB200_ERR_INVALID, "NULL argument");
    const size_t esz = is_double ? 8 : 4;
    CB_TRY(ensure(c, c->seed_mat[0], (size_t)n * esz));  // scratch: the forward pass
    CB_TRY(ensure(c, c->bg_ws, munc_ema_workspace_bytes(n)));
    Span sp(c, FAM_MUNC);
    CU_TRY(launch_munc_ema(x, c->seed_mat[0].p, out, n, is_double ? 1 : 0, alpha, c->bg_ws.p, c->stream));
    c->launches += 6;
    return CB200_OK;
}

int cb200_host_ema(cb200_ctx *c, const void *x, int64_t n, int32_t is_double, double alpha, void *out) {
    DeviceGuard _dg(c ? c->device : 0);
    if (!c) return fail(CB200_ERR_INVALID, "ctx is NULL");
    if (n <= 0) return n < 0 ? fail(CB200_ERR_INVALID, "n must be nonnegative") : CB200_OK;
    if (!x || !out) return fail(CB200_ERR_INVALID, "NULL argument");
    const size_t bytes = (size_t)n * (is_double ? 8 : 4);
    CB_TRY(ensure(c, c->seed_mat[1], bytes));
    CB_TRY(ensure(c, c->seed_mat[2], bytes));
    CB_TRY(h2d(c, c->seed_mat[1].p, x, bytes));
    CB_TRY(cb200_ema(c, c->seed_mat[1].p, n, is_double, alpha, c->seed_mat[2].p));
    CB_TRY(d2h(c, out, c->seed_mat[2].p, bytes));
    CU_TRY(cudaStreamSynchronize(c->stream));
    return CB200_OK;
}

// ---- bedGraph text ---------------------------------------------------------------------
static int bedgraph_core(cb200_ctx *c, const char *chrom, int64_t n, const int64_t *starts, const int64_t *ends,
                         int64_t start0, int64_t step, int64_t end_clip, const float *values, int64_t value_stride,
                         int64_t *bytes) {
    if (!chrom) return fail(CB200_ERR_INVALID, "chrom is NULL");
    const size_t cl = strlen(chrom);
    if (cl == 0 || cl > (size_t)BG_MAX_CHROM) return fail(CB200_ERR_INVALID, "chromosome name must be 1..%d bytes", BG_MAX_CHROM);
    if (n < 0 || value_stride <= 0) return fail(CB200_ERR_INVALID, "bad row count or value stride");
    if ((starts == nullptr) != (ends == nullptr)) return fail(CB200_ERR_INVALID, "starts and ends are given together or not at all");
    *bytes = 0;
    if (n == 0) return CB200_OK;
    if (!values) return fail(CB200_ERR_INVALID, "values is NULL");
    BedGraphArgs a{};
    memcpy(a.chrom, chrom, cl);
    a.chrom_len = (int32_t)cl;
    a.n = n;
    a.starts = reinterpret_cast<const long long *>(starts);
    a.ends = reinterpret_cast<const long long *>(ends);
    a.start0 = start0; a.step = step; a.end_clip = end_clip;
    a.values = values;
    a.value_stride = value_stride;
    const int64_t tiles = bedgraph_tiles(n);
    CB_TRY(ensure(c, c->wr_tiles, (size_t)(tiles + 1) * 8));
    long long *tile_bytes = static_cast<long long *>(c->wr_tiles.p);
    {
        Span sp(c, FAM_WRITER);
        CU_TRY(launch_bedgraph_lengths(a, tile_bytes, c->stream));
    }
    c->launches += 2;
    long long total = 0;
    CU_TRY(cudaMemcpyAsync(&total, tile_bytes + tiles, 8, cudaMemcpyDeviceToHost, c->stream));
    CU_TRY(cudaStreamSynchronize(c->stream));
    CB_TRY(ensure(c, c->wr_text, (size_t)total + 16));
    {
        Span sp(c, FAM_WRITER);
        CU_TRY(launch_bedgraph_write(a, tile_bytes, static_cast<char *>(c->wr_text.p), total, c->stream));
    }
    c->launches += 1;
    *bytes = total;
    return CB200_OK;
}

int cb200_bedgraph_chunk(cb200_ctx *c, const char *chrom, int64_t n, const int64_t *starts, const int64_t *ends,
                         int64_t start0, int64_t step, int64_t end_clip, const float *values, int64_t value_stride,
                         const char **text, int64_t *bytes) {
    DeviceGuard _dg(c ? c->device : 0);
    if (!c || !text || !bytes) return fail(CB200_ERR_INVALID, "NULL argument");
    *text = nullptr;
    CB_TRY(bedgraph_core(c, chrom, n, starts, ends, start0, step, end_clip, values, value_stride, bytes));
    *text = static_cast<const char *>(c->wr_text.p);
    return CB200_OK;
}

int cb200_host_bedgraph_chunk(cb200_ctx *c, const char *chrom, int64_t n, const int64_t *starts, const int64_t *ends,
                              int64_t start0, int64_t step, int64_t end_clip, const float *values,
                              int64_t value_stride, const char **text, int64_t *bytes) {
    DeviceGuard _dg(c ? c->device : 0);
    if (!c || !text || !bytes) return fail(CB200_ERR_INVALID, "NULL argument");
    *text = nullptr;
    *bytes = 0;
    if (n < 0 || value_stride <= 0) return fail(CB200_ERR_INVALID, "bad row count or value stride");
    if (n == 0) return CB200_OK;
    if (!values) return fail(CB200_ERR_INVALID, "values is NULL");
    if ((starts == nullptr) != (ends == nullptr)) return fail(CB200_ERR_INVALID, "starts and ends are given together or not at all");
    // only the strided values travel (the level column of a [n][2] state is gathered on the host side of the copy)
    CB_TRY(ensure(c, c->wr_vals, (size_t)n * 4));
    CU_TRY(cudaMemcpy2DAsync(c->wr_vals.p, 4, values, (size_t)value_stride * 4, 4, (size_t)n, cudaMemcpyHostToDevice,
                             c->stream));
    const int64_t *ds = nullptr, *de = nullptr;
    if (starts) {
        CB_TRY(ensure(c, c->wr_starts, (size_t)n * 8));
        CB_TRY(ensure(c, c->wr_ends, (size_t)n * 8));
        CB_TRY(h2d(c, c->wr_starts.p, starts, (size_t)n * 8));
        CB_TRY(h2d(c, c->wr_ends.p, ends, (size_t)n * 8));
        ds = static_cast<const int64_t *>(c->wr_starts.p);
        de = static_cast<const int64_t *>(c->wr_ends.p);
    }
    int64_t total = 0;
    CB_TRY(bedgraph_core(c, chrom, n, ds, de, start0, step, end_clip, static_cast<const float *>(c->wr_vals.p), 1, &total));
    if ((size_t)total + 16 > c->wr_host_cap) {
        if (c->wr_host) CU_TRY(cudaFreeHost(c->wr_host));
        c->wr_host = nullptr;
        c->wr_host_cap = 0;
        const size_t want = (size_t)total + (size_t)total / 4 + 4096;
        CU_TRY(cudaMallocHost(reinterpret_cast<void **>(&c->wr_host), want));
        c->wr_host_cap = want;
    }
    CB_TRY(d2h(c, c->wr_host, c->wr_text.p, (size_t)total));
    CU_TRY(cudaStreamSynchronize(c->stream));
    *text = c->wr_host;
    *bytes = total;
    return CB200_OK;
}

}  // extern "C"
