// consenrich_b200/csrc/apn_kernels.cu -- forward filter with adaptive process noise (APN).
//
// With ECM_useAPN and no processQScale the reference scales the process noise of bin k+1 by a factor
// that depends on bin k's innovation statistic D_k (cconsenrich.pyx:510-527, 688-703): a nonlinear
// feedback from the filtered state into the model, so the pass is not an associative scan and no
// parallel-in-time kernel reproduces it.  It is still a recursion over the FOLDED per-bin statistics
// (32 bytes per bin, whatever the number of tracks), so it runs on the device as what it is: one
// sequential chain.  A single warp walks the chromosome in chunks of 32 bins; all lanes stream the
// next chunk's statistics in (registers -> shared memory) and the finished chunk's tracks out
// (coalesced), lane 0 runs the reference's recursion -- the same kf2_step / kf1_step the scan kernels
// replay, followed by the reference's APN update in its own operation order (IEEE sqrt and division).
// ~250 cycles per bin: 0.3 s for chr19 at 25 bp, against 0.5 s x (tracks / 10) on a host core.
#include <cuda_runtime.h>

#include "ssm_kernels.cuh"

namespace cb200 {

namespace {

struct ApnState {
    double scale;  // apnScale
};

// the reference's update of apnScale after bin k (pyx:510-527 with proc_noise = 0.5 (Q00 + Q11);
// pyx:688-703 with proc_noise = apnScale q0)
__device__ __forceinline__ void apn_update(const ApnArgs &a, double &apn, float d32, double proc_noise) {
    const double dk = (double)d32;
    if (dk > a.apn_thresh && proc_noise < a.apn_max_q) {
        apn *= sqrt(a.apn_scale * (dk - a.apn_thresh) + a.apn_pc);
    } else if (dk <= a.apn_thresh && proc_noise > a.apn_min_q) {
        apn *= 1.0 / sqrt(a.apn_scale * (a.apn_thresh - dk) + a.apn_pc);
    }
    const double pn = apn * a.q_diag;
    if (pn < a.apn_min_q)
        apn = a.apn_min_q / a.q_diag;
    else if (pn > a.apn_max_q)
        apn = a.apn_max_q / a.q_diag;
}

template <int DIM>
__global__ void __launch_bounds__(32) apn_forward_kernel(const ApnArgs a) {
    __shared__ double2 sA[32], sB[32];
    __shared__ float sLam[32];
    __shared__ float4 oP[32], oQ[32];
    __shared__ float2 oX[32];
    __shared__ float oD[32];
    const int lane = threadIdx.x;
    Kf2 s2{r32(a.state_init), 0.0, r32(a.cov_init), 0.0, 0.0, r32(a.cov_init)};
    State1 s1{a.state_init, a.cov_init};
    double apn = 1.0, sum_d = 0.0, sum_nll = 0.0;
    NllAcc acc;
    nll_acc_init(acc);
    const bool per_bin = a.nll_in_d != 0;
    const int64_t chunks = (a.n + 31) / 32;
    double2 rA = make_double2(0.0, 0.0), rB = make_double2(0.0, 0.0);
    float rLam = 1.0f;
    if (lane < a.n) {
        rA = a.SA[lane];
        rB = a.SB[lane];
        if (a.use_lambda) rLam = a.lam[lane];
    }
    for (int64_t c = 0; c < chunks; ++c) {
        const int64_t k = c * 32 + lane;
        sA[lane] = rA;
        sB[lane] = rB;
        sLam[lane] = rLam;
        __syncwarp();
        if (k + 32 < a.n) {  // the next chunk streams in underneath the recursion
            rA = a.SA[k + 32];
            rB = a.SB[k + 32];
            if (a.use_lambda) rLam = a.lam[k + 32];
        }
        if (lane == 0) {
            const int cnt = (int)min((int64_t)32, a.n - c * 32);
            for (int i = 0; i < cnt; ++i) {
                const double2 s01 = sA[i], s2l = sB[i];
                const double lam = a.use_lambda ? clampd((double)sLam[i], a.lam_min, a.lam_max) : 1.0;
                BinOut o;
                float d32;
                if (DIM == 2) {
                    kf2_step<false>(s2, a.M, apn, lam, s01.x, s01.y, s2l.x, s2l.y, a.m, a.inv_m, a.mlog2pi, a.want_nll != 0,
                                    per_bin, o, acc);
                    d32 = (float)o.stat;
                    oX[i] = make_float2((float)s2.x0, (float)s2.x1);
                    oP[i] = make_float4((float)s2.P00, (float)s2.P01, (float)s2.P10, (float)s2.P11);
                    oQ[i] = make_float4((float)o.Q00, (float)o.Q01, (float)o.Q10, (float)o.Q11);
                    apn_update(a, apn, d32, 0.5 * (o.Q00 + o.Q11));
                } else {
                    const double proc_noise = apn * a.M.q00;
                    kf1_step(s1, proc_noise, lam, s01.x, s01.y, s2l.x, s2l.y, a.m, a.inv_m, a.mlog2pi, a.want_nll != 0, per_bin,
                             o, acc);
                    d32 = (float)o.stat;
                    oX[i] = make_float2((float)s1.x, 0.0f);
                    oP[i] = make_float4((float)s1.P, 0.0f, 0.0f, 0.0f);
                    oQ[i] = make_float4((float)o.Q00, 0.0f, 0.0f, 0.0f);
                    apn_update(a, apn, d32, proc_noise);
                }
                oD[i] = d32;
                sum_d += (double)d32;
                sum_nll += o.nll;
                if (a.want_nll && !per_bin && (i & 7) == 7) nll_acc_renorm(acc, a.use_lambda != 0);
            }
            if (a.want_nll && !per_bin) nll_acc_renorm(acc, a.use_lambda != 0);
        }
        __syncwarp();
        if (k < a.n) {
            if (a.do_store) {
                if (DIM == 2) {
                    reinterpret_cast<float2 *>(a.xf)[k] = oX[lane];
                    reinterpret_cast<float4 *>(a.Pf)[k] = oP[lane];
                    if (k > 0) reinterpret_cast<float4 *>(a.Qf)[k - 1] = oQ[lane];
                } else {
                    a.xf[k] = oX[lane].x;
                    a.Pf[k] = oP[lane].x;
                    if (k > 0) a.Qf[k - 1] = oQ[lane].x;
                }
            }
            if (a.D) a.D[k] = oD[lane];
        }
        __syncwarp();
    }
    if (lane == 0 && a.sums) {
        if (a.want_nll && !per_bin) sum_nll += nll_acc_finish(acc, a.m, a.mlog2pi);
        a.sums[0] = sum_d;
        a.sums[1] = sum_nll;
    }
}

}  // namespace

cudaError_t launch_apn_forward(int dim, const ApnArgs &a, cudaStream_t st) {
    if (a.n <= 0) return cudaSuccess;
    if (dim == 2)
        apn_forward_kernel<2><<<1, 32, 0, st>>>(a);
    else
        apn_forward_kernel<1><<<1, 32, 0, st>>>(a);
    return cudaGetLastError();
}

}  // namespace cb200
