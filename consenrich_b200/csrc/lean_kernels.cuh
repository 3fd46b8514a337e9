// consenrich_b200/csrc/lean_kernels.cuh -- the ECM's inner sweeps on run-major private tracks.
//
// Inside cfixedBackgroundECM (cconsenrich.pyx:8151-8335) the forward filter's tracks are read by
// nothing but the backward pass of the same sweep, and -- when the process precision kappa is the
// only multiplier being fitted, the CLI default -- the smoothed tracks by nothing but the kappa
// update.  None of them has to exist in the reference's public layouts.  These kernels keep every
// per-bin track of the inner sweeps in a RUN-MAJOR layout instead: a warp owns a segment of
// 32 runs x L consecutive bins (one run per lane), and element i of all 32 runs is one contiguous
// row, so every load and store of the sequential per-run recursions is a fully coalesced warp
// access straight from / to registers -- no shared-memory staging, no bank conflicts, no barriers.
//
//   position of bin k:   seg = k / (32 L),  lane = (k mod 32 L) / L,  i = k mod L
//                        index = seg * 32 L + i * 32 + lane
//
// A forward pass is three launches (no spinning between CTAs):
//   lean_fwd_compose   every run composes its filtering element (Sarkka & Garcia-Fernandez); a warp scan
//                      and a scan over the CTA's 8 warps leave per-run exclusive elements and one
//                      aggregate per GROUP (CTA: 8 segments = 256 runs)
//   lean_fwd_prefix    one CTA scans the group aggregates into per-group start states
//   lean_fwd_replay    every run replays the reference's own recursion (float32 rounding points
//                      included, ssm_math.cuh: kf2_step) from its exact start state, writes the
//                      compact forward track (x, P upper triangle, Q upper triangle: 32 B per bin),
//                      and composes the smoother's run elements on the way
// and a backward pass two:
//   lean_bwd_suffix    one CTA scans the groups' smoothing aggregates from the far end
//   lean_bwd_replay    every run replays the reference's RTS recursion and forms kappa_{k+1} the
//                      moment smoothed bins k, k+1 and their lag-one covariance are in registers
//                      (LEAN_PUBLIC: writes stateSmoothed / stateCovarSmoothed / lagCovSmoothed in
//                      the reference's layouts instead -- the one pass whose result the call returns)
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "ssm_math.cuh"

namespace cb200 {

constexpr int LEAN_THREADS = 256;
constexpr int LEAN_WARPS = LEAN_THREADS / 32;  // segments per CTA = one GROUP: the unit of the cross-CTA scan
constexpr int LEAN_SCAN_THREADS = 512;  // the single-CTA group scans
constexpr int LEAN_MIN_BINS = 4096;     // below this the look-back kernels serve the call

struct LeanGeom {
    int64_t n;     // bins
    int32_t logL;  // run length L = 1 << logL (32 or 64)
    int32_t W;     // segments (warps): ceil(n / (32 L))
    int32_t G;     // groups (CTAs of LEAN_WARPS segments): ceil(W / LEAN_WARPS)
    int32_t Gp;    // pitch of the per-group arrays (G rounded up to 32)
    __host__ __device__ int64_t L() const { return (int64_t)1 << logL; }
    __host__ __device__ int64_t seg_bins() const { return (int64_t)32 << logL; }
    // positions of the run-major arrays: whole groups (the last group's empty segments are addressable)
    __host__ __device__ int64_t npad() const { return (int64_t)G * LEAN_WARPS * seg_bins(); }
    // run-major position of bin k
    __host__ __device__ int64_t index(int64_t k) const {
        const int64_t sb = seg_bins();
        const int64_t r = k & (sb - 1);
        return (k - r) + ((r & (L() - 1)) << 5) + (r >> logL);
    }
};

inline LeanGeom lean_geom(int64_t n, int logL) {
    LeanGeom g;
    g.n = n;
    g.logL = logL;
    const int64_t sb = (int64_t)32 << logL;
    g.W = (int32_t)((n + sb - 1) / sb);
    g.G = (g.W + LEAN_WARPS - 1) / LEAN_WARPS;
    g.Gp = (g.G + 31) / 32 * 32;
    return g;
}

// scratch of a forward pass.  Per-group arrays are structures of arrays (component j of group g at
// [j * Gp + g]): the group scan reads them coalesced.
struct LeanFwdScratch {
    double *fagg;   // [14][Gp]          filtering aggregate of the group
    double *fex;    // [G * 8][14][32]   per-run exclusive element within the group
    double *fpref;  // [5][Gp]           Gaussian at the start of the group
    double *partials;  // [G]
    int32_t *counter;  // zero between launches
};

// one set of forward tracks (the ECM keeps two: the pass that closes an iteration runs ahead into
// the spare set)
struct LeanTrack {
    float4 *A;     // [npad] x0 x1 P00 P01
    float4 *B;     // [npad] P11 Q00 Q01 Q11   (Q of the same bin, float32 as the reference stores it)
    double *sagg;  // [9][Gp]         smoothing aggregate of the group
    double *sex;   // [G * 8][9][32]  per-run exclusive (from the far end of the group) smoothing element
};

// A chromosome split over several GPUs (contiguous bin ranges, one shard per rank): what a shard knows
// about its neighbours.  The shards exchange ONE payload per pass (cabi.cu: cb200_split_*):
//   forward : the shard's filtering aggregate (14 f64) + kappa and qScale of its first bin
//   backward: the shard's smoothing aggregate (9 f64) + the filtered Gaussian of its last bin (5 f64)
struct LeanShard {
    int32_t is_first, is_last;  // of the chromosome (an unsplit chromosome: both 1)
    const double *fwd_next;     // device: forward payload of the NEXT shard ([14] kappa, [15] qScale of its first
                                // bin, as that shard's forward pass of this sweep saw them); is_last == 0
    const double *bwd_prev;     // device: backward payload of the PREVIOUS shard ([9..13] filtered Gaussian of its
                                // last bin, float32 values); is_first == 0, kappa-carrying backward replay only
    const double *gathered;     // device [world][LEAN_PAYLOAD] or nullptr: the payloads of this pass; the group scan
    int32_t rank, world;        // pushes its start state through the aggregates of the shards before it
};
constexpr int LEAN_PAYLOAD = 16;  // doubles per shard per pass

struct LeanFwdArgs {
    LeanGeom g;
    const double2 *SA, *SB;  // run-major fold statistics {S0,S1}, {S2,SL}
    const float *kap, *qs;   // run-major multipliers (qs may be nullptr)
    LeanFwdScratch sc;
    LeanTrack trk;
    double *sums;            // device double[2] or nullptr: {0, sum NLL}
    double m, inv_m, mlog2pi;
    Model2 M;                // F = [[1, F01], [0, 1]] only
    double state_init, cov_init, kap_min, kap_max;
    int32_t want_nll, do_store;
    LeanShard sh;
};

struct LeanBwdArgs {
    LeanGeom g;
    LeanTrack trk;
    double *ssuf;            // [5][Gp] smoothed Gaussian just beyond the group
    const float *qs;         // run-major processQScale or nullptr
    float *kap_out;          // run-major, written at the position of bin k+1
    float *xs, *Ps, *lag;    // LEAN_PUBLIC: the reference's layouts
    int64_t lag_rows;
    Model2 M;
    double nu, kap_lo, kap_hi;
    double qi00, qi01, qi10, qi11;  // Q0^-1
    LeanShard sh;
    float *kap_discard;             // device float: where a non-last shard drops the kappa of the bin after its
                                    // last one (the next shard computes that multiplier itself)
};

// all return the cudaError_t of the launch; a forward pass = compose, prefix, replay in this order, a
// backward pass = suffix, replay (a publishing replay re-uses the suffix states of the sweep before it)
cudaError_t lean_fwd_compose(const LeanFwdArgs &a, cudaStream_t st);
cudaError_t lean_fwd_prefix(const LeanFwdArgs &a, cudaStream_t st);
cudaError_t lean_fwd_replay(const LeanFwdArgs &a, cudaStream_t st);
cudaError_t lean_bwd_suffix(const LeanBwdArgs &a, cudaStream_t st);
cudaError_t lean_bwd_replay(const LeanBwdArgs &a, bool publish, cudaStream_t st);
// Split chromosomes: the shard's payload of a pass -- its whole aggregate (ordered reduction of the group
// aggregates; forward: filtering element, 14 f64; backward: smoothing element, 9 f64) and, forward, [14], [15] =
// kappa, qScale of its first bin, backward, [9..13] = the filtered Gaussian of its last bin.
cudaError_t lean_shard_payload(const LeanFwdArgs &a, const LeanTrack &trk, bool backward, double *payload, cudaStream_t st);
// run-major <-> linear copies of per-bin float vectors (fill: value of the padding positions)
cudaError_t lean_gather_f32(const float *linear, float *run_major, const LeanGeom &g, float fill, cudaStream_t st);
cudaError_t lean_scatter_f32(const float *run_major, float *linear, const LeanGeom &g, cudaStream_t st);
cudaError_t lean_fill_f32(float *run_major, const LeanGeom &g, float value, cudaStream_t st);
cudaError_t lean_configure();

}  // namespace cb200
