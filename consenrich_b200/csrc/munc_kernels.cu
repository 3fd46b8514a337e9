// Dense [tracks x intervals] kernels of the observation-noise (MUNC) stage that feeds the path
// (SURVEY 8f, next #3).  First of them: the centred rolling mean with an exclusion mask that turns
// per-cell local evidence into the local variance track, cMuncSmoothDenseLocalEvidence
// (cconsenrich.pyx:5547-5740).
//
// The reference slides one running sum along each row (add the entering cell, subtract the leaving
// one).  Here a CTA takes a tile of TILE consecutive outputs of one row, loads the TILE + window cells
// its windows cover into shared memory as float64 (0 for masked cells) next to their 0/1 counts, forms
// tile-local prefix sums (thread-sequential segments + one block scan of the segment totals) and
// reads every window as a difference of two prefixes.  The float64 window sums agree with the
// reference's running sum to ~1e-13 relative, far below the float32 rounding of the output.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "munc_kernels.cuh"
#include "ssm_math.cuh"  // cb_div: the correctly rounded quotient without the special-case path

namespace cb200 {
namespace {

constexpr int RM_THREADS = 256;

__device__ __forceinline__ bool mask_allows(const uint8_t *mask, int mode, int64_t mask_ld, int64_t j, int64_t k) {
    // _muncSeedMaskAllowsCell with nonzeroMeansActive = False (pyx:4746-4766): nonzero excludes
    if (mode == 0) return true;
    const uint8_t v = mode == 1 ? mask[k] : mask[j * mask_ld + k];
    return v == 0;
}

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

__global__ void __launch_bounds__(RM_THREADS)
rolling_mean_kernel(const float *__restrict__ local, const uint8_t *__restrict__ mask, int mask_mode, int64_t n,
                    int64_t ld, int64_t mask_ld, int64_t W, double eps, int tile, float *__restrict__ out, int64_t out_ld,
                    int *__restrict__ invalid) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int64_t j = blockIdx.y;
    const int64_t i0 = (int64_t)blockIdx.x * tile;
    if (i0 >= n) return;
    const int64_t i1 = min(i0 + (int64_t)tile, n);  // outputs [i0, i1)
    int64_t lo, hi, t;
    window_of(i0, n, W, lo, t);
    window_of(i1 - 1, n, W, t, hi);  // cells [lo, hi) cover every window of the tile
    const int cells = (int)(hi - lo);
    double *ps = reinterpret_cast<double *>(smem_raw);             // [cells + 1] exclusive prefix of the values
    int *pc = reinterpret_cast<int *>(ps + (cells + 1 + 1) / 2 * 2);  // [cells + 1] exclusive prefix of the counts
    __shared__ double seg_sum[RM_THREADS];
    __shared__ int seg_cnt[RM_THREADS];
    const float *row = local + j * ld;

    // coalesced load, four cells per thread in flight (values are fetched whether masked or not, so that
    // nothing waits for the mask); masked cells contribute 0 to the sums and to the counts
    int bad = 0;
    constexpr int LD_UNROLL = 4;
    for (int c0 = threadIdx.x; c0 < cells; c0 += RM_THREADS * LD_UNROLL) {
        float f[LD_UNROLL];
        bool on[LD_UNROLL];
#pragma unroll
        for (int u = 0; u < LD_UNROLL; ++u) {
            const int c = c0 + u * RM_THREADS;
            const int64_t k = lo + (c < cells ? c : 0);
            f[u] = row[k];
            on[u] = c < cells && mask_allows(mask, mask_mode, mask_ld, j, k);
        }
#pragma unroll
        for (int u = 0; u < LD_UNROLL; ++u) {
            const int c = c0 + u * RM_THREADS;
            if (c < cells) {
                if (on[u] && (!(f[u] > 0.0f) || f[u] == INFINITY)) bad = 1;  // not positive and finite (pyx:5569-5571)
                ps[c + 1] = on[u] ? (double)f[u] : 0.0;
                pc[c + 1] = on[u] ? 1 : 0;
            }
        }
    }
    __syncthreads();
    // each thread owns a contiguous segment: local inclusive scan
    const int per = (cells + RM_THREADS - 1) / RM_THREADS;
    const int s0 = min((int)threadIdx.x * per, cells), s1 = min(s0 + per, cells);
    double run = 0.0;
    int cnt = 0;
    for (int c = s0; c < s1; ++c) {
        run += ps[c + 1];
        cnt += pc[c + 1];
        ps[c + 1] = run;
        pc[c + 1] = cnt;
    }
    seg_sum[threadIdx.x] = run;
    seg_cnt[threadIdx.x] = cnt;
    if (bad) atomicOr(invalid, 1);
    __syncthreads();
    // exclusive scan of the segment totals (256 values: one warp-strided pass is plenty)
    if (threadIdx.x < 32) {
        double a = 0.0;
        int b = 0;
        // lane l scans segments [8 l, 8 l + 8)
        double loc_s[8];
        int loc_c[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            loc_s[u] = a;
            loc_c[u] = b;
            a += seg_sum[threadIdx.x * 8 + u];
            b += seg_cnt[threadIdx.x * 8 + u];
        }
        double ex_s = a;
        int ex_c = b;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const double os = __shfl_up_sync(0xffffffffu, ex_s, d);
            const int oc = __shfl_up_sync(0xffffffffu, ex_c, d);
            if ((int)threadIdx.x >= d) {
                ex_s += os;
                ex_c += oc;
            }
        }
        ex_s -= a;  // exclusive over lanes
        ex_c -= b;
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            seg_sum[threadIdx.x * 8 + u] = ex_s + loc_s[u];
            seg_cnt[threadIdx.x * 8 + u] = ex_c + loc_c[u];
        }
    }
    __syncthreads();
    const double off_s = seg_sum[threadIdx.x];
    const int off_c = seg_cnt[threadIdx.x];
    for (int c = s0; c < s1; ++c) {
        ps[c + 1] += off_s;
        pc[c + 1] += off_c;
    }
    if (threadIdx.x == 0) {
        ps[0] = 0.0;
        pc[0] = 0;
    }
    __syncthreads();

    float *orow = out + j * out_ld;
    const int64_t half = W / 2;
    // away from the ends of the row every window is [i - half, i - half + W): 32-bit tile-relative indices
    const bool interior = i0 >= half && (i1 - 1 - half) + W <= n;
    for (int64_t i = i0 + threadIdx.x; i < i1; i += RM_THREADS) {
        int bl, br;
        if (interior) {
            bl = (int)(i - i0);
            br = bl + (int)W;
        } else {
            int64_t l, r;
            window_of(i, n, W, l, r);
            bl = (int)(l - lo);
            br = (int)(r - lo);
        }
        const int count = pc[br] - pc[bl];
        double v = count > 0 ? cb_div(ps[br] - ps[bl], (double)count) : (double)row[i];
        if (v < eps) v = eps;
        orow[i] = (float)v;
    }
}

}  // namespace

int munc_rolling_tile(int64_t window) {
    // outputs per CTA: at least as many as the window so that a cell is loaded at most ~twice
    int tile = 2048;
    while (tile < window) tile *= 2;
    return tile;
}

size_t munc_rolling_smem(int tile, int64_t window) {
    const size_t cells = (size_t)tile + (size_t)window + 2;
    return cells * 8 + cells * 4 + 16;
}

cudaError_t launch_munc_rolling_mean(const float *local, const uint8_t *mask, int mask_mode, int64_t m, int64_t n,
                                     int64_t ld, int64_t mask_ld, int64_t window, double eps, float *out, int64_t out_ld,
                                     int *invalid, cudaStream_t st) {
    if (m <= 0 || n <= 0) return cudaSuccess;
    const int tile = munc_rolling_tile(window);
    const size_t smem = munc_rolling_smem(tile, window);
    cudaError_t e = cudaFuncSetAttribute(rolling_mean_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    e = cudaMemsetAsync(invalid, 0, sizeof(int), st);
    if (e != cudaSuccess) return e;
    const dim3 grid((unsigned)((n + tile - 1) / tile), (unsigned)m);
    rolling_mean_kernel<<<grid, RM_THREADS, smem, st>>>(local, mask, mask_mode, n, ld, mask_ld, window, eps, tile, out, out_ld,
                                                        invalid);
    return cudaGetLastError();
}

}  // namespace cb200
