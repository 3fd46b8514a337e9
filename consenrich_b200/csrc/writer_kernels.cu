// consenrich_b200/csrc/writer_kernels.cu -- bedGraph text on the device.
//
// The reference appends every chromosome's tracks with pandas:
//     df[["Chromosome", "Start", "End", col]].to_csv(path, sep="\t", header=False, index=False,
//                                                    float_format="%.4f", lineterminator="\n")
// (consenrich.py:9797-9805): one row "chrom\tstart\tend\tvalue\n" per interval, the float32 track value
// printed by C's "%.4f" of its exact double value (round-half-even on the exact binary value), NaN as
// the empty string, +-inf as "inf" / "-inf".  At 124 M intervals per track that formatting is minutes of
// host time; here it is byte work on the device, bit-identical to the reference writer:
//
//   bedgraph_len_kernel    length of every row, summed per tile of BG_ROWS rows
//   bedgraph_offsets_kernel (one CTA) exclusive scan of the tile sums -> byte offset of every tile
//   bedgraph_write_kernel  every thread prints its row into the tile's shared-memory text at its offset
//                          (block scan of the lengths), then the CTA copies the text out with aligned
//                          32-bit stores
//
// "%.4f" exactly: for a float32 v, v * 10^4 = v * 625 * 16 has at most 24 + 10 significant bits, so the
// double product is EXACT and rint() (round-half-even) of it is the correctly rounded scaled value; from
// 2^24 upwards every float32 is an integer and the digits come from 128-bit integer arithmetic.
#include <cuda_runtime.h>
#include <stdint.h>

#include "writer_kernels.cuh"

namespace cb200 {

namespace {

constexpr int BG_ROWS = 256;      // rows per tile = threads per CTA
constexpr int BG_MAX_VALUE = 48;  // "-" + 39 digits + "." + 4 digits, rounded up
constexpr int BG_MAX_ROW = BG_MAX_CHROM + 1 + 20 + 1 + 20 + 1 + BG_MAX_VALUE + 1;

__device__ __forceinline__ int digits_u64(unsigned long long v) {
    int d = 1;
    while (v >= 10ull) {
        v /= 10ull;
        ++d;
    }
    return d;
}

// writes v in decimal ending just before `end`; returns the first written position
__device__ __forceinline__ char *put_u64_rev(char *end, unsigned long long v) {
    do {
        *--end = (char)('0' + (int)(v % 10ull));
        v /= 10ull;
    } while (v);
    return end;
}

// "%.4f" of a float32 (as pandas prints a float32 column).  Writes into buf (BG_MAX_VALUE bytes) when
// buf != nullptr; returns the length either way.
__device__ int format_value(float vf, char *buf) {
    const unsigned bits = __float_as_uint(vf);
    const bool neg = (bits >> 31) != 0;
    const unsigned expo = (bits >> 23) & 0xffu, mant = bits & 0x7fffffu;
    if (expo == 0xffu) {
        if (mant) return 0;  // NaN: na_rep = ""
        if (buf) {
            int p = 0;
            if (neg) buf[p++] = '-';
            buf[p++] = 'i'; buf[p++] = 'n'; buf[p++] = 'f';
        }
        return neg ? 4 : 3;
    }
    char tmp[BG_MAX_VALUE];
    char *end = tmp + BG_MAX_VALUE, *p = end;
    const double a = fabs((double)vf);
    if (a < 16777216.0) {  // below 2^24: scaled value exact in double, fits 64 bits
        const unsigned long long q = (unsigned long long)rint(a * 10000.0);
        const unsigned long long ip = q / 10000ull;
        unsigned fp = (unsigned)(q % 10000ull);
        for (int i = 0; i < 4; ++i) {
            *--p = (char)('0' + (int)(fp % 10u));
            fp /= 10u;
        }
        *--p = '.';
        p = put_u64_rev(p, ip);
    } else {  // an integer: 1.mant x 2^(expo - 127), up to 2^128
        unsigned __int128 big = (unsigned __int128)(mant | 0x800000u) << (expo - 150u);
        *--p = '0'; *--p = '0'; *--p = '0'; *--p = '0';
        *--p = '.';
        do {
            *--p = (char)('0' + (int)(big % 10));
            big /= 10;
        } while (big);
    }
    if (neg) *--p = '-';
    const int len = (int)(end - p);
    if (buf)
        for (int i = 0; i < len; ++i) buf[i] = p[i];
    return len;
}

struct RowSpec {
    long long start, end;
    float value;
};

__device__ __forceinline__ RowSpec row_of(const BedGraphArgs &a, int64_t k) {
    RowSpec r;
    r.start = a.starts ? a.starts[k] : a.start0 + k * a.step;
    r.end = a.ends ? a.ends[k] : r.start + a.step;
    if (!a.ends && a.end_clip > 0 && r.end > a.end_clip) r.end = a.end_clip;
    r.value = a.values[k * a.value_stride];
    return r;
}

__device__ __forceinline__ int row_len(const BedGraphArgs &a, const RowSpec &r) {
    // Start / End are printed as signed integers (they never are negative on this path; "-" is handled anyway)
    const int ls = digits_u64((unsigned long long)(r.start < 0 ? -r.start : r.start)) + (r.start < 0);
    const int le = digits_u64((unsigned long long)(r.end < 0 ? -r.end : r.end)) + (r.end < 0);
    return a.chrom_len + 1 + ls + 1 + le + 1 + format_value(r.value, nullptr) + 1;
}

__device__ __forceinline__ int block_exclusive_scan(int v, int *total, int *sh /* BG_ROWS / 32 + 1 */) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int inc = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int o = __shfl_up_sync(0xffffffffu, inc, d);
        if (lane >= d) inc += o;
    }
    if (lane == 31) sh[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        int w = lane < BG_ROWS / 32 ? sh[lane] : 0;
#pragma unroll
        for (int d = 1; d < BG_ROWS / 32; d <<= 1) {
            const int o = __shfl_up_sync(0xffffffffu, w, d);
            if (lane >= d) w += o;
        }
        if (lane < BG_ROWS / 32) sh[lane] = w;  // inclusive over warps
    }
    __syncthreads();
    const int base = warp > 0 ? sh[warp - 1] : 0;
    *total = sh[BG_ROWS / 32 - 1];
    return base + inc - v;
}

__global__ void __launch_bounds__(BG_ROWS) bedgraph_len_kernel(const BedGraphArgs a, long long *tile_bytes) {
    __shared__ int sh[BG_ROWS / 32 + 1];
    const int64_t k = (int64_t)blockIdx.x * BG_ROWS + threadIdx.x;
    const int len = k < a.n ? row_len(a, row_of(a, k)) : 0;
    int total;
    block_exclusive_scan(len, &total, sh);
    if (threadIdx.x == 0) tile_bytes[blockIdx.x] = total;
}

// in place: tile_bytes[t] -> byte offset of tile t; tile_bytes[tiles] = total
__global__ void __launch_bounds__(1024) bedgraph_offsets_kernel(long long *tile_bytes, int64_t tiles) {
    __shared__ long long sh[33];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t per = (tiles + 1023) / 1024;
    const int64_t t0 = (int64_t)tid * per;
    long long mine = 0;
    for (int64_t i = 0; i < per; ++i)
        if (t0 + i < tiles) mine += tile_bytes[t0 + i];
    long long inc = mine;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const long long o = __shfl_up_sync(0xffffffffu, inc, d);
        if (lane >= d) inc += o;
    }
    if (lane == 31) sh[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        long long w = sh[lane];
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const long long o = __shfl_up_sync(0xffffffffu, w, d);
            if (lane >= d) w += o;
        }
        sh[lane] = w;
    }
    __syncthreads();
    long long run = (warp > 0 ? sh[warp - 1] : 0) + inc - mine;
    const long long total = sh[31];
    for (int64_t i = 0; i < per; ++i) {
        if (t0 + i < tiles) {
            const long long b = tile_bytes[t0 + i];
            tile_bytes[t0 + i] = run;
            run += b;
        }
    }
    if (tid == 0) tile_bytes[tiles] = total;
}

__global__ void __launch_bounds__(BG_ROWS) bedgraph_write_kernel(const BedGraphArgs a, const long long *tile_offset,
                                                                char *out, long long out_cap) {
    __shared__ int sh[BG_ROWS / 32 + 1];
    __shared__ __align__(16) char text[BG_ROWS * BG_MAX_ROW];
    const int64_t k = (int64_t)blockIdx.x * BG_ROWS + threadIdx.x;
    RowSpec r{};
    int len = 0;
    if (k < a.n) {
        r = row_of(a, k);
        len = row_len(a, r);
    }
    int total;
    const int off = block_exclusive_scan(len, &total, sh);
    if (k < a.n) {
        char *p = text + off;
        for (int i = 0; i < a.chrom_len; ++i) *p++ = a.chrom[i];
        *p++ = '\t';
        {
            char tmp[24];
            char *e = tmp + 24;
            char *s = put_u64_rev(e, (unsigned long long)(r.start < 0 ? -r.start : r.start));
            if (r.start < 0) *--s = '-';
            while (s < e) *p++ = *s++;
        }
        *p++ = '\t';
        {
            char tmp[24];
            char *e = tmp + 24;
            char *s = put_u64_rev(e, (unsigned long long)(r.end < 0 ? -r.end : r.end));
            if (r.end < 0) *--s = '-';
            while (s < e) *p++ = *s++;
        }
        *p++ = '\t';
        p += format_value(r.value, p);
        *p++ = '\n';
    }
    __syncthreads();
    // ---- copy the tile's text out: aligned 32-bit stores, single bytes at the two ragged ends ----
    const long long g0 = tile_offset[blockIdx.x];
    if (g0 + total > out_cap) return;  // never write past the caller's buffer (the caller checks the total)
    char *dst = out + g0;
    const int head = (int)((4 - (reinterpret_cast<uintptr_t>(dst) & 3)) & 3);
    const int nhead = head < total ? head : total;
    if (threadIdx.x < nhead) dst[threadIdx.x] = text[threadIdx.x];
    const int words = (total - nhead) >> 2;
    for (int w = threadIdx.x; w < words; w += BG_ROWS) {
        const char *s = text + nhead + 4 * w;
        const unsigned v = (unsigned)(unsigned char)s[0] | ((unsigned)(unsigned char)s[1] << 8) |
                           ((unsigned)(unsigned char)s[2] << 16) | ((unsigned)(unsigned char)s[3] << 24);
        *reinterpret_cast<unsigned *>(dst + nhead + 4 * w) = v;
    }
    const int tail0 = nhead + 4 * words;
    if (threadIdx.x < total - tail0) dst[tail0 + threadIdx.x] = text[tail0 + threadIdx.x];
}

}  // namespace

int64_t bedgraph_tiles(int64_t n) { return (n + BG_ROWS - 1) / BG_ROWS; }
int64_t bedgraph_max_row_bytes() { return BG_MAX_ROW; }

cudaError_t launch_bedgraph_lengths(const BedGraphArgs &a, long long *tile_bytes, cudaStream_t st) {
    const int64_t tiles = bedgraph_tiles(a.n);
    if (tiles <= 0) return cudaSuccess;
    bedgraph_len_kernel<<<(unsigned)tiles, BG_ROWS, 0, st>>>(a, tile_bytes);
    bedgraph_offsets_kernel<<<1, 1024, 0, st>>>(tile_bytes, tiles);
    return cudaGetLastError();
}

cudaError_t launch_bedgraph_write(const BedGraphArgs &a, const long long *tile_offset, char *out, long long out_cap,
                                  cudaStream_t st) {
    const int64_t tiles = bedgraph_tiles(a.n);
    if (tiles <= 0) return cudaSuccess;
    bedgraph_write_kernel<<<(unsigned)tiles, BG_ROWS, 0, st>>>(a, tile_offset, out, out_cap);
    return cudaGetLastError();
}

}  // namespace cb200
