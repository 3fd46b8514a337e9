// consenrich_b200/csrc/writer_kernels.cuh -- launch interface of the bedGraph text kernels
// (writer_kernels.cu; reference writer: consenrich.py:9797-9805).  Device pointers only.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace cb200 {

constexpr int BG_MAX_CHROM = 32;  // bytes of a chromosome name

struct BedGraphArgs {
    char chrom[BG_MAX_CHROM];
    int32_t chrom_len;
    int64_t n;                 // rows (intervals)
    const long long *starts;   // [n] or nullptr: start0 + k step
    const long long *ends;     // [n] or nullptr: start + step (clipped to end_clip when end_clip > 0)
    long long start0, step, end_clip;
    const float *values;       // value of row k at values[k * value_stride] (stride 2 reads the level of a [n][2] state)
    int64_t value_stride;
};

int64_t bedgraph_tiles(int64_t n);
int64_t bedgraph_max_row_bytes();
// tile_bytes: [tiles + 1] int64 on the device.  On return tile_bytes[t] = byte offset of tile t's text,
// tile_bytes[tiles] = total bytes.
cudaError_t launch_bedgraph_lengths(const BedGraphArgs &a, long long *tile_bytes, cudaStream_t st);
cudaError_t launch_bedgraph_write(const BedGraphArgs &a, const long long *tile_offset, char *out, long long out_cap,
                                  cudaStream_t st);

}  // namespace cb200
