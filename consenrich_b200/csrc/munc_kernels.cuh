// Dense kernels of the observation-noise (MUNC) stage: launch interface (see munc_kernels.cu).
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

namespace cb200 {

constexpr int64_t MUNC_MAX_WINDOW = 8192;  // tile + window cells must fit one CTA's shared memory

int munc_rolling_groups(int64_t window);
size_t munc_rolling_smem(int groups);

// out[j][i] = max(eps, mean of the unmasked local[j][k] over the centred window of i) (float32),
// the cell itself where the whole window is masked.  mask_mode: 0 none, 1 per interval [n],
// 2 per cell [m][mask_ld]; nonzero excludes.  *invalid (device int) is set when an unmasked cell is
// not positive and finite.  1 <= window <= MUNC_MAX_WINDOW.
cudaError_t launch_munc_rolling_mean(const float *local, const uint8_t *mask, int mask_mode, int64_t m, int64_t n,
                                     int64_t ld, int64_t mask_ld, int64_t window, double eps, float *out, int64_t out_ld,
                                     int *invalid, cudaStream_t st);

}  // namespace cb200
