// normalize.cuh - Data.__init__ (src/data.py:16-21) on device.
//
//   normalized_abd = sklearn.normalize(abd, "l1")   -> float64: x / sum|x| (zero rows stay zero)
//   weights[i]     = max(normalized_abd[i]) ** 2     -> float64
//   abd            = normalized_abd.astype(float32); tnf likewise
// Tallies are integers < 2^32, so the row sum is exact in int64, the fp64 divide is
// correctly rounded exactly as numpy's, and the fp32 store rounds the same fp64 value:
// results are bit-identical, not merely within the 1e-6 the north star allows.
// max(x_j / s) = max(x_j) / s because correctly rounded division is monotone.
// One warp per row; reads 4 B and writes 4 B per element (HBM-bound, streaming).
#pragma once
#include <stdint.h>
#include <cuda_runtime.h>

namespace pg {

__global__ void __launch_bounds__(256)
normalize_rows_kernel(const uint32_t* __restrict__ raw, int64_t rows, int dim, float* __restrict__ out, double* __restrict__ weights)
{
    const int lane = threadIdx.x & 31;
    const int64_t warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t r = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; r < rows; r += warps) {
        const uint32_t* src = raw + r * dim;
        unsigned long long sum = 0;
        uint32_t mx = 0;
        for (int j = lane; j < dim; j += 32) {
            uint32_t v = __ldg(src + j);
            sum += v;
            mx = max(mx, v);
        }
#pragma unroll
        for (int d = 16; d; d >>= 1) {
            sum += __shfl_xor_sync(0xffffffffu, sum, d);
            mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, d));
        }
        const double norm = sum ? (double)sum : 1.0; // sklearn: zero norms are replaced by 1
        float* dst = out + r * dim;
        for (int j = lane; j < dim; j += 32) dst[j] = (float)((double)__ldg(src + j) / norm);
        if (weights && lane == 0) {
            const double m = (double)mx / norm;
            weights[r] = m * m;
        }
    }
}

} // namespace pg
