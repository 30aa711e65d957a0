// sampler.cuh - step-2 input pipeline on the device: the weighted sampler and the batch gather.
//
// Replaces CustomWeightedRandomSampler.__iter__ (src/utils.py:11-23 -> numpy.random.choice(range(N), size, p, replace))
// and the DataLoader's per-item gather of Data.__getitem__ (src/data.py:27-31, src/pangaea.py:87-89).  numpy's choice is
//   with replacement:     cdf = p.cumsum(); cdf /= cdf[-1]; idx = cdf.searchsorted(random_sample(size), side="right")
//   without replacement:  rounds of the same with the probabilities of the items already drawn set to zero; of each
//                         round's draws only the first occurrence of a value is kept, in draw order.
// The uniforms come from the caller (numpy's own generator, so a seeded run draws exactly what the reference draws); what
// runs here is everything that scales with N: the cumulative sum - one thread, sequentially, because numpy's cumsum is
// sequential and the cdf must round the same way - the binary searches, the first-occurrence filter and the row gather.
#pragma once
#include <stdint.h>
#include <cuda_runtime.h>

namespace pg {

// p[i] = w[i] / total (fp64, as `weights.numpy() / torch.sum(weights).numpy()`)
__global__ void sampler_prob_kernel(const double* __restrict__ w, int64_t n, double total, double* __restrict__ p)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = w[i] / total;
}

// cdf = cumsum(p) in index order, then cdf /= cdf[n - 1]: one thread does the sum (~2 ns per element), the division is parallel
__global__ void sampler_cumsum_kernel(const double* __restrict__ p, int64_t n, double* __restrict__ cdf)
{
    if (blockIdx.x || threadIdx.x) return;
    double acc = 0.0;
    for (int64_t i = 0; i < n; ++i) { acc += p[i]; cdf[i] = acc; }
}
__global__ void sampler_norm_kernel(double* __restrict__ cdf, int64_t n)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const double last = cdf[n - 1]; // (the last entry itself is divided by sampler_norm_last_kernel, after every reader)
    if (i < n - 1) cdf[i] = cdf[i] / last;
}
__global__ void sampler_norm_last_kernel(double* __restrict__ cdf, int64_t n) { cdf[n - 1] = cdf[n - 1] / cdf[n - 1]; }

// idx[j] = searchsorted(cdf, u[j], side="right") = number of entries <= u[j]
__global__ void sampler_search_kernel(const double* __restrict__ cdf, int64_t n, const double* __restrict__ u, int64_t m, int64_t* __restrict__ idx)
{
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= m) return;
    const double x = u[j];
    int64_t lo = 0, hi = n;
    while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        if (cdf[mid] <= x) lo = mid + 1; else hi = mid;
    }
    idx[j] = lo;
}

// first-occurrence filter of one round (numpy: unique(new, return_index=True), indices sorted, take): first[v] = smallest j with new[j] = v
__global__ void sampler_first_kernel(const int64_t* __restrict__ idx, int64_t m, unsigned long long* __restrict__ first)
{
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j < m) atomicMin(first + idx[j], (unsigned long long)j);
}
__global__ void sampler_keep_kernel(const int64_t* __restrict__ idx, int64_t m, const unsigned long long* __restrict__ first, long long* __restrict__ keep)
{
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j < m) keep[j] = first[idx[j]] == (unsigned long long)j ? 1 : 0;
}
// out[base + rank[j]] = idx[j] for kept draws; their probability goes to zero for the next round; first[] is reset
__global__ void sampler_commit_kernel(const int64_t* __restrict__ idx, int64_t m, const long long* __restrict__ keep, const long long* __restrict__ rank,
                                      int64_t base, int64_t* __restrict__ out, double* __restrict__ p, unsigned long long* __restrict__ first)
{
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= m) return;
    if (keep[j]) { out[base + rank[j]] = idx[j]; p[idx[j]] = 0.0; }
    first[idx[j]] = ~0ull;
}

// batch gather: out[j, :] = src[idx[j], :] (rows of the normalised matrices); one warp per row
__global__ void __launch_bounds__(256) gather_rows_kernel(const float* __restrict__ src, int dim, const int64_t* __restrict__ idx, int64_t m, int64_t n_rows,
                                                          float* __restrict__ out, uint32_t* __restrict__ bad)
{
    const int64_t j = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (j >= m) return;
    const int64_t r = idx[j];
    if (r < 0 || r >= n_rows) { if (lane == 0) *bad = 1u; return; }
    for (int c = lane; c < dim; c += 32) out[j * dim + c] = __ldg(src + r * dim + c);
}

} // namespace pg
