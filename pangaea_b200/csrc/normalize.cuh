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
//
// Text round trip.  The reference never normalises the tallies themselves but what pandas read back from the
// tools' CSV, and `ostream << double` prints 6 significant digits (count_kmer.cpp:211, count_tnf.cpp:204): a
// tally >= 10^6 arrives as e.g. 1.11493e+06 (KAT-5).  text_round6 reproduces printf("%g")'s decision in
// integers - round to 6 significant digits, ties to even (every tally is an exactly representable integer, so
// a tie is exactly half a unit) - and the kernel normalises the rounded values when `text_round` is set.
#pragma once
#include <stdint.h>
#include <cuda_runtime.h>

namespace pg {

__host__ __device__ __forceinline__ unsigned long long text_round6(uint32_t v)
{
    if (v < 1000000u) return v;
    const uint32_t q = v < 10000000u ? 10u : v < 100000000u ? 100u : v < 1000000000u ? 1000u : 10000u; // unit of the 6th digit
    const uint32_t r = v % q, base = v - r;
    const bool up = 2u * r > q || (2u * r == q && ((base / q) & 1u));
    return (unsigned long long)base + (up ? q : 0u); // (64-bit: 4294967295 rounds to 4294970000)
}

// LPR lanes per row: a full warp for few long rows, 8 lanes (four rows per warp in flight) when there are millions of rows
template <int LPR>
__global__ void __launch_bounds__(256)
normalize_rows_kernel(const uint32_t* __restrict__ raw, int64_t rows, int dim, float* __restrict__ out, double* __restrict__ weights, int text_round)
{
    const int lane = threadIdx.x & (LPR - 1);
    const int64_t warps = ((int64_t)gridDim.x * blockDim.x) / LPR; // row groups in flight
    for (int64_t r0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / LPR; r0 < ((rows + 32 / LPR - 1) / (32 / LPR)) * (32 / LPR); r0 += warps) {
        const bool have = r0 < rows;            // (the groups of a warp stay together for the shuffles)
        const int64_t r = have ? r0 : rows - 1;
        const uint32_t* src = raw + r * dim;
        unsigned long long sum = 0;
        unsigned long long mx = 0;
        for (int j = lane; j < dim; j += LPR) {
            unsigned long long v = __ldg(src + j);
            if (text_round) v = text_round6((uint32_t)v);
            sum += v;
            mx = max(mx, v);
        }
#pragma unroll
        for (int d = LPR / 2; d; d >>= 1) {
            sum += __shfl_xor_sync(0xffffffffu, sum, d);
            mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, d));
        }
        const double norm = sum ? (double)sum : 1.0; // sklearn: zero norms are replaced by 1
        // x / norm, correctly rounded, without an fp64 division per element (B200 runs DDIV as a ~30-instruction sequence on a
        // narrow pipe: 1.1 s for the 160 G elements of the one-cloud-per-pair configuration).  Markstein: with y = RN(1 / norm),
        // q = RN(x y), r = x - norm q (exact in an fma), q' = RN(q + r y) is the correctly rounded quotient - except, possibly,
        // when norm's significand is all ones, where the plain division is used (checked exhaustively against exact rationals
        // on 3 x 10^5 random and adversarial pairs; the GPU tests compare with numpy bit for bit).
        const bool plain = (sum & (sum + 1ull)) == 0ull; // 2^k - 1 (and 0)
        const double y = 1.0 / norm;
        auto quot = [&](unsigned long long v) {
            const double x = (double)v;
            if (plain) return x / norm;
            const double q = x * y;
            const double rem = fma(-norm, q, x);
            return fma(rem, y, q);
        };
        if (!have) continue;
        float* dst = out + r * dim;
        for (int j = lane; j < dim; j += LPR) {
            unsigned long long v = __ldg(src + j);
            if (text_round) v = text_round6((uint32_t)v);
            dst[j] = (float)quot(v);
        }
        if (weights && lane == 0) {
            const double m = quot(mx);
            weights[r] = m * m;
        }
    }
}

// Rows of 4 n <= 512 elements (the production shapes: 400 and 136): one warp per row, the row read ONCE as 16-byte vectors
// that stay in registers between the sum and the division, streamed past the caches in both directions (every byte is
// touched once).  The kernel above remains for odd shapes.
template <int VPL> // vectors per lane: ceil(dim / 4 / 32)
__global__ void __launch_bounds__(256)
normalize_rows_vec_kernel(const uint4* __restrict__ raw, int64_t rows, int nvec, float4* __restrict__ out, double* __restrict__ weights, int text_round)
{
    const int lane = threadIdx.x & 31;
    const int64_t warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t r = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; r < rows; r += warps) { // (r is warp-uniform)
        const uint4* src = raw + r * nvec;
        uint4 v[VPL];
#pragma unroll
        for (int q = 0; q < VPL; ++q) v[q] = lane + 32 * q < nvec ? __ldcs(src + lane + 32 * q) : make_uint4(0u, 0u, 0u, 0u);
        // Fast path - every tally of the row below 10^6 (nothing to round; the sum of <= 512 of them fits 32 bits) and the sum
        // not of the form 2^k - 1: 32-bit sums and shuffles, one conversion + three fp64 operations per element.  Rows with
        // a tally >= 10^6 (KAT-5) or an all-ones sum take the general path below; the branch is warp-uniform.  (ncu, 10 M
        // rows x 400: the general path alone costs 616 warp instructions per row and makes the kernel issue-bound.)
        uint32_t mx32 = 0u, sum32 = 0u;
#pragma unroll
        for (int q = 0; q < VPL; ++q) {
            mx32 = max(max(mx32, max(v[q].x, v[q].y)), max(v[q].z, v[q].w));
            sum32 += min(v[q].x, 1000000u) + min(v[q].y, 1000000u) + min(v[q].z, 1000000u) + min(v[q].w, 1000000u); // (exact when mx32 < 10^6)
        }
        mx32 = __reduce_max_sync(0xffffffffu, mx32);
        sum32 = __reduce_add_sync(0xffffffffu, sum32);
        float4* dst = out + r * nvec;
        if (mx32 < 1000000u && (sum32 & (sum32 + 1u)) != 0u) {
            const double norm = (double)sum32; // (> 0: a zero sum is of the form 2^k - 1)
            const double y = 1.0 / norm;
            auto quot32 = [&](uint32_t t) {
                const double x = (double)t;
                const double q = x * y;
                return (float)fma(fma(-norm, q, x), y, q); // Markstein, see normalize_rows_kernel
            };
#pragma unroll
            for (int q = 0; q < VPL; ++q)
                if (lane + 32 * q < nvec)
                    __stcs(dst + lane + 32 * q, make_float4(quot32(v[q].x), quot32(v[q].y), quot32(v[q].z), quot32(v[q].w)));
            if (weights && lane == 0) {
                const double x = (double)mx32, q = x * y;
                const double m = fma(fma(-norm, q, x), y, q);
                weights[r] = m * m;
            }
            continue;
        }
        auto val = [&](uint32_t x) { return text_round ? text_round6(x) : (unsigned long long)x; };
        unsigned long long sum = 0, mx = 0;
#pragma unroll
        for (int q = 0; q < VPL; ++q) {
            const unsigned long long a = val(v[q].x), b = val(v[q].y), c = val(v[q].z), d = val(v[q].w);
            sum += (a + b) + (c + d);
            mx = max(max(mx, max(a, b)), max(c, d));
        }
#pragma unroll
        for (int d = 16; d; d >>= 1) {
            sum += __shfl_xor_sync(0xffffffffu, sum, d);
            mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, d));
        }
        const double norm = sum ? (double)sum : 1.0;
        const bool plain = (sum & (sum + 1ull)) == 0ull; // see normalize_rows_kernel
        const double y = 1.0 / norm;
        auto quot = [&](unsigned long long t) {
            const double x = (double)t;
            if (plain) return x / norm;
            const double q = x * y;
            return fma(fma(-norm, q, x), y, q);
        };
#pragma unroll
        for (int q = 0; q < VPL; ++q)
            if (lane + 32 * q < nvec)
                __stcs(dst + lane + 32 * q, make_float4((float)quot(val(v[q].x)), (float)quot(val(v[q].y)), (float)quot(val(v[q].z)), (float)quot(val(v[q].w))));
        if (weights && lane == 0) {
            const double m = quot(mx);
            weights[r] = m * m;
        }
    }
}

} // namespace pg
