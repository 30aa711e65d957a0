// scan.cuh - barcode grouping on device: change flags -> cloud boundaries -> rows.
//
// The reference streams the barcode-sorted file once and flushes a cloud whenever
// the barcode of the pair just appended differs from last_barcode
// (count_kmer.cpp:236-282 interleaved, :181-233 paired).  Stated as data-parallel
// primitives: cloud(r) = exclusive_prefix_sum(PG_READ_CHANGE)[r]; cloud g starts at
// the byte after the read that carried the g-th flag; its length is the difference
// of two starts (the reference's reads_seq.size(), separators included) minus the
// bytes of PG_READ_NOFEAT reads; it is emitted iff its label is non-empty and
// length > min_length (count_kmer.cpp:62); rows are numbered by a second exclusive
// scan over the emit flags, which keeps file order (count_kmer.cpp:283-292).
//
// Three-kernel scan (block counts -> scan of counts -> scatter); flags are 1 B per
// read, so this stage moves ~9 B per read against ~100+ B of bases: not a hot spot.
#pragma once
#include <stdint.h>
#include <cuda_runtime.h>

namespace pg {

constexpr int kScanThreads = 256;
constexpr int kScanItems = 16; // flags per thread
constexpr int kScanTile = kScanThreads * kScanItems;

// exclusive scan of one int per thread across the block; returns block total in `total`
__device__ __forceinline__ int block_exclusive_scan(int v, int& total)
{
    __shared__ int warp_sums[kScanThreads / 32];
    int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    int inc = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, inc, d);
        if (lane >= d) inc += t;
    }
    if (lane == 31) warp_sums[wid] = inc;
    __syncthreads();
    if (wid == 0) {
        int s = lane < kScanThreads / 32 ? warp_sums[lane] : 0;
#pragma unroll
        for (int d = 1; d < kScanThreads / 32; d <<= 1) {
            int t = __shfl_up_sync(0xffffffffu, s, d);
            if (lane >= d) s += t;
        }
        if (lane < kScanThreads / 32) warp_sums[lane] = s; // inclusive
    }
    __syncthreads();
    int warp_off = wid ? warp_sums[wid - 1] : 0;
    total = warp_sums[kScanThreads / 32 - 1];
    __syncthreads(); // warp_sums is reused by the next call
    return warp_off + inc - v;
}

// pass 1: number of set `bit` flags per tile
__global__ void __launch_bounds__(kScanThreads)
flag_count_kernel(const uint8_t* __restrict__ flags, int64_t n, uint32_t bit, int32_t* __restrict__ tile_counts)
{
    int64_t base = (int64_t)blockIdx.x * kScanTile + (int64_t)threadIdx.x * kScanItems;
    int c = 0;
#pragma unroll
    for (int i = 0; i < kScanItems; ++i)
        if (base + i < n) c += (flags[base + i] & bit) ? 1 : 0;
    int total;
    block_exclusive_scan(c, total);
    if (threadIdx.x == 0) tile_counts[blockIdx.x] = total;
}

// pass 2: single block, in-place exclusive scan of the tile counts (64-bit carry kept
// in int64 total; per-tile offsets fit int32 because clouds are numbered in int32)
__global__ void __launch_bounds__(kScanThreads)
tile_scan_kernel(int32_t* __restrict__ tile_counts, int64_t n_tiles, int64_t* __restrict__ total_out)
{
    __shared__ int carry_s;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    for (int64_t base = 0; base < n_tiles; base += kScanThreads) {
        int64_t i = base + threadIdx.x;
        int v = i < n_tiles ? tile_counts[i] : 0;
        int total;
        int ex = block_exclusive_scan(v, total);
        int carry = carry_s;
        if (i < n_tiles) tile_counts[i] = carry + ex;
        __syncthreads();
        if (threadIdx.x == 0) carry_s = carry + total;
        __syncthreads();
    }
    if (threadIdx.x == 0) *total_out = carry_s;
}

// pass 3 (reads): cloud starts + bytes to subtract for NOFEAT reads.
// gstart has n_groups + 1 entries; gstart[0] = 0 and gstart[n_groups] = n_bytes are
// written by thread 0 of block 0.
__global__ void __launch_bounds__(kScanThreads)
group_starts_kernel(const uint8_t* __restrict__ flags, const int64_t* __restrict__ read_off, int64_t n_reads,
                    const int32_t* __restrict__ tile_off, int64_t n_groups, int64_t* __restrict__ gstart,
                    unsigned long long* __restrict__ nofeat_len)
{
    int64_t base = (int64_t)blockIdx.x * kScanTile + (int64_t)threadIdx.x * kScanItems;
    uint8_t f[kScanItems];
    int c = 0;
#pragma unroll
    for (int i = 0; i < kScanItems; ++i) {
        f[i] = base + i < n_reads ? flags[base + i] : 0;
        c += f[i] & 1;
    }
    int total;
    int g = tile_off[blockIdx.x] + block_exclusive_scan(c, total);
#pragma unroll
    for (int i = 0; i < kScanItems; ++i) {
        int64_t r = base + i;
        if (r >= n_reads) break;
        if (f[i] & 2) atomicAdd(&nofeat_len[g], (unsigned long long)(read_off[r + 1] - read_off[r]));
        if (f[i] & 1) {
            ++g;
            if (g < n_groups) gstart[g] = read_off[r + 1];
        }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        gstart[0] = 0;
        gstart[n_groups] = read_off[n_reads];
    }
}

// emit flag per cloud: label non-empty and length > min_length
__global__ void group_emit_kernel(const int64_t* __restrict__ gstart, const unsigned long long* __restrict__ nofeat_len,
                                  const uint8_t* __restrict__ group_keep, int64_t n_groups, int64_t min_length,
                                  uint8_t* __restrict__ emit)
{
    int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n_groups) return;
    int64_t len = gstart[g + 1] - gstart[g] - (int64_t)nofeat_len[g];
    // `reads_seq.size() <= mlen` is an unsigned comparison in the reference: a negative
    // -l converts to a huge size_t and drops every cloud.
    bool ok = group_keep[g] && min_length >= 0 && len > min_length;
    emit[g] = ok ? 1 : 0;
}

// pass 3 (clouds): row numbers in file order
__global__ void __launch_bounds__(kScanThreads)
row_assign_kernel(const uint8_t* __restrict__ emit, int64_t n_groups, const int32_t* __restrict__ tile_off,
                  int32_t* __restrict__ row_of_group, int32_t* __restrict__ group_of_row, int32_t* __restrict__ row_lb)
{
    int64_t base = (int64_t)blockIdx.x * kScanTile + (int64_t)threadIdx.x * kScanItems;
    uint8_t f[kScanItems];
    int c = 0;
#pragma unroll
    for (int i = 0; i < kScanItems; ++i) {
        f[i] = base + i < n_groups ? emit[base + i] : 0;
        c += f[i];
    }
    int total;
    int row = tile_off[blockIdx.x] + block_exclusive_scan(c, total);
#pragma unroll
    for (int i = 0; i < kScanItems; ++i) {
        int64_t g = base + i;
        if (g >= n_groups) break;
        row_lb[g] = row; // rows emitted before cloud g: a lower bound for the row of every later cloud
        if (f[i]) {
            row_of_group[g] = row;
            group_of_row[row] = (int32_t)g;
            ++row;
        } else {
            row_of_group[g] = -1;
        }
    }
}

// ---------------------------------------------------------------------------
// word -> cloud map for the streaming kernels (bucket.cuh, tnf.cuh): wg[j] = cloud of base
// 32 j, with kWordMixed set when another cloud starts inside the word (one word per cloud:
// those take a per-position slow path).  Filled per cloud, so the hot kernels never search
// gstart.  BIG = false: one warp per cloud, clouds of more than kBigCloudWords words are left
// to the BIG = true launch (one block per cloud; the unbarcoded tail can be millions of words).
// ---------------------------------------------------------------------------
constexpr uint32_t kWordMixed = 0x80000000u;
constexpr int64_t kBigCloudWords = 8192;

// !BIG: one warp per cloud; clouds of more than kBigCloudWords words are only noted in big_list (at most n_words /
// kBigCloudWords + 1 of them fit a stream).  BIG: one CTA per listed cloud.  (The BIG pass used to walk ALL clouds looking for
// big ones: 5 900 dependent loads per CTA with one cloud per read pair - 3 ms of the 6.7 ms grouping of a 7 M-cloud batch.)
template <bool BIG>
__global__ void __launch_bounds__(256)
word_groups_kernel(const int64_t* __restrict__ gstart, int64_t n_groups, int64_t n_bytes, uint32_t* __restrict__ wg, uint32_t* __restrict__ big_list,
                   uint32_t* __restrict__ big_n)
{
    const int lane = BIG ? threadIdx.x : (threadIdx.x & 31);
    const int step = BIG ? blockDim.x : 32;
    const int64_t first = BIG ? blockIdx.x : (((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    const int64_t stride = BIG ? gridDim.x : (((int64_t)gridDim.x * blockDim.x) >> 5);
    const int64_t n_items = BIG ? (int64_t)*big_n : n_groups;
    for (int64_t i = first; i < n_items; i += stride) {
        const int64_t g = BIG ? (int64_t)big_list[i] : i;
        const int64_t lo = __ldg(gstart + g), hi = __ldg(gstart + g + 1);
        if (lo >= hi) continue;
        const int64_t w_first = (lo + 31) >> 5, w_last = (hi - 1) >> 5; // words whose first base lies in [lo, hi)
        if (!BIG && (w_last - w_first + 1) > kBigCloudWords) {
            if (lane == 0) big_list[atomicAdd(big_n, 1u)] = (uint32_t)g;
            continue;
        }
        for (int64_t w = w_first + lane; w <= w_last; w += step) {
            const bool mixed = (w == w_last) && hi < min((w + 1) << 5, n_bytes);
            wg[w] = (uint32_t)g | (mixed ? kWordMixed : 0u);
        }
    }
}

} // namespace pg
