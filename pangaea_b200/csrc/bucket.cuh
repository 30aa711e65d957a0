// bucket.cuh - L2-sliced access to the dense k-mer table.
//
// Measured on B200 (tools/microbench.cu, profiles/microbench_r01.txt): random u32 RED /
// gather against a 2 GiB table run at 21 / 40 G ops/s (one 32 B DRAM sector per 4 B
// counter), against a <= 64 MiB table at 190 / 290 G ops/s (L2 hits).  So the dense table
// is cut into slices of 2^kSliceBits counters and every pass over the reads is split in
// two:
//   scatter : walk the packed stream once, compute each window's counter index and
//             radix-partition the indices by slice.  A CTA bins one tile (256 words =
//             8192 windows) in shared memory and appends one contiguous run per slice to
//             that slice's region in HBM (coalesced; one cursor atomic per slice per tile).
//   apply   : sweep the partitioned entries slice by slice; all SMs work on the same
//             64 MiB slice at a time, so the counter updates / look-ups are L2 hits.
// Region sizes come from a histogram pre-pass (ALU only, re-reads 0.375 B/base).
// The stream is processed in segments so the entry buffer stays bounded.
//
// count   entries: u32 = index-in-slice | (run-1) << kSliceBits   (run merging of identical
//                  consecutive windows, see count.cuh)
// feature entries: u64 = index-in-slice | row << 32; the apply pass gathers the count,
//                  bins it (count_kmer.cpp:90-93) and reduces (row, bin) tallies with
//                  warp-aggregated RED into the abundance matrix.
#pragma once
#include "featurize.cuh"
#include "table.cuh"

namespace pg {

constexpr int kSliceBits = 24;           // 2^24 u32 counters = 64 MiB per slice
constexpr int kMaxBuckets = 64;
constexpr int kTileWords = 256;          // one word per thread
constexpr int kTileEntries = kTileWords * 32;

struct BucketGeom {
    int n_buckets;
    uint32_t low_mask; // (1 << kSliceBits) - 1
};

// ---------------------------------------------------------------------------
// windows of one word -> dense indices (registers), shared by hist and scatter
// ---------------------------------------------------------------------------
// COUNT flavour: merges runs of identical consecutive indices; ent[n] = idx, run[n] = length
template <bool MERGE>
__device__ __forceinline__ int word_indices(uint64_t lo, uint64_t hi, uint32_t mlo, uint32_t mhi, int k, uint32_t km,
                                            uint32_t (&ent)[32], uint8_t (&run)[32])
{
    int n = 0;
    const uint64_t wmask = low_mask64(2 * k);
#pragma unroll
    for (int i = 0; i < 32; ++i) {
        const uint32_t mw = __funnelshift_r(mlo, mhi, i);
        if ((mw & km) != km) continue;
        const uint32_t w = (uint32_t)((i ? ((lo >> (2 * i)) | (hi << (64 - 2 * i))) : lo) & wmask);
        const uint32_t idx = dense_index_of_window(w, k);
        if (MERGE && n > 0 && ent[n - 1] == idx) { ++run[n - 1]; continue; }
        ent[n] = idx;
        run[n] = 1;
        ++n;
    }
    return n;
}

// ---------------------------------------------------------------------------
// pre-pass: entries per slice (must mirror the scatter kernels' emission exactly or be
// an upper bound of it)
// ---------------------------------------------------------------------------
template <bool MERGE>
__global__ void __launch_bounds__(256)
bucket_hist_kernel(const uint64_t* __restrict__ codes, const uint32_t* __restrict__ mask, int64_t w0, int64_t w1, int k,
                   unsigned long long* __restrict__ totals, int n_buckets)
{
    __shared__ uint32_t h[kMaxBuckets];
    if (threadIdx.x < kMaxBuckets) h[threadIdx.x] = 0u;
    __syncthreads();
    const uint32_t km = (1u << k) - 1u;
    const uint64_t wmask = low_mask64(2 * k);
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t j = w0 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < w1; j += stride) {
        const uint32_t mlo = __ldg(mask + j);
        if (mlo == 0u) continue;
        const uint32_t mhi = __ldg(mask + j + 1);
        const uint64_t lo = __ldg(codes + j), hi = __ldg(codes + j + 1);
        uint32_t prev = 0xFFFFFFFFu;
#pragma unroll
        for (int i = 0; i < 32; ++i) {
            const uint32_t mw = __funnelshift_r(mlo, mhi, i);
            if ((mw & km) != km) continue;
            const uint32_t w = (uint32_t)((i ? ((lo >> (2 * i)) | (hi << (64 - 2 * i))) : lo) & wmask);
            const uint32_t idx = dense_index_of_window(w, k);
            if (MERGE && idx == prev) continue;
            prev = idx;
            atomicAdd(&h[idx >> kSliceBits], 1u);
        }
    }
    __syncthreads();
    if (threadIdx.x < n_buckets && h[threadIdx.x]) atomicAdd(totals + threadIdx.x, (unsigned long long)h[threadIdx.x]);
}

// totals -> region bases (exclusive scan, padded to 32 entries so runs start sector-aligned);
// cursors reset; bases[n_buckets] = total capacity
__global__ void bucket_scan_kernel(const unsigned long long* __restrict__ totals, unsigned long long* __restrict__ bases,
                                   unsigned long long* __restrict__ cursors, int n_buckets)
{
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        unsigned long long acc = 0;
        for (int b = 0; b < n_buckets; ++b) {
            bases[b] = acc;
            cursors[b] = 0;
            acc += (totals[b] + 31ull) & ~31ull;
        }
        bases[n_buckets] = acc;
    }
}

// ---------------------------------------------------------------------------
// tile-level radix partition in shared memory, then one run per slice to HBM
// ---------------------------------------------------------------------------
struct ScatterSmem {
    uint32_t cnt[kMaxBuckets];      // entries of this tile per slice
    uint32_t base[kMaxBuckets + 1]; // exclusive scan of cnt
    unsigned long long gbase[kMaxBuckets]; // where this tile's run starts in the slice's region
};

// after every thread has done  pos = atomicAdd(&S.cnt[b], 1)  for its entries and a
// __syncthreads(): scan the counts and claim the global runs.  Ends with __syncthreads().
__device__ __forceinline__ void scatter_claim(ScatterSmem& S, int n_buckets, const unsigned long long* __restrict__ bases,
                                              unsigned long long* __restrict__ cursors)
{
    if (threadIdx.x < 32) {
        // n_buckets <= 64: two elements per lane
        const int b0 = threadIdx.x * 2, b1 = b0 + 1;
        const uint32_t c0 = b0 < n_buckets ? S.cnt[b0] : 0u, c1 = b1 < n_buckets ? S.cnt[b1] : 0u;
        uint32_t inc = c0 + c1;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, inc, d);
            if ((int)threadIdx.x >= d) inc += t;
        }
        const uint32_t ex = inc - (c0 + c1);
        S.base[b0] = ex;
        S.base[b1] = ex + c0;
        if (threadIdx.x == 31) S.base[kMaxBuckets] = inc;
        if (c0) S.gbase[b0] = bases[b0] + atomicAdd(cursors + b0, (unsigned long long)c0);
        if (c1) S.gbase[b1] = bases[b1] + atomicAdd(cursors + b1, (unsigned long long)c1);
    }
    __syncthreads();
}

// slice that staged entry e belongs to (binary search over <= 64 prefix sums)
__device__ __forceinline__ int bucket_of_staged(const ScatterSmem& S, uint32_t e)
{
    int lo = 0, hi = kMaxBuckets; // base[lo] <= e < base[hi]
#pragma unroll
    for (int s = 0; s < 6; ++s) {
        const int mid = (lo + hi) >> 1;
        if (S.base[mid] <= e) lo = mid; else hi = mid;
    }
    return lo;
}

// ---- count ----------------------------------------------------------------
__global__ void __launch_bounds__(256)
bucket_scatter_count_kernel(const uint64_t* __restrict__ codes, const uint32_t* __restrict__ maskC, int64_t w0, int64_t w1, int k,
                            BucketGeom geo, const unsigned long long* __restrict__ bases, unsigned long long* __restrict__ cursors,
                            uint32_t* __restrict__ entries)
{
    __shared__ ScatterSmem S;
    __shared__ uint32_t stage[kTileEntries];
    const uint32_t km = (1u << k) - 1u;
    const int64_t n_tiles = (w1 - w0 + kTileWords - 1) / kTileWords;
    for (int64_t t = blockIdx.x; t < n_tiles; t += gridDim.x) {
        if (threadIdx.x < kMaxBuckets) S.cnt[threadIdx.x] = 0u;
        __syncthreads();
        const int64_t j = w0 + t * kTileWords + threadIdx.x;
        uint32_t ent[32];
        uint8_t run[32];
        uint16_t pos[32];
        int n = 0;
        if (j < w1) {
            const uint32_t mlo = __ldg(maskC + j);
            if (mlo != 0u) n = word_indices<true>(__ldg(codes + j), __ldg(codes + j + 1), mlo, __ldg(maskC + j + 1), k, km, ent, run);
        }
#pragma unroll
        for (int i = 0; i < 32; ++i)
            if (i < n) pos[i] = (uint16_t)atomicAdd(&S.cnt[ent[i] >> kSliceBits], 1u);
        __syncthreads();
        scatter_claim(S, geo.n_buckets, bases, cursors);
#pragma unroll
        for (int i = 0; i < 32; ++i)
            if (i < n) stage[S.base[ent[i] >> kSliceBits] + pos[i]] = (ent[i] & geo.low_mask) | ((uint32_t)(run[i] - 1) << kSliceBits);
        __syncthreads();
        const uint32_t total = S.base[kMaxBuckets];
        for (uint32_t e = threadIdx.x; e < total; e += blockDim.x) {
            const int b = bucket_of_staged(S, e);
            __stcs(entries + S.gbase[b] + (e - S.base[b]), stage[e]);
        }
        __syncthreads();
    }
}

__global__ void __launch_bounds__(256)
bucket_apply_count_kernel(const uint32_t* __restrict__ entries, const unsigned long long* __restrict__ bases,
                          const unsigned long long* __restrict__ cursors, BucketGeom geo, uint32_t* __restrict__ table)
{
    __shared__ unsigned long long sb[kMaxBuckets + 1], sf[kMaxBuckets];
    if (threadIdx.x <= geo.n_buckets) sb[threadIdx.x] = bases[threadIdx.x];
    if (threadIdx.x < geo.n_buckets) sf[threadIdx.x] = cursors[threadIdx.x];
    __syncthreads();
    const unsigned long long cap = sb[geo.n_buckets];
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    int b = 0;
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < cap; i += stride) {
        while (b + 1 < geo.n_buckets && sb[b + 1] <= i) ++b; // i only grows
        if (i - sb[b] >= sf[b]) continue;                     // padding between regions
        const uint32_t e = __ldcs(entries + i);
        atomicAdd(table + (((uint32_t)b << kSliceBits) | (e & geo.low_mask)), (e >> kSliceBits) + 1u);
    }
}

// ---- featurize --------------------------------------------------------------
// Same tile walk as featurize_kernel (featurize.cuh): TNF goes to block-private bins, but
// the 15-mer windows are not looked up here - their (index, row) pairs are partitioned by
// slice for bucket_apply_feat_kernel.  Words that straddle a cloud boundary, and clouds
// beyond the TNF slots, take the direct path of featurize.cuh (rare).
template <int DUMMY = 0>
__global__ void __launch_bounds__(256)
bucket_scatter_feat_kernel(const FeatParams P, int64_t seg_w0, int64_t seg_w1, BucketGeom geo,
                           const unsigned long long* __restrict__ bases, unsigned long long* __restrict__ cursors,
                           unsigned long long* __restrict__ entries)
{
    extern __shared__ uint32_t smem[];
    __shared__ ScatterSmem S;
    uint32_t* stage_idx = smem;                                  // [kTileEntries]
    uint32_t* stage_row = smem + kTileEntries;                   // [kTileEntries]
    uint32_t* bins = smem + 2 * kTileEntries;                    // [kSlots][td]  (TNF only)
    uint16_t* lut_s = reinterpret_cast<uint16_t*>(bins + kSlots * P.td);
    const int lut_n = 1 << (2 * P.tnf_k);
    for (int i = threadIdx.x; i < kSlots * P.td; i += blockDim.x) bins[i] = 0u;
    for (int i = threadIdx.x; i < lut_n; i += blockDim.x) lut_s[i] = P.lut[i];
    __syncthreads();

    const int64_t w_begin = seg_w0 + (int64_t)blockIdx.x * P.words_per_cta;
    const int64_t w_end = min(seg_w1, w_begin + P.words_per_cta);
    if (w_begin >= w_end) return;

    const int k = P.table.k;
    const uint32_t km = (1u << k) - 1u, tm = (1u << P.tnf_k) - 1u;
    const uint32_t tmask = (1u << (2 * P.tnf_k)) - 1u;

    int64_t g_cur = advance_group(P.gstart, P.n_groups, 0, w_begin * 32);
    for (int64_t tile = w_begin; tile < w_end; tile += kTileWords) {
        const int64_t tile_end = min(tile + (int64_t)kTileWords, w_end);
        const int64_t p0 = tile * 32, p1 = min(tile_end * 32, P.n_bytes);
        const int64_t g_lo = advance_group(P.gstart, P.n_groups, g_cur, p0);
        const int64_t g_hi = advance_group(P.gstart, P.n_groups, g_lo, p1 - 1);
        g_cur = g_lo;
        const bool single = (g_hi == g_lo);
        if (threadIdx.x < kMaxBuckets) S.cnt[threadIdx.x] = 0u;
        __syncthreads();

        const int64_t j = tile + threadIdx.x;
        uint32_t ent[32];
        uint8_t run[32];
        uint16_t pos[32];
        int n = 0;
        int32_t row = -1;
        uint32_t mlo = 0u;
        if (j < tile_end) mlo = __ldg(P.maskF + j); // maskF here = maskR: dropped clouds and NOFEAT reads already cleared
        if (mlo != 0u) {
            const uint32_t mhi = __ldg(P.maskF + j + 1);
            const uint64_t lo = __ldg(P.codes + j), hi = __ldg(P.codes + j + 1);
            const int64_t q0 = j * 32;
            int64_t g = single ? g_lo : advance_group(P.gstart, P.n_groups, g_lo, q0);
            const bool uniform = single || (g + 1 >= P.n_groups) || (__ldg(P.gstart + g + 1) > q0 + 31);
            if (uniform) {
                row = __ldg(P.row_of_group + g);
                if (row >= 0) {
                    const int64_t slot = g - g_lo;
                    uint32_t* tnf_row = slot < kSlots ? bins + slot * P.td : P.tnf + (int64_t)row * P.td;
#pragma unroll
                    for (int i = 0; i < 32; ++i) {
                        const uint32_t mw = __funnelshift_r(mlo, mhi, i);
                        if ((mw & tm) == tm) {
                            const uint32_t w4 = (uint32_t)(i ? ((lo >> (2 * i)) | (hi << (64 - 2 * i))) : lo) & tmask;
                            atomicAdd(tnf_row + lut_s[w4], 1u);
                        }
                    }
                    n = word_indices<false>(lo, hi, mlo, mhi, k, km, ent, run);
                }
            } else {
                // a cloud boundary inside the word: direct look-ups (featurize.cuh slow path)
                int64_t next_start = __ldg(P.gstart + g + 1);
                int32_t r = __ldg(P.row_of_group + g);
                for (int i = 0; i < 32; ++i) {
                    const int64_t q = q0 + i;
                    while (g + 1 < P.n_groups && q >= next_start) {
                        ++g;
                        next_start = __ldg(P.gstart + g + 1);
                        r = __ldg(P.row_of_group + g);
                    }
                    if (r < 0 || !((mlo >> i) & 1u)) continue;
                    const int64_t slot = g - g_lo;
                    uint32_t* tnf_row = slot < kSlots ? bins + slot * P.td : P.tnf + (int64_t)r * P.td;
                    // abundance goes straight to the global row (no abundance slots in this kernel)
                    FeatParams Q = P;
                    feat_one<kDense>(Q, lo, hi, mlo, mhi, i, P.abd + (int64_t)r * P.vs, tnf_row, lut_s);
                }
            }
        }
#pragma unroll
        for (int i = 0; i < 32; ++i)
            if (i < n) pos[i] = (uint16_t)atomicAdd(&S.cnt[ent[i] >> kSliceBits], 1u);
        __syncthreads();
        scatter_claim(S, geo.n_buckets, bases, cursors);
#pragma unroll
        for (int i = 0; i < 32; ++i)
            if (i < n) {
                const uint32_t at = S.base[ent[i] >> kSliceBits] + pos[i];
                stage_idx[at] = ent[i] & geo.low_mask;
                stage_row[at] = (uint32_t)row;
            }
        __syncthreads();
        const uint32_t total = S.base[kMaxBuckets];
        for (uint32_t e = threadIdx.x; e < total; e += blockDim.x) {
            const int b = bucket_of_staged(S, e);
            __stcs(entries + S.gbase[b] + (e - S.base[b]), (unsigned long long)stage_idx[e] | ((unsigned long long)stage_row[e] << 32));
        }
        // TNF bins: same carry rule as featurize_kernel
        const bool carry = single && tile_end < w_end && (g_lo + 1 >= P.n_groups || __ldg(P.gstart + g_lo + 1) > p1);
        if (!carry) {
            const int64_t ns = min((int64_t)kSlots, g_hi - g_lo + 1);
            for (int64_t s = 0; s < ns; ++s) {
                const int32_t r = __ldg(P.row_of_group + g_lo + s);
                if (r < 0) continue;
                uint32_t* src = bins + s * P.td;
                uint32_t* dst = P.tnf + (int64_t)r * P.td;
                for (int b = threadIdx.x; b < P.td; b += blockDim.x) {
                    const uint32_t v = src[b];
                    if (v) { atomicAdd(dst + b, v); src[b] = 0u; }
                }
            }
        }
        __syncthreads();
    }
}

// sweep the partitioned (index, row) pairs slice by slice: gather (L2 hit), bin, and
// reduce equal (row, bin) pairs inside the warp before the RED
__global__ void __launch_bounds__(256)
bucket_apply_feat_kernel(const unsigned long long* __restrict__ entries, const unsigned long long* __restrict__ bases,
                         const unsigned long long* __restrict__ cursors, BucketGeom geo, const FeatParams P)
{
    __shared__ unsigned long long sb[kMaxBuckets + 1], sf[kMaxBuckets];
    if (threadIdx.x <= geo.n_buckets) sb[threadIdx.x] = bases[threadIdx.x];
    if (threadIdx.x < geo.n_buckets) sf[threadIdx.x] = cursors[threadIdx.x];
    __syncthreads();
    const unsigned long long cap = sb[geo.n_buckets];
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    const unsigned long long cap32 = (cap + 31ull) & ~31ull; // whole warps stay in the loop for the match
    int b = 0;
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < cap32; i += stride) {
        while (b + 1 < geo.n_buckets && sb[b + 1] <= i) ++b;
        uint32_t key = 0xFFFFFFFFu; // (row * vs + bin) would overflow 32 bits for big matrices: keep row and bin apart
        uint32_t row = 0, bin = 0;
        bool live = false;
        if (i < cap && i - sb[b] < sf[b]) {
            const unsigned long long e = __ldcs(entries + i);
            uint32_t c = __ldg(P.table.counts + (((uint32_t)b << kSliceBits) | ((uint32_t)e & geo.low_mask)));
            if (c != 0u) {
                c &= kCountMask;
                if (c < P.clamp) {
                    bin = abd_bin(P, c);
                    row = (uint32_t)(e >> 32);
                    live = true;
                    key = (row << 10) ^ bin; // match hint only; equality is re-checked below
                }
            }
        }
        // warp aggregation: lanes with the same (row, bin) elect one leader
        const uint32_t peers = __match_any_sync(0xffffffffu, key);
        if (live) {
            // the hint can collide: count only true equals among the peers
            uint32_t n_eq = 0, first = 32;
            uint32_t m = peers;
            while (m) {
                const int l = __ffs(m) - 1;
                m &= m - 1;
                const uint32_t r2 = __shfl_sync(peers, row, l), b2 = __shfl_sync(peers, bin, l);
                if (r2 == row && b2 == bin) { ++n_eq; if (first == 32) first = l; }
            }
            if (first == (threadIdx.x & 31)) atomicAdd(P.abd + (int64_t)row * P.vs + bin, n_eq);
        } else if (peers) {
            // dead lanes matched each other on the sentinel; they still have to take part in the shuffles above? no:
            // shuffles are issued with mask = peers, and dead lanes' peers contain only dead lanes.
        }
    }
}

} // namespace pg
