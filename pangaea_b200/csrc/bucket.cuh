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
//   apply   : sweep the partitioned entries slice by slice.  Chunks are handed out IN
//             ORDER from an atomic ticket, so everything in flight on the 148 SMs lies
//             inside one 64 MiB slice (a static grid-stride loop lets CTAs drift apart
//             until several slices are live and L2 thrashes - measured: 30 G RED/s).
// Regions have a fixed capacity (kRegionSlack x the mean): the counter index is scrambled
// by a bijection (kmer.cuh) so slices fill evenly for any base composition.  A run that
// does not fit (pathological repeats) is applied straight to the table by the scatter
// kernel instead - slower, never wrong.  The stream is processed in segments so the
// entry buffer stays bounded.
//
// count   entries: u32 = index-in-slice | (run-1) << kSliceBits  (run merging of identical
//                  adjacent windows: a poly-G tail is one entry, not 100 serialised REDs)
// feature entries: u64 = index-in-slice | row << 32; the apply pass gathers the count,
//                  bins it (count_kmer.cpp:90-93) and reduces equal (row, bin) pairs inside
//                  the warp before one RED into the abundance matrix.
#pragma once
#include "featurize.cuh"
#include "table.cuh"

namespace pg {

constexpr int kSliceBits = 23;           // 2^23 u32 counters = 32 MiB per slice (64 MiB slices get written back 4.5x: profiles/)
constexpr int kMaxBuckets = 64;
constexpr int kTileWords = 256;          // one word per thread
constexpr int kTileEntries = kTileWords * 32;
constexpr int kChunk = 2048;             // entries per apply ticket (8 per thread)
constexpr unsigned long long kOverflowRun = ~0ull;

struct BucketGeom {
    int n_buckets;
    uint32_t low_mask;                   // (1 << kSliceBits) - 1
    unsigned long long cap;              // entries per region (multiple of 32)
};

// device scratch, one per ctx
struct BucketState {
    unsigned long long cursors[kMaxBuckets]; // entries claimed per region (may exceed cap)
    unsigned long long limits[kMaxBuckets];  // first offset that did not fit (cap if none)
    unsigned long long ticket;
};

__global__ void bucket_reset_kernel(BucketState* st, unsigned long long cap)
{
    if (threadIdx.x < kMaxBuckets) { st->cursors[threadIdx.x] = 0ull; st->limits[threadIdx.x] = cap; }
    if (threadIdx.x == 0) st->ticket = 0ull;
}

// ---------------------------------------------------------------------------
// Shared-memory atomics without divergence.  A conditional atomicAdd compiles to
// BSSY / BRA / ATOMS / BSYNC (ptxas will not predicate ATOMS.POPC.INC); with 64 of them per
// thread that is a third of the kernel.  Instead every lane always issues the atomic and
// windows that emit nothing go to a dummy slot (index kMaxBuckets); same-address shared
// atomics are aggregated by the hardware, so the dummy is not a hot spot
// (profiles/microbench_r01.txt: 4 hot bins run faster than 136 uniform ones).
// ---------------------------------------------------------------------------
__device__ __forceinline__ void smem_inc_if(uint32_t* cnt, uint32_t slot, uint32_t on)
{
    atomicAdd(cnt + (on ? slot : (uint32_t)kMaxBuckets), 1u);
}
// if (on) stage[ fill[slot]++ ] = v      (fill already holds the run's base offset)
__device__ __forceinline__ void smem_push_if(uint32_t* fill, uint32_t slot, uint32_t* stage, uint32_t v, uint32_t on)
{
    const uint32_t at = atomicAdd(fill + (on ? slot : (uint32_t)kMaxBuckets), 1u);
    if (on) stage[at] = v;
}
// same, two payload words into two staging arrays (featurize: index, row)
__device__ __forceinline__ void smem_push2_if(uint32_t* fill, uint32_t slot, uint32_t* stage_a, uint32_t* stage_b, uint32_t va, uint32_t vb, uint32_t on)
{
    const uint32_t at = atomicAdd(fill + (on ? slot : (uint32_t)kMaxBuckets), 1u);
    if (on) { stage_a[at] = va; stage_b[at] = vb; }
}

// ---------------------------------------------------------------------------
// bit i of the result: mask bits i .. i+k-1 are all set (the k-mer window starting at base
// i of this word is valid), i in 0..31, k in 1..32.  Binary decomposition of k: ~4 AND/shift
// pairs on the 64-bit mask for all 32 windows at once instead of a funnel shift + compare
// per window.
// ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t window_valid_mask(uint32_t mlo, uint32_t mhi, int k)
{
    uint64_t v = ((uint64_t)mhi << 32) | mlo; // v(i) = AND of mask[i .. i+len-1]
    uint64_t acc = ~0ull;                     // acc(i) = AND of mask[i .. i+off-1]
    int off = 0;
#pragma unroll
    for (int bit = 0, len = 1; bit < 6; ++bit, len <<= 1) {
        if (k & len) { acc &= v >> off; off += len; }
        v &= v >> len;
    }
    return (uint32_t)acc;
}

// ---------------------------------------------------------------------------
// dense indices of the 32 windows of one word, in registers (static indexing only).
// KT > 0 fixes k at compile time (15, the production value: shifts and masks become
// immediates).  Rolling update instead of re-extracting every window: with w_i the
// LSB-first window at base i (its reverse complement is w_i ^ 0xAAAA.., its forward value
// the group-reversed w_i),
//   rc_{i+1}  = (rc_i >> 2)  | ((c ^ 2) << 2(k-1))        c = code of base i + k
//   fwd_{i+1} = ((fwd_i << 2) | c) & mask
// min_diff == 0 iff two neighbouring windows have the same index (homopolymer runs).
// ---------------------------------------------------------------------------
template <int KT, bool MERGE>
__device__ __forceinline__ void word_indices(uint64_t lo, uint64_t hi, int k_rt, uint32_t (&ent)[32], uint32_t& min_diff)
{
    const int k = KT ? KT : k_rt;
    const uint32_t wmask = (uint32_t)low_mask64(2 * k);
    const uint32_t w0 = (uint32_t)lo & wmask;
    uint32_t f = fwd_of_window32(w0, k);
    uint32_t r = w0 ^ (0xAAAAAAAAu & wmask);
    const uint64_t tail = (lo >> (2 * k)) | (hi << (64 - 2 * k)); // codes of bases k .. k+31 (k in 1..16)
    const int top = 2 * (k - 1);
    min_diff = 0xFFFFFFFFu;
#pragma unroll
    for (int i = 0; i < 32; ++i) {
        ent[i] = dense_index_of_pair(f, r, k);
        if (MERGE && i > 0) min_diff = min(min_diff, ent[i] ^ ent[i - 1]);
        const uint32_t c = (uint32_t)(tail >> (2 * i)) & 3u;
        f = ((f << 2) | c) & wmask;
        r = (r >> 2) | ((c ^ 2u) << top);
    }
}

// windows identical to the window one base earlier (both valid): folded into that entry's
// run length, so a poly-G tail is one entry instead of 100 REDs serialised on one L2 address.
// Only evaluated for the rare words where word_indices saw two equal neighbours.
__device__ __forceinline__ uint32_t continuation_mask(const uint32_t (&ent)[32], uint32_t valid)
{
    uint32_t cont = 0u;
#pragma unroll
    for (int i = 1; i < 32; ++i)
        if (ent[i] == ent[i - 1]) cont |= 1u << i;
    return cont & valid & (valid << 1);
}

// run length of the entry opened at base i: 1 + the continuation bits that follow it
__device__ __forceinline__ uint32_t run_length(uint32_t cont, int i)
{
    const uint32_t following = i < 31 ? (cont >> (i + 1)) : 0u;
    return (uint32_t)__ffs(~following); // 1 + number of trailing ones
}

// ---------------------------------------------------------------------------
// tile-level radix partition in shared memory, then one run per slice to HBM
// ---------------------------------------------------------------------------
struct ScatterSmem {
    uint32_t cnt[kMaxBuckets + 1];         // entries of this tile per slice (+1: dummy slot of non-emitting windows)
    uint32_t fill[kMaxBuckets + 1];        // staging cursor per slice (starts at base[] after the claim; +1 dummy)
    uint32_t base[kMaxBuckets];            // exclusive scan of cnt
    unsigned long long gbase[kMaxBuckets]; // where this tile's run starts in the entry buffer
};

// after every thread has added its entries to S.cnt and a __syncthreads(): scan the counts
// and claim the global runs.  Ends with __syncthreads().
__device__ __forceinline__ void scatter_claim(ScatterSmem& S, const BucketGeom& geo, BucketState* st)
{
    if (threadIdx.x < 32) { // n_buckets <= 64: lane owns slices `lane` and `lane + 32` (staging order is free)
        const int b0 = threadIdx.x, b1 = b0 + 32;
        const uint32_t c0 = b0 < geo.n_buckets ? S.cnt[b0] : 0u, c1 = b1 < geo.n_buckets ? S.cnt[b1] : 0u;
        // both cursor atomics are issued before either result is used: one round trip, not two
        const unsigned long long off0 = c0 ? atomicAdd(&st->cursors[b0], (unsigned long long)c0) : 0ull;
        const unsigned long long off1 = c1 ? atomicAdd(&st->cursors[b1], (unsigned long long)c1) : 0ull;
        uint32_t i0 = c0, i1 = c1;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t t0 = __shfl_up_sync(0xffffffffu, i0, d), t1 = __shfl_up_sync(0xffffffffu, i1, d);
            if ((int)threadIdx.x >= d) { i0 += t0; i1 += t1; }
        }
        const uint32_t total0 = __shfl_sync(0xffffffffu, i0, 31);
        S.base[b0] = i0 - c0;
        S.base[b1] = total0 + i1 - c1;
        S.fill[b0] = i0 - c0;          // the staging cursor of a slice starts at its run's offset
        S.fill[b1] = total0 + i1 - c1;
        if (c0) {
            if (off0 + c0 > geo.cap) { atomicMin(&st->limits[b0], off0); S.gbase[b0] = kOverflowRun; }
            else S.gbase[b0] = (unsigned long long)b0 * geo.cap + off0;
        }
        if (c1) {
            if (off1 + c1 > geo.cap) { atomicMin(&st->limits[b1], off1); S.gbase[b1] = kOverflowRun; }
            else S.gbase[b1] = (unsigned long long)b1 * geo.cap + off1;
        }
    }
    __syncthreads();
}

// ---- count ----------------------------------------------------------------
template <int KT>
__global__ void __launch_bounds__(256, 3)
bucket_scatter_count_kernel(const uint64_t* __restrict__ codes, const uint32_t* __restrict__ maskC, int64_t w0, int64_t w1, int k,
                            BucketGeom geo, BucketState* __restrict__ st, uint32_t* __restrict__ entries, uint32_t* __restrict__ table)
{
    __shared__ ScatterSmem S;
    __shared__ uint32_t stage[kTileEntries];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t *const cnt_base = S.cnt, *const fill_base = S.fill, *const stage_base = stage;
    const int64_t n_tiles = (w1 - w0 + kTileWords - 1) / kTileWords;
    for (int64_t t = blockIdx.x; t < n_tiles; t += gridDim.x) {
        if (threadIdx.x <= kMaxBuckets) S.cnt[threadIdx.x] = 0u;
        __syncthreads();
        const int64_t j = w0 + t * kTileWords + threadIdx.x;
        uint32_t ent[32], valid = 0u, cont = 0u;
        if (j < w1) {
            const uint32_t mlo = __ldg(maskC + j);
            if (mlo != 0u) {
                uint32_t min_diff;
                valid = window_valid_mask(mlo, __ldg(maskC + j + 1), KT ? KT : k);
                word_indices<KT, true>(__ldg(codes + j), __ldg(codes + j + 1), k, ent, min_diff);
                if (min_diff == 0u) cont = continuation_mask(ent, valid);
            }
        }
        const uint32_t start = valid & ~cont;
        if (start != 0u) {
#pragma unroll
            for (int i = 0; i < 32; ++i) smem_inc_if(cnt_base, ent[i] >> kSliceBits, start & (1u << i));
        }
        __syncthreads();
        scatter_claim(S, geo, st);
        if (cont == 0u) { // no run in this word (the usual case): plain entries
            if (start != 0u) {
#pragma unroll
                for (int i = 0; i < 32; ++i) smem_push_if(fill_base, ent[i] >> kSliceBits, stage_base, ent[i] & geo.low_mask, start & (1u << i));
            }
        } else {
#pragma unroll
            for (int i = 0; i < 32; ++i)
                smem_push_if(fill_base, ent[i] >> kSliceBits, stage_base, (ent[i] & geo.low_mask) | ((run_length(cont, i) - 1u) << kSliceBits), start & (1u << i));
        }
        __syncthreads();
        for (int b = warp; b < geo.n_buckets; b += 8) { // one warp copies one slice's run: coalesced, no search
            const uint32_t n = S.cnt[b];
            if (!n) continue;
            const uint32_t* src = stage + S.base[b];
            if (S.gbase[b] != kOverflowRun) {
                uint32_t* dst = entries + S.gbase[b];
                for (uint32_t e = lane; e < n; e += 32) __stcs(dst + e, src[e]);
            } else { // region full: apply the run here
                for (uint32_t e = lane; e < n; e += 32)
                    atomicAdd(table + (((uint32_t)b << kSliceBits) | (src[e] & geo.low_mask)), (src[e] >> kSliceBits) + 1u);
            }
        }
        __syncthreads();
    }
}

// ordered tickets over the filled part of every region: ticket -> (slice, offset)
struct ApplySmem {
    unsigned long long fill[kMaxBuckets];           // valid entries per region
    unsigned long long chunk_base[kMaxBuckets + 1]; // prefix sum of chunks per region
    unsigned long long ticket;
};

__device__ __forceinline__ void apply_prologue(ApplySmem& A, const BucketGeom& geo, const BucketState* st)
{
    if (threadIdx.x == 0) {
        unsigned long long acc = 0;
        for (int b = 0; b < geo.n_buckets; ++b) {
            const unsigned long long f = min(min(st->cursors[b], st->limits[b]), geo.cap);
            A.fill[b] = f;
            A.chunk_base[b] = acc;
            acc += (f + kChunk - 1) / kChunk;
        }
        A.chunk_base[geo.n_buckets] = acc;
    }
    __syncthreads();
}

__global__ void __launch_bounds__(256)
bucket_apply_count_kernel(const uint32_t* __restrict__ entries, BucketGeom geo, BucketState* __restrict__ st, uint32_t* __restrict__ table)
{
    __shared__ ApplySmem A;
    apply_prologue(A, geo, st);
    const unsigned long long n_chunks = A.chunk_base[geo.n_buckets];
    int b = 0;
    for (;;) {
        if (threadIdx.x == 0) A.ticket = atomicAdd(&st->ticket, 1ull);
        __syncthreads();
        const unsigned long long c = A.ticket;
        __syncthreads();
        if (c >= n_chunks) break;
        while (A.chunk_base[b + 1] <= c) ++b; // tickets only grow
        const unsigned long long off = (c - A.chunk_base[b]) * kChunk;
        const uint32_t n = (uint32_t)min((unsigned long long)kChunk, A.fill[b] - off);
        const uint32_t* src = entries + (unsigned long long)b * geo.cap + off;
        uint32_t* slice = table + ((size_t)b << kSliceBits);
        uint32_t e[kChunk / 256];
#pragma unroll
        for (int u = 0; u < kChunk / 256; ++u) {
            const uint32_t i = threadIdx.x + 256u * u;
            e[u] = i < n ? __ldcs(src + i) : 0xFFFFFFFFu;
        }
#pragma unroll
        for (int u = 0; u < kChunk / 256; ++u)
            if (threadIdx.x + 256u * u < n) atomicAdd(slice + (e[u] & geo.low_mask), (e[u] >> kSliceBits) + 1u);
    }
}

// ---- featurize --------------------------------------------------------------
// one (row, bin) tally per live lane; lanes with equal keys elect a leader -> one RED
__device__ __forceinline__ void abd_reduce_warp(const FeatParams& P, bool live, unsigned long long key)
{
    const uint32_t live_mask = __ballot_sync(0xffffffffu, live);
    if (live) {
        const uint32_t peers = __match_any_sync(live_mask, key);
        if ((threadIdx.x & 31) == (uint32_t)(__ffs(peers) - 1))
            atomicAdd(P.abd + (int64_t)(key >> 32) * P.vs + (uint32_t)key, (uint32_t)__popc(peers));
    }
}

// count -> (row, bin) key, false when the k-mer is absent or beyond the histogram
__device__ __forceinline__ bool abd_key(const FeatParams& P, uint32_t c, uint32_t row, unsigned long long& key)
{
    if (c == 0u) return false; // absent k-mers are skipped (count_kmer.cpp:87)
    c &= kCountMask;
    if (c >= P.clamp) return false;
    key = ((unsigned long long)row << 32) | abd_bin(P, c);
    return true;
}

// Same tile walk as featurize_kernel (featurize.cuh): TNF goes to block-private bins, but
// the 15-mer windows are not looked up here - their (index, row) pairs are partitioned by
// slice for bucket_apply_feat_kernel.  Words that straddle a cloud boundary take the
// direct path of featurize.cuh (rare: one word per cloud).  P.maskF is the cleaned
// feature mask: dropped clouds and PG_READ_NOFEAT reads are already zero in it.
template <int KT>
__global__ void __launch_bounds__(256, 2)
bucket_scatter_feat_kernel(const FeatParams P, int64_t seg_w0, int64_t seg_w1, BucketGeom geo, BucketState* __restrict__ st,
                           unsigned long long* __restrict__ entries)
{
    extern __shared__ uint32_t smem[];
    __shared__ ScatterSmem S;
    uint32_t* stage_idx = smem;                                  // [kTileEntries]
    uint32_t* stage_row = smem + kTileEntries;                   // [kTileEntries]
    uint32_t* bins = smem + 2 * kTileEntries;                    // [kSlots][td] + 1 dummy word (TNF only)
    uint16_t* lut_s = reinterpret_cast<uint16_t*>(bins + kSlots * P.td + 1);
    const int lut_n = 1 << (2 * P.tnf_k);
    for (int i = threadIdx.x; i < kSlots * P.td; i += blockDim.x) bins[i] = 0u;
    for (int i = threadIdx.x; i < lut_n; i += blockDim.x) lut_s[i] = P.lut[i];
    __syncthreads();

    const int64_t w_begin = seg_w0 + (int64_t)blockIdx.x * P.words_per_cta;
    const int64_t w_end = min(seg_w1, w_begin + P.words_per_cta);
    if (w_begin >= w_end) return;

    const int k = KT ? KT : P.table.k;
    const uint32_t tmask = (1u << (2 * P.tnf_k)) - 1u;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t *const cnt_base = S.cnt, *const fill_base = S.fill, *const stage_idx_base = stage_idx, *const stage_row_base = stage_row;

    int64_t g_cur = advance_group(P.gstart, P.n_groups, 0, w_begin * 32);
    int64_t tile = w_begin;
    while (tile < w_end) {
        const int64_t tile_end = min(tile + (int64_t)kTileWords, w_end);
        const int64_t p0 = tile * 32, p1 = min(tile_end * 32, P.n_bytes);
        const int64_t g_lo = advance_group(P.gstart, P.n_groups, g_cur, p0);
        const int64_t g_hi = advance_group(P.gstart, P.n_groups, g_lo, p1 - 1);
        g_cur = g_lo;
        const bool single = (g_hi == g_lo);
        if (single && __ldg(P.row_of_group + g_lo) < 0) { // dropped cloud: jump to the tile holding its end
            const int64_t nxt = __ldg(P.gstart + g_lo + 1) >> 5;
            const int64_t jump = w_begin + ((nxt - w_begin) / kTileWords) * kTileWords;
            tile = max(tile + (int64_t)kTileWords, jump);
            continue;
        }
        if (threadIdx.x <= kMaxBuckets) S.cnt[threadIdx.x] = 0u;
        __syncthreads();

        const int64_t j = tile + threadIdx.x;
        uint32_t ent[32], start = 0u, cont = 0u;
        int32_t row = -1;
        uint32_t mlo = 0u;
        if (j < tile_end) mlo = __ldg(P.maskF + j);
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
                    const uint32_t tvalid = window_valid_mask(mlo, mhi, P.tnf_k);
                    if (slot < kSlots) { // block-private bins; invalid windows hit the dummy word after the slots
                        uint32_t* tnf_row = bins + slot * P.td;
                        const uint32_t dummy = (uint32_t)(kSlots * P.td - slot * P.td);
#pragma unroll
                        for (int i = 0; i < 32; ++i) {
                            const uint32_t w4 = (uint32_t)(i ? ((lo >> (2 * i)) | (hi << (64 - 2 * i))) : lo) & tmask;
                            atomicAdd(tnf_row + ((tvalid & (1u << i)) ? (uint32_t)lut_s[w4] : dummy), 1u);
                        }
                    } else {
                        uint32_t* tnf_row = P.tnf + (int64_t)row * P.td;
                        for (int i = 0; i < 32; ++i)
                            if (tvalid & (1u << i)) {
                                const uint32_t w4 = (uint32_t)(i ? ((lo >> (2 * i)) | (hi << (64 - 2 * i))) : lo) & tmask;
                                atomicAdd(tnf_row + lut_s[w4], 1u);
                            }
                    }
                    uint32_t unused;
                    start = window_valid_mask(mlo, mhi, k);
                    word_indices<KT, false>(lo, hi, k, ent, unused);
                }
            } else {
                // a cloud boundary inside the word: direct look-ups, abundance straight to the global row
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
                    feat_one<kDense>(P, lo, hi, mlo, mhi, i, P.abd + (int64_t)r * P.vs, tnf_row, lut_s);
                }
            }
        }
        if (start != 0u) {
#pragma unroll
            for (int i = 0; i < 32; ++i) smem_inc_if(cnt_base, ent[i] >> kSliceBits, start & (1u << i));
        }
        __syncthreads();
        scatter_claim(S, geo, st);
        if (start != 0u) {
#pragma unroll
            for (int i = 0; i < 32; ++i)
                smem_push2_if(fill_base, ent[i] >> kSliceBits, stage_idx_base, stage_row_base, ent[i] & geo.low_mask, (uint32_t)row, start & (1u << i));
        }
        __syncthreads();
        for (int b = warp; b < geo.n_buckets; b += 8) {
            const uint32_t n = S.cnt[b];
            if (!n) continue;
            const uint32_t *si = stage_idx + S.base[b], *sr = stage_row + S.base[b];
            if (S.gbase[b] != kOverflowRun) {
                unsigned long long* dst = entries + S.gbase[b];
                for (uint32_t e = lane; e < n; e += 32) __stcs(dst + e, (unsigned long long)si[e] | ((unsigned long long)sr[e] << 32));
            } else { // region full: look the run up here (whole warp stays in the loop for the reduction)
                for (uint32_t e0 = 0; e0 < n; e0 += 32) {
                    const uint32_t e = e0 + lane;
                    unsigned long long key = 0;
                    bool live = false;
                    if (e < n) live = abd_key(P, __ldg(P.table.counts + (((uint32_t)b << kSliceBits) | si[e])), sr[e], key);
                    abd_reduce_warp(P, live, key);
                }
            }
        }
        // TNF bins: same carry rule as featurize_kernel
        const bool carry = single && tile_end < w_end && (g_lo + 1 >= P.n_groups || __ldg(P.gstart + g_lo + 1) > p1);
        if (!carry) {
            const int64_t ns = min((int64_t)kSlots, g_hi - g_lo + 1);
            for (int64_t s2 = 0; s2 < ns; ++s2) {
                const int32_t r = __ldg(P.row_of_group + g_lo + s2);
                if (r < 0) continue;
                uint32_t* src = bins + s2 * P.td;
                uint32_t* dst = P.tnf + (int64_t)r * P.td;
                for (int b = threadIdx.x; b < P.td; b += blockDim.x) {
                    const uint32_t v = src[b];
                    if (v) { atomicAdd(dst + b, v); src[b] = 0u; }
                }
            }
        }
        __syncthreads();
        tile = tile_end;
    }
}

// sweep the partitioned (index, row) pairs slice by slice: gather (L2 hit), bin, reduce
__global__ void __launch_bounds__(256)
bucket_apply_feat_kernel(const unsigned long long* __restrict__ entries, BucketGeom geo, BucketState* __restrict__ st, const FeatParams P)
{
    __shared__ ApplySmem A;
    apply_prologue(A, geo, st);
    const unsigned long long n_chunks = A.chunk_base[geo.n_buckets];
    int b = 0;
    for (;;) {
        if (threadIdx.x == 0) A.ticket = atomicAdd(&st->ticket, 1ull);
        __syncthreads();
        const unsigned long long c = A.ticket;
        __syncthreads();
        if (c >= n_chunks) break;
        while (A.chunk_base[b + 1] <= c) ++b;
        const unsigned long long off = (c - A.chunk_base[b]) * kChunk;
        const uint32_t n = (uint32_t)min((unsigned long long)kChunk, A.fill[b] - off);
        const unsigned long long* src = entries + (unsigned long long)b * geo.cap + off;
        const uint32_t* slice = P.table.counts + ((size_t)b << kSliceBits);
        unsigned long long e[kChunk / 256];
        uint32_t cnt[kChunk / 256];
#pragma unroll
        for (int u = 0; u < kChunk / 256; ++u) {
            const uint32_t i = threadIdx.x + 256u * u;
            e[u] = i < n ? __ldcs(src + i) : 0ull;
        }
#pragma unroll
        for (int u = 0; u < kChunk / 256; ++u)
            cnt[u] = (threadIdx.x + 256u * u < n) ? __ldg(slice + ((uint32_t)e[u] & geo.low_mask)) : 0u;
#pragma unroll
        for (int u = 0; u < kChunk / 256; ++u) {
            unsigned long long key = 0;
            const bool live = abd_key(P, cnt[u], (uint32_t)(e[u] >> 32), key);
            abd_reduce_warp(P, live, key);
        }
    }
}

// featurize works on a copy of maskF from which dropped clouds are removed, so the scatter
// kernel needs no per-word cloud look-up for them
__global__ void __launch_bounds__(256)
clear_dropped_groups_kernel(const int64_t* __restrict__ gstart, const int32_t* __restrict__ row_of_group, int64_t n_groups,
                            uint32_t* __restrict__ maskR)
{
    for (int64_t g = blockIdx.x; g < n_groups; g += gridDim.x) {
        if (__ldg(row_of_group + g) >= 0) continue;
        const int64_t lo = gstart[g], hi = gstart[g + 1];
        if (lo >= hi) continue;
        const int64_t wlo = lo >> 5, whi = (hi - 1) >> 5;
        for (int64_t w = wlo + threadIdx.x; w <= whi; w += blockDim.x) {
            const int64_t a = max(lo, w << 5), b = min(hi, (w + 1) << 5);
            const uint32_t bits = (uint32_t)(((1ull << (b - a)) - 1ull) << (a - (w << 5)));
            if (bits == 0xFFFFFFFFu) maskR[w] = 0u; else atomicAnd(&maskR[w], ~bits);
        }
    }
}

} // namespace pg
