// bucket.cuh - L2-sliced access to the dense k-mer table.
//
// Measured on B200 (tools/microbench.cu, profiles/microbench_r01.txt): random u32 RED /
// gather against a 2 GiB table run at 21 / 40 G ops/s (one 32 B DRAM sector per 4 B
// counter), against a <= 64 MiB table at 190 / 290 G ops/s (L2 hits).  So the dense table
// is cut into slices of 2^kSliceBits counters and every pass over the reads is split in
// two:
//   scatter : walk the packed stream once, compute each window's counter index and
//             radix-partition the indices by slice.  A CTA bins one tile (one 32-base word
//             per thread) in shared memory - ONE returning shared atomic per window hands out
//             the slot in a fixed-capacity staging row of the window's slice - and appends one
//             contiguous run per slice to that slice's region in HBM (coalesced; one cursor
//             atomic per slice per tile).
//   apply   : sweep the partitioned entries slice by slice.  Chunks are handed out IN
//             ORDER from an atomic ticket, so everything in flight on the 148 SMs lies
//             inside one 32 MiB slice (a static grid-stride loop lets CTAs drift apart
//             until several slices are live and L2 thrashes - measured: 30 G RED/s).
// Regions have a fixed capacity (region_slack x the mean): the counter index is scrambled
// by a bijection (kmer.cuh) so slices fill evenly for any base composition.  Two things can
// overflow, both handled without losing a window - slower, never wrong:
//   * a staging row (a tile whose windows pile into one slice: tandem repeats) - the tile's
//     windows of that slice are applied straight to the table by a second walk of the tile;
//   * a region (pathological repeats) - the run is applied straight from the staging row.
// The stream is processed in segments so the entry buffer stays bounded.
//
// Entry (u32), from y = 8 * (scrambled class id): bits 3..25 = index inside the slice; the other 9 bits (top 6 + low 3)
// depend on the kind of partition (kScatter* below):
//   shared  : ONE partition serves both passes (the default).  The 9 bits are the window's cloud as a delta against the
//             first cloud of the tile (510 = window that is counted but not featurized, 511 = padding).  Runs are padded
//             to a multiple of 32 entries with kInvalidEntry and every aligned group of 32 entries has its base cloud in
//             a side array (4 B per 128 B of entries).  pg_count keeps the buffer; pg_featurize sweeps it again.
//   count   : entries for the count pass alone: the whole y (top 6 bits = slice, low 3 bits zero), runs padded to 4.
//   feature : entries for the featurize pass alone: as shared, but the delta is a ROW delta (rows are known by then).
// The count pass partitions its entries once more (count2.cuh); the featurize apply gathers the count (L2 hit), bins it
// (count_kmer.cpp:90-93), maps cloud -> row and reduces equal (row, bin) pairs inside the warp before one RED into the
// abundance matrix.
#pragma once
#include <type_traits>
#include "featurize.cuh"
#include "scan.cuh"
#include "table.cuh"

namespace pg {

constexpr int kSliceBits = 23;           // 2^23 u32 counters = 32 MiB per slice (64 MiB slices get written back 4.5x: profiles/)
constexpr int kMaxBuckets = 64;
#ifndef PG_CHUNK
#define PG_CHUNK 1024
#endif
constexpr int kChunk = PG_CHUNK;         // entries per block-wide ticket of the L2 count apply (PG_CHUNK / 256 per thread)
#ifndef PG_FEAT_CHUNK
#define PG_FEAT_CHUNK 2048
#endif
constexpr int kFeatChunk = PG_FEAT_CHUNK; // entries per warp-wide ticket of the featurize apply (multiple of 128)
constexpr unsigned long long kOverflowRun = ~0ull;
constexpr uint32_t kInvalidEntry = 0xFFFFFFFFu; // never a real entry: count entries have low bits 000, deltas stop at 510
constexpr uint32_t kEntryIndexBits = 0x03FFFFF8u;
constexpr int kMaxRowDelta = 509;      // 510 = count-only window of the shared partition, 511 = padding

// tile shape of the scatter kernels: one word per thread.  -DPG_FEAT_THREADS=256 builds the feature-format kernels with 256-word
// tiles (four CTAs per SM); measured (profiles/bench_r02_scatter_ab.txt): scatter 19.2 -> 18.3 ms, but runs half as long cost the
// second partition level and the collect kernel more than that.
#ifndef PG_FEAT_THREADS
#define PG_FEAT_THREADS 512
#endif
template <bool FEAT>
struct ScatterCfg {
    static constexpr int kThreads = FEAT ? PG_FEAT_THREADS : 256;
    static constexpr int kTileWords = kThreads;
    // staging slots per slice per tile: mean = 32 * kThreads / 64 (128 / 256) + > 5 sigma of the binomial
    static constexpr int kStageCap = kThreads == 512 ? 352 : 192;
    // words per staging row: + 4 so that consecutive rows start 4 banks apart (all rows fill at the same pace; with
    // a stride that is a multiple of 32 words the lanes of a store would crowd the banks of the current fill level)
    static constexpr int kStageStride = kStageCap + 4;
    static constexpr int kMinCtas = kThreads == 512 ? 2 : 4;
};
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
// y = 8 * dense index of each of the 32 windows that start in one word, handed to fn(i, y).
// KT = 15 (the production k) is the tuned path, ~11 integer ops per window:
//   u_i  = stream bits [2i, 2i+32)               one funnel shift (compile-time amount)
//   w_i  = u_i & (2^30 - 1)                      LSB-first window = reverse half of its reverse complement
//   f_i  = ((f_{i-1} << 2) | top group of w_i)   forward value, rolling
//   x    = bit 15 of w_i set ? w_i ^ 0x2AAAAAAA : f_i
//          (bit 15 = high bit of the middle base, the same bit in w and f; it differs between a
//           15-mer and its reverse complement, so "the one with the bit clear" is the class
//           representative - kmer.cuh:dense_index_of_pair)
//   id   = x - ((x >> 16) << 15)                 squeeze the clear bit out
//   y    = 8 id A mod 2^32 = x (8A) - (x >> 16)(8A << 15): two IMADs; y >> 3 = (id A) mod 2^29 is the
//          scrambled index of kmer.cuh, y >> 26 its slice.
// KT = 0: any k <= 15 at run time through kmer.cuh (same results, not tuned).
// ---------------------------------------------------------------------------
// windows I0 .. I1 - 1 of the word (the rolling forward value restarts at I0)
template <int KT, int I0, int I1, class Fn>
__device__ __forceinline__ void for_window_range(uint64_t lo, uint64_t hi, int k_rt, Fn&& fn)
{
    const uint32_t s0 = (uint32_t)lo, s1 = (uint32_t)(lo >> 32), s2 = (uint32_t)hi;
    auto stream32 = [&](int i) { return i == 0 ? s0 : (i < 16 ? __funnelshift_r(s0, s1, 2 * i) : (i == 16 ? s1 : __funnelshift_r(s1, s2, 2 * i - 32))); };
    if constexpr (KT == 15) {
        constexpr uint32_t M = 0x3FFFFFFFu, C = 0x2AAAAAAAu;
        constexpr uint32_t A8 = kMixA * 8u, K2 = 0u - (A8 << 15);
        uint32_t f = fwd_of_window32(stream32(I0) & M, 15);
#pragma unroll
        for (int i = I0; i < I1; ++i) {
            const uint32_t w = stream32(i) & M;
            if (i > I0) f = ((f << 2) | (w >> 28)) & M;
            const uint32_t x = (w & 0x8000u) ? (w ^ C) : f;
            fn(i, x * A8 + (x >> 16) * K2);
        }
    } else {
        const int k = k_rt;
        const uint32_t wmask = (uint32_t)low_mask64(2 * k);
        const int top = 2 * (k - 1);
        uint32_t f = fwd_of_window32(stream32(I0) & wmask, k);
#pragma unroll
        for (int i = I0; i < I1; ++i) {
            const uint32_t w = stream32(i) & wmask;
            if (i > I0) f = ((f << 2) | ((w >> top) & 3u)) & wmask;
            fn(i, dense_index_of_pair(f, w ^ (0xAAAAAAAAu & wmask), k) << 3);
        }
    }
}

template <int KT, class Fn>
__device__ __forceinline__ void for_each_window(uint64_t lo, uint64_t hi, int k_rt, Fn&& fn)
{
    for_window_range<KT, 0, 32>(lo, hi, k_rt, fn);
}

// same value for one window, not unrolled (slow paths)
__device__ __forceinline__ uint32_t window_y(uint64_t lo, uint64_t hi, int i, int k)
{
    const uint64_t win = i ? ((lo >> (2 * i)) | (hi << (64 - 2 * i))) : lo;
    return dense_index_of_window((uint32_t)win & (uint32_t)low_mask64(2 * k), k) << 3;
}

// ---- abundance tallies -------------------------------------------------------
// count -> (row, bin) key, false when the k-mer is absent or beyond the histogram
__device__ __forceinline__ bool abd_key(const FeatParams& P, uint32_t c, uint32_t row, unsigned long long& key)
{
    if (c == 0u) return false; // absent k-mers are skipped (count_kmer.cpp:87)
    c &= kCountMask;
    if (c >= P.clamp) return false;
    key = ((unsigned long long)row << 32) | abd_bin(P, c);
    return true;
}

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

// look one window up directly (slow paths: no warp cooperation assumed)
__device__ __forceinline__ void abd_direct(const FeatParams& P, uint32_t y, uint32_t row)
{
    unsigned long long key;
    if (abd_key(P, __ldg(P.table.counts + (y >> 3)), row, key))
        atomicAdd(P.abd + (int64_t)(key >> 32) * P.vs + (uint32_t)key, 1u);
}

// a word with a cloud boundary inside it (kWordMixed), or whose row is too far from the tile's
// base row: resolve the cloud per position and look the windows up directly
__device__ __noinline__ void feat_word_direct(const FeatParams& P, int64_t j, uint32_t g0, uint64_t lo, uint64_t hi, uint32_t valid)
{
    const int k = P.table.k;
    int64_t g = g0;
    int64_t next_start = __ldg(P.gstart + g + 1);
    int32_t row = __ldg(P.row_of_group + g);
    for (int i = 0; i < 32; ++i) {
        const int64_t q = j * 32 + i;
        while (g + 1 < P.n_groups && q >= next_start) {
            ++g;
            next_start = __ldg(P.gstart + g + 1);
            row = __ldg(P.row_of_group + g);
        }
        if (row < 0 || !((valid >> i) & 1u)) continue;
        abd_direct(P, window_y(lo, hi, i, k), (uint32_t)row);
    }
}

// ---------------------------------------------------------------------------
// scatter: tile-level radix partition in shared memory, then one run per slice to HBM
// ---------------------------------------------------------------------------
template <bool FEAT>
struct ScatterSmem {
    uint32_t cnt[2][kMaxBuckets + 8];      // windows of the tile per slice (+ dummy slot of non-emitting windows); double-buffered:
                                           // the counters of tile t are still read (overflow check) while those of t + 1 are zeroed
    alignas(16) uint32_t stage[(kMaxBuckets + 1) * ScatterCfg<FEAT>::kStageStride];
};

// ---- bulk copies (TMA engine, no tensor map): shared -> global, completion tracked per thread in bulk groups ----
__device__ __forceinline__ void bulk_store_fence() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ unsigned long long bulk_policy_evict_first()
{
    unsigned long long pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
// bytes: multiple of 16; both addresses 16 B aligned
__device__ __forceinline__ void bulk_store(void* gdst, const void* ssrc, uint32_t bytes, unsigned long long policy)
{
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;"
                 :: "l"(gdst), "r"((uint32_t)__cvta_generic_to_shared(ssrc)), "r"(bytes), "l"(policy) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); } // sources may be overwritten
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

struct ScatterParams {
    const uint64_t* codes;
    const uint32_t* mask;      // count / shared: maskC; feature: maskF (NOFEAT reads cleared)
    int64_t w0, w1;            // segment of the stream, in words
    int k;
    BucketGeom geo;
    BucketState* st;
    uint32_t* entries;
    int32_t* meta;             // feature / shared: base row / base cloud of every aligned group of 32 entries
    uint32_t* table;           // count / shared: the dense counters (overflow paths)
    uint32_t* lost;            // shared: set to 1 when an overflow path applied counts directly - the entries of those
                               // windows are missing from the buffer, so the featurize pass must not reuse it
    uint32_t* sat;             // count / shared: raised when a direct add takes a counter to bit 31 (table.cuh: saturation)
    uint32_t run_pad;          // runs are padded with kInvalidEntry to a multiple of this many entries: 4 (16 B, what the copy engine
                               // needs) or 32 together with `meta` (the round-1 sweep reads one base per aligned group of 32)
    uint2* runs;               // feature / shared (optional): [tile][kMaxBuckets] = (offset of the tile's run inside the region, live entries) -
                               // lets collect.cuh walk the entries in stream order again
};

// MODE of the scatter kernel
//   kScatterCount  : entries for the count pass only (maskC windows)
//   kScatterFeat   : entries for the featurize pass only (feature windows of emitted clouds, 9-bit ROW delta);
//                    words it cannot encode (3 clouds in a word, delta > 510) are looked up directly
//   kScatterShared : ONE partition for both passes.  Every maskC window gets an entry; feature windows carry the
//                    9-bit CLOUD delta against the tile's first cloud (rows are not known yet when counting:
//                    the apply pass maps cloud -> row), count-only windows (lower-case bases, NOFEAT reads) the
//                    reserved delta 511.  Needs maskF subset of maskC (no quality filter) and clouds of >= 64
//                    bytes (<= 2 clouds per word, <= 257 per tile); the host checks both.
enum { kScatterCount = 0, kScatterFeat = 1, kScatterShared = 2 };
constexpr uint32_t kDeltaCountOnly = 510u; // never 511: with all index bits set that would read as kInvalidEntry

__device__ __forceinline__ uint32_t delta_bits(uint32_t delta) { return ((delta >> 3) << 26) | (delta & 7u); }
__device__ __forceinline__ uint32_t delta_of_entry(uint32_t e) { return ((e >> 26) << 3) | (e & 7u); }

template <int KT, int MODE>
__global__ void __launch_bounds__(ScatterCfg<MODE != kScatterCount>::kThreads, ScatterCfg<MODE != kScatterCount>::kMinCtas)
bucket_scatter_kernel(const ScatterParams Q, const FeatParams P)
{
    constexpr bool FEAT = MODE != kScatterCount; // entry format with delta bits, runs padded to 32 + base per group
    using Cfg = ScatterCfg<FEAT>;
    constexpr int CAP = Cfg::kStageCap, STRIDE = Cfg::kStageStride;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    ScatterSmem<FEAT>& S = *reinterpret_cast<ScatterSmem<FEAT>*>(smem_raw);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int k = KT ? KT : Q.k;
    uint32_t* const stage = S.stage;
    const int64_t n_tiles = (Q.w1 - Q.w0 + Cfg::kTileWords - 1) / Cfg::kTileWords;
    const unsigned long long policy = bulk_policy_evict_first(); // the entries are streamed: written once, read once much later
    if (threadIdx.x <= kMaxBuckets) { S.cnt[0][threadIdx.x] = 0u; S.cnt[1][threadIdx.x] = 0u; }
    int cur = 0;

    // Per tile: bin into the staging rows -> barrier -> the 64 lanes of warps 0 and 1 each claim the run of one slice and
    // hand the row to the copy engine (cp.async.bulk shared -> global) -> every warp moves on to the loads of its next
    // tile; the rows are reused once the engine has read them (wait_group.read + the barrier at the top).  Nobody copies
    // with LDS / STG, and there are two barriers per tile instead of four.
    for (int64_t t = blockIdx.x; t < n_tiles; t += gridDim.x, cur ^= 1) {
        uint32_t* const cnt = S.cnt[cur];
        if (warp < 2) bulk_wait_read(); // the engine is done reading the rows of the previous tile
        __syncthreads();

        const int64_t tile0 = Q.w0 + t * Cfg::kTileWords;
        const int64_t j = tile0 + threadIdx.x;
        {   // the tile this CTA walks next is far ahead in the stream: pull its lines into L2 now (no registers held),
            // so that the dependent loads at the top of the next trip are L2 hits instead of DRAM round trips
            const int64_t jn = j + (int64_t)gridDim.x * Cfg::kTileWords;
            if (jn < Q.w1) {
                if ((threadIdx.x & 15) == 0) asm volatile("prefetch.global.L2 [%0];" ::"l"(Q.codes + jn));
                if ((threadIdx.x & 31) == 0) {
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(Q.mask + jn));
                    if (FEAT) asm volatile("prefetch.global.L2 [%0];" ::"l"(P.wg + jn));
                    if (MODE == kScatterShared) asm volatile("prefetch.global.L2 [%0];" ::"l"(P.maskF + jn));
                }
            }
        }
        uint64_t lo = 0, hi = 0;
        // up to three trips over the 32 windows of the word, each with its own delta bits:
        //   (v0, d0) windows of the word's first cloud   (count mode: all windows, no delta)
        //   (v1, d1) windows of the cloud that starts inside the word (one word per cloud)
        //   (v2, d2) shared mode: count-only windows
        uint32_t v0 = 0u, d0 = 0u, v1 = 0u, d1 = 0u, v2 = 0u;
        uint32_t row0 = 0u, row1 = 0u;  // feat mode: rows of the first two trips (overflow redo)
        int32_t tile_base = 0;          // feat: rows emitted before the tile's first cloud; shared: the tile's first cloud
        if (MODE == kScatterFeat) tile_base = __ldg(P.row_lb + (__ldg(P.wg + tile0) & ~kWordMixed));
        if (MODE == kScatterShared) tile_base = (int32_t)(__ldg(P.wg + tile0) & ~kWordMixed);
        if (j < Q.w1) {
            const uint32_t mlo = __ldg(Q.mask + j);
            if (mlo != 0u) {
                const uint32_t mhi = __ldg(Q.mask + j + 1);
                v0 = window_valid_mask(mlo, mhi, k);
                if (MODE == kScatterFeat && v0) {
                    const uint32_t gw = __ldg(P.wg + j);
                    const uint32_t g = gw & ~kWordMixed;
                    int32_t r = __ldg(P.row_of_group + g), r2 = -1;
                    bool slow = false;
                    if (gw & kWordMixed) {
                        // the next cloud starts inside this word; a third one (clouds shorter than a word) -> slow path
                        const int64_t split = __ldg(P.gstart + g + 1) - j * 32;
                        slow = (int64_t)g + 2 < P.n_groups && __ldg(P.gstart + g + 2) < (j + 1) * 32;
                        r2 = __ldg(P.row_of_group + g + 1);
                        v1 = v0 & ~((1u << split) - 1u);
                        v0 &= (1u << split) - 1u;
                    }
                    if (r < 0) v0 = 0u;   // dropped cloud
                    if (r2 < 0) v1 = 0u;
                    if ((v0 && r - tile_base > kMaxRowDelta) || (v1 && r2 - tile_base > kMaxRowDelta)) slow = true;
                    if (slow) {
                        feat_word_direct(P, j, g, __ldg(Q.codes + j), __ldg(Q.codes + j + 1), v0 | v1);
                        v0 = v1 = 0u;
                    } else {
                        row0 = (uint32_t)r; row1 = (uint32_t)r2;
                        d0 = delta_bits((uint32_t)(r - tile_base));
                        d1 = delta_bits((uint32_t)(r2 - tile_base));
                    }
                }
                if (MODE == kScatterShared && v0) {
                    const uint32_t gw = __ldg(P.wg + j);
                    const uint32_t delta = (gw & ~kWordMixed) - (uint32_t)tile_base;
                    const uint32_t mf = __ldg(P.maskF + j), mf1 = __ldg(P.maskF + j + 1);
                    // same mask words as the count mask (no lower case, no NOFEAT read here - the usual case): same windows
                    const uint32_t vf = (mf == mlo && mf1 == mhi) ? v0 : (mf ? (window_valid_mask(mf, mf1, k) & v0) : 0u);
                    v2 = v0 & ~vf;
                    v0 = vf;
                    d0 = delta_bits(delta);
                    if (gw & kWordMixed) {
                        const int64_t split = __ldg(P.gstart + (gw & ~kWordMixed) + 1) - j * 32;
                        v1 = vf & ~((1u << split) - 1u);
                        v0 = vf & ((1u << split) - 1u);
                        d1 = delta_bits(delta + 1u);
                    }
                }
                if (v0 | v1 | v2) { lo = __ldg(Q.codes + j); hi = __ldg(Q.codes + j + 1); }
            }
        }
        // every lane issues the atomic: windows that emit nothing go to the dummy row, so the loop has
        // no divergence bookkeeping (a conditional shared atomic compiles to BSSY/BRA/ATOMS/BSYNC)
        // one pass over the 32 windows of the word: valid windows `v`, delta bits `d` (`dB` for the windows in `selB` when MERGED)
        auto pass = [&](auto merged, const uint32_t v, const uint32_t d, const uint32_t dB, const uint32_t selB) {
            constexpr bool MERGED = decltype(merged)::value;
            // four windows at a time: the four returning atomics are issued back to back, so their latency overlaps
            uint32_t yb[4];
            auto hand_out = [&](int i, uint32_t y) {
                yb[i & 3] = y;
                if ((i & 3) != 3) return;
                uint32_t bk[4], slot[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) bk[q] = (v & (1u << (i - 3 + q))) ? (yb[q] >> 26) : (uint32_t)kMaxBuckets;
#pragma unroll
                for (int q = 0; q < 4; ++q) slot[q] = atomicAdd(cnt + bk[q], 1u);
#pragma unroll
                for (int q = 0; q < 4; ++q) // a row that overflows is redone below
                {
                    PG_CHECK(bk[q] <= (uint32_t)kMaxBuckets && (bk[q] == (uint32_t)kMaxBuckets || (int)bk[q] < Q.geo.n_buckets));
                    const uint32_t dq = MERGED ? ((selB & (1u << (i - 3 + q))) ? dB : d) : d;
                    stage[bk[q] * STRIDE + min(slot[q], (uint32_t)(CAP - 1))] = FEAT ? ((yb[q] & kEntryIndexBits) | dq) : yb[q];
                }
            };
            // ptxas would hoist the 32 loop-invariant window extractions out of the trip loop and spill them: a shuffle
            // from the own lane is the identity, but not one the optimiser can see through (2 SHFL per 32 windows)
            lo = __shfl_sync(__activemask(), lo, lane);
            hi = __shfl_sync(__activemask(), hi, lane);
            for_each_window<KT>(lo, hi, k, hand_out);
        };
        // Words with a cloud boundary are one per cloud: rare with 20 KB clouds (their second cloud gets its own trip, which
        // almost no warp takes), but every tenth word with one cloud per read pair - then nearly every warp holds one and
        // the second trip would double the work of the whole kernel: such warps pick the delta per window in ONE pass.
        const bool merge = FEAT && __any_sync(0xffffffffu, v1 != 0u);
        if (FEAT && merge && (v0 | v1) != 0u) pass(std::true_type{}, v0 | v1, d0, d1, v1);
#pragma unroll 1
        for (int trip = (FEAT && merge) ? 2 : 0; trip < (MODE == kScatterCount ? 1 : MODE == kScatterFeat ? 2 : 3); ++trip) {
            const uint32_t v = trip == 0 ? v0 : trip == 1 ? v1 : v2;
            const uint32_t d = trip == 0 ? d0 : trip == 1 ? d1 : delta_bits(kDeltaCountOnly);
            if (v == 0u) continue; // trips 1 and 2 are rare: one word per cloud / lower-case bases
            pass(std::false_type{}, v, d, 0u, 0u);
        }
        bulk_store_fence(); // the rows were written through the generic proxy; the copy engine reads them through the async proxy
        __syncthreads();
        if (threadIdx.x <= kMaxBuckets) S.cnt[cur ^ 1][threadIdx.x] = 0u; // (last read before the barrier at the top of this trip)

        // ---- claim one run per slice in the entry buffer and hand it to the copy engine: lane <-> slice ----
        if (warp < 2) {
            const int b = 32 * warp + lane;
            uint2 run = make_uint2(0u, 0u);
            if (b < Q.geo.n_buckets) {
                const uint32_t c = cnt[b];
                const uint32_t n = c > (uint32_t)CAP ? 0u : c; // a row that overflowed is redone below
                const uint32_t n_pad = (n + Q.run_pad - 1u) & ~(Q.run_pad - 1u);
                if (n) {
                    uint32_t* row = stage + b * STRIDE;
                    const unsigned long long off = atomicAdd(&Q.st->cursors[b], (unsigned long long)n_pad);
                    if (off + n_pad <= Q.geo.cap) {
                        const unsigned long long gb = (unsigned long long)b * Q.geo.cap + off;
                        PG_CHECK(n <= (uint32_t)CAP && (gb & 3ull) == 0ull && gb + n_pad <= (unsigned long long)(b + 1) * Q.geo.cap);
                        for (uint32_t e = n; e < n_pad; ++e) row[e] = kInvalidEntry; // (CAP is a multiple of the padding: it fits)
                        bulk_store_fence();
                        bulk_store(Q.entries + gb, row, n_pad * 4u, policy);
                        if (FEAT && Q.meta) for (uint32_t i = 0; i < (n_pad >> 5); ++i) Q.meta[(gb >> 5) + i] = tile_base;
                        run = make_uint2((uint32_t)off, n);
                    } else { // region full (pathological repeats): this lane applies / looks up the run itself
                        atomicMin(&Q.st->limits[b], off);
                        for (uint32_t e = 0; e < n; ++e) {
                            const uint32_t v = row[e];
                            const uint32_t idx = ((uint32_t)b << kSliceBits) | ((v >> 3) & Q.geo.low_mask);
                            if (MODE == kScatterFeat) {
                                unsigned long long key;
                                if (abd_key(P, __ldg(P.table.counts + idx), (uint32_t)tile_base + delta_of_entry(v), key))
                                    atomicAdd(P.abd + (int64_t)(key >> 32) * P.vs + (uint32_t)key, 1u);
                            } else {
                                table_add_checked(Q.table + idx, 1u, Q.sat);
                            }
                        }
                        if (MODE == kScatterShared) *Q.lost = 1u;
                    }
                }
            }
            if (FEAT && Q.runs) Q.runs[t * kMaxBuckets + b] = run; // (runs that took an overflow path stay empty: their tallies were made directly)
            bulk_commit();
        }

        // ---- staging rows that overflowed (tandem repeats, poly-G): walk the tile again, those slices go straight to the table ----
        const bool any_ovf = __any_sync(0xffffffffu, cnt[lane] > (uint32_t)CAP || cnt[lane + 32] > (uint32_t)CAP);
        if (any_ovf && (v0 | v1 | v2) != 0u) {
            for (int i = 0; i < 32; ++i) {
                if (!(((v0 | v1 | v2) >> i) & 1u)) continue;
                const uint32_t y = window_y(lo, hi, i, k);
                if (cnt[y >> 26] <= (uint32_t)CAP) continue;
                if (MODE == kScatterFeat) abd_direct(P, y, ((v0 >> i) & 1u) ? row0 : row1);
                else table_add_checked(Q.table + (y >> 3), 1u, Q.sat);
            }
            if (MODE == kScatterShared) *Q.lost = 1u;
        }
    }
    if (warp < 2) bulk_wait_all(); // the rows must outlive the copies
}

// entries per region after a scatter launch (the cursors are reset for the next segment)
__global__ void bucket_save_fill_kernel(const BucketState* st, BucketGeom geo, unsigned long long* fill_out)
{
    if (threadIdx.x < kMaxBuckets)
        fill_out[threadIdx.x] = (int)threadIdx.x < geo.n_buckets ? min(min(st->cursors[threadIdx.x], st->limits[threadIdx.x]), geo.cap) : 0ull;
}

__global__ void bucket_reset_ticket_kernel(BucketState* st) { st->ticket = 0ull; }

// ---------------------------------------------------------------------------
// apply: ordered tickets over the filled part of every region: ticket -> (slice, offset)
// ---------------------------------------------------------------------------
struct ApplySmem {
    unsigned long long fill[kMaxBuckets];           // valid entries per region
    unsigned long long chunk_base[kMaxBuckets + 1]; // prefix sum of chunks per region
    unsigned long long ticket;
};

// saved_fill != nullptr: regions of an earlier scatter launch whose fills were kept (bucket_save_fill_kernel)
template <int CHUNK = kChunk>
__device__ __forceinline__ void apply_prologue(ApplySmem& A, const BucketGeom& geo, const BucketState* st, const unsigned long long* saved_fill = nullptr)
{
    if (threadIdx.x == 0) {
        unsigned long long acc = 0;
        for (int b = 0; b < geo.n_buckets; ++b) {
            const unsigned long long f = saved_fill ? saved_fill[b] : min(min(st->cursors[b], st->limits[b]), geo.cap);
            A.fill[b] = f;
            A.chunk_base[b] = acc;
            acc += (f + CHUNK - 1) / CHUNK;
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
    const int lane = threadIdx.x & 31;
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
            e[u] = i < n ? __ldcs(src + i) : kInvalidEntry;
        }
#pragma unroll
        for (int u = 0; u < kChunk / 256; ++u) {
            const bool live = e[u] != kInvalidEntry;
            // identical neighbours (homopolymer runs: a poly-G tail is ~86 copies of one window, staged next
            // to each other) would serialise on one L2 address: fold them inside the warp first
            const uint32_t next = __shfl_down_sync(0xffffffffu, e[u], 1);
            if (__any_sync(0xffffffffu, live && lane < 31 && next == e[u])) {
                const uint32_t live_mask = __ballot_sync(0xffffffffu, live);
                if (live) {
                    const uint32_t peers = __match_any_sync(live_mask, e[u]);
                    if (lane == __ffs(peers) - 1) atomicAdd(slice + ((e[u] >> 3) & geo.low_mask), (uint32_t)__popc(peers));
                }
            } else if (live) {
                atomicAdd(slice + ((e[u] >> 3) & geo.low_mask), 1u);
            }
        }
    }
}

// sweep the partitioned entries slice by slice: gather (L2 hit), bin, reduce.
// The ticket of the NEXT chunk is requested before the current chunk is processed and read when it
// is done, so the round trip of the global atomic is hidden behind the chunk's own work.
// SHARED: the entries come from the shared partition of the count pass (cloud deltas; cloud -> row here)
template <bool SHARED>
__global__ void __launch_bounds__(256)
bucket_apply_feat_kernel(const uint32_t* __restrict__ entries, const int32_t* __restrict__ meta, BucketGeom geo,
                         BucketState* __restrict__ st, const unsigned long long* __restrict__ saved_fill, const FeatParams P)
{
    // Tickets are taken per WARP (kFeatChunk entries each): no block-wide barrier in the loop, so a warp that drew a slow
    // chunk (more distinct tallies) does not hold seven others back.  The order of the sweep is the order of the tickets.
    __shared__ ApplySmem A;
    apply_prologue<kFeatChunk>(A, geo, st, SHARED ? saved_fill : nullptr);
    const unsigned long long n_chunks = A.chunk_base[geo.n_buckets];
    const int lane = threadIdx.x & 31;
    unsigned long long c = 0ull;
    if (lane == 0) c = atomicAdd(&st->ticket, 1ull);
    c = __shfl_sync(0xffffffffu, c, 0);
    int b = 0;
    while (c < n_chunks) {
        unsigned long long next = 0ull;
        if (lane == 0) next = atomicAdd(&st->ticket, 1ull); // in flight while the chunk is processed
        while (A.chunk_base[b + 1] <= c) ++b;
        const unsigned long long off = (unsigned long long)b * geo.cap + (c - A.chunk_base[b]) * kFeatChunk; // multiple of 32
        const uint32_t n = (uint32_t)min((unsigned long long)kFeatChunk, A.fill[b] - (c - A.chunk_base[b]) * kFeatChunk);
        const uint32_t* slice = P.table.counts + ((size_t)b << kSliceBits);
        for (uint32_t sb = 0; sb < n; sb += 128u) { // 128 entries at a time: 4 per lane in flight
            const uint32_t* src = entries + off + sb;
            const int32_t* row0_of = meta + ((off + sb) >> 5);
            uint32_t e[4], cnt[4];
            int32_t row0[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const uint32_t i = sb + lane + 32u * u;
                e[u] = i < n ? __ldcs(src + lane + 32u * u) : kInvalidEntry;
                row0[u] = i < n ? __ldg(row0_of + u) : 0; // one base row per warp-wide group of 32 entries
            }
#pragma unroll
            for (int u = 0; u < 4; ++u)
                cnt[u] = (e[u] != kInvalidEntry && !(SHARED && delta_of_entry(e[u]) == kDeltaCountOnly)) ? __ldg(slice + ((e[u] >> 3) & geo.low_mask)) : 0u;
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                // the 32 entries of a warp share row0: (delta, bin) identifies the tally inside the warp in 22 bits
                const uint32_t delta = delta_of_entry(e[u]);
                uint32_t c32 = cnt[u] & kCountMask;
                bool live = cnt[u] != 0u && c32 < P.clamp; // absent k-mers are skipped (count_kmer.cpp:87); cnt = 0 for padding
                int32_t row = row0[u] + (int32_t)delta;
                if (SHARED) { // row0 is the tile's first cloud: cloud -> row (lanes of a warp hit 1-3 addresses); dropped clouds fall out here
                    live = live && delta != kDeltaCountOnly;
                    PG_CHECK(!live || (row >= 0 && row < P.n_groups));
                    row = live ? __ldg(P.row_of_group + row) : -1;
                    live = row >= 0;
                }
                const uint32_t key = (delta << 13) | (live ? abd_bin(P, c32) : 0u); // vector_size <= 8192
                const uint32_t live_mask = __ballot_sync(0xffffffffu, live);
                if (live) {
                    const uint32_t peers = __match_any_sync(live_mask, key);
                    if (lane == __ffs(peers) - 1)
                        atomicAdd(P.abd + (int64_t)row * P.vs + (key & 0x1FFFu), (uint32_t)__popc(peers));
                }
            }
        }
        c = __shfl_sync(0xffffffffu, next, 0);
    }
}

} // namespace pg
