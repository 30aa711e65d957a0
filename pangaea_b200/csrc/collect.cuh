// collect.cuh - the featurize pass of the sliced path in two kernels: LOOK UP, then COLLECT.
//
// Round 1 swept the partitioned entries once and did everything per entry: gather the count (L2 hit), bin it, find equal
// (cloud, bin) pairs in the warp (MATCH.ANY) and reduce them into the abundance matrix with global REDs.  Two things were
// wrong with that: the REDs and the MATCH sit on the critical path of a kernel that is bound by the gather rate anyway, and
// how many REDs a warp needs depends on how many distinct bins its 32 entries hold - so the kernel slowed down with table
// depth (37.5 ms -> 55 ms at 8x depth, the 8-GPU scaling loss) and collapsed when clouds are tiny (one cloud per pair:
// every row was read-modify-written from DRAM once per table slice, 64 times).
// Now:
//   lookup  : the same ordered sweep, but all it does is gather, bin and WRITE THE BIN BACK in place of the index bits
//             (coalesced store).  Nothing depends on what the bins are.
//   collect : walks the read stream tile by tile - the order in which the scatter kernel produced the entries.  The
//             scatter kernel leaves a run table (where each tile's run sits in each region), so a tile's 64 runs are 64
//             coalesced reads; all of them belong to the one or two (or, with tiny clouds, few dozen) clouds of the tile,
//             whose histograms live in shared memory.  A warp takes a run, a lane eight consecutive entries; the three
//             most frequent (cloud, bin) keys of the run are tallied in registers, the rest with shared atomics; rows
//             leave shared memory once per tile.  Depth-independent, and a row is written once.
// Entry after lookup: delta bits unchanged (bucket.cuh), bits 3..16 = bin + 1 (0: nothing to tally - absent k-mer, count
// beyond the histogram, count-only window, padding).
#pragma once
#include "bucket.cuh"

namespace pg {

constexpr uint32_t kBinField = 0x3FFFu; // 14 bits: vector_size <= 8192

struct RunRef { uint32_t off, n; }; // a tile's run inside one region: offset in entries (multiple of the run padding), live entries

template <bool SHARED>
__global__ void __launch_bounds__(256)
bucket_lookup_kernel(uint32_t* __restrict__ entries, BucketGeom geo, BucketState* __restrict__ st, const unsigned long long* __restrict__ saved_fill,
                     const FeatParams P)
{
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
        const unsigned long long off = (unsigned long long)b * geo.cap + (c - A.chunk_base[b]) * kFeatChunk;
        const uint32_t n = (uint32_t)min((unsigned long long)kFeatChunk, A.fill[b] - (c - A.chunk_base[b]) * kFeatChunk);
        const uint32_t* slice = P.table.counts + ((size_t)b << kSliceBits);
        for (uint32_t sb = 0; sb < n; sb += 256u) { // 256 entries at a time: 8 gathers per lane in flight
            uint32_t* src = entries + off + sb;
            uint32_t e[8], cnt[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const uint32_t i = sb + lane + 32u * u;
                e[u] = i < n ? __ldcs(src + lane + 32u * u) : kInvalidEntry;
            }
#pragma unroll
            for (int u = 0; u < 8; ++u)
                cnt[u] = (e[u] != kInvalidEntry && !(SHARED && delta_of_entry(e[u]) == kDeltaCountOnly)) ? __ldg(slice + ((e[u] >> 3) & geo.low_mask)) : 0u;
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const uint32_t i = sb + lane + 32u * u;
                const uint32_t c32 = cnt[u] & kCountMask;
                const bool live = cnt[u] != 0u && c32 < P.clamp; // absent k-mers are skipped (count_kmer.cpp:87); cnt = 0 for padding
                const uint32_t field = live ? abd_bin(P, c32) + 1u : 0u;
                if (i < n) __stcs(src + lane + 32u * u, (e[u] & ~kEntryIndexBits) | (field << 3));
            }
        }
        c = __shfl_sync(0xffffffffu, next, 0);
    }
}

// ---- collect ----
constexpr int kCollectThreads = 512;

struct CollectParams {
    const uint32_t* entries;     // after bucket_lookup_kernel
    const RunRef* runs;          // [n_tiles][kMaxBuckets]
    BucketGeom geo;
    int64_t w0, w1;              // the segment, in words
    int tile_words;              // words per scatter tile (ScatterCfg<true>::kTileWords)
    int slots;                   // cloud slots with a histogram in shared memory
    int shared;                  // 1: deltas are CLOUD deltas against the tile's first cloud; 0: ROW deltas against row_lb of that cloud
};

// HOT: the three most frequent (cloud, bin) keys of a run are tallied in registers - pays for clouds of many reads (a run of
// 230 entries holds a few clouds whose k-mers share a few bins); with one cloud per read pair a run spans ~50 clouds and no
// key repeats: the vote and the three compares per entry are dead weight (ncu: 79 thread instructions per window).
template <bool HOT>
__global__ void __launch_bounds__(kCollectThreads)
bucket_collect_kernel(const CollectParams C, const FeatParams P)
{
    extern __shared__ uint32_t csm[];
    uint32_t* hist = csm;                                   // [slots][vs]
    int32_t* slot_row = reinterpret_cast<int32_t*>(csm + (size_t)C.slots * P.vs); // [slots]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t n_tiles = (C.w1 - C.w0 + C.tile_words - 1) / C.tile_words;
    // contiguous tiles per CTA: consecutive tiles share clouds, so a cloud's row is touched by one or two CTAs
    const int64_t per = (n_tiles + gridDim.x - 1) / gridDim.x;
    const int64_t t_begin = (int64_t)blockIdx.x * per, t_end = min(n_tiles, t_begin + per);
    const int n_hist = C.slots * P.vs;
    for (int i = threadIdx.x; i < n_hist; i += kCollectThreads) hist[i] = 0u;
    __syncthreads();
    for (int64_t t = t_begin; t < t_end; ++t) {
        const int64_t tile0 = C.w0 + t * C.tile_words;
        const uint32_t g0 = __ldg(P.wg + tile0) & ~kWordMixed;       // the tile's first cloud
        const int32_t base = C.shared ? (int32_t)g0 : __ldg(P.row_lb + g0);
        if (threadIdx.x < C.slots) {
            int32_t row = -1;
            const int64_t id = (int64_t)base + threadIdx.x;
            if (C.shared) { if (id < P.n_groups) row = __ldg(P.row_of_group + id); }
            else row = (int32_t)id; // (only deltas that exist are ever used)
            slot_row[threadIdx.x] = row;
        }
        __syncthreads();
        const RunRef* rr = C.runs + t * kMaxBuckets;
        for (int b = warp; b < C.geo.n_buckets; b += kCollectThreads / 32) {
            const RunRef r = rr[b];
            if (!r.n) continue;
            const uint32_t* src = C.entries + (unsigned long long)b * C.geo.cap + r.off;
            // the three most frequent keys among the run's first 32 entries are tallied in registers
            uint32_t hot[3] = { 0u, 0u, 0u };
            if (HOT) {
            uint32_t s = lane < r.n ? __ldg(src + lane) : 0u;
            uint32_t key_s = ((s >> 3) & kBinField) ? ((delta_of_entry(s) << 14) | ((s >> 3) & kBinField)) : 0u;
#pragma unroll
            for (int h = 0; h < 3; ++h) {
                const uint32_t peers = __match_any_sync(0xffffffffu, key_s);
                const uint32_t votes = key_s ? ((uint32_t)__popc(peers) << 5 | (31u - lane)) : 0u; // most votes, lowest lane wins
                const uint32_t best = __reduce_max_sync(0xffffffffu, votes);
                const int who = 31 - (int)(best & 31u);
                hot[h] = best >> 5 ? __shfl_sync(0xffffffffu, key_s, who) : 0u;
                if (key_s == hot[h]) key_s = 0u; // out of the next vote
            }
            }
            uint32_t c0 = 0u, c1 = 0u, c2 = 0u;
            for (uint32_t i0 = 8u * lane; i0 < r.n; i0 += 256u) { // a lane takes eight consecutive entries (runs start 16 B aligned and are padded to 32)
                const uint4 a = __ldcs(reinterpret_cast<const uint4*>(src + i0));
                const uint4 d = __ldcs(reinterpret_cast<const uint4*>(src + i0 + 4));
                const uint32_t v[8] = { a.x, a.y, a.z, a.w, d.x, d.y, d.z, d.w };
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const uint32_t f = (v[u] >> 3) & kBinField;
                    if (i0 + u >= r.n || !f) continue;   // (padding carries bin field 0 after the lookup)
                    const uint32_t delta = delta_of_entry(v[u]);
                    const uint32_t key = (delta << 14) | f;
                    if (HOT && key == hot[0]) ++c0;
                    else if (HOT && key == hot[1]) ++c1;
                    else if (HOT && key == hot[2]) ++c2;
                    else if (delta < (uint32_t)C.slots) atomicAdd(&hist[delta * P.vs + (f - 1u)], 1u);
                    else { // more clouds in the tile than slots: straight to the matrix
                        int32_t row = (int32_t)delta + base;
                        if (C.shared) row = row < P.n_groups ? __ldg(P.row_of_group + row) : -1;
                        if (row >= 0) atomicAdd(P.abd + (int64_t)row * P.vs + (f - 1u), 1u);
                    }
                }
            }
            if constexpr (HOT) {
#pragma unroll
            for (int d = 16; d; d >>= 1) {
                c0 += __shfl_xor_sync(0xffffffffu, c0, d);
                c1 += __shfl_xor_sync(0xffffffffu, c1, d);
                c2 += __shfl_xor_sync(0xffffffffu, c2, d);
            }
            if (lane < 3) {
                const uint32_t key = hot[lane], cnt = lane == 0 ? c0 : lane == 1 ? c1 : c2;
                if (key && cnt) {
                    const uint32_t delta = key >> 14, f = key & kBinField;
                    if (delta < (uint32_t)C.slots) atomicAdd(&hist[delta * P.vs + (f - 1u)], cnt);
                    else {
                        int32_t row = (int32_t)delta + base;
                        if (C.shared) row = row < P.n_groups ? __ldg(P.row_of_group + row) : -1;
                        if (row >= 0) atomicAdd(P.abd + (int64_t)row * P.vs + (f - 1u), cnt);
                    }
                }
            }
            }
        }
        __syncthreads();
        // rows leave shared memory: non-zero bins are reduced into the (zeroed) matrix - a cloud that spans tiles or CTAs adds up
        // there.  One warp per slot: the row number is read once and no index is divided by the row length.
        for (int sl = warp; sl < C.slots; sl += kCollectThreads / 32) {
            const int32_t row = slot_row[sl];
            uint32_t* h = hist + sl * P.vs;
            uint32_t* dst = P.abd + (int64_t)max(row, 0) * P.vs;
            for (int i = lane; i < P.vs; i += 32) {
                const uint32_t v = h[i];
                if (!v) continue;
                h[i] = 0u;
                if (row >= 0) atomicAdd(dst + i, v);
            }
        }
        __syncthreads();
    }
}

} // namespace pg
