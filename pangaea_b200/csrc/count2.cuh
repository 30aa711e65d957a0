// count2.cuh - second partition level + shared-memory apply of the count pass.
//
// Why: applying the partitioned entries with one RED per window into an L2-resident slice
// (bucket_apply_count_kernel) is capped by the SM -> L2 atomic issue rate, ~1 lane / clk / SM:
// 190 G RED/s on B200 whatever the table size (profiles/microbench_r01.txt) = 45 ms for the
// 8.6 G windows of the headline workload.  Shared-memory atomics on a 2^15-counter table run
// at 2 370 G/s (profiles/microbench2_r01.txt).  So the entries of every 32 MiB slice are
// partitioned once more, by the top 8 bits of their index inside the slice, into 256
// sub-slices of 2^15 counters = 128 KB - one shared-memory table:
//   split : 64 regions of u32 entries -> 64 x 256 sub-regions of u16 entries (index inside the
//           sub-slice).  Same tile machinery as the scatter kernels (bucket.cuh): one returning
//           shared atomic per entry hands out the slot in a fixed-capacity staging row.
//   apply : one CTA per sub-region chunk: zero a 128 KB table in shared memory, stream the
//           chunk's entries into it with shared atomics, then add the table to the global
//           counters (plain coalesced read-modify-write when the sub-slice has one chunk, RED
//           when a heavy sub-slice - a k-mer repeated millions of times - was cut into several).
// Both overflow cases of bucket.cuh exist here too and are handled the same way (straight to
// the global table): a staging row (poly-A runs pile into one sub-slice) and a sub-region.
// HBM traffic: split reads 4 B and writes 2 B per window, apply reads 2 B per window plus the
// table once per segment (2 x 2 GiB) - against 4 B + one L2 RED per window before.
#pragma once
#include "bucket.cuh"

namespace pg {

constexpr int kSubBits = 15;                                 // counters per sub-slice
constexpr int kSubFan = 1 << (kSliceBits - kSubBits);        // 256 sub-slices per slice
constexpr int kSubWords = 1 << kSubBits;
constexpr int kSplitThreads = 512;
constexpr int kSplitPer = 32;                                // entries per thread
constexpr int kSplitTile = kSplitThreads * kSplitPer;        // 16384 entries: 64 per sub-slice
constexpr int kSplitCap = 128;                               // staging slots per sub-slice per tile (mean 64 + 8 sigma)
constexpr int kSplitStride = 136;                            // u16 per staging row: 68 words, so that rows start on different banks (all
                                                             // rows fill at the same pace: with a stride of 64 words the 32 lanes of a
                                                             // store would pile onto the few banks of the current fill level)
constexpr int kSubApplyThreads = 1024;
constexpr uint32_t kSubOverflow = 0xFFFFFFFFu;
constexpr uint16_t kSubInvalid = 0x8000u;                    // padding entry: runs start 16 B aligned (8 entries); lands in a dummy counter

struct SubGeom {
    int n_sub;              // n_buckets * kSubFan
    uint32_t cap;           // u16 entries per sub-region (multiple of 8)
    uint32_t chunk;         // entries per apply item (multiple of 8)
};

struct SubState {           // device arrays, n_sub entries each (+1 for item_base)
    uint32_t* cursors;      // entries claimed per sub-region (may exceed cap)
    uint32_t* limits;       // first offset that did not fit (cap if none)
    uint32_t* item_base;    // exclusive scan of the chunks per sub-region, n_sub + 1
};

__global__ void sub_reset_kernel(SubState s, SubGeom geo)
{
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < geo.n_sub; i += gridDim.x * blockDim.x) { s.cursors[i] = 0u; s.limits[i] = geo.cap; }
}

struct SplitSmem {
    uint32_t cnt[2][kSubFan + 8];         // +1: dummy row of the padding lanes; double-buffered like ScatterSmem::cnt
    unsigned long long fill[kMaxBuckets];
    unsigned long long tile_base[kMaxBuckets + 1];
    alignas(16) uint16_t stage[(kSubFan + 1) * kSplitStride];
};

__global__ void __launch_bounds__(kSplitThreads, 2)
bucket_split_kernel(const uint32_t* __restrict__ entries, BucketGeom geo, const BucketState* __restrict__ st, SubGeom sg, SubState ss,
                    uint16_t* __restrict__ entries2, uint32_t* __restrict__ table, uint32_t* __restrict__ sat)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    SplitSmem& S = *reinterpret_cast<SplitSmem*>(smem_raw);
    const int lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        unsigned long long acc = 0;
        for (int b = 0; b < geo.n_buckets; ++b) {
            const unsigned long long f = min(min(st->cursors[b], st->limits[b]), geo.cap);
            S.fill[b] = f;
            S.tile_base[b] = acc;
            acc += (f + kSplitTile - 1) / kSplitTile;
        }
        S.tile_base[geo.n_buckets] = acc;
    }
    for (int i = threadIdx.x; i <= kSubFan; i += kSplitThreads) { S.cnt[0][i] = 0u; S.cnt[1][i] = 0u; }
    __syncthreads();
    const unsigned long long n_tiles = S.tile_base[geo.n_buckets];
    const unsigned long long policy = bulk_policy_evict_first();
    uint16_t* const stage = S.stage;
    int b = 0, cur = 0;
    // Same shape as the scatter kernel (bucket.cuh): bin -> barrier -> thread s claims the run of sub-slice s and hands its
    // row to the copy engine (cp.async.bulk) -> on to the next tile; rows are reused after wait_group.read + the top barrier.
    for (unsigned long long t = blockIdx.x; t < n_tiles; t += gridDim.x, cur ^= 1) {
        while (S.tile_base[b + 1] <= t) ++b; // tiles only grow
        uint32_t* const cnt = S.cnt[cur];
        const unsigned long long off = (t - S.tile_base[b]) * kSplitTile;
        const uint32_t n = (uint32_t)min((unsigned long long)kSplitTile, S.fill[b] - off);
        const uint4* src = reinterpret_cast<const uint4*>(entries + (unsigned long long)b * geo.cap + off);
        uint4 e[kSplitPer / 4];
#pragma unroll
        for (int u = 0; u < kSplitPer / 4; ++u) {
            const uint32_t i4 = u * kSplitThreads + threadIdx.x; // coalesced 16 B per lane
            e[u] = (4u * i4 < n) ? __ldcs(src + i4) : make_uint4(0, 0, 0, 0);
        }
        if (threadIdx.x < kSubFan) bulk_wait_read(); // the engine is done reading the rows of the previous tile
        __syncthreads();
#pragma unroll
        for (int u = 0; u < kSplitPer / 4; ++u) {
            const uint32_t i0 = 4u * (u * kSplitThreads + threadIdx.x);
            const uint32_t v[4] = { e[u].x, e[u].y, e[u].z, e[u].w };
            uint32_t sub[4], slot[4]; // the four returning atomics back to back: their latency overlaps
#pragma unroll
            for (int c = 0; c < 4; ++c) sub[c] = (i0 + c < n && v[c] != kInvalidEntry) ? ((v[c] >> (3 + kSubBits)) & (kSubFan - 1)) : (uint32_t)kSubFan;
#pragma unroll
            for (int c = 0; c < 4; ++c) slot[c] = atomicAdd(cnt + sub[c], 1u);
#pragma unroll
            for (int c = 0; c < 4; ++c) stage[sub[c] * kSplitStride + min(slot[c], (uint32_t)(kSplitCap - 1))] = (uint16_t)((v[c] >> 3) & (kSubWords - 1));
        }
        bulk_store_fence();
        __syncthreads();
        for (int i = threadIdx.x; i <= kSubFan; i += kSplitThreads) S.cnt[cur ^ 1][i] = 0u;
        if (threadIdx.x < kSubFan) { // claim one run per sub-region and hand it to the copy engine
            const int s = threadIdx.x;
            const uint32_t c = cnt[s];
            const uint32_t m = c > (uint32_t)kSplitCap ? 0u : c; // a row that overflowed is redone below
            if (m) {
                const int idx = b * kSubFan + s;
                const uint32_t claim = (m + 7u) & ~7u;
                uint16_t* row = stage + s * kSplitStride;
                const uint32_t at = atomicAdd(ss.cursors + idx, claim); // (wraps only beyond 2^32 entries in one sub-slice of one segment)
                if ((unsigned long long)at + claim <= sg.cap) {
                    PG_CHECK(m <= (uint32_t)kSplitCap && (at & 7u) == 0u);
                    for (uint32_t i = m; i < claim; ++i) row[i] = kSubInvalid; // pad the run to whole 16 B quads
                    bulk_store_fence();
                    bulk_store(entries2 + (unsigned long long)idx * sg.cap + at, row, claim * 2u, policy);
                } else { // sub-region full: apply the run here
                    atomicMin(ss.limits + idx, min(at, sg.cap));
                    uint32_t* sub_table = table + ((size_t)b << kSliceBits) + ((size_t)s << kSubBits);
                    for (uint32_t i = 0; i < m; ++i) table_add_checked(sub_table + row[i], 1u, sat);
                }
            }
            bulk_commit();
        }
        // staging rows that overflowed: those sub-slices' entries of this tile go straight to the table (rare: re-read the tile)
        bool any_ovf = false;
#pragma unroll
        for (int j = 0; j < kSubFan / 32; ++j) any_ovf |= cnt[lane + 32 * j] > (uint32_t)kSplitCap;
        if (__any_sync(0xffffffffu, any_ovf)) {
            for (uint32_t i = threadIdx.x; i < n; i += kSplitThreads) {
                const uint32_t v = __ldg(entries + (unsigned long long)b * geo.cap + off + i);
                if (v == kInvalidEntry) continue;
                const uint32_t sub = (v >> (3 + kSubBits)) & (kSubFan - 1);
                if (cnt[sub] > (uint32_t)kSplitCap) table_add_checked(table + ((size_t)b << kSliceBits) + ((v >> 3) & geo.low_mask), 1u, sat);
            }
        }
    }
    if (threadIdx.x < kSubFan) bulk_wait_all(); // the rows must outlive the copies
}

// chunks per sub-region -> exclusive scan (one CTA; n_sub <= 16384)
__global__ void __launch_bounds__(1024)
sub_items_kernel(SubState ss, SubGeom sg)
{
    __shared__ uint32_t warp_sums[32];
    __shared__ uint32_t carry_s;
    if (threadIdx.x == 0) carry_s = 0u;
    __syncthreads();
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (int base = 0; base < sg.n_sub; base += 1024) {
        const int i = base + threadIdx.x;
        uint32_t v = 0u;
        if (i < sg.n_sub) {
            const uint32_t f = min(min(ss.cursors[i], ss.limits[i]), sg.cap);
            v = (f + sg.chunk - 1) / sg.chunk;
        }
        uint32_t inc = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, inc, d);
            if (lane >= d) inc += t;
        }
        if (lane == 31) warp_sums[wid] = inc;
        __syncthreads();
        if (wid == 0) {
            uint32_t s = warp_sums[lane];
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t t = __shfl_up_sync(0xffffffffu, s, d);
                if (lane >= d) s += t;
            }
            warp_sums[lane] = s;
        }
        __syncthreads();
        const uint32_t carry = carry_s;
        const uint32_t ex = carry + (wid ? warp_sums[wid - 1] : 0u) + inc - v;
        if (i < sg.n_sub) ss.item_base[i] = ex;
        __syncthreads();
        if (threadIdx.x == 1023) carry_s = ex + v;
        __syncthreads();
    }
    if (threadIdx.x == 0) ss.item_base[sg.n_sub] = carry_s;
}

__global__ void __launch_bounds__(kSubApplyThreads, 1)
sub_apply_kernel(const uint16_t* __restrict__ entries2, SubGeom sg, SubState ss, uint32_t* __restrict__ table, uint32_t* __restrict__ sat)
{
    extern __shared__ __align__(16) uint32_t tab[]; // kSubWords counters + 1 dummy (padding entries)
    for (int w = threadIdx.x * 4; w < kSubWords; w += kSubApplyThreads * 4) *reinterpret_cast<uint4*>(tab + w) = make_uint4(0, 0, 0, 0);
    __syncthreads();
    const uint32_t n_items = ss.item_base[sg.n_sub];
    constexpr int kDepth = 4;                                  // 128-bit loads in flight per thread
    constexpr int kMergePer = kSubWords / (4 * kSubApplyThreads); // uint4 per thread in the merge (8)
    for (uint32_t item = blockIdx.x; item < n_items; item += gridDim.x) {
        int lo = 0, hi = sg.n_sub; // largest i with item_base[i] <= item
        while (hi - lo > 1) {
            const int mid = (lo + hi) >> 1;
            if (__ldg(ss.item_base + mid) <= item) lo = mid; else hi = mid;
        }
        const int i = lo;
        const uint32_t first = __ldg(ss.item_base + i), n_chunks = __ldg(ss.item_base + i + 1) - first;
        const uint32_t fill = min(min(ss.cursors[i], ss.limits[i]), sg.cap); // multiple of 8
        const uint32_t start = (item - first) * sg.chunk;
        PG_CHECK(start < fill && (fill & 7u) == 0u && fill <= sg.cap);
        const uint32_t n8 = min(sg.chunk, fill - start) >> 3;
        const uint4* src = reinterpret_cast<const uint4*>(entries2 + (unsigned long long)i * sg.cap + start); // 16 B aligned: cap, chunk multiples of 8
        for (uint32_t base = 0; base < n8; base += kDepth * kSubApplyThreads) {
            uint4 v[kDepth];
#pragma unroll
            for (int d = 0; d < kDepth; ++d) {
                const uint32_t i8 = base + d * kSubApplyThreads + threadIdx.x;
                v[d] = i8 < n8 ? __ldcs(src + i8) : make_uint4(0x80008000u, 0x80008000u, 0x80008000u, 0x80008000u);
            }
#pragma unroll
            for (int d = 0; d < kDepth; ++d) {
                const uint32_t w[4] = { v[d].x, v[d].y, v[d].z, v[d].w };
#pragma unroll
                for (int c = 0; c < 4; ++c) { // padding entries (0x8000) land in the dummy counter
                    atomicAdd(tab + min(w[c] & 0xFFFFu, (uint32_t)kSubWords), 1u);
                    atomicAdd(tab + min(w[c] >> 16, (uint32_t)kSubWords), 1u);
                }
            }
        }
        __syncthreads();
        uint4* dst = reinterpret_cast<uint4*>(table + ((size_t)i << kSubBits));
        uint4 t[kMergePer], g[kMergePer];
#pragma unroll
        for (int u = 0; u < kMergePer; ++u) t[u] = reinterpret_cast<const uint4*>(tab)[u * kSubApplyThreads + threadIdx.x];
        if (n_chunks == 1) { // sole owner of the sub-slice in this launch: plain read-modify-write, all loads in flight together
#pragma unroll
            for (int u = 0; u < kMergePer; ++u) g[u] = dst[u * kSubApplyThreads + threadIdx.x];
#pragma unroll
            for (int u = 0; u < kMergePer; ++u) {
                if (t[u].x | t[u].y | t[u].z | t[u].w) {
                    // both terms are < 2^31 (a segment holds < 2^31 windows, the table is <= kCountMax at segment
                    // boundaries), so the sum cannot wrap: clamp it (table.cuh: saturation)
                    g[u].x = min(g[u].x + t[u].x, kCountMax); g[u].y = min(g[u].y + t[u].y, kCountMax);
                    g[u].z = min(g[u].z + t[u].z, kCountMax); g[u].w = min(g[u].w + t[u].w, kCountMax);
                    dst[u * kSubApplyThreads + threadIdx.x] = g[u];
                    reinterpret_cast<uint4*>(tab)[u * kSubApplyThreads + threadIdx.x] = make_uint4(0, 0, 0, 0);
                }
            }
        } else {
#pragma unroll
            for (int u = 0; u < kMergePer; ++u) {
                uint32_t* d4 = reinterpret_cast<uint32_t*>(dst + u * kSubApplyThreads + threadIdx.x);
                if (t[u].x) table_add_checked(d4, t[u].x, sat);
                if (t[u].y) table_add_checked(d4 + 1, t[u].y, sat);
                if (t[u].z) table_add_checked(d4 + 2, t[u].z, sat);
                if (t[u].w) table_add_checked(d4 + 3, t[u].w, sat);
                if (t[u].x | t[u].y | t[u].z | t[u].w) reinterpret_cast<uint4*>(tab)[u * kSubApplyThreads + threadIdx.x] = make_uint4(0, 0, 0, 0);
            }
        }
        __syncthreads();
    }
}

} // namespace pg
