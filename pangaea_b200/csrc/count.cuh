// count.cuh - step 1a: every valid k-mer window of every read bumps its counter.
//
// Replaces `jellyfish count -C -m k` (call sites src/feature.py:76-94; external tool,
// parity unpinned - semantics restated in oracle/pg_oracle.c:pgo_count_read).
//
// Flat streaming kernel: the packed stream is one long array and read boundaries are
// just invalid positions in maskC, so there is no per-read work list and no ragged
// imbalance.  One thread owns the 32 windows that START in its 32-base word: it loads
// codes[j], codes[j+1], maskC[j], maskC[j+1] (coalesced 8 B / 4 B per lane) and slides
// a funnel shift over them.  Identical consecutive indices (homopolymer runs such as
// poly-G tails would otherwise serialise on one L2 atomic unit) are merged in
// registers before the reduction is issued.
//
// HBM roofline: algorithmic 0.375 B/base of stream + 8 B per window (u32 counter
// read-modify-write); real DRAM traffic is one 32 B sector in + out per counter
// that misses L2.
#pragma once
#include "table.cuh"

namespace pg {

// insert `n` occurrences of `key`, the home slot's key already loaded as `cur` (hash mode)
__device__ __forceinline__ void table_add_hash_probed(const TableView& t, uint64_t key, uint32_t n, uint64_t slot, unsigned long long cur)
{
    if (cur == key) { atomicAdd(&t.slots[slot].count, (unsigned long long)n); return; } // the usual case after the first occurrence: one RED
    table_add_hash(t, key, n);                                      // empty or taken home slot: claim / probe
}

template <int MODE>
__global__ void __launch_bounds__(256)
count_kernel(const uint64_t* __restrict__ codes, const uint32_t* __restrict__ maskC, int64_t n_words, TableView t)
{
    const int k = t.k;
    const uint32_t km = (k >= 32) ? 0xFFFFFFFFu : ((1u << k) - 1u);
    const uint64_t wmask = low_mask64(2 * k);
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n_words; j += stride) {
        const uint32_t mlo = maskC[j];
        if (mlo == 0u) continue; // no window can start on an invalid base
        const uint32_t mhi = maskC[j + 1];
        const uint64_t lo = codes[j], hi = codes[j + 1];
        if (MODE == kHash) {
            // eight windows at a time: their home slots are read together (eight sectors in flight per thread instead of
            // one), then every window whose key is already there costs a single RED
#pragma unroll 1
            for (int i0 = 0; i0 < 32; i0 += 8) {
                uint64_t key[8], slot[8];
                unsigned long long cur[8];
                uint32_t ok = 0u;
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const int i = i0 + u;
                    const uint32_t mw = __funnelshift_r(mlo, mhi, i);
                    const uint64_t w = (i ? ((lo >> (2 * i)) | (hi << (64 - 2 * i))) : lo) & wmask;
                    key[u] = canonical_of_window(w, k);
                    slot[u] = mix64(key[u]) & t.capacity_mask;
                    if ((mw & km) == km) ok |= 1u << u;
                }
#pragma unroll
                for (int u = 0; u < 8; ++u) cur[u] = (ok >> u) & 1u ? *((volatile unsigned long long*)&t.slots[slot[u]].key) : 0ull;
#pragma unroll
                for (int u = 0; u < 8; ++u)
                    if ((ok >> u) & 1u) table_add_hash_probed(t, key[u], 1u, slot[u], cur[u]);
            }
            continue;
        }
        uint64_t cur = 0;
        uint32_t run = 0;
#pragma unroll
        for (int i = 0; i < 32; ++i) {
            const uint32_t mw = __funnelshift_r(mlo, mhi, i);
            if ((mw & km) != km) continue;
            const uint64_t w = (i ? ((lo >> (2 * i)) | (hi << (64 - 2 * i))) : lo) & wmask;
            const uint64_t key = dense_index_of_window((uint32_t)w, k);
            if (run && key == cur) { ++run; continue; }
            if (run) table_add_dense(t, (uint32_t)cur, run);
            cur = key;
            run = 1;
        }
        if (run) table_add_dense(t, (uint32_t)cur, run);
    }
}

} // namespace pg
