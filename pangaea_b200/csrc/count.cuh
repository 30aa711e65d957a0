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
        uint64_t cur = 0;
        uint32_t run = 0;
#pragma unroll
        for (int i = 0; i < 32; ++i) {
            const uint32_t mw = __funnelshift_r(mlo, mhi, i);
            if ((mw & km) != km) continue;
            const uint64_t w = (i ? ((lo >> (2 * i)) | (hi << (64 - 2 * i))) : lo) & wmask;
            uint64_t key;
            if (MODE == kDense) key = dense_index_of_window((uint32_t)w, k);
            else key = canonical_of_window(w, k);
            if (run && key == cur) { ++run; continue; }
            if (run) {
                if (MODE == kDense) table_add_dense(t, (uint32_t)cur, run);
                else table_add_hash(t, cur, run);
            }
            cur = key;
            run = 1;
        }
        if (run) {
            if (MODE == kDense) table_add_dense(t, (uint32_t)cur, run);
            else table_add_hash(t, cur, run);
        }
    }
}

} // namespace pg
