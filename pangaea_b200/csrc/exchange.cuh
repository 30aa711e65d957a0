// exchange.cuh - the owner-partitioned k-mer table across ranks (SURVEY.md §8e "table partitioning", the north star's
// all-to-all): owner = hash(canonical k-mer) mod n_ranks.  Count pass: every rank turns its windows into keys, buckets them
// by owner, ONE all-to-all moves the keys, owners insert.  Featurize pass: the same all-to-all for the query keys, owners look
// the counts up, ONE all-to-all back; the counts return to stream order and are binned per cloud.
// This is the multi-GPU form of the HASH table (k > 16), which has no dense view to all-reduce; for k <= 15 the replicated
// dense table + all-reduce moves 40x fewer bytes (pangaea_b200/distributed.py) and stays the default.
// The kernels here are the device side; the collective itself is torch.distributed (NCCL) in distributed.py.
#pragma once
#include "bucket.cuh"

namespace pg {

__device__ __forceinline__ uint32_t owner_of_key(uint64_t key, uint32_t world)
{
    return (uint32_t)((mix64(key ^ 0x5851F42D4C957F2Dull) >> 33) % world); // (a different mix than the slot hash: owners must not correlate with slots)
}

// keys[(j - w0) * 32 + i] = canonical key of the window that starts at base i of word j, kEmptyKey where no valid window starts
__global__ void __launch_bounds__(256) window_keys_kernel(const uint64_t* __restrict__ codes, const uint32_t* __restrict__ mask, int64_t w0, int64_t w1, int k,
                                                          unsigned long long* __restrict__ keys)
{
    const int64_t j = w0 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= w1) return;
    const uint32_t mlo = mask[j], mhi = mask[j + 1];
    const uint32_t valid = mlo ? window_valid_mask(mlo, mhi, k) : 0u;
    const uint64_t lo = valid ? codes[j] : 0ull, hi = valid ? codes[j + 1] : 0ull;
    const uint64_t wmask = low_mask64(2 * k);
    unsigned long long* out = keys + (j - w0) * 32;
#pragma unroll 4
    for (int i = 0; i < 32; ++i) {
        unsigned long long key = kEmptyKey;
        if ((valid >> i) & 1u) {
            const uint64_t w = (i ? ((lo >> (2 * i)) | (hi << (64 - 2 * i))) : lo) & wmask;
            key = canonical_of_window(w, k);
        }
        out[i] = key;
    }
}

// keys per owner (block-private counters, then one atomic per owner and block)
__global__ void __launch_bounds__(256) owner_count_kernel(const unsigned long long* __restrict__ keys, int64_t n, uint32_t world, unsigned long long* __restrict__ counts)
{
    __shared__ unsigned int c[64];
    if (threadIdx.x < 64) c[threadIdx.x] = 0u;
    __syncthreads();
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const unsigned long long key = keys[i];
        if (key != kEmptyKey) atomicAdd(&c[owner_of_key(key, world)], 1u);
    }
    __syncthreads();
    if (threadIdx.x < world && c[threadIdx.x]) atomicAdd(counts + threadIdx.x, (unsigned long long)c[threadIdx.x]);
}

// sorted[cursor[owner]++] = key; dest[i] = where key i went (-1: no key).  Warp-aggregated cursor claims.
__global__ void __launch_bounds__(256) owner_scatter_kernel(const unsigned long long* __restrict__ keys, int64_t n, uint32_t world,
                                                            unsigned long long* __restrict__ cursors, unsigned long long* __restrict__ sorted,
                                                            long long* __restrict__ dest)
{
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int lane = threadIdx.x & 31;
    for (int64_t i0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x - lane; i0 < n; i0 += stride) {
        const int64_t i = i0 + lane;
        const unsigned long long key = i < n ? keys[i] : kEmptyKey;
        const bool live = key != kEmptyKey;
        const uint32_t o = live ? owner_of_key(key, world) : 0xFFFFFFFFu;
        const uint32_t peers = __match_any_sync(0xffffffffu, o);
        const int leader = __ffs(peers) - 1;
        unsigned long long base = 0ull;
        if (live && lane == leader) base = atomicAdd(cursors + o, (unsigned long long)__popc(peers));
        base = __shfl_sync(0xffffffffu, base, leader);
        if (i < n) {
            long long d = -1;
            if (live) { d = (long long)(base + __popc(peers & ((1u << lane) - 1u))); sorted[d] = key; }
            dest[i] = d;
        }
    }
}

__global__ void __launch_bounds__(256) table_add_keys_kernel(TableView t, int mode, const unsigned long long* __restrict__ keys, int64_t n)
{
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const unsigned long long key = keys[i];
        if (key == kEmptyKey) continue;
        if (mode == kDense) table_add_dense(t, (uint32_t)dense_index_of_fwd(key, t.k), 1u);
        else table_add_hash(t, key, 1u);
    }
}

__global__ void __launch_bounds__(256) table_lookup_keys_kernel(TableView t, int mode, const unsigned long long* __restrict__ keys, int64_t n, uint32_t* __restrict__ out)
{
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const unsigned long long key = keys[i];
        uint32_t c = 0u;
        if (key != kEmptyKey) c = mode == kDense ? __ldg(t.counts + dense_index_of_fwd(key, t.k)) : table_get_hash(t, key);
        out[i] = c; // kPresentZero marker kept: the binning masks it
    }
}

// counts back in stream order + abundance tallies: window i of word j belongs to the cloud of that base; equal (row, bin)
// pairs of a warp fold into one RED (bucket.cuh: abd_reduce_warp)
__global__ void __launch_bounds__(256) abd_from_counts_kernel(const FeatParams P, int64_t w0, int64_t w1, const long long* __restrict__ dest,
                                                              const uint32_t* __restrict__ counts_sorted)
{
    const int64_t n = (w1 - w0) * 32;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int lane = threadIdx.x & 31;
    for (int64_t i0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x - lane; i0 < n; i0 += stride) {
        const int64_t i = i0 + lane;                       // lane = base inside the word: a warp is one word
        const int64_t j = w0 + (i0 >> 5);
        bool live = false;
        unsigned long long key = 0ull;
        const long long d = i < n ? dest[i] : -1;
        if (d >= 0) {
            const uint32_t gw = __ldg(P.wg + j);
            int64_t g = gw & ~kWordMixed;
            if (gw & kWordMixed) { // a cloud boundary inside the word: resolve the cloud of this base
                const int64_t p = j * 32 + lane;
                while (g + 1 < P.n_groups && p >= __ldg(P.gstart + g + 1)) ++g;
            }
            const int32_t row = __ldg(P.row_of_group + g);
            if (row >= 0) live = abd_key(P, counts_sorted[d], (uint32_t)row, key);
        }
        abd_reduce_warp(P, live, key);
    }
}

} // namespace pg
