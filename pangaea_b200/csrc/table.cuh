// table.cuh - the global canonical k-mer count table resident in HBM.
//
// Stands for jellyfish's hash + `dump` + the unordered_map that count_kmer rebuilds
// from the dump text (src/feature.py:76-103, src/cpptools/count_kmer.cpp:139-170).
// Two layouts behind one view:
//   DENSE (k <= 16): direct-addressed u32 counters, index = dense_index (kmer.cuh);
//         the degenerate open-addressing table whose hash is the identity and whose
//         probe sequence has length 1.  k = 15 (the production default): 2^29
//         counters = 2 GiB - 1 % of a B200's HBM.
//   HASH  (k <= 31): lock-free open addressing, linear probing, u64 keys claimed with
//         atomicCAS, u32 counters bumped with atomicAdd (no-return => RED).  Key and counter
//         sit in one 16-byte slot, i.e. in the same 32 B DRAM sector: an insert or a look-up
//         that finds its key in the first slot touches one sector, not two.
// Counter updates are fire-and-forget reductions; nothing waits on a round trip.
#pragma once
#include <stdint.h>
#include <cuda_runtime.h>
#include "kmer.cuh"

namespace pg {

constexpr uint64_t kEmptyKey = ~0ull;
// A dump line "KMER\t0" makes the k-mer PRESENT with frequency 0 in the reference
// (count_kmer.cpp:166 then :87-93 -> bin 0).  jellyfish never writes such a line, but
// pg_table_set honours it: the entry is stored as this marker, and readers mask it off.
constexpr uint32_t kPresentZero = 0x80000000u;
constexpr uint32_t kCountMask = 0x7FFFFFFFu;
enum TableMode { kDense = 0, kHash = 1 };

struct HashSlot {
    unsigned long long key;
    uint32_t count;
    uint32_t pad;
};
static_assert(sizeof(HashSlot) == 16, "one slot = half a DRAM sector");

struct TableView {
    uint32_t* counts;              // dense: counters
    HashSlot* slots;               // hash: key + counter per slot
    uint64_t capacity_mask;        // hash: slots - 1
    uint32_t* overflow;            // hash: set to 1 when an insert finds no slot
    int k;
};

__device__ __forceinline__ void table_add_dense(const TableView& t, uint32_t idx, uint32_t n)
{
    atomicAdd(t.counts + idx, n); // result unused -> RED.E.ADD
}

__device__ __forceinline__ void table_add_hash(const TableView& t, uint64_t key, uint32_t n)
{
    uint64_t slot = mix64(key) & t.capacity_mask;
    for (uint64_t probe = 0; probe <= t.capacity_mask; ++probe) {
        unsigned long long cur = *((volatile unsigned long long*)&t.slots[slot].key);
        if (cur == kEmptyKey) cur = atomicCAS(&t.slots[slot].key, kEmptyKey, (unsigned long long)key);
        if (cur == kEmptyKey || cur == key) {
            atomicAdd(&t.slots[slot].count, n);
            return;
        }
        slot = (slot + 1) & t.capacity_mask;
    }
    *t.overflow = 1u;
}

__device__ __forceinline__ uint32_t table_get_hash(const TableView& t, uint64_t key)
{
    uint64_t slot = mix64(key) & t.capacity_mask;
    for (uint64_t probe = 0; probe <= t.capacity_mask; ++probe) {
        const ulonglong2 s2 = __ldg(reinterpret_cast<const ulonglong2*>(t.slots + slot)); // key and counter in one 16 B load
        if (s2.x == key) return (uint32_t)s2.y;
        if (s2.x == kEmptyKey) return 0u;
        slot = (slot + 1) & t.capacity_mask;
    }
    return 0u;
}

// --- host-driven table maintenance (parity tooling; off the hot path) ---------

// kmer2frequency[key] = count  (count_kmer.cpp:166) - keys arrive as forward values
__global__ void table_set_kernel(TableView t, int mode, const uint64_t* __restrict__ keys, const uint32_t* __restrict__ counts, int64_t n)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint64_t v = keys[i] & low_mask64(2 * t.k);
    if (mode == kDense) {
        t.counts[dense_index_of_fwd(v, t.k)] = counts[i] ? counts[i] : kPresentZero;
    } else {
        uint64_t key = canonical_of_fwd(v, t.k);
        uint64_t slot = mix64(key) & t.capacity_mask;
        for (uint64_t probe = 0; probe <= t.capacity_mask; ++probe) {
            unsigned long long cur = atomicCAS(&t.slots[slot].key, kEmptyKey, (unsigned long long)key);
            if (cur == kEmptyKey || cur == key) { t.slots[slot].count = counts[i] ? counts[i] : kPresentZero; return; }
            slot = (slot + 1) & t.capacity_mask;
        }
        *t.overflow = 1u;
    }
}

__global__ void table_get_kernel(TableView t, int mode, const uint64_t* __restrict__ keys, uint32_t* __restrict__ out, int64_t n)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint64_t v = keys[i] & low_mask64(2 * t.k);
    out[i] = (mode == kDense ? t.counts[dense_index_of_fwd(v, t.k)] : table_get_hash(t, canonical_of_fwd(v, t.k))) & kCountMask;
}

// empty hash table: every key = kEmptyKey, every counter = 0
__global__ void hash_clear_kernel(HashSlot* __restrict__ slots, uint64_t n)
{
    uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    const ulonglong2 empty = make_ulonglong2(kEmptyKey, 0ull);
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) reinterpret_cast<ulonglong2*>(slots)[i] = empty;
}

// number of non-zero counters (distinct k-mers); stride_words = 1 for the dense counters, 4 for the counter of a HashSlot
__global__ void table_nonzero_kernel(const uint32_t* __restrict__ counts, uint64_t n, int stride_words, unsigned long long* __restrict__ total)
{
    uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    unsigned long long c = 0;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) c += counts[i * stride_words] != 0;
#pragma unroll
    for (int d = 16; d; d >>= 1) c += __shfl_xor_sync(0xffffffffu, c, d);
    if ((threadIdx.x & 31) == 0 && c) atomicAdd(total, c);
}

// unordered export of (reference canonical key, count) for non-zero counters
__global__ void table_export_kernel(TableView t, int mode, uint64_t n_slots, uint64_t* __restrict__ keys_out,
                                    uint32_t* __restrict__ counts_out, unsigned long long cap, unsigned long long* __restrict__ cursor)
{
    uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_slots; i += stride) {
        uint32_t c = mode == kDense ? t.counts[i] : t.slots[i].count;
        if (!c) continue;
        unsigned long long at = atomicAdd(cursor, 1ull);
        if (at < cap) {
            keys_out[at] = mode == kDense ? key_of_dense_index(i, t.k) : (uint64_t)t.slots[i].key;
            counts_out[at] = c & kCountMask;
        }
    }
}

} // namespace pg
