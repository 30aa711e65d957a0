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
//         atomicCAS, u64 counters bumped with atomicAdd (no-return => RED).  Key and counter
//         sit in one 16-byte slot, i.e. in the same 32 B DRAM sector: an insert or a look-up
//         that finds its key in the first slot touches one sector, not two.
// Counter updates are fire-and-forget reductions; nothing waits on a round trip.
//
// Saturation.  jellyfish reports true counts and count_kmer drops a k-mer whose count / w >= v
// (count_kmer.cpp:90-92).  The dense u32 counters therefore SATURATE at kCountMax = 2^31 - 1
// instead of wrapping (a poly-G 15-mer of a multi-billion-read run gets there): pg_create
// insists on w * v <= 2^31 - 1, so a saturated counter is dropped exactly like the true one.
// How: the count pass works in segments of < 2^31 windows.  The read-modify-write merge of
// the shared-memory sub-tables (count2.cuh) clamps inline; every path that adds with an atomic
// uses table_add_checked, which raises a device flag when a sum reaches bit 31; after each
// segment table_saturate_kernel (a no-op while the flag is down) clamps the table.  So at
// segment boundaries every counter is <= kCountMax and within a segment none can wrap.
// Hash slots count in 64 bits and clamp on read.
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
constexpr uint32_t kCountMax = 0x7FFFFFFFu;              // dense counters saturate here
constexpr unsigned long long kPresentZero64 = 1ull << 63; // the same marker in a hash slot
enum TableMode { kDense = 0, kHash = 1 };

struct HashSlot {
    unsigned long long key;
    unsigned long long count;
};
static_assert(sizeof(HashSlot) == 16, "one slot = half a DRAM sector");

struct TableView {
    uint32_t* counts;              // dense: counters
    HashSlot* slots;               // hash: key + counter per slot
    uint64_t capacity_mask;        // hash: slots - 1
    uint32_t* overflow;            // hash: set to 1 when an insert finds no slot
    uint32_t* sat;                 // dense: set to 1 when a counter reaches bit 31 (table_saturate_kernel then clamps)
    int k;
};

// add n (< 2^31) to a dense counter with an atomic; the adder whose sum reaches bit 31 raises the flag
__device__ __forceinline__ void table_add_checked(uint32_t* counter, uint32_t n, uint32_t* sat)
{
    const uint32_t old = atomicAdd(counter, n);
    if ((old + n) & 0x80000000u) *sat = 1u;
}

__device__ __forceinline__ void table_add_dense(const TableView& t, uint32_t idx, uint32_t n)
{
    table_add_checked(t.counts + idx, n, t.sat);
}

// what a reader sees of a hash slot's 64-bit counter: the u32 convention of the dense table
__device__ __forceinline__ uint32_t hash_count32(unsigned long long c)
{
    if (c & kPresentZero64) return kPresentZero;
    return c > (unsigned long long)kCountMax ? kCountMax : (uint32_t)c;
}

__device__ __forceinline__ void table_add_hash(const TableView& t, uint64_t key, uint32_t n)
{
    uint64_t slot = mix64(key) & t.capacity_mask;
    for (uint64_t probe = 0; probe <= t.capacity_mask; ++probe) {
        unsigned long long cur = *((volatile unsigned long long*)&t.slots[slot].key);
        if (cur == kEmptyKey) cur = atomicCAS(&t.slots[slot].key, kEmptyKey, (unsigned long long)key);
        if (cur == kEmptyKey || cur == key) {
            atomicAdd(&t.slots[slot].count, (unsigned long long)n);
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
        if (s2.x == key) return hash_count32(s2.y);
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
        t.counts[dense_index_of_fwd(v, t.k)] = counts[i] ? min(counts[i], kCountMax) : kPresentZero;
    } else {
        uint64_t key = canonical_of_fwd(v, t.k);
        uint64_t slot = mix64(key) & t.capacity_mask;
        for (uint64_t probe = 0; probe <= t.capacity_mask; ++probe) {
            unsigned long long cur = atomicCAS(&t.slots[slot].key, kEmptyKey, (unsigned long long)key);
            if (cur == kEmptyKey || cur == key) { t.slots[slot].count = counts[i] ? (unsigned long long)counts[i] : kPresentZero64; return; }
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

// number of non-zero counters (distinct k-mers): the dense counters, or the 64-bit counters of the hash slots
__global__ void table_nonzero_kernel(const uint32_t* __restrict__ counts, const HashSlot* __restrict__ slots, uint64_t n, unsigned long long* __restrict__ total)
{
    uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    unsigned long long c = 0;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) c += counts ? counts[i] != 0u : slots[i].count != 0ull;
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
        uint32_t c = mode == kDense ? t.counts[i] : hash_count32(t.slots[i].count);
        if (!c) continue;
        unsigned long long at = atomicAdd(cursor, 1ull);
        if (at < cap) {
            keys_out[at] = mode == kDense ? key_of_dense_index(i, t.k) : (uint64_t)t.slots[i].key;
            counts_out[at] = c & kCountMask;
        }
    }
}

// clamp every dense counter to `limit`.  flag != nullptr: only when *flag is up (raised by table_add_checked);
// the host lowers it afterwards.  Counters that hold the kPresentZero marker (pg_table_set with count 0) never meet
// this kernel: pg_count refuses a table that holds markers.
__global__ void table_saturate_kernel(uint32_t* __restrict__ counts, uint64_t n, uint32_t limit, const uint32_t* __restrict__ flag)
{
    if (flag && *flag == 0u) return;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    uint4* c4 = reinterpret_cast<uint4*>(counts); // n is a multiple of 4 (4^k / 2 or 4^k, k >= 2) or tiny
    const uint64_t n4 = n / 4;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
        uint4 v = c4[i];
        if (v.x > limit || v.y > limit || v.z > limit || v.w > limit) {
            v.x = min(v.x, limit); v.y = min(v.y, limit); v.z = min(v.z, limit); v.w = min(v.w, limit);
            c4[i] = v;
        }
    }
    for (uint64_t i = n4 * 4 + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) counts[i] = min(counts[i], limit);
}

} // namespace pg
