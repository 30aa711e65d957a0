// kmer.cuh - 2-bit k-mer arithmetic on the packed base stream (device + host).
//
// Stream layout ("LSB-first"): base at stream position p lives in bits
// [2*(p%32), 2*(p%32)+1] of codes64[p/32]; one validity bit per base in
// mask32[p/32] bit (p%32).  A k-mer window starting at p is the 2k-bit field
// w = stream[2p .. 2p+2k), i.e. its FIRST base sits in the LOW bits.
//
// Reference arithmetic (citations relative to /root/reference/):
//   - code = (c >> 1) & 3  => A0 C1 T2 G3            src/cpptools/count_kmer.cpp:81
//   - val  = first base in the HIGH bits              count_kmer.cpp:79-81
//   - rc   = reverse 2-bit groups, xor 0xAAAA.., shift  count_kmer.cpp:11-21
//   - key  = min(val, rc)                             count_kmer.cpp:86
// With w as above: val = reverse_groups(w) and rc(val) = w ^ 0xAAAA..(2k bits) -
// the window as stored already IS the reverse half of the reverse complement.
#pragma once
#include <stdint.h>

#ifdef __CUDACC__
#define PG_HD __host__ __device__ __forceinline__
#else
#define PG_HD inline
#endif

namespace pg {

PG_HD uint64_t low_mask64(int bits) { return bits >= 64 ? ~0ull : ((1ull << bits) - 1ull); }

PG_HD uint32_t rev_groups32(uint32_t x)
{
#ifdef __CUDA_ARCH__
    x = __brev(x);
#else
    x = ((x >> 1) & 0x55555555u) | ((x & 0x55555555u) << 1);
    x = ((x >> 2) & 0x33333333u) | ((x & 0x33333333u) << 2);
    x = ((x >> 4) & 0x0F0F0F0Fu) | ((x & 0x0F0F0F0Fu) << 4);
    x = ((x >> 8) & 0x00FF00FFu) | ((x & 0x00FF00FFu) << 8);
    x = (x >> 16) | (x << 16);
#endif
    return ((x >> 1) & 0x55555555u) | ((x & 0x55555555u) << 1);
}

PG_HD uint64_t rev_groups64(uint64_t x)
{
    return ((uint64_t)rev_groups32((uint32_t)x) << 32) | rev_groups32((uint32_t)(x >> 32));
}

// forward value (reference `val`) of an LSB-first window
PG_HD uint64_t fwd_of_window(uint64_t w, int k) { return rev_groups64(w) >> (64 - 2 * k); }
PG_HD uint32_t fwd_of_window32(uint32_t w, int k) { return rev_groups32(w) >> (32 - 2 * k); }

// reference canonical key of an LSB-first window, any k <= 31
PG_HD uint64_t canonical_of_window(uint64_t w, int k)
{
    uint64_t f = fwd_of_window(w, k);
    uint64_t r = w ^ (0xAAAAAAAAAAAAAAAAull & low_mask64(2 * k));
    return f < r ? f : r;
}

// reference canonical key of a forward-form value (host side of table import/export)
PG_HD uint64_t canonical_of_fwd(uint64_t v, int k)
{
    uint64_t r = (rev_groups64(v) >> (64 - 2 * k)) ^ (0xAAAAAAAAAAAAAAAAull & low_mask64(2 * k));
    return v < r ? v : r;
}

// ---------------------------------------------------------------------------
// Dense (direct-addressed) counter index for k <= 16.
//   even k : index = reference canonical key, 4^k counters.
//   odd  k : v and rc(v) have complementary middle bases (codes c and c^2), so
//            exactly one of them has bit k clear; take that one and squeeze the
//            bit out: a bijection {v, rc(v)} -> [0, 4^k / 2).  Halves the table
//            (k = 15: 2^29 counters = 2 GiB) and needs no min().
// ---------------------------------------------------------------------------
PG_HD uint64_t dense_entries(int k) { return (k & 1) ? (1ull << (2 * k - 1)) : (1ull << (2 * k)); }

PG_HD uint32_t dense_index_of_window(uint32_t w, int k)
{
    uint32_t f = fwd_of_window32(w, k);
    uint32_t r = w ^ (0xAAAAAAAAu & (uint32_t)low_mask64(2 * k));
    if (k & 1) {
        uint32_t x = ((f >> k) & 1u) ? r : f;
        return ((x >> (k + 1)) << k) | (x & ((1u << k) - 1u));
    }
    return f < r ? f : r;
}

PG_HD uint64_t dense_index_of_fwd(uint64_t v, int k)
{
    uint64_t r = (rev_groups64(v) >> (64 - 2 * k)) ^ (0xAAAAAAAAAAAAAAAAull & low_mask64(2 * k));
    if (k & 1) {
        uint64_t x = ((v >> k) & 1ull) ? r : v;
        return ((x >> (k + 1)) << k) | (x & ((1ull << k) - 1ull));
    }
    return v < r ? v : r;
}

// dense index -> reference canonical key
PG_HD uint64_t key_of_dense_index(uint64_t idx, int k)
{
    if (k & 1) {
        uint64_t x = ((idx >> k) << (k + 1)) | (idx & ((1ull << k) - 1ull));
        return canonical_of_fwd(x, k);
    }
    return idx;
}

// 64-bit finaliser (splitmix64) - slot hash of the open-addressing table
PG_HD uint64_t mix64(uint64_t h)
{
    h ^= h >> 30; h *= 0xbf58476d1ce4e5b9ull;
    h ^= h >> 27; h *= 0x94d049bb133111ebull;
    h ^= h >> 31;
    return h;
}

} // namespace pg
