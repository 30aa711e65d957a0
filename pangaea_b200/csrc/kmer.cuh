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

// -DPG_DEBUG_BOUNDS: device-side range checks on the partition buffers (compute-sanitizer is not available on the
// measurement pool; `nvcc ... -DPG_DEBUG_BOUNDS -o build/libpg_debug.so`, run the GPU tests with PG_LIB_PATH pointing at it)
#if defined(PG_DEBUG_BOUNDS) && defined(__CUDA_ARCH__)
#include <cstdio>
#define PG_CHECK(cond)                                                                                    \
    do {                                                                                                  \
        if (!(cond)) { printf("PG_CHECK failed: %s (%s:%d)\n", #cond, __FILE__, __LINE__); __trap(); }    \
    } while (0)
#else
#define PG_CHECK(cond) ((void)0)
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
//   even k : class id = reference canonical key, 4^k counters.
//   odd  k : v and rc(v) have complementary middle bases (codes c and c^2), so
//            exactly one of them has bit k clear; take that one and squeeze the
//            bit out: a bijection {v, rc(v)} -> [0, 4^k / 2).  Halves the table
//            (k = 15: 2^29 counters = 2 GiB) and needs no min().
// The class id is then SCRAMBLED by an invertible map on its bit width (multiplication by
// an odd constant mod 2^bits - Fibonacci hashing, whose high bits mix every input bit).  The
// table is swept in contiguous slices (bucket.cuh); scrambling makes every slice
// receive the same share of the k-mers whatever the base composition of the
// genomes (an AT-rich community would otherwise fill some slices 2-4x more than
// others).  It is a bijection, so counts are unchanged; export inverts it.
// ---------------------------------------------------------------------------
PG_HD uint64_t dense_entries(int k) { return (k & 1) ? (1ull << (2 * k - 1)) : (1ull << (2 * k)); }
PG_HD int dense_bits(int k) { return (k & 1) ? 2 * k - 1 : 2 * k; }

constexpr uint32_t kMixA = 0x9E3779B1u; // odd: multiplication mod 2^bits is a bijection; the HIGH bits of the
                                        // product (the slice number) depend on every bit of the class id
constexpr uint32_t mod_inverse32(uint32_t a)
{
    uint32_t x = a; // Newton: doubles the number of correct low bits each round (a*a = 1 mod 8)
    for (int i = 0; i < 5; ++i) x *= 2u - a * x;
    return x;
}
constexpr uint32_t kMixAInv = mod_inverse32(kMixA);
static_assert(kMixA * kMixAInv == 1u, "modular inverse");

PG_HD uint32_t scramble_bits(uint32_t x, int bits)
{
    const uint32_t m = bits >= 32 ? 0xFFFFFFFFu : ((1u << bits) - 1u);
    return (x * kMixA) & m;
}
PG_HD uint32_t unscramble_bits(uint32_t y, int bits)
{
    const uint32_t m = bits >= 32 ? 0xFFFFFFFFu : ((1u << bits) - 1u);
    return (y * kMixAInv) & m;
}

// class id of the k-mer whose forward value is f and whose reverse complement is r
PG_HD uint32_t dense_index_of_pair(uint32_t f, uint32_t r, int k)
{
    uint32_t id;
    if (k & 1) {
        uint32_t x = ((f >> k) & 1u) ? r : f;
        id = ((x >> (k + 1)) << k) | (x & ((1u << k) - 1u));
    } else {
        id = f < r ? f : r;
    }
    return scramble_bits(id, dense_bits(k));
}

PG_HD uint32_t dense_index_of_window(uint32_t w, int k)
{
    return dense_index_of_pair(fwd_of_window32(w, k), w ^ (0xAAAAAAAAu & (uint32_t)low_mask64(2 * k)), k);
}

PG_HD uint64_t dense_index_of_fwd(uint64_t v, int k)
{
    uint64_t r = (rev_groups64(v) >> (64 - 2 * k)) ^ (0xAAAAAAAAAAAAAAAAull & low_mask64(2 * k));
    uint64_t id;
    if (k & 1) {
        uint64_t x = ((v >> k) & 1ull) ? r : v;
        id = ((x >> (k + 1)) << k) | (x & ((1ull << k) - 1ull));
    } else {
        id = v < r ? v : r;
    }
    return scramble_bits((uint32_t)id, dense_bits(k));
}

// dense index -> reference canonical key
PG_HD uint64_t key_of_dense_index(uint64_t idx, int k)
{
    uint64_t id = unscramble_bits((uint32_t)idx, dense_bits(k));
    if (k & 1) {
        uint64_t x = ((id >> k) << (k + 1)) | (id & ((1ull << k) - 1ull));
        return canonical_of_fwd(x, k);
    }
    return id;
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
