// pack.cuh - ASCII bases -> 2-bit code stream + validity masks (one pass, streaming).
//
// One thread owns 32 consecutive bytes (two 128-bit coalesced loads) and emits one
// u64 of codes and one u32 per mask:
//   maskF  "feature" validity : byte is one of upper-case A C G T - the only bytes
//          count_kmer/count_tnf accept (count_kmer.cpp:73, count_tnf.cpp:91);
//   maskC  "count" validity   : A C G T in either case, and (optionally) quality
//          >= min_qual - what `jellyfish count [--min-qual-char]` accepts
//          (reference call sites src/feature.py:76-94).
// The per-read separator byte is not a base, so no k-mer window can span two reads
// (the reference gets the same effect from `line + "N"`, count_kmer.cpp:247).
// HBM roofline: reads 1 B/base (+1 with quality), writes 0.25 + 2*0.125 B/base.
#pragma once
#include <stdint.h>
#include <cuda_runtime.h>

namespace pg {

// bit (c & 31) set for A(1) C(3) G(7) T(20)  (scalar classification of the ragged tail)
constexpr uint32_t kLetterBits = (1u << 1) | (1u << 3) | (1u << 7) | (1u << 20);

// Four ASCII bytes at a time, SIMD within the 32-bit register (~25 integer ops per 4 bases
// instead of ~17 per base):
//   codes : x = (w >> 1) & 0x03030303 holds the 2-bit codes at byte stride; one multiply gathers
//           them into the top byte (fields are 2 bits wide and land on distinct bits: no carries).
//   letter: v = (w & 0xDF) ^ 0x41 per byte is 0x00 / 0x02 / 0x06 / 0x15 for A / C / G / T in either
//           case.  With a, b, c, d = bits 4, 2, 1, 0 of v moved to bit 0 of their byte, those four
//           values are exactly (~a & ~d & (c | ~b)) | (a & b & ~c & d) with bits 7, 6, 5, 3 clear.
//   masks : bit 0 of every byte -> 4 adjacent bits, again by one multiply.
__device__ __forceinline__ void pack_word(uint32_t w, uint32_t q, bool use_q, uint32_t minq, int base_bit,
                                          uint64_t& codes, uint32_t& mF, uint32_t& mC)
{
    const uint32_t x = (w >> 1) & 0x03030303u;
    const uint32_t c8 = (x * 0x01041040u) >> 24;
    const uint32_t v = (w & 0xDFDFDFDFu) ^ 0x41414141u;
    const uint32_t a = v >> 4, b = v >> 2, c = v >> 1, d = v;
    const uint32_t t1 = ~d & (c | ~b), t2 = b & ~c & d;
    const uint32_t f = (~a & t1) | (a & t2);
    const uint32_t h = (v >> 3) | (v >> 5) | (v >> 6) | (v >> 7);
    const uint32_t any = f & ~h & 0x01010101u;
    const uint32_t up = any & ~(w >> 5);
    uint32_t c4 = (any * 0x01020408u) >> 24;
    const uint32_t f4 = (up * 0x01020408u) >> 24;
    if (use_q) {
        uint32_t qm = 0u;
#pragma unroll
        for (int i = 0; i < 4; ++i) qm |= (uint32_t)(((q >> (8 * i)) & 0xFFu) >= minq) << i;
        c4 &= qm;
    }
    codes |= (uint64_t)c8 << (2 * base_bit);
    mF |= f4 << base_bit;
    mC |= c4 << base_bit;
}

// Packs words [w_begin, w_end) of the batch (a word = 32 bytes).  The output arrays hold
// n_words + 2 entries, n_words = ceil(n_bytes / 32); the launch that covers the end of the batch
// passes w_end = n_words + 2 and zeroes the two trailing pad words, so that "next word" reads
// never need a guard.  Ranges let the upload be pipelined: chunk c is packed while chunk c + 1
// is still crossing PCIe (api.cu: pg_extract_features).
__global__ void __launch_bounds__(256)
pack_kernel(const uint8_t* __restrict__ seq, const uint8_t* __restrict__ qual, int64_t n_bytes, int64_t w_begin, int64_t w_end,
            uint32_t minq, uint64_t* __restrict__ codes, uint32_t* __restrict__ maskF, uint32_t* __restrict__ maskC)
{
    const bool use_q = (qual != nullptr) && (minq != 0);
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t j = w_begin + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < w_end; j += stride) {
        uint64_t cw = 0;
        uint32_t mF = 0, mC = 0;
        int64_t b0 = j * 32;
        if (b0 + 32 <= n_bytes) {
            const uint4* p = reinterpret_cast<const uint4*>(seq + b0);
            uint4 a = __ldg(p), b = __ldg(p + 1);
            uint4 qa = make_uint4(0, 0, 0, 0), qb = qa;
            if (use_q) {
                const uint4* pq = reinterpret_cast<const uint4*>(qual + b0);
                qa = __ldg(pq); qb = __ldg(pq + 1);
            }
            pack_word(a.x, qa.x, use_q, minq, 0, cw, mF, mC);
            pack_word(a.y, qa.y, use_q, minq, 4, cw, mF, mC);
            pack_word(a.z, qa.z, use_q, minq, 8, cw, mF, mC);
            pack_word(a.w, qa.w, use_q, minq, 12, cw, mF, mC);
            pack_word(b.x, qb.x, use_q, minq, 16, cw, mF, mC);
            pack_word(b.y, qb.y, use_q, minq, 20, cw, mF, mC);
            pack_word(b.z, qb.z, use_q, minq, 24, cw, mF, mC);
            pack_word(b.w, qb.w, use_q, minq, 28, cw, mF, mC);
        } else if (b0 < n_bytes) { // ragged tail
            for (int i = 0; i < 32 && b0 + i < n_bytes; ++i) {
                uint32_t c = seq[b0 + i];
                uint32_t letter = (kLetterBits >> (c & 31u)) & 1u;
                uint32_t any = letter & (uint32_t)((c & 0xC0u) == 0x40u);
                uint32_t up = any & (uint32_t)((c & 0x20u) == 0u);
                if (use_q) any &= (uint32_t)(qual[b0 + i] >= minq);
                cw |= (uint64_t)((c >> 1) & 3u) << (2 * i);
                mF |= up << i;
                mC |= any << i;
            }
        }
        codes[j] = cw;
        maskF[j] = mF;
        maskC[j] = mC;
    }
}

// Clears maskF over [lo, hi) - used for PG_READ_NOFEAT reads (rare path: pairs whose
// R1/R2 names disagree, count_kmer.cpp:195) so the featurize kernel never sees them.
__global__ void clear_mask_ranges_kernel(const int64_t* __restrict__ read_off, const uint8_t* __restrict__ read_flag,
                                         int64_t n_reads, uint32_t* __restrict__ maskF)
{
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n_reads; r += stride) {
        if (!(read_flag[r] & 2)) continue;
        int64_t lo = read_off[r], hi = read_off[r + 1];
        for (int64_t w = lo >> 5; w <= (hi - 1) >> 5 && lo < hi; ++w) {
            int64_t a = max(lo, w << 5), b = min(hi, (w + 1) << 5); // [a, b) inside word w
            uint32_t bits = (uint32_t)(((1ull << (b - a)) - 1ull) << (a - (w << 5)));
            atomicAnd(&maskF[w], ~bits);
        }
    }
}

} // namespace pg
