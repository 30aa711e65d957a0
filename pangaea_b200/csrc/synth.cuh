// synth.cuh - synthetic linked reads generated directly in HBM (bench input only;
// SURVEY.md §8d).  Same model as pangaea_b200/synth.py: every barcode owns one
// fragment of one genome; pairs fall uniformly inside it, R2 is the reverse
// complement of the far end; substitutions and N at fixed rates.  Genomes are
// uniform-random, so "the genome" is a hash of (genome, position) and needs no
// storage.  Counter-based hashing makes the output a pure function of the seed.
// Layout written: read bytes + '\n' separator per read, R1 then R2 of each pair.
#pragma once
#include <stdint.h>
#include <cuda_runtime.h>
#include "kmer.cuh"

namespace pg {

struct SynthParams {
    int64_t n_pairs;
    int32_t read_len;
    int64_t n_barcodes;
    const int64_t* bc_start;   // n_barcodes + 1: first pair of every barcode
    const int32_t* bc_genome;  // genome of every barcode
    int64_t genome_len;
    int32_t frag_len, insert;
    uint32_t sub_thresh, n_thresh; // rate * 2^32
    uint64_t seed;
    int64_t bc_base, pair_base; // global index of this batch's first barcode / pair: batches and ranks that share a seed draw
                                // different clouds from the SAME community of genomes
};

__device__ __forceinline__ uint32_t synth_base(uint64_t seed, int32_t genome, int64_t pos)
{
    return (uint32_t)(mix64(seed ^ ((uint64_t)genome << 40) ^ (uint64_t)pos) >> 13) & 3u;
}

__global__ void __launch_bounds__(256)
synth_kernel(const SynthParams S, uint8_t* __restrict__ seq, int64_t* __restrict__ read_off, uint8_t* __restrict__ read_flag)
{
    const int rl = S.read_len + 1;
    const int64_t n_reads = 2 * S.n_pairs, n_bytes = n_reads * rl;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const char letters[4] = { 'A', 'C', 'G', 'T' };
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t * 16 < n_bytes; t += stride) {
        uint32_t out[4] = { 0, 0, 0, 0 };
        int64_t cached_read = -1, start = 0;
        int32_t genome = 0;
        for (int b = 0; b < 16; ++b) {
            const int64_t p = t * 16 + b;
            if (p >= n_bytes) break;
            const int64_t r = p / rl;
            const int i = (int)(p - r * rl);
            uint32_t ch = '\n';
            if (i < S.read_len) {
                if (r != cached_read) {
                    cached_read = r;
                    const int64_t pair = r >> 1;
                    int64_t lo = 0, hi = S.n_barcodes; // bc_start[lo] <= pair < bc_start[hi]
                    while (hi - lo > 1) {
                        int64_t mid = (lo + hi) >> 1;
                        if (__ldg(S.bc_start + mid) <= pair) lo = mid; else hi = mid;
                    }
                    genome = __ldg(S.bc_genome + lo);
                    const uint64_t hb = mix64(S.seed ^ 0x9E3779B97F4A7C15ull ^ (uint64_t)(lo + S.bc_base));
                    const uint64_t hp = mix64(S.seed ^ 0xD1B54A32D192ED03ull ^ (uint64_t)(pair + S.pair_base));
                    const int64_t frag = (int64_t)(hb % (uint64_t)(S.genome_len - S.frag_len + 1));
                    start = frag + (int64_t)(hp % (uint64_t)(S.frag_len - S.insert + 1));
                }
                uint32_t code;
                if ((r & 1) == 0) code = synth_base(S.seed, genome, start + i);
                else code = 3u - synth_base(S.seed, genome, start + S.insert - 1 - i); // complement in ACGT order
                const uint64_t e = mix64(S.seed ^ 0xA0761D6478BD642Full ^ (uint64_t)(p + S.pair_base * 2 * rl));
                if ((uint32_t)e < S.sub_thresh) code = (uint32_t)(e >> 40) & 3u;
                ch = letters[code];
                if ((uint32_t)(e >> 32) < S.n_thresh) ch = 'N';
            }
            out[b >> 2] |= ch << (8 * (b & 3));
        }
        if (t * 16 + 16 <= n_bytes) {
            reinterpret_cast<uint4*>(seq)[t] = make_uint4(out[0], out[1], out[2], out[3]);
        } else {
            for (int b = 0; b < 16 && t * 16 + b < n_bytes; ++b) seq[t * 16 + b] = (uint8_t)(out[b >> 2] >> (8 * (b & 3)));
        }
    }
    // offsets + flags: the R2 read of the first pair of every barcode carries PG_READ_CHANGE
    for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r <= n_reads; r += stride) {
        read_off[r] = r * rl;
        if (r < n_reads) read_flag[r] = 0;
    }
}

__global__ void synth_flags_kernel(const SynthParams S, uint8_t* __restrict__ read_flag)
{
    int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= S.n_barcodes) return;
    int64_t first = S.bc_start[b];
    if (first < S.bc_start[b + 1] && first < S.n_pairs) read_flag[2 * first + 1] = 1;
}

} // namespace pg
