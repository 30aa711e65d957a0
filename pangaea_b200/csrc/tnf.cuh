// tnf.cuh - per-cloud tetranucleotide (tnf_k-mer) tallies of the sliced path.
//
// Replaces bin/count_tnf's countKmer (src/cpptools/count_tnf.cpp:78-113: ++map[canonical
// 4-mer] per valid window, columns in ascending canonical-code order, :138-164).
//
// Streaming kernel over the packed bases, one 32-base word per thread.  The hot loop does no
// canonicalisation at all: the 2 tnf_k-bit window code as it sits in the stream IS the bin
// (4^tnf_k block-private shared-memory bins per cloud slot); a window and its reverse
// complement are folded into their common column (lut) only when a cloud's bins leave shared
// memory - once per cloud per CTA, not once per base.  Invalid windows hit a dummy bin so the
// 32 shared atomics of a word are issued without divergence.
//
// A CTA owns a contiguous range of words and keeps the bins of the clouds under its cursor
// (kTnfSlots of them at tnf_k <= 4, fewer for wider windows; clouds beyond them go to global reductions).  Words
// that hold a cloud boundary (scan.cuh: kWordMixed, one per cloud) resolve the cloud per
// position and go straight to the global matrix.
//
// Bound: shared-memory atomic throughput (one per base; profiles/microbench_r01.txt:
// ~1.7-2.3 T/s) - HBM traffic is 0.375 B/base in, 4 * tnf_dim B per row out.
#pragma once
#include "bucket.cuh"

namespace pg {

constexpr int kTnfThreads = 256;
constexpr int kTnfFoldMaxK = 4; // up to this tnf_k the flush folds a cloud's bins into its tnf_dim columns in shared memory (136 at k = 4)
constexpr int kTnfSlots = 32; // clouds of a 256-word tile with private bins: 1 KB each at tnf_k = 4 (one cloud per read pair still fits)

template <int TK>
__global__ void __launch_bounds__(kTnfThreads)
tnf_kernel(const FeatParams P)
{
    extern __shared__ uint32_t smem[];
    const int tk = TK ? TK : P.tnf_k;
    const int nb = 1 << (2 * tk);                                  // raw bins per slot
    const int n_slots = P.tnf_slots;                               // host: as many as fit (api.cu)
    uint32_t* bins = smem;                                         // [n_slots][nb] + 1 dummy
    uint16_t* lut_s = reinterpret_cast<uint16_t*>(bins + n_slots * nb + 1);
    uint32_t* canon = reinterpret_cast<uint32_t*>(lut_s + nb);    // [warps][td] folded rows of the flush (tnf_k <= 4, api.cu)
    uint32_t* direct = canon + (kTnfThreads / 32) * P.td;          // [2] a word of the tile went straight to the global matrix
    const bool fold = tk <= kTnfFoldMaxK && P.tnf_fold != 0;
    for (int i = threadIdx.x; i < n_slots * nb + 1; i += blockDim.x) bins[i] = 0u;
    for (int i = threadIdx.x; i < nb; i += blockDim.x) lut_s[i] = P.lut[i];
    if (fold && threadIdx.x < 2) direct[threadIdx.x] = 0u;
    __syncthreads();
    uint32_t flushes = 0u; // parity of the flag in use

    const int64_t w_begin = (int64_t)blockIdx.x * P.words_per_cta;
    const int64_t w_end = min(P.n_words, w_begin + P.words_per_cta);
    if (w_begin >= w_end) return;
    const uint32_t tmask = (uint32_t)nb - 1u;
    const uint32_t dummy = (uint32_t)(n_slots * nb);

    for (int64_t tile = w_begin; tile < w_end; tile += kTnfThreads) {
        const int64_t tile_end = min(tile + (int64_t)kTnfThreads, w_end);
        const uint32_t g_lo = __ldg(P.wg + tile) & ~kWordMixed;
        const uint32_t gw_last = __ldg(P.wg + tile_end - 1);
        const uint32_t g_hi = (gw_last & ~kWordMixed) + ((gw_last & kWordMixed) ? 1u : 0u); // clouds of the tile: g_lo .. g_hi (a cloud that
                                                                                             // starts inside the last word included)
        const int64_t j = tile + threadIdx.x;
        if (j + kTnfThreads < w_end) { // next tile of this CTA -> L2 while this one is tallied
            if ((threadIdx.x & 15) == 0) asm volatile("prefetch.global.L2 [%0];" ::"l"(P.codes + j + kTnfThreads));
            if ((threadIdx.x & 31) == 0) {
                asm volatile("prefetch.global.L2 [%0];" ::"l"(P.maskF + j + kTnfThreads));
                asm volatile("prefetch.global.L2 [%0];" ::"l"(P.wg + j + kTnfThreads));
            }
        }
        uint32_t mlo = 0u;
        if (j < tile_end) mlo = __ldg(P.maskF + j);
        if (mlo != 0u) {
            const uint32_t gw = __ldg(P.wg + j);
            const uint32_t g = gw & ~kWordMixed;
            const uint32_t mhi = __ldg(P.maskF + j + 1);
            const uint32_t tvalid = window_valid_mask(mlo, mhi, tk);
            if (!(gw & kWordMixed)) {
                const int32_t row = __ldg(P.row_of_group + g);
                if (row >= 0 && tvalid != 0u) {
                    const uint64_t lo = __ldg(P.codes + j), hi = __ldg(P.codes + j + 1);
                    const uint32_t s0 = (uint32_t)lo, s1 = (uint32_t)(lo >> 32), s2 = (uint32_t)hi;
                    const uint32_t slot = g - g_lo;
                    if (slot < (uint32_t)n_slots) { // block-private bins
                        uint32_t* my = bins + slot * nb;
                        const uint32_t dmy = dummy - slot * nb;
#pragma unroll
                        for (int i = 0; i < 32; ++i) {
                            const uint32_t u = i == 0 ? s0 : (i < 16 ? __funnelshift_r(s0, s1, 2 * i) : (i == 16 ? s1 : __funnelshift_r(s1, s2, 2 * i - 32)));
                            atomicAdd(my + ((tvalid & (1u << i)) ? (u & tmask) : dmy), 1u);
                        }
                    } else { // more clouds in this tile than slots (they own no slot: no conflict with the stores of the flush)
                        uint32_t* tnf_row = P.tnf + (int64_t)row * P.td;
                        for (int i = 0; i < 32; ++i)
                            if (tvalid & (1u << i)) {
                                const uint32_t u = (uint32_t)(i ? ((lo >> (2 * i)) | (hi << (64 - 2 * i))) : lo);
                                atomicAdd(tnf_row + lut_s[u & tmask], 1u);
                            }
                    }
                }
            } else if (tvalid != 0u) {
                const uint64_t lo = __ldg(P.codes + j), hi = __ldg(P.codes + j + 1);
                // a cloud boundary inside the word.  The usual case - ONE boundary, both clouds have bins in shared memory - is
                // tallied like any other word (with one cloud per read pair every tenth word is of this kind: 10 % of the
                // positions going to the global matrix one RED at a time cost that configuration 1 s per 300 M pairs)
                const uint32_t slot = g - g_lo;
                // (an empty last cloud - it starts where the stream ends - owns no base and does not count as a third one)
                const bool third = (int64_t)g + 2 < P.n_groups && __ldg(P.gstart + g + 2) < min((j + 1) * 32, P.n_bytes);
                if (!third && slot + 1u < (uint32_t)n_slots) {
                    const int split = (int)(__ldg(P.gstart + g + 1) - j * 32); // first base of the second cloud, 1 .. 31
                    const bool ok0 = __ldg(P.row_of_group + g) >= 0, ok1 = __ldg(P.row_of_group + g + 1) >= 0;
                    const uint32_t s0 = (uint32_t)lo, s1 = (uint32_t)(lo >> 32), s2 = (uint32_t)hi;
                    const uint32_t first = (1u << split) - 1u;
                    const uint32_t live = tvalid & ((ok0 ? first : 0u) | (ok1 ? ~first : 0u));
                    uint32_t* my = bins + slot * nb;
                    const uint32_t dmy = dummy - slot * nb;
#pragma unroll
                    for (int i = 0; i < 32; ++i) {
                        const uint32_t u = i == 0 ? s0 : (i < 16 ? __funnelshift_r(s0, s1, 2 * i) : (i == 16 ? s1 : __funnelshift_r(s1, s2, 2 * i - 32)));
                        atomicAdd(my + ((live & (1u << i)) ? (u & tmask) + (i >= split ? (uint32_t)nb : 0u) : dmy), 1u);
                    }
                } else {
                // three clouds in one word (clouds shorter than 32 bases) or no bins left: resolve the cloud per position
                if (fold) direct[flushes & 1u] = 1u; // these REDs may hit rows the flush would otherwise store
                int64_t gg = g;
                int64_t next_start = __ldg(P.gstart + gg + 1);
                int32_t row = __ldg(P.row_of_group + gg);
                for (int i = 0; i < 32; ++i) {
                    const int64_t q = j * 32 + i;
                    while (gg + 1 < P.n_groups && q >= next_start) {
                        ++gg;
                        next_start = __ldg(P.gstart + gg + 1);
                        row = __ldg(P.row_of_group + gg);
                    }
                    if (row < 0 || !((tvalid >> i) & 1u)) continue;
                    const uint32_t u = (uint32_t)(i ? ((lo >> (2 * i)) | (hi << (64 - 2 * i))) : lo);
                    atomicAdd(P.tnf + (int64_t)row * P.td + lut_s[u & tmask], 1u);
                }
                }
            }
        }
        // slot 0 stays in shared memory while the next tile continues the same single cloud (no barrier needed then)
        const bool carry = g_hi == g_lo && tile_end < w_end && (__ldg(P.wg + tile_end) & ~kWordMixed) == g_lo;
        if (!carry) {
            __syncthreads();
            const uint32_t n_here = g_hi - g_lo + 1u;
            const uint32_t ns = min((uint32_t)n_slots, n_here);
            if (fold) {
                // one warp per cloud slot (the loads of the row numbers of different slots are in flight together - a block-wide
                // loop over the slots serialises them: 27 dependent round trips per tile with one cloud per read pair).  The
                // 4^k raw bins are folded into the tnf_dim columns in shared memory; a cloud that lies strictly inside the
                // tile (not its first, not its last) is complete here and nobody else touches its row: plain coalesced
                // stores, zeros included.  The two clouds at the edges continue in other tiles: reductions.
                const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
                uint32_t* cw = canon + warp * P.td;
                const bool any_direct = direct[flushes & 1u] != 0u;
                if (threadIdx.x == 0) direct[(flushes + 1u) & 1u] = 0u; // (written by the words of the next tile, after the barrier below)
                for (uint32_t s = warp; s < ns; s += kTnfThreads / 32) {
                    const int32_t row = __ldg(P.row_of_group + g_lo + s);
                    if (row < 0) continue; // (dropped clouds were never tallied)
                    uint32_t* src = bins + s * nb;
                    for (int c = lane; c < P.td; c += 32) cw[c] = 0u;
                    __syncwarp();
                    for (int b = lane; b < nb; b += 32) {
                        const uint32_t v = src[b];
                        if (v) { atomicAdd(cw + lut_s[b], v); src[b] = 0u; }
                    }
                    __syncwarp();
                    const bool whole = s > 0u && s + 1u < n_here && !any_direct;
                    uint32_t* dst = P.tnf + (int64_t)row * P.td;
                    for (int c = lane; c < P.td; c += 32) {
                        const uint32_t v = cw[c];
                        if (whole) dst[c] = v;
                        else if (v) atomicAdd(dst + c, v);
                    }
                    __syncwarp();
                }
                ++flushes;
            } else {
                for (uint32_t s = 0; s < ns; ++s) {
                    const int32_t row = __ldg(P.row_of_group + g_lo + s);
                    if (row < 0) continue;
                    uint32_t* src = bins + s * nb;
                    uint32_t* dst = P.tnf + (int64_t)row * P.td;
                    for (int b = threadIdx.x; b < nb; b += blockDim.x) {
                        const uint32_t v = src[b];
                        if (v) { atomicAdd(dst + lut_s[b], v); src[b] = 0u; }
                    }
                }
            }
            __syncthreads();
        }
    }
}

} // namespace pg
