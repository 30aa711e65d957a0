// featurize.cuh - step 1b: per-cloud abundance histogram + TNF in ONE pass - the DIRECT path (hash table for k > 16,
// k = 16, tiny tables, PG_FORCE_DIRECT); for k <= 15 the sliced path of bucket.cuh / tnf.cuh is about 5x faster.
// FeatParams and the small helpers below are shared with that path.
//
// Replaces bin/count_kmer's countKmer (src/cpptools/count_kmer.cpp:55-108: rolling
// canonical k-mer -> global count c -> ++hist[c / w] if c / w < v; k-mers absent
// from the table are skipped) and bin/count_tnf's countKmer (count_tnf.cpp:78-113:
// ++map[canonical 4-mer], columns in ascending canonical-code order) - the reference
// reads the FASTQ twice for these; here both rolling values come from the same
// registers.
//
// Work decomposition.  The packed stream is cut into contiguous ranges, one per
// persistent CTA; a CTA walks its range tile by tile (one 32-base word per thread)
// and keeps the histogram rows of the clouds under its cursor in shared memory
// (block-private bins, kSlots rows of v + tnf_dim u32).  A row leaves shared memory
// only when the cursor leaves its cloud: non-zero bins are reduced into the zeroed
// global matrices with RED (other CTAs may hold other parts of the same cloud).
// Dropped clouds (empty label / too short - the big unbarcoded tail) are skipped
// without touching the table.  Clouds smaller than a tile overflow the slots and fall
// back to direct global reductions.
//
// Table look-ups are the cost: 32 independent gathers per thread are issued before
// any of them is consumed.
//
// HBM roofline: algorithmic 0.375 B/base of stream + 4 B per 15-mer window (counter
// read) + 4*(v + tnf_dim) B per emitted row; real traffic is one 32 B sector per
// look-up that misses L2.
#pragma once
#include "table.cuh"

namespace pg {

constexpr int kFeatThreads = 256;
constexpr int kSlots = 4;

struct FeatParams {
    const uint64_t* codes;
    const uint32_t* maskF;
    int64_t n_words;
    int64_t n_bytes;
    const int64_t* gstart;      // n_groups + 1
    int64_t n_groups;
    const int32_t* row_of_group;
    const uint32_t* wg;         // n_words: cloud of base 32 j | kWordMixed (scan.cuh) - sliced path and tnf.cuh
    const int32_t* row_lb;      // n_groups: rows emitted before cloud g
    int32_t tnf_k, vs, td;
    int32_t tnf_slots;          // tnf.cuh: cloud slots with block-private bins
    int32_t tnf_fold;           // tnf.cuh: flush with one warp per slot, folded columns, stores for whole clouds
    uint32_t ws, clamp;         // clamp = min(ws * vs, 2^32-1): counts >= clamp fall outside the histogram
    uint32_t magic;             // ceil(2^32 / ws) when use_magic
    int32_t use_magic;
    const uint16_t* lut;        // 4^tnf_k entries: LSB-first window -> TNF column
    uint32_t* abd;              // [rows, vs]
    uint32_t* tnf;              // [rows, td]
    int64_t words_per_cta;
    TableView table;
};

// largest g >= g0 with gstart[g] <= p  (gallop + bisect; clouds under a cursor that
// moves forward are found in O(1))
__device__ __forceinline__ int64_t advance_group(const int64_t* __restrict__ gstart, int64_t n_groups, int64_t g0, int64_t p)
{
    int64_t lo = g0, step = 1;
    while (lo + step < n_groups && __ldg(gstart + lo + step) <= p) { lo += step; step <<= 1; }
    int64_t hi = min(lo + step, n_groups);
    while (hi - lo > 1) {
        int64_t mid = (lo + hi) >> 1;
        if (__ldg(gstart + mid) <= p) lo = mid; else hi = mid;
    }
    return lo;
}

__device__ __forceinline__ uint32_t abd_bin(const FeatParams& P, uint32_t c)
{
    // int pos = count / bin_size  (count_kmer.cpp:90); caller guarantees c < clamp
    return P.use_magic ? (uint32_t)(((uint64_t)c * P.magic) >> 32) : c / P.ws;
}

// one position, row pointers already resolved (used by the slow paths)
template <int MODE>
__device__ __forceinline__ void feat_one(const FeatParams& P, uint64_t lo, uint64_t hi, uint32_t mlo, uint32_t mhi, int i,
                                         uint32_t* abd_row, uint32_t* tnf_row, const uint16_t* lut_s)
{
    const int k = P.table.k;
    const uint32_t km = (1u << k) - 1u, tm = (1u << P.tnf_k) - 1u;
    const uint32_t mw = __funnelshift_r(mlo, mhi, i);
    const uint64_t win = i ? ((lo >> (2 * i)) | (hi << (64 - 2 * i))) : lo;
    if ((mw & tm) == tm) atomicAdd(tnf_row + lut_s[(uint32_t)win & ((1u << (2 * P.tnf_k)) - 1u)], 1u);
    if ((mw & km) == km) {
        const uint64_t w = win & low_mask64(2 * k);
        uint32_t c = MODE == kDense ? __ldg(P.table.counts + dense_index_of_window((uint32_t)w, k))
                                    : table_get_hash(P.table, canonical_of_window(w, k));
        if (c != 0u && (c & kCountMask) < P.clamp) atomicAdd(abd_row + abd_bin(P, c & kCountMask), 1u);
    }
}

template <int MODE>
__global__ void __launch_bounds__(kFeatThreads)
featurize_kernel(const FeatParams P)
{
    extern __shared__ uint32_t smem[];
    const int rowlen = P.vs + P.td;
    uint32_t* bins = smem;                                        // [kSlots][rowlen]
    uint16_t* lut_s = reinterpret_cast<uint16_t*>(smem + kSlots * rowlen);
    const int lut_n = 1 << (2 * P.tnf_k);
    for (int i = threadIdx.x; i < kSlots * rowlen; i += blockDim.x) bins[i] = 0u;
    for (int i = threadIdx.x; i < lut_n; i += blockDim.x) lut_s[i] = P.lut[i];
    __syncthreads();

    const int64_t w_begin = (int64_t)blockIdx.x * P.words_per_cta;
    const int64_t w_end = min(P.n_words, w_begin + P.words_per_cta);
    if (w_begin >= w_end) return;

    const int k = P.table.k;
    const uint32_t km = (1u << k) - 1u, tm = (1u << P.tnf_k) - 1u;
    const uint32_t tmask = (1u << (2 * P.tnf_k)) - 1u;
    const uint64_t wmask = low_mask64(2 * k);

    int64_t g_cur = advance_group(P.gstart, P.n_groups, 0, w_begin * 32);
    int64_t tile = w_begin;
    while (tile < w_end) {
        const int64_t tile_end = min(tile + (int64_t)blockDim.x, w_end);
        const int64_t p0 = tile * 32, p1 = min(tile_end * 32, P.n_bytes); // positions [p0, p1)
        const int64_t g_lo = advance_group(P.gstart, P.n_groups, g_cur, p0);
        const int64_t g_hi = advance_group(P.gstart, P.n_groups, g_lo, p1 - 1);
        g_cur = g_lo;
        const bool single = (g_hi == g_lo);
        if (single && __ldg(P.row_of_group + g_lo) < 0) {
            // whole tile inside a dropped cloud: jump the cursor to the tile holding its end
            // (nothing of this cloud is ever accumulated, so the slots stay clean)
            const int64_t nxt = __ldg(P.gstart + g_lo + 1) >> 5;
            const int64_t jump = w_begin + ((nxt - w_begin) / blockDim.x) * blockDim.x;
            tile = max(tile + (int64_t)blockDim.x, jump);
            continue;
        }

        const int64_t j = tile + threadIdx.x;
        uint32_t mlo = 0u;
        if (j < tile_end) mlo = __ldg(P.maskF + j);
        if (mlo != 0u) {
            const uint32_t mhi = __ldg(P.maskF + j + 1);
            const uint64_t lo = __ldg(P.codes + j), hi = __ldg(P.codes + j + 1);
            const int64_t q0 = j * 32;
            int64_t g = single ? g_lo : advance_group(P.gstart, P.n_groups, g_lo, q0);
            const bool uniform = single || (g + 1 >= P.n_groups) || (__ldg(P.gstart + g + 1) > q0 + 31);
            if (uniform) {
                const int32_t row = __ldg(P.row_of_group + g);
                if (row >= 0) {
                    const int64_t slot = g - g_lo;
                    if (slot < kSlots) {
                        // ---- fast path: 32 windows, one cloud, block-private bins ----
                        uint32_t* abd_row = bins + slot * rowlen;
                        uint32_t* tnf_row = abd_row + P.vs;
                        uint32_t cnt[32];
#pragma unroll
                        for (int i = 0; i < 32; ++i) {
                            const uint32_t mw = __funnelshift_r(mlo, mhi, i);
                            cnt[i] = 0u;
                            if ((mw & km) == km) {
                                const uint64_t w = (i ? ((lo >> (2 * i)) | (hi << (64 - 2 * i))) : lo) & wmask;
                                cnt[i] = MODE == kDense ? __ldg(P.table.counts + dense_index_of_window((uint32_t)w, k))
                                                        : table_get_hash(P.table, canonical_of_window(w, k));
                            }
                        }
#pragma unroll
                        for (int i = 0; i < 32; ++i) {
                            const uint32_t mw = __funnelshift_r(mlo, mhi, i);
                            if ((mw & tm) == tm) {
                                const uint32_t w4 = (uint32_t)(i ? ((lo >> (2 * i)) | (hi << (64 - 2 * i))) : lo) & tmask;
                                atomicAdd(tnf_row + lut_s[w4], 1u);
                            }
                        }
                        // abundance bins: neighbouring windows mostly share a bin -> merge runs in registers
                        uint32_t cur = 0xFFFFFFFFu, run = 0u;
#pragma unroll
                        for (int i = 0; i < 32; ++i) {
                            uint32_t c = cnt[i];
                            if (c == 0u) continue; // absent from the table (count_kmer.cpp:87)
                            c &= kCountMask;
                            if (c >= P.clamp) continue;
                            const uint32_t b = abd_bin(P, c);
                            if (b == cur) { ++run; continue; }
                            if (run) atomicAdd(abd_row + cur, run);
                            cur = b; run = 1u;
                        }
                        if (run) atomicAdd(abd_row + cur, run);
                    } else {
                        // cloud beyond the slots (clouds much smaller than a tile): global reductions
                        uint32_t* abd_row = P.abd + (int64_t)row * P.vs;
                        uint32_t* tnf_row = P.tnf + (int64_t)row * P.td;
                        for (int i = 0; i < 32; ++i) feat_one<MODE>(P, lo, hi, mlo, mhi, i, abd_row, tnf_row, lut_s);
                    }
                }
            } else {
                // ---- a cloud boundary falls inside this word: resolve the cloud per position ----
                int64_t next_start = __ldg(P.gstart + g + 1);
                int32_t row = __ldg(P.row_of_group + g);
                for (int i = 0; i < 32; ++i) {
                    const int64_t q = q0 + i;
                    while (g + 1 < P.n_groups && q >= next_start) {
                        ++g;
                        next_start = __ldg(P.gstart + g + 1);
                        row = __ldg(P.row_of_group + g);
                    }
                    if (row < 0 || !((mlo >> i) & 1u)) continue;
                    const int64_t slot = g - g_lo;
                    uint32_t* abd_row = slot < kSlots ? bins + slot * rowlen : P.abd + (int64_t)row * P.vs;
                    uint32_t* tnf_row = slot < kSlots ? bins + slot * rowlen + P.vs : P.tnf + (int64_t)row * P.td;
                    feat_one<MODE>(P, lo, hi, mlo, mhi, i, abd_row, tnf_row, lut_s);
                }
            }
        }
        __syncthreads();

        // keep slot 0 only when the next tile continues the same single cloud
        const bool carry = single && tile_end < w_end && (g_lo + 1 >= P.n_groups || __ldg(P.gstart + g_lo + 1) > p1);
        if (!carry) {
            const int64_t ns = min((int64_t)kSlots, g_hi - g_lo + 1);
            for (int64_t s = 0; s < ns; ++s) {
                const int32_t row = __ldg(P.row_of_group + g_lo + s);
                uint32_t* src = bins + s * rowlen;
                if (row >= 0) {
                    uint32_t* abd_row = P.abd + (int64_t)row * P.vs;
                    uint32_t* tnf_row = P.tnf + (int64_t)row * P.td;
                    for (int b = threadIdx.x; b < rowlen; b += blockDim.x) {
                        const uint32_t v = src[b];
                        if (v) {
                            atomicAdd(b < P.vs ? abd_row + b : tnf_row + (b - P.vs), v);
                            src[b] = 0u;
                        }
                    }
                }
            }
            __syncthreads();
        }
        tile = tile_end;
    }
}

} // namespace pg
