// ingest.cuh - FASTQ text on the device: line index, header parse, barcode sort, record transforms.
//
// SURVEY.md §8f rows 1, 2, 4 - what sits either side of the featurization hot path in run_pangaea:
//   * the barcode sort         awk | LANG=C sort -k1,1 | cut -f2- | tr "\t" "\n"      (src/run_pangaea:237-252)
//   * the host FASTQ loops     getline + getBarcode                                    (count_kmer.cpp:25-53,236-282)
//   * the format converters    preprocess_stlfr / preprocess_tellseq                   (src/cpptools/preprocess_*.cpp)
//   * extract_reads            barcode -> cluster routing of records                   (src/cpptools/extract_reads.cpp)
// All of them are byte shuffles over the same text, so they share one toolkit: the text lives in HBM, a line index is
// built with a scan over newline counts, every record (8 lines of an interleaved pair / 4 lines of a single FASTQ) gets
// one thread for its header logic, output sizes go through an exclusive scan, and warps copy the bytes.  Everything here
// is HBM-bound byte work; nothing is a contraction.
//
// Line semantics are std::getline's / awk's: '\n' ends a line and is not part of it, a last line without '\n' counts.
#pragma once
#include <stdint.h>
#include <cuda_runtime.h>

namespace pg {

constexpr int kTextTile = 16384;  // bytes per CTA of the newline kernels (256 threads x 64 bytes)

// ---------------------------------------------------------------------------
// generic exclusive scan over int64 (three kernels; tiles of 4096 items)
// ---------------------------------------------------------------------------
constexpr int kScan64Threads = 256, kScan64Items = 16, kScan64Tile = kScan64Threads * kScan64Items;

__device__ __forceinline__ long long block_scan64(long long v, long long& total)
{
    __shared__ long long ws[kScan64Threads / 32];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    long long inc = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const long long t = __shfl_up_sync(0xffffffffu, inc, d);
        if (lane >= d) inc += t;
    }
    if (lane == 31) ws[wid] = inc;
    __syncthreads();
    if (wid == 0) {
        long long s = lane < kScan64Threads / 32 ? ws[lane] : 0;
#pragma unroll
        for (int d = 1; d < kScan64Threads / 32; d <<= 1) {
            const long long t = __shfl_up_sync(0xffffffffu, s, d);
            if (lane >= d) s += t;
        }
        if (lane < kScan64Threads / 32) ws[lane] = s;
    }
    __syncthreads();
    const long long off = wid ? ws[wid - 1] : 0;
    total = ws[kScan64Threads / 32 - 1];
    __syncthreads();
    return off + inc - v;
}

__global__ void __launch_bounds__(kScan64Threads) scan64_reduce_kernel(const long long* __restrict__ in, int64_t n, long long* __restrict__ tile_sum)
{
    const int64_t base = (int64_t)blockIdx.x * kScan64Tile + (int64_t)threadIdx.x * kScan64Items;
    long long c = 0;
#pragma unroll
    for (int i = 0; i < kScan64Items; ++i) if (base + i < n) c += in[base + i];
    long long total;
    block_scan64(c, total);
    if (threadIdx.x == 0) tile_sum[blockIdx.x] = total;
}

__global__ void __launch_bounds__(kScan64Threads) scan64_tiles_kernel(long long* __restrict__ tile_sum, int64_t n_tiles, long long* __restrict__ total_out)
{
    __shared__ long long carry_s;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    for (int64_t base = 0; base < n_tiles; base += kScan64Threads) {
        const int64_t i = base + threadIdx.x;
        const long long v = i < n_tiles ? tile_sum[i] : 0;
        long long total;
        const long long ex = block_scan64(v, total);
        const long long carry = carry_s;
        if (i < n_tiles) tile_sum[i] = carry + ex;
        __syncthreads();
        if (threadIdx.x == 0) carry_s = carry + total;
        __syncthreads();
    }
    if (threadIdx.x == 0 && total_out) *total_out = carry_s;
}

// out[i] = exclusive prefix of in[0..i); out may alias in; out[n] is NOT written (the total goes to total_out of the tiles kernel)
__global__ void __launch_bounds__(kScan64Threads) scan64_apply_kernel(const long long* in, int64_t n, const long long* __restrict__ tile_off,
                                                                      long long* out) // (in / out unqualified: they may alias)
{
    const int64_t base = (int64_t)blockIdx.x * kScan64Tile + (int64_t)threadIdx.x * kScan64Items;
    long long v[kScan64Items], c = 0;
#pragma unroll
    for (int i = 0; i < kScan64Items; ++i) { v[i] = base + i < n ? in[base + i] : 0; c += v[i]; }
    long long total;
    long long run = tile_off[blockIdx.x] + block_scan64(c, total);
#pragma unroll
    for (int i = 0; i < kScan64Items; ++i) {
        if (base + i < n) out[base + i] = run;
        run += v[i];
    }
}

// ---------------------------------------------------------------------------
// line index: start offset of every line
// ---------------------------------------------------------------------------
__device__ __forceinline__ int nl_in_word(uint32_t w)
{
    const uint32_t x = w ^ 0x0A0A0A0Au;                               // bytes equal to '\n' become 0
    const uint32_t z = (x - 0x01010101u) & ~x & 0x80808080u;          // exact zero-byte test needs care with borrows:
    // the classic test can flag a byte 0x01 that follows a zero byte; count exactly instead
    (void)z;
    return ((x & 0xFFu) == 0) + ((x & 0xFF00u) == 0) + ((x & 0xFF0000u) == 0) + ((x & 0xFF000000u) == 0);
}

// newlines per tile of kTextTile bytes
__global__ void __launch_bounds__(256) nl_count_kernel(const uint8_t* __restrict__ text, int64_t n, long long* __restrict__ tile_count)
{
    const int64_t base = (int64_t)blockIdx.x * kTextTile + (int64_t)threadIdx.x * 64;
    int c = 0;
    if (base + 64 <= n && ((uintptr_t)(text + base) & 15) == 0) {
        const uint4* p = reinterpret_cast<const uint4*>(text + base);
#pragma unroll
        for (int i = 0; i < 4; ++i) { const uint4 v = p[i]; c += nl_in_word(v.x) + nl_in_word(v.y) + nl_in_word(v.z) + nl_in_word(v.w); }
    } else {
        for (int i = 0; i < 64 && base + i < n; ++i) c += text[base + i] == '\n';
    }
    long long total;
    block_scan64(c, total);
    if (threadIdx.x == 0) tile_count[blockIdx.x] = total;
}

// line_start[k + 1] = position after the k-th newline; tile_off = exclusive scan of nl_count_kernel's output
__global__ void __launch_bounds__(256) nl_fill_kernel(const uint8_t* __restrict__ text, int64_t n, const long long* __restrict__ tile_off,
                                                      long long* __restrict__ line_start)
{
    const int64_t base = (int64_t)blockIdx.x * kTextTile + (int64_t)threadIdx.x * 64;
    uint8_t b[64];
    int c = 0;
#pragma unroll
    for (int i = 0; i < 64; ++i) { b[i] = base + i < n ? text[base + i] : 0; c += b[i] == '\n'; }
    long long total;
    long long k = tile_off[blockIdx.x] + block_scan64(c, total);
#pragma unroll
    for (int i = 0; i < 64; ++i)
        if (b[i] == '\n') line_start[++k] = base + i + 1;
}

// ---------------------------------------------------------------------------
// small string helpers over the text (device versions of std::string::find / find_first_of)
// ---------------------------------------------------------------------------
constexpr long long kNpos = -1;

__device__ __forceinline__ long long find_char(const uint8_t* s, long long len, uint8_t c, long long from)
{
    for (long long i = from; i < len; ++i) if (s[i] == c) return i;
    return kNpos;
}
// std::string::find("BX:Z")
__device__ __forceinline__ long long find_bxz(const uint8_t* s, long long len)
{
    for (long long i = 0; i + 4 <= len; ++i)
        if (s[i] == 'B' && s[i + 1] == 'X' && s[i + 2] == ':' && s[i + 3] == 'Z') return i;
    return kNpos;
}
__device__ __forceinline__ bool is_space_c(uint8_t c) { return c == ' ' || (c >= 9 && c <= 13); } // isspace in the C locale

struct TextLines {
    const uint8_t* text;
    const long long* line_start; // n_lines + 1 entries; line i = [start[i], start[i+1] - 1) (the newline, real or virtual, excluded)
    long long n_lines;
};
__device__ __forceinline__ const uint8_t* line_ptr(const TextLines& T, long long i, long long* len)
{
    *len = T.line_start[i + 1] - 1 - T.line_start[i];
    return T.text + T.line_start[i];
}

// ---------------------------------------------------------------------------
// getBarcode (count_kmer.cpp:25-53 == count_tnf.cpp:23-52 == extract_reads.cpp:11-39) on one header line.
// read_type: 0 undecided, 1 "10x", 2 "stLFR" (latched by the first decisive header of the FILE: the caller resolves it).
// Offsets are relative to the line; std::string::substr clamping and the size_t wrap of npos + 1 are mirrored
// (as csrc/fastq.cpp: HeaderParser::parse does on the host).
// ---------------------------------------------------------------------------
struct Span { long long off, len; };

__device__ __forceinline__ Span clamp_substr(long long len, unsigned long long pos, unsigned long long n)
{
    Span r = { 0, 0 };
    if (pos > (unsigned long long)len) return r;  // the reference would throw; the host parser yields ""
    if (n > (unsigned long long)len - pos) n = (unsigned long long)len - pos;
    r.off = (long long)pos; r.len = (long long)n;
    return r;
}

__device__ __forceinline__ void get_barcode(const uint8_t* s, long long len, int read_type, Span* name, Span* bc)
{
    const unsigned long long NP = ~0ull;
    if (read_type == 2) {
        const long long p1s = find_char(s, len, '#', 0);
        const unsigned long long p1 = p1s < 0 ? NP : (unsigned long long)p1s;
        const unsigned long long from = p1 + 1ull; // npos + 1 == 0
        const long long p2s = from >= (unsigned long long)len ? kNpos : find_char(s, len, '/', (long long)from);
        const unsigned long long p2 = p2s < 0 ? NP : (unsigned long long)p2s;
        *name = clamp_substr(len, 0, p1);
        *bc = clamp_substr(len, p1 + 1ull, p2 - p1 - 1ull);
        if (bc->len == 5 && s[bc->off] == '0' && s[bc->off + 1] == '_' && s[bc->off + 2] == '0' && s[bc->off + 3] == '_' && s[bc->off + 4] == '0') bc->len = 0;
    } else {
        long long e = kNpos;
        for (long long i = 0; i < len; ++i)
            if (s[i] == ' ' || s[i] == '\r' || s[i] == '\t' || s[i] == '\n') { e = i; break; }
        *name = clamp_substr(len, 0, e < 0 ? NP : (unsigned long long)e);
        bc->off = 0; bc->len = 0;
        const long long p1 = find_bxz(s, len);
        if (p1 >= 0) {
            const long long p2s = p1 + 5 >= len ? kNpos : find_char(s, len, '-', p1 + 5);
            const unsigned long long p2 = p2s < 0 ? NP : (unsigned long long)p2s;
            *bc = clamp_substr(len, (unsigned long long)p1 + 5ull, p2 - (unsigned long long)p1 - 5ull);
        }
    }
}

// first decisive header (interleaved: every 8th line; stride 4 for a single FASTQ): packed = (record << 2) | type
__global__ void latch_kernel(TextLines T, int stride, long long n_rec, unsigned long long* __restrict__ packed_min)
{
    const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_rec) return;
    long long len;
    const uint8_t* s = line_ptr(T, r * stride, &len);
    int type = 0;
    if (find_bxz(s, len) >= 0) type = 1;
    else if (find_char(s, len, '#', 0) >= 0) type = 2;
    if (type) atomicMin(packed_min, ((unsigned long long)r << 2) | (unsigned long long)type);
}

// per record of an interleaved file: barcode span (absolute offsets into the text) under the latched read type
__global__ void barcode_kernel(TextLines T, long long n_rec, unsigned long long latch, long long* __restrict__ bc_off, int32_t* __restrict__ bc_len)
{
    const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_rec) return;
    const long long latch_rec = latch == ~0ull ? 0x7FFFFFFFFFFFFFFFll : (long long)(latch >> 2);
    const int type = r < latch_rec ? 0 : (int)(latch & 3ull);
    long long len;
    const uint8_t* s = line_ptr(T, r * 8, &len);
    Span name, bc;
    get_barcode(s, len, type, &name, &bc);
    bc_off[r] = T.line_start[r * 8] + bc.off;
    bc_len[r] = (int32_t)bc.len;
}

__device__ __forceinline__ bool span_equal(const uint8_t* text, long long a_off, int a_len, long long b_off, int b_len)
{
    if (a_len != b_len) return false;
    for (int i = 0; i < a_len; ++i) if (text[a_off + i] != text[b_off + i]) return false;
    return true;
}

// Reads of an interleaved batch: lines 8r + 1 and 8r + 5 (when they exist).  Writes the per-read byte count (len + 1 for
// the separator) and the change flag on R2 (count_kmer.cpp:251: barcode != last_barcode), comparing with the previous
// record's barcode (record 0: with the carried last_barcode, `carry` = carry_len bytes).
__global__ void reads_kernel(TextLines T, long long n_rec, const long long* __restrict__ bc_off, const int32_t* __restrict__ bc_len,
                             const uint8_t* __restrict__ carry, int carry_len, long long* __restrict__ read_bytes /* 2 per record */,
                             uint8_t* __restrict__ read_flag /* 2 per record */, long long* __restrict__ change /* 1 per record */)
{
    const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_rec) return;
    const long long l1 = r * 8 + 1, l2 = r * 8 + 5;
    long long len;
    read_bytes[2 * r] = 0; read_bytes[2 * r + 1] = 0;
    read_flag[2 * r] = 0; read_flag[2 * r + 1] = 0;
    change[r] = 0;
    if (l1 < T.n_lines) { line_ptr(T, l1, &len); read_bytes[2 * r] = len + 1; }
    if (l2 < T.n_lines) {
        line_ptr(T, l2, &len);
        read_bytes[2 * r + 1] = len + 1;
        bool same;
        if (r == 0) {
            same = bc_len[0] == carry_len;
            for (int i = 0; same && i < carry_len; ++i) same = T.text[bc_off[0] + i] == carry[i];
        } else {
            same = span_equal(T.text, bc_off[r], bc_len[r], bc_off[r - 1], bc_len[r - 1]);
        }
        if (!same) { read_flag[2 * r + 1] = 1; change[r] = 1; }
    }
}

// copy the sequence lines into the batch layout: one warp per read slot (2 per record); read_start = exclusive scan of read_bytes
__global__ void __launch_bounds__(256) copy_reads_kernel(TextLines T, long long n_rec, const long long* __restrict__ read_start,
                                                         const long long* __restrict__ read_bytes, uint8_t* __restrict__ seq)
{
    const long long w = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (w >= 2 * n_rec) return;
    const long long nb = read_bytes[w];
    if (!nb) return;
    const long long line = (w >> 1) * 8 + ((w & 1) ? 5 : 1);
    const uint8_t* src = T.text + T.line_start[line];
    uint8_t* dst = seq + read_start[w];
    for (long long i = lane; i < nb - 1; i += 32) dst[i] = src[i];
    if (lane == 0) dst[nb - 1] = '\n';
}

// quality lines in the same layout (0xFF under the separator), only when the ctx filters by quality
__global__ void __launch_bounds__(256) copy_quals_kernel(TextLines T, long long n_rec, const long long* __restrict__ read_start,
                                                         const long long* __restrict__ read_bytes, uint8_t* __restrict__ qual)
{
    const long long w = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (w >= 2 * n_rec) return;
    const long long nb = read_bytes[w];
    if (!nb) return;
    const long long line = (w >> 1) * 8 + ((w & 1) ? 7 : 3);
    long long qlen = 0;
    const uint8_t* src = nullptr;
    if (line < T.n_lines) src = line_ptr(T, line, &qlen);
    uint8_t* dst = qual + read_start[w];
    for (long long i = lane; i < nb; i += 32) dst[i] = (i < nb - 1 && i < qlen) ? src[i] : (uint8_t)0xFF;
}

// compact the labels of the clouds a batch opens: label k (k >= 1) = barcode of the k-th record that carries a change flag
__global__ void label_spans_kernel(long long n_rec, const long long* __restrict__ change, const long long* __restrict__ change_rank,
                                   const long long* __restrict__ bc_off, const int32_t* __restrict__ bc_len, long long* __restrict__ lab_off,
                                   long long* __restrict__ lab_len)
{
    const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_rec || !change[r]) return;
    lab_off[change_rank[r]] = bc_off[r];
    lab_len[change_rank[r]] = bc_len[r];
}
__global__ void label_copy_kernel(const uint8_t* __restrict__ text, long long n_lab, const long long* __restrict__ lab_off, const long long* __restrict__ lab_len,
                                  const long long* __restrict__ lab_start, uint8_t* __restrict__ blob)
{
    const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n_lab) return;
    for (long long i = 0; i < lab_len[k]; ++i) blob[lab_start[k] + i] = text[lab_off[k] + i];
}

// ---------------------------------------------------------------------------
// barcode sort (run_pangaea:237-252).  Order = bytewise order of the line awk prints: tag "\t" line0 "\t" ... "\t" line7,
// tag = first match of BX:Z:[^[:space:]]+ in the header or "~~~" (GNU sort -k1,1 in the C locale, last-resort comparison
// on the whole line; a tag holds no byte <= '\t', so comparing whole lines gives the same order as key-then-line).
// ---------------------------------------------------------------------------
constexpr int kSortKeyBytes = 96; // leading bytes of the comparison string that the radix passes see; deeper ties are fixed on the host

struct SortRec { long long tag_off; int32_t tag_len; }; // tag_len < 0: "~~~"

__device__ __forceinline__ uint8_t sort_string_byte(const TextLines& T, long long rec, const SortRec& sr, long long j, long long rec_end)
{
    // j-th byte of: tag '\t' block, block = the record's bytes with '\n' -> '\t', the final newline dropped
    const long long tl = sr.tag_len < 0 ? 3 : sr.tag_len;
    if (j < tl) return sr.tag_len < 0 ? (uint8_t)'~' : T.text[sr.tag_off + j];
    if (j == tl) return '\t';
    const long long p = T.line_start[rec * 8] + (j - tl - 1);
    if (p >= rec_end) return 0; // past the end: shorter strings sort first
    const uint8_t c = T.text[p];
    return c == '\n' ? (uint8_t)'\t' : c;
}

// end of the record's last line (exclusive, the newline not included)
__device__ __forceinline__ long long record_end(const TextLines& T, long long rec)
{
    const long long last = min(rec * 8 + 8, T.n_lines);
    return T.line_start[last] - 1;
}

__global__ void sort_tag_kernel(TextLines T, long long n_rec, SortRec* __restrict__ recs, long long* __restrict__ out_bytes, uint32_t* __restrict__ bad)
{
    const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_rec) return;
    long long len;
    const uint8_t* s = line_ptr(T, r * 8, &len);
    if (len == 0 || s[0] != '@' || r * 8 + 8 > T.n_lines) *bad = 1u; // awk's /^@/ + 7 getlines: only well-formed input is handled here
    SortRec sr = { 0, -1 };
    for (long long i = 0; i + 5 < len + 0; ++i) { // BX:Z: followed by at least one non-space byte
        if (s[i] == 'B' && s[i + 1] == 'X' && s[i + 2] == ':' && s[i + 3] == 'Z' && s[i + 4] == ':' && !is_space_c(s[i + 5])) {
            long long e = i + 5;
            while (e < len && !is_space_c(s[e])) ++e;
            sr.tag_off = T.line_start[r * 8] + i; sr.tag_len = (int32_t)(e - i);
            break;
        }
    }
    recs[r] = sr;
    out_bytes[r] = record_end(T, r) + 1 - T.line_start[r * 8]; // every line of the record followed by one newline
}

// keysT[j * n_rec + r] = j-th byte of the record's comparison string; col_seen[j * 256 + b] != 0 when byte value b occurs in column j
__global__ void __launch_bounds__(256) sort_keys_kernel(TextLines T, long long n_rec, const SortRec* __restrict__ recs, uint8_t* __restrict__ keysT,
                                                        uint32_t* __restrict__ col_seen)
{
    const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_rec) return;
    const SortRec sr = recs[r];
    const long long e = record_end(T, r);
    for (int j = 0; j < kSortKeyBytes; ++j) {
        const uint8_t b = sort_string_byte(T, r, sr, j, e);
        keysT[(long long)j * n_rec + r] = b;
        if (!col_seen[j * 256 + b]) col_seen[j * 256 + b] = 1u; // benign race: only ever set
    }
}

// ---- stable LSD radix pass over one key column: one warp owns kRadixChunk consecutive positions ----
constexpr int kRadixChunk = 2048;
constexpr int kRadixWarps = 8;

__global__ void __launch_bounds__(kRadixWarps * 32) radix_hist_kernel(const uint8_t* __restrict__ col, const uint32_t* __restrict__ perm, long long n,
                                                                      long long n_chunks, long long* __restrict__ hist /* [256][n_chunks] */)
{
    __shared__ uint32_t cnt[kRadixWarps][256];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const long long chunk = (long long)blockIdx.x * kRadixWarps + w;
    for (int i = lane; i < 256; i += 32) cnt[w][i] = 0u;
    __syncwarp();
    if (chunk < n_chunks) {
        const long long lo = chunk * kRadixChunk, hi = min(n, lo + kRadixChunk);
        for (long long i = lo + lane; i < hi; i += 32) atomicAdd(&cnt[w][col[perm[i]]], 1u);
        __syncwarp();
        for (int d = lane; d < 256; d += 32) hist[(long long)d * n_chunks + chunk] = cnt[w][d];
    }
}

__global__ void __launch_bounds__(kRadixWarps * 32) radix_scatter_kernel(const uint8_t* __restrict__ col, const uint32_t* __restrict__ perm_in, long long n,
                                                                         long long n_chunks, const long long* __restrict__ base /* scanned hist */,
                                                                         uint32_t* __restrict__ perm_out)
{
    __shared__ long long off[kRadixWarps][256];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const long long chunk = (long long)blockIdx.x * kRadixWarps + w;
    if (chunk >= n_chunks) return;
    for (int d = lane; d < 256; d += 32) off[w][d] = base[(long long)d * n_chunks + chunk];
    __syncwarp();
    const long long lo = chunk * kRadixChunk, hi = min(n, lo + kRadixChunk);
    for (long long i0 = lo; i0 < hi; i0 += 32) {
        const long long i = i0 + lane;
        const bool live = i < hi;
        const uint32_t p = live ? perm_in[i] : 0u;
        const uint32_t d = live ? col[p] : 0xFFFFFFFFu;
        const uint32_t peers = __match_any_sync(0xffffffffu, d);
        const uint32_t rank = __popc(peers & ((1u << lane) - 1u));
        if (live) perm_out[off[w][d] + rank] = p;
        __syncwarp();
        if (live && rank == 0) off[w][d] += __popc(peers);
        __syncwarp();
    }
}

__global__ void iota_kernel(uint32_t* __restrict__ p, long long n)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = (uint32_t)i;
}

// tie[i] = 1 when sorted record i has the same kSortKeyBytes-byte key as record i - 1 (the host then compares deeper)
__global__ void sort_ties_kernel(const uint8_t* __restrict__ keysT, const uint32_t* __restrict__ perm, long long n, uint8_t* __restrict__ tie, uint32_t* __restrict__ any)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    bool same = i > 0;
    for (int j = 0; same && j < kSortKeyBytes; ++j) same = keysT[(long long)j * n + perm[i]] == keysT[(long long)j * n + perm[i - 1]];
    tie[i] = same;
    if (same) *any = 1u;
}

__global__ void gather_len_kernel(const long long* __restrict__ len, const uint32_t* __restrict__ perm, long long n, long long* __restrict__ out)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = len[perm[i]];
}

// sorted record i = input record perm[i]; every byte copied, '\t' -> '\n' (the script's tr), a missing final newline added
__global__ void __launch_bounds__(256) sort_copy_kernel(TextLines T, long long n_rec, const uint32_t* __restrict__ perm, const long long* __restrict__ out_start,
                                                        uint8_t* __restrict__ out)
{
    const long long w = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (w >= n_rec) return;
    const long long r = perm[w];
    const long long a = T.line_start[r * 8], e = record_end(T, r);
    uint8_t* dst = out + out_start[w];
    for (long long i = lane; i < e - a; i += 32) {
        const uint8_t c = T.text[a + i];
        dst[i] = c == '\t' ? (uint8_t)'\n' : c;
    }
    if (lane == 0) dst[e - a] = '\n';
}

} // namespace pg
