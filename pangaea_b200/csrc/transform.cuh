// transform.cuh - record transforms over FASTQ text in HBM, on the toolkit of ingest.cuh:
//   stlfr_*     preprocess_stlfr  -n [-l]   (src/cpptools/preprocess_stlfr.cpp:76-115)
//   tellseq_*   preprocess_tellseq          (src/cpptools/preprocess_tellseq.cpp:52-84)
//   extract_*   extract_reads -i            (src/cpptools/extract_reads.cpp:86-124)
// Each is two kernels around an exclusive scan: sizes per line / record, then the bytes.  Where the reference would die
// on malformed input (std::string::at / replace throwing), a device flag is raised and the call fails with PG_ERR_INVALID.
#pragma once
#include "ingest.cuh"

namespace pg {

__device__ __forceinline__ void copy_bytes_warp(uint8_t* dst, const uint8_t* src, long long n, int lane)
{
    for (long long i = lane; i < n; i += 32) dst[i] = src[i];
}

// ---------------------------------------------------------------------------
// preprocess_stlfr -n: header `name#a_b_c/1` -> `name\tBX:Z:a_b_c[-1]` (or `name` when a or b is "0"), the SAME identifier
// in both output files; every other line copied.  Driven by the lines of file 1; a missing line of file 2 is "".
// ---------------------------------------------------------------------------
struct StlfrHeader { long long pos1, bc_len; bool barcoded, bad; };

__device__ __forceinline__ StlfrHeader stlfr_parse(const uint8_t* s, long long len)
{
    StlfrHeader h = { 0, 0, false, false };
    const long long p1 = find_char(s, len, '#', 0);
    if (p1 < 0) { h.bad = true; return h; }          // line1.replace(npos, ...) throws in the reference
    const long long p2 = p1 + 1 >= len ? kNpos : find_char(s, len, '/', p1 + 1);
    const long long bl = (p2 < 0 ? len : p2) - p1 - 1; // substr(pos1 + 1, pos2 - pos1 - 1)
    const uint8_t* b = s + p1 + 1;
    long long u1 = -1, u2 = -1;
    for (long long i = 0; i < bl; ++i)
        if (b[i] == '_') { if (u1 < 0) u1 = i; else { u2 = i; break; } }
    if (u1 < 0 || u2 < 0) { h.bad = true; return h; } // barcode.at(i) walks off the end
    const bool z1 = u1 == 1 && b[0] == '0', z2 = u2 - u1 - 1 == 1 && b[u1 + 1] == '0';
    h.pos1 = p1; h.bc_len = bl;
    h.barcoded = !z1 && !z2; // `bc1.compare("0") && bc2.compare("0") && bc1.compare("0")`: the third part is never looked at
    return h;
}

// out_len[i] = bytes line i contributes to an output file (newline included); the two files differ only in the copied lines
__global__ void stlfr_size_kernel(TextLines A, TextLines B, int library, long long* __restrict__ len1, long long* __restrict__ len2, uint32_t* __restrict__ bad)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= A.n_lines) return;
    long long la, lb = 0;
    const uint8_t* s = line_ptr(A, i, &la);
    if (i < B.n_lines) line_ptr(B, i, &lb);
    if (i % 4 == 0) {
        const StlfrHeader h = stlfr_parse(s, la);
        if (h.bad) { *bad = 1u; len1[i] = len2[i] = 0; return; }
        const long long id = h.barcoded ? h.pos1 + 6 + h.bc_len + (library ? 2 : 0) : h.pos1;
        len1[i] = len2[i] = id + 1;
    } else {
        len1[i] = la + 1; len2[i] = lb + 1;
    }
}

// one warp per line and output file
__global__ void __launch_bounds__(256) stlfr_write_kernel(TextLines A, TextLines B, int library, const long long* __restrict__ off1,
                                                         const long long* __restrict__ off2, uint8_t* __restrict__ out1, uint8_t* __restrict__ out2)
{
    const long long w = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    const long long i = w >> 1;
    const int which = (int)(w & 1);
    if (i >= A.n_lines) return;
    long long la, lb = 0;
    const uint8_t* sa = line_ptr(A, i, &la);
    const uint8_t* sb = sa;
    if (i < B.n_lines) sb = line_ptr(B, i, &lb);
    uint8_t* dst = which ? out2 + off2[i] : out1 + off1[i];
    if (i % 4 == 0) {
        const StlfrHeader h = stlfr_parse(sa, la);
        if (h.bad) return;
        copy_bytes_warp(dst, sa, h.pos1, lane);
        long long at = h.pos1;
        if (h.barcoded) {
            const char tag[7] = "\tBX:Z:";
            if (lane < 6) dst[at + lane] = (uint8_t)tag[lane];
            at += 6;
            copy_bytes_warp(dst + at, sa + h.pos1 + 1, h.bc_len, lane);
            at += h.bc_len;
            if (library) { if (lane == 0) dst[at] = '-'; if (lane == 1) dst[at + 1] = '1'; at += 2; }
        }
        if (lane == 0) dst[at] = '\n';
    } else {
        const uint8_t* s = which ? sb : sa;
        const long long l = which ? lb : la;
        copy_bytes_warp(dst, s, l, lane);
        if (lane == 0) dst[l] = '\n';
    }
}

// ---------------------------------------------------------------------------
// preprocess_tellseq: per record of file 1 (4 lines): header = R1 header up to the first ' ' + "\tBX:Z:" + index read + "-1"; the
// record is written to both files (and the barcode to the whitelist) only when the index read is 18 bases long.
// ---------------------------------------------------------------------------
__global__ void tellseq_size_kernel(TextLines A, TextLines B, TextLines I, long long n_rec, long long* __restrict__ len1, long long* __restrict__ len2,
                                    long long* __restrict__ len_wl)
{
    const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_rec) return;
    long long lh, ls1, lq1, ls2 = 0, lq2 = 0, lbc = 0;
    const uint8_t* h = line_ptr(A, 4 * r, &lh);
    line_ptr(A, 4 * r + 1, &ls1);
    line_ptr(A, 4 * r + 3, &lq1);
    if (4 * r + 1 < B.n_lines) line_ptr(B, 4 * r + 1, &ls2);
    if (4 * r + 3 < B.n_lines) line_ptr(B, 4 * r + 3, &lq2);
    if (4 * r + 1 < I.n_lines) line_ptr(I, 4 * r + 1, &lbc);
    if (lbc != 18) { len1[r] = len2[r] = len_wl[r] = 0; return; } // "Wrong barcode length."
    const long long sp = find_char(h, lh, ' ', 0);
    const long long hl = (sp < 0 ? lh : sp) + 6 + 18 + 2;
    len1[r] = hl + 1 + ls1 + 3 + lq1 + 1;
    len2[r] = hl + 1 + ls2 + 3 + lq2 + 1;
    len_wl[r] = 19;
}

__global__ void __launch_bounds__(256) tellseq_write_kernel(TextLines A, TextLines B, TextLines I, long long n_rec, const long long* __restrict__ len1,
                                                           const long long* __restrict__ off1, const long long* __restrict__ off2,
                                                           const long long* __restrict__ off_wl, uint8_t* __restrict__ out1, uint8_t* __restrict__ out2,
                                                           uint8_t* __restrict__ out_wl)
{
    const long long w = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    const long long r = w >> 1;
    const int which = (int)(w & 1);
    if (r >= n_rec || len1[r] == 0) return;
    long long lh, ls = 0, lq = 0, lbc;
    const uint8_t* h = line_ptr(A, 4 * r, &lh);
    const uint8_t* bc = line_ptr(I, 4 * r + 1, &lbc);
    const TextLines& S = which ? B : A;
    const uint8_t *s = h, *q = h;
    if (4 * r + 1 < S.n_lines) s = line_ptr(S, 4 * r + 1, &ls);
    if (4 * r + 3 < S.n_lines) q = line_ptr(S, 4 * r + 3, &lq);
    const long long sp = find_char(h, lh, ' ', 0);
    const long long nl = sp < 0 ? lh : sp;
    uint8_t* dst = which ? out2 + off2[r] : out1 + off1[r];
    copy_bytes_warp(dst, h, nl, lane);
    long long at = nl;
    const char tag[7] = "\tBX:Z:";
    if (lane < 6) dst[at + lane] = (uint8_t)tag[lane];
    at += 6;
    copy_bytes_warp(dst + at, bc, 18, lane);
    at += 18;
    if (lane == 0) { dst[at] = '-'; dst[at + 1] = '1'; dst[at + 2] = '\n'; }
    at += 3;
    copy_bytes_warp(dst + at, s, ls, lane);
    at += ls;
    if (lane == 0) { dst[at] = '\n'; dst[at + 1] = '+'; dst[at + 2] = '\n'; }
    at += 3;
    copy_bytes_warp(dst + at, q, lq, lane);
    at += lq;
    if (lane == 0) dst[at] = '\n';
    if (!which) {
        uint8_t* wl = out_wl + off_wl[r];
        copy_bytes_warp(wl, bc, 18, lane);
        if (lane == 0) wl[18] = '\n';
    }
}

// ---------------------------------------------------------------------------
// extract_reads -i: pairs whose barcode belongs to a cluster are appended to that cluster's .fq (header rewritten to
// name + "\tBX:Z:" + barcode + "-1", the other seven lines as they are) and their barcode to its .barcode file.
// run_of_rec = index of the record's barcode run (ingest.cuh: change flags); cluster_of_run comes from the host (label -> cluster).
// ---------------------------------------------------------------------------
__global__ void extract_size_kernel(TextLines T, long long n_rec, unsigned long long latch, const long long* __restrict__ run_of_rec,
                                    const int32_t* __restrict__ cluster_of_run, const int32_t* __restrict__ bc_len, long long* __restrict__ fq_len,
                                    long long* __restrict__ bc_out_len, uint8_t* __restrict__ key_lo, uint8_t* __restrict__ key_hi)
{
    const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_rec) return;
    const int32_t c = cluster_of_run[run_of_rec[r]];
    key_lo[r] = (uint8_t)(c < 0 ? 0xFF : (c & 0xFF));
    key_hi[r] = (uint8_t)(c < 0 ? 0xFF : ((c >> 8) & 0xFF));
    if (c < 0 || r * 8 + 7 >= T.n_lines) { fq_len[r] = 0; bc_out_len[r] = 0; return; } // the record is written when its 8th line arrives
    const long long latch_rec = latch == ~0ull ? 0x7FFFFFFFFFFFFFFFll : (long long)(latch >> 2);
    const int type = r < latch_rec ? 0 : (int)(latch & 3ull);
    long long lh;
    const uint8_t* h = line_ptr(T, r * 8, &lh);
    Span name, bc;
    get_barcode(h, lh, type, &name, &bc);
    const long long body = (T.line_start[r * 8 + 8] - 1) - T.line_start[r * 8 + 1]; // lines 1..7 with the newlines between them
    fq_len[r] = name.len + 6 + bc.len + 3 + body + 1;
    bc_out_len[r] = bc_len[r] + 1;
}

// sorted position i holds record perm[i]; fq_off / bc_off = exclusive scans of the lengths in sorted order
__global__ void __launch_bounds__(256) extract_write_kernel(TextLines T, long long n_rec, unsigned long long latch, const uint32_t* __restrict__ perm,
                                                           const long long* __restrict__ fq_len_sorted, const long long* __restrict__ fq_off,
                                                           const long long* __restrict__ bc_off, uint8_t* __restrict__ out_fq, uint8_t* __restrict__ out_bc)
{
    const long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (i >= n_rec || fq_len_sorted[i] == 0) return;
    const long long r = perm[i];
    const long long latch_rec = latch == ~0ull ? 0x7FFFFFFFFFFFFFFFll : (long long)(latch >> 2);
    const int type = r < latch_rec ? 0 : (int)(latch & 3ull);
    long long lh;
    const uint8_t* h = line_ptr(T, r * 8, &lh);
    Span name, bc;
    get_barcode(h, lh, type, &name, &bc);
    uint8_t* dst = out_fq + fq_off[i];
    copy_bytes_warp(dst, h + name.off, name.len, lane);
    long long at = name.len;
    const char tag[7] = "\tBX:Z:";
    if (lane < 6) dst[at + lane] = (uint8_t)tag[lane];
    at += 6;
    copy_bytes_warp(dst + at, h + bc.off, bc.len, lane);
    at += bc.len;
    if (lane == 0) { dst[at] = '-'; dst[at + 1] = '1'; dst[at + 2] = '\n'; }
    at += 3;
    const long long a = T.line_start[r * 8 + 1], e = T.line_start[r * 8 + 8] - 1;
    copy_bytes_warp(dst + at, T.text + a, e - a, lane);
    if (lane == 0) dst[at + (e - a)] = '\n';
    uint8_t* b = out_bc + bc_off[i];
    copy_bytes_warp(b, h + bc.off, bc.len, lane);
    if (lane == 0) b[bc.len] = '\n';
}

// run index of every record = inclusive count of change flags (record 0's flag compares with the empty barcode)
__global__ void run_index_kernel(const long long* __restrict__ change, const long long* __restrict__ change_excl, long long n_rec, long long* __restrict__ run_of_rec)
{
    const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (r < n_rec) run_of_rec[r] = change_excl[r] + change[r];
}

} // namespace pg
