// ingest_api.cuh - C-ABI over ingest.cuh (included at the end of api.cu: it uses the ctx internals).
#pragma once
#include "ingest.cuh"

// exclusive scan of n int64 values (in -> out, may alias); *total (device, may be null) receives the sum
static int scan64(pg_ctx* ctx, const long long* in, int64_t n, long long* out, long long* d_total)
{
    const int64_t n_tiles = std::max<int64_t>(1, (n + kScan64Tile - 1) / kScan64Tile);
    long long* tile = nullptr;
    CK(dmalloc(ctx, &tile, (size_t)n_tiles));
    scan64_reduce_kernel<<<(int)n_tiles, kScan64Threads, 0, ctx->stream>>>(in, n, tile);
    scan64_tiles_kernel<<<1, kScan64Threads, 0, ctx->stream>>>(tile, n_tiles, d_total);
    scan64_apply_kernel<<<(int)n_tiles, kScan64Threads, 0, ctx->stream>>>(in, n, tile, out);
    CK(cudaGetLastError());
    dfree(ctx, tile);
    return PG_OK;
}

// line index of a device text: line_start (n_lines + 1 entries, caller frees with dfree) and the line count
static int build_line_index(pg_ctx* ctx, const uint8_t* d_text, int64_t n, long long** line_start_out, int64_t* n_lines_out)
{
    *line_start_out = nullptr; *n_lines_out = 0;
    const int64_t n_tiles = std::max<int64_t>(1, (n + kTextTile - 1) / kTextTile);
    long long *tile = nullptr, *d_total = (long long*)ctx->d_scalar;
    CK(dmalloc(ctx, &tile, (size_t)n_tiles));
    nl_count_kernel<<<(int)n_tiles, 256, 0, ctx->stream>>>(d_text, n, tile);
    const int64_t n_t2 = std::max<int64_t>(1, (n_tiles + kScan64Tile - 1) / kScan64Tile);
    long long* t2 = nullptr;
    CK(dmalloc(ctx, &t2, (size_t)n_t2));
    scan64_reduce_kernel<<<(int)n_t2, kScan64Threads, 0, ctx->stream>>>(tile, n_tiles, t2);
    scan64_tiles_kernel<<<1, kScan64Threads, 0, ctx->stream>>>(t2, n_t2, d_total);
    scan64_apply_kernel<<<(int)n_t2, kScan64Threads, 0, ctx->stream>>>(tile, n_tiles, t2, tile);
    CK(cudaGetLastError());
    uint8_t last = '\n';
    CK(cudaMemcpyAsync(ctx->h_pin, d_total, sizeof(long long), cudaMemcpyDeviceToHost, ctx->stream));
    if (n > 0) CK(cudaMemcpyAsync(ctx->h_pin + 1, d_text + n - 1, 1, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    const int64_t newlines = ctx->h_pin[0];
    if (n > 0) last = *(const uint8_t*)(ctx->h_pin + 1);
    const int64_t n_lines = newlines + ((n > 0 && last != '\n') ? 1 : 0);
    long long* ls = nullptr;
    CK(dmalloc(ctx, &ls, (size_t)n_lines + 2));
    const long long zero = 0;
    CK(cudaMemcpyAsync(ls, &zero, sizeof(zero), cudaMemcpyHostToDevice, ctx->stream));
    nl_fill_kernel<<<(int)n_tiles, 256, 0, ctx->stream>>>(d_text, n, tile, ls);
    if (n > 0 && last != '\n') { // unterminated last line: a virtual newline at n
        const long long end = n + 1;
        CK(cudaMemcpyAsync(ls + n_lines, &end, sizeof(end), cudaMemcpyHostToDevice, ctx->stream));
    }
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(ctx->stream)); // (`zero` / `end` are stack variables)
    dfree(ctx, tile); dfree(ctx, t2);
    *line_start_out = ls; *n_lines_out = n_lines;
    return PG_OK;
}

struct pg_ingest {
    std::vector<std::string> labels;
    std::vector<uint8_t> keep;
};

extern "C" void pg_ingest_free(pg_ingest* g) { delete g; }
extern "C" int64_t pg_ingest_n_groups(const pg_ingest* g) { return g ? (int64_t)g->labels.size() : -1; }
extern "C" const uint8_t* pg_ingest_group_keep(const pg_ingest* g) { return g ? g->keep.data() : nullptr; }
extern "C" int64_t pg_ingest_group_labels(const pg_ingest* g, char* buf, int64_t cap, int64_t* offsets)
{
    if (!g) return -1;
    int64_t need = 0;
    for (auto& l : g->labels) need += (int64_t)l.size();
    if (!buf || !offsets || cap < need) return need;
    int64_t at = 0;
    size_t i = 0;
    for (auto& l : g->labels) { offsets[i++] = at; memcpy(buf + at, l.data(), l.size()); at += (int64_t)l.size(); }
    offsets[i] = at;
    return need;
}

// Device-side replacement of the host FASTQ loop for plain-text INTERLEAVED input (count_kmer.cpp:236-282 + getBarcode):
// text (host or device memory) -> a packed pg_batch + the labels of its clouds.  See include/pangaea_b200.h.
extern "C" int pg_ingest_text(pg_ctx* ctx, const void* text, int64_t n_bytes, int flags, const char* last_barcode, int64_t last_len,
                              int32_t* read_type_io, int64_t* consumed_out, pg_batch** batch_out, pg_ingest** info_out)
try {
    if (!ctx || !batch_out || !info_out || !consumed_out || !read_type_io || n_bytes < 0 || (n_bytes && !text) || last_len < 0 || (last_len && !last_barcode))
        return fail(ctx, PG_ERR_INVALID, "pg_ingest_text: bad argument");
    *batch_out = nullptr; *info_out = nullptr; *consumed_out = 0;
    CK(cudaSetDevice(ctx->p.device));
    const bool final_chunk = flags & PG_INGEST_FINAL, on_device = flags & PG_INGEST_DEVICE_TEXT;
    const bool want_q = ctx->p.min_qual_char != 0;
    uint8_t* d_text = nullptr;
    std::vector<void*> tmp; // device temporaries, released on every path
    auto done = [&](int rc) {
        for (void* p : tmp) dfree(ctx, p);
        return rc;
    };
#define CKI(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return done(fail(ctx, PG_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_))); } while (0)
    if (on_device) d_text = (uint8_t*)const_cast<void*>(text);
    else {
        // The chunk crosses PCIe on the copy stream into a staging buffer of its own, so the copy runs next to whatever the
        // compute stream still does for the previous batch (its count pass).  The buffer's last user was the ingest call
        // two chunks back, which returned after a stream sync: nothing to wait for.
        Workspace& w = ctx->ws_text[ctx->text_toggle];
        ctx->text_toggle ^= 1;
        if (w.bytes < (size_t)n_bytes + 64) {
            if (w.p) { cudaFree(w.p); w.p = nullptr; w.bytes = 0; }
            const size_t want = (size_t)n_bytes + (size_t)n_bytes / 8 + 64;
            CKI(cudaMalloc(&w.p, want));
            w.bytes = want;
        }
        d_text = (uint8_t*)w.p;
        cudaEvent_t ev = Timed::get(ctx);
        CKI(cudaMemcpyAsync(d_text, text, (size_t)n_bytes, cudaMemcpyHostToDevice, ctx->copy_stream));
        CKI(cudaEventRecord(ev, ctx->copy_stream));
        CKI(cudaStreamWaitEvent(ctx->stream, ev, 0));
        ctx->pool.push_back(ev);
    }
    // PG_INGEST_DEBUG=1: wall time of every phase on stderr (each mark waits for the stream)
    static const bool dbg = getenv("PG_INGEST_DEBUG") && getenv("PG_INGEST_DEBUG")[0] == '1';
    double t_last = 0;
    auto now_ms = []() { timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6; };
    if (dbg) { cudaStreamSynchronize(ctx->stream); t_last = now_ms(); }
    auto mark = [&](const char* what) {
        if (!dbg) return;
        cudaStreamSynchronize(ctx->stream);
        const double t = now_ms();
        fprintf(stderr, "[ingest] %-12s %7.2f ms\n", what, t - t_last);
        t_last = t;
    };
    mark("h2d");
    long long* line_start = nullptr;
    int64_t n_lines = 0;
    int rc = build_line_index(ctx, d_text, n_bytes, &line_start, &n_lines);
    if (rc) return done(rc);
    tmp.push_back(line_start);
    mark("line index");
    // whole records only, unless this is the end of the input (a trailing partial record still yields its reads)
    int64_t n_rec = final_chunk ? (n_lines + 7) / 8 : n_lines / 8;
    if (!final_chunk) { // complete lines only: an unterminated last line belongs to the next chunk
        CKI(cudaMemcpyAsync(ctx->h_pin, d_text + std::max<int64_t>(n_bytes - 1, 0), 1, cudaMemcpyDeviceToHost, ctx->stream));
        CKI(cudaStreamSynchronize(ctx->stream));
        if (n_bytes > 0 && *(const uint8_t*)ctx->h_pin != '\n') n_rec = (n_lines - 1) / 8;
    }
    TextLines T = { d_text, line_start, n_lines };
    pg_ingest* info = new pg_ingest();
    info->labels.emplace_back(last_barcode ? std::string(last_barcode, (size_t)last_len) : std::string());
    auto finish_empty = [&]() {
        info->keep.assign(1, info->labels[0].empty() ? 0 : 1);
        pg_reads none = {};
        pg_batch* b = nullptr;
        int rc2 = alloc_batch(ctx, &none, &b);
        if (!rc2) { cudaMemsetAsync(b->read_off, 0, sizeof(int64_t), ctx->stream); rc2 = pack_range(ctx, b, 0, b->n_words + 2); }
        if (rc2) { delete info; if (b) pg_batch_free(ctx, b); return done(rc2); }
        *batch_out = b; *info_out = info;
        return done(PG_OK);
    };
    if (n_rec == 0) {
        if (final_chunk) return finish_empty();
        delete info; // not even one whole record in this chunk: the caller widens it
        return done(PG_OK);
    }

    // ---- read type (latched once per FILE) and barcodes ----
    unsigned long long latch = ~0ull;
    if (*read_type_io) latch = (unsigned long long)*read_type_io; // record 0, known type
    else {
        unsigned long long* d_latch = (unsigned long long*)ctx->d_scalar;
        CKI(cudaMemsetAsync(d_latch, 0xFF, sizeof(unsigned long long), ctx->stream));
        latch_kernel<<<(int)((n_rec + 255) / 256), 256, 0, ctx->stream>>>(T, 8, n_rec, d_latch);
        CKI(cudaMemcpyAsync(ctx->h_pin, d_latch, sizeof(unsigned long long), cudaMemcpyDeviceToHost, ctx->stream));
        CKI(cudaStreamSynchronize(ctx->stream));
        latch = (unsigned long long)ctx->h_pin[0];
    }
    long long *bc_off = nullptr, *read_bytes = nullptr, *read_start = nullptr, *change = nullptr, *change_rank = nullptr;
    int32_t* bc_len = nullptr;
    uint8_t *flag2 = nullptr, *d_carry = nullptr;
    CKI(dmalloc(ctx, &bc_off, (size_t)n_rec)); tmp.push_back(bc_off);
    CKI(dmalloc(ctx, &bc_len, (size_t)n_rec)); tmp.push_back(bc_len);
    CKI(dmalloc(ctx, &read_bytes, (size_t)2 * n_rec)); tmp.push_back(read_bytes);
    CKI(dmalloc(ctx, &read_start, (size_t)2 * n_rec + 1)); tmp.push_back(read_start);
    CKI(dmalloc(ctx, &change, (size_t)n_rec)); tmp.push_back(change);
    CKI(dmalloc(ctx, &change_rank, (size_t)n_rec)); tmp.push_back(change_rank);
    CKI(dmalloc(ctx, &flag2, (size_t)2 * n_rec)); tmp.push_back(flag2);
    CKI(dmalloc(ctx, &d_carry, (size_t)last_len + 1)); tmp.push_back(d_carry);
    if (last_len) CKI(cudaMemcpyAsync(d_carry, last_barcode, (size_t)last_len, cudaMemcpyHostToDevice, ctx->stream));
    const int g_rec = (int)((n_rec + 255) / 256);
    barcode_kernel<<<g_rec, 256, 0, ctx->stream>>>(T, n_rec, latch, bc_off, bc_len);
    reads_kernel<<<g_rec, 256, 0, ctx->stream>>>(T, n_rec, bc_off, bc_len, d_carry, (int)last_len, read_bytes, flag2, change);
    CKI(cudaGetLastError());
    mark("headers");

    // ---- where does this batch end?  at the last cloud flush, or anywhere inside a cloud labelled "" ----
    int64_t n_use = n_rec;
    std::vector<long long> h_change;
    {
        h_change.resize((size_t)n_rec);
        CKI(cudaMemcpyAsync(h_change.data(), change, (size_t)n_rec * sizeof(long long), cudaMemcpyDeviceToHost, ctx->stream));
        CKI(cudaStreamSynchronize(ctx->stream));
        if (!final_chunk) {
            int64_t last_flush = -1;
            for (int64_t r = n_rec - 1; r >= 0; --r) if (h_change[(size_t)r]) { last_flush = r; break; }
            // barcode of the last record: empty -> the open cloud is labelled "" (dropped whole): cut at the end
            int32_t last_bc_len = 0;
            CKI(cudaMemcpy(&last_bc_len, bc_len + (n_rec - 1), sizeof(int32_t), cudaMemcpyDeviceToHost));
            if (last_bc_len != 0 && last_flush != n_rec - 1) n_use = last_flush + 1; // (0 when the chunk holds no flush: the caller widens it)
        }
    }
    mark("cut point");
    if (n_use == 0) { delete info; *consumed_out = 0; pg_batch* none = nullptr; *batch_out = none; *info_out = nullptr; return done(PG_OK); }

    // ---- the batch: read offsets, sequence bytes, flags ----
    const int64_t n_reads = 2 * n_use;
    CKI(scan64(ctx, read_bytes, n_reads, read_start, (long long*)ctx->d_scalar) == PG_OK ? cudaSuccess : cudaErrorUnknown);
    CKI(cudaMemcpyAsync(ctx->h_pin, ctx->d_scalar, sizeof(long long), cudaMemcpyDeviceToHost, ctx->stream));
    CKI(cudaMemcpyAsync(ctx->h_pin + 1, line_start + std::min<int64_t>(n_use * 8, n_lines), sizeof(long long), cudaMemcpyDeviceToHost, ctx->stream));
    CKI(cudaStreamSynchronize(ctx->stream));
    const int64_t seq_bytes = ctx->h_pin[0];
    *consumed_out = std::min<int64_t>(ctx->h_pin[1], n_bytes);
    pg_reads shape = {};
    shape.n_reads = n_reads; shape.n_bytes = seq_bytes;
    shape.qual = want_q ? (const uint8_t*)1 : nullptr; // (alloc_batch only tests it)
    pg_batch* b = nullptr;
    rc = alloc_batch(ctx, &shape, &b);
    if (rc) { delete info; return done(rc); }
    CKI(cudaMemcpyAsync(b->read_off, read_start, (size_t)n_reads * sizeof(int64_t), cudaMemcpyDeviceToDevice, ctx->stream));
    CKI(cudaMemcpyAsync(b->read_off + n_reads, ctx->d_scalar, sizeof(int64_t), cudaMemcpyDeviceToDevice, ctx->stream));
    CKI(cudaMemcpyAsync(b->read_flag, flag2, (size_t)n_reads, cudaMemcpyDeviceToDevice, ctx->stream));
    const int g_copy = (int)((n_reads * 32 + 255) / 256);
    copy_reads_kernel<<<g_copy, 256, 0, ctx->stream>>>(T, n_use, read_start, read_bytes, b->seq);
    if (want_q) copy_quals_kernel<<<g_copy, 256, 0, ctx->stream>>>(T, n_use, read_start, read_bytes, b->qual);
    CKI(cudaGetLastError());
    mark("copy reads");
    rc = pack_range(ctx, b, 0, b->n_words + 2);
    if (rc) { delete info; pg_batch_free(ctx, b); return done(rc); }
    mark("pack");

    // ---- labels of the clouds this batch opens ----
    int64_t n_lab = 0;
    for (int64_t r = 0; r < n_use; ++r) n_lab += h_change[(size_t)r] != 0;
    if (n_lab) {
        long long *lab_off = nullptr, *lab_len = nullptr, *lab_start = nullptr;
        uint8_t* blob = nullptr;
        CKI(scan64(ctx, change, n_use, change_rank, nullptr) == PG_OK ? cudaSuccess : cudaErrorUnknown);
        CKI(dmalloc(ctx, &lab_off, (size_t)n_lab)); tmp.push_back(lab_off);
        CKI(dmalloc(ctx, &lab_len, (size_t)n_lab)); tmp.push_back(lab_len);
        CKI(dmalloc(ctx, &lab_start, (size_t)n_lab)); tmp.push_back(lab_start);
        label_spans_kernel<<<(int)((n_use + 255) / 256), 256, 0, ctx->stream>>>(n_use, change, change_rank, bc_off, bc_len, lab_off, lab_len);
        CKI(scan64(ctx, lab_len, n_lab, lab_start, (long long*)ctx->d_scalar) == PG_OK ? cudaSuccess : cudaErrorUnknown);
        CKI(cudaMemcpyAsync(ctx->h_pin, ctx->d_scalar, sizeof(long long), cudaMemcpyDeviceToHost, ctx->stream));
        CKI(cudaStreamSynchronize(ctx->stream));
        const int64_t blob_bytes = ctx->h_pin[0];
        CKI(dmalloc(ctx, &blob, (size_t)blob_bytes + 1)); tmp.push_back(blob);
        label_copy_kernel<<<(int)((n_lab + 255) / 256), 256, 0, ctx->stream>>>(d_text, n_lab, lab_off, lab_len, lab_start, blob);
        CKI(cudaGetLastError());
        std::vector<char> h_blob((size_t)blob_bytes + 1);
        std::vector<long long> h_start((size_t)n_lab), h_len((size_t)n_lab);
        CKI(cudaMemcpyAsync(h_blob.data(), blob, (size_t)blob_bytes, cudaMemcpyDeviceToHost, ctx->stream));
        CKI(cudaMemcpyAsync(h_start.data(), lab_start, (size_t)n_lab * sizeof(long long), cudaMemcpyDeviceToHost, ctx->stream));
        CKI(cudaMemcpyAsync(h_len.data(), lab_len, (size_t)n_lab * sizeof(long long), cudaMemcpyDeviceToHost, ctx->stream));
        CKI(cudaStreamSynchronize(ctx->stream));
        info->labels.reserve((size_t)n_lab + 1);
        for (int64_t k = 0; k < n_lab; ++k) info->labels.emplace_back(h_blob.data() + h_start[(size_t)k], (size_t)h_len[(size_t)k]);
    }
    info->keep.resize(info->labels.size());
    for (size_t g = 0; g < info->labels.size(); ++g) info->keep[g] = info->labels[g].empty() ? 0 : 1;
    if (!*read_type_io && latch != ~0ull && (int64_t)(latch >> 2) < n_use) *read_type_io = (int32_t)(latch & 3ull);
    CKI(cudaStreamSynchronize(ctx->stream));
    mark("labels");
#undef CKI
    *batch_out = b; *info_out = info;
    return done(PG_OK);
} catch (...) { return caught("pg_ingest_text"); }

// ---------------------------------------------------------------------------
// barcode sort of an interleaved FASTQ held in host memory (run_pangaea:237-252)
// ---------------------------------------------------------------------------
extern "C" int pg_fastq_sort_by_barcode(pg_ctx* ctx, const char* in, int64_t n_in, char* out, int64_t out_cap, int64_t* n_out)
try {
    if (!ctx || n_in < 0 || (n_in && !in) || !n_out || (out_cap && !out)) return fail(ctx, PG_ERR_INVALID, "pg_fastq_sort_by_barcode: bad argument");
    *n_out = 0;
    if (n_in == 0) return PG_OK;
    CK(cudaSetDevice(ctx->p.device));
    std::vector<void*> tmp;
    auto done = [&](int rc) { for (void* p : tmp) dfree(ctx, p); return rc; };
#define CKS(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return done(fail(ctx, PG_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_))); } while (0)
    uint8_t* d_text = nullptr;
    CKS(dmalloc(ctx, &d_text, (size_t)n_in + 64)); tmp.push_back(d_text);
    CKS(cudaMemcpyAsync(d_text, in, (size_t)n_in, cudaMemcpyHostToDevice, ctx->stream));
    long long* line_start = nullptr;
    int64_t n_lines = 0;
    int rc = build_line_index(ctx, d_text, n_in, &line_start, &n_lines);
    if (rc) return done(rc);
    tmp.push_back(line_start);
    const int64_t n_rec = (n_lines + 7) / 8;
    if (n_rec >= (1ll << 32)) return done(fail(ctx, PG_ERR_INVALID, "pg_fastq_sort_by_barcode: more than 2^32 records in one call"));
    TextLines T = { d_text, line_start, n_lines };
    SortRec* recs = nullptr;
    long long *rec_bytes = nullptr, *out_start = nullptr, *sorted_bytes = nullptr;
    uint8_t *keysT = nullptr, *tie = nullptr;
    uint32_t *col_seen = nullptr, *perm_a = nullptr, *perm_b = nullptr, *d_flags = nullptr;
    CKS(dmalloc(ctx, &recs, (size_t)n_rec)); tmp.push_back(recs);
    CKS(dmalloc(ctx, &rec_bytes, (size_t)n_rec)); tmp.push_back(rec_bytes);
    CKS(dmalloc(ctx, &sorted_bytes, (size_t)n_rec)); tmp.push_back(sorted_bytes);
    CKS(dmalloc(ctx, &out_start, (size_t)n_rec + 1)); tmp.push_back(out_start);
    CKS(dmalloc(ctx, &keysT, (size_t)n_rec * kSortKeyBytes)); tmp.push_back(keysT);
    CKS(dmalloc(ctx, &tie, (size_t)n_rec)); tmp.push_back(tie);
    CKS(dmalloc(ctx, &col_seen, (size_t)kSortKeyBytes * 256)); tmp.push_back(col_seen);
    CKS(dmalloc(ctx, &perm_a, (size_t)n_rec)); tmp.push_back(perm_a);
    CKS(dmalloc(ctx, &perm_b, (size_t)n_rec)); tmp.push_back(perm_b);
    CKS(dmalloc(ctx, &d_flags, 2)); tmp.push_back(d_flags);
    CKS(cudaMemsetAsync(col_seen, 0, (size_t)kSortKeyBytes * 256 * sizeof(uint32_t), ctx->stream));
    CKS(cudaMemsetAsync(d_flags, 0, 2 * sizeof(uint32_t), ctx->stream));
    const int g_rec = (int)((n_rec + 255) / 256);
    sort_tag_kernel<<<g_rec, 256, 0, ctx->stream>>>(T, n_rec, recs, rec_bytes, d_flags);
    sort_keys_kernel<<<g_rec, 256, 0, ctx->stream>>>(T, n_rec, recs, keysT, col_seen);
    iota_kernel<<<g_rec, 256, 0, ctx->stream>>>(perm_a, n_rec);
    CKS(cudaGetLastError());
    std::vector<uint32_t> h_seen((size_t)kSortKeyBytes * 256);
    uint32_t h_flags[2] = { 0, 0 };
    CKS(cudaMemcpyAsync(h_seen.data(), col_seen, h_seen.size() * sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
    CKS(cudaMemcpyAsync(h_flags, d_flags, sizeof(h_flags), cudaMemcpyDeviceToHost, ctx->stream));
    CKS(cudaStreamSynchronize(ctx->stream));
    if (h_flags[0]) return done(fail(ctx, PG_ERR_INVALID, "pg_fastq_sort_by_barcode: the text is not a whole number of 8-line records that start with '@'"));
    // LSD radix over the key columns, last column first; a column in which every record holds the same byte is skipped
    const int64_t n_chunks = (n_rec + kRadixChunk - 1) / kRadixChunk;
    long long* hist = nullptr;
    CKS(dmalloc(ctx, &hist, (size_t)256 * n_chunks)); tmp.push_back(hist);
    const int g_radix = (int)((n_chunks + kRadixWarps - 1) / kRadixWarps);
    int passes = 0;
    for (int j = kSortKeyBytes - 1; j >= 0; --j) {
        int distinct = 0;
        for (int v = 0; v < 256; ++v) distinct += h_seen[(size_t)j * 256 + v] != 0;
        if (distinct <= 1) continue;
        const uint8_t* col = keysT + (size_t)j * n_rec;
        radix_hist_kernel<<<g_radix, kRadixWarps * 32, 0, ctx->stream>>>(col, perm_a, n_rec, n_chunks, hist);
        rc = scan64(ctx, hist, 256 * n_chunks, hist, nullptr);
        if (rc) return done(rc);
        radix_scatter_kernel<<<g_radix, kRadixWarps * 32, 0, ctx->stream>>>(col, perm_a, n_rec, n_chunks, hist, perm_b);
        std::swap(perm_a, perm_b);
        ++passes;
    }
    CKS(cudaGetLastError());
    // records whose first kSortKeyBytes bytes tie are ordered by the rest of the line on the host
    sort_ties_kernel<<<g_rec, 256, 0, ctx->stream>>>(keysT, perm_a, n_rec, tie, d_flags + 1);
    CKS(cudaMemcpyAsync(h_flags, d_flags, sizeof(h_flags), cudaMemcpyDeviceToHost, ctx->stream));
    CKS(cudaStreamSynchronize(ctx->stream));
    if (h_flags[1]) {
        std::vector<uint32_t> h_perm((size_t)n_rec);
        std::vector<uint8_t> h_tie((size_t)n_rec);
        std::vector<long long> h_ls((size_t)n_lines + 1);
        std::vector<SortRec> h_recs((size_t)n_rec);
        CKS(cudaMemcpy(h_perm.data(), perm_a, (size_t)n_rec * 4, cudaMemcpyDeviceToHost));
        CKS(cudaMemcpy(h_tie.data(), tie, (size_t)n_rec, cudaMemcpyDeviceToHost));
        CKS(cudaMemcpy(h_ls.data(), line_start, ((size_t)n_lines + 1) * 8, cudaMemcpyDeviceToHost));
        CKS(cudaMemcpy(h_recs.data(), recs, (size_t)n_rec * sizeof(SortRec), cudaMemcpyDeviceToHost));
        auto cmp_string = [&](uint32_t r) { // the comparison string of record r
            std::string s;
            const SortRec& sr = h_recs[r];
            if (sr.tag_len < 0) s = "~~~"; else s.assign(in + sr.tag_off, (size_t)sr.tag_len);
            s += '\t';
            const long long a = h_ls[(size_t)r * 8], e = h_ls[(size_t)std::min<int64_t>((int64_t)r * 8 + 8, n_lines)] - 1;
            for (long long p = a; p < e; ++p) s += in[p] == '\n' ? '\t' : in[p];
            return s;
        };
        for (int64_t i = 0; i < n_rec;) {
            int64_t j2 = i + 1;
            while (j2 < n_rec && h_tie[(size_t)j2]) ++j2;
            if (j2 - i > 1) {
                std::vector<std::pair<std::string, uint32_t>> grp;
                for (int64_t k = i; k < j2; ++k) grp.emplace_back(cmp_string(h_perm[(size_t)k]), h_perm[(size_t)k]);
                std::stable_sort(grp.begin(), grp.end(), [](const std::pair<std::string, uint32_t>& x, const std::pair<std::string, uint32_t>& y) { return x.first < y.first; });
                for (int64_t k = i; k < j2; ++k) h_perm[(size_t)k] = grp[(size_t)(k - i)].second;
            }
            i = j2;
        }
        CKS(cudaMemcpy(perm_a, h_perm.data(), (size_t)n_rec * 4, cudaMemcpyHostToDevice));
    }
    // output: every record's bytes in sorted order
    gather_len_kernel<<<g_rec, 256, 0, ctx->stream>>>(rec_bytes, perm_a, n_rec, sorted_bytes);
    rc = scan64(ctx, sorted_bytes, n_rec, out_start, (long long*)ctx->d_scalar);
    if (rc) return done(rc);
    CKS(cudaMemcpyAsync(ctx->h_pin, ctx->d_scalar, sizeof(long long), cudaMemcpyDeviceToHost, ctx->stream));
    CKS(cudaStreamSynchronize(ctx->stream));
    const int64_t total = ctx->h_pin[0];
    *n_out = total;
    if (total > out_cap) return done(fail(ctx, PG_ERR_INVALID, "pg_fastq_sort_by_barcode: output buffer too small (need n_out bytes)"));
    uint8_t* d_out = nullptr;
    CKS(dmalloc(ctx, &d_out, (size_t)total + 64)); tmp.push_back(d_out);
    sort_copy_kernel<<<(int)((n_rec * 32 + 255) / 256), 256, 0, ctx->stream>>>(T, n_rec, perm_a, out_start, d_out);
    CKS(cudaGetLastError());
    CKS(cudaMemcpyAsync(out, d_out, (size_t)total, cudaMemcpyDeviceToHost, ctx->stream));
    CKS(cudaStreamSynchronize(ctx->stream));
    (void)passes;
#undef CKS
    return done(PG_OK);
} catch (...) { return caught("pg_fastq_sort_by_barcode"); }

// ---------------------------------------------------------------------------
// step-2 input pipeline: weighted sampler + batch gather (sampler.cuh)
// ---------------------------------------------------------------------------
#include "sampler.cuh"

struct pg_sampler {
    int64_t n = 0;
    double *p = nullptr, *cdf = nullptr;
    unsigned long long* first = nullptr; // first-occurrence scratch of the draws without replacement (all ~0 between rounds)
    bool cdf_valid = false;
};

extern "C" void pg_sampler_free(pg_ctx* ctx, pg_sampler* s)
{
    if (!s) return;
    if (ctx) { cudaSetDevice(ctx->p.device); dfree(ctx, s->p); dfree(ctx, s->cdf); dfree(ctx, s->first); }
    delete s;
}

extern "C" int pg_sampler_create(pg_ctx* ctx, const double* weights, int64_t n, double total, pg_sampler** out)
try {
    if (!ctx || !out || n < 1 || !weights || !(total > 0.0)) return fail(ctx, PG_ERR_INVALID, "pg_sampler_create: bad argument");
    *out = nullptr;
    CK(cudaSetDevice(ctx->p.device));
    pg_sampler* s = new pg_sampler();
    s->n = n;
    double* d_w = nullptr;
    cudaError_t e = dmalloc(ctx, &s->p, (size_t)n);
    if (e == cudaSuccess) e = dmalloc(ctx, &s->cdf, (size_t)n);
    if (e == cudaSuccess) e = dmalloc(ctx, &s->first, (size_t)n);
    if (e == cudaSuccess) e = dmalloc(ctx, &d_w, (size_t)n);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_w, weights, (size_t)n * sizeof(double), cudaMemcpyDefault, ctx->stream); // host or device weights
    if (e == cudaSuccess) e = cudaMemsetAsync(s->first, 0xFF, (size_t)n * sizeof(unsigned long long), ctx->stream);
    if (e == cudaSuccess) {
        sampler_prob_kernel<<<(int)((n + 255) / 256), 256, 0, ctx->stream>>>(d_w, n, total, s->p);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    dfree(ctx, d_w);
    if (e != cudaSuccess) { pg_sampler_free(ctx, s); return fail(ctx, PG_ERR_CUDA, std::string("pg_sampler_create: ") + cudaGetErrorString(e)); }
    *out = s;
    return PG_OK;
} catch (...) { return caught("pg_sampler_create"); }

static int sampler_refresh_cdf(pg_ctx* ctx, pg_sampler* s)
{
    sampler_cumsum_kernel<<<1, 1, 0, ctx->stream>>>(s->p, s->n, s->cdf);
    if (s->n > 1) sampler_norm_kernel<<<(int)((s->n + 255) / 256), 256, 0, ctx->stream>>>(s->cdf, s->n);
    sampler_norm_last_kernel<<<1, 1, 0, ctx->stream>>>(s->cdf, s->n);
    CK(cudaGetLastError());
    s->cdf_valid = true;
    return PG_OK;
}

// with replacement: idx_out[j] = searchsorted(cdf, uniforms[j], "right").  uniforms: host memory; idx_out: device memory (m int64)
extern "C" int pg_sampler_draw(pg_ctx* ctx, pg_sampler* s, const double* uniforms, int64_t m, int64_t* d_idx_out)
try {
    if (!ctx || !s || m < 0 || (m && (!uniforms || !d_idx_out))) return fail(ctx, PG_ERR_INVALID, "pg_sampler_draw: bad argument");
    if (!m) return PG_OK;
    CK(cudaSetDevice(ctx->p.device));
    if (!s->cdf_valid) { int rc = sampler_refresh_cdf(ctx, s); if (rc) return rc; }
    double* d_u = nullptr;
    CK(dmalloc(ctx, &d_u, (size_t)m));
    CK(cudaMemcpyAsync(d_u, uniforms, (size_t)m * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    sampler_search_kernel<<<(int)((m + 255) / 256), 256, 0, ctx->stream>>>(s->cdf, s->n, d_u, m, d_idx_out);
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(ctx->stream));
    dfree(ctx, d_u);
    return PG_OK;
} catch (...) { return caught("pg_sampler_draw"); }

// one round of numpy's choice(replace=False): m uniforms -> the first occurrences among the draws, appended at d_idx_out[n_found...];
// *n_new = how many were appended.  The caller loops until it has `size` indices, drawing size - n_found fresh uniforms per round.
extern "C" int pg_sampler_draw_unique_round(pg_ctx* ctx, pg_sampler* s, const double* uniforms, int64_t m, int64_t n_found, int64_t* d_idx_out,
                                            int64_t* n_new)
try {
    if (!ctx || !s || m < 1 || !uniforms || !d_idx_out || !n_new || n_found < 0) return fail(ctx, PG_ERR_INVALID, "pg_sampler_draw_unique_round: bad argument");
    CK(cudaSetDevice(ctx->p.device));
    int rc = sampler_refresh_cdf(ctx, s); // the probabilities of the items found so far are zero by now
    if (rc) return rc;
    double* d_u = nullptr;
    int64_t* d_new = nullptr;
    long long *keep = nullptr, *rank = nullptr;
    CK(dmalloc(ctx, &d_u, (size_t)m)); CK(dmalloc(ctx, &d_new, (size_t)m)); CK(dmalloc(ctx, &keep, (size_t)m)); CK(dmalloc(ctx, &rank, (size_t)m));
    CK(cudaMemcpyAsync(d_u, uniforms, (size_t)m * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    const int g = (int)((m + 255) / 256);
    sampler_search_kernel<<<g, 256, 0, ctx->stream>>>(s->cdf, s->n, d_u, m, d_new);
    sampler_first_kernel<<<g, 256, 0, ctx->stream>>>(d_new, m, s->first);
    sampler_keep_kernel<<<g, 256, 0, ctx->stream>>>(d_new, m, s->first, keep);
    rc = scan64(ctx, keep, m, rank, (long long*)ctx->d_scalar);
    if (!rc) {
        sampler_commit_kernel<<<g, 256, 0, ctx->stream>>>(d_new, m, keep, rank, n_found, d_idx_out, s->p, s->first);
        s->cdf_valid = false;
        CK(cudaGetLastError());
        CK(cudaMemcpyAsync(ctx->h_pin, ctx->d_scalar, sizeof(long long), cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        *n_new = ctx->h_pin[0];
    }
    dfree(ctx, d_u); dfree(ctx, d_new); dfree(ctx, keep); dfree(ctx, rank);
    return rc;
} catch (...) { return caught("pg_sampler_draw_unique_round"); }

// batch gather on the device: rows idx[0..m) of the normalised matrices (pg_normalize first) into abd_out [m, v] / tnf_out [m, td]
extern "C" int pg_features_gather(pg_ctx* ctx, const pg_features* f, const int64_t* d_idx, int64_t m, float* d_abd_out, float* d_tnf_out)
try {
    if (!ctx || !f || m < 0 || (m && !d_idx)) return fail(ctx, PG_ERR_INVALID, "pg_features_gather: bad argument");
    if (!f->normalized) return fail(ctx, PG_ERR_STATE, "call pg_normalize first");
    if (!m) return PG_OK;
    CK(cudaSetDevice(ctx->p.device));
    uint32_t* bad = (uint32_t*)ctx->d_scalar;
    CK(cudaMemsetAsync(bad, 0, sizeof(uint32_t), ctx->stream));
    const int g = (int)((m * 32 + 255) / 256);
    if (d_abd_out) gather_rows_kernel<<<g, 256, 0, ctx->stream>>>(f->abd, f->vs, d_idx, m, f->rows, d_abd_out, bad);
    if (d_tnf_out) gather_rows_kernel<<<g, 256, 0, ctx->stream>>>(f->tnf, f->td, d_idx, m, f->rows, d_tnf_out, bad);
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(ctx->h_pin, bad, sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    if (*(const uint32_t*)ctx->h_pin) return fail(ctx, PG_ERR_INVALID, "pg_features_gather: row index out of range");
    return PG_OK;
} catch (...) { return caught("pg_features_gather"); }

// ---------------------------------------------------------------------------
// format converters and extract_reads (transform.cuh)
// ---------------------------------------------------------------------------
#include "transform.cuh"

namespace {
struct DevText { // a host text uploaded with its line index
    uint8_t* text = nullptr;
    long long* line_start = nullptr;
    int64_t n = 0, n_lines = 0;
    TextLines lines() const { return TextLines{ text, line_start, n_lines }; }
};
} // namespace

static int dev_text_upload(pg_ctx* ctx, const char* host, int64_t n, DevText* t)
{
    t->n = n;
    CK(dmalloc(ctx, &t->text, (size_t)n + 64));
    if (n) CK(cudaMemcpyAsync(t->text, host, (size_t)n, cudaMemcpyHostToDevice, ctx->stream));
    return build_line_index(ctx, t->text, n, &t->line_start, &t->n_lines);
}
static void dev_text_free(pg_ctx* ctx, DevText* t) { dfree(ctx, t->text); dfree(ctx, t->line_start); *t = DevText(); }

// scan lengths -> offsets, fetch the total
static int offsets_of(pg_ctx* ctx, const long long* len, int64_t n, long long* off, int64_t* total)
{
    int rc = scan64(ctx, len, n, off, (long long*)ctx->d_scalar);
    if (rc) return rc;
    CK(cudaMemcpyAsync(ctx->h_pin, ctx->d_scalar, sizeof(long long), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    *total = ctx->h_pin[0];
    return PG_OK;
}

extern "C" int pg_preprocess_stlfr(pg_ctx* ctx, const char* r1, int64_t n1, const char* r2, int64_t n2, int library, char* out1, int64_t cap1,
                                   int64_t* n_out1, char* out2, int64_t cap2, int64_t* n_out2)
try {
    if (!ctx || n1 < 0 || n2 < 0 || (n1 && !r1) || (n2 && !r2) || !n_out1 || !n_out2) return fail(ctx, PG_ERR_INVALID, "pg_preprocess_stlfr: bad argument");
    *n_out1 = *n_out2 = 0;
    CK(cudaSetDevice(ctx->p.device));
    DevText A, B;
    long long *len1 = nullptr, *len2 = nullptr, *off1 = nullptr, *off2 = nullptr;
    uint8_t *o1 = nullptr, *o2 = nullptr;
    uint32_t* bad = nullptr;
    auto done = [&](int rc) {
        dev_text_free(ctx, &A); dev_text_free(ctx, &B);
        dfree(ctx, len1); dfree(ctx, len2); dfree(ctx, off1); dfree(ctx, off2); dfree(ctx, o1); dfree(ctx, o2); dfree(ctx, bad);
        return rc;
    };
    int rc = dev_text_upload(ctx, r1, n1, &A);
    if (!rc) rc = dev_text_upload(ctx, r2, n2, &B);
    if (rc) return done(rc);
    const int64_t nl = A.n_lines;
    if (nl == 0) return done(PG_OK);
#define CKT(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return done(fail(ctx, PG_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_))); } while (0)
    CKT(dmalloc(ctx, &len1, (size_t)nl)); CKT(dmalloc(ctx, &len2, (size_t)nl)); CKT(dmalloc(ctx, &off1, (size_t)nl)); CKT(dmalloc(ctx, &off2, (size_t)nl));
    CKT(dmalloc(ctx, &bad, 1));
    CKT(cudaMemsetAsync(bad, 0, sizeof(uint32_t), ctx->stream));
    stlfr_size_kernel<<<(int)((nl + 255) / 256), 256, 0, ctx->stream>>>(A.lines(), B.lines(), library, len1, len2, bad);
    CKT(cudaGetLastError());
    int64_t t1 = 0, t2 = 0;
    rc = offsets_of(ctx, len1, nl, off1, &t1);
    if (!rc) rc = offsets_of(ctx, len2, nl, off2, &t2);
    if (rc) return done(rc);
    uint32_t h_bad = 0;
    CKT(cudaMemcpy(&h_bad, bad, sizeof(h_bad), cudaMemcpyDeviceToHost));
    if (h_bad) return done(fail(ctx, PG_ERR_INVALID, "pg_preprocess_stlfr: a header without '#' or a barcode that is not a_b_c (the reference tool aborts on it)"));
    *n_out1 = t1; *n_out2 = t2;
    if (t1 > cap1 || t2 > cap2 || !out1 || !out2) return done(fail(ctx, PG_ERR_INVALID, "pg_preprocess_stlfr: output buffers too small (sizes are in n_out1 / n_out2)"));
    CKT(dmalloc(ctx, &o1, (size_t)t1 + 64)); CKT(dmalloc(ctx, &o2, (size_t)t2 + 64));
    stlfr_write_kernel<<<(int)((nl * 64 + 255) / 256), 256, 0, ctx->stream>>>(A.lines(), B.lines(), library, off1, off2, o1, o2);
    CKT(cudaGetLastError());
    CKT(cudaMemcpyAsync(out1, o1, (size_t)t1, cudaMemcpyDeviceToHost, ctx->stream));
    CKT(cudaMemcpyAsync(out2, o2, (size_t)t2, cudaMemcpyDeviceToHost, ctx->stream));
    CKT(cudaStreamSynchronize(ctx->stream));
    return done(PG_OK);
} catch (...) { return caught("pg_preprocess_stlfr"); }

extern "C" int pg_preprocess_tellseq(pg_ctx* ctx, const char* r1, int64_t n1, const char* r2, int64_t n2, const char* idx, int64_t ni, char* out1,
                                     int64_t cap1, int64_t* n_out1, char* out2, int64_t cap2, int64_t* n_out2, char* out_wl, int64_t cap_wl,
                                     int64_t* n_out_wl)
try {
    if (!ctx || n1 < 0 || n2 < 0 || ni < 0 || (n1 && !r1) || (n2 && !r2) || (ni && !idx) || !n_out1 || !n_out2 || !n_out_wl)
        return fail(ctx, PG_ERR_INVALID, "pg_preprocess_tellseq: bad argument");
    *n_out1 = *n_out2 = *n_out_wl = 0;
    CK(cudaSetDevice(ctx->p.device));
    DevText A, B, I;
    long long *len1 = nullptr, *len2 = nullptr, *lenw = nullptr, *off1 = nullptr, *off2 = nullptr, *offw = nullptr;
    uint8_t *o1 = nullptr, *o2 = nullptr, *ow = nullptr;
    auto done = [&](int rc) {
        dev_text_free(ctx, &A); dev_text_free(ctx, &B); dev_text_free(ctx, &I);
        dfree(ctx, len1); dfree(ctx, len2); dfree(ctx, lenw); dfree(ctx, off1); dfree(ctx, off2); dfree(ctx, offw); dfree(ctx, o1); dfree(ctx, o2); dfree(ctx, ow);
        return rc;
    };
    int rc = dev_text_upload(ctx, r1, n1, &A);
    if (!rc) rc = dev_text_upload(ctx, r2, n2, &B);
    if (!rc) rc = dev_text_upload(ctx, idx, ni, &I);
    if (rc) return done(rc);
    const int64_t n_rec = A.n_lines / 4; // a record is written when its fourth line has been read
    if (n_rec == 0) return done(PG_OK);
    CKT(dmalloc(ctx, &len1, (size_t)n_rec)); CKT(dmalloc(ctx, &len2, (size_t)n_rec)); CKT(dmalloc(ctx, &lenw, (size_t)n_rec));
    CKT(dmalloc(ctx, &off1, (size_t)n_rec)); CKT(dmalloc(ctx, &off2, (size_t)n_rec)); CKT(dmalloc(ctx, &offw, (size_t)n_rec));
    tellseq_size_kernel<<<(int)((n_rec + 255) / 256), 256, 0, ctx->stream>>>(A.lines(), B.lines(), I.lines(), n_rec, len1, len2, lenw);
    CKT(cudaGetLastError());
    int64_t t1 = 0, t2 = 0, tw = 0;
    rc = offsets_of(ctx, len1, n_rec, off1, &t1);
    if (!rc) rc = offsets_of(ctx, len2, n_rec, off2, &t2);
    if (!rc) rc = offsets_of(ctx, lenw, n_rec, offw, &tw);
    if (rc) return done(rc);
    *n_out1 = t1; *n_out2 = t2; *n_out_wl = tw;
    if (t1 > cap1 || t2 > cap2 || tw > cap_wl || (t1 && !out1) || (t2 && !out2) || (tw && !out_wl))
        return done(fail(ctx, PG_ERR_INVALID, "pg_preprocess_tellseq: output buffers too small (sizes are in n_out*)"));
    CKT(dmalloc(ctx, &o1, (size_t)t1 + 64)); CKT(dmalloc(ctx, &o2, (size_t)t2 + 64)); CKT(dmalloc(ctx, &ow, (size_t)tw + 64));
    tellseq_write_kernel<<<(int)((n_rec * 64 + 255) / 256), 256, 0, ctx->stream>>>(A.lines(), B.lines(), I.lines(), n_rec, len1, off1, off2, offw, o1, o2, ow);
    CKT(cudaGetLastError());
    if (t1) CKT(cudaMemcpyAsync(out1, o1, (size_t)t1, cudaMemcpyDeviceToHost, ctx->stream));
    if (t2) CKT(cudaMemcpyAsync(out2, o2, (size_t)t2, cudaMemcpyDeviceToHost, ctx->stream));
    if (tw) CKT(cudaMemcpyAsync(out_wl, ow, (size_t)tw, cudaMemcpyDeviceToHost, ctx->stream));
    CKT(cudaStreamSynchronize(ctx->stream));
    return done(PG_OK);
} catch (...) { return caught("pg_preprocess_tellseq"); }

// ---- extract_reads -i -------------------------------------------------------
struct pg_extract {
    DevText T;
    int64_t n_rec = 0;
    unsigned long long latch = ~0ull;
    long long *bc_off = nullptr, *change = nullptr, *run_of_rec = nullptr;
    int32_t* bc_len = nullptr;
    std::vector<std::string> run_labels; // run 0 = records whose barcode equals "" from the start of the file
    // after pg_extract_route
    uint8_t *out_fq = nullptr, *out_bc = nullptr;
    int64_t fq_total = 0, bc_total = 0;
};

extern "C" void pg_extract_close(pg_ctx* ctx, pg_extract* x)
{
    if (!x) return;
    if (ctx) {
        cudaSetDevice(ctx->p.device);
        dev_text_free(ctx, &x->T);
        dfree(ctx, x->bc_off); dfree(ctx, x->change); dfree(ctx, x->run_of_rec); dfree(ctx, x->bc_len); dfree(ctx, x->out_fq); dfree(ctx, x->out_bc);
    }
    delete x;
}

extern "C" int pg_extract_open(pg_ctx* ctx, const char* text, int64_t n_bytes, pg_extract** out)
try {
    if (!ctx || !out || n_bytes < 0 || (n_bytes && !text)) return fail(ctx, PG_ERR_INVALID, "pg_extract_open: bad argument");
    *out = nullptr;
    CK(cudaSetDevice(ctx->p.device));
    pg_extract* x = new pg_extract();
    x->run_labels.emplace_back("");
    auto bail = [&](int rc) { pg_extract_close(ctx, x); return rc; };
    int rc = dev_text_upload(ctx, text, n_bytes, &x->T);
    if (rc) return bail(rc);
    const int64_t n_rec = (x->T.n_lines + 7) / 8;
    x->n_rec = n_rec;
    if (n_rec == 0) { *out = x; return PG_OK; }
#define CKX(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return bail(fail(ctx, PG_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_))); } while (0)
    TextLines T = x->T.lines();
    unsigned long long* d_latch = (unsigned long long*)ctx->d_scalar;
    CKX(cudaMemsetAsync(d_latch, 0xFF, sizeof(unsigned long long), ctx->stream));
    const int g = (int)((n_rec + 255) / 256);
    latch_kernel<<<g, 256, 0, ctx->stream>>>(T, 8, n_rec, d_latch);
    CKX(cudaMemcpyAsync(ctx->h_pin, d_latch, sizeof(unsigned long long), cudaMemcpyDeviceToHost, ctx->stream));
    CKX(cudaStreamSynchronize(ctx->stream));
    x->latch = (unsigned long long)ctx->h_pin[0];
    long long *read_bytes = nullptr, *change_excl = nullptr;
    uint8_t* flag2 = nullptr;
    CKX(dmalloc(ctx, &x->bc_off, (size_t)n_rec)); CKX(dmalloc(ctx, &x->bc_len, (size_t)n_rec)); CKX(dmalloc(ctx, &x->change, (size_t)n_rec));
    CKX(dmalloc(ctx, &x->run_of_rec, (size_t)n_rec));
    CKX(dmalloc(ctx, &read_bytes, (size_t)2 * n_rec)); CKX(dmalloc(ctx, &flag2, (size_t)2 * n_rec)); CKX(dmalloc(ctx, &change_excl, (size_t)n_rec));
    barcode_kernel<<<g, 256, 0, ctx->stream>>>(T, n_rec, x->latch, x->bc_off, x->bc_len);
    // barcode runs: a record starts a new run when its barcode differs from the record before (record 0: from "");
    // reads_kernel's change flag needs line 8r + 5 to exist, which holds for every record that extract_reads can output
    reads_kernel<<<g, 256, 0, ctx->stream>>>(T, n_rec, x->bc_off, x->bc_len, nullptr, 0, read_bytes, flag2, x->change);
    CKX(cudaGetLastError());
    rc = scan64(ctx, x->change, n_rec, change_excl, (long long*)ctx->d_scalar);
    if (rc) { dfree(ctx, read_bytes); dfree(ctx, flag2); dfree(ctx, change_excl); return bail(rc); }
    run_index_kernel<<<g, 256, 0, ctx->stream>>>(x->change, change_excl, n_rec, x->run_of_rec);
    CKX(cudaMemcpyAsync(ctx->h_pin, ctx->d_scalar, sizeof(long long), cudaMemcpyDeviceToHost, ctx->stream));
    CKX(cudaStreamSynchronize(ctx->stream));
    const int64_t n_lab = ctx->h_pin[0];
    if (n_lab) { // labels of the runs, as pg_ingest_text collects the labels of the clouds
        long long *lab_off = nullptr, *lab_len = nullptr, *lab_start = nullptr;
        uint8_t* blob = nullptr;
        CKX(dmalloc(ctx, &lab_off, (size_t)n_lab)); CKX(dmalloc(ctx, &lab_len, (size_t)n_lab)); CKX(dmalloc(ctx, &lab_start, (size_t)n_lab));
        label_spans_kernel<<<g, 256, 0, ctx->stream>>>(n_rec, x->change, change_excl, x->bc_off, x->bc_len, lab_off, lab_len);
        int64_t blob_bytes = 0;
        rc = offsets_of(ctx, lab_len, n_lab, lab_start, &blob_bytes);
        if (!rc) {
            CKX(dmalloc(ctx, &blob, (size_t)blob_bytes + 1));
            label_copy_kernel<<<(int)((n_lab + 255) / 256), 256, 0, ctx->stream>>>(x->T.text, n_lab, lab_off, lab_len, lab_start, blob);
            std::vector<char> h_blob((size_t)blob_bytes + 1);
            std::vector<long long> h_start((size_t)n_lab), h_len((size_t)n_lab);
            CKX(cudaMemcpyAsync(h_blob.data(), blob, (size_t)blob_bytes, cudaMemcpyDeviceToHost, ctx->stream));
            CKX(cudaMemcpyAsync(h_start.data(), lab_start, (size_t)n_lab * 8, cudaMemcpyDeviceToHost, ctx->stream));
            CKX(cudaMemcpyAsync(h_len.data(), lab_len, (size_t)n_lab * 8, cudaMemcpyDeviceToHost, ctx->stream));
            CKX(cudaStreamSynchronize(ctx->stream));
            for (int64_t k = 0; k < n_lab; ++k) x->run_labels.emplace_back(h_blob.data() + h_start[(size_t)k], (size_t)h_len[(size_t)k]);
        }
        dfree(ctx, lab_off); dfree(ctx, lab_len); dfree(ctx, lab_start); dfree(ctx, blob);
    }
    dfree(ctx, read_bytes); dfree(ctx, flag2); dfree(ctx, change_excl);
    if (rc) return bail(rc);
    *out = x;
    return PG_OK;
} catch (...) { return caught("pg_extract_open"); }

extern "C" int64_t pg_extract_n_runs(const pg_extract* x) { return x ? (int64_t)x->run_labels.size() : -1; }
extern "C" int64_t pg_extract_run_labels(const pg_extract* x, char* buf, int64_t cap, int64_t* offsets)
{
    if (!x) return -1;
    int64_t need = 0;
    for (auto& l : x->run_labels) need += (int64_t)l.size();
    if (!buf || !offsets || cap < need) return need;
    int64_t at = 0;
    size_t i = 0;
    for (auto& l : x->run_labels) { offsets[i++] = at; memcpy(buf + at, l.data(), l.size()); at += (int64_t)l.size(); }
    offsets[i] = at;
    return need;
}

// cluster_of_run[n_runs]: cluster index (0 .. n_clusters - 1) or -1.  fq_start / bc_start: n_clusters + 1 byte offsets of every
// cluster's slice in the two output blobs (pg_extract_copy)
extern "C" int pg_extract_route(pg_ctx* ctx, pg_extract* x, const int32_t* cluster_of_run, int32_t n_clusters, int64_t* fq_start, int64_t* bc_start)
try {
    if (!ctx || !x || !cluster_of_run || n_clusters < 0 || n_clusters > 65000 || !fq_start || !bc_start) return fail(ctx, PG_ERR_INVALID, "pg_extract_route: bad argument");
    CK(cudaSetDevice(ctx->p.device));
    for (int c = 0; c <= n_clusters; ++c) fq_start[c] = bc_start[c] = 0;
    const int64_t n_rec = x->n_rec;
    if (n_rec == 0) return PG_OK;
    const int64_t n_runs = (int64_t)x->run_labels.size();
    std::vector<void*> tmp;
    auto done = [&](int rc) { for (void* p : tmp) dfree(ctx, p); return rc; };
#define CKR(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return done(fail(ctx, PG_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_))); } while (0)
    int32_t* d_cl = nullptr;
    long long *fq_len = nullptr, *bc_out_len = nullptr, *fq_sorted = nullptr, *bc_sorted = nullptr, *fq_off = nullptr, *bc_off2 = nullptr, *hist = nullptr;
    uint8_t *key_lo = nullptr, *key_hi = nullptr;
    uint32_t *perm_a = nullptr, *perm_b = nullptr;
    CKR(dmalloc(ctx, &d_cl, (size_t)n_runs)); tmp.push_back(d_cl);
    CKR(dmalloc(ctx, &fq_len, (size_t)n_rec)); tmp.push_back(fq_len);
    CKR(dmalloc(ctx, &bc_out_len, (size_t)n_rec)); tmp.push_back(bc_out_len);
    CKR(dmalloc(ctx, &fq_sorted, (size_t)n_rec)); tmp.push_back(fq_sorted);
    CKR(dmalloc(ctx, &bc_sorted, (size_t)n_rec)); tmp.push_back(bc_sorted);
    CKR(dmalloc(ctx, &fq_off, (size_t)n_rec)); tmp.push_back(fq_off);
    CKR(dmalloc(ctx, &bc_off2, (size_t)n_rec)); tmp.push_back(bc_off2);
    CKR(dmalloc(ctx, &key_lo, (size_t)n_rec)); tmp.push_back(key_lo);
    CKR(dmalloc(ctx, &key_hi, (size_t)n_rec)); tmp.push_back(key_hi);
    CKR(dmalloc(ctx, &perm_a, (size_t)n_rec)); tmp.push_back(perm_a);
    CKR(dmalloc(ctx, &perm_b, (size_t)n_rec)); tmp.push_back(perm_b);
    CKR(cudaMemcpyAsync(d_cl, cluster_of_run, (size_t)n_runs * sizeof(int32_t), cudaMemcpyHostToDevice, ctx->stream));
    TextLines T = x->T.lines();
    const int g = (int)((n_rec + 255) / 256);
    extract_size_kernel<<<g, 256, 0, ctx->stream>>>(T, n_rec, x->latch, x->run_of_rec, d_cl, x->bc_len, fq_len, bc_out_len, key_lo, key_hi);
    iota_kernel<<<g, 256, 0, ctx->stream>>>(perm_a, n_rec);
    CKR(cudaGetLastError());
    // stable sort of the records by cluster (two 8-bit radix passes): file order is kept inside every cluster
    const int64_t n_chunks = (n_rec + kRadixChunk - 1) / kRadixChunk;
    CKR(dmalloc(ctx, &hist, (size_t)256 * n_chunks)); tmp.push_back(hist);
    const int g_radix = (int)((n_chunks + kRadixWarps - 1) / kRadixWarps);
    for (const uint8_t* col : { (const uint8_t*)key_lo, (const uint8_t*)key_hi }) {
        radix_hist_kernel<<<g_radix, kRadixWarps * 32, 0, ctx->stream>>>(col, perm_a, n_rec, n_chunks, hist);
        int rc = scan64(ctx, hist, 256 * n_chunks, hist, nullptr);
        if (rc) return done(rc);
        radix_scatter_kernel<<<g_radix, kRadixWarps * 32, 0, ctx->stream>>>(col, perm_a, n_rec, n_chunks, hist, perm_b);
        std::swap(perm_a, perm_b);
    }
    gather_len_kernel<<<g, 256, 0, ctx->stream>>>(fq_len, perm_a, n_rec, fq_sorted);
    gather_len_kernel<<<g, 256, 0, ctx->stream>>>(bc_out_len, perm_a, n_rec, bc_sorted);
    CKR(cudaGetLastError());
    int rc = offsets_of(ctx, fq_sorted, n_rec, fq_off, &x->fq_total);
    if (!rc) rc = offsets_of(ctx, bc_sorted, n_rec, bc_off2, &x->bc_total);
    if (rc) return done(rc);
    // per-cluster slices: bytes of cluster c = sum of its records' lengths (host side: the per-record arrays are small next to the text)
    {
        std::vector<long long> h_len((size_t)n_rec), h_bcl((size_t)n_rec);
        std::vector<uint8_t> h_lo((size_t)n_rec), h_hi((size_t)n_rec);
        CKR(cudaMemcpy(h_len.data(), fq_len, (size_t)n_rec * 8, cudaMemcpyDeviceToHost));
        CKR(cudaMemcpy(h_bcl.data(), bc_out_len, (size_t)n_rec * 8, cudaMemcpyDeviceToHost));
        CKR(cudaMemcpy(h_lo.data(), key_lo, (size_t)n_rec, cudaMemcpyDeviceToHost));
        CKR(cudaMemcpy(h_hi.data(), key_hi, (size_t)n_rec, cudaMemcpyDeviceToHost));
        std::vector<int64_t> fq_sz((size_t)n_clusters + 1, 0), bc_sz((size_t)n_clusters + 1, 0);
        for (int64_t r = 0; r < n_rec; ++r) {
            const int c = h_lo[(size_t)r] | (h_hi[(size_t)r] << 8);
            if (c < n_clusters) { fq_sz[(size_t)c] += h_len[(size_t)r]; bc_sz[(size_t)c] += h_bcl[(size_t)r]; }
        }
        int64_t a = 0, b = 0;
        for (int c = 0; c < n_clusters; ++c) { fq_start[c] = a; bc_start[c] = b; a += fq_sz[(size_t)c]; b += bc_sz[(size_t)c]; }
        fq_start[n_clusters] = a; bc_start[n_clusters] = b;
    }
    dfree(ctx, x->out_fq); dfree(ctx, x->out_bc);
    x->out_fq = x->out_bc = nullptr;
    CKR(dmalloc(ctx, &x->out_fq, (size_t)x->fq_total + 64));
    CKR(dmalloc(ctx, &x->out_bc, (size_t)x->bc_total + 64));
    extract_write_kernel<<<(int)((n_rec * 32 + 255) / 256), 256, 0, ctx->stream>>>(T, n_rec, x->latch, perm_a, fq_sorted, fq_off, bc_off2, x->out_fq, x->out_bc);
    CKR(cudaGetLastError());
    CKR(cudaStreamSynchronize(ctx->stream));
    return done(PG_OK);
} catch (...) { return caught("pg_extract_route"); }

extern "C" int pg_extract_copy(pg_ctx* ctx, const pg_extract* x, char* fq_out, char* bc_out)
try {
    if (!ctx || !x) return fail(ctx, PG_ERR_INVALID, "pg_extract_copy: bad argument");
    CK(cudaSetDevice(ctx->p.device));
    if (x->fq_total && fq_out) CK(cudaMemcpyAsync(fq_out, x->out_fq, (size_t)x->fq_total, cudaMemcpyDeviceToHost, ctx->stream));
    if (x->bc_total && bc_out) CK(cudaMemcpyAsync(bc_out, x->out_bc, (size_t)x->bc_total, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return PG_OK;
} catch (...) { return caught("pg_extract_copy"); }


// ---------------------------------------------------------------------------
// owner-partitioned table across ranks (exchange.cuh): the device side of the two all-to-alls
// ---------------------------------------------------------------------------
extern "C" int64_t pg_batch_n_words(const pg_batch* b) { return b ? b->n_words : -1; }

extern "C" int pg_batch_window_keys(pg_ctx* ctx, pg_batch* b, int64_t w0, int64_t w1, int feature_windows, uint64_t* d_keys)
try {
    if (!ctx || !b || w0 < 0 || w1 < w0 || w1 > b->n_words || (w1 > w0 && !d_keys)) return fail(ctx, PG_ERR_INVALID, "pg_batch_window_keys: bad argument");
    if (w1 == w0) return PG_OK;
    CK(cudaSetDevice(ctx->p.device));
    if (feature_windows && !b->grouped) return fail(ctx, PG_ERR_STATE, "pg_batch_window_keys: call pg_featurize2 first (it derives the feature mask)");
    const uint32_t* mask = feature_windows ? (b->maskR ? b->maskR : b->maskF) : b->maskC;
    window_keys_kernel<<<(int)((w1 - w0 + 255) / 256), 256, 0, ctx->stream>>>(b->codes, mask, w0, w1, ctx->p.k, (unsigned long long*)d_keys);
    CK(cudaGetLastError());
    return PG_OK;
} catch (...) { return caught("pg_batch_window_keys"); }

// d_sorted: the keys grouped by owner (owner 0 first); d_dest[i] = position of key i in d_sorted or -1; counts_out[world] on the host
extern "C" int pg_keys_partition(pg_ctx* ctx, const uint64_t* d_keys, int64_t n, int32_t world, uint64_t* d_sorted, int64_t* d_dest, int64_t* counts_out)
try {
    if (!ctx || n < 0 || world < 1 || world > 64 || !counts_out || (n && (!d_keys || !d_sorted || !d_dest))) return fail(ctx, PG_ERR_INVALID, "pg_keys_partition: bad argument");
    for (int r = 0; r < world; ++r) counts_out[r] = 0;
    if (!n) return PG_OK;
    CK(cudaSetDevice(ctx->p.device));
    unsigned long long* d_cnt = nullptr;
    CK(dmalloc(ctx, &d_cnt, 64));
    CK(cudaMemsetAsync(d_cnt, 0, 64 * sizeof(unsigned long long), ctx->stream));
    const int grid = grid_for(n, 256, ctx->sm_count * 8);
    owner_count_kernel<<<grid, 256, 0, ctx->stream>>>((const unsigned long long*)d_keys, n, (uint32_t)world, d_cnt);
    unsigned long long h_cnt[64];
    CK(cudaMemcpyAsync(h_cnt, d_cnt, (size_t)world * sizeof(unsigned long long), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    unsigned long long h_start[64], acc = 0;
    for (int r = 0; r < world; ++r) { counts_out[r] = (int64_t)h_cnt[r]; h_start[r] = acc; acc += h_cnt[r]; }
    CK(cudaMemcpyAsync(d_cnt, h_start, (size_t)world * sizeof(unsigned long long), cudaMemcpyHostToDevice, ctx->stream));
    owner_scatter_kernel<<<grid, 256, 0, ctx->stream>>>((const unsigned long long*)d_keys, n, (uint32_t)world, d_cnt, (unsigned long long*)d_sorted, (long long*)d_dest);
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(ctx->stream));
    dfree(ctx, d_cnt);
    return PG_OK;
} catch (...) { return caught("pg_keys_partition"); }

extern "C" int pg_table_add_keys(pg_ctx* ctx, const uint64_t* d_keys, int64_t n)
try {
    if (!ctx || n < 0 || (n && !d_keys)) return fail(ctx, PG_ERR_INVALID, "pg_table_add_keys: bad argument");
    CK(cudaSetDevice(ctx->p.device));
    { int rc_ = table_ready(ctx); if (rc_) return rc_; }
    if (ctx->zero_markers) return fail(ctx, PG_ERR_STATE, "pg_table_add_keys: the table holds zero-count markers from pg_table_set - call pg_table_clear first");
    int rc = ensure_table(ctx, std::max<int64_t>(n, 1));
    if (rc) return rc;
    ctx->counted = true;
    // launches of < 2^31 keys, each followed by the gated clamp: a dense counter cannot wrap (table.cuh: saturation)
    for (int64_t at = 0; at < n; at += (1ll << 30)) {
        const int64_t m = std::min<int64_t>(1ll << 30, n - at);
        table_add_keys_kernel<<<grid_for(m, 256, ctx->sm_count * 8), 256, 0, ctx->stream>>>(view(ctx), ctx->mode, (const unsigned long long*)d_keys + at, m);
        CK(saturate_if_flagged(ctx, true));
    }
    CK(cudaGetLastError());
    return check_overflow(ctx);
} catch (...) { return caught("pg_table_add_keys"); }

extern "C" int pg_table_lookup_keys(pg_ctx* ctx, const uint64_t* d_keys, int64_t n, uint32_t* d_counts)
try {
    if (!ctx || n < 0 || (n && (!d_keys || !d_counts))) return fail(ctx, PG_ERR_INVALID, "pg_table_lookup_keys: bad argument");
    if (!n) return PG_OK;
    CK(cudaSetDevice(ctx->p.device));
    { int rc_ = table_ready(ctx); if (rc_) return rc_; }
    if (!ctx->have_table()) { CK(cudaMemsetAsync(d_counts, 0, (size_t)n * sizeof(uint32_t), ctx->stream)); return PG_OK; }
    table_lookup_keys_kernel<<<grid_for(n, 256, ctx->sm_count * 8), 256, 0, ctx->stream>>>(view(ctx), ctx->mode, (const unsigned long long*)d_keys, n, d_counts);
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(ctx->stream));
    return PG_OK;
} catch (...) { return caught("pg_table_lookup_keys"); }

// abundance tallies of the windows that start in words [w0, w1) from counts that came back from their owners:
// d_dest (32 per word, from pg_keys_partition) indexes d_counts_sorted (the counts in the order the keys were sent)
extern "C" int pg_features_add_counts(pg_ctx* ctx, pg_features* f, pg_batch* b, int64_t w0, int64_t w1, const int64_t* d_dest, const uint32_t* d_counts_sorted)
try {
    if (!ctx || !f || !b || w0 < 0 || w1 < w0 || w1 > b->n_words) return fail(ctx, PG_ERR_INVALID, "pg_features_add_counts: bad argument");
    if (!f->row_of_group || !b->wg || !b->gstart) return fail(ctx, PG_ERR_STATE, "pg_features_add_counts: the feature set must come from pg_featurize2(PG_FEAT_NO_ABUNDANCE) of this batch");
    if (w1 == w0 || !f->rows) return PG_OK;
    if (!d_dest || !d_counts_sorted) return fail(ctx, PG_ERR_INVALID, "pg_features_add_counts: null arrays");
    CK(cudaSetDevice(ctx->p.device));
    FeatParams P = {};
    P.gstart = b->gstart; P.n_groups = f->n_groups; P.row_of_group = f->row_of_group; P.wg = b->wg;
    P.vs = f->vs; P.td = f->td; P.ws = (uint32_t)ctx->p.window_size;
    const uint64_t clamp64 = (uint64_t)P.ws * (uint64_t)P.vs;
    P.clamp = clamp64 > 0xFFFFFFFFull ? 0xFFFFFFFFu : (uint32_t)clamp64;
    P.use_magic = (P.ws > 1 && (uint64_t)P.ws * clamp64 < (1ull << 32)) ? 1 : 0;
    P.magic = P.use_magic ? (uint32_t)(((1ull << 32) + P.ws - 1) / P.ws) : 0u;
    P.abd = f->abd_raw; P.tnf = f->tnf_raw;
    Timed t(ctx, T_FEAT, 1);
    abd_from_counts_kernel<<<grid_for((w1 - w0) * 32, 256, ctx->sm_count * 8), 256, 0, ctx->stream>>>(P, w0, w1, (const long long*)d_dest, d_counts_sorted);
    CK(cudaGetLastError());
    return PG_OK;
} catch (...) { return caught("pg_features_add_counts"); }
