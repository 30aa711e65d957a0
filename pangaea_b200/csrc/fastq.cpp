// fastq.cpp - host FASTQ reader: barcode-sorted reads -> pg_reads batch + cloud labels.
//
// Replaces the two streaming loops of the reference tools
//   interleaved: src/cpptools/count_kmer.cpp:236-282 == count_tnf.cpp:234-291
//   paired:      src/cpptools/count_kmer.cpp:181-233 == count_tnf.cpp:170-231
// and getBarcode (count_kmer.cpp:25-53).  It only DECIDES (which bytes are sequence,
// where a cloud is flushed, what its label is); all arithmetic on bases happens on the
// GPU.  Differences from the reference reader are deliberate and invisible in the
// output: one pass instead of three (jellyfish, count_kmer, count_tnf each re-read the
// file), large reads instead of a 303-byte gz buffer (lib/gzstream/gzstream.h:47).
//
// Line semantics are std::getline's: '\n' terminates and is dropped, '\r' is kept (it
// becomes an invalid base and counts towards the cloud length), a last line without
// '\n' is delivered, lines are numbered including blank ones.
#include <algorithm>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <exception>
#include <new>
#include <future>
#include <string>
#if defined(__SSE2__)
#include <emmintrin.h>
#endif
#include <thread>
#include <vector>

#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
#include <zlib.h>

#include "../../include/pangaea_b200.h"

namespace {

struct LineReader {
    gzFile f = nullptr;
    std::vector<char> bufs[2];   // one is parsed while a helper thread inflates / reads the next 16 MB into the other
    int cur = 0;
    size_t pos = 0, end = 0;
    bool eof = false;
    bool failed = false; // a read error or a truncated gzip stream: the caller must not take what came so far for the file
    std::string spill; // a line that straddles two chunks
    struct Fill { int n; int errnum; };
    std::future<Fill> pending;   // the read-ahead of bufs[cur ^ 1] (of bufs[0] before the first fill)
    int pending_buf = 0;

    Fill read_into(int b)
    {
        Fill r;
        r.n = gzread(f, bufs[b].data(), (unsigned)bufs[b].size());
        r.errnum = Z_OK;
        if (r.n <= 0) gzerror(f, &r.errnum); // Z_BUF_ERROR: the compressed stream ends early
        return r;
    }
    void read_ahead(int b)
    {
        pending_buf = b;
        pending = std::async(std::launch::async, [this, b]() { return read_into(b); });
    }
    bool open(const char* path)
    {
        f = gzopen(path, "rb"); // transparent for plain text, like the reference's igzstream
        if (!f) return false;
        gzbuffer(f, 1 << 22);
        bufs[0].resize(1 << 24);
        bufs[1].resize(1 << 24);
        read_ahead(0); // (with two files their first chunks inflate side by side)
        return true;
    }
    ~LineReader()
    {
        if (pending.valid()) pending.wait(); // the helper still holds the gzFile
        if (f) gzclose(f);
    }
    bool fill()
    {
        if (eof) return false;
        const Fill r = pending.get();
        cur = pending_buf;
        pos = 0;
        end = r.n > 0 ? (size_t)r.n : 0;
        if (r.n <= 0) {
            eof = true;
            if (r.n < 0 || (r.errnum != Z_OK && r.errnum != Z_STREAM_END)) failed = true;
        } else {
            read_ahead(cur ^ 1);
        }
        return end > 0;
    }
    // returns false at end of file; line is valid until the next call
    bool next(const char** line, size_t* len)
    {
        if (!f) return false;
        spill.clear();
        bool have = false;
        for (;;) {
            if (pos == end && !fill()) break;
            const char* s = bufs[cur].data() + pos;
            const char* nl = (const char*)memchr(s, '\n', end - pos);
            if (nl) {
                size_t n = (size_t)(nl - s);
                pos += n + 1;
                if (!have) { *line = s; *len = n; return true; }
                spill.append(s, n);
                *line = spill.data(); *len = spill.size();
                return true;
            }
            spill.append(s, end - pos);
            have = true;
            pos = end;
        }
        if (!have) return false;
        *line = spill.data(); *len = spill.size();
        return true;
    }
};

const size_t npos = (size_t)-1;

size_t find_char(const char* s, size_t len, char c, size_t from)
{
    if (from >= len) return npos;
    const char* p = (const char*)memchr(s + from, c, len - from);
    return p ? (size_t)(p - s) : npos;
}
size_t find_bxz(const char* s, size_t len)
{
    for (size_t from = 0;;) {
        size_t p = find_char(s, len, 'B', from);
        if (p == npos || p + 4 > len) return npos;
        if (s[p + 1] == 'X' && s[p + 2] == ':' && s[p + 3] == 'Z') return p;
        from = p + 1;
    }
}
// std::string::substr(pos, n) with n clamped; pos > size would throw in the reference
void substr(std::string* out, const char* s, size_t len, size_t pos, size_t n)
{
    if (pos > len) { out->clear(); return; }
    if (n > len - pos) n = len - pos;
    out->assign(s + pos, n);
}

struct HeaderParser {
    int read_type = 0; // 0 undecided, 1 "10x", 2 "stLFR" - latched once (count_kmer.cpp:24,28-33)
    // name may be null (interleaved input: only the paired reader compares the names of R1 and R2)
    void parse(const char* line, size_t len, std::string* name, std::string* bc)
    {
        if (read_type == 0) {
            if (find_bxz(line, len) != npos) read_type = 1;
            else if (find_char(line, len, '#', 0) != npos) read_type = 2;
        }
        if (read_type == 2) { // count_kmer.cpp:36-43
            size_t p1 = find_char(line, len, '#', 0);
            size_t p2 = find_char(line, len, '/', p1 + 1); // npos + 1 == 0, as in the reference
            if (name) substr(name, line, len, 0, p1);
            substr(bc, line, len, p1 + 1, p2 - p1 - 1);
            if (*bc == "0_0_0") bc->clear();
        } else { // count_kmer.cpp:44-51
            if (name) {
                size_t e = npos;
                for (size_t i = 0; i < len; ++i)
                    if (line[i] == ' ' || line[i] == '\r' || line[i] == '\t' || line[i] == '\n') { e = i; break; }
                substr(name, line, len, 0, e);
            }
            bc->clear();
            size_t p1 = find_bxz(line, len);
            if (p1 != npos) {
                size_t p2 = find_char(line, len, '-', p1 + 5);
                substr(bc, line, len, p1 + 5, p2 - p1 - 5);
            }
        }
    }
};

} // namespace

// ---------------------------------------------------------------------------
// host memory of the batches.  Pinned (cudaHostAlloc) when the caller asks for it: the H2D copy of a batch then runs at
// PCIe speed and overlaps the kernels; pageable otherwise (CPU-only use: tests, tools).  Page-locking costs ~0.1 s per
// GB, so freed blocks are kept in a process-wide pool and handed out again (a stream of chunks reuses two or three).
// ---------------------------------------------------------------------------
#include <cuda_runtime.h>
#include <map>
#include <memory>
#include <mutex>

namespace {

struct HostPool {
    std::mutex mu;
    std::multimap<size_t, void*> idle; // pinned blocks only
    size_t idle_bytes = 0;
    static constexpr size_t kKeep = (size_t)24 << 30;
    void* get(size_t bytes, bool pinned, size_t* cap, bool* is_pinned)
    {
        if (bytes == 0) bytes = 64;
        if (pinned) {
            {
                std::lock_guard<std::mutex> g(mu);
                auto it = idle.lower_bound(bytes);
                if (it != idle.end() && it->first <= 2 * bytes + (1u << 20)) {
                    void* q = it->second; *cap = it->first; *is_pinned = true;
                    idle_bytes -= it->first; idle.erase(it);
                    return q;
                }
            }
            void* q = nullptr;
            if (cudaHostAlloc(&q, bytes, cudaHostAllocPortable) == cudaSuccess) { *cap = bytes; *is_pinned = true; return q; }
            cudaGetLastError(); // no device / no room to lock: pageable memory still works, only slower
        }
        void* q = malloc(bytes);
        if (!q) throw std::bad_alloc();
        *cap = bytes; *is_pinned = false;
        return q;
    }
    void put(void* q, size_t cap, bool is_pinned)
    {
        if (!q) return;
        if (!is_pinned) { free(q); return; }
        std::lock_guard<std::mutex> g(mu);
        if (idle_bytes + cap > kKeep) { cudaFreeHost(q); return; }
        idle.insert({ cap, q }); idle_bytes += cap;
    }
};
HostPool& host_pool() { static HostPool* p = new HostPool(); return *p; } // leaked on purpose: no static-destruction order games with CUDA

} // namespace

// growable array that does not zero-fill (a resize of 10 GB would otherwise cost seconds); vector-like otherwise
template <class T>
struct Buf {
    T* p = nullptr;
    size_t n = 0, cap = 0;
    bool want_pinned = false, is_pinned = false;
    Buf() = default;
    Buf(const Buf&) = delete;
    Buf& operator=(const Buf&) = delete;
    ~Buf() { host_pool().put(p, cap * sizeof(T), is_pinned); }
    T* data() { return p; }
    const T* data() const { return p; }
    size_t size() const { return n; }
    T& operator[](size_t i) { return p[i]; }
    const T& operator[](size_t i) const { return p[i]; }
    void reserve(size_t c)
    {
        if (c <= cap) return;
        const size_t nc = std::max(c, cap + cap / 2 + 64);
        size_t got = 0; bool pin = false;
        T* q = (T*)host_pool().get(nc * sizeof(T), want_pinned, &got, &pin);
        if (n) memcpy(q, p, n * sizeof(T));
        host_pool().put(p, cap * sizeof(T), is_pinned);
        p = q; cap = got / sizeof(T); is_pinned = pin;
    }
    void resize_uninit(size_t m) { reserve(m); n = m; }
    void resize(size_t m, T fill)
    {
        reserve(m);
        for (size_t i = n; i < m; ++i) p[i] = fill;
        n = m;
    }
    void push_back(T v) { reserve(n + 1); p[n++] = v; }
    void append(const T* s, size_t m) { reserve(n + m); memcpy(p + n, s, m * sizeof(T)); n += m; }
};

struct pg_fastq {
    Buf<uint8_t> seq, qual, flag;
    std::vector<uint8_t> keep;
    Buf<int64_t> off;
    std::vector<std::string> labels;
    bool want_qual = false;

    explicit pg_fastq(bool pinned = false, const std::string& label0 = std::string())
    {
        seq.want_pinned = qual.want_pinned = flag.want_pinned = off.want_pinned = pinned;
        off.push_back(0);
        labels.push_back(label0); // label of the cloud that is open when this batch starts ("" at the start of a file)
    }
    int64_t add_read(const char* s, size_t n, uint8_t fl)
    {
        seq.append((const uint8_t*)s, n);
        seq.push_back('\n');
        if (want_qual) qual.resize(seq.size(), (uint8_t)0xFF);
        off.push_back((int64_t)seq.size());
        flag.push_back(fl);
        return (int64_t)flag.size() - 1;
    }
    void set_qual(int64_t r, const char* q, size_t n)
    {
        if (!want_qual || r < 0) return;
        size_t len = (size_t)(off[r + 1] - off[r] - 1);
        memcpy(qual.data() + off[r], q, n < len ? n : len);
    }
    // the cloud is flushed after read r: count_kmer.cpp:251-270 / :200-219
    void change_after(int64_t r, const std::string& new_label)
    {
        flag[r] |= PG_READ_CHANGE;
        labels.push_back(new_label);
    }
    void finish()
    {
        keep.resize(labels.size());
        for (size_t g = 0; g < labels.size(); ++g) keep[g] = labels[g].empty() ? 0 : 1;
    }
};

static thread_local std::string g_fq_err;

namespace {

struct MappedFile {
    const char* p = nullptr;
    size_t n = 0;
    int fd = -1;
    bool open(const char* path)
    {
        fd = ::open(path, O_RDONLY);
        if (fd < 0) return false;
        struct stat st;
        if (fstat(fd, &st) != 0 || !S_ISREG(st.st_mode)) return false;
        n = (size_t)st.st_size;
        if (n == 0) return true;
        void* m = mmap(nullptr, n, PROT_READ, MAP_PRIVATE, fd, 0);
        if (m == MAP_FAILED) return false;
        madvise(m, n, MADV_SEQUENTIAL);
        p = (const char*)m;
        return true;
    }
    ~MappedFile()
    {
        if (p) munmap((void*)p, n);
        if (fd >= 0) close(fd);
    }
};

int parallel_threads()
{
    const char* e = getenv("PG_FASTQ_THREADS");
    if (e && atoi(e) > 0) return std::min(atoi(e), 256);
    const unsigned hc = std::thread::hardware_concurrency();
    return (int)std::max(1u, std::min(hc ? hc : 1u, 32u));
}

// run fn(t) for t in [0, T) on T threads; an exception in any of them is rethrown here (never across a thread boundary)
template <class Fn>
void run_threads(int T, Fn&& fn)
{
    std::vector<std::thread> th;
    std::vector<std::exception_ptr> err((size_t)T);
    for (int t = 0; t < T; ++t)
        th.emplace_back([&, t]() { try { fn(t); } catch (...) { err[(size_t)t] = std::current_exception(); } });
    for (auto& x : th) x.join();
    for (auto& e : err) if (e) std::rethrow_exception(e);
}

// 16 bytes per compare: a match subtracts -1 from its byte lane; the lanes are summed every 255 blocks (psadbw)
uint64_t count_newlines(const char* p, size_t lo, size_t hi)
{
    uint64_t c = 0;
    size_t i = lo;
#if defined(__SSE2__)
    const __m128i nl = _mm_set1_epi8('\n'), zero = _mm_setzero_si128();
    while (i + 16 <= hi) {
        __m128i acc = zero;
        const size_t blocks = std::min<size_t>((hi - i) / 16, 255);
        for (size_t b = 0; b < blocks; ++b, i += 16)
            acc = _mm_sub_epi8(acc, _mm_cmpeq_epi8(_mm_loadu_si128(reinterpret_cast<const __m128i*>(p + i)), nl));
        const __m128i sums = _mm_sad_epu8(acc, zero);
        c += (uint64_t)_mm_cvtsi128_si64(sums) + (uint64_t)_mm_cvtsi128_si64(_mm_srli_si128(sums, 8));
    }
#endif
    for (; i < hi; ++i) c += p[i] == '\n';
    return c;
}

uint64_t count_newlines_parallel(const char* p, size_t lo, size_t hi, int T)
{
    if (hi <= lo) return 0;
    if (T < 2 || hi - lo < ((size_t)8 << 20)) return count_newlines(p, lo, hi);
    std::vector<uint64_t> c((size_t)T, 0);
    const size_t span = hi - lo;
    run_threads(T, [&](int t) { c[(size_t)t] = count_newlines(p, lo + span / T * t, t + 1 == T ? hi : lo + span / T * (t + 1)); });
    uint64_t s = 0;
    for (auto v : c) s += v;
    return s;
}

// calls fn(line_number, ptr, len) for every line of [begin, end); the last line may lack its newline.
// FASTQ lines are short (a memchr call per line costs more than the scan itself): 64 bytes per step, the newline positions of
// the block as a bit mask, one callback per set bit.
template <class Fn>
void for_lines(const char* p, size_t begin, size_t end, uint64_t line0, Fn&& fn)
{
    uint64_t n = line0;
    size_t start = begin; // first byte of the line being scanned
    size_t pos = begin;   // no newline in [start, pos)
#if defined(__SSE2__)
    const __m128i nl = _mm_set1_epi8('\n');
    auto mask16 = [&](size_t at) { return (uint64_t)(uint32_t)_mm_movemask_epi8(_mm_cmpeq_epi8(_mm_loadu_si128(reinterpret_cast<const __m128i*>(p + at)), nl)); };
    while (pos + 64 <= end) {
        uint64_t m = mask16(pos) | (mask16(pos + 16) << 16) | (mask16(pos + 32) << 32) | (mask16(pos + 48) << 48);
        while (m) {
            const size_t e = pos + (size_t)__builtin_ctzll(m);
            fn(n++, p + start, e - start);
            start = e + 1;
            m &= m - 1;
        }
        pos += 64;
    }
#endif
    while (start < end) {
        const size_t from = pos > start ? pos : start;
        const char* q = from < end ? (const char*)memchr(p + from, '\n', end - from) : nullptr;
        const size_t e = q ? (size_t)(q - p) : end;
        fn(n++, p + start, e - start);
        start = pos = e + 1;
    }
}

size_t next_line(const char* p, size_t n, size_t pos)
{
    const char* nl = (const char*)memchr(p + pos, '\n', n - pos);
    return nl ? (size_t)(nl - p) + 1 : n;
}

struct Piece {                 // what one thread contributes
    size_t begin = 0, end = 0; // byte range of whole records: lines [line0, ...) with line0 = 0 mod 8
    uint64_t line0 = 0;
    size_t seq_bytes = 0, n_reads = 0;          // pass A
    std::vector<std::pair<int64_t, std::string>> changes; // (read index inside the piece, new label), in order
    bool has_pair = false;
    std::string first_bc, last_bc;
    int64_t first_r2 = -1;     // read index (inside the piece) of the first pair's R2
};

} // namespace

// ---------------------------------------------------------------------------
// A stream of batches over one input (interleaved plain text: mapped and parsed by all cores; gzip / paired files: the
// sequential reader).  The reference streams a file of any size one cloud at a time (count_kmer.cpp:236-282); here the
// unit is a batch of about `target` sequence bytes that always ends where the reference would flush a cloud - right
// after a read that carries PG_READ_CHANGE - or inside a cloud whose label is "" (such a cloud is dropped whole,
// count_kmer.cpp:62, so both halves are).  The two pieces of sequential state, read_type and last_barcode, are carried
// from batch to batch; labels[0] of a batch is the label of the cloud that is open when it starts.
//
// Byte ranges (one rank of a multi-GPU run reads 1/N of the file): a range [lo, hi) is snapped to the same kind of
// point - the first flush at or after the first record boundary >= lo - by looking at the two pairs before the
// boundary, so rank r ends exactly where rank r + 1 starts and nothing is exchanged but the number of newlines before
// `lo` (record boundaries are lines 0 mod 8 counted from the start of the file).
// ---------------------------------------------------------------------------
struct pg_fastq_stream {
    bool want_qual = false, pinned = false;
    bool done = false;
    // ---- mapped mode ----
    bool mapped = false;
    MappedFile mf;
    size_t pos = 0, end = 0;       // cursor / end of the range, both record boundaries (or EOF)
    uint64_t line = 0;             // line number (from the start of the file) at pos
    HeaderParser latch;            // read_type of the file
    uint64_t latch_line = ~0ull;   // line of the first decisive header
    std::string last;              // last_barcode at pos
    // ---- sequential mode ----
    LineReader ra, rb;
    bool paired = false;
    HeaderParser hp;
    uint64_t n1 = 0;               // lines read from file 1
    bool more1 = true, more2 = true;
    bool open_nonempty = false;    // the open cloud holds a read
    bool pair_ok = false;          // paired mode: R1 / R2 of the current record agree (count_kmer.cpp:195)
    std::string pair_bc;
};

namespace {

// barcode of the record (8 lines, interleaved pair) that starts at `rec`; line = its absolute line number
std::string record_barcode(const pg_fastq_stream* s, size_t rec, uint64_t line)
{
    const char* p = s->mf.p;
    const size_t e = next_line(p, s->mf.n, rec);
    size_t len = e - rec;
    if (len && p[e - 1] == '\n') --len;
    std::string name, bc;
    if (line < s->latch_line) { HeaderParser undecided; undecided.parse(p + rec, len, &name, &bc); }
    else { HeaderParser hp = s->latch; hp.parse(p + rec, len, &name, &bc); }
    return bc;
}

// start of the line `k` lines before the line that starts at pos (pos > 0 is a line start)
size_t lines_back(const char* p, size_t pos, int k)
{
    size_t q = pos - 1; // the newline that ends the previous line
    for (int i = 0; i < k; ++i) {
        const void* nl = q ? memrchr(p, '\n', q) : nullptr;
        if (!nl) return 0;
        q = (size_t)((const char*)nl - p);
    }
    return q + 1;
}

// first record boundary at or after byte `at`, given the number of newlines before `at`
void record_boundary(const char* p, size_t n, size_t at, uint64_t newlines_before, size_t* pos_out, uint64_t* line_out)
{
    size_t pos = std::min(at, n);
    uint64_t ln = newlines_before;
    if (pos > 0 && pos < n && p[pos - 1] != '\n') { pos = next_line(p, n, pos); ++ln; } // inside a line: it belongs to what comes before
    while (pos < n && ln % 8 != 0) { pos = next_line(p, n, pos); ++ln; }
    *pos_out = pos; *line_out = ln;
}

// Where a batch / range may start at or after the record boundary B: B itself when the pair before it flushed a cloud or
// left a ""-labelled cloud open; otherwise just after the next pair whose barcode differs (that pair still belongs to
// the open cloud - the reference's off-by-one).  Returns the position, its line number and last_barcode there.
void align_start(const pg_fastq_stream* s, size_t B, uint64_t lineB, size_t* pos_out, uint64_t* line_out, std::string* last_out)
{
    const char* p = s->mf.p;
    const size_t n = s->mf.n;
    if (B == 0 || B >= n) { *pos_out = std::min(B, n); *line_out = lineB; last_out->clear(); return; } // (at EOF `last` is never used)
    const size_t r1 = lines_back(p, B, 8);
    const std::string p1 = record_barcode(s, r1, lineB - 8);
    std::string p2;
    if (lineB >= 16) p2 = record_barcode(s, lines_back(p, r1, 8), lineB - 16);
    if (p1 != p2 || p1.empty()) { *pos_out = B; *line_out = lineB; *last_out = p1; return; }
    size_t pos = B;
    uint64_t ln = lineB;
    std::string last = p1;
    while (pos < n) {
        const std::string bc = record_barcode(s, pos, ln);
        for (int i = 0; i < 8 && pos < n; ++i) { pos = next_line(p, n, pos); ++ln; }
        if (bc != last) { last = bc; break; }
    }
    *pos_out = pos; *line_out = ln; *last_out = last;
}

// parse the whole records of [begin, end) (both record boundaries; line0 = line number at begin) with T threads and
// append them to fq.  `last` is last_barcode at begin and is updated.
void parse_records_parallel(pg_fastq_stream* s, pg_fastq* fq, size_t begin, size_t end, uint64_t line0, std::string* last, int T)
{
    const char* p = s->mf.p;
    if (end <= begin) return;
    T = (int)std::max<size_t>(1, std::min<size_t>((size_t)T, (end - begin) / 4096 + 1));
    // 1. newline counts per range
    std::vector<size_t> cut((size_t)T + 1);
    for (int t = 0; t <= T; ++t) cut[(size_t)t] = begin + (end - begin) / T * t;
    cut[(size_t)T] = end;
    std::vector<uint64_t> newlines((size_t)T, 0);
    run_threads(T, [&](int t) { newlines[(size_t)t] = count_newlines(p, cut[(size_t)t], cut[(size_t)t + 1]); });
    // 2. first record boundary at or after every cut
    std::vector<Piece> pieces((size_t)T);
    {
        uint64_t lines_before = line0;
        std::vector<size_t> start((size_t)T + 1, end);
        std::vector<uint64_t> start_line((size_t)T + 1, 0);
        for (int t = 0; t < T; ++t) {
            size_t ps; uint64_t ln;
            record_boundary(p, end, cut[(size_t)t], lines_before, &ps, &ln);
            start[(size_t)t] = ps; start_line[(size_t)t] = ln;
            lines_before += newlines[(size_t)t];
        }
        start[0] = begin; start_line[0] = line0;
        for (int t = 0; t < T; ++t) {
            pieces[(size_t)t].begin = start[(size_t)t];
            pieces[(size_t)t].end = std::max(start[(size_t)t], start[(size_t)t + 1]);
            pieces[(size_t)t].line0 = start_line[(size_t)t];
        }
        for (int t = 1; t < T; ++t) // a range without any record boundary: empty piece
            if (pieces[(size_t)t].begin < pieces[(size_t)t - 1].end) pieces[(size_t)t].begin = pieces[(size_t)t].end = pieces[(size_t)t - 1].end;
    }
    // 3a. sizes
    run_threads(T, [&](int t) {
        Piece& pc = pieces[(size_t)t];
        for_lines(p, pc.begin, pc.end, pc.line0, [&](uint64_t ln, const char*, size_t len) {
            if (ln % 4 == 1) { pc.seq_bytes += len + 1; ++pc.n_reads; }
        });
    });
    const size_t seq0 = fq->seq.size(), reads0 = fq->flag.size();
    std::vector<size_t> seq_off((size_t)T + 1, seq0), read_off((size_t)T + 1, reads0);
    for (int t = 0; t < T; ++t) {
        seq_off[(size_t)t + 1] = seq_off[(size_t)t] + pieces[(size_t)t].seq_bytes;
        read_off[(size_t)t + 1] = read_off[(size_t)t] + pieces[(size_t)t].n_reads;
    }
    const size_t total_seq = seq_off[(size_t)T], total_reads = read_off[(size_t)T];
    fq->seq.resize_uninit(total_seq);
    if (fq->want_qual) fq->qual.resize_uninit(total_seq);
    fq->flag.resize_uninit(total_reads);
    fq->off.resize_uninit(total_reads + 1);
    // 3b. parse + copy
    const uint64_t latch_line = s->latch_line;
    const int read_type = s->latch.read_type;
    run_threads(T, [&](int t) {
        Piece& pc = pieces[(size_t)t];
        HeaderParser hp;
        hp.read_type = read_type;
        std::string bc, lastp;
        bool have_last = false;
        size_t so = seq_off[(size_t)t];
        int64_t r = (int64_t)read_off[(size_t)t];
        size_t s1 = 0, l1 = 0, s2 = 0, l2 = 0; // start / length of the pending R1 / R2 (for their quality lines)
        bool q1 = false, q2 = false;
        uint8_t* seq = fq->seq.data();
        uint8_t* qual = fq->want_qual ? fq->qual.data() : nullptr;
        auto add_read = [&](const char* sp, size_t len) {
            memcpy(seq + so, sp, len);
            seq[so + len] = '\n';
            if (qual) memset(qual + so, 0xFF, len + 1);
            so += len + 1;
            fq->off[(size_t)r + 1] = (int64_t)so; // (a read's own start is the previous read's end: written by its owner)
            fq->flag[(size_t)r] = 0;
            return r++;
        };
        auto set_qual = [&](size_t start, size_t rl, const char* q, size_t len) {
            if (qual) memcpy(qual + start, q, len < rl ? len : rl);
        };
        for_lines(p, pc.begin, pc.end, pc.line0, [&](uint64_t ln, const char* sp, size_t len) {
            switch ((ln + 1) % 8) {
            case 1:
                if (ln < latch_line) { HeaderParser undecided; undecided.parse(sp, len, nullptr, &bc); }
                else hp.parse(sp, len, nullptr, &bc);
                break;
            case 2: s1 = so; l1 = len; q1 = true; add_read(sp, len); break;
            case 4: if (q1) set_qual(s1, l1, sp, len); q1 = false; break;
            case 6: {
                s2 = so; l2 = len; q2 = true;
                const int64_t r2 = add_read(sp, len);
                if (!have_last) { // decided when the pieces are stitched
                    pc.has_pair = true; pc.first_bc = bc; pc.first_r2 = r2 - (int64_t)read_off[(size_t)t];
                    have_last = true; lastp = bc;
                } else if (bc != lastp) {
                    fq->flag[(size_t)r2] |= PG_READ_CHANGE;
                    pc.changes.emplace_back(r2 - (int64_t)read_off[(size_t)t], bc);
                    lastp = bc;
                }
                break;
            }
            case 0: if (q2) set_qual(s2, l2, sp, len); q2 = false; break;
            default: break;
            }
        });
        pc.last_bc = lastp;
    });
    // 4. stitch last_barcode across the cuts, collect the labels in file order
    for (int t = 0; t < T; ++t) {
        Piece& pc = pieces[(size_t)t];
        if (!pc.has_pair) continue;
        if (pc.first_bc != *last) {
            fq->flag[read_off[(size_t)t] + (size_t)pc.first_r2] |= PG_READ_CHANGE;
            fq->labels.push_back(pc.first_bc);
        }
        for (auto& ch : pc.changes) fq->labels.push_back(ch.second);
        *last = pc.last_bc;
    }
}

// one more pair, sequentially (the tail of a batch: up to the next flush)
void parse_one_record(pg_fastq_stream* s, pg_fastq* fq, std::string* last)
{
    const char* p = s->mf.p;
    const size_t n = std::min(s->mf.n, s->end);
    std::string name, bc;
    int64_t r1 = -1, r2 = -1;
    for (int i = 0; i < 8 && s->pos < n; ++i) {
        const size_t e = next_line(p, n, s->pos);
        size_t len = e - s->pos;
        if (len && p[e - 1] == '\n') --len;
        const char* sp = p + s->pos;
        switch (i) {
        case 0:
            if (s->line < s->latch_line) { HeaderParser undecided; undecided.parse(sp, len, &name, &bc); }
            else { HeaderParser hp = s->latch; hp.parse(sp, len, &name, &bc); }
            break;
        case 1: r1 = fq->add_read(sp, len, 0); break;
        case 3: fq->set_qual(r1, sp, len); break;
        case 5:
            r2 = fq->add_read(sp, len, 0);
            if (bc != *last) { fq->change_after(r2, bc); *last = bc; }
            break;
        case 7: fq->set_qual(r2, sp, len); break;
        default: break;
        }
        s->pos = e; ++s->line;
    }
}

bool ends_at_flush(const pg_fastq* fq)
{
    const size_t n = fq->flag.size();
    return n == 0 || (fq->flag[n - 1] & PG_READ_CHANGE);
}

int next_mapped(pg_fastq_stream* s, int64_t target, pg_fastq** out)
{
    if (s->pos >= s->end) { s->done = true; return PG_OK; }
    std::unique_ptr<pg_fastq> fq(new pg_fastq(s->pinned, s->last));
    fq->want_qual = s->want_qual;
    const char* p = s->mf.p;
    // file bytes that hold `target` sequence bytes: estimated from the record at the cursor
    size_t win = s->end - s->pos;
    if (target > 0) {
        size_t e = s->pos, seq_len = 0;
        for (int i = 0; i < 8 && e < s->end; ++i) { const size_t e2 = next_line(p, s->end, e); if (i == 1 || i == 5) seq_len += e2 - e; e = e2; }
        const double ratio = seq_len ? (double)(e - s->pos) / (double)seq_len : 3.0;
        const double w = (double)target * ratio;
        if (w < (double)win) win = (size_t)w;
    }
    const int T = parallel_threads();
    size_t B = s->end; uint64_t lineB = 0;
    if (s->pos + win < s->end) {
        const uint64_t nl = s->line + count_newlines_parallel(p, s->pos, s->pos + win, T);
        record_boundary(p, s->end, s->pos + win, nl, &B, &lineB);
    }
    if (B >= s->end) { B = s->end; }
    parse_records_parallel(s, fq.get(), s->pos, B, s->line, &s->last, T);
    if (B < s->end) { s->pos = B; s->line = lineB; }
    else s->pos = s->end;
    // up to the next flush point (or stop at once inside a ""-labelled cloud, which is dropped whole)
    while (s->pos < s->end && !ends_at_flush(fq.get()) && !s->last.empty()) parse_one_record(s, fq.get(), &s->last);
    fq->finish();
    *out = fq.release();
    return PG_OK;
}

// ---- sequential mode (gzip, paired files, pipes): same batches, one line at a time ----
int next_sequential(pg_fastq_stream* s, int64_t target, pg_fastq** out)
{
    if (!s->more1 && !s->more2) { s->done = true; return PG_OK; }
    std::unique_ptr<pg_fastq> fq(new pg_fastq(s->pinned, s->last));
    fq->want_qual = s->want_qual;
    const char *l1, *l2; size_t len1, len2;
    std::string nm1, b1, nm2, b2;
    int64_t r1 = -1, r2 = -1;
    auto may_cut = [&]() { return target > 0 && (int64_t)fq->seq.size() >= target && (!s->open_nonempty || s->last.empty()); };
    if (!s->paired) {
        std::string name, bc;
        while (s->more1) {
            if (!s->ra.next(&l1, &len1)) { s->more1 = false; s->more2 = false; break; }
            switch (++s->n1 % 8) {
            case 1: s->hp.parse(l1, len1, &name, &bc); b1 = bc; break;
            case 2: r1 = fq->add_read(l1, len1, 0); s->open_nonempty = true; break;
            case 4: fq->set_qual(r1, l1, len1); r1 = -1; break;
            case 6:
                r2 = fq->add_read(l1, len1, 0); s->open_nonempty = true;
                if (b1 != s->last) { fq->change_after(r2, b1); s->last = b1; s->open_nonempty = false; }
                break;
            case 0: fq->set_qual(r2, l1, len1); r2 = -1; break;
            default: break;
            }
            if (s->n1 % 8 == 0 && may_cut()) break;
        }
        if (s->ra.failed) return PG_ERR_IO;
    } else {
        while (s->more1) {
            if (!s->ra.next(&l1, &len1)) { s->more1 = false; break; }
            if (!s->more2 || !s->rb.next(&l2, &len2)) { s->more2 = false; l2 = ""; len2 = 0; }
            switch (++s->n1 % 4) {
            case 1:
                s->hp.parse(l1, len1, &nm1, &b1);
                s->hp.parse(l2, len2, &nm2, &b2);
                s->pair_ok = nm1 == nm2 && b1 == b2;
                s->pair_bc = b1;
                break;
            case 2:
                if (s->pair_ok) {
                    r1 = fq->add_read(l1, len1, 0);
                    r2 = fq->add_read(l2, len2, 0);
                    s->open_nonempty = true;
                    if (s->pair_bc != s->last) { fq->change_after(r2, s->pair_bc); s->last = s->pair_bc; s->open_nonempty = false; }
                } else { // counted by jellyfish, but appended to no cloud (count_kmer.cpp:195-196)
                    r1 = fq->add_read(l1, len1, PG_READ_NOFEAT);
                    r2 = fq->add_read(l2, len2, PG_READ_NOFEAT);
                }
                break;
            case 0: fq->set_qual(r1, l1, len1); fq->set_qual(r2, l2, len2); r1 = r2 = -1; break;
            default: break;
            }
            if (s->n1 % 4 == 0 && may_cut()) break;
        }
        // records left in file 2 are still k-mer counted (jellyfish reads both files whole)
        if (!s->more1 && s->more2) {
            uint64_t m = s->n1;
            while (s->rb.next(&l2, &len2)) {
                switch (++m % 4) {
                case 2: r2 = fq->add_read(l2, len2, PG_READ_NOFEAT); break;
                case 0: fq->set_qual(r2, l2, len2); r2 = -1; break;
                default: break;
                }
            }
            s->more2 = false;
        }
        if (s->ra.failed || s->rb.failed) return PG_ERR_IO;
    }
    fq->finish();
    *out = fq.release();
    return PG_OK;
}

} // namespace

// dst[0..n) = src[0..n) with all host cores (file mapping -> pinned staging buffer of the device ingest)
extern "C" int pg_parallel_memcpy(void* dst, const void* src, int64_t n)
{
    if (n < 0 || (n && (!dst || !src))) return PG_ERR_INVALID;
    try {
        const int T = (int)std::max<int64_t>(1, std::min<int64_t>(parallel_threads(), n / (4 << 20) + 1));
        run_threads(T, [&](int t) {
            const int64_t lo = n / T * t, hi = t + 1 == T ? n : n / T * (t + 1);
            memcpy((char*)dst + lo, (const char*)src + lo, (size_t)(hi - lo));
        });
        return PG_OK;
    } catch (...) { return PG_ERR_IO; }
}

// file bytes [offset, offset + n) -> dst with all host cores, each thread pread()ing its slice: one kernel copy from the page
// cache (or the disk) straight into the caller's - pinned - buffer, no page faults on a mapping
extern "C" int pg_parallel_pread(const char* path, int64_t offset, int64_t n, void* dst)
{
    if (!path || offset < 0 || n < 0 || (n && !dst)) return PG_ERR_INVALID;
    if (n == 0) return PG_OK;
    try {
        const int fd = ::open(path, O_RDONLY);
        if (fd < 0) return PG_ERR_IO;
        const int T = (int)std::max<int64_t>(1, std::min<int64_t>(parallel_threads(), n / (4 << 20) + 1));
        std::vector<int> ok((size_t)T, 1);
        run_threads(T, [&](int t) {
            int64_t lo = n / T * t;
            const int64_t hi = t + 1 == T ? n : n / T * (t + 1);
            while (lo < hi) {
                const ssize_t got = ::pread(fd, (char*)dst + lo, (size_t)std::min<int64_t>(hi - lo, 1 << 30), (off_t)(offset + lo));
                if (got <= 0) { ok[(size_t)t] = 0; return; }
                lo += got;
            }
        });
        ::close(fd);
        for (int v : ok) if (!v) return PG_ERR_IO;
        return PG_OK;
    } catch (...) { return PG_ERR_IO; }
}

extern "C" int pg_fastq_count_lines(const char* path, int64_t byte_lo, int64_t byte_hi, int64_t* n_newlines)
{
    if (!path || !n_newlines || byte_lo < 0) return PG_ERR_INVALID;
    try {
        MappedFile mf;
        if (!mf.open(path)) return PG_ERR_IO;
        const size_t lo = std::min((size_t)byte_lo, mf.n), hi = byte_hi < 0 ? mf.n : std::min((size_t)byte_hi, mf.n);
        *n_newlines = (int64_t)count_newlines_parallel(mf.p, lo, std::max(lo, hi), parallel_threads());
        return PG_OK;
    } catch (...) { return PG_ERR_IO; }
}

extern "C" int pg_fastq_stream_open(const char* path1, const char* path2, int flags, int64_t byte_lo, int64_t byte_hi, int64_t lines_before_lo,
                                    pg_fastq_stream** out)
{
    if (!path1 || !out || byte_lo < 0 || lines_before_lo < 0) return PG_ERR_INVALID;
    *out = nullptr;
    try {
        std::unique_ptr<pg_fastq_stream> s(new pg_fastq_stream());
        s->want_qual = (flags & PG_FQ_QUAL) != 0;
        s->pinned = (flags & PG_FQ_PINNED) != 0;
        const bool ranged = byte_lo > 0 || byte_hi >= 0;
        const bool paired = path2 && path2[0];
        bool map_ok = !paired && !(getenv("PG_FASTQ_SEQUENTIAL") && getenv("PG_FASTQ_SEQUENTIAL")[0] == '1') && s->mf.open(path1);
        if (map_ok && s->mf.n >= 2 && (unsigned char)s->mf.p[0] == 0x1f && (unsigned char)s->mf.p[1] == 0x8b) map_ok = false; // gzip
        if (!map_ok) {
            if (ranged) return PG_ERR_INVALID; // byte ranges need a mapped plain-text interleaved file
            s->paired = paired;
            if (!s->ra.open(path1)) return PG_ERR_IO;
            if (paired && !s->rb.open(path2)) return PG_ERR_IO;
            *out = s.release();
            return PG_OK;
        }
        s->mapped = true;
        const char* p = s->mf.p;
        const size_t n = s->mf.n;
        {   // read_type, latched by the first decisive header of the file (normally the very first line); headers before it
            // are parsed with an undecided parser, exactly as the sequential loop would
            std::string nm, bc;
            uint64_t ln = 0;
            for (size_t pos = 0; pos < n && s->latch.read_type == 0; ++ln) {
                const size_t e = next_line(p, n, pos);
                size_t len = e - pos;
                if (len && p[e - 1] == '\n') --len;
                if (ln % 8 == 0) {
                    s->latch.parse(p + pos, len, &nm, &bc);
                    if (s->latch.read_type != 0) s->latch_line = ln;
                }
                pos = e;
            }
        }
        const int T = parallel_threads();
        const size_t lo = std::min((size_t)byte_lo, n), hi = byte_hi < 0 ? n : std::max(lo, std::min((size_t)byte_hi, n));
        size_t B; uint64_t lineB;
        record_boundary(p, n, lo, (uint64_t)lines_before_lo, &B, &lineB);
        if (lo == 0) { s->pos = 0; s->line = 0; s->last.clear(); }
        else align_start(s.get(), B, lineB, &s->pos, &s->line, &s->last);
        if (hi >= n) s->end = n;
        else {
            const uint64_t nl_hi = (uint64_t)lines_before_lo + count_newlines_parallel(p, lo, hi, T);
            record_boundary(p, n, hi, nl_hi, &B, &lineB);
            uint64_t le; std::string dummy;
            align_start(s.get(), B, lineB, &s->end, &le, &dummy);
        }
        if (s->pos > s->end) s->pos = s->end;
        *out = s.release();
        return PG_OK;
    } catch (...) { return PG_ERR_IO; }
}

extern "C" int pg_fastq_stream_next(pg_fastq_stream* s, int64_t target_seq_bytes, pg_fastq** out)
{
    if (!s || !out) return PG_ERR_INVALID;
    *out = nullptr;
    if (s->done) return PG_OK;
    try {
        return s->mapped ? next_mapped(s, target_seq_bytes, out) : next_sequential(s, target_seq_bytes, out);
    } catch (...) { return PG_ERR_IO; } // out of memory included: nothing crosses the C boundary
}

extern "C" void pg_fastq_stream_close(pg_fastq_stream* s) { delete s; }

// the whole input as one batch
extern "C" int pg_fastq_parse(const char* path1, const char* path2, int flags, pg_fastq** out)
{
    if (!path1 || !out) return PG_ERR_INVALID;
    *out = nullptr;
    pg_fastq_stream* s = nullptr;
    int rc = pg_fastq_stream_open(path1, path2, flags, 0, -1, 0, &s);
    if (rc != PG_OK) return rc;
    rc = pg_fastq_stream_next(s, 0, out);
    if (rc == PG_OK && !*out) { // empty input: an empty batch, not "end of stream"
        try { pg_fastq* fq = new pg_fastq((flags & PG_FQ_PINNED) != 0); fq->want_qual = (flags & PG_FQ_QUAL) != 0; fq->finish(); *out = fq; }
        catch (...) { rc = PG_ERR_IO; }
    }
    pg_fastq_stream_close(s);
    return rc;
}

extern "C" void pg_fastq_free(pg_fastq* fq) { delete fq; }

extern "C" void pg_fastq_reads(const pg_fastq* fq, pg_reads* out)
{
    out->seq = fq->seq.data();
    out->qual = fq->want_qual ? fq->qual.data() : nullptr;
    out->read_off = fq->off.data();
    out->read_flag = fq->flag.data();
    out->n_reads = (int64_t)fq->flag.size();
    out->n_bytes = (int64_t)fq->seq.size();
}
extern "C" int64_t pg_fastq_n_groups(const pg_fastq* fq) { return (int64_t)fq->labels.size(); }
extern "C" const uint8_t* pg_fastq_group_keep(const pg_fastq* fq) { return fq->keep.data(); }
// all labels at once: offsets[n_groups + 1] into buf (no terminators); returns the bytes needed (call with cap = 0 to size)
extern "C" int64_t pg_fastq_group_labels(const pg_fastq* fq, char* buf, int64_t cap, int64_t* offsets)
{
    int64_t need = 0;
    for (auto& l : fq->labels) need += (int64_t)l.size();
    if (!buf || !offsets || cap < need) return need;
    int64_t at = 0;
    size_t g = 0;
    for (auto& l : fq->labels) { offsets[g++] = at; memcpy(buf + at, l.data(), l.size()); at += (int64_t)l.size(); }
    offsets[g] = at;
    return need;
}

extern "C" const char* pg_fastq_group_label(const pg_fastq* fq, int64_t g)
{
    return (g >= 0 && g < (int64_t)fq->labels.size()) ? fq->labels[g].c_str() : "";
}
