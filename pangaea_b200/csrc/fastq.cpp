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
#include <new>
#include <string>
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
    std::vector<char> buf;
    size_t pos = 0, end = 0;
    bool eof = false;
    std::string spill; // a line that straddles two chunks

    bool open(const char* path)
    {
        f = gzopen(path, "rb"); // transparent for plain text, like the reference's igzstream
        if (!f) return false;
        gzbuffer(f, 1 << 22);
        buf.resize(1 << 24);
        return true;
    }
    ~LineReader() { if (f) gzclose(f); }
    bool fill()
    {
        if (eof) return false;
        int n = gzread(f, buf.data(), (unsigned)buf.size());
        pos = 0;
        end = n > 0 ? (size_t)n : 0;
        if (n <= 0) eof = true;
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
            const char* s = buf.data() + pos;
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
    void parse(const char* line, size_t len, std::string* name, std::string* bc)
    {
        if (read_type == 0) {
            if (find_bxz(line, len) != npos) read_type = 1;
            else if (find_char(line, len, '#', 0) != npos) read_type = 2;
        }
        if (read_type == 2) { // count_kmer.cpp:36-43
            size_t p1 = find_char(line, len, '#', 0);
            size_t p2 = find_char(line, len, '/', p1 + 1); // npos + 1 == 0, as in the reference
            substr(name, line, len, 0, p1);
            substr(bc, line, len, p1 + 1, p2 - p1 - 1);
            if (*bc == "0_0_0") bc->clear();
        } else { // count_kmer.cpp:44-51
            size_t e = npos;
            for (size_t i = 0; i < len; ++i)
                if (line[i] == ' ' || line[i] == '\r' || line[i] == '\t' || line[i] == '\n') { e = i; break; }
            substr(name, line, len, 0, e);
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

// growable array that does not zero-fill (resize_uninit of 10 GB would otherwise cost seconds); vector-like otherwise
template <class T>
struct Buf {
    T* p = nullptr;
    size_t n = 0, cap = 0;
    Buf() = default;
    Buf(const Buf&) = delete;
    Buf& operator=(const Buf&) = delete;
    ~Buf() { free(p); }
    T* data() { return p; }
    const T* data() const { return p; }
    size_t size() const { return n; }
    T& operator[](size_t i) { return p[i]; }
    const T& operator[](size_t i) const { return p[i]; }
    void reserve(size_t c)
    {
        if (c <= cap) return;
        size_t nc = std::max(c, cap + cap / 2 + 64);
        T* q = (T*)realloc(p, nc * sizeof(T));
        if (!q) throw std::bad_alloc();
        p = q; cap = nc;
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
    int64_t pending_qual_read = -1;

    pg_fastq() { off.push_back(0); labels.emplace_back(""); }
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

static int parse_interleaved(pg_fastq* fq, const char* path)
{
    LineReader r;
    if (!r.open(path)) return PG_ERR_IO;
    HeaderParser hp;
    std::string name, bc, last;
    const char* line; size_t len;
    uint64_t n = 0;
    int64_t r1 = -1, r2 = -1;
    while (r.next(&line, &len)) {
        switch (++n % 8) {
        case 1: hp.parse(line, len, &name, &bc); break;
        case 2: r1 = fq->add_read(line, len, 0); break;
        case 4: fq->set_qual(r1, line, len); r1 = -1; break;
        case 6:
            r2 = fq->add_read(line, len, 0);
            if (bc != last) { fq->change_after(r2, bc); last = bc; }
            break;
        case 0: fq->set_qual(r2, line, len); r2 = -1; break;
        default: break;
        }
    }
    return PG_OK;
}

static int parse_paired(pg_fastq* fq, const char* path1, const char* path2)
{
    LineReader a, b;
    if (!a.open(path1)) return PG_ERR_IO;
    if (!b.open(path2)) return PG_ERR_IO;
    HeaderParser hp;
    std::string n1, b1, n2, b2, last;
    const char *l1, *l2; size_t len1, len2;
    uint64_t n = 0;
    int64_t r1 = -1, r2 = -1;
    bool more2 = true;
    while (a.next(&l1, &len1)) {
        if (!more2 || !b.next(&l2, &len2)) { more2 = false; l2 = ""; len2 = 0; }
        switch (++n % 4) {
        case 1:
            hp.parse(l1, len1, &n1, &b1);
            hp.parse(l2, len2, &n2, &b2);
            break;
        case 2:
            if (n1 == n2 && b1 == b2) {
                r1 = fq->add_read(l1, len1, 0);
                r2 = fq->add_read(l2, len2, 0);
                if (b1 != last) { fq->change_after(r2, b1); last = b1; }
            } else { // counted by jellyfish, but appended to no cloud (count_kmer.cpp:195-196)
                r1 = fq->add_read(l1, len1, PG_READ_NOFEAT);
                r2 = fq->add_read(l2, len2, PG_READ_NOFEAT);
            }
            break;
        case 0: fq->set_qual(r1, l1, len1); fq->set_qual(r2, l2, len2); r1 = r2 = -1; break;
        default: break;
        }
    }
    // records left in file 2 are still k-mer counted (jellyfish reads both files whole)
    if (more2) {
        uint64_t m = n;
        while (b.next(&l2, &len2)) {
            switch (++m % 4) {
            case 2: r2 = fq->add_read(l2, len2, PG_READ_NOFEAT); break;
            case 0: fq->set_qual(r2, l2, len2); r2 = -1; break;
            default: break;
            }
        }
    }
    return PG_OK;
}

// ---------------------------------------------------------------------------
// Parallel reader for plain-text interleaved files (the production input, pangaea.py -i).
// The sequential loop above handles ~0.85 GB/s: 30 s for the 25 GB of a 50 M-pair run, against 0.1 s on the GPU.
// Same decisions, taken by T threads:
//   1. the file is mapped and cut into T byte ranges; every thread counts the newlines of its range;
//   2. a prefix sum gives the line number at every cut, so each thread can start at the first RECORD boundary
//      (line number = 0 mod 8) inside its range and stop at the first one of the next range;
//   3. pass A sizes the output (sequence bytes, reads) per thread, pass B parses headers and copies sequence /
//      quality lines straight to their final offsets;
//   4. the two pieces of sequential state are stitched afterwards: read_type (latched by the first decisive header
//      of the file, count_kmer.cpp:28-33 - found by a short sequential scan before the threads start) and
//      last_barcode at each cut (the first pair of a range is compared with the last pair before it).
// Anything unusual (gzip input, paired files, a file too small to matter) takes the sequential reader.
// ---------------------------------------------------------------------------
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

struct Piece {                 // what one thread contributes
    size_t begin = 0, end = 0; // byte range of whole records: lines [line0, ...) with line0 = 0 mod 8
    uint64_t line0 = 0;
    size_t seq_bytes = 0, n_reads = 0;          // pass A
    std::vector<std::pair<int64_t, std::string>> changes; // (read index inside the piece, new label), in order
    bool has_pair = false;
    std::string first_bc, last_bc;
    int64_t first_r2 = -1;     // read index (inside the piece) of the first pair's R2
};

// calls fn(line_number, ptr, len) for every line of [begin, end); the last line may lack its newline
template <class Fn>
void for_lines(const char* p, size_t begin, size_t end, uint64_t line0, Fn&& fn)
{
    uint64_t n = line0;
    size_t pos = begin;
    while (pos < end) {
        const char* nl = (const char*)memchr(p + pos, '\n', end - pos);
        const size_t len = nl ? (size_t)(nl - (p + pos)) : end - pos;
        fn(n++, p + pos, len);
        pos += len + 1;
    }
}

int parallel_threads()
{
    const char* e = getenv("PG_FASTQ_THREADS");
    if (e && atoi(e) > 0) return std::min(atoi(e), 256);
    const unsigned hc = std::thread::hardware_concurrency();
    return (int)std::max(1u, std::min(hc ? hc : 1u, 32u));
}

} // namespace

// returns PG_OK and fills fq, or a negative value when the sequential reader must be used (-100) / on I/O error
static int parse_interleaved_parallel(pg_fastq* fq, const char* path)
{
    const int T = parallel_threads();
    if (T < 2) return -100;
    MappedFile mf;
    if (!mf.open(path)) return -100; // not a regular file (pipe ...): the sequential reader copes
    const char* env_min = getenv("PG_FASTQ_PARALLEL_MIN");
    const size_t min_bytes = env_min ? (size_t)atoll(env_min) : ((size_t)64 << 20);
    if (mf.n < min_bytes || mf.n < 2) return -100;
    if ((unsigned char)mf.p[0] == 0x1f && (unsigned char)mf.p[1] == 0x8b) return -100; // gzip
    const char* p = mf.p;
    const size_t n = mf.n;

    // read_type, latched by the first decisive header of the file (normally the very first line); headers before it
    // are parsed with an undecided parser, exactly as the sequential loop would
    HeaderParser latch;
    uint64_t latch_line = ~0ull;
    {
        std::string nm, bc;
        uint64_t ln = 0;
        for (size_t pos = 0; pos < n && latch.read_type == 0; ++ln) {
            const char* nl = (const char*)memchr(p + pos, '\n', n - pos);
            const size_t len = nl ? (size_t)(nl - (p + pos)) : n - pos;
            if (ln % 8 == 0) {
                latch.parse(p + pos, len, &nm, &bc);
                if (latch.read_type != 0) latch_line = ln;
            }
            pos += len + 1;
        }
    }

    // 1. newline counts per range
    std::vector<size_t> cut(T + 1);
    for (int t = 0; t <= T; ++t) cut[t] = n / T * t;
    cut[T] = n;
    std::vector<uint64_t> newlines(T, 0);
    {
        std::vector<std::thread> th;
        for (int t = 0; t < T; ++t)
            th.emplace_back([&, t]() {
                uint64_t c = 0;
                size_t pos = cut[t];
                while (pos < cut[t + 1]) {
                    const char* nl = (const char*)memchr(p + pos, '\n', cut[t + 1] - pos);
                    if (!nl) break;
                    ++c;
                    pos = (size_t)(nl - p) + 1;
                }
                newlines[t] = c;
            });
        for (auto& x : th) x.join();
    }
    // 2. first record boundary at or after every cut
    std::vector<Piece> pieces(T);
    {
        uint64_t lines_before = 0; // complete lines before cut[t] = newlines before it
        std::vector<size_t> start(T + 1, n);
        std::vector<uint64_t> start_line(T + 1, 0);
        for (int t = 0; t < T; ++t) {
            // first line that STARTS at or after cut[t]
            size_t pos = cut[t];
            uint64_t ln = lines_before;
            if (pos > 0 && p[pos - 1] != '\n') { // inside a line: it belongs to the range before
                const char* nl = (const char*)memchr(p + pos, '\n', n - pos);
                pos = nl ? (size_t)(nl - p) + 1 : n;
                ++ln;
            }
            while (pos < n && ln % 8 != 0) { // advance to a record boundary
                const char* nl = (const char*)memchr(p + pos, '\n', n - pos);
                pos = nl ? (size_t)(nl - p) + 1 : n;
                ++ln;
            }
            start[t] = pos;
            start_line[t] = ln;
            lines_before += newlines[t];
        }
        start[0] = 0; start_line[0] = 0;
        for (int t = 0; t < T; ++t) {
            pieces[t].begin = start[t];
            pieces[t].end = std::max(start[t], start[t + 1]);
            pieces[t].line0 = start_line[t];
        }
        for (int t = 1; t < T; ++t) // a range without any record boundary: empty piece
            if (pieces[t].begin < pieces[t - 1].end) pieces[t].begin = pieces[t].end = pieces[t - 1].end;
    }
    // 3a. sizes
    {
        std::vector<std::thread> th;
        for (int t = 0; t < T; ++t)
            th.emplace_back([&, t]() {
                Piece& pc = pieces[t];
                for_lines(p, pc.begin, pc.end, pc.line0, [&](uint64_t ln, const char*, size_t len) {
                    if (ln % 4 == 1) { pc.seq_bytes += len + 1; ++pc.n_reads; }
                });
            });
        for (auto& x : th) x.join();
    }
    std::vector<size_t> seq_off(T + 1, 0), read_off(T + 1, 0);
    for (int t = 0; t < T; ++t) { seq_off[t + 1] = seq_off[t] + pieces[t].seq_bytes; read_off[t + 1] = read_off[t] + pieces[t].n_reads; }
    const size_t total_seq = seq_off[T], total_reads = read_off[T];
    fq->seq.resize_uninit(total_seq);
    if (fq->want_qual) fq->qual.resize_uninit(total_seq);
    fq->flag.resize_uninit(total_reads);
    fq->off.resize_uninit(total_reads + 1);
    fq->off[0] = 0;
    // 3b. parse + copy
    {
        std::vector<std::thread> th;
        for (int t = 0; t < T; ++t)
            th.emplace_back([&, t]() {
                Piece& pc = pieces[t];
                HeaderParser hp;
                hp.read_type = latch.read_type;
                std::string name, bc, last;
                bool have_last = false;
                size_t so = seq_off[t];
                int64_t r = (int64_t)read_off[t];
                size_t s1 = 0, l1 = 0, s2 = 0, l2 = 0; // start / length of the pending R1 / R2 (for their quality lines)
                bool q1 = false, q2 = false;
                uint8_t* seq = fq->seq.data();
                uint8_t* qual = fq->want_qual ? fq->qual.data() : nullptr;
                auto add_read = [&](const char* s, size_t len) {
                    memcpy(seq + so, s, len);
                    seq[so + len] = '\n';
                    if (qual) memset(qual + so, 0xFF, len + 1);
                    so += len + 1;
                    fq->off[r + 1] = (int64_t)so; // (a read's own start is the previous read's end: written by its owner)
                    fq->flag[r] = 0;
                    return r++;
                };
                auto set_qual = [&](size_t start, size_t rl, const char* q, size_t len) {
                    if (qual) memcpy(qual + start, q, len < rl ? len : rl);
                };
                for_lines(p, pc.begin, pc.end, pc.line0, [&](uint64_t ln, const char* s, size_t len) {
                    switch ((ln + 1) % 8) {
                    case 1:
                        if (ln < latch_line) { HeaderParser undecided; undecided.parse(s, len, &name, &bc); }
                        else hp.parse(s, len, &name, &bc);
                        break;
                    case 2: s1 = so; l1 = len; q1 = true; add_read(s, len); break;
                    case 4: if (q1) set_qual(s1, l1, s, len); q1 = false; break;
                    case 6: {
                        s2 = so; l2 = len; q2 = true;
                        const int64_t r2 = add_read(s, len);
                        if (!have_last) { // decided when the pieces are stitched
                            pc.has_pair = true; pc.first_bc = bc; pc.first_r2 = r2 - (int64_t)read_off[t];
                            have_last = true; last = bc;
                        } else if (bc != last) {
                            fq->flag[r2] |= PG_READ_CHANGE;
                            pc.changes.emplace_back(r2 - (int64_t)read_off[t], bc);
                            last = bc;
                        }
                        break;
                    }
                    case 0: if (q2) set_qual(s2, l2, s, len); q2 = false; break;
                    default: break;
                    }
                });
                pc.last_bc = last;
            });
        for (auto& x : th) x.join();
    }
    // 4. stitch last_barcode across the cuts, collect the labels in file order
    {
        std::string last; // "" before the first pair of the file (count_kmer.cpp:238)
        for (int t = 0; t < T; ++t) {
            Piece& pc = pieces[t];
            if (!pc.has_pair) continue;
            if (pc.first_bc != last) {
                fq->flag[read_off[t] + (size_t)pc.first_r2] |= PG_READ_CHANGE;
                fq->labels.push_back(pc.first_bc);
            }
            for (auto& ch : pc.changes) fq->labels.push_back(ch.second);
            last = pc.last_bc;
        }
    }
    return PG_OK;
}

extern "C" int pg_fastq_parse(const char* path1, const char* path2, int want_qual, pg_fastq** out)
{
    if (!path1 || !out) return PG_ERR_INVALID;
    *out = nullptr;
    pg_fastq* fq = new pg_fastq();
    fq->want_qual = want_qual != 0;
    int rc;
    if (path2 && path2[0]) {
        rc = parse_paired(fq, path1, path2);
    } else {
        rc = parse_interleaved_parallel(fq, path1);
        if (rc == -100) rc = parse_interleaved(fq, path1); // small / gzip / not a regular file
    }
    if (rc != PG_OK) { delete fq; return rc; }
    fq->finish();
    *out = fq;
    return PG_OK;
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
extern "C" const char* pg_fastq_group_label(const pg_fastq* fq, int64_t g)
{
    return (g >= 0 && g < (int64_t)fq->labels.size()) ? fq->labels[g].c_str() : "";
}
