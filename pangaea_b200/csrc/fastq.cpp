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
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

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

struct pg_fastq {
    std::vector<uint8_t> seq, qual, flag, keep;
    std::vector<int64_t> off;
    std::vector<std::string> labels;
    bool want_qual = false;
    int64_t pending_qual_read = -1;

    pg_fastq() { off.push_back(0); labels.emplace_back(""); }
    int64_t add_read(const char* s, size_t n, uint8_t fl)
    {
        seq.insert(seq.end(), (const uint8_t*)s, (const uint8_t*)s + n);
        seq.push_back('\n');
        if (want_qual) qual.resize(seq.size(), 0xFF);
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

extern "C" int pg_fastq_parse(const char* path1, const char* path2, int want_qual, pg_fastq** out)
{
    if (!path1 || !out) return PG_ERR_INVALID;
    *out = nullptr;
    pg_fastq* fq = new pg_fastq();
    fq->want_qual = want_qual != 0;
    int rc = (path2 && path2[0]) ? parse_paired(fq, path1, path2) : parse_interleaved(fq, path1);
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
