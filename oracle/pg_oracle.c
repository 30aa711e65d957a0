/*
 * pg_oracle.c - CPU restatement of Pangaea's read-cloud featurization path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing on the product path may include, link or
 * call this file; only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs use it, and only as the checker.
 *
 * Parity status
 *   - abundance / TNF / grouping / header parsing: PINNED against the compiled,
 *     unmodified reference binaries (oracle/_ref/count_kmer, count_tnf; see
 *     oracle/Makefile) by tests/test_oracle_vs_ref.py and the golden fixtures
 *     under tests/golden/ that those binaries produced.
 *   - global k-mer counting (the `jellyfish count -C` stage, an external
 *     dependency: bioconda `jellyfish`, no version pinned,
 *     /root/reference/environment.yaml:14; call sites src/feature.py:76-94,103):
 *     PARITY UNPINNED.  jellyfish is not in the reference tree and not
 *     installed; pgo_count_* restates its published behaviour (see the comment
 *     on pgo_count_read) and is anchored only on the reference's call sites.
 *
 * All file:line citations are relative to /root/reference/.
 */
#define _GNU_SOURCE
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <zlib.h>

/* ------------------------------------------------------------------ */
/* a1-a3: 2-bit k-mer arithmetic                                        */
/* ------------------------------------------------------------------ */

/* src/cpptools/count_kmer.cpp:73,81 (same in count_tnf.cpp:91,99): only the
 * four upper-case letters are bases; code = (c >> 1) & 3, i.e. A0 C1 T2 G3. */
static inline int pgo_is_base(unsigned char c) { return c == 'A' || c == 'C' || c == 'G' || c == 'T'; }
static inline uint64_t pgo_code(unsigned char c) { return (uint64_t)((c >> 1) & 3); }

/* src/cpptools/count_kmer.cpp:11-21: reverse the 2-bit groups of the word,
 * complement (code ^ 2 for every base), then drop the unused low groups. */
uint64_t pgo_revcomp(uint64_t x, int k)
{
    uint64_t r = 0;
    for (int i = 0; i < 32; ++i) { /* group i goes to group 31-i */
        r = (r << 2) | (x & 3);
        x >>= 2;
    }
    r ^= 0xAAAAAAAAAAAAAAAAULL;
    return r >> (2 * (32 - k));
}

/* src/cpptools/count_kmer.cpp:86 / count_tnf.cpp:104 */
uint64_t pgo_canonical(uint64_t v, int k)
{
    uint64_t rc = pgo_revcomp(v, k);
    return v < rc ? v : rc;
}

/* src/cpptools/count_kmer.cpp:69: (unsigned long)pow(2, 2k) - 1 */
static inline uint64_t pgo_kmask(int k) { return k >= 32 ? 0x7FFFFFFFFFFFFFFFULL /* x86 cvttsd2si overflow - 1 */ : ((1ULL << (2 * k)) - 1); }

/* ------------------------------------------------------------------ */
/* k-mer -> count map (stands for std::unordered_map<u64, ulong>,        */
/* src/cpptools/count_kmer.cpp:139)                                     */
/* ------------------------------------------------------------------ */
typedef struct pgo_table {
    uint64_t* keys; /* key + 1, 0 = free */
    uint64_t* vals;
    uint64_t cap, used;
} pgo_table;

static inline uint64_t pgo_mix(uint64_t h)
{
    h ^= h >> 33; h *= 0xff51afd7ed558ccdULL; h ^= h >> 33; h *= 0xc4ceb9fe1a85ec53ULL; h ^= h >> 33;
    return h;
}

pgo_table* pgo_table_new(void)
{
    pgo_table* t = (pgo_table*)calloc(1, sizeof(*t));
    t->cap = 1u << 16;
    t->keys = (uint64_t*)calloc(t->cap, 8);
    t->vals = (uint64_t*)calloc(t->cap, 8);
    return t;
}
void pgo_table_free(pgo_table* t) { if (t) { free(t->keys); free(t->vals); free(t); } }
uint64_t pgo_table_size(const pgo_table* t) { return t->used; }

static uint64_t* pgo_table_slot(pgo_table* t, uint64_t key, int create);
static void pgo_table_grow(pgo_table* t)
{
    pgo_table old = *t;
    t->cap = old.cap * 2; t->used = 0;
    t->keys = (uint64_t*)calloc(t->cap, 8);
    t->vals = (uint64_t*)calloc(t->cap, 8);
    for (uint64_t i = 0; i < old.cap; ++i)
        if (old.keys[i]) *pgo_table_slot(t, old.keys[i] - 1, 1) = old.vals[i];
    free(old.keys); free(old.vals);
}
static uint64_t* pgo_table_slot(pgo_table* t, uint64_t key, int create)
{
    if (create && (t->used + 1) * 10 > t->cap * 6) pgo_table_grow(t);
    uint64_t m = t->cap - 1, i = pgo_mix(key) & m;
    while (t->keys[i]) {
        if (t->keys[i] == key + 1) return &t->vals[i];
        i = (i + 1) & m;
    }
    if (!create) return NULL;
    t->keys[i] = key + 1; t->used++;
    return &t->vals[i];
}
/* assignment, as `kmer2frequency[key] = freq` (count_kmer.cpp:166) */
void pgo_table_set(pgo_table* t, uint64_t key, uint64_t v) { *pgo_table_slot(t, key, 1) = v; }
void pgo_table_add(pgo_table* t, uint64_t key, uint64_t v) { *pgo_table_slot(t, key, 1) += v; }
/* returns 1 and *v when present (find != end, count_kmer.cpp:87) */
int pgo_table_get(pgo_table* t, uint64_t key, uint64_t* v)
{
    uint64_t* s = pgo_table_slot(t, key, 0);
    if (!s) return 0;
    *v = *s; return 1;
}
/* unordered export; caller sizes the arrays with pgo_table_size */
void pgo_table_items(const pgo_table* t, uint64_t* keys, uint64_t* vals)
{
    uint64_t n = 0;
    for (uint64_t i = 0; i < t->cap; ++i)
        if (t->keys[i]) { keys[n] = t->keys[i] - 1; vals[n] = t->vals[i]; ++n; }
}

/* ------------------------------------------------------------------ */
/* a17: global canonical k-mer counting - jellyfish stand-in            */
/* ------------------------------------------------------------------ */
/*
 * Restates `jellyfish count -C -m k` (+ optional --min-qual-char) as
 * published in the jellyfish 2.x manual; reference call sites
 * src/feature.py:76,79,83 (paired: with --min-qual-char=?) and :94
 * (interleaved, production: without).  PARITY UNPINNED (see file header).
 *   - every sequence line of every FASTQ record is scanned, no barcode logic;
 *   - A/C/G/T in either case are bases, anything else restarts the window;
 *   - with min_qual != 0 a base whose quality byte is < min_qual acts like N;
 *   - -C: both orientations of a k-mer feed one counter.  Which orientation
 *     names the counter is irrelevant downstream because count_kmer
 *     re-canonicalises dump keys (count_kmer.cpp:166); keys here are the
 *     reference's own canonical form (min in A0 C1 T2 G3 code order).
 */
void pgo_count_read(pgo_table* t, const unsigned char* s, int64_t len, const unsigned char* q, int k, int min_qual)
{
    uint64_t mask = pgo_kmask(k), val = 0;
    int64_t run = 0;
    for (int64_t i = 0; i < len; ++i) {
        unsigned char c = s[i] & 0xDF; /* fold a/c/g/t to upper case */
        int ok = pgo_is_base(c);
        if (ok && q && min_qual && q[i] < (unsigned char)min_qual) ok = 0;
        if (!ok) { val = 0; run = 0; continue; }
        val = ((val << 2) & mask) | pgo_code(c);
        if (++run >= k) pgo_table_add(t, pgo_canonical(val, k), 1);
    }
}

/* line reader over zlib: gzread is transparent for plain text, exactly what the
 * reference's igzstream gives it (lib/gzstream/gzstream.C); '\n' is stripped,
 * '\r' is kept, a last line without '\n' is still delivered (std::getline). */
typedef struct { gzFile f; char* buf; size_t cap, len; } pgo_reader;
static int pgo_open(pgo_reader* r, const char* path)
{
    memset(r, 0, sizeof(*r));
    r->f = gzopen(path, "rb");
    if (!r->f) return 0;
    gzbuffer(r->f, 1 << 20);
    r->cap = 1 << 16; r->buf = (char*)malloc(r->cap);
    return 1;
}
static void pgo_close(pgo_reader* r) { if (r->f) gzclose(r->f); free(r->buf); r->f = NULL; r->buf = NULL; }
static int pgo_getline(pgo_reader* r)
{
    r->len = 0;
    if (!r->f) return 0;
    int got = 0;
    for (;;) {
        if (!gzgets(r->f, r->buf + r->len, (int)(r->cap - r->len))) break;
        got = 1;
        r->len += strlen(r->buf + r->len);
        if (r->len && r->buf[r->len - 1] == '\n') { r->buf[--r->len] = 0; return 1; }
        if (r->len + 1 >= r->cap) { r->cap *= 2; r->buf = (char*)realloc(r->buf, r->cap); } else break; /* EOF without \n */
    }
    return got;
}

/* counts every record of a FASTQ (4 lines per record) */
int pgo_count_fastq(pgo_table* t, const char* path, int k, int min_qual)
{
    pgo_reader r;
    if (!pgo_open(&r, path)) return -1;
    uint64_t n = 0;
    char* seq = NULL; size_t seqlen = 0, seqcap = 0;
    while (pgo_getline(&r)) {
        switch (++n % 4) {
        case 2:
            if (r.len + 1 > seqcap) { seqcap = 2 * (r.len + 1); seq = (char*)realloc(seq, seqcap); }
            memcpy(seq, r.buf, r.len + 1); seqlen = r.len;
            if (!min_qual) pgo_count_read(t, (unsigned char*)seq, (int64_t)seqlen, NULL, k, 0);
            break;
        case 0:
            if (min_qual) {
                int64_t m = (int64_t)(r.len < seqlen ? r.len : seqlen);
                pgo_count_read(t, (unsigned char*)seq, m, (unsigned char*)r.buf, k, min_qual);
            }
            break;
        default: break;
        }
    }
    free(seq);
    pgo_close(&r);
    return 0;
}

static void pgo_decode(uint64_t v, int k, char* out)
{
    static const char L[4] = { 'A', 'C', 'T', 'G' };
    for (int i = k - 1; i >= 0; --i) { out[i] = L[v & 3]; v >>= 2; }
    out[k] = 0;
}

/* `jellyfish dump -c -t` text: "KMER<TAB>COUNT\n" (src/feature.py:103) */
int pgo_dump_write(const pgo_table* t, const char* path, int k)
{
    FILE* f = fopen(path, "w");
    if (!f) return -1;
    char km[40];
    for (uint64_t i = 0; i < t->cap; ++i)
        if (t->keys[i]) { pgo_decode(t->keys[i] - 1, k, km); fprintf(f, "%s\t%llu\n", km, (unsigned long long)t->vals[i]); }
    fclose(f);
    return 0;
}

/* a8, src/cpptools/count_kmer.cpp:139-170: each dump line is re-scanned with the
 * same rolling encoder, re-canonicalised, and ASSIGNED (last one wins). */
int pgo_dump_load(pgo_table* t, const char* path, int k)
{
    FILE* f = fopen(path, "r");
    if (!f) return -1;
    char* line = NULL; size_t cap = 0; ssize_t n;
    uint64_t mask = pgo_kmask(k);
    while ((n = getline(&line, &cap, f)) >= 0) {
        if (n && line[n - 1] == '\n') line[--n] = 0;
        char* tab = strchr(line, '\t');
        size_t klen = tab ? (size_t)(tab - line) : (size_t)n;
        uint64_t freq = tab ? (uint64_t)strtol(tab + 1, NULL, 10) : 0;
        uint64_t val = 0; size_t run = 0;
        for (size_t i = 0; i < klen; ++i) {
            unsigned char c = (unsigned char)line[i];
            if (!pgo_is_base(c)) { val = 0; run = 0; continue; }
            val = ((val << 2) & mask) + pgo_code(c);
            if (++run == (size_t)k) { --run; pgo_table_set(t, pgo_canonical(val, k), freq); }
        }
    }
    free(line); fclose(f);
    return 0;
}

/* ------------------------------------------------------------------ */
/* a4: header -> (read name, barcode)                                   */
/* ------------------------------------------------------------------ */
typedef struct { char* p; size_t len, cap; } pgo_str;
static void pgo_str_set(pgo_str* s, const char* p, size_t n)
{
    if (n + 1 > s->cap) { s->cap = 2 * (n + 1); s->p = (char*)realloc(s->p, s->cap); }
    if (n) memcpy(s->p, p, n);
    s->p[n] = 0; s->len = n;
}
static void pgo_str_append(pgo_str* s, const char* p, size_t n)
{
    if (s->len + n + 1 > s->cap) { s->cap = 2 * (s->len + n + 1); s->p = (char*)realloc(s->p, s->cap); }
    memcpy(s->p + s->len, p, n); s->len += n; s->p[s->len] = 0;
}
static int pgo_str_eq(const pgo_str* a, const pgo_str* b) { return a->len == b->len && (a->len == 0 || !memcmp(a->p, b->p, a->len)); }

/* read_type is one process-wide latch in the reference (count_kmer.cpp:24):
 * 0 = undecided, 1 = "10x", 2 = "stLFR". */
typedef struct { int read_type; } pgo_hdr_state;

/* std::string::substr(pos, n) clamps n; pos > size throws -> we return "" and
 * flag it (the reference would abort; never produced by run_pangaea). */
static void pgo_substr(pgo_str* out, const char* s, size_t len, size_t pos, size_t n)
{
    if (pos > len) { pgo_str_set(out, "", 0); return; }
    if (n > len - pos) n = len - pos;
    pgo_str_set(out, s + pos, n);
}
static size_t pgo_find_char(const char* s, size_t len, char c, size_t from)
{
    if (from >= len) return (size_t)-1;
    const char* p = (const char*)memchr(s + from, c, len - from);
    return p ? (size_t)(p - s) : (size_t)-1;
}
static size_t pgo_find_bxz(const char* s, size_t len)
{
    if (len < 4) return (size_t)-1;
    const char* p = (const char*)memmem(s, len, "BX:Z", 4);
    return p ? (size_t)(p - s) : (size_t)-1;
}

/* src/cpptools/count_kmer.cpp:25-53 (identical in count_tnf.cpp:24-52) */
static void pgo_get_barcode(pgo_hdr_state* st, const char* line, size_t len, pgo_str* name, pgo_str* bc)
{
    const size_t npos = (size_t)-1;
    if (st->read_type == 0) {
        if (pgo_find_bxz(line, len) != npos) st->read_type = 1;
        else if (pgo_find_char(line, len, '#', 0) != npos) st->read_type = 2;
    }
    if (st->read_type == 2) {
        size_t p1 = pgo_find_char(line, len, '#', 0);
        size_t p2 = pgo_find_char(line, len, '/', p1 + 1); /* npos + 1 wraps to 0 like size_t */
        pgo_substr(name, line, len, 0, p1);
        pgo_substr(bc, line, len, p1 + 1, p2 - p1 - 1);
        if (bc->len == 5 && !memcmp(bc->p, "0_0_0", 5)) pgo_str_set(bc, "", 0);
    } else {
        size_t e = npos;
        for (size_t i = 0; i < len; ++i)
            if (line[i] == ' ' || line[i] == '\r' || line[i] == '\t' || line[i] == '\n') { e = i; break; }
        pgo_substr(name, line, len, 0, e);
        pgo_str_set(bc, "", 0);
        size_t p1 = pgo_find_bxz(line, len);
        if (p1 != npos) {
            size_t p2 = pgo_find_char(line, len, '-', p1 + 5);
            pgo_substr(bc, line, len, p1 + 5, p2 - p1 - 5);
        }
    }
}

/* exported for unit tests of the header rules */
int pgo_parse_header(int* read_type_io, const char* line, char* name_out, char* bc_out, int outcap)
{
    pgo_hdr_state st = { *read_type_io };
    pgo_str n = { 0 }, b = { 0 };
    pgo_get_barcode(&st, line, strlen(line), &n, &b);
    *read_type_io = st.read_type;
    snprintf(name_out, (size_t)outcap, "%s", n.p ? n.p : "");
    snprintf(bc_out, (size_t)outcap, "%s", b.p ? b.p : "");
    free(n.p); free(b.p);
    return 0;
}

/* ------------------------------------------------------------------ */
/* per-cloud feature rows                                               */
/* ------------------------------------------------------------------ */
typedef struct pgo_rows {
    int64_t n, dim, cap;
    char** labels;
    double* vals; /* n x dim, integer-valued like the reference's vector<double> */
} pgo_rows;

static pgo_rows* pgo_rows_new(int64_t dim)
{
    pgo_rows* r = (pgo_rows*)calloc(1, sizeof(*r));
    r->dim = dim; r->cap = 64;
    r->labels = (char**)calloc((size_t)r->cap, sizeof(char*));
    r->vals = (double*)calloc((size_t)(r->cap * dim), sizeof(double));
    return r;
}
static double* pgo_rows_push(pgo_rows* r, const char* label)
{
    if (r->n == r->cap) {
        r->cap *= 2;
        r->labels = (char**)realloc(r->labels, (size_t)r->cap * sizeof(char*));
        r->vals = (double*)realloc(r->vals, (size_t)(r->cap * r->dim) * sizeof(double));
    }
    r->labels[r->n] = strdup(label);
    double* v = r->vals + r->n * r->dim;
    memset(v, 0, (size_t)r->dim * sizeof(double));
    r->n++;
    return v;
}
void pgo_rows_free(pgo_rows* r)
{
    if (!r) return;
    for (int64_t i = 0; i < r->n; ++i) free(r->labels[i]);
    free(r->labels); free(r->vals); free(r);
}
int64_t pgo_rows_n(const pgo_rows* r) { return r->n; }
int64_t pgo_rows_dim(const pgo_rows* r) { return r->dim; }
const char* pgo_rows_label(const pgo_rows* r, int64_t i) { return r->labels[i]; }
const double* pgo_rows_vals(const pgo_rows* r) { return r->vals; }

/* a10, src/cpptools/count_tnf.cpp:54-76,138-164: the ordered map is seeded with
 * the canonical code of every k-mer, so column j is the j-th smallest canonical
 * code.  lut[code] = column, returns the column count (136 for k = 4). */
int pgo_tnf_lut(int k, int32_t* lut)
{
    int64_t n = 1LL << (2 * k);
    int32_t col = 0;
    for (int64_t v = 0; v < n; ++v) lut[v] = -1;
    for (int64_t v = 0; v < n; ++v)
        if (pgo_canonical((uint64_t)v, k) == (uint64_t)v) lut[v] = col++;
    for (int64_t v = 0; v < n; ++v)
        if (lut[v] < 0) lut[v] = lut[pgo_canonical((uint64_t)v, k)];
    return col;
}

typedef struct {
    int kind; /* 0 = abundance (count_kmer), 1 = tnf (count_tnf) */
    int k, vs, ws;
    int mlen;
    pgo_table* table;
    int32_t* lut;
    pgo_rows* rows;
} pgo_task;

/* a7 + a9 (count_kmer.cpp:55-108) and a7 + a11 (count_tnf.cpp:78-113): one call
 * per flushed cloud.  `seq` is the concatenation "read N read N ...". */
static void pgo_cloud(pgo_task* tk, const pgo_str* seq, const pgo_str* barcode)
{
    /* `reads_seq.size() <= mlen` compares size_t with int: mlen converts to
     * unsigned, so a negative -l drops everything. */
    if (barcode->len == 0 || (uint64_t)seq->len <= (uint64_t)(int64_t)tk->mlen) return;
    double* row = pgo_rows_push(tk->rows, barcode->p);
    uint64_t mask = pgo_kmask(tk->k), val = 0;
    size_t run = 0;
    for (size_t i = 0; i < seq->len; ++i) {
        unsigned char c = (unsigned char)seq->p[i];
        if (!pgo_is_base(c)) { val = 0; run = 0; continue; }
        val = ((val << 2) & mask) + pgo_code(c);
        if (++run == (size_t)tk->k) {
            --run;
            uint64_t key = pgo_canonical(val, tk->k);
            if (tk->kind == 0) {
                uint64_t cnt;
                if (pgo_table_get(tk->table, key, &cnt)) {
                    int pos = (int)(cnt / (uint64_t)(int64_t)tk->ws); /* count_kmer.cpp:90 */
                    if (pos < tk->vs) row[pos] += 1.0;
                }
            } else {
                row[tk->lut[key]] += 1.0;
            }
        }
    }
}

/* a5, interleaved: count_kmer.cpp:236-282 == count_tnf.cpp:234-291 */
static int pgo_run_interleaved(pgo_task* tk, const char* path)
{
    pgo_reader r;
    if (!pgo_open(&r, path)) return 0; /* igzstream on a missing file: empty, exit 0 */
    pgo_hdr_state st = { 0 };
    pgo_str name = { 0 }, bc = { 0 }, last = { 0 }, seq = { 0 };
    pgo_str_set(&bc, "", 0); pgo_str_set(&last, "", 0); pgo_str_set(&seq, "", 0);
    uint64_t n = 0;
    while (pgo_getline(&r)) {
        switch (++n % 8) {
        case 1: pgo_get_barcode(&st, r.buf, r.len, &name, &bc); break;
        case 2: pgo_str_append(&seq, r.buf, r.len); pgo_str_append(&seq, "N", 1); break;
        case 6:
            pgo_str_append(&seq, r.buf, r.len); pgo_str_append(&seq, "N", 1);
            if (!pgo_str_eq(&bc, &last)) {
                pgo_cloud(tk, &seq, &last);       /* the triggering pair stays with the OLD label */
                pgo_str_set(&last, bc.p, bc.len); /* std::move leaves `barcode` to be overwritten at the next header */
                pgo_str_set(&seq, "", 0);
            }
            break;
        default: break;
        }
    }
    pgo_cloud(tk, &seq, &last);
    pgo_close(&r);
    free(name.p); free(bc.p); free(last.p); free(seq.p);
    return 0;
}

/* a6, paired: count_kmer.cpp:181-233 == count_tnf.cpp:170-231 */
static int pgo_run_paired(pgo_task* tk, const char* path1, const char* path2)
{
    pgo_reader r1, r2;
    int ok1 = pgo_open(&r1, path1), ok2 = pgo_open(&r2, path2);
    pgo_hdr_state st = { 0 };
    pgo_str n1 = { 0 }, b1 = { 0 }, n2 = { 0 }, b2 = { 0 }, last = { 0 }, seq = { 0 };
    pgo_str_set(&n1, "", 0); pgo_str_set(&b1, "", 0); pgo_str_set(&n2, "", 0); pgo_str_set(&b2, "", 0);
    pgo_str_set(&last, "", 0); pgo_str_set(&seq, "", 0);
    uint64_t n = 0;
    while (ok1 && pgo_getline(&r1)) {
        if (!ok2 || !pgo_getline(&r2)) r2.len = 0; /* R2 shorter than R1: treated as an empty line */
        const char* l2 = r2.len ? r2.buf : "";
        switch (++n % 4) {
        case 1:
            pgo_get_barcode(&st, r1.buf, r1.len, &n1, &b1);
            pgo_get_barcode(&st, l2, r2.len, &n2, &b2);
            break;
        case 2:
            if (pgo_str_eq(&n1, &n2) && pgo_str_eq(&b1, &b2)) {
                pgo_str_append(&seq, r1.buf, r1.len); pgo_str_append(&seq, "N", 1);
                pgo_str_append(&seq, l2, r2.len); pgo_str_append(&seq, "N", 1);
                if (!pgo_str_eq(&b1, &last)) {
                    pgo_cloud(tk, &seq, &last);
                    pgo_str_set(&last, b1.p, b1.len);
                    pgo_str_set(&b1, "", 0); /* std::move(p1.second): moved-from string is empty in libstdc++ */
                    pgo_str_set(&seq, "", 0);
                }
            }
            break;
        default: break;
        }
    }
    pgo_cloud(tk, &seq, &last);
    if (ok1) pgo_close(&r1);
    if (ok2) pgo_close(&r2);
    free(n1.p); free(b1.p); free(n2.p); free(b2.p); free(last.p); free(seq.p);
    return 0;
}

/* path2 == NULL or "" selects interleaved mode */
pgo_rows* pgo_abundance(const char* path1, const char* path2, pgo_table* table, int k, int mlen, int vs, int ws)
{
    pgo_task tk = { 0, k, vs, ws, mlen, table, NULL, pgo_rows_new(vs) };
    if (path2 && path2[0]) pgo_run_paired(&tk, path1, path2); else pgo_run_interleaved(&tk, path1);
    return tk.rows;
}

pgo_rows* pgo_tnf(const char* path1, const char* path2, int k, int mlen)
{
    int32_t* lut = (int32_t*)malloc(sizeof(int32_t) << (2 * k));
    int dim = pgo_tnf_lut(k, lut);
    pgo_task tk = { 1, k, 0, 0, mlen, NULL, lut, pgo_rows_new(dim) };
    if (path2 && path2[0]) pgo_run_paired(&tk, path1, path2); else pgo_run_interleaved(&tk, path1);
    free(lut);
    return tk.rows;
}
