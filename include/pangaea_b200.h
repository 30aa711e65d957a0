/*
 * pangaea_b200.h - C-ABI of the B200-native read-cloud featurization path.
 *
 * This is the drop-in boundary for Pangaea's step 1 ("feature extraction").  The
 * reference has no FFI on this path: src/feature.py drives three subprocesses
 * (jellyfish, bin/count_kmer, bin/count_tnf) and reads their gz-CSV files back with
 * pandas.  The entry points below are what a maintainer binds (ctypes stub in
 * INTEGRATION.md) in place of those subprocess calls; each one names the reference
 * interface it replaces.  Citations are relative to /root/reference/.
 *
 * Conventions
 *   - plain C types only; every function returns PG_OK (0) or a negative pg_status
 *     and leaves a message for pg_last_error();
 *   - a pg_ctx owns one CUDA device, one stream, the k-mer count table and all
 *     scratch; it has no global mutable state (the reference's `read_type` global,
 *     count_kmer.cpp:24, lives in the parser object instead) and may be driven
 *     from any one thread at a time with the GIL released;
 *   - there is NO CPU fallback: without a CUDA device pg_create fails.
 *
 * Read batches.  A batch is the sequence lines of consecutive FASTQ records, each
 * followed by ONE separator byte (any non-ACGT byte; the parser uses '\n'), i.e.
 * exactly the string the reference builds per cloud with `reads_seq += line + "N"`
 * (count_kmer.cpp:247-250, :199).  read_off[r+1]-read_off[r] = len(read r)+1, so
 * the reference's min-length rule `reads_seq.size() <= mlen` (count_kmer.cpp:62) is
 * a difference of offsets.  read_flag[r] carries the grouping decisions the
 * reference's main() loop takes while streaming (count_kmer.cpp:181-282):
 *   PG_READ_CHANGE  the barcode compared unequal to last_barcode when this read
 *                   was appended: the cloud is flushed AFTER this read (set on the
 *                   R2 read of the triggering pair - the reference's off-by-one);
 *   PG_READ_NOFEAT  read is counted (jellyfish sees every record) but belongs to
 *                   no cloud (R1/R2 name or barcode mismatch, count_kmer.cpp:195).
 * Cloud g = number of PG_READ_CHANGE flags on earlier reads; its label is the
 * barcode of the pair that carried the g-th flag ("" for g = 0).  group_keep[g]
 * is 0 when that label is empty (count_kmer.cpp:62 `barcode.empty()`).
 */
#ifndef PANGAEA_B200_H
#define PANGAEA_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum pg_status {
    PG_OK = 0,
    PG_ERR_INVALID = -1,  /* bad argument */
    PG_ERR_CUDA = -2,     /* CUDA runtime error (no device, OOM, launch failure) */
    PG_ERR_CAPACITY = -3, /* hash table full - raise table_capacity */
    PG_ERR_IO = -4,       /* file could not be opened / read */
    PG_ERR_STATE = -5     /* call out of order (e.g. featurize before count) */
} pg_status;

enum { PG_READ_CHANGE = 1, PG_READ_NOFEAT = 2 };
enum { PG_TABLE_AUTO = 0, PG_TABLE_DENSE = 1, PG_TABLE_HASH = 2, PG_TABLE_NONE = 3 /* no table: a ctx that only normalises (Data.__init__) */ };

/* Replaces the CLI flags of count_kmer (count_kmer.cpp:112-123) / count_tnf
 * (count_tnf.cpp:117-125) as passed by feature.py:107-109,131-133, and the
 * jellyfish flags at feature.py:76-94. */
typedef struct pg_params {
    int32_t device;          /* CUDA ordinal */
    int32_t k;               /* -k, 15: k-mer size of the abundance feature, 1..31 */
    int32_t tnf_k;           /* count_tnf -k, 4: 1..6 */
    int32_t window_size;     /* -w, 10: abundance bin width (count / w) */
    int32_t vector_size;     /* -v, 400: abundance bins */
    int64_t min_length;      /* -l, 2000: clouds with sum(len+1) <= this are dropped */
    int32_t min_qual_char;   /* jellyfish --min-qual-char (0 = off; '?' in the -1/-2 branch) */
    int32_t table_mode;      /* PG_TABLE_*: dense direct-addressed (k <= 16) or open-addressing hash.  Dense counters saturate at
                              * 2^31 - 1 (jellyfish would report the true count; count_kmer.cpp:90-92 drops either), which is why
                              * window_size * vector_size must be <= 2^31 - 1 */
    uint64_t table_capacity; /* hash slots (power of two; 0 = sized from the first batch) */
} pg_params;

typedef struct pg_reads {
    const uint8_t* seq;       /* n_bytes: read bytes + 1 separator per read */
    const uint8_t* qual;      /* same layout, or NULL (only read when min_qual_char != 0) */
    const int64_t* read_off;  /* n_reads + 1 offsets into seq, read_off[0] = 0 */
    const uint8_t* read_flag; /* n_reads: PG_READ_* bits */
    int64_t n_reads;
    int64_t n_bytes;          /* = read_off[n_reads] */
} pg_reads;

typedef struct pg_ctx pg_ctx;
typedef struct pg_batch pg_batch;       /* a read batch resident in HBM (ASCII + 2-bit packed) */
typedef struct pg_features pg_features; /* per-cloud matrices resident in HBM */

/* ---- lifetime -------------------------------------------------------------- */
void pg_default_params(pg_params* p);
int pg_create(const pg_params* p, pg_ctx** out);
void pg_destroy(pg_ctx* ctx);
/* message of the last failure on ctx (ctx may be NULL: failure of pg_create / parser) */
const char* pg_last_error(const pg_ctx* ctx);
int pg_device_count(void);
/* tnf_k -> number of canonical columns (136 for 4); count_tnf.cpp:138-164 */
int pg_tnf_dim(int tnf_k);
int pg_synchronize(pg_ctx* ctx);
/* free / total device memory as the driver reports it, plus what this ctx holds idle in its caches (reusable by it) */
int pg_mem_info(pg_ctx* ctx, int64_t* free_bytes, int64_t* total_bytes);
/* give the idle device memory this ctx caches for reuse (freed matrices, partitions, packed streams: up to 48 GB) back to
 * the driver - e.g. before the VAE of src/pangaea.py:91 starts on the same GPU.  Live batches, feature sets and the k-mer
 * table are untouched. */
int pg_trim(pg_ctx* ctx);
/* the cudaStream_t every launch of this ctx goes to (for CUDA-event timing by the caller) */
void* pg_stream(pg_ctx* ctx);

/* ---- batches --------------------------------------------------------------- */
/* copy a host batch to HBM (async on the ctx stream; pinned memory overlaps) and 2-bit pack it */
int pg_batch_upload(pg_ctx* ctx, const pg_reads* host, pg_batch** out);
/* adopt device pointers without copying (pointers must stay valid; seq 16-byte aligned) and pack */
int pg_batch_adopt(pg_ctx* ctx, const pg_reads* dev, pg_batch** out);
void pg_batch_free(pg_ctx* ctx, pg_batch* b);
/* one batch of a stream: pg_batch_upload + pg_count2 with the copy overlapped chunk by chunk; the table is NOT cleared */
int pg_batch_upload_count(pg_ctx* ctx, const pg_reads* host, int keep_partition, pg_batch** out);
/* release the ASCII bases: the 2-bit stream, masks and read offsets - 0.5 B per base + 9 B per read - stay, which is
 * all pg_featurize needs (a partition kept by pg_count2 stays too: ~4.5 B per k-mer window).  Lets the batches of a file
 * larger than HBM's ASCII capacity wait on the device for the featurize pass. */
int pg_batch_compact(pg_ctx* ctx, pg_batch* b);
int64_t pg_batch_n_groups(const pg_batch* b); /* 1 + number of PG_READ_CHANGE flags */
/* shape of a batch and its arrays back on the host (tests, debugging; a compacted batch has no bases left: seq_out must be NULL) */
void pg_batch_shape(const pg_batch* b, int64_t* n_reads, int64_t* n_bytes);
int pg_batch_download(pg_ctx* ctx, const pg_batch* b, uint8_t* seq_out, int64_t* read_off_out, uint8_t* read_flag_out);

/* ---- step 1a: global canonical k-mer counts --------------------------------- */
/* replaces `jellyfish count -C -m k` (+ `dump`), feature.py:76-94,103, and the dump
 * loader count_kmer.cpp:139-170 (the table simply stays in HBM).  Adds the batch.
 * Dense mode, k <= 15: the partition of the k-mer windows by table slice that this pass computes is the one
 * pg_featurize needs too, so - memory permitting - it is KEPT in the batch (about 0.75 KB per read pair of 2x100 bp,
 * from a buffer the ctx reuses) until pg_batch_free; pg_featurize of the same batch then skips its own partition. */
int pg_count(pg_ctx* ctx, pg_batch* b);
/* keep_partition = 0: do not keep the partition (a stream of batches is counted first and featurized later) */
int pg_count2(pg_ctx* ctx, pg_batch* b, int keep_partition);
int pg_table_clear(pg_ctx* ctx);
/* kmer2frequency[key] = count (count_kmer.cpp:166): assignment, keys in the reference's
 * canonical form or not (re-canonicalised).  A key set with count 0 is PRESENT with
 * frequency 0 (lands in bin 0), as in the reference; pg_table_get cannot tell it from absent. */
int pg_table_set(pg_ctx* ctx, const uint64_t* keys, const uint32_t* counts, int64_t n);
int pg_table_get(pg_ctx* ctx, const uint64_t* keys, uint32_t* counts_out, int64_t n);
/* number of distinct k-mers; then export (key = reference canonical form, ascending) */
int pg_table_size(pg_ctx* ctx, int64_t* n_distinct);
int pg_table_export(pg_ctx* ctx, uint64_t* keys_out, uint32_t* counts_out, int64_t cap, int64_t* n_out);
/* dense mode only: device pointer and entry count of the u32 counter array, so a
 * data-parallel caller can sum tables across ranks (ncclAllReduce) - SURVEY §8e */
int pg_table_dense_view(pg_ctx* ctx, void** dev_ptr, int64_t* n_entries);
/* the caller is still writing the table on another stream (the all-reduce): every later launch of this ctx that reads
 * or writes the table first waits for `cuda_event` (a cudaEvent_t recorded after that work).  The parts of
 * pg_featurize that do not need the table (cloud grouping, TNF) run ahead of it - they overlap the collective. */
int pg_table_wait_event(pg_ctx* ctx, void* cuda_event);
/* dense mode only: counter = min(counter, max_count).  A data-parallel caller clamps every rank's table to
 * (2^31 - 1) / n_ranks before the int32 sum so that the sum cannot wrap; exact as long as
 * window_size * vector_size <= max_count (a clamped count is dropped from the histogram like the true one). */
int pg_table_clamp(pg_ctx* ctx, uint32_t max_count);

/* ---- step 1b: per-cloud abundance histogram + TNF --------------------------- */
/* replaces bin/count_kmer (countKmer, count_kmer.cpp:55-108) and bin/count_tnf
 * (countKmer, count_tnf.cpp:78-113) in one pass over the packed bases.  group_keep:
 * n_groups bytes (host memory), see header comment.  Rows come out in file order. */
int pg_featurize(pg_ctx* ctx, pg_batch* b, const uint8_t* group_keep, int64_t n_groups, pg_features** out);
/* flags = PG_FEAT_NO_ABUNDANCE: cloud grouping, rows and TNF only; the abundance matrix stays zero for pg_features_add_counts
 * (the owner-partitioned multi-GPU table, below: the counts come from other ranks) */
enum { PG_FEAT_NO_ABUNDANCE = 1 };
int pg_featurize2(pg_ctx* ctx, pg_batch* b, const uint8_t* group_keep, int64_t n_groups, int flags, pg_features** out);
void pg_features_free(pg_ctx* ctx, pg_features* f);
int64_t pg_features_rows(const pg_features* f);
int32_t pg_features_abd_dim(const pg_features* f);
int32_t pg_features_tnf_dim(const pg_features* f);
/* cloud index (into the caller's label list) of every emitted row */
int pg_features_row_groups(pg_ctx* ctx, const pg_features* f, int64_t* groups_out);
/* raw tallies, int32 row-major [rows, dim] - what the CSV files held as text */
int pg_features_copy_raw(pg_ctx* ctx, const pg_features* f, int32_t* abd_out, int32_t* tnf_out);

/* ---- owner-partitioned table across ranks (owner = hash(canonical k-mer) mod n_ranks) ------------------
 * The multi-GPU form of the HASH table (k > 16), which has no dense view to all-reduce.  Device side of the two all-to-alls;
 * the collective itself belongs to the caller (torch.distributed / NCCL - pangaea_b200/distributed.py):
 *   count     keys = pg_batch_window_keys(count windows) -> pg_keys_partition -> all-to-all of keys -> pg_table_add_keys
 *   featurize f = pg_featurize2(NO_ABUNDANCE); keys(feature windows) -> pg_keys_partition -> all-to-all of keys ->
 *             pg_table_lookup_keys at the owner -> all-to-all of counts back -> pg_features_add_counts
 * All d_* arguments are device memory.  Keys are the canonical forms of the reference (count_kmer.cpp:86), ~0 = no window. */
int64_t pg_batch_n_words(const pg_batch* b);
int pg_batch_window_keys(pg_ctx* ctx, pg_batch* b, int64_t w0, int64_t w1, int feature_windows, uint64_t* d_keys /* 32 x (w1 - w0) */);
int pg_keys_partition(pg_ctx* ctx, const uint64_t* d_keys, int64_t n, int32_t world, uint64_t* d_sorted, int64_t* d_dest, int64_t* counts_out);
int pg_table_add_keys(pg_ctx* ctx, const uint64_t* d_keys, int64_t n);
int pg_table_lookup_keys(pg_ctx* ctx, const uint64_t* d_keys, int64_t n, uint32_t* d_counts);
int pg_features_add_counts(pg_ctx* ctx, pg_features* f, pg_batch* b, int64_t w0, int64_t w1, const int64_t* d_dest, const uint32_t* d_counts_sorted);

/* ---- step 2 prologue: Data.__init__ (src/data.py:16-21) --------------------- */
/* L1-normalise both matrices (fp64 divide, fp32 store) and weights = max(abd row)^2 (fp64).  The reference normalises
 * what pandas read back from the tools' CSV text, and `ostream << double` keeps 6 significant digits (count_kmer.cpp:211,
 * count_tnf.cpp:204): tallies >= 10^6 are rounded the same way (ties to even) before the division.  The raw tallies
 * (pg_features_copy_raw, DLPack 0/1) stay exact. */
int pg_normalize(pg_ctx* ctx, pg_features* f);
int pg_features_copy_normalized(pg_ctx* ctx, const pg_features* f, float* abd_out, float* tnf_out, double* weights_out);

/* wrap caller-provided raw tallies (host memory, row-major) so pg_normalize can run on them:
 * the `Data(barcodes, abd, tnf)` entry when the matrices come from load_features() */
int pg_features_from_raw(pg_ctx* ctx, const uint32_t* abd, const uint32_t* tnf, int64_t rows, int32_t abd_dim, int32_t tnf_dim, pg_features** out);

/* rows of consecutive batches as one feature set (device-to-device copies; the parts stay valid).  Its row_groups are
 * the parts' batch-local cloud indices. */
int pg_features_concat(pg_ctx* ctx, pg_features* const* parts, int32_t n_parts, pg_features** out);

/* zero-copy hand-off: which = 0 abd_raw(i32) 1 tnf_raw(i32) 2 abd(f32) 3 tnf(f32) 4 weights(f64).
 * Returns a DLManagedTensor* (DLPack v0.8 layout) whose deleter releases a reference
 * on the feature set; wrap it in a PyCapsule named "dltensor". */
void* pg_features_dlpack(pg_ctx* ctx, pg_features* f, int which);
/* raw device pointer of the same buffers (valid until pg_features_free) */
void* pg_features_device_ptr(const pg_features* f, int which);

/* ---- whole path with HOST buffers (the e2e call) ---------------------------- */
/* upload + count + featurize + normalize; the table is cleared first. */
int pg_extract_features(pg_ctx* ctx, const pg_reads* host, const uint8_t* group_keep, int64_t n_groups, pg_features** out);

/* ---- host FASTQ reader (replaces the getline loops, count_kmer.cpp:181-282,
 *      and getBarcode, count_kmer.cpp:25-53) ------------------------------------ */
typedef struct pg_fastq pg_fastq;
enum { PG_FQ_QUAL = 1,    /* keep the quality lines (needed when pg_params.min_qual_char != 0) */
       PG_FQ_PINNED = 2   /* batch buffers in page-locked host memory (create the pg_ctx first: this touches the CUDA device) */ };
/* the whole input as ONE batch.  path2 == NULL: interleaved (-i), else paired (-1/-2).  Plain text or gzip. */
int pg_fastq_parse(const char* path1, const char* path2, int flags, pg_fastq** out);
void pg_fastq_free(pg_fastq* fq);
void pg_fastq_reads(const pg_fastq* fq, pg_reads* out);      /* host pointers owned by fq */
int64_t pg_fastq_n_groups(const pg_fastq* fq);
const uint8_t* pg_fastq_group_keep(const pg_fastq* fq);       /* n_groups bytes */
const char* pg_fastq_group_label(const pg_fastq* fq, int64_t g);
/* all labels at once: concatenated into buf (no terminators), offsets[n_groups + 1].  Returns the bytes needed; nothing
 * is written when cap is smaller (call with buf = NULL to size). */
int64_t pg_fastq_group_labels(const pg_fastq* fq, char* buf, int64_t cap, int64_t* offsets);

/* The same reader as a STREAM of batches - the reference handles files of any size one cloud at a time
 * (count_kmer.cpp:236-282); here the unit is a batch of about target_seq_bytes that ends where the reference would
 * flush a cloud (or inside a cloud labelled "", which is dropped whole).  read_type and last_barcode are carried from
 * batch to batch; label 0 of a batch is the label of the cloud that is open when the batch starts.  Feed every batch
 * to pg_count, then featurize batch by batch: rows of consecutive batches concatenate to the reference's row order.
 * byte_lo / byte_hi (plain-text interleaved files only; 0 / -1 = the whole file): one rank of a multi-GPU run reads
 * the clouds that START in its byte range; lines_before_lo = number of '\n' in [0, byte_lo) (pg_fastq_count_lines,
 * summed over the lower ranks) - record boundaries are line numbers 0 mod 8. */
typedef struct pg_fastq_stream pg_fastq_stream;
int pg_parallel_memcpy(void* dst, const void* src, int64_t n); /* all host cores */
int pg_parallel_pread(const char* path, int64_t offset, int64_t n, void* dst); /* all host cores: file -> (pinned) staging buffer */
int pg_fastq_count_lines(const char* path, int64_t byte_lo, int64_t byte_hi, int64_t* n_newlines);
int pg_fastq_stream_open(const char* path1, const char* path2, int flags, int64_t byte_lo, int64_t byte_hi, int64_t lines_before_lo,
                         pg_fastq_stream** out);
/* *out = NULL at the end of the stream; target_seq_bytes <= 0: everything that is left */
int pg_fastq_stream_next(pg_fastq_stream* s, int64_t target_seq_bytes, pg_fastq** out);
void pg_fastq_stream_close(pg_fastq_stream* s);

/* ---- FASTQ text on the device -------------------------------- */
/* Replaces the host loop count_kmer.cpp:236-282 + getBarcode (count_kmer.cpp:25-53) for plain-text INTERLEAVED input: the
 * text goes to HBM as it is, a line index is built there, headers are parsed there (read type latched by the first
 * decisive header, "0_0_0", npos wrap-around and all), and the packed batch comes out without the bases ever being
 * touched by the host.  text: a chunk that starts at a record boundary (line number 0 mod 8).  The batch ends at the
 * last cloud flush inside the chunk (or at its end when the open cloud is labelled "" or PG_INGEST_FINAL is set);
 * *consumed = bytes of `text` the batch covers - the caller starts the next chunk there.  consumed = 0 and *batch = NULL:
 * the chunk holds no flush, pass a larger one.  last_barcode / *read_type_io carry the reference's two pieces of
 * sequential state from chunk to chunk ("" and 0 at the start of a file; afterwards: the last label of the previous
 * chunk and the value left in *read_type_io). */
enum { PG_INGEST_FINAL = 1, PG_INGEST_DEVICE_TEXT = 2 };
typedef struct pg_ingest pg_ingest; /* labels / keep flags of the clouds of one ingested batch; label 0 = last_barcode */
int pg_ingest_text(pg_ctx* ctx, const void* text, int64_t n_bytes, int flags, const char* last_barcode, int64_t last_len,
                   int32_t* read_type_io, int64_t* consumed, pg_batch** batch, pg_ingest** info);
int64_t pg_ingest_n_groups(const pg_ingest* info);
const uint8_t* pg_ingest_group_keep(const pg_ingest* info);
int64_t pg_ingest_group_labels(const pg_ingest* info, char* buf, int64_t cap, int64_t* offsets); /* as pg_fastq_group_labels */
void pg_ingest_free(pg_ingest* info);
/* Replaces `awk ... | LANG=C sort -k1,1 | cut -f2- | tr "\t" "\n"` (src/run_pangaea:237-252): the interleaved FASTQ in
 * `in` sorted by its BX:Z: tag (untagged pairs last, "~~~"), ties broken by the rest of the record's text exactly as
 * GNU sort's last-resort comparison does, tabs turned into newlines as the script's `tr` does.  Radix sort on the
 * device; *n_out = bytes needed (nothing is written when out_cap is smaller).  Input must be whole 8-line records whose
 * first line starts with '@' (PG_ERR_INVALID otherwise). */
int pg_fastq_sort_by_barcode(pg_ctx* ctx, const char* in, int64_t n_in, char* out, int64_t out_cap, int64_t* n_out);

/* ---- format converters and extract_reads on the device ---- */
/* preprocess_stlfr -n [-l] (src/cpptools/preprocess_stlfr.cpp:76-115; run_pangaea:143 passes -n -l): headers `name#a_b_c/1`
 * become `name\tBX:Z:a_b_c[-1]` (`name` alone when a or b is "0"), the identifier of file 1 going into BOTH outputs; all other
 * lines are copied.  r1 / r2: the two FASTQ texts (host memory).  *n_out = bytes needed; PG_ERR_INVALID when a buffer is too
 * small or when the reference tool would abort (no '#', barcode not a_b_c).  The whitelist mode (without -n) is not here. */
int pg_preprocess_stlfr(pg_ctx* ctx, const char* r1, int64_t n1, const char* r2, int64_t n2, int library, char* out1, int64_t cap1,
                        int64_t* n_out1, char* out2, int64_t cap2, int64_t* n_out2);
/* preprocess_tellseq (src/cpptools/preprocess_tellseq.cpp:52-84): idx = the index reads (I1); records whose index read is not 18
 * bases are dropped; header = R1 header up to the first ' ' + "\tBX:Z:" + index read + "-1" in both outputs, the third line
 * always "+"; out_wl = the barcodes, one per line. */
int pg_preprocess_tellseq(pg_ctx* ctx, const char* r1, int64_t n1, const char* r2, int64_t n2, const char* idx, int64_t ni, char* out1,
                          int64_t cap1, int64_t* n_out1, char* out2, int64_t cap2, int64_t* n_out2, char* out_wl, int64_t cap_wl,
                          int64_t* n_out_wl);
/* extract_reads -i (src/cpptools/extract_reads.cpp:86-124): open parses the interleaved text on the device and reports its
 * barcode runs (consecutive pairs with the same barcode; run 0 = pairs without barcode at the start); the caller maps every
 * run label to a cluster index or -1 (the reference's barcode2cluster, extract_reads.cpp:58-84); route writes, per cluster
 * and in file order, the .fq blob (header rewritten to name + "\tBX:Z:" + barcode + "-1") and the .barcode blob, and returns
 * the n_clusters + 1 byte offsets of the clusters' slices; copy brings the two blobs to the host. */
typedef struct pg_extract pg_extract;
int pg_extract_open(pg_ctx* ctx, const char* text, int64_t n_bytes, pg_extract** out);
int64_t pg_extract_n_runs(const pg_extract* x);
int64_t pg_extract_run_labels(const pg_extract* x, char* buf, int64_t cap, int64_t* offsets);
int pg_extract_route(pg_ctx* ctx, pg_extract* x, const int32_t* cluster_of_run, int32_t n_clusters, int64_t* fq_start, int64_t* bc_start);
int pg_extract_copy(pg_ctx* ctx, const pg_extract* x, char* fq_out, char* bc_out);
void pg_extract_close(pg_ctx* ctx, pg_extract* x);

/* ---- step-2 input pipeline on the device ---------------------- */
/* Replaces CustomWeightedRandomSampler (src/utils.py:11-23: numpy.random.choice(range(N), size, p = w / sum(w), replace)) and the
 * DataLoader's row gather (src/data.py:27-31, src/pangaea.py:87-89).  The uniforms are the caller's (numpy's generator, so
 * a seeded run draws what the reference draws); the cumulative sum (sequential, as numpy's), the binary searches, the
 * first-occurrence filter of the draws without replacement and the gather of the batch rows run on the device.
 * weights: host or device memory, n doubles; total = their sum as the reference computes it (torch.sum). */
typedef struct pg_sampler pg_sampler;
int pg_sampler_create(pg_ctx* ctx, const double* weights, int64_t n, double total, pg_sampler** out);
void pg_sampler_free(pg_ctx* ctx, pg_sampler* s);
/* replace = True: d_idx_out[j] (device, m x int64) = cdf.searchsorted(uniforms[j], side = "right"); uniforms in host memory */
int pg_sampler_draw(pg_ctx* ctx, pg_sampler* s, const double* uniforms, int64_t m, int64_t* d_idx_out);
/* replace = False, one round of numpy's loop: m = size - n_found fresh uniforms; the draws that are first occurrences are
 * appended at d_idx_out[n_found ...] in draw order and their probabilities zeroed; *n_new = how many.  Loop until done. */
int pg_sampler_draw_unique_round(pg_ctx* ctx, pg_sampler* s, const double* uniforms, int64_t m, int64_t n_found, int64_t* d_idx_out,
                                 int64_t* n_new);
/* rows d_idx[0..m) of the normalised matrices into device buffers [m, abd_dim] / [m, tnf_dim] (either may be NULL) */
int pg_features_gather(pg_ctx* ctx, const pg_features* f, const int64_t* d_idx, int64_t m, float* d_abd_out, float* d_tnf_out);

/* ---- synthetic reads on device (bench input) --------------------- */
/* fills DEVICE buffers shaped like a pg_reads batch: n_pairs pairs of 2 x read_len,
 * (read_len + 1) bytes per read.  d_bc_start[n_barcodes + 1] = first pair of every
 * barcode, d_bc_genome[n_barcodes] = its genome (both device memory). */
int pg_synth_generate(pg_ctx* ctx, int64_t n_pairs, int32_t read_len, int64_t n_barcodes, const int64_t* d_bc_start,
                      const int32_t* d_bc_genome, int64_t genome_len, int32_t frag_len, int32_t insert, double sub_rate,
                      double n_rate, uint64_t seed, uint8_t* d_seq, int64_t* d_read_off, uint8_t* d_read_flag);

/* the same with a global index for the batch's first barcode and first pair: batches / ranks that share `seed` draw
 * different clouds from the same community of genomes */
int pg_synth_generate2(pg_ctx* ctx, int64_t n_pairs, int32_t read_len, int64_t n_barcodes, const int64_t* d_bc_start,
                       const int32_t* d_bc_genome, int64_t genome_len, int32_t frag_len, int32_t insert, double sub_rate,
                       double n_rate, uint64_t seed, int64_t bc_base, int64_t pair_base, uint8_t* d_seq, int64_t* d_read_off,
                       uint8_t* d_read_flag);

/* ---- instrumentation -------------------------------------------------------- */
/* device time (ms, CUDA events on the ctx stream) and launch count of the kernels run
 * since the last reset: which = 0 pack, 1 count (apply / direct), 2 group, 3 featurize (apply / direct),
 * 4 normalize, 5 all, 6 count scatter, 7 featurize scatter, 8 TNF, 9 count split (second partition level),
 * 10 featurize collect (collect.cuh: tallies from the looked-up entries, in stream order) */
int pg_timing_reset(pg_ctx* ctx);
int pg_timing_get(pg_ctx* ctx, int which, double* ms_out, int64_t* launches_out);

#ifdef __cplusplus
}
#endif
#endif /* PANGAEA_B200_H */
