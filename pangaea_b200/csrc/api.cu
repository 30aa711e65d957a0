// api.cu - C-ABI (include/pangaea_b200.h) over the sm_100a kernels.
// Host orchestration only: allocation, launches, read-backs.  No CPU compute path.
#include <algorithm>
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <new>
#include <unordered_map>
#include <string>
#include <vector>

#include <cuda_runtime.h>

#include "../../include/pangaea_b200.h"
#include "bucket.cuh"
#include "collect.cuh"
#include "count.cuh"
#include "count2.cuh"
#include "exchange.cuh"
#include "featurize.cuh"
#include "kmer.cuh"
#include "normalize.cuh"
#include "pack.cuh"
#include "scan.cuh"
#include "synth.cuh"
#include "table.cuh"
#include "tnf.cuh"

using namespace pg;

// ---------------------------------------------------------------------------
// objects
// ---------------------------------------------------------------------------
// timing slots: one per kernel family (pg_timing_get `which`)
enum { T_PACK = 0, T_COUNT = 1, T_GROUP = 2, T_FEAT = 3, T_NORM = 4, T_ALL = 5, T_COUNT_SCATTER = 6, T_FEAT_SCATTER = 7, T_TNF = 8, T_COUNT_SPLIT = 9, T_COLLECT = 10, T_SLOTS = 11 };

// grow-only device scratch kept by the ctx: the multi-GB partition buffers are allocated once, not once per step
// (72 GB of cudaMallocAsync / cudaFreeAsync per step made the pool re-map memory: step times of 106 .. 380 ms)
struct Workspace {
    void* p = nullptr;
    size_t bytes = 0;
};

// Block cache for allocations >= 1 MiB (packed stream, cloud maps, feature matrices).  Steps repeat the same sizes,
// so after the first step every request is served from the free list without a driver call.  cudaMallocAsync was
// measured to fall back to a fresh mapping now and then (100+ ms per GB, host and stream both stalled) when 10+ GB per
// step went through the pool.  All work of a ctx is on one stream, so a freed block may be handed out again at once.
struct BigCache {
    std::unordered_map<void*, size_t> live;   // block -> capacity
    std::multimap<size_t, void*> free_blocks; // capacity -> block
    size_t free_bytes = 0;
};

struct pg_ctx {
    pg_params p;
    int td = 0;
    int mode = kDense;
    cudaStream_t stream = nullptr;
    cudaStream_t copy_stream = nullptr; // H2D of the pipelined upload (pg_extract_features)
    int sm_count = 148;
    // table
    uint32_t* counts = nullptr;
    HashSlot* slots = nullptr; // hash mode
    uint64_t n_slots = 0;
    bool no_table = false;     // PG_TABLE_NONE: a ctx that only normalises (Data.__init__) - nothing that needs the table works
    bool have_table() const { return counts != nullptr || slots != nullptr; }
    uint32_t* d_overflow = nullptr;
    uint32_t* d_sat = nullptr;      // raised by a direct add that takes a dense counter to bit 31 (table.cuh: saturation)
    bool zero_markers = false;      // pg_table_set stored a "present with count 0" marker: counting on top of it is refused
    cudaMemPool_t mempool = nullptr; // private pool of the small stream-ordered allocations (the default pool is left alone)
    bool counted = false;
    cudaEvent_t table_event = nullptr; // pending external write to the table (pg_table_wait_event); not owned
    // TNF look-up table
    uint16_t* d_lut = nullptr;
    // pinned read-back scratch
    int64_t* h_pin = nullptr;
    int64_t* d_scalar = nullptr; // 8 x int64 device scalars
    Workspace ws_entries, ws_entries2, ws_feat; // level-1 / level-2 entries of the count pass, entries of an unshared featurize pass
    BigCache big;
    Workspace ws_text[2];                       // device staging of the text chunks of pg_ingest_text: filled on the copy stream while
    int text_toggle = 0;                        // the compute stream still works on the batch before (two, used alternately)
    std::vector<Workspace> stash_pool;          // idle shared-partition buffers (one is taken by a batch in pg_count and returned by pg_batch_free;
                                                // the batches of a stream that keep their partitions hold one each)
    BucketState* d_bucket = nullptr; // cursors / limits / ticket of the L2-sliced path
    double region_slack = 1.5;       // region capacity = slack x mean entries per slice (PG_REGION_SLACK overrides; tests force overflow)
    bool force_direct = false;   // PG_FORCE_DIRECT=1: never use the L2-sliced path (A/B measurements)
    int tnf_overlap = 2;         // PG_TNF_OVERLAP=n: the TNF kernel runs on the second stream with n CTAs per SM, next to the look-up
                                 // sweep (0 = in line on the ctx stream)
    bool no_shared = false;      // PG_NO_SHARED=1: separate partitions for the count and featurize passes (A/B, tests)
    int feat_path = -1;          // featurize pass of the sliced path: 1 = one sweep (gather + MATCH + RED per entry), 0 = lookup + collect
                                 // (collect.cuh), -1 = by cloud size (PG_FEAT_APPLY=0/1 forces one of them)
    bool count_l2 = false;       // PG_COUNT_L2=1: apply the count entries with L2 atomics instead of the shared-memory sub-slices (A/B)
    int64_t count_seg_words = 1ll << 26; // segment of the count pass (2^31 windows: 13 GB of u32 + 6 GB of u16 entries); PG_SEG_WORDS overrides
    int64_t seg_words = 1ll << 26; // words per segment of the featurize pass: 2^31 windows -> 14 GB of u32 entries with the default
                                   // slack; PG_SEG_WORDS / PG_FEAT_SEG_WORDS override (tests force many segments on small inputs)
    std::string err;
    // timing
    struct Span { int which; cudaEvent_t a, b; };
    std::vector<Span> spans;
    std::vector<cudaEvent_t> pool;
    int64_t launches[T_SLOTS] = { 0 };
};

struct pg_batch {
    uint8_t *seq = nullptr, *qual = nullptr, *read_flag = nullptr;
    int64_t* read_off = nullptr;
    bool owns = false;
    bool owns_meta = false; // read_off / read_flag are the batch's own copies although the bases were adopted (pg_batch_compact)
    int64_t n_reads = 0, n_bytes = 0, n_words = 0;
    uint64_t* codes = nullptr;
    uint32_t *maskF = nullptr, *maskC = nullptr;
    int64_t n_groups = -1; // counted on device, lazily
    // cloud structure derived from the read flags alone (group_stage_a): shared by pg_count and pg_featurize
    bool grouped = false;
    int64_t nofeat = 0;                        // reads with PG_READ_NOFEAT
    int64_t* gstart = nullptr;                 // n_groups + 1 cloud starts (byte offsets)
    unsigned long long* nofeat_len = nullptr;  // n_groups: bytes of NOFEAT reads per cloud
    uint32_t* maskR = nullptr;                 // maskF without the NOFEAT reads (only when nofeat > 0)
    uint32_t* wg = nullptr;                    // n_words: word -> cloud map (sliced path)
    int64_t min_group_len = 0;
    // entries of the shared partition: written by pg_count, reused by pg_featurize (bucket.cuh: kScatterShared)
    struct Segment { int64_t w0, w1; uint32_t* entries; int32_t* meta; unsigned long long* fill; uint2* runs; BucketGeom geo; };
    std::vector<Segment> stash;
    Workspace stash_ws;                        // backing store of all segments (borrowed from / returned to the ctx)
    uint32_t* stash_lost = nullptr;            // device flag: an overflow path bypassed the entry buffer
    bool stash_sweep = true;                   // the kept entries are in the layout of the one-sweep featurize pass (runs padded to 32 + bases)
};

struct pg_features {
    std::atomic<int> refs{ 1 };     // the DLPack deleter may run on any thread
    int device = 0;
    int64_t rows = 0;
    int32_t vs = 0, td = 0;
    uint32_t *abd_raw = nullptr, *tnf_raw = nullptr;
    float *abd = nullptr, *tnf = nullptr;
    double* weights = nullptr;
    int32_t* group_of_row = nullptr;
    int32_t* row_of_group = nullptr; // kept only by pg_featurize2(PG_FEAT_NO_ABUNDANCE): pg_features_add_counts maps cloud -> row with it
    int64_t n_groups = 0;
    bool normalized = false;
    bool exported = false; // handed out through DLPack: a consumer may still have work queued on ITS stream when the last reference goes
};

static thread_local std::string g_err;
static thread_local bool g_err_caught = false; // the newest message is in g_err only (caught(): no ctx at hand)

static int fail(pg_ctx* c, int code, const std::string& msg)
{
    g_err = msg;
    g_err_caught = false;
    if (c) c->err = msg;
    return code;
}

// No C++ exception crosses the C-ABI: every int-returning entry point is a function-try-block that ends here
// (std::bad_alloc from the host-side vectors / strings is the realistic one).
static int caught(const char* fn) noexcept
{
    g_err_caught = true;
    try {
        try { throw; }
        catch (const std::bad_alloc&) { g_err = std::string(fn) + ": out of host memory"; return PG_ERR_CUDA; }
        catch (const std::exception& e) { g_err = std::string(fn) + ": " + e.what(); return PG_ERR_STATE; }
        catch (...) { g_err = std::string(fn) + ": unknown C++ exception"; return PG_ERR_STATE; }
    } catch (...) { return PG_ERR_STATE; } // (even the message could not be built)
}

#define CK(call)                                                                                       \
    do {                                                                                               \
        cudaError_t e_ = (call);                                                                       \
        if (e_ != cudaSuccess)                                                                         \
            return fail(ctx, PG_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_));         \
    } while (0)

struct Timed { // CUDA-event span around the launches of one stage, on the ctx stream unless told otherwise
    pg_ctx* c; int which; cudaEvent_t a = nullptr, b = nullptr; cudaStream_t on = nullptr;
    static cudaEvent_t get(pg_ctx* c)
    {
        if (!c->pool.empty()) { cudaEvent_t e = c->pool.back(); c->pool.pop_back(); return e; }
        cudaEvent_t e; cudaEventCreate(&e); return e;
    }
    Timed(pg_ctx* c_, int w, int n_launches, cudaStream_t on_ = nullptr) : c(c_), which(w), on(on_ ? on_ : c_->stream)
    {
        a = get(c); b = get(c);
        cudaEventRecord(a, on);
        c->launches[w] += n_launches;
    }
    ~Timed()
    {
        cudaEventRecord(b, on);
        c->spans.push_back({ which, a, b });
        if (c->spans.size() > 4096) { // nobody is reading the spans: recycle the oldest
            c->pool.push_back(c->spans.front().a); c->pool.push_back(c->spans.front().b);
            c->spans.erase(c->spans.begin());
        }
    }
};

static inline int grid_for(int64_t n, int threads, int cap)
{
    int64_t g = (n + threads - 1) / threads;
    return (int)std::max<int64_t>(1, std::min<int64_t>(g, cap));
}

// scratch of at least `bytes`; contents undefined.  Re-allocation waits for the stream (the old block may be in use).
static cudaError_t ws_get(pg_ctx* ctx, Workspace& w, size_t bytes)
{
    if (w.bytes >= bytes && w.p) return cudaSuccess;
    if (w.p) { cudaStreamSynchronize(ctx->stream); cudaFree(w.p); w.p = nullptr; w.bytes = 0; }
    cudaError_t e = cudaMalloc(&w.p, std::max<size_t>(bytes, 256));
    if (e == cudaSuccess) w.bytes = std::max<size_t>(bytes, 256);
    else { w.p = nullptr; cudaGetLastError(); }
    return e;
}

constexpr size_t kBigBlock = 1u << 20;          // smaller requests stay with cudaMallocAsync
constexpr size_t kBigCacheLimit = 48ull << 30;  // idle bytes kept before the largest blocks are released

static void big_trim(pg_ctx* ctx, size_t keep_bytes)
{
    if (ctx->big.free_bytes <= keep_bytes) return;
    cudaStreamSynchronize(ctx->stream); // a cached block may still be read by queued work
    // Largest first: ONE cudaFree brings the cache well under the limit.  Measured against releasing the blocks idle longest
    // (one cloud per pair, 300 M pairs in 43 batches, HBM nearly full): 0.49 s of allocator time per job against 4.0 s - the
    // stale blocks are many and small (a cudaFree each), and within a batch the matrices are freed BEFORE the packed stream,
    // so "oldest" evicts exactly what the next batch asks for again.
    while (ctx->big.free_bytes > keep_bytes && !ctx->big.free_blocks.empty()) {
        auto it = std::prev(ctx->big.free_blocks.end());
        cudaFree(it->second);
        ctx->big.free_bytes -= it->first;
        ctx->big.free_blocks.erase(it);
    }
}

static cudaError_t big_alloc(pg_ctx* ctx, void** p, size_t bytes)
{
    bytes = (bytes + (2u << 20) - 1) & ~(size_t)((2u << 20) - 1);
    auto it = ctx->big.free_blocks.lower_bound(bytes);
    if (it != ctx->big.free_blocks.end() && it->first <= bytes + bytes / 4) {
        *p = it->second;
        ctx->big.live[*p] = it->first;
        ctx->big.free_bytes -= it->first;
        ctx->big.free_blocks.erase(it);
        return cudaSuccess;
    }
    cudaError_t e = cudaMalloc(p, bytes);
    if (e != cudaSuccess) { // out of memory: give the idle blocks back and try once more
        cudaGetLastError();
        big_trim(ctx, 0);
        e = cudaMalloc(p, bytes);
    }
    if (e == cudaSuccess) ctx->big.live[*p] = bytes;
    else cudaGetLastError();
    return e;
}

template <class T>
static cudaError_t dmalloc(pg_ctx* ctx, T** p, size_t n)
{
    const size_t bytes = std::max<size_t>(n, 1) * sizeof(T);
    if (bytes >= kBigBlock) return big_alloc(ctx, (void**)p, bytes);
    return cudaMallocFromPoolAsync((void**)p, bytes, ctx->mempool, ctx->stream);
}
static void dfree(pg_ctx* ctx, void* p)
{
    if (!p) return;
    auto it = ctx->big.live.find(p);
    if (it == ctx->big.live.end()) { cudaFreeAsync(p, ctx->stream); return; }
    ctx->big.free_blocks.insert({ it->second, p });
    ctx->big.free_bytes += it->second;
    ctx->big.live.erase(it);
    big_trim(ctx, kBigCacheLimit);
}

// ---------------------------------------------------------------------------
// TNF column table (count_tnf.cpp:54-76,138-164): column = rank of the canonical code
// ---------------------------------------------------------------------------
static int build_tnf_lut(int tk, std::vector<uint16_t>* lut)
{
    const int n = 1 << (2 * tk);
    std::vector<int> col_of_code(n, -1);
    int cols = 0;
    for (int v = 0; v < n; ++v)
        if (canonical_of_fwd((uint64_t)v, tk) == (uint64_t)v) col_of_code[v] = cols++;
    if (lut) {
        lut->resize(n);
        for (int w = 0; w < n; ++w) // w = LSB-first window as the kernels see it
            (*lut)[w] = (uint16_t)col_of_code[canonical_of_window((uint64_t)w, tk)];
    }
    return cols;
}

extern "C" int pg_tnf_dim(int tnf_k) { return (tnf_k < 1 || tnf_k > 6) ? -1 : build_tnf_lut(tnf_k, nullptr); }

extern "C" int pg_device_count(void)
try {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
    return n;
} catch (...) { return caught("pg_device_count"); }

extern "C" void pg_default_params(pg_params* p)
{
    memset(p, 0, sizeof(*p));
    p->device = 0;
    p->k = 15;            // pangaea.py:140
    p->tnf_k = 4;         // pangaea.py:139
    p->window_size = 10;  // pangaea.py:141
    p->vector_size = 400; // pangaea.py:142
    p->min_length = 2000; // pangaea.py:138
    p->min_qual_char = 0;
    p->table_mode = PG_TABLE_AUTO;
    p->table_capacity = 0;
}

extern "C" const char* pg_last_error(const pg_ctx* ctx) { return (ctx && !g_err_caught) ? ctx->err.c_str() : g_err.c_str(); }

static TableView view(pg_ctx* c)
{
    TableView t;
    t.counts = c->counts; t.slots = c->slots; t.capacity_mask = c->n_slots ? c->n_slots - 1 : 0;
    t.overflow = c->d_overflow; t.sat = c->d_sat; t.k = c->p.k;
    return t;
}

static int alloc_hash(pg_ctx* ctx, uint64_t slots)
{
    CK(cudaSetDevice(ctx->p.device));
    ctx->n_slots = slots;
    CK(cudaMalloc((void**)&ctx->slots, slots * sizeof(HashSlot)));
    hash_clear_kernel<<<ctx->sm_count * 8, 256, 0, ctx->stream>>>(ctx->slots, slots);
    CK(cudaGetLastError());
    return PG_OK;
}

extern "C" int pg_create(const pg_params* p, pg_ctx** out)
try {
    pg_ctx* ctx = nullptr;
    if (!p || !out) return fail(nullptr, PG_ERR_INVALID, "pg_create: null argument");
    *out = nullptr;
    if (p->k < 1 || p->k > 31) return fail(nullptr, PG_ERR_INVALID, "k must be in 1..31");
    if (p->tnf_k < 1 || p->tnf_k > 6) return fail(nullptr, PG_ERR_INVALID, "tnf_k must be in 1..6");
    if (p->window_size < 1) return fail(nullptr, PG_ERR_INVALID, "window_size must be >= 1");
    int mode = p->table_mode == PG_TABLE_AUTO ? (p->k <= 16 ? PG_TABLE_DENSE : PG_TABLE_HASH) : p->table_mode;
    if (mode != PG_TABLE_NONE && (p->vector_size < 1 || p->vector_size > 8192)) return fail(nullptr, PG_ERR_INVALID, "vector_size must be in 1..8192");
    if (p->vector_size < 1) return fail(nullptr, PG_ERR_INVALID, "vector_size must be >= 1");
    // counters saturate at 2^31 - 1: exact as long as every count that lands in a bin is below that (table.cuh)
    if ((uint64_t)p->window_size * (uint64_t)p->vector_size > 0x7FFFFFFFull && mode != PG_TABLE_NONE)
        return fail(nullptr, PG_ERR_INVALID, "window_size * vector_size must be <= 2^31 - 1");
    if (mode == PG_TABLE_DENSE && p->k > 16) return fail(nullptr, PG_ERR_INVALID, "dense table needs k <= 16");
    if (mode != PG_TABLE_DENSE && mode != PG_TABLE_HASH && mode != PG_TABLE_NONE) return fail(nullptr, PG_ERR_INVALID, "bad table_mode");
    if (p->table_capacity & (p->table_capacity - 1)) return fail(nullptr, PG_ERR_INVALID, "table_capacity must be a power of two");
    int n_dev = 0;
    if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev == 0)
        return fail(nullptr, PG_ERR_CUDA, "no CUDA device: pangaea_b200 has no CPU path");
    if (p->device < 0 || p->device >= n_dev) return fail(nullptr, PG_ERR_INVALID, "bad device ordinal");

    ctx = new (std::nothrow) pg_ctx();
    if (!ctx) return fail(nullptr, PG_ERR_INVALID, "out of host memory");
    ctx->p = *p;
    ctx->mode = mode == PG_TABLE_DENSE ? kDense : kHash;
    ctx->no_table = mode == PG_TABLE_NONE;
    std::vector<uint16_t> lut;
    ctx->td = build_tnf_lut(p->tnf_k, &lut);
#define CKC(call)                                                                                        \
    do {                                                                                                 \
        cudaError_t e_ = (call);                                                                         \
        if (e_ != cudaSuccess) {                                                                         \
            int rc_ = fail(nullptr, PG_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_));    \
            pg_destroy(ctx);                                                                             \
            return rc_;                                                                                  \
        }                                                                                                \
    } while (0)
    CKC(cudaSetDevice(p->device));
    CKC(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
    CKC(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
    CKC(cudaDeviceGetAttribute(&ctx->sm_count, cudaDevAttrMultiProcessorCount, p->device));
    {   // a pool of its own (other users of cudaMallocAsync in the process keep the default pool's settings); freed blocks
        // stay in it: steady-state steps allocate without touching the driver
        cudaMemPoolProps props = {};
        props.allocType = cudaMemAllocationTypePinned;
        props.handleTypes = cudaMemHandleTypeNone;
        props.location.type = cudaMemLocationTypeDevice;
        props.location.id = p->device;
        CKC(cudaMemPoolCreate(&ctx->mempool, &props));
        cudaMemPool_t pool = ctx->mempool;
        uint64_t keep = ~0ull;
        CKC(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep));
        // reuse by stream order only: with the opportunistic policies the block a request gets depends on how far the host
        // runs ahead of the GPU, and an unlucky choice ends in a fresh mapping (100+ ms per GB) in the middle of a step
        int off = 0;
        CKC(cudaMemPoolSetAttribute(pool, cudaMemPoolReuseAllowOpportunistic, &off));
        CKC(cudaMemPoolSetAttribute(pool, cudaMemPoolReuseAllowInternalDependencies, &off));
    }
    CKC(cudaMalloc((void**)&ctx->d_lut, lut.size() * sizeof(uint16_t)));
    CKC(cudaMemcpy(ctx->d_lut, lut.data(), lut.size() * sizeof(uint16_t), cudaMemcpyHostToDevice));
    CKC(cudaMalloc((void**)&ctx->d_overflow, sizeof(uint32_t)));
    CKC(cudaMemset(ctx->d_overflow, 0, sizeof(uint32_t)));
    CKC(cudaMalloc((void**)&ctx->d_sat, sizeof(uint32_t)));
    CKC(cudaMemset(ctx->d_sat, 0, sizeof(uint32_t)));
    CKC(cudaMalloc((void**)&ctx->d_scalar, 8 * sizeof(int64_t)));
    CKC(cudaMalloc((void**)&ctx->d_bucket, sizeof(BucketState)));
    { const char* e = getenv("PG_REGION_SLACK"); if (e && atof(e) > 0) ctx->region_slack = atof(e); }
    { const char* e = getenv("PG_FORCE_DIRECT"); ctx->force_direct = e && e[0] == '1'; }
    { const char* e = getenv("PG_SEG_WORDS"); if (e && atoll(e) >= 512) ctx->seg_words = ctx->count_seg_words = atoll(e) / 512 * 512; }
    { const char* e = getenv("PG_COUNT_L2"); ctx->count_l2 = e && e[0] == '1'; }
    { const char* e = getenv("PG_NO_SHARED"); ctx->no_shared = e && e[0] == '1'; }
    { const char* e = getenv("PG_FEAT_APPLY"); if (e && (e[0] == '0' || e[0] == '1')) ctx->feat_path = e[0] - '0'; }
    { const char* e = getenv("PG_TNF_OVERLAP"); if (e) ctx->tnf_overlap = std::max(0, std::min(8, atoi(e))); }
    { const char* e = getenv("PG_FEAT_SEG_WORDS"); if (e && atoll(e) >= 512) ctx->seg_words = atoll(e) / 512 * 512; }
    CKC(cudaMallocHost((void**)&ctx->h_pin, 8 * sizeof(int64_t)));
    if (ctx->no_table) {
        // nothing to allocate
    } else if (ctx->mode == kDense) {
        ctx->n_slots = dense_entries(p->k);
        CKC(cudaMalloc((void**)&ctx->counts, ctx->n_slots * sizeof(uint32_t)));
        CKC(cudaMemsetAsync(ctx->counts, 0, ctx->n_slots * sizeof(uint32_t), ctx->stream));
    } else if (p->table_capacity) {
        int rc = alloc_hash(ctx, p->table_capacity);
        if (rc != PG_OK) { pg_destroy(ctx); return rc; }
    }
    CKC(cudaFuncSetAttribute(featurize_kernel<kDense>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
    CKC(cudaFuncSetAttribute(featurize_kernel<kHash>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
    CKC(cudaFuncSetAttribute(bucket_scatter_kernel<15, kScatterCount>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(ScatterSmem<false>)));
    CKC(cudaFuncSetAttribute(bucket_scatter_kernel<0, kScatterCount>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(ScatterSmem<false>)));
    CKC(cudaFuncSetAttribute(bucket_scatter_kernel<15, kScatterFeat>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(ScatterSmem<true>)));
    CKC(cudaFuncSetAttribute(bucket_scatter_kernel<0, kScatterFeat>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(ScatterSmem<true>)));
    CKC(cudaFuncSetAttribute(bucket_scatter_kernel<15, kScatterShared>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(ScatterSmem<true>)));
    CKC(cudaFuncSetAttribute(bucket_scatter_kernel<0, kScatterShared>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(ScatterSmem<true>)));
    CKC(cudaFuncSetAttribute(bucket_split_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SplitSmem)));
    CKC(cudaFuncSetAttribute(bucket_collect_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    CKC(cudaFuncSetAttribute(bucket_collect_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    CKC(cudaFuncSetAttribute(sub_apply_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSubWords * 4 + 16));
    CKC(cudaFuncSetAttribute(tnf_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
    CKC(cudaFuncSetAttribute(tnf_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
    CKC(cudaStreamSynchronize(ctx->stream));
#undef CKC
    *out = ctx;
    return PG_OK;
} catch (...) { return caught("pg_create"); }

extern "C" void pg_destroy(pg_ctx* ctx)
{
    if (!ctx) return;
    cudaSetDevice(ctx->p.device);
    if (ctx->copy_stream) cudaStreamSynchronize(ctx->copy_stream);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    for (auto& s : ctx->spans) { cudaEventDestroy(s.a); cudaEventDestroy(s.b); }
    for (auto e : ctx->pool) cudaEventDestroy(e);
    for (auto& kv : ctx->big.free_blocks) cudaFree(kv.second); // (live blocks belong to batches / feature sets still around)
    cudaFree(ctx->ws_entries.p); cudaFree(ctx->ws_entries2.p); cudaFree(ctx->ws_feat.p);
    for (auto& w : ctx->stash_pool) cudaFree(w.p);
    cudaFree(ctx->ws_text[0].p); cudaFree(ctx->ws_text[1].p);
    cudaFree(ctx->counts); cudaFree(ctx->slots); cudaFree(ctx->d_lut); cudaFree(ctx->d_overflow); cudaFree(ctx->d_sat); cudaFree(ctx->d_scalar); cudaFree(ctx->d_bucket);
    if (ctx->h_pin) cudaFreeHost(ctx->h_pin);
    if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    if (ctx->mempool) cudaMemPoolDestroy(ctx->mempool);
    delete ctx;
}

extern "C" int pg_synchronize(pg_ctx* ctx)
try {
    if (!ctx) return fail(nullptr, PG_ERR_INVALID, "null ctx");
    CK(cudaSetDevice(ctx->p.device));
    CK(cudaStreamSynchronize(ctx->stream));
    return PG_OK;
} catch (...) { return caught("pg_synchronize"); }

extern "C" void* pg_stream(pg_ctx* ctx) { return ctx ? (void*)ctx->stream : nullptr; }

static size_t available_bytes(pg_ctx* ctx);

extern "C" int pg_mem_info(pg_ctx* ctx, int64_t* free_bytes, int64_t* total_bytes)
try {
    if (!ctx || !free_bytes || !total_bytes) return fail(ctx, PG_ERR_INVALID, "pg_mem_info: bad argument");
    CK(cudaSetDevice(ctx->p.device));
    size_t f = 0, t = 0;
    CK(cudaMemGetInfo(&f, &t));
    size_t idle = 0;
    for (auto& w : ctx->stash_pool) idle += w.bytes;
    *free_bytes = (int64_t)(available_bytes(ctx) + ctx->big.free_bytes + idle);
    *total_bytes = (int64_t)t;
    return PG_OK;
} catch (...) { return caught("pg_mem_info"); }

extern "C" int pg_trim(pg_ctx* ctx)
try {
    if (!ctx) return fail(nullptr, PG_ERR_INVALID, "null ctx");
    CK(cudaSetDevice(ctx->p.device));
    big_trim(ctx, 0); // (synchronises the stream first)
    CK(cudaStreamSynchronize(ctx->stream));
    for (auto& w : ctx->stash_pool) cudaFree(w.p);
    ctx->stash_pool.clear();
    if (ctx->mempool) CK(cudaMemPoolTrimTo(ctx->mempool, 0));
    return PG_OK;
} catch (...) { return caught("pg_trim"); }

// ---------------------------------------------------------------------------
// timing
// ---------------------------------------------------------------------------
extern "C" int pg_timing_reset(pg_ctx* ctx)
try {
    if (!ctx) return fail(nullptr, PG_ERR_INVALID, "null ctx");
    CK(cudaStreamSynchronize(ctx->stream));
    for (auto& s : ctx->spans) { ctx->pool.push_back(s.a); ctx->pool.push_back(s.b); }
    ctx->spans.clear();
    memset(ctx->launches, 0, sizeof(ctx->launches));
    return PG_OK;
} catch (...) { return caught("pg_timing_reset"); }

extern "C" int pg_timing_get(pg_ctx* ctx, int which, double* ms_out, int64_t* launches_out)
try {
    if (!ctx || which < 0 || which >= T_SLOTS) return fail(ctx, PG_ERR_INVALID, "pg_timing_get: bad argument");
    CK(cudaStreamSynchronize(ctx->stream));
    double ms = 0;
    int64_t n = 0;
    for (auto& s : ctx->spans)
        if (which == T_ALL || s.which == which) {
            float t = 0;
            CK(cudaEventElapsedTime(&t, s.a, s.b));
            ms += t;
        }
    for (int i = 0; i < T_SLOTS; ++i)
        if (which == T_ALL || i == which) n += ctx->launches[i];
    if (ms_out) *ms_out = ms;
    if (launches_out) *launches_out = n;
    return PG_OK;
} catch (...) { return caught("pg_timing_get"); }

// ---------------------------------------------------------------------------
// batches
// ---------------------------------------------------------------------------
static int alloc_packed(pg_ctx* ctx, pg_batch* b)
{
    b->n_words = (b->n_bytes + 31) / 32;
    CK(dmalloc(ctx, &b->codes, (size_t)b->n_words + 2));
    CK(dmalloc(ctx, &b->maskF, (size_t)b->n_words + 2));
    CK(dmalloc(ctx, &b->maskC, (size_t)b->n_words + 2));
    return PG_OK;
}

// pack words [w_begin, w_end) (w_end = n_words + 2 covers the two zeroed pad words)
static int pack_range(pg_ctx* ctx, pg_batch* b, int64_t w_begin, int64_t w_end)
{
    if (w_end <= w_begin) return PG_OK;
    {
        Timed t(ctx, T_PACK, 1);
        pack_kernel<<<grid_for(w_end - w_begin, 256, ctx->sm_count * 16), 256, 0, ctx->stream>>>(
            b->seq, b->qual, b->n_bytes, w_begin, w_end, (uint32_t)ctx->p.min_qual_char, b->codes, b->maskF, b->maskC);
    }
    CK(cudaGetLastError());
    return PG_OK;
}

static int pack_batch(pg_ctx* ctx, pg_batch* b)
{
    int rc = alloc_packed(ctx, b);
    if (rc) return rc;
    return pack_range(ctx, b, 0, b->n_words + 2);
}

static int check_reads(pg_ctx* ctx, const pg_reads* r)
{
    if (!ctx || !r) return fail(ctx, PG_ERR_INVALID, "null argument");
    if (r->n_reads < 0 || r->n_bytes < 0) return fail(ctx, PG_ERR_INVALID, "negative batch size");
    if (r->n_reads > 0 && (!r->seq || !r->read_off || !r->read_flag)) return fail(ctx, PG_ERR_INVALID, "null batch arrays");
    if (ctx->p.min_qual_char && !r->qual && r->n_bytes) return fail(ctx, PG_ERR_INVALID, "min_qual_char set but the batch carries no qualities");
    return PG_OK;
}

// device buffers of a host batch (no copies yet)
static int alloc_batch(pg_ctx* ctx, const pg_reads* h, pg_batch** out)
{
    pg_batch* b = new pg_batch();
    b->owns = true;
    b->n_reads = h->n_reads;
    b->n_bytes = h->n_bytes;
    const size_t padded = ((size_t)h->n_bytes + 63) & ~(size_t)31;
    bool ok = dmalloc(ctx, &b->seq, padded) == cudaSuccess && dmalloc(ctx, &b->read_off, (size_t)h->n_reads + 1) == cudaSuccess &&
              dmalloc(ctx, &b->read_flag, (size_t)h->n_reads) == cudaSuccess;
    const bool want_q = ctx->p.min_qual_char && h->qual;
    if (ok && want_q) ok = dmalloc(ctx, &b->qual, padded) == cudaSuccess;
    if (ok) ok = alloc_packed(ctx, b) == PG_OK;
    if (!ok) { pg_batch_free(ctx, b); return fail(ctx, PG_ERR_CUDA, "device allocation of the read batch failed"); }
    *out = b;
    return PG_OK;
}

static int check_host_batch(pg_ctx* ctx, const pg_reads* h)
{
    int rc = check_reads(ctx, h);
    if (rc) return rc;
    if (h->n_reads > 0 && (h->read_off[0] != 0 || h->read_off[h->n_reads] != h->n_bytes))
        return fail(ctx, PG_ERR_INVALID, "read_off must start at 0 and end at n_bytes");
    return PG_OK;
}

extern "C" int pg_batch_upload(pg_ctx* ctx, const pg_reads* h, pg_batch** out)
try {
    if (!out) return fail(ctx, PG_ERR_INVALID, "null out");
    *out = nullptr;
    int rc = check_host_batch(ctx, h);
    if (rc) return rc;
    CK(cudaSetDevice(ctx->p.device));
    pg_batch* b = nullptr;
    rc = alloc_batch(ctx, h, &b);
    if (rc) return rc;
    if (h->n_reads) {
        cudaMemcpyAsync(b->seq, h->seq, (size_t)h->n_bytes, cudaMemcpyHostToDevice, ctx->stream);
        if (b->qual) cudaMemcpyAsync(b->qual, h->qual, (size_t)h->n_bytes, cudaMemcpyHostToDevice, ctx->stream);
        cudaMemcpyAsync(b->read_off, h->read_off, ((size_t)h->n_reads + 1) * sizeof(int64_t), cudaMemcpyHostToDevice, ctx->stream);
        cudaMemcpyAsync(b->read_flag, h->read_flag, (size_t)h->n_reads, cudaMemcpyHostToDevice, ctx->stream);
    } else {
        cudaMemsetAsync(b->read_off, 0, sizeof(int64_t), ctx->stream);
    }
    rc = pack_range(ctx, b, 0, b->n_words + 2);
    if (rc == PG_OK && cudaGetLastError() != cudaSuccess) rc = fail(ctx, PG_ERR_CUDA, "pg_batch_upload: copy failed");
    if (rc) { pg_batch_free(ctx, b); return rc; }
    *out = b;
    return PG_OK;
} catch (...) { return caught("pg_batch_upload"); }

extern "C" int pg_batch_adopt(pg_ctx* ctx, const pg_reads* d, pg_batch** out)
try {
    int rc = check_reads(ctx, d);
    if (rc) return rc;
    if (!out) return fail(ctx, PG_ERR_INVALID, "null out");
    *out = nullptr;
    if (((uintptr_t)d->seq & 15) || (d->qual && ((uintptr_t)d->qual & 15))) return fail(ctx, PG_ERR_INVALID, "seq/qual must be 16-byte aligned");
    CK(cudaSetDevice(ctx->p.device));
    pg_batch* b = new pg_batch();
    b->owns = false;
    b->seq = const_cast<uint8_t*>(d->seq);
    b->qual = ctx->p.min_qual_char ? const_cast<uint8_t*>(d->qual) : nullptr;
    b->read_off = const_cast<int64_t*>(d->read_off);
    b->read_flag = const_cast<uint8_t*>(d->read_flag);
    b->n_reads = d->n_reads;
    b->n_bytes = d->n_bytes;
    rc = pack_batch(ctx, b);
    if (rc) { pg_batch_free(ctx, b); return rc; }
    *out = b;
    return PG_OK;
} catch (...) { return caught("pg_batch_adopt"); }

extern "C" void pg_batch_shape(const pg_batch* b, int64_t* n_reads, int64_t* n_bytes)
{
    if (n_reads) *n_reads = b ? b->n_reads : -1;
    if (n_bytes) *n_bytes = b ? b->n_bytes : -1;
}

extern "C" int pg_batch_download(pg_ctx* ctx, const pg_batch* b, uint8_t* seq_out, int64_t* read_off_out, uint8_t* read_flag_out)
try {
    if (!ctx || !b) return fail(ctx, PG_ERR_INVALID, "pg_batch_download: bad argument");
    if (seq_out && !b->seq) return fail(ctx, PG_ERR_STATE, "pg_batch_download: the batch was compacted, its bases are gone");
    CK(cudaSetDevice(ctx->p.device));
    if (seq_out && b->n_bytes) CK(cudaMemcpyAsync(seq_out, b->seq, (size_t)b->n_bytes, cudaMemcpyDeviceToHost, ctx->stream));
    if (read_off_out) CK(cudaMemcpyAsync(read_off_out, b->read_off, ((size_t)b->n_reads + 1) * sizeof(int64_t), cudaMemcpyDeviceToHost, ctx->stream));
    if (read_flag_out && b->n_reads) CK(cudaMemcpyAsync(read_flag_out, b->read_flag, (size_t)b->n_reads, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return PG_OK;
} catch (...) { return caught("pg_batch_download"); }

static void free_stash(pg_ctx* ctx, pg_batch* b)
{
    b->stash.clear();
    if (b->stash_ws.p) { // back to the ctx (later use is ordered on the same stream)
        ctx->stash_pool.push_back(b->stash_ws);
        b->stash_ws = Workspace();
    }
    dfree(ctx, b->stash_lost);
    b->stash_lost = nullptr;
}

extern "C" void pg_batch_free(pg_ctx* ctx, pg_batch* b)
{
    if (!b || !ctx) return;
    cudaSetDevice(ctx->p.device);
    if (b->owns) { dfree(ctx, b->seq); dfree(ctx, b->qual); dfree(ctx, b->read_off); dfree(ctx, b->read_flag); }
    else if (b->owns_meta) { dfree(ctx, b->read_off); dfree(ctx, b->read_flag); }
    dfree(ctx, b->codes); dfree(ctx, b->maskF); dfree(ctx, b->maskC);
    dfree(ctx, b->gstart); dfree(ctx, b->nofeat_len); dfree(ctx, b->maskR); dfree(ctx, b->wg);
    free_stash(ctx, b);
    delete b;
}

// ---------------------------------------------------------------------------
// table
// ---------------------------------------------------------------------------
static int table_ready(pg_ctx* ctx);

extern "C" int pg_table_clear(pg_ctx* ctx)
try {
    if (!ctx) return fail(nullptr, PG_ERR_INVALID, "null ctx");
    CK(cudaSetDevice(ctx->p.device));
    { int rc_ = table_ready(ctx); if (rc_) return rc_; }
    if (ctx->counts) CK(cudaMemsetAsync(ctx->counts, 0, ctx->n_slots * sizeof(uint32_t), ctx->stream));
    if (ctx->slots) hash_clear_kernel<<<ctx->sm_count * 8, 256, 0, ctx->stream>>>(ctx->slots, ctx->n_slots);
    CK(cudaMemsetAsync(ctx->d_overflow, 0, sizeof(uint32_t), ctx->stream));
    CK(cudaMemsetAsync(ctx->d_sat, 0, sizeof(uint32_t), ctx->stream));
    ctx->counted = false;
    ctx->zero_markers = false;
    return PG_OK;
} catch (...) { return caught("pg_table_clear"); }

static int ensure_table(pg_ctx* ctx, int64_t hint_windows)
{
    if (ctx->no_table) return fail(ctx, PG_ERR_STATE, "this ctx was created with PG_TABLE_NONE: it only normalises");
    if (ctx->have_table()) return PG_OK;
    // hash mode, capacity not given: 2x the number of windows of the first batch, within [2^20, 2^33]
    uint64_t want = 2 * (uint64_t)std::max<int64_t>(hint_windows, 1), slots = 1ull << 20;
    while (slots < want && slots < (1ull << 33)) slots <<= 1;
    return alloc_hash(ctx, slots);
}

static int check_overflow(pg_ctx* ctx)
{
    if (ctx->mode != kHash) return PG_OK;
    uint32_t of = 0;
    CK(cudaMemcpyAsync(ctx->h_pin, ctx->d_overflow, sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    memcpy(&of, ctx->h_pin, sizeof(of));
    if (of) return fail(ctx, PG_ERR_CAPACITY, "k-mer hash table is full: raise pg_params.table_capacity");
    return PG_OK;
}

// the dense table is swept slice by slice when it is much larger than L2 (bucket.cuh)
static bool use_buckets(const pg_ctx* c)
{
    if (c->mode != kDense || c->force_direct) return false;
    const uint64_t nb = c->n_slots >> kSliceBits; // entries carry 8 x index in 32 bits: index < 2^29
    return nb >= 2 && nb <= (uint64_t)kMaxBuckets && dense_bits(c->p.k) <= 29;
}

static BucketGeom bucket_geom(const pg_ctx* ctx, int64_t seg_words)
{
    BucketGeom geo;
    geo.n_buckets = (int)(ctx->n_slots >> kSliceBits);
    geo.low_mask = (1u << kSliceBits) - 1u;
    const double mean = (double)seg_words * 32.0 / geo.n_buckets;
    const unsigned long long cap = (unsigned long long)(mean * ctx->region_slack);
    geo.cap = std::max<unsigned long long>(32ull, ((cap + 31ull) / 32ull) * 32ull);
    return geo;
}

// Which featurize pass suits the batch?  Large clouds (the linked-read case: ~20 KB): one sweep - a warp's 32 entries fold into
// ~10 REDs, 42.6 ms against 37 + 35 ms for lookup + collect at the headline workload.  Tiny clouds (one per read pair, hybrid
// mode): lookup + collect - the sweep read-modify-writes every row from DRAM once per table slice (2.7 s against 0.7 + 0.6 s
// at 300 M pairs).  profiles/bench_r02_c2_v3_*.json, bench_r02_c4_*.json.
static bool use_sweep(const pg_ctx* ctx, int64_t n_bytes, int64_t n_groups)
{
    if (ctx->feat_path >= 0) return ctx->feat_path == 1;
    return n_bytes / std::max<int64_t>(1, n_groups) >= 2048;
}

// grid of a scatter launch: every resident CTA walks tiles with a grid stride
template <int MODE>
static int scatter_launch(pg_ctx* ctx, const ScatterParams& Q, const FeatParams& P)
{
    constexpr bool FEAT = MODE != kScatterCount;
    const int64_t n_tiles = (Q.w1 - Q.w0 + ScatterCfg<FEAT>::kTileWords - 1) / ScatterCfg<FEAT>::kTileWords;
    int occ = 1;
    if (Q.k == 15) CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, bucket_scatter_kernel<15, MODE>, ScatterCfg<FEAT>::kThreads, sizeof(ScatterSmem<FEAT>)));
    else CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, bucket_scatter_kernel<0, MODE>, ScatterCfg<FEAT>::kThreads, sizeof(ScatterSmem<FEAT>)));
    if (occ < 1) return fail(ctx, PG_ERR_CUDA, "scatter kernel does not fit on this device");
    const int grid = (int)std::max<int64_t>(1, std::min<int64_t>(n_tiles, (int64_t)ctx->sm_count * occ));
    bucket_reset_kernel<<<1, kMaxBuckets, 0, ctx->stream>>>(ctx->d_bucket, Q.geo.cap);
    if (Q.k == 15) bucket_scatter_kernel<15, MODE><<<grid, ScatterCfg<FEAT>::kThreads, sizeof(ScatterSmem<FEAT>), ctx->stream>>>(Q, P);
    else bucket_scatter_kernel<0, MODE><<<grid, ScatterCfg<FEAT>::kThreads, sizeof(ScatterSmem<FEAT>), ctx->stream>>>(Q, P);
    CK(cudaGetLastError());
    return PG_OK;
}

static int group_stage_a(pg_ctx* ctx, pg_batch* b, bool want_wg, bool packed = true);

// buffers and geometry of the sliced count pass, set up once per batch; segments are then
// counted one by one (pg_extract_features interleaves them with the upload)
struct CountPlan {
    bool two_level = false;
    bool shared = false;    // the level-1 entries also serve the featurize pass: one buffer per segment, kept in the batch
    int64_t seg_words = 0;
    BucketGeom geo = {};
    SubGeom sg = {};
    SubState ss = {};
    uint32_t* entries = nullptr;
    uint16_t* entries2 = nullptr;
    int split_grid = 1;
};

static void count_plan_free(pg_ctx* ctx, CountPlan& P)
{
    dfree(ctx, P.ss.cursors); dfree(ctx, P.ss.limits); dfree(ctx, P.ss.item_base); // the entry buffers belong to the ctx / the batch
    P = CountPlan();
}

// memory cudaMallocAsync can still hand out: free device memory + what the pool holds but does not use
static size_t available_bytes(pg_ctx* ctx)
{
    size_t free_b = 0, total_b = 0;
    if (cudaMemGetInfo(&free_b, &total_b) != cudaSuccess) return 0;
    cudaMemPool_t pool = ctx->mempool;
    uint64_t reserved = 0, used = 0;
    if (pool &&
        cudaMemPoolGetAttribute(pool, cudaMemPoolAttrReservedMemCurrent, &reserved) == cudaSuccess &&
        cudaMemPoolGetAttribute(pool, cudaMemPoolAttrUsedMemCurrent, &used) == cudaSuccess && reserved > used)
        free_b += (size_t)(reserved - used);
    return free_b;
}

static BucketGeom padded_geom(BucketGeom geo) // feature-format runs are padded to 32 entries
{
    geo.cap = (unsigned long long)((double)geo.cap * 1.08 / 32.0 + 1.0) * 32ull;
    return geo;
}

// One partition for both passes (bucket.cuh: kScatterShared) when the batch allows it: no quality filter (feature
// windows must be a subset of the counted ones), every cloud >= 64 bytes, and room to keep the entries until
// pg_featurize.  Otherwise each pass partitions for itself.
static int count_plan_init(pg_ctx* ctx, pg_batch* b, CountPlan& P, bool packed = true, bool keep = true)
{
    const int64_t n_words = b->n_words;
    P.two_level = !ctx->count_l2;
    P.seg_words = std::max<int64_t>(1, std::min<int64_t>(n_words, P.two_level ? ctx->count_seg_words : ctx->seg_words));
    P.geo = bucket_geom(ctx, P.seg_words);
    P.shared = false;
    if (P.two_level && keep && !ctx->no_shared && !ctx->p.min_qual_char && b->stash.empty() && b->read_flag && b->read_off && n_words > 0) {
        int rc = group_stage_a(ctx, b, true, packed);
        if (rc) return rc;
        const BucketGeom pg = padded_geom(P.geo);
        const size_t n_seg = (size_t)((n_words + P.seg_words - 1) / P.seg_words);
        const size_t runs_bytes = (size_t)((P.seg_words + ScatterCfg<true>::kTileWords - 1) / ScatterCfg<true>::kTileWords) * kMaxBuckets * sizeof(uint2);
        const size_t per_seg = (size_t)pg.cap * pg.n_buckets * 4 + (size_t)pg.cap * pg.n_buckets / 32 * 4 + kMaxBuckets * sizeof(unsigned long long) + runs_bytes;
        // as many leading segments as fit (a 125 M-pair batch has 12 of 14.6 GB each): the rest is partitioned per pass
        // leave room for what this step still allocates (level-2 entries, the scratch partitions of segments that are not
        // kept, the feature matrices) and 16 GB for the caller
        size_t n_keep = n_seg;
        // an idle buffer of the pool: the smallest that is large enough, else the largest (it is grown below)
        int pick = -1;
        for (int i = 0; i < (int)ctx->stash_pool.size(); ++i) {
            const size_t have = ctx->stash_pool[(size_t)i].bytes, best = pick < 0 ? 0 : ctx->stash_pool[(size_t)pick].bytes;
            const bool fits = have >= n_seg * per_seg, best_fits = pick >= 0 && best >= n_seg * per_seg;
            if (pick < 0 || (fits && (!best_fits || have < best)) || (!fits && !best_fits && have > best)) pick = i;
        }
        Workspace idle = pick >= 0 ? ctx->stash_pool[(size_t)pick] : Workspace();
        if (n_seg * per_seg > idle.bytes) { // no idle buffer is large enough: how much may one grow?  (steady-state
            // steps never get here - the memory queries below were seen to stall a step for tens of ms on a fresh box)
            const size_t avail = available_bytes(ctx) + idle.bytes;
            const size_t reserve = 3 * per_seg + ((size_t)16 << 30);
            const size_t budget = std::max<size_t>(idle.bytes, avail > reserve ? (size_t)(0.85 * (double)(avail - reserve)) : 0);
            n_keep = std::min<size_t>(n_seg, budget / per_seg);
        }
        if (const char* e = getenv("PG_STASH_SEGMENTS")) n_keep = std::min<size_t>(n_keep, (size_t)std::max(0, atoi(e))); // tests: force a partial stash
        if (b->min_group_len >= 64 && (packed || !b->nofeat) && n_keep > 0) {
            b->stash_ws = idle; // borrow the buffer (grown if too small)
            if (pick >= 0) ctx->stash_pool.erase(ctx->stash_pool.begin() + pick);
            if (ws_get(ctx, b->stash_ws, n_keep * per_seg) == cudaSuccess) {
                P.shared = true;
                P.geo = pg;
                uint8_t* base = (uint8_t*)b->stash_ws.p;
                const size_t E = (size_t)pg.cap * pg.n_buckets;
                for (size_t i = 0; i < n_keep; ++i) {
                    pg_batch::Segment sgm;
                    sgm.w0 = (int64_t)i * P.seg_words; sgm.w1 = std::min<int64_t>(n_words, sgm.w0 + P.seg_words);
                    sgm.entries = (uint32_t*)(base + i * E * 4);
                    sgm.meta = (int32_t*)(base + n_keep * E * 4 + i * (E / 32) * 4);
                    sgm.fill = (unsigned long long*)(base + n_keep * E * 4 + n_keep * (E / 32) * 4 + i * kMaxBuckets * sizeof(unsigned long long));
                    sgm.runs = (uint2*)(base + n_keep * E * 4 + n_keep * (E / 32) * 4 + n_keep * kMaxBuckets * sizeof(unsigned long long) + i * runs_bytes);
                    sgm.geo = pg;
                    b->stash.push_back(sgm);
                }
            } else {
                free_stash(ctx, b);
            }
        }
    }
    cudaError_t e = cudaSuccess;
    if (P.shared) {
        e = dmalloc(ctx, &b->stash_lost, 1);
        if (e == cudaSuccess) e = cudaMemsetAsync(b->stash_lost, 0, sizeof(uint32_t), ctx->stream);
    }
    const size_t n_seg_all = (size_t)((n_words + P.seg_words - 1) / P.seg_words);
    if (e == cudaSuccess && (!P.shared || b->stash.size() < n_seg_all)) { // segments that are not kept share one scratch buffer
        e = ws_get(ctx, ctx->ws_entries, (size_t)P.geo.cap * P.geo.n_buckets * 4);
        P.entries = (uint32_t*)ctx->ws_entries.p;
    }
    if (e == cudaSuccess && P.two_level) {
        P.sg.n_sub = P.geo.n_buckets * kSubFan;
        const double mean = (double)P.seg_words * 32.0 / P.sg.n_sub;
        // runs are padded to 8 entries: + 3.5 entries per run of 64 x (valid fraction) on average
        P.sg.cap = (uint32_t)std::min<double>(4.0e9, std::max(64.0, std::ceil(mean * 1.08 * ctx->region_slack / 8.0) * 8.0));
        P.sg.chunk = (uint32_t)std::max<double>(131072.0, std::ceil(2.0 * mean / 8.0) * 8.0);
        e = ws_get(ctx, ctx->ws_entries2, (size_t)P.sg.cap * P.sg.n_sub * 2);
        P.entries2 = (uint16_t*)ctx->ws_entries2.p;
        if (e == cudaSuccess) e = dmalloc(ctx, &P.ss.cursors, (size_t)P.sg.n_sub);
        if (e == cudaSuccess) e = dmalloc(ctx, &P.ss.limits, (size_t)P.sg.n_sub);
        if (e == cudaSuccess) e = dmalloc(ctx, &P.ss.item_base, (size_t)P.sg.n_sub + 1);
        int occ = 1;
        if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, bucket_split_kernel, kSplitThreads, sizeof(SplitSmem));
        P.split_grid = ctx->sm_count * std::max(occ, 1);
    }
    if (e != cudaSuccess) { count_plan_free(ctx, P); free_stash(ctx, b); return fail(ctx, PG_ERR_CUDA, std::string("count pass: ") + cudaGetErrorString(e)); }
    return PG_OK;
}

// end of a count segment (< 2^31 windows): counters that reached bit 31 are clamped to kCountMax (table.cuh: saturation).
// gated: the kernel returns at once unless a checked add raised the flag during the segment.
static cudaError_t saturate_if_flagged(pg_ctx* ctx, bool gated)
{
    if (ctx->mode != kDense || !ctx->counts) return cudaSuccess;
    table_saturate_kernel<<<ctx->sm_count * 8, 256, 0, ctx->stream>>>(ctx->counts, ctx->n_slots, kCountMax, gated ? ctx->d_sat : nullptr);
    return cudaMemsetAsync(ctx->d_sat, 0, sizeof(uint32_t), ctx->stream);
}

// count the windows that start in words [w0, w1) (w1 - w0 <= plan.seg_words); words w1, w1 + 1 must be packed too
static int count_segment(pg_ctx* ctx, CountPlan& P, pg_batch* b, int64_t w0, int64_t w1)
{
    ScatterParams Q = {};
    Q.codes = b->codes; Q.mask = b->maskC; Q.k = ctx->p.k; Q.geo = P.geo; Q.st = ctx->d_bucket;
    Q.entries = P.entries; Q.meta = nullptr; Q.table = ctx->counts; Q.lost = nullptr; Q.sat = ctx->d_sat; Q.run_pad = 4u;
    Q.w0 = w0; Q.w1 = w1;
    FeatParams F = {};
    int rc = PG_OK;
    const size_t si = (size_t)(w0 / P.seg_words);
    if (P.shared && si < b->stash.size()) {
        if (b->stash[si].w0 != w0 || b->stash[si].w1 != w1) return fail(ctx, PG_ERR_STATE, "count pass: segment does not match the shared partition plan");
        const pg_batch::Segment& sgm = b->stash[si];
        Q.entries = sgm.entries; Q.lost = b->stash_lost; Q.runs = sgm.runs;
        b->stash_sweep = use_sweep(ctx, b->n_bytes, b->n_groups);
        if (b->stash_sweep) { Q.meta = sgm.meta; Q.run_pad = 32u; } // the sweep reads one base cloud per aligned group of 32 entries
        F.maskF = b->maskR ? b->maskR : b->maskF; F.wg = b->wg; F.gstart = b->gstart; F.n_groups = b->n_groups;
        {
            Timed t(ctx, T_COUNT_SCATTER, 3);
            rc = scatter_launch<kScatterShared>(ctx, Q, F);
            if (!rc) bucket_save_fill_kernel<<<1, kMaxBuckets, 0, ctx->stream>>>(ctx->d_bucket, P.geo, sgm.fill);
        }
    } else {
        Timed t(ctx, T_COUNT_SCATTER, 2);
        rc = scatter_launch<kScatterCount>(ctx, Q, F);
    }
    if (rc) return rc;
    if (P.two_level) {
        {
            Timed t(ctx, T_COUNT_SPLIT, 2);
            sub_reset_kernel<<<16, 1024, 0, ctx->stream>>>(P.ss, P.sg);
            bucket_split_kernel<<<P.split_grid, kSplitThreads, sizeof(SplitSmem), ctx->stream>>>(Q.entries, P.geo, ctx->d_bucket, P.sg, P.ss, P.entries2, ctx->counts, ctx->d_sat);
        }
        Timed t(ctx, T_COUNT, 3);
        sub_items_kernel<<<1, 1024, 0, ctx->stream>>>(P.ss, P.sg);
        sub_apply_kernel<<<ctx->sm_count, kSubApplyThreads, kSubWords * 4 + 16, ctx->stream>>>(P.entries2, P.sg, P.ss, ctx->counts, ctx->d_sat);
        CK(saturate_if_flagged(ctx, true));
    } else {
        Timed t(ctx, T_COUNT, 1);
        bucket_apply_count_kernel<<<ctx->sm_count * 8, 256, 0, ctx->stream>>>(Q.entries, P.geo, ctx->d_bucket, ctx->counts);
        CK(saturate_if_flagged(ctx, false)); // plain REDs here (A/B path): clamp unconditionally
    }
    CK(cudaGetLastError());
    return PG_OK;
}

static int count_bucketed(pg_ctx* ctx, pg_batch* b, bool keep)
{
    CountPlan P;
    free_stash(ctx, b); // counting a batch again: its old entries are stale
    int rc = count_plan_init(ctx, b, P, true, keep);
    for (int64_t w0 = 0; !rc && w0 < b->n_words; w0 += P.seg_words) rc = count_segment(ctx, P, b, w0, std::min(b->n_words, w0 + P.seg_words));
    count_plan_free(ctx, P);
    return rc;
}

extern "C" int pg_count(pg_ctx* ctx, pg_batch* b) { return pg_count2(ctx, b, 1); }

extern "C" int pg_count2(pg_ctx* ctx, pg_batch* b, int keep_partition)
try {
    if (!ctx || !b) return fail(ctx, PG_ERR_INVALID, "null argument");
    CK(cudaSetDevice(ctx->p.device));
    { int rc_ = table_ready(ctx); if (rc_) return rc_; }
    if (ctx->zero_markers) return fail(ctx, PG_ERR_STATE, "pg_count: the table holds zero-count markers from pg_table_set - call pg_table_clear first");
    int rc = ensure_table(ctx, b->n_bytes);
    if (rc) return rc;
    if (b->n_words && use_buckets(ctx)) {
        rc = count_bucketed(ctx, b, keep_partition != 0);
        if (rc) return rc;
    } else if (b->n_words) {
        // launches of < 2^31 windows, each followed by the (gated) clamp: a dense counter cannot wrap (table.cuh: saturation)
        const int64_t step = 1ll << 26;
        for (int64_t w0 = 0; w0 < b->n_words; w0 += step) {
            const int64_t n = std::min(step, b->n_words - w0);
            Timed t(ctx, T_COUNT, 2);
            const int grid = grid_for(n, 256, ctx->sm_count * 8);
            if (ctx->mode == kDense) count_kernel<kDense><<<grid, 256, 0, ctx->stream>>>(b->codes + w0, b->maskC + w0, n, view(ctx));
            else count_kernel<kHash><<<grid, 256, 0, ctx->stream>>>(b->codes + w0, b->maskC + w0, n, view(ctx));
            CK(saturate_if_flagged(ctx, true));
        }
    }
    CK(cudaGetLastError());
    ctx->counted = true;
    return check_overflow(ctx);
} catch (...) { return caught("pg_count2"); }

extern "C" int pg_table_set(pg_ctx* ctx, const uint64_t* keys, const uint32_t* counts, int64_t n)
try {
    if (!ctx || (n > 0 && (!keys || !counts)) || n < 0) return fail(ctx, PG_ERR_INVALID, "pg_table_set: bad argument");
    CK(cudaSetDevice(ctx->p.device));
    { int rc_ = table_ready(ctx); if (rc_) return rc_; }
    int rc = ensure_table(ctx, n);
    if (rc) return rc;
    ctx->counted = true;
    if (!n) return PG_OK;
    for (int64_t i = 0; i < n; ++i) if (counts[i] == 0) { ctx->zero_markers = true; break; }
    uint64_t* dk; uint32_t* dc;
    CK(dmalloc(ctx, &dk, (size_t)n)); CK(dmalloc(ctx, &dc, (size_t)n));
    // one key at a time keeps "last assignment wins" (count_kmer.cpp:166) for duplicate keys:
    // duplicates are resolved on the host, in order, before the scatter
    std::vector<uint64_t> hk(keys, keys + n);
    std::vector<uint32_t> hc(counts, counts + n);
    {
        std::vector<int64_t> order(n);
        for (int64_t i = 0; i < n; ++i) order[i] = i;
        auto canon = [&](int64_t i) { return canonical_of_fwd(hk[i] & low_mask64(2 * ctx->p.k), ctx->p.k); };
        std::stable_sort(order.begin(), order.end(), [&](int64_t a, int64_t b2) { return canon(a) < canon(b2); });
        std::vector<uint64_t> k2; std::vector<uint32_t> c2;
        for (int64_t i = 0; i < n; ++i) {
            const bool last = (i + 1 == n) || canon(order[i + 1]) != canon(order[i]);
            if (last) { k2.push_back(hk[order[i]]); c2.push_back(hc[order[i]]); }
        }
        hk.swap(k2); hc.swap(c2);
    }
    const int64_t m = (int64_t)hk.size();
    CK(cudaMemcpyAsync(dk, hk.data(), m * sizeof(uint64_t), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(dc, hc.data(), m * sizeof(uint32_t), cudaMemcpyHostToDevice, ctx->stream));
    table_set_kernel<<<(int)((m + 255) / 256), 256, 0, ctx->stream>>>(view(ctx), ctx->mode, dk, dc, m);
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(ctx->stream)); // hk/hc are about to go out of scope
    dfree(ctx, dk); dfree(ctx, dc);
    return check_overflow(ctx);
} catch (...) { return caught("pg_table_set"); }

extern "C" int pg_table_get(pg_ctx* ctx, const uint64_t* keys, uint32_t* out, int64_t n)
try {
    if (!ctx || (n > 0 && (!keys || !out)) || n < 0) return fail(ctx, PG_ERR_INVALID, "pg_table_get: bad argument");
    if (!n) return PG_OK;
    CK(cudaSetDevice(ctx->p.device));
    { int rc_ = table_ready(ctx); if (rc_) return rc_; }
    if (!ctx->have_table()) { memset(out, 0, (size_t)n * sizeof(uint32_t)); return PG_OK; }
    uint64_t* dk; uint32_t* dc;
    CK(dmalloc(ctx, &dk, (size_t)n)); CK(dmalloc(ctx, &dc, (size_t)n));
    CK(cudaMemcpyAsync(dk, keys, n * sizeof(uint64_t), cudaMemcpyHostToDevice, ctx->stream));
    table_get_kernel<<<(int)((n + 255) / 256), 256, 0, ctx->stream>>>(view(ctx), ctx->mode, dk, dc, n);
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(out, dc, n * sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    dfree(ctx, dk); dfree(ctx, dc);
    return PG_OK;
} catch (...) { return caught("pg_table_get"); }

extern "C" int pg_table_size(pg_ctx* ctx, int64_t* n_distinct)
try {
    if (!ctx || !n_distinct) return fail(ctx, PG_ERR_INVALID, "pg_table_size: bad argument");
    *n_distinct = 0;
    if (!ctx->have_table()) return PG_OK;
    CK(cudaSetDevice(ctx->p.device));
    { int rc_ = table_ready(ctx); if (rc_) return rc_; }
    CK(cudaMemsetAsync(ctx->d_scalar, 0, sizeof(int64_t), ctx->stream));
    if (ctx->mode == kDense) table_nonzero_kernel<<<ctx->sm_count * 8, 256, 0, ctx->stream>>>(ctx->counts, nullptr, ctx->n_slots, (unsigned long long*)ctx->d_scalar);
    else table_nonzero_kernel<<<ctx->sm_count * 8, 256, 0, ctx->stream>>>(nullptr, ctx->slots, ctx->n_slots, (unsigned long long*)ctx->d_scalar);
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(ctx->h_pin, ctx->d_scalar, sizeof(int64_t), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    *n_distinct = ctx->h_pin[0];
    return PG_OK;
} catch (...) { return caught("pg_table_size"); }

extern "C" int pg_table_export(pg_ctx* ctx, uint64_t* keys_out, uint32_t* counts_out, int64_t cap, int64_t* n_out)
try {
    if (!ctx || !keys_out || !counts_out || cap < 0 || !n_out) return fail(ctx, PG_ERR_INVALID, "pg_table_export: bad argument");
    *n_out = 0;
    if (!ctx->have_table() || !cap) return PG_OK;
    CK(cudaSetDevice(ctx->p.device));
    { int rc_ = table_ready(ctx); if (rc_) return rc_; }
    uint64_t* dk; uint32_t* dc;
    CK(dmalloc(ctx, &dk, (size_t)cap)); CK(dmalloc(ctx, &dc, (size_t)cap));
    CK(cudaMemsetAsync(ctx->d_scalar, 0, sizeof(int64_t), ctx->stream));
    table_export_kernel<<<ctx->sm_count * 8, 256, 0, ctx->stream>>>(view(ctx), ctx->mode, ctx->n_slots, dk, dc, (unsigned long long)cap,
                                                                    (unsigned long long*)ctx->d_scalar);
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(ctx->h_pin, ctx->d_scalar, sizeof(int64_t), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    const int64_t n = std::min<int64_t>(ctx->h_pin[0], cap);
    std::vector<uint64_t> hk(n); std::vector<uint32_t> hc(n);
    CK(cudaMemcpyAsync(hk.data(), dk, n * sizeof(uint64_t), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(hc.data(), dc, n * sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    dfree(ctx, dk); dfree(ctx, dc);
    std::vector<int64_t> order(n);
    for (int64_t i = 0; i < n; ++i) order[i] = i;
    std::sort(order.begin(), order.end(), [&](int64_t a, int64_t b) { return hk[a] < hk[b]; });
    for (int64_t i = 0; i < n; ++i) { keys_out[i] = hk[order[i]]; counts_out[i] = hc[order[i]]; }
    *n_out = n;
    if (ctx->h_pin[0] > cap) return fail(ctx, PG_ERR_INVALID, "pg_table_export: buffer too small");
    return PG_OK;
} catch (...) { return caught("pg_table_export"); }

// order the ctx stream after a pending external write of the table; called by everything that touches the table
static int table_ready(pg_ctx* ctx)
{
    if (ctx->table_event) {
        cudaEvent_t ev = ctx->table_event;
        ctx->table_event = nullptr;
        CK(cudaStreamWaitEvent(ctx->stream, ev, 0));
    }
    return PG_OK;
}

extern "C" int pg_table_wait_event(pg_ctx* ctx, void* cuda_event)
try {
    if (!ctx) return fail(nullptr, PG_ERR_INVALID, "null ctx");
    ctx->table_event = (cudaEvent_t)cuda_event;
    return PG_OK;
} catch (...) { return caught("pg_table_wait_event"); }

extern "C" int pg_table_dense_view(pg_ctx* ctx, void** dev_ptr, int64_t* n_entries)
try {
    if (!ctx || !dev_ptr || !n_entries) return fail(ctx, PG_ERR_INVALID, "pg_table_dense_view: bad argument");
    if (ctx->mode != kDense) return fail(ctx, PG_ERR_STATE, "pg_table_dense_view: table is not dense");
    *dev_ptr = ctx->counts;
    *n_entries = (int64_t)ctx->n_slots;
    ctx->counted = true; // a caller that sums tables across ranks owns the contents
    return PG_OK;
} catch (...) { return caught("pg_table_dense_view"); }

extern "C" int pg_table_clamp(pg_ctx* ctx, uint32_t max_count)
try {
    if (!ctx) return fail(nullptr, PG_ERR_INVALID, "null ctx");
    if (ctx->mode != kDense || !ctx->counts) return fail(ctx, PG_ERR_STATE, "pg_table_clamp: table is not dense");
    if (ctx->zero_markers) return fail(ctx, PG_ERR_STATE, "pg_table_clamp: the table holds zero-count markers from pg_table_set");
    if (max_count > kCountMax) max_count = kCountMax;
    CK(cudaSetDevice(ctx->p.device));
    { int rc_ = table_ready(ctx); if (rc_) return rc_; }
    table_saturate_kernel<<<ctx->sm_count * 8, 256, 0, ctx->stream>>>(ctx->counts, ctx->n_slots, max_count, nullptr);
    CK(cudaGetLastError());
    return PG_OK;
} catch (...) { return caught("pg_table_clamp"); }

// ---------------------------------------------------------------------------
// grouping + featurize
// ---------------------------------------------------------------------------
// exclusive scan of (flags & bit) in tiles; returns device tile offsets (caller frees) and the total
static int scan_flags(pg_ctx* ctx, const uint8_t* d_flags, int64_t n, uint32_t bit, int32_t** tile_off_out, int64_t* total_out)
{
    const int64_t n_tiles = std::max<int64_t>(1, (n + kScanTile - 1) / kScanTile);
    int32_t* tile_off;
    CK(dmalloc(ctx, &tile_off, (size_t)n_tiles));
    flag_count_kernel<<<(int)n_tiles, kScanThreads, 0, ctx->stream>>>(d_flags, n, bit, tile_off);
    tile_scan_kernel<<<1, kScanThreads, 0, ctx->stream>>>(tile_off, n_tiles, ctx->d_scalar);
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(ctx->h_pin, ctx->d_scalar, sizeof(int64_t), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    *total_out = ctx->h_pin[0];
    *tile_off_out = tile_off;
    return PG_OK;
}

extern "C" int64_t pg_batch_n_groups(const pg_batch* b) { return b ? b->n_groups : -1; }

// smallest cloud of the batch, in bytes.  A cloud holds at least the read that carries its PG_READ_CHANGE flag; only the
// LAST one (what follows the last flag) can be empty - it owns no base and is skipped here.
__global__ void min_group_len_kernel(const int64_t* __restrict__ gstart, int64_t n_groups, unsigned long long* __restrict__ out)
{
    int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long len = ~0ull;
    if (g < n_groups && gstart[g + 1] > gstart[g]) len = (unsigned long long)(gstart[g + 1] - gstart[g]);
#pragma unroll
    for (int d = 16; d; d >>= 1) len = min(len, __shfl_xor_sync(0xffffffffu, len, d));
    __shared__ unsigned long long block_min; // one global atomic per CTA: with millions of clouds they all hit one address
    if (threadIdx.x == 0) block_min = ~0ull;
    __syncthreads();
    if ((threadIdx.x & 31) == 0 && len != ~0ull) atomicMin(&block_min, len);
    __syncthreads();
    if (threadIdx.x == 0 && block_min != ~0ull) atomicMin(out, block_min);
}

// Cloud structure that follows from the read flags alone: NOFEAT reads masked out of the feature mask,
// cloud starts, word -> cloud map.  Kept in the batch: pg_count (shared partition) and pg_featurize both use it.
// packed = false: the stream is not packed yet (pipelined upload) - the NOFEAT mask is left for a later call
static int group_stage_a(pg_ctx* ctx, pg_batch* b, bool want_wg, bool packed)
{
    if (b->grouped && (!want_wg || b->wg || !b->n_words) && (!b->nofeat || b->maskR || !packed)) return PG_OK;
    int32_t* tile_off = nullptr;
    int64_t changes = 0;
    int rc = PG_OK;
    Timed t(ctx, T_GROUP, 8);
    if (!b->grouped) {
        // PG_READ_NOFEAT reads are rare; when present their bases are masked out of a working copy of maskF
        rc = scan_flags(ctx, b->read_flag, b->n_reads, PG_READ_NOFEAT, &tile_off, &b->nofeat);
        if (rc) return rc;
        dfree(ctx, tile_off);
        rc = scan_flags(ctx, b->read_flag, b->n_reads, PG_READ_CHANGE, &tile_off, &changes);
        if (rc) return rc;
        b->n_groups = changes + 1;
        const int64_t n_groups = b->n_groups;
        if (n_groups > 0x7FFFFFFFll) { dfree(ctx, tile_off); return fail(ctx, PG_ERR_INVALID, "more than 2^31 clouds in one batch"); }
        CK(dmalloc(ctx, &b->gstart, (size_t)n_groups + 1));
        CK(dmalloc(ctx, &b->nofeat_len, (size_t)n_groups));
        CK(cudaMemsetAsync(b->nofeat_len, 0, (size_t)n_groups * sizeof(unsigned long long), ctx->stream));
        if (b->n_reads == 0) {
            CK(cudaMemsetAsync(b->gstart, 0, 2 * sizeof(int64_t), ctx->stream));
        } else {
            const int64_t n_tiles = std::max<int64_t>(1, (b->n_reads + kScanTile - 1) / kScanTile);
            group_starts_kernel<<<(int)n_tiles, kScanThreads, 0, ctx->stream>>>(b->read_flag, b->read_off, b->n_reads, tile_off, n_groups, b->gstart, b->nofeat_len);
        }
        dfree(ctx, tile_off);
        CK(cudaMemsetAsync(ctx->d_scalar, 0xFF, sizeof(int64_t), ctx->stream));
        min_group_len_kernel<<<(int)((n_groups + 255) / 256), 256, 0, ctx->stream>>>(b->gstart, n_groups, (unsigned long long*)ctx->d_scalar);
        CK(cudaGetLastError());
        CK(cudaMemcpyAsync(ctx->h_pin, ctx->d_scalar, sizeof(int64_t), cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        b->min_group_len = ctx->h_pin[0] < 0 ? 0 : ctx->h_pin[0];
        b->grouped = true;
    }
    if (b->nofeat && !b->maskR && packed) {
        CK(dmalloc(ctx, &b->maskR, (size_t)b->n_words + 2));
        CK(cudaMemcpyAsync(b->maskR, b->maskF, ((size_t)b->n_words + 2) * sizeof(uint32_t), cudaMemcpyDeviceToDevice, ctx->stream));
        clear_mask_ranges_kernel<<<grid_for(b->n_reads, 256, ctx->sm_count * 8), 256, 0, ctx->stream>>>(b->read_off, b->read_flag, b->n_reads, b->maskR);
        CK(cudaGetLastError());
    }
    if (want_wg && !b->wg && b->n_words) { // word -> cloud map of the streaming kernels
        CK(dmalloc(ctx, &b->wg, (size_t)b->n_words));
        const size_t max_big = (size_t)(b->n_words / kBigCloudWords) + 2; // clouds of more than kBigCloudWords words: they are disjoint
        uint32_t* big_list = nullptr; // [0] = how many, then the clouds
        CK(dmalloc(ctx, &big_list, max_big + 1));
        CK(cudaMemsetAsync(big_list, 0, sizeof(uint32_t), ctx->stream));
        word_groups_kernel<false><<<grid_for(b->n_groups * 32, 256, ctx->sm_count * 16), 256, 0, ctx->stream>>>(b->gstart, b->n_groups, b->n_bytes, b->wg, big_list + 1, big_list);
        word_groups_kernel<true><<<(int)std::min<int64_t>((int64_t)max_big, (int64_t)ctx->sm_count * 8), 256, 0, ctx->stream>>>(b->gstart, b->n_groups, b->n_bytes, b->wg, big_list + 1, big_list);
        dfree(ctx, big_list);
        CK(cudaGetLastError());
    }
    return PG_OK;
}

extern "C" int pg_featurize(pg_ctx* ctx, pg_batch* b, const uint8_t* group_keep, int64_t n_groups, pg_features** out)
try {
    return pg_featurize2(ctx, b, group_keep, n_groups, 0, out);
} catch (...) { return caught("pg_featurize"); }

extern "C" int pg_featurize2(pg_ctx* ctx, pg_batch* b, const uint8_t* group_keep, int64_t n_groups, int flags, pg_features** out)
try {
    if (!ctx || !b || !out || n_groups < 1 || !group_keep) return fail(ctx, PG_ERR_INVALID, "pg_featurize: bad argument");
    *out = nullptr;
    const bool no_abd = (flags & PG_FEAT_NO_ABUNDANCE) != 0;
    if (!no_abd && (!ctx->counted || !ctx->have_table())) return fail(ctx, PG_ERR_STATE, "pg_featurize: the k-mer table is empty - call pg_count / pg_table_set first");
    if (n_groups > 0x7FFFFFFFll) return fail(ctx, PG_ERR_INVALID, "more than 2^31 clouds in one batch");
    CK(cudaSetDevice(ctx->p.device));

    uint8_t *d_keep = nullptr, *emit = nullptr;
    int32_t *row_of_group = nullptr, *group_of_row_full = nullptr, *tile_off = nullptr, *row_lb = nullptr;
    int64_t rows = 0;
    int rc = PG_OK;
    pg_features* f = nullptr;
    const bool sliced = use_buckets(ctx) || no_abd; // (no_abd: grouping + the TNF kernel of the sliced path, no look-ups at all)
    cudaEvent_t tnf_fork = nullptr, tnf_join = nullptr; // TNF kernel on the second stream (sliced path)
    auto cleanup = [&]() {
        if (tnf_join) { cudaStreamWaitEvent(ctx->stream, tnf_join, 0); ctx->pool.push_back(tnf_fork); ctx->pool.push_back(tnf_join); tnf_fork = tnf_join = nullptr; }
        dfree(ctx, d_keep); dfree(ctx, emit);
        dfree(ctx, row_of_group); dfree(ctx, group_of_row_full);
        dfree(ctx, row_lb);
    };
#define CKF(call)                                                                                        \
    do {                                                                                                 \
        cudaError_t e_ = (call);                                                                         \
        if (e_ != cudaSuccess) {                                                                         \
            cleanup();                                                                                   \
            if (f) pg_features_free(ctx, f);                                                             \
            return fail(ctx, PG_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_));           \
        }                                                                                                \
    } while (0)

    // ---- grouping --------------------------------------------------------
    rc = group_stage_a(ctx, b, sliced);
    if (rc) return rc;
    if (b->n_groups != n_groups)
        return fail(ctx, PG_ERR_INVALID, "pg_featurize: n_groups (" + std::to_string(n_groups) + ") != 1 + change flags (" + std::to_string(b->n_groups) + ")");
    int64_t* const gstart = b->gstart;
    uint32_t* const wg = b->wg;
    {
        Timed t(ctx, T_GROUP, 4); // emit + scan x 2 + rows
        CKF(dmalloc(ctx, &d_keep, (size_t)n_groups));
        CKF(dmalloc(ctx, &emit, (size_t)n_groups));
        CKF(dmalloc(ctx, &row_of_group, (size_t)n_groups));
        CKF(dmalloc(ctx, &group_of_row_full, (size_t)n_groups));
        CKF(dmalloc(ctx, &row_lb, (size_t)n_groups));
        CKF(cudaMemcpyAsync(d_keep, group_keep, (size_t)n_groups, cudaMemcpyHostToDevice, ctx->stream));
        group_emit_kernel<<<(int)((n_groups + 255) / 256), 256, 0, ctx->stream>>>(gstart, b->nofeat_len, d_keep, n_groups, ctx->p.min_length, emit);
        CKF(cudaGetLastError());
        rc = scan_flags(ctx, emit, n_groups, 1u, &tile_off, &rows);
        if (rc) { cleanup(); return rc; }
        {
            const int64_t n_tiles = std::max<int64_t>(1, (n_groups + kScanTile - 1) / kScanTile);
            row_assign_kernel<<<(int)n_tiles, kScanThreads, 0, ctx->stream>>>(emit, n_groups, tile_off, row_of_group, group_of_row_full, row_lb);
        }
        dfree(ctx, tile_off);
        CKF(cudaGetLastError());
    }

    // ---- output matrices ---------------------------------------------------
    f = new pg_features();
    f->device = ctx->p.device;
    f->rows = rows;
    f->vs = ctx->p.vector_size;
    f->td = ctx->td;
    CKF(dmalloc(ctx, &f->abd_raw, (size_t)rows * f->vs));
    CKF(dmalloc(ctx, &f->tnf_raw, (size_t)rows * f->td));
    CKF(dmalloc(ctx, &f->group_of_row, (size_t)rows));
    CKF(cudaMemsetAsync(f->abd_raw, 0, (size_t)rows * f->vs * sizeof(uint32_t), ctx->stream));
    CKF(cudaMemsetAsync(f->tnf_raw, 0, (size_t)rows * f->td * sizeof(uint32_t), ctx->stream));
    if (rows) CKF(cudaMemcpyAsync(f->group_of_row, group_of_row_full, (size_t)rows * sizeof(int32_t), cudaMemcpyDeviceToDevice, ctx->stream));

    // ---- the pass over the bases -------------------------------------------
    if (rows && b->n_words) {
        FeatParams P;
        P.codes = b->codes; P.maskF = b->maskR ? b->maskR : b->maskF; P.n_words = b->n_words; P.n_bytes = b->n_bytes;
        P.gstart = gstart; P.n_groups = n_groups; P.row_of_group = row_of_group; P.wg = wg; P.row_lb = row_lb;
        P.tnf_k = ctx->p.tnf_k; P.vs = f->vs; P.td = f->td;
        P.ws = (uint32_t)ctx->p.window_size;
        const uint64_t clamp64 = (uint64_t)P.ws * (uint64_t)P.vs;
        P.clamp = clamp64 > 0xFFFFFFFFull ? 0xFFFFFFFFu : (uint32_t)clamp64;
        // q = (c * ceil(2^32 / ws)) >> 32 equals c / ws for every c < 2^32 / ws; c < clamp = ws * vs
        P.use_magic = (P.ws > 1 && (uint64_t)P.ws * clamp64 < (1ull << 32)) ? 1 : 0;
        P.magic = P.use_magic ? (uint32_t)(((1ull << 32) + P.ws - 1) / P.ws) : 0u;
        P.lut = ctx->d_lut;
        P.abd = f->abd_raw; P.tnf = f->tnf_raw;
        P.table = view(ctx);
        if (sliced) {
            // ---- L2-sliced path: TNF kernel, then partition of the window indices and slice-by-slice look-ups ----
            {
                const size_t nb = (size_t)1 << (2 * P.tnf_k);
                // cloud slots with private bins: enough for the clouds an average 8 KB tile spans (4 for 20 KB clouds - more would
                // only cost occupancy: 6.3 -> 8.6 ms - up to 32 for one cloud per read pair), within 32 KB of bins per CTA
                const size_t avg_cloud = (size_t)std::max<int64_t>(1, b->n_bytes / std::max<int64_t>(1, n_groups));
                const size_t want_slots = std::min<size_t>(kTnfSlots, std::max<size_t>(4, (size_t)kTnfThreads * 32 / avg_cloud + 3));
                P.tnf_slots = (int)std::max<size_t>(2, std::min<size_t>(want_slots, (32 * 1024) / (nb * sizeof(uint32_t))));
                // Tiny clouds (one per read pair: ~27 per tile): flush with one warp per cloud, folded columns, stores for the
                // clouds that lie inside the tile - and the kernel ALONE on the GPU: next to the look-ups it takes 5x as long
                // and doubles them (100 M pairs 2x150: TNF 60 ms + look-ups 116 ms in sequence, 301 + 244 ms side by side;
                // profiles/bench_r02_c4_tnf_ab.txt).  Clouds of many tiles: block-wide flush of the one or two slots in use,
                // beside the sweep (12.6 ms hidden at the headline workload).
                const bool tiny = avg_cloud < 2048;
                { const char* e = getenv("PG_TNF_FOLD"); P.tnf_fold = e ? atoi(e) : (tiny ? 1 : 0); } // (A/B switch)
                const size_t smem_t = ((size_t)P.tnf_slots * nb + 2) * sizeof(uint32_t) + 2 * nb
                                      + (P.tnf_k <= kTnfFoldMaxK ? ((size_t)(kTnfThreads / 32) * P.td + 2) * sizeof(uint32_t) : 0);
                // The TNF kernel is bound by shared-memory atomics, the look-up sweep below by L1 gathers: with a few CTAs per SM
                // on the second stream it runs NEXT TO the sweep instead of before it.
                const bool side = ctx->tnf_overlap > 0 && (!tiny || getenv("PG_TNF_OVERLAP"));
                const int64_t max_cta = (int64_t)ctx->sm_count * (side ? ctx->tnf_overlap : 8);
                const int64_t n_cta = std::min<int64_t>(max_cta, (b->n_words + kTnfThreads - 1) / kTnfThreads);
                int64_t wpc = (b->n_words + n_cta - 1) / n_cta;
                wpc = (wpc + kTnfThreads - 1) / kTnfThreads * kTnfThreads;
                P.words_per_cta = wpc;
                const int grid = (int)((b->n_words + wpc - 1) / wpc);
                cudaStream_t ts = ctx->stream;
                if (side) {
                    ts = ctx->copy_stream;
                    tnf_fork = Timed::get(ctx); tnf_join = Timed::get(ctx);
                    CKF(cudaEventRecord(tnf_fork, ctx->stream)); // grouping, the cleared matrices
                    CKF(cudaStreamWaitEvent(ts, tnf_fork, 0));
                }
                {
                    Timed t(ctx, T_TNF, 1, ts);
                    if (P.tnf_k == 4) tnf_kernel<4><<<grid, kTnfThreads, smem_t, ts>>>(P);
                    else tnf_kernel<0><<<grid, kTnfThreads, smem_t, ts>>>(P);
                }
                if (side) CKF(cudaEventRecord(tnf_join, ts));
            }
            if (no_abd) { // the caller adds the abundance tallies itself (pg_features_add_counts): it needs cloud -> row
                f->row_of_group = row_of_group; f->n_groups = n_groups;
                row_of_group = nullptr;
                CKF(cudaGetLastError());
                cleanup();
                *out = f;
                return PG_OK;
            }
            rc = table_ready(ctx); // the table may still be inside an all-reduce: grouping and TNF above did not need it
            if (rc) { cleanup(); pg_features_free(ctx, f); return rc; }
            // entries kept by pg_count (shared partition)?  usable unless an overflow path bypassed the buffer
            bool reuse = !b->stash.empty() && b->stash_lost;
            if (reuse) {
                CKF(cudaMemcpyAsync(ctx->h_pin, b->stash_lost, sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
                CKF(cudaStreamSynchronize(ctx->stream));
                uint32_t lost = 0;
                memcpy(&lost, ctx->h_pin, sizeof(lost));
                reuse = !lost;
            }
            int64_t covered = 0; // the kept segments are the leading ones
            if (reuse) {
                for (auto& sgm : b->stash) {
                    if (sgm.w0 != covered) { covered = -1; break; }
                    covered = sgm.w1;
                }
                if (covered < 0) { reuse = false; covered = 0; }
            }
            // look-ups + tallies of one segment's entries: lookup (ordered sweep, bins written back in place) then collect (stream
            // order, histograms in shared memory) - collect.cuh; PG_FEAT_APPLY=1 keeps the round-1 single sweep for A/B runs
            const size_t avg_cloud = (size_t)std::max<int64_t>(1, b->n_bytes / std::max<int64_t>(1, n_groups));
            const int tile_bytes = ScatterCfg<true>::kTileWords * 32;
            int c_slots = (int)std::min<size_t>(96, std::max<size_t>(4, (size_t)tile_bytes / avg_cloud + 3));
            while (c_slots > 2 && ((size_t)c_slots * P.vs + c_slots) * 4 > 190 * 1024) --c_slots;
            const size_t c_smem = ((size_t)c_slots * P.vs + c_slots) * 4;
            int c_occ = 1;
            const bool c_hot = avg_cloud >= 2048; // (PG_FEAT_APPLY=0 on large clouds; the automatic choice uses collect for tiny clouds only)
            CKF(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&c_occ, bucket_collect_kernel<false>, kCollectThreads, c_smem));
            auto lookup_collect = [&](uint32_t* entries, const int32_t* meta, const uint2* runs, const BucketGeom& geo, const unsigned long long* fill,
                                      int64_t w0, int64_t w1, bool shared_entries, bool sweep) {
                ctx->launches[T_GROUP] += 1; // the ticket reset, booked with the other housekeeping kernels
                bucket_reset_ticket_kernel<<<1, 1, 0, ctx->stream>>>(ctx->d_bucket);
                if (sweep) {
                    Timed t(ctx, T_FEAT, 1);
                    if (shared_entries) bucket_apply_feat_kernel<true><<<ctx->sm_count * 8, 256, 0, ctx->stream>>>(entries, meta, geo, ctx->d_bucket, fill, P);
                    else bucket_apply_feat_kernel<false><<<ctx->sm_count * 8, 256, 0, ctx->stream>>>(entries, meta, geo, ctx->d_bucket, nullptr, P);
                    return;
                }
                {
                    Timed t(ctx, T_FEAT, 1);
                    if (shared_entries) bucket_lookup_kernel<true><<<ctx->sm_count * 8, 256, 0, ctx->stream>>>(entries, geo, ctx->d_bucket, fill, P);
                    else bucket_lookup_kernel<false><<<ctx->sm_count * 8, 256, 0, ctx->stream>>>(entries, geo, ctx->d_bucket, nullptr, P);
                }
                CollectParams C;
                C.entries = entries; C.runs = reinterpret_cast<const RunRef*>(runs); C.geo = geo; C.w0 = w0; C.w1 = w1;
                C.tile_words = ScatterCfg<true>::kTileWords; C.slots = c_slots; C.shared = shared_entries ? 1 : 0;
                const int64_t n_tiles = (w1 - w0 + C.tile_words - 1) / C.tile_words;
                const int grid = (int)std::max<int64_t>(1, std::min<int64_t>(n_tiles, (int64_t)ctx->sm_count * std::max(c_occ, 1)));
                Timed t(ctx, T_COLLECT, 1);
                if (c_hot) bucket_collect_kernel<true><<<grid, kCollectThreads, c_smem, ctx->stream>>>(C, P);
                else bucket_collect_kernel<false><<<grid, kCollectThreads, c_smem, ctx->stream>>>(C, P);
            };
            if (reuse) {
                for (auto& sgm : b->stash) lookup_collect(sgm.entries, sgm.meta, sgm.runs, sgm.geo, sgm.fill, sgm.w0, sgm.w1, true, b->stash_sweep);
            }
            if (covered < b->n_words) { // what pg_count did not keep (or everything): partition for this pass alone
                const int64_t seg_words = std::min<int64_t>(b->n_words - covered, ctx->seg_words);
                const BucketGeom geo = padded_geom(bucket_geom(ctx, seg_words));
                const size_t E = (size_t)geo.cap * geo.n_buckets;
                const size_t runs_bytes = (size_t)((seg_words + ScatterCfg<true>::kTileWords - 1) / ScatterCfg<true>::kTileWords) * kMaxBuckets * sizeof(uint2);
                CKF(ws_get(ctx, ctx->ws_feat, E * 4 + E / 32 * 4 + runs_bytes));
                uint32_t* const feat_entries = (uint32_t*)ctx->ws_feat.p;
                int32_t* const feat_meta = (int32_t*)((uint8_t*)ctx->ws_feat.p + E * 4);
                uint2* const feat_runs = (uint2*)((uint8_t*)ctx->ws_feat.p + E * 4 + E / 32 * 4);
                ScatterParams Q = {};
                Q.codes = b->codes; Q.mask = P.maskF; Q.k = ctx->p.k; Q.geo = geo; Q.st = ctx->d_bucket;
                Q.entries = feat_entries; Q.table = ctx->counts; Q.lost = nullptr; Q.sat = ctx->d_sat; Q.runs = feat_runs; Q.run_pad = 4u;
                const bool sweep = use_sweep(ctx, b->n_bytes, n_groups);
                if (sweep) { Q.meta = feat_meta; Q.run_pad = 32u; }
                for (int64_t w0 = covered; w0 < b->n_words; w0 += seg_words) {
                    Q.w0 = w0; Q.w1 = std::min(b->n_words, w0 + seg_words);
                    {
                        Timed t(ctx, T_FEAT_SCATTER, 2);
                        rc = scatter_launch<kScatterFeat>(ctx, Q, P);
                    }
                    if (rc) { cleanup(); pg_features_free(ctx, f); return rc; }
                    lookup_collect(feat_entries, feat_meta, feat_runs, geo, nullptr, Q.w0, Q.w1, false, sweep);
                }
            }
            CKF(cudaGetLastError());
            if (reuse && !b->stash_sweep) free_stash(ctx, b); // the look-up pass overwrote the kept entries with bins: they serve once
            cleanup();
            *out = f;
            return PG_OK;
        }
        rc = table_ready(ctx);
        if (rc) { cleanup(); pg_features_free(ctx, f); return rc; }
        const size_t smem = (size_t)kSlots * (P.vs + P.td) * sizeof(uint32_t) + ((size_t)2 << (2 * P.tnf_k));
        int occ = 1;
        if (ctx->mode == kDense) CKF(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, featurize_kernel<kDense>, kFeatThreads, smem));
        else CKF(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, featurize_kernel<kHash>, kFeatThreads, smem));
        if (occ < 1) { cleanup(); pg_features_free(ctx, f); return fail(ctx, PG_ERR_INVALID, "vector_size / tnf_k too large for shared memory"); }
        const int64_t n_cta = std::min<int64_t>((int64_t)ctx->sm_count * occ, (b->n_words + kFeatThreads - 1) / kFeatThreads);
        int64_t wpc = (b->n_words + n_cta - 1) / n_cta;
        wpc = (wpc + kFeatThreads - 1) / kFeatThreads * kFeatThreads;
        P.words_per_cta = wpc;
        const int grid = (int)((b->n_words + wpc - 1) / wpc);
        Timed t(ctx, T_FEAT, 1);
        if (ctx->mode == kDense) featurize_kernel<kDense><<<grid, kFeatThreads, smem, ctx->stream>>>(P);
        else featurize_kernel<kHash><<<grid, kFeatThreads, smem, ctx->stream>>>(P);
    }
    CKF(cudaGetLastError());
    cleanup();
#undef CKF
    *out = f;
    return PG_OK;
} catch (...) { return caught("pg_featurize2"); }

// With a live ctx the buffers go back to the pool in stream order (no device-wide sync, and
// the next step's allocations reuse them); the DLPack deleter has no ctx and frees synchronously.
static void features_release(pg_features* f, pg_ctx* ctx = nullptr)
{
    if (f->refs.fetch_sub(1) > 1) return;
    cudaSetDevice(f->device);
    // A DLPack consumer (torch) drops its tensor - and with it the last reference - as soon as Python lets go of it, even
    // while kernels that read the buffer are still queued on the consumer's stream; a stream-ordered free on the ctx stream
    // (or handing the block to the next allocation) would pull the memory from under them.
    if (f->exported) cudaDeviceSynchronize();
    void* bufs[7] = { f->abd_raw, f->tnf_raw, f->abd, f->tnf, f->weights, f->group_of_row, f->row_of_group };
    for (void* p : bufs) {
        if (!p) continue;
        if (ctx) dfree(ctx, p); else cudaFree(p);
    }
    delete f;
}

extern "C" void pg_features_free(pg_ctx* ctx, pg_features* f)
{
    if (f) features_release(f, ctx);
}

extern "C" int64_t pg_features_rows(const pg_features* f) { return f ? f->rows : -1; }
extern "C" int32_t pg_features_abd_dim(const pg_features* f) { return f ? f->vs : -1; }
extern "C" int32_t pg_features_tnf_dim(const pg_features* f) { return f ? f->td : -1; }

extern "C" int pg_features_row_groups(pg_ctx* ctx, const pg_features* f, int64_t* groups_out)
try {
    if (!ctx || !f || (f->rows && !groups_out)) return fail(ctx, PG_ERR_INVALID, "pg_features_row_groups: bad argument");
    if (!f->rows) return PG_OK;
    CK(cudaSetDevice(ctx->p.device));
    std::vector<int32_t> tmp(f->rows);
    CK(cudaMemcpyAsync(tmp.data(), f->group_of_row, (size_t)f->rows * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    for (int64_t i = 0; i < f->rows; ++i) groups_out[i] = tmp[i];
    return PG_OK;
} catch (...) { return caught("pg_features_row_groups"); }

extern "C" int pg_features_copy_raw(pg_ctx* ctx, const pg_features* f, int32_t* abd_out, int32_t* tnf_out)
try {
    if (!ctx || !f) return fail(ctx, PG_ERR_INVALID, "pg_features_copy_raw: bad argument");
    CK(cudaSetDevice(ctx->p.device));
    if (abd_out && f->rows) CK(cudaMemcpyAsync(abd_out, f->abd_raw, (size_t)f->rows * f->vs * 4, cudaMemcpyDeviceToHost, ctx->stream));
    if (tnf_out && f->rows) CK(cudaMemcpyAsync(tnf_out, f->tnf_raw, (size_t)f->rows * f->td * 4, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return PG_OK;
} catch (...) { return caught("pg_features_copy_raw"); }

// ---------------------------------------------------------------------------
// Data.__init__
// ---------------------------------------------------------------------------
extern "C" int pg_normalize(pg_ctx* ctx, pg_features* f)
try {
    if (!ctx || !f) return fail(ctx, PG_ERR_INVALID, "pg_normalize: bad argument");
    if (f->normalized) return PG_OK;
    CK(cudaSetDevice(ctx->p.device));
    CK(dmalloc(ctx, &f->abd, (size_t)f->rows * f->vs));
    CK(dmalloc(ctx, &f->tnf, (size_t)f->rows * f->td));
    CK(dmalloc(ctx, &f->weights, (size_t)f->rows));
    if (f->rows) {
        Timed t(ctx, T_NORM, 2);
        auto launch = [&](const uint32_t* raw, int dim, float* out, double* w) {
            if (dim % 4 == 0 && dim <= 512) { // the production shapes: one warp per row, 16-byte vectors, one read
                const int nvec = dim / 4, vpl = (nvec + 31) / 32;
                const int grid = grid_for(f->rows * 32, 256, ctx->sm_count * 8);
                const uint4* r4 = reinterpret_cast<const uint4*>(raw);
                float4* o4 = reinterpret_cast<float4*>(out);
                if (vpl == 1) normalize_rows_vec_kernel<1><<<grid, 256, 0, ctx->stream>>>(r4, f->rows, nvec, o4, w, 1);
                else if (vpl == 2) normalize_rows_vec_kernel<2><<<grid, 256, 0, ctx->stream>>>(r4, f->rows, nvec, o4, w, 1);
                else if (vpl == 3) normalize_rows_vec_kernel<3><<<grid, 256, 0, ctx->stream>>>(r4, f->rows, nvec, o4, w, 1);
                else normalize_rows_vec_kernel<4><<<grid, 256, 0, ctx->stream>>>(r4, f->rows, nvec, o4, w, 1);
            } else if (f->rows >= (1 << 20)) { // millions of rows: four of them per warp in flight
                normalize_rows_kernel<8><<<grid_for(f->rows * 8, 256, ctx->sm_count * 8), 256, 0, ctx->stream>>>(raw, f->rows, dim, out, w, 1);
            } else {
                normalize_rows_kernel<32><<<grid_for(f->rows * 32, 256, ctx->sm_count * 8), 256, 0, ctx->stream>>>(raw, f->rows, dim, out, w, 1);
            }
        };
        launch(f->abd_raw, f->vs, f->abd, f->weights);
        launch(f->tnf_raw, f->td, f->tnf, nullptr);
    }
    CK(cudaGetLastError());
    f->normalized = true;
    return PG_OK;
} catch (...) { return caught("pg_normalize"); }

extern "C" int pg_features_from_raw(pg_ctx* ctx, const uint32_t* abd, const uint32_t* tnf, int64_t rows, int32_t abd_dim, int32_t tnf_dim, pg_features** out)
try {
    if (!ctx || !out || rows < 0 || abd_dim < 1 || tnf_dim < 1 || (rows && (!abd || !tnf))) return fail(ctx, PG_ERR_INVALID, "pg_features_from_raw: bad argument");
    *out = nullptr;
    CK(cudaSetDevice(ctx->p.device));
    pg_features* f = new pg_features();
    f->device = ctx->p.device; f->rows = rows; f->vs = abd_dim; f->td = tnf_dim;
    cudaError_t e = dmalloc(ctx, &f->abd_raw, (size_t)rows * abd_dim);
    if (e == cudaSuccess) e = dmalloc(ctx, &f->tnf_raw, (size_t)rows * tnf_dim);
    if (e == cudaSuccess) e = dmalloc(ctx, &f->group_of_row, (size_t)rows);
    if (e == cudaSuccess && rows) e = cudaMemcpyAsync(f->abd_raw, abd, (size_t)rows * abd_dim * 4, cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess && rows) e = cudaMemcpyAsync(f->tnf_raw, tnf, (size_t)rows * tnf_dim * 4, cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess && rows) e = cudaMemsetAsync(f->group_of_row, 0, (size_t)rows * 4, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) { pg_features_free(ctx, f); return fail(ctx, PG_ERR_CUDA, std::string("pg_features_from_raw: ") + cudaGetErrorString(e)); }
    *out = f;
    return PG_OK;
} catch (...) { return caught("pg_features_from_raw"); }

extern "C" int pg_features_copy_normalized(pg_ctx* ctx, const pg_features* f, float* abd_out, float* tnf_out, double* weights_out)
try {
    if (!ctx || !f) return fail(ctx, PG_ERR_INVALID, "pg_features_copy_normalized: bad argument");
    if (!f->normalized) return fail(ctx, PG_ERR_STATE, "call pg_normalize first");
    CK(cudaSetDevice(ctx->p.device));
    if (abd_out && f->rows) CK(cudaMemcpyAsync(abd_out, f->abd, (size_t)f->rows * f->vs * 4, cudaMemcpyDeviceToHost, ctx->stream));
    if (tnf_out && f->rows) CK(cudaMemcpyAsync(tnf_out, f->tnf, (size_t)f->rows * f->td * 4, cudaMemcpyDeviceToHost, ctx->stream));
    if (weights_out && f->rows) CK(cudaMemcpyAsync(weights_out, f->weights, (size_t)f->rows * 8, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return PG_OK;
} catch (...) { return caught("pg_features_copy_normalized"); }

extern "C" void* pg_features_device_ptr(const pg_features* f, int which)
{
    if (!f) return nullptr;
    switch (which) {
    case 0: return f->abd_raw;
    case 1: return f->tnf_raw;
    case 2: return f->abd;
    case 3: return f->tnf;
    case 4: return f->weights;
    default: return nullptr;
    }
}

// ---- DLPack (v0.8 ABI, declared locally: no external header needed) -----------
namespace {
struct DLDevice { int32_t device_type; int32_t device_id; };
struct DLDataType { uint8_t code; uint8_t bits; uint16_t lanes; };
struct DLTensor { void* data; DLDevice device; int32_t ndim; DLDataType dtype; int64_t* shape; int64_t* strides; uint64_t byte_offset; };
struct DLManagedTensor { DLTensor dl_tensor; void* manager_ctx; void (*deleter)(DLManagedTensor*); };
struct DlHolder { DLManagedTensor mt; int64_t shape[2]; pg_features* f; };
void dl_deleter(DLManagedTensor* mt)
{
    DlHolder* h = reinterpret_cast<DlHolder*>(mt->manager_ctx);
    features_release(h->f);
    delete h;
}
} // namespace

extern "C" void* pg_features_dlpack(pg_ctx* ctx, pg_features* f, int which)
{
    if (!ctx || !f || which < 0 || which > 4) { fail(ctx, PG_ERR_INVALID, "pg_features_dlpack: bad argument"); return nullptr; }
    if (which >= 2 && !f->normalized) { fail(ctx, PG_ERR_STATE, "call pg_normalize first"); return nullptr; }
    cudaSetDevice(ctx->p.device);
    cudaStreamSynchronize(ctx->stream); // the consumer runs on its own stream
    {   // the consumer's deleter may run on any thread, after the ctx is gone: take the buffers out of the ctx's block
        // cache - from here on they are plain cudaMalloc blocks, released with cudaFree / cudaFreeAsync
        void* bufs[7] = { f->abd_raw, f->tnf_raw, f->abd, f->tnf, f->weights, f->group_of_row, f->row_of_group };
        for (void* p : bufs) if (p) ctx->big.live.erase(p);
    }
    DlHolder* h = new DlHolder();
    h->f = f;
    f->exported = true;
    ++f->refs;
    DLTensor& t = h->mt.dl_tensor;
    t.data = pg_features_device_ptr(f, which);
    t.device = { 2 /* kDLCUDA */, f->device };
    t.byte_offset = 0;
    t.strides = nullptr;
    t.shape = h->shape;
    h->shape[0] = f->rows;
    if (which == 4) { t.ndim = 1; t.dtype = { 2 /* float */, 64, 1 }; }
    else {
        t.ndim = 2;
        h->shape[1] = (which == 0 || which == 2) ? f->vs : f->td;
        t.dtype = which < 2 ? DLDataType{ 0 /* int */, 32, 1 } : DLDataType{ 2, 32, 1 };
    }
    h->mt.manager_ctx = h;
    h->mt.deleter = dl_deleter;
    return &h->mt;
}

// ---------------------------------------------------------------------------
// whole path, host buffers in
// ---------------------------------------------------------------------------
// Sliced (dense k <= 15) path: the batch crosses PCIe in chunks on a second stream while the
// compute stream packs chunk c and counts chunk c - 1 (a window of chunk c - 1 may end in the first
// words of chunk c), so the count pass hides behind the copy.  Featurize needs the complete table
// and starts when the last chunk is counted.
static int upload_and_count_pipelined(pg_ctx* ctx, const pg_reads* h, pg_batch** out, bool keep = true)
{
    pg_batch* b = nullptr;
    int rc = alloc_batch(ctx, h, &b);
    if (rc) return rc;
    CountPlan P;
    cudaEvent_t ev = Timed::get(ctx), ev_meta = Timed::get(ctx);
    auto bail = [&](int code) { // copies may still be in flight into the buffers freed here
        cudaStreamSynchronize(ctx->copy_stream);
        count_plan_free(ctx, P); pg_batch_free(ctx, b); ctx->pool.push_back(ev); ctx->pool.push_back(ev_meta);
        return code;
    };
#define CKB(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return bail(fail(ctx, PG_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_))); } while (0)
    CKB(cudaEventRecord(ev, ctx->stream));             // allocations (stream-ordered) and the table clear come first
    CKB(cudaStreamWaitEvent(ctx->copy_stream, ev, 0));
    // the read offsets / flags go first: the cloud structure (shared partition) is derived from them while chunk 0 is in flight
    CKB(cudaMemcpyAsync(b->read_off, h->read_off, ((size_t)h->n_reads + 1) * sizeof(int64_t), cudaMemcpyHostToDevice, ctx->copy_stream));
    CKB(cudaMemcpyAsync(b->read_flag, h->read_flag, (size_t)h->n_reads, cudaMemcpyHostToDevice, ctx->copy_stream));
    CKB(cudaEventRecord(ev_meta, ctx->copy_stream));
    const int64_t seg = std::max<int64_t>(1, std::min<int64_t>(b->n_words, ctx->count_l2 ? ctx->seg_words : ctx->count_seg_words));
    const int64_t n_chunks = (b->n_words + seg - 1) / seg;
    auto copy_chunk = [&](int64_t c) -> cudaError_t {
        const int64_t w0 = c * seg, w1 = std::min(b->n_words, w0 + seg);
        const size_t lo = (size_t)w0 * 32, hi = std::min<size_t>((size_t)w1 * 32, (size_t)b->n_bytes);
        cudaError_t e = cudaMemcpyAsync(b->seq + lo, h->seq + lo, hi - lo, cudaMemcpyHostToDevice, ctx->copy_stream);
        if (e == cudaSuccess && b->qual) e = cudaMemcpyAsync(b->qual + lo, h->qual + lo, hi - lo, cudaMemcpyHostToDevice, ctx->copy_stream);
        if (e == cudaSuccess) e = cudaEventRecord(ev, ctx->copy_stream); // re-recording is fine: the wait below captured the previous record
        return e;
    };
    CKB(copy_chunk(0));
    CKB(cudaStreamWaitEvent(ctx->stream, ev_meta, 0));
    CKB(cudaStreamWaitEvent(ctx->stream, ev, 0));      // (chunk 0's record; captured now, before the event is re-recorded)
    rc = count_plan_init(ctx, b, P, false, keep);      // derives the cloud structure (host waits for the flags only)
    if (rc) return bail(rc);
    if (P.seg_words != seg) return bail(fail(ctx, PG_ERR_STATE, "pipelined upload: segment size mismatch"));
    for (int64_t c = 0; c < n_chunks; ++c) {
        const int64_t w0 = c * seg, w1 = std::min(b->n_words, w0 + seg);
        if (c > 0) {
            CKB(copy_chunk(c));
            CKB(cudaStreamWaitEvent(ctx->stream, ev, 0));
        }
        rc = pack_range(ctx, b, w0, c + 1 == n_chunks ? b->n_words + 2 : w1);
        if (!rc && c > 0) rc = count_segment(ctx, P, b, (c - 1) * seg, w0);
        if (rc) return bail(rc);
    }
    rc = count_segment(ctx, P, b, (n_chunks - 1) * seg, b->n_words);
    if (rc) return bail(rc);
#undef CKB
    count_plan_free(ctx, P);
    ctx->pool.push_back(ev); ctx->pool.push_back(ev_meta);
    ctx->counted = true;
    *out = b;
    return PG_OK;
}

// upload + count of one batch of a stream; the table is NOT cleared (batches add up)
extern "C" int pg_batch_upload_count(pg_ctx* ctx, const pg_reads* host, int keep_partition, pg_batch** out)
try {
    if (!out) return fail(ctx, PG_ERR_INVALID, "null out");
    *out = nullptr;
    if (!ctx) return fail(nullptr, PG_ERR_INVALID, "null ctx");
    CK(cudaSetDevice(ctx->p.device));
    { int rc_ = table_ready(ctx); if (rc_) return rc_; }
    if (ctx->zero_markers) return fail(ctx, PG_ERR_STATE, "pg_batch_upload_count: the table holds zero-count markers from pg_table_set - call pg_table_clear first");
    int rc = ensure_table(ctx, host ? host->n_bytes : 0);
    if (rc) return rc;
    pg_batch* b = nullptr;
    if (use_buckets(ctx) && host && host->n_reads > 0 && host->n_bytes >= (1 << 20)) {
        rc = check_host_batch(ctx, host);
        if (!rc) rc = upload_and_count_pipelined(ctx, host, &b, keep_partition != 0);
        if (rc) return rc;
    } else {
        rc = pg_batch_upload(ctx, host, &b);
        if (rc) return rc;
        rc = pg_count2(ctx, b, keep_partition);
        if (rc) { pg_batch_free(ctx, b); return rc; }
    }
    *out = b;
    return PG_OK;
} catch (...) { return caught("pg_batch_upload_count"); }

extern "C" int pg_batch_compact(pg_ctx* ctx, pg_batch* b)
try {
    if (!ctx || !b) return fail(ctx, PG_ERR_INVALID, "null argument");
    CK(cudaSetDevice(ctx->p.device));
    if (b->owns) { dfree(ctx, b->seq); dfree(ctx, b->qual); }
    else if (!b->owns_meta) {
        // an adopted batch: the caller may release ALL its buffers now, so the read offsets and flags - which pg_featurize
        // still needs - move into memory of the batch's own (9 B per read)
        int64_t* off = nullptr;
        uint8_t* flag = nullptr;
        CK(dmalloc(ctx, &off, (size_t)b->n_reads + 1));
        CK(dmalloc(ctx, &flag, (size_t)std::max<int64_t>(b->n_reads, 1)));
        CK(cudaMemcpyAsync(off, b->read_off, ((size_t)b->n_reads + 1) * sizeof(int64_t), cudaMemcpyDeviceToDevice, ctx->stream));
        if (b->n_reads) CK(cudaMemcpyAsync(flag, b->read_flag, (size_t)b->n_reads, cudaMemcpyDeviceToDevice, ctx->stream));
        b->read_off = off; b->read_flag = flag; b->owns_meta = true;
        CK(cudaStreamSynchronize(ctx->stream)); // the caller's buffers are free to go when this returns
    }
    b->seq = b->qual = nullptr; // (a kept partition stays: pg_count2's keep_partition decides about it)
    return PG_OK;
} catch (...) { return caught("pg_batch_compact"); }

extern "C" int pg_features_concat(pg_ctx* ctx, pg_features* const* parts, int32_t n_parts, pg_features** out)
try {
    if (!ctx || !out || n_parts < 0 || (n_parts && !parts)) return fail(ctx, PG_ERR_INVALID, "pg_features_concat: bad argument");
    *out = nullptr;
    CK(cudaSetDevice(ctx->p.device));
    int64_t rows = 0;
    int32_t vs = ctx->p.vector_size, td = ctx->td;
    for (int i = 0; i < n_parts; ++i) {
        if (!parts[i]) return fail(ctx, PG_ERR_INVALID, "pg_features_concat: null part");
        if (i == 0) { vs = parts[i]->vs; td = parts[i]->td; }
        if (parts[i]->vs != vs || parts[i]->td != td) return fail(ctx, PG_ERR_INVALID, "pg_features_concat: parts differ in shape");
        rows += parts[i]->rows;
    }
    pg_features* f = new pg_features();
    f->device = ctx->p.device; f->rows = rows; f->vs = vs; f->td = td;
    cudaError_t e = dmalloc(ctx, &f->abd_raw, (size_t)rows * vs);
    if (e == cudaSuccess) e = dmalloc(ctx, &f->tnf_raw, (size_t)rows * td);
    if (e == cudaSuccess) e = dmalloc(ctx, &f->group_of_row, (size_t)rows);
    int64_t at = 0;
    for (int i = 0; i < n_parts && e == cudaSuccess; ++i) {
        const pg_features* q = parts[i];
        if (!q->rows) continue;
        e = cudaMemcpyAsync(f->abd_raw + at * vs, q->abd_raw, (size_t)q->rows * vs * 4, cudaMemcpyDeviceToDevice, ctx->stream);
        if (e == cudaSuccess) e = cudaMemcpyAsync(f->tnf_raw + at * td, q->tnf_raw, (size_t)q->rows * td * 4, cudaMemcpyDeviceToDevice, ctx->stream);
        if (e == cudaSuccess) e = cudaMemcpyAsync(f->group_of_row + at, q->group_of_row, (size_t)q->rows * 4, cudaMemcpyDeviceToDevice, ctx->stream);
        at += q->rows;
    }
    if (e != cudaSuccess) { pg_features_free(ctx, f); return fail(ctx, PG_ERR_CUDA, std::string("pg_features_concat: ") + cudaGetErrorString(e)); }
    *out = f;
    return PG_OK;
} catch (...) { return caught("pg_features_concat"); }

extern "C" int pg_extract_features(pg_ctx* ctx, const pg_reads* host, const uint8_t* group_keep, int64_t n_groups, pg_features** out)
try {
    if (!out) return fail(ctx, PG_ERR_INVALID, "null out");
    *out = nullptr;
    int rc = pg_table_clear(ctx);
    if (rc) return rc;
    pg_batch* b = nullptr;
    if (use_buckets(ctx) && host && host->n_reads > 0 && host->n_bytes >= (1 << 20)) {
        rc = check_host_batch(ctx, host);
        if (!rc) rc = upload_and_count_pipelined(ctx, host, &b);
        if (rc) return rc;
    } else {
        rc = pg_batch_upload(ctx, host, &b);
        if (rc) return rc;
        rc = pg_count(ctx, b);
    }
    pg_features* f = nullptr;
    if (!rc) rc = pg_featurize(ctx, b, group_keep, n_groups, &f);
    if (!rc) rc = pg_normalize(ctx, f);
    pg_batch_free(ctx, b);
    if (rc) { if (f) pg_features_free(ctx, f); return rc; }
    *out = f;
    return PG_OK;
} catch (...) { return caught("pg_extract_features"); }

// ---------------------------------------------------------------------------
// synthetic reads (bench input)
// ---------------------------------------------------------------------------
extern "C" int pg_synth_generate2(pg_ctx* ctx, int64_t n_pairs, int32_t read_len, int64_t n_barcodes, const int64_t* d_bc_start,
                                  const int32_t* d_bc_genome, int64_t genome_len, int32_t frag_len, int32_t insert, double sub_rate,
                                  double n_rate, uint64_t seed, int64_t bc_base, int64_t pair_base, uint8_t* d_seq, int64_t* d_read_off,
                                  uint8_t* d_read_flag)
try {
    if (!ctx || n_pairs < 0 || read_len < 1 || n_barcodes < 1 || !d_bc_start || !d_bc_genome || !d_seq || !d_read_off || !d_read_flag)
        return fail(ctx, PG_ERR_INVALID, "pg_synth_generate: bad argument");
    if (insert < read_len || frag_len < insert || genome_len < frag_len) return fail(ctx, PG_ERR_INVALID, "need read_len <= insert <= frag_len <= genome_len");
    CK(cudaSetDevice(ctx->p.device));
    SynthParams S;
    S.n_pairs = n_pairs; S.read_len = read_len; S.n_barcodes = n_barcodes; S.bc_start = d_bc_start; S.bc_genome = d_bc_genome;
    S.genome_len = genome_len; S.frag_len = frag_len; S.insert = insert;
    S.sub_thresh = (uint32_t)std::min(4294967295.0, sub_rate * 4294967296.0);
    S.n_thresh = (uint32_t)std::min(4294967295.0, n_rate * 4294967296.0);
    S.seed = seed; S.bc_base = bc_base; S.pair_base = pair_base;
    synth_kernel<<<ctx->sm_count * 16, 256, 0, ctx->stream>>>(S, d_seq, d_read_off, d_read_flag);
    synth_flags_kernel<<<(int)((n_barcodes + 255) / 256), 256, 0, ctx->stream>>>(S, d_read_flag);
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(ctx->stream));
    return PG_OK;
} catch (...) { return caught("pg_synth_generate2"); }

extern "C" int pg_synth_generate(pg_ctx* ctx, int64_t n_pairs, int32_t read_len, int64_t n_barcodes, const int64_t* d_bc_start,
                                 const int32_t* d_bc_genome, int64_t genome_len, int32_t frag_len, int32_t insert, double sub_rate,
                                 double n_rate, uint64_t seed, uint8_t* d_seq, int64_t* d_read_off, uint8_t* d_read_flag)
try {
    return pg_synth_generate2(ctx, n_pairs, read_len, n_barcodes, d_bc_start, d_bc_genome, genome_len, frag_len, insert, sub_rate, n_rate, seed, 0, 0,
                              d_seq, d_read_off, d_read_flag);
} catch (...) { return caught("pg_synth_generate"); }

#include "ingest_api.cuh"
