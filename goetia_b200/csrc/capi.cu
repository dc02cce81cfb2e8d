// goetia_b200/csrc/capi.cu -- host side of the C ABI declared in include/goetia_b200.h.
//
// Owns all CUDA state: the storages' tables in HBM, a two-slot chunk pipeline (H2D copy of
// chunk n+1 overlaps the kernels of chunk n on a second stream), and device-resident packed
// batches.  No CPU fallback exists anywhere in this file: if CUDA is unusable every entry
// point fails with an error.
#include <cuda_runtime.h>

#include <algorithm>
#include <cctype>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/goetia_b200.h"
#include "kernels.cuh"
#include "bucket.cuh"
#include "sketch.cuh"

using namespace gt;

// ------------------------------------------------------------------------------------------
// errors
// ------------------------------------------------------------------------------------------
static thread_local std::string g_err;
static int fail(const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_err = buf;
    return -1;
}
#define CU(call)                                                                            \
    do {                                                                                    \
        cudaError_t e__ = (call);                                                           \
        if (e__ != cudaSuccess) {                                                           \
            fail("%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, __LINE__); \
            return -1;                                                                      \
        }                                                                                   \
    } while (0)
#define CUP(call)                                                                           \
    do {                                                                                    \
        cudaError_t e__ = (call);                                                           \
        if (e__ != cudaSuccess) {                                                           \
            fail("%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, __LINE__); \
            return nullptr;                                                                 \
        }                                                                                   \
    } while (0)

// ------------------------------------------------------------------------------------------
// context
// ------------------------------------------------------------------------------------------
struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    // grow-only; caller guarantees the owning stream is idle w.r.t. this buffer when growing
    int reserve(size_t n, cudaStream_t s) {
        if (n <= cap) return 0;
        if (p) {
            CU(cudaStreamSynchronize(s));
            CU(cudaFree(p));
            p = nullptr;
            cap = 0;
        }
        size_t want = n + n / 8 + 256;
        CU(cudaMalloc(&p, want));
        cap = want;
        return 0;
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
    template <class T> T* as() const { return static_cast<T*>(p); }
};

struct Slot {
    cudaStream_t stream = nullptr;
    cudaEvent_t packed = nullptr;    // recorded on `stream` when the chunk is packed
    cudaEvent_t consumed = nullptr;  // recorded on the compute stream when the chunk's buffers are free again
    bool consumed_pending = false;
    DevBuf ascii, words, offsets, flags, coarse, nmask, kcount, koff, partial, hits, out16, out64a, out64b, out8;
};

struct Context {
    bool ready = false;
    int device = -1;
    int sms = 0;
    cudaStream_t main = nullptr;      // compute stream (the library's own unless gt_set_compute_stream gave another)
    cudaStream_t own_main = nullptr;
    cudaStream_t apply = nullptr;     // k_apply of a full entry store runs here while the compute stream fills the other
    cudaStream_t own_apply = nullptr;
    bool apply_is_external = false;
    Slot slot[2];
    unsigned long long* d_scratch = nullptr;  // [0] k-mer total, [1] occupied, [2..7] spare
    DevBuf claim_keys, claim_ords;            // GT_MODE_EXACT claim map (compute stream only)
    unsigned long long* h_scratch = nullptr;  // pinned mirror
};
static Context g_ctx;
static std::mutex g_mu;
static unsigned long long g_launches = 0;  // kernels of this library launched so far (gt_launch_count)

// Optional per-kernel device timing (gt_profile_*): CUDA event pairs around the launches of the
// two insert kernels on the stream they run on; resolved lazily.  Off by default.
enum { PROF_BUCKET = 0, PROF_APPLY = 1, PROF_WALK = 2, PROF_REBUCKET = 3, PROF_APPLY_WIN = 4, PROF_KINDS = 5 };  // 3, 4: inside PROF_APPLY
struct ProfSpan { cudaEvent_t a, b; int kind; };
static bool g_prof_on = false;
static std::vector<ProfSpan> g_prof_open;
static std::vector<cudaEvent_t> g_prof_pool;
static double g_prof_ms[PROF_KINDS] = {0};
static unsigned long long g_prof_n[PROF_KINDS] = {0};
static cudaEvent_t prof_event() {
    cudaEvent_t e = nullptr;
    if (!g_prof_pool.empty()) { e = g_prof_pool.back(); g_prof_pool.pop_back(); }
    else cudaEventCreate(&e);
    return e;
}
struct ProfScope {  // brackets one launch
    cudaStream_t s; int kind; cudaEvent_t a = nullptr;
    ProfScope(int kind_, cudaStream_t s_) : s(s_), kind(kind_) {
        if (g_prof_on) { a = prof_event(); cudaEventRecord(a, s); }
    }
    ~ProfScope() {
        if (a) { cudaEvent_t b = prof_event(); cudaEventRecord(b, s); g_prof_open.push_back({a, b, kind}); }
    }
};
static void prof_resolve() {
    for (auto& sp : g_prof_open) {
        float ms = 0;
        if (cudaEventSynchronize(sp.b) == cudaSuccess && cudaEventElapsedTime(&ms, sp.a, sp.b) == cudaSuccess) {
            g_prof_ms[sp.kind] += ms;
            g_prof_n[sp.kind] += 1;
        }
        g_prof_pool.push_back(sp.a);
        g_prof_pool.push_back(sp.b);
    }
    g_prof_open.clear();
}

static int ensure_ctx() {
    if (g_ctx.ready) return 0;
    return fail("gt_init() has not been called");
}

// chunk size of the host-buffer pipeline (bases).  256 Mi bases = 256 MiB of ASCII per slot.
static uint64_t chunk_bases() {
    static uint64_t v = 0;
    if (!v) {
        const char* e = getenv("GT_CHUNK_BASES");
        v = e ? strtoull(e, nullptr, 10) : (256ull << 20);
        if (v < 1024) v = 1024;
    }
    return v;
}

extern "C" int gt_abi_version(void) { return GT_ABI_VERSION; }
extern "C" const char* gt_last_error(void) { return g_err.c_str(); }

extern "C" int gt_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

extern "C" int gt_init(int device) {
    std::lock_guard<std::mutex> lk(g_mu);
    if (g_ctx.ready) {
        if (g_ctx.device == device) return 0;
        return fail("gt_init: already initialised on device %d (one process drives one GPU)", g_ctx.device);
    }
    int n = 0;
    CU(cudaGetDeviceCount(&n));
    if (device < 0 || device >= n) return fail("gt_init: device %d out of range (%d visible)", device, n);
    CU(cudaSetDevice(device));
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10) return fail("gt_init: device %d is sm_%d%d; this library is built for sm_100a only", device, prop.major, prop.minor);
    g_ctx.sms = prop.multiProcessorCount;
    {
        // the producer kernels get the higher priority so that they keep their SM slots while the
        // (L2-bound, issue-light) apply kernel of the previous store runs next to them
        int lo = 0, hi = 0;
        CU(cudaDeviceGetStreamPriorityRange(&lo, &hi));
        CU(cudaStreamCreateWithPriority(&g_ctx.own_main, cudaStreamNonBlocking, hi));
        CU(cudaStreamCreateWithPriority(&g_ctx.own_apply, cudaStreamNonBlocking, lo));
        g_ctx.apply = g_ctx.own_apply;
    }
    g_ctx.main = g_ctx.own_main;
    for (auto& s : g_ctx.slot) {
        CU(cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking));
        CU(cudaEventCreateWithFlags(&s.packed, cudaEventDisableTiming));
        CU(cudaEventCreateWithFlags(&s.consumed, cudaEventDisableTiming));
    }
    CU(cudaMalloc(&g_ctx.d_scratch, 8 * sizeof(unsigned long long)));
    CU(cudaMemset(g_ctx.d_scratch, 0, 8 * sizeof(unsigned long long)));
    CU(cudaMallocHost(&g_ctx.h_scratch, 8 * sizeof(unsigned long long)));
    g_ctx.device = device;
    g_ctx.ready = true;
    return 0;
}

extern "C" int gt_synchronize(void) {
    if (ensure_ctx()) return -1;
    CU(cudaSetDevice(g_ctx.device));
    CU(cudaDeviceSynchronize());
    return 0;
}

// ------------------------------------------------------------------------------------------
// table sizing: get_n_primes_near_x, storage/storage.hh:146-190 (host only, one-off)
// ------------------------------------------------------------------------------------------
static bool is_prime_u64(uint64_t n) {
    if (n < 2) return false;
    if (n < 4) return true;
    if (n % 2 == 0) return false;
    for (uint64_t i = 3; i * i <= n; i += 2)
        if (n % i == 0) return false;
    return true;
}
extern "C" int gt_primes_near(uint32_t n, uint64_t x, uint64_t* out) {
    if (!out) return fail("gt_primes_near: out is NULL");
    if (x == 1) { out[0] = 1; return 1; }  // storage.hh:169-172
    if (x == 0) return 0;
    uint64_t i = x - 1;
    if (i % 2 == 0) { if (i == 0) return 0; --i; }
    uint32_t k = 0;
    while (k != n) {
        if (is_prime_u64(i)) out[k++] = i;
        if (i == 1) break;
        i -= 2;
    }
    return (int)k;
}

// ------------------------------------------------------------------------------------------
// storage
// ------------------------------------------------------------------------------------------
struct Pending;
static void pending_free(Pending* p);
static int pending_flush_sync(gt_storage* st);
static int pending_discard(gt_storage* st);
struct gt_storage {
    Pending* pend = nullptr;  // write-combining store of not-yet-applied blind inserts (bucket_host.inc)
    int kind = 0;
    int n = 0;
    uint64_t sizes[MAX_TABLES];
    uint64_t ref_bytes[MAX_TABLES];    // bytes the reference allocates
    uint64_t alloc_bytes[MAX_TABLES];  // bytes we allocate (multiple of 16, >= ref_bytes)
    TableSet ts;                       // ts.ptr[i] is indexed by GLOBAL slot (biased by own_lo in a sharded storage)
    unsigned long long* d_n_unique = nullptr;
    // sharding (one process per GPU): this rank holds slots [own_lo, own_hi) of table i in slab[i]
    int rank = 0, world = 1;
    uint64_t own_lo[MAX_TABLES], own_hi[MAX_TABLES];
    void* slab[MAX_TABLES];
    uint64_t hint_kmers = 0;           // gt_storage_hint_kmers: bound for the next device-resident batch (0 = none)
};

static uint64_t ref_table_bytes(int kind, uint64_t size) {
    return kind == GT_STORAGE_BIT ? size / 8 + 1 : kind == GT_STORAGE_BYTE ? size : size / 2 + 1;
}

// bytes of the slots [lo, hi) of a table (lo is a multiple of the slice size, so of 8 and 2)
static uint64_t range_bytes(int kind, uint64_t lo, uint64_t hi, bool last) {
    if (hi <= lo) return 0;
    uint64_t n = hi - lo;
    if (kind == GT_STORAGE_BIT) return last ? hi / 8 + 1 - lo / 8 : n / 8;
    if (kind == GT_STORAGE_BYTE) return n;
    return last ? hi / 2 + 1 - lo / 2 : n / 2;
}
static int slots_per_word(int kind) { return kind == GT_STORAGE_BIT ? 32 : kind == GT_STORAGE_BYTE ? 4 : 8; }

// gt_storage_hint_kmers: the caller's bound for the next device-resident batch, used once
static uint64_t take_kmer_hint(gt_storage* st, uint64_t n_bases) {
    const uint64_t h = st->hint_kmers;
    st->hint_kmers = 0;
    return h && h < n_bases ? h : n_bases;
}

struct PlanHost;
static int make_plan(int kind, const uint64_t* sizes, int n, int world, uint64_t budget, int slice_log2_bytes, PlanHost& P);

static gt_storage* storage_create(int kind, const uint64_t* tablesizes, int n_tables, int rank, int world,
                                  const uint64_t* own_lo, const uint64_t* own_hi) {
    if (ensure_ctx()) return nullptr;
    if (kind < 0 || kind > 2) { fail("gt_storage_create: unknown storage kind %d", kind); return nullptr; }
    if (!tablesizes || n_tables < 1 || n_tables > MAX_TABLES) {
        fail("gt_storage_create: n_tables must be 1..%d", MAX_TABLES);
        return nullptr;
    }
    CUP(cudaSetDevice(g_ctx.device));
    gt_storage* st = new gt_storage();
    st->kind = kind;
    st->n = n_tables;
    st->rank = rank;
    st->world = world;
    memset(&st->ts, 0, sizeof st->ts);
    memset(st->slab, 0, sizeof st->slab);
    st->ts.n = n_tables;
    st->ts.kind = kind;
    for (int i = 0; i < n_tables; ++i) {
        uint64_t d = tablesizes[i];
        if (d == 0 || d > (1ull << 63)) {
            fail("gt_storage_create: table size %llu out of range", (unsigned long long)d);
            gt_storage_destroy(st);
            return nullptr;
        }
        st->sizes[i] = d;
        st->own_lo[i] = own_lo ? own_lo[i] : 0;
        st->own_hi[i] = own_hi ? own_hi[i] : d;
        // bytes of the part held here, laid out exactly as the reference lays out that part of its table
        st->ref_bytes[i] = world == 1 ? ref_table_bytes(kind, d) : range_bytes(kind, st->own_lo[i], st->own_hi[i], st->own_hi[i] == d);
        st->alloc_bytes[i] = (st->ref_bytes[i] + 15) / 16 * 16 + 16;
        st->ts.size[i] = d;
        st->ts.magic[i] = ~0ull / d;
        void* p = nullptr;
        cudaError_t e = cudaMalloc(&p, st->alloc_bytes[i]);
        if (e != cudaSuccess) {
            fail("gt_storage_create: cudaMalloc(%llu) failed: %s", (unsigned long long)st->alloc_bytes[i], cudaGetErrorString(e));
            gt_storage_destroy(st);
            return nullptr;
        }
        st->slab[i] = p;
        // kernels index tables by global slot: bias the base by the first slot held here
        st->ts.ptr[i] = reinterpret_cast<uint32_t*>(reinterpret_cast<uintptr_t>(p) -
                                                    (uintptr_t)(st->own_lo[i] / slots_per_word(kind)) * 4u);
    }
    if (cudaMalloc(&st->d_n_unique, sizeof(unsigned long long)) != cudaSuccess) {
        fail("gt_storage_create: cudaMalloc(counter) failed");
        gt_storage_destroy(st);
        return nullptr;
    }
    if (gt_storage_reset(st) != 0) { gt_storage_destroy(st); return nullptr; }
    return st;
}

extern "C" gt_storage* gt_storage_create(int kind, const uint64_t* tablesizes, int n_tables) {
    return storage_create(kind, tablesizes, n_tables, 0, 1, nullptr, nullptr);
}

extern "C" void gt_storage_destroy(gt_storage* st) {
    if (!st) return;
    cudaDeviceSynchronize();
    for (int i = 0; i < st->n; ++i)
        if (st->slab[i]) cudaFree(st->slab[i]);
    if (st->d_n_unique) cudaFree(st->d_n_unique);
    pending_free(st->pend);
    delete st;
}

extern "C" int gt_storage_reset(gt_storage* st) {
    if (ensure_ctx()) return -1;
    if (!st) return fail("gt_storage_reset: NULL storage");
    CU(cudaDeviceSynchronize());
    if (pending_discard(st)) return -1;
    for (int i = 0; i < st->n; ++i) CU(cudaMemsetAsync(st->slab[i], 0, st->alloc_bytes[i], g_ctx.main));
    CU(cudaMemsetAsync(st->d_n_unique, 0, sizeof(unsigned long long), g_ctx.main));
    CU(cudaStreamSynchronize(g_ctx.main));
    return 0;
}

extern "C" int gt_storage_kind(const gt_storage* st) { return st ? st->kind : -1; }
extern "C" int gt_storage_n_tables(const gt_storage* st) { return st ? st->n : -1; }
extern "C" int gt_storage_tablesizes(const gt_storage* st, uint64_t* out) {
    if (!st || !out) return fail("gt_storage_tablesizes: NULL argument");
    for (int i = 0; i < st->n; ++i) out[i] = st->sizes[i];
    return st->n;
}
extern "C" uint64_t gt_storage_table_bytes(const gt_storage* st, int i) {
    if (!st || i < 0 || i >= st->n) return 0;
    return st->ref_bytes[i];
}
extern "C" void* gt_storage_device_table(gt_storage* st, int i) {
    if (!st || i < 0 || i >= st->n) return nullptr;
    if (pending_flush_sync(st)) return nullptr;
    return st->slab[i];
}

extern "C" int gt_storage_download_table(gt_storage* st, int i, uint8_t* host_dst) {
    if (ensure_ctx()) return -1;
    if (!st || !host_dst || i < 0 || i >= st->n) return fail("gt_storage_download_table: bad argument");
    if (pending_flush_sync(st)) return -1;
    CU(cudaDeviceSynchronize());
    CU(cudaMemcpy(host_dst, st->slab[i], st->ref_bytes[i], cudaMemcpyDeviceToHost));
    return 0;
}

extern "C" int gt_storage_upload_table(gt_storage* st, int i, const uint8_t* host_src) {
    if (ensure_ctx()) return -1;
    if (!st || !host_src || i < 0 || i >= st->n) return fail("gt_storage_upload_table: bad argument");
    if (pending_flush_sync(st)) return -1;
    CU(cudaDeviceSynchronize());
    CU(cudaMemset(st->slab[i], 0, st->alloc_bytes[i]));
    CU(cudaMemcpy(st->slab[i], host_src, st->ref_bytes[i], cudaMemcpyHostToDevice));
    return 0;
}

static int grid_for(uint64_t items, int per_block, int blocks_per_sm) {
    uint64_t want = (items + per_block - 1) / per_block;
    uint64_t cap = (uint64_t)g_ctx.sms * blocks_per_sm;
    if (want < 1) want = 1;
    return (int)std::min(want, cap);
}

extern "C" int gt_storage_stats(gt_storage* st, uint64_t* n_unique, uint64_t* n_occupied) {
    if (ensure_ctx()) return -1;
    if (!st) return fail("gt_storage_stats: NULL storage");
    if (pending_flush_sync(st)) return -1;
    CU(cudaDeviceSynchronize());
    cudaStream_t s = g_ctx.main;
    unsigned long long* d_occ = g_ctx.d_scratch + 1;
    CU(cudaMemsetAsync(d_occ, 0, sizeof(unsigned long long), s));
    // only the slots of the reference's table count; the padding words are always zero
    uint64_t n_words = st->alloc_bytes[0] / 4;
    int grid = grid_for(n_words, 256 * 8, 8);
    const uint32_t* t0 = static_cast<const uint32_t*>(st->slab[0]);
    if (st->kind == 0) k_count_occupied<0><<<grid, 256, 0, s>>>(t0, n_words, st->sizes[0], d_occ);
    else if (st->kind == 1) k_count_occupied<1><<<grid, 256, 0, s>>>(t0, n_words, st->sizes[0], d_occ);
    else k_count_occupied<2><<<grid, 256, 0, s>>>(t0, n_words, st->sizes[0], d_occ); ++g_launches;
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(g_ctx.h_scratch + 1, d_occ, sizeof(unsigned long long), cudaMemcpyDeviceToHost, s));
    CU(cudaMemcpyAsync(g_ctx.h_scratch + 2, st->d_n_unique, sizeof(unsigned long long), cudaMemcpyDeviceToHost, s));
    CU(cudaStreamSynchronize(s));
    if (n_occupied) *n_occupied = g_ctx.h_scratch[1];
    if (n_unique) *n_unique = g_ctx.h_scratch[2];
    return 0;
}

// Position-weighted checksum of table i (kernels.cuh, k_checksum).  Linear in the table's words and indexed by
// GLOBAL word position, so the checksums of the parts of a sharded storage add up (mod 2^64) to the checksum of
// the whole table.
extern "C" int gt_storage_checksum(gt_storage* st, int i, uint64_t* out) {
    if (ensure_ctx()) return -1;
    if (!st || !out || i < 0 || i >= st->n) return fail("gt_storage_checksum: bad argument");
    if (pending_flush_sync(st)) return -1;
    CU(cudaDeviceSynchronize());
    cudaStream_t s = g_ctx.main;
    unsigned long long* d_sum = g_ctx.d_scratch + 6;
    CU(cudaMemsetAsync(d_sum, 0, sizeof(unsigned long long), s));
    const uint64_t n_words = st->alloc_bytes[i] / 4;
    const uint64_t word0 = st->own_lo[i] / (uint64_t)slots_per_word(st->kind);
    k_checksum<<<grid_for(n_words, 256 * 8, 8), 256, 0, s>>>(static_cast<const uint32_t*>(st->slab[i]), n_words, word0, d_sum); ++g_launches;
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(g_ctx.h_scratch + 6, d_sum, sizeof(unsigned long long), cudaMemcpyDeviceToHost, s));
    CU(cudaStreamSynchronize(s));
    *out = g_ctx.h_scratch[6];
    return 0;
}

extern "C" int gt_storage_set_n_unique(gt_storage* st, uint64_t n_unique) {
    if (ensure_ctx()) return -1;
    if (!st) return fail("gt_storage_set_n_unique: NULL storage");
    CU(cudaDeviceSynchronize());
    unsigned long long v = n_unique;
    CU(cudaMemcpy(st->d_n_unique, &v, sizeof v, cudaMemcpyHostToDevice));
    return 0;
}

extern "C" int gt_storage_update_from(gt_storage* dst, const gt_storage* src) {
    if (ensure_ctx()) return -1;
    if (!dst || !src) return fail("gt_storage_update_from: NULL storage");
    if (dst->world > 1 || src->world > 1) return fail("gt_storage_update_from: not available on a sharded storage");
    if (dst->kind != GT_STORAGE_BIT || src->kind != GT_STORAGE_BIT)
        return fail("gt_storage_update_from: only BitStorage can be unioned (bitstorage.cc:103-137)");
    if (dst->n != src->n || memcmp(dst->sizes, src->sizes, sizeof(uint64_t) * dst->n) != 0)
        return fail("both nodegraphs must have same table sizes");
    if (pending_flush_sync(dst) || pending_flush_sync(const_cast<gt_storage*>(src))) return -1;
    CU(cudaDeviceSynchronize());
    for (int i = 0; i < dst->n; ++i) {
        uint64_t n_words = dst->alloc_bytes[i] / 4;
        k_or_tables<<<grid_for(n_words, 256 * 4, 8), 256, 0, g_ctx.main>>>(dst->ts.ptr[i], src->ts.ptr[i], n_words); ++g_launches;
    }
    CU(cudaGetLastError());
    CU(cudaStreamSynchronize(g_ctx.main));
    return 0;
}

// ------------------------------------------------------------------------------------------
// kernel dispatch
// ------------------------------------------------------------------------------------------
template <int OP, int KIND, bool CAN, bool TRACK, int NT>
static int launch_walk_t(const WalkArgs& a, const TableSet& ts, cudaStream_t s) {
    auto kern = k_walk<OP, KIND, CAN, TRACK, NT>;
    static int occ = 0;
    static size_t occ_smem = 0;  // the occupancy depends on K through the halo: cached per shared-memory size
    const int halo_words = ((a.K - 1 + 31) >> 5) + 1;
    const size_t smem = (16 + TILE_THREADS + halo_words) * sizeof(uint64_t);
    if (smem > 48 * 1024) CU(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (!occ || occ_smem != smem) {
        CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, TILE_THREADS, smem));
        if (occ < 1) occ = 1;
        occ_smem = smem;
    }
    uint64_t n_tiles = (a.n_bases + TILE_POS - 1) / TILE_POS;
    if (a.tile_n) n_tiles = std::min(n_tiles - std::min(n_tiles, a.tile_lo), a.tile_n);
    if (n_tiles == 0) return 0;
    int grid = (int)std::min<uint64_t>(n_tiles, (uint64_t)g_ctx.sms * occ);
    {
        ProfScope ps(PROF_WALK, s);
        kern<<<grid, TILE_THREADS, smem, s>>>(a, ts); ++g_launches;
    }
    CU(cudaGetLastError());
    return 0;
}

template <int OP, int KIND, bool CAN, bool TRACK>
static int launch_walk_nt(const WalkArgs& a, const TableSet& ts, cudaStream_t s) {
    if (ts.n == 4) return launch_walk_t<OP, KIND, CAN, TRACK, 4>(a, ts, s);
    return launch_walk_t<OP, KIND, CAN, TRACK, 0>(a, ts, s);
}
template <int OP, int KIND, bool TRACK>
static int launch_walk_can(int shifter, const WalkArgs& a, const TableSet& ts, cudaStream_t s) {
    if (shifter == GT_SHIFTER_CAN) return launch_walk_nt<OP, KIND, true, TRACK>(a, ts, s);
    return launch_walk_nt<OP, KIND, false, TRACK>(a, ts, s);
}
template <int OP, bool TRACK>
static int launch_walk_kind(int shifter, const WalkArgs& a, const TableSet& ts, cudaStream_t s) {
    switch (ts.kind) {
        case 0: return launch_walk_can<OP, 0, TRACK>(shifter, a, ts, s);
        case 1: return launch_walk_can<OP, 1, TRACK>(shifter, a, ts, s);
        default: return launch_walk_can<OP, 2, TRACK>(shifter, a, ts, s);
    }
}
static int launch_hash(int shifter, const WalkArgs& a, cudaStream_t s) {
    TableSet ts;
    memset(&ts, 0, sizeof ts);
    if (shifter == GT_SHIFTER_CAN) return launch_walk_t<OP_HASH, 0, true, false, 0>(a, ts, s);
    return launch_walk_t<OP_HASH, 0, false, false, 0>(a, ts, s);
}

// ------------------------------------------------------------------------------------------
// device-resident batch
// ------------------------------------------------------------------------------------------
struct gt_batch {
    uint64_t n_reads = 0, n_bases = 0, n_words = 0, n_words_alloc = 0, base0 = 0;
    uint64_t* d_words = nullptr;
    uint64_t* d_offsets = nullptr;
    uint8_t* d_flags = nullptr;
    uint32_t* d_coarse = nullptr;
    bool owns = false;
    int cached_K = -1;
    int64_t cached_kmers = 0;
};

static uint64_t halo_alloc_words() { return 2052; }  // covers K up to 65535 plus the +1 vector lane

// pack ASCII already on the device (d_ascii, d_offsets are device pointers; offsets absolute,
// base0 = value of offsets[0]) into b's buffers on stream s.
static int pack_on_device(const uint8_t* d_ascii, const uint64_t* d_offsets, uint64_t n_reads, uint64_t n_bases,
                          uint64_t base0, uint64_t* d_words, uint64_t n_words_alloc, uint8_t* d_flags,
                          uint32_t* d_coarse, cudaStream_t s, uint32_t* d_nmask = nullptr) {
    uint64_t n_words = (n_bases + 31) / 32;
    CU(cudaMemsetAsync(d_flags, 0, n_reads ? n_reads : 1, s));
    if (n_words_alloc > n_words) CU(cudaMemsetAsync(d_words + n_words, 0, (n_words_alloc - n_words) * 8, s));
    if (d_nmask && n_words_alloc > n_words) CU(cudaMemsetAsync(d_nmask + n_words, 0, (n_words_alloc - n_words) * 4, s));
    if (n_words) {
        k_pack<<<grid_for(n_words, 256, 16), 256, 0, s>>>(d_ascii, n_bases, d_offsets, n_reads, base0, d_words, n_words, d_flags, d_nmask); ++g_launches;
        CU(cudaGetLastError());
    }
    if (n_reads) {
        k_coarse<<<grid_for(n_reads, 256, 16), 256, 0, s>>>(d_offsets, n_reads, base0, d_coarse); ++g_launches;
        CU(cudaGetLastError());
    }
    return 0;
}

static WalkArgs make_args(const gt_batch& b, int K) {
    WalkArgs a;
    memset(&a, 0, sizeof a);
    a.words = b.d_words;
    a.n_words_alloc = b.n_words_alloc;
    a.offsets = b.d_offsets;
    a.flags = b.d_flags;
    a.coarse = b.d_coarse;
    a.base0 = b.base0;
    a.n_reads = b.n_reads;
    a.n_bases = b.n_bases;
    a.K = K;
    return a;
}

#include "bucket_host.inc"

// Stage one chunk of host reads [r0, r1) into slot `sl` and pack it.  Fills `view`.
static int stage_chunk(Slot& sl, const char* bases, const uint64_t* offsets, uint64_t r0, uint64_t r1, gt_batch& view,
                       bool want_nmask = false) {
    cudaStream_t s = sl.stream;
    uint64_t base0 = offsets[r0], n_bases = offsets[r1] - base0, n_reads = r1 - r0;
    uint64_t n_words = (n_bases + 31) / 32, n_words_alloc = n_words + halo_alloc_words();
    if (sl.ascii.reserve(n_bases + 64, s)) return -1;
    if (sl.words.reserve(n_words_alloc * 8, s)) return -1;
    if (sl.offsets.reserve((n_reads + 1) * 8, s)) return -1;
    if (sl.flags.reserve(n_reads + 1, s)) return -1;
    if (sl.coarse.reserve(((n_bases >> COARSE_SHIFT) + 2) * 4, s)) return -1;
    if (want_nmask && sl.nmask.reserve(n_words_alloc * 4, s)) return -1;
    if (n_bases) CU(cudaMemcpyAsync(sl.ascii.p, bases + base0, n_bases, cudaMemcpyHostToDevice, s));
    CU(cudaMemcpyAsync(sl.offsets.p, offsets + r0, (n_reads + 1) * 8, cudaMemcpyHostToDevice, s));
    if (pack_on_device(sl.ascii.as<uint8_t>(), sl.offsets.as<uint64_t>(), n_reads, n_bases, base0, sl.words.as<uint64_t>(),
                       n_words_alloc, sl.flags.as<uint8_t>(), sl.coarse.as<uint32_t>(), s,
                       want_nmask ? sl.nmask.as<uint32_t>() : nullptr))
        return -1;
    view = gt_batch();
    view.n_reads = n_reads;
    view.n_bases = n_bases;
    view.n_words = n_words;
    view.n_words_alloc = n_words_alloc;
    view.base0 = base0;
    view.d_words = sl.words.as<uint64_t>();
    view.d_offsets = sl.offsets.as<uint64_t>();
    view.d_flags = sl.flags.as<uint8_t>();
    view.d_coarse = sl.coarse.as<uint32_t>();
    return 0;
}

static int check_reads(const char* who, const char* bases, const uint64_t* offsets, uint64_t n_reads, int K) {
    if (ensure_ctx()) return -1;
    if (K < 1 || K > 65535) return fail("%s: K=%d out of range (1..65535)", who, K);
    if (n_reads && (!bases || !offsets)) return fail("%s: NULL bases/offsets", who);
    if (n_reads >= (1ull << 32)) return fail("%s: more than 2^32-1 reads in one call", who);
    return 0;
}

// split reads into chunks of <= chunk_bases() (always at least one read per chunk)
static void chunk_ranges(const uint64_t* offsets, uint64_t n_reads, std::vector<uint64_t>& cuts) {
    cuts.clear();
    cuts.push_back(0);
    uint64_t lim = chunk_bases();
    uint64_t r = 0;
    while (r < n_reads) {
        uint64_t start = offsets[r];
        // first read whose end exceeds start+lim
        const uint64_t* it = std::upper_bound(offsets + r + 1, offsets + n_reads + 1, start + lim);
        uint64_t r1 = (uint64_t)(it - offsets) - 1;
        if (r1 <= r) r1 = r + 1;
        cuts.push_back(r1);
        r = r1;
    }
}

static int validate_offsets(const char* who, const uint64_t* offsets, uint64_t n_reads) {
    for (uint64_t r = 0; r < n_reads; ++r)
        if (offsets[r + 1] < offsets[r]) return fail("%s: offsets must be non-decreasing (read %llu)", who, (unsigned long long)r);
    return 0;
}

// ------------------------------------------------------------------------------------------
// gt_batch (resident packed reads)
// ------------------------------------------------------------------------------------------
static gt_batch* batch_alloc(uint64_t n_reads, uint64_t n_bases) {
    gt_batch* b = new gt_batch();
    b->owns = true;
    b->n_reads = n_reads;
    b->n_bases = n_bases;
    b->n_words = (n_bases + 31) / 32;
    b->n_words_alloc = b->n_words + halo_alloc_words();
    if (cudaMalloc(&b->d_words, b->n_words_alloc * 8) != cudaSuccess ||
        cudaMalloc(&b->d_offsets, (n_reads + 1) * 8) != cudaSuccess ||
        cudaMalloc(&b->d_flags, n_reads + 1) != cudaSuccess ||
        cudaMalloc(&b->d_coarse, ((n_bases >> COARSE_SHIFT) + 2) * 4) != cudaSuccess) {
        fail("gt_batch: out of device memory (%llu bases)", (unsigned long long)n_bases);
        gt_batch_destroy(b);
        return nullptr;
    }
    return b;
}

extern "C" void gt_batch_destroy(gt_batch* b) {
    if (!b) return;
    if (b->owns) {
        cudaDeviceSynchronize();
        if (b->d_words) cudaFree(b->d_words);
        if (b->d_offsets) cudaFree(b->d_offsets);
        if (b->d_flags) cudaFree(b->d_flags);
        if (b->d_coarse) cudaFree(b->d_coarse);
    }
    delete b;
}

extern "C" gt_batch* gt_batch_pack_dev(const void* d_bases, const void* d_offsets, uint64_t n_reads, uint64_t n_bases) {
    if ((reinterpret_cast<uintptr_t>(d_bases) & 15) || (reinterpret_cast<uintptr_t>(d_offsets) & 7)) {
        fail("gt_batch_pack_dev: d_bases must be 16-byte aligned and d_offsets 8-byte aligned");
        return nullptr;
    }
    if (ensure_ctx()) return nullptr;
    if (n_reads >= (1ull << 32)) { fail("gt_batch_pack_dev: too many reads"); return nullptr; }
    gt_batch* b = batch_alloc(n_reads, n_bases);
    if (!b) return nullptr;
    cudaStream_t s = g_ctx.main;
    // offsets of a device batch are required to start at 0
    if (cudaMemcpyAsync(b->d_offsets, d_offsets, (n_reads + 1) * 8, cudaMemcpyDeviceToDevice, s) != cudaSuccess ||
        pack_on_device(static_cast<const uint8_t*>(d_bases), b->d_offsets, n_reads, n_bases, 0, b->d_words, b->n_words_alloc,
                       b->d_flags, b->d_coarse, s) != 0 ||
        cudaStreamSynchronize(s) != cudaSuccess) {
        if (g_err.empty()) fail("gt_batch_pack_dev: CUDA failure: %s", cudaGetErrorString(cudaGetLastError()));
        gt_batch_destroy(b);
        return nullptr;
    }
    return b;
}

extern "C" gt_batch* gt_batch_pack(const char* bases, const uint64_t* offsets, uint64_t n_reads) {
    if (check_reads("gt_batch_pack", bases, offsets, n_reads, 1)) return nullptr;
    if (validate_offsets("gt_batch_pack", offsets, n_reads)) return nullptr;
    uint64_t base0 = n_reads ? offsets[0] : 0;
    uint64_t n_bases = n_reads ? offsets[n_reads] - base0 : 0;
    gt_batch* b = batch_alloc(n_reads, n_bases);
    if (!b) return nullptr;
    // stage the ASCII through the pipeline slots in chunks, packing straight into the batch
    std::vector<uint64_t> rebased(n_reads + 1);
    for (uint64_t r = 0; r <= n_reads; ++r) rebased[r] = offsets[r] - base0;
    bool ok = cudaMemcpy(b->d_offsets, rebased.data(), (n_reads + 1) * 8, cudaMemcpyHostToDevice) == cudaSuccess;
    ok = ok && cudaMemset(b->d_flags, 0, n_reads + 1) == cudaSuccess;
    ok = ok && cudaMemset(b->d_words + b->n_words, 0, (b->n_words_alloc - b->n_words) * 8) == cudaSuccess;
    // the copies / memsets above ran on the legacy stream; the slot streams are non-blocking, so order them explicitly
    ok = ok && cudaDeviceSynchronize() == cudaSuccess;
    // chunk on 32-base boundaries so each chunk fills whole words
    uint64_t lim = chunk_bases() & ~31ull;
    int which = 0;
    for (uint64_t p = 0; ok && p < n_bases; p += lim, which ^= 1) {
        Slot& sl = g_ctx.slot[which];
        uint64_t nb = std::min(lim, n_bases - p);
        if (sl.ascii.reserve(nb + 64, sl.stream)) { ok = false; break; }
        ok = ok && cudaMemcpyAsync(sl.ascii.p, bases + base0 + p, nb, cudaMemcpyHostToDevice, sl.stream) == cudaSuccess;
        uint64_t nw = (nb + 31) / 32;
        k_pack<<<grid_for(nw, 256, 16), 256, 0, sl.stream>>>(sl.ascii.as<uint8_t>(), nb, b->d_offsets, n_reads, p,
                                                             b->d_words + p / 32, nw, b->d_flags, nullptr); ++g_launches;
        ok = ok && cudaGetLastError() == cudaSuccess;
    }
    if (ok && n_reads) {
        k_coarse<<<grid_for(n_reads, 256, 16), 256, 0, g_ctx.main>>>(b->d_offsets, n_reads, 0, b->d_coarse); ++g_launches;
        ok = cudaGetLastError() == cudaSuccess;
    }
    ok = ok && cudaDeviceSynchronize() == cudaSuccess;
    if (!ok) {
        if (g_err.empty()) fail("gt_batch_pack: CUDA failure: %s", cudaGetErrorString(cudaGetLastError()));
        gt_batch_destroy(b);
        return nullptr;
    }
    return b;
}

extern "C" uint64_t gt_batch_n_reads(const gt_batch* b) { return b ? b->n_reads : 0; }
extern "C" uint64_t gt_batch_n_bases(const gt_batch* b) { return b ? b->n_bases : 0; }

// k-mer total (+ optional per-read counts/status) of a batch on stream s; synchronises s.
static int64_t batch_kmers(const gt_batch& b, int K, uint64_t* d_kcount, uint8_t* d_status, cudaStream_t s) {
    unsigned long long* d_tot = g_ctx.d_scratch;
    CU(cudaMemsetAsync(d_tot, 0, sizeof(unsigned long long), s));
    if (b.n_reads) {
        k_kmer_counts<<<grid_for(b.n_reads, 256, 16), 256, 0, s>>>(b.d_offsets, b.n_reads, K, b.d_flags, d_kcount, d_status, d_tot); ++g_launches;
        CU(cudaGetLastError());
    }
    CU(cudaMemcpyAsync(g_ctx.h_scratch, d_tot, sizeof(unsigned long long), cudaMemcpyDeviceToHost, s));
    CU(cudaStreamSynchronize(s));
    return (int64_t)g_ctx.h_scratch[0];
}

extern "C" int64_t gt_batch_n_kmers(gt_batch* b, int K) {
    if (ensure_ctx()) return -1;
    if (!b) return fail("gt_batch_n_kmers: NULL batch");
    if (b->cached_K == K) return b->cached_kmers;
    int64_t n = batch_kmers(*b, K, nullptr, nullptr, g_ctx.main);
    if (n >= 0) { b->cached_K = K; b->cached_kmers = n; }
    return n;
}

extern "C" int gt_batch_status(gt_batch* b, int K, uint8_t* status) {
    if (ensure_ctx()) return -1;
    if (!b || !status) return fail("gt_batch_status: NULL argument");
    if (b->n_reads == 0) return 0;
    Slot& sl = g_ctx.slot[0];
    if (sl.out8.reserve(b->n_reads, sl.stream)) return -1;
    if (batch_kmers(*b, K, nullptr, sl.out8.as<uint8_t>(), sl.stream) < 0) return -1;
    CU(cudaMemcpy(status, sl.out8.p, b->n_reads, cudaMemcpyDeviceToHost));
    return 0;
}

static int check_mode(const char* who, int mode) {
    if (mode != GT_MODE_BLIND && mode != GT_MODE_FAST && mode != GT_MODE_EXACT) return fail("%s: unknown mode %d", who, mode);
    return 0;
}

// GT_MODE_EXACT claim map for up to `claims` posts, emptied on stream s (kernels.cuh, ClaimMap).
static uint64_t exact_log2_cap() { return std::min<uint64_t>(32, std::max<uint64_t>(10, env_u64("GT_EXACT_LOG2_CAP", 27))); }
static int exact_map(uint64_t claims, cudaStream_t s, ClaimMap& m) {
    uint64_t cap = 1024;
    while (cap < 2 * claims) cap <<= 1;
    if (g_ctx.claim_keys.reserve(cap * 8, s) || g_ctx.claim_ords.reserve(cap * 4, s)) return -1;
    CU(cudaMemsetAsync(g_ctx.claim_keys.p, 0xFF, cap * 8, s));
    CU(cudaMemsetAsync(g_ctx.claim_ords.p, 0xFF, cap * 4, s));
    m.keys = g_ctx.claim_keys.as<unsigned long long>();
    m.ords = g_ctx.claim_ords.as<uint32_t>();
    m.mask = cap - 1;
    return 0;
}

static int launch_insert_unordered(gt_storage* st, int shifter, const gt_batch& b, int K, int mode, uint64_t* d_n_new, cudaStream_t s);
// the direct path: k_walk applies every update to the tables itself
static int launch_insert(gt_storage* st, int shifter, const gt_batch& b, int K, int mode, uint64_t* d_n_new, cudaStream_t s) {
    if (st->world > 1) return fail("a sharded storage takes GT_MODE_BLIND inserts through its exchange only (attach it first)");
    if (direct_begin(st, s)) return -1;
    if (launch_insert_unordered(st, shifter, b, K, mode, d_n_new, s)) return -1;
    return direct_end(st, s);
}
static int launch_insert_unordered(gt_storage* st, int shifter, const gt_batch& b, int K, int mode, uint64_t* d_n_new, cudaStream_t s) {
    WalkArgs a = make_args(b, K);
    a.n_unique = st->d_n_unique;
    a.n_new = d_n_new;
    if (mode == GT_MODE_BLIND) return launch_walk_kind<OP_INSERT, false>(shifter, a, st->ts, s);
    if (mode == GT_MODE_FAST) return launch_walk_kind<OP_INSERT, true>(shifter, a, st->ts, s);
    // GT_MODE_EXACT: ranges of tiles in serial order, each claimed (pass 1) and then inserted (pass 2);
    // a range is as long as the claim map allows (every k-mer may post one claim per table)
    const uint64_t n_tiles = (b.n_bases + TILE_POS - 1) / TILE_POS;
    const uint64_t per = std::max<uint64_t>(1, ((1ull << exact_log2_cap()) / 2) / ((uint64_t)TILE_POS * st->n));
    for (uint64_t t0 = 0; t0 < n_tiles; t0 += per) {
        a.tile_lo = t0;
        a.tile_n = std::min(per, n_tiles - t0);
        const uint64_t positions = std::min<uint64_t>(a.tile_n * TILE_POS, b.n_bases - t0 * TILE_POS);
        if (exact_map(positions * st->n, s, a.claims)) return -1;
        if (launch_walk_kind<OP_CLAIM, false>(shifter, a, st->ts, s)) return -1;
        if (launch_walk_kind<OP_INSERT_EXACT, false>(shifter, a, st->ts, s)) return -1;
    }
    return 0;
}

extern "C" int64_t gt_insert_batch(gt_storage* st, int shifter, int K, gt_batch* b, int mode, void* stream) {
    if (ensure_ctx()) return -1;
    if (!st || !b) return fail("gt_insert_batch: NULL argument");
    if (K < 1 || K > 65535) return fail("gt_insert_batch: K=%d out of range", K);
    if (check_mode("gt_insert_batch", mode)) return -1;
    int64_t n = gt_batch_n_kmers(b, K);
    if (n < 0) return -1;
    cudaStream_t s = stream ? static_cast<cudaStream_t>(stream) : g_ctx.main;
    if (bucket_usable(st, mode, K, (uint64_t)n, (uint64_t)n)) {
        // K1b runs on the compute stream, ordered after what the caller queued on `stream`
        if (s != g_ctx.main) {
            CU(cudaEventRecord(g_ctx.slot[0].packed, s));
            CU(cudaStreamWaitEvent(g_ctx.main, g_ctx.slot[0].packed, 0));
        }
        if (bucket_insert(st, shifter, *b, K, (uint64_t)n)) return -1;
        return n;
    }
    if (mode != GT_MODE_BLIND && pending_flush_sync(st)) return -1;
    if (launch_insert(st, shifter, *b, K, mode, nullptr, s)) return -1;
    return n;
}

// ------------------------------------------------------------------------------------------
// host-buffer batch members
// ------------------------------------------------------------------------------------------
extern "C" int64_t gt_insert_sequences(gt_storage* st, int shifter, int K, const char* bases, const uint64_t* offsets,
                                        uint64_t n_reads, int mode, uint64_t* n_new_per_read, uint8_t* status) {
    if (check_reads("gt_insert_sequences", bases, offsets, n_reads, K)) return -1;
    if (!st) return fail("gt_insert_sequences: NULL storage");
    if (check_mode("gt_insert_sequences", mode)) return -1;
    if (n_new_per_read && mode == GT_MODE_BLIND) return fail("gt_insert_sequences: n_new_per_read needs GT_MODE_FAST or GT_MODE_EXACT");
    if (n_reads == 0) return 0;
    CU(cudaSetDevice(g_ctx.device));
    if (mode != GT_MODE_BLIND && pending_flush_sync(st)) return -1;  // is_new must see every earlier insert
    if (offsets[n_reads] < offsets[0]) return fail("gt_insert_sequences: offsets must be non-decreasing");
    std::vector<uint64_t> cuts;
    chunk_ranges(offsets, n_reads, cuts);
    // cheap estimate of the call's k-mers (exact when every read is >= K); the per-read pass over
    // the offsets is done chunk by chunk below so that it overlaps the GPU work of earlier chunks
    const uint64_t call_bases = offsets[n_reads] - offsets[0];
    const uint64_t call_est = call_bases > n_reads * (uint64_t)(K - 1) ? call_bases - n_reads * (uint64_t)(K - 1) : 0;
    const bool bucket_call = !n_new_per_read && bucket_usable(st, mode, K, 0, call_est);
    // k-mer total accumulates on the device across chunks (scratch[3]); read once at the end
    unsigned long long* d_tot = g_ctx.d_scratch + 3;
    CU(cudaMemsetAsync(d_tot, 0, sizeof(unsigned long long), g_ctx.slot[0].stream));
    CU(cudaStreamSynchronize(g_ctx.slot[0].stream));
    bool any_bucketed = false;
    for (size_t c = 0; c + 1 < cuts.size(); ++c) {
        Slot& sl = g_ctx.slot[c & 1];
        cudaStream_t s = sl.stream;
        uint64_t r0 = cuts[c], r1 = cuts[c + 1], nr = r1 - r0;
        // validate this chunk's offsets and count its k-mers (invalid reads only lower the count)
        uint64_t chunk_kmers = 0;
        for (uint64_t r = r0; r < r1; ++r) {
            if (offsets[r + 1] < offsets[r]) {
                cudaDeviceSynchronize();
                return fail("gt_insert_sequences: offsets must be non-decreasing (read %llu)", (unsigned long long)r);
            }
            uint64_t len = offsets[r + 1] - offsets[r];
            chunk_kmers += len >= (uint64_t)K ? len - (uint64_t)K + 1 : 0;
        }
        const bool bucketed = bucket_call && chunk_kmers <= st->pend->budget_kmers;
        any_bucketed |= bucketed;
        gt_batch view;
        if (sl.consumed_pending) {  // the compute stream may still be reading this slot's buffers
            CU(cudaStreamWaitEvent(s, sl.consumed, 0));
            sl.consumed_pending = false;
        }
        if (stage_chunk(sl, bases, offsets, r0, r1, view)) return -1;
        uint8_t* d_status = nullptr;
        if (status) {
            if (sl.out8.reserve(nr, s)) return -1;
            d_status = sl.out8.as<uint8_t>();
        }
        k_kmer_counts<<<grid_for(nr, 256, 16), 256, 0, s>>>(view.d_offsets, nr, K, view.d_flags, nullptr, d_status, d_tot); ++g_launches;
        CU(cudaGetLastError());
        uint64_t* d_n_new = nullptr;
        if (n_new_per_read) {
            if (sl.out64a.reserve(nr * 8, s)) return -1;
            d_n_new = sl.out64a.as<uint64_t>();
            CU(cudaMemsetAsync(d_n_new, 0, nr * 8, s));
        }
        if (bucketed) {
            // copies + pack ran on the slot stream; hashing/bucketing runs on the compute stream
            CU(cudaEventRecord(sl.packed, s));
            CU(cudaStreamWaitEvent(g_ctx.main, sl.packed, 0));
            if (bucket_insert(st, shifter, view, K, chunk_kmers)) return -1;
            CU(cudaEventRecord(sl.consumed, g_ctx.main));
            sl.consumed_pending = true;
        } else if (mode == GT_MODE_EXACT) {
            // serial semantics: the chunks' inserts run one after the other on the compute stream
            // (the slot streams only stage and pack), sharing one claim map
            CU(cudaEventRecord(sl.packed, s));
            CU(cudaStreamWaitEvent(g_ctx.main, sl.packed, 0));
            if (launch_insert(st, shifter, view, K, mode, d_n_new, g_ctx.main)) return -1;
            CU(cudaEventRecord(sl.consumed, g_ctx.main));
            CU(cudaStreamWaitEvent(s, sl.consumed, 0));
        } else if (launch_insert(st, shifter, view, K, mode, d_n_new, s)) {
            return -1;
        }
        if (status) CU(cudaMemcpyAsync(status + r0, d_status, nr, cudaMemcpyDeviceToHost, s));
        if (n_new_per_read) CU(cudaMemcpyAsync(n_new_per_read + r0, d_n_new, nr * 8, cudaMemcpyDeviceToHost, s));
    }
    CU(cudaStreamSynchronize(g_ctx.slot[0].stream));
    CU(cudaStreamSynchronize(g_ctx.slot[1].stream));
    if (any_bucketed) {
        CU(cudaStreamSynchronize(g_ctx.main));
        g_ctx.slot[0].consumed_pending = g_ctx.slot[1].consumed_pending = false;
    }
    CU(cudaMemcpy(g_ctx.h_scratch + 3, d_tot, sizeof(unsigned long long), cudaMemcpyDeviceToHost));
    return (int64_t)g_ctx.h_scratch[3];
}

// Same walk, reads already in HBM as ASCII (d_bases) with device offsets starting at 0.  d_total (device,
// may be NULL): the k-mers consumed are ADDED to it on the stream.  Nothing here waits for the GPU.
static int insert_sequences_dev_queue(gt_storage* st, int shifter, int K, const void* d_bases, const void* d_offsets,
                                      uint64_t n_reads, uint64_t n_bases, int mode, unsigned long long* d_total);

extern "C" int gt_insert_sequences_dev_async(gt_storage* st, int shifter, int K, const void* d_bases, const void* d_offsets,
                                             uint64_t n_reads, uint64_t n_bases, int mode, void* d_kmer_total) {
    return insert_sequences_dev_queue(st, shifter, K, d_bases, d_offsets, n_reads, n_bases, mode,
                                      static_cast<unsigned long long*>(d_kmer_total));
}

extern "C" int64_t gt_insert_sequences_dev(gt_storage* st, int shifter, int K, const void* d_bases, const void* d_offsets,
                                            uint64_t n_reads, uint64_t n_bases, int mode) {
    if (ensure_ctx()) return -1;
    if (n_reads == 0) return 0;
    unsigned long long* d_tot = g_ctx.d_scratch + 3;
    CU(cudaMemsetAsync(d_tot, 0, sizeof(unsigned long long), g_ctx.main));
    if (insert_sequences_dev_queue(st, shifter, K, d_bases, d_offsets, n_reads, n_bases, mode, d_tot)) return -1;
    CU(cudaMemcpyAsync(g_ctx.h_scratch + 3, d_tot, sizeof(unsigned long long), cudaMemcpyDeviceToHost, g_ctx.main));
    CU(cudaStreamSynchronize(g_ctx.main));
    return (int64_t)g_ctx.h_scratch[3];
}

static int insert_sequences_dev_queue(gt_storage* st, int shifter, int K, const void* d_bases, const void* d_offsets,
                                      uint64_t n_reads, uint64_t n_bases, int mode, unsigned long long* d_tot) {
    if (ensure_ctx()) return -1;
    if (!st) return fail("gt_insert_sequences_dev: NULL storage");
    if (K < 1 || K > 65535) return fail("gt_insert_sequences_dev: K=%d out of range (1..65535)", K);
    if (check_mode("gt_insert_sequences_dev", mode)) return -1;
    if (n_reads >= (1ull << 32)) return fail("gt_insert_sequences_dev: more than 2^32-1 reads in one call");
    if (n_reads == 0) return 0;
    if (!d_bases || !d_offsets) return fail("gt_insert_sequences_dev: NULL device pointer");
    if ((reinterpret_cast<uintptr_t>(d_bases) & 15) || (reinterpret_cast<uintptr_t>(d_offsets) & 7))
        return fail("gt_insert_sequences_dev: d_bases must be 16-byte aligned and d_offsets 8-byte aligned");
    CU(cudaSetDevice(g_ctx.device));
    if (mode != GT_MODE_BLIND && pending_flush_sync(st)) return -1;
    // pack into slot 0's buffers on the compute stream (everything here is stream-ordered on it)
    Slot& sl = g_ctx.slot[0];
    cudaStream_t s = g_ctx.main;
    if (sl.consumed_pending) {
        CU(cudaStreamWaitEvent(s, sl.consumed, 0));
        sl.consumed_pending = false;
    }
    CU(cudaStreamSynchronize(sl.stream));
    uint64_t n_words = (n_bases + 31) / 32, n_words_alloc = n_words + halo_alloc_words();
    if (sl.words.reserve(n_words_alloc * 8, s) || sl.flags.reserve(n_reads + 1, s) ||
        sl.coarse.reserve(((n_bases >> COARSE_SHIFT) + 2) * 4, s))
        return -1;
    const uint64_t* offs = static_cast<const uint64_t*>(d_offsets);
    if (pack_on_device(static_cast<const uint8_t*>(d_bases), offs, n_reads, n_bases, 0, sl.words.as<uint64_t>(), n_words_alloc,
                       sl.flags.as<uint8_t>(), sl.coarse.as<uint32_t>(), s))
        return -1;
    gt_batch view;
    view.n_reads = n_reads;
    view.n_bases = n_bases;
    view.n_words = n_words;
    view.n_words_alloc = n_words_alloc;
    view.base0 = 0;
    view.d_words = sl.words.as<uint64_t>();
    view.d_offsets = const_cast<uint64_t*>(offs);
    view.d_flags = sl.flags.as<uint8_t>();
    view.d_coarse = sl.coarse.as<uint32_t>();
    if (d_tot) {
        k_kmer_counts<<<grid_for(n_reads, 256, 16), 256, 0, s>>>(offs, n_reads, K, view.d_flags, nullptr, nullptr, d_tot); ++g_launches;
        CU(cudaGetLastError());
    }
    // n_bases bounds the k-mers of the batch without a round trip to the host (or the caller's tighter bound)
    const uint64_t kmers_upper = take_kmer_hint(st, n_bases);
    if (bucket_usable(st, mode, K, kmers_upper, kmers_upper)) {
        if (bucket_insert(st, shifter, view, K, kmers_upper)) return -1;
    } else if (launch_insert(st, shifter, view, K, mode, nullptr, s)) {
        return -1;
    }
    return 0;
}

// ------------------------------------------------------------------------------------------
// 2-bit packed HOST buffers: the parsing-to-device pipeline's product (north_star: "FastxParser -> pinned 2-bit
// buffers on CUDA streams").  The packer runs on the host (parser threads), so only 0.25 B/base cross PCIe.
// Layout = the device layout: base p of the batch at bits 2*(p%32) of u64 word p/32, A=0 C=1 G=2 T=3; a/c/g/t are
// folded to upper case; a read holding any other byte gets flags[r] = GT_READ_INVALID (DNA_SIMPLE validation,
// sequences/alphabets.hh:112-130, parsing/readers.hh:162-171) and its codes are unspecified.
// ------------------------------------------------------------------------------------------
extern "C" int gt_host_pack_append(const unsigned char* s, size_t L, uint64_t* words, uint64_t pos);  // pack_simd.cpp (AVX2 / portable)

static void host_pack_range(const char* bases, const uint64_t* offsets, uint64_t n_reads, uint64_t base0, uint64_t w_lo, uint64_t w_hi,
                            uint64_t n_bases, uint64_t* words, uint8_t* flags) {
    if (w_lo >= w_hi) return;
    const uint64_t p_lo = w_lo * 32, p_hi = std::min(w_hi * 32, n_bases);
    memset(words + w_lo, 0, (w_hi - w_lo) * 8);
    // word-aligned ranges: nothing is written outside [w_lo, w_hi)
    if (!gt_host_pack_append(reinterpret_cast<const unsigned char*>(bases) + base0 + p_lo, p_hi - p_lo, words, p_lo)) return;
    for (uint64_t q = p_lo; q < p_hi; ++q) {  // rare: flag every read owning an offending byte
        const unsigned char c = (unsigned char)bases[base0 + q] & 0xDF;
        if (c != 'A' && c != 'C' && c != 'G' && c != 'T') {
            const uint64_t r = (uint64_t)(std::upper_bound(offsets, offsets + n_reads + 1, base0 + q) - offsets) - 1;
            __atomic_store_n(flags + r, (uint8_t)GT_READ_INVALID, __ATOMIC_RELAXED);
        }
    }
}

extern "C" int gt_pack_reads_host(const char* bases, const uint64_t* offsets, uint64_t n_reads, uint64_t* words, uint8_t* flags,
                                  int n_threads) {
    if (n_reads && (!bases || !offsets)) return fail("gt_pack_reads_host: NULL bases/offsets");
    if (!words || !flags) return fail("gt_pack_reads_host: NULL output");
    if (validate_offsets("gt_pack_reads_host", offsets, n_reads)) return -1;
    if (n_reads == 0) return 0;
    const uint64_t base0 = offsets[0], n_bases = offsets[n_reads] - base0, n_words = (n_bases + 31) / 32;
    memset(flags, 0, n_reads);
    int T = n_threads > 0 ? n_threads : (int)std::max(1u, std::thread::hardware_concurrency());
    T = (int)std::min<uint64_t>((uint64_t)T, std::max<uint64_t>(1, n_words >> 16));
    if (T <= 1) {
        host_pack_range(bases, offsets, n_reads, base0, 0, n_words, n_bases, words, flags);
        return 0;
    }
    std::vector<std::thread> th;
    const uint64_t per = (n_words + T - 1) / T;
    for (int t = 0; t < T; ++t) {
        const uint64_t lo = std::min<uint64_t>(n_words, (uint64_t)t * per), hi = std::min<uint64_t>(n_words, lo + per);
        th.emplace_back(host_pack_range, bases, offsets, n_reads, base0, lo, hi, n_bases, words, flags);
    }
    for (auto& x : th) x.join();
    return 0;
}

// One chunk [r0, r1) of a packed host batch -> slot buffers.  The chunk's words start at the word holding its first
// base; the bases of that word that belong to the previous read are covered by a PHANTOM read flagged invalid
// (device offsets = [word start, offsets[r0..r1]]), so the walkers need no extra bounds test.
static int stage_packed_chunk(Slot& sl, const uint64_t* words, const uint64_t* offsets, const uint8_t* flags, uint64_t r0, uint64_t r1,
                              gt_batch& view) {
    cudaStream_t s = sl.stream;
    const uint64_t b0 = offsets[r0] - offsets[0], b1 = offsets[r1] - offsets[0];  // relative to words[0]
    const uint64_t w0 = b0 / 32, w1 = (b1 + 31) / 32, nr = r1 - r0;
    const uint64_t n_bases = b1 - w0 * 32, n_words = w1 - w0, n_words_alloc = n_words + halo_alloc_words();
    if (sl.words.reserve(n_words_alloc * 8, s) || sl.offsets.reserve((nr + 2) * 8, s) || sl.flags.reserve(nr + 2, s) ||
        sl.coarse.reserve(((n_bases >> COARSE_SHIFT) + 2) * 4, s))
        return -1;
    uint64_t* d_words = sl.words.as<uint64_t>();
    uint64_t* d_off = sl.offsets.as<uint64_t>();
    uint8_t* d_flags = sl.flags.as<uint8_t>();
    if (n_words) CU(cudaMemcpyAsync(d_words, words + w0, n_words * 8, cudaMemcpyHostToDevice, s));
    CU(cudaMemsetAsync(d_words + n_words, 0, (n_words_alloc - n_words) * 8, s));
    const uint64_t phantom_start = offsets[0] + w0 * 32;  // absolute, like the offsets that follow
    CU(cudaMemcpyAsync(d_off, &phantom_start, 8, cudaMemcpyHostToDevice, s));  // pageable source: staged before the call returns
    CU(cudaMemcpyAsync(d_off + 1, offsets + r0, (nr + 1) * 8, cudaMemcpyHostToDevice, s));
    CU(cudaMemsetAsync(d_flags, READ_INVALID, 1, s));
    CU(cudaMemcpyAsync(d_flags + 1, flags + r0, nr, cudaMemcpyHostToDevice, s));
    k_coarse<<<grid_for(nr + 1, 256, 16), 256, 0, s>>>(d_off, nr + 1, phantom_start, sl.coarse.as<uint32_t>()); ++g_launches;
    CU(cudaGetLastError());
    view = gt_batch();
    view.n_reads = nr + 1;
    view.n_bases = n_bases;
    view.n_words = n_words;
    view.n_words_alloc = n_words_alloc;
    view.base0 = phantom_start;
    view.d_words = d_words;
    view.d_offsets = d_off;
    view.d_flags = d_flags;
    view.d_coarse = sl.coarse.as<uint32_t>();
    return 0;
}

// dBG::insert_sequence over a batch that the host has already validated and 2-bit packed (gt_pack_reads_host, or the
// FASTX front end's workers).  words[0] bit 0 holds base offsets[0]; GT_MODE_BLIND / GT_MODE_FAST (no per-read outputs).
extern "C" int64_t gt_insert_sequences_packed(gt_storage* st, int shifter, int K, const uint64_t* words, const uint64_t* offsets,
                                               const uint8_t* flags, uint64_t n_reads, int mode) {
    if (ensure_ctx()) return -1;
    if (!st) return fail("gt_insert_sequences_packed: NULL storage");
    if (K < 1 || K > 65535) return fail("gt_insert_sequences_packed: K=%d out of range (1..65535)", K);
    if (mode != GT_MODE_BLIND && mode != GT_MODE_FAST) return fail("gt_insert_sequences_packed: GT_MODE_BLIND or GT_MODE_FAST");
    if (n_reads == 0) return 0;
    if (!words || !offsets || !flags) return fail("gt_insert_sequences_packed: NULL argument");
    if (n_reads >= (1ull << 32) - 1) return fail("gt_insert_sequences_packed: too many reads in one call");
    CU(cudaSetDevice(g_ctx.device));
    if (mode != GT_MODE_BLIND && pending_flush_sync(st)) return -1;
    if (offsets[n_reads] < offsets[0]) return fail("gt_insert_sequences_packed: offsets must be non-decreasing");
    std::vector<uint64_t> cuts;
    chunk_ranges(offsets, n_reads, cuts);
    const uint64_t call_bases = offsets[n_reads] - offsets[0];
    const uint64_t call_est = call_bases > n_reads * (uint64_t)(K - 1) ? call_bases - n_reads * (uint64_t)(K - 1) : 0;
    const bool bucket_call = bucket_usable(st, mode, K, 0, call_est);
    unsigned long long* d_tot = g_ctx.d_scratch + 3;
    CU(cudaMemsetAsync(d_tot, 0, sizeof(unsigned long long), g_ctx.slot[0].stream));
    CU(cudaStreamSynchronize(g_ctx.slot[0].stream));
    bool any_bucketed = false;
    for (size_t c = 0; c + 1 < cuts.size(); ++c) {
        Slot& sl = g_ctx.slot[c & 1];
        cudaStream_t s = sl.stream;
        const uint64_t r0 = cuts[c], r1 = cuts[c + 1];
        uint64_t chunk_kmers = 0;
        for (uint64_t r = r0; r < r1; ++r) {
            if (offsets[r + 1] < offsets[r]) {
                cudaDeviceSynchronize();
                return fail("gt_insert_sequences_packed: offsets must be non-decreasing (read %llu)", (unsigned long long)r);
            }
            const uint64_t len = offsets[r + 1] - offsets[r];
            chunk_kmers += len >= (uint64_t)K ? len - (uint64_t)K + 1 : 0;
        }
        const bool bucketed = bucket_call && chunk_kmers <= st->pend->budget_kmers;
        any_bucketed |= bucketed;
        if (sl.consumed_pending) {
            CU(cudaStreamWaitEvent(s, sl.consumed, 0));
            sl.consumed_pending = false;
        }
        gt_batch view;
        if (stage_packed_chunk(sl, words, offsets, flags, r0, r1, view)) return -1;
        k_kmer_counts<<<grid_for(view.n_reads, 256, 16), 256, 0, s>>>(view.d_offsets, view.n_reads, K, view.d_flags, nullptr, nullptr, d_tot); ++g_launches;
        CU(cudaGetLastError());
        if (bucketed) {
            CU(cudaEventRecord(sl.packed, s));
            CU(cudaStreamWaitEvent(g_ctx.main, sl.packed, 0));
            if (bucket_insert(st, shifter, view, K, chunk_kmers)) return -1;
            CU(cudaEventRecord(sl.consumed, g_ctx.main));
            sl.consumed_pending = true;
        } else if (launch_insert(st, shifter, view, K, mode, nullptr, s)) {
            return -1;
        }
    }
    CU(cudaStreamSynchronize(g_ctx.slot[0].stream));
    CU(cudaStreamSynchronize(g_ctx.slot[1].stream));
    if (any_bucketed) {
        CU(cudaStreamSynchronize(g_ctx.main));
        g_ctx.slot[0].consumed_pending = g_ctx.slot[1].consumed_pending = false;
    }
    CU(cudaMemcpy(g_ctx.h_scratch + 3, d_tot, sizeof(unsigned long long), cudaMemcpyDeviceToHost));
    return (int64_t)g_ctx.h_scratch[3];
}

// The same for a packed batch already resident in HBM: d_words[0] bit 0 = base 0, d_offsets (uint64, starting at 0),
// d_flags (uint8 per read); the kernels may read up to n_words_alloc >= ceil(n_bases / 32) + 1 words (what lies
// beyond the batch's last base is never used).  Queued on the compute stream, no host wait.
extern "C" int gt_insert_packed_dev_async(gt_storage* st, int shifter, int K, const void* d_words, uint64_t n_words_alloc,
                                          const void* d_offsets, const void* d_flags, uint64_t n_reads, uint64_t n_bases, int mode,
                                          void* d_kmer_total) {
    if (ensure_ctx()) return -1;
    if (!st) return fail("gt_insert_packed_dev_async: NULL storage");
    if (K < 1 || K > 65535) return fail("gt_insert_packed_dev_async: K=%d out of range (1..65535)", K);
    if (mode != GT_MODE_BLIND && mode != GT_MODE_FAST) return fail("gt_insert_packed_dev_async: GT_MODE_BLIND or GT_MODE_FAST");
    if (n_reads == 0) return 0;
    if (!d_words || !d_offsets || !d_flags) return fail("gt_insert_packed_dev_async: NULL device pointer");
    if (n_reads >= (1ull << 32)) return fail("gt_insert_packed_dev_async: too many reads in one call");
    if ((reinterpret_cast<uintptr_t>(d_words) & 15) || (reinterpret_cast<uintptr_t>(d_offsets) & 7))
        return fail("gt_insert_packed_dev_async: d_words must be 16-byte aligned and d_offsets 8-byte aligned");
    if (n_words_alloc < (n_bases + 31) / 32 + 1) return fail("gt_insert_packed_dev_async: n_words_alloc too small");
    CU(cudaSetDevice(g_ctx.device));
    if (mode != GT_MODE_BLIND && pending_flush_sync(st)) return -1;
    Slot& sl = g_ctx.slot[0];
    cudaStream_t s = g_ctx.main;
    if (sl.consumed_pending) {
        CU(cudaStreamWaitEvent(s, sl.consumed, 0));
        sl.consumed_pending = false;
    }
    CU(cudaStreamSynchronize(sl.stream));
    if (sl.coarse.reserve(((n_bases >> COARSE_SHIFT) + 2) * 4, s)) return -1;
    const uint64_t* offs = static_cast<const uint64_t*>(d_offsets);
    k_coarse<<<grid_for(n_reads, 256, 16), 256, 0, s>>>(offs, n_reads, 0, sl.coarse.as<uint32_t>()); ++g_launches;
    CU(cudaGetLastError());
    gt_batch view;
    view.n_reads = n_reads;
    view.n_bases = n_bases;
    view.n_words = (n_bases + 31) / 32;
    view.n_words_alloc = n_words_alloc;
    view.base0 = 0;
    view.d_words = const_cast<uint64_t*>(static_cast<const uint64_t*>(d_words));
    view.d_offsets = const_cast<uint64_t*>(offs);
    view.d_flags = const_cast<uint8_t*>(static_cast<const uint8_t*>(d_flags));
    view.d_coarse = sl.coarse.as<uint32_t>();
    if (d_kmer_total) {
        k_kmer_counts<<<grid_for(n_reads, 256, 16), 256, 0, s>>>(offs, n_reads, K, view.d_flags, nullptr, nullptr,
                                                                 static_cast<unsigned long long*>(d_kmer_total)); ++g_launches;
        CU(cudaGetLastError());
    }
    const uint64_t kmers_upper = take_kmer_hint(st, n_bases);
    if (bucket_usable(st, mode, K, kmers_upper, kmers_upper)) {
        if (bucket_insert(st, shifter, view, K, kmers_upper)) return -1;
    } else if (launch_insert(st, shifter, view, K, mode, nullptr, s)) {
        return -1;
    }
    return 0;
}

// ------------------------------------------------------------------------------------------
// sharded storage: one process per GPU, table t cut into per-rank slot ranges (bucket_host.inc)
// ------------------------------------------------------------------------------------------
extern "C" int gt_shard_plan(int kind, const uint64_t* tablesizes, int n_tables, int world, uint64_t budget_kmers,
                             int slice_log2_bytes, int32_t* shift_nb, int32_t* table, int32_t* owner, uint64_t* slot0,
                             uint64_t* slots, uint32_t* cap, uint64_t* own_lo, uint64_t* own_hi) {
    if (kind < 0 || kind > 2 || !tablesizes || n_tables < 1 || n_tables > MAX_TABLES || world < 1 || !shift_nb)
        return fail("gt_shard_plan: bad argument");
    for (int i = 0; i < n_tables; ++i)
        if (tablesizes[i] == 0) return fail("gt_shard_plan: empty table");
    PlanHost P;
    if (make_plan(kind, tablesizes, n_tables, world, budget_kmers, slice_log2_bytes > 0 ? slice_log2_bytes : default_slice_log2_bytes(kind), P)) return -1;
    shift_nb[0] = P.shift;
    shift_nb[1] = P.nb;
    for (int b = 0; b < P.nb; ++b) {
        if (table) table[b] = P.table[b];
        if (owner) owner[b] = P.owner[b];
        if (slot0) slot0[b] = P.slot0[b];
        if (slots) slots[b] = P.slots[b];
        if (cap) cap[b] = P.cap[b];
    }
    for (size_t k = 0; k < P.own_lo.size(); ++k) {
        if (own_lo) own_lo[k] = P.own_lo[k];
        if (own_hi) own_hi[k] = P.own_hi[k];
    }
    return 0;
}

extern "C" gt_storage* gt_storage_create_sharded(int kind, const uint64_t* tablesizes, int n_tables, int rank, int world,
                                                  uint64_t budget_kmers, int slice_log2_bytes) {
    if (ensure_ctx()) return nullptr;
    if (world < 1 || rank < 0 || rank >= world) { fail("gt_storage_create_sharded: rank %d of %d", rank, world); return nullptr; }
    if (kind < 0 || kind > 2 || !tablesizes || n_tables < 1 || n_tables > MAX_TABLES) { fail("gt_storage_create_sharded: bad argument"); return nullptr; }
    if (budget_kmers == 0 || budget_kmers > (1ull << 31)) { fail("gt_storage_create_sharded: budget_kmers must be 1..2^31"); return nullptr; }
    Pending* p = new Pending();
    if (make_plan(kind, tablesizes, n_tables, world, budget_kmers, slice_log2_bytes > 0 ? slice_log2_bytes : default_slice_log2_bytes(kind), p->host)) { delete p; return nullptr; }
    gt_storage* st = storage_create(kind, tablesizes, n_tables, rank, world, p->host.own_lo.data() + (size_t)rank * n_tables,
                                    p->host.own_hi.data() + (size_t)rank * n_tables);
    if (!st) { delete p; return nullptr; }
    p->budget_kmers = budget_kmers;
    st->pend = p;
    return st;
}

extern "C" int gt_storage_local_range(const gt_storage* st, int i, uint64_t* lo, uint64_t* hi) {
    if (!st || i < 0 || i >= st->n || !lo || !hi) return fail("gt_storage_local_range: bad argument");
    *lo = st->own_lo[i];
    *hi = st->own_hi[i];
    return 0;
}

extern "C" int gt_storage_attach_exchange(gt_storage* st, int which, void* outbox, void* inbox, void* fill_send,
                                           void* fill_recv) {
    if (ensure_ctx()) return -1;
    if (!st || !st->pend || st->world < 2) return fail("gt_storage_attach_exchange: not a sharded storage");
    if (which < 0 || which > 1) return fail("gt_storage_attach_exchange: buffer set 0 or 1");
    if (!outbox || !inbox || !fill_send || !fill_recv) return fail("gt_storage_attach_exchange: NULL buffer");
    Pending* p = st->pend;
    Pending::Store& S = p->store[which];
    if (S.built) return fail("gt_storage_attach_exchange: buffer set %d already attached", which);
    const PlanHost& H = p->host;
    const int nb = H.nb, W = st->world, me = st->rank;
    std::vector<int> owned;  // my buckets, increasing b
    std::vector<uint64_t> R(W, 0);
    for (int b = 0; b < nb; ++b) {
        R[H.owner[b]] += H.cap[b];
        if (H.owner[b] == me) owned.push_back(b);
    }
    const int n_owned = (int)owned.size();
    if (!p->attached && pending_alloc_common(st, p, n_owned * W)) return -1;  // allocates set 0's arrays too
    if (which == 1) {
        if (cudaMalloc(&S.d_bptr, nb * sizeof(uint32_t*)) != cudaSuccess ||
            cudaMalloc(&S.d_items, (size_t)n_owned * W * sizeof(ApplyItem)) != cudaSuccess) {
            cudaGetLastError();
            return fail("gt_storage_attach_exchange: out of device memory");
        }
    }
    // outbox: regions of the peers in rank order (skipping me), then my own region
    std::vector<uint64_t> region_start(W, 0);
    uint64_t off = 0;
    for (int q = 0; q < W; ++q) if (q != me) { region_start[q] = off; off += R[q]; }
    region_start[me] = off;
    off += R[me];
    std::vector<uint64_t> in_region(nb, 0);  // offset of bucket b inside its owner's region
    {
        std::vector<uint64_t> run(W, 0);
        for (int b = 0; b < nb; ++b) { in_region[b] = run[H.owner[b]]; run[H.owner[b]] += H.cap[b]; }
    }
    uint32_t* ob = static_cast<uint32_t*>(outbox);
    uint32_t* ib = static_cast<uint32_t*>(inbox);
    std::vector<uint32_t*> ptrs(nb);
    for (int b = 0; b < nb; ++b) ptrs[b] = ob + region_start[H.owner[b]] + in_region[b];
    CU(cudaMemcpy(S.d_bptr, ptrs.data(), nb * sizeof(uint32_t*), cudaMemcpyHostToDevice));
    // apply items: slice-major, then source rank
    std::vector<ApplyItem> items;
    items.reserve((size_t)n_owned * W);
    const uint32_t* fr = static_cast<const uint32_t*>(fill_recv);
    for (int j = 0; j < n_owned; ++j) {
        const int b = owned[j];
        for (int q = 0; q < W; ++q) {
            const uint32_t* src;
            if (q == me) src = ptrs[b];
            else src = ib + (uint64_t)(q < me ? q : q - 1) * R[me] + in_region[b];
            items.push_back(ApplyItem{src, fr + (size_t)q * n_owned + j, H.slot0[b], H.cap[b], (uint32_t)H.table[b]});
        }
    }
    if (pending_set_items(p, items, which)) return -1;
    S.d_bfill = static_cast<uint32_t*>(fill_send);
    p->own_bfill = false;
    CU(cudaMemset(S.d_bfill, 0, (size_t)nb * 4));
    p->entries_total = off;
    S.built = true;
    p->attached = true;
    return 0;
}

// ---- peer-memory transport ---------------------------------------------------------------
extern "C" void* gt_peer_alloc(uint64_t bytes) {
    if (ensure_ctx()) return nullptr;
    CUP(cudaSetDevice(g_ctx.device));
    void* p = nullptr;
    if (bytes == 0) bytes = 16;
    cudaError_t e = cudaMalloc(&p, bytes);
    if (e != cudaSuccess) {
        cudaGetLastError();
        fail("gt_peer_alloc: cudaMalloc(%llu) failed: %s", (unsigned long long)bytes, cudaGetErrorString(e));
        return nullptr;
    }
    CUP(cudaMemset(p, 0, bytes));
    return p;
}
extern "C" int gt_peer_free(void* ptr) {
    if (!ptr) return 0;
    CU(cudaDeviceSynchronize());
    CU(cudaFree(ptr));
    return 0;
}
extern "C" int gt_peer_export(void* ptr, uint8_t handle[64]) {
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handles are 64 bytes");
    if (ensure_ctx()) return -1;
    if (!ptr || !handle) return fail("gt_peer_export: NULL argument");
    cudaIpcMemHandle_t h;
    CU(cudaIpcGetMemHandle(&h, ptr));
    memcpy(handle, &h, 64);
    return 0;
}
extern "C" void* gt_peer_open(const uint8_t handle[64]) {
    if (ensure_ctx()) return nullptr;
    if (!handle) { fail("gt_peer_open: NULL handle"); return nullptr; }
    CUP(cudaSetDevice(g_ctx.device));
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, 64);
    void* p = nullptr;
    CUP(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    return p;
}
extern "C" int gt_peer_close(void* ptr) {
    if (!ptr) return 0;
    CU(cudaDeviceSynchronize());
    CU(cudaIpcCloseMemHandle(ptr));
    return 0;
}

// layout of a rank's inbox (peer transport): world bucket regions of R entries, then world overflow lists
static uint64_t ovf_records() { return std::max<uint64_t>(16, env_u64("GT_OVF_RECORDS", 1u << 18)); }
static uint64_t inbox_region_entries(const PlanHost& H, int rank) {
    uint64_t R = 0;
    for (int b = 0; b < H.nb; ++b)
        if (H.owner[b] == rank) R += H.cap[b];
    return R;
}
static uint64_t inbox_ovf_offset_bytes(const PlanHost& H, int rank) {
    return ((uint64_t)H.world * inbox_region_entries(H, rank) * 4 + 15) / 16 * 16;
}
extern "C" uint64_t gt_storage_inbox_bytes(const gt_storage* st, int rank) {
    if (!st || !st->pend || rank < 0 || rank >= st->world) return 0;
    return inbox_ovf_offset_bytes(st->pend->host, rank) + (uint64_t)st->world * ovf_records() * 8;
}

// The inbox layout of the peer transport as host arithmetic (no GPU): what gt_storage_attach_peers lays out.
static void peer_layout(const PlanHost& H, std::vector<uint64_t>& R, std::vector<uint64_t>& in_region) {
    R.assign(H.world, 0);
    in_region.assign(H.nb, 0);
    for (int b = 0; b < H.nb; ++b) {
        in_region[b] = R[H.owner[b]];
        R[H.owner[b]] += H.cap[b];
    }
}
extern "C" int gt_shard_peer_layout(int kind, const uint64_t* tablesizes, int n_tables, int world, uint64_t budget_kmers,
                                    int slice_log2_bytes, uint64_t* region_entries, uint64_t* in_region,
                                    uint64_t* ovf_offset_bytes, uint64_t* inbox_bytes) {
    if (kind < 0 || kind > 2 || !tablesizes || n_tables < 1 || n_tables > MAX_TABLES || world < 1)
        return fail("gt_shard_peer_layout: bad argument");
    PlanHost P;
    if (make_plan(kind, tablesizes, n_tables, world, budget_kmers, slice_log2_bytes > 0 ? slice_log2_bytes : default_slice_log2_bytes(kind), P)) return -1;
    std::vector<uint64_t> R, ir;
    peer_layout(P, R, ir);
    for (int q = 0; q < world; ++q) {
        if (region_entries) region_entries[q] = R[q];
        if (ovf_offset_bytes) ovf_offset_bytes[q] = inbox_ovf_offset_bytes(P, q);
        if (inbox_bytes) inbox_bytes[q] = inbox_ovf_offset_bytes(P, q) + (uint64_t)world * ovf_records() * 8;
    }
    if (in_region)
        for (int b = 0; b < P.nb; ++b) in_region[b] = ir[b];
    return P.nb;
}

// Common part of the peer transports: region_of_rank[q] / ovf_of_rank[q] = where this rank's entries / overflow records
// for owner q are written by k_bucket (peer memory, or a local staging area the caller ships); own_inbox = this rank's
// inbox (world regions of R_me entries + world overflow lists), which the apply reads.
static int attach_producer_areas(gt_storage* st, int which, const char* who, uint32_t* const* region_of_rank,
                                 unsigned long long* const* ovf_of_rank, const void* own_inbox, void* fill_send, void* fill_recv,
                                 bool staged) {
    Pending* p = st->pend;
    Pending::Store& S = p->store[which];
    if (S.built) return fail("%s: buffer set %d already attached", who, which);
    const PlanHost& H = p->host;
    const int nb = H.nb, W = st->world, me = st->rank;
    std::vector<int> owned;
    std::vector<uint64_t> R, in_region;
    peer_layout(H, R, in_region);
    for (int b = 0; b < nb; ++b)
        if (H.owner[b] == me) owned.push_back(b);
    const int n_owned = (int)owned.size();
    if (!p->attached && pending_alloc_common(st, p, n_owned * W)) return -1;
    if (which == 1) {
        if (cudaMalloc(&S.d_bptr, nb * sizeof(uint32_t*)) != cudaSuccess ||
            cudaMalloc(&S.d_items, (size_t)n_owned * W * sizeof(ApplyItem)) != cudaSuccess) {
            cudaGetLastError();
            return fail("%s: out of device memory", who);
        }
    }
    std::vector<uint32_t*> ptrs(nb);
    for (int b = 0; b < nb; ++b) ptrs[b] = region_of_rank[H.owner[b]] + in_region[b];
    CU(cudaMemcpy(S.d_bptr, ptrs.data(), nb * sizeof(uint32_t*), cudaMemcpyHostToDevice));
    std::vector<ApplyItem> items;
    items.reserve((size_t)n_owned * W);
    const uint32_t* fr = static_cast<const uint32_t*>(fill_recv);
    const uint32_t* mine = static_cast<const uint32_t*>(own_inbox);
    for (int j = 0; j < n_owned; ++j) {
        const int b = owned[j];
        for (int q = 0; q < W; ++q)
            items.push_back(ApplyItem{mine + (uint64_t)q * R[me] + in_region[b], fr + (size_t)q * (n_owned + 1) + j, H.slot0[b],
                                      H.cap[b], (uint32_t)H.table[b]});
    }
    if (pending_set_items(p, items, which)) return -1;
    S.d_bfill = static_cast<uint32_t*>(fill_send);
    p->own_bfill = false;
    // overflow lists: mine for every owner (to post to), the world lists of my own inbox (to apply)
    p->ovf_cap = (uint32_t)ovf_records();
    p->count_stride = (uint32_t)n_owned + 1;
    if (!p->d_bowner) {
        std::vector<uint8_t> own(nb);
        for (int b = 0; b < nb; ++b) own[b] = (uint8_t)H.owner[b];
        CU(cudaMalloc(&p->d_bowner, nb));
        CU(cudaMemcpy(p->d_bowner, own.data(), nb, cudaMemcpyHostToDevice));
    }
    CU(cudaMalloc(&S.d_ovf_ptr, W * sizeof(unsigned long long*)));
    CU(cudaMemcpy(S.d_ovf_ptr, ovf_of_rank, W * sizeof(unsigned long long*), cudaMemcpyHostToDevice));
    S.my_ovf = reinterpret_cast<const unsigned long long*>(static_cast<const char*>(own_inbox) + inbox_ovf_offset_bytes(H, me));
    S.ovf_counts = fr + n_owned;
    S.fill_words = (uint32_t)(nb + W);
    CU(cudaMemset(S.d_bfill, 0, (size_t)S.fill_words * 4));
    p->entries_total = (uint64_t)W * R[me];
    p->staged = staged;
    S.built = true;
    p->attached = true;
    return 0;
}

extern "C" int gt_storage_attach_peers(gt_storage* st, int which, void* const* inbox_of_rank, void* fill_send,
                                        void* fill_recv) {
    if (ensure_ctx()) return -1;
    if (!st || !st->pend || st->world < 2) return fail("gt_storage_attach_peers: not a sharded storage");
    if (which < 0 || which > 1) return fail("gt_storage_attach_peers: buffer set 0 or 1");
    if (!inbox_of_rank || !fill_send || !fill_recv) return fail("gt_storage_attach_peers: NULL buffer");
    const PlanHost& H = st->pend->host;
    const int W = st->world, me = st->rank;
    for (int q = 0; q < W; ++q)
        if (!inbox_of_rank[q]) return fail("gt_storage_attach_peers: inbox of rank %d is NULL", q);
    // bucket b, produced here, lands in region `me` of its owner's inbox; so does this rank's overflow list
    std::vector<uint32_t*> regions(W);
    std::vector<unsigned long long*> lists(W);
    const uint64_t cap = ovf_records();
    for (int q = 0; q < W; ++q) {
        regions[q] = static_cast<uint32_t*>(inbox_of_rank[q]) + (uint64_t)me * inbox_region_entries(H, q);
        lists[q] = reinterpret_cast<unsigned long long*>(static_cast<char*>(inbox_of_rank[q]) + inbox_ovf_offset_bytes(H, q)) + (uint64_t)me * cap;
    }
    return attach_producer_areas(st, which, "gt_storage_attach_peers", regions.data(), lists.data(), inbox_of_rank[me], fill_send,
                                 fill_recv, false);
}

// Staged peer transport (copy engines instead of SM stores): as gt_storage_attach_peers, but what this rank produces for a
// foreign owner q is written to a LOCAL staging area -- stage_of_rank[q]: R_q entries, padded to 16 bytes, then this
// rank's overflow list for q -- which the caller ships into q's inbox (region `rank`, overflow list `rank`) with plain
// device-to-device copies over NVLink (gt_peer_copy_async) while the SMs go on hashing.
extern "C" uint64_t gt_storage_stage_bytes(const gt_storage* st, int rank) {
    if (!st || !st->pend || rank < 0 || rank >= st->world) return 0;
    return (inbox_region_entries(st->pend->host, rank) * 4 + 15) / 16 * 16 + ovf_records() * 8;
}
extern "C" int gt_storage_attach_areas(gt_storage* st, int which, void* own_inbox, void* const* region_of_rank,
                                        void* const* ovf_of_rank, void* fill_send, void* fill_recv) {
    if (ensure_ctx()) return -1;
    if (!st || !st->pend || st->world < 2) return fail("gt_storage_attach_areas: not a sharded storage");
    if (which < 0 || which > 1) return fail("gt_storage_attach_areas: buffer set 0 or 1");
    if (!own_inbox || !region_of_rank || !ovf_of_rank || !fill_send || !fill_recv) return fail("gt_storage_attach_areas: NULL buffer");
    const PlanHost& H = st->pend->host;
    const int W = st->world, me = st->rank;
    std::vector<uint32_t*> regions(W);
    std::vector<unsigned long long*> lists(W);
    const uint64_t cap = ovf_records();
    for (int q = 0; q < W; ++q) {
        if (q == me) {  // this rank's own buckets go where the apply reads them
            regions[q] = static_cast<uint32_t*>(own_inbox) + (uint64_t)me * inbox_region_entries(H, me);
            lists[q] = reinterpret_cast<unsigned long long*>(static_cast<char*>(own_inbox) + inbox_ovf_offset_bytes(H, me)) + (uint64_t)me * cap;
            continue;
        }
        if (!region_of_rank[q] || !ovf_of_rank[q]) return fail("gt_storage_attach_areas: area for rank %d is NULL", q);
        if ((reinterpret_cast<uintptr_t>(region_of_rank[q]) & 15) || (reinterpret_cast<uintptr_t>(ovf_of_rank[q]) & 7))
            return fail("gt_storage_attach_areas: regions must be 16-byte aligned, overflow lists 8-byte aligned");
        regions[q] = static_cast<uint32_t*>(region_of_rank[q]);
        lists[q] = static_cast<unsigned long long*>(ovf_of_rank[q]);
    }
    return attach_producer_areas(st, which, "gt_storage_attach_areas", regions.data(), lists.data(), own_inbox, fill_send, fill_recv,
                                 true);
}
extern "C" int gt_storage_attach_staged(gt_storage* st, int which, void* own_inbox, void* const* stage_of_rank, void* fill_send,
                                         void* fill_recv) {
    if (!st || !st->pend || st->world < 2) return fail("gt_storage_attach_staged: not a sharded storage");
    if (!stage_of_rank) return fail("gt_storage_attach_staged: NULL buffer");
    const PlanHost& H = st->pend->host;
    const int W = st->world, me = st->rank;
    std::vector<void*> regions(W, nullptr), lists(W, nullptr);
    for (int q = 0; q < W; ++q) {
        if (q == me) continue;
        if (!stage_of_rank[q]) return fail("gt_storage_attach_staged: staging area for rank %d is NULL", q);
        regions[q] = stage_of_rank[q];
        lists[q] = static_cast<char*>(stage_of_rank[q]) + (inbox_region_entries(H, q) * 4 + 15) / 16 * 16;
    }
    return gt_storage_attach_areas(st, which, own_inbox, regions.data(), lists.data(), fill_send, fill_recv);
}

// Device-to-device copy on a stream of the caller (both pointers valid in this process: local memory or a peer's
// allocation mapped with gt_peer_open).  Runs on a copy engine, not on the SMs.
extern "C" int gt_peer_copy_async(void* dst, const void* src, uint64_t bytes, void* stream) {
    if (ensure_ctx()) return -1;
    if (bytes == 0) return 0;
    if (!dst || !src) return fail("gt_peer_copy_async: NULL pointer");
    CU(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDefault, static_cast<cudaStream_t>(stream)));
    return 0;
}

// Upper bound of the k-mers in the NEXT device-resident batch bucketed into this storage (gt_insert_sequences_dev[_async],
// gt_insert_packed_dev_async), for callers that know it (equal-length reads): without it the library has to assume one
// k-mer per base, and a sharded storage's round budget -- which sizes every exchange buffer -- has to be given in bases.
extern "C" int gt_storage_hint_kmers(gt_storage* st, uint64_t n_kmers_upper) {
    if (!st) return fail("gt_storage_hint_kmers: NULL storage");
    st->hint_kmers = n_kmers_upper;
    return 0;
}

extern "C" int gt_query_hashes_local_dev(gt_storage* st, const void* d_hashes, uint64_t n, void* d_counts) {
    if (ensure_ctx()) return -1;
    if (!st || (n && (!d_hashes || !d_counts))) return fail("gt_query_hashes_local_dev: NULL argument");
    if (st->pend && st->pend->pending_total()) return fail("gt_query_hashes_local_dev: updates are pending (exchange and apply first)");
    if (n == 0) return 0;
    CU(cudaSetDevice(g_ctx.device));
    cudaStream_t s = g_ctx.main;
    if (g_ctx.apply) {  // queries must see every applied update
        cudaEvent_t& ev = g_ctx.slot[0].packed;
        CU(cudaEventRecord(ev, g_ctx.apply));
        CU(cudaStreamWaitEvent(s, ev, 0));
    }
    OwnRange own;
    memset(&own, 0, sizeof own);
    for (int i = 0; i < st->n; ++i) { own.lo[i] = st->own_lo[i]; own.hi[i] = st->own_hi[i]; }
    const uint64_t* d_h = static_cast<const uint64_t*>(d_hashes);
    int16_t* d_c = static_cast<int16_t*>(d_counts);
    const int grid = grid_for(n, 256 * 4, 8);
    if (st->kind == 0) k_query_hashes_local<0><<<grid, 256, 0, s>>>(d_h, n, st->ts, own, d_c);
    else if (st->kind == 1) k_query_hashes_local<1><<<grid, 256, 0, s>>>(d_h, n, st->ts, own, d_c);
    else k_query_hashes_local<2><<<grid, 256, 0, s>>>(d_h, n, st->ts, own, d_c);
    ++g_launches;
    CU(cudaGetLastError());
    return 0;
}

static int scan_counts(Slot& sl, const uint64_t* d_in, uint64_t n, uint64_t* d_out, cudaStream_t s);
// ---- owner-routed requests (kernels.cuh, k_route_requests ...) -------------------------------------------------------
static int shard_ready(const char* who, gt_storage* st) {
    if (ensure_ctx()) return -1;
    if (!st || !st->pend || st->world < 2) return fail("%s: not a sharded storage", who);
    if (st->pend->pending_total()) return fail("%s: updates are pending (exchange and apply first)", who);
    CU(cudaSetDevice(g_ctx.device));
    if (g_ctx.apply) {  // must see every applied update
        cudaEvent_t& ev = g_ctx.slot[0].packed;
        CU(cudaEventRecord(ev, g_ctx.apply));
        CU(cudaStreamWaitEvent(g_ctx.main, ev, 0));
    }
    return 0;
}
// d_req[n * n_tables] (uint64: table << 59 | slot) and d_owner[n * n_tables] (int32 rank) for n hash values
extern "C" int gt_shard_route_hashes_dev(gt_storage* st, const void* d_hashes, uint64_t n, void* d_req, void* d_owner) {
    if (ensure_ctx()) return -1;
    if (!st || !st->pend || st->world < 2) return fail("gt_shard_route_hashes_dev: not a sharded storage");
    if (n == 0) return 0;
    if (!d_hashes || !d_req || !d_owner) return fail("gt_shard_route_hashes_dev: NULL argument");
    CU(cudaSetDevice(g_ctx.device));
    RouteArgs ra;
    memset(&ra, 0, sizeof ra);
    const PlanHost& H = st->pend->host;
    for (int t = 0; t < st->n; ++t) ra.spr[t] = std::max<uint64_t>(1, H.own_hi[t] - H.own_lo[t]);  // rank 0's share = every rank's
    k_route_requests<<<grid_for(n, 256, 8), 256, 0, g_ctx.main>>>(static_cast<const uint64_t*>(d_hashes), n, st->ts, ra,
                                                                   static_cast<unsigned long long*>(d_req), static_cast<int32_t*>(d_owner));
    ++g_launches;
    CU(cudaGetLastError());
    return 0;
}
// owner side of a routed query: d_answers[i] (uint8) = the slot of request i (bit, or counter value)
extern "C" int gt_shard_answer_dev(gt_storage* st, const void* d_req, uint64_t n_req, void* d_answers) {
    if (shard_ready("gt_shard_answer_dev", st)) return -1;
    if (n_req == 0) return 0;
    if (!d_req || !d_answers) return fail("gt_shard_answer_dev: NULL argument");
    const unsigned long long* r = static_cast<const unsigned long long*>(d_req);
    uint8_t* a = static_cast<uint8_t*>(d_answers);
    const int g = grid_for(n_req, 256, 8);
    if (st->kind == 0) k_answer_requests<0><<<g, 256, 0, g_ctx.main>>>(r, n_req, st->ts, a);
    else if (st->kind == 1) k_answer_requests<1><<<g, 256, 0, g_ctx.main>>>(r, n_req, st->ts, a);
    else k_answer_requests<2><<<g, 256, 0, g_ctx.main>>>(r, n_req, st->ts, a);
    ++g_launches;
    CU(cudaGetLastError());
    return 0;
}
// owner side of a routed insert: apply the requests; with d_ord (uint32 serial ordinals assigned by the sources) also
// d_first[i] = 1 iff request i is the FIRST toucher (smallest ordinal of this call) of a slot that was zero before the call
extern "C" int gt_shard_insert_requests_dev(gt_storage* st, const void* d_req, const void* d_ord, uint64_t n_req, void* d_first) {
    if (shard_ready("gt_shard_insert_requests_dev", st)) return -1;
    if (n_req == 0) return 0;
    if (!d_req) return fail("gt_shard_insert_requests_dev: NULL argument");
    if (d_first && !d_ord) return fail("gt_shard_insert_requests_dev: first-toucher flags need ordinals");
    const unsigned long long* r = static_cast<const unsigned long long*>(d_req);
    const uint32_t* o = static_cast<const uint32_t*>(d_ord);
    uint8_t* f = static_cast<uint8_t*>(d_first);
    cudaStream_t s = g_ctx.main;
    ClaimMap cm;
    memset(&cm, 0, sizeof cm);
    const int g = grid_for(n_req, 256, 8);
    if (o) {
        if (exact_map(n_req, s, cm)) return -1;
        if (st->kind == 0) k_claim_requests<0><<<g, 256, 0, s>>>(r, o, n_req, st->ts, cm);
        else if (st->kind == 1) k_claim_requests<1><<<g, 256, 0, s>>>(r, o, n_req, st->ts, cm);
        else k_claim_requests<2><<<g, 256, 0, s>>>(r, o, n_req, st->ts, cm);
        ++g_launches;
    }
    if (st->kind == 0) k_insert_requests<0><<<g, 256, 0, s>>>(r, o, n_req, st->ts, cm, f);
    else if (st->kind == 1) k_insert_requests<1><<<g, 256, 0, s>>>(r, o, n_req, st->ts, cm, f);
    else k_insert_requests<2><<<g, 256, 0, s>>>(r, o, n_req, st->ts, cm, f);
    ++g_launches;
    CU(cudaGetLastError());
    return 0;
}

// KmerIterator over reads resident in HBM: d_values (uint64, capacity >= n_bases) receives hash_type::value() of every
// k-mer (the canonical minimum for GT_SHIFTER_CAN), reads back to back; returns the number of k-mers (one 8-byte read-back).
extern "C" int64_t gt_hash_values_dev(int shifter, int K, const void* d_bases, const void* d_offsets, uint64_t n_reads, uint64_t n_bases,
                                       void* d_values) {
    if (ensure_ctx()) return -1;
    if (K < 1 || K > 65535) return fail("gt_hash_values_dev: K=%d out of range (1..65535)", K);
    if (n_reads == 0) return 0;
    if (!d_bases || !d_offsets || !d_values) return fail("gt_hash_values_dev: NULL device pointer");
    if ((reinterpret_cast<uintptr_t>(d_bases) & 15) || (reinterpret_cast<uintptr_t>(d_offsets) & 7))
        return fail("gt_hash_values_dev: d_bases must be 16-byte aligned and d_offsets 8-byte aligned");
    if (n_reads >= (1ull << 32)) return fail("gt_hash_values_dev: too many reads");
    CU(cudaSetDevice(g_ctx.device));
    Slot& sl = g_ctx.slot[0];
    cudaStream_t s = g_ctx.main;
    if (sl.consumed_pending) {
        CU(cudaStreamWaitEvent(s, sl.consumed, 0));
        sl.consumed_pending = false;
    }
    CU(cudaStreamSynchronize(sl.stream));
    const uint64_t n_words = (n_bases + 31) / 32, n_words_alloc = n_words + halo_alloc_words();
    if (sl.words.reserve(n_words_alloc * 8, s) || sl.flags.reserve(n_reads + 1, s) || sl.coarse.reserve(((n_bases >> COARSE_SHIFT) + 2) * 4, s) ||
        sl.kcount.reserve(n_reads * 8, s) || sl.koff.reserve(n_reads * 8, s))
        return -1;
    const uint64_t* offs = static_cast<const uint64_t*>(d_offsets);
    if (pack_on_device(static_cast<const uint8_t*>(d_bases), offs, n_reads, n_bases, 0, sl.words.as<uint64_t>(), n_words_alloc,
                       sl.flags.as<uint8_t>(), sl.coarse.as<uint32_t>(), s))
        return -1;
    gt_batch view;
    view.n_reads = n_reads;
    view.n_bases = n_bases;
    view.n_words = n_words;
    view.n_words_alloc = n_words_alloc;
    view.base0 = 0;
    view.d_words = sl.words.as<uint64_t>();
    view.d_offsets = const_cast<uint64_t*>(offs);
    view.d_flags = sl.flags.as<uint8_t>();
    view.d_coarse = sl.coarse.as<uint32_t>();
    const int64_t nk = batch_kmers(view, K, sl.kcount.as<uint64_t>(), nullptr, s);
    if (nk < 0) return -1;
    if (scan_counts(sl, sl.kcount.as<uint64_t>(), n_reads, sl.koff.as<uint64_t>(), s)) return -1;
    WalkArgs a = make_args(view, K);
    a.koff = sl.koff.as<uint64_t>();
    a.fw = static_cast<uint64_t*>(d_values);
    a.rc = nullptr;  // value() only
    if (launch_hash(shifter, a, s)) return -1;
    CU(cudaStreamSynchronize(s));
    return nk;
}

// A sharded storage can hold two buffer sets so that the exchange of one round overlaps the
// hashing of the next: choose the set the next inserts bucket into.
extern "C" int gt_storage_select_store(gt_storage* st, int which) {
    if (!st || !st->pend || st->world < 2) return fail("gt_storage_select_store: not a sharded storage");
    if (which < 0 || which > 1 || !st->pend->store[which].built) return fail("gt_storage_select_store: set %d not attached", which);
    st->pend->cur = which;
    return 0;
}

// K2 over this rank's slices for one buffer set: its own buckets plus what the peers sent
// (fill_recv holds the counts).  Asynchronous on the apply stream (gt_set_apply_stream; the
// compute stream when none was given); empties this rank's cursors of that set.  The caller
// orders it after the exchange and before the set is bucketed into again.
extern "C" int gt_storage_apply_store(gt_storage* st, int which) {
    if (ensure_ctx()) return -1;
    if (!st || !st->pend || st->world < 2) return fail("gt_storage_apply_store: not a sharded storage");
    if (which < 0 || which > 1 || !st->pend->store[which].built) return fail("gt_storage_apply_store: set %d not attached", which);
    Pending::Store& S = st->pend->store[which];
    S.pending_kmers = std::max<uint64_t>(S.pending_kmers, 1);  // peers may have sent even if this rank bucketed nothing
    return store_apply_async(st, which);
}

// Queue the apply of everything pending without waiting (single-GPU storages; on a sharded
// storage: buffer set 0, after its exchange).
extern "C" int gt_storage_apply(gt_storage* st) {
    if (ensure_ctx()) return -1;
    if (!st) return fail("gt_storage_apply: NULL storage");
    if (!st->pend) return 0;
    if (st->world > 1) return gt_storage_apply_store(st, st->pend->cur);
    return pending_flush_async(st);
}

// Run this library's kernels on streams of the caller (e.g. torch streams) so that they order
// with the caller's collectives; NULL restores the library's own stream.
extern "C" int gt_set_compute_stream(void* stream) {
    if (ensure_ctx()) return -1;
    CU(cudaDeviceSynchronize());
    g_ctx.main = stream ? static_cast<cudaStream_t>(stream) : g_ctx.own_main;
    return 0;
}
extern "C" int gt_set_apply_stream(void* stream) {
    if (ensure_ctx()) return -1;
    CU(cudaDeviceSynchronize());
    g_ctx.apply = stream ? static_cast<cudaStream_t>(stream) : g_ctx.own_apply;
    g_ctx.apply_is_external = stream != nullptr;
    return 0;
}

// Apply every pending (write-combined) insert of `st` to its tables and wait for it.
extern "C" int gt_storage_flush(gt_storage* st) {
    if (ensure_ctx()) return -1;
    if (!st) return fail("gt_storage_flush: NULL storage");
    if (pending_flush_sync(st)) return -1;
    return 0;
}

extern "C" int gt_storage_pending_info(gt_storage* st, uint64_t* info) {
    if (ensure_ctx()) return -1;
    if (!st || !info) return fail("gt_storage_pending_info: NULL argument");
    memset(info, 0, 8 * sizeof(uint64_t));
    if (!st->pend) return 0;
    Pending* p = st->pend;
    unsigned long long nd[2] = {0, 0};
    CU(cudaDeviceSynchronize());
    CU(cudaMemcpy(nd, p->d_counters, 16, cudaMemcpyDeviceToHost));
    if (nd[1]) return fail("%llu updates overflowed both their bucket and the overflow / spill list (extremely skewed input: "
                           "lower the k-mers per round or call)", nd[1]);
    info[0] = 1;
    info[1] = (uint64_t)p->host.nb;
    info[2] = (uint64_t)p->plan.shift;
    info[3] = p->budget_kmers;
    info[4] = p->entries_total;
    info[5] = p->pending_total();
    info[6] = nd[0];
    info[7] = p->total_chunks;
    return 0;
}

// Synthetic reads for harnesses (kernels.cuh, k_synth_bases): n_bases ASCII bytes of stream `seed` starting at
// global base index `first_base`, written to device memory on the compute stream.
extern "C" int gt_synth_bases_dev(void* d_out, uint64_t n_bases, uint64_t seed, uint64_t first_base) {
    if (ensure_ctx()) return -1;
    if (n_bases == 0) return 0;
    if (!d_out || (reinterpret_cast<uintptr_t>(d_out) & 15)) return fail("gt_synth_bases_dev: d_out must be a 16-byte aligned device pointer");
    CU(cudaSetDevice(g_ctx.device));
    k_synth_bases<<<grid_for((n_bases + 15) / 16, 256, 8), 256, 0, g_ctx.main>>>(static_cast<uint8_t*>(d_out), n_bases, seed, first_base); ++g_launches;
    CU(cudaGetLastError());
    return 0;
}

extern "C" uint64_t gt_launch_count(void) { return g_launches; }

// Device timing helpers for harnesses: events recorded on the library's compute stream.
static cudaEvent_t g_timer[8] = {nullptr};
extern "C" int gt_timer_record(int slot) {
    if (ensure_ctx()) return -1;
    if (slot < 0 || slot >= 8) return fail("gt_timer_record: slot 0..7");
    if (!g_timer[slot]) CU(cudaEventCreate(&g_timer[slot]));
    // the compute stream is ordered after both staging streams at this point
    for (auto& sl : g_ctx.slot) {
        CU(cudaEventRecord(sl.packed, sl.stream));
        CU(cudaStreamWaitEvent(g_ctx.main, sl.packed, 0));
    }
    CU(cudaEventRecord(g_timer[slot], g_ctx.main));
    return 0;
}
extern "C" double gt_timer_elapsed_ms(int from_slot, int to_slot) {
    if (ensure_ctx()) return -1.0;
    if (from_slot < 0 || from_slot >= 8 || to_slot < 0 || to_slot >= 8 || !g_timer[from_slot] || !g_timer[to_slot]) {
        fail("gt_timer_elapsed_ms: slots not recorded");
        return -1.0;
    }
    float ms = 0;
    if (cudaEventSynchronize(g_timer[to_slot]) != cudaSuccess || cudaEventElapsedTime(&ms, g_timer[from_slot], g_timer[to_slot]) != cudaSuccess) {
        fail("gt_timer_elapsed_ms: %s", cudaGetErrorString(cudaGetLastError()));
        return -1.0;
    }
    return (double)ms;
}
extern "C" int gt_profile_enable(int on) {
    if (ensure_ctx()) return -1;
    CU(cudaDeviceSynchronize());
    prof_resolve();
    g_prof_on = on != 0;
    for (int k = 0; k < PROF_KINDS; ++k) { g_prof_ms[k] = 0; g_prof_n[k] = 0; }
    return 0;
}
extern "C" int gt_profile_get(double* ms3, uint64_t* n3) {
    if (ensure_ctx()) return -1;
    if (!ms3 || !n3) return fail("gt_profile_get: NULL argument");
    CU(cudaDeviceSynchronize());
    prof_resolve();
    for (int k = 0; k < 3; ++k) { ms3[k] = g_prof_ms[k]; n3[k] = g_prof_n[k]; }
    return 0;
}
extern "C" int gt_profile_get_detail(double* ms, uint64_t* n, int count) {
    if (ensure_ctx()) return -1;
    if (!ms || !n || count < 0) return fail("gt_profile_get_detail: bad argument");
    CU(cudaDeviceSynchronize());
    prof_resolve();
    for (int k = 0; k < count; ++k) { ms[k] = k < PROF_KINDS ? g_prof_ms[k] : 0.0; n[k] = k < PROF_KINDS ? g_prof_n[k] : 0; }
    return PROF_KINDS;
}

// exclusive scan of per-read k-mer counts on stream s
static int scan_counts(Slot& sl, const uint64_t* d_in, uint64_t n, uint64_t* d_out, cudaStream_t s) {
    if (n == 0) return 0;
    uint64_t per = (uint64_t)SCAN_BLOCK * SCAN_ITEMS;
    uint64_t nb = (n + per - 1) / per;
    if (sl.partial.reserve(nb * 8, s)) return -1;
    k_scan_partials<<<(unsigned)nb, SCAN_BLOCK, 0, s>>>(d_in, n, sl.partial.as<uint64_t>()); ++g_launches;
    k_scan_top<<<1, 1024, 0, s>>>(sl.partial.as<uint64_t>(), nb); ++g_launches;
    k_scan_final<<<(unsigned)nb, SCAN_BLOCK, 0, s>>>(d_in, n, sl.partial.as<uint64_t>(), d_out); ++g_launches;
    CU(cudaGetLastError());
    return 0;
}

// Shared driver for the per-k-mer output paths (query counts / hashes).
//   what: 0 = query -> counts(int16), 1 = hash -> fw/rc
static int64_t per_kmer_outputs(int what, gt_storage* st, int shifter, int K, const char* bases, const uint64_t* offsets,
                                uint64_t n_reads, int16_t* counts, uint64_t* fw, uint64_t* rc, uint8_t* status) {
    std::vector<uint64_t> cuts;
    chunk_ranges(offsets, n_reads, cuts);
    int64_t total = 0;
    for (size_t c = 0; c + 1 < cuts.size(); ++c) {
        Slot& sl = g_ctx.slot[c & 1];
        cudaStream_t s = sl.stream;
        uint64_t r0 = cuts[c], r1 = cuts[c + 1], nr = r1 - r0;
        gt_batch view;
        if (stage_chunk(sl, bases, offsets, r0, r1, view)) return -1;
        if (sl.kcount.reserve(nr * 8, s) || sl.koff.reserve(nr * 8, s) || sl.out8.reserve(nr, s)) return -1;
        int64_t nk = batch_kmers(view, K, sl.kcount.as<uint64_t>(), sl.out8.as<uint8_t>(), s);
        if (nk < 0) return -1;
        if (scan_counts(sl, sl.kcount.as<uint64_t>(), nr, sl.koff.as<uint64_t>(), s)) return -1;
        WalkArgs a = make_args(view, K);
        a.koff = sl.koff.as<uint64_t>();
        if (what == 0) {
            if (sl.out16.reserve((uint64_t)nk * 2 + 2, s)) return -1;
            a.counts = sl.out16.as<int16_t>();
            if (launch_walk_kind<OP_QUERY, false>(shifter, a, st->ts, s)) return -1;
            if (nk) CU(cudaMemcpyAsync(counts + total, a.counts, (uint64_t)nk * 2, cudaMemcpyDeviceToHost, s));
        } else {
            if (sl.out64a.reserve((uint64_t)nk * 8 + 8, s)) return -1;
            a.fw = sl.out64a.as<uint64_t>();
            if (shifter == GT_SHIFTER_CAN) {
                if (sl.out64b.reserve((uint64_t)nk * 8 + 8, s)) return -1;
                a.rc = sl.out64b.as<uint64_t>();
            }
            if (launch_hash(shifter, a, s)) return -1;
            if (nk) {
                CU(cudaMemcpyAsync(fw + total, a.fw, (uint64_t)nk * 8, cudaMemcpyDeviceToHost, s));
                if (shifter == GT_SHIFTER_CAN && rc) CU(cudaMemcpyAsync(rc + total, a.rc, (uint64_t)nk * 8, cudaMemcpyDeviceToHost, s));
            }
        }
        if (status) CU(cudaMemcpyAsync(status + r0, sl.out8.p, nr, cudaMemcpyDeviceToHost, s));
        total += nk;
    }
    CU(cudaStreamSynchronize(g_ctx.slot[0].stream));
    CU(cudaStreamSynchronize(g_ctx.slot[1].stream));
    return total;
}

extern "C" int64_t gt_query_sequences(gt_storage* st, int shifter, int K, const char* bases, const uint64_t* offsets,
                                       uint64_t n_reads, int16_t* counts, uint8_t* status) {
    if (check_reads("gt_query_sequences", bases, offsets, n_reads, K)) return -1;
    if (!st || !counts) return fail("gt_query_sequences: NULL argument");
    if (st->world > 1) return fail("gt_query_sequences: not available on a sharded storage (query the owner ranks)");
    if (pending_flush_sync(st)) return -1;
    if (validate_offsets("gt_query_sequences", offsets, n_reads)) return -1;
    if (n_reads == 0) return 0;
    CU(cudaSetDevice(g_ctx.device));
    return per_kmer_outputs(0, st, shifter, K, bases, offsets, n_reads, counts, nullptr, nullptr, status);
}

extern "C" int64_t gt_hash_sequences(int shifter, int K, const char* bases, const uint64_t* offsets, uint64_t n_reads,
                                      uint64_t* fw, uint64_t* rc, uint8_t* status) {
    if (check_reads("gt_hash_sequences", bases, offsets, n_reads, K)) return -1;
    if (!fw) return fail("gt_hash_sequences: fw is NULL");
    if (validate_offsets("gt_hash_sequences", offsets, n_reads)) return -1;
    if (n_reads == 0) return 0;
    CU(cudaSetDevice(g_ctx.device));
    return per_kmer_outputs(1, nullptr, shifter, K, bases, offsets, n_reads, nullptr, fw, rc, status);
}

extern "C" int64_t gt_median_count_at_least(gt_storage* st, int shifter, int K, const char* bases, const uint64_t* offsets,
                                             uint64_t n_reads, uint32_t cutoff, uint8_t* pass, uint8_t* status) {
    if (check_reads("gt_median_count_at_least", bases, offsets, n_reads, K)) return -1;
    if (!st || !pass) return fail("gt_median_count_at_least: NULL argument");
    if (st->world > 1) return fail("gt_median_count_at_least: not available on a sharded storage");
    if (pending_flush_sync(st)) return -1;
    if (validate_offsets("gt_median_count_at_least", offsets, n_reads)) return -1;
    if (n_reads == 0) return 0;
    CU(cudaSetDevice(g_ctx.device));
    std::vector<uint64_t> cuts;
    chunk_ranges(offsets, n_reads, cuts);
    unsigned long long* d_tot = g_ctx.d_scratch + 3;
    CU(cudaMemsetAsync(d_tot, 0, sizeof(unsigned long long), g_ctx.slot[0].stream));
    CU(cudaStreamSynchronize(g_ctx.slot[0].stream));
    for (size_t c = 0; c + 1 < cuts.size(); ++c) {
        Slot& sl = g_ctx.slot[c & 1];
        cudaStream_t s = sl.stream;
        uint64_t r0 = cuts[c], r1 = cuts[c + 1], nr = r1 - r0;
        gt_batch view;
        if (stage_chunk(sl, bases, offsets, r0, r1, view)) return -1;
        if (sl.kcount.reserve(nr * 8, s) || sl.hits.reserve(nr * 4, s) || sl.out8.reserve(2 * nr, s)) return -1;
        uint8_t* d_status = sl.out8.as<uint8_t>();
        uint8_t* d_pass = d_status + nr;
        k_kmer_counts<<<grid_for(nr, 256, 16), 256, 0, s>>>(view.d_offsets, nr, K, view.d_flags, sl.kcount.as<uint64_t>(), d_status, d_tot); ++g_launches;
        CU(cudaGetLastError());
        CU(cudaMemsetAsync(sl.hits.p, 0, nr * 4, s));
        WalkArgs a = make_args(view, K);
        a.hits = sl.hits.as<uint32_t>();
        a.cutoff = cutoff;
        if (launch_walk_kind<OP_MEDIAN, false>(shifter, a, st->ts, s)) return -1;
        k_median_decide<<<grid_for(nr, 256, 16), 256, 0, s>>>(sl.kcount.as<uint64_t>(), a.hits, nr, d_pass); ++g_launches;
        CU(cudaGetLastError());
        CU(cudaMemcpyAsync(pass + r0, d_pass, nr, cudaMemcpyDeviceToHost, s));
        if (status) CU(cudaMemcpyAsync(status + r0, d_status, nr, cudaMemcpyDeviceToHost, s));
    }
    CU(cudaStreamSynchronize(g_ctx.slot[0].stream));
    CU(cudaStreamSynchronize(g_ctx.slot[1].stream));
    CU(cudaMemcpy(g_ctx.h_scratch + 3, d_tot, sizeof(unsigned long long), cudaMemcpyDeviceToHost));
    return (int64_t)g_ctx.h_scratch[3];
}

// Random-access roofline probe (kernels.cuh, k_probe_random): ops/s of independent random 32-bit RED.OR (what = 0)
// or random 32 B sector loads (what = 1) over a scratch footprint of footprint_bytes, device-timed (best of 3).
extern "C" double gt_probe_random(int what, uint64_t footprint_bytes, uint64_t n_ops) {
    if (ensure_ctx()) return -1.0;
    if (what < 0 || what > 1 || footprint_bytes < 4096 || n_ops == 0) { fail("gt_probe_random: bad argument"); return -1.0; }
    if (cudaSetDevice(g_ctx.device) != cudaSuccess) { fail("gt_probe_random: cudaSetDevice failed"); return -1.0; }
    uint32_t* buf = nullptr;
    if (cudaMalloc(&buf, footprint_bytes) != cudaSuccess) { cudaGetLastError(); fail("gt_probe_random: out of device memory"); return -1.0; }
    cudaMemsetAsync(buf, 0, footprint_bytes, g_ctx.main);
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    const int grid = g_ctx.sms * 8;
    double best = 0;
    for (int rep = 0; rep < 4; ++rep) {  // rep 0 warms up
        cudaEventRecord(a, g_ctx.main);
        if (what == 0) k_probe_random<0><<<grid, 256, 0, g_ctx.main>>>(buf, footprint_bytes / 4, n_ops, g_ctx.d_scratch + 7);
        else k_probe_random<1><<<grid, 256, 0, g_ctx.main>>>(buf, footprint_bytes / 4, n_ops, g_ctx.d_scratch + 7);
        ++g_launches;
        cudaEventRecord(b, g_ctx.main);
        float ms = 0;
        if (cudaEventSynchronize(b) != cudaSuccess || cudaEventElapsedTime(&ms, a, b) != cudaSuccess) { best = -1; break; }
        if (rep && ms > 0) best = std::max(best, (double)n_ops / (ms * 1e-3));
    }
    cudaEventDestroy(a);
    cudaEventDestroy(b);
    cudaFree(buf);
    if (best <= 0) { fail("gt_probe_random: %s", cudaGetErrorString(cudaGetLastError())); return -1.0; }
    return best;
}

// Same as gt_median_count_at_least, reads already in HBM as ASCII (d_bases, 16-byte aligned) with device offsets
// starting at 0; d_pass (device, uint8 per read) receives the decisions.  Everything is queued on the compute stream
// (no host wait); d_kmer_total (device uint64, may be NULL) is incremented by the k-mers judged.
extern "C" int gt_median_count_at_least_dev(gt_storage* st, int shifter, int K, const void* d_bases, const void* d_offsets,
                                            uint64_t n_reads, uint64_t n_bases, uint32_t cutoff, void* d_pass, void* d_kmer_total) {
    if (ensure_ctx()) return -1;
    if (!st || !d_pass) return fail("gt_median_count_at_least_dev: NULL argument");
    if (st->world > 1) return fail("gt_median_count_at_least_dev: not available on a sharded storage");
    if (K < 1 || K > 65535) return fail("gt_median_count_at_least_dev: K=%d out of range (1..65535)", K);
    if (n_reads >= (1ull << 32)) return fail("gt_median_count_at_least_dev: more than 2^32-1 reads in one call");
    if (n_reads == 0) return 0;
    if (!d_bases || !d_offsets) return fail("gt_median_count_at_least_dev: NULL device pointer");
    if ((reinterpret_cast<uintptr_t>(d_bases) & 15) || (reinterpret_cast<uintptr_t>(d_offsets) & 7))
        return fail("gt_median_count_at_least_dev: d_bases must be 16-byte aligned and d_offsets 8-byte aligned");
    CU(cudaSetDevice(g_ctx.device));
    if (st->pend && st->pend->pending_total() && pending_flush_sync(st)) return -1;
    Slot& sl = g_ctx.slot[0];
    cudaStream_t s = g_ctx.main;
    if (sl.consumed_pending) {
        CU(cudaStreamWaitEvent(s, sl.consumed, 0));
        sl.consumed_pending = false;
    }
    CU(cudaStreamSynchronize(sl.stream));
    if (g_ctx.apply) {  // the query must see every applied update
        CU(cudaEventRecord(sl.packed, g_ctx.apply));
        CU(cudaStreamWaitEvent(s, sl.packed, 0));
    }
    const uint64_t n_words = (n_bases + 31) / 32, n_words_alloc = n_words + halo_alloc_words();
    if (sl.words.reserve(n_words_alloc * 8, s) || sl.flags.reserve(n_reads + 1, s) || sl.coarse.reserve(((n_bases >> COARSE_SHIFT) + 2) * 4, s) ||
        sl.kcount.reserve(n_reads * 8, s) || sl.hits.reserve(n_reads * 4, s))
        return -1;
    const uint64_t* offs = static_cast<const uint64_t*>(d_offsets);
    if (pack_on_device(static_cast<const uint8_t*>(d_bases), offs, n_reads, n_bases, 0, sl.words.as<uint64_t>(), n_words_alloc,
                       sl.flags.as<uint8_t>(), sl.coarse.as<uint32_t>(), s))
        return -1;
    gt_batch view;
    view.n_reads = n_reads;
    view.n_bases = n_bases;
    view.n_words = n_words;
    view.n_words_alloc = n_words_alloc;
    view.base0 = 0;
    view.d_words = sl.words.as<uint64_t>();
    view.d_offsets = const_cast<uint64_t*>(offs);
    view.d_flags = sl.flags.as<uint8_t>();
    view.d_coarse = sl.coarse.as<uint32_t>();
    unsigned long long* d_tot = d_kmer_total ? static_cast<unsigned long long*>(d_kmer_total) : g_ctx.d_scratch + 4;
    k_kmer_counts<<<grid_for(n_reads, 256, 16), 256, 0, s>>>(offs, n_reads, K, view.d_flags, sl.kcount.as<uint64_t>(), nullptr, d_tot); ++g_launches;
    CU(cudaGetLastError());
    CU(cudaMemsetAsync(sl.hits.p, 0, n_reads * 4, s));
    WalkArgs a = make_args(view, K);
    a.hits = sl.hits.as<uint32_t>();
    a.cutoff = cutoff;
    if (launch_walk_kind<OP_MEDIAN, false>(shifter, a, st->ts, s)) return -1;
    k_median_decide<<<grid_for(n_reads, 256, 16), 256, 0, s>>>(sl.kcount.as<uint64_t>(), a.hits, n_reads, static_cast<uint8_t*>(d_pass)); ++g_launches;
    CU(cudaGetLastError());
    return 0;
}

// DiginormFilter over a batch, batch-synchronous: every read of the CALL is judged
// against the table state at the start of the call (median_count_at_least, diginorm.hh:35-68), then the kept
// reads are inserted (filter_sequence, :111-119).  A call of one read reproduces the reference's serial filter
// exactly.  Returns the k-mers of all judged reads (the processor's "time": filter_sequence returns len-K+1
// whether or not the read passes); keep[r] = 1 for reads that passed (they are the ones FilterProcessor writes
// out, processors.hh:389-417); *n_kept = how many.
extern "C" int64_t gt_diginorm_sequences(gt_storage* st, int shifter, int K, const char* bases, const uint64_t* offsets,
                                          uint64_t n_reads, uint32_t cutoff, uint8_t* keep, uint64_t* n_kept) {
    if (check_reads("gt_diginorm_sequences", bases, offsets, n_reads, K)) return -1;
    if (!st || !keep) return fail("gt_diginorm_sequences: NULL argument");
    if (st->world > 1) return fail("gt_diginorm_sequences: not available on a sharded storage");
    if (pending_flush_sync(st)) return -1;
    if (validate_offsets("gt_diginorm_sequences", offsets, n_reads)) return -1;
    if (n_kept) *n_kept = 0;
    if (n_reads == 0) return 0;
    CU(cudaSetDevice(g_ctx.device));
    std::vector<uint64_t> cuts;
    chunk_ranges(offsets, n_reads, cuts);
    const size_t n_chunks = cuts.size() - 1;
    unsigned long long* d_tot = g_ctx.d_scratch + 3;   // k-mers judged
    unsigned long long* d_kept = g_ctx.d_scratch + 5;  // reads kept
    CU(cudaMemsetAsync(d_tot, 0, sizeof(unsigned long long), g_ctx.slot[0].stream));
    CU(cudaMemsetAsync(d_kept, 0, sizeof(unsigned long long), g_ctx.slot[0].stream));
    CU(cudaStreamSynchronize(g_ctx.slot[0].stream));
    // pass 0 judges every chunk; pass 1 stages the chunks again and inserts the kept reads.  With a single
    // chunk (the usual batch) the judged chunk is still resident and is inserted straight away.
    for (int pass = 0; pass < 2; ++pass) {
        for (size_t c = 0; c < n_chunks; ++c) {
            Slot& sl = g_ctx.slot[c & 1];
            cudaStream_t s = sl.stream;
            const uint64_t r0 = cuts[c], r1 = cuts[c + 1], nr = r1 - r0;
            gt_batch view;
            if (stage_chunk(sl, bases, offsets, r0, r1, view)) return -1;
            if (sl.kcount.reserve(nr * 8, s) || sl.hits.reserve(nr * 4, s) || sl.out8.reserve(nr, s)) return -1;
            uint8_t* d_keep = sl.out8.as<uint8_t>();
            if (pass == 0) {
                k_kmer_counts<<<grid_for(nr, 256, 16), 256, 0, s>>>(view.d_offsets, nr, K, view.d_flags, sl.kcount.as<uint64_t>(), nullptr, d_tot); ++g_launches;
                CU(cudaGetLastError());
                CU(cudaMemsetAsync(sl.hits.p, 0, nr * 4, s));
                WalkArgs a = make_args(view, K);
                a.hits = sl.hits.as<uint32_t>();
                a.cutoff = cutoff;
                if (launch_walk_kind<OP_MEDIAN, false>(shifter, a, st->ts, s)) return -1;
                k_diginorm_keep<<<grid_for(nr, 256, 16), 256, 0, s>>>(sl.kcount.as<uint64_t>(), a.hits, nr, view.d_flags, d_keep, d_kept); ++g_launches;
                CU(cudaGetLastError());
                CU(cudaMemcpyAsync(keep + r0, d_keep, nr, cudaMemcpyDeviceToHost, s));
                if (n_chunks > 1) continue;
            } else {
                // the judgement of pass 0 (host `keep`) comes back as flags
                CU(cudaMemcpyAsync(d_keep, keep + r0, nr, cudaMemcpyHostToDevice, s));
                k_flag_unkept<<<grid_for(nr, 256, 16), 256, 0, s>>>(d_keep, nr, view.d_flags); ++g_launches;
                CU(cudaGetLastError());
            }
            if (launch_insert(st, shifter, view, K, GT_MODE_BLIND, nullptr, s)) return -1;
        }
        CU(cudaStreamSynchronize(g_ctx.slot[0].stream));
        CU(cudaStreamSynchronize(g_ctx.slot[1].stream));
        if (n_chunks == 1) break;
    }
    CU(cudaMemcpy(g_ctx.h_scratch + 3, d_tot, sizeof(unsigned long long), cudaMemcpyDeviceToHost));
    CU(cudaMemcpy(g_ctx.h_scratch + 5, d_kept, sizeof(unsigned long long), cudaMemcpyDeviceToHost));
    if (n_kept) *n_kept = g_ctx.h_scratch[5];
    return (int64_t)g_ctx.h_scratch[3];
}

// split reads into chunks of <= lim bases (always at least one read per chunk)
static void chunk_ranges_lim(const uint64_t* offsets, uint64_t n_reads, uint64_t lim, std::vector<uint64_t>& cuts) {
    cuts.clear();
    cuts.push_back(0);
    uint64_t r = 0;
    while (r < n_reads) {
        const uint64_t start = offsets[r];
        const uint64_t* it = std::upper_bound(offsets + r + 1, offsets + n_reads + 1, start + lim);
        uint64_t r1 = (uint64_t)(it - offsets) - 1;
        if (r1 <= r) r1 = r + 1;
        cuts.push_back(r1);
        r = r1;
    }
}

// DiginormFilter::Filter::filter_sequence over a batch with the REFERENCE'S SERIAL SEMANTICS (diginorm.hh:111-119: reads
// are judged and inserted one after the other, so a read sees every earlier kept read).  Parallel rounds that give
// exactly that result:
//   judge    median_count_at_least of every undecided read against the tables (which hold every read kept so far).
//            Counts only grow, so a read that is dropped now would also be dropped in its serial turn: final.
//   claim    every tentatively kept read claims the slots of its k-mers with its read index (minimum wins);
//   verify   a tentative read that holds all its slots shares none with an EARLIER tentative read, so whatever
//            those turn out to be its counts are the serial ones: it is kept and inserted now.  The others are judged
//            again in the next round.  The first tentative read always holds its slots, so every round decides at least
//            one read; reads of unrelated sequence are decided in the first round.
// Same return values as gt_diginorm_sequences.
extern "C" int64_t gt_diginorm_sequences_serial(gt_storage* st, int shifter, int K, const char* bases, const uint64_t* offsets,
                                                 uint64_t n_reads, uint32_t cutoff, uint8_t* keep, uint64_t* n_kept) {
    if (check_reads("gt_diginorm_sequences_serial", bases, offsets, n_reads, K)) return -1;
    if (!st || !keep) return fail("gt_diginorm_sequences_serial: NULL argument");
    if (st->world > 1) return fail("gt_diginorm_sequences_serial: not available on a sharded storage");
    if (pending_flush_sync(st)) return -1;
    if (validate_offsets("gt_diginorm_sequences_serial", offsets, n_reads)) return -1;
    if (n_kept) *n_kept = 0;
    if (n_reads == 0) return 0;
    CU(cudaSetDevice(g_ctx.device));
    // a chunk's claims (one per k-mer and table) must fit the claim map
    const uint64_t lim = std::max<uint64_t>(1024, std::min<uint64_t>(chunk_bases(), ((1ull << exact_log2_cap()) / 2) / (uint64_t)st->n));
    std::vector<uint64_t> cuts;
    chunk_ranges_lim(offsets, n_reads, lim, cuts);
    Slot& sl = g_ctx.slot[0];
    cudaStream_t s = sl.stream;
    CU(cudaStreamSynchronize(g_ctx.slot[1].stream));
    unsigned long long* d_tot = g_ctx.d_scratch + 3;    // k-mers judged
    unsigned long long* d_kept = g_ctx.d_scratch + 5;   // reads kept
    unsigned long long* d_tent = g_ctx.d_scratch + 4;   // tentative keeps of a round
    CU(cudaMemsetAsync(d_tot, 0, 8, s));
    CU(cudaMemsetAsync(d_kept, 0, 8, s));
    for (size_t c = 0; c + 1 < cuts.size(); ++c) {
        const uint64_t r0 = cuts[c], r1 = cuts[c + 1], nr = r1 - r0;
        gt_batch view;
        if (stage_chunk(sl, bases, offsets, r0, r1, view)) return -1;
        // out8: [state | flags of the round | conflict | keep]
        if (sl.kcount.reserve(nr * 8, s) || sl.hits.reserve(nr * 4, s) || sl.out8.reserve(4 * nr, s)) return -1;
        uint8_t* d_state = sl.out8.as<uint8_t>();
        uint8_t* d_rflags = d_state + nr;
        uint8_t* d_conf = d_state + 2 * nr;
        uint8_t* d_keep = d_state + 3 * nr;
        const int g = grid_for(nr, 256, 16);
        k_kmer_counts<<<g, 256, 0, s>>>(view.d_offsets, nr, K, view.d_flags, sl.kcount.as<uint64_t>(), nullptr, d_tot); ++g_launches;
        CU(cudaGetLastError());
        CU(cudaMemsetAsync(d_state, RD_UNDECIDED, nr, s));
        gt_batch sub = view;
        sub.d_flags = d_rflags;
        for (uint64_t round = 0;; ++round) {
            k_rd_select<<<g, 256, 0, s>>>(d_state, view.d_flags, nr, RD_UNDECIDED, d_rflags); ++g_launches;
            CU(cudaMemsetAsync(sl.hits.p, 0, nr * 4, s));
            WalkArgs a = make_args(sub, K);
            a.hits = sl.hits.as<uint32_t>();
            a.cutoff = cutoff;
            if (launch_walk_kind<OP_MEDIAN, false>(shifter, a, st->ts, s)) return -1;
            CU(cudaMemsetAsync(d_tent, 0, 8, s));
            k_rd_judge<<<g, 256, 0, s>>>(sl.kcount.as<uint64_t>(), a.hits, nr, d_state, d_tent); ++g_launches;
            CU(cudaGetLastError());
            CU(cudaMemcpyAsync(g_ctx.h_scratch + 4, d_tent, 8, cudaMemcpyDeviceToHost, s));
            CU(cudaStreamSynchronize(s));
            if (g_ctx.h_scratch[4] == 0) break;
            k_rd_select<<<g, 256, 0, s>>>(d_state, view.d_flags, nr, RD_TENTATIVE, d_rflags); ++g_launches;
            WalkArgs b = make_args(sub, K);
            if (exact_map(view.n_bases * (uint64_t)st->n, s, b.claims)) return -1;
            if (launch_walk_kind<OP_RD_CLAIM, false>(shifter, b, st->ts, s)) return -1;
            CU(cudaMemsetAsync(d_conf, 0, nr, s));
            b.conflict = d_conf;
            if (launch_walk_kind<OP_RD_VERIFY, false>(shifter, b, st->ts, s)) return -1;
            k_rd_settle<<<g, 256, 0, s>>>(d_conf, nr, d_state); ++g_launches;
            k_rd_select<<<g, 256, 0, s>>>(d_state, view.d_flags, nr, RD_KEEP_NOW, d_rflags); ++g_launches;
            CU(cudaGetLastError());
            if (launch_insert(st, shifter, sub, K, GT_MODE_BLIND, nullptr, s)) return -1;
            k_rd_finish<<<g, 256, 0, s>>>(nr, d_state, d_keep, d_kept, 0); ++g_launches;
            CU(cudaGetLastError());
        }
        k_rd_finish<<<g, 256, 0, s>>>(nr, d_state, d_keep, d_kept, 1); ++g_launches;
        CU(cudaGetLastError());
        CU(cudaMemcpyAsync(keep + r0, d_keep, nr, cudaMemcpyDeviceToHost, s));
        CU(cudaStreamSynchronize(s));
    }
    CU(cudaMemcpy(g_ctx.h_scratch + 3, d_tot, 8, cudaMemcpyDeviceToHost));
    CU(cudaMemcpy(g_ctx.h_scratch + 5, d_kept, 8, cudaMemcpyDeviceToHost));
    if (n_kept) *n_kept = g_ctx.h_scratch[5];
    return (int64_t)g_ctx.h_scratch[3];
}

// dBG::insert_and_query_sequence over a batch (dbg.hh:327-340; the input of StreamingSolidFilter, solidifier.hh:58-76):
// every k-mer is inserted and its count AFTER its own insert is returned -- with the reference's serial semantics over
// the whole batch (reads in order, k-mers left to right): kernels.cuh, "Batched serial insert_and_query".  counts are
// laid out like gt_query_sequences.  BitStorage::insert_and_query is always 1 (bitstorage.cc:78-84).
extern "C" int64_t gt_insert_and_query_sequences(gt_storage* st, int shifter, int K, const char* bases, const uint64_t* offsets,
                                                  uint64_t n_reads, int16_t* counts, uint8_t* status) {
    if (check_reads("gt_insert_and_query_sequences", bases, offsets, n_reads, K)) return -1;
    if (!st || !counts) return fail("gt_insert_and_query_sequences: NULL argument");
    if (st->world > 1) return fail("gt_insert_and_query_sequences: not available on a sharded storage");
    if (pending_flush_sync(st)) return -1;
    if (validate_offsets("gt_insert_and_query_sequences", offsets, n_reads)) return -1;
    if (n_reads == 0) return 0;
    CU(cudaSetDevice(g_ctx.device));
    std::vector<uint64_t> cuts;
    chunk_ranges(offsets, n_reads, cuts);
    Slot& sl = g_ctx.slot[0];
    cudaStream_t s = sl.stream;
    CU(cudaStreamSynchronize(g_ctx.slot[1].stream));
    unsigned long long* d_left = g_ctx.d_scratch + 4;
    int64_t total = 0;
    for (size_t c = 0; c + 1 < cuts.size(); ++c) {
        const uint64_t r0 = cuts[c], r1 = cuts[c + 1], nr = r1 - r0;
        gt_batch view;
        if (stage_chunk(sl, bases, offsets, r0, r1, view)) return -1;
        if (sl.kcount.reserve(nr * 8, s) || sl.koff.reserve(nr * 8, s) || sl.out8.reserve(nr, s)) return -1;
        const int64_t nk = batch_kmers(view, K, sl.kcount.as<uint64_t>(), sl.out8.as<uint8_t>(), s);
        if (nk < 0) return -1;
        if (scan_counts(sl, sl.kcount.as<uint64_t>(), nr, sl.koff.as<uint64_t>(), s)) return -1;
        if (sl.out16.reserve((uint64_t)nk * 2 + 2, s)) return -1;
        if (sl.nmask.reserve(view.n_bases + 16, s)) return -1;  // done[] per position
        uint8_t* d_done = sl.nmask.as<uint8_t>();
        CU(cudaMemsetAsync(d_done, 0, view.n_bases + 1, s));
        WalkArgs a = make_args(view, K);
        a.koff = sl.koff.as<uint64_t>();
        a.counts = sl.out16.as<int16_t>();
        a.done = d_done;
        a.n_left = d_left;
        a.n_unique = st->d_n_unique;
        const uint64_t n_tiles = (view.n_bases + TILE_POS - 1) / TILE_POS;
        const uint64_t per = std::max<uint64_t>(1, ((1ull << exact_log2_cap()) / 2) / ((uint64_t)TILE_POS * st->n));
        if (direct_begin(st, s)) return -1;
        for (uint64_t t0 = 0; t0 < n_tiles; t0 += per) {  // groups of tiles in serial order, each run to completion
            a.tile_lo = t0;
            a.tile_n = std::min(per, n_tiles - t0);
            const uint64_t positions = std::min<uint64_t>(a.tile_n * TILE_POS, view.n_bases - t0 * TILE_POS);
            for (uint64_t round = 0;; ++round) {
                if (exact_map(positions * st->n, s, a.claims)) return -1;
                if (launch_walk_kind<OP_IQ_CLAIM, false>(shifter, a, st->ts, s)) return -1;
                CU(cudaMemsetAsync(d_left, 0, 8, s));
                if (launch_walk_kind<OP_IQ_APPLY, false>(shifter, a, st->ts, s)) return -1;
                CU(cudaMemcpyAsync(g_ctx.h_scratch + 4, d_left, 8, cudaMemcpyDeviceToHost, s));
                CU(cudaStreamSynchronize(s));
                if (g_ctx.h_scratch[4] == 0) break;
            }
        }
        if (direct_end(st, s)) return -1;
        if (nk) CU(cudaMemcpyAsync(counts + total, a.counts, (uint64_t)nk * 2, cudaMemcpyDeviceToHost, s));
        if (status) CU(cudaMemcpyAsync(status + r0, sl.out8.p, nr, cudaMemcpyDeviceToHost, s));
        CU(cudaStreamSynchronize(s));
        total += nk;
    }
    return total;
}

// ------------------------------------------------------------------------------------------
// hash-vector members
// ------------------------------------------------------------------------------------------
template <int KIND, bool TRACK>
static void launch_insert_hashes_t(const gt_storage* st, const uint64_t* d_h, uint64_t n, uint8_t* d_new, cudaStream_t s) {
    int grid = grid_for(n, 256 * 4, 8);
    if (st->n == 4) k_insert_hashes<KIND, TRACK, 4><<<grid, 256, 0, s>>>(d_h, n, st->ts, d_new, st->d_n_unique);
    else k_insert_hashes<KIND, TRACK, 0><<<grid, 256, 0, s>>>(d_h, n, st->ts, d_new, st->d_n_unique); ++g_launches;
}

template <int KIND, int NT>
static int launch_insert_hashes_exact_t(const gt_storage* st, const uint64_t* d_h, uint64_t n, uint8_t* d_new, const ClaimMap& cm,
                                        cudaStream_t s) {
    int grid = grid_for(n, 256 * 4, 8);
    k_claim_hashes<KIND, NT><<<grid, 256, 0, s>>>(d_h, n, st->ts, cm); ++g_launches;
    k_insert_hashes_exact<KIND, NT><<<grid, 256, 0, s>>>(d_h, n, st->ts, cm, d_new, st->d_n_unique); ++g_launches;
    CU(cudaGetLastError());
    return 0;
}
static int launch_insert_hashes_exact(const gt_storage* st, const uint64_t* d_h, uint64_t n, uint8_t* d_new, const ClaimMap& cm,
                                      cudaStream_t s) {
    const bool four = st->n == 4;
    switch (st->kind) {
        case 0: return four ? launch_insert_hashes_exact_t<0, 4>(st, d_h, n, d_new, cm, s) : launch_insert_hashes_exact_t<0, 0>(st, d_h, n, d_new, cm, s);
        case 1: return four ? launch_insert_hashes_exact_t<1, 4>(st, d_h, n, d_new, cm, s) : launch_insert_hashes_exact_t<1, 0>(st, d_h, n, d_new, cm, s);
        default: return four ? launch_insert_hashes_exact_t<2, 4>(st, d_h, n, d_new, cm, s) : launch_insert_hashes_exact_t<2, 0>(st, d_h, n, d_new, cm, s);
    }
}

extern "C" int gt_insert_hashes(gt_storage* st, const uint64_t* hashes, uint64_t n, int mode, uint8_t* is_new) {
    if (ensure_ctx()) return -1;
    if (!st || (n && !hashes)) return fail("gt_insert_hashes: NULL argument");
    if (st->world > 1) return fail("gt_insert_hashes: not available on a sharded storage");
    if (check_mode("gt_insert_hashes", mode)) return -1;
    if (is_new && mode == GT_MODE_BLIND) return fail("gt_insert_hashes: is_new needs GT_MODE_FAST or GT_MODE_EXACT");
    // is_new must see every earlier insert; and a direct table writer must not run beside an apply that merges
    // windows into the tables (bucket.cuh, K2w)
    if (pending_flush_sync(st)) return -1;
    if (n == 0) return 0;
    CU(cudaSetDevice(g_ctx.device));
    const uint64_t chunk = 64ull << 20;  // hashes per chunk
    for (uint64_t i0 = 0, c = 0; i0 < n; i0 += chunk, ++c) {
        Slot& sl = g_ctx.slot[c & 1];
        cudaStream_t s = sl.stream;
        uint64_t m = std::min(chunk, n - i0);
        if (sl.out64a.reserve(m * 8, s) || sl.out8.reserve(m, s)) return -1;
        CU(cudaMemcpyAsync(sl.out64a.p, hashes + i0, m * 8, cudaMemcpyHostToDevice, s));
        uint8_t* d_new = is_new ? sl.out8.as<uint8_t>() : nullptr;
        const uint64_t* d_h = sl.out64a.as<uint64_t>();
        bool track = mode != GT_MODE_BLIND;
        if (mode == GT_MODE_EXACT) {
            // chunks in order on the compute stream, each claimed then inserted (see launch_insert)
            CU(cudaEventRecord(sl.packed, s));
            CU(cudaStreamWaitEvent(g_ctx.main, sl.packed, 0));
            const uint64_t per = std::max<uint64_t>(1, ((1ull << exact_log2_cap()) / 2) / (uint64_t)st->n);
            for (uint64_t j0 = 0; j0 < m; j0 += per) {
                const uint64_t mm = std::min(per, m - j0);
                ClaimMap cm;
                if (exact_map(mm * st->n, g_ctx.main, cm)) return -1;
                if (launch_insert_hashes_exact(st, d_h + j0, mm, d_new ? d_new + j0 : nullptr, cm, g_ctx.main)) return -1;
            }
            CU(cudaEventRecord(sl.consumed, g_ctx.main));
            CU(cudaStreamWaitEvent(s, sl.consumed, 0));
        } else
        switch (st->kind * 2 + (track ? 1 : 0)) {
            case 0: launch_insert_hashes_t<0, false>(st, d_h, m, d_new, s); break;
            case 1: launch_insert_hashes_t<0, true>(st, d_h, m, d_new, s); break;
            case 2: launch_insert_hashes_t<1, false>(st, d_h, m, d_new, s); break;
            case 3: launch_insert_hashes_t<1, true>(st, d_h, m, d_new, s); break;
            case 4: launch_insert_hashes_t<2, false>(st, d_h, m, d_new, s); break;
            default: launch_insert_hashes_t<2, true>(st, d_h, m, d_new, s); break;
        }
        CU(cudaGetLastError());
        if (is_new) CU(cudaMemcpyAsync(is_new + i0, d_new, m, cudaMemcpyDeviceToHost, s));
    }
    CU(cudaStreamSynchronize(g_ctx.slot[0].stream));
    CU(cudaStreamSynchronize(g_ctx.slot[1].stream));
    return 0;
}

extern "C" int gt_query_hashes(gt_storage* st, const uint64_t* hashes, uint64_t n, int16_t* counts) {
    if (ensure_ctx()) return -1;
    if (!st || (n && (!hashes || !counts))) return fail("gt_query_hashes: NULL argument");
    if (st->world > 1) return fail("gt_query_hashes: not available on a sharded storage");
    if (pending_flush_sync(st)) return -1;
    if (n == 0) return 0;
    CU(cudaSetDevice(g_ctx.device));
    const uint64_t chunk = 64ull << 20;
    for (uint64_t i0 = 0, c = 0; i0 < n; i0 += chunk, ++c) {
        Slot& sl = g_ctx.slot[c & 1];
        cudaStream_t s = sl.stream;
        uint64_t m = std::min(chunk, n - i0);
        if (sl.out64a.reserve(m * 8, s) || sl.out16.reserve(m * 2, s)) return -1;
        CU(cudaMemcpyAsync(sl.out64a.p, hashes + i0, m * 8, cudaMemcpyHostToDevice, s));
        const uint64_t* d_h = sl.out64a.as<uint64_t>();
        int16_t* d_c = sl.out16.as<int16_t>();
        int grid = grid_for(m, 256 * 4, 8);
        if (st->n == 4) {
            if (st->kind == 0) k_query_hashes<0, 4><<<grid, 256, 0, s>>>(d_h, m, st->ts, d_c);
            else if (st->kind == 1) k_query_hashes<1, 4><<<grid, 256, 0, s>>>(d_h, m, st->ts, d_c);
            else k_query_hashes<2, 4><<<grid, 256, 0, s>>>(d_h, m, st->ts, d_c); ++g_launches;
        } else {
            if (st->kind == 0) k_query_hashes<0, 0><<<grid, 256, 0, s>>>(d_h, m, st->ts, d_c);
            else if (st->kind == 1) k_query_hashes<1, 0><<<grid, 256, 0, s>>>(d_h, m, st->ts, d_c);
            else k_query_hashes<2, 0><<<grid, 256, 0, s>>>(d_h, m, st->ts, d_c); ++g_launches;
        }
        CU(cudaGetLastError());
        CU(cudaMemcpyAsync(counts + i0, d_c, m * 2, cudaMemcpyDeviceToHost, s));
    }
    CU(cudaStreamSynchronize(g_ctx.slot[0].stream));
    CU(cudaStreamSynchronize(g_ctx.slot[1].stream));
    return 0;
}

#include "sketch_host.inc"
#include "fastx_host.inc"
