/*
 * oracle/ref_harness.cc -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * A thin extern "C" driver around the UNMODIFIED goetia reference, compiled from
 * the sources where they lie under /root/reference (see oracle/Makefile) into
 * oracle/_ref/libgoetia_ref.so.  Nothing from the reference is copied: this file
 * only #includes its headers and instantiates the templates that dbg.hh declares
 * `extern template` (dbg.hh:511-614), for the storage x shifter pairs on the
 * k-mer ingest hot path.
 *
 * Used for: (1) generating tests/golden/ vectors, (2) validating the plain-C
 * restatement in oracle/goetia_oracle.c, (3) the "reference" CPU baseline arm of
 * bench.py.  The product library (goetia_b200/libgoetia_b200.so) never links or
 * loads it.
 */
#include <cstdint>
#include <cstring>
#include <memory>
#include <string>
#include <thread>
#include <vector>
#include <chrono>

#include "goetia/dbg.hh"
#include "goetia/diginorm.hh"
#include "goetia/hashing/hashshifter.hh"
#include "goetia/hashing/kmeriterator.hh"
#include "goetia/hashing/smhasher/MurmurHash3.h"
#include "goetia/parsing/readers.hh"
#include "goetia/processors.hh"
#include "goetia/storage/bitstorage.hh"
#include "goetia/storage/bytestorage.hh"
#include "goetia/storage/nibblestorage.hh"
#include "goetia/storage/storage.hh"
#include "goetia/traversal/unitig_walker.hh"

using namespace goetia;

/* explicit instantiation definitions: dbg.hh:511-614 only declares these. */
#define GT_INSTANTIATE(S, H)                                   \
    template class goetia::dBG<goetia::S, goetia::H>;          \
    template class goetia::UnitigWalker<goetia::dBG<goetia::S, goetia::H>>; \
    template class goetia::KmerIterator<goetia::dBG<goetia::S, goetia::H>>;

GT_INSTANTIATE(BitStorage, FwdLemireShifter)
GT_INSTANTIATE(BitStorage, CanLemireShifter)
GT_INSTANTIATE(ByteStorage, FwdLemireShifter)
GT_INSTANTIATE(ByteStorage, CanLemireShifter)
GT_INSTANTIATE(NibbleStorage, FwdLemireShifter)
GT_INSTANTIATE(NibbleStorage, CanLemireShifter)

namespace {

struct RefGraph {
    virtual ~RefGraph() {}
    virtual int64_t insert_sequence(const std::string& s, uint64_t* n_new) = 0;
    virtual int64_t query_sequence(const std::string& s, int16_t* counts) = 0;
    virtual int64_t insert_and_query_sequence(const std::string& s, int16_t* counts) = 0;
    virtual int insert_hash(uint64_t h) = 0;
    virtual int16_t query_hash(uint64_t h) = 0;
    virtual int16_t insert_and_query_hash(uint64_t h) = 0;
    virtual void stats(uint64_t* n_unique, uint64_t* n_occupied) = 0;
    virtual uint8_t** raw() = 0;
    virtual void reset() = 0;
    virtual void save(const std::string& fn) = 0;
    virtual void load(const std::string& fn) = 0;
    virtual int64_t process_file(const std::string& fn, uint64_t* n_seqs, uint64_t* n_skipped) = 0;
    virtual int64_t advance_trace(const std::string& fn, uint64_t interval, bool strict, uint32_t min_length,
                                  uint64_t* out, uint64_t cap) = 0;
    virtual int median_count_at_least(const std::string& s, unsigned cutoff) = 0;
    virtual int64_t insert_reads_mt(const char* bases, const uint64_t* offsets, uint64_t n_reads,
                                    int n_threads) = 0;
    std::vector<uint64_t> sizes;
    int storage_kind, shifter_kind, K;
};

template <class S> struct median_helper {
    template <class G> static int run(const std::string&, unsigned, G&) { return -1; }
};
template <> struct median_helper<ByteStorage> {
    template <class G> static int run(const std::string& s, unsigned cutoff, G& g) {
        return DiginormFilter<G>::median_count_at_least(s, cutoff, g) ? 1 : 0;
    }
};
template <> struct median_helper<NibbleStorage> {
    template <class G> static int run(const std::string& s, unsigned cutoff, G& g) {
        return DiginormFilter<G>::median_count_at_least(s, cutoff, g) ? 1 : 0;
    }
};

template <class S, class H>
struct RefGraphImpl : RefGraph {
    typedef dBG<S, H> graph_t;
    std::shared_ptr<S> storage;
    std::shared_ptr<graph_t> g;

    RefGraphImpl(int K_, const std::vector<uint64_t>& sz) {
        sizes = sz;
        K = K_;
        storage = std::make_shared<S>(sz);
        g = graph_t::build(storage, (uint16_t)K_);
    }
    int64_t insert_sequence(const std::string& s, uint64_t* n_new) override {
        try {
            if (n_new) return (int64_t)g->insert_sequence(s, *n_new);
            return (int64_t)g->insert_sequence(s);
        } catch (SequenceLengthException&) { return -1; }
    }
    int64_t query_sequence(const std::string& s, int16_t* counts) override {
        try {
            auto c = g->query_sequence(s);
            std::memcpy(counts, c.data(), c.size() * sizeof(int16_t));
            return (int64_t)c.size();
        } catch (SequenceLengthException&) { return -1; }
    }
    int64_t insert_and_query_sequence(const std::string& s, int16_t* counts) override {
        try {
            auto c = g->insert_and_query_sequence(s);
            std::memcpy(counts, c.data(), c.size() * sizeof(int16_t));
            return (int64_t)c.size();
        } catch (SequenceLengthException&) { return -1; }
    }
    int insert_hash(uint64_t h) override { return storage->insert(h) ? 1 : 0; }
    int16_t query_hash(uint64_t h) override { return storage->query(h); }
    int16_t insert_and_query_hash(uint64_t h) override { return storage->insert_and_query(h); }
    void stats(uint64_t* n_unique, uint64_t* n_occupied) override {
        *n_unique = g->n_unique();
        *n_occupied = g->n_occupied();
    }
    uint8_t** raw() override { return g->get_raw(); }
    void reset() override { g->reset(); }
    void save(const std::string& fn) override { g->save(fn); }
    void load(const std::string& fn) override { g->load(fn); }
    int64_t process_file(const std::string& fn, uint64_t* n_seqs, uint64_t* n_skipped) override {
        /* the reference's own streaming driver: processors.hh:112-127 */
        auto proc = graph_t::Processor::build(g, 100000, false);
        auto parser = FastxParser<DNA_SIMPLE>::build(fn, false, 0);
        auto res = proc->process(parser);
        *n_seqs = std::get<0>(res);
        *n_skipped = parser->n_skipped();
        return (int64_t)std::get<1>(res);
    }
    int64_t advance_trace(const std::string& fn, uint64_t interval, bool strict, uint32_t min_length,
                          uint64_t* out, uint64_t cap) override {
        /* FileProcessor::advance (processors.hh:208-229) called until nothing remains; every return value
         * <n_sequences, time_total, remaining> is recorded */
        auto proc = graph_t::Processor::build(g, interval, false);
        auto parser = FastxParser<DNA_SIMPLE>::build(fn, strict, min_length);
        uint64_t n = 0;
        bool remaining = true;
        while (remaining) {
            auto res = proc->advance(parser);
            remaining = std::get<2>(res);
            if (n < cap) {
                out[3 * n] = std::get<0>(res);
                out[3 * n + 1] = std::get<1>(res);
                out[3 * n + 2] = remaining ? 1 : 0;
            }
            ++n;
        }
        return (int64_t)n;
    }
    int median_count_at_least(const std::string& s, unsigned cutoff) override {
        return median_helper<S>::run(s, cutoff, *g);
    }
    int64_t insert_reads_mt(const char* bases, const uint64_t* offsets, uint64_t n_reads,
                            int n_threads) override {
        /* BASELINE.md section 3 mode A (n_threads==1) / mode B: one dBG
         * reference-copy per thread (dbg.hh:97-101) over one shared storage. */
        std::vector<uint64_t> totals(n_threads, 0);
        auto work = [&](int t) {
            graph_t local(*g);
            uint64_t tot = 0;
            for (uint64_t r = t; r < n_reads; r += n_threads) {
                std::string s(bases + offsets[r], offsets[r + 1] - offsets[r]);
                try { tot += local.insert_sequence(s); } catch (SequenceLengthException&) {}
            }
            totals[t] = tot;
        };
        if (n_threads <= 1) { n_threads = 1; totals.resize(1); work(0); }
        else {
            std::vector<std::thread> th;
            for (int t = 0; t < n_threads; ++t) th.emplace_back(work, t);
            for (auto& x : th) x.join();
        }
        uint64_t tot = 0;
        for (auto v : totals) tot += v;
        return (int64_t)tot;
    }
};

template <class S>
RefGraph* make_graph(int shifter_kind, int K, const std::vector<uint64_t>& sz) {
    if (shifter_kind == 0) return new RefGraphImpl<S, FwdLemireShifter>(K, sz);
    return new RefGraphImpl<S, CanLemireShifter>(K, sz);
}

}  // namespace

extern "C" {

/* storage_kind: 0 BitStorage, 1 ByteStorage, 2 NibbleStorage; shifter_kind: 0 Fwd, 1 Can */
void* ref_dbg_create(int storage_kind, int shifter_kind, int K, const uint64_t* sizes, int n_tables) {
    std::vector<uint64_t> sz(sizes, sizes + n_tables);
    RefGraph* g = nullptr;
    if (storage_kind == 0) g = make_graph<BitStorage>(shifter_kind, K, sz);
    else if (storage_kind == 1) g = make_graph<ByteStorage>(shifter_kind, K, sz);
    else if (storage_kind == 2) g = make_graph<NibbleStorage>(shifter_kind, K, sz);
    if (g) { g->storage_kind = storage_kind; g->shifter_kind = shifter_kind; }
    return g;
}
void ref_dbg_destroy(void* p) { delete static_cast<RefGraph*>(p); }

int64_t ref_dbg_insert_sequence(void* p, const char* seq, uint64_t len, uint64_t* n_new) {
    return static_cast<RefGraph*>(p)->insert_sequence(std::string(seq, len), n_new);
}
int64_t ref_dbg_query_sequence(void* p, const char* seq, uint64_t len, int16_t* counts) {
    return static_cast<RefGraph*>(p)->query_sequence(std::string(seq, len), counts);
}
int64_t ref_dbg_insert_and_query_sequence(void* p, const char* seq, uint64_t len, int16_t* counts) {
    return static_cast<RefGraph*>(p)->insert_and_query_sequence(std::string(seq, len), counts);
}
int ref_dbg_insert_hash(void* p, uint64_t h) { return static_cast<RefGraph*>(p)->insert_hash(h); }
int16_t ref_dbg_query_hash(void* p, uint64_t h) { return static_cast<RefGraph*>(p)->query_hash(h); }
int16_t ref_dbg_insert_and_query_hash(void* p, uint64_t h) {
    return static_cast<RefGraph*>(p)->insert_and_query_hash(h);
}
void ref_dbg_stats(void* p, uint64_t* n_unique, uint64_t* n_occupied) {
    static_cast<RefGraph*>(p)->stats(n_unique, n_occupied);
}
/* bytes the reference allocates per table: bitstorage.hh:143-156, bytestorage.hh:125-134,
 * nibblestorage.hh:166-177 */
uint64_t ref_dbg_table_bytes(void* p, int i) {
    RefGraph* g = static_cast<RefGraph*>(p);
    uint64_t s = g->sizes[i];
    if (g->storage_kind == 0) return s / 8 + 1;
    if (g->storage_kind == 1) return s;
    return s / 2 + 1;
}
const uint8_t* ref_dbg_table(void* p, int i) { return static_cast<RefGraph*>(p)->raw()[i]; }
void ref_dbg_reset(void* p) { static_cast<RefGraph*>(p)->reset(); }
int ref_dbg_save(void* p, const char* fn) {
    try { static_cast<RefGraph*>(p)->save(fn); return 0; } catch (...) { return -1; }
}
int ref_dbg_load(void* p, const char* fn) {
    try { static_cast<RefGraph*>(p)->load(fn); return 0; } catch (...) { return -1; }
}
int64_t ref_dbg_process_file(void* p, const char* fn, uint64_t* n_seqs, uint64_t* n_skipped) {
    return static_cast<RefGraph*>(p)->process_file(fn, n_seqs, n_skipped);
}
int64_t ref_dbg_advance_trace(void* p, const char* fn, uint64_t interval, int strict, uint32_t min_length, uint64_t* out,
                              uint64_t cap) {
    try {
        return static_cast<RefGraph*>(p)->advance_trace(fn, interval, strict != 0, min_length, out, cap);
    } catch (...) { return -1; }
}
int ref_dbg_median_count_at_least(void* p, const char* seq, uint64_t len, unsigned cutoff) {
    try {
        return static_cast<RefGraph*>(p)->median_count_at_least(std::string(seq, len), cutoff);
    } catch (SequenceLengthException&) { return -2; }
}
/* returns total k-mers consumed; *seconds = wall time of the insert loop only */
int64_t ref_dbg_insert_reads(void* p, const char* bases, const uint64_t* offsets, uint64_t n_reads,
                             int n_threads, double* seconds) {
    auto t0 = std::chrono::steady_clock::now();
    int64_t n = static_cast<RefGraph*>(p)->insert_reads_mt(bases, offsets, n_reads, n_threads);
    auto t1 = std::chrono::steady_clock::now();
    if (seconds) *seconds = std::chrono::duration<double>(t1 - t0).count();
    return n;
}

/* KmerIterator over one sequence; fw/rc filled (rc only for shifter_kind 1). Returns #k-mers or -1 */
int64_t ref_hash_sequence(int shifter_kind, int K, const char* seq, uint64_t len, uint64_t* fw,
                          uint64_t* rc) {
    std::string s(seq, len);
    try {
        int64_t n = 0;
        if (shifter_kind == 0) {
            KmerIterator<FwdLemireShifter> it(s, (uint16_t)K);
            while (!it.done()) { auto h = it.next(); fw[n++] = h.value(); }
        } else {
            KmerIterator<CanLemireShifter> it(s, (uint16_t)K);
            while (!it.done()) { auto h = it.next(); fw[n] = h.fw_hash; rc[n] = h.rc_hash; ++n; }
        }
        return n;
    } catch (SequenceLengthException&) { return -1; }
}

/* static hash of the first K characters (hashshifter.hh:177-194) */
int ref_hash_kmer(int shifter_kind, int K, const char* seq, uint64_t len, uint64_t* fw, uint64_t* rc) {
    std::string s(seq, len);
    try {
        if (shifter_kind == 0) {
            auto h = FwdLemireShifter::hash(s, (uint16_t)K);
            *fw = h.value(); *rc = 0;
        } else {
            auto h = CanLemireShifter::hash(s, (uint16_t)K);
            *fw = h.fw_hash; *rc = h.rc_hash;
        }
        return 0;
    } catch (...) { return -1; }
}

int ref_primes_near(uint32_t n, uint64_t x, uint64_t* out) {
    auto v = get_n_primes_near_x(n, x);
    for (size_t i = 0; i < v.size(); ++i) out[i] = v[i];
    return (int)v.size();
}

void ref_murmur3_x64_128(const void* key, int len, uint32_t seed, uint64_t* out2) {
    murmurhash::MurmurHash3_x64_128(key, len, seed, out2);
}

/* 256-entry character table as the reference holds it (characterhash.h:27-113) */
void ref_char_table(uint64_t* out256) {
    CyclicHash<uint64_t> h(1);
    for (int i = 0; i < 256; ++i) out256[i] = h.hasher.hashvalues[i];
}

/* FastxParser<DNA_SIMPLE> walk: returns number of records, fills n_skipped; callback-free:
 * writes concatenated validated sequences into buf (if non-null, capacity cap) + offsets */
int64_t ref_parse_file(const char* fn, int strict, uint32_t min_length, char* buf, uint64_t cap,
                       uint64_t* offsets, uint64_t max_records, uint64_t* n_skipped) {
    auto parser = FastxParser<DNA_SIMPLE>::build(std::string(fn), (bool)strict, min_length);
    uint64_t n = 0, pos = 0;
    if (offsets) offsets[0] = 0;
    while (!parser->is_complete()) {
        std::optional<Record> rec;
        try { rec = parser->next(); } catch (...) { return -2; }
        if (!rec) continue;
        const std::string& s = rec.value().sequence;
        if (buf) {
            if (pos + s.size() > cap || n >= max_records) return -3;
            std::memcpy(buf + pos, s.data(), s.size());
        }
        pos += s.size();
        ++n;
        if (offsets) offsets[n] = pos;
    }
    *n_skipped = parser->n_skipped();
    return (int64_t)n;
}

}  // extern "C"
