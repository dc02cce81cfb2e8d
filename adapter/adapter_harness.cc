/* adapter/adapter_harness.cc -- compiles the drop-in boundary for real: the UNMODIFIED reference templates
 * goetia::dBG<StorageType, ShifterType> (dbg.hh:39-41), KmerIterator, UnitigWalker and
 * FileProcessor / InserterProcessor (processors.hh:112-127, 304-343) instantiated over the GPU-backed StorageType of
 * adapter/goetia_gpu_storage.hh, and a small extern "C" driver so that tests/test_gpu_adapter.py can run them.
 *
 * Built by adapter/Makefile against the reference's headers where they lie under /root/reference, linked with the
 * reference's own hashing / parsing objects (oracle/_ref/obj, the same unmodified translation units the CPU baseline
 * uses) and with goetia_b200/libgoetia_b200.so.  Output: adapter/_build/libgoetia_adapter.so (git-ignored, travels to
 * the GPU box).  No reference source is copied.
 */
#include <cstdint>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

#include "goetia/dbg.hh"
#include "goetia/hashing/hashshifter.hh"
#include "goetia/hashing/kmeriterator.hh"
#include "goetia/parsing/readers.hh"
#include "goetia/processors.hh"
#include "goetia/traversal/unitig_walker.hh"

#include "goetia_gpu_storage.hh"

using namespace goetia;

/* the explicit instantiations a maintainer would add next to dbg.hh:511-614 */
#define GAD_INSTANTIATE(S, H)                                              \
    template class goetia::dBG<goetia::S, goetia::H>;                      \
    template class goetia::UnitigWalker<goetia::dBG<goetia::S, goetia::H>>; \
    template class goetia::KmerIterator<goetia::dBG<goetia::S, goetia::H>>;

GAD_INSTANTIATE(GpuBitStorage, FwdLemireShifter)
GAD_INSTANTIATE(GpuBitStorage, CanLemireShifter)
GAD_INSTANTIATE(GpuByteStorage, FwdLemireShifter)
GAD_INSTANTIATE(GpuByteStorage, CanLemireShifter)
GAD_INSTANTIATE(GpuNibbleStorage, FwdLemireShifter)
GAD_INSTANTIATE(GpuNibbleStorage, CanLemireShifter)

namespace {

thread_local std::string g_err;

struct Graph {
    virtual ~Graph() {}
    virtual int64_t process_file(const std::string& fn, uint64_t* n_seqs) = 0;
    virtual int64_t insert_sequence(const std::string& s, uint64_t* n_new) = 0;
    virtual int64_t query_sequence(const std::string& s, int16_t* counts) = 0;
    virtual int64_t insert_and_query_sequence(const std::string& s, int16_t* counts) = 0;
    virtual void stats(uint64_t* n_unique, uint64_t* n_occupied) = 0;
    virtual uint64_t table_bytes(int i) = 0;
    virtual const uint8_t* table(int i) = 0;
    virtual void save(const std::string& fn) = 0;
    virtual uint16_t load(const std::string& fn) = 0;
    virtual void reset() = 0;
    virtual void defer(uint64_t n) = 0;
};

template <class S, class H>
struct GraphImpl : Graph {
    typedef dBG<S, H> graph_t;
    std::shared_ptr<S> storage;
    std::shared_ptr<graph_t> g;
    uint16_t K;

    GraphImpl(int K_, const std::vector<uint64_t>& sizes) : K((uint16_t)K_) {
        storage = std::make_shared<S>(sizes);
        g = graph_t::build(storage, K);
    }
    int64_t process_file(const std::string& fn, uint64_t* n_seqs) override {
        // the reference's own per-read driver: FileProcessor<InserterProcessor<dBG>>::process (processors.hh:112-127)
        // (graph_t::Processor = InserterProcessor<dBG>, dbg.hh:436)
        auto proc = graph_t::Processor::build(g, 100000, false);
        auto parser = FastxParser<DNA_SIMPLE>::build(fn, false, 0);
        auto res = proc->process(parser);
        if (n_seqs) *n_seqs = std::get<0>(res);
        return (int64_t)std::get<1>(res);
    }
    int64_t insert_sequence(const std::string& s, uint64_t* n_new) override {
        if (n_new) return (int64_t)g->insert_sequence(s, *n_new);
        return (int64_t)g->insert_sequence(s);
    }
    int64_t query_sequence(const std::string& s, int16_t* counts) override {
        auto c = g->query_sequence(s);
        std::memcpy(counts, c.data(), c.size() * sizeof(int16_t));
        return (int64_t)c.size();
    }
    int64_t insert_and_query_sequence(const std::string& s, int16_t* counts) override {
        auto c = g->insert_and_query_sequence(s);
        std::memcpy(counts, c.data(), c.size() * sizeof(int16_t));
        return (int64_t)c.size();
    }
    void stats(uint64_t* n_unique, uint64_t* n_occupied) override {
        *n_unique = g->n_unique();
        *n_occupied = g->n_occupied();
    }
    uint64_t table_bytes(int i) override { return storage->table_bytes((size_t)i); }
    const uint8_t* table(int i) override { return g->get_raw()[i]; }  // dBG::get_raw (dbg.hh:224) -> get_raw_tables()
    void save(const std::string& fn) override { g->save(fn); }
    uint16_t load(const std::string& fn) override {
        uint16_t k = 0;
        storage->load(fn, k);
        return k;
    }
    void reset() override { g->reset(); }
    void defer(uint64_t n) override { storage->defer_inserts((size_t)n); }
};

template <class S>
Graph* make_shifter(int can, int K, const std::vector<uint64_t>& sizes) {
    if (can) return new GraphImpl<S, CanLemireShifter>(K, sizes);
    return new GraphImpl<S, FwdLemireShifter>(K, sizes);
}

}  // namespace

#define GAD_TRY(expr, bad)                          \
    try {                                           \
        expr;                                       \
    } catch (std::exception & e) {                  \
        g_err = e.what();                           \
        return bad;                                 \
    }

extern "C" {

const char* gad_last_error(void) { return g_err.c_str(); }

void* gad_create(int kind, int can, int K, const uint64_t* sizes, int n_tables) {
    std::vector<uint64_t> sz(sizes, sizes + n_tables);
    GAD_TRY(return kind == 0 ? make_shifter<GpuBitStorage>(can, K, sz)
                             : kind == 1 ? make_shifter<GpuByteStorage>(can, K, sz) : make_shifter<GpuNibbleStorage>(can, K, sz),
            nullptr)
}
void gad_destroy(void* h) { delete static_cast<Graph*>(h); }
int gad_defer(void* h, uint64_t n) { GAD_TRY(static_cast<Graph*>(h)->defer(n); return 0, -1) }
int64_t gad_process_file(void* h, const char* path, uint64_t* n_seqs) { GAD_TRY(return static_cast<Graph*>(h)->process_file(path, n_seqs), -1) }
int64_t gad_insert_sequence(void* h, const char* seq, uint64_t len, uint64_t* n_new) {
    GAD_TRY(return static_cast<Graph*>(h)->insert_sequence(std::string(seq, len), n_new), -1)
}
int64_t gad_query_sequence(void* h, const char* seq, uint64_t len, int16_t* counts) {
    GAD_TRY(return static_cast<Graph*>(h)->query_sequence(std::string(seq, len), counts), -1)
}
int64_t gad_insert_and_query_sequence(void* h, const char* seq, uint64_t len, int16_t* counts) {
    GAD_TRY(return static_cast<Graph*>(h)->insert_and_query_sequence(std::string(seq, len), counts), -1)
}
int gad_stats(void* h, uint64_t* n_unique, uint64_t* n_occupied) { GAD_TRY(static_cast<Graph*>(h)->stats(n_unique, n_occupied); return 0, -1) }
uint64_t gad_table_bytes(void* h, int i) { GAD_TRY(return static_cast<Graph*>(h)->table_bytes(i), 0) }
const uint8_t* gad_table(void* h, int i) { GAD_TRY(return static_cast<Graph*>(h)->table(i), nullptr) }
int gad_save(void* h, const char* path) { GAD_TRY(static_cast<Graph*>(h)->save(path); return 0, -1) }
int gad_load(void* h, const char* path) { GAD_TRY(return (int)static_cast<Graph*>(h)->load(path), -1) }
int gad_reset(void* h) { GAD_TRY(static_cast<Graph*>(h)->reset(); return 0, -1) }

}  // extern "C"
