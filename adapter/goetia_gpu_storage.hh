/* adapter/goetia_gpu_storage.hh -- the reference-side binding of libgoetia_b200: a StorageType for the UNMODIFIED
 * goetia::dBG<StorageType, ShifterType> template (include/goetia/dbg.hh:39-41).
 *
 * This is the file a goetia maintainer drops into include/goetia/storage/ (INTEGRATION.md section 2).  It models the
 * StorageType concept exactly as BitStorage / ByteStorage / NibbleStorage do -- Storage<uint64_t> (storage/storage.hh:
 * 98-143) + Tagged<S> (meta.hh:59-71) + StorageTraits<S> (bitstorage.hh:68-76) -- on top of the C ABI of
 * include/goetia_b200.h, which owns all CUDA state.  It is compiled against the reference's own headers by
 * adapter/Makefile (adapter_harness.cc instantiates goetia::dBG<GpuBitStorage, CanLemireShifter> and drives it through
 * the reference's own InserterProcessor::process); nothing of the reference is copied or modified.
 *
 * Semantics
 *   insert / query / insert_and_query (one hash): the reference's, one k-mer at a time (gt_insert_hashes with
 *     GT_MODE_EXACT = the serial first-toucher rule; bitstorage.hh:195-219, bytestorage.cc:60-150, nibblestorage.cc:60-130).
 *   defer_inserts(n): write-behind mode for the throughput of an unmodified caller.  insert() then appends the hash to a
 *     host buffer (and returns true); the buffer is sent as ONE batch (GT_MODE_EXACT, so n_unique_kmers stays the serial
 *     count) when it holds n hashes and before anything reads the storage.  Tables, n_occupied and n_unique_kmers are
 *     identical to the immediate mode; only insert()'s own return value is not meaningful while deferring.
 *   insert_many / query_many / insert_sequences / query_sequences: the batch members SURVEY.md section 8b asks for.
 *   get_raw_tables(): host mirrors, byte-identical to the CPU storages' tables after the same inserts.
 *   save / load: OXLI v4 files, byte-identical to bitstorage.cc:151-307 / bytestorage.cc:480-536 / nibblestorage.cc:132-278.
 *   Errors: the C ABI's <0 + gt_last_error() is re-thrown as GoetiaException / GoetiaFileException (goetia.hh:140-178).
 */
#ifndef GOETIA_GPU_STORAGE_HH
#define GOETIA_GPU_STORAGE_HH

#include <cstdint>
#include <cstring>
#include <fstream>
#include <memory>
#include <string>
#include <tuple>
#include <vector>

#include "goetia/goetia.hh"
#include "goetia/meta.hh"
#include "goetia/storage/storage.hh"

#include "goetia_b200.h"

namespace goetia {

template <int Kind> class GpuStorage;  // Kind = GT_STORAGE_BIT / GT_STORAGE_BYTE / GT_STORAGE_NIBBLE

template <int Kind>
struct StorageTraits<GpuStorage<Kind>> {
    static constexpr bool is_probabilistic = true;
    static constexpr bool is_counting      = Kind != GT_STORAGE_BIT;
    static constexpr int  bits_per_slot    = Kind == GT_STORAGE_BIT ? 1 : Kind == GT_STORAGE_BYTE ? 8 : 4;

    typedef std::tuple<uint64_t, uint16_t> params_type;
    static constexpr params_type default_params = std::make_tuple(1'000'000, 4);  // bitstorage.hh:75
};

template <int Kind>
class GpuStorage : public Storage<uint64_t>,
                   public Tagged<GpuStorage<Kind>>
{
protected:
    std::vector<uint64_t>                    _tablesizes;
    std::shared_ptr<gt_storage>              _h;             // deleter = gt_storage_destroy
    mutable std::vector<uint64_t>            _pending;       // write-behind buffer (defer_inserts)
    size_t                                   _defer = 0;
    mutable std::vector<std::vector<byte_t>> _mirror;        // host copies handed out by get_raw_tables()
    mutable std::vector<byte_t*>             _mirror_ptrs;

    static void check(int64_t rc) {
        if (rc < 0) throw GoetiaException(gt_last_error());
    }

    void create() {
        static bool inited = false;
        if (!inited) {
            check(gt_init(0));
            inited = true;
        }
        gt_storage* p = gt_storage_create(Kind, _tablesizes.data(), (int)_tablesizes.size());
        if (!p) throw GoetiaException(gt_last_error());
        _h.reset(p, gt_storage_destroy);
    }

    // OXLI "ht_type" byte of the reference's files (storage.hh:65-71)
    static constexpr unsigned char saved_type() {
        return Kind == GT_STORAGE_BIT ? SAVED_HASHBITS : Kind == GT_STORAGE_BYTE ? SAVED_COUNTING_HT : SAVED_SMALLCOUNT;
    }

public:
    using Storage<uint64_t>::value_type;
    using Traits = StorageTraits<GpuStorage<Kind>>;

    GpuStorage(uint64_t max_table, uint16_t N)
        : GpuStorage(get_n_primes_near_x(N, max_table))
    {
    }

    explicit GpuStorage(const std::vector<uint64_t>& tablesizes)
        : _tablesizes(tablesizes)
    {
        create();
    }

    static std::shared_ptr<GpuStorage> build(uint64_t max_table, uint16_t N) {
        return std::make_shared<GpuStorage>(max_table, N);
    }
    static std::shared_ptr<GpuStorage> build(const typename Traits::params_type& params) {
        return std::make_shared<GpuStorage>(std::get<0>(params), std::get<1>(params));
    }
    std::shared_ptr<GpuStorage> clone() const {  // same sizes, empty (bitstorage.hh:136-138)
        return std::make_shared<GpuStorage>(_tablesizes);
    }

    gt_storage* handle() const { return _h.get(); }

    // ---- write-behind -------------------------------------------------------------------------------------------
    void defer_inserts(size_t batch_hashes) {
        flush();
        _defer = batch_hashes;
        _pending.reserve(batch_hashes);
    }
    void flush() const {
        if (_pending.empty()) return;
        check(gt_insert_hashes(_h.get(), _pending.data(), _pending.size(), GT_MODE_EXACT, nullptr));
        _pending.clear();
    }

    // ---- Storage<uint64_t> (storage.hh:116-127) ---------------------------------------------------------------------
    const bool insert(value_type khash) override {
        if (_defer) {
            _pending.push_back(khash);
            if (_pending.size() >= _defer) flush();
            return true;
        }
        uint8_t is_new = 0;
        check(gt_insert_hashes(_h.get(), &khash, 1, GT_MODE_EXACT, &is_new));
        return is_new != 0;
    }

    const count_t query(value_type khash) const override {
        flush();
        count_t c = 0;
        check(gt_query_hashes(_h.get(), &khash, 1, &c));
        return c;
    }

    // bitstorage.cc:78-84 (always 1), bytestorage.cc:142-150 / nibblestorage.cc:102-109 (1 if new, else the count after)
    const count_t insert_and_query(value_type khash) override {
        flush();
        uint8_t is_new = 0;
        check(gt_insert_hashes(_h.get(), &khash, 1, GT_MODE_EXACT, &is_new));
        if (Kind == GT_STORAGE_BIT || is_new) return 1;
        count_t c = 0;
        check(gt_query_hashes(_h.get(), &khash, 1, &c));
        return c;
    }

    const uint64_t n_occupied() const override {
        flush();
        uint64_t u = 0, o = 0;
        check(gt_storage_stats(_h.get(), &u, &o));
        return o;
    }
    const uint64_t n_unique_kmers() const override {
        flush();
        uint64_t u = 0, o = 0;
        check(gt_storage_stats(_h.get(), &u, &o));
        return u;
    }

    double estimated_fp() {  // bitstorage.hh:184-192: occupancy of every table multiplied up
        double fp = n_occupied() / (double)_tablesizes[0];
        return pow(fp, (double)_tablesizes.size());
    }

    byte_t** get_raw_tables() override {
        flush();
        const size_t n = _tablesizes.size();
        _mirror.resize(n);
        _mirror_ptrs.resize(n);
        for (size_t i = 0; i < n; ++i) {
            _mirror[i].resize(gt_storage_table_bytes(_h.get(), (int)i));
            check(gt_storage_download_table(_h.get(), (int)i, _mirror[i].data()));
            _mirror_ptrs[i] = _mirror[i].data();
        }
        return _mirror_ptrs.data();
    }

    void reset() override {
        _pending.clear();
        check(gt_storage_reset(_h.get()));
    }

    std::vector<uint64_t> get_tablesizes() const { return _tablesizes; }
    const size_t n_tables() const { return _tablesizes.size(); }
    const uint64_t table_bytes(size_t i) const { return gt_storage_table_bytes(_h.get(), (int)i); }

    // ---- batch members ---------------------------------------------------------------------
    void insert_many(const uint64_t* hashes, uint64_t n, uint8_t* is_new /* may be NULL */) {
        flush();
        check(gt_insert_hashes(_h.get(), hashes, n, is_new ? GT_MODE_EXACT : GT_MODE_BLIND, is_new));
    }
    void query_many(const uint64_t* hashes, uint64_t n, count_t* counts) const {
        flush();
        check(gt_query_hashes(_h.get(), hashes, n, counts));
    }
    // reads = concatenated bytes + offsets[n_reads + 1]; shifter = GT_SHIFTER_FWD / GT_SHIFTER_CAN; returns k-mers consumed
    uint64_t insert_sequences(int shifter, uint16_t K, const char* bases, const uint64_t* offsets, uint64_t n_reads,
                              int mode = GT_MODE_BLIND, uint64_t* n_new_per_read = nullptr) {
        flush();
        const int64_t n = gt_insert_sequences(_h.get(), shifter, K, bases, offsets, n_reads, mode, n_new_per_read, nullptr);
        check(n);
        return (uint64_t)n;
    }
    uint64_t query_sequences(int shifter, uint16_t K, const char* bases, const uint64_t* offsets, uint64_t n_reads,
                             count_t* counts) const {
        flush();
        const int64_t n = gt_query_sequences(_h.get(), shifter, K, bases, offsets, n_reads, counts, nullptr);
        check(n);
        return (uint64_t)n;
    }

    // ---- OXLI v4 files ------------------------------------------------------------------------------------------------
    void save(std::string filename, uint16_t ksize) override {
        byte_t** tables = get_raw_tables();
        std::ofstream out(filename.c_str(), std::ios::binary);
        if (!out.is_open()) throw GoetiaFileException("cannot open " + filename + " for writing");
        const unsigned char version = SAVED_FORMAT_VERSION, ht_type = saved_type();
        const unsigned int save_ksize = ksize;
        const unsigned char save_n_tables = (unsigned char)_tablesizes.size();
        const unsigned long long occupied = n_occupied();
        out.write(SAVED_SIGNATURE, 4);
        out.write((const char*)&version, 1);
        out.write((const char*)&ht_type, 1);
        if (Kind == GT_STORAGE_BYTE) {
            const unsigned char use_bigcount = 0;  // off by default (storage.hh:110); bytestorage.cc:497-503
            out.write((const char*)&use_bigcount, 1);
        }
        out.write((const char*)&save_ksize, sizeof(save_ksize));
        out.write((const char*)&save_n_tables, sizeof(save_n_tables));
        out.write((const char*)&occupied, sizeof(occupied));
        for (size_t i = 0; i < _tablesizes.size(); ++i) {
            const unsigned long long size = _tablesizes[i];
            out.write((const char*)&size, sizeof(size));
            out.write((const char*)tables[i], (std::streamsize)table_bytes(i));
        }
        if (Kind == GT_STORAGE_BYTE) {
            const unsigned long long n_bigcounts = 0;
            out.write((const char*)&n_bigcounts, sizeof(n_bigcounts));
        }
        if (out.fail()) throw GoetiaFileException("error writing " + filename);
    }

    void load(std::string filename, uint16_t& ksize) override {
        std::ifstream in(filename.c_str(), std::ios::binary);
        if (!in.is_open()) throw GoetiaFileException("cannot open " + filename);
        char sig[4];
        unsigned char version = 0, ht_type = 0;
        in.read(sig, 4);
        in.read((char*)&version, 1);
        in.read((char*)&ht_type, 1);
        if (in.fail() || std::memcmp(sig, SAVED_SIGNATURE, 4) != 0)
            throw GoetiaFileException("Does not start with signature for a oxli binary file: " + filename);
        if (version != SAVED_FORMAT_VERSION) throw GoetiaFileException("Incorrect file format version");
        if (ht_type != saved_type()) throw GoetiaFileException("Incorrect file format type");
        if (Kind == GT_STORAGE_BYTE) {
            unsigned char use_bigcount = 0;
            in.read((char*)&use_bigcount, 1);
        }
        unsigned int save_ksize = 0;
        unsigned char save_n_tables = 0;
        unsigned long long occupied = 0;
        in.read((char*)&save_ksize, sizeof(save_ksize));
        in.read((char*)&save_n_tables, sizeof(save_n_tables));
        in.read((char*)&occupied, sizeof(occupied));
        if (in.fail()) throw GoetiaFileException("truncated header: " + filename);
        std::vector<uint64_t> sizes;
        std::vector<std::vector<byte_t>> tables;
        for (unsigned i = 0; i < save_n_tables; ++i) {
            unsigned long long size = 0;
            in.read((char*)&size, sizeof(size));
            const uint64_t nbytes = Kind == GT_STORAGE_BIT ? size / 8 + 1 : Kind == GT_STORAGE_BYTE ? size : size / 2 + 1;
            tables.emplace_back(nbytes);
            in.read((char*)tables.back().data(), (std::streamsize)nbytes);
            if (in.fail()) throw GoetiaFileException("truncated table: " + filename);
            sizes.push_back(size);
        }
        _pending.clear();
        if (sizes != _tablesizes) {
            _tablesizes = sizes;
            create();
        }
        for (size_t i = 0; i < tables.size(); ++i) check(gt_storage_upload_table(_h.get(), (int)i, tables[i].data()));
        check(gt_storage_set_n_unique(_h.get(), 0));  // not part of the file; the reference leaves it at 0 too
        ksize = (uint16_t)save_ksize;
    }

    // BitStorage::update_from (bitstorage.cc:103-137)
    void update_from(const GpuStorage& other) {
        flush();
        other.flush();
        check(gt_storage_update_from(_h.get(), other._h.get()));
    }
};

typedef GpuStorage<GT_STORAGE_BIT>    GpuBitStorage;
typedef GpuStorage<GT_STORAGE_BYTE>   GpuByteStorage;
typedef GpuStorage<GT_STORAGE_NIBBLE> GpuNibbleStorage;

}  // namespace goetia

#endif
