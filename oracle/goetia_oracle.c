/*
 * oracle/goetia_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * A scalar, single-threaded, plain-C restatement of goetia's k-mer ingest hot path
 * (SURVEY.md section 8a).  It exists so that the CUDA path can be checked bit-for-bit
 * on a box where /root/reference does not exist.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load it; the product library
 * never does.
 *
 * Parity pinning (see DESIGN.md "Oracle"):
 *   - dBG / storage / hashing: PINNED.  Checked against the reference's known-answer
 *     vector (tests/test_hashing.py:15-19) and against golden vectors generated from the
 *     compiled, unmodified reference (oracle/_ref, recipe in oracle/Makefile) -- see
 *     tests/golden/make_golden.py and tests/test_oracle.py.
 *   - SourmashSketch: PARITY UNPINNED.  The arithmetic lives in libsourmash 3.4.0 (Rust,
 *     environment_dev.yml:23), which is not under /root/reference and cannot be built
 *     here.  What is restated is its published KmerMinHash::add_sequence algorithm;
 *     only the MurmurHash3_x64_128 step is pinned (against the reference's vendored
 *     src/goetia/hashing/smhasher/MurmurHash3.cc:262 and sourmash's documented
 *     hash_murmur("ACG") == 1731421407650554201).
 *
 * Every function cites the reference file:line it follows (paths relative to
 * /root/reference).
 */
#define _POSIX_C_SOURCE 200809L
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>
#include <time.h>

/* ------------------------------------------------------------------------------------
 * Character table.  include/goetia/hashing/rollinghash/characterhash.h:27-113 holds a fixed
 * 256 x u64 table (never re-randomised: the ctor ignores maxval, :21).  Reads that reach
 * the dBG through FastxParser<DNA_SIMPLE> contain only A,C,G,T (parsing/readers.hh:162-171),
 * so the four entries below are the only ones the hot path can touch.  Values read out of
 * the compiled reference (re-checked by tests/test_oracle.py
 * against tests/golden/char_table.json).
 * ---------------------------------------------------------------------------------- */
#define ORC_T_A 16664410744025174816ull
#define ORC_T_C 15956807086001210932ull
#define ORC_T_G 9404339731978646439ull
#define ORC_T_T 836480985777824379ull

/* returns 0..3 for ACGT, -1 otherwise.  Complement code = 3 - code
 * (sequences/alphabets.hh:132-145: A<->T, C<->G). */
static inline int orc_code(unsigned char c) {
    switch (c) {
        case 'A': return 0;
        case 'C': return 1;
        case 'G': return 2;
        case 'T': return 3;
        default: return -1;
    }
}
static const uint64_t ORC_T[4] = {ORC_T_A, ORC_T_C, ORC_T_G, ORC_T_T};

static inline uint64_t rotl64(uint64_t x, unsigned r) {
    r &= 63u;
    return r ? (x << r) | (x >> (64u - r)) : x;
}
static inline uint64_t rotr64(uint64_t x, unsigned r) {
    r &= 63u;
    return r ? (x >> r) | (x << (64u - r)) : x;
}

/* ------------------------------------------------------------------------------------
 * CyclicHash<uint64_t> -- hashing/rollinghash/cyclichash.h
 *   eat            :105-108   H = rotl(H,1) ^ T[c]
 *   update         :85-92     H = rotl(H,1) ^ rotl(T[out], K%64) ^ T[in]
 *   reverse_update :96-101    H = rotr(H ^ rotl(T[out], K%64) ^ T[in], 1)
 * (fastleftshiftn :41-43 is rotl by n%64; K%64==0 is a shift-by-64 there, which on x86
 *  behaves as rotate-by-0, the mathematically consistent answer -- restated as rotl 0.)
 * ---------------------------------------------------------------------------------- */
typedef struct {
    uint64_t h;
    unsigned r; /* K % 64 */
} orc_cyclic;

static inline void cyc_eat(orc_cyclic* c, int code) { c->h = rotl64(c->h, 1) ^ ORC_T[code]; }
static inline void cyc_update(orc_cyclic* c, int out, int in) {
    c->h = rotl64(c->h, 1) ^ rotl64(ORC_T[out], c->r) ^ ORC_T[in];
}
static inline void cyc_reverse_update(orc_cyclic* c, int out, int in) {
    c->h = rotr64(c->h ^ rotl64(ORC_T[out], c->r) ^ ORC_T[in], 1);
}

/*
 * KmerIterator<Shifter> over one sequence -- hashing/kmeriterator.hh:64-76 (hash_base of the
 * first K characters), :109-123 (next: shift_right(seq[i-1], seq[i+K-1])).
 * Forward policy:  hashing/rollinghashshifter.hh:69-79 (hash_base), :103-106 (shift_right).
 * Canonical policy: :183-197 (fwd eats s[i], rc eats comp(s[K-1-i])), :203-208
 *   (shift_right: hasher.update(out,in); rc_hasher.reverse_update(comp(in), comp(out))).
 * canonical value = min(fw, rc) -- hashing/canonical.hh:124-126.
 *
 * can == 0: fills fw[]; rc may be NULL.   can == 1: fills fw[] and rc[].
 * Returns the number of k-mers (len-K+1); -1 if len < K (SequenceLengthException,
 * kmeriterator.hh:57-59); -2 if a non-ACGT byte is met.
 */
int64_t orc_hash_sequence(int can, int K, const char* seq, uint64_t len, uint64_t* fw, uint64_t* rc) {
    if (K <= 0 || len < (uint64_t)K) return -1;
    for (uint64_t i = 0; i < len; ++i)
        if (orc_code((unsigned char)seq[i]) < 0) return -2;
    orc_cyclic f = {0, (unsigned)K % 64u}, r = {0, (unsigned)K % 64u};
    for (int i = 0; i < K; ++i) {
        cyc_eat(&f, orc_code((unsigned char)seq[i]));
        if (can) cyc_eat(&r, 3 - orc_code((unsigned char)seq[K - 1 - i]));
    }
    uint64_t n = len - (uint64_t)K + 1;
    fw[0] = f.h;
    if (can) rc[0] = r.h;
    for (uint64_t i = 1; i < n; ++i) {
        int out = orc_code((unsigned char)seq[i - 1]);
        int in = orc_code((unsigned char)seq[i + K - 1]);
        cyc_update(&f, out, in);
        fw[i] = f.h;
        if (can) {
            cyc_reverse_update(&r, 3 - in, 3 - out);
            rc[i] = r.h;
        }
    }
    return (int64_t)n;
}

/* ------------------------------------------------------------------------------------
 * Table sizing -- storage/storage.hh:146-163 (is_prime, trial division) and :166-190
 * (get_n_primes_near_x: the n largest primes strictly below x, descending; x==1 -> {1}).
 * ---------------------------------------------------------------------------------- */
static int orc_is_prime(uint64_t n) {
    if (n < 2) return 0;
    if (n == 2) return 1;
    if (n % 2 == 0) return 0;
    for (uint64_t i = 3; i * i <= n; i += 2)
        if (n % i == 0) return 0;
    return 1;
}
int orc_primes_near(uint32_t n, uint64_t x, uint64_t* out) {
    int k = 0;
    if (x == 1) { out[0] = 1; return 1; }
    if (x == 0) return 0;
    uint64_t i = x - 1;
    if (i % 2 == 0) { if (i == 0) return 0; i--; }
    while ((uint32_t)k != n) {
        if (orc_is_prime(i)) out[k++] = i;
        if (i == 1) break;
        i -= 2;
    }
    return k;
}

/* ------------------------------------------------------------------------------------
 * Storages.  kind: 0 BitStorage, 1 ByteStorage, 2 NibbleStorage.
 * Allocation sizes: bitstorage.hh:143-156 (size/8+1), bytestorage.hh:125-134 (size),
 * nibblestorage.hh:166-177 (size/2+1).
 * ---------------------------------------------------------------------------------- */
#define ORC_MAX_TABLES 32
typedef struct {
    int kind;
    int n_tables;
    uint64_t sizes[ORC_MAX_TABLES];
    uint64_t nbytes[ORC_MAX_TABLES];
    uint8_t* tables[ORC_MAX_TABLES];
    uint64_t n_unique;
    uint64_t n_occupied;
} orc_storage;

static uint64_t orc_table_nbytes(int kind, uint64_t size) {
    if (kind == 0) return size / 8 + 1;
    if (kind == 1) return size;
    return size / 2 + 1;
}

orc_storage* orc_storage_create(int kind, const uint64_t* sizes, int n_tables) {
    if (kind < 0 || kind > 2 || n_tables < 1 || n_tables > ORC_MAX_TABLES) return NULL;
    orc_storage* s = (orc_storage*)calloc(1, sizeof(orc_storage));
    s->kind = kind;
    s->n_tables = n_tables;
    for (int i = 0; i < n_tables; ++i) {
        s->sizes[i] = sizes[i];
        s->nbytes[i] = orc_table_nbytes(kind, sizes[i]);
        s->tables[i] = (uint8_t*)calloc(s->nbytes[i], 1);
        if (!s->tables[i]) return NULL;
    }
    return s;
}
void orc_storage_destroy(orc_storage* s) {
    if (!s) return;
    for (int i = 0; i < s->n_tables; ++i) free(s->tables[i]);
    free(s);
}
void orc_storage_reset(orc_storage* s) {
    for (int i = 0; i < s->n_tables; ++i) memset(s->tables[i], 0, s->nbytes[i]);
    s->n_unique = 0;
    s->n_occupied = 0;
}
uint64_t orc_storage_table_bytes(orc_storage* s, int i) { return s->nbytes[i]; }
uint8_t* orc_storage_table(orc_storage* s, int i) { return s->tables[i]; }
void orc_storage_stats(orc_storage* s, uint64_t* n_unique, uint64_t* n_occupied) {
    *n_unique = s->n_unique;
    *n_occupied = s->n_occupied;
}

/* BitStorage::insert -- storage/bitstorage.hh:195-219 */
static int bit_insert(orc_storage* s, uint64_t h) {
    int is_new = 0;
    for (int i = 0; i < s->n_tables; ++i) {
        uint64_t bin = h % s->sizes[i];
        uint8_t bit = (uint8_t)(1u << (bin % 8));
        uint8_t orig = s->tables[i][bin / 8];
        s->tables[i][bin / 8] = orig | bit;
        if (!(orig & bit)) {
            if (i == 0) s->n_occupied++;
            is_new = 1;
        }
    }
    if (is_new) s->n_unique++;
    return is_new;
}
/* BitStorage::query -- src/goetia/storage/bitstorage.cc:87-100 */
static int16_t bit_query(const orc_storage* s, uint64_t h) {
    for (int i = 0; i < s->n_tables; ++i) {
        uint64_t bin = h % s->sizes[i];
        if (!(s->tables[i][bin / 8] & (1u << (bin % 8)))) return 0;
    }
    return 1;
}

/* ByteStorage::insert -- src/goetia/storage/bytestorage.cc:60-113 (bigcount off, the
 * default: storage.hh:110).  NB: is_new / occupied are only examined while is_new is still
 * false (:69-79), so occupied counts table-0 zero bins. */
static int byte_insert(orc_storage* s, uint64_t h) {
    int is_new = 0;
    for (int i = 0; i < s->n_tables; ++i) {
        uint64_t bin = h % s->sizes[i];
        uint8_t cur = s->tables[i][bin];
        if (!is_new && cur == 0) {
            is_new = 1;
            if (i == 0) s->n_occupied++;
        }
        if (cur < 255) s->tables[i][bin] = (uint8_t)(cur + 1);
    }
    if (is_new) s->n_unique++;
    return is_new;
}
/* ByteStorage::query -- bytestorage.cc:116-139 */
static int16_t byte_query(const orc_storage* s, uint64_t h) {
    int16_t m = 255;
    for (int i = 0; i < s->n_tables; ++i) {
        int16_t c = s->tables[i][h % s->sizes[i]];
        if (c < m) m = c;
    }
    return m;
}

/* NibbleStorage addressing -- storage/nibblestorage.hh:109-122: byte (bin/2); odd bin ->
 * low nibble (mask 15, shift 0), even bin -> high nibble (mask 240, shift 4).
 * insert -- src/goetia/storage/nibblestorage.cc:60-100 (saturate at 15). */
static int nib_insert(orc_storage* s, uint64_t h) {
    int is_new = 0;
    for (int i = 0; i < s->n_tables; ++i) {
        uint64_t bin = h % s->sizes[i];
        uint64_t idx = bin / 2;
        uint8_t mask = (bin % 2) ? 15 : 240;
        uint8_t shift = (bin % 2) ? 0 : 4;
        uint8_t cur = (uint8_t)((s->tables[i][idx] & mask) >> shift);
        if (!is_new && cur == 0) {
            is_new = 1;
            if (i == 0) s->n_occupied++;
        }
        if (cur == 15) continue;
        uint8_t nc = (uint8_t)((cur + 1) << shift);
        s->tables[i][idx] = (uint8_t)((s->tables[i][idx] & ~mask) | (nc & mask));
    }
    if (is_new) s->n_unique++;
    return is_new;
}
/* NibbleStorage::query -- nibblestorage.cc:112-130 */
static int16_t nib_query(const orc_storage* s, uint64_t h) {
    uint8_t m = 15;
    for (int i = 0; i < s->n_tables; ++i) {
        uint64_t bin = h % s->sizes[i];
        uint8_t mask = (bin % 2) ? 15 : 240;
        uint8_t shift = (bin % 2) ? 0 : 4;
        uint8_t c = (uint8_t)((s->tables[i][bin / 2] & mask) >> shift);
        if (c < m) m = c;
    }
    return m;
}

int orc_insert(orc_storage* s, uint64_t h) {
    return s->kind == 0 ? bit_insert(s, h) : s->kind == 1 ? byte_insert(s, h) : nib_insert(s, h);
}
int16_t orc_query(const orc_storage* s, uint64_t h) {
    return s->kind == 0 ? bit_query(s, h) : s->kind == 1 ? byte_query(s, h) : nib_query(s, h);
}
/* insert_and_query: Bit always 1 (bitstorage.cc:78-84); Byte/Nibble: 1 if new else query
 * (bytestorage.cc:142-150, nibblestorage.cc:102-109). */
int16_t orc_insert_and_query(orc_storage* s, uint64_t h) {
    int is_new = orc_insert(s, h);
    if (s->kind == 0 || is_new) return 1;
    return orc_query(s, h);
}

void orc_insert_hashes(orc_storage* s, const uint64_t* h, uint64_t n, uint8_t* is_new) {
    for (uint64_t i = 0; i < n; ++i) {
        int r = orc_insert(s, h[i]);
        if (is_new) is_new[i] = (uint8_t)r;
    }
}
void orc_query_hashes(const orc_storage* s, const uint64_t* h, uint64_t n, int16_t* counts) {
    for (uint64_t i = 0; i < n; ++i) counts[i] = orc_query(s, h[i]);
}

/* ------------------------------------------------------------------------------------
 * dBG<Storage, Shifter> sequence members -- include/goetia/dbg.hh
 *   insert_sequence(seq)            :296-305  returns len-K+1
 *   insert_sequence(seq, n_new)     :307-318
 *   query_sequence(seq)             :349-362
 *   insert_and_query_sequence(seq)  :327-340
 * The value inserted is hash_type::value(): fw for Fwd, min(fw,rc) for Can
 * (dbg.hh:144-146, canonical.hh:124-126).
 * mode: 0 insert, 1 query, 2 insert_and_query.  counts may be NULL for mode 0.
 * ---------------------------------------------------------------------------------- */
static int64_t orc_walk(orc_storage* s, int can, int K, const char* seq, uint64_t len, int mode,
                        uint64_t* n_new, int16_t* counts) {
    if (K <= 0 || len < (uint64_t)K) return -1;
    for (uint64_t i = 0; i < len; ++i)
        if (orc_code((unsigned char)seq[i]) < 0) return -2;
    orc_cyclic f = {0, (unsigned)K % 64u}, r = {0, (unsigned)K % 64u};
    for (int i = 0; i < K; ++i) {
        cyc_eat(&f, orc_code((unsigned char)seq[i]));
        if (can) cyc_eat(&r, 3 - orc_code((unsigned char)seq[K - 1 - i]));
    }
    uint64_t n = len - (uint64_t)K + 1, nn = 0;
    for (uint64_t i = 0; i < n; ++i) {
        if (i) {
            int out = orc_code((unsigned char)seq[i - 1]);
            int in = orc_code((unsigned char)seq[i + K - 1]);
            cyc_update(&f, out, in);
            if (can) cyc_reverse_update(&r, 3 - in, 3 - out);
        }
        uint64_t v = can ? (f.h < r.h ? f.h : r.h) : f.h;
        if (mode == 0) nn += (uint64_t)orc_insert(s, v);
        else if (mode == 1) counts[i] = orc_query(s, v);
        else counts[i] = orc_insert_and_query(s, v);
    }
    if (n_new) *n_new = nn;
    return (int64_t)n;
}
int64_t orc_insert_sequence(orc_storage* s, int can, int K, const char* seq, uint64_t len, uint64_t* n_new) {
    return orc_walk(s, can, K, seq, len, 0, n_new, NULL);
}
int64_t orc_query_sequence(orc_storage* s, int can, int K, const char* seq, uint64_t len, int16_t* counts) {
    return orc_walk(s, can, K, seq, len, 1, NULL, counts);
}
int64_t orc_insert_and_query_sequence(orc_storage* s, int can, int K, const char* seq, uint64_t len,
                                      int16_t* counts) {
    return orc_walk(s, can, K, seq, len, 2, NULL, counts);
}

/*
 * The reference's streaming loop: FileProcessor::advance + InserterProcessor::process_sequence
 * (processors.hh:208-229, :304-331) -- reads in order, each through insert_sequence; reads
 * shorter than K are swallowed and contribute 0 k-mers.  bases = concatenated reads,
 * offsets[n_reads+1].  n_new_per_read may be NULL.  Returns total k-mers consumed;
 * *seconds (if non-NULL) = wall time of the loop.
 */
int64_t orc_insert_reads(orc_storage* s, int can, int K, const char* bases, const uint64_t* offsets,
                         uint64_t n_reads, uint64_t* n_new_per_read, double* seconds) {
    struct timespec t0, t1;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    int64_t total = 0;
    for (uint64_t r = 0; r < n_reads; ++r) {
        uint64_t nn = 0;
        int64_t n = orc_walk(s, can, K, bases + offsets[r], offsets[r + 1] - offsets[r], 0, &nn, NULL);
        if (n_new_per_read) n_new_per_read[r] = (n > 0) ? nn : 0;
        if (n > 0) total += n;
    }
    clock_gettime(CLOCK_MONOTONIC, &t1);
    if (seconds) *seconds = (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
    return total;
}

/* counts laid out back to back: read r's counts start at kmer_offsets[r] where
 * kmer_offsets[r+1]-kmer_offsets[r] = max(0, len_r-K+1).  Returns total k-mers. */
int64_t orc_query_reads(orc_storage* s, int can, int K, const char* bases, const uint64_t* offsets,
                        uint64_t n_reads, int16_t* counts) {
    int64_t total = 0;
    for (uint64_t r = 0; r < n_reads; ++r) {
        int64_t n = orc_walk(s, can, K, bases + offsets[r], offsets[r + 1] - offsets[r], 1, NULL,
                             counts + total);
        if (n > 0) total += n;
    }
    return total;
}

/*
 * DiginormFilter::median_count_at_least -- include/goetia/diginorm.hh:35-68.
 * min_req = unsigned(0.5 + float(n_kmers)/2) (:40); true iff at least min_req k-mers have
 * query >= cutoff (the early exits do not change the answer).
 * Returns 1/0, -1 if len < K, -2 for a non-ACGT byte.
 */
int orc_median_count_at_least(orc_storage* s, int can, int K, const char* seq, uint64_t len,
                              unsigned cutoff) {
    if (K <= 0 || len < (uint64_t)K) return -1;
    uint64_t n = len - (uint64_t)K + 1;
    int16_t* counts = (int16_t*)malloc(n * sizeof(int16_t));
    int64_t rc = orc_walk(s, can, K, seq, len, 1, NULL, counts);
    if (rc < 0) { free(counts); return (int)rc; }
    unsigned min_req = (unsigned)(0.5 + (float)n / 2);
    unsigned hit = 0;
    for (uint64_t i = 0; i < n; ++i)
        if ((unsigned)counts[i] >= cutoff) ++hit;
    free(counts);
    return hit >= min_req ? 1 : 0;
}

/*
 * DiginormFilter::Filter::filter_sequence streamed over reads -- diginorm.hh:111-119 with the
 * batch-synchronous rule of SURVEY.md section 8a: reads are judged in batches of `batch`
 * against the table state at batch start, then the passing (kept) reads of the batch are
 * inserted.  batch == 1 is exactly the reference.  keep[r] = 1 if the read was kept
 * (median < cutoff) and inserted.  Returns number kept.
 */
int64_t orc_diginorm_reads(orc_storage* s, int can, int K, const char* bases, const uint64_t* offsets,
                           uint64_t n_reads, unsigned cutoff, uint64_t batch, uint8_t* keep) {
    int64_t kept = 0;
    if (batch == 0) batch = 1;
    for (uint64_t b0 = 0; b0 < n_reads; b0 += batch) {
        uint64_t b1 = b0 + batch < n_reads ? b0 + batch : n_reads;
        for (uint64_t r = b0; r < b1; ++r) {
            int m = orc_median_count_at_least(s, can, K, bases + offsets[r], offsets[r + 1] - offsets[r], cutoff);
            keep[r] = (m == 0) ? 1 : 0; /* too-short / invalid reads are dropped */
        }
        for (uint64_t r = b0; r < b1; ++r)
            if (keep[r]) {
                orc_walk(s, can, K, bases + offsets[r], offsets[r + 1] - offsets[r], 0, NULL, NULL);
                ++kept;
            }
    }
    return kept;
}

/* ------------------------------------------------------------------------------------
 * MurmurHash3_x64_128 -- the published (public-domain, Austin Appleby) algorithm, as the
 * reference vendors it at src/goetia/hashing/smhasher/MurmurHash3.cc:262 and as libsourmash
 * uses it for hash_murmur (sketches/sourmash/sourmash.hpp:19-21).  Restated from the
 * published description; pinned by tests against oracle/_ref and the documented values.
 * ---------------------------------------------------------------------------------- */
static inline uint64_t fmix64(uint64_t k) {
    k ^= k >> 33;
    k *= 0xff51afd7ed558ccdull;
    k ^= k >> 33;
    k *= 0xc4ceb9fe1a85ec53ull;
    k ^= k >> 33;
    return k;
}
void orc_murmur3_x64_128(const void* key, int len, uint32_t seed, uint64_t* out) {
    const uint8_t* data = (const uint8_t*)key;
    const int nblocks = len / 16;
    uint64_t h1 = seed, h2 = seed;
    const uint64_t c1 = 0x87c37b91114253d5ull, c2 = 0x4cf5ad432745937full;
    for (int i = 0; i < nblocks; ++i) {
        uint64_t k1, k2;
        memcpy(&k1, data + 16 * i, 8);
        memcpy(&k2, data + 16 * i + 8, 8);
        k1 *= c1; k1 = rotl64(k1, 31); k1 *= c2; h1 ^= k1;
        h1 = rotl64(h1, 27); h1 += h2; h1 = h1 * 5 + 0x52dce729;
        k2 *= c2; k2 = rotl64(k2, 33); k2 *= c1; h2 ^= k2;
        h2 = rotl64(h2, 31); h2 += h1; h2 = h2 * 5 + 0x38495ab5;
    }
    const uint8_t* tail = data + nblocks * 16;
    uint64_t k1 = 0, k2 = 0;
    int t = len & 15;
    for (int i = t - 1; i >= 8; --i) k2 ^= (uint64_t)tail[i] << (8 * (i - 8));
    if (t > 8) { k2 *= c2; k2 = rotl64(k2, 33); k2 *= c1; h2 ^= k2; }
    for (int i = (t > 8 ? 7 : t - 1); i >= 0; --i) k1 ^= (uint64_t)tail[i] << (8 * i);
    if (t > 0) { k1 *= c1; k1 = rotl64(k1, 31); k1 *= c2; h1 ^= k1; }
    h1 ^= (uint64_t)len; h2 ^= (uint64_t)len;
    h1 += h2; h2 += h1;
    h1 = fmix64(h1); h2 = fmix64(h2);
    h1 += h2; h2 += h1;
    out[0] = h1; out[1] = h2;
}

/* ------------------------------------------------------------------------------------
 * SourmashSketch::Sketch -- include/goetia/sketches/sourmash_sketch.hh:24-82.
 * PARITY UNPINNED (see header).  Restated algorithm (libsourmash 3.4.0
 * KmerMinHash::add_sequence(seq, force=true) as called at sourmash_sketch.hh:79 through
 * sketches/sourmash/sourmash.hpp:92):
 *   len < K: nothing; upper-case; each window with a byte outside ACGT is skipped;
 *   word = min(kmer, revcomp(kmer)) as byte strings; h = MurmurHash3_x64_128(word, K, seed)[0];
 *   add_hash: drop if max_hash != 0 and h > max_hash; nothing is ever kept if num == 0 and
 *   max_hash == 0; keep an ascending duplicate-free vector, truncated to the num smallest
 *   when num != 0.
 * max_hash_from_scaled -- sourmash_sketch.hh:52-61.
 * ---------------------------------------------------------------------------------- */
uint64_t orc_max_hash_from_scaled(uint64_t scaled) {
    if (scaled == 0) return 0;
    if (scaled == 1) return UINT64_MAX;
    return (uint64_t)((double)UINT64_MAX / (double)scaled);
}

typedef struct {
    int K;
    uint32_t seed;
    uint32_t num;
    uint64_t max_hash;
    uint64_t* mins;
    uint64_t n, cap;
} orc_sketch;

orc_sketch* orc_sketch_create(uint32_t num, int K, uint32_t seed, uint64_t max_hash) {
    orc_sketch* s = (orc_sketch*)calloc(1, sizeof(orc_sketch));
    s->K = K; s->seed = seed; s->num = num; s->max_hash = max_hash;
    s->cap = 1024; s->mins = (uint64_t*)malloc(s->cap * sizeof(uint64_t));
    return s;
}
void orc_sketch_destroy(orc_sketch* s) { if (s) { free(s->mins); free(s); } }
uint64_t orc_sketch_size(const orc_sketch* s) { return s->n; }
const uint64_t* orc_sketch_mins(const orc_sketch* s) { return s->mins; }

void orc_sketch_add_hash(orc_sketch* s, uint64_t h) {
    if (s->max_hash != 0 && h > s->max_hash) return;
    if (s->num == 0 && s->max_hash == 0) return;
    uint64_t cur_max = s->n ? s->mins[s->n - 1] : UINT64_MAX;
    if (!(s->n == 0 || h <= s->max_hash || h <= cur_max || s->n < s->num)) return;
    /* binary search for first element >= h */
    uint64_t lo = 0, hi = s->n;
    while (lo < hi) { uint64_t mid = (lo + hi) / 2; if (s->mins[mid] < h) lo = mid + 1; else hi = mid; }
    if (lo < s->n && s->mins[lo] == h) return;
    if (s->n == s->cap) { s->cap *= 2; s->mins = (uint64_t*)realloc(s->mins, s->cap * sizeof(uint64_t)); }
    memmove(s->mins + lo + 1, s->mins + lo, (s->n - lo) * sizeof(uint64_t));
    s->mins[lo] = h;
    s->n++;
    if (s->num != 0 && s->n > s->num) s->n--;
}

static inline unsigned char orc_upper(unsigned char c) { return (c >= 'a' && c <= 'z') ? (unsigned char)(c - 32) : c; }

/* returns len-K+1 as Sketch::insert_sequence does (sourmash_sketch.hh:80), 0 if len < K */
int64_t orc_sketch_add_sequence(orc_sketch* s, const char* seq, uint64_t len) {
    int K = s->K;
    if (len < (uint64_t)K) return 0;
    char* fwd = (char*)malloc(K);
    char* rev = (char*)malloc(K);
    uint64_t n = len - (uint64_t)K + 1;
    for (uint64_t i = 0; i < n; ++i) {
        int ok = 1;
        for (int j = 0; j < K; ++j) {
            unsigned char c = orc_upper((unsigned char)seq[i + j]);
            int code = orc_code(c);
            if (code < 0) { ok = 0; break; }
            fwd[j] = (char)c;
            rev[K - 1 - j] = "TGCA"[code];
        }
        if (!ok) continue;
        const char* word = memcmp(fwd, rev, K) <= 0 ? fwd : rev;
        uint64_t out[2];
        orc_murmur3_x64_128(word, K, s->seed, out);
        orc_sketch_add_hash(s, out[0]);
    }
    free(fwd); free(rev);
    return (int64_t)n;
}

int64_t orc_sketch_add_reads(orc_sketch* s, const char* bases, const uint64_t* offsets, uint64_t n_reads,
                             double* seconds) {
    struct timespec t0, t1;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    int64_t total = 0;
    for (uint64_t r = 0; r < n_reads; ++r)
        total += orc_sketch_add_sequence(s, bases + offsets[r], offsets[r + 1] - offsets[r]);
    clock_gettime(CLOCK_MONOTONIC, &t1);
    if (seconds) *seconds = (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
    return total;
}

/* FNV-1a-64 over a byte range: the table checksum used by the golden vectors
 * (SURVEY.md section 8c). */
uint64_t orc_fnv1a(const uint8_t* p, uint64_t n) {
    uint64_t h = 14695981039346656037ull;
    for (uint64_t i = 0; i < n; ++i) { h ^= p[i]; h *= 1099511628211ull; }
    return h;
}

/* ------------------------------------------------------------------------------------
 * OXLI v4 table files -- save: bitstorage.cc:151-189 (SAVED_HASHBITS=2),
 * bytestorage.cc:480-536 (SAVED_COUNTING_HT=1, extra use_bigcount byte + trailing
 * n_bigcounts u64), nibblestorage.cc:132-164 (SAVED_SMALLCOUNT=7).  Layout:
 *   "OXLI" | u8 version=4 | u8 type | [u8 use_bigcount: Byte only] | u32 ksize | u8 n_tables |
 *   u64 occupied_bins | per table: u64 size, raw bytes | [u64 n_bigcounts=0: Byte only]
 * Raw bytes per table as SAVED (not as allocated): Bit size/8+1, Byte size, Nibble size/2+1.
 * Returns 0 / -1.
 * ---------------------------------------------------------------------------------- */
int orc_storage_save(const orc_storage* s, const char* fn, uint16_t ksize) {
    FILE* f = fopen(fn, "wb");
    if (!f) return -1;
    uint8_t version = 4, type = s->kind == 0 ? 2 : s->kind == 1 ? 1 : 7;
    fwrite("OXLI", 1, 4, f);
    fwrite(&version, 1, 1, f);
    fwrite(&type, 1, 1, f);
    if (s->kind == 1) { uint8_t big = 0; fwrite(&big, 1, 1, f); }
    uint32_t k32 = ksize;
    fwrite(&k32, 4, 1, f);
    uint8_t nt = (uint8_t)s->n_tables;
    fwrite(&nt, 1, 1, f);
    fwrite(&s->n_occupied, 8, 1, f);
    for (int i = 0; i < s->n_tables; ++i) {
        fwrite(&s->sizes[i], 8, 1, f);
        fwrite(s->tables[i], 1, s->nbytes[i], f);
    }
    if (s->kind == 1) { uint64_t nb = 0; fwrite(&nb, 8, 1, f); }
    fclose(f);
    return 0;
}
