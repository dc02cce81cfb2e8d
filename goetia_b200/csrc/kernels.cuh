// goetia_b200/csrc/kernels.cuh -- sm_100a device code for the k-mer ingest path.
//
// Nothing here is GEMM-shaped: every kernel is integer/bit work bounded by HBM sector
// traffic (random 32 B sectors of the hash tables), so the design rules that matter are
// coalesced + vectorised input loads, shared-memory staging of the packed read tile,
// fire-and-forget reductions (RED) where no return value is needed, and persistent grids
// sized from the SM count.  No tensor cores, by design.
//
// Data layout in HBM
//   packed reads : flat 2-bit stream, base p of the batch at bits 2*(p%32) of u64 word p/32,
//                  A=0 C=1 G=2 T=3 (complement = 3-code); reads are concatenated and
//                  described by offsets[n_reads+1] (in bases), flags[n_reads] (bit1 = read
//                  holds a non-ACGT byte) and coarse[p/256] = index of the read holding
//                  base 256*(p/256).
//   BitStorage   : table i = size_i bits, bin b -> bit (b%8) of byte b/8  == bit (b%32) of
//                  little-endian u32 word b/32 (bitstorage.hh:199-203)
//   ByteStorage  : table i = size_i bytes, bin b -> byte b (bytestorage.cc:66-67)
//   NibbleStorage: bin b -> byte b/2, odd b = low nibble, even b = high nibble
//                  (nibblestorage.hh:109-122)
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace gt {

constexpr int MAX_TABLES = 32;
constexpr int TILE_THREADS = 256;
constexpr int POS_PER_THREAD = 32;                       // one packed u64 word per thread
constexpr int TILE_POS = TILE_THREADS * POS_PER_THREAD;  // 8192 k-mer start positions per tile
constexpr int COARSE_SHIFT = 8;                          // coarse read index every 256 bases

constexpr uint8_t READ_SHORT = 1, READ_INVALID = 2;

// Lemire character table entries for A,C,G,T (hashing/rollinghash/characterhash.h:27-113;
// the only four a validated read can touch -- SURVEY.md section 8a row a1).
__host__ __device__ __forceinline__ uint64_t lemire_T(int code) {
    return code == 0 ? 16664410744025174816ull
         : code == 1 ? 15956807086001210932ull
         : code == 2 ? 9404339731978646439ull
                     : 836480985777824379ull;
}

__host__ __device__ __forceinline__ uint64_t rotl64(uint64_t x, unsigned r) {
    r &= 63u;
    return r ? (x << r) | (x >> (64u - r)) : x;
}
__host__ __device__ __forceinline__ uint64_t rotl1(uint64_t x) { return (x << 1) | (x >> 63); }
__host__ __device__ __forceinline__ uint64_t rotr1(uint64_t x) { return (x >> 1) | (x << 63); }

// Exact h % d for every 64-bit h and 1 <= d <= 2^63, with m = floor((2^64-1)/d):
// q' = mulhi(h, m) is floor(h/d) or one less, so one conditional subtract fixes it.
// (Replaces the run-time 64-bit division of bitstorage.hh:199 / bytestorage.cc:66.)
__host__ __device__ __forceinline__ uint64_t fastmod_u64(uint64_t h, uint64_t d, uint64_t m) {
#ifdef __CUDA_ARCH__
    uint64_t q = __umul64hi(h, m);
#else
    uint64_t q = (uint64_t)(((unsigned __int128)h * m) >> 64);
#endif
    uint64_t r = h - q * d;
    return r >= d ? r - d : r;
}

struct TableSet {
    int n;
    int kind;
    uint64_t size[MAX_TABLES];
    uint64_t magic[MAX_TABLES];
    uint32_t* ptr[MAX_TABLES];
};

// ------------------------------------------------------------------------------------------
// Storage primitives.  MODE: 0 = blind (no return value wanted), 1 = tracked (is_new).
// ------------------------------------------------------------------------------------------
template <int KIND, bool TRACK>
__device__ __forceinline__ bool slot_insert(uint32_t* __restrict__ tbl, uint64_t bin) {
    if constexpr (KIND == 0) {
        // BitStorage::insert, bitstorage.hh:198-211: atomic OR of one bit.
        uint32_t* w = tbl + (bin >> 5);
        uint32_t mask = 1u << (bin & 31);
        if constexpr (TRACK) {
            uint32_t old = atomicOr(w, mask);
            return !(old & mask);
        } else {
            atomicOr(w, mask);  // result unused -> RED.E.OR (fire and forget)
            return false;
        }
    } else {
        // ByteStorage::insert (bytestorage.cc:60-113) / NibbleStorage::insert
        // (nibblestorage.cc:60-100): saturating increment of an 8- / 4-bit field.  No
        // sub-word atomics exist, so CAS on the containing 32-bit word; the final value is
        // exactly min(max, hits) however the updates interleave.
        uint32_t* w;
        unsigned sh;
        uint32_t fmax;
        if constexpr (KIND == 1) {
            w = tbl + (bin >> 2);
            sh = (unsigned)(bin & 3) * 8u;
            fmax = 255u;
        } else {
            w = tbl + (bin >> 3);
            sh = (unsigned)((bin >> 1) & 3) * 8u + ((bin & 1) ? 0u : 4u);
            fmax = 15u;
        }
        uint32_t old = __ldcg(w);
        uint32_t f;
        while (true) {
            f = (old >> sh) & fmax;
            if (f == fmax) break;
            uint32_t prev = atomicCAS(w, old, old + (1u << sh));
            if (prev == old) break;
            old = prev;
        }
        return f == 0;
    }
}

template <int KIND>
__device__ __forceinline__ uint32_t slot_query(const uint32_t* __restrict__ tbl, uint64_t bin) {
    if constexpr (KIND == 0) {
        return (__ldg(tbl + (bin >> 5)) >> (bin & 31)) & 1u;
    } else if constexpr (KIND == 1) {
        return __ldg(reinterpret_cast<const uint8_t*>(tbl) + bin);
    } else {
        uint32_t b = __ldg(reinterpret_cast<const uint8_t*>(tbl) + (bin >> 1));
        return (bin & 1) ? (b & 15u) : (b >> 4);
    }
}

// Storage::insert over all tables.  Returns is_new under the "atomic winner" rule when TRACK.
template <int KIND, bool TRACK, int NT>
__device__ __forceinline__ bool storage_insert(const TableSet& ts, uint64_t h) {
    bool is_new = false;
    if constexpr (NT > 0) {
#pragma unroll
        for (int i = 0; i < NT; ++i) {
            uint64_t bin = fastmod_u64(h, ts.size[i], ts.magic[i]);
            is_new |= slot_insert<KIND, TRACK>(ts.ptr[i], bin);
        }
    } else {
        for (int i = 0; i < ts.n; ++i) {
            uint64_t bin = fastmod_u64(h, ts.size[i], ts.magic[i]);
            is_new |= slot_insert<KIND, TRACK>(ts.ptr[i], bin);
        }
    }
    return is_new;
}

// Storage::query: AND of bits (bitstorage.cc:87-100) or min of counters clamped to max
// (bytestorage.cc:116-139, nibblestorage.cc:112-130).
template <int KIND, int NT>
__device__ __forceinline__ uint32_t storage_query(const TableSet& ts, uint64_t h) {
    uint32_t v[NT > 0 ? NT : 1];
    uint32_t acc = KIND == 0 ? 1u : KIND == 1 ? 255u : 15u;
    if constexpr (NT > 0) {
#pragma unroll
        for (int i = 0; i < NT; ++i)
            v[i] = slot_query<KIND>(ts.ptr[i], fastmod_u64(h, ts.size[i], ts.magic[i]));
#pragma unroll
        for (int i = 0; i < NT; ++i) acc = KIND == 0 ? (acc & v[i]) : min(acc, v[i]);
    } else {
        for (int i = 0; i < ts.n; ++i) {
            uint32_t x = slot_query<KIND>(ts.ptr[i], fastmod_u64(h, ts.size[i], ts.magic[i]));
            acc = KIND == 0 ? (acc & x) : min(acc, x);
        }
    }
    return acc;
}

// ------------------------------------------------------------------------------------------
// GT_MODE_EXACT: the serial "first toucher" rule.  The reference inserts
// k-mers one at a time; k-mer j is new iff some table's slot is still zero when j arrives, i.e.
// the slot was zero before this launch AND no earlier k-mer of the launch touches it.  Two
// passes over the same positions reproduce that for any interleaving of threads:
//   pass 1 (OP_CLAIM)        every k-mer whose slot (t, bin) is zero posts its serial ordinal
//                            into a claim map keyed by (t, bin); the map keeps the minimum;
//   pass 2 (OP_INSERT_EXACT) a k-mer is new iff it holds the winning claim of one of its
//                            slots; then the slots are updated (blind).  Pass 2 never reads a
//                            table, so its updates cannot disturb another thread's decision.
// The claim map is an open-addressed table in HBM: keys by CAS, ordinals by atomicMin.
// ------------------------------------------------------------------------------------------
struct ClaimMap {
    unsigned long long* keys;  // CLAIM_EMPTY when free
    uint32_t* ords;            // 0xFFFFFFFF when free
    uint64_t mask;             // capacity - 1 (capacity is a power of two >= 2 x claims)
};
constexpr unsigned long long CLAIM_EMPTY = ~0ull;

__device__ __forceinline__ uint64_t claim_key(int table, uint64_t bin) { return (bin << 5) | (uint64_t)table; }  // MAX_TABLES == 32
__device__ __forceinline__ uint64_t claim_mix(uint64_t k) {
    k ^= k >> 33; k *= 0xff51afd7ed558ccdull; k ^= k >> 33; k *= 0xc4ceb9fe1a85ec53ull; k ^= k >> 33;
    return k;
}
__device__ __forceinline__ void claim_post(const ClaimMap& m, uint64_t key, uint32_t ord) {
    uint64_t s = claim_mix(key) & m.mask;
    while (true) {
        unsigned long long cur = m.keys[s];
        if (cur == CLAIM_EMPTY) cur = atomicCAS(m.keys + s, CLAIM_EMPTY, (unsigned long long)key);
        if (cur == CLAIM_EMPTY || cur == key) {
            atomicMin(m.ords + s, ord);
            return;
        }
        s = (s + 1) & m.mask;
    }
}
// ordinal holding the claim of `key`, or 0xFFFFFFFF when nobody claimed it
__device__ __forceinline__ uint32_t claim_winner(const ClaimMap& m, uint64_t key) {
    uint64_t s = claim_mix(key) & m.mask;
    while (true) {
        unsigned long long cur = m.keys[s];
        if (cur == key) return m.ords[s];
        if (cur == CLAIM_EMPTY) return 0xFFFFFFFFu;
        s = (s + 1) & m.mask;
    }
}

// pass 1 for one k-mer
template <int KIND, int NT>
__device__ __forceinline__ void storage_claim(const TableSet& ts, const ClaimMap& m, uint64_t h, uint32_t ord) {
    const int nt = NT > 0 ? NT : ts.n;
#pragma unroll
    for (int i = 0; i < nt; ++i) {
        const uint64_t bin = fastmod_u64(h, ts.size[i], ts.magic[i]);
        if (slot_query<KIND>(ts.ptr[i], bin) == 0) claim_post(m, claim_key(i, bin), ord);
    }
}
// pass 2 for one k-mer: Storage::insert with the serial is_new
template <int KIND, int NT>
__device__ __forceinline__ bool storage_insert_exact(const TableSet& ts, const ClaimMap& m, uint64_t h, uint32_t ord) {
    const int nt = NT > 0 ? NT : ts.n;
    bool is_new = false;
#pragma unroll
    for (int i = 0; i < nt; ++i) {
        const uint64_t bin = fastmod_u64(h, ts.size[i], ts.magic[i]);
        is_new |= claim_winner(m, claim_key(i, bin)) == ord;
        slot_insert<KIND, false>(ts.ptr[i], bin);
    }
    return is_new;
}

// ------------------------------------------------------------------------------------------
// Batched serial insert_and_query (dbg.hh:327-340 over Storage::insert_and_query, bytestorage.cc:142-150,
// nibblestorage.cc:102-109).  The reference inserts k-mers one at a time and returns each k-mer's count AFTER its own
// insert, so k-mer j sees every earlier k-mer of the batch that shares one of its slots (SURVEY.md section 8a:
// count_j = min_i min(max, pre_i + rank_j(i, bin) + 1)).  Rounds reproduce that for any thread interleaving: in a round
// every unprocessed k-mer posts its position as a claim on each of its slots (the map keeps the minimum); a k-mer that
// holds ALL its slots has no unprocessed predecessor on any of them, so the tables show exactly the serial state it
// would see: it reads its counts, increments, and is done.  The smallest unprocessed position always wins, so every
// round makes progress; the number of rounds is the largest multiplicity of a slot within the batch.
// ------------------------------------------------------------------------------------------
template <int KIND, int NT>
__device__ __forceinline__ void storage_claim_always(const TableSet& ts, const ClaimMap& m, uint64_t h, uint32_t ord) {
    const int nt = NT > 0 ? NT : ts.n;
#pragma unroll
    for (int i = 0; i < nt; ++i) claim_post(m, claim_key(i, fastmod_u64(h, ts.size[i], ts.magic[i])), ord);
}
template <int KIND, int NT>
__device__ __forceinline__ bool storage_holds_all(const TableSet& ts, const ClaimMap& m, uint64_t h, uint32_t ord) {
    const int nt = NT > 0 ? NT : ts.n;
    bool all = true;
#pragma unroll
    for (int i = 0; i < nt; ++i) all &= claim_winner(m, claim_key(i, fastmod_u64(h, ts.size[i], ts.magic[i]))) == ord;
    return all;
}
// Storage::insert_and_query of a k-mer that holds all its slots this round: the count after its own insert
template <int KIND, int NT>
__device__ __forceinline__ uint32_t storage_insert_and_query(const TableSet& ts, uint64_t h, bool& is_new) {
    const int nt = NT > 0 ? NT : ts.n;
    constexpr uint32_t fmax = KIND == 0 ? 1u : KIND == 1 ? 255u : 15u;
    uint32_t acc = fmax;
    is_new = false;  // Storage::insert's return value: some table's slot was empty (counts towards n_unique)
#pragma unroll
    for (int i = 0; i < nt; ++i) {
        const uint64_t bin = fastmod_u64(h, ts.size[i], ts.magic[i]);
        // the slot is this k-mer's alone in this round; neighbours in the same word may change, hence the atomic update
        uint32_t v;
        if constexpr (KIND == 0) v = 1u;
        else if constexpr (KIND == 1) v = __ldcg(reinterpret_cast<const uint8_t*>(ts.ptr[i]) + bin);
        else {
            const uint32_t b = __ldcg(reinterpret_cast<const uint8_t*>(ts.ptr[i]) + (bin >> 1));
            v = (bin & 1) ? (b & 15u) : (b >> 4);
        }
        if constexpr (KIND != 0) {
            is_new |= v == 0u;
            v = min(fmax, v + 1u);
        }
        acc = min(acc, v);
        const bool was_empty = slot_insert<KIND, KIND == 0>(ts.ptr[i], bin);
        if constexpr (KIND == 0) is_new |= was_empty;
    }
    return acc;
}

// ------------------------------------------------------------------------------------------
// K0: validate + 2-bit pack.  One thread per output word (32 bases = 2 x 16 B loads).
// Folds a/c/g/t to upper case exactly as DNA_SIMPLE::_validate (sequences/alphabets.hh:112-130);
// any other byte flags its read GT_READ_INVALID (the parser would skip it:
// parsing/readers.hh:162-171).
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t pack8(uint64_t x, uint32_t& bad, uint32_t& badbits) {
    // per byte: code = ((c>>1)&3) ^ ((c>>2)&1)  maps A,C,G,T (either case) to 0,1,2,3
    uint64_t code = ((x >> 1) & 0x0303030303030303ull) ^ ((x >> 2) & 0x0101010101010101ull);
    // validity: (c & 0xDF) must be one of 'A' 'C' 'G' 'T'
    uint64_t u = x & 0xDFDFDFDFDFDFDFDFull;
    const uint64_t L = 0x7F7F7F7F7F7F7F7Full;
    uint64_t ok = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        uint64_t pat = (k == 0 ? 0x41ull : k == 1 ? 0x43ull : k == 2 ? 0x47ull : 0x54ull) * 0x0101010101010101ull;
        uint64_t t = u ^ pat;
        ok |= ~(((t & L) + L) | t | L);  // 0x80 in exactly the zero bytes of t
    }
    bad |= (ok != 0x8080808080808080ull);
    {   // gather "byte j is invalid" into bit j
        uint64_t y = (~ok >> 7) & 0x0101010101010101ull;
        y = (y | (y >> 7)) & 0x0003000300030003ull;
        y = (y | (y >> 14)) & 0x0000000F0000000Full;
        y = (y | (y >> 28)) & 0xFFull;
        badbits = (uint32_t)y;
    }
    // gather the 8 two-bit codes into 16 bits
    code = (code | (code >> 6)) & 0x000F000F000F000Full;
    code = (code | (code >> 12)) & 0x000000FF000000FFull;
    code = (code | (code >> 24)) & 0xFFFFull;
    return (uint32_t)code;
}

__device__ __forceinline__ uint64_t find_read(const uint64_t* __restrict__ offsets, uint64_t n_reads, uint64_t p) {
    // largest r with offsets[r] <= p and offsets[r+1] > p
    uint64_t lo = 0, hi = n_reads;  // invariant: offsets[lo] <= p < offsets[hi]
    while (hi - lo > 1) {
        uint64_t mid = (lo + hi) >> 1;
        if (__ldg(offsets + mid) <= p) lo = mid; else hi = mid;
    }
    return lo;
}

__global__ void __launch_bounds__(256) k_pack(const uint8_t* __restrict__ ascii, uint64_t n_bases,
                                               const uint64_t* __restrict__ offsets, uint64_t n_reads,
                                               uint64_t base0, uint64_t* __restrict__ words, uint64_t n_words,
                                               uint8_t* __restrict__ flags, uint32_t* __restrict__ nmask) {
    for (uint64_t w = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; w < n_words;
         w += (uint64_t)gridDim.x * blockDim.x) {
        uint64_t p = w * 32;
        uint64_t x[4];
        if (p + 32 <= n_bases) {
            const uint4* src = reinterpret_cast<const uint4*>(ascii + p);
            uint4 a = __ldcs(src), b = __ldcs(src + 1);
            x[0] = (uint64_t)a.x | ((uint64_t)a.y << 32);
            x[1] = (uint64_t)a.z | ((uint64_t)a.w << 32);
            x[2] = (uint64_t)b.x | ((uint64_t)b.y << 32);
            x[3] = (uint64_t)b.z | ((uint64_t)b.w << 32);
        } else {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                uint64_t v = 0;
                for (int j = 0; j < 8; ++j) {
                    uint64_t q = p + 8 * k + j;
                    uint64_t c = q < n_bases ? ascii[q] : (uint64_t)'A';
                    v |= c << (8 * j);
                }
                x[k] = v;
            }
        }
        uint32_t bad = 0, nm = 0;
        uint64_t out = 0;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            uint32_t bb;
            out |= (uint64_t)pack8(x[k], bad, bb) << (16 * k);
            nm |= bb << (8 * k);
        }
        words[w] = out;
        if (nmask) nmask[w] = nm;
        if (bad && flags) {
            // rare path: flag every read owning an offending byte
            for (int j = 0; j < 32; ++j) {
                uint64_t q = p + j;
                if (q >= n_bases) break;
                uint8_t c = (uint8_t)(x[j >> 3] >> (8 * (j & 7))) & 0xDF;
                if (c != 'A' && c != 'C' && c != 'G' && c != 'T')
                    flags[find_read(offsets, n_reads, q + base0)] = READ_INVALID;
            }
        }
    }
}

// coarse[g] = index of the read holding base g*256.  One thread per read.
__global__ void __launch_bounds__(256) k_coarse(const uint64_t* __restrict__ offsets, uint64_t n_reads,
                                                 uint64_t base0, uint32_t* __restrict__ coarse) {
    for (uint64_t r = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; r < n_reads;
         r += (uint64_t)gridDim.x * blockDim.x) {
        uint64_t s = offsets[r] - base0, e = offsets[r + 1] - base0;
        for (uint64_t g = (s + 255) >> COARSE_SHIFT; (g << COARSE_SHIFT) < e; ++g) coarse[g] = (uint32_t)r;
    }
}

// per read: k-mer count (0 for short / invalid reads) and status byte; block-reduced total.
__global__ void __launch_bounds__(256) k_kmer_counts(const uint64_t* __restrict__ offsets, uint64_t n_reads, int K,
                                                      const uint8_t* __restrict__ flags, uint64_t* __restrict__ kcount,
                                                      uint8_t* __restrict__ status, unsigned long long* __restrict__ total) {
    uint64_t mine = 0;
    for (uint64_t r = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; r < n_reads;
         r += (uint64_t)gridDim.x * blockDim.x) {
        uint64_t len = offsets[r + 1] - offsets[r];
        uint8_t st = flags ? (flags[r] & READ_INVALID) : 0;
        if (len < (uint64_t)K) st |= READ_SHORT;
        uint64_t c = st ? 0 : len - (uint64_t)K + 1;
        if (kcount) kcount[r] = c;
        if (status) status[r] = st;
        mine += c;
    }
    for (int o = 16; o; o >>= 1) mine += __shfl_down_sync(0xffffffffu, mine, o);
    if ((threadIdx.x & 31) == 0 && mine) atomicAdd(total, (unsigned long long)mine);
}

// ------------------------------------------------------------------------------------------
// Exclusive scan of u64 (read k-mer counts -> output offsets).  Three small passes; n is
// the number of reads, so this is never on the critical path.
// ------------------------------------------------------------------------------------------
constexpr int SCAN_BLOCK = 256;
constexpr int SCAN_ITEMS = 8;
__global__ void __launch_bounds__(SCAN_BLOCK) k_scan_partials(const uint64_t* __restrict__ in, uint64_t n,
                                                               uint64_t* __restrict__ partial) {
    __shared__ uint64_t sm[SCAN_BLOCK / 32];
    uint64_t base = (uint64_t)blockIdx.x * SCAN_BLOCK * SCAN_ITEMS;
    uint64_t s = 0;
    for (int k = 0; k < SCAN_ITEMS; ++k) {
        uint64_t i = base + (uint64_t)k * SCAN_BLOCK + threadIdx.x;
        if (i < n) s += in[i];
    }
    for (int o = 16; o; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint64_t t = 0;
        for (int w = 0; w < SCAN_BLOCK / 32; ++w) t += sm[w];
        partial[blockIdx.x] = t;
    }
}
__global__ void __launch_bounds__(1024) k_scan_top(uint64_t* __restrict__ partial, uint64_t n_blocks) {
    // one CTA of 1024 threads; thread t owns a contiguous chunk of the block partials
    __shared__ uint64_t warp_sums[32];
    const uint64_t per = (n_blocks + 1023) / 1024;
    const uint64_t b0 = (uint64_t)threadIdx.x * per;
    const uint64_t b1 = b0 + per < n_blocks ? b0 + per : n_blocks;
    uint64_t s = 0;
    for (uint64_t b = b0; b < b1; ++b) s += partial[b];
    uint64_t incl = s;
    for (int o = 1; o < 32; o <<= 1) {
        uint64_t t = __shfl_up_sync(0xffffffffu, incl, o);
        if ((threadIdx.x & 31) >= o) incl += t;
    }
    if ((threadIdx.x & 31) == 31) warp_sums[threadIdx.x >> 5] = incl;
    __syncthreads();
    uint64_t run = incl - s;
    for (int w = 0; w < (int)(threadIdx.x >> 5); ++w) run += warp_sums[w];
    for (uint64_t b = b0; b < b1; ++b) { uint64_t v = partial[b]; partial[b] = run; run += v; }
}
__global__ void __launch_bounds__(SCAN_BLOCK) k_scan_final(const uint64_t* __restrict__ in, uint64_t n,
                                                            const uint64_t* __restrict__ partial,
                                                            uint64_t* __restrict__ out) {
    // thread t owns SCAN_ITEMS consecutive items so a block-level scan of thread sums suffices
    __shared__ uint64_t warp_sums[SCAN_BLOCK / 32];
    uint64_t base = (uint64_t)blockIdx.x * SCAN_BLOCK * SCAN_ITEMS + (uint64_t)threadIdx.x * SCAN_ITEMS;
    uint64_t v[SCAN_ITEMS], s = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) { v[k] = (base + k < n) ? in[base + k] : 0; s += v[k]; }
    uint64_t incl = s;
    for (int o = 1; o < 32; o <<= 1) {
        uint64_t t = __shfl_up_sync(0xffffffffu, incl, o);
        if ((threadIdx.x & 31) >= o) incl += t;
    }
    if ((threadIdx.x & 31) == 31) warp_sums[threadIdx.x >> 5] = incl;
    __syncthreads();
    uint64_t woff = 0;
    for (int w = 0; w < (int)(threadIdx.x >> 5); ++w) woff += warp_sums[w];
    uint64_t run = partial[blockIdx.x] + woff + incl - s;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) {
        if (base + k < n) out[base + k] = run;
        run += v[k];
    }
}

// ------------------------------------------------------------------------------------------
// K1+K2/K3: the tile walker.  One CTA = 8192 consecutive k-mer start positions of the flat
// packed stream; the tile's words (plus a K-1 halo) are staged in shared memory with 16 B
// loads; each thread owns one packed word = 32 consecutive windows, seeds the two cyclic
// hashes from the first window (K "eat" steps, cyclichash.h:105-108) and then rolls
// (update :85-92 / reverse_update :96-101, as CanLemirePolicy::shift_right does,
// rollinghashshifter.hh:203-208).  Windows that cross a read boundary, or lie in a skipped
// read, are hashed (the roll must continue) but not used.
//
// OP: 0 insert, 1 query -> counts, 2 hash -> fw/rc out, 3 median hits (count >= cutoff per read),
//     4 / 5 the two passes of GT_MODE_EXACT over the tiles [tile_lo, tile_lo + tile_n)
// ------------------------------------------------------------------------------------------
//     6 / 7 the claim / apply rounds of the batched serial insert_and_query (see k_walk), 8 / 9 the claim / verify
//     passes of the serial-equivalent diginorm (read-level conflicts)
enum { OP_INSERT = 0, OP_QUERY = 1, OP_HASH = 2, OP_MEDIAN = 3, OP_CLAIM = 4, OP_INSERT_EXACT = 5,
       OP_IQ_CLAIM = 6, OP_IQ_APPLY = 7, OP_RD_CLAIM = 8, OP_RD_VERIFY = 9 };

struct WalkArgs {
    const uint64_t* words;     // packed bases
    uint64_t n_words_alloc;    // words that may be read (>= ceil(n_bases/32))
    const uint64_t* offsets;   // n_reads+1, in bases, absolute (base0 is subtracted)
    const uint8_t* flags;      // n_reads
    const uint32_t* coarse;    // ceil(n_bases/256)
    uint64_t base0;            // offsets[first read of this batch]
    uint64_t n_reads;
    uint64_t n_bases;
    int K;
    // outputs
    const uint64_t* koff;          // OP_QUERY / OP_HASH: exclusive scan of per-read k-mer counts
    int16_t* counts;               // OP_QUERY
    uint64_t* fw;                  // OP_HASH
    uint64_t* rc;                  // OP_HASH (CAN only)
    uint32_t* hits;                // OP_MEDIAN: per read
    uint32_t cutoff;               // OP_MEDIAN
    unsigned long long* n_unique;  // OP_INSERT tracked
    uint64_t* n_new;               // OP_INSERT tracked, per read (optional)
    // GT_MODE_EXACT: tile range of this launch (tile_n == 0: all tiles); ordinal = position - tile_lo * TILE_POS
    uint64_t tile_lo, tile_n;
    ClaimMap claims;
    // OP_IQ_*: done[p] = 1 once the k-mer at position p has been inserted-and-queried; n_left counts the k-mers a round
    // had to leave for the next one.  OP_RD_VERIFY: conflict[r] = 1 when read r shares a slot with an earlier claimed read.
    uint8_t* done;
    uint8_t* conflict;
    unsigned long long* n_left;
};

template <int OP, int KIND, bool CAN, bool TRACK, int NT>
__global__ void __launch_bounds__(TILE_THREADS)
k_walk(const __grid_constant__ WalkArgs a, const __grid_constant__ TableSet ts) {
    extern __shared__ __align__(16) uint64_t smem[];
    // tab[0..3] = {T[c], rotl(T[3-c],K)} for the incoming base, tab[4..7] = {rotl(T[c],K), T[3-c]} for the outgoing
    ulonglong2* tab = reinterpret_cast<ulonglong2*>(smem);
    uint64_t* sw = smem + 16;
    const int K = a.K;
    const int halo_words = ((K - 1 + 31) >> 5) + 1;
    const int tile_words = TILE_THREADS + halo_words;
    const int tid = threadIdx.x;
    if (tid < 4) {
        tab[tid] = make_ulonglong2(lemire_T(tid), rotl64(lemire_T(3 - tid), (unsigned)K));
        tab[4 + tid] = make_ulonglong2(rotl64(lemire_T(tid), (unsigned)K), lemire_T(3 - tid));
    }
    uint64_t n_tiles = (a.n_bases + TILE_POS - 1) / TILE_POS;
    if (a.tile_n && a.tile_lo + a.tile_n < n_tiles) n_tiles = a.tile_lo + a.tile_n;
    unsigned long long block_new = 0, iq_new = 0;
    constexpr bool COUNT_NEW = (OP == OP_INSERT && TRACK) || OP == OP_INSERT_EXACT;
    constexpr bool COUNT_LEFT = OP == OP_IQ_APPLY;

    for (uint64_t tile = a.tile_lo + blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        __syncthreads();  // previous tile's smem fully consumed (and tab visible)
        const uint64_t w0 = tile * TILE_THREADS;
        // stage: 16 B vector loads of the packed words (2 words per load)
        for (int i = tid * 2; i < tile_words; i += TILE_THREADS * 2) {
            uint64_t gi = w0 + i;
            if (gi + 1 < a.n_words_alloc) {
                ulonglong2 v = *reinterpret_cast<const ulonglong2*>(a.words + gi);
                sw[i] = v.x;
                if (i + 1 < tile_words) sw[i + 1] = v.y;
            } else {
                sw[i] = gi < a.n_words_alloc ? a.words[gi] : 0;
                if (i + 1 < tile_words) sw[i + 1] = 0;
            }
        }
        __syncthreads();

        const uint64_t p0 = tile * TILE_POS + (uint64_t)tid * POS_PER_THREAD;
        if (p0 < a.n_bases) {
            // seed both hashes from the window at p0
            uint64_t fw = 0, rc = 0;
            for (int j = 0; j < K; ++j) {
                int cf = (int)((sw[tid + (j >> 5)] >> (2 * (j & 31))) & 3);
                fw = rotl1(fw) ^ tab[cf].x;
                if (CAN) {
                    int jr = K - 1 - j;
                    int cr = (int)((sw[tid + (jr >> 5)] >> (2 * (jr & 31))) & 3);
                    rc = rotl1(rc) ^ tab[4 + cr].y;
                }
            }
            // outgoing bases: my own word; incoming bases: 64 bits starting at base p0+K-1
            const uint64_t wo = sw[tid];
            uint64_t win;
            {
                int a0 = tid + ((K - 1) >> 5);
                unsigned sh = 2u * (unsigned)((K - 1) & 31);
                win = sh ? (sw[a0] >> sh) | (sw[a0 + 1] << (64u - sh)) : sw[a0];
            }
            // locate the read holding p0
            uint64_t r = __ldg(a.coarse + (p0 >> COARSE_SHIFT));
            uint64_t rend = __ldg(a.offsets + r + 1) - a.base0;
            {
                int steps = 0;
                while (rend <= p0) {
                    if (++steps > 8) {  // pathological run of tiny reads: finish by bisection
                        r = find_read(a.offsets, a.n_reads, p0 + a.base0);
                        rend = __ldg(a.offsets + r + 1) - a.base0;
                        break;
                    }
                    ++r;
                    rend = __ldg(a.offsets + r + 1) - a.base0;
                }
            }
            uint64_t rstart = __ldg(a.offsets + r) - a.base0;
            bool rok = !(__ldg(a.flags + r) & READ_INVALID);
            uint32_t acc = 0;  // per-read accumulator (n_new or median hits) flushed on read change

#pragma unroll 4
            for (int i = 0; i < POS_PER_THREAD; ++i) {
                const uint64_t p = p0 + i;
                if (p >= a.n_bases) break;
                if (i) {
                    int out = (int)((wo >> (2 * (i - 1))) & 3);
                    int in = (int)((win >> (2 * i)) & 3);
                    ulonglong2 ti = tab[in], to = tab[4 + out];
                    fw = rotl1(fw) ^ to.x ^ ti.x;
                    if (CAN) rc = rotr1(rc ^ ti.y ^ to.y);
                }
                if (p >= rend) {
                    if ((OP == OP_MEDIAN || COUNT_NEW) && acc) {
                        if (OP == OP_MEDIAN) atomicAdd(a.hits + r, acc);
                        else if (a.n_new) atomicAdd((unsigned long long*)(a.n_new + r), (unsigned long long)acc);
                        acc = 0;
                    }
                    do {
                        ++r;
                        rstart = rend;
                        rend = __ldg(a.offsets + r + 1) - a.base0;
                    } while (p >= rend);
                    rok = !(__ldg(a.flags + r) & READ_INVALID);
                }
                if (rok && p + (uint64_t)K <= rend) {
                    const uint64_t h = CAN ? (fw < rc ? fw : rc) : fw;  // Canonical::value(), canonical.hh:124-126
                    if constexpr (OP == OP_INSERT) {
                        bool nw = storage_insert<KIND, TRACK, NT>(ts, h);
                        if (TRACK) acc += nw, block_new += nw;
                    } else if constexpr (OP == OP_CLAIM) {
                        storage_claim<KIND, NT>(ts, a.claims, h, (uint32_t)(p - a.tile_lo * TILE_POS));
                    } else if constexpr (OP == OP_INSERT_EXACT) {
                        bool nw = storage_insert_exact<KIND, NT>(ts, a.claims, h, (uint32_t)(p - a.tile_lo * TILE_POS));
                        acc += nw, block_new += nw;
                    } else if constexpr (OP == OP_IQ_CLAIM) {
                        if (!a.done[p]) storage_claim_always<KIND, NT>(ts, a.claims, h, (uint32_t)p);
                    } else if constexpr (OP == OP_IQ_APPLY) {
                        if (!a.done[p]) {
                            if (storage_holds_all<KIND, NT>(ts, a.claims, h, (uint32_t)p)) {
                                bool nw;
                                a.counts[a.koff[r] + (p - rstart)] = (int16_t)storage_insert_and_query<KIND, NT>(ts, h, nw);
                                a.done[p] = 1;
                                iq_new += nw;
                            } else {
                                ++block_new;  // left for the next round
                            }
                        }
                    } else if constexpr (OP == OP_RD_CLAIM) {
                        storage_claim_always<KIND, NT>(ts, a.claims, h, (uint32_t)r);
                    } else if constexpr (OP == OP_RD_VERIFY) {
                        if (!storage_holds_all<KIND, NT>(ts, a.claims, h, (uint32_t)r)) a.conflict[r] = 1;
                    } else if constexpr (OP == OP_QUERY) {
                        a.counts[a.koff[r] + (p - rstart)] = (int16_t)storage_query<KIND, NT>(ts, h);
                    } else if constexpr (OP == OP_HASH) {
                        uint64_t o = a.koff[r] + (p - rstart);
                        if (CAN && !a.rc) a.fw[o] = h;  // value() only (routed queries on a sharded storage)
                        else {
                            a.fw[o] = fw;
                            if (CAN) a.rc[o] = rc;
                        }
                    } else {
                        acc += storage_query<KIND, NT>(ts, h) >= a.cutoff;
                    }
                }
            }
            if ((OP == OP_MEDIAN || COUNT_NEW) && acc) {
                if (OP == OP_MEDIAN) atomicAdd(a.hits + r, acc);
                else if (a.n_new) atomicAdd((unsigned long long*)(a.n_new + r), (unsigned long long)acc);
            }
        }
    }
    if constexpr (COUNT_NEW || COUNT_LEFT) {
        for (int o = 16; o; o >>= 1) block_new += __shfl_down_sync(0xffffffffu, block_new, o);
        if ((tid & 31) == 0 && block_new) atomicAdd(COUNT_LEFT ? a.n_left : a.n_unique, block_new);
    }
    if constexpr (OP == OP_IQ_APPLY) {  // k-mers that were new to the storage count towards n_unique, as Storage::insert does
        for (int o = 16; o; o >>= 1) iq_new += __shfl_down_sync(0xffffffffu, iq_new, o);
        if ((tid & 31) == 0 && iq_new && a.n_unique) atomicAdd(a.n_unique, iq_new);
    }
}

// median_count_at_least decision per read (diginorm.hh:40): min_req = unsigned(0.5 + float(n)/2)
__global__ void __launch_bounds__(256) k_median_decide(const uint64_t* __restrict__ kcount, const uint32_t* __restrict__ hits,
                                                        uint64_t n_reads, uint8_t* __restrict__ pass) {
    for (uint64_t r = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; r < n_reads;
         r += (uint64_t)gridDim.x * blockDim.x) {
        uint64_t n = kcount[r];
        unsigned min_req = (unsigned)(0.5 + (double)((float)n / 2.0f));
        pass[r] = (n > 0 && hits[r] >= min_req) ? 1 : 0;
    }
}

// DiginormFilter::Filter::filter_sequence (diginorm.hh:111-119), the keep decision of a batch: a read is kept
// (and then inserted) iff it is walkable and median_count_at_least is false.  Reads that are not kept get the
// READ_INVALID flag so that the insert walk that follows skips them.
__global__ void __launch_bounds__(256) k_diginorm_keep(const uint64_t* __restrict__ kcount, const uint32_t* __restrict__ hits,
                                                        uint64_t n_reads, uint8_t* __restrict__ flags, uint8_t* __restrict__ keep,
                                                        unsigned long long* __restrict__ n_kept) {
    unsigned long long mine = 0;
    for (uint64_t r = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; r < n_reads;
         r += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t n = kcount[r];
        const unsigned min_req = (unsigned)(0.5 + (double)((float)n / 2.0f));
        const bool k = n > 0 && hits[r] < min_req;
        keep[r] = k ? 1 : 0;
        if (!k) flags[r] |= READ_INVALID;
        mine += k;
    }
    for (int o = 16; o; o >>= 1) mine += __shfl_down_sync(0xffffffffu, mine, o);
    if ((threadIdx.x & 31) == 0 && mine) atomicAdd(n_kept, mine);
}

// Serial-equivalent diginorm (capi.cu, gt_diginorm_sequences_serial): per-read state machine helpers.
//   state: 0 undecided, 1 kept (inserted), 2 dropped, 3 tentative keep of this round, 4 kept in this round (to insert)
enum { RD_UNDECIDED = 0, RD_KEPT = 1, RD_DROPPED = 2, RD_TENTATIVE = 3, RD_KEEP_NOW = 4 };
// flags_out[r] = 0 for the reads in state `want` (and valid), READ_INVALID otherwise: the walkers then see only those reads
__global__ void __launch_bounds__(256) k_rd_select(const uint8_t* __restrict__ state, const uint8_t* __restrict__ flags, uint64_t n_reads,
                                                    uint8_t want, uint8_t* __restrict__ flags_out) {
    for (uint64_t r = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; r < n_reads; r += (uint64_t)gridDim.x * blockDim.x)
        flags_out[r] = (state[r] == want && !(flags[r] & READ_INVALID)) ? 0 : READ_INVALID;
}
// after the median walk over the undecided reads: dropped (final: counts only grow) or tentative keep
__global__ void __launch_bounds__(256) k_rd_judge(const uint64_t* __restrict__ kcount, const uint32_t* __restrict__ hits, uint64_t n_reads,
                                                   uint8_t* __restrict__ state, unsigned long long* __restrict__ n_tentative) {
    unsigned long long mine = 0;
    for (uint64_t r = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; r < n_reads; r += (uint64_t)gridDim.x * blockDim.x) {
        if (state[r] != RD_UNDECIDED) continue;
        const uint64_t n = kcount[r];
        const unsigned min_req = (unsigned)(0.5 + (double)((float)n / 2.0f));
        const bool keep = n > 0 && hits[r] < min_req;
        state[r] = keep ? RD_TENTATIVE : RD_DROPPED;
        mine += keep;
    }
    for (int o = 16; o; o >>= 1) mine += __shfl_down_sync(0xffffffffu, mine, o);
    if ((threadIdx.x & 31) == 0 && mine) atomicAdd(n_tentative, mine);
}
// after the verify walk: a tentative read without a conflict is kept now; the others are judged again next round
__global__ void __launch_bounds__(256) k_rd_settle(const uint8_t* __restrict__ conflict, uint64_t n_reads, uint8_t* __restrict__ state) {
    for (uint64_t r = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; r < n_reads; r += (uint64_t)gridDim.x * blockDim.x)
        if (state[r] == RD_TENTATIVE) state[r] = conflict[r] ? RD_UNDECIDED : RD_KEEP_NOW;
}
__global__ void __launch_bounds__(256) k_rd_finish(uint64_t n_reads, uint8_t* __restrict__ state, uint8_t* __restrict__ keep,
                                                    unsigned long long* __restrict__ n_kept, int final_pass) {
    unsigned long long mine = 0;
    for (uint64_t r = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; r < n_reads; r += (uint64_t)gridDim.x * blockDim.x) {
        if (state[r] == RD_KEEP_NOW) state[r] = RD_KEPT;
        if (final_pass) {
            keep[r] = state[r] == RD_KEPT ? 1 : 0;
            mine += state[r] == RD_KEPT;
        }
    }
    if (final_pass) {
        for (int o = 16; o; o >>= 1) mine += __shfl_down_sync(0xffffffffu, mine, o);
        if ((threadIdx.x & 31) == 0 && mine) atomicAdd(n_kept, mine);
    }
}

__global__ void __launch_bounds__(256) k_flag_unkept(const uint8_t* __restrict__ keep, uint64_t n_reads, uint8_t* __restrict__ flags) {
    for (uint64_t r = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; r < n_reads; r += (uint64_t)gridDim.x * blockDim.x)
        if (!keep[r]) flags[r] |= READ_INVALID;
}

// ------------------------------------------------------------------------------------------
// Hash-vector entry points (Storage::insert / query on raw hash values; the shape of the
// reference's own storage micro-benchmark, src/goetia/benchmarks/bench_storage.cc:17-61).
// ------------------------------------------------------------------------------------------
template <int KIND, bool TRACK, int NT>
__global__ void __launch_bounds__(256) k_insert_hashes(const uint64_t* __restrict__ hashes, uint64_t n,
                                                        const __grid_constant__ TableSet ts,
                                                        uint8_t* __restrict__ is_new, unsigned long long* n_unique) {
    unsigned long long mine = 0;
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        bool nw = storage_insert<KIND, TRACK, NT>(ts, __ldcs(hashes + i));
        if (TRACK) {
            if (is_new) is_new[i] = nw;
            mine += nw;
        }
    }
    if (TRACK) {
        for (int o = 16; o; o >>= 1) mine += __shfl_down_sync(0xffffffffu, mine, o);
        if ((threadIdx.x & 31) == 0 && mine) atomicAdd(n_unique, mine);
    }
}

// GT_MODE_EXACT over a hash vector: ordinal = index in the launch (see storage_claim).
template <int KIND, int NT>
__global__ void __launch_bounds__(256) k_claim_hashes(const uint64_t* __restrict__ hashes, uint64_t n,
                                                       const __grid_constant__ TableSet ts, const ClaimMap m) {
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
        storage_claim<KIND, NT>(ts, m, hashes[i], (uint32_t)i);
}
template <int KIND, int NT>
__global__ void __launch_bounds__(256) k_insert_hashes_exact(const uint64_t* __restrict__ hashes, uint64_t n,
                                                              const __grid_constant__ TableSet ts, const ClaimMap m,
                                                              uint8_t* __restrict__ is_new, unsigned long long* n_unique) {
    unsigned long long mine = 0;
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        bool nw = storage_insert_exact<KIND, NT>(ts, m, hashes[i], (uint32_t)i);
        if (is_new) is_new[i] = nw;
        mine += nw;
    }
    for (int o = 16; o; o >>= 1) mine += __shfl_down_sync(0xffffffffu, mine, o);
    if ((threadIdx.x & 31) == 0 && mine) atomicAdd(n_unique, mine);
}

template <int KIND, int NT>
__global__ void __launch_bounds__(256) k_query_hashes(const uint64_t* __restrict__ hashes, uint64_t n,
                                                       const __grid_constant__ TableSet ts, int16_t* __restrict__ counts) {
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
        counts[i] = (int16_t)storage_query<KIND, NT>(ts, __ldcs(hashes + i));
}

// Storage::query restricted to the slot ranges [lo_t, hi_t) this rank holds (sharded storage):
// tables whose slot lies elsewhere contribute the neutral element, so the MIN over all ranks'
// answers is the reference's AND / min over all tables.
struct OwnRange { uint64_t lo[MAX_TABLES], hi[MAX_TABLES]; };
template <int KIND>
__global__ void __launch_bounds__(256) k_query_hashes_local(const uint64_t* __restrict__ hashes, uint64_t n,
                                                             const __grid_constant__ TableSet ts,
                                                             const __grid_constant__ OwnRange own, int16_t* __restrict__ counts) {
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t h = __ldcs(hashes + i);
        uint32_t acc = KIND == 0 ? 1u : KIND == 1 ? 255u : 15u;
        for (int t = 0; t < ts.n; ++t) {
            const uint64_t bin = fastmod_u64(h, ts.size[t], ts.magic[t]);
            if (bin >= own.lo[t] && bin < own.hi[t]) {
                const uint32_t x = slot_query<KIND>(ts.ptr[t], bin);
                acc = KIND == 0 ? (acc & x) : min(acc, x);
            }
        }
        counts[i] = (int16_t)acc;
    }
}

// ------------------------------------------------------------------------------------------
// Owner-routed requests on a sharded storage ("queries: same forward route; owners answer with the
// bit / count, 1 B per message"; tracked inserts: first-toucher flags travel back the same way).
// A request = (table << 59 | global slot).  Rank q holds slots [q * spr[t], (q + 1) * spr[t]) of table t.
// ------------------------------------------------------------------------------------------
struct RouteArgs { uint64_t spr[MAX_TABLES]; };
__global__ void __launch_bounds__(256) k_route_requests(const uint64_t* __restrict__ hashes, uint64_t n, const __grid_constant__ TableSet ts,
                                                         const __grid_constant__ RouteArgs ra, unsigned long long* __restrict__ req,
                                                         int32_t* __restrict__ owner) {
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t h = hashes[i];
        for (int t = 0; t < ts.n; ++t) {
            const uint64_t bin = fastmod_u64(h, ts.size[t], ts.magic[t]);
            req[i * ts.n + t] = ((unsigned long long)t << 59) | bin;
            owner[i * ts.n + t] = (int32_t)(bin / ra.spr[t]);
        }
    }
}
template <int KIND>
__global__ void __launch_bounds__(256) k_answer_requests(const unsigned long long* __restrict__ req, uint64_t n, const __grid_constant__ TableSet ts,
                                                          uint8_t* __restrict__ ans) {
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        const unsigned long long r = req[i];
        ans[i] = (uint8_t)slot_query<KIND>(ts.ptr[(int)(r >> 59)], r & ((1ull << 59) - 1));
    }
}
// tracked insert of routed requests, serial first-toucher rule over the ordinals the sources assigned (GT_MODE_EXACT's
// two passes, kernels.cuh "GT_MODE_EXACT", on messages instead of k-mers)
template <int KIND>
__global__ void __launch_bounds__(256) k_claim_requests(const unsigned long long* __restrict__ req, const uint32_t* __restrict__ ord, uint64_t n,
                                                         const __grid_constant__ TableSet ts, const ClaimMap m) {
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        const unsigned long long r = req[i];
        const int t = (int)(r >> 59);
        const uint64_t bin = r & ((1ull << 59) - 1);
        if (slot_query<KIND>(ts.ptr[t], bin) == 0) claim_post(m, claim_key(t, bin), ord[i]);
    }
}
template <int KIND>
__global__ void __launch_bounds__(256) k_insert_requests(const unsigned long long* __restrict__ req, const uint32_t* __restrict__ ord, uint64_t n,
                                                          const __grid_constant__ TableSet ts, const ClaimMap m, uint8_t* __restrict__ first) {
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        const unsigned long long r = req[i];
        const int t = (int)(r >> 59);
        const uint64_t bin = r & ((1ull << 59) - 1);
        if (first) first[i] = ord && claim_winner(m, claim_key(t, bin)) == ord[i] ? 1 : 0;
        slot_insert<KIND, false>(ts.ptr[t], bin);
    }
}

// number of non-zero slots of a table (n_occupied == non-zero slots of table 0, because the
// reference bumps _occupied_bins exactly when a table-0 slot leaves zero: bitstorage.hh:205-208,
// bytestorage.cc:71-77, nibblestorage.cc:75-81).
template <int KIND>
__global__ void __launch_bounds__(256) k_count_occupied(const uint32_t* __restrict__ tbl, uint64_t n_words,
                                                         uint64_t n_slots, unsigned long long* out) {
    unsigned long long mine = 0;
    for (uint64_t w = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; w < n_words; w += (uint64_t)gridDim.x * blockDim.x) {
        uint32_t x = tbl[w];
        if (KIND == 0) {
            mine += __popc(x);
        } else if (KIND == 1) {
            mine += ((x & 0xffu) != 0) + ((x & 0xff00u) != 0) + ((x & 0xff0000u) != 0) + ((x & 0xff000000u) != 0);
        } else {
            uint32_t y = (x | (x >> 1) | (x >> 2) | (x >> 3)) & 0x11111111u;
            mine += __popc(y);
        }
    }
    (void)n_slots;
    for (int o = 16; o; o >>= 1) mine += __shfl_down_sync(0xffffffffu, mine, o);
    if ((threadIdx.x & 31) == 0 && mine) atomicAdd(out, mine);
}

// Position-weighted checksum of a table: sum over 32-bit words of word * mix(index) mod 2^64 (mix = the
// 64-bit finaliser of the word index, forced odd).  Order-independent to compute, sensitive to every bit and
// to where it sits; used to compare multi-GB tables (two insert paths, shards vs whole) without moving them.
__host__ __device__ __forceinline__ uint64_t checksum_weight(uint64_t i) {
    uint64_t k = i + 0x9e3779b97f4a7c15ull;
    k ^= k >> 33; k *= 0xff51afd7ed558ccdull; k ^= k >> 33; k *= 0xc4ceb9fe1a85ec53ull; k ^= k >> 33;
    return k | 1ull;
}
__global__ void __launch_bounds__(256) k_checksum(const uint32_t* __restrict__ tbl, uint64_t n_words, uint64_t word0,
                                                   unsigned long long* out) {
    unsigned long long mine = 0;
    for (uint64_t w = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; w < n_words; w += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t x = tbl[w];
        if (x) mine += (unsigned long long)x * checksum_weight(word0 + w);
    }
    for (int o = 16; o; o >>= 1) mine += __shfl_down_sync(0xffffffffu, mine, o);
    if ((threadIdx.x & 31) == 0 && mine) atomicAdd(out, mine);
}

// ------------------------------------------------------------------------------------------
// Counter-based synthetic reads for harnesses (bench.py, the full-size parity tests): the base at GLOBAL
// index i of stream `seed` is a pure function of (seed, i), so the same reads exist on the CPU (numpy:
// goetia_b200/synth.py), on one GPU and on any rank of a sharded run.  word w = i / 32:
//   z = splitmix64(seed + (w + 1) * 0x9E3779B97F4A7C15);  code(i) = (z >> 2*(i%32)) & 3;  byte = "ACGT"[code]
// ------------------------------------------------------------------------------------------
__host__ __device__ __forceinline__ uint64_t synth_word(uint64_t seed, uint64_t w) {
    uint64_t z = seed + (w + 1) * 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
// one thread = 16 output bytes (out must be 16-byte aligned)
__global__ void __launch_bounds__(256) k_synth_bases(uint8_t* __restrict__ out, uint64_t n_bases, uint64_t seed, uint64_t first) {
    const uint64_t n16 = (n_bases + 15) / 16;
    for (uint64_t g = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; g < n16; g += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t i0 = first + g * 16;
        uint64_t w = i0 >> 5, z = synth_word(seed, w);
        uint32_t v[4] = {0, 0, 0, 0};
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            const uint64_t i = i0 + j;
            if ((i >> 5) != w) { w = i >> 5; z = synth_word(seed, w); }
            const uint32_t code = (uint32_t)(z >> (2 * (i & 31))) & 3u;
            const uint32_t ch = (0x54474341u >> (8 * code)) & 0xffu;  // "ACGT"
            v[j >> 2] |= ch << (8 * (j & 3));
        }
        if (g * 16 + 16 <= n_bases) {
            *reinterpret_cast<uint4*>(out + g * 16) = make_uint4(v[0], v[1], v[2], v[3]);
        } else {
            for (uint64_t j = 0; g * 16 + j < n_bases; ++j) out[g * 16 + j] = (uint8_t)(v[j >> 2] >> (8 * (j & 3)));
        }
    }
}

// Roofline probes (harness helpers): independent random 32-bit RED.OR (WHAT = 0) or random 32-byte sector loads
// (WHAT = 1) over a footprint of n_words 32-bit words -- the "random-sector roofline" R_rand the insert / query
// rates are quoted against, measured in the same run on the same device.
template <int WHAT>
__global__ void __launch_bounds__(256) k_probe_random(uint32_t* __restrict__ buf, uint64_t n_words, uint64_t n_ops, unsigned long long* sink) {
    unsigned long long acc = 0;
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n_ops; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t z = synth_word(0x5eedull, i);
        const uint64_t w = (uint64_t)(((unsigned __int128)z * n_words) >> 64);
        if (WHAT == 0) atomicOr(buf + w, 1u << (z & 31));
        else acc += __ldg(buf + (w & ~7ull));
    }
    if (WHAT == 1 && acc == 0x123456789abcdefull) *sink = acc;  // keeps the loads alive
}

// BitStorage::update_from (bitstorage.cc:103-137): dst |= src
__global__ void __launch_bounds__(256) k_or_tables(uint32_t* __restrict__ dst, const uint32_t* __restrict__ src, uint64_t n_words) {
    for (uint64_t w = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; w < n_words; w += (uint64_t)gridDim.x * blockDim.x)
        dst[w] |= src[w];
}

}  // namespace gt
