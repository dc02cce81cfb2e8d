// goetia_b200/csrc/bucket.cuh -- the write-combining insert path (K1b + K2).
//
// Why: a blind insert is n_tables random read-modify-writes per k-mer.  Sent straight to
// tables that are much larger than L2 they run at the DRAM random-sector rate (measured on
// B200: 20 G RED.OR/s over a 4 GB footprint) while the same RED.OR on an L2-resident
// footprint (<= 64 MB) runs at 193 G/s (scripts/microbench.cu, profiles/).  So updates are
// not applied where they are produced.  Every table is cut into SLICES of 2^shift slots
// (32 MB of table by default); the hashing kernel appends each update -- a 4-byte
// slice-local slot offset -- to the slice's BUCKET in HBM, and the apply kernel walks the
// buckets slice by slice, so that all RED/CAS traffic of a slice hits one L2-resident window
// of the table and each table sector travels to and from DRAM once per flush instead of
// once per update.  OR and saturating add commute, so the final table bytes are exactly
// those of the reference's serial loop (sequential-equivalence rule).
//
// DRAM bytes per k-mer (4 tables): 16 B bucket write + 16 B bucket read + the table streamed
// once per flush, against 256 B algorithmic (4 x (32 B sector read + 32 B write-back)).
//
//   K1b k_bucket      : stage packed tile -> roll both cyclic hashes -> n_tables reductions ->
//                       append to per-slice staging rows in shared memory -> one global
//                       cursor reservation per (sub-step, slice) -> 16 B run copies.
//   K2  k_apply       : CTA = one 4096-entry chunk of one bucket, buckets in blockIdx (= slice) order;
//                       16 B streaming loads of entries, RED.OR / CAS into the slice.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <type_traits>

#include "kernels.cuh"

namespace gt {

constexpr int BK_MAX_BUCKETS = 1024;  // all tables together (3 x 4 KB of shared memory)
constexpr int AP_THREADS = 256;
constexpr int AP_PER_THREAD = 32;
constexpr int AP_CHUNK = AP_THREADS * AP_PER_THREAD;  // entries per apply CTA

// What k_bucket needs: where each bucket's entries go.  With one GPU every bucket lives in the
// storage's own pending store; in a sharded storage (one process per GPU, table t cut into
// per-rank slot ranges) the buckets of slices owned by a peer point into that peer's outbox
// region (or straight into peer memory), and own_lo/own_hi give this rank's slot range.
struct ProducePlan {
    int n_tables;
    int shift;                        // log2(slots per slice)
    int n_buckets;                    // all tables, all owners
    uint32_t first[MAX_TABLES + 1];   // first bucket id of table t
    uint64_t own_lo[MAX_TABLES];      // slots [own_lo, own_hi) of table t are held by this rank
    uint64_t own_hi[MAX_TABLES];
    // cheap exact h % size for tables of >= 2^28 slots (see bin_of): dsh = size << rs >= 2^32,
    // m32 = floor((2^64-1) / dsh) < 2^32; rs < 0 selects the generic 64-bit fastmod
    uint64_t dsh[MAX_TABLES];
    uint64_t negd[MAX_TABLES];        // 2^64 - dsh: h - q * d is one multiply-add on h + q * negd
    uint32_t m32[MAX_TABLES];
    int32_t rs[MAX_TABLES];
    uint32_t slot_mask;               // 2^shift - 1
    uint32_t stage_cap;               // shared-memory staging entries per bucket and sub-step (multiple of 4)
    uint32_t piece;                   // entries a CTA takes from a bucket's cursor at a time (0: exact-size requests)
    uint32_t* const* bptr;            // [n_buckets] entry array of each bucket
    const uint32_t* bcap;             // [n_buckets] capacity in entries (multiple of 4)
    uint32_t* bfill;                  // [n_buckets] cursor; can exceed bcap (excess handled at once)
    unsigned long long* n_direct;     // updates that overflowed a bucket and were applied directly
    unsigned long long* n_dropped;    // overflowed updates of foreign slots that found the overflow list full too (an error)
    // Sharded storages, peer transport: an update that overflows the bucket of a slice held by rank q is
    // posted as a full (table, slot) record to this rank's overflow list in q's inbox (rare path).
    // An update that overflows the bucket of a slice held HERE is not applied by k_bucket itself (the apply of the
    // other store may be merging windows into the tables with plain read-modify-writes, see K2w); it is parked
    // in the store's spill list and applied after that store's windows.
    unsigned long long* own_spill;    // [own_spill_cap] (table, slot) records
    uint32_t* own_spill_fill;
    uint32_t own_spill_cap;
    const uint8_t* bowner;            // [n_buckets] owner rank of each bucket (NULL: no overflow lists)
    unsigned long long* const* ovf_ptr;  // [world] this rank's list in the inbox of rank q
    uint32_t* ovf_fill;               // [world] local cursors (travel with the bucket fill counts)
    uint32_t ovf_cap;                 // records per list
};

// overflow record: table in the top 5 bits (MAX_TABLES == 32), global slot below (table sizes <= 2^59 here)
__device__ __forceinline__ unsigned long long ovf_pack(int t, uint64_t bin) { return ((unsigned long long)t << 59) | bin; }

// an update this rank cannot apply itself and cannot bucket: post it to the owner's overflow list
__device__ __forceinline__ void post_foreign(const ProducePlan& bp, int t, uint32_t b, uint64_t bin, unsigned long long& dropped) {
    if (bp.bowner) {
        const int q = bp.bowner[b];
        const uint32_t k = atomicAdd(bp.ovf_fill + q, 1u);
        if (k < bp.ovf_cap) {
            bp.ovf_ptr[q][k] = ovf_pack(t, bin);
            return;
        }
    }
    ++dropped;
}

// What k_apply walks: one item per (slice, source rank) in slice order.
struct ApplyItem {
    const uint32_t* src;   // entries
    const uint32_t* fill;  // device location of the entry count (a cursor, or a received count)
    uint64_t slot0;        // first slot of the slice in table coordinates
    uint32_t cap;
    uint32_t table;
};

constexpr uint32_t BK_PAD = 0xFFFFFFFFu;  // filler entry (runs are padded to 16 B); never a valid offset (shift <= 31)

// h % d, bit-exact, for d with dsh = d << rs in [2^32, 2^63]: the reciprocal m32 = floor((2^64-1)/dsh)
// then fits 32 bits, so mulhi64(h, m32) is two 32x32 products; it is floor(h/dsh) or one less
// (same argument as fastmod_u64), one conditional subtract gives h % dsh, and h % d follows by
// rs (<= 4) more conditional subtracts of d << j.
__device__ __forceinline__ uint64_t bin_of(uint64_t h, uint64_t d, uint64_t magic, uint64_t dsh, uint32_t m32, int rs) {
    if (rs < 0) return fastmod_u64(h, d, magic);
    const uint32_t hl = (uint32_t)h, hh = (uint32_t)(h >> 32);
    const uint64_t u = (uint64_t)hh * m32 + __umulhi(hl, m32);  // < 2^64
    const uint32_t q = (uint32_t)(u >> 32);
    const uint64_t qd = (uint64_t)q * (uint32_t)dsh + ((uint64_t)(q * (uint32_t)(dsh >> 32)) << 32);
    uint64_t r = h - qd;
    if (r >= dsh) r -= dsh;
    for (int j = rs - 1; j >= 0; --j) {  // rs is uniform (per table); 0 iterations for tables of >= 2^32 slots
        const uint64_t dj = d << j;
        if (r >= dj) r -= dj;
    }
    return r;
}

// ------------------------------------------------------------------------------------------
// K1b.  One CTA = 8192 consecutive k-mer start positions of the flat packed stream (as k_walk),
// processed in 4 sub-steps of 8 positions per thread.  Every update is appended to its
// bucket's staging row in shared memory (one returning shared atomic per update); after each
// sub-step the rows are copied out: one global cursor reservation per (sub-step, bucket),
// runs padded to 16 B and written with 16 B stores.  A row that is full (skewed input) spills
// the update straight to the bucket / the table, so nothing is ever lost.
// ------------------------------------------------------------------------------------------
constexpr int BK_SUB = 8;  // positions per thread and sub-step
constexpr int BK_OWN = BK_MAX_BUCKETS / TILE_THREADS;  // buckets a lane may own (4)
constexpr uint32_t BK_NONE = 0xFFFFFFFFu;

// an update of a slot held here whose bucket is full is parked in the store's spill list (ProducePlan::own_spill):
// k_bucket never writes a table itself, because the apply of the OTHER store may be merging windows into it
// (a plain read-modify-write, see K2w) at the same time
template <int KIND>
__device__ __forceinline__ void apply_own_overflow(const TableSet& ts, const ProducePlan& bp, int t, uint64_t bin,
                                                   unsigned long long& direct, unsigned long long& dropped) {
    const uint32_t k = atomicAdd(bp.own_spill_fill, 1u);
    if (k < bp.own_spill_cap) {
        bp.own_spill[k] = ovf_pack(t, bin);
        ++direct;
    } else {
        ++dropped;
    }
}

template <int KIND>
__device__ __forceinline__ void bucket_spill(const TableSet& ts, const ProducePlan& bp, int t, uint32_t b, uint32_t off,
                                             unsigned long long& direct, unsigned long long& dropped) {
    // slow path: a 16 B group holding this one entry, or -- bucket full -- the overflow route.  The cursor is only
    // bumped while it is below the capacity, so that skewed input (every update spilling) cannot wrap it.
    const uint32_t cap = __ldg(bp.bcap + b);
    const uint32_t g = __ldcg(bp.bfill + b) < cap ? atomicAdd(bp.bfill + b, 4u) : cap;
    if (g < cap) {
        *reinterpret_cast<uint4*>(bp.bptr[b] + g) = make_uint4(off, BK_PAD, BK_PAD, BK_PAD);
        return;
    }
    const uint64_t bin = ((uint64_t)(b - bp.first[t]) << bp.shift) + off;
    if (bin >= bp.own_lo[t] && bin < bp.own_hi[t]) {
        apply_own_overflow<KIND>(ts, bp, t, bin, direct, dropped);
    } else {
        post_foreign(bp, t, b, bin, dropped);
    }
}

// Bulk (TMA) copy of a staging row to its bucket: shared::cta -> global, 16-byte granules, tracked by the
// issuing thread's bulk async-group.  The destination may be peer memory (NVLink) in a sharded storage.
__device__ __forceinline__ void bulk_row_out(void* gdst, const void* ssrc, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                 :: "l"(gdst), "r"((uint32_t)__cvta_generic_to_shared(ssrc)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
// generic-proxy writes to shared memory (the appends) must be made visible to the async proxy that reads the rows
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// BIG: every table has >= 2^32 slots (rs == 0), so the reducer is the straight 32-bit-reciprocal one with no
// run-time case analysis (the common case for tables larger than L2: C3's 8e9-bit tables).
template <bool BIG>
__device__ __forceinline__ uint64_t bin_of_t(uint64_t h, const TableSet& ts, const ProducePlan& bp, int t) {
    if constexpr (BIG) {
        const uint64_t d = bp.dsh[t];  // == ts.size[t]
        const uint64_t nd = bp.negd[t];
        const uint32_t m = bp.m32[t];
        const uint32_t q = (uint32_t)(((uint64_t)(uint32_t)(h >> 32) * m + __umulhi((uint32_t)h, m)) >> 32);
        // h - q * d = h + q * (2^64 - d) modulo 2^64: one wide multiply-add and one 32-bit one
        const uint64_t lo = (uint64_t)q * (uint32_t)nd + h;
        const uint64_t r = lo + ((uint64_t)(q * (uint32_t)(nd >> 32)) << 32);  // < 2 d
        const uint64_t r2 = r - d;
        return (int64_t)r2 < 0 ? r : r2;  // d <= 2^63 and r < 2 d: r - d is "negative" exactly when r < d
    } else {
        return bin_of(h, ts.size[t], ts.magic[t], bp.dsh[t], bp.m32[t], bp.rs[t]);
    }
}

template <int KIND, bool CAN, int NT, bool BIG>
__global__ void __launch_bounds__(TILE_THREADS, 4)  // <= 64 registers: three CTAs per SM and room for apply CTAs beside them
k_bucket(const __grid_constant__ WalkArgs a, const __grid_constant__ TableSet ts, const __grid_constant__ ProducePlan bp) {
    extern __shared__ __align__(16) uint64_t smem[];
    const int K = a.K;
    const int halo_words = ((K - 1 + 31) >> 5) + 1;
    const int tile_words = TILE_THREADS + halo_words;
    const int nb = bp.n_buckets;
    const int nt = NT > 0 ? NT : bp.n_tables;
    const uint32_t C = bp.stage_cap;                                 // staging entries per bucket (multiple of 4)
    ulonglong2* tab = reinterpret_cast<ulonglong2*>(smem);           // 8 x 16 B: per-base roll constants
    ulonglong2* tab2 = tab + 8;                                      // 16 x 16 B: per-base-pair seed constants {fw, rc}
    uint64_t* sw = smem + 48;                                        // packed tile + halo
    uint32_t* stage = reinterpret_cast<uint32_t*>(sw + tile_words + (tile_words & 1));  // [nb][C], 16-byte aligned rows
    // cnt[2][nb]: entries staged in every bucket's row, for this / the next sub-step
    uint32_t* cnt = stage + (size_t)nb * C;
    // capacity and entry array of every bucket, copied here once: the copy-out reads them for every row and sub-step,
    // and a global load there is latency nobody hides (ncu: the capacity test was the kernel's top stall line)
    uint32_t* s_cap = cnt + 2 * nb;
    uint32_t** s_dst = reinterpret_cast<uint32_t**>(s_cap + nb + (nb & 1));
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid < 4) {
        tab[tid] = make_ulonglong2(lemire_T(tid), rotl64(lemire_T(3 - tid), (unsigned)K));
        tab[4 + tid] = make_ulonglong2(rotl64(lemire_T(tid), (unsigned)K), lemire_T(3 - tid));
    }
    if (tid < 16) {
        // two bases c0 (first), c1: forward eats c0 then c1; the reverse strand sees comp(c1) first
        const int c0 = tid & 3, c1 = tid >> 2;
        tab2[tid] = make_ulonglong2(rotl1(lemire_T(c0)) ^ lemire_T(c1), lemire_T(3 - c0) ^ rotl1(lemire_T(3 - c1)));
    }
    for (int b = tid; b < 2 * nb; b += TILE_THREADS) cnt[b] = 0;
    for (int b = tid; b < nb; b += TILE_THREADS) {
        s_cap[b] = __ldg(bp.bcap + b);
        s_dst[b] = bp.bptr[b];
    }
    const uint64_t n_tiles = (a.n_bases + TILE_POS - 1) / TILE_POS;
    unsigned long long direct = 0, dropped = 0;
    int phase = 0;  // which half of cnt[] the current sub-step appends to
    const uint32_t R = bp.piece, C4 = (C + 3u) & ~3u;
    uint32_t c_pos[BK_OWN], c_left[BK_OWN], c_nxt[BK_OWN];  // this lane's buckets: write position, room, next piece
#pragma unroll
    for (int rr = 0; rr < BK_OWN; ++rr) { c_pos[rr] = 0; c_left[rr] = 0; c_nxt[rr] = BK_NONE; }

    for (uint64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        __syncthreads();  // previous tile's words consumed; tab / cnt visible
        const uint64_t w0 = tile * TILE_THREADS;
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
        const bool live = p0 < a.n_bases;
        uint64_t fw = 0, rc = 0, wo = 0, win = 0, r = 0, rend = 0;
        bool rok = false;
        if (live) {
            // seed both hashes from the window at p0 (word-aligned), two bases per step:
            //   fw = sum_j rotl(T[c_j], K-1-j)  (K "eat" steps, cyclichash.h:105-108)
            //   rc = sum_j rotl(T[comp c_j], j) (rollinghashshifter.hh:183-197)
            const int pairs = K >> 1;
            if (K & 1) {
                const int j = K - 1;
                const int c = (int)((sw[tid + (j >> 5)] >> (2 * (j & 31))) & 3);
                if (CAN) rc = tab[4 + c].y;  // T[comp c]; the loop below rotates it by K-1
            }
            for (int g = 0; g < pairs; ++g) {
                const int nf = (int)((sw[tid + (g >> 4)] >> (4 * (g & 15))) & 15);
                fw = rotl64(fw, 2) ^ tab2[nf].x;
                if (CAN) {
                    const int gr = pairs - 1 - g;
                    const int nr = (int)((sw[tid + (gr >> 4)] >> (4 * (gr & 15))) & 15);
                    rc = rotl64(rc, 2) ^ tab2[nr].y;
                }
            }
            if (K & 1) {
                const int j = K - 1;
                const int c = (int)((sw[tid + (j >> 5)] >> (2 * (j & 31))) & 3);
                fw = rotl1(fw) ^ tab[c].x;
            }
            wo = sw[tid];
            {
                int a0 = tid + ((K - 1) >> 5);
                unsigned sh = 2u * (unsigned)((K - 1) & 31);
                win = sh ? (sw[a0] >> sh) | (sw[a0 + 1] << (64u - sh)) : sw[a0];
            }
            r = __ldg(a.coarse + (p0 >> COARSE_SHIFT));
            rend = __ldg(a.offsets + r + 1) - a.base0;
            int steps = 0;
            while (rend <= p0) {
                if (++steps > 8) {
                    r = find_read(a.offsets, a.n_reads, p0 + a.base0);
                    rend = __ldg(a.offsets + r + 1) - a.base0;
                    break;
                }
                ++r;
                rend = __ldg(a.offsets + r + 1) - a.base0;
            }
            rok = !(__ldg(a.flags + r) & READ_INVALID);
        }
        // vmask bit i: position p0 + i starts a k-mer that lies inside one valid read (dbg.hh:296-305 walks every read
        // on its own; here the reads are back to back in one stream).  Worked out once per thread from the read
        // boundaries, so that the 32 steps below test a bit instead of comparing 64-bit positions.
        uint32_t vmask = 0;
        if (live) {
            const uint32_t n_here = (uint32_t)min((uint64_t)POS_PER_THREAD, a.n_bases - p0);
            uint32_t i = 0;
            while (true) {
                const uint64_t left = rend - (p0 + i);                 // positions of read r from p0 + i on (>= 0)
                const uint32_t seg_end = left < (uint64_t)(n_here - i) ? i + (uint32_t)left : n_here;  // read r covers [i, seg_end)
                if (rok && left >= (uint64_t)K) {
                    const uint64_t starts = left - (uint64_t)K + 1;   // valid k-mer starts from i on
                    const uint32_t hi = starts < (uint64_t)(n_here - i) ? i + (uint32_t)starts : n_here;  // bits [i, hi)
                    if (hi > i) vmask |= (hi - i >= 32u ? 0xFFFFFFFFu : ((1u << (hi - i)) - 1u)) << i;
                }
                i = seg_end;
                if (i >= n_here) break;
                ++r;
                rend = __ldg(a.offsets + r + 1) - a.base0;
                rok = !(__ldg(a.flags + r) & READ_INVALID);
            }
        }

        for (int sub = 0; sub < POS_PER_THREAD / BK_SUB; ++sub) {
            uint32_t* cn = cnt + phase * nb;
            // ---- append: roll, reduce, stage ------------------------------------------------------
            if (vmask >> (sub * BK_SUB)) {  // a k-mer starts here or later in this thread's run (else the hashes are not needed any more)
                const uint32_t slot_mask = bp.slot_mask;
#pragma unroll 4
                for (int ii = 0; ii < BK_SUB; ++ii) {
                    const int i = sub * BK_SUB + ii;
                    if (i) {
                        int out = (int)((wo >> (2 * (i - 1))) & 3);
                        int in = (int)((win >> (2 * i)) & 3);
                        ulonglong2 ti = tab[in], to = tab[4 + out];
                        fw = rotl1(fw) ^ to.x ^ ti.x;
                        if (CAN) rc = rotr1(rc ^ ti.y ^ to.y);
                    }
                    if ((vmask >> i) & 1u) {
                        const uint64_t h = CAN ? (fw < rc ? fw : rc) : fw;  // Canonical::value(), canonical.hh:124-126
                        uint32_t full = 0;  // tables whose row was full (skewed input): handled behind the loop, off the hot path
#pragma unroll
                        for (int t = 0; t < nt; ++t) {
                            const uint64_t bin = bin_of_t<BIG>(h, ts, bp, t);
                            const uint32_t b = bp.first[t] + (uint32_t)(bin >> bp.shift);
                            const uint32_t idx = atomicAdd(&cn[b], 1u);
                            if (idx < C) stage[b * C + idx] = (uint32_t)bin & slot_mask;
                            else full |= 1u << t;
                        }
                        if (full) {
                            for (int t = 0; t < nt; ++t) {
                                if (!((full >> t) & 1u)) continue;
                                const uint64_t bin = bin_of_t<BIG>(h, ts, bp, t);
                                bucket_spill<KIND>(ts, bp, t, bp.first[t] + (uint32_t)(bin >> bp.shift), (uint32_t)bin & slot_mask, direct, dropped);
                            }
                        }
                    }
                }
            }
            fence_async_smem();
            __syncthreads();
            // ---- copy the rows out.  Lane l of warp w owns buckets w + 8*l + 256*rr (rr < 4) and keeps their write
            // position in registers: room in the global bucket is taken R entries at a time and the next piece is
            // requested one flush ahead, so the latency of the global cursor atomic is hidden behind the next append
            // phase (R == 0: exact-size requests).  The owner pads its row to a multiple of 16 bytes and sends it
            // with ONE bulk copy (two when the run straddles a piece boundary); nobody else touches the row.
            uint32_t* cz = cnt + (phase ^ 1) * nb;
            bool issued = false;
#pragma unroll
            for (int rr = 0; rr < BK_OWN; ++rr) {
                const int mine = warp + rr * TILE_THREADS + 8 * lane;  // TILE_THREADS / 32 == 8 warps
                if (mine >= nb) break;
                const uint32_t row0 = (uint32_t)mine * C;
                const uint32_t n = min(cn[mine], C);
                cz[mine] = 0;  // the other half is idle now: rewind it for the next sub-step
                if (n == 0) continue;
                const uint32_t n4 = (n + 3u) & ~3u;
                uint32_t* row = stage + row0;
                for (uint32_t e = n; e < n4; ++e) row[e] = BK_PAD;
                uint32_t d0 = 0, l0 = n4, d1 = 0, l1 = 0;
                if (R == 0) {
                    d0 = atomicAdd(bp.bfill + mine, n4);
                } else {
                    l0 = min(c_left[rr], n4);
                    d0 = c_pos[rr];
                    c_pos[rr] += l0;
                    c_left[rr] -= l0;
                    l1 = n4 - l0;
                    if (l1) {  // continue in the next piece (R >= the longest run)
                        if (c_nxt[rr] == BK_NONE) c_nxt[rr] = atomicAdd(bp.bfill + mine, R);
                        d1 = c_nxt[rr];
                        c_nxt[rr] = BK_NONE;
                        c_pos[rr] = d1 + l1;
                        c_left[rr] = R - l1;
                    }
                    if (c_left[rr] < C4 && c_nxt[rr] == BK_NONE) c_nxt[rr] = atomicAdd(bp.bfill + mine, R);
                }
                const uint32_t cap = s_cap[mine];
                uint32_t* dst = s_dst[mine];
                if ((l0 == 0 || d0 + l0 <= cap) && (l1 == 0 || d1 + l1 <= cap)) {
                    fence_async_smem();  // the pad entries above
                    if (l0) bulk_row_out(dst + d0, row, l0 * 4u);
                    if (l1) bulk_row_out(dst + d1, row + l0, l1 * 4u);
                    issued = true;
                } else {  // bucket (nearly) full -- skewed input: group by group, the excess takes the overflow route
                    int t = 0;
                    while ((uint32_t)mine >= bp.first[t + 1]) ++t;
                    for (uint32_t e = 0; e < n4; e += 4) {
                        const uint4 v = *reinterpret_cast<const uint4*>(row + e);
                        const uint32_t at = e < l0 ? d0 + e : d1 + (e - l0);
                        if (at < cap) {
                            *reinterpret_cast<uint4*>(dst + at) = v;
                            continue;
                        }
                        const uint32_t offs[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            if (offs[k] == BK_PAD) continue;
                            const uint64_t bin = ((uint64_t)((uint32_t)mine - bp.first[t]) << bp.shift) + offs[k];
                            if (bin >= bp.own_lo[t] && bin < bp.own_hi[t]) apply_own_overflow<KIND>(ts, bp, t, bin, direct, dropped);
                            else post_foreign(bp, t, (uint32_t)mine, bin, dropped);
                        }
                    }
                }
            }
            if (issued) {
                bulk_commit();
                bulk_wait_read();  // the rows are appended to again after the barrier below
            }
            __syncthreads();
            phase ^= 1;
        }
    }
    // what is left of the pieces this CTA took is filled with pad entries (the apply skips them)
    if (R) {
#pragma unroll
        for (int rr = 0; rr < BK_OWN; ++rr) {
            const int mine = warp + rr * TILE_THREADS + 8 * lane;
            if (mine >= nb) continue;
            const uint32_t cap = s_cap[mine];
            uint32_t* dst = s_dst[mine];
            const uint4 pad = make_uint4(BK_PAD, BK_PAD, BK_PAD, BK_PAD);
            for (uint32_t e = c_pos[rr]; e < c_pos[rr] + c_left[rr] && e < cap; e += 4) *reinterpret_cast<uint4*>(dst + e) = pad;
            if (c_nxt[rr] != BK_NONE)
                for (uint32_t e = c_nxt[rr]; e < c_nxt[rr] + R && e < cap; e += 4) *reinterpret_cast<uint4*>(dst + e) = pad;
        }
    }
    if (direct) atomicAdd(bp.n_direct, direct);
    if (dropped) atomicAdd(bp.n_dropped, dropped);
}

// ------------------------------------------------------------------------------------------
// K2
// ------------------------------------------------------------------------------------------
// chunk_start[n_buckets+1]: first CTA of each bucket, host-built from the capacities, so the
// grid needs no device-side planning; CTAs past a bucket's fill exit at once.
// The plain apply (global atomics): used when little is pending relative to the tables (store_apply_async).
template <int KIND>
__global__ void __launch_bounds__(AP_THREADS)
k_apply(const __grid_constant__ TableSet ts, const ApplyItem* __restrict__ items, const uint32_t* __restrict__ chunk_start,
        int n_items) {
    __shared__ uint32_t s_b;
    if (threadIdx.x == 0) {
        // item of this CTA: largest b with chunk_start[b] <= blockIdx.x
        uint32_t lo = 0, hi = (uint32_t)n_items;
        while (hi - lo > 1) {
            uint32_t mid = (lo + hi) >> 1;
            if (__ldg(chunk_start + mid) <= blockIdx.x) lo = mid; else hi = mid;
        }
        s_b = lo;
    }
    __syncthreads();
    const uint32_t b = s_b;
    const ApplyItem it = items[b];
    const uint32_t fill = min(__ldcg(it.fill), it.cap);
    const uint32_t e0 = (blockIdx.x - __ldg(chunk_start + b)) * (uint32_t)AP_CHUNK;
    if (e0 >= fill) return;
    uint32_t* tbl = ts.ptr[it.table] + (KIND == 0 ? (it.slot0 >> 5) : KIND == 1 ? (it.slot0 >> 2) : (it.slot0 >> 3));
    const uint32_t* src = it.src;
    const uint32_t n = min((uint32_t)AP_CHUNK, fill - e0);
    // 16 B streaming loads where the chunk is full and aligned; scalar tail otherwise
    if (n == AP_CHUNK && ((reinterpret_cast<uintptr_t>(src + e0) & 15) == 0)) {
        const uint4* v = reinterpret_cast<const uint4*>(src + e0);
#pragma unroll
        for (int k = 0; k < AP_PER_THREAD / 4; ++k) {
            uint4 x = __ldcs(v + k * AP_THREADS + threadIdx.x);
            if (x.x != BK_PAD) slot_insert<KIND, false>(tbl, x.x);
            if (x.y != BK_PAD) slot_insert<KIND, false>(tbl, x.y);
            if (x.z != BK_PAD) slot_insert<KIND, false>(tbl, x.z);
            if (x.w != BK_PAD) slot_insert<KIND, false>(tbl, x.w);
        }
    } else {
        for (uint32_t e = threadIdx.x; e < n; e += AP_THREADS) {
            uint32_t x = __ldcs(src + e0 + e);
            if (x != BK_PAD) slot_insert<KIND, false>(tbl, x);
        }
    }
}

template <int KIND>
__device__ __forceinline__ void counter_addr(uint32_t off, uint32_t& word, uint32_t& sh) {
    if constexpr (KIND == 1) {
        word = off >> 2;
        sh = (off & 3u) * 8u;
    } else {
        word = off >> 3;
        sh = ((off >> 1) & 3u) * 8u + ((off & 1u) ? 0u : 4u);
    }
}

// ------------------------------------------------------------------------------------------
// K2w: the apply through shared-memory WINDOWS (all three storages).
//
// k_apply above sends one L2 atomic per update.  That is a hard ceiling (measured on B200: ~200 G RED/s chip-wide,
// and every RED of a warp is its own L1->L2 request, so the SM's load/store path is busy 32 cycles per warp
// instruction -- which is also what slows k_bucket down when the two kernels share an SM), and for the counting
// storages a saturating increment needs a CAS (load + compare-and-swap, ~26 G/s).  Shared-memory atomics are an
// order of magnitude cheaper, so a slice is cut once more into WINDOWS of 2^wshift slots that fit shared memory:
//
//   k_rebucket2  partitions the entries of a slice (all sources) by window through per-window staging rows in shared
//                memory (below): runs of window-local offsets appended to the window's SUB-BUCKET;
//   k_apply_win  one CTA per window: zero the window image in shared memory, stream the sub-bucket with 16 B loads
//                and OR / saturating-add into shared memory, then merge the image into the table ONCE with 16 B
//                loads and stores (table |= image; per-byte / per-nibble saturating add for the counting storages).
//
// The image of a counting window holds the hits of this apply per counter, saturating at the counter's maximum, so
// table' = min(max, table + min(max, hits)) = min(max, table + hits): exactly the reference's one-at-a-time
// saturating increments (bytestorage.cc:60-113, nibblestorage.cc:60-100) in any order.  Slices that are not larger
// than a window skip k_rebucket2 (n_win == 1: the level-1 entries already are window-local).
// The merge is a plain read-modify-write, so nothing else may write the tables while an apply runs: k_bucket
// parks the updates that overflow a bucket of its own slots in the store's spill list (ProducePlan::own_spill,
// applied after the store's windows), a sub-bucket that overflows applies the excess with global atomics from
// k_rebucket2 (which runs before any window of that apply), and the direct writers (k_walk, k_insert_hashes) are
// ordered against the apply stream (direct_begin / direct_end).
// ------------------------------------------------------------------------------------------
constexpr int AW_THREADS = 512;
constexpr int WIN_MAX_PER_SLICE = 1024;

struct SliceWin {
    uint32_t* sub;         // two-level: n_win sub-buckets of cap2 entries each (window-local offsets)
    uint32_t* sub_fill;    // two-level: n_win cursors
    uint64_t slot0;        // first slot of the slice in table coordinates
    uint32_t cap2;
    uint32_t n_win;        // windows of this slice that exist (the last slice of a table may be short)
    uint32_t words;        // 32-bit table words of the slice that exist (multiple of 4)
    uint32_t table;
    uint32_t win0;         // id of the slice's first window among all windows applied here
};

template <int KIND>
__device__ __forceinline__ uint32_t* slice_words_ptr(const TableSet& ts, uint32_t table, uint64_t slot0) {
    return ts.ptr[table] + (KIND == 0 ? (slot0 >> 5) : KIND == 1 ? (slot0 >> 2) : (slot0 >> 3));
}


// ------------------------------------------------------------------------------------------
// k_rebucket2: the level-2 partition by STAGING ROWS (as k_bucket does at level 1).
//
// (Its predecessor, a counting sort per 8192-entry chunk, was bound by shared-memory wavefronts and barriers -- ncu,
// profiles/r2_ncu_full_v1.txt: 16 wavefronts per warp and entry for histogram atomic, rank lookup, scatter, gather and
// base lookup, six barriers per chunk; 10.4 ms per 2.9 G entries against 3.7 ms for its 24 GB of DRAM traffic.)  A persistent
// CTA keeps one row of C entries per window in shared memory; an entry costs ONE returning shared atomic (the
// row cursor) and ONE shared store.  After every chunk (16 entries per thread) each row is sent to its sub-bucket by its owner
// thread: one global cursor reservation and ONE bulk (TMA) copy shared -> global of the row's multiple-of-4
// prefix; the 0..3 entries behind it move to the front of the row (one 16-byte load and store) and go out with
// the next chunk, so sub-buckets hold no pad entries except where a CTA leaves a slice (rows are then padded to
// 16 bytes and emptied).  The chunks of an item are dealt round-robin to the CTAs, item after item: all CTAs work
// on the same slice at about the same time, so the slice's n_win output streams stay together in DRAM; the next
// chunk is loaded into registers while the current one is staged.  A row that is full (skewed input) and a
// sub-bucket that is full send the update straight to the table.
// shared memory: cnt[nw_max + 1] (cnt[nw_max]: where pad entries count themselves) | stage[nw_max][C]
// C/4 is odd: consecutive rows then start 4 banks apart modulo 32 (a multiple of 8 would put all rows on 2 or 4
// bank groups: rows fill at the same pace, so the append positions of a warp would collide).
// ------------------------------------------------------------------------------------------
constexpr int RB2_PER_THREAD = 16;  // entries a thread stages per chunk: a chunk is T * 16 entries (T threads per CTA)

// this CTA's walk over the chunks: item after item, chunk j of an item goes to CTA (chunks of earlier items + j) % grid
struct ChunkWalk {
    const ApplyItem* items;
    int n_items, b, ips, slice;
    uint32_t j, nch, rot, fill, chunk;
    const uint32_t* src;
    __device__ __forceinline__ void open_item() {  // b is valid: read its fill, find this CTA's first chunk in it
        const ApplyItem it = items[b];
        fill = min(__ldcg(it.fill), it.cap);
        src = it.src;
        nch = (fill + chunk - 1) / chunk;
        j = (blockIdx.x + gridDim.x - rot) % gridDim.x;
        slice = b / ips;
    }
    __device__ __forceinline__ void start(const ApplyItem* items_, int n_items_, int ips_, uint32_t chunk_) {
        chunk = chunk_; items = items_; n_items = n_items_; ips = ips_; b = 0; slice = 0; rot = 0; j = nch = fill = 0; src = nullptr;
        if (n_items > 0) { open_item(); settle(); }
    }
    __device__ __forceinline__ void settle() {  // move on to the next item until this CTA has a chunk in it
        while (b < n_items && j >= nch) {
            rot = (rot + nch) % gridDim.x;
            if (++b < n_items) open_item();
        }
    }
    __device__ __forceinline__ bool done() const { return b >= n_items; }
    __device__ __forceinline__ void next() { j += gridDim.x; settle(); }
};

template <int T>
__device__ __forceinline__ void rb2_load(const ChunkWalk& w, uint32_t (&v)[RB2_PER_THREAD], int tid) {
    const uint32_t e0 = w.j * w.chunk;
    const uint32_t n = min(w.chunk, w.fill - e0);
    const uint32_t* src = w.src + e0;
    if (n == w.chunk && ((reinterpret_cast<uintptr_t>(src) & 15) == 0)) {
#pragma unroll
        for (int k = 0; k < RB2_PER_THREAD / 4; ++k) {
            const uint4 x = __ldcs(reinterpret_cast<const uint4*>(src) + k * T + tid);
            v[4 * k] = x.x; v[4 * k + 1] = x.y; v[4 * k + 2] = x.z; v[4 * k + 3] = x.w;
        }
    } else {
#pragma unroll
        for (int k = 0; k < RB2_PER_THREAD; ++k) {
            const uint32_t e = (uint32_t)k * T + tid;
            v[k] = e < n ? __ldcs(src + e) : BK_PAD;
        }
    }
}

template <int KIND, int T, int PASSES>  // T threads; PASSES * T / 4 >= windows per slice
__global__ void __launch_bounds__(T, 1024 / T)
k_rebucket2(const __grid_constant__ TableSet ts, const ApplyItem* __restrict__ items, int n_items, int items_per_slice,
            const SliceWin* __restrict__ slices, int wshift, int nw_max, uint32_t C, uint32_t wmask, uint32_t stage_off) {
    // wmask = 2^wshift - 1 and stage_off = (nw_max + 4) & ~3 come as parameters: constant-bank operands cost no
    // registers, and the compiler otherwise re-derives them inside the predicated append (64-register budget)
    extern __shared__ __align__(16) uint32_t rb_sm[];
    uint32_t* cnt = rb_sm;                    // [nw_max + 1], padded to a multiple of 4 words
    uint32_t* stage = rb_sm + stage_off;
    const int tid = threadIdx.x;
    for (int w = tid; w < nw_max; w += T) cnt[w] = 0;
    // pads count themselves in cnt[nw_max], which starts at 2^31: their index is never < C (no store) and, read as a
    // signed number, never >= C (no overflow flag)
    if (tid == 0) cnt[nw_max] = 0x80000000u;
    __syncthreads();
    const uint32_t nwm = (uint32_t)nw_max;
    int cur_slice = -1;
    SliceWin sl;
    uint32_t* tbl = nullptr;

    // Send the rows out: 4 lanes per row, 128 rows per pass.  The group's first lane takes the row's multiple-of-4
    // prefix, reserves room in the sub-bucket (all passes' reservations are issued before any is waited for), the four
    // lanes copy it with 16-byte loads and stores, and the 0..3 entries behind it move to the front of the row.
    // (A bulk copy per row would cost more here: UBLKCP is a uniform-datapath instruction, so per-lane copies are
    // serialised by a ~10-instruction loop per lane -- 2 G of the 4.7 G warp instructions of the first version.)
    // final: the CTA leaves the slice -- rows are padded to 16 bytes and emptied.
    auto flush = [&](bool final) {
        const int q = tid & 3, lead = tid & 28;
        const uint32_t C4 = C >> 2;
        uint4* stage16 = reinterpret_cast<uint4*>(stage);
        uint32_t g_[PASSES], n4_[PASSES];
#pragma unroll
        for (int i = 0; i < PASSES; ++i) {
            const uint32_t row = (uint32_t)(tid >> 2) + i * (T / 4);
            uint32_t n4 = 0, g = 0;
            if (q == 0 && row < sl.n_win) {
                const uint32_t n = min(cnt[row], C);
                n4 = n & ~3u;
                if (final && n4 != n) {
                    uint32_t* r = stage + (size_t)row * C;
                    for (uint32_t e = n; e < n4 + 4; ++e) r[e] = BK_PAD;  // C is a multiple of 4
                    n4 += 4;
                }
                cnt[row] = final ? 0u : n - n4;
                if (n4) g = atomicAdd(sl.sub_fill + row, n4);
            }
            g_[i] = g;
            n4_[i] = n4;
        }
        __syncwarp();
#pragma unroll
        for (int i = 0; i < PASSES; ++i) {
            const uint32_t row = (uint32_t)(tid >> 2) + i * (T / 4);
            const uint32_t n4 = __shfl_sync(0xffffffffu, n4_[i], lead), g = __shfl_sync(0xffffffffu, g_[i], lead);
            if (n4) {
                uint4* r16 = stage16 + (size_t)row * C4;
                uint32_t* dst = sl.sub + (size_t)row * sl.cap2;
                if (g + n4 <= sl.cap2) {
                    uint4* d16 = reinterpret_cast<uint4*>(dst + g);
                    for (uint32_t u = q; u < (n4 >> 2); u += 4) d16[u] = r16[u];
                } else if (q == 0) {  // sub-bucket full (skewed input): what does not fit goes straight to the table
                    const uint32_t* r = stage + (size_t)row * C;
                    for (uint32_t e = 0; e < n4; ++e) {
                        const uint32_t x = r[e];
                        if (g + e < sl.cap2) dst[g + e] = x;
                        else if (x != BK_PAD) slot_insert<KIND, false>(tbl, ((uint64_t)row << wshift) + x);
                    }
                }
            }
            __syncwarp();
            if (!final && n4 && n4 < C) {  // one word per lane: the row's first granule <- the granule behind the prefix
                uint32_t* r = stage + (size_t)row * C;
                r[q] = r[n4 + q];
            }
        }
    };

    ChunkWalk cw;
    cw.start(items, n_items, items_per_slice, (uint32_t)T * RB2_PER_THREAD);
    uint32_t v[RB2_PER_THREAD], nx[RB2_PER_THREAD];
    if (!cw.done()) rb2_load<T>(cw, nx, tid);
    while (!cw.done()) {
        const int s = cw.slice;
        if (s != cur_slice) {
            if (cur_slice >= 0) {
                flush(true);
                __syncthreads();
            }
            cur_slice = s;
            sl = slices[s];
            tbl = slice_words_ptr<KIND>(ts, sl.table, sl.slot0);
        }
#pragma unroll
        for (int k = 0; k < RB2_PER_THREAD; ++k) v[k] = nx[k];
        cw.next();
        if (!cw.done()) rb2_load<T>(cw, nx, tid);  // in flight while this chunk is staged
        uint32_t ovf = 0;
#pragma unroll
        for (int k0 = 0; k0 < RB2_PER_THREAD; k0 += 4) {  // four row cursors in flight before the first store needs one
            uint32_t w[4], idx[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                w[k] = min(v[k0 + k] >> wshift, nwm);  // pads (0xFFFFFFFF) go to cnt[nw_max]
                idx[k] = atomicAdd(&cnt[w[k]], 1u);
            }
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                if (idx[k] < C) stage[w[k] * C + idx[k]] = v[k0 + k] & wmask;
                if ((int32_t)idx[k] >= (int32_t)C) ovf |= 1u << (k0 + k);
            }
        }
        if (ovf) {  // row full (skewed input): straight to the table
#pragma unroll
            for (int k = 0; k < RB2_PER_THREAD; ++k)
                if (ovf & (1u << k)) slot_insert<KIND, false>(tbl, v[k]);
        }
        fence_async_smem();
        __syncthreads();
        flush(false);
        __syncthreads();
    }
    if (cur_slice >= 0) flush(true);
}

// EXACT: the saturating increment as a compare-and-swap loop.  !EXACT (counting storages): a plain add of 1 << field --
// right as long as no field of the image receives more hits than it can hold (checked afterwards, see k_apply_win).
template <int KIND, bool EXACT>
__device__ __forceinline__ void window_update(uint32_t* __restrict__ win, uint32_t off) {
    if constexpr (KIND == 0) {
        atomicOr(win + (off >> 5), 1u << (off & 31));
    } else {
        uint32_t word, sh;
        counter_addr<KIND>(off, word, sh);
        if constexpr (!EXACT) {
            atomicAdd(win + word, 1u << sh);
        } else {
            constexpr uint32_t fmax = KIND == 1 ? 255u : 15u;
            uint32_t old = win[word];
            while (((old >> sh) & fmax) != fmax) {
                const uint32_t prev = atomicCAS(win + word, old, old + (1u << sh));
                if (prev == old) break;
                old = prev;
            }
        }
    }
}

template <int KIND>
__device__ __forceinline__ uint32_t window_merge(uint32_t t, uint32_t d) {
    if constexpr (KIND == 0) return t | d;
    else if constexpr (KIND == 1) return __vaddus4(t, d);
    else {
        const uint32_t lo = __vminu4((t & 0x0f0f0f0fu) + (d & 0x0f0f0f0fu), 0x0f0f0f0fu);
        const uint32_t hi = __vminu4(((t >> 4) & 0x0f0f0f0fu) + ((d >> 4) & 0x0f0f0f0fu), 0x0f0f0f0fu);
        return lo | (hi << 4);
    }
}

// sum of the counters held in one image word
template <int KIND>
__device__ __forceinline__ uint32_t field_sum(uint32_t w) {
    if constexpr (KIND == 1) return __dp4a(w, 0x01010101u, 0u);
    else return __dp4a(w & 0x0f0f0f0fu, 0x01010101u, __dp4a((w >> 4) & 0x0f0f0f0fu, 0x01010101u, 0u));
}

// One CTA per window.  TWO_LEVEL: the window's sub-bucket (k_rebucket2's output); otherwise the slice is the window
// and its sources are the level-1 items [slice * items_per_slice, +items_per_slice).
//
// Counting storages: the image is first built OPTIMISTICALLY with plain shared-memory adds of 1 << field (as cheap as
// the Bloom filter's OR; the CAS loop costs three times as much).  Every add raises the sum of the image's fields by
// exactly one unless it finds its field at the maximum -- then the field wraps to 0 and carries into its neighbour,
// and the sum drops by at least max - 1.  So "sum of all fields == entries applied" holds exactly when no field was
// asked to hold more than it can (a counter hit more than 255 / 15 times by ONE apply: poly-A input); the CTA checks
// it before the merge and otherwise rebuilds the image with the saturating CAS.  Either way the image holds
// min(max, hits) and the merge adds it to the table with per-field saturation.
template <int KIND, bool TWO_LEVEL>
__global__ void __launch_bounds__(AW_THREADS)
k_apply_win(const __grid_constant__ TableSet ts, const ApplyItem* __restrict__ items, int items_per_slice,
            const SliceWin* __restrict__ slices, const uint32_t* __restrict__ win_slice, int wshift) {
    extern __shared__ __align__(16) uint32_t win[];
    __shared__ uint32_t s_red[2 * (AW_THREADS / 32)];
    const uint32_t s = __ldg(win_slice + blockIdx.x);
    const SliceWin sl = slices[s];
    const uint32_t d = blockIdx.x - sl.win0;
    constexpr int spw_log2 = KIND == 0 ? 5 : KIND == 1 ? 2 : 3;
    const uint32_t wwords = 1u << (wshift - spw_log2);  // words of a full window
    const uint32_t w_lo = d * wwords;
    if (w_lo >= sl.words) return;
    const uint32_t nwords = min(wwords, sl.words - w_lo);  // multiple of 4
    const int n_src = TWO_LEVEL ? 1 : items_per_slice;
    uint32_t total = 0;
    if constexpr (TWO_LEVEL) {
        total = min(__ldcg(sl.sub_fill + d), sl.cap2);
    } else {
        for (int q = 0; q < n_src; ++q) {
            const ApplyItem it = items[s * (uint32_t)items_per_slice + q];
            total += min(__ldcg(it.fill), it.cap);
        }
    }
    if (total == 0) return;  // nothing for this window: the table is not touched
    const int tid = threadIdx.x;
    uint4* win4 = reinterpret_cast<uint4*>(win);

    // zero the image, then stream the entries into it; returns the entries this thread applied (pads excluded)
    auto build = [&](auto exact_tag) -> uint32_t {
        constexpr bool EXACT = decltype(exact_tag)::value;
        uint32_t applied = 0;
        for (uint32_t i = tid; i < nwords / 4; i += AW_THREADS) win4[i] = make_uint4(0, 0, 0, 0);
        __syncthreads();
        auto one = [&](uint32_t x) {
            if (x != BK_PAD) {
                window_update<KIND, EXACT>(win, x);
                if constexpr (KIND != 0 && !EXACT) ++applied;
            }
        };
        auto four = [&](const uint4& v) { one(v.x); one(v.y); one(v.z); one(v.w); };
        for (int q = 0; q < n_src; ++q) {
            const uint32_t* src;
            uint32_t n;
            if constexpr (TWO_LEVEL) {
                src = sl.sub + (size_t)d * sl.cap2;
                n = total;
            } else {
                const ApplyItem it = items[s * (uint32_t)items_per_slice + q];
                src = it.src;
                n = min(__ldcg(it.fill), it.cap);
            }
            // 16 B streaming loads over the aligned body, scalar head / tail
            const uint32_t head = min(n, (uint32_t)((16 - (reinterpret_cast<uintptr_t>(src) & 15)) & 15) / 4);
            if ((uint32_t)tid < head) one(__ldcs(src + tid));
            const uint32_t body = (n - head) / 4;
            const uint4* v = reinterpret_cast<const uint4*>(src + head);
            uint32_t i = tid;
            for (; i + 3 * AW_THREADS < body; i += 4 * AW_THREADS) {
                const uint4 a = __ldcs(v + i), b = __ldcs(v + i + AW_THREADS), c = __ldcs(v + i + 2 * AW_THREADS), e = __ldcs(v + i + 3 * AW_THREADS);
                four(a); four(b); four(c); four(e);
            }
            for (; i < body; i += AW_THREADS) four(__ldcs(v + i));
            const uint32_t tail0 = head + body * 4;
            if (tail0 + tid < n) one(__ldcs(src + tail0 + tid));
        }
        __syncthreads();
        return applied;
    };

    if constexpr (KIND == 0) {
        build(std::true_type{});
    } else {
        uint32_t applied = build(std::false_type{});
        uint32_t sum = 0;
        for (uint32_t i = tid; i < nwords / 4; i += AW_THREADS) {
            const uint4 w = win4[i];
            sum += field_sum<KIND>(w.x) + field_sum<KIND>(w.y) + field_sum<KIND>(w.z) + field_sum<KIND>(w.w);
        }
        for (int o = 16; o; o >>= 1) {
            applied += __shfl_down_sync(0xffffffffu, applied, o);
            sum += __shfl_down_sync(0xffffffffu, sum, o);
        }
        if ((tid & 31) == 0) { s_red[tid >> 5] = applied; s_red[AW_THREADS / 32 + (tid >> 5)] = sum; }
        __syncthreads();
        uint32_t a_all = 0, s_all = 0;
        for (int w = 0; w < AW_THREADS / 32; ++w) { a_all += s_red[w]; s_all += s_red[AW_THREADS / 32 + w]; }
        if (a_all != s_all) {  // some counter was hit more often than it can count: saturate properly
            __syncthreads();
            build(std::true_type{});
        }
    }
    // merge the image into the table: one coalesced read-modify-write of the window
    uint4* tbl4 = reinterpret_cast<uint4*>(slice_words_ptr<KIND>(ts, sl.table, sl.slot0) + w_lo);
    for (uint32_t i = tid; i < nwords / 4; i += AW_THREADS) {
        const uint4 dlt = win4[i];
        if ((dlt.x | dlt.y | dlt.z | dlt.w) == 0u) continue;
        uint4 t = __ldcs(tbl4 + i);
        t.x = window_merge<KIND>(t.x, dlt.x);
        t.y = window_merge<KIND>(t.y, dlt.y);
        t.z = window_merge<KIND>(t.z, dlt.z);
        t.w = window_merge<KIND>(t.w, dlt.w);
        __stcs(tbl4 + i, t);
    }
}

// Overflow lists received from the peers (rare path): world lists of up to `cap` (table, slot) records.
template <int KIND>
__global__ void __launch_bounds__(256) k_apply_overflow(const __grid_constant__ TableSet ts, const unsigned long long* __restrict__ lists,
                                                         const uint32_t* __restrict__ counts, uint32_t count_stride, uint32_t cap, int world) {
    for (int q = 0; q < world; ++q) {
        const uint32_t n = min(__ldcg(counts + (size_t)q * count_stride), cap);
        const unsigned long long* src = lists + (size_t)q * cap;
        for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
            const unsigned long long rec = src[i];
            slot_insert<KIND, false>(ts.ptr[(int)(rec >> 59)], rec & ((1ull << 59) - 1));
        }
    }
}

}  // namespace gt
