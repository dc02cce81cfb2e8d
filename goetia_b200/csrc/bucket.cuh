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

#include "kernels.cuh"

namespace gt {

constexpr int BK_MAX_BUCKETS = 1024;  // all tables together (3 x 4 KB of shared memory)
constexpr int AP_THREADS = 256;
constexpr int AP_PER_THREAD = 16;
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
    uint32_t m32[MAX_TABLES];
    int32_t rs[MAX_TABLES];
    uint32_t stage_cap;               // shared-memory staging entries per bucket and sub-step (multiple of 4)
    uint32_t piece;                   // entries a CTA takes from a bucket's cursor at a time (0: exact-size requests)
    uint32_t* const* bptr;            // [n_buckets] entry array of each bucket
    const uint32_t* bcap;             // [n_buckets] capacity in entries (multiple of 4)
    uint32_t* bfill;                  // [n_buckets] cursor; can exceed bcap (excess handled at once)
    unsigned long long* n_direct;     // updates that overflowed a bucket and were applied directly
    unsigned long long* n_dropped;    // overflowed updates of foreign slots that found the overflow list full too (an error)
    // Sharded storages, peer transport: an update that overflows the bucket of a slice held by rank q is
    // posted as a full (table, slot) record to this rank's overflow list in q's inbox (rare path).
    // Counting storages: an update that overflows the bucket of a slice held HERE is not applied by k_bucket
    // itself (a saturating CAS must not run beside the optimistic k_apply_add of the other store, see K2 for
    // the counting storages); it is parked in the store's spill list and applied after that store's buckets.
    unsigned long long* own_spill;    // [own_spill_cap] (table, slot) records; NULL for BitStorage
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

// an update of a slot held here whose bucket is full: BitStorage applies it on the spot (OR commutes with
// everything); the counting storages park it (see ProducePlan::own_spill)
template <int KIND>
__device__ __forceinline__ void apply_own_overflow(const TableSet& ts, const ProducePlan& bp, int t, uint64_t bin,
                                                   unsigned long long& direct, unsigned long long& dropped) {
    if constexpr (KIND == 0) {
        slot_insert<0, false>(ts.ptr[t], bin);
        ++direct;
    } else {
        const uint32_t k = atomicAdd(bp.own_spill_fill, 1u);
        if (k < bp.own_spill_cap) {
            bp.own_spill[k] = ovf_pack(t, bin);
            ++direct;
        } else {
            ++dropped;
        }
    }
}

template <int KIND>
__device__ __forceinline__ void bucket_spill(const TableSet& ts, const ProducePlan& bp, int t, uint32_t b, uint32_t off,
                                             unsigned long long& direct, unsigned long long& dropped) {
    // slow path: a 16 B group holding this one entry, or -- bucket full -- the table itself
    const uint32_t g = atomicAdd(bp.bfill + b, 4u), cap = __ldg(bp.bcap + b);
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

template <int KIND, bool CAN, int NT>
__global__ void __launch_bounds__(TILE_THREADS, 4)  // <= 64 registers: leaves room for two k_apply CTAs per SM beside three of these
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
    uint32_t* stage = reinterpret_cast<uint32_t*>(sw + tile_words + (tile_words & 1));  // [nb][C]
    uint32_t* cnt = stage + (size_t)nb * C;                          // [2][nb] appends of this / the next sub-step
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
    const uint64_t n_tiles = (a.n_bases + TILE_POS - 1) / TILE_POS;
    const uint32_t slot_mask = (uint32_t)((1ull << bp.shift) - 1);
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

        for (int sub = 0; sub < POS_PER_THREAD / BK_SUB; ++sub) {
            uint32_t* cn = cnt + phase * nb;
            // ---- append: roll, reduce, stage ------------------------------------------------------
            if (live) {
#pragma unroll 4
                for (int ii = 0; ii < BK_SUB; ++ii) {
                    const int i = sub * BK_SUB + ii;
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
                        do {
                            ++r;
                            rend = __ldg(a.offsets + r + 1) - a.base0;
                        } while (p >= rend);
                        rok = !(__ldg(a.flags + r) & READ_INVALID);
                    }
                    if (rok && p + (uint64_t)K <= rend) {
                        const uint64_t h = CAN ? (fw < rc ? fw : rc) : fw;  // Canonical::value(), canonical.hh:124-126
#pragma unroll
                        for (int t = 0; t < nt; ++t) {
                            const uint64_t bin = bin_of(h, ts.size[t], ts.magic[t], bp.dsh[t], bp.m32[t], bp.rs[t]);
                            const uint32_t b = bp.first[t] + (uint32_t)(bin >> bp.shift);
                            const uint32_t off = (uint32_t)bin & slot_mask;
                            const uint32_t pos = atomicAdd(&cn[b], 1u);
                            if (pos < C) stage[b * C + pos] = off;
                            else bucket_spill<KIND>(ts, bp, t, b, off, direct, dropped);
                        }
                    }
                }
            }
            __syncthreads();
            // ---- copy the rows out.  Lane l of warp w owns buckets w + 8*l + 256*rr (rr < 4) and keeps
            // their write position in registers: room in the global bucket is taken R entries at a
            // time and the next piece is requested one flush ahead, so the latency of the global
            // cursor atomic is hidden behind the next append phase (R == 0: exact-size requests).
            uint32_t* cz = cnt + (phase ^ 1) * nb;
#pragma unroll
            for (int rr = 0; rr < BK_OWN; ++rr) {
                const int base = warp + rr * TILE_THREADS;
                if (base >= nb) break;
                const int mine = base + 8 * lane;  // TILE_THREADS / 32 == 8 warps
                uint32_t n = 0, d0 = 0, l0 = 0, d1 = 0, l1 = 0;
                if (mine < nb) {
                    n = min(cn[mine], C);
                    cz[mine] = 0;  // the other half is idle now: clear it for the next sub-step
                    const uint32_t n4 = (n + 3u) & ~3u;
                    if (R == 0) {
                        if (n4) { d0 = atomicAdd(bp.bfill + mine, n4); l0 = n4; }
                    } else if (n4) {
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
                }
                const unsigned have = __ballot_sync(0xffffffffu, n != 0);
                for (unsigned m = have; m; m &= m - 1) {
                    const int l = __ffs(m) - 1;
                    const uint32_t bn = __shfl_sync(0xffffffffu, n, l);
                    const uint32_t bd0 = __shfl_sync(0xffffffffu, d0, l), bl0 = __shfl_sync(0xffffffffu, l0, l);
                    const uint32_t bd1 = __shfl_sync(0xffffffffu, d1, l);
                    const uint32_t b = (uint32_t)(base + 8 * l);
                    const uint32_t cap = __ldg(bp.bcap + b);
                    uint32_t* dst = bp.bptr[b];
                    const uint32_t* row = stage + b * C;
                    for (uint32_t e = 4 * lane; e < bn; e += 128) {
                        uint4 v = *reinterpret_cast<const uint4*>(row + e);
                        if (e + 1 >= bn) v.y = BK_PAD;
                        if (e + 2 >= bn) v.z = BK_PAD;
                        if (e + 3 >= bn) v.w = BK_PAD;
                        const uint32_t at = e < bl0 ? bd0 + e : bd1 + (e - bl0);
                        if (at < cap) {
                            *reinterpret_cast<uint4*>(dst + at) = v;
                        } else {  // bucket full (skewed input): apply here, correctness never depends on capacity
                            int t = 0;
                            while (b >= bp.first[t + 1]) ++t;
                            const uint32_t offs[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                            for (int k = 0; k < 4; ++k) {
                                if (offs[k] == BK_PAD) continue;
                                const uint64_t bin = ((uint64_t)(b - bp.first[t]) << bp.shift) + offs[k];
                                if (bin >= bp.own_lo[t] && bin < bp.own_hi[t]) {
                                    apply_own_overflow<KIND>(ts, bp, t, bin, direct, dropped);
                                } else {
                                    post_foreign(bp, t, b, bin, dropped);
                                }
                            }
                        }
                    }
                }
            }
            __syncthreads();
            phase ^= 1;
        }
    }
    // what is left of the pieces this CTA took is filled with pad entries (k_apply skips them)
    if (R) {
#pragma unroll
        for (int rr = 0; rr < BK_OWN; ++rr) {
            const int mine = warp + rr * TILE_THREADS + 8 * lane;
            if (mine >= nb) continue;
            const uint32_t cap = __ldg(bp.bcap + mine);
            uint32_t* dst = bp.bptr[mine];
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
// only_failed != NULL: apply only the slices whose conservation check failed (replay pass of the counting
// storages, see below); items_per_slice consecutive items belong to one slice.
template <int KIND>
__global__ void __launch_bounds__(AP_THREADS)
k_apply(const __grid_constant__ TableSet ts, const ApplyItem* __restrict__ items, const uint32_t* __restrict__ chunk_start,
        int n_items, const unsigned long long* __restrict__ only_failed = nullptr, int items_per_slice = 1) {
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
    if (only_failed && __ldcg(only_failed + b / (uint32_t)items_per_slice) == 0ull) return;
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

// ------------------------------------------------------------------------------------------
// K2 for the counting storages (ByteStorage / NibbleStorage).
//
// A saturating increment by CAS costs a load plus a compare-and-swap per update and runs at a fraction of the
// fire-and-forget atomic rate (measured on B200: ~26 G updates/s against ~200 G RED/s; an atomicAdd that
// returns its old value manages ~55 G/s).  But saturation is rare, and 32-bit word arithmetic is linear: as long
// as no counter of a slice is hit while it stands at its maximum, a plain RED.ADD of 1 << shift IS the saturating
// increment.  So a slice is applied optimistically and checked by a conservation law:
//   k_slice_sums<-1>   delta[slice] -= sum of all counters of the slice          (before)
//   k_apply_red        every update is one RED.ADD; delta[slice] -= updates applied
//   k_slice_sums<+1>   delta[slice] += sum of all counters of the slice          (after)
// An add that lands on a counter below its maximum raises the slice's counter sum by exactly 1; one that lands on
// a counter AT its maximum wraps it (and carries into the neighbour, or off the word) and changes the sum by
// <= 1 - max.  Hence delta[slice] == 0 <=> no add wrapped <=> the slice already holds min(max, hits).  For the
// (rare) slices with delta != 0:
//   k_apply_red<UNDO>  subtracts every update again -- exact whatever the wraps did to neighbouring counters,
//                      because add and subtract cancel mod 2^32;
//   k_apply(only_failed) replays the updates with the saturating CAS.
// A slice that failed is remembered as hot and goes straight to the CAS replay in the following applies (4-bit
// counters saturate routinely once a table fills up); every 8th apply clears the flags and tries again.
// GT_APPLY_CAS=1 forces the plain CAS apply.
// Nothing else writes the tables meanwhile: k_bucket parks its own overflow (ProducePlan::own_spill) and the direct
// writers are ordered against the apply stream (direct_begin / direct_end).  Final bytes are bit-exact.
// Items of one slice are consecutive (items_per_slice = ranks that produced for it): slice = item / items_per_slice.
// ------------------------------------------------------------------------------------------
struct SliceDesc {
    uint64_t slot0, slots;
    uint32_t table;
};

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

// sum of the 4 byte counters / 8 nibble counters of a table word
template <int KIND>
__device__ __forceinline__ uint32_t counter_sum(uint32_t x) {
    if constexpr (KIND == 1) return __dp4a(x, 0x01010101u, 0u);
    else return __dp4a(x & 0x0f0f0f0fu, 0x01010101u, __dp4a((x >> 4) & 0x0f0f0f0fu, 0x01010101u, 0u));
}

// grid (x, n_slices): delta[slice] += SIGN * (sum of the counters of the slice)
template <int KIND, int SIGN>
__global__ void __launch_bounds__(256) k_slice_sums(const __grid_constant__ TableSet ts, const SliceDesc* __restrict__ slices,
                                                     unsigned long long* __restrict__ delta) {
    const SliceDesc sd = slices[blockIdx.y];
    constexpr uint32_t spw = KIND == 1 ? 4u : 8u;
    const uint32_t* tbl = ts.ptr[sd.table] + sd.slot0 / spw;  // slices start on a word boundary
    const uint64_t n_words = (sd.slots + spw - 1) / spw;      // a table's last word may be partial: its padding is zero
    unsigned long long mine = 0;
    const uint4* v = reinterpret_cast<const uint4*>(tbl);
    const uint64_t n4 = (reinterpret_cast<uintptr_t>(tbl) & 15) == 0 ? n_words / 4 : 0;
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n4; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint4 x = __ldcg(v + i);
        mine += counter_sum<KIND>(x.x) + counter_sum<KIND>(x.y) + counter_sum<KIND>(x.z) + counter_sum<KIND>(x.w);
    }
    for (uint64_t w = n4 * 4 + blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; w < n_words; w += (uint64_t)gridDim.x * blockDim.x)
        mine += counter_sum<KIND>(__ldcg(tbl + w));
    for (int o = 16; o; o >>= 1) mine += __shfl_down_sync(0xffffffffu, mine, o);
    if ((threadIdx.x & 31) == 0 && mine) atomicAdd(delta + blockIdx.y, SIGN > 0 ? mine : 0ull - mine);
}

// UNDO = false: the optimistic pass.  UNDO = true: subtract the same updates again, only in the slices whose
// check failed (delta != 0 after k_slice_sums<+1>); delta is left untouched so that the replay sees it too.
// hot[slice] != 0: the slice saturated in a recent apply; it is left to the CAS replay straight away (no adds,
// nothing to undo) until the flags are cleared again (every 8th apply retries the optimistic path).
template <int KIND, bool UNDO>
__global__ void __launch_bounds__(AP_THREADS)
k_apply_red(const __grid_constant__ TableSet ts, const ApplyItem* __restrict__ items, const uint32_t* __restrict__ chunk_start,
            int n_items, int items_per_slice, unsigned long long* __restrict__ delta, const uint32_t* __restrict__ hot) {
    static_assert(KIND == 1 || KIND == 2, "counting storages only");
    __shared__ uint32_t s_b;
    if (threadIdx.x == 0) {
        uint32_t lo = 0, hi = (uint32_t)n_items;
        while (hi - lo > 1) {
            uint32_t mid = (lo + hi) >> 1;
            if (__ldg(chunk_start + mid) <= blockIdx.x) lo = mid; else hi = mid;
        }
        s_b = lo;
    }
    __syncthreads();
    const uint32_t b = s_b;
    if (__ldcg(hot + b / (uint32_t)items_per_slice) != 0u) return;
    if (UNDO && __ldcg(delta + b / (uint32_t)items_per_slice) == 0ull) return;
    const ApplyItem it = items[b];
    const uint32_t fill = min(__ldcg(it.fill), it.cap);
    const uint32_t e0 = (blockIdx.x - __ldg(chunk_start + b)) * (uint32_t)AP_CHUNK;
    if (e0 >= fill) return;
    uint32_t* tbl = ts.ptr[it.table] + (KIND == 1 ? (it.slot0 >> 2) : (it.slot0 >> 3));
    const uint32_t* src = it.src;
    const uint32_t n = min((uint32_t)AP_CHUNK, fill - e0);
    uint32_t applied = 0;
    auto one = [&](uint32_t off) {
        if (off == BK_PAD) return;
        uint32_t word, sh;
        counter_addr<KIND>(off, word, sh);
        atomicAdd(tbl + word, UNDO ? 0u - (1u << sh) : (1u << sh));  // result unused: RED.ADD
        ++applied;
    };
    if (n == AP_CHUNK && ((reinterpret_cast<uintptr_t>(src + e0) & 15) == 0)) {
        const uint4* v = reinterpret_cast<const uint4*>(src + e0);
#pragma unroll
        for (int k = 0; k < AP_PER_THREAD / 4; ++k) {
            const uint4 x = __ldcs(v + k * AP_THREADS + threadIdx.x);
            one(x.x); one(x.y); one(x.z); one(x.w);
        }
    } else {
        for (uint32_t e = threadIdx.x; e < n; e += AP_THREADS) one(__ldcs(src + e0 + e));
    }
    if (!UNDO) {
        for (int o = 16; o; o >>= 1) applied += __shfl_down_sync(0xffffffffu, applied, o);
        if ((threadIdx.x & 31) == 0 && applied) atomicAdd(delta + b / (uint32_t)items_per_slice, 0ull - (unsigned long long)applied);
    }
}

// after the undo pass: slices that failed the check become hot; hot slices (old and new) are what the replay applies
__global__ void __launch_bounds__(256) k_mark_hot(unsigned long long* __restrict__ delta, uint32_t* __restrict__ hot, int n_slices) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n_slices) return;
    const uint32_t h = (hot[s] != 0u || delta[s] != 0ull) ? 1u : 0u;
    hot[s] = h;
    delta[s] = h;
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
