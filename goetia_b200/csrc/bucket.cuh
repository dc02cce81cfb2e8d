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
//   K1b k_bucket      : stage packed tile -> roll both cyclic hashes -> n_tables fastmods ->
//                       CTA-local counting sort by slice in shared memory -> one global
//                       cursor reservation per (tile, slice) -> coalesced run copies.
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
    uint32_t* const* bptr;            // [n_buckets] entry array of each bucket
    const uint32_t* bcap;             // [n_buckets] capacity in entries
    uint32_t* bfill;                  // [n_buckets] cursor; can exceed bcap (excess handled at once)
    unsigned long long* n_direct;     // updates that overflowed a bucket and were applied directly
    unsigned long long* n_dropped;    // overflowed updates of slots this rank does not hold (an error)
};

// What k_apply walks: one item per (slice, source rank) in slice order.
struct ApplyItem {
    const uint32_t* src;   // entries
    const uint32_t* fill;  // device location of the entry count (a cursor, or a received count)
    uint64_t slot0;        // first slot of the slice in table coordinates
    uint32_t cap;
    uint32_t table;
};

// ------------------------------------------------------------------------------------------
// K1b
// ------------------------------------------------------------------------------------------
template <int KIND, bool CAN, int NT>
__global__ void __launch_bounds__(TILE_THREADS, 2)
k_bucket(const __grid_constant__ WalkArgs a, const __grid_constant__ TableSet ts, const __grid_constant__ ProducePlan bp) {
    extern __shared__ __align__(16) uint64_t smem[];
    const int K = a.K;
    const int halo_words = ((K - 1 + 31) >> 5) + 1;
    const int tile_words = TILE_THREADS + halo_words;
    const int nb = bp.n_buckets;
    const int nt = NT > 0 ? NT : bp.n_tables;
    ulonglong2* tab = reinterpret_cast<ulonglong2*>(smem);           // 8 x 16 B
    uint64_t* hs = smem + 16;                                        // TILE_POS hashes, [i][tid]
    uint64_t* sw = hs + TILE_POS;                                    // packed tile + halo
    uint32_t* sorted = reinterpret_cast<uint32_t*>(sw + tile_words + (tile_words & 1));  // TILE_POS offsets
    uint32_t* hist = sorted + TILE_POS;                              // [nb] updates of this tile per bucket
    uint32_t* cur = hist + nb;                                       // [nb] scatter cursor (run end after scatter)
    uint32_t* gpos = cur + nb;                                       // [nb] reserved position in the global bucket
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid < 4) {
        tab[tid] = make_ulonglong2(lemire_T(tid), rotl64(lemire_T(3 - tid), (unsigned)K));
        tab[4 + tid] = make_ulonglong2(rotl64(lemire_T(tid), (unsigned)K), lemire_T(3 - tid));
    }
    const uint64_t n_tiles = (a.n_bases + TILE_POS - 1) / TILE_POS;
    const uint64_t slot_mask = (1ull << bp.shift) - 1;
    unsigned long long direct = 0, dropped = 0;

    for (uint64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        __syncthreads();  // previous tile fully flushed; tab visible
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
        for (int b = tid; b < nb; b += TILE_THREADS) hist[b] = 0;
        __syncthreads();

        // ---- phase 1: roll, count per bucket, park the hash values ---------------------------
        uint32_t vm = 0;  // bit i: window p0+i is a k-mer of a usable read
        const uint64_t p0 = tile * TILE_POS + (uint64_t)tid * POS_PER_THREAD;
        if (p0 < a.n_bases) {
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
            const uint64_t wo = sw[tid];
            uint64_t win;
            {
                int a0 = tid + ((K - 1) >> 5);
                unsigned sh = 2u * (unsigned)((K - 1) & 31);
                win = sh ? (sw[a0] >> sh) | (sw[a0 + 1] << (64u - sh)) : sw[a0];
            }
            uint64_t r = __ldg(a.coarse + (p0 >> COARSE_SHIFT));
            uint64_t rend = __ldg(a.offsets + r + 1) - a.base0;
            {
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
            }
            bool rok = !(__ldg(a.flags + r) & READ_INVALID);
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
                    do {
                        ++r;
                        rend = __ldg(a.offsets + r + 1) - a.base0;
                    } while (p >= rend);
                    rok = !(__ldg(a.flags + r) & READ_INVALID);
                }
                if (rok && p + (uint64_t)K <= rend) {
                    const uint64_t h = CAN ? (fw < rc ? fw : rc) : fw;
                    hs[i * TILE_THREADS + tid] = h;
                    vm |= 1u << i;
#pragma unroll
                    for (int t = 0; t < nt; ++t) {
                        uint64_t bin = fastmod_u64(h, ts.size[t], ts.magic[t]);
                        atomicAdd(&hist[bp.first[t] + (uint32_t)(bin >> bp.shift)], 1u);
                    }
                }
            }
        }
        __syncthreads();

        // ---- reserve room in the global buckets; per-table exclusive scan of the counts -------
        for (int b = tid; b < nb; b += TILE_THREADS) {
            uint32_t c = hist[b];
            gpos[b] = c ? atomicAdd(bp.bfill + b, c) : 0u;
        }
        for (int t = warp; t < nt; t += TILE_THREADS / 32) {
            const int b0 = (int)bp.first[t], b1 = (int)bp.first[t + 1];
            uint32_t run = 0;
            for (int b = b0; b < b1; b += 32) {
                uint32_t c = (b + lane < b1) ? hist[b + lane] : 0u, incl = c;
                for (int o = 1; o < 32; o <<= 1) {
                    uint32_t v = __shfl_up_sync(0xffffffffu, incl, o);
                    if (lane >= o) incl += v;
                }
                if (b + lane < b1) cur[b + lane] = run + incl - c;
                run += __shfl_sync(0xffffffffu, incl, 31);
            }
        }
        __syncthreads();

        // ---- per table: scatter offsets into bucket order, then copy the runs out ---------------
        for (int t = 0; t < nt; ++t) {
            const uint64_t size = ts.size[t], magic = ts.magic[t];
            const uint32_t fb = bp.first[t];
            uint32_t m = vm;
            while (m) {
                int i = __ffs(m) - 1;
                m &= m - 1;
                uint64_t bin = fastmod_u64(hs[i * TILE_THREADS + tid], size, magic);
                uint32_t pos = atomicAdd(&cur[fb + (uint32_t)(bin >> bp.shift)], 1u);
                sorted[pos] = (uint32_t)(bin & slot_mask);
            }
            __syncthreads();
            for (uint32_t b = fb + warp; b < bp.first[t + 1]; b += TILE_THREADS / 32) {
                const uint32_t n = hist[b];
                if (!n) continue;
                const uint32_t src = cur[b] - n, g = gpos[b], cap = __ldg(bp.bcap + b);
                uint32_t* dst = bp.bptr[b];
                for (uint32_t e = lane; e < n; e += 32) {
                    uint32_t off = sorted[src + e];
                    if (g + e < cap) {
                        dst[g + e] = off;
                    } else {  // bucket full (skewed input): apply here, correctness never depends on capacity
                        const uint64_t bin = ((uint64_t)(b - fb) << bp.shift) + off;
                        if (bin >= bp.own_lo[t] && bin < bp.own_hi[t]) {
                            slot_insert<KIND, false>(ts.ptr[t], bin);
                            ++direct;
                        } else {
                            ++dropped;
                        }
                    }
                }
            }
            __syncthreads();
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
            slot_insert<KIND, false>(tbl, x.x);
            slot_insert<KIND, false>(tbl, x.y);
            slot_insert<KIND, false>(tbl, x.z);
            slot_insert<KIND, false>(tbl, x.w);
        }
    } else {
        for (uint32_t e = threadIdx.x; e < n; e += AP_THREADS) slot_insert<KIND, false>(tbl, __ldcs(src + e0 + e));
    }
}

}  // namespace gt
