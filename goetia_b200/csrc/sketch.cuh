// goetia_b200/csrc/sketch.cuh -- K4: scaled / bottom-k MinHash (SourmashSketch) on sm_100a.
//
// Per window: skip if it holds a non-ACGT base (add_sequence(force=true),
// sketches/sourmash_sketch.hh:71-81), take the lexicographically smaller of the k-mer and its
// reverse complement, expand it back to ASCII, MurmurHash3_x64_128(word, K, seed)[0]
// (the hash libsourmash applies: sketches/sourmash/sourmash.hpp:19-21), keep it iff
// h <= limit.  Kept hashes go straight into a device-resident open-addressing hash set
// (64-bit atomicCAS), which performs the dedupe; ALU-bound, HBM traffic is only the packed
// input (0.25 B/base) plus ~0.1 % of windows touching the set.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "kernels.cuh"

namespace gt {

constexpr uint64_t SET_EMPTY = ~0ull;

__device__ __forceinline__ uint64_t fmix64(uint64_t k) {
    k ^= k >> 33;
    k *= 0xff51afd7ed558ccdull;
    k ^= k >> 33;
    k *= 0xc4ceb9fe1a85ec53ull;
    k ^= k >> 33;
    return k;
}

// four 2-bit codes (first base in the low bits of x4) -> four ASCII bytes, first base lowest
__device__ __forceinline__ uint32_t ascii4(uint32_t x4) {
    uint32_t sel = (x4 & 3u) | ((x4 & 0xCu) << 2) | ((x4 & 0x30u) << 4) | ((x4 & 0xC0u) << 6);
    return __byte_perm(0x54474341u /* "ACGT" */, 0u, sel);
}

// reverse the order of the 2-bit groups of a 64-bit word
__device__ __forceinline__ uint64_t pair_reverse64(uint64_t x) {
    uint64_t y = __brevll(x);
    return ((y & 0x5555555555555555ull) << 1) | ((y >> 1) & 0x5555555555555555ull);
}

// A window of up to 64 bases, little-endian packed (first base in the low bits of lo).
struct Win {
    uint64_t lo, hi;
};

template <int NW>
__device__ __forceinline__ Win win_revcomp(Win x, int K) {
    // reverse complement: complement every code (3-c == ~c on 2 bits), reverse group order,
    // shift down so the first base of the result sits at bit 0
    Win r;
    if constexpr (NW == 1) {
        unsigned sh = 64u - 2u * (unsigned)K;
        r.lo = pair_reverse64(~x.lo) >> sh;
        r.hi = 0;
    } else {
        uint64_t a = pair_reverse64(~x.hi), b = pair_reverse64(~x.lo);  // 128-bit value (a = low half after reversal)
        unsigned sh = 128u - 2u * (unsigned)K;                          // 0..62 for K in 33..64
        if (sh == 0) { r.lo = a; r.hi = b; }
        else { r.lo = (a >> sh) | (b << (64u - sh)); r.hi = b >> sh; }
    }
    return r;
}

// lexicographic "less or equal" of two windows == numeric compare of their big-endian forms;
// comparing from the first base on: find the lowest differing 2-bit group.
template <int NW>
__device__ __forceinline__ bool win_lex_le(Win a, Win b) {
    uint64_t d = a.lo ^ b.lo;
    uint64_t xa = a.lo, xb = b.lo;
    if (NW == 2 && d == 0) { d = a.hi ^ b.hi; xa = a.hi; xb = b.hi; }
    if (d == 0) return true;
    int bit = __ffsll((long long)d) - 1;
    int g = bit & ~1;
    return ((xa >> g) & 3) < ((xb >> g) & 3);
}

// MurmurHash3_x64_128 (published algorithm) of the K ASCII bytes of window w; returns h1.
template <int NW>
__device__ __forceinline__ uint64_t murmur_window(Win w, int K, uint32_t seed) {
    const uint64_t c1 = 0x87c37b91114253d5ull, c2 = 0x4cf5ad432745937full;
    uint64_t h1 = seed, h2 = seed;
    // byte j of the key is base j; 8 bases (16 bits) make one u64 of the key
    auto key64 = [&](int q) -> uint64_t {  // bytes 8q .. 8q+7 (bases beyond K read as garbage; masked by caller)
        uint32_t bits;
        int sh = 16 * q;
        if (sh < 64) bits = (uint32_t)(w.lo >> sh) & 0xffffu;
        else bits = (uint32_t)(w.hi >> (sh - 64)) & 0xffffu;
        uint32_t a = ascii4(bits & 0xffu), b = ascii4(bits >> 8);
        return (uint64_t)a | ((uint64_t)b << 32);
    };
    const int nblocks = K >> 4;
#pragma unroll
    for (int i = 0; i < (NW == 1 ? 2 : 4); ++i) {
        if (i < nblocks) {
            uint64_t k1 = key64(2 * i), k2 = key64(2 * i + 1);
            k1 *= c1; k1 = (k1 << 31) | (k1 >> 33); k1 *= c2; h1 ^= k1;
            h1 = (h1 << 27) | (h1 >> 37); h1 += h2; h1 = h1 * 5 + 0x52dce729;
            k2 *= c2; k2 = (k2 << 33) | (k2 >> 31); k2 *= c1; h2 ^= k2;
            h2 = (h2 << 31) | (h2 >> 33); h2 += h1; h2 = h2 * 5 + 0x38495ab5;
        }
    }
    const int t = K & 15;
    if (t) {
        uint64_t k1 = key64(2 * nblocks), k2 = t > 8 ? key64(2 * nblocks + 1) : 0;
        if (t < 8) k1 &= (1ull << (8 * t)) - 1;
        if (t > 8) {
            if (t < 16) k2 &= (1ull << (8 * (t - 8))) - 1;
            k2 *= c2; k2 = (k2 << 33) | (k2 >> 31); k2 *= c1; h2 ^= k2;
        }
        k1 *= c1; k1 = (k1 << 31) | (k1 >> 33); k1 *= c2; h1 ^= k1;
    }
    h1 ^= (uint64_t)K; h2 ^= (uint64_t)K;
    h1 += h2; h2 += h1;
    h1 = fmix64(h1); h2 = fmix64(h2);
    h1 += h2;
    return h1;
}

struct SetView {
    uint64_t* slots;
    uint64_t mask;              // capacity - 1 (capacity is a power of two)
    unsigned long long* count;  // live entries
    uint64_t max_count;         // refuse inserts beyond this load and raise *overflow
    int* overflow;
    int* has_max;               // the value ~0ull itself is a member (cannot live in a slot)
};

__device__ __forceinline__ void set_insert(const SetView& s, uint64_t h) {
    if (h == SET_EMPTY) { *s.has_max = 1; return; }
    if (*reinterpret_cast<volatile unsigned long long*>(s.count) >= s.max_count) { *s.overflow = 1; return; }
    uint64_t idx = (h ^ (h >> 29)) & s.mask;
    for (uint64_t probe = 0; probe <= s.mask; ++probe) {
        unsigned long long old = atomicCAS(reinterpret_cast<unsigned long long*>(s.slots + idx), SET_EMPTY, h);
        if (old == SET_EMPTY) {
            unsigned long long c = atomicAdd(s.count, 1ull);
            if (c + 1 >= s.max_count) *s.overflow = 1;  // everything so far is stored; caller grows + replays
            return;
        }
        if (old == h) return;
        idx = (idx + 1) & s.mask;
    }
    *s.overflow = 1;
}

struct SketchArgs {
    const uint64_t* words;
    uint64_t n_words_alloc;
    const uint32_t* nmask;    // bit j of nmask[w] = base 32w+j is not ACGT
    const uint64_t* offsets;
    const uint32_t* coarse;
    uint64_t base0, n_reads, n_bases;
    int K;
    uint32_t seed;
    const uint64_t* limit;    // device scalar: keep h <= *limit
};

// Same tiling as k_walk: CTA = 8192 window starts, packed words + N-mask staged in smem.
template <int NW>
__global__ void __launch_bounds__(TILE_THREADS) k_sketch(const __grid_constant__ SketchArgs a, const __grid_constant__ SetView set) {
    extern __shared__ __align__(16) uint64_t smem[];
    const int K = a.K;
    const int halo_words = ((K - 1 + 31) >> 5) + 1;
    const int tile_words = TILE_THREADS + halo_words;
    uint64_t* sw = smem;
    uint32_t* sn = reinterpret_cast<uint32_t*>(smem + tile_words + (tile_words & 1));
    const int tid = threadIdx.x;
    const uint64_t n_tiles = (a.n_bases + TILE_POS - 1) / TILE_POS;
    const uint64_t limit = *a.limit;
    const uint64_t kmask_lo = (K >= 32) ? ~0ull : ((1ull << (2 * K)) - 1);
    const uint64_t kmask_hi = (K <= 32) ? 0ull : (K >= 64 ? ~0ull : ((1ull << (2 * (K - 32))) - 1));

    for (uint64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        __syncthreads();
        const uint64_t w0 = tile * TILE_THREADS;
        for (int i = tid; i < tile_words; i += TILE_THREADS) {
            uint64_t gi = w0 + i;
            bool in = gi < a.n_words_alloc;
            sw[i] = in ? a.words[gi] : 0;
            sn[i] = in ? a.nmask[gi] : 0;
        }
        __syncthreads();
        const uint64_t p0 = tile * TILE_POS + (uint64_t)tid * POS_PER_THREAD;
        if (p0 >= a.n_bases) continue;

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
        const uint64_t b0 = sw[tid], b1 = sw[tid + 1], b2 = NW == 2 ? sw[tid + 2] : 0;
        // invalid-base bits for positions p0 .. p0+95
        const uint64_t n01 = (uint64_t)sn[tid] | ((uint64_t)sn[tid + 1] << 32);
        const uint64_t n2 = NW == 2 ? (uint64_t)sn[tid + 2] : 0;

#pragma unroll 2
        for (int i = 0; i < POS_PER_THREAD; ++i) {
            const uint64_t p = p0 + i;
            if (p >= a.n_bases) break;
            while (p >= rend) { ++r; rend = __ldg(a.offsets + r + 1) - a.base0; }
            if (p + (uint64_t)K > rend) continue;
            // N check over [p, p+K)
            uint64_t nm = n01 >> i;
            if (NW == 2 && i) nm |= n2 << (64 - i);
            bool bad;
            if (K < 64) bad = (nm & ((1ull << K) - 1)) != 0;
            else bad = nm != 0;
            if (bad) continue;
            Win x;
            const unsigned sh = 2u * (unsigned)i;
            x.lo = sh ? (b0 >> sh) | (b1 << (64u - sh)) : b0;
            x.hi = 0;
            if (NW == 2) x.hi = sh ? (b1 >> sh) | (b2 << (64u - sh)) : b1;
            x.lo &= kmask_lo;
            x.hi &= kmask_hi;
            Win rcw = win_revcomp<NW>(x, K);
            Win c = win_lex_le<NW>(x, rcw) ? x : rcw;
            uint64_t h = murmur_window<NW>(c, K, a.seed);
            if (h <= limit) set_insert(set, h);
        }
    }
}

// add raw hash values (MinHash::add_hash)
__global__ void __launch_bounds__(256) k_set_add(const uint64_t* __restrict__ h, uint64_t n, const uint64_t* limit,
                                                  const __grid_constant__ SetView set) {
    const uint64_t lim = *limit;
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        uint64_t v = h[i];
        if (v <= lim) set_insert(set, v);
    }
}

// compact the live slots into out[] (order arbitrary); warp-aggregated append
__global__ void __launch_bounds__(256) k_set_extract(const uint64_t* __restrict__ slots, uint64_t cap, uint64_t* __restrict__ out,
                                                      unsigned long long* __restrict__ n_out) {
    for (uint64_t i0 = (blockIdx.x * (uint64_t)blockDim.x + threadIdx.x) & ~31ull; i0 < cap; i0 += (uint64_t)gridDim.x * blockDim.x) {
        uint64_t i = i0 + (threadIdx.x & 31);
        uint64_t v = i < cap ? slots[i] : SET_EMPTY;
        unsigned m = __ballot_sync(0xffffffffu, v != SET_EMPTY);
        if (m) {
            int lane = threadIdx.x & 31;
            unsigned long long base = 0;
            if (lane == 0) base = atomicAdd(n_out, (unsigned long long)__popc(m));
            base = __shfl_sync(0xffffffffu, base, 0);
            if (v != SET_EMPTY) out[base + __popc(m & ((1u << lane) - 1))] = v;
        }
    }
}

// re-insert every live slot of an old table into a new (larger) one
__global__ void __launch_bounds__(256) k_set_rehash(const uint64_t* __restrict__ old_slots, uint64_t old_cap,
                                                     const __grid_constant__ SetView set) {
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < old_cap; i += (uint64_t)gridDim.x * blockDim.x) {
        uint64_t v = old_slots[i];
        if (v != SET_EMPTY) set_insert(set, v);
    }
}

}  // namespace gt
