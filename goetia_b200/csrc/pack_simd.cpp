// goetia_b200/csrc/pack_simd.cpp -- host-side 2-bit packing of read bases (the FASTX workers' share of the device
// pipeline: the batches cross PCIe as 0.25 B/base).  Plain C++ for the host compiler; AVX2 where the CPU has it.
//
// Layout = the device layout: base p of a stream sits at bits 2*(p%32) of u64 word p/32,
// A=0 C=1 G=2 T=3, either case (DNA_SIMPLE folds acgt to upper case: sequences/alphabets.hh:112-130); any other
// byte makes the read invalid (the parser skips it: parsing/readers.hh:162-171).
#include <initializer_list>
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#if defined(__x86_64__)
#include <immintrin.h>
#endif

namespace {

// 8 bytes -> 16 bits of codes; *bad set when a byte is not one of ACGTacgt (same arithmetic as k_pack on the device)
inline uint32_t pack8(uint64_t x, bool& bad) {
    uint64_t code = ((x >> 1) & 0x0303030303030303ull) ^ ((x >> 2) & 0x0101010101010101ull);
    const uint64_t u = x & 0xDFDFDFDFDFDFDFDFull, L = 0x7F7F7F7F7F7F7F7Full;
    uint64_t ok = 0;
    for (uint64_t pat : {0x41ull, 0x43ull, 0x47ull, 0x54ull}) {
        const uint64_t t = u ^ (pat * 0x0101010101010101ull);
        ok |= ~(((t & L) + L) | t | L);
    }
    bad |= ok != 0x8080808080808080ull;
    code = (code | (code >> 6)) & 0x000F000F000F000Full;
    code = (code | (code >> 12)) & 0x000000FF000000FFull;
    code = (code | (code >> 24)) & 0xFFFFull;
    return (uint32_t)code;
}

// OR `nbases` codes (2 bits each, nbases <= 32) into the stream at base position pos
inline void put(uint64_t* words, uint64_t pos, uint64_t v, unsigned nbases) {
    const unsigned sh = 2u * (unsigned)(pos & 31);
    words[pos >> 5] |= v << sh;
    if (sh && 2u * nbases > 64u - sh) words[(pos >> 5) + 1] |= v >> (64u - sh);
}

int append_scalar(const unsigned char* s, size_t L, uint64_t* words, uint64_t pos) {
    bool bad = false;
    size_t i = 0;
    for (; i + 32 <= L; i += 32) {
        uint64_t v = 0;
        for (int k = 0; k < 4; ++k) {
            uint64_t x;
            memcpy(&x, s + i + 8 * k, 8);
            v |= (uint64_t)pack8(x, bad) << (16 * k);
        }
        put(words, pos + i, v, 32);
    }
    if (i < L) {
        unsigned char tail[32];
        memset(tail, 'A', sizeof tail);
        memcpy(tail, s + i, L - i);
        uint64_t v = 0;
        for (int k = 0; k < 4; ++k) {
            uint64_t x;
            memcpy(&x, tail + 8 * k, 8);
            v |= (uint64_t)pack8(x, bad) << (16 * k);
        }
        const unsigned n = (unsigned)(L - i);
        put(words, pos + i, n < 32 ? v & ((1ull << (2 * n)) - 1) : v, n);
    }
    return bad ? 1 : 0;
}

#if defined(__x86_64__)
// 32 bytes -> 64 bits of codes
__attribute__((target("avx2"))) inline uint64_t pack32_avx2(__m256i x, __m256i& ok_acc) {
    const __m256i m3 = _mm256_set1_epi8(3), m1 = _mm256_set1_epi8(1);
    const __m256i code = _mm256_xor_si256(_mm256_and_si256(_mm256_srli_epi16(x, 1), m3), _mm256_and_si256(_mm256_srli_epi16(x, 2), m1));
    const __m256i u = _mm256_and_si256(x, _mm256_set1_epi8((char)0xDF));
    const __m256i ok = _mm256_or_si256(_mm256_or_si256(_mm256_cmpeq_epi8(u, _mm256_set1_epi8('A')), _mm256_cmpeq_epi8(u, _mm256_set1_epi8('C'))),
                                       _mm256_or_si256(_mm256_cmpeq_epi8(u, _mm256_set1_epi8('G')), _mm256_cmpeq_epi8(u, _mm256_set1_epi8('T'))));
    ok_acc = _mm256_and_si256(ok_acc, ok);
    // c0 + 4 c1 per 16-bit lane, then (..) + 16 (..) per 32-bit lane: one byte of codes per 4 bases
    const __m256i p16 = _mm256_maddubs_epi16(code, _mm256_set1_epi16(0x0401));
    const __m256i p32 = _mm256_madd_epi16(p16, _mm256_set1_epi32(0x00100001));
    // gather the low byte of each 32-bit lane: 4 bytes per 128-bit half
    const __m256i sh = _mm256_setr_epi8(0, 4, 8, 12, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1,
                                        0, 4, 8, 12, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1);
    const __m256i g = _mm256_shuffle_epi8(p32, sh);
    const uint64_t lo = (uint32_t)_mm256_extract_epi32(g, 0), hi = (uint32_t)_mm256_extract_epi32(g, 4);
    return lo | (hi << 32);
}

__attribute__((target("avx2"))) int append_avx2(const unsigned char* s, size_t L, uint64_t* words, uint64_t pos) {
    __m256i ok = _mm256_set1_epi8((char)0xFF);
    size_t i = 0;
    for (; i + 32 <= L; i += 32) {
        const __m256i x = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(s + i));
        put(words, pos + i, pack32_avx2(x, ok), 32);
    }
    if (i < L) {
        unsigned char tail[32];
        memset(tail, 'A', sizeof tail);
        memcpy(tail, s + i, L - i);
        const __m256i x = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(tail));
        const unsigned n = (unsigned)(L - i);
        const uint64_t v = pack32_avx2(x, ok);
        put(words, pos + i, n < 32 ? v & ((1ull << (2 * n)) - 1) : v, n);
    }
    return _mm256_movemask_epi8(ok) == -1 ? 0 : 1;
}
#endif

using append_fn = int (*)(const unsigned char*, size_t, uint64_t*, uint64_t);

append_fn pick() {
#if defined(__x86_64__)
    const char* e = getenv("GT_HOST_PACK_SCALAR");  // tests compare the two paths
    if (!(e && *e == '1') && __builtin_cpu_supports("avx2")) return append_avx2;
#endif
    return append_scalar;
}

}  // namespace

// OR the 2-bit codes of s[0, L) into `words` (zero where nothing has been written yet; at least (pos + L + 31) / 32 + 1
// words long) at base position pos.  Returns 0 when every byte is one of ACGTacgt, 1 otherwise (the words then hold
// garbage for this read: the caller clears them again).
// (Library-internal: not part of the C ABI of include/goetia_b200.h.)
extern "C" __attribute__((visibility("hidden"))) int gt_host_pack_append(const unsigned char* s, size_t L, uint64_t* words, uint64_t pos) {
    static const append_fn fn = pick();
    return fn(s, L, words, pos);
}
