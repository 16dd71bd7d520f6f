// dcn_host_pack.cpp -- see dcn_host_pack.h.
#include "dcn_host_pack.h"

#include <string.h>

#if defined(__x86_64__)
#include <immintrin.h>
#define DCN_X86 1
#endif

namespace dcn {

static inline bool is_acgt(uint8_t b) {
    uint8_t u = b & 0xDFu;
    return u == 'A' || u == 'C' || u == 'G' || u == 'T';
}

// 32 bases -> 2 code words + 2 mask halves
static inline void pack32_scalar(const uint8_t *p, uint32_t *codes, uint16_t *inv) {
    for (int h = 0; h < 2; h++) {
        uint32_t c = 0, m = 0;
        for (int j = 0; j < 16; j++) {
            uint8_t b = p[16 * h + j];
            c |= (uint32_t)((b >> 1) & 3u) << (2 * j);
            m |= (uint32_t)(!is_acgt(b)) << j;
        }
        codes[h] = c;
        inv[h] = (uint16_t)m;
    }
}

#ifdef DCN_X86
__attribute__((target("avx2,bmi2"))) static void pack_blocks_avx2(const uint8_t *p, uint64_t n_blocks, uint32_t *codes,
                                                                  uint16_t *inv, std::vector<uint64_t> *bad32,
                                                                  std::vector<uint32_t> *bad_mask) {
    // letter expected for each low nibble of the byte: 1 -> 'A', 3 -> 'C', 4 -> 'T', 7 -> 'G'
    const __m256i lut = _mm256_setr_epi8(-1, 0x41, -1, 0x43, 0x54, -1, -1, 0x47, -1, -1, -1, -1, -1, -1, -1, -1,
                                         -1, 0x41, -1, 0x43, 0x54, -1, -1, 0x47, -1, -1, -1, -1, -1, -1, -1, -1);
    const __m256i m0f = _mm256_set1_epi8(0x0F), mdf = _mm256_set1_epi8((char)0xDF);
    const uint64_t M = 0x0606060606060606ULL;
    for (uint64_t i = 0; i < n_blocks; i++) {
        const uint8_t *q = p + 32 * i;
        __m256i v = _mm256_loadu_si256(reinterpret_cast<const __m256i *>(q));
        __m256i want = _mm256_shuffle_epi8(lut, _mm256_and_si256(v, m0f));
        uint32_t ok = (uint32_t)_mm256_movemask_epi8(_mm256_cmpeq_epi8(want, _mm256_and_si256(v, mdf)));
        uint32_t bad = ~ok;
        if (inv) memcpy(inv + 2 * i, &bad, 4);
        if (bad && bad32) { bad32->push_back(i); if (bad_mask) bad_mask->push_back(bad); }
        uint64_t x0, x1, x2, x3;
        memcpy(&x0, q, 8); memcpy(&x1, q + 8, 8); memcpy(&x2, q + 16, 8); memcpy(&x3, q + 24, 8);
        uint64_t c = _pext_u64(x0, M) | (_pext_u64(x1, M) << 16) | (_pext_u64(x2, M) << 32) | (_pext_u64(x3, M) << 48);
        memcpy(codes + 2 * i, &c, 8);
    }
}

// AVX-512 (BW): 64 bases per step.  Codes: (byte >> 1) & 3, then two multiply-adds fold 4 bytes into one
// (c0 + 4 c1 + 16 c2 + 64 c3) and a narrowing move keeps the low byte of every dword.  Mask: compare
// (byte & 0xDF) with the letter its low nibble stands for.  `nt` = streaming stores: the output is a
// pinned staging buffer the GPU's copy engine reads next, so keep it out of the cores' caches.
__attribute__((target("avx512f,avx512bw,avx512vl"))) static void pack_blocks_avx512(const uint8_t *p, uint64_t n_blocks64,
                                                                                   uint32_t *codes, uint16_t *inv, bool nt,
                                                                                   std::vector<uint64_t> *bad32,
                                                                                   std::vector<uint32_t> *bad_mask) {
    const __m512i lut = _mm512_broadcast_i32x4(_mm_setr_epi8(-1, 0x41, -1, 0x43, 0x54, -1, -1, 0x47, -1, -1, -1, -1, -1, -1, -1, -1));
    const __m512i m0f = _mm512_set1_epi8(0x0F), mdf = _mm512_set1_epi8((char)0xDF), m03 = _mm512_set1_epi8(0x03);
    const __m512i mul8 = _mm512_set1_epi16(0x0401), mul16 = _mm512_set1_epi32(0x00100001);
    for (uint64_t i = 0; i < n_blocks64; i++) {
        // runs ahead across 4 KB pages.  T2, not NTA: measured on the host CPUs of this pool, the non-temporal hint
        // costs 30-60 % of the packing rate at every thread count (it keeps the L2 streamer from helping)
        _mm_prefetch(reinterpret_cast<const char *>(p + 64 * i + 4096), _MM_HINT_T2);
        __m512i v = _mm512_loadu_si512(p + 64 * i);
        __m512i want = _mm512_shuffle_epi8(lut, _mm512_and_si512(v, m0f));
        uint64_t bad = ~_mm512_cmpeq_epi8_mask(want, _mm512_and_si512(v, mdf));
        if (bad && bad32) {   // rare: reads are almost all ACGT
            if ((uint32_t)bad) { bad32->push_back(2 * i); if (bad_mask) bad_mask->push_back((uint32_t)bad); }
            if (bad >> 32) { bad32->push_back(2 * i + 1); if (bad_mask) bad_mask->push_back((uint32_t)(bad >> 32)); }
        }
        __m512i c = _mm512_and_si512(_mm512_srli_epi16(v, 1), m03);
        __m512i c4 = _mm512_maddubs_epi16(c, mul8);          // c0 + 4 c1 per 16-bit lane
        __m512i c8 = _mm512_madd_epi16(c4, mul16);           // + 16 * (c2 + 4 c3) per 32-bit lane
        __m128i out = _mm512_cvtepi32_epi8(c8);              // 16 bytes = 64 bases
        if (nt) {
            _mm_stream_si128(reinterpret_cast<__m128i *>(codes + 4 * i), out);
            if (inv) _mm_stream_si64(reinterpret_cast<long long *>(inv + 4 * i), (long long)bad);
        } else {
            _mm_storeu_si128(reinterpret_cast<__m128i *>(codes + 4 * i), out);
            if (inv) memcpy(inv + 4 * i, &bad, 8);
        }
    }
    if (nt) _mm_sfence();
}

static int simd_level() {   // 0 scalar, 1 AVX2 + BMI2, 2 AVX-512 BW
    static const int lvl = []() {
        if (__builtin_cpu_supports("avx512f") && __builtin_cpu_supports("avx512bw") && __builtin_cpu_supports("avx512vl")) return 2;
        if (__builtin_cpu_supports("avx2") && __builtin_cpu_supports("bmi2")) return 1;
        return 0;
    }();
    return lvl;
}
#endif

// ------------------------------------------------------------------ packing + newline flags in one pass
void pack_records(const uint8_t *bases, uint64_t a0, uint64_t nb, const uint64_t *off0, uint32_t nr, uint32_t k,
                  uint32_t prefix_len, uint32_t *codes, uint16_t *inv, uint32_t *nl, std::vector<uint64_t> &bad32,
                  std::vector<uint32_t> *bad_mask) {
    bad32.clear();
    if (bad_mask) bad_mask->clear();
    pack_ascii(bases + a0, nb, codes, inv, 1, &bad32, bad_mask);
    memset(nl, 0, ((size_t)nr + 31) / 32 * 4);
    auto rec_end = [&](uint32_t q) {   // one past the last byte the reference looks at
        const uint64_t len = off0[q + 1] - off0[q];
        return off0[q] + ((prefix_len > 0 && len > prefix_len) ? prefix_len : len);   // src/filter_common.rs:222-226
    };
    uint32_t q0 = 0;
    for (const uint64_t bi : bad32) {
        const uint64_t B = a0 + 32 * bi;   // block [B, B + 32); rec_end is non-decreasing in q
        uint32_t lo = q0, hi = nr;         // first record with rec_end > B
        while (lo < hi) {
            const uint32_t mid = lo + (hi - lo) / 2;
            if (rec_end(mid) > B) hi = mid; else lo = mid + 1;
        }
        q0 = lo;
        for (uint32_t q = lo; q < nr; q++) {
            const uint64_t e = rec_end(q);
            if (e > B + 32) break;
            if (off0[q + 1] - off0[q] < (uint64_t)k) continue;                               // :217-219
            if (bases[e - 1] == (uint8_t)'\n') nl[q / 32] |= 1u << (q % 32);                // :229
        }
    }
}

// ------------------------------------------------------------------ equal-length test of a chunk's records
#ifdef DCN_X86
__attribute__((target("avx2"))) static uint64_t first_odd_avx2(const uint64_t *off, uint64_t n, uint64_t len0) {
    const __m256i want = _mm256_set1_epi64x((long long)len0);
    uint64_t r = 0;
    for (; r + 16 <= n; r += 16) {   // 16 records per step; left at the first step that holds an odd one
        __m256i acc = _mm256_setzero_si256();
        for (int j = 0; j < 16; j += 4) {
            const __m256i a = _mm256_loadu_si256(reinterpret_cast<const __m256i *>(off + r + j));
            const __m256i b = _mm256_loadu_si256(reinterpret_cast<const __m256i *>(off + r + j + 1));
            acc = _mm256_or_si256(acc, _mm256_xor_si256(_mm256_sub_epi64(b, a), want));
        }
        if (!_mm256_testz_si256(acc, acc)) break;
    }
    return r;
}
#endif

bool offsets_equal_length(const uint64_t *off, uint64_t n, uint64_t len0) {
    uint64_t r = 0;
#ifdef DCN_X86
    if (simd_level() >= 1) r = first_odd_avx2(off, n, len0);
#endif
    for (; r < n; r++)
        if (off[r + 1] - off[r] != len0) return false;
    return true;
}

bool pack_has_simd() {
#ifdef DCN_X86
    return simd_level() > 0;
#else
    return false;
#endif
}

void pack_ascii(const uint8_t *bases, uint64_t n, uint32_t *codes, uint16_t *inv, int simd, std::vector<uint64_t> *bad32,
                std::vector<uint32_t> *bad_mask) {
    const uint64_t full = n / 32;
    uint64_t done = 0;
#ifdef DCN_X86
    // simd: 1 = best available, 2 = force the AVX2 path (tests), 0 = scalar
    if (simd == 1 && simd_level() == 2) {
        const bool nt = ((reinterpret_cast<uintptr_t>(codes) & 15u) == 0) && ((reinterpret_cast<uintptr_t>(inv) & 7u) == 0);
        pack_blocks_avx512(bases, n / 64, codes, inv, nt, bad32, bad_mask);
        done = (n / 64) * 2;
    } else if (simd && simd_level() >= 1) {
        pack_blocks_avx2(bases, full, codes, inv, bad32, bad_mask);
        done = full;
    }
#else
    (void)simd;
#endif
    uint16_t m2[2];
    for (uint64_t i = done; i < full; i++) {
        pack32_scalar(bases + 32 * i, codes + 2 * i, m2);
        if (inv) { inv[2 * i] = m2[0]; inv[2 * i + 1] = m2[1]; }
        if (bad32 && (m2[0] | m2[1])) { bad32->push_back(i); if (bad_mask) bad_mask->push_back((uint32_t)m2[0] | ((uint32_t)m2[1] << 16)); }
    }
    if (n % 32) {
        uint8_t tail[32];
        memset(tail, 0, sizeof(tail));   // byte 0: code 0, not ACGT
        memcpy(tail, bases + 32 * full, n % 32);
        pack32_scalar(tail, codes + 2 * full, m2);
        if (inv) { inv[2 * full] = m2[0]; inv[2 * full + 1] = m2[1]; }
        if (bad32) { bad32->push_back(full); if (bad_mask) bad_mask->push_back((uint32_t)m2[0] | ((uint32_t)m2[1] << 16)); }   // the padding is flagged non-ACGT: always listed
    }
}

}  // namespace dcn
