// dcn_core.cuh -- arithmetic primitives of the filter hot path, usable from device code and
// from the host-side emulation harness (tests only).  Reference semantics: SURVEY.md Appendix A;
// src/filter_common.rs:211-310 (filter flavour), src/minimizers.rs:125-191 (index flavour).
#pragma once
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define DCN_HD __host__ __device__ __forceinline__
#else
#define DCN_HD inline
#endif

namespace dcn {

// ------------------------------------------------------------------ bit helpers
DCN_HD uint32_t popc32(uint32_t x) {
#ifdef __CUDA_ARCH__
    return (uint32_t)__popc(x);
#else
    return (uint32_t)__builtin_popcount(x);
#endif
}
DCN_HD uint32_t brev32(uint32_t x) {
#ifdef __CUDA_ARCH__
    return __brev(x);
#else
    x = ((x >> 1) & 0x55555555u) | ((x & 0x55555555u) << 1);
    x = ((x >> 2) & 0x33333333u) | ((x & 0x33333333u) << 2);
    x = ((x >> 4) & 0x0F0F0F0Fu) | ((x & 0x0F0F0F0Fu) << 4);
    return __builtin_bswap32(x);
#endif
}
// low 32 bits of (hi:lo) >> s, s in [0,31]
DCN_HD uint32_t fshr(uint32_t lo, uint32_t hi, uint32_t s) {
#ifdef __CUDA_ARCH__
    return __funnelshift_r(lo, hi, s);
#else
    s &= 31u;
    return s ? (lo >> s) | (hi << (32 - s)) : lo;
#endif
}
DCN_HD uint32_t rotl32(uint32_t x, uint32_t r) {
#ifdef __CUDA_ARCH__
    return __funnelshift_l(x, x, r);
#else
    r &= 31u;
    return r ? (x << r) | (x >> (32 - r)) : x;
#endif
}
DCN_HD uint32_t rotr32(uint32_t x, uint32_t r) {
#ifdef __CUDA_ARCH__
    return __funnelshift_r(x, x, r);
#else
    r &= 31u;
    return r ? (x >> r) | (x << (32 - r)) : x;
#endif
}
DCN_HD uint32_t rot16(uint32_t x) {
#ifdef __CUDA_ARCH__
    return __byte_perm(x, 0, 0x1032);
#else
    return (x << 16) | (x >> 16);
#endif
}
DCN_HD uint64_t mulhi64(uint64_t a, uint64_t b) {
#ifdef __CUDA_ARCH__
    return __umul64hi(a, b);
#else
    return (uint64_t)(((unsigned __int128)a * b) >> 64);
#endif
}
DCN_HD uint64_t rotl64(uint64_t x, int r) { return (x << r) | (x >> (64 - r)); }
DCN_HD uint64_t bswap64(uint64_t x) {
#ifdef __CUDA_ARCH__
    uint32_t lo = (uint32_t)x, hi = (uint32_t)(x >> 32);
    return ((uint64_t)__byte_perm(lo, 0, 0x0123) << 32) | __byte_perm(hi, 0, 0x0123);
#else
    return __builtin_bswap64(x);
#endif
}

// ------------------------------------------------------------------ xxh3
// xxhash-rust xxh3_64(&v.to_le_bytes()), seed 0: XXH3_len_4to8_64b + rrmxmx.
DCN_HD uint64_t xxh3_u64(uint64_t v) {
    uint64_t x = (v >> 32) | (v << 32);
    x ^= 0xC73AB174C5ECD5A2ULL;
    x ^= rotl64(x, 49) ^ rotl64(x, 24);
    x *= 0x9FB21C651E98DF25ULL;
    x ^= (x >> 35) + 8;
    x *= 0x9FB21C651E98DF25ULL;
    return x ^ (x >> 28);
}
// 16-byte input (k > 32): XXH3_len_9to16_64b + avalanche.
DCN_HD uint64_t xxh3_u128(uint64_t lo, uint64_t hi) {
    uint64_t a = lo ^ 0x6782737BEA4239B9ULL, b = hi ^ 0xAF56BC3B0996523AULL;
    uint64_t acc = 16 + bswap64(a) + b + ((a * b) ^ mulhi64(a, b));
    acc ^= acc >> 37;
    acc *= 0x165667919E3779F9ULL;
    return acc ^ (acc >> 32);
}

// ------------------------------------------------------------------ ntHash seeds
// indexed by the 2-bit code A=0 C=1 T=2 G=3; complement = code ^ 2
DCN_HD uint32_t nt_f(uint32_t c) {
    return c == 0 ? 0x95c60474u : c == 1 ? 0x62a02b4cu : c == 2 ? 0x82572324u : 0x4be24456u;
}

// ------------------------------------------------------------------ canonical k-mer (A.4)
// v holds k bases, base i at bits 2i (k <= 32).  Reverse complement in the same layout.
DCN_HD uint64_t revcomp_2bit(uint64_t v, int k) {
    uint32_t lo = (uint32_t)v, hi = (uint32_t)(v >> 32);
    uint32_t rl = brev32(hi), rh = brev32(lo);              // bit-reverse 64
    rl = ((rl & 0x55555555u) << 1) | ((rl >> 1) & 0x55555555u);  // restore bit order inside pairs
    rh = ((rh & 0x55555555u) << 1) | ((rh >> 1) & 0x55555555u);
    uint64_t r = ((uint64_t)rh << 32) | rl;                 // base i now at pair 31-i
    r ^= 0xAAAAAAAAAAAAAAAAULL;                             // complement: code ^ 2
    return r >> (64 - 2 * k);                               // base i at pair k-1-i
}

// ------------------------------------------------------------------ classification (A.6)
// src/filter_common.rs:84-112.  round() = half away from zero; `as usize` saturates, NaN -> 0.
DCN_HD uint64_t required_hits(uint32_t abs_thr, double rel_thr, uint64_t total) {
    uint64_t rel = 0;
    if (total != 0) {
        double r = ::round(rel_thr * (double)total);
        if (!(r > 0.0)) rel = 0;
        else if (r >= 18446744073709551616.0) rel = ~0ULL;
        else rel = (uint64_t)r;
        if (rel < 1) rel = 1;
    }
    return (uint64_t)abs_thr > rel ? (uint64_t)abs_thr : rel;
}
DCN_HD bool meets_criteria(uint64_t hits, uint64_t total, uint32_t abs_thr, double rel_thr, int deplete) {
    uint64_t req = required_hits(abs_thr, rel_thr, total);
    return deplete ? hits < req : hits >= req;
}

// ------------------------------------------------------------------ HBM index table
// Open addressing, 32-byte buckets of four u64 keys (one DRAM sector per probe), linear probing
// over buckets.  bucket(h) = mulhi(h, n_buckets): keys are xxh3 outputs, already uniform.
// Replaces FxHashSet<u64> (src/index.rs:98-105; probed at src/filter_common.rs:144,185).
static constexpr uint64_t DCN_EMPTY = 0xFFFFFFFFFFFFFFFFULL;
static constexpr int DCN_BUCKET = 4;

struct TableView {
    const uint64_t *slots;   // n_buckets * 4
    uint64_t n_buckets;
    int has_empty_key;       // DCN_EMPTY itself is a member of the set
};

DCN_HD uint64_t table_bucket(uint64_t h, uint64_t n_buckets) { return mulhi64(h, n_buckets); }

}  // namespace dcn
