// dcn_generic.cuh -- any (k, w): one thread per unit, sequential over the unit's bases.
//
// The fused tile kernel is specialised for the default index parameters (k = 31, w = 15).  Indexes
// built with other parameters (k up to 56 on the filter side, src/filter_common.rs:269-272; u128
// k-mer values above 32, :289-297) take this path: same arithmetic, no
// shared-memory staging, distinct hits through the global (hash, unit) set.  It is a correctness
// path, not a fast one.
#pragma once
#include "dcn_tile.cuh"

namespace dcn {

struct U128 { uint64_t lo, hi; };
DCN_HD bool u128_less(U128 a, U128 b) { return a.hi < b.hi || (a.hi == b.hi && a.lo < b.lo); }

DCN_HD bool generic_is_acgt(uint8_t b) {
    uint32_t u = b & 0xDFu;
    return u == 'A' || u == 'C' || u == 'G' || u == 'T';
}

// canonical k-mer value at seq[p .. p+k) (codes = (byte >> 1) & 3) and its xxh3
DCN_HD uint64_t generic_kmer_hash(const uint8_t *seq, uint64_t p, int k) {
    U128 fw = {0, 0}, rc = {0, 0};
    for (int i = 0; i < k; i++) {
        uint64_t c = (seq[p + i] >> 1) & 3u;
        uint64_t d = ((seq[p + k - 1 - i] >> 1) & 3u) ^ 2u;
        if (i < 32) { fw.lo |= c << (2 * i); rc.lo |= d << (2 * i); }
        else { fw.hi |= c << (2 * (i - 32)); rc.hi |= d << (2 * (i - 32)); }
    }
    U128 v = u128_less(fw, rc) ? fw : rc;
    return k <= 32 ? xxh3_u64(v.lo) : xxh3_u128(v.lo, v.hi);
}

// filter-flavour extraction + lookup + distinct count for one unit; returns (hits, total)
DCN_HD void generic_unit(const FilterParams &P, int k, int w, const DedupView &dd, uint32_t u, uint32_t &hits_out,
                         uint32_t &total_out) {
    uint32_t hits = 0, total = 0;
    const uint32_t l = (uint32_t)(k + w - 1);
    uint16_t ring[256];   // ntHash keys (upper 16 bits) of the last w k-mers
    for (uint32_t r = u * P.rpu; r < (u + 1) * P.rpu; r++) {
        const uint64_t gs = P.rec_off[r] - P.base0;
        const uint64_t len = P.rec_off[r + 1] - P.base0 - gs;
        if (len < (uint64_t)k) continue;                                   // src/filter_common.rs:217-219
        uint64_t n = (P.prefix_len > 0 && len > P.prefix_len) ? P.prefix_len : len;   // :222-226
        const uint8_t *seq = P.bases + gs;
        if (n > 0 && seq[n - 1] == (uint8_t)'\n') n--;                      // :229
        if (n < l) continue;
        uint32_t fw = 0, rc = 0;
        for (int i = 0; i < k; i++) {
            uint32_t c = (seq[i] >> 1) & 3u;
            fw ^= rotl32(nt_f(c), (uint32_t)(k - 1 - i));
            rc ^= rotl32(nt_f(c ^ 2u), (uint32_t)i);
        }
        uint32_t tg = 0;
        for (uint32_t i = 0; i < l; i++) tg += (seq[i] >> 2) & 1u;          // T/G <=> bit 1 of the code
        uint64_t prev = ~0ULL;
        const uint64_t nk = n - (uint64_t)k + 1;
        for (uint64_t p = 0; p < nk; p++) {
            if (p > 0) {
                uint32_t oc = (seq[p - 1] >> 1) & 3u, ic = (seq[p + k - 1] >> 1) & 3u;
                fw = rotl32(fw, 1) ^ rotl32(nt_f(oc), (uint32_t)k) ^ nt_f(ic);
                rc = rotr32(rc ^ nt_f(oc ^ 2u) ^ rotl32(nt_f(ic ^ 2u), (uint32_t)k), 1);
            }
            ring[p % (uint64_t)w] = (uint16_t)((fw + rc) >> 16);
            if (p + 1 < (uint64_t)w) continue;
            const uint64_t j = p + 1 - (uint64_t)w;                         // window start
            if (j > 0) tg += ((seq[j + l - 1] >> 2) & 1u) - ((seq[j - 1] >> 2) & 1u);
            uint64_t left = j, right = j;
            uint32_t kl = 0x10000u, kr = 0x10000u;
            for (uint64_t q = j; q <= p; q++) {
                uint32_t key = ring[q % (uint64_t)w];
                if (key < kl) { kl = key; left = q; }
                if (key <= kr) { kr = key; right = q; }
            }
            const uint64_t pick = (2 * tg > l) ? left : right;
            if (j > 0 && pick == prev) continue;
            prev = pick;
            bool ok = true;
            for (int i = 0; i < k && ok; i++) ok = generic_is_acgt(seq[pick + i]);   // :275-286
            if (!ok) continue;
            total++;
            uint64_t h = generic_kmer_hash(seq, pick, k);
            if (table_contains(P.table, h) && dedup_insert(dd, h, u)) hits++;
        }
    }
    hits_out = hits;
    total_out = total;
}

#ifdef __CUDACC__
__global__ void __launch_bounds__(128)
filter_generic_kernel(FilterParams P, int k, int w, DedupView dd) {
    for (uint32_t u = blockIdx.x * blockDim.x + threadIdx.x; u < P.n_units; u += gridDim.x * blockDim.x) {
        uint32_t hits, total;
        generic_unit(P, k, w, dd, u, hits, total);
        P.hits[u] = hits;
        P.total[u] = total;
        P.keep[u] = meets_criteria(hits, total, P.abs_thr, P.rel_thr, P.deplete) ? 1 : 0;
    }
}
#endif

}  // namespace dcn
