// dcn_generic.cuh -- any (k, w): one thread per chunk of windows, sequential over the chunk.
//
// The fused tile kernel is specialised for the default index parameters (k = 31, w = 15).  Indexes
// built with other parameters (k up to 56 on the filter side, src/filter_common.rs:269-272; u128
// k-mer values above 32, :289-297; k up to 57 at index time, src/main.rs:166) and the materialising
// extraction entry point (B3, dcn_extract) take this path: same arithmetic,
// no shared-memory staging.  A record is cut into chunks of `cstride` windows; a chunk recomputes
// the window before its first one so the consecutive-duplicate rule (A.3 step 5) sees its
// predecessor.  Output order inside a record is window order, so per-chunk counts + an exclusive
// scan give the CSR layout the reference returns (Vec<u64> hashes, Vec<u32> positions).
#pragma once
#include "dcn_tile.cuh"

namespace dcn {

static constexpr uint32_t DCN_GENERIC_CSTRIDE = 256;   // windows per chunk
static constexpr int DCN_MAX_W = 255;                   // IndexHeader.window_size is a u8 (src/index.rs:17-22)

struct U128 { uint64_t lo, hi; };
DCN_HD bool u128_less(U128 a, U128 b) { return a.hi < b.hi || (a.hi == b.hi && a.lo < b.lo); }

DCN_HD bool generic_is_acgt(uint8_t b) {
    uint32_t u = b & 0xDFu;
    return u == 'A' || u == 'C' || u == 'G' || u == 'T';
}

template <int FLAV>
DCN_HD uint32_t generic_code(uint8_t b) {
    return FLAV == FLAVOUR_INDEX ? code_index_flavour(b) : ((uint32_t)b >> 1) & 3u;
}

// canonical k-mer value at seq[p .. p+k) and its xxh3 (A.4, A.5).  Only all-ACGT k-mers get here,
// for which the lossy code (byte >> 1) & 3 is the code in both flavours.
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

// src/minimizers.rs:73-121 as a table over base counts, 64 values per axis (k <= 57)
DCN_HD bool generic_entropy_ok(const uint32_t *pass, const uint8_t *seq, uint64_t p, int k) {
    if (!pass) return true;
    uint32_t nA = 0, nC = 0, nG = 0;
    for (int i = 0; i < k; i++) {
        uint32_t c = (seq[p + i] >> 1) & 3u;
        nA += c == 0; nC += c == 1; nG += c == 3;
    }
    uint32_t idx = (nA * 64u + nC) * 64u + nG;
    return (pass[idx >> 5] >> (idx & 31u)) & 1u;
}

// effective length of a record (src/filter_common.rs:217-229 / src/minimizers.rs:135-137)
template <int FLAV>
DCN_HD uint64_t generic_eff_len(const uint8_t *bases, uint64_t gstart, uint64_t len, uint32_t prefix_len, int k) {
    if (len < (uint64_t)k) return 0;
    if (FLAV == FLAVOUR_INDEX) return len;
    uint64_t n = (prefix_len > 0 && len > prefix_len) ? prefix_len : len;
    if (n > 0 && bases[gstart + n - 1] == (uint8_t)'\n') n--;
    return n;
}

DCN_HD uint64_t generic_chunks_of(uint64_t eff_len, int k, int w, uint32_t cstride) {
    const uint64_t l = (uint64_t)(k + w - 1);
    if (eff_len < l) return 0;
    return (eff_len - l + 1 + cstride - 1) / cstride;
}

// Calls emit(position, hash) for every minimizer that survives the ACGT (and entropy) filter among
// the windows [w_begin, w_end) of the effective sequence seq[0 .. n), in window order.
template <int FLAV, class Emit>
DCN_HD void generic_span(const uint8_t *seq, uint64_t n, int k, int w, uint64_t w_begin, uint64_t w_end,
                         const uint32_t *entropy_pass, Emit emit) {
    const uint32_t l = (uint32_t)(k + w - 1);
    if (n < l || w_begin >= w_end) return;
    uint16_t ring[DCN_MAX_W + 1];   // ntHash keys (upper 16 bits) of the last w k-mers
    const uint64_t j0 = w_begin > 0 ? w_begin - 1 : 0;   // carry window: computed, never emitted
    uint32_t fw = 0, rc = 0;
    for (int i = 0; i < k; i++) {
        uint32_t c = generic_code<FLAV>(seq[j0 + i]);
        fw ^= rotl32(nt_f(c), (uint32_t)(k - 1 - i));
        rc ^= rotl32(nt_f(c ^ 2u), (uint32_t)i);
    }
    uint32_t tg = 0;
    for (uint32_t i = 0; i < l; i++) tg += (generic_code<FLAV>(seq[j0 + i]) >> 1) & 1u;   // T/G <=> bit 1 of the code
    uint64_t prev = ~0ULL;
    const uint64_t p_end = w_end + (uint64_t)w - 1;   // one past the last k-mer start needed
    for (uint64_t p = j0; p < p_end; p++) {
        if (p > j0) {
            uint32_t oc = generic_code<FLAV>(seq[p - 1]), ic = generic_code<FLAV>(seq[p + k - 1]);
            fw = rotl32(fw, 1) ^ rotl32(nt_f(oc), (uint32_t)k) ^ nt_f(ic);
            rc = rotr32(rc ^ nt_f(oc ^ 2u) ^ rotl32(nt_f(ic ^ 2u), (uint32_t)k), 1);
        }
        ring[p % (uint64_t)w] = (uint16_t)((fw + rc) >> 16);
        if (p + 1 < j0 + (uint64_t)w) continue;
        const uint64_t j = p + 1 - (uint64_t)w;                         // window start
        if (j > j0) tg += ((generic_code<FLAV>(seq[j + l - 1]) >> 1) & 1u) - ((generic_code<FLAV>(seq[j - 1]) >> 1) & 1u);
        uint64_t left = j, right = j;
        uint32_t kl = 0x10000u, kr = 0x10000u;
        for (uint64_t q = j; q <= p; q++) {
            uint32_t key = ring[q % (uint64_t)w];
            if (key < kl) { kl = key; left = q; }
            if (key <= kr) { kr = key; right = q; }
        }
        const uint64_t pick = (2 * tg > l) ? left : right;               // A.3 step 4
        const bool dup = j > 0 && pick == prev;                          // A.3 step 5
        prev = pick;
        if (dup || j < w_begin) continue;
        bool ok = true;
        for (int i = 0; i < k && ok; i++) ok = generic_is_acgt(seq[pick + i]);   // src/filter_common.rs:275-286
        if (!ok) continue;
        if (FLAV == FLAVOUR_INDEX && !generic_entropy_ok(entropy_pass, seq, pick, k)) continue;
        emit(pick, generic_kmer_hash(seq, pick, k));
    }
}

// chunk g of the batch -> (record, first window, one past last window); rec_chunk_off is the
// exclusive scan of the per-record chunk counts (n_rec + 1 entries)
DCN_HD uint32_t generic_find_record(const uint64_t *rec_chunk_off, uint32_t n_rec, uint64_t g) {
    uint32_t lo = 0, hi = n_rec;   // largest r with rec_chunk_off[r] <= g
    while (hi - lo > 1) {
        uint32_t mid = lo + (hi - lo) / 2;
        if (rec_chunk_off[mid] <= g) lo = mid; else hi = mid;
    }
    return lo;
}

struct GenericBatch {
    const uint8_t *bases;       // bases[x - base0] readable for base0 <= x < rec_off[n_rec]
    uint64_t base0;
    const uint64_t *rec_off;    // absolute offsets, n_rec + 1
    uint32_t n_rec;
    uint32_t prefix_len;
    int k, w;
    uint32_t cstride;
    const uint32_t *entropy_pass;
    const uint64_t *rec_chunk_off;
};

template <int FLAV, class Emit>
DCN_HD void generic_chunk(const GenericBatch &B, uint64_t g, uint32_t &rec_out, Emit emit) {
    const uint32_t r = generic_find_record(B.rec_chunk_off, B.n_rec, g);
    rec_out = r;
    const uint64_t c = g - B.rec_chunk_off[r];
    const uint64_t gs = B.rec_off[r] - B.base0;
    const uint64_t len = B.rec_off[r + 1] - B.base0 - gs;
    const uint64_t n = generic_eff_len<FLAV>(B.bases, gs, len, B.prefix_len, B.k);
    const uint64_t nwin = n - (uint64_t)(B.k + B.w - 1) + 1;   // n >= l: the record has >= 1 chunk
    const uint64_t wb = c * B.cstride;
    const uint64_t we = wb + B.cstride < nwin ? wb + B.cstride : nwin;
    generic_span<FLAV>(B.bases + gs, n, B.k, B.w, wb, we, B.entropy_pass, emit);
}

#ifdef __CUDACC__
// chunks per record (input of the exclusive scan)
template <int FLAV>
__global__ void generic_rec_chunks_kernel(GenericBatch B, uint64_t *rec_chunks) {
    for (uint32_t r = blockIdx.x * blockDim.x + threadIdx.x; r <= B.n_rec; r += gridDim.x * blockDim.x) {
        uint64_t nc = 0;
        if (r < B.n_rec) {
            const uint64_t gs = B.rec_off[r] - B.base0, len = B.rec_off[r + 1] - B.base0 - gs;
            nc = generic_chunks_of(generic_eff_len<FLAV>(B.bases, gs, len, B.prefix_len, B.k), B.k, B.w, B.cstride);
        }
        rec_chunks[r] = nc;
    }
}

// B3 pass 1: surviving minimizers per chunk
template <int FLAV>
__global__ void __launch_bounds__(128)
generic_count_kernel(GenericBatch B, uint64_t n_chunks, uint64_t *chunk_count) {
    for (uint64_t g = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; g <= n_chunks; g += (uint64_t)gridDim.x * blockDim.x) {
        uint64_t n = 0;
        uint32_t r;
        if (g < n_chunks) generic_chunk<FLAV>(B, g, r, [&](uint64_t, uint64_t) { n++; });
        chunk_count[g] = n;
    }
}

// B3 pass 2: write hashes / positions at the scanned offsets
template <int FLAV>
__global__ void __launch_bounds__(128)
generic_write_kernel(GenericBatch B, uint64_t n_chunks, const uint64_t *__restrict__ chunk_off, uint64_t *out_hashes,
                     uint32_t *out_pos) {
    for (uint64_t g = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; g < n_chunks; g += (uint64_t)gridDim.x * blockDim.x) {
        uint64_t at = chunk_off[g];
        uint32_t r;
        generic_chunk<FLAV>(B, g, r, [&](uint64_t pos, uint64_t h) {
            out_hashes[at] = h;
            if (out_pos) out_pos[at] = (uint32_t)pos;
            at++;
        });
    }
}

// CSR offsets of the records: out_off[r] = chunk_off[rec_chunk_off[r]]
__global__ void generic_rec_off_kernel(const uint64_t *__restrict__ rec_chunk_off, const uint64_t *__restrict__ chunk_off,
                                       uint32_t n_rec, uint64_t *out_off) {
    for (uint32_t r = blockIdx.x * blockDim.x + threadIdx.x; r <= n_rec; r += gridDim.x * blockDim.x)
        out_off[r] = chunk_off[rec_chunk_off[r]];
}

// B1 for any (k, w): extraction + lookup + distinct count per chunk, totals added per unit
__global__ void __launch_bounds__(128)
generic_filter_kernel(GenericBatch B, uint64_t n_chunks, uint32_t rpu, TableView table, DedupView dd, uint32_t *hits,
                      uint32_t *total) {
    for (uint64_t g = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; g < n_chunks; g += (uint64_t)gridDim.x * blockDim.x) {
        uint32_t nt = 0, nh = 0, r = 0;
        // the unit is known only after the record lookup inside generic_chunk: find it first
        const uint32_t unit = generic_find_record(B.rec_chunk_off, B.n_rec, g) / rpu;
        generic_chunk<FLAVOUR_FILTER>(B, g, r, [&](uint64_t, uint64_t h) {
            nt++;
            if (table_contains(table, h) && dedup_insert(dd, h, unit)) nh++;
        });
        if (nt) atomicAdd(&total[unit], nt);
        if (nh) atomicAdd(&hits[unit], nh);
    }
}

// keep flag of every unit once all chunks have added their counts
__global__ void generic_finalize_kernel(uint32_t n_units, const uint32_t *__restrict__ hits, const uint32_t *__restrict__ total,
                                        uint32_t abs_thr, double rel_thr, int deplete, uint8_t *keep) {
    for (uint32_t u = blockIdx.x * blockDim.x + threadIdx.x; u < n_units; u += gridDim.x * blockDim.x)
        keep[u] = meets_criteria(hits[u], total[u], abs_thr, rel_thr, deplete) ? 1 : 0;
}
#endif

}  // namespace dcn
