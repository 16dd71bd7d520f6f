// dcn_warp.cuh -- the warp-autonomous short-read path of the fused extract -> lookup -> classify kernel.
//
// One WARP owns a tile: a 16-byte-aligned span of 1536 bases of the concatenated base stream that holds whole
// units (a record, or a pair of mates).  Lane l owns the 48 consecutive bases [48 l, 48 l + 48) of the span, as
// three blocks of 16: it converts them (2-bit codes + non-ACGT mask), rolls the canonical ntHash over its 48 k-mer
// starts without re-seeding, and picks the minimizer of its 48 windows (van Herk / Gil-Werman over w = 15, 16
// windows per block).  Only two things cross a lane boundary: the code words of the next lane (k-mers and strand
// counts reach 44 bases ahead) and its first 14 hashes; both go through the warp's own slice of shared memory.
// Picks are compacted into a per-warp list, hashed (xxh3 of the canonical 2-bit k-mer), probed in the HBM table,
// and counted per unit by the same warp.  Nothing in a tile is shared with another warp, so the phases are joined
// by __syncwarp() only: no CTA barrier, no block-wide scan, and 32 warps per SM run 32 tiles in 32 different phases.
// The tile's bytes arrive by one TMA bulk copy (cp.async.bulk + mbarrier) issued by the warp itself one tile ahead.
//
// Reference semantics: src/filter_common.rs:211-310 (extraction), :129-198 (distinct hits),
// :84-112 (thresholds).  Same arithmetic as dcn_tile.cuh, another mapping of the work onto threads; the CTA-tile
// code stays for what a warp tile cannot hold (units of more than PKCAP picks, long units, index builds).
#pragma once
#include <stddef.h>

#include "dcn_tile.cuh"

namespace dcn {

#ifndef DCN_PICKS_IN_FLIGHT
#define DCN_PICKS_IN_FLIGHT 1   // probes a lane has outstanding in P6 (measured: 1 -> 298.9, 2 -> 294.6, 3 -> 268.6, 4 -> 252.5 Gbp/s)
#endif
#ifndef DCN_PICKS_LAST_3
#define DCN_PICKS_LAST_3 0      // A/B knob: one more pick per lane in the iteration that then finishes the tile's list (measured: slower)
#endif
#ifndef DCN_LONG_PICKS_IN_FLIGHT
#define DCN_LONG_PICKS_IN_FLIGHT 2   // the same in the probe loop of a long unit's chunk
#endif

struct WG {
    static constexpr int K = 31, W = 15, L = 45;
    static constexpr int NL = 32;              // lanes
    static constexpr int LB = 48;              // bases = k-mer starts = window starts per lane
    static constexpr int NBLK = 3;             // blocks of 16 per lane
    static constexpr int TB = NL * LB;         // 1536 bases a tile may reference from its origin
    static constexpr int NSLOT = TB / 16;      // 96 blocks
    static constexpr int MAXR = 64;            // records per tile
    static constexpr int PKCAP = 320;          // picks per pass; a denser run of units is split, a denser single unit goes to the CTA path
    static constexpr int NBW = TB / 32;        // words of the per-position bit arrays
    static constexpr int HXP = 8;              // pitch (words) of the rows that hand a lane's first hashes to its left neighbour
};

// tile descriptor written by the planner: units [a, b) live in [origin, origin + TB) (origin relative to base0).
// A chunk of a LONG unit (more than DCN_MAX_SHORT bases) is a tile too: bit 63 of origin set, a = record,
// b = la | carry << 4 | nw << 5 (first computed window, tile-local; carry window present; own windows).
struct alignas(16) WTile { uint64_t origin; uint32_t a, b; };
static constexpr uint64_t WTILE_LONG = 1ull << 63;
// own windows of a long-unit chunk: 15 (alignment) + 1 (carry window) + WCS + L - 1 <= TB
static constexpr uint32_t DCN_WCS = 1472;

struct WarpTables {                    // one per CTA
    u32x2 tb0[256];                    // 4-base ntHash aggregates (fw, rc), as TileSmem::tb0
    u32x2 tio[16];                     // rolling step indexed by outgoing | incoming << 2
    uint16_t req[256];                 // required_hits(total), total < 256
};

struct alignas(16) u32x4a { uint32_t x, y, z, w; };   // 16-byte aligned: one LDS.128 / STS.128

// Shared memory of one warp.  The first region is used three ways in turn:
//   hash / slice phases:  hx (each lane's first 14 hashes for its left neighbour) | per-block results rel4, em
//   compaction:           pk_pos over hx (dead)                                    | rel4, em still read
//   probe / count:        pk_pos | pk_hash over the rest (rel4 / em dead)
struct alignas(16) WarpSmem {
    static constexpr int HX_BYTES = (WG::NL + 1) * WG::HXP * 4;   // 1056
    static constexpr int POS_BYTES = WG::PKCAP * 2;               // 640 <= HX_BYTES
    static constexpr int REGION = POS_BYTES + WG::PKCAP * 8;      // 3200
    alignas(16) unsigned char region[REGION];
    alignas(16) uint8_t stage[WG::TB + 16];   // ASCII bytes of the tile (bulk-copy destination)
    uint32_t codes[WG::NSLOT + 4];     // 16 bases x 2 bit per word
    uint32_t inv[WG::NBW + 2];         // non-ACGT bits
    uint32_t brk[WG::NBW + 2];         // position is a record start or lies outside every effective sequence
    uint32_t dead[WG::NBW + 2];        // no window may start here
    uint16_t ufirst[WG::MAXR + 2];     // first pick index of each unit
    uint16_t ustartpos[WG::MAXR + 2];
    uint16_t lastpick[WG::NL];         // last window's pick of each lane (0xFFFF: that window is invalid)
    uint32_t npicks, nhits;            // nhits: long chunks, entries of the compacted hit list
    unsigned long long mbar;           // transaction barrier of the bulk copy

    DCN_HD uint32_t *hx() { return reinterpret_cast<uint32_t *>(region); }
    DCN_HD uint16_t *pk_pos() { return reinterpret_cast<uint16_t *>(region); }
    DCN_HD const uint16_t *pk_pos() const { return reinterpret_cast<const uint16_t *>(region); }
    DCN_HD uint64_t *pk_hash() { return reinterpret_cast<uint64_t *>(region + POS_BYTES); }
    DCN_HD u32x4a *rel4() { return reinterpret_cast<u32x4a *>(region + HX_BYTES); }                       // per block: pick positions, 8 bits per window
    DCN_HD uint32_t *em() { return reinterpret_cast<uint32_t *>(region + HX_BYTES + WG::NSLOT * 16); }    // per block: emit mask | exclusive pick offset << 16
};
static_assert(WarpSmem::POS_BYTES <= WarpSmem::HX_BYTES, "the pick positions replace the hash rows only");
static_assert(WarpSmem::HX_BYTES + WG::NSLOT * 20 <= WarpSmem::REGION, "block results fit beside the hash rows");

struct WarpPriv {
    uint32_t sL[2], eL[2], eff[2];   // up to two records of the tile per lane (tile-local start, end, effective end)
    uint32_t fw, rc;          // rolling state: k-mer at the start of block 1 after the seed phase
    uint32_t hp[8];           // block 0 hashes (upper 16 bits), two per word
    uint32_t vfirst;          // the lane's first window is valid
    uint32_t pickoff;
    u32x4a lrel[3];           // long chunks only: the lane's block results, read before the pick list overwrites them
    uint32_t lem[3];
};

// ------------------------------------------------------------------ tables
DCN_HD void winit_tables(int t, int nt, WarpTables &T, uint32_t abs_thr, double rel_thr) {
    for (int b = t; b < 256; b += nt) {
        uint32_t fw = 0, rc = 0;
        for (int m = 0; m < 4; m++) {
            uint32_t c = ((uint32_t)b >> (2 * m)) & 3u;
            fw ^= rotl32(nt_f(c), (uint32_t)(30 - m));
            rc ^= rotl32(nt_f(c ^ 2u), (uint32_t)m);
        }
        T.tb0[b].x = fw; T.tb0[b].y = rc;
        const uint64_t r = required_hits(abs_thr, rel_thr, (uint64_t)b);
        T.req[b] = (uint16_t)(r > 0xFFFFull ? 0xFFFFull : r);
    }
    if (t < 16) {
        uint32_t oc = (uint32_t)t & 3u, ic = (uint32_t)t >> 2;
        T.tio[t].x = rotl32(nt_f(oc), WG::K) ^ nt_f(ic);
        T.tio[t].y = nt_f(oc ^ 2u) ^ rotl32(nt_f(ic ^ 2u), WG::K);
    }
}

// ------------------------------------------------------------------ convert
// four ASCII bytes -> four 2-bit codes (packed-seq lossy code (b >> 1) & 3, src/filter_common.rs:238) + non-ACGT bits
DCN_HD void convert4(uint32_t x, uint32_t &packed8, uint32_t &inv4) {
    const uint32_t c = (x >> 1) & 0x03030303u;
    const uint32_t c0b = c & 0x01010101u, c1b = (c >> 1) & 0x01010101u;
    const uint32_t e = 0x41414141u + c0b * 2u + c1b * 0x13u - (c0b & c1b) * 0xFu;   // the letter the code stands for
    const uint32_t d = (x & 0xDFDFDFDFu) ^ e;
    const uint32_t nz = (((d & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | d) & 0x80808080u;
    inv4 = (nz * 0x00204081u) >> 28;
    packed8 = (c * 0x01041040u) >> 24;
}

DCN_HD uint32_t pack_hi16(uint32_t even, uint32_t odd) {   // upper halves of two words -> one word
#ifdef __CUDA_ARCH__
    return __byte_perm(even, odd, 0x7632);
#else
    return (even >> 16) | (odd & 0xFFFF0000u);
#endif
}

// hashes of the 16 k-mers of one block, given the rolling state at its first k-mer; with CROSS the state
// is advanced to the first k-mer of the next block.  W0 = the block's codes, W1 / W2 = the next two words.
template <bool CROSS>
DCN_HD void whash_block(const WarpTables &T, uint32_t &fw, uint32_t &rc, uint32_t W0, uint32_t W1, uint32_t W2,
                        uint32_t (&hp)[8]) {
    const uint32_t in = fshr(W1, W2, 2u * (uint32_t)(WG::K - 16));        // base K + j at bits 2j
    const uint32_t mE = (W0 & 0x33333333u) | ((in & 0x33333333u) << 2);
    const uint32_t mO = ((W0 >> 2) & 0x33333333u) | (in & 0xCCCCCCCCu);
    uint32_t h[16];
    h[0] = fw + rc;
#pragma unroll
    for (int j = 0; j < 16; j++) {
        if (j == 15 && !CROSS) break;
        const uint32_t idx = (((j & 1) ? mO : mE) >> (4 * (j >> 1))) & 15u;
        const u32x2 e = T.tio[idx];
        fw = rotl32(fw, 1) ^ e.x;
        rc = rotr32(rc ^ e.y, 1);
        if (j < 15) h[j + 1] = fw + rc;
    }
#pragma unroll
    for (int j = 0; j < 8; j++) hp[j] = pack_hi16(h[2 * j], h[2 * j + 1]);
}

// ------------------------------------------------------------------ window minima of one block
// q[0..14]: upper 16 bits of the 30 hashes the block's 16 windows cover, two per word.  c0..c3: the code words of
// the 64 bases from the block start.  `slot` = block index in the tile (position / 16).
// Same selection as phase_slide.
DCN_HD void wslide_block(const uint32_t (&q)[15], uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                         const uint32_t *brk, const uint32_t *dead, int slot, uint32_t (&rel4)[4], uint32_t &emask,
                         uint32_t &valid16, uint32_t &rel15) {
    uint32_t key[30], oL[16], oR[16];
    // leftmost smallest: min over (h16 << 16 | i)
#pragma unroll
    for (int j = 0; j < 15; j++) {
        key[2 * j] = (q[j] << 16) | (uint32_t)(2 * j);
        key[2 * j + 1] = (q[j] & 0xFFFF0000u) | (uint32_t)(2 * j + 1);
    }
#pragma unroll
    for (int i = 13; i >= 0; i--) key[i] = umin32(key[i], key[i + 1]);
#pragma unroll
    for (int i = 16; i < 30; i++) key[i] = umin32(key[i], key[i - 1]);
    oL[0] = key[0];
#pragma unroll
    for (int i = 1; i < 15; i++) oL[i] = umin32(key[i], key[14 + i]);
    oL[15] = key[29];
    // rightmost smallest: min over (h16 << 16 | 31 - i); the low byte of the result is 31 - position
#pragma unroll
    for (int j = 0; j < 15; j++) {
        key[2 * j] = (q[j] << 16) | (uint32_t)(31 - 2 * j);
        key[2 * j + 1] = (q[j] & 0xFFFF0000u) | (uint32_t)(30 - 2 * j);
    }
#pragma unroll
    for (int i = 13; i >= 0; i--) key[i] = umin32(key[i], key[i + 1]);
#pragma unroll
    for (int i = 16; i < 30; i++) key[i] = umin32(key[i], key[i - 1]);
    oR[0] = key[0];
#pragma unroll
    for (int i = 1; i < 15; i++) oR[i] = umin32(key[i], key[14 + i]);
    oR[15] = key[29];

    // leftmost and rightmost smallest differ only where the window minimum is tied (~2e-4 of windows): the strand that
    // chooses between them (below) is only worked out for a block that holds such a window
    uint32_t l4[4], r4[4], tied = 0;
#pragma unroll
    for (int g = 0; g < 4; g++) {
        l4[g] = pack_low_bytes(oL[4 * g], oL[4 * g + 1], oL[4 * g + 2], oL[4 * g + 3]);
        r4[g] = pack_low_bytes(oR[4 * g], oR[4 * g + 1], oR[4 * g + 2], oR[4 * g + 3]) ^ 0x1F1F1F1Fu;
        tied |= l4[g] ^ r4[g];
    }
    if (tied) {
        // canonical strand: #(T|G) > #(A|C) over the L bases of the window; T/G <=> bit 1 of the code
        uint32_t cnt = popc32(c0 & 0xAAAAAAAAu) + popc32(c1 & 0xAAAAAAAAu) +
                       popc32(c2 & (0xAAAAAAAAu & ((1u << (2 * ((WG::L - 32) & 15))) - 1u)));
        constexpr int LW = (WG::L - 1) / 16, LS = 2 * ((WG::L - 1) % 16);
        static_assert(LW == 2, "L = 45");
        const uint32_t xw_lo = (c2 >> 1) & 0x55555555u, xw_hi = (c3 >> 1) & 0x55555555u;
        const uint32_t xin = fshr(xw_lo, xw_hi, (uint32_t)LS) & ~3u;        // lane j = T/G flag of base L - 1 + j
        const uint32_t xout = ((c0 >> 1) & 0x55555555u) << 2;               // lane j = flag of base j - 1
        constexpr uint32_t THR = (uint32_t)(WG::L + 1) / 2;
#pragma unroll
        for (int g = 0; g < 4; g++) {
            const uint32_t spi = (((xin >> (8 * g)) & 0xFFu) * 0x00041041u) & 0x01010101u;
            const uint32_t spo = (((xout >> (8 * g)) & 0xFFu) * 0x00041041u) & 0x01010101u;
            const uint32_t cnt4 = cnt * 0x01010101u + spi * 0x01010101u - spo * 0x01010101u;
            cnt = cnt4 >> 24;
            const uint32_t canon = ((cnt4 + (0x80u - THR) * 0x01010101u) >> 7) & 0x01010101u;
            const uint32_t msk = canon * 0xFFu;
            l4[g] = (l4[g] & msk) | (r4[g] & ~msk);
        }
    }
    uint32_t neq = 0, prev_hi = 0;
#pragma unroll
    for (int g = 0; g < 4; g++) {
        const uint32_t rel = l4[g];
        rel4[g] = rel;
        const uint32_t before = (rel << 8) | (g == 0 ? (rel & 0xFFu) : prev_hi);
        const uint32_t xd = rel ^ before;
        const uint32_t nz = ((((xd & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | xd) >> 7) & 0x01010101u;
        neq |= (((nz * 0x00204081u) >> 21) & 0xFu) << (4 * g);
        prev_hi = rel >> 24;
    }
    rel15 = prev_hi;

    // window j is valid iff j is not dead and no break bit lies in (j, j + L - 1]
    const uint64_t bw = bits64_at(brk, slot);
    const uint32_t dead16 = (uint32_t)bits64_at(dead, slot) & 0xFFFFu;
    uint64_t x = bw >> 1;
    x |= x >> 1; x |= x >> 2; x |= x >> 4; x |= x >> 8; x |= x >> 16;   // reach 31
    x |= x >> (WG::L - 2 - 31);                                        // reach L - 2 = 43
    const uint32_t v16 = ~((uint32_t)x | dead16) & 0xFFFFu;
    const uint32_t first16 = (uint32_t)bw & ~dead16 & 0xFFFFu;
    valid16 = v16;
    emask = v16 & (first16 | ((v16 << 1) & neq));   // bit 0 is completed by the caller (needs the previous block's last pick)
}

// ------------------------------------------------------------------ picks
DCN_HD bool wpick_valid(const WarpSmem &s, uint32_t p) {
    const uint32_t w0 = p >> 5, sh = p & 31u;
    const uint32_t bits = fshr(s.inv[w0], s.inv[w0 + 1], sh);
    return (bits & ((1u << WG::K) - 1u)) == 0;
}
DCN_HD uint64_t wpick_hash(const WarpSmem &s, uint32_t p) {
    const uint32_t w0 = p >> 4, sh = 2u * (p & 15u);
    const uint32_t a = s.codes[w0], b = s.codes[w0 + 1], c = s.codes[w0 + 2];
    const uint64_t fw = (((uint64_t)fshr(b, c, sh) << 32) | fshr(a, b, sh)) & ((1ULL << (2 * WG::K)) - 1ULL);
    const uint64_t rc = revcomp_2bit(fw, WG::K);
    return xxh3_u64(fw < rc ? fw : rc);
}

// NF picks of one lane (list entries idx0, idx0 + 32, ...): hash and request all of them, then test
template <int NF, bool EXTRACT>
DCN_HD void wprobe_picks(const FilterParams &P, WarpSmem &s, uint32_t idx0, uint32_t npicks) {
    uint16_t *pk_pos = s.pk_pos();
    uint64_t *pk_hash = s.pk_hash();
    uint32_t pp[NF];
    bool v[NF];
    uint64_t h[NF], bk[NF];
    Bucket k[NF];
#pragma unroll
    for (int f = 0; f < NF; f++) {
        const uint32_t idx = idx0 + (uint32_t)f * WG::NL;
        pp[f] = idx < npicks ? pk_pos[idx] : 0u;
        v[f] = idx < npicks && wpick_valid(s, pp[f]);
        h[f] = 0; bk[f] = 0;
        k[f].k0 = k[f].k1 = k[f].k2 = k[f].k3 = 0;
        if (v[f]) {
            h[f] = wpick_hash(s, pp[f]);
            if (!EXTRACT) { bk[f] = table_bucket(h[f], P.table.n_buckets); k[f] = load_bucket(P.table.slots, bk[f]); }
        }
    }
#pragma unroll
    for (int f = 0; f < NF; f++) {
        const uint32_t idx = idx0 + (uint32_t)f * WG::NL;
        if (v[f]) {
            pk_hash[idx] = h[f];
            pk_pos[idx] = (uint16_t)(pp[f] | 0x4000u | (!EXTRACT && table_contains_from(P.table, h[f], bk[f], k[f]) ? 0x8000u : 0u));
        }
    }
}

// ------------------------------------------------------------------ where a tile's bases come from
struct WSrc {
    const uint8_t *stage;      // ASCII: the warp's staged bytes, position 0 of the run at stage[0]; nullptr when packed
    uint32_t stage_bytes;      // readable bytes from `stage`
    const uint32_t *codes;     // packed form: word of the run's position 0 (global)
    const uint16_t *inv;
    uint64_t words;            // packed words readable from `codes` / `inv`
};

// effective length of a record (src/filter_common.rs:217-229): raw-length guard, prefix, one '\n'
DCN_HD uint32_t w_eff_len(const FilterParams &P, const WSrc &src, uint32_t r, uint32_t sL, uint32_t len) {
    if (len < (uint32_t)WG::K) return 0;
    const uint32_t n = (P.prefix_len > 0 && len > P.prefix_len) ? P.prefix_len : len;
    bool nl;
    if (src.stage) nl = src.stage[sL + n - 1] == (uint8_t)'\n';
    else { const uint32_t nb = P.nl_bit0 + r; nl = P.nl_bits && ((P.nl_bits[nb >> 5] >> (nb & 31u)) & 1u) != 0; }
    return nl ? n - 1 : n;
}

// ------------------------------------------------------------------ one run of whole short units
// Units [u_begin, u_end): at most MAXR records, every base inside [origin, origin + TB) (origin relative to
// base0, multiple of 16).  Ex provides par(f) = run f for every lane, then __syncwarp; scan = warp exclusive sum;
// ballot / match64 / bcast64 = warp votes; after_scan(ok) is called once the picks are counted (the device uses
// it to start the next tile's bulk copy into the stage the convert phase has drained).
// Returns false, nothing written, if the run emits more than PKCAP picks.
// With LONG the run is one chunk of a long unit instead: windows [la, la + carry + nw) of the tile are its own (the
// first one only carries the previous chunk's last pick for the consecutive-duplicate rule), no record table, every
// window may emit (the pick list keeps positions only: 1600 fit), and the picks go through the global (hash, unit)
// set `dd` and per-unit atomics instead of the per-unit count (filter_long_chunk is the CTA-tile form of the same).
struct WLong { uint32_t unit, la, carry, nw; const DedupView *dd; uint64_t reg_lo; uint32_t reg_sz; };   // reg_*: the unit's region of the set (dedup_region), 0 = none

// With EXTRACT (B3, get_minimizer_hashes_and_positions src/filter_common.rs:211-310; rpu = 1) nothing is probed: the
// picks' hashes and record-relative positions go to a block of the temp arrays (ex.xalloc), per-record valid counts
// feed the CSR offsets (ExtractOut, as extract_tiles_kernel).
template <bool PACKED, bool LONG, bool EXTRACT, class Ex>
DCN_HD bool warp_run(Ex &ex, const WarpTables &T, WarpSmem &s, const FilterParams &P, const WSrc &src,
                     uint64_t origin, uint32_t u_begin, uint32_t u_end, bool last_run, const WLong &lg) {
    using Priv = WarpPriv;
    const uint32_t r_begin = u_begin * P.rpu;
    const uint32_t n_rec_t = LONG ? 0u : (u_end - u_begin) * P.rpu;
    const uint32_t n_units_t = LONG ? 0u : u_end - u_begin;
    // windows may start in [dead_lo, dead_hi); sequence exists in [span_lo, span_hi)
    const uint32_t span_lo = LONG ? lg.la : (uint32_t)(P.rec_off[r_begin] - P.base0 - origin);
    const uint32_t span_hi = LONG ? lg.la + lg.carry + lg.nw + (uint32_t)WG::L - 1u : (uint32_t)(P.rec_off[r_begin + n_rec_t] - P.base0 - origin);
    const uint32_t dead_lo = span_lo, dead_hi = LONG ? lg.la + lg.carry + lg.nw : span_hi;

    // ---- P1: record boundaries (registers), start state of the bit arrays, convert
    ex.par([&](int l, Priv &pv) {
        for (int i = l; i < WG::NBW + 2; i += WG::NL) {
            s.dead[i] = outside_mask(32u * (uint32_t)i, dead_lo, dead_hi);
            s.brk[i] = outside_mask(32u * (uint32_t)i, span_lo, span_hi);
        }
#pragma unroll
        for (int j = 0; j < 2; j++) {
            const uint32_t i = (uint32_t)l + 32u * (uint32_t)j;
            pv.sL[j] = pv.eL[j] = pv.eff[j] = 0;
            if (!LONG && i < n_rec_t) {
                const uint32_t sL = (uint32_t)(P.rec_off[r_begin + i] - P.base0 - origin);
                const uint32_t eL = (uint32_t)(P.rec_off[r_begin + i + 1] - P.base0 - origin);
                pv.sL[j] = sL; pv.eL[j] = eL;
                pv.eff[j] = sL + w_eff_len(P, src, r_begin + i, sL, eL - sL);
            }
        }
        uint16_t *ih = reinterpret_cast<uint16_t *>(s.inv);   // 48 non-ACGT bits of the lane: three halfwords of its own
#pragma unroll 1
        for (int v = 0; v < 3; v++) {   // a loop, not three copies: 32 warps in 32 different places share one instruction cache
            uint32_t codes = 0, inv16 = 0xFFFFu;
            if (PACKED) {
                const uint64_t wi = 3ull * (uint64_t)l + (uint64_t)v;
                if (wi < src.words) { codes = src.codes[wi]; inv16 = src.inv[wi]; }
            } else {
                const uint32_t off = 48u * (uint32_t)l + 16u * (uint32_t)v;
                if (off + 16u <= src.stage_bytes) {
                    const u32x4a qv = *reinterpret_cast<const u32x4a *>(src.stage + off);
                    const uint32_t w[4] = {qv.x, qv.y, qv.z, qv.w};
                    inv16 = 0;
#pragma unroll
                    for (int i = 0; i < 4; i++) {
                        uint32_t p8, i4;
                        convert4(w[i], p8, i4);
                        codes |= p8 << (8 * i);
                        inv16 |= i4 << (4 * i);
                    }
                }
            }
            s.codes[3 * l + v] = codes;
            ih[3 * l + v] = (uint16_t)inv16;
        }
        if (l < 4) s.codes[WG::NSLOT + l] = 0;
        if (l < 2) s.inv[WG::NBW + l] = 0;
    });

    // ---- P2: record structure into the bit arrays; seed + hashes of block 0; its first 14 go to the left neighbour
    ex.par([&](int l, Priv &pv) {
#pragma unroll
        for (int j = 0; j < 2; j++) {
            const uint32_t i = (uint32_t)l + 32u * (uint32_t)j;
            if (i < n_rec_t) {
                set_bit(s.brk, pv.sL[j]);
                if ((i & (P.rpu - 1u)) == 0) s.ustartpos[i >> (P.rpu - 1u)] = (uint16_t)pv.sL[j];   // rpu is 1 or 2
                if (pv.eff[j] < pv.eL[j]) { set_bits(s.dead, pv.eff[j], pv.eL[j]); set_bits(s.brk, pv.eff[j], pv.eL[j]); }
            }
        }
        if (l == 0) s.ustartpos[n_units_t] = (uint16_t)span_hi;   // P7 takes a unit's length from two consecutive starts
        if (LONG && l == 0 && !lg.carry) set_bit(s.brk, lg.la);   // record start: its first window always emits
        // k-mer at the lane's first base: bases 0..30 = own words 0 and 1 (less its last base)
        const uint32_t c0 = s.codes[3 * l], c1 = s.codes[3 * l + 1], c2 = s.codes[3 * l + 2];
        uint32_t fw = 0, rc = 0;
#pragma unroll
        for (int g = 0; g < 8; g++) {
            uint32_t byte = ((g < 4 ? c0 : c1) >> (8 * (g & 3))) & 0xFFu;
            if (g == 7) byte &= 0x3Fu;
            const u32x2 e = T.tb0[byte];
            fw ^= rotr32(e.x, 4 * g);
            rc ^= rotl32(e.y, 4 * g);
        }
        fw ^= rotr32(nt_f(0u), 1);       // slot 31 of the table entry (code 0) is not part of the k-mer
        rc ^= rotl32(nt_f(2u), 31);
        whash_block<true>(T, fw, rc, c0, c1, c2, pv.hp);
        pv.fw = fw; pv.rc = rc;
        u32x4a *row = reinterpret_cast<u32x4a *>(&s.hx()[l * WG::HXP]);
        u32x4a a, b;
        a.x = pv.hp[0]; a.y = pv.hp[1]; a.z = pv.hp[2]; a.w = pv.hp[3];
        b.x = pv.hp[4]; b.y = pv.hp[5]; b.z = pv.hp[6]; b.w = pv.hp[7];
        row[0] = a; row[1] = b;
    });

    // ---- P3: the three blocks in turn: hashes of the next block (block 2: the right neighbour's first hashes), window
    // minima, consecutive-duplicate rule against the block before; per-block results go to shared memory
    ex.par([&](int l, Priv &pv) {
        uint32_t W0 = s.codes[3 * l], W1 = s.codes[3 * l + 1], W2 = s.codes[3 * l + 2], W3 = s.codes[3 * l + 3];
        uint32_t fw = pv.fw, rc = pv.rc;
        uint32_t q[15], prev_last = 0xFFFFu;   // last pick of the previous block, relative to ITS first base
#pragma unroll
        for (int j = 0; j < 8; j++) q[j] = pv.hp[j];
#pragma unroll 1
        for (int r = 0; r < 3; r++) {
            const uint32_t W4 = s.codes[3 * l + r + 4];
            uint32_t hb[8];
            if (r < 2) {
                whash_block<true>(T, fw, rc, W1, W2, W3, hb);   // (block 2's crossing step is one step too many: harmless)
            } else {   // lane 31 reads the pad row: its last windows are dead
                const u32x4a *row = reinterpret_cast<const u32x4a *>(&s.hx()[(l + 1) * WG::HXP]);
                const u32x4a a = row[0], b = row[1];
                hb[0] = a.x; hb[1] = a.y; hb[2] = a.z; hb[3] = a.w; hb[4] = b.x; hb[5] = b.y; hb[6] = b.z; hb[7] = b.w;
            }
#pragma unroll
            for (int j = 0; j < 7; j++) q[8 + j] = hb[j];
            u32x4a rel;
            uint32_t rel4[4], emask, v16, r15;
            wslide_block(q, W0, W1, W2, W3, s.brk, s.dead, 3 * l + r, rel4, emask, v16, r15);
            // consecutive-duplicate rule across the lane's own block boundaries (A.3 step 5)
            if ((v16 & 1u) && !(emask & 1u) && prev_last != 0xFFFFu && prev_last != 16u + (rel4[0] & 0xFFu)) emask |= 1u;
            if (r == 0) pv.vfirst = v16 & 1u;
            prev_last = (v16 & 0x8000u) ? r15 : 0xFFFFu;
            rel.x = rel4[0]; rel.y = rel4[1]; rel.z = rel4[2]; rel.w = rel4[3];
            s.rel4()[3 * l + r] = rel;
            s.em()[3 * l + r] = emask;
            W0 = W1; W1 = W2; W2 = W3; W3 = W4;
#pragma unroll
            for (int j = 0; j < 8; j++) q[j] = hb[j];
        }
        s.lastpick[l] = prev_last != 0xFFFFu ? (uint16_t)(48u * (uint32_t)l + 32u + prev_last) : (uint16_t)0xFFFFu;
    });

    // ---- P4 + scan: first window of each lane against the left neighbour's last pick; pick offsets of every block
    ex.scan([&](int l, Priv &pv) {
                uint32_t e0 = s.em()[3 * l];
                if (pv.vfirst && !(e0 & 1u)) {
                    const uint32_t lp = l > 0 ? s.lastpick[l - 1] : 0xFFFFu;
                    const uint32_t mine = 48u * (uint32_t)l + (s.rel4()[3 * l].x & 0xFFu);
                    if (lp != 0xFFFFu && lp != mine) { e0 |= 1u; s.em()[3 * l] = e0; }
                }
                return popc32(e0) + popc32(s.em()[3 * l + 1]) + popc32(s.em()[3 * l + 2]);
            },
            [&](int l, Priv &pv, uint32_t excl, uint32_t total) {
                pv.pickoff = excl;
                uint32_t off = excl;
#pragma unroll
                for (int r = 0; r < 3; r++) {
                    const uint32_t e = s.em()[3 * l + r];
                    s.em()[3 * l + r] = e | (off << 16);
                    off += popc32(e);
                }
                if (l == 0) s.npicks = total;
            });
    const uint32_t npicks = s.npicks;
    if (!LONG && npicks > (uint32_t)WG::PKCAP) { ex.after_scan(false); return false; }
    ex.after_scan(last_run);

    if (LONG) {
        // ---- P5 (long): block results into registers first: up to 1477 pick positions overwrite them
        ex.par([&](int l, Priv &pv) {
#pragma unroll
            for (int r = 0; r < 3; r++) { pv.lrel[r] = s.rel4()[3 * l + r]; pv.lem[r] = s.em()[3 * l + r]; }
        });
        ex.par([&](int l, Priv &pv) {
#pragma unroll
            for (int r = 0; r < 3; r++) {
                const u32x4a rel = pv.lrel[r];
                uint32_t em = pv.lem[r] & 0xFFFFu, idx = pv.lem[r] >> 16;
                const uint64_t relA = rel.x | ((uint64_t)rel.y << 32), relB = rel.z | ((uint64_t)rel.w << 32);
                const uint32_t base = 48u * (uint32_t)l + 16u * (uint32_t)r;
                while (em) {
#ifdef __CUDA_ARCH__
                    const int i = __ffs((int)em) - 1;
#else
                    const int i = __builtin_ctz(em);
#endif
                    em &= em - 1;
                    const uint32_t rl = (uint32_t)(((i & 8) ? relB : relA) >> (8 * (i & 7))) & 0xFFu;
                    s.pk_pos()[idx++] = (uint16_t)(base + rl);
                }
            }
        });
        // ---- P6 (long): probe every pick (two per lane in flight), then the hits alone through the (hash, unit) set.
        // A set insert is two or three dependent round trips (compare-and-swap, again when the slot holds an entry of
        // an earlier call); done inside the probe loop every iteration paid them as soon as one lane had a hit.  Now the
        // probe loop only flags the hits, a ballot pass compacts them (~14 % of the picks of a read that shares a
        // quarter of its minimizers with the index) and the inserts run over full lanes: 3 + 2 round trips per chunk
        // instead of 9.
        constexpr uint32_t HCAP = 2u * (uint32_t)(2 * (WG::NBW + 2));   // hit positions that fit brk | dead (free after the slide phase)
        uint16_t *hl = reinterpret_cast<uint16_t *>(s.brk);
#if defined(DCN_EMU_HCAP) && !defined(__CUDA_ARCH__)
        const uint32_t hcap = (DCN_EMU_HCAP) < HCAP ? (DCN_EMU_HCAP) : HCAP;   // host emulation: a test shrinks the list to reach the leftover branch
#else
        const uint32_t hcap = HCAP;
#endif
        static_assert(offsetof(WarpSmem, dead) == offsetof(WarpSmem, brk) + sizeof(uint32_t) * (WG::NBW + 2), "brk | dead contiguous");
        ex.par([&](int l, Priv &) {
            uint16_t *pk_pos = s.pk_pos();
            uint32_t n_valid = 0;   // this lane's picks
            constexpr int NF = DCN_LONG_PICKS_IN_FLIGHT;
            for (uint32_t idx0 = (uint32_t)l; idx0 < npicks; idx0 += NF * WG::NL) {
                uint32_t pp[NF];
                bool v[NF];
                uint64_t h[NF], bk[NF];
                Bucket k[NF];
#pragma unroll
                for (int f = 0; f < NF; f++) {
                    const uint32_t idx = idx0 + (uint32_t)f * WG::NL;
                    pp[f] = idx < npicks ? pk_pos[idx] : 0u;
                    v[f] = idx < npicks && wpick_valid(s, pp[f]);
                    h[f] = 0; bk[f] = 0;
                    k[f].k0 = k[f].k1 = k[f].k2 = k[f].k3 = 0;
                    if (v[f]) { h[f] = wpick_hash(s, pp[f]); bk[f] = table_bucket(h[f], P.table.n_buckets); k[f] = load_bucket(P.table.slots, bk[f]); }
                }
#pragma unroll
                for (int f = 0; f < NF; f++) {
                    const uint32_t idx = idx0 + (uint32_t)f * WG::NL;
                    if (v[f] && table_contains_from(P.table, h[f], bk[f], k[f])) pk_pos[idx] = (uint16_t)(pp[f] | 0x8000u);
                    n_valid += (uint32_t)v[f];
                }
            }
            s.ufirst[l] = (uint16_t)n_valid;      // (<= 47 per lane; the unit tables are free in a long chunk)
        });
        ex.par([&](int l, Priv &) {
            uint16_t *pk_pos = s.pk_pos();
            uint32_t nh = 0;
            for (uint32_t base = 0; base < npicks; base += WG::NL) {
                const uint32_t idx = base + (uint32_t)l;
                const uint32_t pp = idx < npicks ? pk_pos[idx] : 0u;
                const bool hit = (pp & 0x8000u) != 0;
                const uint32_t m = ex.ballot(l, hit);
                const uint32_t at = nh + popc32(m & ((1u << l) - 1u));
                if (hit && at < hcap) { hl[at] = (uint16_t)(pp & 0x7FFFu); pk_pos[idx] = (uint16_t)(pp & 0x7FFFu); }   // (beyond the list: the flag stays)
                nh += popc32(m);
            }
            if (l == 0) s.nhits = nh < hcap ? nh : hcap;
        });
        ex.par([&](int l, Priv &) {
            const uint16_t *pk_pos = s.pk_pos();
            const uint32_t nlist = s.nhits;
            uint32_t n_fresh = 0;
            for (uint32_t i = (uint32_t)l; i < nlist; i += WG::NL)
                n_fresh += (uint32_t)dedup_insert(*lg.dd, wpick_hash(s, hl[i]), lg.unit, lg.reg_lo, lg.reg_sz);
            for (uint32_t idx = (uint32_t)l; idx < npicks; idx += WG::NL) {   // a chunk with more than HCAP hits: the rest, lane by lane
                const uint32_t pp = pk_pos[idx];
                if (pp & 0x8000u) n_fresh += (uint32_t)dedup_insert(*lg.dd, wpick_hash(s, pp & 0x7FFFu), lg.unit, lg.reg_lo, lg.reg_sz);
            }
            s.ustartpos[l] = (uint16_t)n_fresh;
        });
        ex.par([&](int l, Priv &) {
            if (l == 0) {
                uint32_t nv = 0, nf = 0;
                for (int i = 0; i < WG::NL; i++) { nv += s.ufirst[i]; nf += s.ustartpos[i]; }
                ex.global_add(&P.total[lg.unit], nv);
                ex.global_add(&P.hits[lg.unit], nf);
            }
        });
        return true;
    }

    // ---- P5: compact the picks (the positions take over the hx rows: every lane is past its last hx read);
    // unit -> pick range from the per-block offsets
    ex.par([&](int l, Priv &) {
#pragma unroll 1
        for (int r = 0; r < 3; r++) {
            const u32x4a rel = s.rel4()[3 * l + r];
            const uint32_t e = s.em()[3 * l + r];
            uint32_t em = e & 0xFFFFu, idx = e >> 16;
            const uint64_t relA = rel.x | ((uint64_t)rel.y << 32), relB = rel.z | ((uint64_t)rel.w << 32);
            const uint32_t base = 48u * (uint32_t)l + 16u * (uint32_t)r;
            while (em) {
#ifdef __CUDA_ARCH__
                const int i = __ffs((int)em) - 1;
#else
                const int i = __builtin_ctz(em);
#endif
                em &= em - 1;
                const uint32_t rl = (uint32_t)(((i & 8) ? relB : relA) >> (8 * (i & 7))) & 0xFFu;
                s.pk_pos()[idx++] = (uint16_t)(base + rl);
            }
        }
        for (uint32_t u = (uint32_t)l; u < n_units_t; u += WG::NL) {
            const uint32_t pos = s.ustartpos[u];
            const uint32_t tt = pos >> 4, ii = pos & 15u;
            uint32_t first = npicks;   // an empty unit at the very end of the tile starts past the last block
            if (tt < (uint32_t)WG::NSLOT) { const uint32_t e = s.em()[tt]; first = (e >> 16) + popc32(e & ((1u << ii) - 1u)); }
            s.ufirst[u] = (uint16_t)first;
            if (u == 0) s.ufirst[n_units_t] = (uint16_t)npicks;
        }
    });

    // ---- P6: hash every pick and probe the table, DCN_PICKS_IN_FLIGHT picks per lane in flight: hash + request all
    // of them, then test (the hashes take over rel4 / em).  pk_pos: position | valid << 14 | in-index << 15.
    // A tile of 2x150 pairs emits 141 +- 10 picks: at two per lane that is two full iterations and a third for a
    // dozen lanes.  Measured (quick bench, Gbp/s): one per lane 298.9 (kept), two 294.6 - 296.0, three 268.6, four 252.5; two, and three
    // in the iteration that then finishes the list (DCN_PICKS_LAST_3: two round trips per tile instead of three) 266.3;
    // one, and two in the last 276.8.  Fewer round trips do not pay: a second copy of the probe body in the loop costs
    // more in instruction fetch than a round trip costs in latency (the body is ~2800 instructions and 32 warps are in
    // 32 different places of it), and more requests per warp at 74 % of the request ceiling only lengthen the queues.
    ex.par([&](int l, Priv &) {
        constexpr int NF = DCN_PICKS_IN_FLIGHT;
        uint32_t base = 0;
        while (base < npicks) {
            const uint32_t rem = npicks - base;
            if (DCN_PICKS_LAST_3 && !EXTRACT && rem > (uint32_t)NF * WG::NL && rem <= (uint32_t)(NF + 1) * WG::NL) {
                wprobe_picks<NF + 1, EXTRACT>(P, s, base + (uint32_t)l, npicks);
                base += (uint32_t)(NF + 1) * WG::NL;
            } else {
                wprobe_picks<NF, EXTRACT>(P, s, base + (uint32_t)l, npicks);
                base += (uint32_t)NF * WG::NL;
            }
        }
    });

    if (EXTRACT) {
        // ---- P7 (extract): the warp walks its records; a record's picks (list order = position order) go to the
        // tile's block of the temp arrays with the position made relative to the record
        const uint64_t tbase = ex.xalloc(npicks);
        ex.par([&](int l, Priv &) {
            const uint32_t lane = (uint32_t)l;
            const uint16_t *pk_pos = s.pk_pos();
            const uint64_t *pk_hash = s.pk_hash();
            for (uint32_t u = 0; u < n_units_t; u++) {
                const uint32_t a = s.ufirst[u], b = s.ufirst[u + 1], sL = s.ustartpos[u];
                uint32_t cnt = 0;
                for (uint32_t base = a; base < b; base += 32u) {
                    const uint32_t idx = base + lane;
                    bool valid = false;
                    if (idx < b) {
                        const uint32_t pp = pk_pos[idx];
                        valid = (pp & 0x4000u) != 0;
                        const uint64_t at = tbase + idx;
                        if (at < P.xo.tmp_cap) {
                            P.xo.tmp_p[at] = ((pp & 0x3FFFu) - sL) | (valid ? 0x80000000u : 0u);
                            if (valid) P.xo.tmp_h[at] = pk_hash[idx];
                        }
                    }
                    cnt += popc32(ex.ballot(l, valid));
                }
                if (lane == 0) {
                    P.xo.rec_cnt[u_begin + u] = cnt;
                    P.xo.rec_tmp[u_begin + u] = ((tbase + a) << 16) | (uint64_t)(b - a);
                }
            }
        });
        return true;
    }

    // ---- P7: distinct hits per unit (src/filter_common.rs:143-145) and the threshold test: the warp walks its
    // units; a unit's picks are consecutive list entries, 32 per pass (see filter_short_tile for the later-pass rules)
    ex.par([&](int l, Priv &) {
        const uint32_t lane = (uint32_t)l, lt = (1u << lane) - 1u;
        const uint16_t *pk_pos = s.pk_pos();
        const uint64_t *pk_hash = s.pk_hash();
        for (uint32_t u = 0; u < n_units_t; u++) {
            const uint32_t a = s.ufirst[u], b = s.ufirst[u + 1];
            uint32_t hits = 0, total = 0, vprev = 0;
            uint64_t hprev = 0;
            for (uint32_t base = a; base < b; base += 32u) {
                const uint32_t idx = base + lane;
                bool valid = false, found = false;
                uint64_t h = 0;
                if (idx < b) {
                    const uint32_t pp = pk_pos[idx];
                    valid = (pp & 0x4000u) != 0;
                    found = (pp & 0x8000u) != 0;
                    if (valid) h = pk_hash[idx];
                }
                const uint32_t vmask = ex.ballot(l, valid);
                const uint32_t same = ex.match64(l, h, valid);
                bool fresh = valid && found && (same & vmask & lt) == 0;
                if (base > a) {
                    uint32_t cand = ex.ballot(l, fresh);
                    while (cand) {
                        const uint32_t cl = popc32((cand & (0u - cand)) - 1u);
                        cand &= cand - 1u;
                        const uint64_t hv = ex.bcast64(l, h, cl);
                        const uint32_t hitprev = ex.ballot(l, ((vprev >> lane) & 1u) != 0 && hprev == hv);
                        if (lane == cl && hitprev) fresh = false;
                    }
                    if (fresh && base > a + 32u) {
                        for (uint32_t j = a; j < base - 32u && fresh; j++)
                            if (pk_hash[j] == h && (pk_pos[j] & 0x4000u)) fresh = false;
                    }
                }
                hits += popc32(ex.ballot(l, fresh));
                total += popc32(vmask);
                hprev = h; vprev = vmask;
            }
            if (lane == 0) {
                const uint32_t gu = u_begin + u;
                P.total[gu] = total;
                P.hits[gu] = hits;
                bool keep;
                if (total < 256u) { const uint32_t req = T.req[total]; keep = P.deplete ? hits < req : hits >= req; }
                else keep = meets_criteria(hits, total, P.abs_thr, P.rel_thr, P.deplete);
                P.keep[gu] = keep ? 1 : 0;
                ex.tally(P.rpu, (uint32_t)s.ustartpos[u + 1] - (uint32_t)s.ustartpos[u], keep);   // the six summary counters (a13)
            }
        }
    });
    return true;
}

// A tile: its units as one run, split in halves while a run emits more picks than one pass can hold.  A single
// unit that still does not fit is handed to `overflow(u)` (the CTA-tile path holds 1024 picks per unit).
// `stage0` = staged ASCII of the tile (position 0 = tile origin), `stage_bytes` = valid bytes in it.
template <bool PACKED, bool EXTRACT, class Ex, class Ovf>
DCN_HD void warp_tile(Ex &ex, const WarpTables &T, WarpSmem &s, const FilterParams &P, const WTile &tile,
                      uint32_t stage_bytes, Ovf overflow) {
    uint32_t lo = tile.a, span = tile.b - tile.a;
    while (lo < tile.b) {
        const uint32_t hi = lo + span < tile.b ? lo + span : tile.b;
        const uint64_t first = P.rec_off[(uint64_t)lo * P.rpu] - P.base0;
        const uint64_t origin = lo == tile.a ? tile.origin : (first & ~15ull);
        const uint32_t delta = (uint32_t)(origin - tile.origin);
        WSrc src;
        src.stage = PACKED ? nullptr : s.stage + delta;
        // whole 16-byte vectors: what follows the tile's last byte in its last vector is stale, and outside every record
        const uint32_t sb16 = (stage_bytes + 15u) & ~15u;
        src.stage_bytes = sb16 > delta ? sb16 - delta : 0u;
        src.codes = PACKED ? P.pk_codes + (origin >> 4) : nullptr;
        src.inv = PACKED ? P.pk_inv + (origin >> 4) : nullptr;
        const uint64_t n_rel = P.n_bases - P.base0;
        src.words = PACKED ? ((n_rel + 15) >> 4) - (origin >> 4) : 0;
        WLong none;
        none.unit = none.la = none.carry = none.nw = 0; none.dd = nullptr; none.reg_lo = 0; none.reg_sz = 0;
        if (warp_run<PACKED, false, EXTRACT>(ex, T, s, P, src, origin, lo, hi, hi == tile.b, none)) {
            lo = hi;
        } else if (hi - lo == 1) {
            overflow(lo);
            if (hi == tile.b) ex.after_scan(true);
            lo = hi;
        } else {
            span = (hi - lo + 1) / 2;
        }
    }
}

// One chunk of a long unit (descriptor: WTILE_LONG set): see WLong.
template <bool PACKED, class Ex>
DCN_HD void warp_long_tile(Ex &ex, const WarpTables &T, WarpSmem &s, const FilterParams &P, const DedupView &dd,
                           const WTile &tile, uint32_t stage_bytes) {
    const uint64_t origin = tile.origin & ~WTILE_LONG;
    WLong lg;
    lg.unit = tile.a / P.rpu; lg.la = tile.b & 15u; lg.carry = (tile.b >> 4) & 1u; lg.nw = tile.b >> 5; lg.dd = &dd;
    lg.reg_lo = 0; lg.reg_sz = 0;
    if (dd.per16)
        dedup_region(dd, P.rec_off[(uint64_t)lg.unit * P.rpu] - P.base0, P.rec_off[(uint64_t)(lg.unit + 1) * P.rpu] - P.base0, lg.reg_lo, lg.reg_sz);
    WSrc src;
    src.stage = PACKED ? nullptr : s.stage;
    src.stage_bytes = (stage_bytes + 15u) & ~15u;
    src.codes = PACKED ? P.pk_codes + (origin >> 4) : nullptr;
    src.inv = PACKED ? P.pk_inv + (origin >> 4) : nullptr;
    const uint64_t n_rel = P.n_bases - P.base0;
    src.words = PACKED ? ((n_rel + 15) >> 4) - (origin >> 4) : 0;
    warp_run<PACKED, true, false>(ex, T, s, P, src, origin, 0u, 0u, true, lg);
}

// chunk descriptors of one long record with `eff_len` effective bases starting at gs (relative to base0)
template <class Emit>
DCN_HD uint32_t wplan_long_record(uint32_t rec, uint64_t gs, uint64_t eff_len, Emit emit) {
    if (eff_len < (uint64_t)WG::L) return 0;
    const uint64_t nwin = eff_len - (uint64_t)WG::L + 1;
    const uint32_t nc = (uint32_t)((nwin + DCN_WCS - 1) / DCN_WCS);
    for (uint32_t c = 0; c < nc; c++) {
        const uint64_t w0 = (uint64_t)c * DCN_WCS;
        const uint32_t nw = (uint32_t)(nwin - w0 < DCN_WCS ? nwin - w0 : DCN_WCS), carry = c > 0 ? 1u : 0u;
        const uint64_t a = gs + w0 - carry, origin = a & ~15ull;
        WTile t;
        t.origin = origin | WTILE_LONG; t.a = rec; t.b = (uint32_t)(a - origin) | (carry << 4) | (nw << 5);
        emit(c, t);
    }
    return nc;
}
DCN_HD uint32_t wplan_long_chunks_of(uint64_t eff_len) {
    if (eff_len < (uint64_t)WG::L) return 0;
    return (uint32_t)((eff_len - (uint64_t)WG::L + 1 + DCN_WCS - 1) / DCN_WCS);
}

// ------------------------------------------------------------------ planner
// Greedy: a tile takes units while they fit (bases, records) and are short; long units are skipped (long path).
// One thread per segment of the batch (units whose first base lies in [seg * SEG, (seg + 1) * SEG)).
static constexpr uint32_t DCN_WSEG = 1u << 16;

// units a tile may hold: MAXR records, and at most the 32 the planner's warp looks at in one step
DCN_HD uint32_t wplan_max_units(uint32_t rpu) { const uint32_t m = (uint32_t)WG::MAXR / rpu; return m < 32u ? m : 32u; }

DCN_HD uint32_t wplan_first_unit(const uint64_t *rec_off, uint64_t base0, uint32_t rpu, uint32_t n_units, uint64_t pos) {
    uint32_t lo = 0, hi = n_units;   // first unit with start >= pos
    while (lo < hi) {
        const uint32_t mid = lo + (hi - lo) / 2;
        if (rec_off[(uint64_t)mid * rpu] - base0 < pos) lo = mid + 1; else hi = mid;
    }
    return lo;
}

// Walks the segment's units; calls emit(origin, a, b) per tile; returns the tile count.
template <class Emit>
DCN_HD uint32_t wplan_segment(const uint64_t *rec_off, uint64_t base0, uint32_t rpu, uint32_t n_units, uint64_t seg,
                              Emit emit) {
    const uint64_t lo_pos = seg * DCN_WSEG, hi_pos = lo_pos + DCN_WSEG;
    uint32_t u = wplan_first_unit(rec_off, base0, rpu, n_units, lo_pos);
    const uint32_t maxu = wplan_max_units(rpu);
    uint32_t n = 0;
    while (u < n_units) {
        const uint64_t start = rec_off[(uint64_t)u * rpu] - base0;
        if (start >= hi_pos) break;
        uint64_t end = rec_off[(uint64_t)(u + 1) * rpu] - base0;
        if (end - start > DCN_MAX_SHORT) { u++; continue; }
        const uint64_t origin = start & ~15ull;
        uint32_t v = u + 1;
        while (v < n_units && v - u < maxu) {
            const uint64_t s2 = end;   // == start of unit v
            if (s2 >= hi_pos) break;
            const uint64_t e2 = rec_off[(uint64_t)(v + 1) * rpu] - base0;
            if (e2 - s2 > DCN_MAX_SHORT || e2 - origin > (uint64_t)WG::TB) break;
            end = e2; v++;
        }
        emit(origin, u, v);
        n++;
        u = v;
    }
    return n;
}

}  // namespace dcn
