// dcn_tile.cuh -- the tile pipeline of the fused extract -> lookup -> classify kernel.
//
// One CTA (256 threads) owns a tile: a 16-byte-aligned span of the concatenated base stream that
// holds up to 4080 window starts.  Thread t owns the 16 bases [16t, 16t+16) of the span: it
// converts them (2-bit codes + non-ACGT mask), hashes the 16 k-mers that start there (rolling
// canonical ntHash, seeded from per-thread aggregates so no thread re-reads 31 bases), and picks
// the minimizer of the 16 windows that start there (van Herk / Gil-Werman min over w = 15 with
// R = w + 1 = 16 outputs per thread).  Picks are compacted into a shared-memory list and then
// processed one pick per thread: ACGT filter, canonical 2-bit k-mer, xxh3, one 32-byte probe
// of the HBM table, per-unit distinct-hit count, threshold test.
//
// Every phase is a function of (thread id, shared state, per-thread private state) so the same
// code runs on the device (phases separated by __syncthreads) and, for tests only, on the host
// (phases run as loops over t).  Reference semantics: SURVEY.md Appendix A.
#pragma once
#include "dcn_core.cuh"
#include "dcn_plan.cuh"

namespace dcn {

enum Flavour { FLAVOUR_FILTER = 0, FLAVOUR_INDEX = 1 };

template <int K_, int W_>
struct Geo {
    static_assert(W_ == 15, "fast path is specialised for w = 15 (R = w + 1 = 16 windows per thread)");
    static_assert(K_ >= 17 && K_ <= 31, "fast path needs 17 <= k <= 31 (k-mer spans exactly 3 threads' words)");
    static constexpr int K = K_, W = W_, L = K_ + W_ - 1;
#ifndef DCN_NT
#define DCN_NT 256
#endif
    static constexpr int NT = DCN_NT;           // threads per CTA
    static constexpr int NV = NT + 2;           // 16-byte vectors loaded per tile
    static constexpr int WCAP = NT * 16 - 16;   // window starts per tile (thread 255 only hashes)
    static constexpr int BCAP = WCAP + L - 1;   // bases a tile may reference
    static constexpr int NBW = NT / 2 + 4;      // 32-bit words of the per-position bit arrays
#ifndef DCN_MAXR
#define DCN_MAXR (2 * DCN_NT)
#endif
    static constexpr int MAXR = DCN_MAXR;       // records per sub-batch
    static constexpr int PKCAP = 4 * NT < 1024 ? 1024 : 4 * NT;   // picks per pass (>= the 980 windows of one short unit); a denser run of units is split (see filter_tile)
    static constexpr int HP = 20;               // hrow pitch in words (16 data + 4 pad: conflict-free LDS.128)
};

struct alignas(16) Bucket { uint64_t k0, k1, k2, k3; };   // one 32-byte bucket of the HBM table
struct u32x2 { uint32_t x, y; };
struct u32x4 { uint32_t x, y, z, w; };

template <class G>
struct alignas(16) TileSmem {
    u32x4 ag[G::NV + 6];              // per vector: E_fw, E_rc (own role), N_fw, N_rc (right-neighbour role)
    union {                           // hrow is dead once the window minima are taken; picks reuse it
        uint32_t hrow[G::NT * G::HP]; // ntHash (upper 16 bits) of the 16 k-mers of each thread
        uint32_t pk_pos[G::NT * G::HP];   // local position | in-index << 30 | valid << 31
    };
    uint64_t pk_hash[G::PKCAP];       // xxh3 of each pick
    u32x2 tb0[256];                   // 4-base aggregate table (fw, rc)
    u32x2 tio[16];                    // rolling-step table indexed by outgoing code | incoming code << 2 (fw, rc parts)
    u32x2 tsb[4];                     // single base at the neighbour-role rotation
    uint32_t codes[G::NV + 6];        // 16 bases x 2 bit per word
    uint32_t inv[(G::NV + 6 + 1) / 2 + 2];  // non-ACGT bits, 32 positions per word
    uint32_t brk[G::NBW + 2];         // position is a record start or lies outside every effective sequence
    uint32_t dead[G::NBW + 2];        // no window may start here
    uint32_t lastpick[G::NT];
    uint32_t poff[G::NT];             // exclusive pick offset
    uint32_t emk[G::NT];              // emit mask
    uint16_t ufirst[G::MAXR + 2];     // first pick index of each unit
    uint16_t ustartpos[G::MAXR + 2];
    uint32_t rec_se[G::MAXR];         // record start | end << 16 (tile-local), loaded one phase early
    uint16_t rec_eff[G::MAXR];        // end of the record's effective sequence (tile-local)
    uint32_t wsum[16];
    uint32_t npicks;
    uint32_t next_tile[2];            // tile claims of filter_fused_kernel: next tile, tile after next
    uint16_t req[256];                // required_hits(total) for total < 256 (src/filter_common.rs:84-96), filled once per CTA
};
template <class G>
struct TilePriv {
    uint32_t c0;           // own 16 codes
    uint32_t efw, erc;     // own aggregates
    uint32_t h[16];        // ntHash of own 16 k-mers
    uint32_t rel4[4];      // pick position relative to 16t, 8 bits per window
    uint32_t emask;        // windows that emit a pick
    uint32_t valid16;
    uint32_t pickoff;
};

// ------------------------------------------------------------------ tables
template <class G>
DCN_HD void init_tables(int t, TileSmem<G> &s) {
    // tb0[b]: bases c0..c3 of byte b at group 0: fw = XOR rotl(F[c_m], 30-m), rc = XOR rotl(F[c_m^2], m)
    for (int b = t; b < 256; b += G::NT) {
        uint32_t fw = 0, rc = 0;
        for (int m = 0; m < 4; m++) {
            uint32_t c = ((uint32_t)b >> (2 * m)) & 3u;
            fw ^= rotl32(nt_f(c), (uint32_t)(30 - m));
            rc ^= rotl32(nt_f(c ^ 2u), (uint32_t)m);
        }
        s.tb0[b].x = fw; s.tb0[b].y = rc;
    }
    if (t < 16) {   // one step of the rolling hash: fw = rotl(fw, 1) ^ x, rc = rotr(rc ^ y, 1)
        uint32_t oc = (uint32_t)t & 3u, ic = (uint32_t)t >> 2;
        s.tio[t].x = rotl32(nt_f(oc), G::K) ^ nt_f(ic);
        s.tio[t].y = nt_f(oc ^ 2u) ^ rotl32(nt_f(ic ^ 2u), G::K);
    }
    if (t < 4) {
        uint32_t c = (uint32_t)t;
        s.tsb[t].x = rotl32(nt_f(c), 15);              s.tsb[t].y = rotl32(nt_f(c ^ 2u), 15);
    }
}

// required hits per total, so that the per-unit threshold test is a table lookup instead of f64 arithmetic
template <class G>
DCN_HD void init_required(int t, TileSmem<G> &s, uint32_t abs_thr, double rel_thr) {
    for (int i = t; i < 256; i += G::NT) {
        const uint64_t r = required_hits(abs_thr, rel_thr, (uint64_t)i);
        s.req[i] = (uint16_t)(r > 0xFFFFull ? 0xFFFFull : r);   // hits <= total < 256: saturation keeps every comparison exact
    }
}

// ------------------------------------------------------------------ phase 1: load + convert
// FILTER flavour: packed-seq lossy code (b >> 1) & 3, src/filter_common.rs:238.
// INDEX flavour: IUPAC map first (src/minimizers.rs:24-43), then the same packing.
DCN_HD uint32_t code_index_flavour(uint32_t b) {
    // letters only differ from the lossy code when non-ACGT: R,S,K,D,V,G -> G(3); A,W -> A(0); T -> T(2); else C(1)
    uint32_t idx = b & 0x1fu;
    bool letter = (b & 0xC0u) == 0x40u && idx >= 1 && idx <= 26;
    const uint32_t MG = (1u << 18) | (1u << 19) | (1u << 11) | (1u << 4) | (1u << 22) | (1u << 7);
    const uint32_t MA = (1u << 1) | (1u << 23);
    const uint32_t MT = (1u << 20);
    if (!letter) return 1u;
    if ((MG >> idx) & 1u) return 3u;
    if ((MA >> idx) & 1u) return 0u;
    if ((MT >> idx) & 1u) return 2u;
    return 1u;
}

// Where a tile's bases come from: ASCII bytes (device-resident batches, index builds) or the
// host-packed form the host-pointer pipeline ships over PCIe (2-bit codes, 16 bases per u32 word
// in packed-seq order, plus one non-ACGT bit per base: 0.375 B/bp instead of 1 B/bp).  Offsets are
// relative to base0; word i of `codes` / `inv` covers relative bases [16 i, 16 i + 16).
struct TileSrc {
    const uint8_t *bases;
    const uint32_t *codes;
    const uint16_t *inv;
    uint64_t n_bases;   // readable extent (relative)
};

template <class G, int FLAV, bool PACKED>
DCN_HD void convert_vector(int v, TileSmem<G> &s, const TileSrc &src, uint64_t origin,
                           uint32_t &codes_out, uint32_t &efw_out, uint32_t &erc_out) {
    uint64_t g = origin + 16ull * (uint64_t)v;
    const uint64_t n_bases = src.n_bases;
    uint32_t codes = 0, inv16 = 0;
    if (PACKED) {
        // the packer pads the last word (codes 0, non-ACGT bits 1), like the byte path below
        if (g < n_bases) { codes = src.codes[g >> 4]; inv16 = src.inv[g >> 4]; }
        else inv16 = 0xFFFFu;
    } else {
    const uint8_t *bases = src.bases;
    uint32_t w[4];
    if (g + 16 <= n_bases) {
#ifdef __CUDA_ARCH__
        uint4 q = __ldg(reinterpret_cast<const uint4 *>(bases + g));
#else
        u32x4 q = *reinterpret_cast<const u32x4 *>(bases + g);
#endif
        w[0] = q.x; w[1] = q.y; w[2] = q.z; w[3] = q.w;
    } else {
        for (int i = 0; i < 4; i++) {
            uint32_t x = 0;
            for (int b = 0; b < 4; b++) {
                uint64_t a = g + (uint64_t)(4 * i + b);
                uint32_t byte = a < n_bases ? bases[a] : 0u;
                x |= byte << (8 * b);
            }
            w[i] = x;
        }
    }
#pragma unroll
    for (int i = 0; i < 4; i++) {
        uint32_t x = w[i];
        uint32_t c = (x >> 1) & 0x03030303u;
        // exact ACGT/acgt test: (byte & 0xDF) must equal the letter its own code stands for
        uint32_t c0b = c & 0x01010101u, c1b = (c >> 1) & 0x01010101u;
        uint32_t e = 0x41414141u + c0b * 2u + c1b * 0x13u - (c0b & c1b) * 0xFu;
        uint32_t d = (x & 0xDFDFDFDFu) ^ e;
        uint32_t nz = (((d & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | d) & 0x80808080u;
        uint32_t inv4 = (nz * 0x00204081u) >> 28;
        if (FLAV == FLAVOUR_INDEX) {
            uint32_t cc = 0;
            for (int b = 0; b < 4; b++) cc |= code_index_flavour((x >> (8 * b)) & 0xffu) << (8 * b);
            c = cc;
        }
        uint32_t packed8 = (c * 0x01041040u) >> 24;
        codes |= packed8 << (8 * i);
        inv16 |= inv4 << (4 * i);
    }
    }
    // aggregates over the 16 bases: E_fw = XOR rotl(F[c_i], 30-i), E_rc = XOR rotl(F[c_i^2], i)
    uint32_t efw = 0, erc = 0;
#pragma unroll
    for (int gI = 0; gI < 4; gI++) {
        u32x2 e = s.tb0[(codes >> (8 * gI)) & 0xffu];
        efw ^= rotr32(e.x, 4 * gI);
        erc ^= rotl32(e.y, 4 * gI);
    }
    u32x2 last = s.tsb[codes >> 30];
    u32x4 a;
    a.x = efw; a.y = erc;
    a.z = rot16(efw ^ last.x);   // role "bases 16..30 of the k-mer that starts one thread to the left"
    a.w = rot16(erc ^ last.y);
    s.ag[v] = a;
    s.codes[v] = codes;
    reinterpret_cast<uint16_t *>(s.inv)[v] = (uint16_t)inv16;
    codes_out = codes; efw_out = efw; erc_out = erc;
}

template <class G, int FLAV, bool PACKED>
DCN_HD void phase_convert(int t, TileSmem<G> &s, TilePriv<G> &pv, const TileSrc &src, uint64_t origin) {
    convert_vector<G, FLAV, PACKED>(t, s, src, origin, pv.c0, pv.efw, pv.erc);
    if (t < G::NV - G::NT) {
        uint32_t a, b, c;
        convert_vector<G, FLAV, PACKED>(G::NT + t, s, src, origin, a, b, c);
    }
    if (t < 6) {  // zero pad words read by the last threads
        s.codes[G::NV + t] = 0;
        if (t < 2) s.inv[(G::NV + 1) / 2 + t] = 0;
    }
}

// ------------------------------------------------------------------ phase 2: rolling ntHash
template <class G>
DCN_HD void phase_hash(int t, TileSmem<G> &s, TilePriv<G> &pv) {
    const uint32_t c0 = pv.c0, c1 = s.codes[t + 1], c2 = s.codes[t + 2];
    u32x4 nb = s.ag[t + 1];
    // k-mer at 16t covers own bases 0..15 and the neighbour's bases 0..K-17.  The aggregates are
    // built for K = 31 (neighbour contributes 15 bases); for K < 31 peel the surplus bases off.
    uint32_t fw = pv.efw ^ nb.z, rc = pv.erc ^ nb.w;
    if (G::K < 31) {
        // aggregates carry rotation (30 - i) for fw; a K-mer needs (K-1-i): rotate right by 31-K.
        // Remove neighbour bases K-16 .. 14 (indices 16+j in k-mer coordinates).
        for (int j = G::K - 16; j < 15; j++) {
            uint32_t c = (c1 >> (2 * j)) & 3u;
            fw ^= rotl32(nt_f(c), (uint32_t)((30 - 16 - j) & 31));
            rc ^= rotl32(nt_f(c ^ 2u), (uint32_t)(16 + j));
        }
        fw = rotr32(fw, 31 - G::K);
    }
    pv.h[0] = fw + rc;
    // step j (k-mer 16t+j -> 16t+j+1) drops base j and takes base K+j.  The two code streams are
    // interleaved once, so that a step's table index (out | in << 2) is one shift + one mask:
    // even steps read nibble j/2 of mE, odd steps nibble (j-1)/2 of mO.
    const uint32_t in = fshr(c1, c2, 2u * (uint32_t)(G::K - 16));        // base K+j at bits 2j
    const uint32_t mE = (c0 & 0x33333333u) | ((in & 0x33333333u) << 2);
    const uint32_t mO = ((c0 >> 2) & 0x33333333u) | (in & 0xCCCCCCCCu);
#pragma unroll
    for (int i = 1; i < 16; i++) {
        const int j = i - 1;
        const uint32_t idx = (((j & 1) ? mO : mE) >> (4 * (j >> 1))) & 15u;
        const u32x2 e = s.tio[idx];
        fw = rotl32(fw, 1) ^ e.x;
        rc = rotr32(rc ^ e.y, 1);
        pv.h[i] = fw + rc;
    }
    // only the upper 16 bits take part in the window comparison: keep them masked
    uint32_t *row = &s.hrow[t * G::HP];
#pragma unroll
    for (int i = 0; i < 16; i++) { pv.h[i] &= 0xFFFF0000u; row[i] = pv.h[i]; }
}

// ------------------------------------------------------------------ phase 3: window minima
DCN_HD uint32_t umin32(uint32_t a, uint32_t b) { return a < b ? a : b; }
DCN_HD uint32_t umax32(uint32_t a, uint32_t b) { return a > b ? a : b; }

// low bytes of four words -> one word
DCN_HD uint32_t pack_low_bytes(uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
#ifdef __CUDA_ARCH__
    return __byte_perm(__byte_perm(a, b, 0x0040), __byte_perm(c, d, 0x0040), 0x5410);
#else
    return (a & 0xFFu) | ((b & 0xFFu) << 8) | ((c & 0xFFu) << 16) | ((d & 0xFFu) << 24);
#endif
}

// 64-bit window of a bit array starting at bit 16*t
DCN_HD uint64_t bits64_at(const uint32_t *arr, int t) {
    int w0 = t >> 1;
    uint32_t sh = (uint32_t)(t & 1) * 16u;
    uint32_t a = arr[w0], b = arr[w0 + 1], c = arr[w0 + 2];
    uint32_t lo = fshr(a, b, sh), hi = fshr(b, c, sh);
    return ((uint64_t)hi << 32) | lo;
}

template <class G>
DCN_HD void phase_slide(int t, TileSmem<G> &s, TilePriv<G> &pv) {
    pv.emask = 0; pv.valid16 = 0;
    pv.rel4[0] = pv.rel4[1] = pv.rel4[2] = pv.rel4[3] = 0;   // thread NT-1 owns no window
    if (t >= G::NT - 1) { s.lastpick[t] = 0xFFFFFFFFu; return; }

    uint32_t hv[30];
#pragma unroll
    for (int i = 0; i < 16; i++) hv[i] = pv.h[i];
    const uint32_t *nrow = &s.hrow[(t + 1) * G::HP];
#pragma unroll
    for (int i = 0; i < 14; i++) hv[16 + i] = nrow[i];

    // left keys: (h >> 16) << 16 | i  -> min = smallest hash, leftmost;  SURVEY A.3 steps 1-2
    uint32_t key[30], oL[16], oR[16];
#pragma unroll
    for (int i = 0; i < 30; i++) key[i] = hv[i] + (uint32_t)i;   // hv is masked: + == |
    {
#pragma unroll
        for (int i = 13; i >= 0; i--) key[i] = umin32(key[i], key[i + 1]);      // suffix min of block A
#pragma unroll
        for (int i = 16; i < 30; i++) key[i] = umin32(key[i], key[i - 1]);      // prefix min of block B
        oL[0] = key[0];
#pragma unroll
        for (int i = 1; i < 15; i++) oL[i] = umin32(key[i], key[14 + i]);
        oL[15] = key[29];
    }
    // right keys: ~(h >> 16) << 16 | i -> max = smallest hash, rightmost;  A.3 step 3
#pragma unroll
    for (int i = 0; i < 30; i++) key[i] = (0xFFFF0000u + (uint32_t)i) - hv[i];   // == (~hv & 0xFFFF0000) | i
    {
#pragma unroll
        for (int i = 13; i >= 0; i--) key[i] = umax32(key[i], key[i + 1]);
#pragma unroll
        for (int i = 16; i < 30; i++) key[i] = umax32(key[i], key[i - 1]);
        oR[0] = key[0];
#pragma unroll
        for (int i = 1; i < 15; i++) oR[i] = umax32(key[i], key[14 + i]);
        oR[15] = key[29];
    }

    // pick positions (low byte of the keys: the index 0..29), left and right flavour, four windows per word.  They differ
    // only where the window minimum is tied (~2e-4 of windows): the strand that chooses between them (A.3 step 4) is only
    // worked out for a thread that holds such a window
    uint32_t l4[4], r4[4], tied = 0;
#pragma unroll
    for (int g = 0; g < 4; g++) {
        l4[g] = pack_low_bytes(oL[4 * g], oL[4 * g + 1], oL[4 * g + 2], oL[4 * g + 3]);
        r4[g] = pack_low_bytes(oR[4 * g], oR[4 * g + 1], oR[4 * g + 2], oR[4 * g + 3]);
        tied |= l4[g] ^ r4[g];
    }
    if (tied) {
        // canonical strand: #(T|G) > #(A|C) over the L bases of the window.  T/G <=> bit 1 of the 2-bit code.
        const uint32_t c0 = pv.c0, c1 = s.codes[t + 1], c2 = s.codes[t + 2], c3 = s.codes[t + 3];
        uint32_t cnt = 0;
        {
            // bases 0 .. L-1
            cnt = popc32(c0 & 0xAAAAAAAAu) + popc32(c1 & 0xAAAAAAAAu);
            constexpr int rem = G::L - 32;  // bases 32 .. L-1 live in c2 (and c3 if L > 48)
            if (rem >= 16) {
                cnt += popc32(c2 & 0xAAAAAAAAu);
                constexpr int rem3 = rem - 16;
                if (rem3 > 0) cnt += popc32(c3 & (rem3 >= 16 ? 0xAAAAAAAAu : (0xAAAAAAAAu & ((1u << (2 * (rem3 & 15))) - 1u))));
            } else if (rem > 0) {
                cnt += popc32(c2 & (0xAAAAAAAAu & ((1u << (2 * (rem & 15))) - 1u)));
            }
        }
        // The 16 windows of the thread, four at a time in byte lanes.  cnt(i) = cnt(0) + sum_{j<=i} (in(j) - out(j)),
        // in(j) = T/G flag of base L-1+j, out(j) = flag of base j-1: the flags sit in the 2-bit lanes of the code
        // words; a multiply spreads four of them into four bytes, a second multiply prefix-sums the bytes.
        constexpr int LW = (G::L - 1) / 16, LS = 2 * ((G::L - 1) % 16);
        const uint32_t xw_lo = ((LW == 0 ? c0 : LW == 1 ? c1 : c2) >> 1) & 0x55555555u;
        const uint32_t xw_hi = ((LW == 0 ? c1 : LW == 1 ? c2 : c3) >> 1) & 0x55555555u;
        const uint32_t xin = fshr(xw_lo, xw_hi, (uint32_t)LS) & ~3u;        // lane j = in(j); lane 0 belongs to cnt(0)
        const uint32_t xout = ((c0 >> 1) & 0x55555555u) << 2;               // lane j = out(j); lane 0 = 0
        constexpr uint32_t THR = (uint32_t)(G::L + 1) / 2;                  // canonical <=> cnt >= THR (L is odd)
#pragma unroll
        for (int g = 0; g < 4; g++) {
            const uint32_t spi = (((xin >> (8 * g)) & 0xFFu) * 0x00041041u) & 0x01010101u;
            const uint32_t spo = (((xout >> (8 * g)) & 0xFFu) * 0x00041041u) & 0x01010101u;
            const uint32_t cnt4 = cnt * 0x01010101u + spi * 0x01010101u - spo * 0x01010101u;   // cnt of windows 4g .. 4g+3
            cnt = cnt4 >> 24;
            const uint32_t canon = ((cnt4 + (0x80u - THR) * 0x01010101u) >> 7) & 0x01010101u;
            const uint32_t msk = canon * 0xFFu;                                                  // 0xFF in the bytes of canonical windows
            l4[g] = (l4[g] & msk) | (r4[g] & ~msk);                                              // A.3 step 4
        }
    }
    uint32_t neq = 0, prev_hi = 0;
#pragma unroll
    for (int g = 0; g < 4; g++) {
        const uint32_t rel = l4[g];
        pv.rel4[g] = rel;
        // pick(i) != pick(i-1): compare every byte with the one before it (window 0 with itself)
        const uint32_t before = (rel << 8) | (g == 0 ? (rel & 0xFFu) : prev_hi);
        const uint32_t xd = rel ^ before;
        const uint32_t nz = ((((xd & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | xd) >> 7) & 0x01010101u;
        neq |= (((nz * 0x00204081u) >> 21) & 0xFu) << (4 * g);
        prev_hi = rel >> 24;
    }
    const uint32_t rel15 = prev_hi;

    // window validity from the record structure: window j is valid iff j is not dead and no
    // break bit lies in (j, j + L - 1].
    uint64_t bw = bits64_at(s.brk, t);
    uint32_t dead16 = (uint32_t)bits64_at(s.dead, t) & 0xFFFFu;
    uint64_t m = bw >> 1;
    // smear right by L-2: bit j of x = OR of m[j .. j+L-2]
    uint64_t x = m;
    {
        int have = 0;  // current smear reach
        // doubling while reach*2+1 <= L-2
        int target = G::L - 2;
        int sh = 1;
        while (have + sh <= target) { x |= x >> sh; have += sh; sh <<= 1; }
        if (have < target) { x |= x >> (target - have); }
    }
    uint32_t invalid16 = ((uint32_t)x | dead16) & 0xFFFFu;
    uint32_t valid16 = ~invalid16 & 0xFFFFu;
    uint32_t first16 = (uint32_t)bw & ~dead16 & 0xFFFFu;      // first window of a record: always emits
    pv.valid16 = valid16;
    // emit(j) = valid(j) && (first(j) || (valid(j-1) && pick(j) != pick(j-1)));  A.3 step 5
    pv.emask = valid16 & (first16 | ((valid16 << 1) & neq));  // bit 0 completed in phase_emit_fix
    s.lastpick[t] = (valid16 & 0x8000u) ? (uint32_t)(16 * t) + rel15 : 0xFFFFFFFFu;
}

// bit 0 of the emit mask needs the last pick of the thread to the left
template <class G>
DCN_HD uint32_t phase_emit_fix(int t, TileSmem<G> &s, TilePriv<G> &pv) {
    if (t < G::NT - 1 && (pv.valid16 & 1u) && !(pv.emask & 1u)) {
        uint32_t lp = t > 0 ? s.lastpick[t - 1] : 0xFFFFFFFFu;
        uint32_t mine = (uint32_t)(16 * t) + (pv.rel4[0] & 0xFFu);
        if (lp != 0xFFFFFFFFu && lp != mine) pv.emask |= 1u;
    }
    return popc32(pv.emask);
}

// ------------------------------------------------------------------ phase 4: compact picks
template <class G>
DCN_HD void phase_emit(int t, TileSmem<G> &s, TilePriv<G> &pv, uint32_t excl, uint32_t total,
                       uint32_t cap = (uint32_t)G::PKCAP) {
    pv.pickoff = excl;
    s.poff[t] = excl;
    s.emk[t] = pv.emask;
    if (t == 0) s.npicks = total;
    if (total > cap) return;   // the caller splits the run and retries
    uint32_t em = pv.emask;
    uint32_t idx = pv.pickoff;
    const uint64_t relA = pv.rel4[0] | ((uint64_t)pv.rel4[1] << 32), relB = pv.rel4[2] | ((uint64_t)pv.rel4[3] << 32);
    while (em) {   // one iteration per emitted pick (no dynamic register indexing: pv stays in registers)
#ifdef __CUDA_ARCH__
        const int i = __ffs((int)em) - 1;
#else
        const int i = __builtin_ctz(em);
#endif
        em &= em - 1;
        const uint32_t rel = (uint32_t)(((i & 8) ? relB : relA) >> (8 * (i & 7))) & 0xFFu;
        s.pk_pos[idx++] = (uint32_t)(16 * t) + rel;
    }
}

// ------------------------------------------------------------------ pick -> canonical k-mer -> xxh3
template <class G>
DCN_HD bool pick_kmer_valid(const TileSmem<G> &s, uint32_t p) {
    // non-ACGT bits of [p, p+K): src/filter_common.rs:275-286 / src/minimizers.rs:157-160
    uint32_t w0 = p >> 5, sh = p & 31u;
    uint32_t a = s.inv[w0], b = s.inv[w0 + 1];
    uint32_t bits = fshr(a, b, sh);
    return (bits & ((G::K >= 32) ? 0xFFFFFFFFu : ((1u << (G::K & 31)) - 1u))) == 0;
}

template <class G>
DCN_HD uint64_t pick_kmer_fw(const TileSmem<G> &s, uint32_t p) {
    uint32_t w0 = p >> 4, sh = 2u * (p & 15u);
    uint32_t a = s.codes[w0], b = s.codes[w0 + 1], c = s.codes[w0 + 2];
    uint32_t lo = fshr(a, b, sh), hi = fshr(b, c, sh);
    uint64_t v = ((uint64_t)hi << 32) | lo;
    return G::K >= 32 ? v : v & ((1ULL << (2 * (G::K & 31))) - 1ULL);
}

template <class G>
DCN_HD uint64_t pick_hash(const TileSmem<G> &s, uint32_t p) {
    uint64_t fw = pick_kmer_fw<G>(s, p);
    uint64_t rc = revcomp_2bit(fw, G::K);
    return xxh3_u64(fw < rc ? fw : rc);
}

// ------------------------------------------------------------------ table probe
DCN_HD Bucket load_bucket(const uint64_t *slots, uint64_t b) {
    Bucket r;
    const uint64_t *p = slots + 4 * b;
#ifdef __CUDA_ARCH__
#if defined(DCN_L2_HINT) && DCN_L2_HINT == 64
    asm volatile("ld.global.nc.L1::no_allocate.L2::64B.v4.u64 {%0,%1,%2,%3}, [%4];"
                 : "=l"(r.k0), "=l"(r.k1), "=l"(r.k2), "=l"(r.k3) : "l"(p));
#elif defined(DCN_L2_HINT) && DCN_L2_HINT == 128
    asm volatile("ld.global.nc.L1::no_allocate.L2::128B.v4.u64 {%0,%1,%2,%3}, [%4];"
                 : "=l"(r.k0), "=l"(r.k1), "=l"(r.k2), "=l"(r.k3) : "l"(p));
#elif defined(DCN_L2_HINT) && DCN_L2_HINT == 1
    asm volatile("ld.global.cg.v4.u64 {%0,%1,%2,%3}, [%4];"
                 : "=l"(r.k0), "=l"(r.k1), "=l"(r.k2), "=l"(r.k3) : "l"(p));
#elif defined(DCN_L2_HINT) && DCN_L2_HINT == 2
    asm volatile("ld.global.nc.L1::no_allocate.L2::evict_first.v4.u64 {%0,%1,%2,%3}, [%4];"
                 : "=l"(r.k0), "=l"(r.k1), "=l"(r.k2), "=l"(r.k3) : "l"(p));
#else
    asm volatile("ld.global.nc.L1::no_allocate.v4.u64 {%0,%1,%2,%3}, [%4];"
                 : "=l"(r.k0), "=l"(r.k1), "=l"(r.k2), "=l"(r.k3) : "l"(p));
#endif
#else
    r.k0 = p[0]; r.k1 = p[1]; r.k2 = p[2]; r.k3 = p[3];
#endif
    return r;
}
// continue a probe whose first bucket has been loaded
DCN_HD bool table_contains_from(const TableView &tv, uint64_t h, uint64_t b, Bucket k) {
    if (h == DCN_EMPTY) return tv.has_empty_key != 0;
    for (;;) {
        if (k.k0 == h || k.k1 == h || k.k2 == h || k.k3 == h) return true;
        // slots of a bucket fill in order (an insert takes the first empty slot and nothing is ever deleted):
        // the bucket has room, i.e. the probe sequence ends here, iff its last slot is empty
        if (k.k3 == DCN_EMPTY) return false;
        if (++b == tv.n_buckets) b = 0;
        k = load_bucket(tv.slots, b);
    }
}
DCN_HD bool table_contains(const TableView &tv, uint64_t h) {
    uint64_t b = table_bucket(h, tv.n_buckets);
    return table_contains_from(tv, h, b, load_bucket(tv.slots, b));
}

// set bits [a, b) of a 32-bit-word bit array (shared memory on device)
DCN_HD void set_bits(uint32_t *arr, uint32_t a, uint32_t b) {
    while (a < b) {
        uint32_t w = a >> 5, lo = a & 31u;
        uint32_t end = (w + 1) << 5;
        if (end > b) end = b;
        uint32_t width = end - a;
        uint32_t mask = width == 32 ? 0xFFFFFFFFu : (((1u << width) - 1u) << lo);
#ifdef __CUDA_ARCH__
        atomicOr(&arr[w], mask);
#else
        arr[w] |= mask;
#endif
        a = end;
    }
}

DCN_HD void set_bit(uint32_t *arr, uint32_t p) {
#ifdef __CUDA_ARCH__
    atomicOr(&arr[p >> 5], 1u << (p & 31u));
#else
    arr[p >> 5] |= 1u << (p & 31u);
#endif
}

// bits of the 32-position word that starts at `lo` whose position is < a or >= b
DCN_HD uint32_t outside_mask(uint32_t lo, uint32_t a, uint32_t b) {
    uint32_t m = 0;
    if (lo < a) m = (a - lo >= 32u) ? 0xFFFFFFFFu : ((1u << (a - lo)) - 1u);
    if (lo + 32u > b) m |= (b <= lo) ? 0xFFFFFFFFu : (0xFFFFFFFFu << (b - lo));
    return m;
}
// Start state of the per-position bit arrays, written by all threads (one word each, no atomics):
// windows may only start in [dead_lo, dead_hi), sequence only exists in [brk_lo, brk_hi).
template <class G>
DCN_HD void init_structure_words(int t, TileSmem<G> &s, uint32_t dead_lo, uint32_t dead_hi, uint32_t brk_lo, uint32_t brk_hi) {
    for (int i = t; i < G::NBW + 2; i += G::NT) {
        s.dead[i] = outside_mask(32u * (uint32_t)i, dead_lo, dead_hi);
        s.brk[i] = outside_mask(32u * (uint32_t)i, brk_lo, brk_hi);
    }
}

// ------------------------------------------------------------------ filter parameters
// B3 (extraction only) through the tile pipeline: a tile reserves a block of the temp arrays for its
// picks, the per-record valid counts go through a scan, and a copy kernel compacts the blocks into CSR.
struct ExtractOut {
    uint64_t *tmp_h;                 // xxh3 of each pick (block of the tile, list order)
    uint32_t *tmp_p;                 // position in the record's effective sequence | valid << 31
    uint64_t tmp_cap;
    unsigned long long *cursor;      // next free temp entry (may run past tmp_cap: the host retries)
    uint64_t *rec_cnt;               // valid picks of each record
    uint64_t *rec_tmp;               // temp start of each record << 16 | picks (valid or not)
};

struct FilterParams {
    const uint8_t *bases;     // concatenated ASCII records (device), 16-byte aligned; nullptr when packed
    const uint32_t *pk_codes; // host-packed form (see TileSrc): 2-bit codes ...
    const uint16_t *pk_inv;   // ... and non-ACGT bits; both cover [base0, n_bases) rounded up to 16
    const uint32_t *nl_bits;  // packed form only: bit nl_bit0 + r = record r ends its effective prefix in '\n'
    uint32_t nl_bit0;
    uint64_t base0;           // absolute offset of bases[0] (multiple of 16); rec_off is absolute
    uint64_t n_bases;         // absolute end offset: bases[x - base0] is readable for base0 <= x < n_bases
    const uint64_t *rec_off;  // n_rec + 1 offsets (device)
    uint32_t n_rec;
    uint32_t rpu;             // records per unit: 1 (single) or 2 (pair: records 2i, 2i+1)
    uint32_t n_units;
    uint32_t prefix_len;      // src/filter_common.rs:222-226
    uint32_t abs_thr;
    double rel_thr;
    int deplete;
    TableView table;
    uint8_t *keep;            // per unit
    uint32_t *hits;
    uint32_t *total;
    ExtractOut xo;            // extraction mode only
};

// effective length of a record (src/filter_common.rs:217-229): raw-length guard, prefix, one '\n'.
// `r` is the record's index in P.rec_off, `gs` its start relative to base0.
template <class G>
DCN_HD uint64_t filter_eff_len(const FilterParams &P, uint32_t r, uint64_t gs, uint64_t len) {
    if (len < (uint64_t)G::K) return 0;
    uint64_t n = (P.prefix_len > 0 && len > P.prefix_len) ? P.prefix_len : len;
    const uint32_t nb = P.nl_bit0 + r;
    const bool nl = P.bases ? P.bases[gs + n - 1] == (uint8_t)'\n' : (P.nl_bits && ((P.nl_bits[nb >> 5] >> (nb & 31u)) & 1u) != 0);
    return nl ? n - 1 : n;
}
DCN_HD TileSrc filter_src(const FilterParams &P) {
    TileSrc src;
    src.bases = P.bases; src.codes = P.pk_codes; src.inv = P.pk_inv; src.n_bases = P.n_bases - P.base0;
    return src;
}

// ------------------------------------------------------------------ short-unit tile driver
// Units [u_begin, u_end) are whole (record or pair) and fit the tile: every base of every unit
// lies in [origin, origin + BCAP).  Ex provides par(f) = run f for every thread then barrier,
// scan(get, put) = block-wide exclusive sum, match64 / ballot = warp votes.
// Barriers per tile: 7.  The last phase has no trailing barrier: the first phase of the next
// tile touches none of the arrays it reads.
// Record boundaries of the tile: loaded in the same phase as the base vectors (so the dependent
// rec_off -> last-byte loads overlap the tile's own DRAM latency) ...
template <class G>
DCN_HD void phase_structure_load(int t, TileSmem<G> &s, const FilterParams &P, uint32_t r_begin, uint32_t n_rec_t,
                                 uint64_t origin) {
    for (uint32_t i = (uint32_t)t; i < n_rec_t; i += G::NT) {
        uint64_t gs = P.rec_off[r_begin + i] - P.base0, ge = P.rec_off[r_begin + i + 1] - P.base0;
        uint32_t sL = (uint32_t)(gs - origin), eL = (uint32_t)(ge - origin);
        uint32_t eff = sL + (uint32_t)filter_eff_len<G>(P, r_begin + i, gs, eL - sL);
        s.rec_se[i] = sL | (eL << 16);
        s.rec_eff[i] = (uint16_t)eff;
    }
}
// ... and turned into the per-position bit arrays one phase later.
template <class G>
DCN_HD void phase_structure(int t, TileSmem<G> &s, uint32_t rpu, uint32_t n_rec_t) {
    for (uint32_t i = (uint32_t)t; i < n_rec_t; i += G::NT) {
        const uint32_t se = s.rec_se[i];
        const uint32_t sL = se & 0xFFFFu, eL = se >> 16, eff = s.rec_eff[i];
        set_bit(s.brk, sL);
        if (i % rpu == 0) s.ustartpos[i / rpu] = (uint16_t)sL;
        if (eff < eL) { set_bits(s.dead, eff, eL); set_bits(s.brk, eff, eL); }
        // positions before the first record and after the last one were marked by init_structure_words
    }
}

// first pick index of every unit: the picks emitted by windows that start before the unit's first base
template <class G>
DCN_HD void phase_unit_first(int t, TileSmem<G> &s, uint32_t n_units_t, uint32_t npicks) {
    for (uint32_t u = (uint32_t)t; u < n_units_t; u += G::NT) {
        uint32_t pos = s.ustartpos[u];
        uint32_t tt = pos >> 4, ii = pos & 15u;
        // a unit that starts in the last 16-position chunk or beyond has no window in this tile
        bool tail = tt >= (uint32_t)(G::NT - 1);
        s.ufirst[u] = (uint16_t)(tail ? npicks : s.poff[tt] + popc32(s.emk[tt] & ((1u << ii) - 1u)));
        if (u == 0) s.ufirst[n_units_t] = (uint16_t)npicks;
    }
}

// Returns false (after one barrier-consistent decision, nothing written) when the run emits more
// than PKCAP picks: the caller then splits the run.  A single short unit (<= 1024 bases) never does.
enum TileMode { MODE_FILTER = 0, MODE_EXTRACT = 1 };

template <class G, bool PACKED, int MODE, class Ex>
DCN_HD bool filter_short_tile(Ex &ex, TileSmem<G> &s, const FilterParams &P, uint32_t u_begin, uint32_t u_end) {
    using Priv = TilePriv<G>;
    const uint32_t r_begin = u_begin * P.rpu;
    const uint32_t n_rec_t = (u_end - u_begin) * P.rpu;
    const uint32_t n_units_t = u_end - u_begin;
    const uint64_t first_start = P.rec_off[r_begin] - P.base0;
    const uint64_t origin = first_start & ~15ull;   // relative to base0, like every offset below
    const TileSrc src = filter_src(P);
    const uint32_t span_lo = (uint32_t)(first_start - origin);
    const uint32_t span_hi = (uint32_t)(P.rec_off[r_begin + n_rec_t] - P.base0 - origin);

    ex.par([&](int t, Priv &pv) {
        init_structure_words<G>(t, s, span_lo, span_hi, span_lo, span_hi);
        phase_structure_load<G>(t, s, P, r_begin, n_rec_t, origin);
        phase_convert<G, FLAVOUR_FILTER, PACKED>(t, s, pv, src, origin);
    });
    ex.par([&](int t, Priv &pv) {
        phase_structure<G>(t, s, P.rpu, n_rec_t);
        phase_hash<G>(t, s, pv);
    });
    ex.par([&](int t, Priv &pv) {
        ex.midtile_prefetch(t);
        phase_slide<G>(t, s, pv);
    });
    ex.scan([&](int t, Priv &pv) { return phase_emit_fix<G>(t, s, pv); },
            [&](int t, Priv &pv, uint32_t excl, uint32_t total) { phase_emit<G>(t, s, pv, excl, total); });
    const uint32_t npicks = s.npicks;
    if (npicks > (uint32_t)G::PKCAP) { ex.barrier(); return false; }

    if (MODE == MODE_EXTRACT) {
        // B3: hash every pick (no probe); thread 0 reserves the tile's block of the temp arrays
        ex.par([&](int t, Priv &) {
            phase_unit_first<G>(t, s, n_units_t, npicks);
            if (t == 0) {
                const uint64_t base = ex.global_add64(P.xo.cursor, npicks);
                s.wsum[10] = (uint32_t)base; s.wsum[11] = (uint32_t)(base >> 32);
            }
            for (uint32_t idx = (uint32_t)t; idx < npicks; idx += G::NT) {
                const uint32_t pp = s.pk_pos[idx];
                if (pick_kmer_valid<G>(s, pp & 0xFFFFu)) {
                    s.pk_hash[idx] = pick_hash<G>(s, pp & 0xFFFFu);
                    s.pk_pos[idx] = pp | 0x80000000u;
                }
            }
        });
        // one warp per record: its picks (list order = position order) go to the tile's block, with the
        // position made relative to the record; the valid count feeds the CSR offsets
        ex.par_nosync([&](int t, Priv &) {
            const uint32_t lane = (uint32_t)t & 31u;
            const uint64_t tbase = (uint64_t)s.wsum[10] | ((uint64_t)s.wsum[11] << 32);
            for (uint32_t u = (uint32_t)t >> 5; u < n_units_t; u += (uint32_t)G::NT / 32u) {
                const uint32_t a = s.ufirst[u], b = s.ufirst[u + 1], sL = s.ustartpos[u];
                uint32_t cnt = 0;
                for (uint32_t base = a; base < b; base += 32u) {
                    const uint32_t idx = base + lane;
                    bool valid = false;
                    if (idx < b) {
                        const uint32_t pp = s.pk_pos[idx];
                        valid = (pp & 0x80000000u) != 0;
                        const uint64_t at = tbase + idx;
                        if (at < P.xo.tmp_cap) {
                            P.xo.tmp_p[at] = ((pp & 0xFFFFu) - sL) | (pp & 0x80000000u);
                            if (valid) P.xo.tmp_h[at] = s.pk_hash[idx];
                        }
                    }
                    cnt += popc32(ex.ballot(t, valid));
                }
                if (lane == 0) {
                    P.xo.rec_cnt[u_begin + u] = cnt;
                    P.xo.rec_tmp[u_begin + u] = ((tbase + a) << 16) | (uint64_t)(b - a);
                }
            }
        });
        return true;
    }

    // hash every pick and probe the table: two picks per thread are in flight at a time (hash A,
    // request A, hash B, request B, then test A and B), the answer is kept as bit 30 of the pick;
    // unit -> pick-range tables on the side.  (Staging the buckets in shared memory with cp.async so
    // that the duplicate test overlaps the HBM latency was measured 7 % slower: round-1 notes.)
    ex.par([&](int t, Priv &) {
        phase_unit_first<G>(t, s, n_units_t, npicks);
        for (uint32_t idx = (uint32_t)t; idx < npicks; idx += 2 * G::NT) {
            const uint32_t idxB = idx + G::NT;
            uint32_t ppA = s.pk_pos[idx], ppB = idxB < npicks ? s.pk_pos[idxB] : 0u;
            const bool vA = pick_kmer_valid<G>(s, ppA & 0xFFFFu);
            const bool vB = idxB < npicks && pick_kmer_valid<G>(s, ppB & 0xFFFFu);
            uint64_t hA = 0, hB = 0, bA = 0, bB = 0;
            Bucket kA, kB;
            kA.k0 = kA.k1 = kA.k2 = kA.k3 = 0; kB = kA;
            if (vA) {
                hA = pick_hash<G>(s, ppA & 0xFFFFu);
                bA = table_bucket(hA, P.table.n_buckets);
                kA = load_bucket(P.table.slots, bA);
            }
            if (vB) {
                hB = pick_hash<G>(s, ppB & 0xFFFFu);
                bB = table_bucket(hB, P.table.n_buckets);
                kB = load_bucket(P.table.slots, bB);
            }
            if (vA) {
                s.pk_hash[idx] = hA;
                s.pk_pos[idx] = ppA | 0x80000000u | (table_contains_from(P.table, hA, bA, kA) ? 0x40000000u : 0u);
            }
            if (vB) {
                s.pk_hash[idxB] = hB;
                s.pk_pos[idxB] = ppB | 0x80000000u | (table_contains_from(P.table, hB, bB, kB) ? 0x40000000u : 0u);
            }
        }
    });
    // Distinct hits per unit (src/filter_common.rs:143-145: contains && seen.insert) and the threshold
    // test, one warp per unit: the unit's picks are consecutive list entries, 32 per pass.  A pick is a
    // duplicate iff an EARLIER valid pick of the same unit has the same hash: within a pass one warp
    // match answers that; a pick of a later pass (units with more than 32 picks) is compared with the
    // previous pass through a broadcast per candidate, with older passes by a scan of the list.
    // No trailing barrier: the first phase of the next tile touches none of the arrays read here.
    ex.par_nosync([&](int t, Priv &) {
        const uint32_t lane = (uint32_t)t & 31u, lt = (1u << lane) - 1u;
        for (uint32_t u = (uint32_t)t >> 5; u < n_units_t; u += (uint32_t)G::NT / 32u) {
            const uint32_t a = s.ufirst[u], b = s.ufirst[u + 1];
            uint32_t hits = 0, total = 0, vprev = 0;
            uint64_t hprev = 0;
            for (uint32_t base = a; base < b; base += 32u) {
                const uint32_t idx = base + lane;
                bool valid = false, found = false;
                uint64_t h = 0;
                if (idx < b) {
                    const uint32_t pp = s.pk_pos[idx];
                    valid = (pp & 0x80000000u) != 0;
                    found = (pp & 0x40000000u) != 0;
                    if (valid) h = s.pk_hash[idx];
                }
                const uint32_t vmask = ex.ballot(t, valid);
                const uint32_t same = ex.match64(t, h, valid);
                bool fresh = valid && found && (same & vmask & lt) == 0;
                if (base > a) {
                    uint32_t cand = ex.ballot(t, fresh);
                    while (cand) {   // warp-uniform: one round per candidate of this pass
                        const uint32_t l = popc32((cand & (0u - cand)) - 1u);
                        cand &= cand - 1u;
                        const uint64_t hv = ex.bcast64(t, h, l);
                        const uint32_t hitprev = ex.ballot(t, ((vprev >> lane) & 1u) != 0 && hprev == hv);
                        if (lane == l && hitprev) fresh = false;
                    }
                    if (fresh && base > a + 32u) {   // passes before the previous one
                        const uint32_t hlo = (uint32_t)h;
                        const uint32_t *h32 = reinterpret_cast<const uint32_t *>(s.pk_hash);
                        for (uint32_t j = a; j < base - 32u && fresh; j++)
                            if (h32[2 * j] == hlo && s.pk_hash[j] == h && (s.pk_pos[j] & 0x80000000u)) fresh = false;
                    }
                }
                hits += popc32(ex.ballot(t, fresh));
                total += popc32(vmask);
                hprev = h; vprev = vmask;
            }
            if (lane == 0) {
                const uint32_t gu = u_begin + u;
                P.total[gu] = total;
                P.hits[gu] = hits;
                bool keep;
                if (total < 256u) { const uint32_t req = s.req[total]; keep = P.deplete ? hits < req : hits >= req; }
                else keep = meets_criteria(hits, total, P.abs_thr, P.rel_thr, P.deplete);
                P.keep[gu] = keep ? 1 : 0;
            }
        }
    });
    return true;
}

// a run of short units, split in halves while it emits more picks than one pass can hold
template <class G, bool PACKED, int MODE, class Ex>
DCN_HD void filter_short_run(Ex &ex, TileSmem<G> &s, const FilterParams &P, uint32_t u_begin, uint32_t u_end) {
    uint32_t lo = u_begin;
    uint32_t span = u_end - u_begin;
    while (lo < u_end) {
        uint32_t hi = lo + span < u_end ? lo + span : u_end;
        if (filter_short_tile<G, PACKED, MODE>(ex, s, P, lo, hi)) {
            if (hi < u_end) ex.barrier();  // the next pass rewrites tables the last phase still reads
            lo = hi;
        } else {
            span = (hi - lo + 1) / 2;      // hi - lo >= 2 here: one short unit always fits
        }
    }
}

// All units whose first base lies in one tile: runs of short units go through
// filter_short_tile; long units are skipped here (they are cut into chunks by the long path).
template <class G, bool PACKED, int MODE, class Ex>
DCN_HD void filter_tile(Ex &ex, TileSmem<G> &s, const FilterParams &P, const PlanCfg &cfg, uint32_t n_long,
                        uint32_t u_first, uint32_t u_end) {
    const uint64_t rpu = P.rpu;
    // common case: the batch has no long unit and the tile's records fit one pass
    if (n_long == 0 && (uint64_t)(u_end - u_first) * rpu <= (uint64_t)G::MAXR) {
        filter_short_run<G, PACKED, MODE>(ex, s, P, u_first, u_end);
        return;
    }
    uint32_t u = u_first;
    bool any = false;
    while (u < u_end) {
        uint64_t len = P.rec_off[(uint64_t)(u + 1) * rpu] - P.rec_off[(uint64_t)u * rpu];
        if (len > cfg.max_short) { u++; continue; }
        uint32_t v = u + 1;
        while (v < u_end && (uint64_t)(v - u + 1) * rpu <= (uint64_t)G::MAXR &&
               P.rec_off[(uint64_t)(v + 1) * rpu] - P.rec_off[(uint64_t)v * rpu] <= cfg.max_short)
            v++;
        filter_short_run<G, PACKED, MODE>(ex, s, P, u, v);
        ex.barrier();  // a following pass rewrites the tables the last phase still reads
        u = v;
        any = true;
    }
    // a tile of long units only: the caller's tile-claim slots (filter_fused_kernel) rely on at least one barrier per tile
    if (!any) ex.barrier();
}

// =====================================================================================
// Chunked path: one record cut into chunks of CSTRIDE window starts.  Used for units longer than
// DCN_MAX_SHORT in the filter ("long path", any chunk on any CTA; distinct hits are established
// through a global (hash, unit) set) and for every record of an index build.
// =====================================================================================
struct ChunkDesc { uint32_t rec, chunk; };

template <class G>
struct ChunkGeo {
    // a chunk loads from a 16-byte aligned origin (<= 15 bases of slack) and computes one extra
    // "carry" window before its first own window, so that the consecutive-duplicate rule
    //  sees its predecessor
    static constexpr uint32_t CSTRIDE = (uint32_t)G::WCAP - 16u;
    static constexpr uint32_t PICKCAP = (uint32_t)(G::NT * G::HP);   // pk_pos capacity when it owns all of hrow
    static_assert(PICKCAP >= CSTRIDE + 1, "every window of a chunk may emit");
};

// number of chunks of a record with `eff_len` effective bases
template <class G>
DCN_HD uint32_t chunks_of(uint64_t eff_len) {
    if (eff_len < (uint64_t)G::L) return 0;
    uint64_t nwin = eff_len - (uint64_t)G::L + 1;
    return (uint32_t)((nwin + ChunkGeo<G>::CSTRIDE - 1) / ChunkGeo<G>::CSTRIDE);
}

// Runs the per-position phases for chunk `c` of the record whose effective sequence is
// [gs, gs + eff_len) (offsets relative to base0).  On return pk_pos[0 .. npicks) holds the picks'
// local positions and *origin_out the offset of local position 0.  Returns npicks.
template <class G, int FLAV, bool PACKED, class Ex>
DCN_HD uint32_t chunk_picks(Ex &ex, TileSmem<G> &s, const TileSrc &src, uint64_t gs,
                            uint64_t eff_len, uint32_t c, uint64_t *origin_out) {
    using Priv = TilePriv<G>;
    const uint64_t nwin = eff_len - (uint64_t)G::L + 1;
    const uint64_t w0 = (uint64_t)c * ChunkGeo<G>::CSTRIDE;
    const uint32_t nw = (uint32_t)(nwin - w0 < ChunkGeo<G>::CSTRIDE ? nwin - w0 : ChunkGeo<G>::CSTRIDE);
    const uint32_t carry = c > 0 ? 1u : 0u;
    const uint64_t a = gs + w0 - carry;          // first window computed by this chunk
    const uint64_t origin = a & ~15ull;
    const uint32_t la = (uint32_t)(a - origin);
    const uint64_t eff_end = gs + eff_len;
    *origin_out = origin;

    const uint64_t e64 = eff_end - origin;                           // bases beyond the effective sequence
    const uint32_t e = e64 < (uint64_t)((G::NBW + 2) * 32) ? (uint32_t)e64 : (uint32_t)((G::NBW + 2) * 32);
    ex.par([&](int t, Priv &pv) {
        // windows of this chunk start in [la, la + carry + nw); later ones belong to later chunks
        init_structure_words<G>(t, s, la, la + carry + nw, la, e);
        phase_convert<G, FLAV, PACKED>(t, s, pv, src, origin);
    });
    ex.par([&](int t, Priv &pv) {
        if (t == 0) {
            s.wsum[8] = 0; s.wsum[9] = 0;                            // per-chunk tallies of the consumer phase
            if (!carry) set_bit(s.brk, la);                          // record start: first window always emits
        }
        phase_hash<G>(t, s, pv);
    });
    ex.par([&](int t, Priv &pv) { phase_slide<G>(t, s, pv); });
    ex.scan([&](int t, Priv &pv) { return phase_emit_fix<G>(t, s, pv); },
            [&](int t, Priv &pv, uint32_t excl, uint32_t total) { phase_emit<G>(t, s, pv, excl, total, ChunkGeo<G>::PICKCAP); });
    return s.npicks;
}

// ------------------------------------------------------------------ global (hash, unit) set
// 16-byte entries {hash, epoch << 32 | unit + 1}; all-zero = empty.  Exact: the full key is stored.  The set is never
// cleared between calls: every call inserts under a new EPOCH and an entry of another epoch counts as free (it is
// replaced by a second compare-and-swap), so the long path no longer pays a memset of 4 bytes per long base per call.
struct DedupView {
    unsigned __int128 *slots;
    uint64_t cap;         // entries (any size: the start slot is mulhi(mix, cap))
    uint32_t *overflow;   // set when an insert gives up
    uint32_t epoch;       // >= 1
    // 0: one set for the whole launch (the slot of (hash, unit) is anywhere in it).  n > 0: every unit has a REGION of the
    // set of its own, n slots per 16 bases of the unit, at the unit's place in the batch (dedup_region): the entries of a
    // long read then sit in a few tens of KB that stay in L2 while its chunks are being processed, instead of costing one
    // random DRAM request per hit
    uint32_t per16;
};

// region of the set that belongs to the unit spanning bases [ub, ue) of the launch (relative to base0)
DCN_HD void dedup_region(const DedupView &d, uint64_t ub, uint64_t ue, uint64_t &lo, uint32_t &sz) {
    lo = (ub >> 4) * d.per16;
    const uint64_t n = ((ue >> 4) - (ub >> 4)) * d.per16;
    sz = n > 0xFFFFFFFFull ? 0xFFFFFFFFu : (uint32_t)n;
}

// reg_sz > 0: the unit's own region [reg_lo, reg_lo + reg_sz) of the set (linear probing wraps inside it)
DCN_HD bool dedup_insert(const DedupView &d, uint64_t h, uint32_t unit, uint64_t reg_lo = 0, uint32_t reg_sz = 0) {
    const unsigned __int128 val = ((unsigned __int128)(((uint64_t)d.epoch << 32) | ((uint64_t)unit + 1)) << 64) | h;
    const uint64_t lo = reg_sz ? reg_lo : 0, hi = reg_sz ? reg_lo + reg_sz : d.cap;
    uint64_t slot = reg_sz ? reg_lo + mulhi64(h, (uint64_t)reg_sz) : mulhi64(h ^ ((uint64_t)unit * 0x9E3779B97F4A7C15ULL), d.cap);
    const uint32_t max_probes = reg_sz && reg_sz < 4096u ? reg_sz : 4096u;
    for (uint32_t probes = 0; probes < max_probes; probes++) {
#ifdef __CUDA_ARCH__
        unsigned __int128 old = atomicCAS(&d.slots[slot], (unsigned __int128)0, val);
#else
        unsigned __int128 old = d.slots[slot];
        if (old == 0) d.slots[slot] = val;
#endif
        if (old == 0) return true;
        if (old == val) return false;
        if ((uint32_t)(old >> 96) != d.epoch) {   // left by an earlier call: free
#ifdef __CUDA_ARCH__
            const unsigned __int128 old2 = atomicCAS(&d.slots[slot], old, val);
#else
            const unsigned __int128 old2 = d.slots[slot];
            if (old2 == old) d.slots[slot] = val;
#endif
            if (old2 == old) return true;
            if (old2 == val) return false;   // the same (hash, unit) got there first
            // somebody else's entry of this call took the slot meanwhile: occupied
        }
        if (++slot == hi) slot = lo;
    }
    *d.overflow = 1;
    return false;
}

// ------------------------------------------------------------------ long path of the filter
template <class G, bool PACKED, class Ex>
DCN_HD void filter_long_chunk(Ex &ex, TileSmem<G> &s, const FilterParams &P, const DedupView &dd, ChunkDesc cd) {
    using Priv = TilePriv<G>;
    const uint64_t gs = P.rec_off[cd.rec] - P.base0;
    const uint64_t len = P.rec_off[cd.rec + 1] - P.base0 - gs;
    const uint64_t eff_len = filter_eff_len<G>(P, cd.rec, gs, len);
    const uint32_t unit = cd.rec / P.rpu;
    uint64_t origin;
    const uint32_t npicks = chunk_picks<G, FLAVOUR_FILTER, PACKED>(ex, s, filter_src(P), gs, eff_len, cd.chunk, &origin);
    ex.par([&](int t, Priv &) {
        // two picks per thread in flight (hash A, request A, hash B, request B, then test and record A and B)
        const uint32_t rounds = (npicks + G::NT - 1) / G::NT;
        for (uint32_t r = 0; r < rounds; r += 2) {
            const uint32_t idxA = r * G::NT + (uint32_t)t, idxB = idxA + G::NT;
            bool vA = false, vB = false;
            uint64_t hA = 0, hB = 0, bA = 0, bB = 0;
            Bucket kA, kB;
            kA.k0 = kA.k1 = kA.k2 = kA.k3 = 0; kB = kA;
            if (idxA < npicks) {
                const uint32_t p = s.pk_pos[idxA] & 0xFFFFu;
                vA = pick_kmer_valid<G>(s, p);
                if (vA) { hA = pick_hash<G>(s, p); bA = table_bucket(hA, P.table.n_buckets); kA = load_bucket(P.table.slots, bA); }
            }
            if (idxB < npicks) {
                const uint32_t p = s.pk_pos[idxB] & 0xFFFFu;
                vB = pick_kmer_valid<G>(s, p);
                if (vB) { hB = pick_hash<G>(s, p); bB = table_bucket(hB, P.table.n_buckets); kB = load_bucket(P.table.slots, bB); }
            }
            const bool fA = vA && table_contains_from(P.table, hA, bA, kA) && dedup_insert(dd, hA, unit);
            const bool fB = vB && table_contains_from(P.table, hB, bB, kB) && dedup_insert(dd, hB, unit);
            ex.tally2(t, vA, fA, &s.wsum[8], &s.wsum[9]);
            ex.tally2(t, vB, fB, &s.wsum[8], &s.wsum[9]);
        }
    });
    ex.par([&](int t, Priv &) {
        if (t == 0) {
            ex.global_add(&P.total[unit], s.wsum[8]);
            ex.global_add(&P.hits[unit], s.wsum[9]);
        }
    });
}

// ------------------------------------------------------------------ index-flavour extraction
struct IndexParams {
    const uint8_t *bases;
    uint64_t base0, n_bases;
    const uint64_t *rec_off;
    uint32_t n_rec;
    const uint32_t *entropy_pass;  // bitmap over (cA, cC, cG) or nullptr when the threshold is 0
    uint64_t *out;                 // hashes, unordered (duplicates are removed by the sort/unique that follows)
    uint64_t out_cap;
    unsigned long long *out_count; // cursor; may exceed out_cap (overflow is detected by the host)
};

// src/minimizers.rs:73-121 reduced to a table: the scaled entropy of an all-ACGT k-mer depends
// only on its base counts.  v = forward 2-bit k-mer (A=0 C=1 T=2 G=3).
template <class G>
DCN_HD bool entropy_ok(const uint32_t *pass, uint64_t v) {
    if (!pass) return true;
    const uint64_t M = 0x5555555555555555ULL;
    uint64_t lo = v & M, hi = (v >> 1) & M;
    uint32_t nC = popc32((uint32_t)(lo & ~hi)) + popc32((uint32_t)((lo & ~hi) >> 32));
    uint32_t nG = popc32((uint32_t)(lo & hi)) + popc32((uint32_t)((lo & hi) >> 32));
    uint32_t nT = popc32((uint32_t)(~lo & hi)) + popc32((uint32_t)((~lo & hi) >> 32));
    uint32_t nA = (uint32_t)G::K - nC - nG - nT;
    uint32_t idx = (nA * 32u + nC) * 32u + nG;   // counts <= K <= 31
    return (pass[idx >> 5] >> (idx & 31u)) & 1u;
}

template <class G, class Ex>
DCN_HD void index_chunk(Ex &ex, TileSmem<G> &s, const IndexParams &P, ChunkDesc cd) {
    using Priv = TilePriv<G>;
    const uint64_t gs = P.rec_off[cd.rec] - P.base0;
    const uint64_t len = P.rec_off[cd.rec + 1] - P.base0 - gs;
    const uint64_t eff_len = len < (uint64_t)G::K ? 0 : len;   // src/minimizers.rs:135-137
    uint64_t origin;
    TileSrc src;
    src.bases = P.bases; src.codes = nullptr; src.inv = nullptr; src.n_bases = P.n_bases - P.base0;
    const uint32_t npicks = chunk_picks<G, FLAVOUR_INDEX, false>(ex, s, src, gs, eff_len, cd.chunk, &origin);
    ex.par([&](int t, Priv &) {
        const uint32_t rounds = (npicks + G::NT - 1) / G::NT;
        for (uint32_t r = 0; r < rounds; r++) {
            const uint32_t idx = r * G::NT + (uint32_t)t;
            bool valid = false;
            uint64_t h = 0;
            if (idx < npicks) {
                uint32_t p = s.pk_pos[idx] & 0xFFFFu;
                if (pick_kmer_valid<G>(s, p)) {
                    uint64_t fw = pick_kmer_fw<G>(s, p);
                    if (entropy_ok<G>(P.entropy_pass, fw)) {
                        uint64_t rc = revcomp_2bit(fw, G::K);
                        h = xxh3_u64(fw < rc ? fw : rc);
                        valid = true;
                    }
                }
            }
            ex.append64(t, valid, h, P.out, P.out_cap, P.out_count);
        }
    });
}

}  // namespace dcn
