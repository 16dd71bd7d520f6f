// dcn_emu.cpp -- TEST-ONLY host emulation of the CUDA tile pipeline.
//
// Compiles deacon_server_b200/csrc/dcn_tile.cuh with g++ and runs every phase as a loop over the
// 256 "threads" of a CTA, so the kernel's logic (bit tricks, rolling hash seeding, van Herk
// window minima, pick compaction, per-unit distinct-hit count) can be checked against the oracle
// on a machine without a GPU.  It is never part of the product library and bench.py never loads it.
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <vector>

static uint32_t emu_hcap = 0xFFFFFFFFu;   // test hook: capacity of the long path's compacted hit list (dcn_warp.cuh)
#define DCN_EMU_HCAP emu_hcap

#include "../../deacon_server_b200/csrc/dcn_plan.cuh"
#include "../../deacon_server_b200/csrc/dcn_tile.cuh"
#include "../../deacon_server_b200/csrc/dcn_warp.cuh"
#include "../../deacon_server_b200/csrc/dcn_generic.cuh"
#include "../../deacon_server_b200/csrc/dcn_host_pack.h"

using namespace dcn;

template <class G, class PrivT = TilePriv<G>>
struct HostExec {
    std::vector<PrivT> pv;
    HostExec() : pv(G::NT) {}
    int after_scan_calls = 0;
    void after_scan(bool go) { after_scan_calls += go ? 1 : 0; }
    unsigned long long *xcursor = nullptr;
    uint64_t xalloc(uint32_t n) { uint64_t at = *xcursor; *xcursor += n; return at; }
    uint64_t tally_n = 0, tally_bp = 0, tally_n_kept = 0, tally_bp_kept = 0;
    void tally(uint32_t nrec, uint32_t len, bool keep) { tally_n += nrec; tally_bp += len; if (keep) { tally_n_kept += nrec; tally_bp_kept += len; } }

    // ---- warp votes.  A phase runs thread after thread here, so a vote cannot see the other
    // lanes' operands of the same pass.  The phase is therefore re-run from a snapshot until the
    // recorded operands stop changing: pass n answers vote k from the operands recorded in pass
    // n-1, which are final for every vote whose inputs depend on fewer than n earlier votes.
    std::vector<uint64_t> cur_v, prev_v;
    std::vector<uint32_t> calls;
    size_t vote_slot(int t) {
        size_t c = calls[t]++;
        size_t slot = c * G::NT + (size_t)t;
        if (cur_v.size() <= slot) cur_v.resize(slot + 16 * G::NT, 0);
        return slot;
    }
    uint64_t prev_at(size_t slot) const { return slot < prev_v.size() ? prev_v[slot] : 0; }
    uint32_t ballot(int t, bool p) {
        size_t slot = vote_slot(t);
        cur_v[slot] = p;
        uint32_t m = 0;
        size_t w = slot - (size_t)(t & 31);
        for (int l = 0; l < 32; l++) m |= (uint32_t)(prev_at(w + l) & 1) << l;
        return m;
    }
    uint64_t bcast64(int t, uint64_t v, uint32_t src) {
        size_t slot = vote_slot(t);
        cur_v[slot] = v;
        return prev_at(slot - (size_t)(t & 31) + src);
    }
    uint32_t match64(int t, uint64_t v, bool) {
        size_t slot = vote_slot(t);
        cur_v[slot] = v;
        uint32_t m = 0;
        size_t w = slot - (size_t)(t & 31);
        for (int l = 0; l < 32; l++) m |= (uint32_t)(prev_at(w + l) == prev_at(slot)) << l;
        return m;
    }

    void *smem_ptr = nullptr;
    size_t smem_bytes = 0;

    template <class F>
    void par(F f) {
        std::vector<PrivT> saved_pv = pv;
        std::vector<uint8_t> saved_smem;
        if (smem_ptr) saved_smem.assign((uint8_t *)smem_ptr, (uint8_t *)smem_ptr + smem_bytes);
        prev_v.clear();
        for (int pass = 0; pass < 64; pass++) {
            cur_v.clear();
            calls.assign(G::NT, 0);
            for (int t = 0; t < G::NT; t++) f(t, pv[t]);
            bool used = false;
            for (auto c : calls) used |= c != 0;
            if (!used) return;                       // no votes: one pass is the phase
            cur_v.resize(std::max(cur_v.size(), prev_v.size()), 0);
            prev_v.resize(cur_v.size(), 0);
            if (pass > 0 && cur_v == prev_v) return;  // fixpoint: this pass saw final operands everywhere
            prev_v = cur_v;
            pv = saved_pv;
            if (smem_ptr) memcpy(smem_ptr, saved_smem.data(), smem_bytes);
        }
        abort();  // votes did not converge: a bug in the phase or in this emulation
    }
    template <class F>
    void par_nosync(F f) { par(f); }
    void barrier() {}
    void midtile_prefetch(int) {}
    template <class Get, class Put>
    void scan(Get get, Put put) {
        std::vector<uint32_t> v(G::NT);
        for (int t = 0; t < G::NT; t++) v[t] = get(t, pv[t]);
        uint32_t total = 0;
        std::vector<uint32_t> ex(G::NT);
        for (int t = 0; t < G::NT; t++) { ex[t] = total; total += v[t]; }
        for (int t = 0; t < G::NT; t++) put(t, pv[t], ex[t], total);
    }
    void tally2(int, bool a, bool b, uint32_t *ca, uint32_t *cb) { *ca += a; *cb += b; }
    void global_add(uint32_t *p, uint32_t v) { *p += v; }
    uint64_t global_add64(unsigned long long *p, uint64_t v) { uint64_t o = *p; *p += v; return o; }
    void append64(int, bool valid, uint64_t v, uint64_t *out, uint64_t cap, unsigned long long *count) {
        if (!valid) return;
        unsigned long long pos = (*count)++;
        if (pos < cap) out[pos] = v;
    }
};

static uint64_t dedup_cap_override = 0;
static int emu_impl = 0;   // 0: warp tiles (filter_warp_kernel + filter_tail_kernel), 1: CTA tiles (filter_fused_kernel)
static uint64_t emu_ovf_units = 0, emu_wtiles = 0;
struct WEmuGeo { static constexpr int NT = 32; };

extern "C" {

void emu_set_dedup_cap(uint64_t cap) { dedup_cap_override = cap; }
void emu_set_hit_list_cap(uint32_t cap) { emu_hcap = cap ? cap : 0xFFFFFFFFu; }
void emu_set_impl(int impl) { emu_impl = impl; }
uint64_t emu_last_overflow_units() { return emu_ovf_units; }
uint64_t emu_last_wtiles() { return emu_wtiles; }

// bucketed table, same layout as the device table (dcn_core.cuh)
int emu_table_build(const uint64_t *keys, uint64_t n, double load, uint64_t **slots_out, uint64_t *nb_out,
                    int *has_empty) {
    uint64_t nb = table_buckets_for(n, load);
    uint64_t *slots = (uint64_t *)aligned_alloc(32, nb * 4 * sizeof(uint64_t));
    for (uint64_t i = 0; i < nb * 4; i++) slots[i] = DCN_EMPTY;
    *has_empty = 0;
    for (uint64_t i = 0; i < n; i++) {
        uint64_t h = keys[i];
        if (h == DCN_EMPTY) { *has_empty = 1; continue; }
        uint64_t b = table_bucket(h, nb);
        for (bool done = false; !done;) {
            for (int s = 0; s < 4 && !done; s++) {
                uint64_t &slot = slots[4 * b + s];
                if (slot == h) done = true;
                else if (slot == DCN_EMPTY) { slot = h; done = true; }
            }
            if (!done && ++b == nb) b = 0;
        }
    }
    *slots_out = slots;
    *nb_out = nb;
    return 0;
}
void emu_free(void *p) { free(p); }

}  // extern "C"

// mirrors dcn_filter_batch_device for k=31, w=15; returns 0 or a negative error
template <bool PACKED>
static int emu_filter_batch_t(const uint64_t *slots, uint64_t nb, int has_empty, const uint8_t *bases_in,
                     const uint64_t *rec_off, uint32_t n_rec, int paired, uint32_t prefix_len, uint32_t abs_thr,
                     double rel_thr, int deplete, uint8_t *keep, uint32_t *hits, uint32_t *total) {
    using G = Geo<31, 15>;
    uint64_t n_bases = rec_off[n_rec];
    // device buffers are 16-byte aligned; copy into an aligned buffer of exactly n_bases bytes
    uint8_t *bases = (uint8_t *)aligned_alloc(16, ((n_bases + 15) / 16 + 1) * 16);
    memcpy(bases, bases_in, n_bases);
    FilterParams P;
    P.bases = bases; P.pk_codes = nullptr; P.pk_inv = nullptr; P.nl_bits = nullptr; P.nl_bit0 = 0;
    P.base0 = 0; P.n_bases = n_bases; P.rec_off = rec_off; P.n_rec = n_rec;
    std::vector<uint32_t> pk_codes(2 * ((n_bases + 31) / 32) + 2), nl_bits((n_rec + 31) / 32 + 1, 0);
    std::vector<uint16_t> pk_inv(2 * ((n_bases + 31) / 32) + 2);
    if (PACKED) {   // what dcn_filter_batch's ingest stage ships instead of the bytes
        pack_ascii(bases, n_bases, pk_codes.data(), pk_inv.data(), 1);
        for (uint32_t r = 0; r < n_rec; r++) {
            uint64_t len = rec_off[r + 1] - rec_off[r];
            if (len < (uint64_t)G::K) continue;
            uint64_t n = (prefix_len > 0 && len > prefix_len) ? prefix_len : len;
            if (bases[rec_off[r] + n - 1] == (uint8_t)'\n') nl_bits[r >> 5] |= 1u << (r & 31);
        }
        P.bases = nullptr; P.pk_codes = pk_codes.data(); P.pk_inv = pk_inv.data(); P.nl_bits = nl_bits.data();
    }
    P.rpu = paired ? 2 : 1; P.n_units = n_rec / P.rpu;
    P.prefix_len = prefix_len; P.abs_thr = abs_thr; P.rel_thr = rel_thr; P.deplete = deplete;
    P.table.slots = slots; P.table.n_buckets = nb; P.table.has_empty_key = has_empty;
    P.keep = keep; P.hits = hits; P.total = total;

    // plan (device: prep kernels)
    PlanCfg cfg;
    uint32_t max_short = 0, n_long = 0;
    for (uint32_t u = 0; u < P.n_units; u++) plan_unit_stats(rec_off, P.rpu, u, max_short, n_long);
    cfg = plan_make_cfg<G>(max_short);
    uint32_t n_tiles = plan_num_tiles(n_bases, cfg);
    std::vector<uint32_t> tile_first(n_tiles + 1, 0), tile_end(n_tiles + 1, 0);
    for (uint32_t u = 0; u < P.n_units; u++) plan_unit_tiles(rec_off, 0, P.rpu, P.n_units, u, cfg, tile_first.data(), tile_end.data());

    auto *s = new TileSmem<G>();
    memset(s, 0xA5, sizeof(*s));  // shared memory starts out as garbage on the device
    HostExec<G> ex;
    ex.smem_ptr = s; ex.smem_bytes = sizeof(*s);
    ex.par([&](int t, TilePriv<G> &) { init_tables<G>(t, *s); init_required<G>(t, *s, abs_thr, rel_thr); });
    if (emu_impl == 1) {
        for (uint32_t tile = 0; tile < n_tiles; tile++)
            if (tile_first[tile] < tile_end[tile])
                filter_tile<G, PACKED, MODE_FILTER>(ex, *s, P, cfg, n_long, tile_first[tile], tile_end[tile]);
    } else {
        // mirrors wplan_kernel + filter_warp_kernel (one "warp" runs every tile) + the overflow part of filter_tail_kernel
        std::vector<WTile> wt;
        const uint64_t n_seg = (n_bases + DCN_WSEG - 1) / DCN_WSEG;
        for (uint64_t seg = 0; seg < n_seg; seg++) {
            const uint32_t cnt = wplan_segment(rec_off, 0, P.rpu, P.n_units, seg, [](uint64_t, uint32_t, uint32_t) {});
            const size_t before = wt.size();
            wplan_segment(rec_off, 0, P.rpu, P.n_units, seg, [&](uint64_t o, uint32_t a, uint32_t b) { WTile t; t.origin = o; t.a = a; t.b = b; wt.push_back(t); });
            if (wt.size() - before != cnt) abort();
        }
        if (wt.size() > n_bases / (uint64_t)(WG::TB - 15 - (int)DCN_MAX_SHORT) + (uint64_t)n_rec / (WG::MAXR / 2) + n_bases / DCN_MAX_SHORT + n_bases / DCN_WSEG + 16) abort();
        emu_wtiles = wt.size();
        auto *T = new WarpTables();
        auto *ws = new WarpSmem();
        memset(T, 0xA5, sizeof(*T));
        memset(ws, 0xA5, sizeof(*ws));
        HostExec<WEmuGeo, WarpPriv> wex;
        wex.smem_ptr = ws; wex.smem_bytes = sizeof(*ws);
        for (int t = 0; t < 1024; t++) winit_tables(t, 1024, *T, abs_thr, rel_thr);
        std::vector<uint32_t> ovf;
        for (size_t i = wt.size(); i-- > 0;) {   // any order
            const WTile &t = wt[i];
            const uint64_t left = n_bases - t.origin;
            const uint32_t need = left < (uint64_t)WG::TB ? (uint32_t)left : (uint32_t)WG::TB;
            if (!PACKED) memcpy(ws->stage, bases + t.origin, need);   // bulk copy + tail bytes; the rest of the stage keeps the previous tile
            const int calls0 = wex.after_scan_calls;
            warp_tile<PACKED, false>(wex, *T, *ws, P, t, need, [&](uint32_t u) { ovf.push_back(u); });
            if (wex.after_scan_calls != calls0 + 1) abort();   // the device starts the next tile's copy exactly once per tile
        }
        emu_ovf_units = ovf.size();
        if (ovf.size() > n_bases / WG::PKCAP + 16) abort();
        for (uint32_t u : ovf) filter_short_run<G, PACKED, MODE_FILTER>(ex, *s, P, u, u + 1);
        delete T;
        delete ws;
    }
    int rc = 0;
    if (n_long) {  // mirrors prep_long[_warp]_kernel + the chunk loop (CTA tiles: filter_fused_kernel; warp tiles: filter_warp_kernel) + finalize_long_kernel
        std::vector<uint32_t> long_units;
        std::vector<ChunkDesc> desc;
        std::vector<WTile> ltiles;
        uint64_t long_bases = 0;
        for (uint32_t u = 0; u < P.n_units; u++) {
            uint64_t len = rec_off[(uint64_t)(u + 1) * P.rpu] - rec_off[(uint64_t)u * P.rpu];
            if (len <= DCN_MAX_SHORT) continue;
            long_units.push_back(u);
            long_bases += len;
            hits[u] = 0; total[u] = 0;
            for (uint32_t r = u * P.rpu; r < (u + 1) * P.rpu; r++) {
                uint64_t gs = rec_off[r], rl = rec_off[r + 1] - gs;
                const uint64_t eff = filter_eff_len<G>(P, r, gs, rl);
                uint32_t nc = chunks_of<G>(eff);
                for (uint32_t c = 0; c < nc; c++) desc.push_back(ChunkDesc{r, c});
                const uint32_t nw = wplan_long_record(r, gs, eff, [&](uint32_t, const WTile &t) { ltiles.push_back(t); });
                if (nw != wplan_long_chunks_of(eff)) abort();
            }
        }
        uint64_t cap = std::max<uint64_t>(4096, long_bases / 4);
        // the warp-tile path gives every long unit a region of the set of its own (enqueue_filter does so for batches that
        // are mostly long units; here always, so that the mixed cases exercise it too)
        const bool local = emu_impl != 1 && !dedup_cap_override;
        if (local) cap = ((n_bases >> 4) + 2) * 4;
        if (dedup_cap_override) cap = dedup_cap_override;
        std::vector<unsigned __int128> slots(cap, 0);
        uint32_t overflow = 0;
        DedupView dd{slots.data(), cap, &overflow, 7u, local ? 4u : 0u};
        // the set is never cleared between calls (epoch tags): run the whole long path twice over the same slots, the
        // first time under another epoch, so that the pass that counts finds every slot it wants taken by a stale entry
        for (uint32_t epoch = 6; epoch <= 7; epoch++) {
            dd.epoch = epoch;
            for (uint32_t u : long_units) { hits[u] = 0; total[u] = 0; }
            if (emu_impl == 1) {
                for (size_t i = desc.size(); i-- > 0;) filter_long_chunk<G, PACKED>(ex, *s, P, dd, desc[i]);
            } else {
                auto *T = new WarpTables();
                auto *ws = new WarpSmem();
                memset(ws, 0xA5, sizeof(*ws));
                HostExec<WEmuGeo, WarpPriv> wex;
                wex.smem_ptr = ws; wex.smem_bytes = sizeof(*ws);
                for (int t = 0; t < 1024; t++) winit_tables(t, 1024, *T, abs_thr, rel_thr);
                for (size_t i = ltiles.size(); i-- > 0;) {   // any order
                    const uint64_t origin = ltiles[i].origin & ~WTILE_LONG;
                    const uint64_t left = n_bases - origin;
                    const uint32_t need = left < (uint64_t)WG::TB ? (uint32_t)left : (uint32_t)WG::TB;
                    if (!PACKED) memcpy(ws->stage, bases + origin, need);
                    warp_long_tile<PACKED>(wex, *T, *ws, P, dd, ltiles[i], need);
                }
                delete T;
                delete ws;
            }
            if (overflow) break;
        }
        for (uint32_t u : long_units)
            keep[u] = meets_criteria(hits[u], total[u], abs_thr, rel_thr, deplete) ? 1 : 0;
        if (overflow) rc = -6;
    }
    delete s;
    free(bases);
    return rc;
}

extern "C" {

int emu_filter_batch(const uint64_t *slots, uint64_t nb, int has_empty, const uint8_t *bases_in,
                     const uint64_t *rec_off, uint32_t n_rec, int paired, uint32_t prefix_len, uint32_t abs_thr,
                     double rel_thr, int deplete, uint8_t *keep, uint32_t *hits, uint32_t *total, int packed) {
    if (packed)
        return emu_filter_batch_t<true>(slots, nb, has_empty, bases_in, rec_off, n_rec, paired, prefix_len, abs_thr, rel_thr,
                                        deplete, keep, hits, total);
    return emu_filter_batch_t<false>(slots, nb, has_empty, bases_in, rec_off, n_rec, paired, prefix_len, abs_thr, rel_thr,
                                     deplete, keep, hits, total);
}

// mirrors tile_extract_device (B3, k=31, w=15, every record <= DCN_MAX_SHORT): wplan + extract_warp_kernel +
// extract_tail_kernel + scan + extract_compact_kernel.  -> number of minimizers, or -6 when out_cap is too small
long long emu_tile_extract(const uint8_t *bases_in, const uint64_t *rec_off, uint32_t n_rec, uint32_t prefix_len, uint64_t *out_h,
                           uint32_t *out_p, uint64_t *out_off, uint64_t out_cap) {
    using G = Geo<31, 15>;
    const uint64_t n_bases = rec_off[n_rec];
    uint8_t *bases = (uint8_t *)aligned_alloc(16, ((n_bases + 15) / 16 + 1) * 16);
    memcpy(bases, bases_in, n_bases);
    FilterParams P;
    memset(&P, 0, sizeof(P));
    P.bases = bases; P.base0 = 0; P.n_bases = n_bases; P.rec_off = rec_off; P.n_rec = n_rec; P.rpu = 1; P.n_units = n_rec;
    P.prefix_len = prefix_len;
    const uint64_t cap = n_bases + 64 * 1024;
    std::vector<uint64_t> tmp_h(cap), rec_cnt(n_rec + 1, 0), rec_tmp(n_rec + 1, 0);
    std::vector<uint32_t> tmp_p(cap);
    unsigned long long cursor = 0;
    P.xo.tmp_h = tmp_h.data(); P.xo.tmp_p = tmp_p.data(); P.xo.tmp_cap = cap; P.xo.cursor = &cursor;
    P.xo.rec_cnt = rec_cnt.data(); P.xo.rec_tmp = rec_tmp.data();
    auto *T = new WarpTables();
    auto *ws = new WarpSmem();
    memset(ws, 0xA5, sizeof(*ws));
    HostExec<WEmuGeo, WarpPriv> wex;
    wex.smem_ptr = ws; wex.smem_bytes = sizeof(*ws); wex.xcursor = &cursor;
    for (int t = 0; t < 1024; t++) winit_tables(t, 1024, *T, 2, 0.01);
    std::vector<uint32_t> ovf;
    const uint64_t n_seg = (n_bases + DCN_WSEG - 1) / DCN_WSEG;
    for (uint64_t seg = 0; seg < n_seg; seg++)
        wplan_segment(rec_off, 0, 1, n_rec, seg, [&](uint64_t o, uint32_t a, uint32_t b) {
            WTile t; t.origin = o; t.a = a; t.b = b;
            const uint64_t left = n_bases - o;
            const uint32_t need = left < (uint64_t)WG::TB ? (uint32_t)left : (uint32_t)WG::TB;
            memcpy(ws->stage, bases + o, need);
            warp_tile<false, true>(wex, *T, *ws, P, t, need, [&](uint32_t u) { ovf.push_back(u); });
        });
    if (!ovf.empty()) {
        auto *s = new TileSmem<G>();
        memset(s, 0xA5, sizeof(*s));
        HostExec<G> ex;
        ex.smem_ptr = s; ex.smem_bytes = sizeof(*s);
        ex.par([&](int t, TilePriv<G> &) { init_tables<G>(t, *s); });
        for (uint32_t u : ovf) filter_short_run<G, false, MODE_EXTRACT>(ex, *s, P, u, u + 1);
        delete s;
    }
    emu_ovf_units = ovf.size();
    uint64_t m = 0;
    for (uint32_t r = 0; r < n_rec; r++) { out_off[r] = m; m += rec_cnt[r]; }
    out_off[n_rec] = m;
    long long rc = (long long)m;
    if (m > out_cap) rc = -6;
    else
        for (uint32_t r = 0; r < n_rec; r++) {
            const uint64_t start = rec_tmp[r] >> 16, n = rec_tmp[r] & 0xFFFFu;
            uint64_t at = out_off[r];
            for (uint64_t i = 0; i < n; i++)
                if (tmp_p[start + i] & 0x80000000u) { out_h[at] = tmp_h[start + i]; out_p[at] = tmp_p[start + i] & 0x7FFFFFFFu; at++; }
            if (at != out_off[r + 1]) rc = -100;
        }
    delete T; delete ws;
    free(bases);
    return rc;
}

void emu_pack_ascii(const uint8_t *bases, uint64_t n, uint32_t *codes, uint16_t *inv, int simd) { pack_ascii(bases, n, codes, inv, simd); }
int emu_pack_has_simd() { return pack_has_simd() ? 1 : 0; }

// mirrors dcn_index_build_device's extraction (k=31, w=15): unordered hashes, duplicates included
long long emu_index_extract(const uint8_t *bases_in, const uint64_t *rec_off, uint32_t n_rec, const uint32_t *entropy_pass,
                            uint64_t *out, uint64_t out_cap) {
    using G = Geo<31, 15>;
    uint64_t n_bases = rec_off[n_rec];
    uint8_t *bases = (uint8_t *)aligned_alloc(16, ((n_bases + 15) / 16 + 1) * 16);
    memcpy(bases, bases_in, n_bases);
    unsigned long long count = 0;
    IndexParams P;
    P.bases = bases; P.base0 = 0; P.n_bases = n_bases; P.rec_off = rec_off; P.n_rec = n_rec;
    P.entropy_pass = entropy_pass; P.out = out; P.out_cap = out_cap; P.out_count = &count;
    auto *s = new TileSmem<G>();
    memset(s, 0xA5, sizeof(*s));
    HostExec<G> ex;
    ex.smem_ptr = s; ex.smem_bytes = sizeof(*s);
    ex.par([&](int t, TilePriv<G> &) { init_tables<G>(t, *s); });
    for (uint32_t r = 0; r < n_rec; r++) {
        uint64_t rl = rec_off[r + 1] - rec_off[r];
        uint32_t nc = chunks_of<G>(rl < (uint64_t)G::K ? 0 : rl);
        for (uint32_t c = 0; c < nc; c++) index_chunk<G>(ex, *s, P, ChunkDesc{r, c});
    }
    delete s;
    free(bases);
    return (long long)count;
}

}  // extern "C"

// ---- generic (k, w) path: mirrors generic_rec_chunks_kernel + scan + generic_count/write/filter kernels
template <int FLAV>
static std::vector<uint64_t> emu_rec_chunk_off(const uint8_t *bases, const uint64_t *rec_off, uint32_t n_rec, uint32_t prefix,
                                               int k, int w, uint32_t cstride) {
    std::vector<uint64_t> off(n_rec + 1, 0);
    for (uint32_t r = 0; r < n_rec; r++) {
        uint64_t gs = rec_off[r], len = rec_off[r + 1] - gs;
        off[r + 1] = off[r] + generic_chunks_of(generic_eff_len<FLAV>(bases, gs, len, prefix, k), k, w, cstride);
    }
    return off;
}

template <int FLAV>
static long long emu_generic_extract_t(const uint8_t *bases, const uint64_t *rec_off, uint32_t n_rec, int k, int w,
                                       uint32_t prefix, const uint32_t *entropy_pass, uint32_t cstride, uint64_t *out_h,
                                       uint32_t *out_p, uint64_t *out_off, uint64_t cap) {
    std::vector<uint64_t> rco = emu_rec_chunk_off<FLAV>(bases, rec_off, n_rec, prefix, k, w, cstride);
    GenericBatch B;
    B.bases = bases; B.base0 = 0; B.rec_off = rec_off; B.n_rec = n_rec; B.prefix_len = prefix; B.k = k; B.w = w;
    B.cstride = cstride; B.entropy_pass = entropy_pass; B.rec_chunk_off = rco.data();
    const uint64_t n_chunks = rco[n_rec];
    std::vector<uint64_t> coff(n_chunks + 1, 0);
    for (uint64_t g = 0; g < n_chunks; g++) {   // pass 1
        uint64_t n = 0;
        uint32_t r;
        generic_chunk<FLAV>(B, g, r, [&](uint64_t, uint64_t) { n++; });
        coff[g + 1] = coff[g] + n;
    }
    for (uint32_t r = 0; r <= n_rec; r++) out_off[r] = coff[rco[r]];
    if (coff[n_chunks] > cap) return -6;
    for (uint64_t g = n_chunks; g-- > 0;) {     // pass 2, any order
        uint64_t at = coff[g];
        uint32_t r;
        generic_chunk<FLAV>(B, g, r, [&](uint64_t pos, uint64_t h) { out_h[at] = h; out_p[at] = (uint32_t)pos; at++; });
        if (at != coff[g + 1]) return -100;
    }
    return (long long)coff[n_chunks];
}

extern "C" long long emu_generic_extract(int flavour, const uint8_t *bases, const uint64_t *rec_off, uint32_t n_rec, int k, int w,
                                         uint32_t prefix, const uint32_t *entropy_pass, uint32_t cstride, uint64_t *out_h,
                                         uint32_t *out_p, uint64_t *out_off, uint64_t cap) {
    if (flavour == 1) return emu_generic_extract_t<FLAVOUR_INDEX>(bases, rec_off, n_rec, k, w, 0, entropy_pass, cstride, out_h, out_p, out_off, cap);
    return emu_generic_extract_t<FLAVOUR_FILTER>(bases, rec_off, n_rec, k, w, prefix, nullptr, cstride, out_h, out_p, out_off, cap);
}

extern "C" int emu_generic_filter(const uint64_t *slots, uint64_t nb, int has_empty, const uint8_t *bases, const uint64_t *rec_off,
                                  uint32_t n_rec, int paired, uint32_t prefix, int k, int w, uint32_t cstride, uint32_t abs_thr,
                                  double rel_thr, int deplete, uint8_t *keep, uint32_t *hits, uint32_t *total) {
    const uint32_t rpu = paired ? 2 : 1, n_units = n_rec / rpu;
    std::vector<uint64_t> rco = emu_rec_chunk_off<FLAVOUR_FILTER>(bases, rec_off, n_rec, prefix, k, w, cstride);
    GenericBatch B;
    B.bases = bases; B.base0 = 0; B.rec_off = rec_off; B.n_rec = n_rec; B.prefix_len = prefix; B.k = k; B.w = w;
    B.cstride = cstride; B.entropy_pass = nullptr; B.rec_chunk_off = rco.data();
    TableView tv{slots, nb, has_empty};
    uint64_t cap = 4096;
    while (cap < 4 * rec_off[n_rec] / ((uint64_t)w + 1)) cap <<= 1;
    std::vector<unsigned __int128> set(cap, 0);
    uint32_t overflow = 0;
    DedupView dd{set.data(), cap, &overflow, 7u};
    for (uint32_t u = 0; u < n_units; u++) hits[u] = total[u] = 0;
    for (uint64_t g = rco[n_rec]; g-- > 0;) {
        uint32_t r = 0, nt = 0, nh = 0;
        const uint32_t unit = generic_find_record(rco.data(), n_rec, g) / rpu;
        generic_chunk<FLAVOUR_FILTER>(B, g, r, [&](uint64_t, uint64_t h) {
            nt++;
            if (table_contains(tv, h) && dedup_insert(dd, h, unit)) nh++;
        });
        total[unit] += nt; hits[unit] += nh;
    }
    for (uint32_t u = 0; u < n_units; u++) keep[u] = meets_criteria(hits[u], total[u], abs_thr, rel_thr, deplete) ? 1 : 0;
    return overflow ? -6 : 0;
}
