// dcn_emu.cpp -- TEST-ONLY host emulation of the CUDA tile pipeline.
//
// Compiles deacon_server_b200/csrc/dcn_tile.cuh with g++ and runs every phase as a loop over the
// 256 "threads" of a CTA, so the kernel's logic (bit tricks, rolling hash seeding, van Herk
// window minima, pick compaction, per-unit distinct-hit count) can be checked against the oracle
// on a machine without a GPU.  It is never part of the product library and bench.py never loads it.
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "../../deacon_server_b200/csrc/dcn_plan.cuh"
#include "../../deacon_server_b200/csrc/dcn_tile.cuh"

using namespace dcn;

template <class G>
struct HostExec {
    std::vector<TilePriv<G>> pv;
    HostExec() : pv(G::NT) {}
    template <class F>
    void par(F f) {
        for (int t = 0; t < G::NT; t++) f(t, pv[t]);
    }
    template <class Get, class Put>
    void scan(Get get, Put put) {
        std::vector<uint32_t> v(G::NT);
        for (int t = 0; t < G::NT; t++) v[t] = get(t, pv[t]);
        uint32_t total = 0;
        std::vector<uint32_t> ex(G::NT);
        for (int t = 0; t < G::NT; t++) { ex[t] = total; total += v[t]; }
        for (int t = 0; t < G::NT; t++) put(t, pv[t], ex[t], total);
    }
    void ballot2(int, uint32_t idx, bool valid, bool hit, uint32_t *vm, uint32_t *hm) {
        if (idx >= (uint32_t)G::PKCAP) return;
        if (valid) vm[idx >> 5] |= 1u << (idx & 31);
        if (hit) hm[idx >> 5] |= 1u << (idx & 31);
    }
};

extern "C" {

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

// mirrors dcn_filter_batch_device for k=31, w=15; returns 0 or a negative error
int emu_filter_batch(const uint64_t *slots, uint64_t nb, int has_empty, const uint8_t *bases_in,
                     const uint64_t *rec_off, uint32_t n_rec, int paired, uint32_t prefix_len, uint32_t abs_thr,
                     double rel_thr, int deplete, uint8_t *keep, uint32_t *hits, uint32_t *total) {
    using G = Geo<31, 15>;
    uint64_t n_bases = rec_off[n_rec];
    // device buffers are 16-byte aligned; copy into an aligned buffer of exactly n_bases bytes
    uint8_t *bases = (uint8_t *)aligned_alloc(16, ((n_bases + 15) / 16 + 1) * 16);
    memcpy(bases, bases_in, n_bases);
    FilterParams P;
    P.bases = bases; P.base0 = 0; P.n_bases = n_bases; P.rec_off = rec_off; P.n_rec = n_rec;
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
    ex.par([&](int t, TilePriv<G> &) { init_tables<G>(t, *s); });
    int rc = 0;
    for (uint32_t tile = 0; tile < n_tiles; tile++)
        filter_tile<G>(ex, *s, P, cfg, tile_first[tile], tile_end[tile]);
    if (n_long) rc = -100;  // long units are handled by a different kernel (not emulated here)
    delete s;
    free(bases);
    return rc;
}

}  // extern "C"
