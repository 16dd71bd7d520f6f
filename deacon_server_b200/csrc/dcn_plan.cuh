// dcn_plan.cuh -- how a batch of records is cut into tiles (shared by the prep kernels and the
// host-side emulation used in tests).
#pragma once
#include "dcn_core.cuh"

namespace dcn {

// Units (a record, or a pair of mates) of at most DCN_MAX_SHORT bases are processed whole inside
// one tile ("short path"); longer units are cut into chunks ("long path").
static constexpr uint32_t DCN_MAX_SHORT = 1024;

struct PlanCfg {
    uint32_t S;          // tile stride: tile i owns the units whose first base lies in [i*S, (i+1)*S)
    uint32_t max_short;  // longest short unit of the batch
};

DCN_HD uint64_t table_buckets_for(uint64_t n_keys, double load) {
    if (!(load > 0.05)) load = 0.05;
    if (load > 0.9) load = 0.9;
    double nb = (double)n_keys / (4.0 * load);
    uint64_t r = (uint64_t)nb + 1;
    return r < 64 ? 64 : r;
}

DCN_HD void plan_unit_stats(const uint64_t *rec_off, uint32_t rpu, uint32_t u, uint32_t &max_short,
                            uint32_t &n_long) {
    uint64_t len = rec_off[(uint64_t)(u + 1) * rpu] - rec_off[(uint64_t)u * rpu];
    if (len > DCN_MAX_SHORT) n_long++;
    else if ((uint32_t)len > max_short) max_short = (uint32_t)len;
}

// A run of short units that starts in tile i spans at most 15 (alignment) + S - 1 + max_short bases.
template <class G>
DCN_HD PlanCfg plan_make_cfg(uint32_t max_short) {
    PlanCfg c;
    c.max_short = max_short;
    c.S = (uint32_t)G::BCAP - 14u - max_short;
    return c;
}

DCN_HD uint32_t plan_num_tiles(uint64_t n_bases, const PlanCfg &c) { return (uint32_t)(n_bases / c.S) + 1u; }

// tile_first/tile_end must be zero-initialised; a tile that owns no unit keeps first == end == 0.
// Offsets are absolute; base0 (a multiple of 16) is the absolute offset of bases[0].
DCN_HD void plan_unit_tiles(const uint64_t *rec_off, uint64_t base0, uint32_t rpu, uint32_t n_units, uint32_t u,
                            const PlanCfg &c, uint32_t *tile_first, uint32_t *tile_end) {
    uint64_t tile = (rec_off[(uint64_t)u * rpu] - base0) / c.S;
    if (u == 0 || (rec_off[(uint64_t)(u - 1) * rpu] - base0) / c.S != tile) tile_first[tile] = u;
    if (u + 1 == n_units || (rec_off[(uint64_t)(u + 1) * rpu] - base0) / c.S != tile) tile_end[tile] = u + 1;
}

}  // namespace dcn
