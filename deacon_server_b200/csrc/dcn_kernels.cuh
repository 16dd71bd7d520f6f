// dcn_kernels.cuh -- device execution policy and the __global__ kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>

#include "dcn_tile.cuh"
#include "dcn_warp.cuh"

namespace dcn {

// Execution policy for the device: one phase = every thread runs it, then a CTA barrier.
template <class G>
struct DevExec {
    TilePriv<G> pv;
    uint32_t *wsum;

    template <class F>
    __device__ __forceinline__ void par(F f) {
        f((int)threadIdx.x, pv);
        __syncthreads();
    }
    template <class F>
    __device__ __forceinline__ void par_nosync(F f) { f((int)threadIdx.x, pv); }
    __device__ __forceinline__ void barrier() { __syncthreads(); }
    __device__ __forceinline__ uint32_t ballot(int, bool p) { return __ballot_sync(0xFFFFFFFFu, p); }
    __device__ __forceinline__ uint64_t bcast64(int, uint64_t v, uint32_t src) {
        return (uint64_t)__shfl_sync(0xFFFFFFFFu, (unsigned long long)v, (int)src);
    }
    // lanes of this warp holding the same 64-bit value
    __device__ __forceinline__ uint32_t match64(int, uint64_t v, bool) { return __match_any_sync(0xFFFFFFFFu, (unsigned long long)v); }
    // add the number of lanes with a / b set to two shared-memory counters
    __device__ __forceinline__ void tally2(int t, bool a, bool b, uint32_t *ca, uint32_t *cb) {
        uint32_t ma = __ballot_sync(0xFFFFFFFFu, a), mb = __ballot_sync(0xFFFFFFFFu, b);
        if ((t & 31) == 0) {
            if (ma) atomicAdd(ca, (uint32_t)__popc(ma));
            if (mb) atomicAdd(cb, (uint32_t)__popc(mb));
        }
    }
    __device__ __forceinline__ void global_add(uint32_t *p, uint32_t v) { if (v) atomicAdd(p, v); }
    __device__ __forceinline__ uint64_t global_add64(unsigned long long *p, uint64_t v) { return atomicAdd(p, (unsigned long long)v); }
    // warp-aggregated append to a global array
    __device__ __forceinline__ void append64(int t, bool valid, uint64_t v, uint64_t *out, uint64_t cap,
                                             unsigned long long *count) {
        const uint32_t m = __ballot_sync(0xFFFFFFFFu, valid);
        if (!m) return;
        const int lane = t & 31;
        unsigned long long base = 0;
        if (lane == 0) base = atomicAdd(count, (unsigned long long)__popc(m));
        base = __shfl_sync(0xFFFFFFFFu, base, 0);
        if (valid) {
            unsigned long long pos = base + (unsigned long long)__popc(m & ((1u << lane) - 1u));
            if (pos < cap) out[pos] = v;
        }
    }
    // block-wide exclusive sum of a u32 (two packed 16-bit counters in our use)
    template <class Get, class Put>
    __device__ __forceinline__ void scan(Get get, Put put) {
        const int t = (int)threadIdx.x, lane = t & 31, warp = t >> 5;
        const uint32_t v = get(t, pv);
        uint32_t x = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            uint32_t y = __shfl_up_sync(0xFFFFFFFFu, x, d);
            if (lane >= d) x += y;
        }
        if (lane == 31) wsum[warp] = x;
        __syncthreads();
        uint32_t base = 0, total = 0;
#pragma unroll
        for (int w = 0; w < G::NT / 32; w++) {
            uint32_t sv = wsum[w];
            total += sv;
            if (w < warp) base += sv;
        }
        put(t, pv, base + x - v, total);
        __syncthreads();
    }
    // Records of the tile this CTA processes next: their offsets are pulled into L2 from the middle of
    // the current tile (the tile range itself was loaded one tile ahead), so the next tile's first
    // dependent loads (rec_off -> last byte of each record) do not start from DRAM.
    const uint64_t *pf_off = nullptr;
    uint32_t pf_lo = 0, pf_hi = 0;
    __device__ __forceinline__ void midtile_prefetch(int t) {
        const uint32_t i = pf_lo + 16u * (uint32_t)t;   // 16 offsets per 128-byte line
        if (t < 16 && i <= pf_hi) asm volatile("prefetch.global.L2 [%0];" ::"l"(pf_off + i));
    }
};

// device-side batch statistics, written by the prep kernels
struct BatchStats {
    uint32_t max_short;      // longest short unit
    uint32_t n_long;         // units longer than DCN_MAX_SHORT
    uint32_t n_chunks;       // chunk descriptors written
    uint32_t overflow;       // an internal table overflowed
    uint32_t n_long_listed;  // entries of the long-unit list
    uint32_t tile_claims;    // tiles claimed beyond the first wave (filter_fused_kernel)
    unsigned long long long_bases;  // bases in long units
    uint32_t chunk_claims;   // long-path chunks claimed beyond the first wave (filter_tail_kernel)
    uint32_t n_wtiles;       // warp tiles written by the planner (filter_warp_kernel)
    uint32_t n_ovf;          // units handed from the warp path to the CTA path (more picks than a warp pass holds)
    uint32_t ovf_claims;
};
static_assert(sizeof(BatchStats) <= 64, "the plan header is 64 bytes");

// ------------------------------------------------------------------ prep: unit statistics
__global__ void prep_stats_kernel(const uint64_t *__restrict__ rec_off, uint32_t rpu, uint32_t n_units,
                                  BatchStats *st) {  // lengths only: independent of base0
    uint32_t mx = 0, nl = 0;
    unsigned long long lb = 0;
    for (uint32_t u = blockIdx.x * blockDim.x + threadIdx.x; u < n_units; u += gridDim.x * blockDim.x) {
        uint32_t before = nl;
        plan_unit_stats(rec_off, rpu, u, mx, nl);
        if (nl != before) lb += rec_off[(uint64_t)(u + 1) * rpu] - rec_off[(uint64_t)u * rpu];
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        mx = max(mx, __shfl_xor_sync(0xFFFFFFFFu, mx, d));
        nl += __shfl_xor_sync(0xFFFFFFFFu, nl, d);
        lb += __shfl_xor_sync(0xFFFFFFFFu, lb, d);
    }
    if ((threadIdx.x & 31) == 0) {
        if (mx) atomicMax(&st->max_short, mx);
        if (nl) { atomicAdd(&st->n_long, nl); atomicAdd(&st->long_bases, lb); }
    }
}

// ------------------------------------------------------------------ prep: chunk descriptors
// Filter long path: one thread per unit; long units are listed, their outputs zeroed, and every
// record of theirs is cut into chunks.
template <class G>
__global__ void prep_long_kernel(FilterParams P, BatchStats *st, uint32_t *long_units, ChunkDesc *desc,
                                 uint32_t desc_cap) {
    for (uint32_t u = blockIdx.x * blockDim.x + threadIdx.x; u < P.n_units; u += gridDim.x * blockDim.x) {
        uint64_t len = P.rec_off[(uint64_t)(u + 1) * P.rpu] - P.rec_off[(uint64_t)u * P.rpu];
        if (len <= DCN_MAX_SHORT) continue;
        long_units[atomicAdd(&st->n_long_listed, 1u)] = u;
        P.hits[u] = 0; P.total[u] = 0;
        for (uint32_t r = u * P.rpu; r < (u + 1) * P.rpu; r++) {
            uint64_t gs = P.rec_off[r] - P.base0, rl = P.rec_off[r + 1] - P.base0 - gs;
            uint32_t nc = chunks_of<G>(filter_eff_len<G>(P, r, gs, rl));
            if (!nc) continue;
            uint32_t at = atomicAdd(&st->n_chunks, nc);
            for (uint32_t c = 0; c < nc; c++)
                if (at + c < desc_cap) desc[at + c] = ChunkDesc{r, c};
                else st->overflow = 1;
        }
    }
}

// Warp-tile form of the same: long units are listed and zeroed, every record of theirs is cut into chunks of
// DCN_WCS windows, appended to the warp-tile list behind the short tiles (the planner ran before).
template <class G>
__global__ void prep_long_warp_kernel(FilterParams P, BatchStats *st, uint32_t *long_units, WTile *tiles, uint32_t tile_cap) {
    for (uint32_t u = blockIdx.x * blockDim.x + threadIdx.x; u < P.n_units; u += gridDim.x * blockDim.x) {
        uint64_t len = P.rec_off[(uint64_t)(u + 1) * P.rpu] - P.rec_off[(uint64_t)u * P.rpu];
        if (len <= DCN_MAX_SHORT) continue;
        long_units[atomicAdd(&st->n_long_listed, 1u)] = u;
        P.hits[u] = 0; P.total[u] = 0;
        for (uint32_t r = u * P.rpu; r < (u + 1) * P.rpu; r++) {
            const uint64_t gs = P.rec_off[r] - P.base0, rl = P.rec_off[r + 1] - P.base0 - gs;
            const uint64_t eff = filter_eff_len<G>(P, r, gs, rl);
            const uint32_t nc = wplan_long_chunks_of(eff);
            if (!nc) continue;
            const uint32_t at = atomicAdd(&st->n_wtiles, nc);
            wplan_long_record(r, gs, eff, [&](uint32_t c, const WTile &t) {
                if (at + c < tile_cap) tiles[at + c] = t; else st->overflow = 1;
            });
        }
    }
}

// Index build: every record is cut into chunks.  A warp per record, the lanes write its descriptors (a thread per record
// wrote the 32 000 descriptors of a chromosome one after the other: 1.2 ms of the 50 ms build of a 24-contig reference).
template <class G>
__global__ void prep_index_chunks_kernel(const uint64_t *__restrict__ rec_off, uint32_t n_rec, BatchStats *st,
                                         ChunkDesc *desc, uint32_t desc_cap) {
    const uint32_t lane = threadIdx.x & 31u, warps = (gridDim.x * blockDim.x) >> 5;
    for (uint32_t r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; r < n_rec; r += warps) {
        uint64_t rl = rec_off[r + 1] - rec_off[r];
        uint32_t nc = chunks_of<G>(rl < (uint64_t)G::K ? 0 : rl);
        if (!nc) continue;   // (warp-uniform)
        uint32_t at = 0;
        if (lane == 0) at = atomicAdd(&st->n_chunks, nc);
        at = __shfl_sync(0xFFFFFFFFu, at, 0);
        for (uint32_t c = lane; c < nc; c += 32u)
            if (at + c < desc_cap) desc[at + c] = ChunkDesc{r, c};
            else st->overflow = 1;
    }
}

// keep flag of the long units once every chunk has added its counts
// `counters` (optional): the long units' share of the summary counters (warp-tile path; the CTA path runs stats_kernel)
__global__ void finalize_long_kernel(FilterParams P, const BatchStats *st, const uint32_t *long_units, unsigned long long *counters) {
    unsigned long long n_all = 0, n_kept = 0, bp_all = 0, bp_kept = 0;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < st->n_long_listed; i += gridDim.x * blockDim.x) {
        uint32_t u = long_units[i];
        const bool keep = meets_criteria(P.hits[u], P.total[u], P.abs_thr, P.rel_thr, P.deplete);
        P.keep[u] = keep ? 1 : 0;
        const unsigned long long len = P.rec_off[(uint64_t)(u + 1) * P.rpu] - P.rec_off[(uint64_t)u * P.rpu];
        n_all += P.rpu; bp_all += len;
        if (keep) { n_kept += P.rpu; bp_kept += len; }
    }
    if (counters && n_all) {
        atomicAdd(&counters[0], n_all); atomicAdd(&counters[1], n_all - n_kept); atomicAdd(&counters[2], bp_all);
        atomicAdd(&counters[3], bp_kept); atomicAdd(&counters[4], bp_all - bp_kept); atomicAdd(&counters[5], n_kept);
    }
}

// ------------------------------------------------------------------ index-flavour extraction kernel
template <class G>
__global__ void __launch_bounds__(G::NT, 1024 / G::NT)
extract_index_kernel(IndexParams P, const BatchStats *st, const ChunkDesc *__restrict__ desc) {
    extern __shared__ __align__(16) unsigned char dcn_smem_raw[];
    TileSmem<G> &s = *reinterpret_cast<TileSmem<G> *>(dcn_smem_raw);
    DevExec<G> ex;
    ex.wsum = s.wsum;
    init_tables<G>((int)threadIdx.x, s);
    __syncthreads();
    const uint32_t n_chunks = st->n_chunks;
    for (uint32_t w = blockIdx.x; w < n_chunks; w += gridDim.x) index_chunk<G>(ex, s, P, desc[w]);
}

// ------------------------------------------------------------------ prep: tile ownership
template <class G>
__global__ void prep_tiles_kernel(const uint64_t *__restrict__ rec_off, uint32_t rpu, uint32_t n_units,
                                  uint64_t base0, const BatchStats *st, uint32_t *tile_first, uint32_t *tile_end) {
    const PlanCfg cfg = plan_make_cfg<G>(st->max_short);
    for (uint32_t u = blockIdx.x * blockDim.x + threadIdx.x; u < n_units; u += gridDim.x * blockDim.x)
        plan_unit_tiles(rec_off, base0, rpu, n_units, u, cfg, tile_first, tile_end);
}

// ------------------------------------------------------------------ long path: chunks of long units
// The first wave takes chunk = (last CTA first, so that short and long work interleave in the fused kernel); chunks
// beyond it are claimed from a counter like the tiles: a chunk's cost follows its hit density (host-derived chunks hit
// on every probe, random ones never).  Measured on the config-3 batch: 127.7 -> 143.2 Gbp/s against the static map.
// Every chunk passes barriers (the phases of chunk_picks), which order the two claim slots.
template <class G, bool PACKED>
__device__ __forceinline__ void long_chunks_loop(DevExec<G> &ex, TileSmem<G> &s, const FilterParams &P, const DedupView &dd,
                                                 const ChunkDesc *__restrict__ desc, const BatchStats *st) {
    const uint32_t n_chunks = st->n_chunks;
    unsigned int *chunk_ctr = const_cast<unsigned int *>(&st->chunk_claims);
    if (threadIdx.x == 0) s.next_tile[0] = gridDim.x + atomicAdd(chunk_ctr, 1u);
    uint32_t cpar = 0;
    __syncthreads();
    uint32_t w = gridDim.x - 1 - blockIdx.x;
    while (w < n_chunks) {
        const uint32_t nw = s.next_tile[cpar];
        if (threadIdx.x == 0) s.next_tile[cpar ^ 1u] = nw < n_chunks ? gridDim.x + atomicAdd(chunk_ctr, 1u) : 0xFFFFFFFFu;
        cpar ^= 1u;
        filter_long_chunk<G, PACKED>(ex, s, P, dd, desc[w]);
        w = nw;
    }
}

// ------------------------------------------------------------------ the fused filter kernel
// Persistent CTAs; the first wave takes tile = CTA index, later tiles are claimed from a counter.  4 CTAs per SM (64 registers, ~53 KB shared memory each).
#ifndef DCN_CTAS_PER_SM
#define DCN_CTAS_PER_SM (1024 / DCN_NT)
#endif
template <class G, bool PACKED>
__global__ void __launch_bounds__(G::NT, DCN_CTAS_PER_SM)
filter_fused_kernel(FilterParams P, const BatchStats *st, const uint32_t *__restrict__ tile_first,
                    const uint32_t *__restrict__ tile_end, DedupView dd, const ChunkDesc *__restrict__ desc) {
    extern __shared__ __align__(16) unsigned char dcn_smem_raw[];
    TileSmem<G> &s = *reinterpret_cast<TileSmem<G> *>(dcn_smem_raw);
    DevExec<G> ex;
    ex.wsum = s.wsum;
    init_tables<G>((int)threadIdx.x, s);
    init_required<G>((int)threadIdx.x, s, P.abs_thr, P.rel_thr);
    // Tiles beyond the first wave are claimed from a counter (BatchStats::tile_claims, cleared with the plan) instead of
    // dealt out round-robin: a CTA that drew cheap tiles takes more of them (static map: 185.3 Gbp/s, claims: 197.8).
    // The claim for the tile after next is issued by thread 0 at the top of a tile and read by everyone at the top of
    // the next one (the barriers of the tile in between order the two slots).
    unsigned int *tile_ctr = const_cast<unsigned int *>(&st->tile_claims);
    if (threadIdx.x == 0) s.next_tile[0] = gridDim.x + atomicAdd(tile_ctr, 1u);
    uint32_t par = 0;
    __syncthreads();
    const PlanCfg cfg = plan_make_cfg<G>(st->max_short);
    const uint32_t n_tiles = plan_num_tiles(P.n_bases - P.base0, cfg);
    const uint32_t n_long = st->n_long;
    // Software pipeline over this CTA's tiles: the unit range of tile i+1 is loaded and its bases are
    // pulled into L2 while tile i is processed.
    uint32_t tile = blockIdx.x, a = 0, b = 0;
    if (tile < n_tiles) { a = tile_first[tile]; b = tile_end[tile]; }
    ex.pf_off = P.rec_off;
    while (tile < n_tiles) {
        const uint32_t nt = s.next_tile[par];
        if (threadIdx.x == 0) s.next_tile[par ^ 1u] = nt < n_tiles ? gridDim.x + atomicAdd(tile_ctr, 1u) : 0xFFFFFFFFu;
        par ^= 1u;
        uint32_t a2 = 0, b2 = 0;
        if (nt < n_tiles) {
            a2 = tile_first[nt]; b2 = tile_end[nt];
            // tile nt owns the units that start in [nt * S, (nt + 1) * S): at most S + max_short + 15 bases
            const uint64_t lo = (uint64_t)nt * cfg.S, n_rel = P.n_bases - P.base0;
            const uint64_t off = lo + 128ull * threadIdx.x;
            if (128u * threadIdx.x < cfg.S + cfg.max_short + 16u && off < n_rel) {
                if (PACKED) {   // 128 bases = 32 bytes of codes + 16 bytes of mask
                    if ((threadIdx.x & 3u) == 0) asm volatile("prefetch.global.L2 [%0];" ::"l"(P.pk_codes + (off >> 4)));
                    if ((threadIdx.x & 7u) == 0) asm volatile("prefetch.global.L2 [%0];" ::"l"(P.pk_inv + (off >> 4)));
                } else {
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(P.bases + off));
                }
            }
        }
        ex.pf_lo = a2 * P.rpu; ex.pf_hi = b2 * P.rpu;   // a2 == b2 == 0: one harmless line
        if (a < b) filter_tile<G, PACKED, MODE_FILTER>(ex, s, P, cfg, n_long, a, b);
        else __syncthreads();   // an empty tile has no barrier of its own to order the claim slots
        tile = nt; a = a2; b = b2;
    }
    if (n_long) {  // long units: chunks, spread over the CTAs in reverse so short and long work interleave
        __syncthreads();
        long_chunks_loop<G, PACKED>(ex, s, P, dd, desc, st);
    }
}

// ------------------------------------------------------------------ warp-tile planner
// One warp per 64 KB segment of the batch (units whose first base lies in the segment), the greedy walk of
// wplan_segment done 32 units at a time: the lanes load the offsets of 32 consecutive units (coalesced); every lane
// works out the tile that would START at its unit (run of short units of this segment, then a binary search over
// the lanes' end offsets for what fits 1536 bases from the lane's aligned origin); the warp then hops from tile
// start to tile start through those counts (shuffles only) and the starting lanes write their tiles.  One round
// trip to memory plans ~5 tiles, where the first version planned one: a small batch (a 16 MB chunk of the
// host-pointer pipeline is 256 segments of 43 tiles) is bound by that latency, not by throughput.  The segment's
// first unit is found by a 32-ary search (3-4 round trips instead of 16).  Tiles are staged in shared memory and
// appended to the list 64 or more at a time (one atomic per flush); the list order is irrelevant, tiles are
// claimed from a counter.  Long units (skipped here, cut into chunks by prep_long_warp_kernel) are counted on the
// way: n_long, long_bases.
__global__ void __launch_bounds__(256)
wplan_kernel(const uint64_t *__restrict__ rec_off, uint32_t rpu, uint32_t n_units, uint64_t base0, uint64_t n_rel,
             BatchStats *st, WTile *tiles, uint32_t tile_cap, uint32_t *promise_broken) {
    __shared__ WTile buf[8][96];
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const uint32_t FULL = 0xFFFFFFFFu;
    const uint64_t n_seg = (n_rel + DCN_WSEG - 1) / DCN_WSEG;
    const uint64_t n_warps = (uint64_t)gridDim.x * 8u;
    const uint32_t maxu = wplan_max_units(rpu);
    const uint64_t HUGE = ~0ull >> 1;
    for (uint64_t seg = (uint64_t)blockIdx.x * 8u + warp; seg < n_seg; seg += n_warps) {
        const uint64_t lo_pos = seg * DCN_WSEG, hi_pos = lo_pos + DCN_WSEG;
        // first unit with start >= lo_pos (wplan_first_unit), 32 probes per step
        uint32_t u = 0;
        {
            uint32_t lo = 0, hi = n_units;
            while (lo < hi) {
                const uint32_t mid = lo + (uint32_t)(((uint64_t)(hi - lo) * (lane + 1u)) / 33u);   // lo <= mid < hi, non-decreasing in lane
                const bool less = rec_off[(uint64_t)mid * rpu] - base0 < lo_pos;
                const uint32_t c = (uint32_t)__popc(__ballot_sync(FULL, less));             // probes 0 .. c-1 lie before the segment
                const uint32_t below = __shfl_sync(FULL, mid, (int)((c + 31u) & 31u)), above = __shfl_sync(FULL, mid, (int)(c & 31u));
                if (c > 0u) lo = below + 1u;
                if (c < 32u) hi = above;
            }
            u = lo;
        }
        uint32_t nbuf = 0, n_long = 0;
        unsigned long long long_bases = 0;   // per lane; summed over the warp at the end
        auto flush = [&]() {
            uint32_t at = 0;
            if (lane == 0) at = atomicAdd(&st->n_wtiles, nbuf);
            at = __shfl_sync(FULL, at, 0);
            for (uint32_t i = lane; i < nbuf; i += 32u) {
                if (at + i < tile_cap) tiles[at + i] = buf[warp][i];
                else st->overflow = 1;
            }
            __syncwarp();
            nbuf = 0;
        };
        while (u < n_units) {
            const uint64_t idx = (uint64_t)u + lane;
            const uint64_t s_l = idx <= n_units ? rec_off[idx * rpu] - base0 : HUGE;
            uint64_t e_l = __shfl_down_sync(FULL, (unsigned long long)s_l, 1);
            if (lane == 31u) e_l = idx + 1 <= n_units ? rec_off[(idx + 1) * rpu] - base0 : HUGE;
            const uint64_t s0 = __shfl_sync(FULL, (unsigned long long)s_l, 0);
            if (s0 >= hi_pos) break;
            const bool inseg = idx < n_units && s_l < hi_pos;
            const uint64_t len = e_l - s_l;
            const bool is_long = inseg && len > DCN_MAX_SHORT;
            const bool ok = inseg && !is_long;
            const uint32_t okm = __ballot_sync(FULL, ok), longm = __ballot_sync(FULL, is_long), insegm = __ballot_sync(FULL, inseg);
            // the tile that would start at this lane's unit: short units of the segment in a row (at most maxu) ...
            const uint32_t nok = ~(okm >> lane);                                   // (bits beyond the window read as "not ok")
            uint32_t run = nok ? (uint32_t)__ffs((int)nok) - 1u : 32u;
            if (run > maxu) run = maxu;
            // ... of which the first c end within 1536 bases of the tile's aligned origin (a short unit always fits alone)
            const uint64_t origin = s_l & ~15ull, limit = origin + (uint64_t)WG::TB;
            uint32_t c = ok ? 1u : 0u;
#pragma unroll
            for (uint32_t step = 16u; step >= 1u; step >>= 1) {
                const uint32_t t = c + step;
                const uint32_t from = lane + t - 1u;
                const uint64_t ev = __shfl_sync(FULL, (unsigned long long)e_l, (int)(from & 31u));
                if (ok && t <= run && from < 32u && ev <= limit) c = t;
            }
            // the count is final if the tile is full or the unit that ended it is in view; otherwise the walk restarts there
            const uint32_t finalm = __ballot_sync(FULL, c == maxu || lane + c < 32u);
            uint32_t p = 0, startm = 0, skipm = 0;
            bool seg_done = false;
            while (p < 32u) {
                if (!((insegm >> p) & 1u)) { seg_done = true; break; }              // past the batch or the segment
                if ((longm >> p) & 1u) { skipm |= 1u << p; p++; continue; }
                if (!((finalm >> p) & 1u)) break;
                startm |= 1u << p;
                p += __shfl_sync(FULL, c, (int)p);
            }
            if ((skipm >> lane) & 1u) long_bases += len;
            n_long += (uint32_t)__popc(skipm);
            if ((startm >> lane) & 1u) {
                WTile t;
                t.origin = origin; t.a = u + lane; t.b = u + lane + c;
                buf[warp][nbuf + (uint32_t)__popc(startm & ((1u << lane) - 1u))] = t;
            }
            nbuf += (uint32_t)__popc(startm);
            u += p;                                                                  // p >= 1: lane 0's unit is in view and its count final
            __syncwarp();
            if (nbuf >= 64u) flush();
            if (seg_done) break;
        }
        if (nbuf) flush();
        if (n_long) {   // (warp-uniform)
#pragma unroll
            for (int d = 16; d >= 1; d >>= 1) long_bases += __shfl_xor_sync(FULL, long_bases, d);
            if (lane == 0) {
                atomicAdd(&st->n_long, n_long); atomicAdd(&st->long_bases, long_bases);
                if (promise_broken) *reinterpret_cast<volatile uint32_t *>(promise_broken) = 1u;   // pinned host word (dcn_filter_batch_device_hint)
            }
        }
    }
}

// ------------------------------------------------------------------ the warp-tile filter kernel (short units)
// One persistent CTA of 32 warps per SM; every warp runs tiles on its own (dcn_warp.cuh): the first wave takes tile =
// global warp index, later tiles are claimed from a counter two tiles ahead.  The tile's bytes are brought into the
// warp's stage by one bulk copy (cp.async.bulk -> mbarrier complete_tx) issued by lane 0 as soon as the previous
// tile's convert phase has drained the stage, i.e. the copy of tile i + 1 runs under the hash / slice / probe / count
// phases of tile i.
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, uint32_t phase) {
    asm volatile(
        "{\n\t"
        ".reg .pred P1;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
        "@P1 bra DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "DONE:\n\t"
        "}" ::"r"(smem_u32(bar)), "r"(phase) : "memory");
}
__device__ __forceinline__ void bulk_copy_g2s(void *dst, const void *src, uint32_t bytes, unsigned long long *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// src/local_filter.rs:347-371 (single) / 488-525 (pair): seqs and bp in / kept / filtered
__device__ __forceinline__ void add_summary(unsigned long long *counters, unsigned long long n_all, unsigned long long n_kept,
                                            unsigned long long bp_all, unsigned long long bp_kept) {
    atomicAdd(&counters[0], n_all);                  // total_seqs
    atomicAdd(&counters[1], n_all - n_kept);         // filtered_seqs
    atomicAdd(&counters[2], bp_all);                 // total_bp
    atomicAdd(&counters[3], bp_kept);                // output_bp
    atomicAdd(&counters[4], bp_all - bp_kept);       // filtered_bp
    atomicAdd(&counters[5], n_kept);                 // output_seq_counter
}

#ifndef DCN_WARPS
#define DCN_WARPS 32            // warps per CTA of filter_warp_kernel (one CTA per SM): 32 x 64 registers
#endif
#define DCN_WCTAS (32 / DCN_WARPS)   // CTAs per SM of the warp-tile kernels: 32 warps x 64 registers per SM either way

// Per-warp pipeline state that is only touched between tiles lives in shared memory, not in registers: the
// descriptors of the next two tiles (brought in by cp.async, so that no register waits for them).
struct WarpPipe {
    WTile d1, d2;   // next tile, tile after next
    unsigned long long bp_all, bp_kept;   // summary counters of the units this warp classified (lane 0 adds, flushed once at the end)
    uint32_t n_all, n_kept, xleft, pad_;
    unsigned long long xbase;             // extraction: the warp's current block of the temp arrays, entries left in it
};
static constexpr uint32_t DCN_XBLK = 4096;   // temp entries a warp reserves with one atomic (extract_warp_kernel)

struct WarpDevExec {
    WarpPriv pv;
    int lane;
    // what after_scan() needs to start the next tile
    const FilterParams *P;
    WarpSmem *s;
    WarpPipe *pipe;
    const WTile *tiles;
    unsigned int *tile_ctr;
    uint32_t id1, id2, c3;  // ids of the next two tiles; c3: claim for the one after (lane 0)
    uint32_t n_tiles, total_warps;
    bool packed;

    template <class F>
    __device__ __forceinline__ void par(F f) { f(lane, pv); __syncwarp(); }
    __device__ __forceinline__ uint32_t ballot(int, bool p) { return __ballot_sync(0xFFFFFFFFu, p); }
    __device__ __forceinline__ uint64_t bcast64(int, uint64_t v, uint32_t src) {
        return (uint64_t)__shfl_sync(0xFFFFFFFFu, (unsigned long long)v, (int)src);
    }
    __device__ __forceinline__ uint32_t match64(int, uint64_t v, bool) { return __match_any_sync(0xFFFFFFFFu, (unsigned long long)v); }
    __device__ __forceinline__ void global_add(uint32_t *p, uint32_t v) { if (v) atomicAdd(p, v); }
    // extraction: room for n picks in the temp arrays, from the warp's own block (a new block when it does not fit)
    __device__ __forceinline__ uint64_t xalloc(uint32_t n) {
        if (lane == 0) {
            if (n > pipe->xleft) {
                const uint32_t blk = n > DCN_XBLK ? n : DCN_XBLK;
                pipe->xbase = atomicAdd(P->xo.cursor, (unsigned long long)blk);
                pipe->xleft = blk;
            }
            pipe->xleft -= n;
            pipe->xbase += n;
        }
        __syncwarp();
        const uint64_t at = pipe->xbase - n;
        __syncwarp();
        return at;
    }
    __device__ __forceinline__ void tally(uint32_t nrec, uint32_t len, bool keep) {   // lane 0 only
        pipe->bp_all += len; pipe->n_all += nrec;
        if (keep) { pipe->bp_kept += len; pipe->n_kept += nrec; }
    }
    template <class Get, class Put>
    __device__ __forceinline__ void scan(Get get, Put put) {
        const uint32_t v = get(lane, pv);
        uint32_t x = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, x, d);
            if (lane >= d) x += y;
        }
        const uint32_t total = __shfl_sync(0xFFFFFFFFu, x, 31);
        put(lane, pv, x - v, total);
        __syncwarp();
    }
    // bytes of tile t that the bulk copy brings (multiple of 16); the up-to-15 bytes after them are copied by lanes
    __device__ __forceinline__ uint32_t tile_need(uint64_t origin) const {
        const uint64_t left = P->n_bases - P->base0 - origin;
        return left < (uint64_t)WG::TB ? (uint32_t)left : (uint32_t)WG::TB;   // (origin without the WTILE_LONG bit)
    }
    __device__ __forceinline__ void issue_copy(const WTile &tt) {
        WTile t = tt;
        const bool is_long = (t.origin & WTILE_LONG) != 0;
        t.origin &= ~WTILE_LONG;
        if (packed) {   // packed input is read straight from global memory: pull the next tile's words into L2
            const uint64_t w0 = t.origin >> 4;
            if (lane < 4) asm volatile("prefetch.global.L2 [%0];" ::"l"(P->pk_codes + w0 + 32u * (uint32_t)lane));
            else if (lane < 6) asm volatile("prefetch.global.L2 [%0];" ::"l"(P->pk_inv + w0 + 64u * (uint32_t)(lane - 4)));
        } else if (lane == 0) {
            const uint32_t bytes = tile_need(t.origin) & ~15u;   // t.origin: flag bit already cleared
            mbar_expect_tx(&s->mbar, bytes);
            if (bytes) bulk_copy_g2s(s->stage, P->bases + t.origin, bytes, &s->mbar);
        }
        // the record offsets of the tile: one 128-byte line holds 16 (a chunk of a long unit needs none)
        const uint32_t r0 = t.a * P->rpu, r1 = t.b * P->rpu;
        if (!is_long && lane >= 8 && lane < 14 && r0 + 16u * (uint32_t)(lane - 8) <= r1)
            asm volatile("prefetch.global.L2 [%0];" ::"l"(P->rec_off + r0 + 16u * (uint32_t)(lane - 8)));
    }
    __device__ __forceinline__ void fetch_desc(WTile *dst, uint32_t id) {   // lane 0: 16 bytes global -> shared, asynchronously
        asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst)), "l"(tiles + id) : "memory");
    }
    __device__ __forceinline__ void after_scan(bool go) {
        if (!go) return;
        if (id1 < n_tiles) issue_copy(pipe->d1);
        c3 = 0xFFFFFFFFu;
        if (lane == 0 && id2 < n_tiles) {
            fetch_desc(&pipe->d2, id2);
            c3 = total_warps + atomicAdd(tile_ctr, 1u);
        }
    }
};

// WITH_LONG: the batch holds long units (their chunks are listed behind the short tiles); a batch without any runs the
// instantiation that does not carry that code (measured: 2 % on the headline config, registers and instruction cache).
template <bool PACKED, bool WITH_LONG>
__global__ void __launch_bounds__(DCN_WARPS * 32, DCN_WCTAS)
filter_warp_kernel(FilterParams P, BatchStats *st, const WTile *__restrict__ tiles, uint32_t *ovf_list, uint32_t ovf_cap,
                   unsigned long long *counters, DedupView dd) {
    extern __shared__ __align__(16) unsigned char dcn_smem_raw[];
    WarpTables &T = *reinterpret_cast<WarpTables *>(dcn_smem_raw);
    const int warp = (int)(threadIdx.x >> 5), lane = (int)(threadIdx.x & 31u);
    constexpr size_t OFF_PIPE = (sizeof(WarpTables) + 15) & ~(size_t)15, OFF_WARPS = OFF_PIPE + DCN_WARPS * sizeof(WarpPipe);
    WarpPipe &pipe = reinterpret_cast<WarpPipe *>(dcn_smem_raw + OFF_PIPE)[warp];
    WarpSmem &s = reinterpret_cast<WarpSmem *>(dcn_smem_raw + OFF_WARPS)[warp];
    winit_tables((int)threadIdx.x, (int)blockDim.x, T, P.abs_thr, P.rel_thr);
    if (lane == 0) {
        mbar_init(&s.mbar, 1u);
        pipe.bp_all = pipe.bp_kept = 0; pipe.n_all = pipe.n_kept = 0; pipe.xleft = 0; pipe.xbase = 0;
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();   // the only CTA barrier of the kernel

    WarpDevExec ex;
    ex.lane = lane; ex.P = &P; ex.s = &s; ex.pipe = &pipe; ex.tiles = tiles; ex.packed = PACKED;
    ex.tile_ctr = &st->tile_claims;
    ex.n_tiles = st->n_wtiles;
    ex.total_warps = gridDim.x * DCN_WARPS;
    const uint32_t n_tiles = ex.n_tiles;

    // first wave: tile = global warp index; two more tiles claimed up front
    uint32_t id0 = blockIdx.x * DCN_WARPS + (uint32_t)warp;
    uint32_t claim = 0;
    if (lane == 0 && id0 < n_tiles) claim = ex.total_warps + atomicAdd(ex.tile_ctr, 2u);
    claim = __shfl_sync(0xFFFFFFFFu, claim, 0);
    ex.id1 = id0 < n_tiles ? claim : 0xFFFFFFFFu;
    ex.id2 = id0 < n_tiles ? claim + 1u : 0xFFFFFFFFu;
    ex.c3 = 0xFFFFFFFFu;
    WTile d0;
    d0.origin = 0; d0.a = d0.b = 0;
    if (id0 < n_tiles) { d0 = tiles[id0]; ex.issue_copy(d0); }
    if (lane == 0 && ex.id1 < n_tiles) ex.fetch_desc(&pipe.d1, ex.id1);
    bool have = id0 < n_tiles;
    uint32_t phase = 0;
    while (have) {
        const bool is_long = (d0.origin & WTILE_LONG) != 0;
        const uint64_t origin0 = d0.origin & ~WTILE_LONG;
        const uint32_t need = ex.tile_need(origin0);
        asm volatile("cp.async.wait_all;" ::: "memory");   // the descriptors requested during the previous tile (lane 0)
        __syncwarp();
        if (!PACKED) {
            mbar_wait(&s.mbar, phase);
            phase ^= 1u;
            const uint32_t bulk = need & ~15u;
            if ((uint32_t)lane < need - bulk) s.stage[bulk + (uint32_t)lane] = P.bases[origin0 + bulk + (uint32_t)lane];
            __syncwarp();
        }
        if (WITH_LONG && is_long) {   // a chunk of a long unit (listed behind the short tiles by prep_long_warp_kernel)
            warp_long_tile<PACKED>(ex, T, s, P, dd, d0, need);
        } else {
            warp_tile<PACKED, false>(ex, T, s, P, d0, need, [&](uint32_t u) {
                if (lane == 0) {
                    const uint32_t at = atomicAdd(&st->n_ovf, 1u);
                    if (at < ovf_cap) ovf_list[at] = u; else st->overflow = 1;
                }
            });
        }
        // rotate the pipeline: d1 (in shared memory since the previous tile) becomes the current tile
        have = ex.id1 < n_tiles;
        asm volatile("cp.async.wait_all;" ::: "memory");
        __syncwarp();
        d0 = pipe.d1;
        __syncwarp();
        if (lane == 0) pipe.d1 = pipe.d2;
        ex.id1 = ex.id2;
        ex.id2 = __shfl_sync(0xFFFFFFFFu, ex.c3, 0);
    }
    if (lane == 0 && pipe.n_all) add_summary(counters, pipe.n_all, pipe.n_kept, pipe.bp_all, pipe.bp_kept);
}

// B3 on warp tiles: the same skeleton, ASCII input, no long units (dcn_extract sends batches with a record above
// DCN_MAX_SHORT bases through the generic kernels); records of more than a warp pass's picks are listed for
// extract_tail_kernel.
__global__ void __launch_bounds__(DCN_WARPS * 32, DCN_WCTAS)
extract_warp_kernel(FilterParams P, BatchStats *st, const WTile *__restrict__ tiles, uint32_t *ovf_list, uint32_t ovf_cap) {
    constexpr bool PACKED = false;
    extern __shared__ __align__(16) unsigned char dcn_smem_raw[];
    WarpTables &T = *reinterpret_cast<WarpTables *>(dcn_smem_raw);
    const int warp = (int)(threadIdx.x >> 5), lane = (int)(threadIdx.x & 31u);
    constexpr size_t OFF_PIPE = (sizeof(WarpTables) + 15) & ~(size_t)15, OFF_WARPS = OFF_PIPE + DCN_WARPS * sizeof(WarpPipe);
    WarpPipe &pipe = reinterpret_cast<WarpPipe *>(dcn_smem_raw + OFF_PIPE)[warp];
    WarpSmem &s = reinterpret_cast<WarpSmem *>(dcn_smem_raw + OFF_WARPS)[warp];
    winit_tables((int)threadIdx.x, (int)blockDim.x, T, P.abs_thr, P.rel_thr);
    if (lane == 0) {
        mbar_init(&s.mbar, 1u);
        pipe.bp_all = pipe.bp_kept = 0; pipe.n_all = pipe.n_kept = 0; pipe.xleft = 0; pipe.xbase = 0;
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();   // the only CTA barrier of the kernel

    WarpDevExec ex;
    ex.lane = lane; ex.P = &P; ex.s = &s; ex.pipe = &pipe; ex.tiles = tiles; ex.packed = PACKED;
    ex.tile_ctr = &st->tile_claims;
    ex.n_tiles = st->n_wtiles;
    ex.total_warps = gridDim.x * DCN_WARPS;
    const uint32_t n_tiles = ex.n_tiles;

    // first wave: tile = global warp index; two more tiles claimed up front
    uint32_t id0 = blockIdx.x * DCN_WARPS + (uint32_t)warp;
    uint32_t claim = 0;
    if (lane == 0 && id0 < n_tiles) claim = ex.total_warps + atomicAdd(ex.tile_ctr, 2u);
    claim = __shfl_sync(0xFFFFFFFFu, claim, 0);
    ex.id1 = id0 < n_tiles ? claim : 0xFFFFFFFFu;
    ex.id2 = id0 < n_tiles ? claim + 1u : 0xFFFFFFFFu;
    ex.c3 = 0xFFFFFFFFu;
    WTile d0;
    d0.origin = 0; d0.a = d0.b = 0;
    if (id0 < n_tiles) { d0 = tiles[id0]; ex.issue_copy(d0); }
    if (lane == 0 && ex.id1 < n_tiles) ex.fetch_desc(&pipe.d1, ex.id1);
    bool have = id0 < n_tiles;
    uint32_t phase = 0;
    while (have) {
        const bool is_long = (d0.origin & WTILE_LONG) != 0;
        const uint64_t origin0 = d0.origin & ~WTILE_LONG;
        const uint32_t need = ex.tile_need(origin0);
        asm volatile("cp.async.wait_all;" ::: "memory");   // the descriptors requested during the previous tile (lane 0)
        __syncwarp();
        if (!PACKED) {
            mbar_wait(&s.mbar, phase);
            phase ^= 1u;
            const uint32_t bulk = need & ~15u;
            if ((uint32_t)lane < need - bulk) s.stage[bulk + (uint32_t)lane] = P.bases[origin0 + bulk + (uint32_t)lane];
            __syncwarp();
        }
        if (!is_long) {
            warp_tile<PACKED, true>(ex, T, s, P, d0, need, [&](uint32_t u) {
                if (lane == 0) {
                    const uint32_t at = atomicAdd(&st->n_ovf, 1u);
                    if (at < ovf_cap) ovf_list[at] = u; else st->overflow = 1;
                }
            });
        }
        // rotate the pipeline: d1 (in shared memory since the previous tile) becomes the current tile
        have = ex.id1 < n_tiles;
        asm volatile("cp.async.wait_all;" ::: "memory");
        __syncwarp();
        d0 = pipe.d1;
        __syncwarp();
        if (lane == 0) pipe.d1 = pipe.d2;
        ex.id1 = ex.id2;
        ex.id2 = __shfl_sync(0xFFFFFFFFu, ex.c3, 0);
    }
}

// records the warp extraction could not hold (more than a warp pass's picks): the CTA-tile extraction, one record at a time
template <class G>
__global__ void __launch_bounds__(G::NT, 1024 / G::NT)
extract_tail_kernel(FilterParams P, const BatchStats *st, const uint32_t *__restrict__ ovf_list) {
    const uint32_t n_ovf = st->n_ovf;
    if (n_ovf == 0) return;
    extern __shared__ __align__(16) unsigned char dcn_smem_raw[];
    TileSmem<G> &s = *reinterpret_cast<TileSmem<G> *>(dcn_smem_raw);
    DevExec<G> ex;
    ex.wsum = s.wsum;
    init_tables<G>((int)threadIdx.x, s);
    __syncthreads();
    ex.pf_off = P.rec_off;
    for (uint32_t i = blockIdx.x; i < n_ovf; i += gridDim.x) {
        const uint32_t u = ovf_list[i];
        filter_short_run<G, false, MODE_EXTRACT>(ex, s, P, u, u + 1u);
        __syncthreads();
    }
}

// ------------------------------------------------------------------ the CTA-tile tail of the warp kernel
// What a warp tile cannot hold: single units that emit more picks than a warp pass takes (filter_short_run holds
// 1024 per unit: enough for any short unit) and the chunks of long units.  Launched after filter_warp_kernel on the
// same stream; with nothing listed every CTA leaves at once.
template <class G, bool PACKED>
__global__ void __launch_bounds__(G::NT, DCN_CTAS_PER_SM)
filter_tail_kernel(FilterParams P, const BatchStats *st, const uint32_t *__restrict__ ovf_list, DedupView dd,
                   const ChunkDesc *__restrict__ desc, unsigned long long *counters) {
    const uint32_t n_ovf = st->n_ovf, n_long = desc ? st->n_long : 0u;   // no chunk list: the long units went through warp tiles
    if (n_ovf == 0 && n_long == 0) return;
    extern __shared__ __align__(16) unsigned char dcn_smem_raw[];
    TileSmem<G> &s = *reinterpret_cast<TileSmem<G> *>(dcn_smem_raw);
    DevExec<G> ex;
    ex.wsum = s.wsum;
    init_tables<G>((int)threadIdx.x, s);
    init_required<G>((int)threadIdx.x, s, P.abs_thr, P.rel_thr);
    __syncthreads();
    ex.pf_off = P.rec_off;
    for (uint32_t i = blockIdx.x; i < n_ovf; i += gridDim.x) {
        const uint32_t u = ovf_list[i];
        filter_short_run<G, PACKED, MODE_FILTER>(ex, s, P, u, u + 1u);
        __syncthreads();
        if (threadIdx.x == 0) {   // the unit's share of the summary counters (the warp kernel left it out)
            const unsigned long long len = P.rec_off[(uint64_t)(u + 1) * P.rpu] - P.rec_off[(uint64_t)u * P.rpu];
            const bool kept = P.keep[u] != 0;
            add_summary(counters, P.rpu, kept ? P.rpu : 0u, len, kept ? len : 0ull);
        }
    }
    if (n_long) {
        __syncthreads();
        long_chunks_loop<G, PACKED>(ex, s, P, dd, desc, st);
    }
}

// ------------------------------------------------------------------ B3: extraction through the tile pipeline
// get_minimizer_hashes_and_positions (src/filter_common.rs:211-310) for a batch of short records
// (every record <= DCN_MAX_SHORT bases, k = 31, w = 15): phases 1-4 of the fused kernel, then hashes
// and record-relative positions into per-tile blocks of the temp arrays.
template <class G>
__global__ void __launch_bounds__(G::NT, 1024 / G::NT)
extract_tiles_kernel(FilterParams P, const BatchStats *st, const uint32_t *__restrict__ tile_first,
                     const uint32_t *__restrict__ tile_end) {
    extern __shared__ __align__(16) unsigned char dcn_smem_raw[];
    TileSmem<G> &s = *reinterpret_cast<TileSmem<G> *>(dcn_smem_raw);
    DevExec<G> ex;
    ex.wsum = s.wsum;
    init_tables<G>((int)threadIdx.x, s);
    __syncthreads();
    const PlanCfg cfg = plan_make_cfg<G>(st->max_short);
    const uint32_t n_tiles = plan_num_tiles(P.n_bases - P.base0, cfg);
    ex.pf_off = P.rec_off;
    for (uint32_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const uint32_t a = tile_first[tile], b = tile_end[tile];
        if (a < b) {
            filter_tile<G, false, MODE_EXTRACT>(ex, s, P, cfg, 0u, a, b);
            __syncthreads();   // the extraction phases end without a barrier; the next tile rewrites wsum / the pick list
        }
    }
}

// CSR compaction: a half warp per record moves its valid picks from the tile block to out_off[r] .. (a 150-base read has
// ~14 picks: with a whole warp per record more than half of the lanes had nothing to move)
__global__ void extract_compact_kernel(ExtractOut xo, const uint64_t *__restrict__ out_off, uint32_t n_rec,
                                       uint64_t *__restrict__ out_h, uint32_t *__restrict__ out_p) {
    const uint32_t lane = threadIdx.x & 31u, sub = lane & 15u, half = lane >> 4;
    const uint32_t groups = (gridDim.x * blockDim.x) >> 4;
    const uint32_t n_it = (n_rec + groups - 1) / groups;   // the same trip count for both halves of a warp (ballots below)
    uint32_t r = (blockIdx.x * blockDim.x + threadIdx.x) >> 4;
    for (uint32_t it = 0; it < n_it; it++, r += groups) {
        uint64_t start = 0, at = 0;
        uint32_t n = 0;
        if (r < n_rec) {
            const uint64_t rt = xo.rec_tmp[r];
            start = rt >> 16; n = (uint32_t)(rt & 0xFFFFu);
            at = out_off[r];
        }
        const uint32_t n_other = __shfl_xor_sync(0xFFFFFFFFu, n, 16);
        const uint32_t n_max = n > n_other ? n : n_other;
        for (uint32_t base = 0; base < n_max; base += 16u) {
            const uint32_t i = base + sub;
            uint32_t pp = 0;
            if (i < n) pp = xo.tmp_p[start + i];
            const bool valid = (pp & 0x80000000u) != 0;
            const uint32_t m = (__ballot_sync(0xFFFFFFFFu, valid) >> (16u * half)) & 0xFFFFu;
            if (valid) {
                const uint64_t o = at + (uint64_t)__popc(m & ((1u << sub) - 1u));
                out_h[o] = xo.tmp_h[start + i];
                if (out_p) out_p[o] = pp & 0x7FFFFFFFu;
            }
            at += (uint64_t)__popc(m);
        }
    }
}

// ------------------------------------------------------------------ B2: lookup on pre-hashed records
// unpaired_should_keep / paired_should_keep (src/remote_filter.rs:230-301): one hash list per
// record.  One warp per record: lanes take 32 consecutive hashes (coalesced 8-byte loads), each
// lane probes one 32-byte bucket, distinct hits by warp match within the 32 and a scan of the
// record's earlier hashes for records of up to DCN_MAX_SHORT hashes; longer records go through
// the global (hash, record) set.  No shared memory, so 64 warps per SM keep ~2000 probes in flight.
__global__ void __launch_bounds__(256)
lookup_kernel(const uint64_t *__restrict__ hashes, const uint64_t *__restrict__ rec_off, uint32_t n_rec,
              TableView table, DedupView dd, uint32_t abs_thr, double rel_thr, int deplete,
              uint8_t *__restrict__ keep, uint32_t *__restrict__ hits_out, uint32_t *__restrict__ total_out,
              uint8_t *__restrict__ hit_flags) {   // optional: 1 where a hash is a counted (first, in-index) hit
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t lt = (1u << lane) - 1u;
    const uint32_t warps = (gridDim.x * blockDim.x) >> 5;
    // Software pipeline over the warp's records, three stages deep: while record r is probed, the first 32 hashes of
    // record r + warps and the offsets of record r + 2 warps are already requested, so a warp keeps three dependent-load
    // chains (offsets -> hashes -> bucket) in flight.
    uint32_t r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    uint64_t a = 0, b = 0, h0 = 0, an = 0, bn = 0;
    if (r < n_rec) {
        a = rec_off[r]; b = rec_off[r + 1];
        if (a + lane < b) h0 = hashes[a + lane];
    }
    if (r + warps < n_rec) { an = rec_off[r + warps]; bn = rec_off[r + warps + 1]; }
    while (r < n_rec) {
        const uint32_t rn = r + warps, rnn = rn + warps;
        uint64_t ann = 0, bnn = 0, hn = 0;
        if (rnn < n_rec) { ann = rec_off[rnn]; bnn = rec_off[rnn + 1]; }
        if (rn < n_rec && an + lane < bn) hn = hashes[an + lane];
        const bool big = b - a > DCN_MAX_SHORT;
        uint32_t hits = 0;
        for (uint64_t base = a; base < b; base += 32) {
            const uint64_t i = base + lane;
            const bool in = i < b;
            uint64_t h = 0;
            bool found = false;
            if (in) {
                h = base == a ? h0 : hashes[i];
                found = table_contains(table, h);
            }
            bool fresh;
            if (big) {
                fresh = found && dedup_insert(dd, h, r);
            } else {
                const uint32_t inmask = __ballot_sync(0xFFFFFFFFu, in);
                const uint32_t same = __match_any_sync(0xFFFFFFFFu, (unsigned long long)h);
                bool dup = (same & inmask & lt) != 0;
                if (found && !dup) {
                    const uint32_t hlo = (uint32_t)h;
                    for (uint64_t j = a; j < base && !dup; j++) {
                        uint64_t e = hashes[j];
                        if ((uint32_t)e == hlo) dup = e == h;
                    }
                }
                fresh = found && !dup;
            }
            if (hit_flags && in) hit_flags[i] = fresh ? 1 : 0;
            hits += (uint32_t)__popc(__ballot_sync(0xFFFFFFFFu, fresh));
        }
        if (lane == 0) {
            const uint64_t total = b - a;
            hits_out[r] = hits;
            total_out[r] = (uint32_t)total;
            keep[r] = meets_criteria(hits, total, abs_thr, rel_thr, deplete) ? 1 : 0;
        }
        r = rn; a = an; b = bn; h0 = hn; an = ann; bn = bnn;
    }
}

// ------------------------------------------------------------------ record offsets of a chunk of equal-length records
// (host ingest: such a chunk's rec_off is first + i * len, so it is written here instead of being copied over PCIe)
__global__ void uniform_offsets_kernel(uint64_t *__restrict__ off, uint32_t n, uint64_t first, uint64_t len) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) off[i] = first + (uint64_t)i * len;
}

// ------------------------------------------------------------------ sparse non-ACGT mask of the host-packed form
// The packer threads ship (32-base block, 32-bit mask) pairs for the blocks that hold a non-ACGT byte; the dense bit
// array the kernels read is a memset plus this.
// `base`: block index of inv32[0] (the packer threads list blocks relative to their chunk: 0; a caller's list counts from
// the start of the batch)
__global__ void inv_scatter_kernel(uint32_t *__restrict__ inv32, const uint2 *__restrict__ exc, uint32_t n, uint32_t base) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) inv32[exc[i].x - base] = exc[i].y;
}

// ------------------------------------------------------------------ index build: finish a radix sort of the top 40 bits
// The keys are XXH3 outputs: after a stable radix sort of bits 24 .. 63 (five 8-bit passes instead of eight) two keys are
// out of order only if they share those 40 bits -- ~70 000 pairs among 4 x 10^8.  A thread that finds its key smaller
// than the one before it (an inversion) and is the first such thread of its run of equal prefixes lists the run; a
// second kernel sorts the listed runs in place, one thread each (insertion sort; runs of two or three distinct keys,
// plus their duplicates).  Two kernels, because deciding who owns a run while somebody sorts it would be a race.  Runs of one repeated key -- a
// minimizer of a repeat occurs thousands of times in the picks of a genome -- have no inversion and cost nothing.  A
// run that reaches more than DCN_RUN_MAX keys from an inversion (keys that are not hashes: an .idx file may hold
// anything) raises `too_long` and the caller sorts all 64 bits instead.
static constexpr uint32_t DCN_RUN_MAX = 48;
// pass 1 (reads only): the runs that need sorting, listed by their first key's index
__global__ void find_prefix_runs_kernel(const uint64_t *__restrict__ keys, uint64_t n, uint64_t *__restrict__ runs, uint32_t runs_cap,
                                        uint32_t *n_runs, uint32_t *too_long) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x + 1; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t v = keys[i];
        if (keys[i - 1] <= v) continue;                            // in order (the common case: one load pair per key)
        const uint64_t p = v >> 24;
        // back to the start of the run; another inversion on the way means an earlier thread lists the run
        uint64_t s0 = i - 1;
        bool mine = true;
        while (s0 > 0 && i - s0 <= DCN_RUN_MAX && (keys[s0 - 1] >> 24) == p) {
            if (keys[s0 - 1] > keys[s0]) { mine = false; break; }
            s0--;
        }
        if (!mine) continue;
        if (i - s0 > DCN_RUN_MAX) { *too_long = 1u; continue; }
        const uint32_t at = atomicAdd(n_runs, 1u);
        if (at < runs_cap) runs[at] = s0; else *too_long = 1u;
    }
}
// pass 2: one thread per listed run sorts it in place (the runs are disjoint)
__global__ void sort_prefix_runs_kernel(uint64_t *__restrict__ keys, uint64_t n, const uint64_t *__restrict__ runs, const uint32_t *n_runs,
                                        uint32_t runs_cap, uint32_t *too_long) {
    const uint32_t nr = *n_runs < runs_cap ? *n_runs : runs_cap;
    for (uint32_t r = blockIdx.x * blockDim.x + threadIdx.x; r < nr; r += gridDim.x * blockDim.x) {
        const uint64_t s0 = runs[r], p = keys[s0] >> 24;
        uint64_t e = s0 + 1;
        while (e < n && e - s0 <= DCN_RUN_MAX && (keys[e] >> 24) == p) e++;
        if (e - s0 > DCN_RUN_MAX) { *too_long = 1u; continue; }
        for (uint64_t a = s0 + 1; a < e; a++) {
            const uint64_t x = keys[a];
            uint64_t b2 = a;
            while (b2 > s0 && keys[b2 - 1] > x) { keys[b2] = keys[b2 - 1]; b2--; }
            keys[b2] = x;
        }
    }
}

// ------------------------------------------------------------------ arena form of the host pipeline: one claim's blob -> its places
// A packer thread's claim crosses PCIe as ONE copy (codes | newline flags | exception list or dense mask | offsets) into a
// staging buffer; this kernel moves the parts to the claim's places in the arena (device-to-device, a few microseconds)
// -- three or four small copies per claim cost the copy engine more than their bytes (measured: 223 copies per call,
// 12.5 ms of engine time for 575 MB).  The dense non-ACGT array was cleared by a memset enqueued before (sparse form).
struct ClaimUnpack {
    const uint32_t *codes; uint32_t *dst_codes; uint64_t n_words;
    const uint32_t *nl; uint32_t *dst_nl; uint32_t nl_words;
    const uint2 *exc; int64_t n_exc;            // sparse form: (block, mask) pairs; < 0: dense mask at `inv`
    const uint32_t *inv; uint32_t *dst_inv32;
    const uint64_t *off; uint64_t *dst_off; uint32_t n_off;   // offsets shipped (n_off entries), or generated:
    uint64_t off_first, off_len; uint32_t n_gen;              // dst_off[i] = off_first + i * off_len, i < n_gen
};
__global__ void claim_unpack_kernel(ClaimUnpack u) {
    const uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x, nt = (uint64_t)gridDim.x * blockDim.x;
    const uint64_t n4 = u.n_words / 4;   // (both sides are 16-byte aligned: claims start at 64-base cuts)
    const uint4 *s4 = reinterpret_cast<const uint4 *>(u.codes);
    uint4 *d4 = reinterpret_cast<uint4 *>(u.dst_codes);
    for (uint64_t i = tid; i < n4; i += nt) d4[i] = s4[i];
    for (uint64_t i = 4 * n4 + tid; i < u.n_words; i += nt) u.dst_codes[i] = u.codes[i];
    for (uint64_t i = tid; i < u.nl_words; i += nt) u.dst_nl[i] = u.nl[i];
    if (u.n_exc >= 0) { for (uint64_t i = tid; i < (uint64_t)u.n_exc; i += nt) u.dst_inv32[u.exc[i].x] = u.exc[i].y; }
    else { for (uint64_t i = tid; i < u.n_words / 2; i += nt) u.dst_inv32[i] = u.inv[i]; }
    for (uint64_t i = tid; i < u.n_off; i += nt) u.dst_off[i] = u.off[i];
    for (uint64_t i = tid; i < u.n_gen; i += nt) u.dst_off[i] = u.off_first + i * u.off_len;
}

// ------------------------------------------------------------------ summary counters (a13)
// src/local_filter.rs:347-371 (single) / 488-525 (pair): seqs and bp in / kept / filtered.
__global__ void stats_kernel(const uint64_t *__restrict__ rec_off, uint32_t rpu, uint32_t n_units,
                             const uint8_t *__restrict__ keep, unsigned long long *counters) {
    unsigned long long bp_all = 0, bp_kept = 0, n_all = 0, n_kept = 0;
    for (uint32_t u = blockIdx.x * blockDim.x + threadIdx.x; u < n_units; u += gridDim.x * blockDim.x) {
        unsigned long long len = rec_off[(uint64_t)(u + 1) * rpu] - rec_off[(uint64_t)u * rpu];
        bp_all += len; n_all += rpu;
        if (keep[u]) { bp_kept += len; n_kept += rpu; }
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        bp_all += __shfl_xor_sync(0xFFFFFFFFu, bp_all, d);
        bp_kept += __shfl_xor_sync(0xFFFFFFFFu, bp_kept, d);
        n_all += __shfl_xor_sync(0xFFFFFFFFu, n_all, d);
        n_kept += __shfl_xor_sync(0xFFFFFFFFu, n_kept, d);
    }
    if ((threadIdx.x & 31) == 0 && n_all) {
        atomicAdd(&counters[0], n_all);                  // total_seqs
        atomicAdd(&counters[1], n_all - n_kept);         // filtered_seqs
        atomicAdd(&counters[2], bp_all);                 // total_bp
        atomicAdd(&counters[3], bp_kept);                // output_bp
        atomicAdd(&counters[4], bp_all - bp_kept);       // filtered_bp
        atomicAdd(&counters[5], n_kept);                 // output_seq_counter
    }
}

// one call's share of the summary counters -> the ctx's counters
// `st` (optional): a launch whose distinct-hit set overflowed does not commit -- its caller repeats it with a larger set
// (the launches of the arena pipeline find out when they are retired, not before the commit is enqueued)
__global__ void commit_counters_kernel(const unsigned long long *__restrict__ call_cnt, unsigned long long *counters, const BatchStats *st) {
    if (st && st->overflow) return;
    if (threadIdx.x < 6 && call_cnt[threadIdx.x]) atomicAdd(&counters[threadIdx.x], call_cnt[threadIdx.x]);
}

// ------------------------------------------------------------------ table build (K4)
__global__ void table_fill_kernel(uint64_t *slots, uint64_t n_slots) {
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n_slots; i += (uint64_t)gridDim.x * blockDim.x)
        slots[i] = DCN_EMPTY;
}

// Insert keys (duplicates allowed).  counts[0] += newly inserted keys, counts[1] = DCN_EMPTY seen.
__global__ void table_insert_kernel(uint64_t *slots, uint64_t n_buckets, const uint64_t *__restrict__ keys,
                                    uint64_t n_keys, unsigned long long *counts) {
    unsigned long long added = 0;
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n_keys; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t h = keys[i];
        if (h == DCN_EMPTY) { counts[1] = 1; continue; }
        uint64_t b = table_bucket(h, n_buckets);
        bool done = false;
        while (!done) {
            unsigned long long *bp = reinterpret_cast<unsigned long long *>(slots + 4 * b);
#pragma unroll
            for (int sI = 0; sI < 4 && !done; sI++) {
                unsigned long long cur = bp[sI];
                if (cur == h) { done = true; break; }
                if (cur == DCN_EMPTY) {
                    unsigned long long old = atomicCAS(&bp[sI], (unsigned long long)DCN_EMPTY, (unsigned long long)h);
                    if (old == DCN_EMPTY) { added++; done = true; }
                    else if (old == h) done = true;
                }
            }
            if (!done && ++b == n_buckets) b = 0;
        }
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) added += __shfl_xor_sync(0xFFFFFFFFu, added, d);
    if ((threadIdx.x & 31) == 0 && added) atomicAdd(&counts[0], added);
}

// ------------------------------------------------------------------ random-access ceiling probe
// Every thread issues `per_thread` independent probes at pseudo-random buckets; a probe reads SECTORS consecutive
// 32-byte sectors of a SECTORS * 32-byte aligned bucket (1 = the table's layout: 4 keys; 4 = what a 128-byte / 16-key
// bucket would cost).  Four probes in flight per thread.
template <int SECTORS>
__global__ void random_access_kernel(const uint64_t *__restrict__ slots, uint64_t n_buckets, uint32_t per_thread,
                                     unsigned long long *sink) {
    uint64_t x = (blockIdx.x * (uint64_t)blockDim.x + threadIdx.x) * 0x9E3779B97F4A7C15ULL + 0x1234567ULL;
    unsigned long long acc = 0;
    const uint64_t n_wide = n_buckets / SECTORS;
    for (uint32_t i = 0; i < per_thread; i += 4) {
        Bucket bk[4][SECTORS];
#pragma unroll
        for (int j = 0; j < 4; j++) {
            x = xxh3_u64(x + i + j);
            const uint64_t b0 = table_bucket(x, n_wide) * SECTORS;
#pragma unroll
            for (int q = 0; q < SECTORS; q++) bk[j][q] = load_bucket(slots, b0 + q);
        }
#pragma unroll
        for (int j = 0; j < 4; j++)
#pragma unroll
            for (int q = 0; q < SECTORS; q++) acc += bk[j][q].k0 ^ bk[j][q].k1 ^ bk[j][q].k2 ^ bk[j][q].k3;
    }
    if (acc == 0x0123456789ABCDEFULL) *sink = acc;
}

// The 128-byte bucket read the way a cooperative probe would do it: four adjacent lanes take the four sectors of one
// line in the same instruction (one line request per probe instead of four sector requests).
__global__ void random_access_line_kernel(const uint64_t *__restrict__ slots, uint64_t n_buckets, uint32_t per_group,
                                          unsigned long long *sink) {
    const uint64_t group = (blockIdx.x * (uint64_t)blockDim.x + threadIdx.x) >> 2;
    const uint32_t q = threadIdx.x & 3u;
    uint64_t x = group * 0x9E3779B97F4A7C15ULL + 0x1234567ULL;
    unsigned long long acc = 0;
    const uint64_t n_wide = n_buckets / 4;
    for (uint32_t i = 0; i < per_group; i += 4) {
        Bucket bk[4];
#pragma unroll
        for (int j = 0; j < 4; j++) {
            x = xxh3_u64(x + i + j);
            bk[j] = load_bucket(slots, table_bucket(x, n_wide) * 4 + q);
        }
#pragma unroll
        for (int j = 0; j < 4; j++) acc += bk[j].k0 ^ bk[j].k1 ^ bk[j].k2 ^ bk[j].k3;
    }
    if (acc == 0x0123456789ABCDEFULL) *sink = acc;
}

}  // namespace dcn
