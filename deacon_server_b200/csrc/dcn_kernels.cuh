// dcn_kernels.cuh -- device execution policy and the __global__ kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>

#include "dcn_tile.cuh"

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
#ifdef DCN_DYNAMIC_CHUNKS
    uint32_t chunk_claims, pad2;    // long-path chunks claimed beyond the first wave (inside the plan's 64-byte header)
#endif
};

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

// Index build: every record is cut into chunks.
template <class G>
__global__ void prep_index_chunks_kernel(const uint64_t *__restrict__ rec_off, uint32_t n_rec, BatchStats *st,
                                         ChunkDesc *desc, uint32_t desc_cap) {
    for (uint32_t r = blockIdx.x * blockDim.x + threadIdx.x; r < n_rec; r += gridDim.x * blockDim.x) {
        uint64_t rl = rec_off[r + 1] - rec_off[r];
        uint32_t nc = chunks_of<G>(rl < (uint64_t)G::K ? 0 : rl);
        if (!nc) continue;
        uint32_t at = atomicAdd(&st->n_chunks, nc);
        for (uint32_t c = 0; c < nc; c++)
            if (at + c < desc_cap) desc[at + c] = ChunkDesc{r, c};
            else st->overflow = 1;
    }
}

// keep flag of the long units once every chunk has added its counts
__global__ void finalize_long_kernel(FilterParams P, const BatchStats *st, const uint32_t *long_units) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < st->n_long_listed; i += gridDim.x * blockDim.x) {
        uint32_t u = long_units[i];
        P.keep[u] = meets_criteria(P.hits[u], P.total[u], P.abs_thr, P.rel_thr, P.deplete) ? 1 : 0;
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
        const uint32_t n_chunks = st->n_chunks;
#ifdef DCN_DYNAMIC_CHUNKS
        // Build variant, off by default until its parity run exists (make EXTRA=-DDCN_DYNAMIC_CHUNKS; DESIGN.md 9:
        // 127.7 -> 143.2 Gbp/s on the config-3 batch in one measurement): chunks beyond the first wave are
        // claimed from a counter like the tiles above (a chunk's cost follows its hit density).  Every chunk passes
        // barriers (the phases of chunk_picks), which order the two claim slots.
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
#else
        for (uint32_t w = gridDim.x - 1 - blockIdx.x; w < n_chunks; w += gridDim.x)
            filter_long_chunk<G, PACKED>(ex, s, P, dd, desc[w]);
#endif
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

// CSR compaction: one warp per record moves its valid picks from the tile block to out_off[r] ..
__global__ void extract_compact_kernel(ExtractOut xo, const uint64_t *__restrict__ out_off, uint32_t n_rec,
                                       uint64_t *__restrict__ out_h, uint32_t *__restrict__ out_p) {
    const uint32_t lane = threadIdx.x & 31u, warps = (gridDim.x * blockDim.x) >> 5;
    for (uint32_t r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; r < n_rec; r += warps) {
        const uint64_t rt = xo.rec_tmp[r], start = rt >> 16;
        const uint32_t n = (uint32_t)(rt & 0xFFFFu);
        uint64_t at = out_off[r];
        for (uint32_t base = 0; base < n; base += 32u) {
            const uint32_t i = base + lane;
            uint32_t pp = 0;
            if (i < n) pp = xo.tmp_p[start + i];
            const bool valid = (pp & 0x80000000u) != 0;
            const uint32_t m = __ballot_sync(0xFFFFFFFFu, valid);
            if (valid) {
                const uint64_t o = at + (uint64_t)__popc(m & ((1u << lane) - 1u));
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
// Every thread issues `per_thread` independent 32-byte loads at pseudo-random buckets.
__global__ void random_access_kernel(const uint64_t *__restrict__ slots, uint64_t n_buckets, uint32_t per_thread,
                                     unsigned long long *sink) {
    uint64_t x = (blockIdx.x * (uint64_t)blockDim.x + threadIdx.x) * 0x9E3779B97F4A7C15ULL + 0x1234567ULL;
    unsigned long long acc = 0;
    for (uint32_t i = 0; i < per_thread; i += 4) {
        Bucket bk[4];
#pragma unroll
        for (int j = 0; j < 4; j++) {
            x = xxh3_u64(x + i + j);
            bk[j] = load_bucket(slots, table_bucket(x, n_buckets));
        }
#pragma unroll
        for (int j = 0; j < 4; j++) acc += bk[j].k0 ^ bk[j].k1 ^ bk[j].k2 ^ bk[j].k3;
    }
    if (acc == 0x0123456789ABCDEFULL) *sink = acc;
}

}  // namespace dcn
