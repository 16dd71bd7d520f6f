// dcn_api.cu -- the C ABI (include/deacon_cuda.h) over the sm_100a kernels.
//
// There is deliberately no CPU path in this file: every entry point either runs CUDA kernels or
// returns an error.
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <sched.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>
#include <cub/device/device_select.cuh>
#include <math.h>

#include "../../include/deacon_cuda.h"
#include "dcn_kernels.cuh"
#include "dcn_generic.cuh"
#include "dcn_host_pack.h"

using namespace dcn;

namespace {

thread_local std::string g_create_error;

struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
    // only used when the buffer holds a distinct-hit set (DedupView): bytes cleared once, epoch of the last call
    size_t set_cleared = 0;
    uint32_t set_epoch = 0;
    cudaError_t ensure(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        set_cleared = 0; set_epoch = 0;
        if (p) { cudaError_t e = cudaFree(p); p = nullptr; cap = 0; if (e != cudaSuccess) return e; }
        size_t want = bytes + bytes / 8 + 256;
        cudaError_t e = cudaMalloc(&p, want);
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
    template <class T> T *as() const { return reinterpret_cast<T *>(p); }
};

struct HostBuf {  // pinned staging
    void *p = nullptr;
    size_t cap = 0;
    cudaError_t ensure(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        if (p) { cudaFreeHost(p); p = nullptr; cap = 0; }
        size_t want = bytes + bytes / 8 + 256;
        cudaError_t e = cudaMallocHost(&p, want);
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() { if (p) cudaFreeHost(p); p = nullptr; cap = 0; }
    template <class T> T *as() const { return reinterpret_cast<T *>(p); }
};

struct Slot {  // one stage of the host-pointer pipeline
    DevBuf in, out, plan, longs, dedup;   // in: bases | off (ASCII) or codes | inv | off | nl (packed); out: hits | total | keep
    HostBuf h_in, h_out;                  // pinned mirrors of `in` (packed mode only) and `out`
    cudaStream_t stream = nullptr;
    cudaEvent_t ev_start = nullptr, ev_h2d = nullptr, ev_kernel = nullptr, ev_done = nullptr;
    bool busy = false, packed = false;
    uint32_t u0 = 0, u1 = 0;
};

// what the fused kernel reads: ASCII bytes, or the host-packed form (dcn_host_pack.h)
struct FilterInput {
    const uint8_t *bases = nullptr;
    const uint32_t *codes = nullptr;
    const uint16_t *inv = nullptr;
    const uint32_t *nl = nullptr;
    uint32_t nl_bit0 = 0;
};

// Device arena of the two-route host-pointer pipeline (filter_pipeline_arena): the whole batch in ONE coordinate system
// (ASCII bytes, packed codes / non-ACGT bits / newline flags, record offsets, results), so that a kernel launch can
// cover any contiguous range of units that has arrived, however many copies brought it.
struct ArenaState {
    DevBuf ascii, codes, inv, nl, off, out;
    static const int NL = 8;
    Slot launch[NL];              // plan / longs / dedup / h_out / stream / events of one kernel launch
    std::mutex launch_m[NL];
    cudaStream_t copy_stream = nullptr;   // the ASCII route's copies, in order
    cudaEvent_t ev_sub[4] = {nullptr, nullptr, nullptr, nullptr};   // pacing of the ASCII copies
    cudaEvent_t ev_front[NL] = {};        // "ASCII copied up to here", one per launch slot
    BatchStats *h_after = nullptr;        // pinned, NL entries: stats of a launch with long units (overflow flag of its distinct-hit set)
};

}  // namespace

struct dcn_ctx {
    ArenaState *arena = nullptr;
    int device = 0;
    int sm_count = 148;
    std::string err;
    cudaStream_t stream = nullptr;
    // resident index
    DevBuf table;
    uint64_t n_buckets = 0, n_keys = 0;
    int has_empty = 0;
    uint8_t k = 0, w = 0;
    double load = 0.5;
    uint32_t *h_promise = nullptr;   // pinned: set by the device when a dcn_filter_batch_device_hint promise was broken
    int fused_impl = 0;   // 0: warp tiles (filter_warp_kernel + filter_tail_kernel); 1: CTA tiles (filter_fused_kernel), DCN_FUSED_IMPL=cta
    // scratch
    DevBuf plan;       // BatchStats + tile_first + tile_end (one memset clears all three)
    DevBuf counters;   // 6 x u64 ProcessingStats + 2 x u64 table-build counters
    DevBuf keys_stage;
    DevBuf longs, dedup;   // long-path scratch of the device-pointer API
    // index build
    DevBuf ib_bases, ib_off, ib_desc, ib_keys, ib_alt, ib_tmp, ib_entropy, ib_stats, ib_runs;
    uint64_t ib_n = 0;     // the working key set: sorted unique keys of the last build / decode / union / diff (in ib_keys)
    uint8_t ws_k = 0, ws_w = 0;   // its header (src/index.rs:17-22)
    DevBuf ws_table, ws_flags;    // scratch table for set difference
    // generic (k, w) path and B3 extraction: staging, chunk plan, CSR outputs
    DevBuf gx_bases, gx_off, gx_rc, gx_cc, gx_tmp, gx_h, gx_p, gx_oo, gx_entropy;
    static const int NSLOT = 4;
    Slot slot[NSLOT];
    // host ingest (filter_pipeline): pack_threads = 0 ships everything as ASCII over PCIe
    int pack_threads = -1;   // -1: decide at first use (DCN_PACK_THREADS, or the CPUs this process may use - 4, at most 16)
    std::vector<Slot> pslot;   // the stages of the packer threads (four each), created by the threads themselves
    std::vector<Slot> aslot;   // arena form: the packer threads' blobs (h_in), exception lists (in), copy streams and events
    float t_pack = 0;
    // A chunk is either packed by a host thread (the CPU reads 1 B/bp, PCIe carries 0.43 B/bp) or shipped as ASCII
    // (PCIe carries 1 B/bp, no CPU work); `pack_fraction` caps the share of the batch the first route may take.
    double pack_gbps = 0;       // packing rate of the last call that packed (ASCII GB/s), for reporting
    double pack_fraction = -1;  // < 0: automatic (pinned caller buffers: dynamic split, pageable: 1); DCN_PACK_FRACTION overrides
    uint64_t n_packed_chunks = 0, n_ascii_chunks = 0, n_uniform_chunks = 0;
    std::atomic<uint64_t> launches{0};
    float t_h2d = 0, t_kernel = 0, t_d2h = 0;
    uint64_t bytes_h2d = 0, bytes_d2h = 0;   // what the last host-pointer filter call moved over PCIe
    // CUDA-event pairs around every launch of the fused kernel (ring), for dcn_fused_time_take
    static const int KEV = 256;
    cudaEvent_t kev0[KEV], kev1[KEV];
    uint32_t kev_head = 0, kev_count = 0;

    std::mutex err_m;   // the pipeline's packer threads report errors too
    int fail(int code, const char *what, cudaError_t e = cudaSuccess) {
        std::lock_guard<std::mutex> g(err_m);
        err = what;
        if (e != cudaSuccess) { err += ": "; err += cudaGetErrorString(e); }
        return code;
    }
};

#define CK(call)                                                            \
    do {                                                                    \
        cudaError_t e_ = (call);                                            \
        if (e_ != cudaSuccess) return ctx->fail(DCN_ERR_CUDA, #call, e_);   \
    } while (0)

using G31 = Geo<31, 15>;

// A distinct-hit set of `entries` 16-byte slots in `buf` for one call: cleared only the first time the bytes are used
// (and when the 32-bit epoch wraps); after that a new epoch makes every older entry count as free.
static cudaError_t open_dedup_set(DevBuf &buf, uint64_t entries, cudaStream_t st, DedupView &dd, uint32_t *overflow_flag) {
    cudaError_t e = buf.ensure(entries * 16);
    if (e != cudaSuccess) return e;
    if (buf.set_cleared < entries * 16 || buf.set_epoch == 0xFFFFFFFFu) {
        if ((e = cudaMemsetAsync(buf.p, 0, buf.cap, st)) != cudaSuccess) return e;
        buf.set_cleared = buf.cap;
        buf.set_epoch = 0;
    }
    dd.slots = buf.as<unsigned __int128>(); dd.cap = entries; dd.overflow = overflow_flag; dd.epoch = ++buf.set_epoch; dd.per16 = 0;
    return cudaSuccess;
}

static size_t warp_kernel_smem() { return ((sizeof(WarpTables) + 15) & ~(size_t)15) + DCN_WARPS * (sizeof(WarpPipe) + sizeof(WarpSmem)); }

// warp-tile plan: the 64-byte header + the tile list.  A tile ends because the next unit does not fit (it then spans
// more than TB - 15 - DCN_MAX_SHORT bases), because it holds MAXR records, because a long unit follows, or at a
// segment end.
static uint64_t wplan_tile_cap(uint64_t n_rel, uint32_t n_rec) {
    return n_rel / (uint64_t)(WG::TB - 15 - (int)DCN_MAX_SHORT) + (uint64_t)n_rec / (WG::MAXR / 2) + n_rel / DCN_MAX_SHORT + n_rel / DCN_WSEG + 16 +
           n_rel / DCN_WCS + n_rel / DCN_MAX_SHORT;   // + the chunks of long units: one per DCN_WCS windows and one partial per record
}
static uint64_t wplan_ovf_cap(uint64_t n_rel) { return n_rel / WG::PKCAP + 16; }   // a unit of more than PKCAP picks has more than PKCAP bases

static size_t plan_bytes(uint64_t n_rel_bases) {
    // smallest stride the planner can choose -> most tiles
    uint64_t s_min = (uint64_t)G31::BCAP - 14u - DCN_MAX_SHORT;
    uint64_t n_tiles = n_rel_bases / s_min + 2;
    return 64 + (size_t)n_tiles * 2 * sizeof(uint32_t);
}

// src/minimizers.rs:73-121 evaluated on the host for every base-count triple: the same f32
// operations in the same order (p = count / total; entropy -= p * log2f(p); entropy / 2 >= thr).
// `stride` values per axis: 32 for the k = 31 tile kernel, 64 for the generic path (k <= 57).
static void build_entropy_bitmap(int k, float thr, uint32_t stride, std::vector<uint32_t> &bits) {
    bits.assign((size_t)stride * stride * stride / 32, 0);
    for (int a = 0; a <= k; a++)
        for (int c = 0; a + c <= k; c++)
            for (int g = 0; a + c + g <= k; g++) {
                int counts[4] = {a, c, g, k - a - c - g};   // A, C, G, T order of the reference
                volatile float entropy = 0.0f;
                float total_f = (float)k;
                if (k >= 10) {
                    for (int i = 0; i < 4; i++)
                        if (counts[i] > 0) {
                            volatile float p = (float)counts[i] / total_f;
                            volatile float term = p * log2f(p);
                            entropy = entropy - term;
                        }
                }
                float scaled = k < 10 ? 1.0f : entropy / 2.0f;
                if (scaled >= thr) {
                    uint32_t idx = ((uint32_t)a * stride + (uint32_t)c) * stride + (uint32_t)g;
                    bits[idx >> 5] |= 1u << (idx & 31);
                }
            }
}

// (k, w) accepted by the reference: k + w - 1 odd (src/index.rs:186-194 and the upstream assert),
// k <= 56 when filtering (src/filter_common.rs:269-272), k <= 57 at index time (src/main.rs:166).
static int check_kw(dcn_ctx *ctx, int k, int w, int flavour) {
    const int kmax = flavour == DCN_FLAVOUR_INDEX ? 57 : 56;
    if (k < 1 || k > kmax) return ctx->fail(DCN_ERR_UNSUPPORTED, flavour == DCN_FLAVOUR_INDEX ? "k must be in 1..=57" : "k must be in 1..=56 for filtering");
    if (w < 1 || w > DCN_MAX_W) return ctx->fail(DCN_ERR_UNSUPPORTED, "w must be in 1..=255");
    if (((k + w - 1) & 1) == 0) return ctx->fail(DCN_ERR_ARG, "k + w - 1 must be odd");
    return DCN_OK;
}

// chunk plan of the generic path: per-record chunk counts -> exclusive scan (in place); one sync
static int generic_plan(dcn_ctx *ctx, GenericBatch &B, int flavour, uint64_t *rc, DevBuf &tmp, cudaStream_t st, uint64_t *n_chunks) {
    const int pb = 256;
    const int pg = (int)std::max<uint64_t>(1, std::min<uint64_t>(((uint64_t)B.n_rec + 1 + pb - 1) / pb, (uint64_t)ctx->sm_count * 8));
    if (flavour == DCN_FLAVOUR_INDEX) generic_rec_chunks_kernel<FLAVOUR_INDEX><<<pg, pb, 0, st>>>(B, rc);
    else generic_rec_chunks_kernel<FLAVOUR_FILTER><<<pg, pb, 0, st>>>(B, rc);
    size_t tb = 0;
    CK(cub::DeviceScan::ExclusiveSum(nullptr, tb, rc, rc, (int64_t)B.n_rec + 1, st));
    CK(tmp.ensure(tb));
    CK(cub::DeviceScan::ExclusiveSum(tmp.p, tb, rc, rc, (int64_t)B.n_rec + 1, st));
    ctx->launches += 3;
    CK(cudaMemcpyAsync(n_chunks, rc + B.n_rec, sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    B.rec_chunk_off = rc;
    return DCN_OK;
}

static int grid_for(dcn_ctx *ctx, uint64_t n, int block) {
    return (int)std::max<uint64_t>(1, std::min<uint64_t>((n + block - 1) / block, (uint64_t)ctx->sm_count * 16));
}

// Generic extraction into device CSR buffers (ctx->gx_h / gx_p / gx_oo).  *n_out = minimizers.
// If `cap_limit` is non-zero and exceeded, only the offsets are produced (the caller reports overflow).
static int generic_extract_device(dcn_ctx *ctx, int flavour, const uint8_t *d_bases, const uint64_t *d_off, uint32_t n_rec,
                                  int k, int w, uint32_t prefix_len, float entropy_thr, bool want_pos, uint64_t cap_limit,
                                  cudaStream_t st, uint64_t *n_out, bool *written) {
    *n_out = 0; *written = false;
    GenericBatch B;
    B.bases = d_bases; B.base0 = 0; B.rec_off = d_off; B.n_rec = n_rec; B.prefix_len = prefix_len;
    B.k = k; B.w = w; B.cstride = DCN_GENERIC_CSTRIDE; B.entropy_pass = nullptr; B.rec_chunk_off = nullptr;
    if (flavour == DCN_FLAVOUR_INDEX && entropy_thr != 0.0f) {
        std::vector<uint32_t> bits;
        build_entropy_bitmap(k, entropy_thr, 64, bits);
        CK(ctx->gx_entropy.ensure(bits.size() * 4));
        CK(cudaMemcpyAsync(ctx->gx_entropy.p, bits.data(), bits.size() * 4, cudaMemcpyHostToDevice, st));
        CK(cudaStreamSynchronize(st));   // `bits` dies with this scope
        B.entropy_pass = ctx->gx_entropy.as<uint32_t>();
    }
    CK(ctx->gx_rc.ensure(((size_t)n_rec + 1) * 8));
    CK(ctx->gx_oo.ensure(((size_t)n_rec + 1) * 8));
    uint64_t n_chunks = 0;
    int rc = generic_plan(ctx, B, flavour, ctx->gx_rc.as<uint64_t>(), ctx->gx_tmp, st, &n_chunks);
    if (rc) return rc;
    CK(ctx->gx_cc.ensure((n_chunks + 1) * 8));
    uint64_t *cc = ctx->gx_cc.as<uint64_t>();
    const int g1 = grid_for(ctx, n_chunks + 1, 128);
    if (flavour == DCN_FLAVOUR_INDEX) generic_count_kernel<FLAVOUR_INDEX><<<g1, 128, 0, st>>>(B, n_chunks, cc);
    else generic_count_kernel<FLAVOUR_FILTER><<<g1, 128, 0, st>>>(B, n_chunks, cc);
    size_t tb = 0;
    CK(cub::DeviceScan::ExclusiveSum(nullptr, tb, cc, cc, (int64_t)n_chunks + 1, st));
    CK(ctx->gx_tmp.ensure(tb));
    CK(cub::DeviceScan::ExclusiveSum(ctx->gx_tmp.p, tb, cc, cc, (int64_t)n_chunks + 1, st));
    generic_rec_off_kernel<<<grid_for(ctx, (uint64_t)n_rec + 1, 256), 256, 0, st>>>(ctx->gx_rc.as<uint64_t>(), cc, n_rec, ctx->gx_oo.as<uint64_t>());
    ctx->launches += 4;
    uint64_t m = 0;
    CK(cudaMemcpyAsync(&m, cc + n_chunks, sizeof(m), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    CK(cudaGetLastError());
    *n_out = m;
    if (cap_limit && m > cap_limit) return DCN_OK;
    CK(ctx->gx_h.ensure(std::max<uint64_t>(m, 1) * 8));
    if (want_pos) CK(ctx->gx_p.ensure(std::max<uint64_t>(m, 1) * 4));
    if (n_chunks) {
        const int g2 = grid_for(ctx, n_chunks, 128);
        uint32_t *pp = want_pos ? ctx->gx_p.as<uint32_t>() : nullptr;
        if (flavour == DCN_FLAVOUR_INDEX) generic_write_kernel<FLAVOUR_INDEX><<<g2, 128, 0, st>>>(B, n_chunks, cc, ctx->gx_h.as<uint64_t>(), pp);
        else generic_write_kernel<FLAVOUR_FILTER><<<g2, 128, 0, st>>>(B, n_chunks, cc, ctx->gx_h.as<uint64_t>(), pp);
        ctx->launches += 1;
        CK(cudaGetLastError());
    }
    *written = true;
    return DCN_OK;
}

// B1 for indexes whose (k, w) is not the specialised (31, 15)
static int enqueue_filter_generic(dcn_ctx *ctx, DevBuf &plan, DevBuf &tmp, DevBuf &dedup, const uint8_t *d_bases, uint64_t base0,
                                  uint64_t n_bases_abs, const uint64_t *d_off, uint32_t n_rec, uint32_t rpu, uint32_t n_units,
                                  uint32_t prefix_len, uint32_t abs_thr, double rel_thr, int deplete, uint8_t *d_keep,
                                  uint32_t *d_hits, uint32_t *d_total, cudaStream_t st) {
    int rc = check_kw(ctx, ctx->k, ctx->w, DCN_FLAVOUR_FILTER);
    if (rc) return rc;
    GenericBatch B;
    B.bases = d_bases; B.base0 = base0; B.rec_off = d_off; B.n_rec = n_rec; B.prefix_len = prefix_len;
    B.k = ctx->k; B.w = ctx->w; B.cstride = DCN_GENERIC_CSTRIDE; B.entropy_pass = nullptr; B.rec_chunk_off = nullptr;
    CK(plan.ensure(64 + ((size_t)n_rec + 1) * 8));
    BatchStats *d_stats = plan.as<BatchStats>();
    uint64_t *rcoff = reinterpret_cast<uint64_t *>(plan.as<uint8_t>() + 64);
    uint64_t n_chunks = 0;
    if ((rc = generic_plan(ctx, B, DCN_FLAVOUR_FILTER, rcoff, tmp, st, &n_chunks))) return rc;
    TableView tv;
    tv.slots = ctx->table.as<uint64_t>(); tv.n_buckets = ctx->n_buckets; tv.has_empty_key = ctx->has_empty;
    const uint64_t n_rel = n_bases_abs - base0;
    // expected picks ~ 2 / (w + 1) per base; twice that many slots, grown x4 on overflow
    uint64_t dedup_cap = std::max<uint64_t>(4096, 4 * n_rel / ((uint64_t)ctx->w + 1));
    for (int attempt = 0; attempt < 4; attempt++) {
        CK(cudaMemsetAsync(d_stats, 0, sizeof(BatchStats), st));
        CK(cudaMemsetAsync(d_hits, 0, (size_t)n_units * 4, st));
        CK(cudaMemsetAsync(d_total, 0, (size_t)n_units * 4, st));
        DedupView dd;
        CK(open_dedup_set(dedup, dedup_cap, st, dd, &d_stats->overflow));
        if (n_chunks) {
            generic_filter_kernel<<<grid_for(ctx, n_chunks, 128), 128, 0, st>>>(B, n_chunks, rpu, tv, dd, d_hits, d_total);
            ctx->launches += 1;
        }
        BatchStats after;
        CK(cudaMemcpyAsync(&after, d_stats, sizeof(after), cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        CK(cudaGetLastError());
        if (!after.overflow) break;
        if (attempt == 3) return ctx->fail(DCN_ERR_OVERFLOW, "distinct-hit set overflowed after 4 attempts");
        dedup_cap *= 4;
    }
    generic_finalize_kernel<<<grid_for(ctx, n_units, 256), 256, 0, st>>>(n_units, d_hits, d_total, abs_thr, rel_thr, deplete, d_keep);
    stats_kernel<<<grid_for(ctx, n_units, 256), 256, 0, st>>>(d_off, rpu, n_units, d_keep, ctx->counters.as<unsigned long long>());
    ctx->launches += 2;
    CK(cudaGetLastError());
    return DCN_OK;
}

// Enqueue the whole filter pipeline for one device-resident batch on `st`.
// `longs` / `dedup` are scratch for the long path (units > DCN_MAX_SHORT bases).
static int enqueue_filter(dcn_ctx *ctx, DevBuf &plan, DevBuf &longs, DevBuf &dedup, const FilterInput &in,
                          uint64_t base0, uint64_t n_bases_abs, const uint64_t *d_off, uint32_t n_rec, int paired,
                          uint32_t prefix_len, uint32_t abs_thr, double rel_thr, int deplete, uint8_t *d_keep,
                          uint32_t *d_hits, uint32_t *d_total, cudaStream_t st, const BatchStats *host_stats = nullptr,
                          bool time_fused = true, bool promised_short = false,   // false: no event pair around the fused kernel (the ring is not thread-safe)
                          BatchStats *async_after = nullptr, uint32_t dedup_grow = 1) {
    // async_after (pinned host memory): a batch with long units is enqueued without waiting for it -- the batch's stats
    // (with the overflow flag of the distinct-hit set) are copied there behind the kernels and the CALLER looks at them once
    // the stream has got that far; on overflow it calls again with dedup_grow = 4 (and no async_after: that call retries
    // by itself).  Without it the call waits for the long path and retries with a set four times the size.
    if (!ctx->table.p) return ctx->fail(DCN_ERR_NO_INDEX, "no index resident: call dcn_index_upload first");
    const uint32_t rpu = paired ? 2u : 1u;
    if (paired && (n_rec & 1u)) return ctx->fail(DCN_ERR_ARG, "paired batch needs an even record count");
    const uint32_t n_units = n_rec / rpu;
    if (n_units == 0) return DCN_OK;
    const uint8_t *d_bases = in.bases;
    if (ctx->k != 31 || ctx->w != 15) {   // the tile kernel is specialised for the default parameters
        if (!d_bases) return ctx->fail(DCN_ERR_ARG, "packed input is only implemented for k=31, w=15");
        return enqueue_filter_generic(ctx, plan, longs, dedup, d_bases, base0, n_bases_abs, d_off, n_rec, rpu, n_units,
                                      prefix_len, abs_thr, rel_thr, deplete, d_keep, d_hits, d_total, st);
    }
    if ((d_bases && (reinterpret_cast<uintptr_t>(d_bases) & 15u)) || (base0 & 15u))
        return ctx->fail(DCN_ERR_ARG, "d_bases must be 16-byte aligned");

    const uint64_t n_rel = n_bases_abs - base0;
    const bool warp_impl = ctx->fused_impl == 0;
    // plan buffer: 64-byte header (BatchStats), then the tile plan.  CTA tiles: tile_first | tile_end (cleared per call);
    // warp tiles: the tile list | the list of units handed to the CTA path (only the header is cleared).
    const uint64_t wtile_cap = wplan_tile_cap(n_rel, n_rec), wovf_cap = wplan_ovf_cap(n_rel);
    const size_t pbytes = warp_impl ? 128 + (size_t)wtile_cap * sizeof(WTile) + (size_t)wovf_cap * 4 : plan_bytes(n_rel);
    CK(plan.ensure(pbytes));
    const uint64_t n_tiles_max = (plan_bytes(n_rel) - 64) / (2 * sizeof(uint32_t));
    BatchStats *d_stats = plan.as<BatchStats>();
    uint32_t *tile_first = reinterpret_cast<uint32_t *>(plan.as<uint8_t>() + 64);
    uint32_t *tile_end = tile_first + n_tiles_max;
    // warp tiles: bytes 64..127 hold this call's share of the six summary counters; it is added to the ctx's counters
    // when the call's last attempt is enqueued (a retry after a distinct-hit set overflow starts it from zero again)
    unsigned long long *call_cnt = reinterpret_cast<unsigned long long *>(plan.as<uint8_t>() + 64);
    WTile *wtiles = reinterpret_cast<WTile *>(plan.as<uint8_t>() + 128);
    uint32_t *wovf = reinterpret_cast<uint32_t *>(plan.as<uint8_t>() + 128 + (size_t)wtile_cap * sizeof(WTile));

    FilterParams P;
    P.bases = d_bases; P.pk_codes = in.codes; P.pk_inv = in.inv; P.nl_bits = in.nl; P.nl_bit0 = in.nl_bit0; P.base0 = base0; P.n_bases = n_bases_abs;
    P.rec_off = d_off; P.n_rec = n_rec; P.rpu = rpu; P.n_units = n_units;
    P.prefix_len = prefix_len; P.abs_thr = abs_thr; P.rel_thr = rel_thr; P.deplete = deplete;
    P.table.slots = ctx->table.as<uint64_t>(); P.table.n_buckets = ctx->n_buckets; P.table.has_empty_key = ctx->has_empty;
    P.keep = d_keep; P.hits = d_hits; P.total = d_total;

    const int pb = 256;
    const int pg = (int)std::min<uint64_t>((n_units + pb - 1) / pb, (uint64_t)ctx->sm_count * 8);
    const size_t smem = sizeof(TileSmem<G31>);
    const uint64_t tiles_lb = n_rel / G31::BCAP + 1;
    const int grid = (int)std::min<uint64_t>(tiles_lb, (uint64_t)ctx->sm_count * DCN_CTAS_PER_SM);
    // Warp-tile grid: one CTA per SM.  DCN_TILES_PER_WARP = n (A/B knob, off by default) gives a small batch only as many
    // CTAs as leave every warp n tiles, so that kernels of different streams run side by side on disjoint SMs.  Measured
    // (16.8 Mbp launches, tools/chunk_cost.py): one stream 150 us -> 193 us (n = 4) -> 298 us (n = 8); thirteen streams of
    // the chunk pipeline: no gain beyond noise.  Small launches are slow for another reason -- all warps of an SM are in
    // the same phase, so probe latency and ALU work do not overlap -- and the cure is fewer, larger launches
    // (filter_pipeline_arena), not narrower ones.
    static const uint64_t tiles_per_warp = []() { const char *e = getenv("DCN_TILES_PER_WARP"); return e ? (uint64_t)std::max(0, atoi(e)) : 0ull; }();
    int wgrid = (int)std::min<uint64_t>((n_rel / WG::TB + DCN_WARPS) / DCN_WARPS, (uint64_t)ctx->sm_count * DCN_WCTAS);
    if (tiles_per_warp) wgrid = (int)std::max<uint64_t>(1, std::min<uint64_t>((uint64_t)wgrid, (n_rel / WG::TB + 1 + DCN_WARPS * tiles_per_warp - 1) / (DCN_WARPS * tiles_per_warp)));
    const uint64_t n_seg = (n_rel + DCN_WSEG - 1) / DCN_WSEG;
    const int sg = (int)std::max<uint64_t>(1, std::min<uint64_t>((n_seg + 7) / 8, (uint64_t)ctx->sm_count * 8));   // 8 warps per CTA, one segment per warp

    // A batch can only contain a long unit if it holds more than DCN_MAX_SHORT bases; otherwise the
    // stats readback (one small sync) is skipped.
    uint64_t dedup_cap = 0;
    for (int attempt = 0; attempt < 4; attempt++) {
        CK(cudaMemsetAsync(plan.p, 0, warp_impl ? 128 : pbytes, st));
        if (warp_impl) {   // the planner also counts the long units
            wplan_kernel<<<sg, 256, 0, st>>>(d_off, rpu, n_units, base0, n_rel, d_stats, wtiles, (uint32_t)std::min<uint64_t>(wtile_cap, 0xFFFFFFFFull),
                                             promised_short ? ctx->h_promise : nullptr);
            ctx->launches += 1;
        } else {
            prep_stats_kernel<<<pg, pb, 0, st>>>(d_off, rpu, n_units, d_stats);
            prep_tiles_kernel<G31><<<pg, pb, 0, st>>>(d_off, rpu, n_units, base0, d_stats, tile_first, tile_end);
            ctx->launches += 2;
        }
        BatchStats hs;
        memset(&hs, 0, sizeof(hs));
        if (host_stats) {
            hs = *host_stats;   // the host-pointer pipeline knows the unit lengths: no readback, no sync
        } else if (promised_short && warp_impl) {
            // the caller vouches for short units only: nothing to read back (the planner reports a broken promise)
        } else if (n_rel > DCN_MAX_SHORT) {
            CK(cudaMemcpyAsync(&hs, d_stats, sizeof(hs), cudaMemcpyDeviceToHost, st));
            CK(cudaStreamSynchronize(st));
        }
        DedupView dd;
        dd.slots = nullptr; dd.cap = 0; dd.overflow = &d_stats->overflow; dd.epoch = 1; dd.per16 = 0;
        uint32_t *long_units = nullptr;
        ChunkDesc *desc = nullptr;
        if (hs.n_long) {
            // distinct hits of long units go through a global (hash, unit) set: only hits are inserted and
            // picks are ~0.13 per base, so 0.25 entries per base is at most half full; it grows x4 on
            // overflow (exactness is kept by retrying, never by dropping).  The set is cleared per call:
            // its size is what the long path pays up front (4 bytes per long base).
            // Warp tiles, batch mostly long units: per-unit regions of the set, laid out by position in the batch
            // (DedupView::per16; 4 slots per 16 bases = the same 0.25 per base); a batch with few long units keeps the one
            // small set (regions would reserve 4 bytes per base of the short units too).  DCN_DEDUP_LOCAL=0 forces the latter.
            static const bool local_ok = []() { const char *e = getenv("DCN_DEDUP_LOCAL"); return !e || atoi(e) != 0; }();
            const bool local = warp_impl && local_ok && hs.long_bases >= n_rel / 4;
            static const uint32_t shrink = []() { const char *e = getenv("DCN_DEDUP_SHRINK"); return e ? (uint32_t)std::max(1, atoi(e)) : 1u; }();   // tests: start too small
            if (!dedup_cap) dedup_cap = (local ? std::max<uint64_t>(1, 4 / shrink) : std::max<uint64_t>(64, std::max<uint64_t>(4096, hs.long_bases / 4) / shrink)) * dedup_grow;   // local: slots per 16 bases
            const uint32_t desc_cap = (uint32_t)(hs.long_bases / ChunkGeo<G31>::CSTRIDE + (uint64_t)hs.n_long * rpu + 16);
            CK(open_dedup_set(dedup, local ? ((n_rel >> 4) + 2) * dedup_cap : dedup_cap, st, dd, &d_stats->overflow));
            if (local) dd.per16 = (uint32_t)dedup_cap;
            CK(longs.ensure((size_t)hs.n_long * 4 + 64 + (size_t)desc_cap * sizeof(ChunkDesc)));
            long_units = longs.as<uint32_t>();
            desc = reinterpret_cast<ChunkDesc *>(longs.as<uint8_t>() + (((size_t)hs.n_long * 4 + 63) & ~(size_t)63));
            if (warp_impl) prep_long_warp_kernel<G31><<<pg, pb, 0, st>>>(P, d_stats, long_units, wtiles, (uint32_t)std::min<uint64_t>(wtile_cap, 0xFFFFFFFFull));
            else prep_long_kernel<G31><<<pg, pb, 0, st>>>(P, d_stats, long_units, desc, desc_cap);
            ctx->launches += 1;
        }
        const uint32_t ke = ctx->kev_head % dcn_ctx::KEV;
        if (time_fused) CK(cudaEventRecord(ctx->kev0[ke], st));
        if (warp_impl) {
            // short units: warp tiles; then the CTA-tile tail (units of more than a warp pass's picks, long chunks).  With
            // no long unit in the batch the tail only has work on pathological input, so a few CTAs are enough.
            const uint32_t ocap = (uint32_t)std::min<uint64_t>(wovf_cap, 0xFFFFFFFFull);
            unsigned long long *cnt = call_cnt;
            const size_t wsm = warp_kernel_smem();
            if (hs.n_long) {
                if (in.codes) filter_warp_kernel<true, true><<<wgrid, DCN_WARPS * 32, wsm, st>>>(P, d_stats, wtiles, wovf, ocap, cnt, dd);
                else filter_warp_kernel<false, true><<<wgrid, DCN_WARPS * 32, wsm, st>>>(P, d_stats, wtiles, wovf, ocap, cnt, dd);
            } else {
                if (in.codes) filter_warp_kernel<true, false><<<wgrid, DCN_WARPS * 32, wsm, st>>>(P, d_stats, wtiles, wovf, ocap, cnt, dd);
                else filter_warp_kernel<false, false><<<wgrid, DCN_WARPS * 32, wsm, st>>>(P, d_stats, wtiles, wovf, ocap, cnt, dd);
            }
            const int tgrid = std::min(grid, 16);   // overflow units only (pathological input): the long chunks are warp tiles
            if (in.codes) filter_tail_kernel<G31, true><<<tgrid, G31::NT, smem, st>>>(P, d_stats, wovf, dd, desc, call_cnt);
            else filter_tail_kernel<G31, false><<<tgrid, G31::NT, smem, st>>>(P, d_stats, wovf, dd, desc, call_cnt);
            ctx->launches += 1;
        } else if (in.codes) filter_fused_kernel<G31, true><<<grid, G31::NT, smem, st>>>(P, d_stats, tile_first, tile_end, dd, desc);
        else filter_fused_kernel<G31, false><<<grid, G31::NT, smem, st>>>(P, d_stats, tile_first, tile_end, dd, desc);
        if (time_fused) {
            CK(cudaEventRecord(ctx->kev1[ke], st));
            ctx->kev_head++;
            if (ctx->kev_count < dcn_ctx::KEV) ctx->kev_count++;
        }
        ctx->launches += 1;
        if (hs.n_long) {
            finalize_long_kernel<<<std::max(1, (int)std::min<uint32_t>((hs.n_long + 255) / 256, 1024)), 256, 0, st>>>(P, d_stats, long_units, warp_impl ? call_cnt : nullptr);
            ctx->launches += 1;
            if (async_after && warp_impl) {
                CK(cudaMemcpyAsync(async_after, d_stats, sizeof(BatchStats), cudaMemcpyDeviceToHost, st));
                break;
            }
            BatchStats after;
            CK(cudaMemcpyAsync(&after, d_stats, sizeof(after), cudaMemcpyDeviceToHost, st));
            CK(cudaStreamSynchronize(st));
            if (after.overflow) {
                if (attempt == 3) return ctx->fail(DCN_ERR_OVERFLOW, "distinct-hit set overflowed after 4 attempts");
                dedup_cap *= 4;
                continue;
            }
        }
        break;
    }
    if (warp_impl) commit_counters_kernel<<<1, 32, 0, st>>>(call_cnt, ctx->counters.as<unsigned long long>(), async_after ? d_stats : nullptr);   // the warp-tile kernels counted on the way
    else stats_kernel<<<pg, pb, 0, st>>>(d_off, rpu, n_units, d_keep, ctx->counters.as<unsigned long long>());
    ctx->launches += 1;
    CK(cudaGetLastError());
    return DCN_OK;
}

extern "C" {

int dcn_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
    return n;
}

dcn_ctx *dcn_ctx_create(int device) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n <= 0) {
        g_create_error = std::string("no CUDA device available (there is no CPU fallback): ") + cudaGetErrorString(e);
        return nullptr;
    }
    if (device < 0 || device >= n) { g_create_error = "device index out of range"; return nullptr; }
    if ((e = cudaSetDevice(device)) != cudaSuccess) { g_create_error = cudaGetErrorString(e); return nullptr; }
    cudaDeviceProp prop;
    if ((e = cudaGetDeviceProperties(&prop, device)) != cudaSuccess) { g_create_error = cudaGetErrorString(e); return nullptr; }
    if (prop.major < 10) {
        g_create_error = "this library is built for sm_100a (B200) only; device is sm_" + std::to_string(prop.major * 10 + prop.minor);
        return nullptr;
    }
    dcn_ctx *ctx = new dcn_ctx();
    memset(ctx->kev0, 0, sizeof(ctx->kev0));
    memset(ctx->kev1, 0, sizeof(ctx->kev1));
    ctx->device = device;
    ctx->sm_count = prop.multiProcessorCount;
    bool ok = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) == cudaSuccess;
    for (int i = 0; ok && i < dcn_ctx::NSLOT; i++) {
        Slot &s = ctx->slot[i];
        ok = ok && cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking) == cudaSuccess;
        ok = ok && cudaEventCreate(&s.ev_start) == cudaSuccess && cudaEventCreate(&s.ev_h2d) == cudaSuccess;
        ok = ok && cudaEventCreate(&s.ev_kernel) == cudaSuccess && cudaEventCreate(&s.ev_done) == cudaSuccess;
    }
    for (int i = 0; ok && i < dcn_ctx::KEV; i++)
        ok = ok && cudaEventCreate(&ctx->kev0[i]) == cudaSuccess && cudaEventCreate(&ctx->kev1[i]) == cudaSuccess;
    ok = ok && ctx->counters.ensure(8 * sizeof(unsigned long long)) == cudaSuccess;
    ok = ok && cudaMemset(ctx->counters.p, 0, 8 * sizeof(unsigned long long)) == cudaSuccess;
    ok = ok && cudaMallocHost(reinterpret_cast<void **>(&ctx->h_promise), 64) == cudaSuccess;
    if (ok) *ctx->h_promise = 0;
    ok = ok && cudaFuncSetAttribute(filter_fused_kernel<G31, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    (int)sizeof(TileSmem<G31>)) == cudaSuccess;
    ok = ok && cudaFuncSetAttribute(filter_fused_kernel<G31, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    (int)sizeof(TileSmem<G31>)) == cudaSuccess;
    ok = ok && cudaFuncSetAttribute(filter_warp_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)warp_kernel_smem()) == cudaSuccess;
    ok = ok && cudaFuncSetAttribute(filter_warp_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)warp_kernel_smem()) == cudaSuccess;
    ok = ok && cudaFuncSetAttribute(filter_warp_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)warp_kernel_smem()) == cudaSuccess;
    ok = ok && cudaFuncSetAttribute(filter_warp_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)warp_kernel_smem()) == cudaSuccess;
    ok = ok && cudaFuncSetAttribute(filter_tail_kernel<G31, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    (int)sizeof(TileSmem<G31>)) == cudaSuccess;
    ok = ok && cudaFuncSetAttribute(filter_tail_kernel<G31, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    (int)sizeof(TileSmem<G31>)) == cudaSuccess;
    if (const char *fi = getenv("DCN_FUSED_IMPL")) ctx->fused_impl = strcmp(fi, "cta") == 0 ? 1 : 0;
    ok = ok && cudaFuncSetAttribute(extract_warp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)warp_kernel_smem()) == cudaSuccess;
    ok = ok && cudaFuncSetAttribute(extract_tail_kernel<G31>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    (int)sizeof(TileSmem<G31>)) == cudaSuccess;
    ok = ok && cudaFuncSetAttribute(extract_tiles_kernel<G31>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    (int)sizeof(TileSmem<G31>)) == cudaSuccess;
    ok = ok && cudaFuncSetAttribute(extract_index_kernel<G31>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    (int)sizeof(TileSmem<G31>)) == cudaSuccess;
    if (!ok) {
        g_create_error = std::string("context setup failed: ") + cudaGetErrorString(cudaGetLastError());
        dcn_ctx_destroy(ctx);
        return nullptr;
    }
    return ctx;
}

void dcn_ctx_destroy(dcn_ctx *ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaDeviceSynchronize();
    ctx->table.release(); ctx->plan.release(); ctx->counters.release(); ctx->keys_stage.release();
    ctx->longs.release(); ctx->dedup.release();
    ctx->ib_bases.release(); ctx->ib_off.release(); ctx->ib_desc.release(); ctx->ib_keys.release();
    ctx->ib_alt.release(); ctx->ib_tmp.release(); ctx->ib_entropy.release(); ctx->ib_stats.release(); ctx->ib_runs.release();
    ctx->gx_bases.release(); ctx->gx_off.release(); ctx->gx_rc.release(); ctx->gx_cc.release(); ctx->gx_tmp.release();
    ctx->gx_h.release(); ctx->gx_p.release(); ctx->gx_oo.release(); ctx->gx_entropy.release();
    ctx->ws_table.release(); ctx->ws_flags.release();
    for (size_t i = 0; i < dcn_ctx::NSLOT + ctx->pslot.size() + ctx->aslot.size(); i++) {
        Slot &s = i < dcn_ctx::NSLOT ? ctx->slot[i] : i < dcn_ctx::NSLOT + ctx->pslot.size() ? ctx->pslot[i - dcn_ctx::NSLOT] : ctx->aslot[i - dcn_ctx::NSLOT - ctx->pslot.size()];
        s.in.release(); s.out.release(); s.plan.release(); s.longs.release(); s.dedup.release();
        s.h_in.release(); s.h_out.release();
        if (s.stream) cudaStreamDestroy(s.stream);
        if (s.ev_start) cudaEventDestroy(s.ev_start);
        if (s.ev_h2d) cudaEventDestroy(s.ev_h2d);
        if (s.ev_kernel) cudaEventDestroy(s.ev_kernel);
        if (s.ev_done) cudaEventDestroy(s.ev_done);
    }
    if (ArenaState *ar = ctx->arena) {
        ar->ascii.release(); ar->codes.release(); ar->inv.release(); ar->nl.release(); ar->off.release(); ar->out.release();
        for (int i = 0; i < ArenaState::NL; i++) {
            Slot &s = ar->launch[i];
            s.plan.release(); s.longs.release(); s.dedup.release(); s.h_out.release();
            if (s.stream) cudaStreamDestroy(s.stream);
            if (s.ev_start) cudaEventDestroy(s.ev_start);
            if (s.ev_kernel) cudaEventDestroy(s.ev_kernel);
            if (s.ev_done) cudaEventDestroy(s.ev_done);
            if (ar->ev_front[i]) cudaEventDestroy(ar->ev_front[i]);
        }
        for (int i = 0; i < 4; i++) if (ar->ev_sub[i]) cudaEventDestroy(ar->ev_sub[i]);
        if (ar->h_after) cudaFreeHost(ar->h_after);
        if (ar->copy_stream) cudaStreamDestroy(ar->copy_stream);
        delete ar;
    }
    for (int i = 0; i < dcn_ctx::KEV; i++) {
        if (ctx->kev0[i]) cudaEventDestroy(ctx->kev0[i]);
        if (ctx->kev1[i]) cudaEventDestroy(ctx->kev1[i]);
    }
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    if (ctx->h_promise) cudaFreeHost(ctx->h_promise);
    delete ctx;
}

const char *dcn_last_error(dcn_ctx *ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

void *dcn_host_alloc(size_t bytes) {
    void *p = nullptr;
    if (cudaMallocHost(&p, bytes ? bytes : 1) != cudaSuccess) return nullptr;
    return p;
}
void dcn_host_free(void *p) { if (p) cudaFreeHost(p); }

// ---------------------------------------------------------------------------- index residency
int dcn_index_set_load_factor(dcn_ctx *ctx, double load) {
    if (!ctx) return DCN_ERR_ARG;
    if (!(load >= 0.05 && load <= 0.9)) return ctx->fail(DCN_ERR_ARG, "load factor must be in [0.05, 0.9]");
    ctx->load = load;
    return DCN_OK;
}

static int table_begin(dcn_ctx *ctx, uint64_t n_keys, uint8_t k, uint8_t w, cudaStream_t st) {
    CK(cudaSetDevice(ctx->device));
    ctx->n_buckets = table_buckets_for(n_keys, ctx->load);
    ctx->n_keys = 0; ctx->has_empty = 0; ctx->k = k; ctx->w = w;
    CK(ctx->table.ensure(ctx->n_buckets * 4 * sizeof(uint64_t)));
    table_fill_kernel<<<ctx->sm_count * 8, 256, 0, st>>>(ctx->table.as<uint64_t>(), ctx->n_buckets * 4);
    CK(cudaMemsetAsync(ctx->counters.as<unsigned long long>() + 6, 0, 2 * sizeof(unsigned long long), st));
    ctx->launches += 1;
    return DCN_OK;
}
static int table_insert(dcn_ctx *ctx, const uint64_t *d_keys, uint64_t n, cudaStream_t st) {
    if (!n) return DCN_OK;
    int grid = (int)std::min<uint64_t>((n + 255) / 256, (uint64_t)ctx->sm_count * 16);
    table_insert_kernel<<<grid, 256, 0, st>>>(ctx->table.as<uint64_t>(), ctx->n_buckets, d_keys, n,
                                              ctx->counters.as<unsigned long long>() + 6);
    ctx->launches += 1;
    CK(cudaGetLastError());
    return DCN_OK;
}
static int table_end(dcn_ctx *ctx, cudaStream_t st) {
    unsigned long long c[2] = {0, 0};
    CK(cudaMemcpyAsync(c, ctx->counters.as<unsigned long long>() + 6, sizeof(c), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    ctx->has_empty = c[1] != 0;
    ctx->n_keys = c[0] + (c[1] ? 1 : 0);
    return DCN_OK;
}

int dcn_index_upload_device(dcn_ctx *ctx, const uint64_t *d_keys, uint64_t n_keys, uint8_t k, uint8_t w, void *stream) {
    if (!ctx) return DCN_ERR_ARG;
    if (!d_keys && n_keys) return ctx->fail(DCN_ERR_ARG, "null key pointer");
    cudaStream_t st = (cudaStream_t)stream;
    int rc = table_begin(ctx, n_keys, k, w, st);
    if (rc) return rc;
    if ((rc = table_insert(ctx, d_keys, n_keys, st))) return rc;
    return table_end(ctx, st);
}

int dcn_index_upload(dcn_ctx *ctx, const uint64_t *keys, uint64_t n_keys, uint8_t k, uint8_t w) {
    if (!ctx) return DCN_ERR_ARG;
    if (!keys && n_keys) return ctx->fail(DCN_ERR_ARG, "null key pointer");
    cudaStream_t st = ctx->stream;
    int rc = table_begin(ctx, n_keys, k, w, st);
    if (rc) return rc;
    const uint64_t CH = 32ull << 20;  // keys per staging chunk (256 MB)
    CK(ctx->keys_stage.ensure(std::min(CH, std::max<uint64_t>(n_keys, 1)) * sizeof(uint64_t)));
    for (uint64_t i = 0; i < n_keys; i += CH) {
        uint64_t n = std::min(CH, n_keys - i);
        CK(cudaMemcpyAsync(ctx->keys_stage.p, keys + i, n * sizeof(uint64_t), cudaMemcpyHostToDevice, st));
        if ((rc = table_insert(ctx, ctx->keys_stage.as<uint64_t>(), n, st))) return rc;
        CK(cudaStreamSynchronize(st));  // staging buffer is reused
    }
    return table_end(ctx, st);
}

int dcn_index_info(dcn_ctx *ctx, uint64_t *n_keys, uint8_t *k, uint8_t *w, uint64_t *table_bytes) {
    if (!ctx) return DCN_ERR_ARG;
    if (!ctx->table.p) return ctx->fail(DCN_ERR_NO_INDEX, "no index resident");
    if (n_keys) *n_keys = ctx->n_keys;
    if (k) *k = ctx->k;
    if (w) *w = ctx->w;
    if (table_bytes) *table_bytes = ctx->n_buckets * 4 * sizeof(uint64_t);
    return DCN_OK;
}

// ---------------------------------------------------------------------------- B1 filter
// a promise given to dcn_filter_batch_device_hint that the device found broken: reported once, by the next call
static int check_promise(dcn_ctx *ctx) {
    if (ctx->h_promise && *reinterpret_cast<volatile uint32_t *>(ctx->h_promise)) {
        *ctx->h_promise = 0;
        return ctx->fail(DCN_ERR_ARG, "an earlier dcn_filter_batch_device_hint call held a unit longer than the promised max_unit_len: "
                                      "its long units were not classified");
    }
    return DCN_OK;
}

int dcn_filter_batch_device_hint(dcn_ctx *ctx, const uint8_t *d_bases, const uint64_t *d_rec_off, uint32_t n_rec,
                                 uint64_t n_bases, int paired, uint32_t prefix_len, uint32_t abs_thr, double rel_thr,
                                 int deplete, uint8_t *d_keep, uint32_t *d_hits, uint32_t *d_total, void *stream,
                                 uint32_t max_unit_len) {
    if (!ctx) return DCN_ERR_ARG;
    int rc = check_promise(ctx);
    if (rc) return rc;
    CK(cudaSetDevice(ctx->device));
    cudaStream_t st = (cudaStream_t)stream;   // NULL = the legacy default stream, like any CUDA API
    FilterInput in;
    in.bases = d_bases;
    const bool promised = max_unit_len > 0 && max_unit_len <= DCN_MAX_SHORT;
    return enqueue_filter(ctx, ctx->plan, ctx->longs, ctx->dedup, in, 0, n_bases, d_rec_off, n_rec, paired,
                          prefix_len, abs_thr, rel_thr, deplete, d_keep, d_hits, d_total, st, nullptr, true, promised);
}

int dcn_filter_batch_device(dcn_ctx *ctx, const uint8_t *d_bases, const uint64_t *d_rec_off, uint32_t n_rec,
                            uint64_t n_bases, int paired, uint32_t prefix_len, uint32_t abs_thr, double rel_thr,
                            int deplete, uint8_t *d_keep, uint32_t *d_hits, uint32_t *d_total, void *stream) {
    return dcn_filter_batch_device_hint(ctx, d_bases, d_rec_off, n_rec, n_bases, paired, prefix_len, abs_thr, rel_thr, deplete,
                                        d_keep, d_hits, d_total, stream, 0);
}

static size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

static int pack_threads_of(dcn_ctx *ctx) {   // -> host threads available for packing (0 = ship ASCII only)
    if (ctx->pack_threads < 0) {
        const char *e = getenv("DCN_PACK_THREADS");
        // default: leave four hardware threads to the caller, the enqueueing thread and the driver's own threads
        // (measured on a 16-vCPU box: 10-14 packers 70-72 Gbp/s end to end; with all 16 the enqueueing thread starves)
        int hc = (int)std::thread::hardware_concurrency();
        cpu_set_t cs;   // a process bound to the CPUs next to its GPU (one rank per GPU) counts only those
        if (sched_getaffinity(0, sizeof(cs), &cs) == 0 && CPU_COUNT(&cs) > 0) hc = std::min(hc, (int)CPU_COUNT(&cs));
        // ... and a container's CPU quota counts too (cgroup v2 cpu.max, v1 cfs quota / period)
        auto quota_cpus = []() -> int {
            long long q = -1, per = 100000;
            if (FILE *f = fopen("/sys/fs/cgroup/cpu.max", "r")) {
                char buf[64] = {0};
                if (fscanf(f, "%63s %lld", buf, &per) >= 1 && strcmp(buf, "max") != 0) q = atoll(buf);
                fclose(f);
            } else if (FILE *f1 = fopen("/sys/fs/cgroup/cpu/cpu.cfs_quota_us", "r")) {
                if (fscanf(f1, "%lld", &q) != 1) q = -1;
                fclose(f1);
                if (FILE *f2 = fopen("/sys/fs/cgroup/cpu/cpu.cfs_period_us", "r")) {
                    if (fscanf(f2, "%lld", &per) != 1) per = 100000;
                    fclose(f2);
                }
            }
            return (q > 0 && per > 0) ? (int)std::max<long long>(1, (q + per - 1) / per) : 0;
        }();
        if (quota_cpus > 0) hc = std::min(hc, quota_cpus);
        int n = e ? atoi(e) : std::min(std::max(hc - 4, 1), 12);   // beyond ~12 the packers share out the host's DRAM bandwidth, not cores (measured: 12 = 16)
        ctx->pack_threads = std::max(0, std::min(n, 256));
    }
    return ctx->pack_threads;
}

// Host-pointer form (SURVEY.md 8f.1 "pinned double-buffered streams"), CHUNK form.  It serves the single-route cases
// (no packer threads; caller-packed input, in 64 / 128 MB chunks) and what the arena form declines (a handful of huge
// units, batches beyond DCN_ARENA_MAX_MB, DCN_PIPELINE=chunks); when both routes run, filter_pipeline_arena above takes
// the call: same routes, same claims from the two ends, but kernels over whatever contiguous range has arrived instead
// of one kernel chain per chunk.  The batch is cut into unit-aligned atoms of
// 4 MB; a chunk (a run of atoms) goes through a pipeline stage with its own stream: copy in, kernels, results out
// through a pinned blob, scattered into the caller's arrays when the stage is reused.  A chunk reaches the GPU by
// one of two routes:
//   ASCII   the bytes are copied as they are (1 B/bp over PCIe, no CPU work), the GPU converts them;
//   packed  a host thread packs the chunk (2-bit codes + non-ACGT bits, plus record offsets and newline flags:
//           what PackedSeqVec::from_ascii and the mask loop of src/filter_common.rs:238-258 compute) into one
//           pinned blob and 0.25 B/bp cross PCIe (sparse non-ACGT list; 0.375 with the dense mask).
// The copy engine and the host cores work at the same time (measured on the round-1 box: a pinned H2D stream keeps
// 55 GB/s beside 12 packing threads doing 64 GB/s, tools/hybrid_probe.py), so with pinned caller buffers the two
// routes share a batch dynamically.  The calling thread ships 32 MB ASCII chunks from the FRONT of the atom list
// through NSLOT stages, at the pace of the link.  Packer threads (alive for the call) claim atoms from its BACK,
// up to 16 MB at a time while plenty are left and single atoms near the end, pack them, and enqueue them on stages
// and streams of their own; the call ends when the two fronts meet.  Stages complete in any order: results are
// scattered by unit index.  Pageable caller buffers take the packed route only (a direct copy would be staged by
// the driver at ~8 GB/s).
struct HostSrc {   // caller's host buffers: ASCII, or already packed (dcn_filter_batch_packed)
    const uint8_t *bases = nullptr;
    const uint32_t *codes = nullptr;
    const uint16_t *inv = nullptr;
    const uint32_t *nl = nullptr;
    const uint32_t *exc = nullptr;   // packed, sparse form (dcn_filter_batch_packed_sparse): (block, mask) pairs instead of `inv`
    uint64_t n_exc = 0;
};

namespace {

struct ChunkPlan {
    uint32_t u0, u1, nr, nu;
    uint64_t a0, b1, nb;   // chunk origin (64-aligned), end, bases covered
    // blob layouts
    size_t n_words, o_inv, o_off_p, o_nl, in_packed, o_off_a, in_ascii, out_bytes;
    // sparse packed form (what the packer threads ship): codes | newline flags | exceptions | offsets on the wire and on
    // the device, where the dense non-ACGT bit array follows (cleared by a memset, the listed blocks written by a kernel)
    size_t s_nl, s_exc, s_off, s_inv, in_sparse, dev_sparse;
    uint32_t exc_cap;   // exception slots: 32-base blocks that hold a non-ACGT byte; a chunk with more goes dense
};

static void sparse_layout(ChunkPlan &c) {
    c.s_nl = align_up(c.n_words * 4, 8);
    c.s_exc = c.s_nl + align_up(((size_t)c.nr + 31) / 32 * 4 + 4, 8);
    c.exc_cap = (uint32_t)std::min<uint64_t>(c.n_words / 64 + 4, 0x7FFFFFFFu);   // 1 block in 32: beyond that the dense mask is smaller
    c.s_off = c.s_exc + (size_t)c.exc_cap * 8;
    c.in_sparse = c.s_off + ((size_t)c.nr + 1) * 8;
    c.s_inv = align_up(c.in_sparse, 16);
    c.dev_sparse = c.s_inv + c.n_words * 2 + 16;
}

struct ChunkStats {
    BatchStats st;
    bool uniform;        // every record has length rec_len0: rec_off is first + i * rec_len0
    uint64_t rec_len0;
};

// unit-length statistics of a chunk (what prep_stats_kernel computes on the device) and the equal-length test
static ChunkStats chunk_stats(const uint64_t *off0, uint32_t nu, uint32_t rpu) {
    ChunkStats c;
    memset(&c.st, 0, sizeof(c.st));
    const uint64_t nr = (uint64_t)nu * rpu;
    c.rec_len0 = nr ? off0[1] - off0[0] : 0;
    c.uniform = nr > 0 && offsets_equal_length(off0, nr, c.rec_len0);
    if (c.uniform) {   // the unit statistics follow from the one length
        const uint64_t len = c.rec_len0 * rpu;
        if (len > DCN_MAX_SHORT) { c.st.n_long = nu; c.st.long_bases = len * nu; }
        else c.st.max_short = (uint32_t)len;
        return c;
    }
    for (uint32_t u = 0; u < nu; u++) {
        const uint64_t len = off0[(uint64_t)(u + 1) * rpu] - off0[(uint64_t)u * rpu];
        if (len > DCN_MAX_SHORT) { c.st.n_long++; c.st.long_bases += len; }
        else if ((uint32_t)len > c.st.max_short) c.st.max_short = (uint32_t)len;
    }
    return c;
}

}  // namespace

// ---------------------------------------------------------------------------- two-route pipeline, arena form
// What the chunk pipeline below could not do: the GPU side of a 16 MB chunk is a kernel of two or three tiles per
// warp, which runs at a third of the big-launch rate (every warp of the SM is in the same phase, so the probe
// phase and the ALU phases do not overlap, and three waves are paid for 2.3), and a 1.5 Gbp batch was 120 such
// launches: 16 ms of SM time for 5.4 ms of work -- the limiter of the whole call (tools/e2e_timeline.py).  Here the
// copies keep their granularity (the link and the packers are fed as before) but land in ONE device arena that
// mirrors the batch, and kernels are launched over whatever contiguous range of units has arrived: 32 MB runs of the
// ASCII front, 64 MB or more of the packed back.  A launch waits for its copies through events, never on the host.
//   ASCII route   the calling thread copies atoms [head, ..) in order on one stream, two copies ahead of the link;
//   packed route  packer threads claim atoms from the back, pack [p0, p1) -- the 64-byte-aligned cuts just below
//                 the claim's first unit and below its upper neighbour's first unit, so the claims' words tile the
//                 arena exactly and only the batch's last claim is padded -- and copy codes, newline flags and the
//                 exception list (or the dense mask) to their places; the thread that completes a run of
//                 DCN_LAUNCH_ATOMS copied atoms below the last launch launches it.
// Atoms start at multiples of 32 units, so every claim owns whole words of the newline-flag array.
static int filter_pipeline_arena(dcn_ctx *ctx, const uint8_t *bases, const uint64_t *rec_off, uint32_t n_rec, int paired,
                                 uint32_t prefix_len, uint32_t abs_thr, double rel_thr, int deplete,
                                 uint8_t *keep, uint32_t *hits, uint32_t *total,
                                 int n_packers, bool ascii_route, double pack_share, uint64_t chunk_bases) {
    const uint32_t rpu = paired ? 2u : 1u;
    const uint32_t n_units = n_rec / rpu;
    const uint64_t A0 = rec_off[0] & ~63ull, B1 = rec_off[(uint64_t)n_units * rpu];
    const uint64_t nb_total = B1 - A0;
    static const int trace_level = []() { const char *e = getenv("DCN_HOST_TRACE"); return e ? atoi(e) : 0; }();
    static const int launch_atoms = []() { const char *e = getenv("DCN_LAUNCH_ATOMS"); return e ? std::max(1, atoi(e)) : 16; }();
    static const int launch_atoms_ascii = []() { const char *e = getenv("DCN_LAUNCH_ATOMS_ASCII"); return e ? std::max(1, atoi(e)) : 8; }();
    static const int ascii_atoms = []() { const char *e = getenv("DCN_ASCII_ATOMS"); return e ? std::max(1, atoi(e)) : 2; }();
    static const bool sparse_wire = []() { const char *e = getenv("DCN_SPARSE_MASK"); return !e || atoi(e) != 0; }();
    static const int PST = []() { const char *e = getenv("DCN_PACKER_STAGES"); return e ? std::max(1, std::min(8, atoi(e))) : 3; }();
    static const int ascii_ahead = []() { const char *e = getenv("DCN_ASCII_AHEAD"); return e ? std::max(1, std::min(4, atoi(e))) : 2; }();
    auto now_ms = [] { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    const double t_call0 = now_ms();

    // ---- atoms: ~4 MB of whole units, starting at multiples of 32 units
    const uint64_t atom_bases = std::max<uint64_t>(chunk_bases / 8, 1);
    std::vector<uint32_t> atom_u;
    atom_u.push_back(0);
    for (uint32_t u0 = 0; u0 < n_units;) {
        const uint64_t b0 = rec_off[(uint64_t)u0 * rpu];
        uint32_t lo = u0 + 1, hi = n_units;
        while (lo < hi) {
            const uint32_t mid = lo + (hi - lo + 1) / 2;
            if (rec_off[(uint64_t)mid * rpu] - b0 <= atom_bases) lo = mid; else hi = mid - 1;
        }
        if (lo < n_units) {
            lo = std::max(u0 + 32u, lo & ~31u);
            if (lo > n_units || n_units - lo < 32u) lo = n_units;
        }
        atom_u.push_back(lo);
        u0 = lo;
    }
    const int n_atoms = (int)atom_u.size() - 1;
    if (n_atoms < 8) return 1;   // a handful of huge units (atoms hold at least 32): not handled, the chunk pipeline cuts finer
    const int pack_budget = (int)std::min<double>(n_atoms, pack_share * n_atoms + 0.5);
    n_packers = std::min(n_packers, pack_budget);
    if (n_packers <= 0) return 1;   // not handled: the caller falls back to the chunk pipeline
    const int packer_grab_max = 4;

    // ---- the arena
    if (!ctx->arena) ctx->arena = new ArenaState();
    ArenaState &ar = *ctx->arena;
    const uint64_t n_words_total = 2 * ((nb_total + 64 + 31) / 32);
    if ((ascii_route && ar.ascii.ensure(nb_total + 256) != cudaSuccess) || ar.codes.ensure(n_words_total * 4 + 64) != cudaSuccess ||
        ar.inv.ensure(n_words_total * 2 + 64) != cudaSuccess || ar.nl.ensure(((size_t)n_rec / 32 + 4) * 4) != cudaSuccess ||
        ar.off.ensure(((size_t)n_rec + 1) * 8) != cudaSuccess || ar.out.ensure((size_t)n_units * 9 + 64) != cudaSuccess)
        return ctx->fail(DCN_ERR_NOMEM, "arena allocation failed", cudaGetLastError());
    if (!ar.copy_stream) {
        CK(cudaMallocHost(reinterpret_cast<void **>(&ar.h_after), sizeof(BatchStats) * ArenaState::NL));
        memset(ar.h_after, 0, sizeof(BatchStats) * ArenaState::NL);
        CK(cudaStreamCreateWithFlags(&ar.copy_stream, cudaStreamNonBlocking));
        for (int i = 0; i < 4; i++) CK(cudaEventCreateWithFlags(&ar.ev_sub[i], cudaEventDisableTiming));
        for (int i = 0; i < ArenaState::NL; i++) {
            Slot &s = ar.launch[i];
            CK(cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking));
            CK(cudaEventCreate(&s.ev_start)); CK(cudaEventCreate(&s.ev_kernel)); CK(cudaEventCreate(&s.ev_done));
            CK(cudaEventCreateWithFlags(&ar.ev_front[i], cudaEventDisableTiming));
        }
    }
    // launch slots: plan buffers for the largest range a launch takes (a launch that needs more grows its buffer: a
    // cudaFree in the middle of a call stalls everything, so it must not happen in the steady state)
    static const uint64_t range_cap = []() { const char *e = getenv("DCN_LAUNCH_CAP_MB"); return (e ? strtoull(e, nullptr, 10) : 160ull) << 20; }();
    {
        const uint64_t cap_rel = std::min<uint64_t>(nb_total, range_cap + (uint64_t)atom_bases) + 4096;
        const uint32_t cap_rec = (uint32_t)std::min<uint64_t>(n_rec, (uint64_t)((double)n_rec * 2.0 * (double)cap_rel / (double)std::max<uint64_t>(nb_total, 1)) + 64);
        const size_t pbytes = 128 + (size_t)wplan_tile_cap(cap_rel, cap_rec) * sizeof(WTile) + (size_t)wplan_ovf_cap(cap_rel) * 4;
        for (int i = 0; i < ArenaState::NL; i++)
            if (ar.launch[i].plan.ensure(pbytes) != cudaSuccess) return ctx->fail(DCN_ERR_NOMEM, "plan allocation failed", cudaGetLastError());
    }
    uint8_t *const d_ascii = ar.ascii.as<uint8_t>();
    uint32_t *const d_codes = ar.codes.as<uint32_t>();
    uint16_t *const d_inv = ar.inv.as<uint16_t>();
    uint32_t *const d_nl = ar.nl.as<uint32_t>();
    uint64_t *const d_off = ar.off.as<uint64_t>();
    uint32_t *const d_hits = ar.out.as<uint32_t>();
    uint32_t *const d_total = d_hits + n_units;
    uint8_t *const d_keep = reinterpret_cast<uint8_t *>(d_total + n_units);

    const bool out_pinned = [&] {
        const void *outs[3] = {keep, hits, total};
        for (const void *o : outs) {
            cudaPointerAttributes pa;
            const bool pinned = cudaPointerGetAttributes(&pa, o) == cudaSuccess && pa.type == cudaMemoryTypeHost;
            cudaGetLastError();
            if (!pinned) return false;
        }
        return true;
    }();

    ctx->t_h2d = ctx->t_kernel = ctx->t_d2h = ctx->t_pack = 0;
    ctx->bytes_h2d = ctx->bytes_d2h = 0;
    cudaEvent_t ev_call = nullptr;
    if (trace_level >= 2) {
        cudaEventCreate(&ev_call);
        cudaEventRecord(ev_call, ar.copy_stream);
    }

    // ---- shared state
    std::mutex m;
    int head = 0, tail = n_atoms, taken_by_packers = 0, in_flight = 0;
    int back_launched = n_atoms;          // packed atoms [back_launched, n_atoms) are launched
    int first_rc = DCN_OK;
    unsigned launch_seq = 0;
    std::vector<uint8_t> astate((size_t)n_atoms, 0);            // 0 unclaimed (or ASCII), 1 being packed, 2 copied
    std::vector<cudaEvent_t> aev((size_t)n_atoms, nullptr);     // the event after the atom's claim was copied
    std::vector<int> claim_end((size_t)n_atoms, 0);             // at a claim's first atom: one past its last
    std::vector<BatchStats> cstat((size_t)n_atoms);             // at a claim's first atom: unit statistics
    std::atomic<uint64_t> n_h2d{0}, n_d2h{0}, n_launch_p{0}, n_launch_a{0}, n_packed_claims{0}, n_ascii_copies{0}, n_uniform{0}, packed_bases{0};
    double kernel_ms_sum = 0, d2h_ms_sum = 0;   // under m
    double pack_busy_ms = 0, pack_wait_ms = 0;
    auto set_rc = [&](int rc) { std::lock_guard<std::mutex> g(m); if (rc && !first_rc) first_rc = rc; };

    // what a launch slot is running: enough to enqueue it again (a launch whose distinct-hit set overflowed is repeated)
    struct LInfo { int a_lo = 0, a_hi = 0; bool packed = false, has_long = false; BatchStats hs; };
    LInfo linfo[ArenaState::NL];

    // kernels + results of the launch in slot idx; `again`: the repeat after an overflow (larger set, waits for the long path itself)
    auto enqueue_launch = [&](int idx, bool again) -> int {
        Slot &s = ar.launch[idx];
        const LInfo &L = linfo[idx];
        const uint32_t u_lo = atom_u[(size_t)L.a_lo], u_hi = atom_u[(size_t)L.a_hi], nu = u_hi - u_lo, nr = nu * rpu;
        const uint64_t r0 = (uint64_t)u_lo * rpu;
        const uint64_t base0 = rec_off[r0] & ~63ull, n_abs = rec_off[(uint64_t)u_hi * rpu];
        FilterInput in;
        if (L.packed) {
            in.codes = d_codes + (base0 - A0) / 16;
            in.inv = d_inv + (base0 - A0) / 16;
            in.nl = d_nl + r0 / 32;
            in.nl_bit0 = (uint32_t)(r0 % 32);
        } else {
            in.bases = d_ascii + (base0 - A0);
        }
        CK(cudaEventRecord(s.ev_start, s.stream));
        if (L.has_long && !again) ar.h_after[idx].overflow = 0;
        const int rc = enqueue_filter(ctx, s.plan, s.longs, s.dedup, in, base0, n_abs, d_off + r0, nr, paired, prefix_len, abs_thr, rel_thr, deplete,
                                      d_keep + u_lo, d_hits + u_lo, d_total + u_lo, s.stream, &L.hs, false, false,
                                      L.has_long && !again ? &ar.h_after[idx] : nullptr, again ? 4u : 1u);
        if (rc) return rc;
        CK(cudaEventRecord(s.ev_kernel, s.stream));
        if (out_pinned) {
            CK(cudaMemcpyAsync(hits + u_lo, d_hits + u_lo, (size_t)nu * 4, cudaMemcpyDeviceToHost, s.stream));
            CK(cudaMemcpyAsync(total + u_lo, d_total + u_lo, (size_t)nu * 4, cudaMemcpyDeviceToHost, s.stream));
            CK(cudaMemcpyAsync(keep + u_lo, d_keep + u_lo, nu, cudaMemcpyDeviceToHost, s.stream));
        } else {
            uint8_t *o = s.h_out.as<uint8_t>();
            CK(cudaMemcpyAsync(o, d_hits + u_lo, (size_t)nu * 4, cudaMemcpyDeviceToHost, s.stream));
            CK(cudaMemcpyAsync(o + (size_t)nu * 4, d_total + u_lo, (size_t)nu * 4, cudaMemcpyDeviceToHost, s.stream));
            CK(cudaMemcpyAsync(o + (size_t)nu * 8, d_keep + u_lo, nu, cudaMemcpyDeviceToHost, s.stream));
        }
        n_d2h += (uint64_t)nu * 9;
        CK(cudaEventRecord(s.ev_done, s.stream));
        s.busy = true; s.u0 = u_lo; s.u1 = u_hi; s.packed = L.packed;
        return DCN_OK;
    };

    auto retire = [&](int idx) -> int {   // (the slot's mutex is held)
        Slot &s = ar.launch[idx];
        if (!s.busy) return DCN_OK;
        CK(cudaEventSynchronize(s.ev_done));
        if (linfo[idx].has_long && ar.h_after[idx].overflow) {
            // the distinct-hit set of this launch was too small: nothing was committed; once more, four times the size
            ar.h_after[idx].overflow = 0;
            const int rc = enqueue_launch(idx, true);
            if (rc) return rc;
            CK(cudaEventSynchronize(s.ev_done));
        }
        const uint32_t nu = s.u1 - s.u0;
        if (!out_pinned) {
            const uint8_t *o = s.h_out.as<uint8_t>();
            memcpy(hits + s.u0, o, (size_t)nu * 4);
            memcpy(total + s.u0, o + (size_t)nu * 4, (size_t)nu * 4);
            memcpy(keep + s.u0, o + (size_t)nu * 8, nu);
        }
        float a = 0, b = 0;
        cudaEventElapsedTime(&a, s.ev_start, s.ev_kernel);
        cudaEventElapsedTime(&b, s.ev_kernel, s.ev_done);
        { std::lock_guard<std::mutex> g(m); kernel_ms_sum += a; d2h_ms_sum += b; }
        if (trace_level >= 2 && ev_call) {   // device timeline of the launch, ms since the call's first enqueue
            float t0 = 0;
            cudaEventElapsedTime(&t0, ev_call, s.ev_start);
            fprintf(stderr, "[dcn launch done] units %u..%u %s: kernels %.3f .. %.3f, results back %.3f\n", s.u0, s.u1, s.packed ? "packed" : "ascii", t0, t0 + a, t0 + a + b);
        }
        s.busy = false;
        return DCN_OK;
    };

    // one launch over atoms [a_lo, a_hi) of one representation, after `waits`
    auto launch_range = [&](unsigned seq, int a_lo, int a_hi, bool packed, const BatchStats &hs, const std::vector<cudaEvent_t> &waits) -> int {
        const int idx = (int)(seq % ArenaState::NL);
        std::lock_guard<std::mutex> lg(ar.launch_m[idx]);
        Slot &s = ar.launch[idx];
        int rc = retire(idx);
        if (rc) return rc;
        const uint32_t u_lo = atom_u[(size_t)a_lo], u_hi = atom_u[(size_t)a_hi], nu = u_hi - u_lo;
        if (nu == 0) return DCN_OK;
        const uint64_t base0 = rec_off[(uint64_t)u_lo * rpu] & ~63ull, n_abs = rec_off[(uint64_t)u_hi * rpu];
        if (!out_pinned && s.h_out.ensure((size_t)nu * 9) != cudaSuccess) return ctx->fail(DCN_ERR_NOMEM, "staging allocation failed", cudaGetLastError());
        if (hs.n_long) {
            // long units: the distinct-hit set and the long-unit list of this slot at the size the largest range needs,
            // once -- launches of varying size would otherwise grow them (cudaFree: a device-wide stall) call after call
            const uint64_t cap_rel = std::max<uint64_t>(n_abs - base0, std::min<uint64_t>(nb_total, range_cap + (uint64_t)atom_bases));
            if (s.dedup.ensure((size_t)std::max<uint64_t>(4096, cap_rel / 4 + 64) * 16) != cudaSuccess ||
                s.longs.ensure((size_t)(cap_rel / DCN_MAX_SHORT + 64) * 4 + 64 + (size_t)(cap_rel / ChunkGeo<G31>::CSTRIDE + cap_rel / DCN_MAX_SHORT * rpu + 80) * sizeof(ChunkDesc)) != cudaSuccess)
                return ctx->fail(DCN_ERR_NOMEM, "long-path scratch allocation failed", cudaGetLastError());
        }
        for (cudaEvent_t e : waits) if (e) CK(cudaStreamWaitEvent(s.stream, e, 0));
        LInfo &L = linfo[idx];
        L.a_lo = a_lo; L.a_hi = a_hi; L.packed = packed; L.has_long = hs.n_long != 0; L.hs = hs;
        rc = enqueue_launch(idx, false);
        if (rc) return rc;
        (packed ? n_launch_p : n_launch_a)++;
        if (trace_level >= 2)
            fprintf(stderr, "[dcn launch] %.2f ms: %s atoms %d..%d units %u..%u (%.1f MB)\n", now_ms() - t_call0, packed ? "packed" : "ascii", a_lo, a_hi,
                    u_lo, u_hi, (n_abs - base0) / 1e6);
        return DCN_OK;
    };

    // no launch of this call is still running (racy reads of `busy` are harmless: the answer only tunes the batching)
    auto gpu_idle = [&]() -> bool {
        for (int i = 0; i < ArenaState::NL; i++)
            if (ar.launch[i].busy && cudaEventQuery(ar.launch[i].ev_done) == cudaErrorNotReady) return false;
        cudaGetLastError();
        return true;
    };
    static const int launch_atoms_idle = []() { const char *e = getenv("DCN_LAUNCH_ATOMS_IDLE"); return e ? std::max(1, atoi(e)) : 4; }();

    // (m held) the copied run of packed atoms just below the last launch; taken when it is long enough (DCN_LAUNCH_ATOMS; a
    // few atoms are enough while the GPU has nothing to do: batching only pays when launches queue up), or when it is all
    // there will be
    struct Range { int a_lo = 0, a_hi = 0; unsigned seq = 0; BatchStats hs; std::vector<cudaEvent_t> waits; };
    auto take_packed_range = [&](bool final_call, Range &r) -> bool {
        int a = back_launched;
        while (a > 0 && astate[(size_t)a - 1] == 2) a--;
        const int pending = back_launched - a;
        const bool all_in = final_call || (head >= tail && in_flight == 0) || (taken_by_packers >= pack_budget && in_flight == 0);
        if (pending <= 0) return false;
        if (pending < launch_atoms && !(all_in && (a == 0 || astate[(size_t)a - 1] == 0)) && !(pending >= launch_atoms_idle && gpu_idle())) return false;
        // at most range_cap bases per launch (whole claims, from the top): the launch slots' plan buffers are sized for that
        const uint64_t top = rec_off[(uint64_t)atom_u[(size_t)back_launched] * rpu];
        int lowest = -1;
        for (int i = a; i < back_launched; i = claim_end[(size_t)i])
            if (lowest < 0 && top - rec_off[(uint64_t)atom_u[(size_t)i] * rpu] <= range_cap) lowest = i;
        if (lowest < 0) {   // the top claim alone is larger (a record longer than the cap): it goes by itself
            for (int i = a; i < back_launched; i = claim_end[(size_t)i]) lowest = i;
        }
        a = lowest;
        r.a_lo = a; r.a_hi = back_launched; r.seq = launch_seq++;
        memset(&r.hs, 0, sizeof(r.hs));
        r.waits.clear();
        for (int i = a; i < back_launched; i = claim_end[(size_t)i]) {
            const BatchStats &c = cstat[(size_t)i];
            r.hs.n_long += c.n_long; r.hs.long_bases += c.long_bases; r.hs.max_short = std::max(r.hs.max_short, c.max_short);
            if (r.waits.empty() || r.waits.back() != aev[(size_t)i]) r.waits.push_back(aev[(size_t)i]);
        }
        if (back_launched < n_atoms) r.waits.push_back(aev[(size_t)back_launched]);   // the upper neighbour packed the last bases of this run's last unit
        back_launched = a;
        return true;
    };

    if ((int)ctx->aslot.size() < 8 * n_packers) ctx->aslot.resize((size_t)8 * n_packers);   // (8: room for any DCN_PACKER_STAGES)

    auto packer = [&](int t) {
        double pack_ms = 0, wait_ms = 0, ship_ms = 0, t_begin = now_ms() - t_call0, t_first = 0, t_last = 0;
        uint64_t my_bases = 0;
        std::vector<uint64_t> bad32;
        std::vector<uint32_t> bad_mask;
        auto body = [&]() -> int {
            CK(cudaSetDevice(ctx->device));
            for (int i = 0; i < PST; i++) {
                Slot &s = ctx->aslot[(size_t)(8 * t + i)];
                if (!s.stream) CK(cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking));
                if (!s.ev_h2d) CK(cudaEventCreateWithFlags(&s.ev_h2d, cudaEventDisableTiming));
            }
            bool first_claim = true;
            for (int flip = 0;; flip = (flip + 1) % PST) {
                int a_lo, a_hi;
                {
                    std::lock_guard<std::mutex> g(m);
                    if (first_rc || tail <= head || taken_by_packers >= pack_budget) break;
                    int grab = std::max(1, std::min<int>(packer_grab_max, (tail - head) / n_packers));
                    // the threads pack at one pace: first claims of 1, 2, 3, 4 atoms take them out of step, so their copies
                    // reach the link spread out instead of twelve at a time
                    if (first_claim) { grab = std::min(grab, 1 + t % packer_grab_max); first_claim = false; }
                    grab = std::min(grab, std::min(tail - head, pack_budget - taken_by_packers));
                    a_hi = tail; a_lo = tail -= grab; taken_by_packers += grab; in_flight++;
                    for (int i = a_lo; i < a_hi; i++) astate[(size_t)i] = 1;
                }
                const uint32_t u0 = atom_u[(size_t)a_lo], u1 = atom_u[(size_t)a_hi], nu = u1 - u0, nr = nu * rpu;
                const uint64_t r0 = (uint64_t)u0 * rpu;
                const uint64_t *off0 = rec_off + r0;
                const uint64_t p0 = off0[0] & ~63ull;
                const uint64_t p1 = a_hi == n_atoms ? B1 : (rec_off[(uint64_t)u1 * rpu] & ~63ull);   // (multiple of 32 from p0 unless it is the batch's end)
                const uint64_t nb = p1 - p0;
                const size_t n_words = 2 * (size_t)((nb + 31) / 32);
                const size_t nl_words = ((size_t)nr + 31) / 32;
                const uint32_t exc_cap = (uint32_t)std::min<uint64_t>(n_words / 64 + 4, 0x7FFFFFFFu);
                // blob: codes | newline flags | exception list or dense mask | offsets
                const size_t o_nl = align_up(n_words * 4, 8), o_x = o_nl + align_up(nl_words * 4 + 4, 8);
                const size_t o_off = o_x + std::max<size_t>((size_t)exc_cap * 8, align_up(n_words * 2, 8));
                const size_t blob = o_off + ((size_t)nr + 1) * 8;
                Slot &s = ctx->aslot[(size_t)(8 * t + flip)];
                if (s.busy) {   // the blob's previous copies have left the host
                    const double w0 = now_ms();
                    CK(cudaEventSynchronize(s.ev_h2d));
                    wait_ms += now_ms() - w0;
                    s.busy = false;
                }
                const double t0 = now_ms();
                if (s.h_in.ensure(std::max(blob, (size_t)(packer_grab_max * atom_bases) / 2)) != cudaSuccess ||
                    s.in.ensure(std::max(blob, (size_t)(packer_grab_max * atom_bases) / 2)) != cudaSuccess)
                    return ctx->fail(DCN_ERR_NOMEM, "pinned staging allocation failed", cudaGetLastError());
                uint8_t *hin = s.h_in.as<uint8_t>();
                const ChunkStats cs = chunk_stats(off0, nu, rpu);
                uint32_t *h_codes = reinterpret_cast<uint32_t *>(hin);
                uint32_t *h_nl = reinterpret_cast<uint32_t *>(hin + o_nl);
                uint16_t *h_inv = reinterpret_cast<uint16_t *>(hin + o_x);
                int64_t n_exc = -1;
                if (sparse_wire) pack_records(bases, p0, nb, off0, nr, ctx->k, prefix_len, h_codes, nullptr, h_nl, bad32, &bad_mask);
                else pack_records(bases, p0, nb, off0, nr, ctx->k, prefix_len, h_codes, h_inv, h_nl, bad32);
                // records that end beyond p1 (the last < 64 bases of the claim's last unit are packed by the upper neighbour):
                // their newline flags are looked at directly (src/filter_common.rs:217-229)
                for (uint32_t q = nr; q-- > 0;) {
                    const uint64_t len = off0[q + 1] - off0[q];
                    const uint64_t e = off0[q] + ((prefix_len > 0 && len > prefix_len) ? prefix_len : len);
                    if (off0[q + 1] <= p1) break;          // everything from here down lies inside [p0, p1)
                    if (e <= p1 || len < (uint64_t)ctx->k) continue;
                    if (bases[e - 1] == (uint8_t)'\n') h_nl[q / 32] |= 1u << (q % 32);
                }
                if (sparse_wire) {
                    if (bad32.size() <= exc_cap) {
                        n_exc = (int64_t)bad32.size();
                        uint32_t *ex = reinterpret_cast<uint32_t *>(hin + o_x);
                        for (size_t i = 0; i < bad32.size(); i++) { ex[2 * i] = (uint32_t)bad32[i]; ex[2 * i + 1] = bad_mask[i]; }
                    } else {   // N-rich claim: the dense mask is smaller
                        memset(h_inv, 0, n_words * 2);
                        for (size_t i = 0; i < bad32.size(); i++) memcpy(h_inv + 2 * bad32[i], &bad_mask[i], 4);
                    }
                }
                // offsets (when they are shipped) right behind what the claim really uses of the blob: one copy carries it all
                const size_t o_off_w = n_exc >= 0 ? align_up(o_x + (size_t)n_exc * 8, 8) : align_up(o_x + n_words * 2, 8);
                if (!cs.uniform) memcpy(hin + o_off_w, off0, ((size_t)nr + 1) * 8);
                const size_t wire = o_off_w + (cs.uniform ? 0 : ((size_t)nr + 1) * 8);
                const double t1 = now_ms();
                pack_ms += t1 - t0;
                my_bases += nb;
                // ---- one copy into the stage's device buffer, one kernel to the claim's places in the arena
                uint64_t moved = 0;
                if (wire) {
                    const uint8_t *dst = s.in.as<uint8_t>();
                    CK(cudaMemcpyAsync(s.in.p, hin, wire, cudaMemcpyHostToDevice, s.stream));
                    moved += wire;
                    uint16_t *inv_dst = d_inv + (p0 - A0) / 16;
                    if (n_exc >= 0 && n_words) CK(cudaMemsetAsync(inv_dst, 0, n_words * 2, s.stream));
                    ClaimUnpack cu;
                    cu.codes = reinterpret_cast<const uint32_t *>(dst); cu.dst_codes = d_codes + (p0 - A0) / 16; cu.n_words = n_words;
                    cu.nl = reinterpret_cast<const uint32_t *>(dst + o_nl); cu.dst_nl = d_nl + r0 / 32; cu.nl_words = (uint32_t)nl_words;
                    cu.exc = reinterpret_cast<const uint2 *>(dst + o_x); cu.n_exc = n_exc;
                    cu.inv = reinterpret_cast<const uint32_t *>(dst + o_x); cu.dst_inv32 = reinterpret_cast<uint32_t *>(inv_dst);
                    cu.off = reinterpret_cast<const uint64_t *>(dst + o_off_w); cu.dst_off = d_off + r0;
                    cu.n_off = cs.uniform ? 0u : nr + 1; cu.n_gen = cs.uniform ? nr + 1 : 0u;
                    cu.off_first = off0[0]; cu.off_len = cs.rec_len0;
                    claim_unpack_kernel<<<grid_for(ctx, std::max<uint64_t>(n_words / 4, (uint64_t)nr + 1), 256), 256, 0, s.stream>>>(cu);
                    ctx->launches += 1;
                    if (cs.uniform) n_uniform++;
                }
                CK(cudaEventRecord(s.ev_h2d, s.stream));
                CK(cudaGetLastError());
                s.busy = true;
                n_h2d += moved;
                n_packed_claims++;
                Range r;
                bool go;
                {
                    std::lock_guard<std::mutex> g(m);
                    for (int i = a_lo; i < a_hi; i++) { astate[(size_t)i] = 2; aev[(size_t)i] = s.ev_h2d; }
                    claim_end[(size_t)a_lo] = a_hi; cstat[(size_t)a_lo] = cs.st;
                    in_flight--;
                    go = take_packed_range(false, r);
                }
                if (go) { const int rc = launch_range(r.seq, r.a_lo, r.a_hi, true, r.hs, r.waits); if (rc) return rc; }
                t_last = now_ms() - t_call0;
                ship_ms += t_last - (t1 - t_call0);
                if (t_first == 0) t_first = t_last;
            }
            return DCN_OK;
        };
        const int rc = body();
        if (rc) set_rc(rc);
        packed_bases += my_bases;
        std::lock_guard<std::mutex> g(m);
        pack_busy_ms += pack_ms; pack_wait_ms += wait_ms;
        if (trace_level >= 1)
            fprintf(stderr, "[dcn packer %d] begin %.2f first copy %.2f last %.2f end %.2f ms; packing %.2f enqueueing %.2f waiting %.2f ms; %.1f MB\n",
                    t, t_begin, t_first, t_last, now_ms() - t_call0, pack_ms, ship_ms, wait_ms, my_bases / 1e6);
    };
    std::vector<std::thread> packers;
    for (int t = 0; t < n_packers; t++) {
        try { packers.emplace_back(packer, t); } catch (const std::exception &) { break; }
    }
    if (packers.empty()) ascii_route = true;

    // ---- the calling thread: the ASCII front
    int rc = DCN_OK;
    double main_wait_ms = 0;
    {
        int k_sub = 0, launched_front = 0, copied_front = 0;
        BatchStats pend;
        memset(&pend, 0, sizeof(pend));
        auto launch_front = [&]() -> int {
            unsigned seq;
            { std::lock_guard<std::mutex> g(m); seq = launch_seq++; }
            const int idx = (int)(seq % ArenaState::NL);
            CK(cudaEventRecord(ar.ev_front[idx], ar.copy_stream));
            std::vector<cudaEvent_t> waits{ar.ev_front[idx]};
            const int r = launch_range(seq, launched_front, copied_front, false, pend, waits);
            launched_front = copied_front;
            memset(&pend, 0, sizeof(pend));
            return r;
        };
        while (ascii_route && rc == DCN_OK) {
            if (k_sub >= ascii_ahead) {   // no more than a few copies ahead of the link: the packed claims' copies queue on the same engine
                const double w0 = now_ms();
                const cudaError_t e = cudaEventSynchronize(ar.ev_sub[(k_sub - ascii_ahead) % 4]);
                if (e != cudaSuccess) { rc = ctx->fail(DCN_ERR_CUDA, "cudaEventSynchronize(ev_sub)", e); break; }
                main_wait_ms += now_ms() - w0;
            }
            int a_lo, a_hi;
            {
                std::lock_guard<std::mutex> g(m);
                if (first_rc || head >= tail) break;
                const bool packers_done = taken_by_packers >= pack_budget;
                const int grab = packers_done ? std::min(8, tail - head) : std::max(1, std::min<int>(ascii_atoms, (tail - head) / 6));
                a_lo = head;
                a_hi = head = std::min<int>(tail, head + grab);
            }
            const uint32_t u0 = atom_u[(size_t)a_lo], u1 = atom_u[(size_t)a_hi], nu = u1 - u0, nr = nu * rpu;
            const uint64_t r0 = (uint64_t)u0 * rpu;
            const uint64_t x0 = a_lo == 0 ? A0 : rec_off[r0], x1 = rec_off[(uint64_t)u1 * rpu];
            const ChunkStats cs = chunk_stats(rec_off + r0, nu, rpu);
            auto enq = [&]() -> int {
                if (x1 > x0) CK(cudaMemcpyAsync(d_ascii + (x0 - A0), bases + x0, (size_t)(x1 - x0), cudaMemcpyHostToDevice, ar.copy_stream));
                n_h2d += x1 - x0;
                if (cs.uniform) {
                    uniform_offsets_kernel<<<grid_for(ctx, (uint64_t)nr + 1, 256), 256, 0, ar.copy_stream>>>(d_off + r0, nr + 1, rec_off[r0], cs.rec_len0);
                    ctx->launches += 1;
                    n_uniform++;
                } else {
                    CK(cudaMemcpyAsync(d_off + r0, rec_off + r0, ((size_t)nr + 1) * 8, cudaMemcpyHostToDevice, ar.copy_stream));
                    n_h2d += ((uint64_t)nr + 1) * 8;
                }
                CK(cudaEventRecord(ar.ev_sub[k_sub % 4], ar.copy_stream));
                return DCN_OK;
            };
            if ((rc = enq())) break;
            k_sub++;
            n_ascii_copies++;
            pend.n_long += cs.st.n_long; pend.long_bases += cs.st.long_bases; pend.max_short = std::max(pend.max_short, cs.st.max_short);
            copied_front = a_hi;
            if (copied_front - launched_front >= launch_atoms_ascii || (copied_front - launched_front >= ascii_atoms && gpu_idle())) rc = launch_front();
        }
        if (rc == DCN_OK && copied_front > launched_front) rc = launch_front();
    }
    if (rc) set_rc(rc);
    for (auto &t : packers) t.join();
    {   // whatever the packers copied and nobody launched (their budget ran out before the fronts met, or an error stopped them)
        Range r;
        bool go;
        { std::lock_guard<std::mutex> g(m); go = !first_rc && take_packed_range(true, r); }
        while (go) {
            const int r2 = launch_range(r.seq, r.a_lo, r.a_hi, true, r.hs, r.waits);
            if (r2) { set_rc(r2); break; }
            std::lock_guard<std::mutex> g(m);
            go = take_packed_range(true, r);
        }
    }
    for (int i = 0; i < ArenaState::NL; i++) {
        std::lock_guard<std::mutex> lg(ar.launch_m[i]);
        const int r2 = retire(i);
        if (r2) set_rc(r2);
    }
    rc = first_rc;
    if (rc) {
        cudaDeviceSynchronize();
        for (int i = 0; i < ArenaState::NL; i++) ar.launch[i].busy = false;
    }
    for (auto &sl : ctx->aslot) sl.busy = false;   // every copy was waited for by a launch that has been retired
    ctx->t_kernel = (float)kernel_ms_sum; ctx->t_d2h = (float)d2h_ms_sum;
    ctx->bytes_h2d = n_h2d; ctx->bytes_d2h = n_d2h;
    ctx->n_packed_chunks += n_packed_claims; ctx->n_ascii_chunks += n_ascii_copies; ctx->n_uniform_chunks += n_uniform;
    if (ev_call) cudaEventDestroy(ev_call);
    const double call_ms = now_ms() - t_call0;
    if (trace_level >= 1)
        fprintf(stderr, "[dcn host] arena: %d atoms, %d packers: %llu packed claims / %llu ascii copies, %llu + %llu launches; call %.2f ms; "
                "front waited %.2f ms for the link; packers: packing %.2f ms, waiting for a blob %.2f ms (sums over threads); h2d %.1f MB\n",
                n_atoms, n_packers, (unsigned long long)n_packed_claims.load(), (unsigned long long)n_ascii_copies.load(),
                (unsigned long long)n_launch_p.load(), (unsigned long long)n_launch_a.load(), call_ms, main_wait_ms, pack_busy_ms, pack_wait_ms,
                n_h2d.load() / 1e6);
    if (packed_bases) {
        ctx->t_pack = (float)call_ms;
        if (pack_busy_ms > 0) ctx->pack_gbps = (double)packed_bases / 1e6 / pack_busy_ms * std::max(1, n_packers);
    }
    return rc;
}

static int filter_pipeline(dcn_ctx *ctx, const HostSrc &src, const uint64_t *rec_off, uint32_t n_rec, int paired,
                           uint32_t prefix_len, uint32_t abs_thr, double rel_thr, int deplete,
                           uint8_t *keep, uint32_t *hits, uint32_t *total) {
    const uint8_t *bases = src.bases;
    const bool prepacked = src.codes != nullptr;
    if (!rec_off || (!bases && !prepacked && n_rec && rec_off[n_rec] > rec_off[0])) return ctx->fail(DCN_ERR_ARG, "null input pointer");
    if (!keep || !hits || !total) return ctx->fail(DCN_ERR_ARG, "null output pointer");
    if (prepacked && (ctx->k != 31 || ctx->w != 15)) return ctx->fail(DCN_ERR_UNSUPPORTED, "packed input is only implemented for k=31, w=15");
    if (!ctx->table.p) return ctx->fail(DCN_ERR_NO_INDEX, "no index resident: call dcn_index_upload first");
    CK(cudaSetDevice(ctx->device));
    const uint32_t rpu = paired ? 2u : 1u;
    if (paired && (n_rec & 1u)) return ctx->fail(DCN_ERR_ARG, "paired batch needs an even record count");
    const uint32_t n_units = n_rec / rpu;
    if (n_units == 0) return DCN_OK;

    static const uint64_t chunk_bases_env = []() {
        const char *e = getenv("DCN_CHUNK_MB");
        uint64_t mb = e ? strtoull(e, nullptr, 10) : 32;
        if (mb < 1) mb = 1;
        return mb << 20;
    }();
    // Caller-packed input runs the chunk form (one route): its chunks are larger, because a kernel over 32 Mbp runs at half
    // the rate of one over 128 Mbp (tools/chunk_cost.py) and at 0.25 - 0.375 B/bp the copy of a chunk is short anyway.
    // Measured (bench.py, Gbp/s, dense mask / sparse list): 32 MB 119 / 137, 64 MB 131 / 144, 128 MB 128 / 172, 256 MB 122 / 161.
    static const uint64_t packed_chunk_env = []() { const char *e = getenv("DCN_PACKED_CHUNK_MB"); return e ? std::max<uint64_t>(1, strtoull(e, nullptr, 10)) << 20 : 0ull; }();
    const uint64_t chunk_bases = !prepacked || getenv("DCN_CHUNK_MB") ? chunk_bases_env
                               : packed_chunk_env ? packed_chunk_env : (src.exc ? 128ull << 20 : 64ull << 20);

    // ---- routes
    int n_packers = 0;           // host threads packing atoms
    bool ascii_route = true;     // the enqueueing thread may ship chunks as ASCII
    double pack_share = 0;       // share of the atoms the packers may take
    const uint64_t total_bases = rec_off[(uint64_t)n_units * rpu] - rec_off[0];
    if (!prepacked && ctx->k == 31 && ctx->w == 15 && pack_threads_of(ctx) > 0) {
        cudaPointerAttributes pa;
        const bool pinned = cudaPointerGetAttributes(&pa, bases) == cudaSuccess && pa.type == cudaMemoryTypeHost;
        cudaGetLastError();
        static const double env_frac = []() { const char *e = getenv("DCN_PACK_FRACTION"); return e ? atof(e) : -1.0; }();
        const double forced = ctx->pack_fraction >= 0 ? ctx->pack_fraction : env_frac;
        if (forced >= 0) { pack_share = std::min(1.0, forced); ascii_route = pack_share < 1.0; }
        else if (!pinned) { pack_share = 1.0; ascii_route = false; }
        else pack_share = total_bases >= 6 * chunk_bases ? 1.0 : 0.0;   // dynamic split; not worth the threads for a few chunks
        if (pack_share > 0) n_packers = pack_threads_of(ctx);
    }

    // ---- the two routes together: the arena form (kernels over whatever has arrived); the chunk form below serves the
    // single-route cases (no packers, caller-packed input) and batches beyond the arena's size limit
    static const bool arena_on = []() { const char *e = getenv("DCN_PIPELINE"); return !e || strcmp(e, "chunks") != 0; }();
    static const uint64_t arena_max = []() { const char *e = getenv("DCN_ARENA_MAX_MB"); return (e ? strtoull(e, nullptr, 10) : 8192ull) << 20; }();
    if (arena_on && n_packers > 0 && total_bases <= arena_max) {
        const int rc = filter_pipeline_arena(ctx, bases, rec_off, n_rec, paired, prefix_len, abs_thr, rel_thr, deplete, keep, hits, total,
                                             n_packers, ascii_route, pack_share, chunk_bases);
        if (rc <= 0) return rc;   // (1: not handled)
    }

    // ---- plan: the batch is cut into unit-aligned ATOMS of 4 MB.  The ASCII route ships up to `atoms_per_chunk`
    // consecutive atoms as one chunk (one copy, one kernel); a packer thread takes up to half a chunk at a time while
    // plenty of atoms are left and single atoms near the end, so the two routes meet with at most one atom's packing
    // time (~0.7 ms) of imbalance instead of a chunk's.
    const uint32_t atoms_per_chunk = n_packers > 0 ? 8 : 1;
    const uint64_t atom_bases = std::max<uint64_t>(chunk_bases / atoms_per_chunk, 1);
    std::vector<uint32_t> atom_u;   // atom i covers units [atom_u[i], atom_u[i + 1])
    atom_u.push_back(0);
    for (uint32_t u0 = 0; u0 < n_units;) {
        // largest u1 with rec_off[u1*rpu] - rec_off[u0*rpu] <= atom_bases (at least one unit)
        const uint64_t b0 = rec_off[(uint64_t)u0 * rpu];
        uint32_t lo = u0 + 1, hi = n_units;
        while (lo < hi) {
            uint32_t mid = lo + (hi - lo + 1) / 2;
            if (rec_off[(uint64_t)mid * rpu] - b0 <= atom_bases) lo = mid; else hi = mid - 1;
        }
        atom_u.push_back(lo);
        u0 = lo;
    }
    const int n_atoms = (int)atom_u.size() - 1;
    auto make_chunk = [&](int a_first, int a_last) {   // atoms [a_first, a_last)
        ChunkPlan c;
        c.u0 = atom_u[(size_t)a_first]; c.u1 = atom_u[(size_t)a_last]; c.nu = c.u1 - c.u0; c.nr = c.nu * rpu;
        c.b1 = rec_off[(uint64_t)c.u1 * rpu];
        c.a0 = rec_off[(uint64_t)c.u0 * rpu] & ~63ull;   // chunk origin: keeps 16-byte loads and 32-base pack blocks aligned
        c.nb = c.b1 - c.a0;
        c.n_words = 2 * ((c.nb + 31) / 32);
        // packed blob: codes | non-ACGT bits | newline flags | offsets (last: an equal-length chunk does not ship them)
        c.o_inv = c.n_words * 4; c.o_nl = align_up(c.o_inv + c.n_words * 2, 8);
        c.o_off_p = c.o_nl + align_up(((size_t)c.nr + 31) / 32 * 4 + 4, 8);   // + one word: the caller-packed form's flags start mid-word
        c.in_packed = c.o_off_p + ((size_t)c.nr + 1) * 8;
        c.o_off_a = align_up(c.nb + 16, 16); c.in_ascii = c.o_off_a + ((size_t)c.nr + 1) * 8;
        c.out_bytes = (size_t)c.nu * 9;
        sparse_layout(c);
        return c;
    };
    // the largest chunks of the two routes, for sizing the stages once (records of >= 64 bases on average assumed;
    // a chunk that needs more grows its stage)
    auto layout_for = [&](uint64_t nb_cap) {
        ChunkPlan c;
        memset(&c, 0, sizeof(c));
        c.nb = std::min<uint64_t>(nb_cap, total_bases + 64);
        c.nr = (uint32_t)std::min<uint64_t>(c.nb / 64 + 2, n_rec); c.nu = c.nr / rpu + 1;
        c.n_words = 2 * ((c.nb + 31) / 32);
        // packed blob: codes | non-ACGT bits | newline flags | offsets (last: an equal-length chunk does not ship them)
        c.o_inv = c.n_words * 4; c.o_nl = align_up(c.o_inv + c.n_words * 2, 8);
        c.o_off_p = c.o_nl + align_up(((size_t)c.nr + 31) / 32 * 4 + 4, 8);   // + one word: the caller-packed form's flags start mid-word
        c.in_packed = c.o_off_p + ((size_t)c.nr + 1) * 8;
        c.o_off_a = align_up(c.nb + 16, 16); c.in_ascii = c.o_off_a + ((size_t)c.nr + 1) * 8;
        c.out_bytes = (size_t)c.nu * 9;
        sparse_layout(c);
        return c;
    };
    static const int env_grab = []() { const char *e = getenv("DCN_PACKER_GRAB"); return e ? std::max(1, atoi(e)) : 0; }();
    const int packer_grab_max = env_grab ? env_grab : std::max<int>(1, (int)atoms_per_chunk / 2);
    const ChunkPlan big_ascii = layout_for(chunk_bases + 4096), big_packed = layout_for((uint64_t)packer_grab_max * atom_bases + 4096);
    const int pack_budget = (int)std::min<double>(n_atoms, pack_share * n_atoms + 0.5);   // atoms the packers may take
    n_packers = std::min(n_packers, pack_budget);
    if (n_packers == 0) ascii_route = true;

    ctx->t_h2d = ctx->t_kernel = ctx->t_d2h = ctx->t_pack = 0;
    ctx->bytes_h2d = ctx->bytes_d2h = 0;
    static const int trace_level = []() { const char *e = getenv("DCN_HOST_TRACE"); return e ? atoi(e) : 0; }();
    cudaEvent_t ev_call = nullptr;
    if (trace_level >= 2) {
        cudaEventCreate(&ev_call);
        cudaEventRecord(ev_call, ctx->slot[0].stream);
    }
    auto now_ms = [] { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    const double t_call0 = now_ms();
    // pinned output arrays receive the device's results directly
    const bool out_pinned = [&] {
        const void *outs[3] = {keep, hits, total};
        for (const void *o : outs) {
            cudaPointerAttributes pa;
            const bool pinned = cudaPointerGetAttributes(&pa, o) == cudaSuccess && pa.type == cudaMemoryTypeHost;
            cudaGetLastError();
            if (!pinned) return false;
        }
        return true;
    }();

    // what one thread of the pipeline adds up; merged into the ctx under `m` when the thread is done
    struct Acc {
        uint64_t h2d = 0, d2h = 0, n_packed = 0, n_ascii = 0, n_uniform = 0, packed_bases = 0;
        float t_h2d = 0, t_kernel = 0, t_d2h = 0;
        double wait_ms = 0, pack_ms = 0;   // waiting for a stage to come back; packing
        double ship_ms = 0, t_begin = 0, t_first_ship = 0, t_last_ship = 0, t_end = 0;   // host-side trace (ms since the call began)
    };
    auto retire = [&](Slot &s, Acc &acc) -> int {  // wait for a stage and scatter its results (disjoint unit ranges per stage)
        if (!s.busy) return DCN_OK;
        const double w0 = now_ms();
        CK(cudaEventSynchronize(s.ev_done));
        acc.wait_ms += now_ms() - w0;
        const uint32_t nu = s.u1 - s.u0;
        if (!out_pinned) {   // results came back through the stage's pinned blob
            const uint8_t *o = s.h_out.as<uint8_t>();
            memcpy(hits + s.u0, o, (size_t)nu * 4);
            memcpy(total + s.u0, o + (size_t)nu * 4, (size_t)nu * 4);
            memcpy(keep + s.u0, o + (size_t)nu * 8, nu);
        }
        float a = 0, b = 0, c = 0;
        cudaEventElapsedTime(&a, s.ev_start, s.ev_h2d);
        cudaEventElapsedTime(&b, s.ev_h2d, s.ev_kernel);
        cudaEventElapsedTime(&c, s.ev_kernel, s.ev_done);
        acc.t_h2d += a; acc.t_kernel += b; acc.t_d2h += c;
        if (trace_level >= 2 && ev_call) {   // device timeline of the chunk, ms since the call's first enqueue
            float t0 = 0;
            cudaEventElapsedTime(&t0, ev_call, s.ev_start);
            fprintf(stderr, "[dcn chunk] units %u..%u %s start %.3f h2d_end %.3f kernel_end %.3f done %.3f\n", s.u0, s.u1,
                    s.packed ? "packed" : "ascii", t0, t0 + a, t0 + a + b, t0 + a + b + c);
        }
        s.busy = false;
        return DCN_OK;
    };

    // Enqueue one chunk on stage `s`: copies in, the kernels, results out.  `route`: 0 ASCII bytes, 1 the caller's packed
    // arrays, 2 the blob a packer thread has just written to s.h_in.  Called by the enqueueing thread (stages ctx->slot)
    // and by every packer thread (its own two stages): it touches nothing shared but atomics and the device.
    // n_exc >= 0 (route 2 only): the blob is in the sparse form with that many exceptions; < 0: dense
    auto ship = [&](Slot &s, const ChunkPlan &c, int route, const ChunkStats &cs, Acc &acc, bool time_fused, int64_t n_exc = -1) -> int {
        const uint64_t *off0 = rec_off + (uint64_t)c.u0 * rpu;
        // sized for the largest chunk the stage is likely to see, not for this one: growing a buffer later costs a
        // cudaFree / cudaFreeHost (a device-wide sync) in the middle of some call
        const ChunkPlan &big = route == 2 ? big_packed : big_ascii;
        const size_t exc_room = route == 1 && src.exc ? 16 + (c.n_words / 2 + 1) * 8 : 0;   // caller's sparse list: at most one entry per 32-base block
        if (s.in.ensure(std::max(route ? std::max(c.in_packed, c.dev_sparse) + 8 : c.in_ascii, route ? std::max(big.in_packed, big.dev_sparse) + 8 : big.in_ascii) + exc_room) != cudaSuccess ||
            s.out.ensure(std::max(c.out_bytes, big.out_bytes)) != cudaSuccess || s.h_out.ensure(std::max(c.out_bytes, big.out_bytes)) != cudaSuccess)
            return ctx->fail(DCN_ERR_NOMEM, "staging allocation failed", cudaGetLastError());
        uint8_t *din = s.in.as<uint8_t>();
        FilterInput in;
        const uint64_t *d_off;
        // offsets of the chunk on the device: copied, or generated when the records all have one length (then rec_off
        // is an arithmetic sequence: 8 bytes per record stay off PCIe, 5 % of an ASCII chunk of 150-base reads)
        auto ship_offsets = [&](uint8_t *dst, const void *host_src) -> cudaError_t {
            if (cs.uniform) {
                uniform_offsets_kernel<<<grid_for(ctx, (uint64_t)c.nr + 1, 256), 256, 0, s.stream>>>(reinterpret_cast<uint64_t *>(dst), c.nr + 1, off0[0], cs.rec_len0);
                ctx->launches += 1;
                acc.n_uniform++;
                return cudaGetLastError();
            }
            acc.h2d += ((uint64_t)c.nr + 1) * 8;
            return cudaMemcpyAsync(dst, host_src, ((size_t)c.nr + 1) * 8, cudaMemcpyHostToDevice, s.stream);
        };
        CK(cudaEventRecord(s.ev_start, s.stream));
        if (route == 1) {   // slices of the caller's packed arrays, copied as they are
            const uint64_t r_first = (uint64_t)c.u0 * rpu;
            CK(cudaMemcpyAsync(din, src.codes + c.a0 / 16, c.n_words * 4, cudaMemcpyHostToDevice, s.stream));
            acc.h2d += c.n_words * 4;
            if (src.exc) {
                // sparse form: the listed blocks of this chunk (the list is ascending) are copied behind the blob and
                // scattered into the cleared dense array; 0.25 B/bp cross PCIe instead of 0.375
                const uint64_t blk0 = c.a0 / 32, blk1 = blk0 + c.n_words / 2;
                auto first_at_least = [&](uint64_t b) {
                    uint64_t lo = 0, hi = src.n_exc;
                    while (lo < hi) { const uint64_t mid = lo + (hi - lo) / 2; if (src.exc[2 * mid] < b) lo = mid + 1; else hi = mid; }
                    return lo;
                };
                const uint64_t e0 = first_at_least(blk0), e1 = first_at_least(blk1);
                CK(cudaMemsetAsync(din + c.o_inv, 0, c.n_words * 2, s.stream));
                if (e1 > e0) {
                    const size_t o_exc = align_up(c.in_packed, 8);
                    CK(cudaMemcpyAsync(din + o_exc, src.exc + 2 * e0, (size_t)(e1 - e0) * 8, cudaMemcpyHostToDevice, s.stream));
                    acc.h2d += (e1 - e0) * 8;
                    inv_scatter_kernel<<<grid_for(ctx, e1 - e0, 128), 128, 0, s.stream>>>(
                        reinterpret_cast<uint32_t *>(din + c.o_inv), reinterpret_cast<const uint2 *>(din + o_exc), (uint32_t)(e1 - e0), (uint32_t)blk0);
                    ctx->launches += 1;
                }
            } else {
                CK(cudaMemcpyAsync(din + c.o_inv, src.inv + c.a0 / 16, c.n_words * 2, cudaMemcpyHostToDevice, s.stream));
                acc.h2d += c.n_words * 2;
            }
            CK(ship_offsets(din + c.o_off_p, off0));
            if (src.nl) {
                const uint64_t w0 = r_first / 32, w1 = (r_first + c.nr + 31) / 32;
                CK(cudaMemcpyAsync(din + c.o_nl, src.nl + w0, (size_t)(w1 - w0) * 4, cudaMemcpyHostToDevice, s.stream));
                acc.h2d += (w1 - w0) * 4;
                in.nl = reinterpret_cast<const uint32_t *>(din + c.o_nl);
                in.nl_bit0 = (uint32_t)(r_first % 32);
            }
            in.codes = reinterpret_cast<const uint32_t *>(din);
            in.inv = reinterpret_cast<const uint16_t *>(din + c.o_inv);
            d_off = reinterpret_cast<const uint64_t *>(din + c.o_off_p);
        } else if (route == 2 && n_exc >= 0) {
            // sparse form: codes, newline flags and the exception list in one copy (+ the offsets unless they are generated);
            // the dense non-ACGT bits the kernel reads are rebuilt on the device: cleared, then the listed blocks written
            const uint8_t *hin = s.h_in.as<uint8_t>();
            const size_t wire = cs.uniform ? c.s_exc + (size_t)n_exc * 8 : c.in_sparse;
            CK(cudaMemcpyAsync(din, hin, wire, cudaMemcpyHostToDevice, s.stream));
            acc.h2d += wire;
            if (cs.uniform) CK(ship_offsets(din + c.s_off, nullptr));
            CK(cudaMemsetAsync(din + c.s_inv, 0, c.n_words * 2, s.stream));
            if (n_exc) {
                inv_scatter_kernel<<<grid_for(ctx, (uint64_t)n_exc, 128), 128, 0, s.stream>>>(
                    reinterpret_cast<uint32_t *>(din + c.s_inv), reinterpret_cast<const uint2 *>(din + c.s_exc), (uint32_t)n_exc, 0u);
                ctx->launches += 1;
            }
            acc.n_packed++;
            in.codes = reinterpret_cast<const uint32_t *>(din);
            in.inv = reinterpret_cast<const uint16_t *>(din + c.s_inv);
            in.nl = reinterpret_cast<const uint32_t *>(din + c.s_nl);
            d_off = reinterpret_cast<const uint64_t *>(din + c.s_off);
        } else if (route == 2) {
            const uint8_t *hin = s.h_in.as<uint8_t>();
            if (cs.uniform) {   // codes, non-ACGT bits and newline flags in one copy; the offsets are generated
                CK(cudaMemcpyAsync(din, hin, c.o_off_p, cudaMemcpyHostToDevice, s.stream));
                acc.h2d += c.o_off_p;
                CK(ship_offsets(din + c.o_off_p, nullptr));
            } else {
                CK(cudaMemcpyAsync(din, hin, c.in_packed, cudaMemcpyHostToDevice, s.stream));
                acc.h2d += c.in_packed;
            }
            acc.n_packed++;
            in.codes = reinterpret_cast<const uint32_t *>(din);
            in.inv = reinterpret_cast<const uint16_t *>(din + c.o_inv);
            in.nl = reinterpret_cast<const uint32_t *>(din + c.o_nl);
            d_off = reinterpret_cast<const uint64_t *>(din + c.o_off_p);
        } else {
            acc.n_ascii++;
            if (c.nb) CK(cudaMemcpyAsync(din, bases + c.a0, (size_t)c.nb, cudaMemcpyHostToDevice, s.stream));
            acc.h2d += c.nb;
            CK(ship_offsets(din + c.o_off_a, off0));
            in.bases = din;
            d_off = reinterpret_cast<const uint64_t *>(din + c.o_off_a);
        }
        CK(cudaEventRecord(s.ev_h2d, s.stream));
        uint8_t *dout = s.out.as<uint8_t>();
        const int rc = enqueue_filter(ctx, s.plan, s.longs, s.dedup, in, c.a0, c.b1, d_off, c.nr, paired, prefix_len, abs_thr, rel_thr, deplete,
                                      dout + (size_t)c.nu * 8, reinterpret_cast<uint32_t *>(dout), reinterpret_cast<uint32_t *>(dout + (size_t)c.nu * 4),
                                      s.stream, &cs.st, time_fused);
        if (rc) return rc;
        CK(cudaEventRecord(s.ev_kernel, s.stream));
        if (out_pinned) {   // straight into the caller's arrays: no staging blob, no scatter on this thread
            CK(cudaMemcpyAsync(hits + c.u0, dout, (size_t)c.nu * 4, cudaMemcpyDeviceToHost, s.stream));
            CK(cudaMemcpyAsync(total + c.u0, dout + (size_t)c.nu * 4, (size_t)c.nu * 4, cudaMemcpyDeviceToHost, s.stream));
            CK(cudaMemcpyAsync(keep + c.u0, dout + (size_t)c.nu * 8, c.nu, cudaMemcpyDeviceToHost, s.stream));
        } else {
            CK(cudaMemcpyAsync(s.h_out.p, dout, c.out_bytes, cudaMemcpyDeviceToHost, s.stream));
        }
        acc.d2h += c.out_bytes;
        CK(cudaEventRecord(s.ev_done, s.stream));
        s.busy = true; s.u0 = c.u0; s.u1 = c.u1; s.packed = route != 0;
        return DCN_OK;
    };

    // ---- shared state of the two routes
    std::mutex m;
    int head = 0, tail = n_atoms, taken_by_packers = 0;   // ASCII takes atoms from head, a packer from tail
    int first_rc = DCN_OK;                                  // first failure of any thread; stops the others
    Acc sum;
    double pack_busy_ms = 0, pack_wait_ms = 0;              // summed over the packer threads
    auto merge = [&](const Acc &a, int rc) {
        std::lock_guard<std::mutex> g(m);
        sum.h2d += a.h2d; sum.d2h += a.d2h; sum.n_packed += a.n_packed; sum.n_ascii += a.n_ascii; sum.n_uniform += a.n_uniform;
        sum.packed_bases += a.packed_bases;
        sum.t_h2d += a.t_h2d; sum.t_kernel += a.t_kernel; sum.t_d2h += a.t_d2h;
        if (rc && !first_rc) first_rc = rc;
    };
    // stages per packer thread: enough that a thread never waits for its own earlier atoms, whose copies queue behind
    // the ASCII chunks already handed to the copy engine (2 stages: the threads idled a quarter of the call)
    static const int PST = []() { const char *e = getenv("DCN_PACKER_STAGES"); return e ? std::max(1, std::min(8, atoi(e))) : 4; }();
    if ((int)ctx->pslot.size() < PST * n_packers) ctx->pslot.resize((size_t)PST * n_packers);

    // A packer thread is a small pipeline of its own: claim atoms from the back of the batch (a few at a time while
    // plenty are left, one at a time near the end so all threads finish together), pack them into the pinned blob of
    // one of its stages, enqueue copy + kernels + results on that stage's stream, and collect the stage's previous
    // results before the blob is written again.  Nothing goes through the enqueueing thread: at ~90 us of CUDA API
    // calls per chunk, one thread enqueueing every packed atom was the limit of the whole pipeline (64 Gbp/s).
    auto packer = [&](int t) {
        Acc acc;
        int rc = DCN_OK;
        std::vector<uint64_t> bad32;
        std::vector<uint32_t> bad_mask;
        static const bool sparse_wire = []() { const char *e = getenv("DCN_SPARSE_MASK"); return !e || atoi(e) != 0; }();
        auto body = [&]() -> int {
            acc.t_begin = now_ms() - t_call0;
            CK(cudaSetDevice(ctx->device));
            for (int i = 0; i < PST; i++) {
                Slot &s = ctx->pslot[(size_t)(PST * t + i)];
                if (s.stream) continue;
                CK(cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking));
                CK(cudaEventCreate(&s.ev_start)); CK(cudaEventCreate(&s.ev_h2d));
                CK(cudaEventCreate(&s.ev_kernel)); CK(cudaEventCreate(&s.ev_done));
            }
            for (int i = 0; i < PST; i++) {   // all of the thread's staging in its first call (pinning memory is slow)
                Slot &s = ctx->pslot[(size_t)(PST * t + i)];
                if (s.h_in.ensure(std::max(big_packed.in_packed, big_packed.in_sparse)) != cudaSuccess ||
                    s.in.ensure(std::max(big_packed.in_packed, big_packed.dev_sparse) + 8) != cudaSuccess ||
                    s.out.ensure(big_packed.out_bytes) != cudaSuccess || s.h_out.ensure(big_packed.out_bytes) != cudaSuccess)
                    return ctx->fail(DCN_ERR_NOMEM, "staging allocation failed", cudaGetLastError());
            }
            for (int flip = 0;; flip = (flip + 1) % PST) {
                int a_lo, a_hi;
                {
                    std::lock_guard<std::mutex> g(m);
                    if (first_rc || tail <= head || taken_by_packers >= pack_budget) break;
                    int grab = std::max(1, std::min<int>(packer_grab_max, (tail - head) / n_packers));
                    grab = std::min(grab, std::min(tail - head, pack_budget - taken_by_packers));
                    a_hi = tail; a_lo = tail -= grab; taken_by_packers += grab;
                }
                const ChunkPlan c = make_chunk(a_lo, a_hi);
                Slot &s = ctx->pslot[(size_t)(PST * t + flip)];
                int r = retire(s, acc);   // the blob's previous copy has left the host once its results are back
                if (r) return r;
                const double t0 = now_ms();
                if (s.h_in.ensure(std::max(std::max(c.in_packed, c.in_sparse), std::max(big_packed.in_packed, big_packed.in_sparse))) != cudaSuccess)
                    return ctx->fail(DCN_ERR_NOMEM, "pinned staging allocation failed", cudaGetLastError());
                uint8_t *hin = s.h_in.as<uint8_t>();
                const uint64_t *off0 = rec_off + (uint64_t)c.u0 * rpu;
                const ChunkStats cs = chunk_stats(off0, c.nu, rpu);
                uint32_t *h_codes = reinterpret_cast<uint32_t *>(hin);
                int64_t n_exc = -1;
                if (sparse_wire) {
                    // codes and newline flags into the blob; the non-ACGT bits only as (block, mask) pairs of the blocks that
                    // have any (real reads: a handful per chunk): 0.25 B/bp cross PCIe instead of 0.375
                    uint32_t *h_nl = reinterpret_cast<uint32_t *>(hin + c.s_nl);
                    pack_records(bases, c.a0, c.nb, off0, c.nr, ctx->k, prefix_len, h_codes, nullptr, h_nl, bad32, &bad_mask);
                    if (bad32.size() <= c.exc_cap) {
                        n_exc = (int64_t)bad32.size();
                        uint32_t *ex = reinterpret_cast<uint32_t *>(hin + c.s_exc);
                        for (size_t i = 0; i < bad32.size(); i++) { ex[2 * i] = (uint32_t)bad32[i]; ex[2 * i + 1] = bad_mask[i]; }
                        if (!cs.uniform) memcpy(hin + c.s_off, off0, ((size_t)c.nr + 1) * 8);
                    } else {   // N-rich chunk: the dense form is smaller; rebuild it from the list, flags moved to their dense place
                        uint16_t *h_inv = reinterpret_cast<uint16_t *>(hin + c.o_inv);
                        memmove(hin + c.o_nl, h_nl, ((size_t)c.nr + 31) / 32 * 4);
                        memset(h_inv, 0, c.n_words * 2);
                        for (size_t i = 0; i < bad32.size(); i++) memcpy(h_inv + 2 * bad32[i], &bad_mask[i], 4);
                        if (!cs.uniform) memcpy(hin + c.o_off_p, off0, ((size_t)c.nr + 1) * 8);
                    }
                } else {
                    if (!cs.uniform) memcpy(hin + c.o_off_p, off0, ((size_t)c.nr + 1) * 8);
                    uint16_t *h_inv = reinterpret_cast<uint16_t *>(hin + c.o_inv);
                    uint32_t *h_nl = reinterpret_cast<uint32_t *>(hin + c.o_nl);
                    // codes, non-ACGT bits and newline flags in one pass over the atom (dcn_host_pack.h)
                    pack_records(bases, c.a0, c.nb, off0, c.nr, ctx->k, prefix_len, h_codes, h_inv, h_nl, bad32);
                }
                const double t1 = now_ms();
                acc.pack_ms += t1 - t0;
                acc.packed_bases += c.nb;
                if ((r = ship(s, c, 2, cs, acc, false, n_exc))) return r;
                acc.t_last_ship = now_ms() - t_call0;
                acc.ship_ms += acc.t_last_ship - (t1 - t_call0);
                if (acc.t_first_ship == 0) acc.t_first_ship = acc.t_last_ship;
            }
            acc.t_end = now_ms() - t_call0;
            for (int i = 0; i < PST; i++) {
                const int r = retire(ctx->pslot[(size_t)(PST * t + i)], acc);
                if (r) return r;
            }
            return DCN_OK;
        };
        rc = body();
        merge(acc, rc);
        std::lock_guard<std::mutex> g(m);
        pack_busy_ms += acc.pack_ms; pack_wait_ms += acc.wait_ms;
        if (trace_level >= 1)
            fprintf(stderr, "[dcn packer %d] begin %.2f first ship %.2f last ship %.2f claims done %.2f end %.2f ms; packing %.2f shipping %.2f waiting %.2f ms; %.1f MB\n",
                    t, acc.t_begin, acc.t_first_ship, acc.t_last_ship, acc.t_end, now_ms() - t_call0, acc.pack_ms, acc.ship_ms, acc.wait_ms, acc.packed_bases / 1e6);
    };
    std::vector<std::thread> packers;
    for (int t = 0; t < n_packers; t++) {
        try {
            packers.emplace_back(packer, t);
        } catch (const std::exception &) {   // thread limit reached: go on with the threads there are
            break;
        }
    }
    if (packers.empty() && !ascii_route) ascii_route = true;   // nobody to pack: everything goes as ASCII

    // ---- the enqueueing thread: ASCII chunks (or the caller's packed arrays) from the front, at the pace of the link
    Acc main_acc;
    int which = 0, rc = DCN_OK;
    while (ascii_route && rc == DCN_OK) {
        Slot &s = ctx->slot[which];
        const double h0 = now_ms();
        if ((rc = retire(s, main_acc))) break;  // the stage's previous chunk (NSLOT chunks ago)
        const double h1 = now_ms();
        if (n_packers > 0) {
            // Beside packers the route is decided as late as possible: no more than two ASCII copies are handed to the
            // copy engine ahead of time (one running, one queued: the link never idles), and the chunks shrink as the
            // two fronts close in, so that when they meet little is left queued while the packer threads sit idle.
            Slot &prev2 = ctx->slot[(which + dcn_ctx::NSLOT - 2) % dcn_ctx::NSLOT];
            if (prev2.busy) {
                const double w0 = now_ms();
                const cudaError_t e = cudaEventSynchronize(prev2.ev_h2d);   // (no CK here: the packer threads must be joined)
                if (e != cudaSuccess) { rc = ctx->fail(DCN_ERR_CUDA, "cudaEventSynchronize(ev_h2d)", e); break; }
                main_acc.wait_ms += now_ms() - w0;
            }
        }
        int a_lo, a_hi;
        {
            std::lock_guard<std::mutex> g(m);
            if (first_rc || head >= tail) break;
            // beside packers an ASCII chunk is a few atoms only: the packed chunks' small copies queue behind it on the copy engine
            static const int ascii_atoms = []() { const char *e = getenv("DCN_ASCII_ATOMS"); return e ? std::max(1, atoi(e)) : 2; }();   // measured: 8 atoms 89.5, 2 atoms 97.7 Gbp/s
            const int grab = n_packers > 0 ? std::max(1, std::min<int>(std::min<int>((int)atoms_per_chunk, ascii_atoms), (tail - head) / 6)) : (int)atoms_per_chunk;
            a_lo = head;
            a_hi = head = std::min<int>(tail, head + grab);
        }
        const double h2 = now_ms();
        const ChunkPlan c = make_chunk(a_lo, a_hi);
        const ChunkStats cs = chunk_stats(rec_off + (uint64_t)c.u0 * rpu, c.nu, rpu);
        const double h3 = now_ms();
        if ((rc = ship(s, c, prepacked ? 1 : 0, cs, main_acc, true))) break;
        if (trace_level >= 3)
            fprintf(stderr, "[dcn main] units %u..%u host ms since call: loop %.3f retired %.3f claimed %.3f stats %.3f shipped %.3f\n",
                    c.u0, c.u1, h0 - t_call0, h1 - t_call0, h2 - t_call0, h3 - t_call0, now_ms() - t_call0);
        which = (which + 1) % dcn_ctx::NSLOT;
    }
    for (int i = 0; i < dcn_ctx::NSLOT; i++) {   // oldest first
        int r2 = retire(ctx->slot[(which + i) % dcn_ctx::NSLOT], main_acc);
        if (!rc) rc = r2;
    }
    merge(main_acc, rc);
    for (auto &t : packers) t.join();
    rc = first_rc;
    if (rc) {
        cudaDeviceSynchronize();
        for (auto &sl : ctx->slot) sl.busy = false;
        for (auto &sl : ctx->pslot) sl.busy = false;
    }
    ctx->t_h2d = sum.t_h2d; ctx->t_kernel = sum.t_kernel; ctx->t_d2h = sum.t_d2h;
    ctx->bytes_h2d = sum.h2d; ctx->bytes_d2h = sum.d2h;
    ctx->n_packed_chunks += sum.n_packed; ctx->n_ascii_chunks += sum.n_ascii; ctx->n_uniform_chunks += sum.n_uniform;
    if (ev_call) cudaEventDestroy(ev_call);
    const double call_ms = now_ms() - t_call0;
    if (trace_level >= 1)
        fprintf(stderr, "[dcn host] %d atoms, %d packers: %llu packed / %llu ascii chunks; call %.2f ms; enqueuer waited %.2f ms on its stages; "
                "packers: packing %.2f ms, waiting for a stage %.2f ms (sums over threads); h2d %.1f MB\n",
                n_atoms, n_packers, (unsigned long long)sum.n_packed, (unsigned long long)sum.n_ascii, call_ms, main_acc.wait_ms,
                pack_busy_ms, pack_wait_ms, sum.h2d / 1e6);
    if (sum.packed_bases) {
        ctx->t_pack = (float)call_ms;
        if (pack_busy_ms > 0) ctx->pack_gbps = (double)sum.packed_bases / 1e6 / pack_busy_ms * std::max(1, n_packers);
    }
    return rc;
}

int dcn_filter_batch(dcn_ctx *ctx, const uint8_t *bases, const uint64_t *rec_off, uint32_t n_rec, int paired,
                     uint32_t prefix_len, uint32_t abs_thr, double rel_thr, int deplete,
                     uint8_t *keep, uint32_t *hits, uint32_t *total) {
    if (!ctx) return DCN_ERR_ARG;
    HostSrc src;
    src.bases = bases;
    return filter_pipeline(ctx, src, rec_off, n_rec, paired, prefix_len, abs_thr, rel_thr, deplete, keep, hits, total);
}

int dcn_filter_batch_packed(dcn_ctx *ctx, const uint32_t *codes, const uint16_t *inv, const uint32_t *nl_bits,
                            const uint64_t *rec_off, uint32_t n_rec, int paired, uint32_t prefix_len, uint32_t abs_thr,
                            double rel_thr, int deplete, uint8_t *keep, uint32_t *hits, uint32_t *total) {
    if (!ctx) return DCN_ERR_ARG;
    if (!codes || !inv) return ctx->fail(DCN_ERR_ARG, "null input pointer");
    if (n_rec && rec_off && rec_off[0] != 0) return ctx->fail(DCN_ERR_ARG, "rec_off[0] must be 0");
    HostSrc src;
    src.codes = codes; src.inv = inv; src.nl = nl_bits;
    return filter_pipeline(ctx, src, rec_off, n_rec, paired, prefix_len, abs_thr, rel_thr, deplete, keep, hits, total);
}

int dcn_filter_batch_packed_sparse(dcn_ctx *ctx, const uint32_t *codes, const uint32_t *exc, uint64_t n_exc, const uint32_t *nl_bits,
                                   const uint64_t *rec_off, uint32_t n_rec, int paired, uint32_t prefix_len, uint32_t abs_thr,
                                   double rel_thr, int deplete, uint8_t *keep, uint32_t *hits, uint32_t *total) {
    if (!ctx) return DCN_ERR_ARG;
    if (!codes || (!exc && n_exc)) return ctx->fail(DCN_ERR_ARG, "null input pointer");
    if (n_rec && rec_off && rec_off[0] != 0) return ctx->fail(DCN_ERR_ARG, "rec_off[0] must be 0");
    for (uint64_t i = 1; i < n_exc; i++)
        if (exc[2 * i] <= exc[2 * i - 2]) return ctx->fail(DCN_ERR_ARG, "the exception list must be in ascending block order");
    static const uint32_t none[2] = {0xFFFFFFFFu, 0u};
    HostSrc src;
    src.codes = codes; src.exc = n_exc ? exc : none; src.n_exc = n_exc; src.nl = nl_bits;
    return filter_pipeline(ctx, src, rec_off, n_rec, paired, prefix_len, abs_thr, rel_thr, deplete, keep, hits, total);
}

int dcn_pack_records_sparse(const uint8_t *bases, const uint64_t *rec_off, uint32_t n_rec, uint8_t k, uint32_t prefix_len,
                            uint32_t *codes, uint32_t *exc, uint64_t exc_cap, uint64_t *n_exc, uint32_t *nl_bits) {
    if (!rec_off || !codes || !n_exc || !nl_bits || (!exc && exc_cap) || (!bases && n_rec && rec_off[n_rec] > 0)) return DCN_ERR_ARG;
    std::vector<uint64_t> bad32;
    std::vector<uint32_t> bad_mask;
    pack_records(bases, 0, n_rec ? rec_off[n_rec] : 0, rec_off, n_rec, k, prefix_len, codes, nullptr, nl_bits, bad32, &bad_mask);
    *n_exc = bad32.size();
    if (bad32.size() > exc_cap) return DCN_ERR_OVERFLOW;
    for (size_t i = 0; i < bad32.size(); i++) { exc[2 * i] = (uint32_t)bad32[i]; exc[2 * i + 1] = bad_mask[i]; }
    return DCN_OK;
}

int dcn_newline_bits(const uint8_t *bases, const uint64_t *rec_off, uint32_t n_rec, uint8_t k, uint32_t prefix_len,
                     uint32_t *nl_bits) {
    if (!rec_off || !nl_bits || (!bases && n_rec && rec_off[n_rec] > rec_off[0])) return DCN_ERR_ARG;
    for (uint32_t w = 0; w < (n_rec + 31) / 32; w++) nl_bits[w] = 0;
    for (uint32_t r = 0; r < n_rec; r++) {
        const uint64_t len = rec_off[r + 1] - rec_off[r];
        if (len < (uint64_t)k) continue;                                                        // src/filter_common.rs:217-219
        const uint64_t n = (prefix_len > 0 && len > prefix_len) ? prefix_len : len;             // :222-226
        if (bases[rec_off[r] + n - 1] == (uint8_t)'\n') nl_bits[r >> 5] |= 1u << (r & 31u);    // :229
    }
    return DCN_OK;
}

int dcn_host_pack_fraction(dcn_ctx *ctx, double fraction) {
    if (!ctx) return DCN_ERR_ARG;
    if (fraction > 1.0) return ctx->fail(DCN_ERR_ARG, "fraction must be <= 1 (negative = automatic)");
    ctx->pack_fraction = fraction;
    return DCN_OK;
}

int dcn_host_pack_threads(dcn_ctx *ctx, int n_threads) {
    if (!ctx) return DCN_ERR_ARG;
    if (n_threads < 0 || n_threads > 256) return ctx->fail(DCN_ERR_ARG, "n_threads must be in 0..=256");
    ctx->pack_threads = n_threads;
    ctx->pack_gbps = 0;

    return DCN_OK;
}

int dcn_pack_records(const uint8_t *bases, const uint64_t *rec_off, uint32_t n_rec, uint8_t k, uint32_t prefix_len,
                     uint32_t *codes, uint16_t *inv, uint32_t *nl_bits) {
    if (!rec_off || !codes || !inv || !nl_bits || (!bases && n_rec && rec_off[n_rec] > 0)) return DCN_ERR_ARG;
    std::vector<uint64_t> bad32;
    pack_records(bases, 0, n_rec ? rec_off[n_rec] : 0, rec_off, n_rec, k, prefix_len, codes, inv, nl_bits, bad32);
    return DCN_OK;
}

int dcn_pack_ascii(const uint8_t *bases, uint64_t n_bases, uint32_t *codes, uint16_t *inv) {
    if ((!bases && n_bases) || !codes || !inv) return DCN_ERR_ARG;
    pack_ascii(bases, n_bases, codes, inv, 1);
    return DCN_OK;
}

// ---------------------------------------------------------------------------- B2 lookup
static int lookup_device(dcn_ctx *ctx, const uint64_t *d_hashes, const uint64_t *d_rec_off, uint32_t n_rec,
                         uint32_t abs_thr, double rel_thr, int deplete, uint8_t *d_keep, uint32_t *d_hits,
                         uint32_t *d_total, uint8_t *d_flags, void *stream) {
    if (!ctx) return DCN_ERR_ARG;
    if (!ctx->table.p) return ctx->fail(DCN_ERR_NO_INDEX, "no index resident: call dcn_index_upload first");
    if (n_rec == 0) return DCN_OK;
    CK(cudaSetDevice(ctx->device));
    cudaStream_t st = (cudaStream_t)stream;
    TableView tv;
    tv.slots = ctx->table.as<uint64_t>(); tv.n_buckets = ctx->n_buckets; tv.has_empty_key = ctx->has_empty;
    CK(ctx->plan.ensure(256));
    BatchStats *d_stats = ctx->plan.as<BatchStats>();
    const int pb = 256;
    const int pg = (int)std::min<uint64_t>(((uint64_t)n_rec + pb - 1) / pb, (uint64_t)ctx->sm_count * 8);
    uint64_t dedup_cap = 0;
    for (int attempt = 0; attempt < 4; attempt++) {
        CK(cudaMemsetAsync(d_stats, 0, sizeof(BatchStats), st));
        prep_stats_kernel<<<pg, pb, 0, st>>>(d_rec_off, 1, n_rec, d_stats);   // records with > 1024 hashes
        BatchStats hs;
        CK(cudaMemcpyAsync(&hs, d_stats, sizeof(hs), cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        DedupView dd;
        dd.slots = nullptr; dd.cap = 0; dd.overflow = &d_stats->overflow; dd.epoch = 1; dd.per16 = 0;
        if (hs.n_long) {
            if (!dedup_cap) dedup_cap = std::max<uint64_t>(4096, hs.long_bases * 2);
            CK(open_dedup_set(ctx->dedup, dedup_cap, st, dd, &d_stats->overflow));
        }
        const int grid = (int)std::min<uint64_t>(((uint64_t)n_rec * 32 + 255) / 256, (uint64_t)ctx->sm_count * 8);
        lookup_kernel<<<grid, 256, 0, st>>>(d_hashes, d_rec_off, n_rec, tv, dd, abs_thr, rel_thr, deplete, d_keep, d_hits, d_total, d_flags);
        ctx->launches += 2;
        CK(cudaGetLastError());
        if (!hs.n_long) break;
        BatchStats after;
        CK(cudaMemcpyAsync(&after, d_stats, sizeof(after), cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        if (!after.overflow) break;
        if (attempt == 3) return ctx->fail(DCN_ERR_OVERFLOW, "distinct-hit set overflowed after 4 attempts");
        dedup_cap *= 4;
    }
    return DCN_OK;
}

int dcn_lookup_batch_device(dcn_ctx *ctx, const uint64_t *d_hashes, const uint64_t *d_rec_off, uint32_t n_rec,
                            uint32_t abs_thr, double rel_thr, int deplete, uint8_t *d_keep, uint32_t *d_hits,
                            uint32_t *d_total, void *stream) {
    return lookup_device(ctx, d_hashes, d_rec_off, n_rec, abs_thr, rel_thr, deplete, d_keep, d_hits, d_total, nullptr, stream);
}

int dcn_lookup_batch(dcn_ctx *ctx, const uint64_t *hashes, const uint64_t *rec_off, uint32_t n_rec, uint32_t abs_thr,
                     double rel_thr, int deplete, uint8_t *keep, uint32_t *hits, uint32_t *total) {
    return dcn_lookup_batch_flags(ctx, hashes, rec_off, n_rec, abs_thr, rel_thr, deplete, keep, hits, total, nullptr);
}

int dcn_lookup_batch_flags(dcn_ctx *ctx, const uint64_t *hashes, const uint64_t *rec_off, uint32_t n_rec, uint32_t abs_thr,
                           double rel_thr, int deplete, uint8_t *keep, uint32_t *hits, uint32_t *total, uint8_t *hit_flags) {
    if (!ctx) return DCN_ERR_ARG;
    if (!rec_off || !keep || !hits || !total) return ctx->fail(DCN_ERR_ARG, "null pointer");
    if (n_rec == 0) return DCN_OK;
    if (rec_off[0] != 0) return ctx->fail(DCN_ERR_ARG, "rec_off[0] must be 0");
    const uint64_t n_hash = rec_off[n_rec];
    if (n_hash && !hashes) return ctx->fail(DCN_ERR_ARG, "null hash pointer");
    CK(cudaSetDevice(ctx->device));
    Slot &s = ctx->slot[0];
    cudaStream_t st = s.stream;
    const size_t o_off = align_up(n_hash * 8, 8), o_tot = (size_t)n_rec * 4, o_keep = (size_t)n_rec * 8;
    CK(s.in.ensure(o_off + ((size_t)n_rec + 1) * 8));
    const size_t o_flags = align_up((size_t)n_rec * 9, 16);
    CK(s.out.ensure(o_flags + (hit_flags ? n_hash : 0)));
    uint8_t *din = s.in.as<uint8_t>(), *dout = s.out.as<uint8_t>();
    if (n_hash) CK(cudaMemcpyAsync(din, hashes, n_hash * 8, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(din + o_off, rec_off, (size_t)(n_rec + 1) * 8, cudaMemcpyHostToDevice, st));
    int rc = lookup_device(ctx, reinterpret_cast<uint64_t *>(din), reinterpret_cast<uint64_t *>(din + o_off), n_rec, abs_thr,
                           rel_thr, deplete, dout + o_keep, reinterpret_cast<uint32_t *>(dout),
                           reinterpret_cast<uint32_t *>(dout + o_tot), hit_flags ? dout + o_flags : nullptr, st);
    if (rc) return rc;
    if (hit_flags && n_hash) CK(cudaMemcpyAsync(hit_flags, dout + o_flags, n_hash, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(keep, dout + o_keep, n_rec, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(hits, dout, (size_t)n_rec * 4, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(total, dout + o_tot, (size_t)n_rec * 4, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    return DCN_OK;
}

// ---------------------------------------------------------------------------- B3 extraction
// Fast path: filter flavour, k = 31, w = 15, no record above DCN_MAX_SHORT bases -> the tile pipeline.
// Leaves the CSR in ctx->gx_h / gx_p / gx_oo like generic_extract_device.
static int tile_extract_device(dcn_ctx *ctx, const uint8_t *d_bases, const uint64_t *d_off, uint32_t n_rec, uint64_t n_bases,
                               uint32_t prefix_len, bool want_pos, uint64_t cap_limit, cudaStream_t st, uint64_t *n_out,
                               bool *written) {
    *n_out = 0; *written = false;
    const bool warp_impl = ctx->fused_impl == 0;
    const uint64_t wtile_cap = wplan_tile_cap(n_bases, n_rec), wovf_cap = wplan_ovf_cap(n_bases);
    const size_t pbytes = warp_impl ? 128 + (size_t)wtile_cap * sizeof(WTile) + (size_t)wovf_cap * 4 : plan_bytes(n_bases);
    CK(ctx->plan.ensure(pbytes));
    const uint64_t n_tiles_max = (plan_bytes(n_bases) - 64) / (2 * sizeof(uint32_t));
    BatchStats *d_stats = ctx->plan.as<BatchStats>();
    uint32_t *tile_first = reinterpret_cast<uint32_t *>(ctx->plan.as<uint8_t>() + 64);
    uint32_t *tile_end = tile_first + n_tiles_max;
    WTile *wtiles = reinterpret_cast<WTile *>(ctx->plan.as<uint8_t>() + 128);
    uint32_t *wovf = reinterpret_cast<uint32_t *>(ctx->plan.as<uint8_t>() + 128 + (size_t)wtile_cap * sizeof(WTile));
    CK(ctx->gx_rc.ensure(((size_t)n_rec + 1) * 8));     // valid picks per record, then (after the scan) unused
    CK(ctx->gx_oo.ensure(((size_t)n_rec + 1) * 8));     // CSR offsets
    CK(ctx->gx_cc.ensure(((size_t)n_rec + 1) * 8 + 64)); // temp start | pick count per record (+ the cursor)
    FilterParams P;
    memset(&P, 0, sizeof(P));
    P.bases = d_bases; P.base0 = 0; P.n_bases = n_bases; P.rec_off = d_off; P.n_rec = n_rec; P.rpu = 1; P.n_units = n_rec;
    P.prefix_len = prefix_len;
    unsigned long long *d_cursor = reinterpret_cast<unsigned long long *>(ctx->gx_cc.as<uint8_t>() + ((size_t)n_rec + 1) * 8);
    const int pb = 256;
    const int pg = (int)std::min<uint64_t>(((uint64_t)n_rec + pb - 1) / pb, (uint64_t)ctx->sm_count * 8);
    const int grid = (int)std::min<uint64_t>(n_bases / G31::BCAP + 1, (uint64_t)ctx->sm_count * (1024 / G31::NT));
    const int wgrid = (int)std::min<uint64_t>((n_bases / WG::TB + DCN_WARPS) / DCN_WARPS, (uint64_t)ctx->sm_count * DCN_WCTAS);
    const uint64_t n_seg = (n_bases + DCN_WSEG - 1) / DCN_WSEG;
    const int sg = (int)std::max<uint64_t>(1, std::min<uint64_t>((n_seg + 7) / 8, (uint64_t)ctx->sm_count * 8));
    // ~0.095 picks per base for 150-base records; the warp kernel's warps take the temp arrays in blocks of DCN_XBLK
    uint64_t cap = (uint64_t)((double)n_bases * 0.13) + 4096 + (warp_impl ? (uint64_t)wgrid * DCN_WARPS * DCN_XBLK : 0);
    unsigned long long used = 0;
    for (int attempt = 0; attempt < 2; attempt++) {
        CK(ctx->ib_alt.ensure(cap * 8));     // temp hashes
        CK(ctx->ib_tmp.ensure(cap * 4));     // temp positions
        CK(cudaMemsetAsync(ctx->plan.p, 0, warp_impl ? 128 : pbytes, st));
        CK(cudaMemsetAsync(d_cursor, 0, 8, st));
        CK(cudaMemsetAsync(ctx->gx_rc.p, 0, ((size_t)n_rec + 1) * 8, st));
        P.xo.tmp_h = ctx->ib_alt.as<uint64_t>(); P.xo.tmp_p = ctx->ib_tmp.as<uint32_t>(); P.xo.tmp_cap = cap;
        P.xo.cursor = d_cursor; P.xo.rec_cnt = ctx->gx_rc.as<uint64_t>(); P.xo.rec_tmp = ctx->gx_cc.as<uint64_t>();
        if (warp_impl) {
            wplan_kernel<<<sg, 256, 0, st>>>(d_off, 1, n_rec, 0, n_bases, d_stats, wtiles, (uint32_t)std::min<uint64_t>(wtile_cap, 0xFFFFFFFFull), nullptr);
            extract_warp_kernel<<<wgrid, DCN_WARPS * 32, warp_kernel_smem(), st>>>(P, d_stats, wtiles, wovf, (uint32_t)std::min<uint64_t>(wovf_cap, 0xFFFFFFFFull));
            extract_tail_kernel<G31><<<std::min(grid, 16), G31::NT, sizeof(TileSmem<G31>), st>>>(P, d_stats, wovf);
        } else {
            prep_stats_kernel<<<pg, pb, 0, st>>>(d_off, 1, n_rec, d_stats);
            prep_tiles_kernel<G31><<<pg, pb, 0, st>>>(d_off, 1, n_rec, 0, d_stats, tile_first, tile_end);
            extract_tiles_kernel<G31><<<grid, G31::NT, sizeof(TileSmem<G31>), st>>>(P, d_stats, tile_first, tile_end);
        }
        ctx->launches += 3;
        CK(cudaMemcpyAsync(&used, d_cursor, 8, cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        CK(cudaGetLastError());
        if (used <= cap) break;
        if (attempt == 1) return ctx->fail(DCN_ERR_OVERFLOW, "pick buffer overflowed twice");
        cap = used + 1024 + (warp_impl ? (uint64_t)wgrid * DCN_WARPS * DCN_XBLK : 0);
    }
    // records that own no window keep count 0 (memset); exclusive scan -> CSR offsets
    size_t tb = 0;
    CK(cub::DeviceScan::ExclusiveSum(nullptr, tb, ctx->gx_rc.as<uint64_t>(), ctx->gx_oo.as<uint64_t>(), (int64_t)n_rec + 1, st));
    CK(ctx->gx_tmp.ensure(tb));
    CK(cub::DeviceScan::ExclusiveSum(ctx->gx_tmp.p, tb, ctx->gx_rc.as<uint64_t>(), ctx->gx_oo.as<uint64_t>(), (int64_t)n_rec + 1, st));
    uint64_t m = 0;
    CK(cudaMemcpyAsync(&m, ctx->gx_oo.as<uint64_t>() + n_rec, 8, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    *n_out = m;
    ctx->launches += 1;
    if (cap_limit && m > cap_limit) return DCN_OK;
    CK(ctx->gx_h.ensure(std::max<uint64_t>(m, 1) * 8));
    if (want_pos) CK(ctx->gx_p.ensure(std::max<uint64_t>(m, 1) * 4));
    extract_compact_kernel<<<grid_for(ctx, (uint64_t)n_rec * 32, 256), 256, 0, st>>>(P.xo, ctx->gx_oo.as<uint64_t>(), n_rec, ctx->gx_h.as<uint64_t>(),
                                                                                  want_pos ? ctx->gx_p.as<uint32_t>() : nullptr);
    ctx->launches += 1;
    CK(cudaGetLastError());
    *written = true;
    return DCN_OK;
}

int dcn_extract_device(dcn_ctx *ctx, int flavour, const uint8_t *d_bases, const uint64_t *d_rec_off, uint32_t n_rec,
                       uint64_t n_bases, uint8_t k, uint8_t w, uint32_t prefix_len, float entropy_thr, uint64_t *d_out_hashes,
                       uint32_t *d_out_pos, uint64_t *d_out_off, uint64_t out_cap, uint64_t *n_out, void *stream) {
    if (!ctx) return DCN_ERR_ARG;
    if (flavour != DCN_FLAVOUR_FILTER && flavour != DCN_FLAVOUR_INDEX) return ctx->fail(DCN_ERR_ARG, "unknown flavour");
    if (!d_rec_off || !d_out_off || !n_out || (!d_out_hashes && out_cap)) return ctx->fail(DCN_ERR_ARG, "null pointer");
    int rc = check_kw(ctx, k, w, flavour);
    if (rc) return rc;
    CK(cudaSetDevice(ctx->device));
    cudaStream_t st = (cudaStream_t)stream;
    *n_out = 0;
    bool tiles = flavour == DCN_FLAVOUR_FILTER && k == 31 && w == 15 && n_rec > 0 && n_bases > 0 &&
                 (reinterpret_cast<uintptr_t>(d_bases) & 15u) == 0 && !getenv("DCN_EXTRACT_GENERIC");
    if (tiles) {   // the tile pipeline takes whole short records only: ask the device for the longest one
        CK(ctx->plan.ensure(256));
        BatchStats hs;
        CK(cudaMemsetAsync(ctx->plan.p, 0, sizeof(BatchStats), st));
        prep_stats_kernel<<<grid_for(ctx, n_rec, 256), 256, 0, st>>>(d_rec_off, 1, n_rec, ctx->plan.as<BatchStats>());
        CK(cudaMemcpyAsync(&hs, ctx->plan.p, sizeof(hs), cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        ctx->launches += 1;
        tiles = hs.n_long == 0;
    }
    uint64_t m = 0;
    bool written = false;
    if (tiles)
        rc = tile_extract_device(ctx, d_bases, d_rec_off, n_rec, n_bases, prefix_len, d_out_pos != nullptr, out_cap ? out_cap : 1, st, &m, &written);
    else
        rc = generic_extract_device(ctx, flavour, d_bases, d_rec_off, n_rec, k, w, flavour == DCN_FLAVOUR_FILTER ? prefix_len : 0,
                                    entropy_thr, d_out_pos != nullptr, out_cap ? out_cap : 1, st, &m, &written);
    if (rc) return rc;
    *n_out = m;
    CK(cudaMemcpyAsync(d_out_off, ctx->gx_oo.p, ((size_t)n_rec + 1) * 8, cudaMemcpyDeviceToDevice, st));
    if (m > out_cap) return ctx->fail(DCN_ERR_OVERFLOW, "out_cap too small: *n_out holds the required capacity");
    if (written && m) {
        CK(cudaMemcpyAsync(d_out_hashes, ctx->gx_h.p, m * 8, cudaMemcpyDeviceToDevice, st));
        if (d_out_pos) CK(cudaMemcpyAsync(d_out_pos, ctx->gx_p.p, m * 4, cudaMemcpyDeviceToDevice, st));
    }
    return DCN_OK;
}

int dcn_extract(dcn_ctx *ctx, int flavour, const uint8_t *bases, const uint64_t *rec_off, uint32_t n_rec, uint8_t k,
                uint8_t w, uint32_t prefix_len, float entropy_thr, uint64_t *out_hashes, uint32_t *out_pos,
                uint64_t *out_off, uint64_t out_cap) {
    if (!ctx) return DCN_ERR_ARG;
    if (flavour != DCN_FLAVOUR_FILTER && flavour != DCN_FLAVOUR_INDEX) return ctx->fail(DCN_ERR_ARG, "unknown flavour");
    if (!rec_off || !out_off || (!out_hashes && out_cap)) return ctx->fail(DCN_ERR_ARG, "null pointer");
    int rc = check_kw(ctx, k, w, flavour);
    if (rc) return rc;
    if (n_rec && rec_off[0] != 0) return ctx->fail(DCN_ERR_ARG, "rec_off[0] must be 0");
    const uint64_t n_bases = n_rec ? rec_off[n_rec] : 0;
    if (n_bases && !bases) return ctx->fail(DCN_ERR_ARG, "null input pointer");
    CK(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    CK(ctx->gx_bases.ensure(n_bases + 64));
    CK(ctx->gx_off.ensure(((size_t)n_rec + 1) * 8));
    if (n_bases) CK(cudaMemcpyAsync(ctx->gx_bases.p, bases, n_bases, cudaMemcpyHostToDevice, st));
    if (n_rec) CK(cudaMemcpyAsync(ctx->gx_off.p, rec_off, ((size_t)n_rec + 1) * 8, cudaMemcpyHostToDevice, st));
    else CK(cudaMemsetAsync(ctx->gx_off.p, 0, 8, st));
    uint64_t m = 0;
    bool written = false;
    bool tiles = flavour == DCN_FLAVOUR_FILTER && k == 31 && w == 15 && n_rec > 0 && n_bases > 0 && !getenv("DCN_EXTRACT_GENERIC");
    for (uint32_t r = 0; tiles && r < n_rec; r++) tiles = rec_off[r + 1] - rec_off[r] <= DCN_MAX_SHORT;
    if (tiles)
        rc = tile_extract_device(ctx, ctx->gx_bases.as<uint8_t>(), ctx->gx_off.as<uint64_t>(), n_rec, n_bases, prefix_len,
                                 out_pos != nullptr, out_cap ? out_cap : 1, st, &m, &written);
    else
        rc = generic_extract_device(ctx, flavour, ctx->gx_bases.as<uint8_t>(), ctx->gx_off.as<uint64_t>(), n_rec, k, w,
                                    flavour == DCN_FLAVOUR_FILTER ? prefix_len : 0, entropy_thr, out_pos != nullptr, out_cap ? out_cap : 1,
                                    st, &m, &written);
    if (rc) return rc;
    CK(cudaMemcpyAsync(out_off, ctx->gx_oo.p, ((size_t)n_rec + 1) * 8, cudaMemcpyDeviceToHost, st));
    if (written && m && m <= out_cap) {   // (a size query passes out_cap = 0 and no output arrays: nothing to copy then)
        CK(cudaMemcpyAsync(out_hashes, ctx->gx_h.p, m * 8, cudaMemcpyDeviceToHost, st));
        if (out_pos) CK(cudaMemcpyAsync(out_pos, ctx->gx_p.p, m * 4, cudaMemcpyDeviceToHost, st));
    }
    CK(cudaStreamSynchronize(st));
    if (m > out_cap) return ctx->fail(DCN_ERR_OVERFLOW, "out_cap too small: out_off[n_rec] holds the required capacity");
    return DCN_OK;
}

// FxHashSet::extend (src/index.rs:267-284) == radix sort + unique of the n_picks hashes in ib_alt
static int index_sort_unique(dcn_ctx *ctx, uint64_t n_picks, uint8_t k, uint8_t w, int make_resident, uint64_t *n_keys_out,
                             cudaStream_t st) {
    ctx->ws_k = k; ctx->ws_w = w; ctx->ib_n = 0;
    if (n_picks == 0) {
        if (make_resident) return dcn_index_upload_device(ctx, nullptr, 0, k, w, st);
        return DCN_OK;
    }
    CK(ctx->ib_stats.ensure(128));   // (no-op after any build or decode: the buffer is allocated with slack)
    unsigned long long *d_count = reinterpret_cast<unsigned long long *>(ctx->ib_stats.as<uint8_t>() + 64);
    CK(ctx->ib_keys.ensure(n_picks * sizeof(uint64_t)));
    size_t tmp_sort = 0, tmp_sel = 0;
    cub::DoubleBuffer<uint64_t> db(ctx->ib_alt.as<uint64_t>(), ctx->ib_keys.as<uint64_t>());
    CK(cub::DeviceRadixSort::SortKeys(nullptr, tmp_sort, db, (int64_t)n_picks, 0, 64, st));
    CK(cub::DeviceSelect::Unique(nullptr, tmp_sel, (const uint64_t *)nullptr, (uint64_t *)nullptr, (unsigned long long *)nullptr, (int64_t)n_picks, st));
    CK(ctx->ib_tmp.ensure(std::max(tmp_sort, tmp_sel)));
    // five radix passes over bits 24 .. 63, then the few runs of keys that share those bits are sorted in place
    // (sort_prefix_runs_kernel); keys that are not hash-like (long runs) get the full eight passes
    static const bool prefix_sort = []() { const char *e = getenv("DCN_PREFIX_SORT"); return !e || atoi(e) != 0; }();
    bool full = !prefix_sort;
    if (!full) {
        uint32_t *d_flag = reinterpret_cast<uint32_t *>(ctx->ib_stats.as<uint8_t>() + 96), *d_nruns = d_flag + 1;
        const uint32_t runs_cap = 1u << 20;
        CK(ctx->ib_runs.ensure((size_t)runs_cap * 8));
        CK(cudaMemsetAsync(d_flag, 0, 8, st));
        CK(cub::DeviceRadixSort::SortKeys(ctx->ib_tmp.p, tmp_sort, db, (int64_t)n_picks, 24, 64, st));
        find_prefix_runs_kernel<<<grid_for(ctx, n_picks, 256), 256, 0, st>>>(db.Current(), n_picks, ctx->ib_runs.as<uint64_t>(), runs_cap, d_nruns, d_flag);
        sort_prefix_runs_kernel<<<grid_for(ctx, runs_cap, 128), 128, 0, st>>>(db.Current(), n_picks, ctx->ib_runs.as<uint64_t>(), d_nruns, runs_cap, d_flag);
        uint32_t flag = 0;
        CK(cudaMemcpyAsync(&flag, d_flag, 4, cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        ctx->launches += 7;
        full = flag != 0;
    }
    if (full) CK(cub::DeviceRadixSort::SortKeys(ctx->ib_tmp.p, tmp_sort, db, (int64_t)n_picks, 0, 64, st));
    uint64_t *sorted = db.Current();
    uint64_t *uniq = db.Alternate();
    CK(cub::DeviceSelect::Unique(ctx->ib_tmp.p, tmp_sel, sorted, uniq, d_count, (int64_t)n_picks, st));
    ctx->launches += 8;
    unsigned long long n_unique = 0;
    CK(cudaMemcpyAsync(&n_unique, d_count, sizeof(n_unique), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    if (uniq != ctx->ib_keys.as<uint64_t>()) {  // keep the result in ib_keys
        std::swap(ctx->ib_keys, ctx->ib_alt);
    }
    ctx->ib_n = n_unique;
    if (n_keys_out) *n_keys_out = n_unique;
    if (make_resident) return dcn_index_upload_device(ctx, ctx->ib_keys.as<uint64_t>(), n_unique, k, w, st);
    return DCN_OK;
}

// Index-flavour extraction of every record into ctx->ib_alt (unordered, duplicates included).
static int index_extract_device(dcn_ctx *ctx, const uint8_t *d_bases, const uint64_t *d_rec_off, uint32_t n_rec,
                                uint64_t n_bases, uint8_t k, uint8_t w, float entropy_thr, cudaStream_t st, uint64_t *n_out) {
    *n_out = 0;
    CK(ctx->ib_stats.ensure(64 + 16));
    if (k != 31 || w != 15) {   // generic extraction (ordered CSR, offsets unused)
        uint64_t m = 0;
        bool written = false;
        int rc0 = generic_extract_device(ctx, DCN_FLAVOUR_INDEX, d_bases, d_rec_off, n_rec, k, w, 0, entropy_thr, false, 0, st, &m, &written);
        if (rc0) return rc0;
        if (m) {
            CK(ctx->ib_alt.ensure(m * sizeof(uint64_t)));
            CK(cudaMemcpyAsync(ctx->ib_alt.p, ctx->gx_h.p, m * sizeof(uint64_t), cudaMemcpyDeviceToDevice, st));
        }
        *n_out = m;
        return DCN_OK;
    }
    if (reinterpret_cast<uintptr_t>(d_bases) & 15u) return ctx->fail(DCN_ERR_ARG, "d_bases must be 16-byte aligned");
    const uint32_t *d_entropy = nullptr;
    if (entropy_thr != 0.0f) {
        std::vector<uint32_t> bits;
        build_entropy_bitmap(k, entropy_thr, 32, bits);
        CK(ctx->ib_entropy.ensure(bits.size() * 4));
        CK(cudaMemcpyAsync(ctx->ib_entropy.p, bits.data(), bits.size() * 4, cudaMemcpyHostToDevice, st));
        CK(cudaStreamSynchronize(st));
        d_entropy = ctx->ib_entropy.as<uint32_t>();
    }
    const uint32_t desc_cap = (uint32_t)(n_bases / ChunkGeo<G31>::CSTRIDE + n_rec + 16);
    CK(ctx->ib_desc.ensure((size_t)desc_cap * sizeof(ChunkDesc)));
    BatchStats *d_stats = ctx->ib_stats.as<BatchStats>();
    unsigned long long *d_count = reinterpret_cast<unsigned long long *>(ctx->ib_stats.as<uint8_t>() + 64);

    // typical density is 0.1255 picks per base; low-complexity sequence can reach 1 per window
    uint64_t cap = (uint64_t)((double)n_bases * 0.16) + 4096;
    unsigned long long n_picks = 0;
    for (int attempt = 0; attempt < 2; attempt++) {
        CK(ctx->ib_alt.ensure(cap * sizeof(uint64_t)));
        CK(cudaMemsetAsync(ctx->ib_stats.p, 0, 64 + 16, st));
        const int pb = 256;
        const int pg = (int)std::max<uint64_t>(1, std::min<uint64_t>(((uint64_t)n_rec * 32 + pb - 1) / pb, (uint64_t)ctx->sm_count * 8));   // a warp per record
        prep_index_chunks_kernel<G31><<<pg, pb, 0, st>>>(d_rec_off, n_rec, d_stats, reinterpret_cast<ChunkDesc *>(ctx->ib_desc.p), desc_cap);
        IndexParams P;
        P.bases = d_bases; P.base0 = 0; P.n_bases = n_bases; P.rec_off = d_rec_off; P.n_rec = n_rec;
        P.entropy_pass = d_entropy; P.out = ctx->ib_alt.as<uint64_t>(); P.out_cap = cap; P.out_count = d_count;
        extract_index_kernel<G31><<<ctx->sm_count * (1024 / G31::NT), G31::NT, sizeof(TileSmem<G31>), st>>>(P, d_stats, reinterpret_cast<ChunkDesc *>(ctx->ib_desc.p));
        ctx->launches += 2;
        CK(cudaMemcpyAsync(&n_picks, d_count, sizeof(n_picks), cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        CK(cudaGetLastError());
        if (n_picks <= cap) break;
        if (attempt == 1) return ctx->fail(DCN_ERR_OVERFLOW, "minimizer buffer overflowed twice");
        cap = n_picks + 1024;
    }
    *n_out = n_picks;
    return DCN_OK;
}

int dcn_index_build_device(dcn_ctx *ctx, const uint8_t *d_bases, const uint64_t *d_rec_off, uint32_t n_rec,
                           uint64_t n_bases, uint8_t k, uint8_t w, float entropy_thr, int make_resident,
                           uint64_t *n_keys_out, void *stream) {
    if (!ctx) return DCN_ERR_ARG;
    int rc0 = check_kw(ctx, k, w, DCN_FLAVOUR_INDEX);
    if (rc0) return rc0;
    CK(cudaSetDevice(ctx->device));
    cudaStream_t st = (cudaStream_t)stream;
    ctx->ib_n = 0;
    if (n_keys_out) *n_keys_out = 0;
    uint64_t n_picks = 0;
    if ((rc0 = index_extract_device(ctx, d_bases, d_rec_off, n_rec, n_bases, k, w, entropy_thr, st, &n_picks))) return rc0;
    return index_sort_unique(ctx, n_picks, k, w, make_resident, n_keys_out, st);
}

int dcn_index_build(dcn_ctx *ctx, const uint8_t *bases, const uint64_t *rec_off, uint32_t n_rec, uint8_t k, uint8_t w,
                    float entropy_thr, int make_resident, uint64_t *n_keys_out) {
    if (!ctx) return DCN_ERR_ARG;
    if (!rec_off || (!bases && n_rec && rec_off[n_rec] > 0)) return ctx->fail(DCN_ERR_ARG, "null input pointer");
    CK(cudaSetDevice(ctx->device));
    const uint64_t n_bases = n_rec ? rec_off[n_rec] : 0;
    if (n_rec && rec_off[0] != 0) return ctx->fail(DCN_ERR_ARG, "rec_off[0] must be 0");
    CK(ctx->ib_bases.ensure(n_bases + 64));
    CK(ctx->ib_off.ensure((size_t)(n_rec + 1) * 8));
    if (n_bases) CK(cudaMemcpyAsync(ctx->ib_bases.p, bases, n_bases, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(ctx->ib_off.p, rec_off, (size_t)(n_rec + 1) * 8, cudaMemcpyHostToDevice, ctx->stream));
    return dcn_index_build_device(ctx, ctx->ib_bases.as<uint8_t>(), ctx->ib_off.as<uint64_t>(), n_rec, n_bases, k, w,
                                  entropy_thr, make_resident, n_keys_out, ctx->stream);
}

int dcn_index_build_keys(dcn_ctx *ctx, uint64_t *out_keys, uint64_t cap) {
    if (!ctx) return DCN_ERR_ARG;
    if (cap < ctx->ib_n) return ctx->fail(DCN_ERR_ARG, "output buffer smaller than the key count");
    CK(cudaSetDevice(ctx->device));
    if (ctx->ib_n) CK(cudaMemcpy(out_keys, ctx->ib_keys.p, ctx->ib_n * sizeof(uint64_t), cudaMemcpyDeviceToHost));
    return DCN_OK;
}
const uint64_t *dcn_index_build_keys_device(dcn_ctx *ctx) { return ctx && ctx->ib_n ? ctx->ib_keys.as<uint64_t>() : nullptr; }

// ---------------------------------------------------------------------------- .idx codec + set algebra on the GPU
// bincode-2 varint (src/index.rs:57-72 via bincode config::standard): < 251 one byte; 0xFB + u16; 0xFC + u32; 0xFD + u64
static bool host_read_varint(const uint8_t *p, uint64_t len, uint64_t &pos, uint64_t &v) {
    if (pos >= len) return false;
    const uint8_t t = p[pos++];
    if (t < 251) { v = t; return true; }
    const int n = t == 0xFB ? 2 : t == 0xFC ? 4 : t == 0xFD ? 8 : 0;   // 0xFE (u128) cannot hold a u64 field
    if (!n || pos + (uint64_t)n > len) return false;
    v = 0;
    for (int i = 0; i < n; i++) v |= (uint64_t)p[pos + i] << (8 * i);
    pos += (uint64_t)n;
    return true;
}

namespace {
struct NotInTable {   // predicate of the set difference: key absent from the scratch table
    TableView tv;
    __device__ bool operator()(const uint64_t &key) const { return !table_contains(tv, key); }
};

// body of a .idx file whose every key is a 9-byte token (0xFD + 8 bytes LE): token i starts at byte 9 i.
// bad[0] is set when a tag is not 0xFD (the caller then falls back to the sequential scan).
__global__ void idx_decode9_kernel(const uint8_t *__restrict__ body, uint64_t n, uint64_t *__restrict__ keys, uint32_t *bad) {
    const uint64_t *w = reinterpret_cast<const uint64_t *>(body);   // 8-byte aligned, padded by 16 bytes
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t at = 9 * i;
        if (body[at] != 0xFDu) *bad = 1;
        const uint64_t a = at + 1, q = a >> 3;
        const uint32_t sh = (uint32_t)(a & 7u) * 8u;
        const uint64_t lo = w[q], hi = w[q + 1];
        keys[i] = sh ? (lo >> sh) | (hi << (64u - sh)) : lo;
    }
}

__device__ __forceinline__ uint32_t varint_len(uint64_t v) { return v < 251 ? 1u : v < (1ull << 16) ? 3u : v < (1ull << 32) ? 5u : 9u; }

__global__ void idx_token_len_kernel(const uint64_t *__restrict__ keys, uint64_t n, uint64_t *__restrict__ len) {
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i <= n; i += (uint64_t)gridDim.x * blockDim.x)
        len[i] = i < n ? varint_len(keys[i]) : 0;
}
__global__ void idx_count_short_kernel(const uint64_t *__restrict__ keys, uint64_t n, unsigned long long *n_short) {
    unsigned long long c = 0;
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
        c += keys[i] < (1ull << 32);
    for (int d = 16; d > 0; d >>= 1) c += __shfl_xor_sync(0xFFFFFFFFu, c, d);
    if ((threadIdx.x & 31) == 0 && c) atomicAdd(n_short, c);
}
// off == nullptr: every token is 9 bytes
__global__ void idx_encode_kernel(const uint64_t *__restrict__ keys, uint64_t n, const uint64_t *__restrict__ off, uint8_t *__restrict__ out) {
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t v = keys[i];
        uint8_t *p = out + (off ? off[i] : 9 * i);
        const uint32_t l = off ? varint_len(v) : 9u;
        if (l == 1) { p[0] = (uint8_t)v; continue; }
        p[0] = l == 3 ? 0xFBu : l == 5 ? 0xFCu : 0xFDu;
        for (uint32_t b = 0; b + 1 < l; b++) p[1 + b] = (uint8_t)(v >> (8 * b));
    }
}
}  // namespace

// keys of a decoded .idx body -> device buffer `dst` (n entries)
static int idx_body_to_device(dcn_ctx *ctx, const uint8_t *body, uint64_t body_len, uint64_t n, DevBuf &dst, cudaStream_t st) {
    CK(dst.ensure(std::max<uint64_t>(n, 1) * 8));
    if (n == 0) return DCN_OK;
    if (body_len == 9 * n) {   // the usual layout: xxh3 values below 2^32 are a 2^-32 event
        CK(ctx->gx_bases.ensure(body_len + 32));
        CK(ctx->ib_stats.ensure(64 + 16));
        uint32_t *d_bad = reinterpret_cast<uint32_t *>(ctx->ib_stats.as<uint8_t>() + 72);
        CK(cudaMemsetAsync(d_bad, 0, 4, st));
        CK(cudaMemsetAsync(ctx->gx_bases.as<uint8_t>() + (body_len & ~7ull), 0, 24, st));   // the last word pair is read whole
        CK(cudaMemcpyAsync(ctx->gx_bases.p, body, body_len, cudaMemcpyHostToDevice, st));
        idx_decode9_kernel<<<grid_for(ctx, n, 256), 256, 0, st>>>(ctx->gx_bases.as<uint8_t>(), n, dst.as<uint64_t>(), d_bad);
        ctx->launches += 1;
        uint32_t bad = 0;
        CK(cudaMemcpyAsync(&bad, d_bad, 4, cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        CK(cudaGetLastError());
        if (!bad) return DCN_OK;
    }
    // short tokens present: sequential scan of the stream on the host (the codec is host-side in the reference too)
    std::vector<uint64_t> keys(n);
    uint64_t pos = 0;
    for (uint64_t i = 0; i < n; i++)
        if (!host_read_varint(body, body_len, pos, keys[i])) return ctx->fail(DCN_ERR_ARG, "Failed to deserialise hash: truncated or malformed index body");
    CK(cudaMemcpyAsync(dst.p, keys.data(), n * 8, cudaMemcpyHostToDevice, st));
    CK(cudaStreamSynchronize(st));
    return DCN_OK;
}

// working set := working set \ B, B = n_b keys (any order, duplicates allowed) in device memory
static int working_set_subtract(dcn_ctx *ctx, const uint64_t *d_b, uint64_t n_b, cudaStream_t st) {
    if (ctx->ib_n == 0 || n_b == 0) return DCN_OK;
    const uint64_t nb = table_buckets_for(n_b, 0.5);
    CK(ctx->ws_table.ensure(nb * 4 * sizeof(uint64_t)));
    CK(ctx->ws_flags.ensure(64));
    unsigned long long *cnt = ctx->ws_flags.as<unsigned long long>();
    CK(cudaMemsetAsync(cnt, 0, 64, st));
    table_fill_kernel<<<ctx->sm_count * 8, 256, 0, st>>>(ctx->ws_table.as<uint64_t>(), nb * 4);
    table_insert_kernel<<<grid_for(ctx, n_b, 256), 256, 0, st>>>(ctx->ws_table.as<uint64_t>(), nb, d_b, n_b, cnt);
    unsigned long long c[2] = {0, 0};
    CK(cudaMemcpyAsync(c, cnt, sizeof(c), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    NotInTable pred;
    pred.tv.slots = ctx->ws_table.as<uint64_t>(); pred.tv.n_buckets = nb; pred.tv.has_empty_key = c[1] != 0;
    CK(ctx->ib_alt.ensure(ctx->ib_n * 8));
    size_t tb = 0;
    unsigned long long *d_n = cnt + 4;
    CK(cub::DeviceSelect::If(nullptr, tb, ctx->ib_keys.as<uint64_t>(), ctx->ib_alt.as<uint64_t>(), d_n, (int64_t)ctx->ib_n, pred, st));
    CK(ctx->ib_tmp.ensure(tb));
    CK(cub::DeviceSelect::If(ctx->ib_tmp.p, tb, ctx->ib_keys.as<uint64_t>(), ctx->ib_alt.as<uint64_t>(), d_n, (int64_t)ctx->ib_n, pred, st));
    ctx->launches += 4;
    unsigned long long kept = 0;
    CK(cudaMemcpyAsync(&kept, d_n, 8, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    CK(cudaGetLastError());
    std::swap(ctx->ib_keys, ctx->ib_alt);   // a stable selection of a sorted set stays sorted
    ctx->ib_n = kept;
    return DCN_OK;
}

int dcn_idx_decode(dcn_ctx *ctx, const uint8_t *file, uint64_t len, int mode, int make_resident, uint8_t *version,
                   uint8_t *k, uint8_t *w, uint64_t *n_in_file, uint64_t *n_set) {
    if (!ctx) return DCN_ERR_ARG;
    if (!file || len < 3) return ctx->fail(DCN_ERR_ARG, "Failed to deserialise index header");
    if (mode < DCN_SET_REPLACE || mode > DCN_SET_SUBTRACT) return ctx->fail(DCN_ERR_ARG, "unknown mode");
    if (version) *version = file[0];
    if (k) *k = file[1];
    if (w) *w = file[2];
    if (file[0] != 2) return ctx->fail(DCN_ERR_ARG, "Unsupported index format version (src/index.rs:34-40): only version 2");
    uint64_t pos = 3, count = 0;
    if (!host_read_varint(file, len, pos, count)) return ctx->fail(DCN_ERR_ARG, "Failed to deserialise minimizer count");
    if (n_in_file) *n_in_file = count;
    if (count > len) return ctx->fail(DCN_ERR_ARG, "minimizer count exceeds the file size");
    if (mode != DCN_SET_REPLACE && (file[1] != ctx->ws_k || file[2] != ctx->ws_w))
        return ctx->fail(DCN_ERR_ARG, "Incompatible headers: k, w differ from the first index (src/index.rs:474-485, 626-640)");
    CK(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    int rc;
    if (mode == DCN_SET_REPLACE) {
        if ((rc = idx_body_to_device(ctx, file + pos, len - pos, count, ctx->ib_alt, st))) return rc;
        rc = index_sort_unique(ctx, count, file[1], file[2], 0, nullptr, st);
    } else if (mode == DCN_SET_UNION) {
        if ((rc = idx_body_to_device(ctx, file + pos, len - pos, count, ctx->gx_h, st))) return rc;
        const uint64_t tot = ctx->ib_n + count;
        CK(ctx->ib_alt.ensure(std::max<uint64_t>(tot, 1) * 8));
        if (ctx->ib_n) CK(cudaMemcpyAsync(ctx->ib_alt.p, ctx->ib_keys.p, ctx->ib_n * 8, cudaMemcpyDeviceToDevice, st));
        if (count) CK(cudaMemcpyAsync(ctx->ib_alt.as<uint64_t>() + ctx->ib_n, ctx->gx_h.p, count * 8, cudaMemcpyDeviceToDevice, st));
        rc = index_sort_unique(ctx, tot, ctx->ws_k, ctx->ws_w, 0, nullptr, st);
    } else {
        if ((rc = idx_body_to_device(ctx, file + pos, len - pos, count, ctx->gx_h, st))) return rc;
        rc = working_set_subtract(ctx, ctx->gx_h.as<uint64_t>(), count, st);
    }
    if (rc) return rc;
    if (n_set) *n_set = ctx->ib_n;
    if (make_resident) {
        if ((rc = dcn_index_upload_device(ctx, ctx->ib_keys.as<uint64_t>(), ctx->ib_n, ctx->ws_k, ctx->ws_w, st))) return rc;
        // a load for filtering: the decode / sort scratch (file bytes, two key-sized buffers: ~25 B per key beside the
        // 16 B per key of the table) is dead weight from here on; the sorted key set itself stays for `index info`,
        // union / diff and dcn_idx_encode until the caller releases it (dcn_working_set_release)
        CK(cudaStreamSynchronize(st));
        ctx->gx_bases.release(); ctx->ib_alt.release(); ctx->ib_tmp.release(); ctx->gx_h.release();
    }
    return DCN_OK;
}

int dcn_working_set_release(dcn_ctx *ctx) {
    if (!ctx) return DCN_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    CK(cudaDeviceSynchronize());
    ctx->ib_keys.release(); ctx->ib_alt.release(); ctx->ib_tmp.release(); ctx->ib_bases.release(); ctx->ib_off.release();
    ctx->ib_desc.release(); ctx->gx_bases.release(); ctx->gx_h.release(); ctx->ws_table.release(); ctx->ws_flags.release();
    ctx->ib_n = 0;
    return DCN_OK;
}

int dcn_index_diff_sequences(dcn_ctx *ctx, const uint8_t *bases, const uint64_t *rec_off, uint32_t n_rec, uint64_t *n_set) {
    if (!ctx) return DCN_ERR_ARG;
    if (!rec_off || (!bases && n_rec && rec_off[n_rec] > 0)) return ctx->fail(DCN_ERR_ARG, "null input pointer");
    if (!ctx->ws_k) return ctx->fail(DCN_ERR_NO_INDEX, "no working key set: build or decode an index first");
    if (n_rec && rec_off[0] != 0) return ctx->fail(DCN_ERR_ARG, "rec_off[0] must be 0");
    CK(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    const uint64_t n_bases = n_rec ? rec_off[n_rec] : 0;
    CK(ctx->ib_bases.ensure(n_bases + 64));
    CK(ctx->ib_off.ensure(((size_t)n_rec + 1) * 8));
    if (n_bases) CK(cudaMemcpyAsync(ctx->ib_bases.p, bases, n_bases, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(ctx->ib_off.p, rec_off, ((size_t)n_rec + 1) * 8, cudaMemcpyHostToDevice, st));
    // the working set lives in ib_keys; extraction writes ib_alt, which the subtraction then reuses as its output:
    // move the extracted hashes out of the way first
    uint64_t m = 0;
    int rc = index_extract_device(ctx, ctx->ib_bases.as<uint8_t>(), ctx->ib_off.as<uint64_t>(), n_rec, n_bases, ctx->ws_k, ctx->ws_w,
                                  0.0f, st, &m);   // entropy 0.0: src/index.rs:378-381
    if (rc) return rc;
    CK(ctx->gx_h.ensure(std::max<uint64_t>(m, 1) * 8));
    if (m) CK(cudaMemcpyAsync(ctx->gx_h.p, ctx->ib_alt.p, m * 8, cudaMemcpyDeviceToDevice, st));
    if ((rc = working_set_subtract(ctx, ctx->gx_h.as<uint64_t>(), m, st))) return rc;
    if (n_set) *n_set = ctx->ib_n;
    return DCN_OK;
}

int dcn_idx_encode(dcn_ctx *ctx, uint8_t *out, uint64_t cap, uint64_t *len) {
    if (!ctx || !len) return DCN_ERR_ARG;
    if (!ctx->ws_k) return ctx->fail(DCN_ERR_NO_INDEX, "no working key set: build or decode an index first");
    CK(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    const uint64_t n = ctx->ib_n;
    uint8_t head[12] = {2, ctx->ws_k, ctx->ws_w};
    uint64_t hl = 3;
    if (n < 251) head[hl++] = (uint8_t)n;
    else {
        const int nb = n < (1ull << 16) ? 2 : n < (1ull << 32) ? 4 : 8;
        head[hl++] = nb == 2 ? 0xFB : nb == 4 ? 0xFC : 0xFD;
        for (int i = 0; i < nb; i++) head[hl++] = (uint8_t)(n >> (8 * i));
    }
    uint64_t body = 9 * n;
    const uint64_t *d_off = nullptr;
    if (n) {
        CK(ctx->ws_flags.ensure(64));
        unsigned long long *d_short = ctx->ws_flags.as<unsigned long long>();
        CK(cudaMemsetAsync(d_short, 0, 8, st));
        idx_count_short_kernel<<<grid_for(ctx, n, 256), 256, 0, st>>>(ctx->ib_keys.as<uint64_t>(), n, d_short);
        unsigned long long n_short = 0;
        CK(cudaMemcpyAsync(&n_short, d_short, 8, cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        if (n_short) {   // variable token lengths: lengths -> exclusive scan -> offsets
            CK(ctx->gx_cc.ensure((n + 1) * 8));
            uint64_t *l = ctx->gx_cc.as<uint64_t>();
            idx_token_len_kernel<<<grid_for(ctx, n + 1, 256), 256, 0, st>>>(ctx->ib_keys.as<uint64_t>(), n, l);
            size_t tb = 0;
            CK(cub::DeviceScan::ExclusiveSum(nullptr, tb, l, l, (int64_t)n + 1, st));
            CK(ctx->gx_tmp.ensure(tb));
            CK(cub::DeviceScan::ExclusiveSum(ctx->gx_tmp.p, tb, l, l, (int64_t)n + 1, st));
            CK(cudaMemcpyAsync(&body, l + n, 8, cudaMemcpyDeviceToHost, st));
            CK(cudaStreamSynchronize(st));
            d_off = l;
            ctx->launches += 3;
        }
    }
    *len = hl + body;
    if (!out || cap < hl + body) return ctx->fail(DCN_ERR_OVERFLOW, "output buffer too small: *len holds the required size");
    memcpy(out, head, hl);
    if (n) {
        CK(ctx->gx_bases.ensure(body + 16));
        idx_encode_kernel<<<grid_for(ctx, n, 256), 256, 0, st>>>(ctx->ib_keys.as<uint64_t>(), n, d_off, ctx->gx_bases.as<uint8_t>());
        ctx->launches += 2;
        CK(cudaMemcpyAsync(out + hl, ctx->gx_bases.p, body, cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        CK(cudaGetLastError());
    }
    return DCN_OK;
}

int dcn_working_set_info(dcn_ctx *ctx, uint64_t *n_keys, uint8_t *k, uint8_t *w) {
    if (!ctx) return DCN_ERR_ARG;
    if (!ctx->ws_k) return ctx->fail(DCN_ERR_NO_INDEX, "no working key set: build or decode an index first");
    if (n_keys) *n_keys = ctx->ib_n;
    if (k) *k = ctx->ws_k;
    if (w) *w = ctx->ws_w;
    return DCN_OK;
}

int dcn_index_make_resident(dcn_ctx *ctx) {
    if (!ctx) return DCN_ERR_ARG;
    if (!ctx->ws_k) return ctx->fail(DCN_ERR_NO_INDEX, "no working key set: build or decode an index first");
    CK(cudaSetDevice(ctx->device));
    return dcn_index_upload_device(ctx, ctx->ib_keys.as<uint64_t>(), ctx->ib_n, ctx->ws_k, ctx->ws_w, ctx->stream);
}

// ---------------------------------------------------------------------------- counters
int dcn_stats_get(dcn_ctx *ctx, uint64_t counters[6]) {
    if (!ctx || !counters) return DCN_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(counters, ctx->counters.p, 6 * sizeof(uint64_t), cudaMemcpyDeviceToHost));
    return check_promise(ctx);   // everything enqueued has run: a promise broken by any earlier call shows here
}
int dcn_stats_accumulate_device(dcn_ctx *ctx, const uint64_t *d_rec_off, uint32_t n_rec, int paired, const uint8_t *d_keep,
                                void *stream) {
    if (!ctx) return DCN_ERR_ARG;
    if (!d_rec_off || !d_keep) return ctx->fail(DCN_ERR_ARG, "null pointer");
    const uint32_t rpu = paired ? 2u : 1u;
    if (paired && (n_rec & 1u)) return ctx->fail(DCN_ERR_ARG, "paired batch needs an even record count");
    const uint32_t n_units = n_rec / rpu;
    if (!n_units) return DCN_OK;
    CK(cudaSetDevice(ctx->device));
    stats_kernel<<<grid_for(ctx, n_units, 256), 256, 0, (cudaStream_t)stream>>>(d_rec_off, rpu, n_units, d_keep,
                                                                               ctx->counters.as<unsigned long long>());
    ctx->launches += 1;
    CK(cudaGetLastError());
    return DCN_OK;
}

int dcn_stats_reset(dcn_ctx *ctx) {
    if (!ctx) return DCN_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    CK(cudaDeviceSynchronize());
    CK(cudaMemset(ctx->counters.p, 0, 6 * sizeof(uint64_t)));
    return DCN_OK;
}

// ---------------------------------------------------------------------------- measurement
int dcn_last_timing(dcn_ctx *ctx, float *h2d_ms, float *kernel_ms, float *d2h_ms) {
    if (!ctx) return DCN_ERR_ARG;
    if (h2d_ms) *h2d_ms = ctx->t_h2d;
    if (kernel_ms) *kernel_ms = ctx->t_kernel;
    if (d2h_ms) *d2h_ms = ctx->t_d2h;
    return DCN_OK;
}

int dcn_last_transfer_bytes(dcn_ctx *ctx, uint64_t *h2d_bytes, uint64_t *d2h_bytes) {
    if (!ctx) return DCN_ERR_ARG;
    if (h2d_bytes) *h2d_bytes = ctx->bytes_h2d;
    if (d2h_bytes) *d2h_bytes = ctx->bytes_d2h;
    return DCN_OK;
}

int dcn_measure_random_access_wide(dcn_ctx *ctx, int sectors, uint64_t *n_probes, float *ms) {
    if (!ctx || !ms || !n_probes) return DCN_ERR_ARG;
    if (sectors != 1 && sectors != 2 && sectors != 4 && sectors != -4) return ctx->fail(DCN_ERR_ARG, "sectors per probe: 1, 2, 4, or -4 (four lanes share a 128-byte line)");
    if (!ctx->table.p) return ctx->fail(DCN_ERR_NO_INDEX, "no index resident");
    CK(cudaSetDevice(ctx->device));
    const int threads = 256, grid = ctx->sm_count * 8;
    uint64_t per_thread = *n_probes / ((uint64_t)threads * grid);
    per_thread = std::max<uint64_t>(4, (per_thread + 3) / 4 * 4);
    *n_probes = per_thread * threads * grid / (sectors == -4 ? 4 : 1);
    cudaEvent_t a, b;
    CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    CK(cudaEventRecord(a, ctx->stream));
    unsigned long long *sink = ctx->counters.as<unsigned long long>() + 7;
    if (sectors == -4) random_access_line_kernel<<<grid, threads, 0, ctx->stream>>>(ctx->table.as<uint64_t>(), ctx->n_buckets, (uint32_t)per_thread, sink);
    else if (sectors == 1) random_access_kernel<1><<<grid, threads, 0, ctx->stream>>>(ctx->table.as<uint64_t>(), ctx->n_buckets, (uint32_t)per_thread, sink);
    else if (sectors == 2) random_access_kernel<2><<<grid, threads, 0, ctx->stream>>>(ctx->table.as<uint64_t>(), ctx->n_buckets, (uint32_t)per_thread, sink);
    else random_access_kernel<4><<<grid, threads, 0, ctx->stream>>>(ctx->table.as<uint64_t>(), ctx->n_buckets, (uint32_t)per_thread, sink);
    CK(cudaEventRecord(b, ctx->stream));
    CK(cudaEventSynchronize(b));
    CK(cudaEventElapsedTime(ms, a, b));
    cudaEventDestroy(a); cudaEventDestroy(b);
    ctx->launches += 1;
    return DCN_OK;
}

int dcn_measure_random_access(dcn_ctx *ctx, uint64_t *n_probes, float *ms) { return dcn_measure_random_access_wide(ctx, 1, n_probes, ms); }

int dcn_last_pack_ms(dcn_ctx *ctx, float *pack_ms) {
    if (!ctx || !pack_ms) return DCN_ERR_ARG;
    *pack_ms = ctx->t_pack;
    return DCN_OK;
}

uint64_t dcn_launch_count(dcn_ctx *ctx) { return ctx ? ctx->launches.load() : 0; }

int dcn_fused_time_take(dcn_ctx *ctx, float *total_ms, uint32_t *n_launches) {
    if (!ctx || !total_ms || !n_launches) return DCN_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    float sum = 0;
    for (uint32_t i = 0; i < ctx->kev_count; i++) {
        uint32_t e = (ctx->kev_head - 1 - i) % dcn_ctx::KEV;
        CK(cudaEventSynchronize(ctx->kev1[e]));
        float ms = 0;
        CK(cudaEventElapsedTime(&ms, ctx->kev0[e], ctx->kev1[e]));
        sum += ms;
    }
    *total_ms = sum;
    *n_launches = ctx->kev_count;
    ctx->kev_count = 0;
    return DCN_OK;
}

}  // extern "C"
