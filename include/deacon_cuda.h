/*
 * deacon_cuda.h -- C ABI of the B200-native filter hot path of Deacon (deacon-server fork).
 *
 * This is the drop-in boundary: plain pointers and sizes, no CUDA/torch types.  Every entry point
 * cites the reference interface (file:line under /root/reference) it replaces; INTEGRATION.md shows
 * the Rust `extern "C"` block and the call sites a maintainer would change.
 *
 * Conventions
 *   - Return value: 0 = ok, negative = error (DCN_ERR_*); dcn_last_error(ctx) gives the text.
 *     Nothing here aborts; there is NO CPU fallback -- without a CUDA device dcn_ctx_create fails.
 *   - The caller owns every host buffer (inputs and pre-sized outputs); the library owns device
 *     memory inside dcn_ctx.  A dcn_ctx is bound to one GPU; calls on one ctx are serialised by the
 *     caller (the reference's server holds a mutex for a whole request, src/server.rs:121,145);
 *     different ctxs are independent and may be driven from different host threads.
 *   - Records are passed the way the reference's batch engine holds them (src/remote_filter.rs:727-755):
 *     one concatenated byte buffer `bases` + `rec_off[n_rec + 1]` byte offsets.  With `paired != 0`
 *     records 2i and 2i+1 are the mates of pair i and per-unit outputs have n_rec / 2 entries.
 *   - `*_device` variants take device pointers (inputs already in HBM, outputs left in HBM) and a
 *     cudaStream_t passed as void* (NULL = the legacy default stream); work is enqueued on that
 *     stream.  A batch that may contain a unit longer than 1024 bases costs one small stream sync.
 */
#ifndef DEACON_CUDA_H
#define DEACON_CUDA_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct dcn_ctx dcn_ctx;

enum {
    DCN_OK = 0,
    DCN_ERR_CUDA = -1,        /* a CUDA runtime call failed (text in dcn_last_error) */
    DCN_ERR_ARG = -2,         /* invalid argument */
    DCN_ERR_NO_INDEX = -3,    /* no index resident in this ctx */
    DCN_ERR_UNSUPPORTED = -4, /* (k, w) outside what the kernels implement */
    DCN_ERR_NOMEM = -5,
    DCN_ERR_OVERFLOW = -6     /* an internal table overflowed even after retries */
};

enum { DCN_FLAVOUR_FILTER = 0, DCN_FLAVOUR_INDEX = 1 };

/* ---- context ------------------------------------------------------------------------------- */
int dcn_device_count(void);
/* Replaces nothing in the reference (it has no device); one ctx per GPU. */
dcn_ctx *dcn_ctx_create(int device);
void dcn_ctx_destroy(dcn_ctx *ctx);
/* ctx may be NULL: returns the error text of the last failed dcn_ctx_create on this thread. */
const char *dcn_last_error(dcn_ctx *ctx);
/* Pinned host memory for callers that want full PCIe rate on the *_batch entry points. */
void *dcn_host_alloc(size_t bytes);
void dcn_host_free(void *p);

/* ---- B4: index residency --------------------------------------------------------------------
 * Replaces load_minimizer_hashes' insert loop (src/index.rs:98-105), the Arc<FxHashSet<u64>> held by
 * FilterProcessor (src/local_filter.rs:156,631) and the server's static INDEX (src/server.rs:17,68-86).
 * `keys` need not be sorted or unique.  k, w are the IndexHeader fields (src/index.rs:17-22). */
int dcn_index_upload(dcn_ctx *ctx, const uint64_t *keys, uint64_t n_keys, uint8_t k, uint8_t w);
int dcn_index_upload_device(dcn_ctx *ctx, const uint64_t *d_keys, uint64_t n_keys, uint8_t k, uint8_t w,
                            void *stream);
/* n_keys = distinct keys resident (what `deacon index info` prints, src/index.rs:311-340). */
int dcn_index_info(dcn_ctx *ctx, uint64_t *n_keys, uint8_t *k, uint8_t *w, uint64_t *table_bytes);
/* Table load factor used by the next upload (default 0.5; 4 keys per 32-byte bucket). */
int dcn_index_set_load_factor(dcn_ctx *ctx, double load);

/* ---- B1: batch classify on raw sequences ------------------------------------------------------
 * Replaces FilterProcessor::should_keep_sequence / should_keep_pair (src/local_filter.rs:221-285) and
 * the par_iter extraction + check pair of the batch engine (src/remote_filter.rs:762-790, 963-986,
 * 1219-1241): per unit (keep, hit_count, total_minimizers).  k, w come from the resident index.
 * abs_thr / rel_thr / deplete / prefix_len: src/filter_common.rs:84-112, 222-226. */
int dcn_filter_batch(dcn_ctx *ctx, const uint8_t *bases, const uint64_t *rec_off, uint32_t n_rec, int paired,
                     uint32_t prefix_len, uint32_t abs_thr, double rel_thr, int deplete,
                     uint8_t *keep, uint32_t *hits, uint32_t *total);
/* d_bases must be 16-byte aligned; n_bases = rec_off[n_rec] (known to the caller, saves a sync). */
int dcn_filter_batch_device(dcn_ctx *ctx, const uint8_t *d_bases, const uint64_t *d_rec_off, uint32_t n_rec,
                            uint64_t n_bases, int paired, uint32_t prefix_len, uint32_t abs_thr, double rel_thr,
                            int deplete, uint8_t *d_keep, uint32_t *d_hits, uint32_t *d_total, void *stream);
/* The same with the caller's promise that no unit (a record, or a pair of mates) is longer than `max_unit_len`
 * bases -- the reference's callers know this from the parser that produced rec_off.  With max_unit_len <= 1024 the
 * call only enqueues work on `stream`: nothing is read back, no stream synchronisation (dcn_filter_batch_device has
 * to learn from the device whether the batch holds long units, which costs one small synchronisation per call).
 * A broken promise is detected on the device; the units that were too long are left unclassified and the NEXT
 * filter / stats call on this ctx returns DCN_ERR_ARG.  max_unit_len = 0 or > 1024: same as dcn_filter_batch_device. */
int dcn_filter_batch_device_hint(dcn_ctx *ctx, const uint8_t *d_bases, const uint64_t *d_rec_off, uint32_t n_rec,
                                 uint64_t n_bases, int paired, uint32_t prefix_len, uint32_t abs_thr, double rel_thr,
                                 int deplete, uint8_t *d_keep, uint32_t *d_hits, uint32_t *d_total, void *stream,
                                 uint32_t max_unit_len);

/* Host ingest of dcn_filter_batch.  For k=31, w=15 indexes part of the batch can be packed on
 * `n_threads` host threads (2-bit codes + non-ACGT bits: what PackedSeqVec::from_ascii and the mask loop of
 * src/filter_common.rs:238-258 compute per record on the CPU) into pinned staging, so 0.4 B/bp cross PCIe instead
 * of 1 B/bp; the rest is copied as ASCII and converted on the GPU.  Results are identical.
 * Default threads: DCN_PACK_THREADS, else the CPUs this process may use (affinity mask, cgroup quota) - 4, at least 1
 * and at most 12; 0 = never pack. */
int dcn_host_pack_threads(dcn_ctx *ctx, int n_threads);
/* Share of the batch the packing route may take.  Negative (default) = automatic.  Pinned caller buffers: the two
 * routes split the batch dynamically -- the copy engine ships ASCII chunks from the front of the batch while the
 * packer threads work from its back, until they meet (batches under 6 chunks of 32 MB go ASCII only).  Pageable
 * buffers: 1 (a direct copy would be staged by the driver at a few GB/s).  A value in [0, 1] caps the packers'
 * share instead (1 = packed only, 0 = ASCII only).  DCN_PACK_FRACTION sets the same from the environment. */
int dcn_host_pack_fraction(dcn_ctx *ctx, double fraction);
/* The packer itself (no GPU needed): codes[i] = bases [16i, 16i+16) as (byte >> 1) & 3, base j at bits 2j;
 * inv[i] bit j = base 16i+j is not one of ACGTacgt.  Both arrays hold 2 * ceil(n_bases / 32) entries;
 * positions past n_bases are padded with code 0 / non-ACGT. */
int dcn_pack_ascii(const uint8_t *bases, uint64_t n_bases, uint32_t *codes, uint16_t *inv);

/* ---- B1 with host-packed input ("host ingest at rate") --------------------------------
 * Same as dcn_filter_batch for callers that already hold the batch in the packed form (a FASTQ parser can
 * emit it while it touches the bytes anyway): `codes` / `inv` as written by dcn_pack_ascii over the whole
 * concatenated buffer (base 0 of the batch = bit 0 of word 0), `nl_bits` as written by dcn_newline_bits (NULL
 * = no record ends in a newline), `rec_off` in bases as before.  0.43 B/bp cross PCIe; k=31, w=15 only. */
int dcn_filter_batch_packed(dcn_ctx *ctx, const uint32_t *codes, const uint16_t *inv, const uint32_t *nl_bits,
                            const uint64_t *rec_off, uint32_t n_rec, int paired, uint32_t prefix_len, uint32_t abs_thr,
                            double rel_thr, int deplete, uint8_t *keep, uint32_t *hits, uint32_t *total);
/* The same with the non-ACGT bits as a sparse list instead of the dense `inv` array (what the library's own packer
 * threads put on the wire): exc[2 i] = index of a 32-base block of the batch (base position / 32), exc[2 i + 1] = the
 * 32 non-ACGT bits of that block; every block that holds a non-ACGT byte (or padding behind the last base) is listed,
 * once, in ascending order (else DCN_ERR_ARG).  0.25 instead of 0.375 bytes per base cross PCIe; the dense array the
 * kernels read is rebuilt on the device.  Replaces the same loop (src/filter_common.rs:238-258). */
int dcn_filter_batch_packed_sparse(dcn_ctx *ctx, const uint32_t *codes, const uint32_t *exc, uint64_t n_exc,
                                   const uint32_t *nl_bits, const uint64_t *rec_off, uint32_t n_rec, int paired,
                                   uint32_t prefix_len, uint32_t abs_thr, double rel_thr, int deplete, uint8_t *keep,
                                   uint32_t *hits, uint32_t *total);
/* bit r of nl_bits[(n_rec + 31) / 32] = record r (raw length >= k) ends its effective prefix in '\n'
 * (the one byte of the ASCII form the packed form cannot tell from other non-ACGT bytes; src/filter_common.rs:229). */
int dcn_newline_bits(const uint8_t *bases, const uint64_t *rec_off, uint32_t n_rec, uint8_t k, uint32_t prefix_len,
                     uint32_t *nl_bits);
/* dcn_pack_ascii over bases [0, rec_off[n_rec]) and dcn_newline_bits in ONE pass over the bytes (no GPU needed): what
 * the library's own packer threads run per chunk, and what a parser that wants dcn_filter_batch_packed should call.
 * The flags cost nothing extra: only records that end inside a 32-base block holding a non-ACGT byte are looked at.
 * codes / inv: 2 * ceil(rec_off[n_rec] / 32) entries each; nl_bits: ceil(n_rec / 32) words. */
int dcn_pack_records(const uint8_t *bases, const uint64_t *rec_off, uint32_t n_rec, uint8_t k, uint32_t prefix_len,
                     uint32_t *codes, uint16_t *inv, uint32_t *nl_bits);
/* dcn_pack_records for dcn_filter_batch_packed_sparse: the non-ACGT bits as (block, mask) pairs in `exc` (2 * exc_cap
 * words).  *n_exc = number of listed blocks; DCN_ERR_OVERFLOW (nothing written to exc) when exc_cap is smaller. */
int dcn_pack_records_sparse(const uint8_t *bases, const uint64_t *rec_off, uint32_t n_rec, uint8_t k, uint32_t prefix_len,
                            uint32_t *codes, uint32_t *exc, uint64_t exc_cap, uint64_t *n_exc, uint32_t *nl_bits);

/* ---- B2: batch classify on pre-hashed records --------------------------------------------------
 * Replaces unpaired_should_keep / paired_should_keep (src/remote_filter.rs:230-301), i.e. the body of
 * the server's POST handlers (src/server.rs:120-164).  A paired request carries one pooled hash list
 * per pair (src/filter_common.rs:312-348), so both forms are "one hash list per record" here. */
int dcn_lookup_batch(dcn_ctx *ctx, const uint64_t *hashes, const uint64_t *rec_off, uint32_t n_rec,
                     uint32_t abs_thr, double rel_thr, int deplete,
                     uint8_t *keep, uint32_t *hits, uint32_t *total);
/* Same, plus hit_flags[i] = 1 where hashes[i] is a counted hit (in the index and first of its value in the record):
 * what `--debug` needs to print the matching k-mers (sequence_matches / pair_matches, src/filter_common.rs:129-198). */
int dcn_lookup_batch_flags(dcn_ctx *ctx, const uint64_t *hashes, const uint64_t *rec_off, uint32_t n_rec,
                           uint32_t abs_thr, double rel_thr, int deplete,
                           uint8_t *keep, uint32_t *hits, uint32_t *total, uint8_t *hit_flags);
int dcn_lookup_batch_device(dcn_ctx *ctx, const uint64_t *d_hashes, const uint64_t *d_rec_off, uint32_t n_rec,
                            uint32_t abs_thr, double rel_thr, int deplete,
                            uint8_t *d_keep, uint32_t *d_hits, uint32_t *d_total, void *stream);

/* ---- B3: extraction only ------------------------------------------------------------------------
 * Replaces get_minimizer_hashes_and_positions (src/filter_common.rs:211, flavour FILTER) and
 * compute_minimizer_hashes / fill_minimizer_hashes (src/minimizers.rs:53,125, flavour INDEX).
 * Output is CSR: out_off[n_rec + 1] offsets into out_hashes / out_pos (positions relative to the
 * record's effective sequence; out_pos may be NULL).  Returns DCN_ERR_OVERFLOW if out_cap is too small;
 * out_off[n_rec] then still holds the required capacity. */
int dcn_extract(dcn_ctx *ctx, int flavour, const uint8_t *bases, const uint64_t *rec_off, uint32_t n_rec,
                uint8_t k, uint8_t w, uint32_t prefix_len, float entropy_thr,
                uint64_t *out_hashes, uint32_t *out_pos, uint64_t *out_off, uint64_t out_cap);

/* Device-pointer form (the GPU-resident client of the batch engine: extraction -> dcn_lookup_batch_device).
 * d_bases 16-byte aligned; *n_out = minimizers found (also when DCN_ERR_OVERFLOW is returned). */
int dcn_extract_device(dcn_ctx *ctx, int flavour, const uint8_t *d_bases, const uint64_t *d_rec_off, uint32_t n_rec,
                       uint64_t n_bases, uint8_t k, uint8_t w, uint32_t prefix_len, float entropy_thr,
                       uint64_t *d_out_hashes, uint32_t *d_out_pos, uint64_t *d_out_off, uint64_t out_cap,
                       uint64_t *n_out, void *stream);

/* ---- index build (config 4) ---------------------------------------------------------------------
 * Replaces index::build's extraction + FxHashSet::extend (src/index.rs:225-284): index-flavour
 * extraction of every record, radix sort + unique on the GPU.  The sorted unique key set stays in the
 * ctx; fetch it with dcn_index_build_keys (for write_minimizers, src/index.rs:130-164) and/or make it
 * the resident index with make_resident != 0. */
int dcn_index_build(dcn_ctx *ctx, const uint8_t *bases, const uint64_t *rec_off, uint32_t n_rec,
                    uint8_t k, uint8_t w, float entropy_thr, int make_resident, uint64_t *n_keys_out);
int dcn_index_build_device(dcn_ctx *ctx, const uint8_t *d_bases, const uint64_t *d_rec_off, uint32_t n_rec,
                           uint64_t n_bases, uint8_t k, uint8_t w, float entropy_thr, int make_resident,
                           uint64_t *n_keys_out, void *stream);
int dcn_index_build_keys(dcn_ctx *ctx, uint64_t *out_keys, uint64_t cap);
/* device pointer to the sorted unique keys of the last build (valid until the next build) */
const uint64_t *dcn_index_build_keys_device(dcn_ctx *ctx);

/* ---- .idx container and set algebra on the working key set ---------------------------
 * A ctx holds one "working key set": the sorted unique keys of the last dcn_index_build, dcn_idx_decode, union
 * or difference, with its (k, w).  dcn_index_build_keys[_device] read it; dcn_idx_encode serialises it.
 *
 * dcn_idx_decode replaces load_minimizer_hashes' decode + insert loop (src/index.rs:80-107): `file` is a whole
 * .idx file (3-byte header, bincode varint count, varint keys).  The usual all-9-byte-token body is decoded by a
 * kernel straight from the uploaded bytes; bodies with short tokens are scanned sequentially on the host.
 *   mode DCN_SET_REPLACE   working set := file            (index info / load)
 *        DCN_SET_UNION     working set := working ∪ file  (index::union, src/index.rs:563-664)
 *        DCN_SET_SUBTRACT  working set := working \ file  (index::diff, index - index, src/index.rs:421-537)
 * Headers must agree for UNION / SUBTRACT ("Incompatible headers", src/index.rs:474-485, 626-640). */
enum { DCN_SET_REPLACE = 0, DCN_SET_UNION = 1, DCN_SET_SUBTRACT = 2 };
int dcn_idx_decode(dcn_ctx *ctx, const uint8_t *file, uint64_t len, int mode, int make_resident,
                   uint8_t *version, uint8_t *k, uint8_t *w, uint64_t *n_in_file, uint64_t *n_set);
/* stream_diff_fastx (src/index.rs:311-419): working set -= minimizers of these records (index flavour,
 * entropy 0, the working set's k and w). */
int dcn_index_diff_sequences(dcn_ctx *ctx, const uint8_t *bases, const uint64_t *rec_off, uint32_t n_rec, uint64_t *n_set);
/* write_minimizers (src/index.rs:130-164): header + varint count + varint keys (ascending).  Returns
 * DCN_ERR_OVERFLOW with *len = required size when cap is too small (out may be NULL to query). */
int dcn_idx_encode(dcn_ctx *ctx, uint8_t *out, uint64_t cap, uint64_t *len);
/* `deacon index info` (src/index.rs:539-560): size and header of the working key set. */
int dcn_working_set_info(dcn_ctx *ctx, uint64_t *n_keys, uint8_t *k, uint8_t *w);
/* Frees the working key set and every build / decode scratch buffer (a ctx that only filters from here on keeps
 * the resident table alone).  dcn_idx_decode(make_resident = 1) already drops its decode and sort scratch itself. */
int dcn_working_set_release(dcn_ctx *ctx);
/* Make the working key set the resident (probed) index. */
int dcn_index_make_resident(dcn_ctx *ctx);

/* ---- a13: the six summary counters (ProcessingStats, src/local_filter.rs:179-187) ----------------
 * counters[6] = {total_seqs, filtered_seqs, total_bp, output_bp, filtered_bp, output_seq_counter},
 * accumulated over every dcn_filter_batch call since the last reset.  These are the values the
 * multi-GPU driver all-reduces over NCCL. */
int dcn_stats_get(dcn_ctx *ctx, uint64_t counters[6]);
int dcn_stats_reset(dcn_ctx *ctx);
/* The client of the batch engine owns the reads and therefore the counters (src/remote_filter.rs:793-835): add the
 * counters of a batch whose decisions came from dcn_lookup_batch_device.  d_rec_off = offsets of the READS. */
int dcn_stats_accumulate_device(dcn_ctx *ctx, const uint64_t *d_rec_off, uint32_t n_rec, int paired, const uint8_t *d_keep,
                                void *stream);

/* ---- measurement helpers ------------------------------------------------------------------------
 * Milliseconds of the last host-pointer call: H2D copies, kernels, D2H copies (CUDA events). */
int dcn_last_timing(dcn_ctx *ctx, float *h2d_ms, float *kernel_ms, float *d2h_ms);
/* Bytes the last dcn_filter_batch / dcn_filter_batch_packed call copied over PCIe in each direction.  Less than
 * the caller's buffers when a chunk's records all have one length: its rec_off is then an arithmetic sequence
 * and is written by a kernel on the device instead of being copied (8 bytes per record saved). */
int dcn_last_transfer_bytes(dcn_ctx *ctx, uint64_t *h2d_bytes, uint64_t *d2h_bytes);
/* Host packing time (ms, wall clock) of the last dcn_filter_batch call; 0 when it shipped ASCII. */
int dcn_last_pack_ms(dcn_ctx *ctx, float *pack_ms);
/* Random 32-byte-sector read ceiling over the resident table: about *n_probes independent loads;
 * on return *n_probes is the exact number issued and *ms the kernel time. */
int dcn_measure_random_access(dcn_ctx *ctx, uint64_t *n_probes, float *ms);
/* The same with `sectors` (1, 2 or 4) consecutive 32-byte sectors read per probe from a sectors * 32-byte aligned
 * bucket: what a 64- or 128-byte (8- / 16-key) bucket layout would pay per probe (bucket-width A/B). */
int dcn_measure_random_access_wide(dcn_ctx *ctx, int sectors, uint64_t *n_probes, float *ms);
/* Number of kernel launches issued by this ctx so far. */
uint64_t dcn_launch_count(dcn_ctx *ctx);
/* Sum of the CUDA-event durations of the fused filter kernel's launches since the last take (at most
 * the last 256), measured on the stream they were launched on; waits for them to finish. */
int dcn_fused_time_take(dcn_ctx *ctx, float *total_ms, uint32_t *n_launches);

#ifdef __cplusplus
}
#endif
#endif
