/*
 * deacon_oracle.h -- CPU restatement of Deacon's filter hot path.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load it.  The product path
 * (deacon_server_b200/csrc, libdeacon_cuda.so) never links or calls anything here.
 *
 * PARITY STATUS: "parity unpinned" for minimizer SELECTION.  The arithmetic that picks
 * minimizer positions lives in third-party crates that are not vendored in /root/reference
 * and cannot be built here (no cargo/rustc): simd-minimizers 1.3.0, packed-seq 3.2.1
 * (Cargo.lock:1954-1961, 1380-1388).  Their published algorithm is restated below
 * (SURVEY.md Appendix A.2-A.4).  It is anchored on (a) every behavioural known-answer
 * test the reference holds for this path (tests/filter_tests.rs; see tests/test_oracle_golden.py),
 * (b) the XXH3-64 specification, checked against the `xxhash` Python module, and
 * (c) an independent pure-Python restatement (oracle/py_oracle.py).
 * The hash (xxh3), the 2-bit packing, the ACGT filter, the classification rule and the
 * .idx codec follow code that IS in the reference tree and are pinned by (a)+(b).
 */
#ifndef DEACON_ORACLE_H
#define DEACON_ORACLE_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* xxhash-rust 0.8.15 xxh3_64(&v.to_le_bytes()), seed 0 (src/filter_common.rs:296,305). */
uint64_t dcno_xxh3_u64(uint64_t v);
uint64_t dcno_xxh3_u128(uint64_t lo, uint64_t hi);

/* Canonical ntHash of the k-mer starting at codes[0] (closed form). */
uint32_t dcno_nthash_closed(const uint8_t *codes, int k);

/* simd_minimizers::canonical_minimizer_positions on 2-bit codes (A=0,C=1,T=2,G=3).
 * out_pos must hold max(0, n-(k+w-1)+1) entries.  Returns number of positions. */
size_t dcno_minimizer_positions(const uint8_t *codes, size_t n, int k, int w, uint32_t *out_pos);
/* brute-force twin (closed-form hash, full rescan); tests compare the two */
size_t dcno_minimizer_positions_brute(const uint8_t *codes, size_t n, int k, int w, uint32_t *out_pos);

/* src/filter_common.rs:211-310 get_minimizer_hashes_and_positions.
 * out_hashes/out_pos must hold `len` entries.  Returns count. */
size_t dcno_extract_filter(const uint8_t *seq, size_t len, size_t prefix_len, int k, int w,
                           uint64_t *out_hashes, uint32_t *out_pos);

/* src/minimizers.rs:125-191 fill_minimizer_hashes (IUPAC map + entropy filter). */
size_t dcno_extract_index(const uint8_t *seq, size_t len, int k, int w, float entropy_thr,
                          uint64_t *out_hashes);

/* src/minimizers.rs:73-121 calculate_scaled_entropy. */
float dcno_scaled_entropy(const uint8_t *kmer, int k);

/* src/filter_common.rs:84-112. */
uint64_t dcno_required_hits(uint64_t abs_thr, double rel_thr, uint64_t total);
int dcno_meets_criteria(uint64_t hits, uint64_t total, uint64_t abs_thr, double rel_thr, int deplete);

/* FxHashSet<u64> stand-in (exact membership). */
typedef struct dcno_set dcno_set;
dcno_set *dcno_set_new(uint64_t expected);
void dcno_set_free(dcno_set *s);
void dcno_set_insert_many(dcno_set *s, const uint64_t *keys, uint64_t n, int threads);
int dcno_set_contains(const dcno_set *s, uint64_t key);
uint64_t dcno_set_len(const dcno_set *s);
/* copies the keys out (unsorted); out must hold dcno_set_len entries */
void dcno_set_keys(const dcno_set *s, uint64_t *out);

/* src/filter_common.rs:129-198 sequence_matches / pair_matches on pre-hashed records
 * + src/remote_filter.rs:230-301 unpaired_/paired_should_keep (batch form).
 * rec_off has n_rec+1 entries into `hashes`. */
void dcno_lookup_batch(const dcno_set *idx, const uint64_t *hashes, const uint64_t *rec_off,
                       uint32_t n_rec, uint64_t abs_thr, double rel_thr, int deplete,
                       uint8_t *keep, uint32_t *hits, uint32_t *total, int threads);

/* src/local_filter.rs:221-285 should_keep_sequence / should_keep_pair over a batch of raw
 * records.  paired!=0: records 2i, 2i+1 are mates and outputs have n_rec/2 entries. */
void dcno_filter_batch(const dcno_set *idx, const uint8_t *bases, const uint64_t *rec_off,
                       uint32_t n_rec, int paired, uint64_t prefix_len, int k, int w,
                       uint64_t abs_thr, double rel_thr, int deplete,
                       uint8_t *keep, uint32_t *hits, uint32_t *total, int threads);

/* src/index.rs:167-308 build: union of compute_minimizer_hashes over records. */
void dcno_index_build(dcno_set *dst, const uint8_t *bases, const uint64_t *rec_off,
                      uint32_t n_rec, int k, int w, float entropy_thr, int threads);

/* .idx container, src/index.rs:17-22,57-72,130-164 (bincode 2 standard config = varint LE).
 * encode: returns bytes written (buffer must hold 3+9+9*n).  decode: returns 0 on success. */
size_t dcno_idx_encode(const uint64_t *keys, uint64_t n, uint8_t k, uint8_t w, uint8_t *out);
int dcno_idx_decode_header(const uint8_t *buf, size_t len, uint8_t *version, uint8_t *k, uint8_t *w,
                           uint64_t *count, size_t *body_off);
int dcno_idx_decode_keys(const uint8_t *buf, size_t len, size_t body_off, uint64_t count, uint64_t *out);

#ifdef __cplusplus
}
#endif
#endif
