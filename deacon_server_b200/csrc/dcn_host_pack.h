// dcn_host_pack.h -- host-side ingest helpers of the host-pointer pipeline: the
// 2-bit packing + non-ACGT bitmask the reference computes per record on the CPU
// (PackedSeqVec::from_ascii and the mask loop, src/filter_common.rs:238-258), done here once per
// batch by the pipeline's packer threads (filter_pipeline, dcn_api.cu) straight into pinned staging
// buffers, so that 0.375 B/bp instead of 1 B/bp cross PCIe.  Plain C++ (no CUDA): also compiled into the test-only host emulation.
#pragma once
#include <stdint.h>

#include <vector>

namespace dcn {

// Packs bases[0 .. n): codes[i] holds bases [16 i, 16 i + 16) as 2-bit codes (byte >> 1) & 3, base j
// at bits 2 j (packed-seq order A=0 C=1 T=2 G=3); inv[i] bit j = base 16 i + j is not one of
// ACGTacgt.  Both arrays must hold 2 * ceil(n / 32) entries; positions >= n are padded with code 0,
// non-ACGT 1.  simd = 0 forces the scalar loop (tests compare the two).  `bad32`, when given, receives
// (appended, ascending) the index of every 32-base block that holds a non-ACGT byte or padding: the only
// places a record-terminating newline can be, so the caller's newline-flag pass visits those and nothing else.
// `inv` may be null (sparse form): then the non-ACGT bits exist only as `bad_mask` -- the 32-bit mask of every block
// listed in `bad32`, in the same order (a batch of real reads has a handful: 4 bytes per listed block cross PCIe
// instead of 1 bit per base).
void pack_ascii(const uint8_t *bases, uint64_t n, uint32_t *codes, uint16_t *inv, int simd,
                std::vector<uint64_t> *bad32 = nullptr, std::vector<uint32_t> *bad_mask = nullptr);
bool pack_has_simd();

// One pass over a chunk: packs bases[a0 .. a0 + nb) (pack_ascii) and writes the newline flags of its `nr` records
// (off0[0 .. nr], absolute offsets into `bases`, off0[0] >= a0): bit q of nl = record q has raw length >= k and the
// last byte of its effective prefix is '\n' (src/filter_common.rs:217-229).  A record can only end in '\n' inside
// a 32-base block where the packer saw a non-ACGT byte, so only the records ending inside the listed blocks are
// looked at (a per-record pass over the chunk costs 28 % of the packing time).  `bad32` is scratch.
void pack_records(const uint8_t *bases, uint64_t a0, uint64_t nb, const uint64_t *off0, uint32_t nr, uint32_t k,
                  uint32_t prefix_len, uint32_t *codes, uint16_t *inv, uint32_t *nl, std::vector<uint64_t> &bad32,
                  std::vector<uint32_t> *bad_mask = nullptr);

// true iff off[r + 1] - off[r] == len0 for every r < n (off holds n + 1 entries): a chunk whose records all have
// one length gets its offsets written on the device instead of copied.  Memory-speed (AVX2) scan, early exit.
bool offsets_equal_length(const uint64_t *off, uint64_t n, uint64_t len0);

}  // namespace dcn
