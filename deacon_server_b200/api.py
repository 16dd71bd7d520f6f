"""Host-side mirror of the reference's interface for the filter path, over the C ABI.

Names and argument meaning follow the reference (paths relative to /root/reference):
  IndexHeader / load_minimizer_hashes / write_minimizers ........ src/index.rs:17-164
  calculate_required_hits / meets_filtering_criteria ............ src/filter_common.rs:84-112
  DeaconGpu.should_keep_sequence / should_keep_pair ............. src/local_filter.rs:221-285
  DeaconGpu.unpaired_should_keep / paired_should_keep ........... src/remote_filter.rs:230-301
  DeaconGpu.get_minimizer_hashes_and_positions .................. src/filter_common.rs:211-310
  DeaconGpu.compute_minimizer_hashes ............................ src/minimizers.rs:53-68

All sequence work runs on the GPU through libdeacon_cuda.so; nothing here computes minimizers,
hashes or lookups on the CPU.
"""
from __future__ import annotations

import ctypes as C
import math
from dataclasses import dataclass

import numpy as np

from . import _lib
from ._lib import DeaconCudaError


# --------------------------------------------------------------------------- .idx container
@dataclass
class IndexHeader:
    """src/index.rs:17-54"""
    format_version: int = 2
    kmer_length: int = 31
    window_size: int = 15

    def validate(self) -> None:
        if self.format_version != 2:
            raise ValueError(f"Unsupported index format version: {self.format_version}")


def _read_varint(buf: np.ndarray, off: int):
    t = int(buf[off])
    if t < 251:
        return t, off + 1
    n = {0xFB: 2, 0xFC: 4, 0xFD: 8}.get(t)
    if n is None or off + 1 + n > len(buf):
        raise ValueError("malformed varint in index file")
    return int.from_bytes(buf[off + 1:off + 1 + n].tobytes(), "little"), off + 1 + n


def decode_index(data: bytes | np.ndarray):
    """bincode-2 standard-config stream: 3 x u8 header, varint count, varint u64 keys."""
    buf = np.frombuffer(data, np.uint8) if not isinstance(data, np.ndarray) else data
    if len(buf) < 4:
        raise ValueError("Failed to deserialise index header")
    header = IndexHeader(int(buf[0]), int(buf[1]), int(buf[2]))
    header.validate()
    count, off = _read_varint(buf, 3)
    body = buf[off:]
    # fast path: every key >= 2^32 -> 0xFD + 8 bytes (true for almost every xxh3 value)
    if len(body) == 9 * count and (count == 0 or bool((body[0::9] == 0xFD).all())):
        keys = body.reshape(count, 9)[:, 1:].copy().view("<u8").reshape(count)
        return keys, header
    keys = np.empty(count, np.uint64)
    p = 0
    for i in range(count):
        keys[i], p = _read_varint(body, p)
    return keys, header


def encode_index(keys: np.ndarray, header: IndexHeader) -> bytes:
    keys = np.ascontiguousarray(keys, np.uint64)

    def varint(v: int) -> bytes:
        if v < 251:
            return bytes([v])
        if v < 1 << 16:
            return b"\xfb" + v.to_bytes(2, "little")
        if v < 1 << 32:
            return b"\xfc" + v.to_bytes(4, "little")
        return b"\xfd" + v.to_bytes(8, "little")

    head = bytes([header.format_version, header.kmer_length, header.window_size]) + varint(len(keys))
    big = keys >= np.uint64(1 << 32)
    if bool(big.all()):
        out = np.empty((len(keys), 9), np.uint8)
        out[:, 0] = 0xFD
        out[:, 1:] = keys.astype("<u8").view(np.uint8).reshape(len(keys), 8)
        return head + out.tobytes()
    return head + b"".join(varint(int(v)) for v in keys)


def load_minimizer_hashes(path) -> tuple[np.ndarray, IndexHeader]:
    """src/index.rs:80-107 -> (keys, header).  Keys are returned as an array (set semantics are
    established when they are uploaded into the GPU table)."""
    with open(path, "rb") as f:
        data = np.fromfile(f, np.uint8)
    return decode_index(data)


def write_minimizers(keys: np.ndarray, header: IndexHeader, output_path) -> None:
    """src/index.rs:130-164.  Keys are written sorted ascending (the reference's order is its
    hash-set iteration order, which its loader ignores: src/index.rs:101-105)."""
    with open(output_path, "wb") as f:
        f.write(encode_index(np.sort(np.ascontiguousarray(keys, np.uint64)), header))


# --------------------------------------------------------------------------- host classification
def calculate_required_hits(abs_threshold: int, rel_threshold: float, total_minimizers: int) -> int:
    """src/filter_common.rs:84-96 (f64::round = half away from zero)."""
    if total_minimizers == 0:
        rel_required = 0
    else:
        x = rel_threshold * float(total_minimizers)
        if not x > 0.0:
            r = 0
        else:
            r = math.floor(x)
            if x - r >= 0.5:
                r += 1
        rel_required = max(int(r), 1)
    return max(abs_threshold, rel_required)


def meets_filtering_criteria(hit_count, total_minimizers, abs_threshold, rel_threshold, deplete) -> bool:
    """src/filter_common.rs:99-112"""
    required = calculate_required_hits(abs_threshold, rel_threshold, total_minimizers)
    return hit_count < required if deplete else hit_count >= required


# --------------------------------------------------------------------------- GPU context
def _concat(records):
    arrs = [np.frombuffer(bytes(r), np.uint8) if not isinstance(r, np.ndarray) else r.astype(np.uint8, copy=False)
            for r in records]
    off = np.zeros(len(arrs) + 1, np.uint64)
    if arrs:
        off[1:] = np.cumsum([len(a) for a in arrs], dtype=np.uint64)
    bases = np.concatenate(arrs) if arrs and int(off[-1]) else np.zeros(0, np.uint8)
    return np.ascontiguousarray(bases, np.uint8), off


class DeaconGpu:
    """One GPU context (dcn_ctx): resident index + batch operators.  No CPU fallback."""

    def __init__(self, device: int = 0):
        self._lib = _lib.load()
        self._ctx = self._lib.dcn_ctx_create(device)
        if not self._ctx:
            raise DeaconCudaError(-1, self._lib.dcn_last_error(None).decode())
        self.device = device
        self.header: IndexHeader | None = None

    def close(self):
        if getattr(self, "_ctx", None):
            self._lib.dcn_ctx_destroy(self._ctx)
            self._ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc: int):
        if rc != 0:
            raise DeaconCudaError(rc, self._lib.dcn_last_error(self._ctx).decode())

    # ---- B4
    def set_load_factor(self, load: float):
        self._check(self._lib.dcn_index_set_load_factor(self._ctx, load))

    def index_upload(self, keys: np.ndarray, header: IndexHeader | None = None):
        header = header or IndexHeader()
        keys = np.ascontiguousarray(keys, np.uint64)
        self._check(self._lib.dcn_index_upload(self._ctx, keys.ctypes.data, len(keys), header.kmer_length,
                                               header.window_size))
        self.header = header

    def index_upload_device(self, d_keys, header: IndexHeader | None = None, stream: int = 0):
        """d_keys: torch.uint64/int64 CUDA tensor."""
        header = header or IndexHeader()
        self._check(self._lib.dcn_index_upload_device(self._ctx, d_keys.data_ptr(), d_keys.numel(), header.kmer_length,
                                                      header.window_size, stream))
        self.header = header

    def load_index(self, path):
        keys, header = load_minimizer_hashes(path)
        self.index_upload(keys, header)
        return header

    def index_info(self):
        n, k, w, tb = C.c_uint64(), C.c_uint8(), C.c_uint8(), C.c_uint64()
        self._check(self._lib.dcn_index_info(self._ctx, C.byref(n), C.byref(k), C.byref(w), C.byref(tb)))
        return {"n_keys": n.value, "kmer_length": k.value, "window_size": w.value, "table_bytes": tb.value}

    # ---- B1
    def filter_batch(self, bases: np.ndarray, rec_off: np.ndarray, paired=False, prefix_length=0, abs_threshold=2,
                     rel_threshold=0.01, deplete=False):
        """-> (keep u8[], hits u32[], total u32[]) per record (or per pair)."""
        bases = np.ascontiguousarray(bases, np.uint8)
        rec_off = np.ascontiguousarray(rec_off, np.uint64)
        n_rec = len(rec_off) - 1
        nu = n_rec // 2 if paired else n_rec
        keep = np.zeros(max(nu, 1), np.uint8)
        hits = np.zeros(max(nu, 1), np.uint32)
        total = np.zeros(max(nu, 1), np.uint32)
        bptr = bases.ctypes.data if len(bases) else keep.ctypes.data
        self._check(self._lib.dcn_filter_batch(self._ctx, bptr, rec_off.ctypes.data, n_rec, int(paired),
                                               prefix_length, abs_threshold, rel_threshold, int(deplete),
                                               keep.ctypes.data, hits.ctypes.data, total.ctypes.data))
        return keep[:nu], hits[:nu], total[:nu]

    def filter_batch_ptr(self, bases_ptr, off_ptr, n_rec, paired, prefix_length, abs_threshold, rel_threshold, deplete,
                         keep_ptr, hits_ptr, total_ptr):
        """Raw host-pointer form (pinned buffers from dcn_host_alloc / torch pin_memory)."""
        self._check(self._lib.dcn_filter_batch(self._ctx, bases_ptr, off_ptr, n_rec, int(paired), prefix_length,
                                               abs_threshold, rel_threshold, int(deplete), keep_ptr, hits_ptr, total_ptr))

    def filter_batch_device(self, d_bases, d_rec_off, n_rec, n_bases, d_keep, d_hits, d_total, paired=False,
                            prefix_length=0, abs_threshold=2, rel_threshold=0.01, deplete=False, stream: int = 0,
                            max_unit_len: int = 0):
        """Inputs and outputs are CUDA tensors (anything with .data_ptr()); asynchronous on `stream`.
        max_unit_len (<= 1024): the caller's promise that no record / pair is longer -- the call then only
        enqueues (dcn_filter_batch_device_hint); 0 = unknown, the library asks the device (one small sync)."""
        self._check(self._lib.dcn_filter_batch_device_hint(
            self._ctx, d_bases.data_ptr(), d_rec_off.data_ptr(), n_rec, n_bases, int(paired), prefix_length,
            abs_threshold, rel_threshold, int(deplete), d_keep.data_ptr(), d_hits.data_ptr(), d_total.data_ptr(),
            stream, int(max_unit_len)))

    def filter_batch_packed(self, codes, inv, nl_bits, rec_off, paired=False, prefix_length=0, abs_threshold=2,
                            rel_threshold=0.01, deplete=False):
        """filter_batch for a batch already in the packed form (pack_ascii / newline_bits)."""
        rec_off = np.ascontiguousarray(rec_off, np.uint64)
        n_rec = len(rec_off) - 1
        nu = n_rec // 2 if paired else n_rec
        keep, hits, total = np.zeros(max(nu, 1), np.uint8), np.zeros(max(nu, 1), np.uint32), np.zeros(max(nu, 1), np.uint32)
        self._check(self._lib.dcn_filter_batch_packed(
            self._ctx, codes.ctypes.data, inv.ctypes.data, nl_bits.ctypes.data if nl_bits is not None else None,
            rec_off.ctypes.data, n_rec, int(paired), prefix_length, abs_threshold, rel_threshold, int(deplete),
            keep.ctypes.data, hits.ctypes.data, total.ctypes.data))
        return keep[:nu], hits[:nu], total[:nu]

    def filter_batch_packed_ptr(self, codes_ptr, inv_ptr, nl_ptr, off_ptr, n_rec, paired, prefix_length, abs_threshold,
                                rel_threshold, deplete, keep_ptr, hits_ptr, total_ptr):
        self._check(self._lib.dcn_filter_batch_packed(self._ctx, codes_ptr, inv_ptr, nl_ptr, off_ptr, n_rec, int(paired),
                                                      prefix_length, abs_threshold, rel_threshold, int(deplete), keep_ptr,
                                                      hits_ptr, total_ptr))

    def filter_batch_packed_sparse(self, codes, exc, nl_bits, rec_off, paired=False, prefix_length=0, abs_threshold=2,
                                   rel_threshold=0.01, deplete=False):
        """filter_batch for a batch in the sparse packed form (pack_records_sparse): exc = (block, mask) pairs, u32[n, 2]."""
        rec_off = np.ascontiguousarray(rec_off, np.uint64)
        exc = np.ascontiguousarray(exc, np.uint32).reshape(-1, 2)
        n_rec = len(rec_off) - 1
        nu = n_rec // 2 if paired else n_rec
        keep, hits, total = np.zeros(max(nu, 1), np.uint8), np.zeros(max(nu, 1), np.uint32), np.zeros(max(nu, 1), np.uint32)
        self._check(self._lib.dcn_filter_batch_packed_sparse(
            self._ctx, codes.ctypes.data, exc.ctypes.data if len(exc) else None, len(exc),
            nl_bits.ctypes.data if nl_bits is not None else None, rec_off.ctypes.data, n_rec, int(paired), prefix_length,
            abs_threshold, rel_threshold, int(deplete), keep.ctypes.data, hits.ctypes.data, total.ctypes.data))
        return keep[:nu], hits[:nu], total[:nu]

    def filter_batch_packed_sparse_ptr(self, codes_ptr, exc_ptr, n_exc, nl_ptr, off_ptr, n_rec, paired, prefix_length,
                                       abs_threshold, rel_threshold, deplete, keep_ptr, hits_ptr, total_ptr):
        self._check(self._lib.dcn_filter_batch_packed_sparse(self._ctx, codes_ptr, exc_ptr, int(n_exc), nl_ptr, off_ptr, n_rec,
                                                             int(paired), prefix_length, abs_threshold, rel_threshold,
                                                             int(deplete), keep_ptr, hits_ptr, total_ptr))

    def host_pack_threads(self, n: int):
        """Host threads that pack chunks for filter_batch (0 = ship ASCII over PCIe)."""
        self._check(self._lib.dcn_host_pack_threads(self._ctx, int(n)))

    def host_pack_fraction(self, fraction: float):
        """Share of chunks packed on the host (negative = automatic balance with the PCIe copy engine)."""
        self._check(self._lib.dcn_host_pack_fraction(self._ctx, float(fraction)))

    def should_keep_sequence(self, seq, **kw):
        """FilterProcessor::should_keep_sequence (src/local_filter.rs:221-252) -> (keep, hits, total)."""
        bases, off = _concat([seq])
        k, h, t = self.filter_batch(bases, off, paired=False, **kw)
        return bool(k[0]), int(h[0]), int(t[0])

    def should_keep_pair(self, seq1, seq2, **kw):
        """FilterProcessor::should_keep_pair (src/local_filter.rs:254-285)."""
        bases, off = _concat([seq1, seq2])
        k, h, t = self.filter_batch(bases, off, paired=True, **kw)
        return bool(k[0]), int(h[0]), int(t[0])

    # ---- B2
    def lookup_batch(self, hashes: np.ndarray, rec_off: np.ndarray, abs_threshold=2, rel_threshold=0.01, deplete=False):
        hashes = np.ascontiguousarray(hashes, np.uint64)
        rec_off = np.ascontiguousarray(rec_off, np.uint64)
        n = len(rec_off) - 1
        keep = np.zeros(max(n, 1), np.uint8)
        hits = np.zeros(max(n, 1), np.uint32)
        total = np.zeros(max(n, 1), np.uint32)
        hptr = hashes.ctypes.data if len(hashes) else keep.ctypes.data
        self._check(self._lib.dcn_lookup_batch(self._ctx, hptr, rec_off.ctypes.data, n, abs_threshold, rel_threshold,
                                               int(deplete), keep.ctypes.data, hits.ctypes.data, total.ctypes.data))
        return keep[:n], hits[:n], total[:n]

    def lookup_batch_flags(self, hashes: np.ndarray, rec_off: np.ndarray, abs_threshold=2, rel_threshold=0.01, deplete=False):
        """lookup_batch + per-hash flag: counted hit (in the index, first of its value in the record)."""
        hashes = np.ascontiguousarray(hashes, np.uint64)
        rec_off = np.ascontiguousarray(rec_off, np.uint64)
        n = len(rec_off) - 1
        keep, hits, total = np.zeros(max(n, 1), np.uint8), np.zeros(max(n, 1), np.uint32), np.zeros(max(n, 1), np.uint32)
        flags = np.zeros(max(len(hashes), 1), np.uint8)
        hptr = hashes.ctypes.data if len(hashes) else keep.ctypes.data
        self._check(self._lib.dcn_lookup_batch_flags(self._ctx, hptr, rec_off.ctypes.data, n, abs_threshold, rel_threshold,
                                                     int(deplete), keep.ctypes.data, hits.ctypes.data, total.ctypes.data,
                                                     flags.ctypes.data))
        return keep[:n], hits[:n], total[:n], flags[:len(hashes)]

    def _should_keep(self, recs, kmer_length, abs_threshold, rel_threshold, deplete, debug, paired):
        lists = [np.asarray(rec[0], np.uint64) for rec in recs]
        off = np.zeros(len(lists) + 1, np.uint64)
        if lists:
            off[1:] = np.cumsum([len(x) for x in lists], dtype=np.uint64)
        hashes = np.concatenate(lists) if lists and int(off[-1]) else np.zeros(0, np.uint64)
        if not debug:
            k, h, t = self.lookup_batch(hashes, off, abs_threshold, rel_threshold, deplete)
            return [(bool(k[i]), int(h[i]), int(t[i]), []) for i in range(len(lists))]
        k, h, t, fl = self.lookup_batch_flags(hashes, off, abs_threshold, rel_threshold, deplete)
        out = []
        for i, rec in enumerate(recs):
            kmers = []
            positions, seqs = rec[1], rec[2]
            for j in np.flatnonzero(fl[int(off[i]):int(off[i + 1])]):
                if j >= len(positions):
                    continue
                pos = int(positions[j])
                if paired:      # src/filter_common.rs:186-195: one sequence per hash (always empty in practice)
                    if j >= len(seqs):
                        continue
                    seq = bytes(seqs[j])
                    if pos + kmer_length > len(seq):
                        continue
                else:           # src/filter_common.rs:146-151
                    seq = bytes(seqs)
                kmers.append(seq[pos:pos + kmer_length].decode("utf-8", "replace"))
            out.append((bool(k[i]), int(h[i]), int(t[i]), kmers))
        return out

    def unpaired_should_keep(self, input_minimizers_and_positions, kmer_length, abs_threshold, rel_threshold, deplete,
                             debug=False):
        """src/remote_filter.rs:230-264: list of (hashes, positions, seq) -> list of (keep, hits, total, kmers)."""
        return self._should_keep(input_minimizers_and_positions, kmer_length, abs_threshold, rel_threshold, deplete, debug, False)

    def paired_should_keep(self, input_minimizers_and_positions, kmer_length, abs_threshold, rel_threshold, deplete,
                           debug=False):
        """src/remote_filter.rs:266-301: list of (pooled hashes, positions, sequences) -> the same tuples."""
        return self._should_keep(input_minimizers_and_positions, kmer_length, abs_threshold, rel_threshold, deplete, debug, True)

    def should_keep_sequence_debug(self, record_id: str, seq, prefix_length=0, abs_threshold=2, rel_threshold=0.01, deplete=False):
        """The `--debug` line of the default engine for one single-end record (src/local_filter.rs:351-363):
        extraction (B3) + lookup with hit flags (B2) on the GPU, k-mer strings cut from the sequence on the host."""
        hdr = self.header or IndexHeader()
        h, p = self.get_minimizer_hashes_and_positions(seq, prefix_length, hdr.kmer_length, hdr.window_size)
        eff = bytes(seq)
        if prefix_length > 0 and len(eff) > prefix_length:
            eff = eff[:prefix_length]
        if eff.endswith(b"\n"):
            eff = eff[:-1]
        (keep, hits, total, kmers), = self.unpaired_should_keep([(h, p, eff)], hdr.kmer_length, abs_threshold, rel_threshold,
                                                                 deplete, debug=True)
        line = f"DEBUG: {record_id} hits={hits}/{total} keep={'true' if keep else 'false'} kmers=[{','.join(kmers)}]"
        return keep, hits, total, kmers, line

    # ---- B3 / index build
    def extract(self, bases, rec_off, flavour=0, k=31, w=15, prefix_length=0, entropy_threshold=0.0, cap=None):
        bases = np.ascontiguousarray(bases, np.uint8)
        rec_off = np.ascontiguousarray(rec_off, np.uint64)
        n = len(rec_off) - 1
        cap = int(cap if cap is not None else max(16, len(bases)))
        hashes = np.zeros(cap, np.uint64)
        pos = np.zeros(cap, np.uint32)
        out_off = np.zeros(n + 1, np.uint64)
        bptr = bases.ctypes.data if len(bases) else hashes.ctypes.data
        self._check(self._lib.dcn_extract(self._ctx, flavour, bptr, rec_off.ctypes.data, n, k, w, prefix_length,
                                          entropy_threshold, hashes.ctypes.data, pos.ctypes.data, out_off.ctypes.data, cap))
        m = int(out_off[n])
        return hashes[:m], pos[:m], out_off

    def extract_device(self, d_bases, d_rec_off, n_rec, n_bases, d_hashes, d_pos, d_off, k=31, w=15, prefix_length=0,
                       flavour=0, entropy_threshold=0.0, stream: int = 0) -> int:
        """B3 on device-resident records into device CSR buffers -> number of minimizers."""
        n = C.c_uint64()
        self._check(self._lib.dcn_extract_device(
            self._ctx, flavour, d_bases.data_ptr(), d_rec_off.data_ptr(), n_rec, n_bases, k, w, prefix_length, entropy_threshold,
            d_hashes.data_ptr(), d_pos.data_ptr() if d_pos is not None else None, d_off.data_ptr(), d_hashes.numel(),
            C.byref(n), stream))
        return n.value

    def lookup_batch_device(self, d_hashes, d_rec_off, n_rec, d_keep, d_hits, d_total, abs_threshold=2, rel_threshold=0.01,
                            deplete=False, stream: int = 0):
        self._check(self._lib.dcn_lookup_batch_device(self._ctx, d_hashes.data_ptr(), d_rec_off.data_ptr(), n_rec, abs_threshold,
                                                      rel_threshold, int(deplete), d_keep.data_ptr(), d_hits.data_ptr(),
                                                      d_total.data_ptr(), stream))

    def stats_accumulate_device(self, d_rec_off, n_rec, paired, d_keep, stream: int = 0):
        self._check(self._lib.dcn_stats_accumulate_device(self._ctx, d_rec_off.data_ptr(), n_rec, int(paired), d_keep.data_ptr(), stream))

    def get_minimizer_hashes_and_positions(self, seq, prefix_length, kmer_length, window_size):
        """src/filter_common.rs:211 -> (hashes, positions)."""
        bases, off = _concat([seq])
        h, p, _ = self.extract(bases, off, 0, kmer_length, window_size, prefix_length)
        return h, p

    def compute_minimizer_hashes(self, seq, kmer_length, window_size, entropy_threshold=0.0):
        """src/minimizers.rs:53"""
        bases, off = _concat([seq])
        h, _, _ = self.extract(bases, off, 1, kmer_length, window_size, 0, entropy_threshold)
        return h

    def index_build(self, bases, rec_off, k=31, w=15, entropy_threshold=0.0, make_resident=True) -> np.ndarray:
        """index::build (src/index.rs:167-308) minus file I/O -> sorted unique keys."""
        bases = np.ascontiguousarray(bases, np.uint8)
        rec_off = np.ascontiguousarray(rec_off, np.uint64)
        n = C.c_uint64()
        bptr = bases.ctypes.data if len(bases) else rec_off.ctypes.data
        self._check(self._lib.dcn_index_build(self._ctx, bptr, rec_off.ctypes.data, len(rec_off) - 1, k, w,
                                              entropy_threshold, int(make_resident), C.byref(n)))
        keys = np.zeros(max(1, n.value), np.uint64)
        self._check(self._lib.dcn_index_build_keys(self._ctx, keys.ctypes.data, len(keys)))
        if make_resident:
            self.header = IndexHeader(2, k, w)
        return keys[:n.value]

    # ---- .idx codec + set algebra on the GPU (working key set of the ctx)
    SET_REPLACE, SET_UNION, SET_SUBTRACT = 0, 1, 2

    def idx_decode(self, data, mode: int = 0, make_resident: bool = False):
        """load_minimizer_hashes' decode (src/index.rs:80-107) on the GPU -> (header, keys in file, keys in the set)."""
        buf = np.frombuffer(data, np.uint8) if not isinstance(data, np.ndarray) else np.ascontiguousarray(data, np.uint8)
        ver, k, w, nf, ns = C.c_uint8(), C.c_uint8(), C.c_uint8(), C.c_uint64(), C.c_uint64()
        ptr = buf.ctypes.data if len(buf) else None
        self._check(self._lib.dcn_idx_decode(self._ctx, ptr, len(buf), mode, int(make_resident), C.byref(ver), C.byref(k),
                                             C.byref(w), C.byref(nf), C.byref(ns)))
        hdr = IndexHeader(ver.value, k.value, w.value)
        if make_resident:
            self.header = hdr
        return hdr, nf.value, ns.value

    def working_set_info(self) -> dict:
        n, k, w = C.c_uint64(), C.c_uint8(), C.c_uint8()
        self._check(self._lib.dcn_working_set_info(self._ctx, C.byref(n), C.byref(k), C.byref(w)))
        return {"n_keys": n.value, "kmer_length": k.value, "window_size": w.value}

    def working_set_release(self):
        """Free the working key set and all build / decode scratch (the resident table stays)."""
        self._check(self._lib.dcn_working_set_release(self._ctx))

    def working_keys(self) -> np.ndarray:
        """Sorted unique keys of the working set."""
        n = self.working_set_info()["n_keys"]
        keys = np.zeros(max(1, n), np.uint64)
        self._check(self._lib.dcn_index_build_keys(self._ctx, keys.ctypes.data, len(keys)))
        return keys[:n]

    def idx_encode(self) -> bytes:
        """write_minimizers (src/index.rs:130-164) of the working set, encoded on the GPU."""
        ln = C.c_uint64()
        rc = self._lib.dcn_idx_encode(self._ctx, None, 0, C.byref(ln))
        if rc not in (0, -6):
            self._check(rc)
        out = np.zeros(ln.value, np.uint8)
        self._check(self._lib.dcn_idx_encode(self._ctx, out.ctypes.data, len(out), C.byref(ln)))
        return out.tobytes()

    def index_union(self, data) -> int:
        """index::union step (src/index.rs:626-650): working set |= keys of this .idx -> size of the set."""
        return self.idx_decode(data, self.SET_UNION)[2]

    def index_diff(self, data) -> int:
        """index::diff, index - index (src/index.rs:515-528): working set -= keys of this .idx."""
        return self.idx_decode(data, self.SET_SUBTRACT)[2]

    def index_diff_sequences(self, bases, rec_off) -> int:
        """stream_diff_fastx (src/index.rs:311-419): working set -= minimizers of these records."""
        bases = np.ascontiguousarray(bases, np.uint8)
        rec_off = np.ascontiguousarray(rec_off, np.uint64)
        n = C.c_uint64()
        bptr = bases.ctypes.data if len(bases) else rec_off.ctypes.data
        self._check(self._lib.dcn_index_diff_sequences(self._ctx, bptr, rec_off.ctypes.data, len(rec_off) - 1, C.byref(n)))
        return n.value

    def index_make_resident(self):
        self._check(self._lib.dcn_index_make_resident(self._ctx))

    # ---- counters / measurement
    def stats(self) -> dict:
        c = (C.c_uint64 * 6)()
        self._check(self._lib.dcn_stats_get(self._ctx, c))
        names = ("total_seqs", "filtered_seqs", "total_bp", "output_bp", "filtered_bp", "output_seq_counter")
        return dict(zip(names, [int(x) for x in c]))

    def stats_reset(self):
        self._check(self._lib.dcn_stats_reset(self._ctx))

    def last_timing(self):
        a, b, c = C.c_float(), C.c_float(), C.c_float()
        self._check(self._lib.dcn_last_timing(self._ctx, C.byref(a), C.byref(b), C.byref(c)))
        return {"h2d_ms": a.value, "kernel_ms": b.value, "d2h_ms": c.value}

    def last_transfer_bytes(self):
        """(h2d, d2h) bytes the last host-pointer filter call moved over PCIe."""
        a, b = C.c_uint64(), C.c_uint64()
        self._check(self._lib.dcn_last_transfer_bytes(self._ctx, C.byref(a), C.byref(b)))
        return a.value, b.value

    def last_pack_ms(self) -> float:
        a = C.c_float()
        self._check(self._lib.dcn_last_pack_ms(self._ctx, C.byref(a)))
        return a.value

    def measure_random_access(self, n_probes: int):
        n, ms = C.c_uint64(n_probes), C.c_float()
        self._check(self._lib.dcn_measure_random_access(self._ctx, C.byref(n), C.byref(ms)))
        return n.value, ms.value

    def measure_random_access_wide(self, sectors: int, n_probes: int):
        n, ms = C.c_uint64(n_probes), C.c_float()
        self._check(self._lib.dcn_measure_random_access_wide(self._ctx, int(sectors), C.byref(n), C.byref(ms)))
        return n.value, ms.value

    def fused_time_take(self):
        """(total ms, launches) of the fused kernel since the last call (CUDA events, launching stream)."""
        ms, n = C.c_float(), C.c_uint32()
        self._check(self._lib.dcn_fused_time_take(self._ctx, C.byref(ms), C.byref(n)))
        return ms.value, n.value

    def index_build_device(self, d_bases, d_rec_off, n_rec, n_bases, k=31, w=15, entropy_threshold=0.0,
                           make_resident=True, stream: int = 0) -> int:
        """GPU index build from device-resident sequences -> number of distinct keys."""
        n = C.c_uint64()
        self._check(self._lib.dcn_index_build_device(self._ctx, d_bases.data_ptr(), d_rec_off.data_ptr(), n_rec, n_bases,
                                                     k, w, entropy_threshold, int(make_resident), C.byref(n), stream))
        if make_resident:
            self.header = IndexHeader(2, k, w)
        return n.value

    def index_build_keys_ptr(self) -> int:
        return int(self._lib.dcn_index_build_keys_device(self._ctx) or 0)

    def launch_count(self) -> int:
        return int(self._lib.dcn_launch_count(self._ctx))


def pack_ascii(bases: np.ndarray):
    """Host packer of the ingest stage (no GPU): -> (codes u32[], inv u16[]), see dcn_pack_ascii."""
    bases = np.ascontiguousarray(bases, np.uint8)
    n = len(bases)
    nw = 2 * ((n + 31) // 32)
    codes = np.zeros(max(nw, 1), np.uint32)
    inv = np.zeros(max(nw, 1), np.uint16)
    bptr = bases.ctypes.data if n else codes.ctypes.data
    rc = _lib.load().dcn_pack_ascii(bptr, n, codes.ctypes.data, inv.ctypes.data)
    if rc:
        raise DeaconCudaError(rc, "dcn_pack_ascii failed")
    return codes[:nw], inv[:nw]


def newline_bits(bases: np.ndarray, rec_off: np.ndarray, k: int = 31, prefix_length: int = 0) -> np.ndarray:
    """Per-record flag of the packed ingest form (dcn_newline_bits): the effective prefix ends in a newline."""
    bases = np.ascontiguousarray(bases, np.uint8)
    rec_off = np.ascontiguousarray(rec_off, np.uint64)
    n = len(rec_off) - 1
    out = np.zeros(max(1, (n + 31) // 32), np.uint32)
    bptr = bases.ctypes.data if len(bases) else out.ctypes.data
    rc = _lib.load().dcn_newline_bits(bptr, rec_off.ctypes.data, n, k, prefix_length, out.ctypes.data)
    if rc:
        raise DeaconCudaError(rc, "dcn_newline_bits failed")
    return out


def pack_records(bases: np.ndarray, rec_off: np.ndarray, k: int = 31, prefix_length: int = 0):
    """pack_ascii over bases[:rec_off[-1]] and newline_bits in one pass (dcn_pack_records) -> (codes, inv, nl_bits)."""
    bases = np.ascontiguousarray(bases, np.uint8)
    rec_off = np.ascontiguousarray(rec_off, np.uint64)
    n = len(rec_off) - 1
    nb = int(rec_off[-1]) if n else 0
    nw = 2 * ((nb + 31) // 32)
    codes = np.zeros(max(nw, 1), np.uint32)
    inv = np.zeros(max(nw, 1), np.uint16)
    nl = np.zeros(max(1, (n + 31) // 32), np.uint32)
    bptr = bases.ctypes.data if len(bases) else codes.ctypes.data
    rc = _lib.load().dcn_pack_records(bptr, rec_off.ctypes.data, n, k, prefix_length, codes.ctypes.data, inv.ctypes.data,
                                      nl.ctypes.data)
    if rc:
        raise DeaconCudaError(rc, "dcn_pack_records failed")
    return codes[:nw], inv[:nw], nl


def pack_records_sparse(bases: np.ndarray, rec_off: np.ndarray, k: int = 31, prefix_length: int = 0):
    """dcn_pack_records_sparse -> (codes u32[], exc u32[n, 2] = (32-base block, non-ACGT mask) ascending, nl_bits)."""
    bases = np.ascontiguousarray(bases, np.uint8)
    rec_off = np.ascontiguousarray(rec_off, np.uint64)
    n = len(rec_off) - 1
    nb = int(rec_off[-1]) if n else 0
    nw = 2 * ((nb + 31) // 32)
    codes = np.zeros(max(nw, 1), np.uint32)
    nl = np.zeros(max(1, (n + 31) // 32), np.uint32)
    bptr = bases.ctypes.data if len(bases) else codes.ctypes.data
    cap = max(64, nw // 64)
    while True:
        exc = np.zeros((cap, 2), np.uint32)
        n_exc = C.c_uint64()
        rc = _lib.load().dcn_pack_records_sparse(bptr, rec_off.ctypes.data, n, k, prefix_length, codes.ctypes.data, exc.ctypes.data,
                                                 cap, C.byref(n_exc), nl.ctypes.data)
        if rc == -6 and n_exc.value > cap:   # DCN_ERR_OVERFLOW: the list is longer (N-rich input)
            cap = int(n_exc.value)
            continue
        if rc:
            raise DeaconCudaError(rc, "dcn_pack_records_sparse failed")
        return codes[:nw], exc[:n_exc.value], nl
