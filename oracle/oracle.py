"""ctypes front-end of the CPU oracle (oracle/deacon_oracle.c).

TEST INFRASTRUCTURE ONLY: import this from tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs, never from deacon_server_b200 (the product path).
"parity unpinned" for minimizer selection -- see deacon_oracle.h.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libdeacon_oracle.so")

_u8p = C.POINTER(C.c_uint8)
_u32p = C.POINTER(C.c_uint32)
_u64p = C.POINTER(C.c_uint64)


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "deacon_oracle.c")
    hdr = os.path.join(_HERE, "deacon_oracle.h")
    if (not force and os.path.exists(_LIB_PATH)
            and os.path.getmtime(_LIB_PATH) >= max(os.path.getmtime(src), os.path.getmtime(hdr))):
        return _LIB_PATH
    subprocess.check_call(["make", "-C", _HERE, "-B", "libdeacon_oracle.so"],
                          stdout=subprocess.DEVNULL)
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    build()
    L = C.CDLL(_LIB_PATH)
    L.dcno_xxh3_u64.restype = C.c_uint64
    L.dcno_xxh3_u64.argtypes = [C.c_uint64]
    L.dcno_xxh3_u128.restype = C.c_uint64
    L.dcno_xxh3_u128.argtypes = [C.c_uint64, C.c_uint64]
    L.dcno_nthash_closed.restype = C.c_uint32
    L.dcno_nthash_closed.argtypes = [_u8p, C.c_int]
    for name in ("dcno_minimizer_positions", "dcno_minimizer_positions_brute"):
        f = getattr(L, name)
        f.restype = C.c_size_t
        f.argtypes = [_u8p, C.c_size_t, C.c_int, C.c_int, _u32p]
    L.dcno_extract_filter.restype = C.c_size_t
    L.dcno_extract_filter.argtypes = [_u8p, C.c_size_t, C.c_size_t, C.c_int, C.c_int, _u64p, _u32p]
    L.dcno_extract_index.restype = C.c_size_t
    L.dcno_extract_index.argtypes = [_u8p, C.c_size_t, C.c_int, C.c_int, C.c_float, _u64p]
    L.dcno_scaled_entropy.restype = C.c_float
    L.dcno_scaled_entropy.argtypes = [_u8p, C.c_int]
    L.dcno_required_hits.restype = C.c_uint64
    L.dcno_required_hits.argtypes = [C.c_uint64, C.c_double, C.c_uint64]
    L.dcno_meets_criteria.restype = C.c_int
    L.dcno_meets_criteria.argtypes = [C.c_uint64, C.c_uint64, C.c_uint64, C.c_double, C.c_int]
    L.dcno_set_new.restype = C.c_void_p
    L.dcno_set_new.argtypes = [C.c_uint64]
    L.dcno_set_free.argtypes = [C.c_void_p]
    L.dcno_set_insert_many.argtypes = [C.c_void_p, _u64p, C.c_uint64, C.c_int]
    L.dcno_set_contains.restype = C.c_int
    L.dcno_set_contains.argtypes = [C.c_void_p, C.c_uint64]
    L.dcno_set_len.restype = C.c_uint64
    L.dcno_set_len.argtypes = [C.c_void_p]
    L.dcno_set_keys.argtypes = [C.c_void_p, _u64p]
    L.dcno_lookup_batch.argtypes = [C.c_void_p, _u64p, _u64p, C.c_uint32, C.c_uint64, C.c_double,
                                    C.c_int, _u8p, _u32p, _u32p, C.c_int]
    L.dcno_filter_batch.argtypes = [C.c_void_p, _u8p, _u64p, C.c_uint32, C.c_int, C.c_uint64,
                                    C.c_int, C.c_int, C.c_uint64, C.c_double, C.c_int,
                                    _u8p, _u32p, _u32p, C.c_int]
    L.dcno_index_build.argtypes = [C.c_void_p, _u8p, _u64p, C.c_uint32, C.c_int, C.c_int,
                                   C.c_float, C.c_int]
    L.dcno_idx_encode.restype = C.c_size_t
    L.dcno_idx_encode.argtypes = [_u64p, C.c_uint64, C.c_uint8, C.c_uint8, _u8p]
    L.dcno_idx_decode_header.restype = C.c_int
    L.dcno_idx_decode_header.argtypes = [_u8p, C.c_size_t, _u8p, _u8p, _u8p, _u64p,
                                         C.POINTER(C.c_size_t)]
    L.dcno_idx_decode_keys.restype = C.c_int
    L.dcno_idx_decode_keys.argtypes = [_u8p, C.c_size_t, C.c_size_t, C.c_uint64, _u64p]
    _lib = L
    return L


def _p(a, t):
    return a.ctypes.data_as(t)


def _bytes(seq) -> np.ndarray:
    if isinstance(seq, str):
        seq = seq.encode()
    if isinstance(seq, (bytes, bytearray)):
        return np.frombuffer(bytes(seq), dtype=np.uint8).copy() if len(seq) else np.zeros(0, np.uint8)
    return np.ascontiguousarray(seq, dtype=np.uint8)


def xxh3_u64(v: int) -> int:
    return lib().dcno_xxh3_u64(v & 0xFFFFFFFFFFFFFFFF)


def xxh3_u128(v: int) -> int:
    return lib().dcno_xxh3_u128(v & 0xFFFFFFFFFFFFFFFF, (v >> 64) & 0xFFFFFFFFFFFFFFFF)


def pack_codes(seq) -> np.ndarray:
    """packed-seq lossy 2-bit code: (byte >> 1) & 3 -> A=0 C=1 T=2 G=3."""
    return ((_bytes(seq) >> 1) & 3).astype(np.uint8)


def nthash_closed(codes: np.ndarray, k: int) -> int:
    codes = np.ascontiguousarray(codes, np.uint8)
    return lib().dcno_nthash_closed(_p(codes, _u8p), k)


def minimizer_positions(codes: np.ndarray, k: int, w: int, brute: bool = False) -> np.ndarray:
    codes = np.ascontiguousarray(codes, np.uint8)
    out = np.zeros(max(1, len(codes)), np.uint32)
    f = lib().dcno_minimizer_positions_brute if brute else lib().dcno_minimizer_positions
    n = f(_p(codes, _u8p), len(codes), k, w, _p(out, _u32p))
    return out[:n].copy()


def extract_filter(seq, k: int = 31, w: int = 15, prefix_len: int = 0):
    """get_minimizer_hashes_and_positions (src/filter_common.rs:211) -> (hashes, positions)."""
    b = _bytes(seq)
    h = np.zeros(max(1, len(b)), np.uint64)
    p = np.zeros(max(1, len(b)), np.uint32)
    n = lib().dcno_extract_filter(_p(b, _u8p), len(b), prefix_len, k, w, _p(h, _u64p), _p(p, _u32p))
    return h[:n].copy(), p[:n].copy()


def extract_index(seq, k: int = 31, w: int = 15, entropy: float = 0.0) -> np.ndarray:
    """compute_minimizer_hashes (src/minimizers.rs:53)."""
    b = _bytes(seq)
    h = np.zeros(max(1, len(b)), np.uint64)
    n = lib().dcno_extract_index(_p(b, _u8p), len(b), k, w, entropy, _p(h, _u64p))
    return h[:n].copy()


def scaled_entropy(kmer, k: int | None = None) -> float:
    b = _bytes(kmer)
    return float(lib().dcno_scaled_entropy(_p(b, _u8p), len(b) if k is None else k))


def required_hits(abs_thr: int, rel_thr: float, total: int) -> int:
    return lib().dcno_required_hits(abs_thr, rel_thr, total)


def meets_criteria(hits: int, total: int, abs_thr: int, rel_thr: float, deplete: bool) -> bool:
    return bool(lib().dcno_meets_criteria(hits, total, abs_thr, rel_thr, int(deplete)))


class IndexSet:
    """FxHashSet<u64> stand-in."""

    def __init__(self, keys=None, threads: int = 1, expected: int = 0):
        self._h = lib().dcno_set_new(max(expected, 0 if keys is None else len(keys)))
        if keys is not None and len(keys):
            self.insert(keys, threads)

    def insert(self, keys, threads: int = 1):
        keys = np.ascontiguousarray(keys, np.uint64)
        lib().dcno_set_insert_many(self._h, _p(keys, _u64p), len(keys), threads)

    def __contains__(self, key: int) -> bool:
        return bool(lib().dcno_set_contains(self._h, int(key)))

    def __len__(self) -> int:
        return int(lib().dcno_set_len(self._h))

    def keys(self) -> np.ndarray:
        out = np.zeros(max(1, len(self)), np.uint64)
        lib().dcno_set_keys(self._h, _p(out, _u64p))
        return np.sort(out[:len(self)])

    def __del__(self):
        try:
            if self._h:
                lib().dcno_set_free(self._h)
                self._h = None
        except Exception:
            pass


def concat_records(records):
    """list of bytes/str -> (bases u8[], rec_off u64[n+1])."""
    arrs = [_bytes(r) for r in records]
    off = np.zeros(len(arrs) + 1, np.uint64)
    if arrs:
        off[1:] = np.cumsum([len(a) for a in arrs], dtype=np.uint64)
    bases = np.concatenate(arrs) if arrs else np.zeros(0, np.uint8)
    return np.ascontiguousarray(bases, np.uint8), off


def index_build(records, k=31, w=15, entropy=0.0, threads=1) -> IndexSet:
    bases, off = records if isinstance(records, tuple) else concat_records(records)
    s = IndexSet(expected=int(0.14 * len(bases)) + 64)   # pre-sized: ~0.126 minimizers per base
    if len(bases) == 0:
        bases = np.zeros(1, np.uint8)
    lib().dcno_index_build(s._h, _p(bases, _u8p), _p(off, _u64p), len(off) - 1, k, w, entropy, threads)
    return s


def filter_batch(idx: IndexSet, bases, rec_off, paired=False, prefix_len=0, k=31, w=15,
                 abs_thr=2, rel_thr=0.01, deplete=False, threads=1):
    bases = np.ascontiguousarray(bases, np.uint8)
    rec_off = np.ascontiguousarray(rec_off, np.uint64)
    n_rec = len(rec_off) - 1
    n_unit = n_rec // 2 if paired else n_rec
    keep = np.zeros(max(1, n_unit), np.uint8)
    hits = np.zeros(max(1, n_unit), np.uint32)
    total = np.zeros(max(1, n_unit), np.uint32)
    if len(bases) == 0:
        bases = np.zeros(1, np.uint8)
    lib().dcno_filter_batch(idx._h, _p(bases, _u8p), _p(rec_off, _u64p), n_rec, int(paired),
                            prefix_len, k, w, abs_thr, rel_thr, int(deplete),
                            _p(keep, _u8p), _p(hits, _u32p), _p(total, _u32p), threads)
    return keep[:n_unit], hits[:n_unit], total[:n_unit]


def lookup_batch(idx: IndexSet, hashes, rec_off, abs_thr=2, rel_thr=0.01, deplete=False, threads=1):
    hashes = np.ascontiguousarray(hashes, np.uint64)
    rec_off = np.ascontiguousarray(rec_off, np.uint64)
    n = len(rec_off) - 1
    keep = np.zeros(max(1, n), np.uint8)
    hits = np.zeros(max(1, n), np.uint32)
    total = np.zeros(max(1, n), np.uint32)
    if len(hashes) == 0:
        hashes = np.zeros(1, np.uint64)
    lib().dcno_lookup_batch(idx._h, _p(hashes, _u64p), _p(rec_off, _u64p), n, abs_thr, rel_thr,
                            int(deplete), _p(keep, _u8p), _p(hits, _u32p), _p(total, _u32p), threads)
    return keep[:n], hits[:n], total[:n]


def idx_encode(keys, k: int, w: int) -> bytes:
    keys = np.ascontiguousarray(keys, np.uint64)
    out = np.zeros(3 + 9 + 9 * len(keys), np.uint8)
    kk = keys if len(keys) else np.zeros(1, np.uint64)
    n = lib().dcno_idx_encode(_p(kk, _u64p), len(keys), k, w, _p(out, _u8p))
    return out[:n].tobytes()


def idx_decode(buf: bytes):
    """-> (version, k, w, keys)."""
    b = np.frombuffer(buf, np.uint8).copy()
    ver, k, w = C.c_uint8(), C.c_uint8(), C.c_uint8()
    cnt, off = C.c_uint64(), C.c_size_t()
    rc = lib().dcno_idx_decode_header(_p(b, _u8p), len(b), C.byref(ver), C.byref(k), C.byref(w),
                                      C.byref(cnt), C.byref(off))
    if rc == -2:
        raise ValueError(f"Unsupported index format version: {ver.value}")
    if rc:
        raise ValueError("truncated index header")
    keys = np.zeros(max(1, cnt.value), np.uint64)
    if lib().dcno_idx_decode_keys(_p(b, _u8p), len(b), off.value, cnt.value, _p(keys, _u64p)):
        raise ValueError("truncated index body")
    return ver.value, k.value, w.value, keys[:cnt.value]
