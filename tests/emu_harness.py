"""Loads the TEST-ONLY host emulation of the CUDA tile pipeline (tests/emu/dcn_emu.cpp)."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "emu")
_FLAGS = os.environ.get("DCN_EMU_FLAGS", "").split()          # e.g. -DDCN_NT=128: kernel-geometry experiments
_SO = os.path.join(_HERE, "libdcn_emu%s.so" % ("_" + "_".join(f.strip("-").replace("=", "") for f in _FLAGS) if _FLAGS else ""))
_CSRC = os.path.join(os.path.dirname(_HERE), "..", "deacon_server_b200", "csrc")
u8p, u32p, u64p = C.POINTER(C.c_uint8), C.POINTER(C.c_uint32), C.POINTER(C.c_uint64)


def build():
    srcs = [os.path.join(_HERE, "dcn_emu.cpp")] + [os.path.join(_CSRC, f) for f in (
        "dcn_host_pack.cpp", "dcn_core.cuh", "dcn_plan.cuh", "dcn_tile.cuh", "dcn_warp.cuh", "dcn_generic.cuh", "dcn_host_pack.h")]
    if os.path.exists(_SO) and os.path.getmtime(_SO) >= max(os.path.getmtime(s) for s in srcs):
        return _SO
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-pthread", "-Wno-unknown-pragmas", *_FLAGS, "-o", _SO, srcs[0], srcs[1]])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
    return _lib


def _p(a, t):
    return a.ctypes.data_as(t)


def filter_batch(keys, bases, off, paired=False, prefix=0, abs_thr=2, rel=0.01, deplete=False, load=0.5, packed=False):
    L = lib()
    keys = np.ascontiguousarray(keys, np.uint64)
    slots, nb, he = u64p(), C.c_uint64(), C.c_int()
    kk = keys if len(keys) else np.zeros(1, np.uint64)
    L.emu_table_build(_p(kk, u64p), C.c_uint64(len(keys)), C.c_double(load), C.byref(slots), C.byref(nb), C.byref(he))
    n_rec = len(off) - 1
    nu = n_rec // 2 if paired else n_rec
    keep = np.zeros(max(nu, 1), np.uint8)
    hits = np.zeros(max(nu, 1), np.uint32)
    tot = np.zeros(max(nu, 1), np.uint32)
    b = bases if len(bases) else np.zeros(1, np.uint8)
    rc = L.emu_filter_batch(slots, nb, he, _p(b, u8p), _p(off, u64p), C.c_uint32(n_rec), int(paired), C.c_uint32(prefix),
                            C.c_uint32(abs_thr), C.c_double(rel), int(deplete), _p(keep, u8p), _p(hits, u32p), _p(tot, u32p),
                            int(packed))
    L.emu_free(slots)
    return rc, keep[:nu], hits[:nu], tot[:nu]


def pack_ascii(bases, simd=True):
    n = len(bases)
    nw = 2 * ((n + 31) // 32)
    codes, inv = np.zeros(max(nw, 1), np.uint32), np.zeros(max(nw, 1), np.uint16)
    b = bases if n else np.zeros(1, np.uint8)
    lib().emu_pack_ascii(_p(b, u8p), C.c_uint64(n), _p(codes, u32p), inv.ctypes.data_as(C.POINTER(C.c_uint16)), int(simd))
    return codes[:nw], inv[:nw]


def index_extract(bases, off, entropy_bitmap=None):
    """-> unordered hashes (with duplicates) of every record, index flavour."""
    L = lib()
    L.emu_index_extract.restype = C.c_longlong
    n_rec = len(off) - 1
    cap = max(16, len(bases))
    out = np.zeros(cap, np.uint64)
    b = bases if len(bases) else np.zeros(1, np.uint8)
    eb = None if entropy_bitmap is None else _p(np.ascontiguousarray(entropy_bitmap, np.uint32), u32p)
    n = L.emu_index_extract(_p(b, u8p), _p(off, u64p), C.c_uint32(n_rec), eb, _p(out, u64p), C.c_uint64(cap))
    assert 0 <= n <= cap
    return out[:n]


def set_impl(name):
    """'warp' (default): warp tiles + CTA tail, as the product runs; 'cta': the CTA-tile fused kernel."""
    lib().emu_set_impl(1 if name == "cta" else 0)


def last_overflow_units():
    lib().emu_last_overflow_units.restype = C.c_uint64
    return int(lib().emu_last_overflow_units())


def tile_extract(bases, off, prefix=0, cap=None):
    """B3 through the warp-tile extraction (k = 31, w = 15, short records) -> (hashes, positions, out_off) or (rc, None, off)."""
    L = lib()
    L.emu_tile_extract.restype = C.c_longlong
    n_rec = len(off) - 1
    cap = max(16, len(bases)) if cap is None else cap
    oh, op, oo = np.zeros(cap, np.uint64), np.zeros(cap, np.uint32), np.zeros(n_rec + 1, np.uint64)
    b = bases if len(bases) else np.zeros(1, np.uint8)
    n = L.emu_tile_extract(_p(b, u8p), _p(off, u64p), C.c_uint32(n_rec), C.c_uint32(prefix), _p(oh, u64p), _p(op, u32p), _p(oo, u64p),
                           C.c_uint64(cap))
    if n < 0:
        return int(n), None, oo
    return oh[:n], op[:n], oo


def set_dedup_cap(cap):
    lib().emu_set_dedup_cap(C.c_uint64(cap))


def set_hit_list_cap(cap):
    """Long path of the warp-tile kernel: entries of the compacted hit list (0 = the real size)."""
    lib().emu_set_hit_list_cap(C.c_uint32(cap))


def generic_extract(bases, off, flavour=0, k=31, w=15, prefix=0, entropy_bitmap=None, cstride=256, cap=None):
    """-> (hashes, positions, out_off): the generic (k, w) extraction (B3), CSR per record."""
    L = lib()
    L.emu_generic_extract.restype = C.c_longlong
    n_rec = len(off) - 1
    cap = max(16, len(bases)) if cap is None else cap
    oh, op, oo = np.zeros(cap, np.uint64), np.zeros(cap, np.uint32), np.zeros(n_rec + 1, np.uint64)
    b = bases if len(bases) else np.zeros(1, np.uint8)
    eb = None if entropy_bitmap is None else _p(np.ascontiguousarray(entropy_bitmap, np.uint32), u32p)
    n = L.emu_generic_extract(int(flavour), _p(b, u8p), _p(off, u64p), C.c_uint32(n_rec), int(k), int(w), C.c_uint32(prefix), eb,
                              C.c_uint32(cstride), _p(oh, u64p), _p(op, u32p), _p(oo, u64p), C.c_uint64(cap))
    if n < 0:
        return int(n), None, oo
    return oh[:n], op[:n], oo


def generic_filter(keys, bases, off, k, w, paired=False, prefix=0, abs_thr=2, rel=0.01, deplete=False, cstride=256):
    L = lib()
    keys = np.ascontiguousarray(keys, np.uint64)
    slots, nb, he = u64p(), C.c_uint64(), C.c_int()
    kk = keys if len(keys) else np.zeros(1, np.uint64)
    L.emu_table_build(_p(kk, u64p), C.c_uint64(len(keys)), C.c_double(0.5), C.byref(slots), C.byref(nb), C.byref(he))
    n_rec = len(off) - 1
    nu = n_rec // 2 if paired else n_rec
    keep, hits, tot = np.zeros(max(nu, 1), np.uint8), np.zeros(max(nu, 1), np.uint32), np.zeros(max(nu, 1), np.uint32)
    b = bases if len(bases) else np.zeros(1, np.uint8)
    rc = L.emu_generic_filter(slots, nb, he, _p(b, u8p), _p(off, u64p), C.c_uint32(n_rec), int(paired), C.c_uint32(prefix), int(k),
                              int(w), C.c_uint32(cstride), C.c_uint32(abs_thr), C.c_double(rel), int(deplete), _p(keep, u8p),
                              _p(hits, u32p), _p(tot, u32p))
    L.emu_free(slots)
    return rc, keep[:nu], hits[:nu], tot[:nu]
