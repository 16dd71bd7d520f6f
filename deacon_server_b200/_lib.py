"""ctypes binding of libdeacon_cuda.so (include/deacon_cuda.h).

The product path has no CPU fallback: if the CUDA library is missing or no B200 is visible the
import / context creation raises.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("DCN_LIB") or os.path.join(_HERE, "libdeacon_cuda.so")   # DCN_LIB: kernel-variant experiments

u8p = C.POINTER(C.c_uint8)
u32p = C.POINTER(C.c_uint32)
u64p = C.POINTER(C.c_uint64)
f32p = C.POINTER(C.c_float)

# name -> (restype, argtypes); mirrors include/deacon_cuda.h one to one
SIGNATURES = {
    "dcn_device_count": (C.c_int, []),
    "dcn_ctx_create": (C.c_void_p, [C.c_int]),
    "dcn_ctx_destroy": (None, [C.c_void_p]),
    "dcn_last_error": (C.c_char_p, [C.c_void_p]),
    "dcn_host_alloc": (C.c_void_p, [C.c_size_t]),
    "dcn_host_free": (None, [C.c_void_p]),
    "dcn_index_upload": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint8, C.c_uint8]),
    "dcn_index_upload_device": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint8, C.c_uint8, C.c_void_p]),
    "dcn_index_info": (C.c_int, [C.c_void_p, u64p, u8p, u8p, u64p]),
    "dcn_index_set_load_factor": (C.c_int, [C.c_void_p, C.c_double]),
    "dcn_filter_batch": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_int, C.c_uint32, C.c_uint32,
                                   C.c_double, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    "dcn_filter_batch_device": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint64, C.c_int,
                                          C.c_uint32, C.c_uint32, C.c_double, C.c_int, C.c_void_p, C.c_void_p,
                                          C.c_void_p, C.c_void_p]),
    "dcn_filter_batch_device_hint": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint64, C.c_int,
                                               C.c_uint32, C.c_uint32, C.c_double, C.c_int, C.c_void_p, C.c_void_p,
                                               C.c_void_p, C.c_void_p, C.c_uint32]),
    "dcn_filter_batch_packed": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_int,
                                          C.c_uint32, C.c_uint32, C.c_double, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    "dcn_filter_batch_packed_sparse": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p, C.c_uint32, C.c_int,
                                                C.c_uint32, C.c_uint32, C.c_double, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    "dcn_pack_records_sparse": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint8, C.c_uint32, C.c_void_p, C.c_void_p, C.c_uint64,
                                         C.c_void_p, C.c_void_p]),
    "dcn_newline_bits": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint8, C.c_uint32, C.c_void_p]),
    "dcn_pack_records": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint8, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p]),
    "dcn_host_pack_threads": (C.c_int, [C.c_void_p, C.c_int]),
    "dcn_host_pack_fraction": (C.c_int, [C.c_void_p, C.c_double]),
    "dcn_pack_ascii": (C.c_int, [C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p]),
    "dcn_last_pack_ms": (C.c_int, [C.c_void_p, f32p]),
    "dcn_lookup_batch": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_double, C.c_int,
                                   C.c_void_p, C.c_void_p, C.c_void_p]),
    "dcn_lookup_batch_flags": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_double, C.c_int,
                                         C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "dcn_lookup_batch_device": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_double,
                                          C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "dcn_extract": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint8, C.c_uint8,
                              C.c_uint32, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64]),
    "dcn_extract_device": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint64, C.c_uint8, C.c_uint8,
                                     C.c_uint32, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, u64p, C.c_void_p]),
    "dcn_index_build": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint8, C.c_uint8, C.c_float,
                                  C.c_int, u64p]),
    "dcn_index_build_device": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint64, C.c_uint8,
                                         C.c_uint8, C.c_float, C.c_int, u64p, C.c_void_p]),
    "dcn_index_build_keys": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint64]),
    "dcn_index_build_keys_device": (C.c_void_p, [C.c_void_p]),
    "dcn_idx_decode": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint64, C.c_int, C.c_int, u8p, u8p, u8p, u64p, u64p]),
    "dcn_index_diff_sequences": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, u64p]),
    "dcn_idx_encode": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint64, u64p]),
    "dcn_working_set_info": (C.c_int, [C.c_void_p, u64p, u8p, u8p]),
    "dcn_working_set_release": (C.c_int, [C.c_void_p]),
    "dcn_index_make_resident": (C.c_int, [C.c_void_p]),
    "dcn_stats_get": (C.c_int, [C.c_void_p, u64p]),
    "dcn_stats_reset": (C.c_int, [C.c_void_p]),
    "dcn_stats_accumulate_device": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint32, C.c_int, C.c_void_p, C.c_void_p]),
    "dcn_last_timing": (C.c_int, [C.c_void_p, f32p, f32p, f32p]),
    "dcn_last_transfer_bytes": (C.c_int, [C.c_void_p, u64p, u64p]),
    "dcn_measure_random_access": (C.c_int, [C.c_void_p, u64p, f32p]),
    "dcn_measure_random_access_wide": (C.c_int, [C.c_void_p, C.c_int, u64p, f32p]),
    "dcn_launch_count": (C.c_uint64, [C.c_void_p]),
    "dcn_fused_time_take": (C.c_int, [C.c_void_p, f32p, u32p]),
}

_lib = None


class DeaconCudaError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"deacon_cuda error {code}: {msg}")
        self.code = code


def load():
    """Load libdeacon_cuda.so and bind every symbol of the header.  Raises if it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(make -C deacon_server_b200/csrc).  There is no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    old_variant = bool(os.environ.get("DCN_LIB")) and bool(os.environ.get("DCN_LIB_OLD"))   # kernel A/B against an older build only
    for name, (res, args) in SIGNATURES.items():
        if old_variant and not hasattr(lib, name):
            continue
        fn = getattr(lib, name)  # AttributeError if the .so does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib
