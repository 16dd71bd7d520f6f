"""deacon-server_b200: B200-native (sm_100a) filter hot path of Deacon behind a C ABI.

Python is only the test / benchmark host here; the product is deacon_server_b200/libdeacon_cuda.so
(sources in csrc/, interface in include/deacon_cuda.h).
"""
from .api import (DeaconGpu, IndexHeader, calculate_required_hits, load_minimizer_hashes,  # noqa: F401
                  meets_filtering_criteria, write_minimizers)
from ._lib import DeaconCudaError, LIB_PATH, load  # noqa: F401

__all__ = ["DeaconGpu", "IndexHeader", "DeaconCudaError", "load", "LIB_PATH", "load_minimizer_hashes",
           "write_minimizers", "calculate_required_hits", "meets_filtering_criteria"]
