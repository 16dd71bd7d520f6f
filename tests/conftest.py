import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")
    # the product library and the host driver are built artefacts (git-ignored): build them if a fresh checkout has none
    pkg = os.path.join(ROOT, "deacon_server_b200")
    if not (os.path.exists(os.path.join(pkg, "libdeacon_cuda.so")) and os.path.exists(os.path.join(pkg, "deacon-b200"))):
        import __graft_entry__
        __graft_entry__.build()


@pytest.fixture(scope="session")
def gpu():
    """One dcn_ctx on cuda:0.  Fails loudly (no CPU fallback) if the library or the GPU is missing."""
    import deacon_server_b200 as d
    g = d.DeaconGpu(0)
    yield g
    g.close()
