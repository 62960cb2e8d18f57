import os
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
for entry in (str(ROOT), str(ROOT / "tests")):
    if entry not in sys.path:
        sys.path.insert(0, entry)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box via gpurun)")


@pytest.fixture(scope="session", autouse=True)
def _built():
    """Build the CUDA library and the oracle once per session if they are missing."""
    lib = ROOT / "helio_b200" / "libhelio_voxel_cuda.so"
    oracle = ROOT / "oracle" / "libhvx_oracle.so"
    if not lib.exists() or not oracle.exists():
        import __graft_entry__ as entry
        entry.build()
    yield


