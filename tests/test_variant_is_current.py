"""The jittered stress build links the same objects as the product library; a stale one (built before the last ABI change)
would fail the GPU suite's stress tests for the wrong reason.  CPU check: if it is built, it exports what the product does."""
import ctypes as C
from pathlib import Path

import pytest

from helio_b200 import _ffi

VARIANT = Path(__file__).resolve().parent.parent / "build" / "variants" / "libhvx_jitter2.so"


def test_stress_variant_exports_the_whole_abi():
    if not VARIANT.exists():
        pytest.skip("build/variants/libhvx_jitter2.so is not built")
    lib = C.CDLL(str(VARIANT))
    missing = [name for name in _ffi.EXPORTS if not hasattr(lib, name)]
    assert not missing, f"stale stress variant (make -C helio_b200/csrc ../../build/variants/libhvx_jitter2.so): lacks {missing}"
