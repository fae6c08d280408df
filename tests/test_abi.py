"""CPU checks of the boundary: the C-ABI library loads and exports every symbol include/dcp.h declares, and
fails loudly (no CPU fallback) when no CUDA device is present."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared(header):
    src = open(os.path.join(ROOT, "include", header)).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(dcph?_[a-z0-9_]+)\s*\(", src)))


def test_dcp_exports_every_declared_symbol():
    import dycore_b200  # noqa: F401
    from dycore_b200 import device
    L = ctypes.CDLL(device.lib_path())
    names = _declared("dcp.h")
    assert len(names) >= 25
    for n in names:
        assert hasattr(L, n), f"libdcp.so does not export {n}"
    assert sorted(device.EXPORTS) == names


def test_harness_exports_every_declared_symbol():
    L = ctypes.CDLL(os.path.join(ROOT, "lib", "libdcp_harness.so"))
    for n in _declared("dcp_harness.h"):
        assert hasattr(L, n), n


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from dycore_b200 import device
    with pytest.raises(device.DcpError, match="no CUDA device|CUDA"):
        device.Context(0)
