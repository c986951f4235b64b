"""CPU: the C-ABI library loads, exports every symbol include/vinet_b200.h declares, and its struct
layouts agree with the ctypes binding (no compute calls)."""
import ctypes
import os
import re

from vinet_b200 import lib as L

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "vinet_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(vinet_[a-z0-9_]+)\s*\(", src)))


def test_library_loads_and_exports_every_declared_symbol():
    lib = L.get()          # raises on a missing .so or an ABI (sizeof) mismatch
    dll = ctypes.CDLL(L.LIB_PATH)
    names = _declared()
    assert len(names) >= 30
    for n in names:
        assert hasattr(dll, n), "missing export: " + n
        assert n in L.SIGNATURES, "binding lacks " + n
    assert "sm_100a" in lib.version()


def test_struct_sizes_match():
    lib = L.get()
    sizes = (ctypes.c_int64 * len(L.ABI_STRUCTS))()
    n = lib.fn["vinet_abi_sizes"](sizes, len(L.ABI_STRUCTS))
    assert n == len(L.ABI_STRUCTS)
    assert list(sizes) == [ctypes.sizeof(t) for t in L.ABI_STRUCTS]


def test_no_cpu_fallback():
    import pytest
    import torch
    from vinet_b200 import VideoSaliencyModel, kldiv
    m = VideoSaliencyModel(num_clips=8)
    with pytest.raises(RuntimeError):
        m(torch.zeros(1, 3, 8, 32, 32))
    with pytest.raises(RuntimeError):
        kldiv(torch.rand(1, 4, 4), torch.rand(1, 4, 4))
