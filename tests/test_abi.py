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


def test_conv_tiling_query_is_host_only_and_returns_valid_tilings():
    """vinet_conv_tiling (the library's per-layer N tiling, asked by the engine before it packs weights) runs without a GPU
    and must return a tiling the kernels accept: block_n a multiple of 16 and <= 256, n_tiles * block_n covering N."""
    import ctypes as C
    from vinet_b200 import lib as L
    lib = L.get()
    layers = [  # B, T0, T1, H, W, Cin, Cout, k, st, pt  (the B=8 32x224x384 workload and two sweep shapes)
        (8, 32, 0, 112, 192, 64, 64, (7, 1, 1), 2, 3), (8, 16, 0, 56, 96, 64, 192, (1, 3, 3), 1, 0),
        (8, 16, 0, 56, 96, 192, 192, (3, 1, 1), 1, 1), (8, 16, 0, 28, 48, 96, 128, (1, 3, 3), 1, 0),
        (8, 8, 0, 14, 24, 160, 320, (1, 3, 3), 1, 0), (8, 4, 0, 7, 12, 192, 384, (1, 3, 3), 1, 0),
        (8, 4, 8, 14, 24, 832, 480, (3, 3, 3), 3, 0), (8, 4, 16, 28, 48, 480, 192, (5, 3, 3), 5, 0),
        (8, 4, 16, 56, 96, 192, 64, (5, 3, 3), 5, 0), (8, 4, 0, 112, 192, 64, 32, (2, 3, 3), 2, 0),
        (1, 8, 0, 112, 192, 64, 192, (1, 3, 3), 1, 0), (2, 4, 0, 32, 48, 192, 192, (3, 1, 1), 1, 1),
        (8, 16, 0, 28, 48, 256, 32, (1, 1, 1), 1, 0)]
    for B, T0, T1, H, W, Cin, Cout, k, st, pt in layers:
        Ti = T0 + T1
        To = (Ti + 2 * pt - k[0]) // st + 1
        pad = 1 if k[1] == 3 else 0
        taps = [(a, b, c) for a in range(k[0]) for b in range(k[1]) for c in range(k[2])]
        d = L.Conv()
        d.kernel, d.N, d.out_dtype = L.KERNEL_TMA, Cout, L.BF16
        g = d.g
        g.mode, g.dtype, g.B, g.Tr, g.Hr, g.Wr, g.row_tstep, g.row_toff = L.GATHER_FPROP, L.BF16, B, To, H, W, 1, 0
        g.Ts, g.Hs, g.Ws, g.Cs, g.ntaps = Ti, H, W, Cin, len(taps)
        for i, (a, b, c) in enumerate(taps):
            g.tap[i][0], g.tap[i][1], g.tap[i][2] = a, b, c
        g.st, g.sh, g.sw, g.pt, g.ph, g.pw = st, 1, 1, pt, pad, pad
        g.src[0].ptr, g.src[0].ld, g.src[0].T = 4096, Cin, T0
        if T1:
            g.src[1].ptr, g.src[1].ld, g.src[1].T = 8192, Cin, T1
        bn, nt = C.c_int32(), C.c_int32()
        lib.call("vinet_conv_tiling", C.byref(d), L.ENGINE_TC, C.byref(bn), C.byref(nt))
        assert 16 <= bn.value <= 256 and bn.value % 16 == 0, (Cout, bn.value)
        assert nt.value >= 1 and nt.value * bn.value >= Cout and (nt.value - 1) * bn.value < Cout + 16, (Cout, bn.value, nt.value)
