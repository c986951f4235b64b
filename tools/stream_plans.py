"""Print the streaming-kernel plan (csrc/conv_stream.cu cost model) for the conv layers of the B=8 32x224x384 workload.
CPU only: python tools/stream_plans.py"""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["VINET_STREAM_DEBUG"] = "1"
from vinet_b200 import lib as L

LAYERS = [  # name, B, T0, T1, H, W, Cin, Cout, k, st, pt
    ("base1.0.conv_t", 8, 32, 0, 112, 192, 64, 64, (7, 1, 1), 2, 3),
    ("base1.3.conv_s", 8, 16, 0, 56, 96, 64, 192, (1, 3, 3), 1, 0),
    ("base1.3.conv_t", 8, 16, 0, 56, 96, 192, 192, (3, 1, 1), 1, 1),
    ("3b.b1.conv_s", 8, 16, 0, 28, 48, 96, 128, (1, 3, 3), 1, 0),
    ("3b.b1.conv_t", 8, 16, 0, 28, 48, 128, 128, (3, 1, 1), 1, 1),
    ("3c.b1.conv_s", 8, 16, 0, 28, 48, 128, 192, (1, 3, 3), 1, 0),
    ("3c.b2.conv_s", 8, 16, 0, 28, 48, 32, 96, (1, 3, 3), 1, 0),
    ("4f.b1.conv_s", 8, 8, 0, 14, 24, 160, 320, (1, 3, 3), 1, 0),
    ("4f.b1.conv_t", 8, 8, 0, 14, 24, 320, 320, (3, 1, 1), 1, 1),
    ("5c.b1.conv_s", 8, 4, 0, 7, 12, 192, 384, (1, 3, 3), 1, 0),
    ("convtsp1", 8, 4, 0, 7, 12, 1024, 832, (1, 3, 3), 1, 0),
    ("convtsp2", 8, 4, 8, 14, 24, 832, 480, (3, 3, 3), 3, 0),
    ("convtsp3", 8, 4, 16, 28, 48, 480, 192, (5, 3, 3), 5, 0),
    ("convtsp4.0", 8, 4, 16, 56, 96, 192, 64, (5, 3, 3), 5, 0),
    ("convtsp4.3", 8, 4, 0, 112, 192, 64, 32, (2, 3, 3), 2, 0),
    ("convtsp4.6", 8, 2, 0, 224, 384, 32, 32, (2, 1, 1), 2, 0),
]


def fill(g, mode, B, Tr, H, W, rts, rto, Ts, Cs, taps, st, pt, ph, pw, T0):
    g.mode, g.dtype, g.B, g.Tr, g.Hr, g.Wr, g.row_tstep, g.row_toff = mode, L.BF16, B, Tr, H, W, rts, rto
    g.Ts, g.Hs, g.Ws, g.Cs, g.ntaps = Ts, H, W, Cs, len(taps)
    for i, (a, b, c) in enumerate(taps):
        g.tap[i][0], g.tap[i][1], g.tap[i][2] = a, b, c
    g.st, g.sh, g.sw, g.pt, g.ph, g.pw = st, 1, 1, pt, ph, pw
    g.src[0].ptr, g.src[0].ld, g.src[0].T, g.src[0].xform = 4096, Cs, T0, 0
    if Ts > T0:
        g.src[1].ptr, g.src[1].ld, g.src[1].T, g.src[1].xform = 8192, Cs, Ts - T0, 0


lib = L.get()
for name, B, T0, T1, H, W, Cin, Cout, k, st, pt in LAYERS:
    Ti = T0 + T1
    To = (Ti + 2 * pt - k[0]) // st + 1
    ph = pw = 1 if k[1] == 3 else 0
    taps = [(a, b, c) for a in range(k[0]) for b in range(k[1]) for c in range(k[2])]
    d = L.Conv()
    d.kernel, d.N, d.accumulate, d.out_dtype = L.KERNEL_TMA, Cout, 0, L.BF16
    fill(d.g, L.GATHER_FPROP, B, To, H, W, 1, 0, Ti, Cin, taps, st, pt, ph, pw, T0)
    bn, nt = C.c_int32(), C.c_int32()
    sys.stderr.write("%-16s fprop: " % name)
    sys.stderr.flush()
    lib.call("vinet_conv_tiling", C.byref(d), L.ENGINE_TC, C.byref(bn), C.byref(nt))
    sys.stderr.write("   -> block_n %d x %d\n" % (bn.value, nt.value))
    for rho in range(st):
        dts = [dt for dt in range(k[0]) if dt % st == rho]
        t0 = (rho - pt) % st
        frames = len(range(t0, Ti, st))
        if not dts or not frames:
            continue
        ptaps = [(dt, b, c) for dt in dts for b in range(k[1]) for c in range(k[2])]
        d = L.Conv()
        d.kernel, d.N, d.accumulate, d.out_dtype = L.KERNEL_TMA, Cin, 0, L.BF16
        fill(d.g, L.GATHER_DGRAD, B, frames, H, W, st, t0, To, Cout, ptaps, st, pt, ph, pw, To)
        sys.stderr.write("%-16s dgrad%d: " % (name, rho))
        sys.stderr.flush()
        lib.call("vinet_conv_tiling", C.byref(d), L.ENGINE_TC, C.byref(bn), C.byref(nt))
        sys.stderr.write("   -> block_n %d x %d\n" % (bn.value, nt.value))
        if rho == 0 and st > 2:
            break
