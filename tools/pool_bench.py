"""Time the max-pool backward variants (atomic scatter / generic gather / compile-time gather) on the pools of the B=8 workload."""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from vinet_b200 import lib as L

lib = L.get()
st = torch.cuda.current_stream().cuda_stream
# name, (B,T,H,W,C), k, s, p, overwrite
CASES = [("base1.1", (8, 16, 112, 192, 64), (1, 3, 3), (1, 2, 2), (0, 1, 1), 1),
         ("maxp2", (8, 16, 56, 96, 192), (1, 3, 3), (1, 2, 2), (0, 1, 1), 0),
         ("maxp3", (8, 16, 28, 48, 480), (3, 3, 3), (2, 2, 2), (1, 1, 1), 0),
         ("3c.pool", (8, 16, 28, 48, 256), (3, 3, 3), (1, 1, 1), (1, 1, 1), 0),
         ("4f.pool", (8, 8, 14, 24, 528), (3, 3, 3), (1, 1, 1), (1, 1, 1), 0),
         ("5c.pool", (8, 4, 7, 12, 832), (3, 3, 3), (1, 1, 1), (1, 1, 1), 0)]
only = sys.argv[1] if len(sys.argv) > 1 else ""
modes = [int(m) for m in sys.argv[2].split(",")] if len(sys.argv) > 2 else [5, 3, 1]
for name, (B, T, H, W, Cn), k, s, p, ow in CASES:
    if only and only != name:
        continue
    To, Ho, Wo = [(n + 2 * pp - kk) // ss + 1 for n, kk, ss, pp in zip((T, H, W), k, s, p)]
    x = torch.randn(B, T, H, W, Cn, device="cuda").clamp_min(0).to(torch.bfloat16)
    out = torch.empty(B, To, Ho, Wo, Cn, dtype=torch.bfloat16, device="cuda")
    idx = torch.empty(B, To, Ho, Wo, Cn, dtype=torch.uint8, device="cuda")
    gout = torch.randn(B, To, Ho, Wo, Cn, device="cuda").to(torch.bfloat16)
    gin = torch.zeros(B, T, H, W, Cn, dtype=torch.bfloat16, device="cuda")
    d = L.Pool()
    d.x, d.ldx, d.dtype, d.xform = x.data_ptr(), Cn, L.BF16, L.XF_IDENT
    d.B, d.Ti, d.Hi, d.Wi, d.C = B, T, H, W, Cn
    (d.kt, d.kh, d.kw), (d.st, d.sh, d.sw), (d.pt, d.ph, d.pw) = k, s, p
    d.To, d.Ho, d.Wo, d.out, d.ldo, d.out_dtype, d.idx = To, Ho, Wo, out.data_ptr(), Cn, L.BF16, idx.data_ptr()
    lib.call("vinet_maxpool_fwd", C.byref(d), st)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        lib.call("vinet_maxpool_fwd", C.byref(d), st)
    e1.record()
    torch.cuda.synchronize()
    fwd_us = e0.elapsed_time(e1) / 5 * 1e3
    d.gout, d.ldgo, d.gin, d.ldgi, d.gout_dtype, d.gin_dtype, d.gin_overwrite = gout.data_ptr(), Cn, gin.data_ptr(), Cn, L.BF16, L.BF16, ow
    res = []
    for mode in modes:
        lib.call("vinet_debug_set", 3, mode)
        for _ in range(2):
            lib.call("vinet_maxpool_bwd", C.byref(d), st)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            lib.call("vinet_maxpool_bwd", C.byref(d), st)
        e1.record()
        torch.cuda.synchronize()
        res.append("mode %d: %7.1f us" % (mode, e0.elapsed_time(e1) / 5 * 1e3))
    lib.call("vinet_debug_set", 3, 1)
    mb = (gin.numel() * 2 * (1 if ow else 2) + gout.numel() * 2 + idx.numel()) / 1e6
    fmb = (x.numel() * 2 + out.numel() * 2 + idx.numel()) / 1e6
    print("%-8s in %s: fwd %7.1f us (min %.0f us) | bwd %s   (min traffic %.0f MB = %.0f us at 6.5 TB/s)"
          % (name, (B, T, H, W, Cn), fwd_us, fmb / 6.5, "  ".join(res), mb, mb / 6.5), flush=True)
