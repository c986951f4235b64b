"""Time the whole-layer BatchNorm kernels (forward fused, backward fused) on layer shapes of the B=8 workload, warm (back to
back, as in the step).  Knobs via env: VINET_BN_MAXB, VINET_BN_MINRPT, VINET_BN_ATOM."""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from vinet_b200 import lib as L

lib = L.get()
st = torch.cuda.current_stream().cuda_stream
SHAPES = [(5505024, 64), (688128, 192), (172032, 96), (172032, 128), (172032, 32), (21504, 128), (21504, 320), (21504, 64), (2688, 384), (2688, 128)]
for rows, Cn in SHAPES:
    raw = torch.randn(rows, Cn, device="cuda").to(torch.bfloat16)
    out = torch.empty_like(raw)
    g = torch.randn(rows, Cn, device="cuda").to(torch.bfloat16)
    dy = torch.empty_like(raw)
    gamma, beta = torch.ones(Cn, device="cuda"), torch.zeros(Cn, device="cuda")
    rm, rv = torch.zeros(Cn, device="cuda"), torch.ones(Cn, device="cuda")
    ss, stt = torch.empty(2, Cn, device="cuda"), torch.empty(2, Cn, device="cuda")
    sums, bsums = torch.zeros(2 * Cn + 4, dtype=torch.float64, device="cuda"), torch.zeros(2 * Cn + 4, dtype=torch.float64, device="cuda")
    dgam, dbet = torch.empty(Cn, device="cuda"), torch.empty(Cn, device="cuda")
    sd = L.BnStats(); sd.y, sd.ld, sd.dtype, sd.rows, sd.C, sd.sums = raw.data_ptr(), Cn, L.BF16, rows, Cn, sums.data_ptr()
    f = L.BnFinalize(); f.sums, f.rows, f.C, f.gamma, f.beta, f.eps, f.momentum = sums.data_ptr(), rows, Cn, gamma.data_ptr(), beta.data_ptr(), 1e-3, 1e-3
    f.running_mean, f.running_var, f.training = rm.data_ptr(), rv.data_ptr(), 1
    f.scale, f.shift, f.mean, f.invstd = ss[0].data_ptr(), ss[1].data_ptr(), stt[0].data_ptr(), stt[1].data_ptr()
    a = L.BnApply(); a.y, a.ldy, a.dtype, a.rows, a.C, a.relu = raw.data_ptr(), Cn, L.BF16, rows, Cn, 1
    a.scale, a.shift, a.out, a.ldo, a.out_dtype = ss[0].data_ptr(), ss[1].data_ptr(), out.data_ptr(), Cn, L.BF16
    b = L.BnBwd(); b.g, b.ldg, b.y, b.ldy, b.dtype, b.rows, b.C, b.relu = g.data_ptr(), Cn, raw.data_ptr(), Cn, L.BF16, rows, Cn, 1
    b.scale, b.shift, b.mean, b.invstd = ss[0].data_ptr(), ss[1].data_ptr(), stt[0].data_ptr(), stt[1].data_ptr()
    b.gamma, b.sums, b.dgamma, b.dbeta, b.dy, b.lddy, b.dy_dtype, b.training, b.g_dtype = gamma.data_ptr(), bsums.data_ptr(), dgam.data_ptr(), dbet.data_ptr(), dy.data_ptr(), Cn, L.BF16, 1, L.BF16

    def timeit(fn, n=20):
        for _ in range(3):
            fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n * 1e3
    t_f = timeit(lambda: lib.call("vinet_bn_fwd_fused", C.byref(sd), C.byref(f), C.byref(a), st))
    t_s = timeit(lambda: (lib.call("vinet_bn_stats_finalize", C.byref(sd), C.byref(f), st), lib.call("vinet_bn_apply", C.byref(a), st)))
    fa, aa = (L.BnFinalize * 1)(), (L.BnApply * 1)()
    C.memmove(C.byref(fa[0]), C.byref(f), C.sizeof(L.BnFinalize)); C.memmove(C.byref(aa[0]), C.byref(a), C.sizeof(L.BnApply))
    sums.fill_(1.0)
    t_as = timeit(lambda: lib.call("vinet_bn_apply_stats_multi", fa, aa, 1, st))
    t_am = timeit(lambda: lib.call("vinet_bn_apply_multi", aa, 1, st))
    t_a1 = timeit(lambda: lib.call("vinet_bn_apply", C.byref(a), st))
    print("   apply only: bn_apply %7.1f us, apply_multi %7.1f us, apply_stats_multi %7.1f us" % (t_a1, t_am, t_as))
    sums.zero_()
    t_b = timeit(lambda: lib.call("vinet_bn_bwd_fused", C.byref(b), st))
    t_b2 = timeit(lambda: (lib.call("vinet_bn_bwd_reduce", C.byref(b), st), lib.call("vinet_bn_bwd_apply", C.byref(b), st)))
    mb = rows * Cn * 2 / 1e6
    print("rows %8d C %4d (%7.1f MB): fwd fused %7.1f us (%.2f TB/s)  split %7.1f us | bwd fused %7.1f us (%.2f TB/s)  split %7.1f us"
          % (rows, Cn, mb, t_f, 3 * mb / t_f, t_s, t_b, 5 * mb / t_b, t_b2), flush=True)
