"""Fused vs materialised 2x up-sampling in front of the decoder convolutions (B=8 layers of the north-star workload):
times fprop and the weight gradient of each layer with source 0 read THROUGH relu + bilinear 2x (VINET_XF_UP2, interpolating
producer warps) against the same convolution on the materialised hi-res tensor + the standalone upsample kernel.
  python tools/up2_bench.py [layer ...] [--iters N] [--only-fused]   (ncu target: --only-fused --iters 1)"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from vinet_b200 import lib as L
from vinet_b200.engine import Act, ConvGeom, Engine

LAYERS = {  # name: B, T0 (up-sampled frames), T1 (skip frames), hi-res H, W, Cin, Cout, kt
    "convtsp2": (8, 4, 8, 14, 24, 832, 480, 3),
    "convtsp3": (8, 4, 16, 28, 48, 480, 192, 5),
    "convtsp4.0": (8, 4, 16, 56, 96, 192, 64, 5),
    "convtsp4.3": (8, 4, 0, 112, 192, 64, 32, 2),
}
args = [a for a in sys.argv[1:] if not a.startswith("--")]
iters = int(sys.argv[sys.argv.index("--iters") + 1]) if "--iters" in sys.argv else 5
only_fused = "--only-fused" in sys.argv
if "--iters" in sys.argv:
    args = [a for a in args if a != sys.argv[sys.argv.index("--iters") + 1]]
names = args or list(LAYERS)
dev = torch.device("cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timed(fn):
    ts = []
    for _ in range(iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return min(ts)


for name in names:
    B, T0, T1, H, W, Cin, Cout, kt = LAYERS[name]
    e = Engine("bf16")
    e.begin(dev, True, True)
    g = torch.Generator().manual_seed(0)
    z = e.new_act("z", B, T0, H // 2, W // 2, Cin)
    z.buf.copy_(torch.randn(z.buf.shape, generator=g))
    u = Act(z.buf, B, T0, H, W, Cin, 0, L.XF_RELU | L.XF_UP2)
    u.up2, u.name = True, "z"
    e.want_grad(u, "z.up")
    srcs_f = [u]
    if T1:
        y = e.new_act("y", B, T1, H, W, Cin)
        y.buf.copy_(torch.randn(y.buf.shape, generator=g))
        srcs_f.append(y)
    w = (torch.randn(Cout, Cin, kt, 3, 3, generator=g) / (Cin * kt * 9) ** 0.5).to(dev)
    geom = ConvGeom((kt, 3, 3), (kt, 1, 1), (0, 1, 1))
    To, Ho, Wo = geom.out_dims(T0 + T1, H, W)
    out = e.new_act("o", B, To, Ho, Wo, Cout)
    dy = torch.randn(out.buf.shape, device=dev).to(e.tdtype)
    gflop = 2.0 * B * To * Ho * Wo * kt * 9 * Cin * Cout / 1e9
    res = {}
    for mode in (["fused"] if only_fused else ["fused", "materialised"]):
        if mode == "materialised":
            m = e.materialize_up2(u)
            srcs = [m] + srcs_f[1:]
            up_d = L.Upsample()
            up_d.z, up_d.ldz, up_d.dtype, up_d.relu, up_d.B, up_d.T, up_d.h, up_d.w, up_d.C = z.ptr(), z.ld, e.dt, 1, B, T0, H // 2, W // 2, Cin
            up_d.u, up_d.ldu, up_d.u_dtype = m.ptr(), m.ld, e.dt
            res["upsample_fwd"] = timed(lambda: e.call("vinet_upsample_fwd", up_d))
        else:
            srcs = srcs_f
        holder = {}
        e.profile = []

        def fprop():
            holder["bwd"] = e.conv("c." + mode, srcs, w, geom, out)
        e.profile = None
        fprop()          # packs the weights
        torch.cuda.synchronize()
        n0 = L.get().fn["vinet_up2_launch_count"]()
        res[mode + " fprop"] = timed(fprop)
        # weight gradient only: no source needs a data gradient
        for s in srcs:
            s.needs_grad = False

        def wgrad():
            holder["bwd"](dy.data_ptr(), out.C)
        wgrad(); e.unpack_flush(); torch.cuda.synchronize()
        res[mode + " wgrad"] = timed(wgrad)
        e.unpack_flush()
        res[mode + " up2 launches"] = L.get().fn["vinet_up2_launch_count"]() - n0
    print("%-11s %7.1f GFLOP | " % (name, gflop) + " | ".join("%s %.3f" % (k, v) if isinstance(v, float) else "%s %d" % (k, v) for k, v in res.items()), flush=True)
