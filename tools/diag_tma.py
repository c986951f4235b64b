"""GPU diagnostic: TMA-fed tcgen05 conv kernels (fprop / dgrad / wgrad) vs the fp32 SIMT engine, then a
quick A/B timing of TMA vs register-gather feeding on the big layers of the 32x224x384 workload.

Runs the checks in a subprocess with a timeout so a trap / hang cannot take the caller down.
  python tools/diag_tma.py            (parent)   |   python tools/diag_tma.py child [perf]
"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def build(precision, use_tma, B, T0, T1, H, W, Cin, Cout, k, st, pad, seed=0):
    import torch
    from vinet_b200.engine import ConvGeom, Engine
    dev = torch.device("cuda")
    e = Engine(precision)
    e.use_tma = use_tma
    e.begin(dev, True, True)
    g = torch.Generator().manual_seed(seed)
    srcs = []
    for i, T in enumerate((T0, T1)):
        if T == 0:
            continue
        x = e.new_act("x%d" % i, B, T, H, W, Cin)
        x.buf.copy_(torch.randn(x.buf.shape, generator=g).to(torch.bfloat16).float())
        srcs.append(x)
    w = torch.randn(Cout, Cin, *k, generator=g) / (Cin * k[0] * k[1] * k[2]) ** 0.5
    w = w.to(torch.bfloat16).float().to(dev)
    geom = ConvGeom(k, (st, 1, 1), pad)
    To, Ho, Wo = geom.out_dims(T0 + T1, H, W)
    out = e.new_act("o", B, To, Ho, Wo, Cout)
    return e, srcs, w, geom, out, g


def run(precision, use_tma, *case):
    import torch
    e, srcs, w, geom, out, g = build(precision, use_tma, *case)
    bwd = e.conv("c", srcs, w, geom, out)
    torch.cuda.synchronize()
    res = {"out": out.buf.float().cpu()}
    dy = torch.randn(out.buf.shape, generator=g).to(torch.bfloat16).float().cuda().to(e.tdtype)
    for gb in e.grad_bufs:
        gb.zero_()
    bwd(dy.data_ptr(), out.C)
    e.unpack_flush()
    torch.cuda.synchronize()
    res["dW"] = e.param_grads["c.weight"].float().cpu()
    for i, s in enumerate(srcs):
        res["dx%d" % i] = s.grad.float().cpu()
    return res


CASES = [  # B, T0, T1, H, W, Cin, Cout, k, stride_t, pad
    (1, 1, 0, 8, 16, 64, 16, (1, 1, 1), 1, (0, 0, 0)),       # one tile, one k-block
    (1, 2, 0, 16, 16, 256, 256, (1, 1, 1), 1, (0, 0, 0)),    # 4 k-blocks, N=256
    (2, 3, 0, 7, 12, 832, 384, (1, 1, 1), 1, (0, 0, 0)),     # base4 geometry: 84-row boxes, 2 N tiles, 13 k-blocks
    (1, 2, 0, 10, 13, 24, 40, (1, 3, 3), 1, (0, 1, 1)),      # ragged frame, Cin=24 (partial 64-block, 2 MMAs)
    (1, 2, 0, 14, 24, 96, 208, (1, 3, 3), 1, (0, 1, 1)),     # Cin=96: 2 channel blocks per tap
    (2, 4, 0, 14, 24, 208, 208, (3, 1, 1), 1, (1, 0, 0)),    # temporal conv with padding (skipped taps)
    (1, 8, 0, 16, 32, 64, 64, (7, 1, 1), 2, (3, 0, 0)),      # stem conv_t: temporal stride 2, 7 taps
    (2, 1, 2, 6, 5, 16, 24, (3, 3, 3), 3, (0, 1, 1)),        # decoder: T-concat, stride == kernel
    (1, 4, 8, 14, 24, 192, 480, (3, 3, 3), 3, (0, 1, 1)),    # convtsp2-like: concat, 2 N tiles, 81 k-blocks
    (1, 4, 16, 28, 48, 64, 64, (5, 3, 3), 5, (0, 1, 1)),     # convtsp4.0-like
    (1, 2, 0, 56, 96, 64, 192, (1, 3, 3), 1, (0, 1, 1)),     # base1.3.conv_s geometry (32x4 boxes)
    (1, 2, 0, 32, 64, 32, 32, (2, 1, 1), 2, (0, 0, 0)),      # convtsp4.6-like
    (2, 5, 0, 12, 20, 32, 48, (3, 3, 3), 1, (1, 1, 1)),      # full 3x3x3 stride 1: halo + 3 live output frames
    (2, 16, 0, 28, 48, 128, 192, (1, 3, 3), 1, (0, 1, 1)),   # 3c.b1.conv_s: streamed weights, 2 N tiles, many items per CTA
    (2, 16, 0, 28, 48, 192, 192, (3, 1, 1), 1, (1, 0, 0)),   # 3c.b1.conv_t: resident weights, runs of output frames
    (3, 9, 0, 20, 40, 64, 64, (7, 1, 1), 2, (3, 0, 0)),      # stem conv_t with an odd frame count
    (1, 24, 0, 16, 36, 64, 64, (7, 1, 1), 2, (3, 0, 0)),     # stem conv_t, 12 output frames: strided temporal-halo tiles (partial)
    (2, 32, 0, 8, 16, 64, 64, (7, 1, 1), 2, (3, 0, 0)),      # stem conv_t, 16 output frames: one full temporal-halo tile
]


def child_check():
    ok_all = True
    for c in CASES:
        try:
            a = run("bf16", True, *c)
            b = run("fp32", False, *c)
        except Exception as ex:  # noqa
            print("case %s EXC %s" % (c, str(ex)[:300]), flush=True)
            ok_all = False
            break
        line = "case %s:" % (c,)
        for k_ in a:
            ref = b[k_]
            err = (a[k_] - ref).abs()
            scale = ref.abs().max().item() + 1e-20
            rel = err.max().item() / scale
            frac = (err > 2e-2 * scale).float().mean().item()
            line += " %s rel %.2e bad %.4f |" % (k_, rel, frac)
            tol = 2e-3 if k_ == "dW" else 3e-2       # out / dx are stored in bf16; dW is an fp32 sum of bf16 products
            if not rel <= tol:
                ok_all = False
                line += " <-- FAIL"
        print(line, flush=True)
    print("TMA %s" % ("ALL_OK" if ok_all else "FAILED"), flush=True)


PERF = [  # name, case (B=8 layers of the north-star workload)
    ("base1.0.conv_s", None),
    ("base1.3.conv_s", (8, 16, 0, 56, 96, 64, 192, (1, 3, 3), 1, (0, 1, 1))),
    ("base1.3.conv_t", (8, 16, 0, 56, 96, 192, 192, (3, 1, 1), 1, (1, 0, 0))),
    ("base1.0.conv_t", (8, 32, 0, 112, 192, 64, 64, (7, 1, 1), 2, (3, 0, 0))),
    ("3c.b1.conv_s", (8, 16, 0, 28, 48, 128, 192, (1, 3, 3), 1, (0, 1, 1))),
    ("3c.b1.conv_t", (8, 16, 0, 28, 48, 192, 192, (3, 1, 1), 1, (1, 0, 0))),
    ("3c.b0 1x1", (8, 16, 0, 28, 48, 256, 128, (1, 1, 1), 1, (0, 0, 0))),
    ("convtsp1", (8, 4, 0, 7, 12, 1024, 832, (1, 3, 3), 1, (0, 1, 1))),
    ("convtsp2", (8, 4, 8, 14, 24, 832, 480, (3, 3, 3), 3, (0, 1, 1))),
    ("convtsp3", (8, 4, 16, 28, 48, 480, 192, (5, 3, 3), 5, (0, 1, 1))),
    ("convtsp4.0", (8, 4, 16, 56, 96, 192, 64, (5, 3, 3), 5, (0, 1, 1))),
    ("convtsp4.3", (8, 4, 0, 112, 192, 64, 32, (2, 3, 3), 2, (0, 1, 1))),
]


def build_stem(use_tma, B=8, T=32, H=224, W=384):
    import torch
    from vinet_b200 import model as M
    from vinet_b200.engine import ConvGeom, Engine
    e = Engine("bf16")
    e.use_tma = use_tma
    e.begin(torch.device("cuda"), True, True)
    x = torch.randn(B, T, 3, H, W, device="cuda").permute(0, 2, 1, 3, 4)
    xin = M.pack_input(e, x)
    w = (torch.randn(64, 3, 1, 7, 7, device="cuda") / 147 ** 0.5).to(torch.bfloat16).float()
    geom = ConvGeom((1, 7, 7), (1, 2, 2), (0, 3, 3))
    out = e.new_act("o", B, T, H // 2, W // 2, 64)
    return e, [xin], w, geom, out, None


def child_perf():
    import torch
    flush = torch.empty(192 * 1024 * 1024, dtype=torch.uint8, device="cuda")
    print("%-16s %8s | %21s | %21s | %21s" % ("layer", "GFLOP", "fprop ms (TF/s)", "dgrad ms (TF/s)", "wgrad ms (TF/s)"))
    from vinet_b200 import lib as L
    for name, c in PERF:
        for use_tma in (True, False):      # True: streaming kernel where eligible; False: conv_tma.cu's per-tap boxes
            L.get().call("vinet_debug_set", 2, 1 if use_tma else 0)
            e, srcs, w, geom, out, g = build_stem(True) if c is None else build("bf16", True, *c)
            e.profile = []
            e.l2_flush = flush
            dy = torch.randn(out.buf.shape, device="cuda").to(e.tdtype)
            for it in range(3):
                e.profile = []
                bwd = e.conv("c", srcs, w, geom, out, cin_real=3 if c is None else None)
                e.gwritten = set()          # data gradients as first writers (plain stores), like most layers of the model
                bwd(dy.data_ptr(), out.C)
            torch.cuda.synchronize()
            agg = {}
            for label, kind, flops, e0, e1, _kern in e.profile:
                ms, fl = agg.get(kind, (0.0, 0.0))
                agg[kind] = (ms + e0.elapsed_time(e1), fl + flops)
            gf = agg["fprop"][1] / 1e9
            cells = []
            for kind in ("fprop", "dgrad", "wgrad"):
                ms, fl = agg.get(kind, (0.0, 0.0))
                cells.append("%8.3f (%7.1f)" % (ms, fl / 1e9 / ms if ms > 0 else 0.0))
            print("%-16s %8.1f | %s  [%s]" % (name, gf, " | ".join(cells), "stream" if use_tma else "per-tap"), flush=True)
            del e
            torch.cuda.empty_cache()


def main():
    if len(sys.argv) > 1 and sys.argv[1] == "child":
        if len(sys.argv) > 2 and sys.argv[2] == "perf":
            child_perf()
        else:
            child_check()
        return
    for mode in ([], ["perf"]):
        try:
            r = subprocess.run([sys.executable, os.path.abspath(__file__), "child"] + mode, timeout=420,
                               capture_output=True, text=True)
            print(r.stdout[-8000:])
            if r.returncode != 0:
                print("exit %d stderr: %s" % (r.returncode, r.stderr[-2500:]))
        except subprocess.TimeoutExpired as ex:
            print("TIMEOUT", (ex.stdout or b"")[-3000:])
        sys.stdout.flush()


if __name__ == "__main__":
    main()
