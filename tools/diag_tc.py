"""GPU diagnostic: tcgen05 conv kernels vs the fp32 SIMT engine on plain GEMM-shaped problems.

Each (debug-variant) runs in its own subprocess with a timeout, so a trap / hang in one encoding cannot
poison the others.  Usage:  python tools/diag_tc.py            (parent)   |   python tools/diag_tc.py child <dbg>
"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def child(dbg):
    import torch
    from vinet_b200 import lib as L
    from vinet_b200.engine import Act, ConvGeom, Engine
    lib = L.get()
    lib.call("vinet_debug_set", 0, dbg)
    dev = torch.device("cuda")
    gen = torch.Generator().manual_seed(0)

    def run(precision, B, T, H, W, Cin, Cout, k, seed_w):
        e = Engine(precision)
        e.begin(dev, True, True)
        x = e.new_act("x", B, T, H, W, Cin)
        g = torch.Generator().manual_seed(1)
        x.buf.copy_(torch.randn(x.buf.shape, generator=g).to(torch.bfloat16).float())
        w = (torch.randn(Cout, Cin, *k, generator=torch.Generator().manual_seed(seed_w)) / (Cin * k[0] * k[1] * k[2]) ** 0.5)
        w = w.to(torch.bfloat16).float().to(dev)
        geom = ConvGeom(k, (1, 1, 1), (k[0] // 2, k[1] // 2, k[2] // 2))
        out = e.new_act("o", B, T, H, W, Cout)
        bwd = e.conv("c", [x], w, geom, out)
        torch.cuda.synchronize()
        res = {"out": out.buf.float().cpu()}
        dy = torch.randn(out.buf.shape, generator=g).to(torch.bfloat16).float().to(dev).to(e.tdtype)
        for gb in e.grad_bufs:
            gb.zero_()
        bwd(dy.data_ptr(), Cout)
        e.unpack_flush()
        torch.cuda.synchronize()
        res["dW"] = e.param_grads["c.weight"].float().cpu()
        res["dx"] = x.grad.float().cpu()
        return res

    cases = [  # B,T,H,W,Cin,Cout,k
        (1, 1, 8, 16, 64, 16, (1, 1, 1)),      # M=128 K=64  N=16: one tile, one k-block
        (1, 1, 8, 16, 64, 64, (1, 1, 1)),
        (1, 1, 8, 16, 128, 128, (1, 1, 1)),    # 2 k-blocks
        (1, 2, 16, 16, 256, 256, (1, 1, 1)),   # M=512, 4 k-blocks, N=256
        (1, 1, 10, 13, 24, 40, (1, 3, 3)),     # ragged M, Cin=24, taps
        (2, 4, 14, 24, 192, 480, (3, 3, 3)),   # 2 N tiles, 81 k-blocks
        (1, 2, 28, 48, 16, 32, (1, 3, 3)),
    ]
    ok_all = True
    for c in cases:
        try:
            a = run("bf16", *c, 5)
            b = run("fp32", *c, 5)
        except Exception as ex:  # noqa
            print("dbg=%d case %s EXC %s" % (dbg, c, str(ex)[:200]), flush=True)
            ok_all = False
            break
        line = "dbg=%d case %s:" % (dbg, c)
        for k_ in ("out", "dW", "dx"):
            ref = b[k_]
            err = (a[k_] - ref).abs()
            scale = ref.abs().max().item() + 1e-20
            rel = err.max().item() / scale
            frac = (err > 2e-2 * scale).float().mean().item()
            line += " %s rel %.2e bad %.3f |" % (k_, rel, frac)
            if rel > 3e-2:
                ok_all = False
        print(line, flush=True)
    print("dbg=%d %s" % (dbg, "ALL_OK" if ok_all else "FAILED"), flush=True)


def main():
    if len(sys.argv) > 1 and sys.argv[1] == "child":
        child(int(sys.argv[2]))
        return
    variants = [int(v) for v in sys.argv[1:]] or [0, 1, 4, 2]
    for dbg in variants:
        try:
            r = subprocess.run([sys.executable, os.path.abspath(__file__), "child", str(dbg)], timeout=240,
                               capture_output=True, text=True)
            print(r.stdout[-6000:])
            if r.returncode != 0:
                print("dbg=%d exit %d stderr: %s" % (dbg, r.returncode, r.stderr[-1500:]))
        except subprocess.TimeoutExpired:
            print("dbg=%d TIMEOUT" % dbg)
        sys.stdout.flush()


if __name__ == "__main__":
    main()
