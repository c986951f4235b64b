"""Run one conv layer of the workload (fprop + dgrad + wgrad) a few times: the target of `ncu --set full`."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import torch
import diag_tma as D

name = sys.argv[1] if len(sys.argv) > 1 else "base1.3.conv_s"
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 3
case = dict(D.PERF)[name]
e, srcs, w, geom, out, g = D.build_stem(True) if case is None else D.build("bf16", True, *case)
dy = torch.randn(out.buf.shape, device="cuda").to(e.tdtype)
for it in range(iters):
    bwd = e.conv("c", srcs, w, geom, out, cin_real=3 if case is None else None)
    bwd(dy.data_ptr(), out.C)
    e.unpack_flush()
torch.cuda.synchronize()
print("done", name)
