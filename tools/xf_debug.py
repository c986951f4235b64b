"""Development aid: run the fusion token block on the GPU and on the numpy spec, compare every pooled buffer by name."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import scenarios as S
from oracle.kernel_spec import Spec
from vinet_b200 import engine as E

d = int(sys.argv[1]) if len(sys.argv) > 1 else 512
engines = []
orig = S.make_engine


def mk(device, precision, backend=None):
    e = orig(device, precision, backend)
    engines.append(e)
    return e


S.make_engine = mk
if len(sys.argv) > 2:          # xf_debug.py <layers> <p>: the AViNet transformer block with dropout
    engines_ = []
    real_engine = S.Engine

    class Rec(real_engine):
        def __init__(self, *a, **k):
            super().__init__(*a, **k)
            engines.append(self)
    S.Engine = Rec
    got = S.xf_block("cuda", "fp32", layers=d, p=float(sys.argv[2]))
    ref = S.xf_block("cpu", "fp32", Spec(), layers=d, p=float(sys.argv[2]))
else:
    got = S.fusion_block("cuda", "fp32", d=d)
    ref = S.fusion_block("cpu", "fp32", Spec(), d=d)
g, c = engines
for k in c.pool:
    if k in g.pool and not c.pool[k].is_floating_point():
        a, b = g.pool[k].cpu(), c.pool[k]
        print("%-60s differing %d of %d   %s %s" % (k, int((a != b).sum()), b.numel(), a.flatten()[:6].tolist(), b.flatten()[:6].tolist()))
for k in c.pool:
    if k in g.pool and c.pool[k].is_floating_point():
        a, b = g.pool[k].detach().float().cpu(), c.pool[k].detach().float()
        print("%-60s %.3e  (|ref| %.3e)" % (k, ((a - b).norm() / (b.norm() + 1e-30)).item(), b.norm().item()))
for k in sorted(ref):
    a, b = got[k].float(), ref[k].float()
    print("RES %-60s %.3e" % (k, ((a - b).norm() / (b.norm() + 1e-30)).item()))
if len(sys.argv) > 2:
    sys.exit(0)
hk = [k for k in c.pool if k.endswith(".h")][0]
hg, hc = g.pool[hk].cpu(), c.pool[hk]
flips = ((hg > 0) != (hc > 0))
print("ReLU mask flips behind linear1:", int(flips.sum()), "of", flips.numel(), "| |h| at the flips:", torch.maximum(hg, hc)[flips].tolist()[:8])
