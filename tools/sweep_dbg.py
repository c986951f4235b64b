import sys, os
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import torch
from oracle import torch_oracle as O
from vinet_b200 import VideoSaliencyModel, kldiv, lib as L
def build(T, prec):
    ref = O.ViNetOracle(T); O.randomize_(ref, 31)
    m = VideoSaliencyModel(num_clips=T); m.load_state_dict(ref.state_dict())
    return m.cuda().set_precision(prec)
for shape in [(8, 2, 128, 192), (32, 3, 160, 224)]:
    T, B, H, W = shape
    d = O.make_inputs(B, T, H, W, 31); x, gt = d["x"].cuda(), d["gt"].cuda()
    out = {}
    for tag, prec, key2 in [("stream", "bf16", 1), ("pertap", "bf16", 0), ("simt", "bf16_simt", 1)]:
        L.get().call("vinet_debug_set", 2, key2)
        m = build(T, prec).train()
        p = m(x); l = kldiv(p, gt); l.backward()
        out[tag] = (p.detach().float().cpu(), float(l.detach()), {n: q.grad.float().cpu() for n, q in m.named_parameters()})
        del m
    L.get().call("vinet_debug_set", 2, 1)
    for a in ("stream", "pertap"):
        pa, la, ga = out[a]; pb, lb, gb = out["simt"]
        errs = sorted((float((ga[n] - gb[n]).norm() / (gb[n].norm() + 1e-30)), n) for n in ga if n.startswith("decoder."))
        print(shape, a, "pred maxdiff %.3e (max %.3f) loss %.6f vs %.6f  worst decoder grad" % ((pa - pb).abs().max(), pb.abs().max(), la, lb), errs[-1])
    pa, pb = out["stream"][0], out["pertap"][0]
    print(shape, "stream vs pertap pred maxdiff %.3e" % (pa - pb).abs().max())
