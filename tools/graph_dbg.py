import sys
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import torch
from oracle import torch_oracle as O
from vinet_b200 import VideoSaliencyModel, kldiv, GraphedTrainStep
T, B, H, W = 8, 2, 64, 96
d = O.make_inputs(B, T, H, W, 5); x, gt = d["x"].cuda(), d["gt"].cuda()
res = {}
for mode in ("eager", "graph"):
    ref = O.ViNetOracle(T); O.randomize_(ref, 5)
    m = VideoSaliencyModel(num_clips=T); m.load_state_dict(ref.state_dict()); m = m.cuda().set_precision("fp32").train()
    opt = torch.optim.Adam(m.parameters(), lr=1e-4, fused=True, capturable=True)
    if mode == "graph":
        step = GraphedTrainStep(m, kldiv, opt, x, gt)
        for _ in range(3): step(x, gt)
    else:
        for _ in range(3):
            opt.zero_grad(set_to_none=True); l = kldiv(m(x), gt); l.backward(); opt.step()
    torch.cuda.synchronize()
    sd = {k: v.detach().float().cpu().clone() for k, v in m.state_dict().items()}
    m.eval()
    with torch.no_grad():
        p1 = m(x).float().cpu()
    m.set_precision("fp32")      # fresh engine, no caches
    m.eval()
    with torch.no_grad():
        p2 = m(x).float().cpu()
    res[mode] = (sd, p1, p2)
sa, sb = res["eager"][0], res["graph"][0]
worst = sorted(((sa[k] - sb[k]).abs().max().item() / (sa[k].abs().max().item() + 1e-12), k) for k in sa)[-5:]
print("worst state diffs", worst)
print("eval same-engine: eager vs graph", (res["eager"][1] - res["graph"][1]).abs().max().item())
print("eval fresh-engine: eager vs graph", (res["eager"][2] - res["graph"][2]).abs().max().item())
print("graph: same-engine vs fresh-engine", (res["graph"][1] - res["graph"][2]).abs().max().item())
print("eager: same-engine vs fresh-engine", (res["eager"][1] - res["eager"][2]).abs().max().item())
