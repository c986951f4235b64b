"""Parity report: every golden case (executed-reference vectors) under each precision mode of the engine; prints the
saliency-map max relative error and the kldiv relative error (north-star gates: 1e-3 / 1e-5)."""
import glob
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

from oracle import torch_oracle as O
from vinet_b200 import VideoSaliencyModel, kldiv

GOLD = os.path.join(ROOT, "tests", "golden")
modes = sys.argv[1].split(",") if len(sys.argv) > 1 else ["fp32", "bf16x3", "bf16x6", "bf16"]
for p in sorted(glob.glob(os.path.join(GOLD, "vinet_*.json"))):
    meta = json.load(open(p))
    z = np.load(p[:-5] + ".npz")
    d = O.make_inputs(meta["B"], meta["T"], meta["H"], meta["W"], meta["seed"])
    x, gt = d["x"].cuda(), d["gt"].cuda()
    for prec in modes:
        ref = O.ViNetOracle(meta["T"], meta.get("num_hier", 3))
        O.randomize_(ref, meta["seed"])
        m = VideoSaliencyModel(num_clips=meta["T"], num_hier=meta.get("num_hier", 3))
        m.load_state_dict(ref.state_dict())
        m = m.cuda().set_precision(prec)
        if meta["train"]:
            m.train()
            pred = m(x)
        else:
            m.eval()
            with torch.no_grad():
                pred = m(x)
        loss = kldiv(pred.detach(), gt).item()
        pr = pred.detach().cpu().numpy()
        rel = np.abs(pr - z["pred"]) / np.abs(z["pred"])
        want = float(z["loss_kldiv"])
        bad = np.unravel_index(rel.argmax(), rel.shape)
        print("%-20s %-7s map max-rel %.3e (mean %.2e; at pred %.3e ref %.3e)  abs-max %.3e  kldiv rel %.2e"
              % (os.path.basename(p)[:-5], prec, rel.max(), rel.mean(), pr[bad], z["pred"][bad], np.abs(pr - z["pred"]).max(), abs(loss - want) / want), flush=True)
        del m
