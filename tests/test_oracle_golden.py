"""CPU: the oracle restatement reproduces the golden vectors generated from the executed reference
(oracle/make_golden.py).  This is what pins the oracle (the reference has no tests)."""
import glob
import json
import os

import numpy as np
import pytest
import torch

from oracle import arch as oarch
from oracle import torch_oracle as O
from vinet_b200 import arch as parch

GOLD = os.path.join(os.path.dirname(__file__), "golden")
CASES = sorted(os.path.basename(p)[:-5] for p in glob.glob(os.path.join(GOLD, "*.json")))


def _close(a, b, rtol):
    return abs(a - b) <= rtol * abs(b) + 1e-12


def test_arch_tables_agree():
    assert oarch.MIXED == parch.MIXED and oarch.STAGES == parch.STAGES
    assert oarch.DECODER_HEAD == parch.DECODER_HEAD and oarch.SOUNDNET == parch.SOUNDNET
    for t in (8, 16, 32, 48):
        assert oarch.decoder_tail(t) == parch.decoder_tail(t)
    for h in (0, 1, 2, 3):
        assert oarch.decoder_head(h) == parch.decoder_head(h)
    assert parch.decoder_head(3) == parch.DECODER_HEAD


def test_losses_golden():
    z = np.load(os.path.join(GOLD, "losses.npz"))
    for tag in O.LOSS_CASES:
        s, gt, fix = O.make_loss_inputs(tag)
        cs = z[f"{tag}/checksum"]
        got = [float(s.double().sum()), float(s.double().abs().sum()), float(gt.double().sum()),
               float(gt.double().abs().sum()), float(fix.double().sum()), float(fix.double().abs().sum())]
        assert np.allclose(cs, got, rtol=1e-12), "seeded loss inputs are not reproducible here"
        s.requires_grad_(True)
        for nm, fn, tgt in [("kldiv", O.kldiv, gt), ("cc", O.cc, gt), ("sim", O.similarity, gt), ("nss", O.nss, fix)]:
            v = fn(s, tgt)
            (g,) = torch.autograd.grad(v, s)
            assert _close(v.item(), float(z[f"{tag}/{nm}"]), 1e-5), nm       # north_star: 1e-5 on loss scalars
            ref = z[f"{tag}/{nm}_grad"]
            gg = g.numpy() if tag == "a" else g.numpy()[:, ::7, ::5]
            assert np.allclose(gg, ref, rtol=1e-4, atol=1e-9), nm


@pytest.mark.parametrize("name", CASES)
def test_model_golden(name):
    meta = json.load(open(os.path.join(GOLD, name + ".json")))
    if meta["kind"] != "vinet" and os.environ.get("VINET_FAST_TESTS"):
        pytest.skip("fast mode")
    z = np.load(os.path.join(GOLD, name + ".npz"))
    T, B, H, W, seed = meta["T"], meta["B"], meta["H"], meta["W"], meta["seed"]
    model = {"vinet": lambda: O.ViNetOracle(T, meta.get("num_hier", 3)), "avinet": lambda: O.AViNetOracle(T),
             "avinet_xf": lambda: O.set_dropout(O.AViNetOracle(T, use_transformer=True), 0.0),
             "fusion": lambda: O.set_dropout(O.AVFusionOracle(num_clips=T), 0.0)}[meta["kind"]]()
    sd = model.state_dict()
    assert list(sd.keys()) == meta["keys"]
    assert [list(v.shape) for v in sd.values()] == meta["shapes"]
    O.randomize_(model, seed)
    d = O.make_inputs(B, T, H, W, seed, audio=(meta["kind"] != "vinet"))
    xs = d["x"].double()
    assert np.allclose([float(xs.sum()), float(xs.abs().sum())], meta["x_checksum"], rtol=1e-12)
    w = torch.cat([p.detach().flatten() for p in model.parameters()]).double()
    assert np.allclose([float(w.sum()), float(w.abs().sum())], meta["w_checksum"], rtol=1e-12)
    args = (d["x"],) if meta["kind"] == "vinet" else (d["x"], d["audio"])
    if meta["train"]:
        model.train()
        pred = model(*args)
        loss = O.kldiv(pred, d["gt"])
        loss.backward()
        # north_star tolerances: 1e-3 relative on maps, 1e-5 on loss scalars (the oracle is far inside)
        assert np.allclose(pred.detach().numpy(), z["pred"], rtol=1e-4, atol=1e-6)
        assert _close(loss.item(), float(z["loss_kldiv"]), 1e-5)
        named = dict(model.named_parameters())
        for k, dig in meta["grad_digest"].items():
            if dig is None:
                assert named[k].grad is None, k
                continue
            if k.startswith("audionet.conv") and k.endswith(".bias"):
                continue       # a bias in front of a train-mode BatchNorm: the true gradient is 0, what is stored is summation noise
            g = named[k].grad.double().flatten()
            assert _close(float(g.norm()), dig[0], 2e-3), (k, float(g.norm()), dig[0])
        for k in z.files:
            if k.startswith("grad/"):
                g = named[k[5:]].grad.numpy()
                assert np.allclose(g, z[k], rtol=2e-3, atol=1e-5 * np.abs(z[k]).max()), k
            if k.startswith("stat/"):
                assert np.allclose(model.state_dict()[k[5:]].numpy(), z[k], rtol=1e-4, atol=1e-6), k
    else:
        model.eval()
        with torch.no_grad():
            pred = model(*args)
        assert np.allclose(pred.numpy(), z["pred"], rtol=1e-4, atol=1e-6)
        assert _close(O.kldiv(pred, d["gt"]).item(), float(z["loss_kldiv"]), 1e-5)
