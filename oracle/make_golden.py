"""TEST INFRASTRUCTURE ONLY — generate tests/golden/*.npz by executing the UNMODIFIED reference.

Run in the build container (needs /root/reference):   python -m oracle.make_golden
For every case it (1) builds the reference model and this repo's restatement, (2) asserts that
state_dict keys/shapes are identical, (3) loads the same seeded weights into both, (4) runs both
on the same seeded inputs and asserts they agree, and (5) stores the REFERENCE's outputs
(saliency maps, loss scalars, parameter-gradient digests, running statistics) as the golden
vectors.  Inputs/weights are not stored: they are regenerated from seeds by
``oracle.torch_oracle.make_inputs`` / ``randomize_`` and guarded by stored checksums.
"""
import json
import os
import sys

import numpy as np
import torch

from . import ref_loader
from . import torch_oracle as O

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

# name, kind, num_clips, B, H, W, train, seed
CASES = [
    ("vinet_t8_train", "vinet", 8, 2, 64, 96, True, 0),
    ("vinet_t16_eval", "vinet", 16, 1, 64, 64, False, 1),
    ("vinet_t32_train", "vinet", 32, 1, 64, 96, True, 2),
    ("vinet_t32_eval", "vinet", 32, 2, 96, 64, False, 3),
    ("vinet_t48_eval", "vinet", 48, 1, 64, 64, False, 4),
    ("avinet_t32_train", "avinet", 32, 1, 224, 384, True, 5),
    # ablation decoders (--num_hier 0/1/2, model.py:501,564,627): name, kind, T, B, H, W, train, seed, num_hier
    ("vinet_hier0_train", "vinet", 32, 2, 64, 96, True, 6, 0),
    ("vinet_hier1_eval", "vinet", 32, 1, 64, 96, False, 7, 1),
    ("vinet_hier2_train", "vinet", 32, 2, 96, 64, True, 8, 2),
    # transformer fusion variants (model.py:211-221,239-247 and model.py:116-189), train mode with every dropout probability set
    # to 0 on the reference at run time (torch's Philox stream is not part of the contract; see oracle.torch_oracle.set_dropout)
    ("avinet_xf_train", "avinet_xf", 32, 1, 224, 384, True, 9),
    ("fusion_train", "fusion", 32, 1, 224, 384, True, 10),
]


def checksum(t):
    t = t.detach().double()
    return [float(t.sum()), float(t.abs().sum())]


def grad_digest(model):
    """Per-parameter (L2 norm, sum, value at 3 fixed flat indices) — small but position-sensitive."""
    out = {}
    for k, p in model.named_parameters():
        if p.grad is None:
            out[k] = None
            continue
        g = p.grad.detach().double().flatten()
        n = g.numel()
        idx = [0, n // 3, n - 1]
        out[k] = [float(g.norm()), float(g.sum())] + [float(g[i]) for i in idx]
    return out


def run_case(name, kind, T, B, H, W, train, seed, num_hier=3):
    torch.manual_seed(0)
    if kind == "vinet":
        ref = ref_loader.build_vinet(T, num_hier)
        mine = O.ViNetOracle(T, num_hier)
    elif kind == "avinet":
        ref = ref_loader.build_avinet()
        mine = O.AViNetOracle(T)
    elif kind == "avinet_xf":
        ref = O.set_dropout(ref_loader.build_avinet(use_transformer=True), 0.0)
        mine = O.set_dropout(O.AViNetOracle(T, use_transformer=True), 0.0)
    else:
        ref = O.set_dropout(ref_loader.build_fusion(), 0.0)
        mine = O.set_dropout(O.AVFusionOracle(num_clips=T), 0.0)
    rs, ms = ref.state_dict(), mine.state_dict()
    assert list(rs.keys()) == list(ms.keys()), "state_dict key order differs from the reference"
    for k in rs:
        assert rs[k].shape == ms[k].shape and rs[k].dtype == ms[k].dtype, k
    O.randomize_(mine, seed)
    ref.load_state_dict(mine.state_dict())
    d = O.make_inputs(B, T, H, W, seed, audio=(kind != "vinet"))
    args = (d["x"],) if kind == "vinet" else (d["x"], d["audio"])
    _, ref_loss = ref_loader.load()
    rec = {}
    meta = {"kind": kind, "T": T, "B": B, "H": H, "W": W, "train": train, "seed": seed, "num_hier": num_hier,
            "x_checksum": checksum(d["x"]), "gt_checksum": checksum(d["gt"]),
            "w_checksum": checksum(torch.cat([p.detach().flatten() for p in mine.parameters()])),
            "keys": list(rs.keys()), "shapes": [list(v.shape) for v in rs.values()]}
    if train:
        ref.train(); mine.train()
        for p in list(ref.parameters()) + list(mine.parameters()):
            p.grad = None
        pr = ref(*args); pm = mine(*args)
        lr = ref_loss.kldiv(pr, d["gt"]); lm = O.kldiv(pm, d["gt"])
        lr.backward(); lm.backward()
        assert torch.allclose(pr, pm, rtol=0, atol=1e-6), (pr - pm).abs().max()
        assert abs(lr.item() - lm.item()) <= 1e-6 * abs(lr.item())
        gr, gm = grad_digest(ref), grad_digest(mine)
        for k in gr:
            assert (gr[k] is None) == (gm[k] is None), k
            if gr[k] is not None:
                assert abs(gr[k][0] - gm[k][0]) <= 1e-4 * abs(gr[k][0]) + 1e-9, (k, gr[k], gm[k])
        rec["pred"] = pr.detach().numpy()
        rec["loss_kldiv"] = np.float64(lr.item())
        meta["grad_digest"] = gr
        # a few complete gradient tensors + BN running stats after the step
        sd = ref.state_dict()
        pfx = "" if kind == "vinet" else "visual_model."
        full = [pfx + "backbone.base1.0.conv_s.weight", pfx + "backbone.base1.0.bn_s.weight",
                pfx + "backbone.base2.0.branch2.1.conv_t.weight", pfx + "backbone.base4.1.branch3.1.bn.bias",
                pfx + "decoder.convtsp4.3.weight"]
        if kind in ("avinet", "avinet_xf"):
            full += ["bilinear.bias", "audionet.conv1.weight", "audionet.batchnorm7.weight"]
        if kind == "avinet_xf":
            full += ["conv_in_1x1.weight", "conv_out_1x1.bias", "transformer.transformer_encoder.layers.0.self_attn.in_proj_bias",
                     "transformer.transformer_encoder.layers.0.self_attn.out_proj.bias",
                     "transformer.transformer_encoder.layers.2.norm2.weight", "transformer.transformer_encoder.layers.1.linear1.bias"]
        if kind == "fusion":
            full += ["audionet.conv1.weight", "conv_in_1x1.bias", "audio_conv_1x1.bias",
                     "transformer.transformer_encoder.layers.0.self_attn.in_proj_bias",
                     "transformer.transformer_encoder.layers.2.norm1.bias", "transformer.transformer_encoder.layers.1.linear2.bias"]
        named = dict(ref.named_parameters())
        for k in full:
            rec["grad/" + k] = named[k].grad.detach().numpy()
        for k in [pfx + "backbone.base1.0.bn_s.running_mean", pfx + "backbone.base1.0.bn_s.running_var",
                  pfx + "backbone.base3.2.branch1.1.bn_t.running_var"]:
            rec["stat/" + k] = sd[k].numpy()
    else:
        ref.eval(); mine.eval()
        with torch.no_grad():
            pr = ref(*args); pm = mine(*args)
        assert torch.allclose(pr, pm, rtol=0, atol=1e-6), (pr - pm).abs().max()
        rec["pred"] = pr.numpy()
        rec["loss_kldiv"] = np.float64(ref_loss.kldiv(pr, d["gt"]).item())
    meta["pred_stats"] = [float(pr.min()), float(pr.max()), float(pr.mean()), float(pr.std())]
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **rec)
    with open(os.path.join(OUT, name + ".json"), "w") as f:
        json.dump(meta, f)
    print(name, "ok", meta["pred_stats"], float(rec["loss_kldiv"]))


def run_losses():
    """Loss scalars + input gradients of the reference's loss.py on seeded maps."""
    _, L = ref_loader.load()
    rec = {}
    for tag in O.LOSS_CASES:
        s, gt, fix = O.make_loss_inputs(tag)
        s.requires_grad_(True)
        rec[f"{tag}/checksum"] = np.array(checksum(s) + checksum(gt) + checksum(fix))
        for nm, fn, mine, tgt in [("kldiv", L.kldiv, O.kldiv, gt), ("cc", L.cc, O.cc, gt),
                                  ("sim", L.similarity, O.similarity, gt), ("nss", L.nss, O.nss, fix)]:
            v = fn(s, tgt)
            (gr,) = torch.autograd.grad(v, s)
            v2 = mine(s, tgt)
            (gr2,) = torch.autograd.grad(v2, s)
            assert abs(v.item() - v2.item()) <= 1e-6 * abs(v.item()) + 1e-9, (nm, v.item(), v2.item())
            assert torch.allclose(gr, gr2, rtol=1e-4, atol=1e-9), nm
            rec[f"{tag}/{nm}"] = np.float64(v.item())
            rec[f"{tag}/{nm}_grad"] = gr.numpy() if tag == "a" else gr.numpy()[:, ::7, ::5].copy()
    np.savez_compressed(os.path.join(OUT, "losses.npz"), **rec)
    print("losses ok", {k: float(v) for k, v in rec.items() if v.ndim == 0})


def main():
    if not ref_loader.available():
        sys.exit("reference not available: goldens can only be regenerated in the build container")
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(os.cpu_count())
    run_losses()
    only = sys.argv[1:]
    for c in CASES:
        if not only or c[0] in only:
            run_case(*c)


if __name__ == "__main__":
    main()
