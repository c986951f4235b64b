"""GPU: the TENSOR-CORE parity mode — the very tcgen05 / TMA kernels the benchmark measures (conv_stream_kernel,
conv_wgrad_halo_kernel, conv_gemm_tma_kernel, conv_wgrad_tma_kernel), gated at the north-star tolerances against the
executed reference's golden vectors: saliency maps 1e-3 relative, loss scalars 1e-5 relative.

`set_precision("bf16x6")` keeps activations / gradients in fp32 and issues every convolution GEMM as the sum of tcgen05 launches
over the bf16 expansions of both operands (x = x0 + x1 + x2, w = w0 + w1 + w2, the six terms with i + j <= 2), accumulated in
fp32 — see include/vinet_b200.h (vinet_split_bf16).  "bf16x3" (two-term expansion, three launches) is the cheaper variant; its
2^-16 truncation is reported here but only gated loosely.  Reference entry: /root/reference/model.py:103-112 (forward),
loss.py:13-38 (kldiv).
"""
import copy
import glob
import json
import os

import numpy as np
import pytest
import torch

from oracle import torch_oracle as O
from vinet_b200 import VideoSaliencyModel, kldiv

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")
CASES = sorted(os.path.basename(p)[:-5] for p in glob.glob(os.path.join(GOLD, "vinet_*.json")))
TC_KERNELS = {"conv_stream_kernel", "conv_wgrad_halo_kernel", "conv_gemm_tma_kernel", "conv_wgrad_tma_kernel"}


def _build(meta, precision):
    ref = O.ViNetOracle(meta["T"], meta.get("num_hier", 3))
    O.randomize_(ref, meta["seed"])
    m = VideoSaliencyModel(num_clips=meta["T"], num_hier=meta.get("num_hier", 3))
    m.load_state_dict(ref.state_dict())
    return ref, m.cuda().set_precision(precision)


def _run(m, x, gt, train):
    """forward (+ kldiv + backward); also returns the set of CUDA kernels that served the conv launches."""
    eng = m._engine_for(x.device)
    eng.profile = []
    if train:
        m.train()
        pred = m(x)
        loss = kldiv(pred, gt)
        loss.backward()
    else:
        m.eval()
        with torch.no_grad():
            pred = m(x)
        loss = kldiv(pred, gt)
    torch.cuda.synchronize()
    kernels = {r[5] for r in eng.profile}
    eng.profile = None
    return pred.detach(), loss.detach(), kernels


@pytest.mark.parametrize("name", CASES)
def test_tensorcore_parity_mode_matches_reference_golden(name):
    meta = json.load(open(os.path.join(GOLD, name + ".json")))
    z = np.load(os.path.join(GOLD, name + ".npz"))
    _, m = _build(meta, "bf16x6")
    d = O.make_inputs(meta["B"], meta["T"], meta["H"], meta["W"], meta["seed"])
    x, gt = d["x"].cuda(), d["gt"].cuda()
    pred, loss, kernels = _run(m, x, gt, meta["train"])
    # every convolution ran on a tcgen05 kernel (no FFMA engine, no register-gather fallback)
    assert kernels and kernels <= TC_KERNELS, kernels
    assert "conv_stream_kernel" in kernels and "conv_gemm_tma_kernel" in kernels
    if meta["train"]:
        assert "conv_wgrad_halo_kernel" in kernels
    p = pred.cpu().numpy()
    want = float(z["loss_kldiv"])
    rel = np.abs(p - z["pred"]) / np.abs(z["pred"])
    print("%s bf16x6: map max-rel %.3e, kldiv %.8f vs %.8f (rel %.2e)" % (name, rel.max(), loss.item(), want, abs(loss.item() - want) / want))
    assert np.allclose(p, z["pred"], rtol=1e-3, atol=1e-6), rel.max()
    assert abs(loss.item() - want) <= 1e-5 * abs(want), (loss.item(), want)
    if not meta["train"]:
        return
    named = dict(m.named_parameters())
    errs = []
    for k, dig in meta["grad_digest"].items():
        g = named[k].grad
        assert g is not None, k
        errs.append(abs(float(g.double().norm()) - dig[0]) / (dig[0] + 1e-30))
    assert np.median(errs) < 3e-2 and max(errs) < 2e-1, (np.median(errs), max(errs))
    # full gradient tensors the goldens hold: the decoder tightly, the backbone at the level two correct fp32
    # implementations agree through ~60 train-mode BatchNorm layers with tiny batches (tests/test_plan_cpu.py)
    worst_bb = 0.0
    for k in z.files:
        if not k.startswith("grad/"):
            continue
        g = named[k[5:]].grad.cpu().numpy()
        rl2 = np.linalg.norm(g - z[k]) / (np.linalg.norm(z[k]) + 1e-30)
        if k.startswith("grad/decoder"):
            assert rl2 < 1e-2, (k, rl2)
        else:
            worst_bb = max(worst_bb, rl2)
            assert rl2 < 1e-1, (k, rl2)
    print("%s bf16x6: worst backbone gradient rel-L2 vs golden %.3e" % (name, worst_bb))
    for k in z.files:
        if k.startswith("stat/"):
            assert np.allclose(m.state_dict()[k[5:]].cpu().numpy(), z[k], rtol=1e-3, atol=1e-5), k


def test_tensorcore_parity_full_size_clip():
    """One 32x224x384 clip (BASELINE.json's shape), train mode, forward + kldiv + backward on the tcgen05 kernels in the
    split-precision mode against the PyTorch oracle in strict fp32 on the same GPU: map 1e-3, loss 1e-5; gradients against an
    fp64 oracle with stock fp32 PyTorch as the yardstick."""
    T, B, H, W = 32, 1, 224, 384
    meta = {"T": T, "seed": 21}
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        ref, m = _build(meta, "bf16x6")
        d = O.make_inputs(B, T, H, W, 21)
        x, gt = d["x"].cuda(), d["gt"].cuda()
        ref = ref.cuda().train()
        ref64 = copy.deepcopy(ref).double()
        pr = ref(x); lr = O.kldiv(pr, gt); lr.backward()
        p64 = ref64(x.double()); O.kldiv(p64, gt.double()).backward()
        pm, lm, kernels = _run(m, x, gt, True)
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old
    assert kernels <= TC_KERNELS and {"conv_stream_kernel", "conv_wgrad_halo_kernel"} <= kernels, kernels
    rel = ((pm - pr).abs() / pr.abs()).max().item()
    print("full-size bf16x6: map max-rel %.3e vs torch fp32, kldiv %.8f vs %.8f" % (rel, lm.item(), lr.item()))
    assert torch.allclose(pm, pr, rtol=1e-3, atol=1e-6), rel
    assert abs(lm.item() - lr.item()) <= 1e-5 * abs(lr.item()), (lm.item(), lr.item())
    rp, r64 = dict(ref.named_parameters()), dict(ref64.named_parameters())
    e_mine = sorted(float((q.grad.double() - r64[n].grad).norm() / (r64[n].grad.norm() + 1e-30)) for n, q in m.named_parameters())
    e_ref = sorted(float((rp[n].grad.double() - r64[n].grad).norm() / (r64[n].grad.norm() + 1e-30)) for n, _ in m.named_parameters())
    print("full-size bf16x6 grad rel-L2 vs fp64: engine median %.3e max %.3e ; torch fp32 median %.3e max %.3e"
          % (e_mine[len(e_mine) // 2], e_mine[-1], e_ref[len(e_ref) // 2], e_ref[-1]))
    assert e_mine[len(e_mine) // 2] <= 2 * e_ref[len(e_ref) // 2] + 2e-3
    errs = sorted((float((q.grad - rp[n].grad).norm() / (rp[n].grad.norm() + 1e-30)), n) for n, q in m.named_parameters())
    assert all(e < 2e-3 for e, n in errs if n.startswith("decoder.convtsp4")), [e for e in errs if e[1].startswith("decoder.")][-3:]


def test_two_term_split_is_close_but_reported_separately():
    """bf16x3 (three launches per GEMM): truncation 2^-16 per product.  Reported, gated loosely (map 1e-2, loss 1e-3): the
    parity claim rests on bf16x6."""
    name = "vinet_t32_train"
    meta = json.load(open(os.path.join(GOLD, name + ".json")))
    z = np.load(os.path.join(GOLD, name + ".npz"))
    _, m = _build(meta, "bf16x3")
    d = O.make_inputs(meta["B"], meta["T"], meta["H"], meta["W"], meta["seed"])
    pred, loss, kernels = _run(m, d["x"].cuda(), d["gt"].cuda(), True)
    assert kernels <= TC_KERNELS
    p = pred.cpu().numpy()
    want = float(z["loss_kldiv"])
    rel = (np.abs(p - z["pred"]) / np.abs(z["pred"])).max()
    print("%s bf16x3: map max-rel %.3e, kldiv rel %.2e" % (name, rel, abs(loss.item() - want) / want))
    assert rel < 1e-2 and abs(loss.item() - want) <= 1e-3 * want


def test_avinet_tensorcore_parity_mode_matches_reference_golden():
    from vinet_b200 import VideoAudioSaliencyModel
    name = "avinet_t32_train"
    meta = json.load(open(os.path.join(GOLD, name + ".json")))
    z = np.load(os.path.join(GOLD, name + ".npz"))
    ref = O.AViNetOracle(meta["T"])
    O.randomize_(ref, meta["seed"])
    m = VideoAudioSaliencyModel(num_clips=meta["T"], soundnet_weights=False)
    m.load_state_dict(ref.state_dict())
    m = m.cuda().set_precision("bf16x6").train()
    d = O.make_inputs(meta["B"], meta["T"], meta["H"], meta["W"], meta["seed"], audio=True)
    pred = m(d["x"].cuda(), d["audio"].cuda())
    loss = kldiv(pred, d["gt"].cuda())
    loss.backward()
    p = pred.detach().cpu().numpy()
    want = float(z["loss_kldiv"])
    rel = (np.abs(p - z["pred"]) / np.abs(z["pred"])).max()
    print("avinet bf16x6: map max-rel %.3e, kldiv rel %.2e" % (rel, abs(loss.item() - want) / want))
    assert np.allclose(p, z["pred"], rtol=1e-3, atol=1e-6), rel
    assert abs(loss.item() - want) <= 1e-5 * abs(want), (loss.item(), want)
    named = dict(m.named_parameters())
    k = "grad/visual_model.decoder.convtsp4.3.weight"
    g = named[k[5:]].grad.cpu().numpy()
    assert np.allclose(g, z[k], rtol=5e-3, atol=5e-4 * np.abs(z[k]).max()), k
