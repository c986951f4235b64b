"""CPU: the host-side plan (descriptors, slices, tape order, backward formulas) executed with the numpy
kernel specification (oracle/kernel_spec.py) against PyTorch autograd / the oracle."""
import copy

import numpy as np
import pytest
import torch

import scenarios as S
from oracle import torch_oracle as O
from oracle.kernel_spec import Spec
from vinet_b200 import VideoSaliencyModel
from vinet_b200 import loss as PL


def test_conv_concat_relu_upsample_vs_torch():
    for kw in ({}, {"T0": 1, "T1": 4, "kt": 5, "Cin": 24, "Cout": 16, "H": 4, "W": 7}):
        a = S.conv_up("cpu", "fp32", Spec(), **kw)
        b = S.conv_up_torch(**kw)
        assert not S.compare(a, b, 1e-4), S.compare(a, b, 1e-4)


def test_upsample_fused_into_next_conv_vs_torch():
    """a6: the second decoder conv reads the first stage's output through relu + the 2x bilinear up-sampling in its input stage
    (VINET_XF_UP2); border-exact against PyTorch autograd, and identical to the plan that materialises the hi-res tensor."""
    for kw in ({}, {"B": 1, "T0": 2, "T1": 3, "kt": 5, "h": 3, "w": 4, "C0": 8, "C1": 16, "C2": 8}, {"h": 1, "w": 2, "T1": 2}):
        a = S.up_fused("cpu", "fp32", Spec(), **kw)
        assert a["_materialized"] == ["c2"], a["_materialized"]          # only the scenario's explicit output materialisation
        b = S.up_fused_torch(**kw)
        assert not S.compare(a, b, 1e-4), S.compare(a, b, 1e-4)
        c = S.up_fused("cpu", "fp32", Spec(), fuse=False, **kw)
        assert c["_materialized"] == ["c1", "c2"]
        assert not S.compare(a, c, 1e-6), S.compare(a, c, 1e-6)


def test_model_state_dict_layout_matches_oracle():
    for t, h in [(8, 3), (16, 3), (32, 3), (48, 3), (32, 0), (32, 1), (32, 2)]:
        a, b = VideoSaliencyModel(num_clips=t, num_hier=h).state_dict(), O.ViNetOracle(t, h).state_dict()
        assert list(a.keys()) == list(b.keys())
        assert all(a[k].shape == b[k].shape and a[k].dtype == b[k].dtype for k in a)


@pytest.mark.parametrize("num_hier", [0, 1, 2])
def test_ablation_decoders_plan(num_hier):
    """--num_hier 0/1/2 (DecoderConvUpNoHier / 1Hier / 2Hier, /root/reference/model.py:501-688) from the same descriptor-driven
    plan: forward and every parameter gradient against the oracle (skip tensors the decoder does not read still get their
    gradient from the backbone chain)."""
    T, B, H, W = 32, 1, 64, 64
    ref = O.ViNetOracle(T, num_hier)
    O.randomize_(ref, 40 + num_hier)
    ref.train()
    m = VideoSaliencyModel(num_clips=T, num_hier=num_hier)
    m.load_state_dict(ref.state_dict())
    m.set_precision("fp32")
    m.__dict__["_backend"] = Spec()
    m.train()
    d = O.make_inputs(B, T, H, W, 40 + num_hier)
    pr = ref(d["x"]); O.kldiv(pr, d["gt"]).backward()
    pm = m(d["x"]); O.kldiv(pm, d["gt"]).backward()
    assert torch.allclose(pm, pr, rtol=1e-3, atol=1e-6), (pm - pr).abs().max()
    rp = dict(ref.named_parameters())
    for n, q in m.named_parameters():
        assert q.grad is not None, n
        if n.startswith("decoder.convtsp4"):
            assert float((q.grad - rp[n].grad).norm() / (rp[n].grad.norm() + 1e-30)) < 2e-2, n


def _spec_model(T, ref):
    m = VideoSaliencyModel(num_clips=T)
    m.load_state_dict(ref.state_dict())
    m.set_precision("fp32")
    m.__dict__["_backend"] = Spec()
    return m


@pytest.mark.parametrize("T", [8, 16, 32, 48])
def test_full_plan_eval_forward(T):
    ref = O.ViNetOracle(T)
    O.randomize_(ref, T)
    ref.eval()
    m = _spec_model(T, ref).eval()
    d = O.make_inputs(1, T, 64, 64, T)
    with torch.no_grad():
        pr, pm = ref(d["x"]), m(d["x"])
    assert pm.shape == pr.shape == (1, 64, 64)
    assert torch.allclose(pm, pr, rtol=1e-3, atol=1e-5), (pm - pr).abs().max()


def test_full_plan_train_forward_backward():
    """Gradients: this random-weight net with tiny BatchNorm batches is chaotic in fp32 (one ReLU flip moves
    every upstream gradient; the fp32 PyTorch oracle itself sits 0.4-5e-2 away from an fp64 run, and which
    layer flips first differs between two correct fp32 implementations).  The yardstick is therefore the
    fp64 oracle with a bar at the fp32 oracle's own worst-case distance; exact per-op checks live in the
    scenario tests above."""
    T, B, H, W = 8, 2, 64, 96
    ref = O.ViNetOracle(T)
    O.randomize_(ref, 0)
    ref.train()
    ref64 = copy.deepcopy(ref).double()
    m = _spec_model(T, ref).train()
    d = O.make_inputs(B, T, H, W, 0)
    p32 = ref(d["x"]); O.kldiv(p32, d["gt"]).backward()
    p64 = ref64(d["x"].double()); O.kldiv(p64, d["gt"].double()).backward()
    pm = m(d["x"]); lm = O.kldiv(pm, d["gt"]); lm.backward()
    assert (pm - p64).abs().max() <= 3 * (p32 - p64).abs().max() + 1e-6
    r64, r32 = dict(ref64.named_parameters()), dict(ref.named_parameters())
    e32, em = [], []
    for n, q in m.named_parameters():
        assert q.grad is not None, n
        g64 = r64[n].grad
        e32.append(float((r32[n].grad - g64).norm() / g64.norm()))
        em.append(float((q.grad - g64).norm() / g64.norm()))
    assert np.median(em) <= max(3 * np.median(e32), 3e-2), (np.median(em), np.median(e32))
    assert max(em) <= max(3 * max(e32), 6e-2), (max(em), max(e32))
    # the decoder (closest to the loss, before any chaos can build up) must agree tightly
    for (n, q), a, b in zip(m.named_parameters(), e32, em):
        if n.startswith("decoder.convtsp4"):
            assert b <= 2 * a + 1e-5, (n, a, b)
    # BatchNorm running statistics and counters follow nn.BatchNorm3d semantics
    sr, sm = ref.state_dict(), m.state_dict()
    for k in sr:
        if "running" in k:
            assert torch.allclose(sm[k], sr[k], rtol=1e-4, atol=1e-6), k
        if "num_batches" in k:
            assert int(sm[k]) == int(sr[k]) == 1


def test_loss_wrappers_shapes_and_values():
    PL._backend = Spec()
    try:
        s, gt, fix = O.make_loss_inputs("a")
        s.requires_grad_(True)
        for mine, ref, tgt in [(PL.kldiv, O.kldiv, gt), (PL.cc, O.cc, gt), (PL.similarity, O.similarity, gt), (PL.nss, O.nss, fix)]:
            v = mine(s, tgt)
            assert v.shape == ()
            (g,) = torch.autograd.grad(v, s)
            v2 = ref(s, tgt)
            (g2,) = torch.autograd.grad(v2, s)
            assert abs(v.item() - v2.item()) <= 1e-5 * abs(v2.item()) + 1e-9
            assert torch.allclose(g, g2, rtol=1e-4, atol=1e-9)

        class A:
            kldiv, cc, sim, l1 = True, True, False, False
            kldiv_coeff, cc_coeff, sim_coeff, batch_size = 1.0, -1.0, -1.0, 3
        out = PL.loss_func(s, gt, A)
        assert out.shape == (1,)
        assert abs(out.item() - (O.kldiv(s, gt) - O.cc(s, gt)).item()) < 1e-5
    finally:
        PL._backend = None


def test_packed_weights_follow_a_fused_optimizer_step():
    """torch.optim.Adam(fused=True) updates parameters WITHOUT bumping Tensor._version, so the engine's cache of packed conv
    weights must not key on the version alone: after an optimizer step the same engine has to compute with the NEW weights
    (what a fresh engine computes), both in the next training forward and in an eval forward."""
    T = 8
    ref = O.ViNetOracle(T)
    O.randomize_(ref, 3)
    m = _spec_model(T, ref).train()
    d = O.make_inputs(1, T, 64, 64, 3)
    opt = torch.optim.Adam(m.parameters(), lr=1e-2, fused=True)
    v0 = m.decoder.convtsp1[0].weight._version
    loss = O.kldiv(m(d["x"]), d["gt"])
    loss.backward()
    opt.step()
    assert m.decoder.convtsp1[0].weight._version == v0, "this torch bumps versions in fused Adam: the test no longer bites"
    m.eval()
    with torch.no_grad():
        same_engine = m(d["x"])
    fresh = _spec_model(T, m).eval()          # same (updated) parameters, empty caches
    with torch.no_grad():
        want = fresh(d["x"])
    assert torch.allclose(same_engine, want, rtol=1e-5, atol=1e-7), (same_engine - want).abs().max()


def test_weight_cache_follows_writes_behind_autograd():
    """`p.data.copy_()` bumps neither Tensor._version nor the optimizer epoch; `invalidate_weight_cache()` (called by
    broadcast_parameters) must make the next forward use the new weights, and `param.data = new` (moved storage) is detected
    by the cache stamp itself."""
    T = 8
    ref = O.ViNetOracle(T)
    O.randomize_(ref, 4)
    m = _spec_model(T, ref).eval()
    d = O.make_inputs(1, T, 64, 64, 4)
    other = O.ViNetOracle(T)
    O.randomize_(other, 9)
    with torch.no_grad():
        first = m(d["x"])
        for (n, p), (_, q) in zip(m.named_parameters(), other.named_parameters()):
            p.data.copy_(q)
        for (n, p), (_, q) in zip(m.named_buffers(), other.named_buffers()):
            p.data.copy_(q)
        m.invalidate_weight_cache()
        second = m(d["x"])
        want = _spec_model(T, other).eval()(d["x"])
    assert (first - second).abs().max() > 1e-3
    assert torch.allclose(second, want, rtol=1e-5, atol=1e-7), (second - want).abs().max()
    # moved storage: no explicit invalidation needed
    with torch.no_grad():
        for p, q in zip(m.parameters(), ref.parameters()):
            p.data = q.detach().clone()
        for p, q in zip(m.buffers(), ref.buffers()):
            p.data = q.detach().clone()
        third = m(d["x"])
    assert torch.allclose(third, first, rtol=1e-5, atol=1e-7), (third - first).abs().max()


def test_grad_arena_accumulates_and_second_backward_raises():
    """with the flat gradient arena, param.grad aliases the arena after the first backward; a second forward/backward
    without zero_grad must ACCUMULATE (g1 + g2) exactly like plain autograd, and a second backward through the same forward
    must raise instead of returning partial gradients."""
    T = 8
    ref = O.ViNetOracle(T)
    O.randomize_(ref, 6)
    d1, d2 = O.make_inputs(1, T, 64, 64, 6), O.make_inputs(1, T, 64, 64, 7)
    grads = {}
    for arena in (False, True):
        m = _spec_model(T, ref).train()
        if arena:
            m.enable_grad_arena()
        for d in (d1, d2):
            loss = O.kldiv(m(d["x"]), d["gt"])
            loss.backward()
        grads[arena] = {n: p.grad.clone() for n, p in m.named_parameters()}
    for n in grads[False]:
        a, b = grads[False][n], grads[True][n]
        assert torch.allclose(a, b, rtol=1e-4, atol=1e-6 * float(a.abs().max()) + 1e-12), n
    m = _spec_model(T, ref).train()
    out = m(d1["x"])
    loss = O.kldiv(out, d1["gt"])
    loss.backward(retain_graph=True)
    with pytest.raises(RuntimeError):
        loss.backward()


def test_split_precision_plan_tracks_fp64():
    """Host logic of the tensor-core parity mode ("bf16x6": operand splitting, term order, accumulate masks of the data gradient,
    shared packed weight gradient) on the numpy kernel spec: forward and gradients sit as close to an fp64 oracle run as the
    fp32 plan does."""
    T, B, H, W = 8, 1, 64, 64
    ref = O.ViNetOracle(T)
    O.randomize_(ref, 0)
    ref.train()
    ref64 = copy.deepcopy(ref).double()
    d = O.make_inputs(B, T, H, W, 0)
    p64 = ref64(d["x"].double()); O.kldiv(p64, d["gt"].double()).backward()
    r64 = dict(ref64.named_parameters())
    errs = {}
    for prec in ("fp32", "bf16x6"):
        m = _spec_model(T, ref).set_precision(prec)
        m.__dict__["_backend"] = Spec()
        m.train()
        pm = m(d["x"]); O.kldiv(pm, d["gt"]).backward()
        eg = [float((q.grad - r64[n].grad).norm() / (r64[n].grad.norm() + 1e-30)) for n, q in m.named_parameters()]
        errs[prec] = (float(((pm.detach() - p64.detach()).abs() / p64.detach().abs()).max()), float(np.median(eg)))
    assert errs["bf16x6"][0] <= 3 * errs["fp32"][0] + 1e-5, errs
    assert errs["bf16x6"][1] <= 3 * errs["fp32"][1] + 1e-3, errs


# ----------------------------------------------------------------------------- transformer fusion variants
def test_transformer_variants_state_dict_layout_matches_oracle():
    """Same keys, key ORDER, shapes and dtypes as the reference (the oracle's layout is asserted equal to the executed reference's
    by oracle/make_golden.py): `load_state_dict` of a reference checkpoint works unchanged."""
    from vinet_b200 import VideoAudioSaliencyFusionModel, VideoAudioSaliencyModel
    pairs = [(VideoAudioSaliencyModel(use_transformer=True, soundnet_weights=False), O.AViNetOracle(32, use_transformer=True)),
             (VideoAudioSaliencyFusionModel(soundnet_weights=False), O.AVFusionOracle())]
    for mine, ref in pairs:
        a, b = mine.state_dict(), ref.state_dict()
        assert list(a.keys()) == list(b.keys())
        assert all(a[k].shape == b[k].shape and a[k].dtype == b[k].dtype for k in a)
        assert torch.equal(a["transformer.pos_encoder.pe"], b["transformer.pos_encoder.pe"])
        mine.load_state_dict(b)


def test_avinet_transformer_block_plan_vs_torch():
    """use_transformer=True (model.py:239-247): conv_in_1x1 -> 32 channel tokens of 336 features -> encoder -> conv_out_1x1; forward,
    input gradient and every parameter gradient against PyTorch autograd (dropout 0)."""
    a, b = S.xf_block("cpu", "fp32", Spec()), S.xf_block_torch()
    assert set(a) == set(b)
    assert not S.compare(a, b, 1e-4), S.compare(a, b, 1e-4)


def test_fusion_model_token_block_plan_vs_torch():
    """VideoAudioSaliencyFusionModel (model.py:151-183): visual + audio tokens, encoder, [tokens | mean audio token] decoder input,
    with a pending BatchNorm + ReLU on the backbone feature (read transform of the token GEMM)."""
    a, b = S.fusion_block("cpu", "fp32", Spec()), S.fusion_block_torch()
    assert set(a) == set(b)
    assert not S.compare(a, b, 1e-4), S.compare(a, b, 1e-4)


def test_transformer_dropout_plan():
    """Train-mode dropout (p = 0.1 like nn.TransformerEncoderLayer): masks keep ~90 %, kept values are scaled by 1/(1-p), every
    forward draws fresh masks from the device-side counter, eval mode launches none, and the backward stays consistent (finite)."""
    from oracle.kernel_spec import dropout_keep
    keep = dropout_keep(200000, 0.1, 1234, 7, 3)
    assert abs(keep.mean() - 0.9) < 5e-3
    assert (dropout_keep(200000, 0.1, 1234, 8, 3) != keep).mean() > 0.1          # another step: another mask
    assert (dropout_keep(200000, 0.1, 1234, 7, 4) != keep).mean() > 0.1          # another site: another mask
    base = S.xf_block("cpu", "fp32", Spec(), layers=1)
    d1 = S.xf_block("cpu", "fp32", Spec(), layers=1, p=0.1)
    ev = S.xf_block("cpu", "fp32", Spec(), layers=1, p=0.1, training=False)
    assert all(torch.isfinite(v).all() for v in d1.values())
    assert (d1["out"] - base["out"]).abs().max() > 1e-3                           # dropout changed the result ...
    assert not S.compare(ev, base, 1e-6, skip=()), "eval mode must not drop"       # ... and is off in eval mode
    x = torch.arange(1, 1001, dtype=torch.float32)
    y, mask = torch.empty(1000), torch.empty(1000, dtype=torch.uint8)
    rng = torch.tensor([5, 0], dtype=torch.int64)
    sp = Spec()
    sp.call("vinet_rng_advance", rng.data_ptr(), None)
    assert rng.tolist() == [5, 1]
    sp.call("vinet_dropout_fwd", x.data_ptr(), y.data_ptr(), mask.data_ptr(), 1000, 0.25, rng.data_ptr(), 2, None)
    assert torch.allclose(y, torch.where(mask.bool(), x / 0.75, torch.zeros(())), rtol=1e-6, atol=0) and 0.6 < mask.float().mean() < 0.9
    g, relu_ref = torch.ones(1000), x - 500                    # (kept alive: the call takes raw addresses)
    sp.call("vinet_dropout_bwd", g.data_ptr(), g.data_ptr(), mask.data_ptr(), relu_ref.data_ptr(), 1000, 0.25, None)
    assert torch.allclose(g, torch.where(mask.bool() & (x > 500), torch.full((), 1 / 0.75), torch.zeros(())), rtol=1e-6, atol=0)


def test_only_gradient_gemms_are_order_free():
    """vinet_bgemm_t.accumulate bit 1 (the library may split K across CTAs and combine with atomics) is set on the products of the
    backward pass only: every forward product of the transformer block (they all carry a bias term, or are batched over
    clips x heads) keeps a fixed reduction order, so forward results stay bit-reproducible."""
    seen = []

    class Logging(Spec):
        def bgemm(self, d, stream):
            seen.append((int(d.accumulate), bool(d.bias1) or bool(d.bias2), d.nb1 * d.nb2, int(d.relu)))
            Spec.bgemm(self, d, stream)
    S.xf_block("cpu", "fp32", Logging(), layers=1)
    free = [s for s in seen if s[0] & 2]
    assert len(free) >= 10 and all(not bias and not relu for _, bias, _, relu in free)
    first_free = next(i for i, s in enumerate(seen) if s[0] & 2)
    assert all(s[0] & 2 for s in seen[first_free:]), "once the tape runs, every product is a gradient"
    assert all((bias or nb > 1) for _, bias, nb, _ in seen[:first_free]), "forward products"
