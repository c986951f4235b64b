"""Small engine-level scenarios run identically on (cpu, numpy kernel spec) and (cuda, real kernels)."""
import torch
import torch.nn.functional as F

from vinet_b200 import lib as L
from vinet_b200 import model as M
from vinet_b200.engine import ConvGeom, Engine


def ncdhw(t):
    return t.detach().float().permute(0, 4, 1, 2, 3).contiguous().cpu()


def bf16r(t):
    return t.to(torch.bfloat16).float()


def make_engine(device, precision, backend=None):
    e = Engine(precision, backend=backend)
    e.begin(torch.device(device), True, True)
    return e


def fill_act(e, name, B, T, H, W, C, gen, affine, gdtype=None):
    a = e.new_act(name, B, T, H, W, C, affine=affine, gdtype=gdtype)
    a.buf.copy_(bf16r(torch.randn(a.buf.shape, generator=gen)))
    if affine:
        a.scale.copy_(torch.rand(C, generator=gen) + 0.5)
        a.shift.copy_(torch.randn(C, generator=gen) * 0.3)
        a.xform = L.XF_AFFINE_RELU
    return a


def run_tape(e, out, gen):
    for g in e.grad_bufs:
        g.zero_()
    e.gwritten = {g.data_ptr() for g in e.grad_bufs}      # everything is initialised: writers accumulate
    go = bf16r(torch.randn(out.grad.shape, generator=gen))
    out.grad.copy_(go)
    for fn in reversed(e.tape):
        fn()
    e.unpack_flush()          # weight-gradient unpacks are queued and launched in batches (Engine.backward does this)
    if e.device.type == "cuda":
        torch.cuda.synchronize()
    return go


def conv_up(device, precision, backend=None, B=2, T0=1, T1=2, H=6, W=5, Cin=16, Cout=24, kt=3, seed=0):
    """decoder stage: conv over a T-concat of (identity, affine+relu) sources -> relu -> 2x upsample."""
    gen = torch.Generator().manual_seed(seed)
    e = make_engine(device, precision, backend)
    u = fill_act(e, "u", B, T0, H, W, Cin, gen, False)
    y = fill_act(e, "y", B, T1, H, W, Cin, gen, True)
    w = bf16r(torch.randn(Cout, Cin, kt, 3, 3, generator=gen) * (2.0 / (Cin * kt * 9)) ** 0.5).to(device)
    out = e.conv_relu_up("c", [u, y], w, ConvGeom((kt, 3, 3), (kt, 1, 1), (0, 1, 1)))
    hi = e.materialize_up2(out)        # the stage's output is virtual (read through the up-sampling): the standalone kernel writes it
    run_tape(e, out, gen)
    return {"out": ncdhw(hi.buf), "dW": e.param_grads["c.weight"].cpu(), "du": ncdhw(u.grad), "dy": ncdhw(y.grad)}


def up_fused(device, precision, backend=None, B=2, T0=1, T1=2, h=5, w=6, C0=16, C1=24, C2=16, kt=3, seed=0, fuse=True):
    """Two decoder stages (model.py:286-298): conv(1,3,3) -> relu -> 2x up -> [T-concat with a skip tensor] -> conv(kt,3,3)/(kt,1,1)
    -> relu -> 2x up.  The second convolution reads the first stage's output THROUGH the relu + up-sampling in its input stage
    (fprop and weight gradient); fuse=False materialises the hi-res tensor first (the pre-fusion plan) for an A/B comparison."""
    gen = torch.Generator().manual_seed(seed)
    e = make_engine(device, precision, backend)
    x = fill_act(e, "x", B, T0, h, w, C0, gen, False)
    y = fill_act(e, "y", B, T1, 2 * h, 2 * w, C1, gen, False)
    w1 = bf16r(torch.randn(C1, C0, 1, 3, 3, generator=gen) * (2.0 / (C0 * 9)) ** 0.5).to(device)
    w2 = bf16r(torch.randn(C2, C1, kt, 3, 3, generator=gen) * (2.0 / (C1 * kt * 9)) ** 0.5).to(device)
    u1 = e.conv_relu_up("c1", [x], w1, ConvGeom((1, 3, 3), (1, 1, 1), (0, 1, 1)))
    if not fuse:
        u1 = e.materialize_up2(u1)
    e.profile = [] if backend is None and device != "cpu" else None
    n_up2 = L.get().fn["vinet_up2_launch_count"]() if backend is None else 0
    u2 = e.conv_relu_up("c2", [u1, y], w2, ConvGeom((kt, 3, 3), (kt, 1, 1), (0, 1, 1)))
    hi = e.materialize_up2(u2)
    run_tape(e, u2, gen)
    kernels = sorted({r[5] for r in e.profile}) if e.profile is not None else []
    e.profile = None
    return {"out": ncdhw(hi.buf), "dW1": e.param_grads["c1.weight"].cpu(), "dW2": e.param_grads["c2.weight"].cpu(),
            "dx": ncdhw(x.grad), "dy": ncdhw(y.grad), "_kernels": kernels, "_materialized": list(e.up2_mat_log),
            "_up2_launches": (L.get().fn["vinet_up2_launch_count"]() - n_up2) if backend is None else 0}


def up_fused_torch(B=2, T0=1, T1=2, h=5, w=6, C0=16, C1=24, C2=16, kt=3, seed=0):
    gen = torch.Generator().manual_seed(seed)

    def rnd(*s):
        return bf16r(torch.randn(*s, generator=gen))
    xb = rnd(B, T0, h, w, C0)
    yb = rnd(B, T1, 2 * h, 2 * w, C1)
    w1 = bf16r(torch.randn(C1, C0, 1, 3, 3, generator=gen) * (2.0 / (C0 * 9)) ** 0.5).requires_grad_(True)
    w2 = bf16r(torch.randn(C2, C1, kt, 3, 3, generator=gen) * (2.0 / (C1 * kt * 9)) ** 0.5).requires_grad_(True)
    x = xb.permute(0, 4, 1, 2, 3).contiguous().requires_grad_(True)
    y = yb.permute(0, 4, 1, 2, 3).contiguous().requires_grad_(True)
    u1 = F.interpolate(F.relu(F.conv3d(x, w1, None, 1, (0, 1, 1))), scale_factor=(1, 2, 2), mode="trilinear")
    u2 = F.interpolate(F.relu(F.conv3d(torch.cat([u1, y], 2), w2, None, (kt, 1, 1), (0, 1, 1))), scale_factor=(1, 2, 2), mode="trilinear")
    go = bf16r(torch.randn(B, u2.shape[2], u2.shape[3], u2.shape[4], C2, generator=gen)).permute(0, 4, 1, 2, 3)
    u2.backward(go)
    return {"out": u2.detach(), "dW1": w1.grad, "dW2": w2.grad, "dx": x.grad, "dy": y.grad}


def conv_up_torch(B=2, T0=1, T1=2, H=6, W=5, Cin=16, Cout=24, kt=3, seed=0):
    """The same scenario in plain PyTorch fp32 (the torch reference of the kernels)."""
    gen = torch.Generator().manual_seed(seed)

    def rnd(*s):
        return bf16r(torch.randn(*s, generator=gen))
    ub = rnd(B, T0, H, W, Cin)
    yb = rnd(B, T1, H, W, Cin)
    sc = torch.rand(Cin, generator=gen) + 0.5
    sh = torch.randn(Cin, generator=gen) * 0.3
    w = bf16r(torch.randn(Cout, Cin, kt, 3, 3, generator=gen) * (2.0 / (Cin * kt * 9)) ** 0.5).requires_grad_(True)
    xu = ub.permute(0, 4, 1, 2, 3).contiguous().requires_grad_(True)
    xy = F.relu(yb.permute(0, 4, 1, 2, 3) * sc.view(1, -1, 1, 1, 1) + sh.view(1, -1, 1, 1, 1))
    xy.retain_grad() if xy.requires_grad else None
    xy = xy.detach().requires_grad_(True)
    z = F.conv3d(torch.cat([xu, xy], 2), w, None, (kt, 1, 1), (0, 1, 1))
    o = F.interpolate(F.relu(z), scale_factor=(1, 2, 2), mode="trilinear")
    go = bf16r(torch.randn(B, T0 and o.shape[2], o.shape[3], o.shape[4], Cout, generator=gen)).permute(0, 4, 1, 2, 3)
    o.backward(go)
    return {"out": o.detach(), "dW": w.grad, "du": xu.grad, "dy": xy.grad}


def mixed(device, precision, backend=None, name="3b", B=2, T=3, H=5, W=4, seed=0):
    from oracle import torch_oracle as O
    gen = torch.Generator().manual_seed(seed)
    ref = O.Inception(name)
    O.randomize_(ref, seed + 1)
    m = M.Mixed(name)
    m.load_state_dict(ref.state_dict())
    with torch.no_grad():
        for p in m.parameters():
            if p.dim() == 5:
                p.copy_(bf16r(p))
    m.to(device)
    cin = M.arch.MIXED[name][0]
    e = make_engine(device, precision, backend)
    x = fill_act(e, "x", B, T, H, W, cin, gen, True)
    out = M._mixed(e, "mx", x, m)
    run_tape(e, out, gen)
    res = {"out": ncdhw(out.buf), "dx": ncdhw(x.grad)}
    for k, v in e.param_grads.items():
        res["g/" + k] = v.detach().cpu()
    return res


def stem(device, precision, backend=None, B=1, T=8, H=32, W=32, seed=0):
    """pack_input + SepConv3d(3,64,k7,s2,p3) + MaxPool(1,3,3)/(1,2,2): strided convs, channel padding."""
    from oracle import torch_oracle as O
    gen = torch.Generator().manual_seed(seed)
    ref = O.SepConv(3, 64, 7, 2, 3)
    O.randomize_(ref, seed + 2)
    m = M.SepConv3d(3, 64, 7, 2, 3)
    m.load_state_dict(ref.state_dict())
    with torch.no_grad():
        for p in m.parameters():
            if p.dim() == 5:
                p.copy_(bf16r(p))
    m.to(device)
    e = make_engine(device, precision, backend)
    x = bf16r(torch.randn(B, T, 3, H, W, generator=gen)).permute(0, 2, 1, 3, 4).to(device)
    xin = M.pack_input(e, x)
    a = M._sepconv(e, "stem", [xin], m, cin_real=3)
    out = e.maxpool("pool", a, (1, 3, 3), (1, 2, 2), (0, 1, 1))
    run_tape(e, out, gen)
    res = {"out": ncdhw(out.buf), "raw": ncdhw(a.buf)}
    for k, v in e.param_grads.items():
        res["g/" + k] = v.detach().cpu()
    return res


def compare(a, b, rtol, what="", skip=()):
    """Per entry: relative L2 error <= rtol and max|a-b| <= 20*rtol*max|b| (a single ReLU-mask flip between
    two correct implementations moves a few elements by a lot, never the bulk). Returns the failures."""
    bad = []
    for k in b:
        if k.startswith("_") or any(s in k for s in skip):
            continue
        ref = b[k].float()
        got = a[k].float().reshape(ref.shape)
        err = (got - ref).abs().max().item()
        scale = ref.abs().max().item() + 1e-20
        l2 = ((got - ref).norm() / (ref.norm() + 1e-20)).item()
        if not (l2 <= rtol and err <= 20 * rtol * scale):
            bad.append("%s%s: relL2 %.3e maxerr %.3e vs scale %.3e" % (what, k, l2, err, scale))
    return bad


def audio_fuse(device, precision, backend=None, B=2, seed=0):
    """SoundNet stack + max-pool/bilinear fusion of AViNet, standalone (inputs: waveform + a y0 feature)."""
    from oracle import torch_oracle as O
    from vinet_b200 import avmodel as AV
    gen = torch.Generator().manual_seed(seed)
    ref = O.SoundNetOracle()
    O.randomize_(ref, seed + 3)
    net = AV.SoundNet()
    net.load_state_dict(ref.state_dict())
    net.to(device)
    bil = AV.BilinearParams(42, 3, 336)
    with torch.no_grad():
        bil.weight.copy_(torch.randn(336, 42, 3, generator=gen) * 0.1)
        bil.bias.copy_(torch.randn(336, generator=gen) * 0.1)
    bil.to(device)
    e = make_engine(device, precision, backend)
    audio = O.make_inputs(B, 8, 32, 32, seed, audio=True)["audio"].to(device)
    a, ga = AV.soundnet_plan(e, "audionet.", net, audio)
    y0 = fill_act(e, "y0", B, 4, 7, 12, 1024, gen, True, gdtype=torch.float32)
    out = AV.avfuse_plan(e, "bilinear", y0, a, ga, bil)
    run_tape(e, out, gen)
    res = {"a": a.detach().cpu().clone(), "out": ncdhw(out.buf), "dy0": ncdhw(y0.grad)}
    for k, v in e.param_grads.items():
        res["g/" + k] = v.detach().cpu()
    return res


def audio_fuse_torch(B=2, seed=0):
    from oracle import torch_oracle as O
    gen = torch.Generator().manual_seed(seed)
    ref = O.SoundNetOracle()
    O.randomize_(ref, seed + 3)
    ref.train()
    w = (torch.randn(336, 42, 3, generator=gen) * 0.1).requires_grad_(True)
    b = (torch.randn(336, generator=gen) * 0.1).requires_grad_(True)
    audio = O.make_inputs(B, 8, 32, 32, seed, audio=True)["audio"]
    a = ref(audio)
    yb = bf16r(torch.randn(B, 4, 7, 12, 1024, generator=gen))
    sc = torch.rand(1024, generator=gen) + 0.5
    sh = torch.randn(1024, generator=gen) * 0.3
    y0 = F.relu(yb.permute(0, 4, 1, 2, 3) * sc.view(1, -1, 1, 1, 1) + sh.view(1, -1, 1, 1, 1)).requires_grad_(True)
    f = F.bilinear(F.max_pool3d(y0, (4, 1, 1), (2, 1, 2)).flatten(2), a.flatten(2), w, b)
    f = f.view(B, 1024, 4, 7, 12)
    go = bf16r(torch.randn(B, 4, 7, 12, 1024, generator=gen)).permute(0, 4, 1, 2, 3)
    f.backward(go)
    res = {"a": a.detach().flatten(2), "out": f.detach(), "dy0": y0.grad, "g/bilinear.weight": w.grad, "g/bilinear.bias": b.grad}
    for n, p in ref.named_parameters():
        if p.grad is not None:
            res["g/audionet." + n] = p.grad.reshape(p.grad.shape[:3]) if p.grad.dim() == 4 else p.grad
    return res


# ----------------------------------------------------------------------------- transformer fusion block (AViNet use_transformer=True)
class _XfRef(torch.nn.Module):
    """conv_in_1x1 -> tokens = channels -> Transformer -> conv_out_1x1 (model.py:239-247), plain torch."""

    def __init__(self, layers):
        super().__init__()
        from oracle import torch_oracle as O
        self.conv_in_1x1 = torch.nn.Conv3d(1024, 32, 1)
        self.conv_out_1x1 = torch.nn.Conv3d(32, 1024, 1)
        self.transformer = O.TransformerOracle(336, 336, 4, layers, 32)

    def forward(self, f):
        t = self.conv_in_1x1(f).flatten(2).permute(1, 0, 2)
        t = self.transformer(t).permute(1, 0, 2)
        return self.conv_out_1x1(t.reshape(t.size(0), t.size(1), 4, 7, 12))


def _xf_ref(layers, seed):
    from oracle import torch_oracle as O
    ref = O.set_dropout(O.randomize_(_XfRef(layers), seed + 5), 0.0)
    return ref.train()


def xf_block(device, precision, backend=None, B=2, layers=2, seed=0, p=0.0, training=True):
    """The transformer fusion block of AViNet standalone: a random fused feature in, decoder-input gradient back."""
    from oracle import torch_oracle as O
    from vinet_b200 import xfmr as X
    gen = torch.Generator().manual_seed(seed)
    ref = _xf_ref(layers, seed)

    class Holder(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.conv_in_1x1 = M.ConvParams(1024, 32, (1, 1, 1), bias=True)
            self.conv_out_1x1 = M.ConvParams(32, 1024, (1, 1, 1), bias=True)
            self.transformer = X.Transformer(336, hidden_size=336, nhead=4, num_encoder_layers=layers, max_len=32)
    m = Holder()
    m.load_state_dict(ref.state_dict())
    O.set_dropout(m, p)
    m.to(device)
    e = Engine(precision, backend=backend)
    e.begin(torch.device(device), training, True)
    fused = fill_act(e, "fused", B, 4, 7, 12, 1024, gen, False, gdtype=torch.float32)
    out = X.avinet_transformer_plan(e, m, fused)
    run_tape(e, out, gen)
    res = {"out": ncdhw(out.buf), "dfused": ncdhw(fused.grad)}
    for k, v in e.param_grads.items():
        res["g/" + k] = v.detach().cpu().clone()
    return res


def xf_block_torch(B=2, layers=2, seed=0):
    gen = torch.Generator().manual_seed(seed)
    ref = _xf_ref(layers, seed)
    f = bf16r(torch.randn(B, 4, 7, 12, 1024, generator=gen)).permute(0, 4, 1, 2, 3).requires_grad_(True)
    out = ref(f)
    go = bf16r(torch.randn(B, 4, 7, 12, 1024, generator=gen)).permute(0, 4, 1, 2, 3)
    out.backward(go)
    res = {"out": out.detach(), "dfused": f.grad}
    for n, prm in ref.named_parameters():
        res["g/" + n] = prm.grad
    return res


# ----------------------------------------------------------------------------- VideoAudioSaliencyFusionModel's token block
class _FuRef(torch.nn.Module):
    """model.py:151-183 between the backbone / SoundNet outputs and the decoder input, plain torch (d = C' features)."""

    def __init__(self, d, layers):
        super().__init__()
        from oracle import torch_oracle as O
        self.conv_in_1x1 = torch.nn.Conv3d(1024, d, 1)
        self.audio_conv_1x1 = torch.nn.Conv2d(1024, d, 1)
        self.transformer = O.TransformerOracle(d, d, 4, layers, 339)

    def forward(self, y0, a):
        v = self.conv_in_1x1(y0).flatten(2)
        au = self.audio_conv_1x1(a).flatten(2)
        t = self.transformer(torch.cat((v, au), 2).permute(2, 0, 1)).permute(1, 2, 0)
        vf = t[..., :336].reshape(t.size(0), t.size(1), 4, 7, 12)
        af = t[..., 336:].mean(2).view(t.size(0), t.size(1), 1, 1, 1).repeat(1, 1, 4, 7, 12)
        return torch.cat((vf, af), 1)


def _fu_ref(d, layers, seed):
    from oracle import torch_oracle as O
    return O.set_dropout(O.randomize_(_FuRef(d, layers), seed + 9), 0.0).train()


def fusion_block(device, precision, backend=None, B=2, d=64, layers=1, seed=0):
    from oracle import torch_oracle as O
    from vinet_b200 import xfmr as X
    gen = torch.Generator().manual_seed(seed)
    ref = _fu_ref(d, layers, seed)

    class Holder(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.conv_in_1x1 = M.ConvParams(1024, d, (1, 1, 1), bias=True)
            self.audio_conv_1x1 = M.ConvParams(1024, d, (1, 1), bias=True)
            self.transformer = X.Transformer(d, hidden_size=d, nhead=4, num_encoder_layers=layers, max_len=339)
    m = Holder()
    m.load_state_dict(ref.state_dict())
    O.set_dropout(m, 0.0)
    m.to(device)
    e = make_engine(device, precision, backend)
    y0 = fill_act(e, "y0", B, 4, 7, 12, 1024, gen, True, gdtype=torch.float32)
    a = torch.randn(B, 1024, 3, generator=gen).to(device)
    ga = torch.zeros_like(a)
    out = X.fusion_plan(e, m, y0, a, ga)
    run_tape(e, out, gen)
    res = {"out": ncdhw(out.buf), "dy0": ncdhw(y0.grad), "da": ga.cpu().clone()}
    for k, v in e.param_grads.items():
        res["g/" + k] = v.detach().cpu().clone()
    return res


def fusion_block_torch(B=2, d=64, layers=1, seed=0):
    gen = torch.Generator().manual_seed(seed)
    ref = _fu_ref(d, layers, seed)
    yb = bf16r(torch.randn(B, 4, 7, 12, 1024, generator=gen))
    sc = torch.rand(1024, generator=gen) + 0.5
    sh = torch.randn(1024, generator=gen) * 0.3
    y0 = F.relu(yb.permute(0, 4, 1, 2, 3) * sc.view(1, -1, 1, 1, 1) + sh.view(1, -1, 1, 1, 1)).requires_grad_(True)
    a = torch.randn(B, 1024, 3, generator=gen).requires_grad_(True)
    out = ref(y0, a.unsqueeze(-1))
    go = bf16r(torch.randn(B, 4, 7, 12, 2 * d, generator=gen)).permute(0, 4, 1, 2, 3)
    out.backward(go)
    res = {"out": out.detach(), "dy0": y0.grad, "da": a.grad}
    for n, prm in ref.named_parameters():
        res["g/" + n] = prm.grad
    return res
