"""TEST INFRASTRUCTURE ONLY — plain PyTorch fp32 restatement of the ViNet/AViNet hot path.

A table-driven re-statement of what the reference computes, built from stock ``torch.nn`` ops
so it runs on the CPU (here and on the GPU box's host cores) and, for cross-checks, on CUDA in
fp32 with TF32 disabled.  The module tree uses the reference's attribute names so that
``state_dict()`` keys/shapes are interchangeable with the reference.

It is pinned against the executed reference by ``oracle/make_golden.py`` (state_dict key/shape
equality, forward outputs, parameter gradients, running-stat updates and the loss scalars) and
re-checked against the committed vectors by ``tests/test_oracle_golden.py``.

Reference sites restated here:
  BasicConv3d ............ model_utils.py:128-139     SepConv3d ........ model_utils.py:141-160
  Mixed_* ................ model_utils.py:162-420     BackBoneS3D ...... model.py:690-743
  DecoderConvUp{,8,16,48}  model.py:251-499           VideoSaliencyModel model.py:72-112
  SoundNet ............... model.py:746-825           VideoAudioSaliencyModel model.py:191-249
  kldiv/cc/similarity/nss  loss.py:13-120             loss_func/get_loss utils.py:9-39
"""
import math

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import arch

BN_EPS, BN_MOM = 1e-3, 1e-3          # model_utils.py:132
EPS = 2.2204e-16                     # loss.py:35


class _Seq(nn.Sequential):
    pass


class UnitConv(nn.Module):
    """conv(1x1x1, no bias) -> BN -> ReLU   (model_utils.py:128-139)."""

    def __init__(self, cin, cout):
        super().__init__()
        self.conv = nn.Conv3d(cin, cout, 1, 1, 0, bias=False)
        self.bn = nn.BatchNorm3d(cout, eps=BN_EPS, momentum=BN_MOM)

    def forward(self, x):
        return F.relu(self.bn(self.conv(x)))


class SepConv(nn.Module):
    """(1,k,k) conv -> BN -> ReLU -> (k,1,1) conv -> BN -> ReLU   (model_utils.py:141-160)."""

    def __init__(self, cin, cout, k, s, p):
        super().__init__()
        self.conv_s = nn.Conv3d(cin, cout, (1, k, k), (1, s, s), (0, p, p), bias=False)
        self.bn_s = nn.BatchNorm3d(cout, eps=BN_EPS, momentum=BN_MOM)
        self.conv_t = nn.Conv3d(cout, cout, (k, 1, 1), (s, 1, 1), (p, 0, 0), bias=False)
        self.bn_t = nn.BatchNorm3d(cout, eps=BN_EPS, momentum=BN_MOM)

    def forward(self, x):
        x = F.relu(self.bn_s(self.conv_s(x)))
        return F.relu(self.bn_t(self.conv_t(x)))


class Inception(nn.Module):
    """Four-branch S3D block; outputs concatenated on channels (model_utils.py:162-190)."""

    def __init__(self, name):
        super().__init__()
        cin, b0, b1r, b1, b2r, b2, b3 = arch.MIXED[name]
        self.branch0 = _Seq(UnitConv(cin, b0))
        self.branch1 = _Seq(UnitConv(cin, b1r), SepConv(b1r, b1, 3, 1, 1))
        self.branch2 = _Seq(UnitConv(cin, b2r), SepConv(b2r, b2, 3, 1, 1))
        self.branch3 = _Seq(nn.MaxPool3d(3, 1, 1), UnitConv(cin, b3))

    def forward(self, x):
        return torch.cat([self.branch0(x), self.branch1(x), self.branch2(x), self.branch3(x)], 1)


class Backbone(nn.Module):
    """S3D encoder returning the four hierarchy tensors [y0,y1,y2,y3] (model.py:690-743)."""

    def __init__(self):
        super().__init__()
        self.base1 = _Seq(SepConv(3, 64, 7, 2, 3), nn.MaxPool3d((1, 3, 3), (1, 2, 2), (0, 1, 1)),
                          UnitConv(64, 64), SepConv(64, 192, 3, 1, 1))
        self.maxp2 = nn.MaxPool3d((1, 3, 3), (1, 2, 2), (0, 1, 1))
        self.base2 = _Seq(*[Inception(n) for n in arch.STAGES["base2"]])
        self.maxp3 = nn.MaxPool3d(3, 2, 1)
        self.base3 = _Seq(*[Inception(n) for n in arch.STAGES["base3"]])
        self.maxt4 = nn.MaxPool3d((2, 1, 1), (2, 1, 1))
        self.maxp4 = nn.MaxPool3d((1, 2, 2), (1, 2, 2))
        self.base4 = _Seq(*[Inception(n) for n in arch.STAGES["base4"]])

    def forward(self, x):
        y3 = self.base1(x)
        y2 = self.base2(self.maxp2(y3))
        y1 = self.base3(self.maxp3(y2))
        y0 = self.base4(self.maxp4(self.maxt4(y1)))
        return [y0, y1, y2, y3]


class Decoder(nn.Module):
    """conv -> ReLU -> 2x bilinear, T-concat with the skip tensor, time-collapsing convs, sigmoid
    (model.py:251-311 and the 8/16/48-frame variants)."""

    def __init__(self, num_clips=32, num_hier=3):
        super().__init__()
        self.num_hier = num_hier
        if num_hier != 3:
            num_clips = 32       # the ablation decoders exist for 32-frame clips only (model.py:84-90 ignores num_clips)
        self.upsampling = nn.Upsample(scale_factor=(1, 2, 2), mode="trilinear")
        heads = []
        for cin, cout, kt in arch.decoder_head(num_hier):
            heads.append([nn.Conv3d(cin, cout, (kt, 3, 3), (kt, 1, 1), (0, 1, 1), bias=False),
                          nn.ReLU(), self.upsampling])
        tail = []
        for item in arch.decoder_tail(num_clips):
            if item == "relu":
                tail.append(nn.ReLU())
            elif item == "up":
                tail.append(self.upsampling)
            elif item == "sigmoid":
                tail.append(nn.Sigmoid())
            else:
                _, cin, cout, k, s, p, bias = item
                tail.append(nn.Conv3d(cin, cout, k, s, (0, p, p), bias=bias))
        self.convtsp1 = _Seq(*heads[0])
        self.convtsp2 = _Seq(*heads[1])
        self.convtsp3 = _Seq(*heads[2])
        self.convtsp4 = _Seq(*(heads[3] + tail))

    def forward(self, y0, y1=None, y2=None, y3=None):
        h = self.num_hier
        z = self.convtsp1(y0)
        z = self.convtsp2(torch.cat((z, y1), 2) if h >= 1 else z)
        z = self.convtsp3(torch.cat((z, y2), 2) if h >= 2 else z)
        z = self.convtsp4(torch.cat((z, y3), 2) if h >= 3 else z)
        return z.view(z.size(0), z.size(3), z.size(4))


class ViNetOracle(nn.Module):
    def __init__(self, num_clips=32, num_hier=3):
        super().__init__()
        self.backbone = Backbone()
        self.num_hier = num_hier
        self.decoder = Decoder(num_clips, num_hier)

    def forward(self, x):
        ys = self.backbone(x)
        return self.decoder(*ys[:self.num_hier + 1])


class SoundNetOracle(nn.Module):
    """SoundNet-8 trunk, layers 1..7 (model.py:746-825); conv8_* exist as unused parameters."""

    def __init__(self):
        super().__init__()
        for i, (cin, cout, k, p, pool) in enumerate(arch.SOUNDNET, 1):
            setattr(self, f"conv{i}", nn.Conv2d(cin, cout, (k, 1), (2, 1), (p, 0)))
            setattr(self, f"batchnorm{i}", nn.BatchNorm2d(cout, eps=1e-5, momentum=0.1))
        self.conv8_objs = nn.Conv2d(1024, 1000, (8, 1), (2, 1))
        self.conv8_scns = nn.Conv2d(1024, 401, (8, 1), (2, 1))

    def forward(self, w):
        x = w
        for i, (_, _, _, _, pool) in enumerate(arch.SOUNDNET, 1):
            x = F.relu(getattr(self, f"batchnorm{i}")(getattr(self, f"conv{i}")(x)))
            if pool > 1:
                x = F.max_pool2d(x, (pool, 1), (pool, 1))
        return x


class PositionalEncodingOracle(nn.Module):
    """model.py:8-26: sinusoidal table added to the (S, B, F) tokens; the module's Dropout is never applied (model.py:24-26)."""

    def __init__(self, feat_size, dropout=0.1, max_len=4):
        super().__init__()
        self.dropout = nn.Dropout(p=dropout)
        pos = torch.arange(0, max_len, dtype=torch.float).unsqueeze(1)
        div = torch.exp(torch.arange(0, feat_size, 2).float() * (-math.log(10000.0) / feat_size))
        pe = torch.zeros(max_len, feat_size)
        pe[:, 0::2], pe[:, 1::2] = torch.sin(pos * div), torch.cos(pos * div)
        self.register_buffer("pe", pe.unsqueeze(1))

    def forward(self, x):
        return x + self.pe


class TransformerOracle(nn.Module):
    """model.py:28-69 in the encoder-only form every model of the reference instantiates (num_decoder_layers=-1, spatial_dim=-1)."""

    def __init__(self, feat_size, hidden_size, nhead, num_encoder_layers, max_len):
        super().__init__()
        self.pos_encoder = PositionalEncodingOracle(feat_size, max_len=max_len)
        self.transformer_encoder = nn.TransformerEncoder(nn.TransformerEncoderLayer(feat_size, nhead, hidden_size), num_encoder_layers)

    def forward(self, tokens):
        return self.transformer_encoder(self.pos_encoder(tokens))


def set_dropout(model, p):
    """Dropout probability of every nn.Dropout and attention module of `model` (works on the reference, the oracle and the
    drop-in's parameter holders alike).  The parity cases run with p = 0: torch's Philox stream is not reproduced."""
    for m in model.modules():
        if isinstance(m, nn.Dropout):
            m.p = p
        elif isinstance(m, nn.MultiheadAttention):
            m.dropout = p
    return model


class AViNetOracle(nn.Module):
    """SoundNet || backbone -> MaxPool3d((4,1,1),s=(2,1,2)) -> Bilinear(42,3,336) [-> conv_in_1x1 -> Transformer over the C'
    channel tokens -> conv_out_1x1] -> decoder (model.py:191-249)."""

    def __init__(self, num_clips=32, use_transformer=False, transformer_in_channel=32, num_encoder_layers=3, nhead=4):
        super().__init__()
        self.use_transformer = use_transformer
        self.visual_model = ViNetOracle(num_clips)
        if use_transformer:
            self.conv_in_1x1 = nn.Conv3d(1024, transformer_in_channel, 1)
            self.conv_out_1x1 = nn.Conv3d(32, 1024, 1)
            self.transformer = TransformerOracle(4 * 7 * 12, 4 * 7 * 12, nhead, num_encoder_layers, transformer_in_channel)
        self.audionet = SoundNetOracle()
        self.bilinear = nn.Bilinear(42, 3, 4 * 7 * 12)

    def forward(self, x, audio):
        a = self.audionet(audio)
        y0, y1, y2, y3 = self.visual_model.backbone(x)
        y0 = F.max_pool3d(y0, (4, 1, 1), (2, 1, 2))
        f = self.bilinear(y0.flatten(2), a.flatten(2))
        f = f.view(f.size(0), f.size(1), 4, 7, 12)
        if self.use_transformer:                      # model.py:239-247
            t = self.conv_in_1x1(f).flatten(2).permute(1, 0, 2)
            t = self.transformer(t).permute(1, 0, 2)
            f = self.conv_out_1x1(t.reshape(t.size(0), t.size(1), 4, 7, 12))
        return self.visual_model.decoder(f, y1, y2, y3)


class AVFusionOracle(nn.Module):
    """VideoAudioSaliencyFusionModel (model.py:116-189): 336 visual tokens (conv_in_1x1 of y0) and 3 audio tokens (audio_conv_1x1 of
    the SoundNet output) of C' features through one Transformer; decoder input = [visual tokens | mean audio token, broadcast]."""

    def __init__(self, transformer_in_channel=512, num_encoder_layers=3, nhead=4, num_clips=32):
        super().__init__()
        self.visual_model = ViNetOracle(num_clips)
        self.conv_in_1x1 = nn.Conv3d(1024, transformer_in_channel, 1)
        self.transformer = TransformerOracle(transformer_in_channel, transformer_in_channel, nhead, num_encoder_layers, 4 * 7 * 12 + 3)
        self.audionet = SoundNetOracle()
        self.audio_conv_1x1 = nn.Conv2d(1024, transformer_in_channel, 1)
        self.bilinear = nn.Bilinear(42, 3, 4 * 7 * 12)      # registered but unused, as in the reference (model.py:149)

    def forward(self, x, audio):
        a = self.audio_conv_1x1(self.audionet(audio)).flatten(2)
        y0, y1, y2, y3 = self.visual_model.backbone(x)
        v = self.conv_in_1x1(y0).flatten(2)
        t = self.transformer(torch.cat((v, a), 2).permute(2, 0, 1)).permute(1, 2, 0)
        n = 4 * 7 * 12
        vf = t[..., :n].reshape(t.size(0), t.size(1), 4, 7, 12)
        af = t[..., n:].mean(2).view(t.size(0), t.size(1), 1, 1, 1).repeat(1, 1, 4, 7, 12)
        return self.visual_model.decoder(torch.cat((vf, af), 1), y1, y2, y3)


# ----------------------------------------------------------------------------- losses (loss.py)
def _flat(x):
    return x.reshape(x.size(0), -1)


def kldiv(s, g):
    """loss.py:13-38."""
    s = _flat(s); g = _flat(g)
    s = s / s.sum(1, keepdim=True)
    g = g / g.sum(1, keepdim=True)
    return (g * torch.log(EPS + g / (s + EPS))).sum(1).mean()


def _minmax(x):
    """loss.py:41-51."""
    mn = x.min(1, keepdim=True)[0]
    mx = x.max(1, keepdim=True)[0]
    return (x - mn) / (mx - mn)


def similarity(s, g):
    """loss.py:53-78."""
    s = _minmax(_flat(s)); g = _minmax(_flat(g))
    s = s / s.sum(1, keepdim=True)
    g = g / g.sum(1, keepdim=True)
    return torch.min(s, g).sum(1).mean()


def cc(s, g):
    """loss.py:80-99 (torch.std is the unbiased estimator)."""
    s = _flat(s); g = _flat(g)
    s = (s - s.mean(1, keepdim=True)) / s.std(1, keepdim=True)
    g = (g - g.mean(1, keepdim=True)) / g.std(1, keepdim=True)
    return ((s * g).sum(1) / torch.sqrt((s * s).sum(1) * (g * g).sum(1))).mean()


def nss(s, fix):
    """loss.py:101-120 (same-size branch only)."""
    s = _flat(s); fix = _flat(fix)
    s = (s - s.mean(1, keepdim=True)) / (s.std(1, keepdim=True) + EPS)
    return ((s * fix).sum(1) / fix.sum(1)).mean()


def loss_func(pred, gt, kldiv_coeff=1.0, cc_coeff=None, sim_coeff=None):
    """utils.py:9-39 with the default flags (only kldiv enabled); returns a shape-(1,) tensor."""
    def one(p, g):
        loss = torch.zeros(1, dtype=p.dtype, device=p.device)
        if kldiv_coeff is not None:
            loss = loss + kldiv_coeff * kldiv(p, g)
        if cc_coeff is not None:
            loss = loss + cc_coeff * cc(p, g)
        if sim_coeff is not None:
            loss = loss + sim_coeff * similarity(p, g)
        return loss
    if pred.dim() == 4:
        p = pred.permute(1, 0, 2, 3); g = gt.permute(1, 0, 2, 3)
        return sum(one(p[i], g[i]) for i in range(p.size(0))) / p.size(0)
    return one(pred, gt)


# ----------------------------------------------------------------------------- seeded helpers
def randomize_(model, seed=0, head_gain=4.0):
    """Deterministic non-trivial weights: default Kaiming init gives a near-constant 0.535 output
    (SURVEY.md fact 7), so BN affine/running stats are randomised and the last conv is scaled up."""
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for name, m in model.named_modules():
            if isinstance(m, (nn.BatchNorm3d, nn.BatchNorm2d)):
                m.weight.copy_(0.5 + torch.rand(m.weight.shape, generator=g))
                m.bias.copy_(0.2 * torch.randn(m.bias.shape, generator=g))
                m.running_mean.copy_(0.1 * torch.randn(m.running_mean.shape, generator=g))
                m.running_var.copy_(0.5 + torch.rand(m.running_var.shape, generator=g))
            elif isinstance(m, (nn.LayerNorm, nn.Linear, nn.MultiheadAttention)):
                for p in m._parameters.values():      # transformer variants: the three encoder layers start as deep copies
                    if p is not None:
                        p.copy_(torch.randn(p.shape, generator=g) * (0.3 if p.dim() == 1 else (1.0 / p.shape[-1]) ** 0.5)
                                + (1.0 if isinstance(m, nn.LayerNorm) and p is m.weight else 0.0))
            elif isinstance(m, (nn.Conv3d, nn.Conv2d, nn.Bilinear)):
                fan_in = m.weight[0].numel()
                m.weight.copy_(torch.randn(m.weight.shape, generator=g) * (2.0 / fan_in) ** 0.5)
                if m.bias is not None:
                    m.bias.copy_(0.1 * torch.randn(m.bias.shape, generator=g))
        last = None
        for m in model.modules():
            if isinstance(m, nn.Conv3d) and m.out_channels == 1:
                last = m
        if last is not None:       # zero-mean head weights keep the sigmoid away from saturation
            last.weight.sub_(last.weight.mean()).mul_(head_gain)
            last.bias.zero_()
    return model


def make_inputs(B, T, H, W, seed=0, audio=False):
    """Synthetic clip in the caller's layout: a non-contiguous (B,3,T,H,W) permute of (B,T,3,H,W)
    memory (train.py:204-205), a strictly positive gt map and a binary fixation map."""
    g = torch.Generator().manual_seed(1000 + seed)
    x = torch.randn(B, T, 3, H, W, generator=g).permute(0, 2, 1, 3, 4)
    gt = torch.rand(B, H, W, generator=g) + 1e-3
    fix = (torch.rand(B, H, W, generator=g) > 0.98).float()
    out = {"x": x, "gt": gt, "fix": fix}
    if audio:
        a = torch.zeros(B, 1, arch.AUDIO_LEN, 1)
        n = 47040
        s0 = (arch.AUDIO_LEN - n) // 2
        a[:, 0, s0:s0 + n, 0] = 0.05 * torch.randn(B, n, generator=g) * torch.hann_window(n)
        out["audio"] = a
    return out


LOSS_CASES = {"a": (3, 32, 48), "b": (2, 224, 384)}


def make_loss_inputs(tag):
    """Seeded (s_map, gt, fixation) triples for the loss goldens (tests/golden/losses.npz)."""
    B, H, W = LOSS_CASES[tag]
    g = torch.Generator().manual_seed(77 + ord(tag))
    s = torch.rand(B, H, W, generator=g)
    gt = torch.rand(B, H, W, generator=g) ** 3 + 1e-4
    fix = (torch.rand(B, H, W, generator=g) > 0.97).float()
    return s, gt, fix
