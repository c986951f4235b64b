"""Drop-in ``VideoAudioSaliencyModel`` (AViNet, reference model.py:191-249; ``use_transformer=True`` through vinet_b200/xfmr.py).

SoundNet (model.py:746-825) runs as seven fp32 conv1d + fused BN/ReLU/MaxPool kernels (csrc/audio.cu);
its (B,1024,3) output is fused with the max-pooled top backbone feature by the bilinear kernel
(model.py:229-237) and the result replaces y0 at the decoder input.
"""
import ctypes as C
import os

import torch
import torch.nn as nn

from . import arch
from . import lib as L
from .engine import Act
from .model import BNParams, ConvParams, VideoSaliencyModel, _PlanModule, backbone_plan, decoder_plan, pack_input
from . import xfmr


class SoundNet(nn.Module):
    """Parameter layout of the reference SoundNet (model.py:746-791), including the unused conv8_* heads."""

    def __init__(self):
        super().__init__()
        for i, (cin, cout, k, p, pool) in enumerate(arch.SOUNDNET, 1):
            setattr(self, "conv%d" % i, ConvParams(cin, cout, (k, 1), bias=True))
            setattr(self, "batchnorm%d" % i, BNParams(cout, eps=1e-5, momentum=0.1))
        self.conv8_objs = ConvParams(1024, 1000, (8, 1), bias=True)
        self.conv8_scns = ConvParams(1024, 401, (8, 1), bias=True)


class BilinearParams(nn.Module):
    def __init__(self, in1, in2, out):
        super().__init__()
        bound = 1.0 / in1 ** 0.5
        self.weight = nn.Parameter(torch.empty(out, in1, in2).uniform_(-bound, bound))
        self.bias = nn.Parameter(torch.empty(out).uniform_(-bound, bound))


def soundnet_plan(e, pfx, net, audio):
    """SoundNet.forward (model.py:793-825) -> fp32 (B,1024,3) tensor; backward reads ``ga`` (grad of it)."""
    B = audio.shape[0]
    assert audio.shape[1:] == (1, arch.AUDIO_LEN, 1) and audio.dtype == torch.float32
    x = audio.contiguous().view(B, 1, arch.AUDIO_LEN)
    st = e.stream()
    layers = []
    for i, (cin, cout, k, p, pool) in enumerate(arch.SOUNDNET, 1):
        conv, bn = getattr(net, "conv%d" % i), getattr(net, "batchnorm%d" % i)
        Lin = x.shape[2]
        Lout = (Lin + 2 * p - k) // 2 + 1
        y = e.buf("%sy%d" % (pfx, i), (B, cout, Lout), torch.float32)
        d = L.Conv1d()
        d.x, d.w, d.bias = x.data_ptr(), conv.weight.data_ptr(), conv.bias.data_ptr()
        d.B, d.Cin, d.Lin, d.Cout, d.Lout, d.k, d.stride, d.pad, d.y = B, cin, Lin, cout, Lout, k, 2, p, y.data_ptr()
        e.lib.call("vinet_conv1d_fwd", C.byref(d), st)
        out = e.buf("%so%d" % (pfx, i), (B, cout, Lout // pool), torch.float32)
        stat = e.buf("%sstat%d" % (pfx, i), (2, cout), torch.float32)
        b = L.Bn1d()
        b.y, b.B, b.C, b.L, b.pool = y.data_ptr(), B, cout, Lout, pool
        b.gamma, b.beta, b.eps, b.momentum = bn.weight.data_ptr(), bn.bias.data_ptr(), bn.eps, bn.momentum
        b.running_mean, b.running_var, b.training = bn.running_mean.data_ptr(), bn.running_var.data_ptr(), int(e.training)
        b.mean, b.invstd, b.out = stat[0].data_ptr(), stat[1].data_ptr(), out.data_ptr()
        e.lib.call("vinet_bn1d_fwd", C.byref(b), st)
        if e.training:
            bn.num_batches_tracked += 1
        layers.append((i, d, b, x, y, out, conv, bn))
        x = out
    ga = e.buf(pfx + "ga", tuple(x.shape), torch.float32) if e.record else None
    if e.record:
        def backward():
            g = ga
            for i, d, b, xin, y, out, conv, bn in reversed(layers):
                dy = e.buf("%sdy%d" % (pfx, i), tuple(y.shape), torch.float32)
                dgam, dbet = e.grad_tensor("%sbatchnorm%d.weight" % (pfx, i), bn.weight), e.grad_tensor("%sbatchnorm%d.bias" % (pfx, i), bn.bias)
                b.gout, b.dy, b.dgamma, b.dbeta = g.data_ptr(), dy.data_ptr(), dgam.data_ptr(), dbet.data_ptr()
                e.lib.call("vinet_bn1d_bwd", C.byref(b), e.stream())
                gw, gb = e.grad_tensor("%sconv%d.weight" % (pfx, i), conv.weight), e.grad_tensor("%sconv%d.bias" % (pfx, i), conv.bias)
                dx = e.buf("%sdx%d" % (pfx, i), tuple(xin.shape), torch.float32) if i > 1 else None
                d.dy, d.dx, d.dw, d.dbias = dy.data_ptr(), (dx.data_ptr() if dx is not None else None), gw.data_ptr(), gb.data_ptr()
                e.lib.call("vinet_conv1d_bwd", C.byref(d), e.stream())
                e.param_grads["%sconv%d.weight" % (pfx, i)] = gw
                e.param_grads["%sconv%d.bias" % (pfx, i)] = gb
                e.param_grads["%sbatchnorm%d.weight" % (pfx, i)] = dgam
                e.param_grads["%sbatchnorm%d.bias" % (pfx, i)] = dbet
                g = dx
        e.tape.append(backward)
    return x, ga


def avfuse_plan(e, name, y0, a, ga, bil):
    """MaxPool3d((4,1,1),stride=(2,1,2)) + nn.Bilinear(42,3,336) + view(B,1024,4,7,12) (model.py:235-237)."""
    assert (y0.T, y0.H, y0.W, y0.C) == (4, 7, 12, 1024), "AViNet is hard-wired to 32x224x384 clips (model.py:230,237)"
    B = y0.B
    out = e.new_act(name + ".fused", B, 4, 7, 12, 1024, gdtype=torch.float32)
    vbuf = e.buf(name + ".v", (B, 1024, 42), torch.float32)
    d = L.AvFuse()
    d.y0, d.ld, d.dtype, d.xform = y0.ptr(), y0.ld, e.dt, y0.xform
    d.scale = None if y0.scale is None else y0.scale.data_ptr()
    d.shift = None if y0.shift is None else y0.shift.data_ptr()
    d.audio, d.w, d.bias, d.B, d.C = a.data_ptr(), bil.weight.data_ptr(), bil.bias.data_ptr(), B, 1024
    d.vbuf, d.out, d.ldo, d.out_dtype = vbuf.data_ptr(), out.ptr(), out.ld, e.dt
    e.lib.call("vinet_avfuse_fwd", C.byref(d), e.stream())
    if e.record:
        def backward():
            gw, gb = e.grad_tensor(name + ".weight", bil.weight), e.grad_tensor(name + ".bias", bil.bias)
            assert out.gdt == L.F32 and y0.gdt == L.F32, "the AV fusion kernel keeps fp32 gradients"
            e.ensure_init(y0)
            d.gout, d.ldgo, d.gy0, d.ldgy0 = out.gptr(), out.ldg, y0.gptr(), y0.ldg
            d.gaudio, d.dw, d.dbias = ga.data_ptr(), gw.data_ptr(), gb.data_ptr()
            e.lib.call("vinet_avfuse_bwd", C.byref(d), e.stream())
            e.param_grads[name + ".weight"] = gw
            e.param_grads[name + ".bias"] = gb
        e.tape.append(backward)
    return out


class VideoAudioSaliencyModel(_PlanModule):
    """AViNet (model.py:191-249). ``soundnet_weights``: None -> load './soundnet8_final.pth' relative to
    the cwd exactly like the reference (model.py:224); a path -> load that file; False -> keep the random
    initialisation (tests / synthetic benchmarks on a box without the checkpoint)."""

    _n_extra = 1

    def __init__(self, use_transformer=False, transformer_in_channel=32, num_encoder_layers=3, nhead=4,
                 use_upsample=True, num_hier=3, num_clips=32, soundnet_weights=None):
        super().__init__()
        self.use_transformer = use_transformer
        self.visual_model = VideoSaliencyModel(transformer_in_channel=transformer_in_channel, nhead=nhead,
                                               use_upsample=use_upsample, num_hier=num_hier, num_clips=num_clips)
        if use_transformer:                   # model.py:211-221, registered in the reference's order (state_dict key order)
            self.conv_in_1x1 = ConvParams(1024, transformer_in_channel, (1, 1, 1), bias=True)
            self.conv_out_1x1 = ConvParams(32, 1024, (1, 1, 1), bias=True)
            self.transformer = xfmr.Transformer(4 * 7 * 12, hidden_size=4 * 7 * 12, nhead=nhead, num_encoder_layers=num_encoder_layers,
                                                num_decoder_layers=-1, max_len=transformer_in_channel)
        self.audionet = SoundNet()
        if soundnet_weights is not False:
            path = "./soundnet8_final.pth" if soundnet_weights is None else soundnet_weights
            self.audionet.load_state_dict(torch.load(path))
            print("Loaded SoundNet Weights")
        for param in self.audionet.parameters():
            param.requires_grad = True
        self.bilinear = BilinearParams(42, 3, 4 * 7 * 12)

    def _plan_uses(self, name):
        return "conv8_" not in name          # parameters without a gradient in the reference as well

    def forward(self, x, audio):
        return self._call_plan(x, audio)

    def _run_plan(self, e, record, x, audio):
        e.generation += 1
        e.begin(x.device, self.training, record)
        # the audio branch is independent of the backbone until the fusion: its ~20 small launches run on the side stream
        audio_side = e.fork()
        e.on_side(True)
        a, ga = soundnet_plan(e, "audionet.", self.audionet, audio)
        if audio_side:
            e._side_active = False          # keep the fork open: the backbone below forks / joins the same side stream itself
        xin = pack_input(e, x)
        y0, y1, y2, y3 = backbone_plan(e, "visual_model.backbone.", self.visual_model.backbone, xin,
                                       y0_gdtype=torch.float32)
        fused = avfuse_plan(e, "bilinear", y0, a, ga, self.bilinear)
        if self.use_transformer:
            fused = xfmr.avinet_transformer_plan(e, self, fused)
        e.mark_backward_point("decoder")
        out = decoder_plan(e, "visual_model.decoder.", self.visual_model.decoder, fused, y1, y2, y3)
        e.end_forward()
        return out


class VideoAudioSaliencyFusionModel(_PlanModule):
    """Drop-in ``VideoAudioSaliencyFusionModel`` (model.py:116-189): visual and audio tokens through one Transformer
    (vinet_b200/xfmr.py).  Same constructor and ``state_dict`` as the reference; ``use_transformer`` is stored and ignored like
    there; ``soundnet_weights`` as in ``VideoAudioSaliencyModel``.  The registered-but-unused ``bilinear`` gets no gradient."""

    _n_extra = 1

    def __init__(self, use_transformer=True, transformer_in_channel=512, num_encoder_layers=3, nhead=4, use_upsample=True, num_hier=3,
                 num_clips=32, soundnet_weights=None):
        super().__init__()
        self.use_transformer = use_transformer
        self.visual_model = VideoSaliencyModel(transformer_in_channel=transformer_in_channel, nhead=nhead,
                                               use_upsample=use_upsample, num_hier=num_hier, num_clips=num_clips)
        self.conv_in_1x1 = ConvParams(1024, transformer_in_channel, (1, 1, 1), bias=True)
        self.transformer = xfmr.Transformer(transformer_in_channel, hidden_size=transformer_in_channel, nhead=nhead,
                                            num_encoder_layers=num_encoder_layers, num_decoder_layers=-1, max_len=4 * 7 * 12 + 3)
        self.audionet = SoundNet()
        self.audio_conv_1x1 = ConvParams(1024, transformer_in_channel, (1, 1), bias=True)
        if soundnet_weights is not False:
            path = "./soundnet8_final.pth" if soundnet_weights is None else soundnet_weights
            self.audionet.load_state_dict(torch.load(path))
            print("Loaded SoundNet Weights")
        for param in self.audionet.parameters():
            param.requires_grad = True
        self.bilinear = BilinearParams(42, 3, 4 * 7 * 12)

    def _plan_uses(self, name):
        return "conv8_" not in name and not name.startswith("bilinear.")

    def forward(self, x, audio):
        return self._call_plan(x, audio)

    def _run_plan(self, e, record, x, audio):
        e.generation += 1
        e.begin(x.device, self.training, record)
        audio_side = e.fork()
        e.on_side(True)
        a, ga = soundnet_plan(e, "audionet.", self.audionet, audio)
        if audio_side:
            e._side_active = False
        xin = pack_input(e, x)
        y0, y1, y2, y3 = backbone_plan(e, "visual_model.backbone.", self.visual_model.backbone, xin, y0_gdtype=torch.float32)
        assert (y0.T, y0.H, y0.W, y0.C) == (4, 7, 12, 1024), "the fusion model is hard-wired to 32x224x384 clips (model.py:142,174)"
        e.join()                 # the audio tokens are read next
        fused = xfmr.fusion_plan(e, self, y0, a, ga)
        e.mark_backward_point("decoder")
        out = decoder_plan(e, "visual_model.decoder.", self.visual_model.decoder, fused, y1, y2, y3)
        e.end_forward()
        return out
