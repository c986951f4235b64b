"""Transformer fusion of the audio-visual variants: the reference's ``Transformer`` /
``PositionalEncoding`` (model.py:8-69, encoder-only configuration: ``num_decoder_layers=-1``, ``spatial_dim=-1``, the only one any
model instantiates), its use in ``VideoAudioSaliencyModel(use_transformer=True)`` (model.py:211-221, 239-247) and in
``VideoAudioSaliencyFusionModel`` (model.py:116-189).

Parameters live in torch's own ``nn.TransformerEncoder`` (a parameter HOLDER here: same ``state_dict`` keys, shapes and
initialisation as the reference; its ``forward`` is never called).  All computation is a plan of C-ABI calls (csrc/xfmr.cu):
tokens are kept sequence-major, row ``s * B + b`` of an fp32 ``[S * B, d]`` matrix - exactly the memory order of the reference's
``(S, B, d)`` tensors - and every linear layer, attention product, 1x1 convolution with bias, permute / flatten / cat / mean / repeat
and bias-gradient reduction is one ``vinet_bgemm`` over strided views.  Dropout (p = 0.1 at four places per encoder layer in train
mode, like ``nn.TransformerEncoderLayer``) uses the library's counter-based generator: statistically equivalent to, not bit-identical
with, torch's Philox stream (parity tests run with p = 0 or in eval mode).
"""
import math
import threading
from types import SimpleNamespace

import torch
import torch.nn as nn

from . import lib as L


class PositionalEncoding(nn.Module):
    """The reference's sinusoidal table (model.py:8-26) as the buffer ``pe`` of shape (max_len, 1, feat): token s, feature 2i / 2i+1
    = sin / cos(s * 10000^(-2i / feat)).  The module's Dropout exists but is never applied (model.py:24-26)."""

    def __init__(self, feat_size, dropout=0.1, max_len=4):
        super().__init__()
        self.dropout = nn.Dropout(p=dropout)
        token = torch.arange(max_len, dtype=torch.float32)[:, None]
        freq = torch.exp(torch.arange(0, feat_size, 2, dtype=torch.float32) * (-math.log(10000.0) / feat_size))
        table = torch.zeros(max_len, feat_size)
        table[:, 0::2], table[:, 1::2] = torch.sin(token * freq), torch.cos(token * freq)
        self.register_buffer("pe", table[:, None, :].contiguous())


class Transformer(nn.Module):
    """Parameter layout of the reference ``Transformer`` (model.py:28-46) in its encoder-only form."""

    def __init__(self, feat_size, hidden_size=256, nhead=4, num_encoder_layers=3, max_len=4, num_decoder_layers=-1, num_queries=4,
                 spatial_dim=-1):
        super().__init__()
        if num_decoder_layers != -1 or spatial_dim != -1:
            raise NotImplementedError("only the encoder-only Transformer the reference's models instantiate (model.py:138-145, 215-222)")
        self.pos_encoder = PositionalEncoding(feat_size, max_len=max_len)
        self.spatial_dim, self.use_decoder = spatial_dim, False
        self.feat_size, self.hidden_size, self.nhead, self.max_len = feat_size, hidden_size, nhead, max_len
        self.transformer_encoder = nn.TransformerEncoder(nn.TransformerEncoderLayer(feat_size, nhead, hidden_size), num_encoder_layers)


def _ptr(t, off=0):
    return t.data_ptr() + off * t.element_size()


# set while a backward closure runs: gradient reductions may be split along K (vinet_bgemm_t.accumulate bit 1).  Per thread:
# nn.DataParallel runs the replicas' forwards and backwards on one thread per device.
_STATE = threading.local()


def _backward_pass(fn):
    def run():
        _STATE.order_free = True
        try:
            fn()
        finally:
            _STATE.order_free = False
    return run


def gemm(e, M, N, K, A, sA, B, sB, Cp, sC, nb=(1, 1), alpha=1.0, a_dtype=L.F32, b_dtype=L.F32, c_dtype=L.F32, bias1=None, bias2=None,
         relu=0, acc=0, a_xf=None, a_xf_on_m=0):
    """One vinet_bgemm launch.  A / B / Cp are device addresses; sA = (sm, sk[, sb1, sb2]), sB = (sn, sk[, ...]), sC = (sm, sn[, ...])
    in elements; bias1 = (addr, sm, sn); bias2 = (addr, sm, sn, sb1, sb2); a_xf = (scale addr or None, shift addr or None, relu)."""
    d = L.Bgemm()
    sA, sB, sC = (tuple(sA) + (0, 0))[:4], (tuple(sB) + (0, 0))[:4], (tuple(sC) + (0, 0))[:4]
    d.A, d.sAm, d.sAk, d.sAb1, d.sAb2, d.a_dtype = A, sA[0], sA[1], sA[2], sA[3], a_dtype
    d.B, d.sBn, d.sBk, d.sBb1, d.sBb2, d.b_dtype = B, sB[0], sB[1], sB[2], sB[3], b_dtype
    d.C, d.sCm, d.sCn, d.sCb1, d.sCb2, d.c_dtype = Cp, sC[0], sC[1], sC[2], sC[3], c_dtype
    d.M, d.N, d.K, d.nb1, d.nb2, d.alpha, d.relu, d.accumulate = M, N, K, nb[0], nb[1], alpha, relu, int(acc) | (2 if getattr(_STATE, "order_free", False) else 0)
    if a_xf is not None:
        d.a_scale, d.a_shift, d.a_relu = a_xf
        d.a_xf_on_m = a_xf_on_m
    if bias1 is not None:
        d.bias1, d.s1m, d.s1n = bias1
    if bias2 is not None:
        b2 = (tuple(bias2) + (0, 0))[:5]
        d.bias2, d.s2m, d.s2n, d.s2b1, d.s2b2 = b2
    e.call("vinet_bgemm", d)


def _on_device(t, dev):
    """`t` lives on the engine's device (an engine begun on torch.device("cuda") holds tensors that report cuda:<current>)."""
    return t is not None and t.device.type == dev.type and (dev.index is None or t.device.index == dev.index)


def ones(e):
    t = e.pool.get("xfmr.one")
    if not _on_device(t, e.device):
        t = e.pool["xfmr.one"] = torch.ones(4, dtype=torch.float32, device=e.device)
    return t


def colsum(e, x, rows, n, out):
    """out[n] = sum over rows of x[rows, n] (bias gradients): a GEMM against a broadcast one."""
    gemm(e, n, 1, rows, _ptr(x), (1, n), _ptr(ones(e)), (0, 0), _ptr(out), (1, 0))


def linear(e, x, rows, k, w, b, out, n, relu=0):
    """out[rows, n] = x[rows, k] @ w[n, k]^T + b (F.linear)."""
    gemm(e, rows, n, k, _ptr(x), (k, 1), _ptr(w), (k, 1), _ptr(out), (n, 1), bias1=(_ptr(b), 0, 1), relu=relu)


def linear_bwd(e, pname, x, rows, k, w, b, dy, n, dx=None, dx_acc=0):
    """Gradients of ``linear``: dw = dy^T x, db = colsum(dy), and (optionally) dx (+)= dy @ w."""
    gw, gb = e.grad_tensor(pname + "weight", w), e.grad_tensor(pname + "bias", b)
    gemm(e, n, k, rows, _ptr(dy), (1, n), _ptr(x), (1, k), _ptr(gw), (k, 1))
    colsum(e, dy, rows, n, gb)
    e.param_grads[pname + "weight"], e.param_grads[pname + "bias"] = gw, gb
    if dx is not None:
        gemm(e, rows, k, n, _ptr(dy), (n, 1), _ptr(w), (1, k), _ptr(dx), (k, 1), acc=dx_acc)


def _rng(e):
    t = e.pool.get("xfmr.rng")
    if not _on_device(t, e.device):
        t = e.pool["xfmr.rng"] = torch.tensor([torch.initial_seed() & 0x7FFFFFFF, 0], dtype=torch.int64, device=e.device)
    return t


def _dropout(e, name, x, n, p, salt, out=None):
    """Train-mode nn.Dropout on n fp32 elements; returns (result tensor, mask or None).  p == 0: no launch."""
    if p <= 0.0:
        return x, None
    mask = e.buf(name + ".mask", (n,), torch.uint8)
    y = x if out is None else out
    e.lib.call("vinet_dropout_fwd", _ptr(x), _ptr(y), _ptr(mask), n, p, _ptr(_rng(e)), salt, e.stream())
    return y, mask


def _dropout_bwd(e, g, out, mask, relu_ref, n, p):
    e.lib.call("vinet_dropout_bwd", _ptr(g), _ptr(out), None if mask is None else _ptr(mask), None if relu_ref is None else _ptr(relu_ref),
               n, p, e.stream())


def encoder_plan(e, pfx, tr, x, S, B):
    """nn.TransformerEncoder(layers of nn.TransformerEncoderLayer(d, nhead, ff), post-norm, ReLU) on x = fp32 [S * B, d] (the
    positional table is already added).  Returns (y, gy, gx): the output, the buffer the caller's backward fills with dL/dy, and
    the buffer that holds dL/dx once this plan's backward has run (None without a tape)."""
    enc = tr.transformer_encoder
    d, H = tr.feat_size, tr.nhead
    dh, R = d // H, S * B
    assert d % H == 0 and d <= 512 and tr.hidden_size <= 512, "csrc/xfmr.cu: LayerNorm rows of at most 512 features"
    scale = 1.0 / math.sqrt(dh)
    f32 = torch.float32
    train = e.training
    if train and any(l.dropout.p > 0 or l.dropout1.p > 0 or l.dropout2.p > 0 or l.self_attn.dropout > 0 for l in enc.layers):
        e.lib.call("vinet_rng_advance", _ptr(_rng(e)), e.stream())
    saved = []
    for i, l in enumerate(enc.layers):
        n = "%s%d." % (pfx, i)
        at = l.self_attn
        ff = l.linear1.out_features
        pa, p1, pf, p2 = (at.dropout, l.dropout1.p, l.dropout.p, l.dropout2.p) if train else (0.0, 0.0, 0.0, 0.0)
        qkv = e.buf(n + "qkv", (R, 3 * d), f32)
        linear(e, x, R, d, at.in_proj_weight, at.in_proj_bias, qkv, 3 * d)
        P = e.buf(n + "p", (B, H, S, S), f32)
        hs = (B * 3 * d, 1, 3 * d, dh)                     # a head's (token, feature) view of the packed q | k | v rows
        gemm(e, S, S, dh, _ptr(qkv), hs, _ptr(qkv, d), hs, _ptr(P), (S, 1, H * S * S, S * S), nb=(B, H), alpha=scale)
        e.lib.call("vinet_softmax_fwd", _ptr(P), B * H * S, S, e.stream())
        Pd, mask_a = _dropout(e, n + "pd", P, B * H * S * S, pa, 4 * i, out=e.buf(n + "pd", (B, H, S, S), f32) if pa > 0 else None)
        ctx = e.buf(n + "ctx", (R, d), f32)
        gemm(e, S, dh, S, _ptr(Pd), (S, 1, H * S * S, S * S), _ptr(qkv, 2 * d), (1, B * 3 * d, 3 * d, dh), _ptr(ctx), (B * d, 1, d, dh), nb=(B, H))
        ao = e.buf(n + "ao", (R, d), f32)
        linear(e, ctx, R, d, at.out_proj.weight, at.out_proj.bias, ao, d)
        _, mask_1 = _dropout(e, n + "ao", ao, R * d, p1, 4 * i + 1)
        x1, z1, st1 = e.buf(n + "x1", (R, d), f32), e.buf(n + "z1", (R, d), f32), e.buf(n + "st1", (R, 2), f32)
        ln1 = L.AddLn()
        ln1.x, ln1.y, ln1.z, ln1.stat, ln1.gamma, ln1.beta = _ptr(x), _ptr(ao), _ptr(z1), _ptr(st1), _ptr(l.norm1.weight), _ptr(l.norm1.bias)
        ln1.eps, ln1.n, ln1.rows, ln1.out = l.norm1.eps, d, R, _ptr(x1)
        e.call("vinet_add_layernorm_fwd", ln1)
        h = e.buf(n + "h", (R, ff), f32)
        linear(e, x1, R, d, l.linear1.weight, l.linear1.bias, h, ff, relu=1)
        hd, mask_f = _dropout(e, n + "hd", h, R * ff, pf, 4 * i + 2, out=e.buf(n + "hd", (R, ff), f32) if pf > 0 else None)
        f = e.buf(n + "f", (R, d), f32)
        linear(e, hd, R, ff, l.linear2.weight, l.linear2.bias, f, d)
        _, mask_2 = _dropout(e, n + "f", f, R * d, p2, 4 * i + 3)
        x2, z2, st2 = e.buf(n + "x2", (R, d), f32), e.buf(n + "z2", (R, d), f32), e.buf(n + "st2", (R, 2), f32)
        ln2 = L.AddLn()
        ln2.x, ln2.y, ln2.z, ln2.stat, ln2.gamma, ln2.beta = _ptr(x1), _ptr(f), _ptr(z2), _ptr(st2), _ptr(l.norm2.weight), _ptr(l.norm2.bias)
        ln2.eps, ln2.n, ln2.rows, ln2.out = l.norm2.eps, d, R, _ptr(x2)
        e.call("vinet_add_layernorm_fwd", ln2)
        # what the backward reads (the descriptors hold raw addresses: `keep` pins the buffers behind them)
        saved.append(SimpleNamespace(n=n, l=l, xin=x, qkv=qkv, P=P, Pd=Pd, mask_a=mask_a, ctx=ctx, mask_1=mask_1, x1=x1, ln1=ln1, h=h, hd=hd,
                                     mask_f=mask_f, mask_2=mask_2, ln2=ln2, p=(pa, p1, pf, p2), ff=ff, keep=(z1, st1, z2, st2, ao, f, x2)))
        x = x2
    if not e.record:
        return x, None, None
    gy = e.buf(pfx + "gy", (R, d), f32)
    gx_box = [None]

    def backward():
        g = gy
        for sv in reversed(saved):
            n, l, xin, qkv, P, Pd, ctx, x1, ln1, ln2, h, hd, ff = sv.n, sv.l, sv.xin, sv.qkv, sv.P, sv.Pd, sv.ctx, sv.x1, sv.ln1, sv.ln2, sv.h, sv.hd, sv.ff
            mask_a, mask_1, mask_f, mask_2, (pa, p1, pf, p2) = sv.mask_a, sv.mask_1, sv.mask_f, sv.mask_2, sv.p
            at, pn = l.self_attn, n
            # ---- norm2(x1 + dropout2(linear2(dropout(relu(linear1(x1))))))
            dz2 = e.buf(n + "dz2", (R, d), f32)
            gg2, gb2 = e.grad_tensor(pn + "norm2.weight", l.norm2.weight, zero=True), e.grad_tensor(pn + "norm2.bias", l.norm2.bias, zero=True)
            ln2.gout, ln2.dz, ln2.dgamma, ln2.dbeta = _ptr(g), _ptr(dz2), _ptr(gg2), _ptr(gb2)
            e.call("vinet_add_layernorm_bwd", ln2)
            e.param_grads[pn + "norm2.weight"], e.param_grads[pn + "norm2.bias"] = gg2, gb2
            df = dz2
            if mask_2 is not None:
                df = e.buf(n + "dtmp", (R, d), f32)
                _dropout_bwd(e, dz2, df, mask_2, None, R * d, p2)
            dh_ = e.buf(n + "dh", (R, ff), f32)
            linear_bwd(e, pn + "linear2.", hd, R, ff, l.linear2.weight, l.linear2.bias, df, d, dx=dh_)
            _dropout_bwd(e, dh_, dh_, mask_f, h, R * ff, pf)                    # dropout mask and the ReLU of linear1
            linear_bwd(e, pn + "linear1.", x1, R, d, l.linear1.weight, l.linear1.bias, dh_, ff, dx=dz2, dx_acc=1)   # dz2 = dL/dx1 now
            # ---- norm1(x + dropout1(out_proj(attention(x))))
            dz1 = e.buf(n + "dz1", (R, d), f32)
            gg1, gb1 = e.grad_tensor(pn + "norm1.weight", l.norm1.weight, zero=True), e.grad_tensor(pn + "norm1.bias", l.norm1.bias, zero=True)
            ln1.gout, ln1.dz, ln1.dgamma, ln1.dbeta = _ptr(dz2), _ptr(dz1), _ptr(gg1), _ptr(gb1)
            e.call("vinet_add_layernorm_bwd", ln1)
            e.param_grads[pn + "norm1.weight"], e.param_grads[pn + "norm1.bias"] = gg1, gb1
            dao = dz1
            if mask_1 is not None:
                dao = e.buf(n + "dtmp", (R, d), f32)
                _dropout_bwd(e, dz1, dao, mask_1, None, R * d, p1)
            dctx = e.buf(n + "dctx", (R, d), f32)
            linear_bwd(e, pn + "self_attn.out_proj.", ctx, R, d, at.out_proj.weight, at.out_proj.bias, dao, d, dx=dctx)
            # ---- softmax(q k^T / sqrt(dh)) v per (clip, head)
            dqkv = e.buf(n + "dqkv", (R, 3 * d), f32)
            dP = e.buf(n + "dp", (B, H, S, S), f32)
            ps, cs, hs = (S, 1, H * S * S, S * S), (B * d, 1, d, dh), (B * 3 * d, 1, 3 * d, dh)
            gemm(e, S, S, dh, _ptr(dctx), cs, _ptr(qkv, 2 * d), hs, _ptr(dP), ps, nb=(B, H))                                    # dP = dctx v^T
            gemm(e, S, dh, S, _ptr(Pd), (1, S) + ps[2:], _ptr(dctx), (1, B * d, d, dh), _ptr(dqkv, 2 * d), hs, nb=(B, H))      # dv = P^T dctx
            if mask_a is not None:
                _dropout_bwd(e, dP, dP, mask_a, None, B * H * S * S, pa)
            e.lib.call("vinet_softmax_bwd", _ptr(P), _ptr(dP), B * H * S, S, e.stream())
            gemm(e, S, dh, S, _ptr(dP), ps, _ptr(qkv, d), (1, B * 3 * d, 3 * d, dh), _ptr(dqkv), hs, nb=(B, H), alpha=scale)      # dq = dS k
            gemm(e, S, dh, S, _ptr(dP), (1, S) + ps[2:], _ptr(qkv), (1, B * 3 * d, 3 * d, dh), _ptr(dqkv, d), hs, nb=(B, H), alpha=scale)  # dk = dS^T q
            gw, gb = e.grad_tensor(pn + "self_attn.in_proj_weight", at.in_proj_weight), e.grad_tensor(pn + "self_attn.in_proj_bias", at.in_proj_bias)
            gemm(e, 3 * d, d, R, _ptr(dqkv), (1, 3 * d), _ptr(xin), (1, d), _ptr(gw), (d, 1))
            colsum(e, dqkv, R, 3 * d, gb)
            e.param_grads[pn + "self_attn.in_proj_weight"], e.param_grads[pn + "self_attn.in_proj_bias"] = gw, gb
            gemm(e, R, d, 3 * d, _ptr(dqkv), (3 * d, 1), _ptr(at.in_proj_weight), (1, d), _ptr(dz1), (d, 1), acc=1)               # dz1 = dL/dx now
            g = dz1
        gx_box[0] = g

    e.tape.append(_backward_pass(backward))
    return x, gy, gx_box


def avinet_transformer_plan(e, m, fused):
    """``use_transformer=True`` (model.py:239-247): conv_in_1x1 (1024 -> C' channels, bias) on the fused feature, the C' channels
    become the TOKENS and the 4*7*12 = 336 positions their features (flatten(2) + permute(1,0,2)), positional table, encoder,
    and conv_out_1x1 (C' -> 1024, bias) back to the decoder's input.  Three launches besides the encoder: the permutes are strides."""
    B, Pn, Cin = fused.B, fused.T * fused.H * fused.W, fused.C
    tr, cin, cout = m.transformer, m.conv_in_1x1, m.conv_out_1x1
    S, Cout = cin.weight.shape[0], cout.weight.shape[0]
    assert Pn == tr.feat_size and S == tr.max_len and fused.xform == L.XF_IDENT and Cin == cin.weight.shape[1] and cout.weight.shape[1] == S
    f32 = torch.float32
    x0 = e.buf("xf.x0", (S * B, Pn), f32)
    gemm(e, S, Pn, Cin, _ptr(cin.weight), (Cin, 1), fused.ptr(), (fused.ld, 1, Pn * fused.ld), _ptr(x0), (B * Pn, 1, Pn), nb=(B, 1),
         b_dtype=e.dt, bias1=(_ptr(cin.bias), 1, 0), bias2=(_ptr(tr.pos_encoder.pe), Pn, 1))
    box = {}
    if e.record:
        def conv_in_backward():
            gx = box["gx"][0]
            assert fused.gdt == L.F32, "the transformer fusion keeps fp32 gradients"
            gw, gb = e.grad_tensor("conv_in_1x1.weight", cin.weight), e.grad_tensor("conv_in_1x1.bias", cin.bias)
            gemm(e, S, Cin, B * Pn, _ptr(gx), (B * Pn, 1), fused.ptr(), (1, fused.ld), _ptr(gw), (Cin, 1), b_dtype=e.dt)
            gemm(e, S, 1, B * Pn, _ptr(gx), (B * Pn, 1), _ptr(ones(e)), (0, 0), _ptr(gb), (1, 0))
            first = e.first_write(fused)
            gemm(e, Pn, Cin, S, _ptr(gx), (1, B * Pn, Pn), _ptr(cin.weight), (1, Cin), fused.gptr(), (fused.ldg, 1, Pn * fused.ldg), nb=(B, 1),
                 acc=0 if first else 1)
            e.param_grads["conv_in_1x1.weight"], e.param_grads["conv_in_1x1.bias"] = gw, gb
        e.tape.append(_backward_pass(conv_in_backward))
    y, gy, box["gx"] = encoder_plan(e, "transformer.transformer_encoder.layers.", tr, x0, S, B)
    out = e.new_act("xf.out", B, fused.T, fused.H, fused.W, Cout, gdtype=f32)
    gemm(e, Pn, Cout, S, _ptr(y), (1, B * Pn, Pn), _ptr(cout.weight), (S, 1), out.ptr(), (out.ld, 1, Pn * out.ld), nb=(B, 1), c_dtype=e.dt,
         bias1=(_ptr(cout.bias), 0, 1))
    if e.record:
        def conv_out_backward():
            assert out.gdt == L.F32
            g, ldg = out.gptr(), out.ldg
            gw, gb = e.grad_tensor("conv_out_1x1.weight", cout.weight), e.grad_tensor("conv_out_1x1.bias", cout.bias)
            gemm(e, S, Pn, Cout, _ptr(cout.weight), (1, S), g, (ldg, 1, Pn * ldg), _ptr(gy), (B * Pn, 1, Pn), nb=(B, 1))
            gemm(e, Cout, S, B * Pn, g, (1, ldg), _ptr(y), (B * Pn, 1), _ptr(gw), (S, 1))
            gemm(e, Cout, 1, B * Pn, g, (1, ldg), _ptr(ones(e)), (0, 0), _ptr(gb), (1, 0))
            e.param_grads["conv_out_1x1.weight"], e.param_grads["conv_out_1x1.bias"] = gw, gb
        e.tape.append(_backward_pass(conv_out_backward))
    return out


def _act_xf(a):
    """Pending transform of an activation as vinet_bgemm's A-side read transform."""
    if a.xform == L.XF_IDENT:
        return None
    assert not (a.xform & L.XF_UP2)
    aff = bool(a.xform & 2)
    return (a.scale.data_ptr() if aff else None, a.shift.data_ptr() if aff else None, a.xform & 1)


def fusion_plan(e, m, y0, a, ga):
    """VideoAudioSaliencyFusionModel.forward between the backbone and the decoder (model.py:151-186): 336 visual tokens
    (conv_in_1x1 of y0, one per position) and 3 audio tokens (audio_conv_1x1 of the SoundNet output) of C' features, written
    straight into one sequence-major token matrix (flatten / cat / permute are strides), positional table, encoder, and the
    decoder input [visual tokens | mean audio token broadcast over the 336 positions] written straight into an NDHWC activation."""
    B, Pn, Cin = y0.B, y0.T * y0.H * y0.W, y0.C
    tr, cv, ca = m.transformer, m.conv_in_1x1, m.audio_conv_1x1
    d = cv.weight.shape[0]
    na = a.shape[2]
    S = Pn + na
    assert d == tr.feat_size and S == tr.max_len and Cin == cv.weight.shape[1] == ca.weight.shape[1] and tuple(a.shape) == (B, Cin, na)
    f32 = torch.float32
    xf = _act_xf(y0)
    pe = tr.pos_encoder.pe
    x0 = e.buf("fu.x0", (S * B, d), f32)
    gemm(e, Pn, d, Cin, y0.ptr(), (y0.ld, 1, Pn * y0.ld), _ptr(cv.weight), (Cin, 1), _ptr(x0), (B * d, 1, d), nb=(B, 1), a_dtype=e.dt, a_xf=xf,
         bias1=(_ptr(cv.bias), 0, 1), bias2=(_ptr(pe), d, 1))
    gemm(e, na, d, Cin, _ptr(a), (1, na, Cin * na), _ptr(ca.weight), (Cin, 1), _ptr(x0, Pn * B * d), (B * d, 1, d), nb=(B, 1),
         bias1=(_ptr(ca.bias), 0, 1), bias2=(_ptr(pe, Pn * d), d, 1))
    box = {}
    if e.record:
        def tokens_backward():
            gx = box["gx"][0]
            assert y0.gdt == L.F32, "the transformer fusion keeps fp32 gradients"
            gwv, gbv = e.grad_tensor("conv_in_1x1.weight", cv.weight), e.grad_tensor("conv_in_1x1.bias", cv.bias)
            gwa, gba = e.grad_tensor("audio_conv_1x1.weight", ca.weight), e.grad_tensor("audio_conv_1x1.bias", ca.bias)
            for b in range(B):       # dWv[c, k] = sum over (clip, position): the two operands order that pair differently -> one launch per clip
                gemm(e, Cin, d, Pn, y0.ptr() + b * Pn * y0.ld * y0.buf.element_size(), (1, y0.ld), _ptr(gx, b * d), (1, B * d), _ptr(gwv), (1, Cin),
                     a_dtype=e.dt, a_xf=xf, a_xf_on_m=1, acc=int(b > 0))
            colsum(e, gx, Pn * B, d, gbv)
            for j in range(na):      # dWa[c, ch] = sum over clips of the j-th audio token's gradient x the j-th SoundNet column
                gemm(e, d, Cin, B, _ptr(gx, (Pn + j) * B * d), (1, d), _ptr(a, j), (na, Cin * na), _ptr(gwa), (Cin, 1), acc=int(j > 0))
            gemm(e, d, 1, na * B, _ptr(gx, Pn * B * d), (1, d), _ptr(ones(e)), (0, 0), _ptr(gba), (1, 0))
            first = e.first_write(y0)
            gemm(e, Pn, Cin, d, _ptr(gx), (B * d, 1, d), _ptr(cv.weight), (1, Cin), y0.gptr(), (y0.ldg, 1, Pn * y0.ldg), nb=(B, 1), acc=0 if first else 1)
            gemm(e, na, Cin, d, _ptr(gx, Pn * B * d), (B * d, 1, d), _ptr(ca.weight), (1, Cin), _ptr(ga), (1, na, Cin * na), nb=(B, 1))
            e.param_grads["conv_in_1x1.weight"], e.param_grads["conv_in_1x1.bias"] = gwv, gbv
            e.param_grads["audio_conv_1x1.weight"], e.param_grads["audio_conv_1x1.bias"] = gwa, gba
        e.tape.append(_backward_pass(tokens_backward))
    y, gy, box["gx"] = encoder_plan(e, "transformer.transformer_encoder.layers.", tr, x0, S, B)
    out = e.new_act("fu.out", B, y0.T, y0.H, y0.W, 2 * d, gdtype=f32)
    esz = out.buf.element_size()
    gemm(e, Pn, d, 0, None, (0, 0), None, (0, 0), out.ptr(), (out.ld, 1, Pn * out.ld), nb=(B, 1), c_dtype=e.dt, bias2=(_ptr(y), B * d, 1, d))
    gemm(e, Pn, d, na, _ptr(ones(e)), (0, 0), _ptr(y, Pn * B * d), (1, B * d, d), out.ptr() + d * esz, (out.ld, 1, Pn * out.ld), nb=(B, 1),
         c_dtype=e.dt, alpha=1.0 / na)
    if e.record:
        def split_backward():
            assert out.gdt == L.F32
            g, ldg = out.gptr(), out.ldg
            gemm(e, Pn, d, 0, None, (0, 0), None, (0, 0), _ptr(gy), (B * d, 1, d), nb=(B, 1), bias2=(g, ldg, 1, Pn * ldg))
            gemm(e, na, d, Pn, _ptr(ones(e)), (0, 0), g + 4 * d, (1, ldg, Pn * ldg), _ptr(gy, Pn * B * d), (B * d, 1, d), nb=(B, 1), alpha=1.0 / na)
        e.tape.append(_backward_pass(split_backward))
    return out
