"""Drop-in ``VideoSaliencyModel`` / ``VideoAudioSaliencyModel`` (reference model.py:72-112, :191-249).

Same constructor keywords, ``forward()`` signatures, sub-module attribute tree and ``state_dict`` layout
as the reference, so ``train.py`` / ``generate_result*.py`` can import these classes instead of the
reference's.  The sub-modules are *parameter holders only* (they have no ``forward``): the computation
is one fused plan executed by ``vinet_b200.engine.Engine`` through the C-ABI kernels, wrapped in a single
``torch.autograd.Function`` so ``loss.backward()`` / optimizers / DDP see ordinary ``.grad`` tensors.

Reference sites: BasicConv3d/SepConv3d/Mixed_* model_utils.py:128-420; BackBoneS3D model.py:690-743;
DecoderConvUp{,8,16,48} model.py:251-499; SoundNet model.py:746-825.
"""
import math
import os

import torch
import torch.nn as nn

from . import arch
from . import lib as L
from .engine import Act, ConvGeom, Engine, WinAct

BN_EPS, BN_MOM = 1e-3, 1e-3   # model_utils.py:132


# ----------------------------------------------------------------------------- parameter holders
class ConvParams(nn.Module):
    """Holds a Conv3d/Conv2d weight (+bias) with PyTorch's default initialisation; never computes."""

    def __init__(self, cin, cout, k, bias=False):
        super().__init__()
        self.weight = nn.Parameter(torch.empty(cout, cin, *k))
        self.bias = nn.Parameter(torch.empty(cout)) if bias else None
        nn.init.kaiming_uniform_(self.weight, a=math.sqrt(5))
        if bias:
            bound = 1.0 / math.sqrt(cin * math.prod(k))
            nn.init.uniform_(self.bias, -bound, bound)


class BNParams(nn.Module):
    def __init__(self, c, eps=BN_EPS, momentum=BN_MOM):
        super().__init__()
        self.eps, self.momentum = eps, momentum
        self.weight = nn.Parameter(torch.ones(c))
        self.bias = nn.Parameter(torch.zeros(c))
        self.register_buffer("running_mean", torch.zeros(c))
        self.register_buffer("running_var", torch.ones(c))
        self.register_buffer("num_batches_tracked", torch.tensor(0, dtype=torch.long))


class Marker(nn.Module):
    """Parameter-less placeholder keeping nn.Sequential indices equal to the reference's
    (MaxPool3d / ReLU / Upsample / Sigmoid entries)."""

    def __init__(self, what):
        super().__init__()
        self.what = what

    def extra_repr(self):
        return self.what


class BasicConv3d(nn.Module):
    def __init__(self, cin, cout):
        super().__init__()
        self.conv = ConvParams(cin, cout, (1, 1, 1))
        self.bn = BNParams(cout)


class SepConv3d(nn.Module):
    def __init__(self, cin, cout, k, stride, padding):
        super().__init__()
        self.k, self.stride, self.padding = k, stride, padding
        self.conv_s = ConvParams(cin, cout, (1, k, k))
        self.bn_s = BNParams(cout)
        self.conv_t = ConvParams(cout, cout, (k, 1, 1))
        self.bn_t = BNParams(cout)


class Mixed(nn.Module):
    def __init__(self, name):
        super().__init__()
        self.name = name
        cin, b0, b1r, b1, b2r, b2, b3 = arch.MIXED[name]
        self.branch0 = nn.Sequential(BasicConv3d(cin, b0))
        self.branch1 = nn.Sequential(BasicConv3d(cin, b1r), SepConv3d(b1r, b1, 3, 1, 1))
        self.branch2 = nn.Sequential(BasicConv3d(cin, b2r), SepConv3d(b2r, b2, 3, 1, 1))
        self.branch3 = nn.Sequential(Marker("MaxPool3d(3,1,1)"), BasicConv3d(cin, b3))


class BackBoneS3D(nn.Module):
    def __init__(self):
        super().__init__()
        self.base1 = nn.Sequential(SepConv3d(3, 64, 7, 2, 3), Marker("MaxPool3d((1,3,3),(1,2,2),(0,1,1))"),
                                   BasicConv3d(64, 64), SepConv3d(64, 192, 3, 1, 1))
        self.maxp2 = Marker("MaxPool3d((1,3,3),(1,2,2),(0,1,1))")
        self.base2 = nn.Sequential(*[Mixed(n) for n in arch.STAGES["base2"]])
        self.maxp3 = Marker("MaxPool3d(3,2,1)")
        self.base3 = nn.Sequential(*[Mixed(n) for n in arch.STAGES["base3"]])
        self.maxt4 = Marker("MaxPool3d((2,1,1),(2,1,1))")
        self.maxp4 = Marker("MaxPool3d((1,2,2),(1,2,2))")
        self.base4 = nn.Sequential(*[Mixed(n) for n in arch.STAGES["base4"]])


class DecoderConvUp(nn.Module):
    """Parameter layout of DecoderConvUp / DecoderConvUp8 / 16 / 48 (model.py:251-499) and of the ablation decoders
    DecoderConvUpNoHier / 1Hier / 2Hier (model.py:501-688; num_hier < 3, 32-frame clips only like the reference)."""

    def __init__(self, num_clips=32, num_hier=3):
        super().__init__()
        if num_hier != 3:
            num_clips = 32       # model.py:84-90: the ablation decoders ignore num_clips
        self.num_clips, self.num_hier = num_clips, num_hier
        self.upsampling = Marker("Upsample((1,2,2),trilinear)")
        heads = [[ConvParams(cin, cout, (kt, 3, 3)), Marker("ReLU"), self.upsampling]
                 for cin, cout, kt in arch.decoder_head(num_hier)]
        tail = []
        for item in arch.decoder_tail(num_clips):
            if isinstance(item, str):
                tail.append(self.upsampling if item == "up" else Marker(item))
            else:
                _, cin, cout, k, _, _, bias = item
                tail.append(ConvParams(cin, cout, k, bias=bias))
        self.convtsp1 = nn.Sequential(*heads[0])
        self.convtsp2 = nn.Sequential(*heads[1])
        self.convtsp3 = nn.Sequential(*heads[2])
        self.convtsp4 = nn.Sequential(*(heads[3] + tail))


# ----------------------------------------------------------------------------- the fused plan
def _sepconv_s(e, pfx, srcs, m, cin_real=None):
    """First half of SepConv3d (model_utils.py:144-146): (1,k,k) conv + BN + ReLU -> the intermediate activation."""
    a0 = srcs[0]
    k, s, p = m.k, m.stride, m.padding
    gs = ConvGeom((1, k, k), (1, s, s), (0, p, p))
    To, Ho, Wo = gs.out_dims(a0.T, a0.H, a0.W)
    cout = m.conv_s.weight.shape[0]
    mid = e.new_act(pfx + ".s", a0.B, To, Ho, Wo, cout)
    e.conv_bn(pfx + ".conv_s", pfx + ".bn_s", srcs, m.conv_s.weight, m.bn_s, gs, mid, cin_real=cin_real)
    return mid


def _sepconv_t(e, pfx, mid, m, out=None, defer_bn=False):
    """Second half of SepConv3d (model_utils.py:148-150): (k,1,1) conv + BN + ReLU.
    defer_bn: the only consumer is a max-pool, which applies scale / shift / ReLU on read - the engine may skip the apply pass."""
    k, s, p = m.k, m.stride, m.padding
    gt = ConvGeom((k, 1, 1), (s, 1, 1), (p, 0, 0))
    To2, _, _ = gt.out_dims(mid.T, mid.H, mid.W)
    if out is None and not defer_bn:
        out = e.new_act(pfx + ".t", mid.B, To2, mid.H, mid.W, m.conv_t.weight.shape[0])
    return e.conv_bn(pfx + ".conv_t", pfx + ".bn_t", [mid], m.conv_t.weight, m.bn_t, gt, out, defer_to=pfx + ".t")


def _sepconv(e, pfx, srcs, m, out=None, cin_real=None, defer_bn=False):
    return _sepconv_t(e, pfx, _sepconv_s(e, pfx, srcs, m, cin_real=cin_real), m, out, defer_bn=defer_bn)


_G1 = ConvGeom((1, 1, 1), (1, 1, 1), (0, 0, 0))


def _basic(e, pfx, x, m, out=None):
    if out is None:
        out = e.new_act(pfx + ".o", x.B, x.T, x.H, x.W, m.conv.weight.shape[0])
    e.conv_bn(pfx + ".conv", pfx + ".bn", [x], m.conv.weight, m.bn, _G1, out)
    return out


def _mixed(e, pfx, x, m, gdtype=None):
    """Mixed_3b..5c (model_utils.py:162-420).  Launch order differs from the reference's textual order (pool
    branch first, the three 1x1 convs on x fused into one GEMM); the arithmetic per branch is unchanged."""
    cin, b0, b1r, b1, b2r, b2, b3 = arch.MIXED[m.name]
    out = e.new_act(pfx + ".cat", x.B, x.T, x.H, x.W, b0 + b1 + b2 + b3, gdtype=gdtype)
    # The BatchNorm layers that are ready together run as ONE multi-layer launch per pass (Engine.bn_begin / bn_flush):
    # {branch0, branch1.0, branch2.0, branch3.1}, then both conv_s, then both conv_t - 3 launch groups instead of 7 layers.
    e.bn_begin()
    # Independent launches of the block alternate between the main and a side stream (Engine.fork / on_side / join): the pool
    # branch beside the fused 1x1 GEMM, branch1 beside branch2 - small layers that each leave most of the 148 SMs idle.
    e.fork()
    e.on_side(True)
    # pool branch first: its backward (a scatter-add) then runs after the convs' data gradients have written x.grad
    p = e.maxpool(pfx + ".branch3.pool", x, (3, 3, 3), (1, 1, 1), (1, 1, 1))
    _basic(e, pfx + ".branch3.1", p, m.branch3[1], out.slice(b0 + b1 + b2, b3))
    e.on_side(False)
    t1 = e.new_act(pfx + ".branch1.0.o", x.B, x.T, x.H, x.W, b1r)
    t2 = e.new_act(pfx + ".branch2.0.o", x.B, x.T, x.H, x.W, b2r)
    e.conv_bn_group(pfx + ".fused1x1", x, [
        (pfx + ".branch0.0.conv", pfx + ".branch0.0.bn", m.branch0[0].conv.weight, m.branch0[0].bn, out.slice(0, b0)),
        (pfx + ".branch1.0.conv", pfx + ".branch1.0.bn", m.branch1[0].conv.weight, m.branch1[0].bn, t1),
        (pfx + ".branch2.0.conv", pfx + ".branch2.0.bn", m.branch2[0].conv.weight, m.branch2[0].bn, t2)])
    e.join()
    e.bn_flush()
    e.fork()
    m1 = _sepconv_s(e, pfx + ".branch1.1", [t1], m.branch1[1])
    e.on_side(True)
    m2 = _sepconv_s(e, pfx + ".branch2.1", [t2], m.branch2[1])
    e.join()
    e.bn_flush()
    e.fork()
    _sepconv_t(e, pfx + ".branch1.1", m1, m.branch1[1], out.slice(b0, b1))
    e.on_side(True)
    _sepconv_t(e, pfx + ".branch2.1", m2, m.branch2[1], out.slice(b0 + b1, b2))
    e.join()
    e.bn_end()
    return out


def backbone_plan(e, pfx, bb, x, y0_gdtype=None, windows=None):
    """BackBoneS3D.forward (model.py:720-743): returns [y0, y1, y2, y3] as Acts.
    windows = (b, L): inference over overlapping sliding windows (generate_result.py:55-73).  `x` then holds b + L - 1 single
    frames; the per-frame (1,7,7) stem convolution runs ONCE per frame and the temporal stem convolution reads window i as frames
    i .. i+L-1 of its output (batch pitch = one frame) - consecutive windows share L-1 of their L stem frames."""
    if windows is None:
        a = _sepconv(e, pfx + "base1.0", [x], bb.base1[0], cin_real=3, defer_bn=True)   # consumed by the max-pool below only
    else:
        b, Lc = windows
        mid = _sepconv_s(e, pfx + "base1.0", [x], bb.base1[0], cin_real=3)
        assert mid.T == 1 and mid.B == b + Lc - 1 and mid.choff == 0 and not e.record and mid.xform == L.XF_IDENT
        win = Act(mid.buf, b, Lc, mid.H, mid.W, mid.C)
        win.ldb = mid.H * mid.W * mid.ld
        a = _sepconv_t(e, pfx + "base1.0", win, bb.base1[0])
    a = e.maxpool(pfx + "base1.1", a, (1, 3, 3), (1, 2, 2), (0, 1, 1))
    a = _basic(e, pfx + "base1.2", a, bb.base1[2])
    y3 = _sepconv(e, pfx + "base1.3", [a], bb.base1[3])
    a = e.maxpool(pfx + "maxp2", y3, (1, 3, 3), (1, 2, 2), (0, 1, 1))
    for i, m in enumerate(bb.base2):
        a = _mixed(e, pfx + "base2.%d" % i, a, m)
    y2 = a
    a = e.maxpool(pfx + "maxp3", y2, (3, 3, 3), (2, 2, 2), (1, 1, 1))
    for i, m in enumerate(bb.base3):
        a = _mixed(e, pfx + "base3.%d" % i, a, m)
    y1 = a
    # maxt4 (2,1,1) followed by maxp4 (1,2,2) == one (2,2,2)/2 max pool
    a = e.maxpool(pfx + "maxt4p4", y1, (2, 2, 2), (2, 2, 2), (0, 0, 0))
    for i, m in enumerate(bb.base4):
        a = _mixed(e, pfx + "base4.%d" % i, a, m, gdtype=y0_gdtype if i == len(bb.base4) - 1 else None)
    return [a, y1, y2, y3]


def decoder_plan(e, pfx, dec, y0, y1, y2, y3):
    """DecoderConvUp*.forward (model.py:286-311; ablations :543-562, :606-625, :669-688): returns the (B,H,W) fp32 saliency
    map tensor.  Stage i concatenates the skip tensor y_i along time only when i <= num_hier."""
    heads = [dec.convtsp1[0], dec.convtsp2[0], dec.convtsp3[0], dec.convtsp4[0]]
    names = ["convtsp1.0", "convtsp2.0", "convtsp3.0", "convtsp4.0"]
    skips = [None] + [y if i < dec.num_hier else None for i, y in enumerate((y1, y2, y3))]
    z = y0
    for m, nm, skip in zip(heads, names, skips):
        kt = m.weight.shape[2]
        srcs = [z] if skip is None else [z, skip]
        z = e.conv_relu_up(pfx + nm, srcs, m.weight, ConvGeom((kt, 3, 3), (kt, 1, 1), (0, 1, 1)))
    tail = arch.decoder_tail(dec.num_clips)
    convs = [(i + 3, it) for i, it in enumerate(tail) if not isinstance(it, str)]
    # Tail: conv A (kt,3,3) -> relu -> up [-> conv B (kt,1,1) (+bias) -> relu] -> 1x1x1 head -> sigmoid.  The last 2x up-sampling
    # writes the largest tensor of the decoder and feeds only pointwise-in-space work, so it never runs as such:
    #   * conv B is linear and pointwise in space, the interpolation weights sum to one: B(up(x)) + b == up(B(x) + b).  B therefore
    #     runs on the LOW-RES grid (a quarter of the rows) and the head interpolates its output (then relu, dot, sigmoid);
    #   * without conv B (T = 8 / 16) the head interpolates relu(A's output) directly.
    idx, (_, cin, cout, k, s, p, bias) = convs[0]
    geomA = ConvGeom(k, s, (0, p, p))
    wA = dec.convtsp4[idx].weight
    if len(convs) == 3:
        a = e.conv_relu_act(pfx + "convtsp4.%d" % idx, [z], wA, geomA)
        idx, (_, cin, cout, k, s, p, bias) = convs[1]
        m = dec.convtsp4[idx]
        z, conv_bwd = e.conv_relu(pfx + "convtsp4.%d" % idx, [a], m.weight, ConvGeom(k, s, (0, p, p)), bias=m.bias)
        order = "post"
    else:
        z, conv_bwd = e.conv_relu(pfx + "convtsp4.%d" % idx, [z], wA, geomA)
        order = "pre"
    idx, _ = convs[-1]
    m = dec.convtsp4[idx]
    return e.head(pfx + "convtsp4.%d" % idx, z, m.weight, m.bias, conv_bwd, up2=order)


class _PlanFunction(torch.autograd.Function):
    """One autograd node for the whole model: forward runs the plan, backward runs the tape."""

    @staticmethod
    def forward(ctx, model, record, names, x, *extra_and_params):
        e = model._engine_for(x.device)
        out = model._run_plan(e, record, x, *extra_and_params[:model._n_extra])
        ctx.engine, ctx.names, ctx.gen, ctx.n_extra = e, names, e.generation, model._n_extra
        return out

    @staticmethod
    def backward(ctx, gout):
        e = ctx.engine
        if ctx.gen != e.generation:
            raise RuntimeError("vinet_b200: backward() of a stale forward (one outstanding forward per model)")
        if ctx.gen == e.consumed_gen:
            raise RuntimeError("vinet_b200: second backward() through the same forward: the activation tape is consumed by the "
                               "first one (retain_graph is not supported; run the forward again)")
        grads = e.backward(gout.contiguous().float())
        return (None, None, None, None) + (None,) * ctx.n_extra + tuple(grads.pop(n, None) for n in ctx.names)


class _PlanModule(nn.Module):
    """Shared machinery: engine cache (per device), parameter list, autograd wiring."""

    precision = "bf16"
    _n_extra = 0

    def set_precision(self, precision):
        """"bf16": throughput mode (bf16 storage, tcgen05).  "bf16x3": tensor-core parity mode (fp32 storage, every conv = three
        tcgen05 launches over bf16-split operands; "bf16x6" adds a third term).  "fp32": FFMA reference parity mode."""
        assert precision in Engine.PRECISIONS
        self.precision = precision
        self.__dict__.pop("_engines", None)
        return self

    def _engine_for(self, device):
        engines = self.__dict__.setdefault("_engines", {})
        key = (str(device), self.precision)
        if key not in engines:
            engines[key] = Engine(self.precision, backend=self.__dict__.get("_backend"))
            engines[key].generation = 0
        return engines[key]

    # ---- flat gradient arena (multi-GPU: ONE all-reduce on one buffer instead of DDP's per-tensor bucket copies) ----
    def enable_grad_arena(self, on=True):
        """Write every parameter gradient into one flat fp32 buffer; ``param.grad`` tensors become views of it."""
        self.__dict__["_use_arena"] = bool(on)
        self.__dict__.pop("_arenas", None)
        return self

    def _arena_for(self, device, named):
        arenas = self.__dict__.setdefault("_arenas", {})
        key = (str(device), tuple(n for n, _ in named))
        if key not in arenas:
            layout, off = {}, 0
            for n, p in named:
                layout[n] = (off, p.numel())
                off += (p.numel() + 3) // 4 * 4          # 16-byte aligned slices
            arenas[key] = (torch.zeros(off, dtype=torch.float32, device=device), layout, dict(named))
        return arenas[key]

    def sync_gradients(self, group=None):
        """Average the gradients of the last backward over the process group: one NCCL all-reduce of the flat arena
        (124.4 MB for ViNet).  BatchNorm statistics stay per replica (the reference's nn.DataParallel semantics).
        With ``overlap_gradient_sync()`` the decoder's slice (3/4 of the bytes, complete after the first fifth of the backward pass)
        has already been reduced on a side stream while the backbone's backward ran; only the rest is reduced here."""
        import torch.distributed as dist
        arenas = self.__dict__.get("_arenas")
        assert arenas, "sync_gradients() needs enable_grad_arena() and a backward pass"
        world = dist.get_world_size(group)
        early = self.__dict__.pop("_early_sync", None)
        for flat, *_ in arenas.values():
            parts = [flat]
            if early is not None and early[0] is flat:
                lo, hi = early[1], early[2]
                parts = [t for t in (flat[:lo], flat[hi:]) if t.numel()]
            for t in parts:
                if dist.get_backend(group) == "nccl":
                    dist.all_reduce(t, op=dist.ReduceOp.AVG, group=group)
                else:
                    dist.all_reduce(t, group=group)
                    t.div_(world)
        if early is not None:
            torch.cuda.current_stream(early[0].device).wait_event(early[3])

    def overlap_gradient_sync(self, on=True, group=None):
        """Start the all-reduce of the decoder's gradients as soon as the decoder's backward has finished, on a side stream, so
        that it overlaps the backbone's backward (train.py's loop has no such hook: this is an opt-in of the flat-arena path,
        used by bench.py at N > 1; inside a captured step the fork / join become graph edges).  Needs ``enable_grad_arena()``,
        ``zero_grad(set_to_none=True)`` discipline and ``sync_gradients()`` after every backward."""
        self.__dict__["_overlap_sync"] = (group,) if on else None
        return self

    def _decoder_backward_done(self, e):
        """Engine call-back (tape marker in front of the decoder plan): every decoder gradient sits in its arena slice."""
        cfg = self.__dict__.get("_overlap_sync")
        if cfg is None or e.arena is None or e.device.type != "cuda" or e.arena_bypassed:
            return
        import torch.distributed as dist
        if not (dist.is_available() and dist.is_initialized()) or dist.get_backend(cfg[0]) != "nccl":
            return
        flat, layout = e.arena[0], e.arena[1]
        spans = [(off, off + (n + 3) // 4 * 4) for name, (off, n) in layout.items() if ".decoder." in "." + name]
        if not spans:
            return
        lo, hi = min(a for a, _ in spans), min(max(b for _, b in spans), flat.numel())
        if any(not (lo <= off < hi) == (".decoder." in "." + name) for name, (off, n) in layout.items()):
            return                                                   # the decoder's parameters are not one contiguous run
        main = torch.cuda.current_stream(e.device)
        side = self.__dict__.get("_sync_stream")
        if side is None or side.device != e.device:
            side = self.__dict__["_sync_stream"] = torch.cuda.Stream(device=e.device)
        side.wait_stream(main)
        with torch.cuda.stream(side):
            dist.all_reduce(flat[lo:hi], op=dist.ReduceOp.AVG, group=cfg[0])
            ev = torch.cuda.Event()
            ev.record(side)
        self.__dict__["_early_sync"] = (flat, lo, hi, ev)

    def broadcast_parameters(self, src=0, group=None):
        """Rank `src`'s parameters and buffers to every rank (what DistributedDataParallel does at construction)."""
        import torch.distributed as dist
        for t in list(self.parameters()) + list(self.buffers()):
            dist.broadcast(t.detach(), src, group=group)      # detach() shares the version counter (t.data would not bump it)
        return self.invalidate_weight_cache()

    def invalidate_weight_cache(self):
        """Parameters were written behind autograd's back (``p.data.copy_``, EMA updates, a custom in-place optimizer): make
        every engine re-pack its cached tensor-core weight copies on the next forward."""
        for e in self.__dict__.get("_engines", {}).values():
            e.weights_dirty = True
        return self

    def _replica_parameters(self):
        """nn.DataParallel replica (train.py:182-184): ``_parameters`` is empty, the broadcast copies live in
        ``_former_parameters`` of every sub-module and are plain attributes with a grad_fn."""
        named = []
        for prefix, mod in self.named_modules():
            for k, t in getattr(mod, "_former_parameters", {}).items():
                if t is not None:
                    named.append(((prefix + "." if prefix else "") + k, t))
        return named

    def _call_plan(self, x, *extra):
        if x.device.type != "cuda" and self.__dict__.get("_backend") is None:
            raise RuntimeError("vinet_b200 has no CPU path: move the model and inputs to a CUDA device")
        replica = bool(getattr(self, "_is_replica", False))
        params = self._replica_parameters() if replica else self.named_parameters()
        named = [(n, p) for n, p in params if p.requires_grad and self._plan_uses(n)]
        names = tuple(n for n, _ in named)
        record = torch.is_grad_enabled() and len(named) > 0
        e = self._engine_for(x.device)
        e.replica = replica
        e.arena = self._arena_for(x.device, named) if (record and self.__dict__.get("_use_arena") and not replica) else None
        e.backward_point_cb = (lambda name, e=e: self._decoder_backward_done(e)) if (e.arena is not None and self.__dict__.get("_overlap_sync")) else None
        return _PlanFunction.apply(self, record, names, x, *extra, *[p for _, p in named])

    def _plan_uses(self, name):
        return True


class VideoSaliencyModel(_PlanModule):
    """ViNet (model.py:72-112): ``num_hier`` 3 (default; clip lengths 8/16/32/48) and the ablation decoders 0/1/2
    (model.py:501-688).  ``use_upsample=False`` raises like the reference does (its DecoderConvT does not exist)."""

    def __init__(self, transformer_in_channel=32, nhead=4, use_upsample=True, num_hier=3, num_clips=32):
        super().__init__()
        if not use_upsample:
            raise NameError("name 'DecoderConvT' is not defined")      # model.py:101 — same failure as the reference
        if num_hier not in (0, 1, 2, 3):
            raise AttributeError("'VideoSaliencyModel' object has no attribute 'decoder'")   # model.py:84-99 builds none either
        if num_hier == 3 and num_clips not in (8, 16, 32, 48):
            raise AttributeError("'VideoSaliencyModel' object has no attribute 'decoder'")
        self.backbone = BackBoneS3D()
        self.num_hier = num_hier
        self.decoder = DecoderConvUp(num_clips, num_hier)

    def forward(self, x):
        return self._call_plan(x)

    def forward_windows(self, frames, windows):
        """Inference over `windows` overlapping clips of a frame sequence: frames is (windows + L - 1, 3, H, W),
        window i = frames i .. i+L-1; returns (windows, H, W) like `forward` on the stacked clips, with the per-frame stem
        convolution computed once per frame.  bf16 tensor-core engine, eval mode, no autograd."""
        assert not self.training and not torch.is_grad_enabled() and self.precision == "bf16", "forward_windows: eval / no_grad / bf16"
        Lc = self.decoder.num_clips
        assert frames.dim() == 4 and frames.shape[0] == windows + Lc - 1 and frames.shape[1] == 3
        self.__dict__["_windows"] = (windows, Lc)
        try:
            return self._call_plan(frames.unsqueeze(2))
        finally:
            self.__dict__.pop("_windows", None)

    def _run_plan(self, e, record, x, prefix=""):
        e.generation += 1
        e.begin(x.device, self.training, record)
        xin = pack_input(e, x)
        ys = backbone_plan(e, prefix + "backbone.", self.backbone, xin, windows=self.__dict__.get("_windows"))
        e.mark_backward_point("decoder")     # reached in the backward pass once every decoder gradient has been written
        out = decoder_plan(e, prefix + "decoder.", self.decoder, *ys)
        e.end_forward()
        return out


def pack_input(e, x):
    """(B,3,T,H,W) fp32, any strides (train.py:205 passes a permuted view) -> NDHWC, C padded 3->8."""
    assert x.dim() == 5 and x.shape[1] == 3 and x.dtype == torch.float32, "expected a (B,3,T,H,W) fp32 clip"
    B, _, T, H, W = x.shape
    assert H % 32 == 0 and W % 32 == 0, "H and W must be multiples of 32 (the reference fails otherwise)"
    win = e.eng == L.ENGINE_TC and e.use_tma      # TMA engine: rows padded with zero columns (3 left, >= 5 right)
    # bf16 tensor-core mode: FOUR channels per pixel - the forward stem convolution reads the clip in place through un-swizzled
    # UMMA descriptors and its weight gradient as 16-pixel windows (VINET_KLAYOUT_WIN4, csrc/conv_stream.cu); the parity modes
    # keep the 8-channel fp32 clip they split into bf16 planes
    win4 = (win and not e.split and e.dt == L.BF16 and "vinet_conv_win4_fused" in e.lib.fn and not os.environ.get("VINET_NO_WIN4"))
    d = L.PackInput()
    d.x = x.data_ptr()
    d.sb, d.sc, d.st, d.sh, d.sw = x.stride()
    d.B, d.C, d.T, d.H, d.W, d.cpad, d.out_dtype = B, 3, T, H, W, 8, e.dt
    if win4:
        wl, Wp = 3, W + 16                        # 16-pixel windows of the last output column stay inside the row
        buf = e.buf("input.packed4", (B, T, H, Wp, 4), torch.bfloat16)
        d.out, d.out4, d.wl, d.Wp = None, buf.data_ptr(), wl, Wp
        e.call("vinet_pack_input", d)
        return WinAct(buf, B, T, H, W, wl, Wp, cpp=4)
    wl, Wp = (3, W + 8) if win else (0, W)
    buf = e.buf("input.packed", (B, T, H, Wp, 8), e.tdtype)
    d.out = buf.data_ptr()
    d.wl, d.Wp = (wl, Wp) if win else (0, 0)
    e.call("vinet_pack_input", d)
    return WinAct(buf, B, T, H, W, wl, Wp) if win else Act(buf, B, T, H, W, 8)
