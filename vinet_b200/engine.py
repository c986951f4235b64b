"""Host-side executor of the hot path: turns layers into C-ABI kernel calls (no torch compute ops).

PyTorch is used for device memory (the caching allocator), streams and parameters only; every
arithmetic step is a kernel of ``libvinet_b200.so`` reached through ``vinet_b200.lib``.

Data model
  * ``Act``  — a channel view of an NDHWC buffer together with the *pending transform* of its producer
    (BatchNorm scale/shift and/or ReLU, applied by whoever reads it) and, when training, an fp32 buffer
    holding the gradient w.r.t. the activated value.
  * forward ops append closures to ``Engine.tape``; ``Engine.backward()`` runs them in reverse.  This is
    a purpose-built replacement for autograd on this path: conv dgrad/wgrad, BN, pool,
    upsample, head and loss backward are all explicit kernels.

Precision modes
  * ``bf16``   : bf16 storage, tcgen05 tensor-core GEMMs (fp32 accumulation in TMEM)   — throughput mode
  * ``bf16x3`` : fp32 storage; every conv operand is split into bf16 terms (x = x0 + x1) and the convolution is the sum of
                 the three tcgen05 launches x0*w0 + x0*w1 + x1*w0 accumulated in fp32     — tensor-core PARITY mode: the very
                 kernels of the throughput mode gated at the north-star tolerances (``bf16x6``: three terms, six launches)
  * ``fp32``   : fp32 storage, fp32 FFMA GEMMs                                          — reference parity mode
"""
import ctypes as C
import os

import torch

from . import lib as L


_POISON = bool(os.environ.get("VINET_POISON"))

# Packed (bf16, swizzled) copies of the conv weights are cached per parameter.  Tensor._version alone cannot tell when a
# parameter changed: fused optimizers (torch.optim.Adam(fused=True), the fast path every trainer uses) update parameters
# WITHOUT bumping it.  Every optimizer step therefore also advances this epoch (global post-step hook), and a cache entry
# is valid only for the (version, epoch) it was packed at.
_WEIGHT_EPOCH = [0]


def _on_optimizer_step(*_args, **_kwargs):
    _WEIGHT_EPOCH[0] += 1


try:
    from torch.optim.optimizer import register_optimizer_step_post_hook as _reg_hook
    _reg_hook(_on_optimizer_step)
except Exception:      # pragma: no cover - very old torch: fall back to re-packing on every recorded forward
    _WEIGHT_EPOCH = None


def _wstamp(w):
    # the storage address is part of the stamp: `param.data = new` / `.to()` keep the Parameter object but move its memory
    return (w._version, _WEIGHT_EPOCH[0], w.data_ptr()) if _WEIGHT_EPOCH is not None else None


def cdiv(a, b):
    return -(-a // b)


def round_up(a, b):
    return cdiv(a, b) * b


class Act:
    def __init__(self, buf, B, T, H, W, C_, choff=0, xform=L.XF_IDENT, scale=None, shift=None):
        self.buf, self.B, self.T, self.H, self.W, self.C, self.choff = buf, B, T, H, W, C_, choff
        self.ld = buf.shape[-1]
        self.xform, self.scale, self.shift = xform, scale, shift
        self.ldb = 0          # elements between clips; 0 = dense.  One frame: overlapping sliding windows (inference driver)
        self.up2 = False      # VIRTUAL 2x up-sampled activation: buf holds [B,T,H/2,W/2,ld], xform has XF_UP2, H and W are the
        self.mat = None       # hi-res extents consumers see; mat = its materialised copy, made only for consumers that cannot fuse
        self.grad = None      # [B,T,H,W,ldg] in the engine's storage type (bf16 / fp32), or fp32 when forced
        self.gchoff = 0
        self.gdt = L.F32
        self.needs_grad = False

    @property
    def rows(self):
        return self.B * self.T * self.H * self.W

    def ptr(self):
        return self.buf.data_ptr() + self.choff * self.buf.element_size()

    def gptr(self):
        return self.grad.data_ptr() + self.gchoff * self.grad.element_size()

    @property
    def ldg(self):
        return self.grad.shape[-1]

    def slice(self, c0, c):
        a = Act(self.buf, self.B, self.T, self.H, self.W, c, self.choff + c0, self.xform,
                None if self.scale is None else self.scale[c0:c0 + c],
                None if self.shift is None else self.shift[c0:c0 + c])
        a.grad, a.gchoff, a.needs_grad, a.gdt = self.grad, self.gchoff + c0, self.needs_grad, self.gdt
        a.up2 = self.up2
        return a


class WinAct(Act):
    """The packed input clip stored with explicit zero columns ([B,T,H,Wp,8], image at columns wl..wl+W) so that
    the TMA-fed stem conv can address every row as overlapping sliding windows (VINET_KLAYOUT_WIN8)."""

    def __init__(self, buf, B, T, H, W, wl, Wp, cpp=8):
        super().__init__(buf, B, T, H, W, 8)
        self.wl, self.Wp = wl, Wp
        self.cpp = cpp        # channels per stored pixel: 8 (VINET_KLAYOUT_WIN8) or 4 (bf16 tensor-core mode, VINET_KLAYOUT_WIN4)


def _ptr(t):
    return None if t is None else t.data_ptr()


def _taps(kt, kh, kw):
    return [(a, b, c) for a in range(kt) for b in range(kh) for c in range(kw)]


def _fill_taps(field, taps):
    for i, (a, b, c) in enumerate(taps):
        field[i][0], field[i][1], field[i][2], field[i][3] = a, b, c, 0


class ConvGeom:
    def __init__(self, k, s, p):
        (self.kt, self.kh, self.kw), (self.st, self.sh, self.sw), (self.pt, self.ph, self.pw) = k, s, p

    def out_dims(self, T, H, W):
        return ((T + 2 * self.pt - self.kt) // self.st + 1, (H + 2 * self.ph - self.kh) // self.sh + 1,
                (W + 2 * self.pw - self.kw) // self.sw + 1)


_G1 = ConvGeom((1, 1, 1), (1, 1, 1), (0, 0, 0))


class Engine:
    PRECISIONS = ("bf16", "fp32", "bf16_simt", "bf16x3", "bf16x6")

    def __init__(self, precision="bf16", backend=None):
        # "bf16_simt": bf16 storage with the fp32 FFMA engine — the on-device cross-check of the tcgen05 path
        assert precision in self.PRECISIONS
        self.precision = precision
        # split-precision parity mode: activations / gradients stay fp32, conv operands are bf16 expansions (vinet_split_bf16)
        self.split = {"bf16x3": 2, "bf16x6": 3}.get(precision, 0)
        self.eng = L.ENGINE_TC if (precision == "bf16" or self.split) else L.ENGINE_SIMT
        self.dt = L.F32 if (precision == "fp32" or self.split) else L.BF16
        self.tdtype = torch.float32 if self.dt == L.F32 else torch.bfloat16
        self.opdt = L.BF16 if self.split else self.dt      # storage type of the conv operands (sources and dY)
        # (operand part, weight part) of every launch of one convolution; parts i + j < nparts
        self.terms = [(i, j) for n in range(self.split) for i in range(n + 1) for j in [n - i]] if self.split else [(0, 0)]
        self.lib = backend if backend is not None else L.get()   # tests may inject the numpy kernel spec
        self.pool = {}
        self.wcache = {}
        self.tape = []
        self.grad_bufs = []
        self.zero_list = []
        self.param_grads = {}
        self.device = None
        self.training = False
        self.record = False
        self.use_tma = backend is None   # False: force the register-gather tcgen05 kernels (tests / A-B timing / numpy spec)
        self._bn_pending = None  # deferred BatchNorm tails of layers that are ready together (bn_begin / bn_flush)
        self.pack_descs = {}     # cache key -> (vinet_pack_t, 16-byte chunks): everything refresh_packed_weights() re-packs
        self._pack_table = None
        self.weights_dirty = False   # set by GraphedTrainStep: parameters changed without a Tensor._version bump
        self._repack_all = False
        self.arena = None       # optional flat fp32 gradient arena: (flat tensor, {param name: (offset, numel)}, {name: parameter})
        self.par_branches = True
        self._side, self._side_active, self._forked = None, False, False
        self.realloc_count = 0  # pooled buffers re-allocated because a forward came with another shape
        self.dwp_arena, self.dwp_layout, self.dwp_dirty, self.unpack_queue = None, {}, False, []
        self.replica = False    # nn.DataParallel replica: its weights are fresh broadcast copies every forward (no multi-repack)
        self.consumed_gen = -1  # generation whose tape has been run: a second backward through it must fail loudly
        self.up2_mat_log = []   # names of the virtual up-sampled activations that had to be materialised (tests assert on it)
        self.backward_point_cb = None   # callable(name): invoked from the backward pass at the points marked in the forward plan
        self.arena_bypassed = False     # a gradient of this pass was handed out as a scratch tensor instead of its arena slice
        self.profile = None     # bench.py: list of (layer, kind, flops, start_event, end_event) per conv launch
        self.l2_flush = None

    # ------------------------------------------------------------------ plumbing
    def stream(self):
        if self.device.type != "cuda":
            return None
        if self._side_active:
            return self._side.cuda_stream
        return torch.cuda.current_stream(self.device).cuda_stream

    # ---- branch concurrency: independent convolutions of a Mixed block (branch1 / branch2 chains, the pool branch) are small
    # launches that leave most SMs idle; issuing one of each pair on a side stream lets them overlap.  Only KERNEL LAUNCHES move
    # to the side stream (torch's current stream is untouched, so every allocation stays on the main stream); fork / join are
    # events, which a CUDA-graph capture turns into plain dependency edges.
    def par_enabled(self):
        return (self.device is not None and self.device.type == "cuda" and self.profile is None and self.par_branches
                and not os.environ.get("VINET_NO_PAR_BRANCH"))

    def fork(self):
        """Make the side stream wait for everything issued so far on the main stream."""
        if not self.par_enabled():
            return False
        if self._side is None or self._side.device != self.device:
            self._side = torch.cuda.Stream(device=self.device)
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(self.device))
        self._side.wait_event(ev)
        self._forked = True
        return True

    def on_side(self, flag):
        """Route the following kernel launches to the side stream (True) or back to the main stream (False)."""
        self._side_active = bool(flag) and self._forked

    def join(self):
        """The main stream waits for the side stream; launches go to the main stream again."""
        self._side_active = False
        if not self._forked:
            return
        ev = torch.cuda.Event()
        ev.record(self._side)
        torch.cuda.current_stream(self.device).wait_event(ev)
        self._forked = False

    def call(self, name, desc):
        self.lib.call(name, C.byref(desc), self.stream())

    def buf(self, name, shape, dtype, zero=False):
        """Pooled scratch tensor.  zero=True: allocated zero-filled ONCE; its consumer kernel clears what it reads
        (BN statistics, packed weight gradients), so accumulate-into buffers need no per-step memset."""
        t = self.pool.get(name)
        if t is None or tuple(t.shape) != tuple(shape) or t.dtype != dtype or t.device != self.device:
            if t is not None:
                self.realloc_count += 1      # a captured CUDA graph that used the old buffer is stale now (graph.py checks)
            if zero:
                t = torch.zeros(shape, dtype=dtype, device=self.device)
                self.pool[name] = t
                return t
            t = torch.empty(shape, dtype=dtype, device=self.device)
            if _POISON and t.is_floating_point():
                t.fill_(float("nan"))       # debug aid: reads of never-written scratch surface as NaN
            self.pool[name] = t
        return t

    def grad_tensor(self, name, like, zero=False):
        """Where the gradient of parameter `name` is written: its slice of the flat arena when one is installed (all
        gradients then live in ONE buffer: a single NCCL all-reduce, no per-tensor bucket copies), else a fresh tensor."""
        if self.arena is not None:
            hit = self.arena[1].get(name)
            if hit is not None and self.arena[0].device == like.device:
                t = self.arena[0][hit[0]:hit[0] + hit[1]].view(like.shape)
                # autograd adopted the arena view as param.grad on the previous backward: if that gradient is still alive
                # (gradient accumulation, zero_grad(set_to_none=False)), writing the slice would destroy it and AccumulateGrad
                # would add the tensor to itself.  Hand out a scratch tensor instead: autograd then accumulates INTO the arena.
                p = self.arena[2].get(name) if len(self.arena) > 2 else None
                if p is None or p.grad is None or p.grad.data_ptr() != t.data_ptr():
                    return t.zero_() if zero else t
                self.arena_bypassed = True
        return torch.zeros_like(like) if zero else torch.empty_like(like)

    def timed(self, label, kind, flops, fn):
        """Run one kernel call; when profiling, bracket it with CUDA events on the launching stream (after
        evicting L2 by overwriting a buffer larger than it)."""
        if self.profile is None:
            return fn()
        if self.l2_flush is not None:
            self.memset(self.l2_flush)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        kern = self.lib.fn["vinet_last_kernel"]().decode() if "vinet_last_kernel" in self.lib.fn else ""
        self.profile.append((label, kind, flops, e0, e1, kern))

    def memset(self, t):
        self.lib.call("vinet_memset_async", t.data_ptr(), 0, t.numel() * t.element_size(), self.stream())

    def begin(self, device, training, record):
        self.device, self.training, self.record = device, training, record
        self.tape, self.grad_bufs, self.zero_list, self.param_grads = [], [], [], {}
        self.gwritten = set()
        self.bn_counters = []
        self._bn_pending = None
        self._repack_all, self.weights_dirty = self.weights_dirty, False
        if self.eng == L.ENGINE_TC and not self.replica:
            self.refresh_packed_weights()
        # BatchNorm batch statistics taken in the convolution epilogue (bf16 tensor-core training): one fp64 arena for all
        # layers of the step, bump-allocated in plan order and zeroed by ONE memset per step
        self.epi_stats = (self.eng == L.ENGINE_TC and not self.split and training and self.use_tma
                          and "vinet_bn_apply_stats_multi" in self.lib.fn and not os.environ.get("VINET_NO_EPI_STATS"))
        self.stats_off = 0
        self.epi_stats_1x1 = not os.environ.get("VINET_NO_EPI_STATS_1X1")
        if self.epi_stats:
            self.stats_arena = self.buf("bn.stats_arena", (1 << 16,), torch.float64)
            self.memset(self.stats_arena)

    def stats_slice(self, n):
        """Device pointer of 2*n zeroed doubles ([sum | sum of squares][n channels]) for one convolution of this step."""
        off = self.stats_off
        self.stats_off += 2 * n
        assert self.stats_off <= self.stats_arena.numel(), "BatchNorm statistics arena exhausted"
        return self.stats_arena.data_ptr() + 8 * off

    def end_forward(self):
        """nn.BatchNorm's num_batches_tracked counters of every layer touched by this forward, one launch."""
        if self.bn_counters:
            torch._foreach_add_(self.bn_counters, 1)
            self.bn_counters = []

    def new_act(self, name, B, T, H, W, C_, xform=L.XF_IDENT, affine=False, dtype=None, gdtype=None):
        buf = self.buf(name, (B, T, H, W, C_), dtype or self.tdtype)
        sc = sh = None
        if affine:
            ss = self.buf(name + ".ss", (2, C_), torch.float32)
            sc, sh = ss[0], ss[1]
        a = Act(buf, B, T, H, W, C_, 0, xform, sc, sh)
        if self.record:
            self.want_grad(a, name, gdtype)
        return a

    def want_grad(self, a, name, dtype=None):
        """Gradient w.r.t. the activated value, in the storage type (bf16 mode: bf16 gradients)."""
        dtype = dtype or self.tdtype
        a.grad = self.buf(name + ".grad", (a.B, a.T, a.H, a.W, a.C), dtype)
        a.gchoff, a.needs_grad = 0, True
        a.gdt = L.F32 if dtype == torch.float32 else L.BF16
        self.grad_bufs.append(a.grad)

    # Gradient buffers are never zeroed up front: the first writer of a buffer STORES (a convolution's data
    # gradient covers every element), later writers accumulate.  Scatter-style writers (max-pool backward)
    # and partial-coverage writers ask for an initialised buffer instead.
    def first_write(self, a):
        """True when `a` is the first (and full-coverage) writer of its gradient buffer."""
        key = a.grad.data_ptr()
        if key in self.gwritten:
            return False
        self.gwritten.add(key)
        if a.gchoff == 0 and a.C == a.ldg:
            return True
        self.memset(a.grad)
        return False

    def ensure_init(self, a):
        key = a.grad.data_ptr()
        if key not in self.gwritten:
            self.gwritten.add(key)
            self.memset(a.grad)

    # ------------------------------------------------------------------ descriptors
    def _src(self, s, a):
        s.ptr, s.scale, s.shift, s.ld, s.T, s.xform = a.ptr(), _ptr(a.scale), _ptr(a.shift), a.ld, a.T, a.xform
        s.ldb = a.ldb

    def _gather_fprop(self, g, srcs, geom, cs, To, Ho, Wo, taps=None):
        a0 = srcs[0]
        g.mode, g.dtype = L.GATHER_FPROP, self.opdt
        g.B, g.Tr, g.Hr, g.Wr, g.row_tstep, g.row_toff = a0.B, To, Ho, Wo, 1, 0
        g.Ts, g.Hs, g.Ws, g.Cs = sum(s.T for s in srcs), a0.H, a0.W, cs
        if taps is None:
            taps = _taps(geom.kt, geom.kh, geom.kw)
        g.ntaps = len(taps)
        _fill_taps(g.tap, taps)
        g.st, g.sh, g.sw, g.pt, g.ph, g.pw = geom.st, geom.sh, geom.sw, geom.pt, geom.ph, geom.pw
        self._src(g.src[0], srcs[0])
        if len(srcs) > 1:
            self._src(g.src[1], srcs[1])
        else:
            g.src[1].ptr, g.src[1].T = None, 0

    def _fill_win_gather(self, g, aw, geom, To, Ho, Wo, taps, npx):
        """FPROP gather over the padded clip as overlapping pixel windows: one K block per kernel row dh holds the (dw, c) window
        of npx pixels x aw.cpp channels (64 elements: 8 pixels of the 8-channel clip or 16 of the 4-channel one; 32 elements:
        the 8-pixel windows the streaming kernel reads in place from the 4-channel clip)."""
        g.mode, g.dtype = L.GATHER_FPROP, self.opdt
        g.B, g.Tr, g.Hr, g.Wr, g.row_tstep, g.row_toff = aw.B, To, Ho, Wo, 1, 0
        g.Ts, g.Hs, g.Ws, g.Cs, g.ntaps = aw.T, aw.H, Wo, npx * aw.cpp, len(taps)
        _fill_taps(g.tap, taps)
        g.st, g.sh, g.sw, g.pt, g.ph, g.pw = 1, geom.sh, 1, 0, geom.ph, 0
        s0 = g.src[0]
        s0.ptr, s0.scale, s0.shift, s0.ld, s0.T, s0.xform = aw.buf.data_ptr(), None, None, aw.cpp * geom.sw, aw.T, L.XF_IDENT
        s0.ldh = aw.Wp * aw.cpp
        g.src[1].ptr, g.src[1].T = None, 0

    @staticmethod
    def tiling(n):
        n16 = round_up(n, 16)
        n_tiles = cdiv(n16, 256)
        return round_up(cdiv(n16, n_tiles), 16), n_tiles

    def conv_tiling(self, d, n):
        """(block_n, n_tiles) for the conv described by `d` (gather / outputs / kernel filled in).  The TMA-fed tensor-core
        kernels pick their own N tiling per layer (csrc/conv_stream.cu); everything else uses the default."""
        if self.eng == L.ENGINE_TC and d.kernel == L.KERNEL_TMA and "vinet_conv_tiling" in self.lib.fn:
            bn, nt = C.c_int32(0), C.c_int32(0)
            d.N = n
            self.lib.call("vinet_conv_tiling", C.byref(d), self.eng, C.byref(bn), C.byref(nt))
            return bn.value, nt.value
        return self.tiling(n)

    def packed_weight(self, w, key, mode, taps, cs, n, layout=L.KLAYOUT_DENSE, tiling=None, part=0, cslice=None):
        """bf16-swizzled (TC) or fp32 (SIMT) GEMM B operand of a conv weight, cached per parameter version.
        part: which term of the weight's bf16 expansion (split-precision parity mode); cslice = (c0, c): only the reduction
        channels [c0, c0+c) (input channels for FPROP, output channels for DGRAD) — `cs` is then c."""
        block_n, n_tiles = tiling if tiling is not None else self.tiling(n)
        ck = (key, mode, tuple(taps), self.eng, layout, block_n, n_tiles, part, cslice)
        if layout != L.KLAYOUT_DENSE:
            k_blocks = len(taps) * cdiv(cs, L.TC_BLOCK_K)
        else:
            k_blocks = cdiv(len(taps) * cs, L.TC_BLOCK_K)
        hit = self.wcache.get(ck)
        if hit is not None and hit[0] is not None and hit[0] == _wstamp(w) and hit[1].device == w.device and hit[3] is w and not self._repack_all:
            return hit[1], block_n, n_tiles, k_blocks
        nbytes = self.lib.fn["vinet_packed_weight_bytes"](self.eng, n, block_n, n_tiles, k_blocks)
        out = hit[1] if hit is not None and hit[1].numel() == nbytes and hit[1].device == w.device else \
            torch.empty(nbytes, dtype=torch.uint8, device=w.device)
        d = L.Pack()
        wc = w.detach()
        assert wc.is_contiguous() and wc.dtype == torch.float32
        d.w, d.Cout, d.Cin, d.kt, d.kh, d.kw = wc.data_ptr(), *wc.shape
        woff = 0
        if cslice is not None:
            c0, c = cslice
            kk = wc.shape[2] * wc.shape[3] * wc.shape[4]
            if mode == L.GATHER_FPROP:          # a range of input channels: same strides, shifted base
                woff, d.ld_cin, d.Cin = 4 * c0 * kk, wc.shape[1], c
            else:                               # a range of output channels (the outermost dimension)
                woff, d.Cout = 4 * c0 * wc.shape[1] * kk, c
            d.w = wc.data_ptr() + woff
        d.cs, d.mode, d.ntaps = cs, mode, len(taps)
        _fill_taps(d.tap, taps)
        d.engine, d.block_n, d.n_tiles, d.k_blocks, d.out = self.eng, block_n, n_tiles, k_blocks, out.data_ptr()
        d.layout, d.part = layout, part
        self.call("vinet_pack_weights", d)
        self.wcache[ck] = (_wstamp(w), out, None, w)
        if self.eng == L.ENGINE_TC and "vinet_pack_weights_multi" in self.lib.fn:
            old = self.pack_descs.get(ck)
            if old is None or old[0].w != d.w or old[0].out != d.out:
                self._pack_table = None                       # a new entry / moved storage: the device table must be rebuilt
            self.pack_descs[ck] = (d, n_tiles * k_blocks * block_n * 8, woff)
        return out, block_n, n_tiles, k_blocks

    def refresh_packed_weights(self):
        """Parameters changed (optimizer step): re-pack EVERY cached weight with one launch instead of one per conv call."""
        if not self.pack_descs or self.device is None or self.device.type != "cuda":
            return
        stale = self._repack_all or any(self.wcache[ck][0] != _wstamp(self.wcache[ck][3]) for ck in self.pack_descs)
        if not stale:
            return
        for key, hit in self.wcache.items():                   # fused 1x1 groups: refresh the concatenated weights in place
            if isinstance(key, str) and key.endswith(".wcat"):
                with torch.no_grad():
                    torch.cat([w.detach() for w in hit[2]], 0, out=hit[1])
                self.wcache[key] = (tuple(_wstamp(w) for w in hit[2]), hit[1], hit[2])
        for ck, (dsc, _, woff) in self.pack_descs.items():      # parameters whose storage moved (param.data = ..., .to())
            ptr = self.wcache[ck][3].data_ptr() + woff
            if dsc.w != ptr:
                dsc.w = ptr
                self._pack_table = None
        if self._pack_table is None:
            keys = list(self.pack_descs)
            n = len(keys)
            tab = (L.Pack * n)()
            begin, tot = [], 0
            for i, ck in enumerate(keys):
                C.memmove(C.byref(tab[i]), C.byref(self.pack_descs[ck][0]), C.sizeof(L.Pack))
                begin.append(tot)
                tot += self.pack_descs[ck][1]
            tab_dev = torch.frombuffer(bytearray(bytes(tab)), dtype=torch.uint8).to(self.device)
            beg_dev = torch.tensor(begin, dtype=torch.int64, device=self.device)
            self._pack_table = (keys, tab_dev, beg_dev, tot)
        keys, tab_dev, beg_dev, tot = self._pack_table
        self.lib.call("vinet_pack_weights_multi", tab_dev.data_ptr(), beg_dev.data_ptr(), len(keys), tot, self.stream())
        for ck in keys:
            hit = self.wcache[ck]
            self.wcache[ck] = (_wstamp(hit[3]), hit[1], None, hit[3])
        self._repack_all = False

    # ---- packed weight gradients: one zeroed arena per backward pass, unpacked by a handful of batched launches ----
    def dwp_buf(self, name, rows, lddw):
        """Zeroed fp32 [rows, lddw] accumulator of one convolution's packed weight gradient.  All of a step's accumulators live in
        ONE arena that the backward pass clears with a single memset (they used to cost ~66 memset launches per step); the arena
        layout is discovered on the first backward pass and fixed afterwards."""
        n = rows * lddw
        lay = self.dwp_layout.get(name)
        if self.dwp_arena is not None and lay is not None and lay[1] == n:
            return self.dwp_arena[lay[0]:lay[0] + n].view(rows, lddw)
        self.dwp_layout[name] = (None, n)          # not (or no longer) in the arena: private buffer now, arena rebuilt next pass
        self.dwp_dirty = True
        t = self.buf(name + ".dwp", (rows, lddw), torch.float32)
        self.memset(t)
        return t

    def dwp_begin(self):
        if self.dwp_dirty and self.dwp_layout:
            off = 0
            for k, (_, n) in self.dwp_layout.items():
                self.dwp_layout[k] = (off, n)
                off += (n + 63) // 64 * 64
            self.dwp_arena = torch.empty(off, dtype=torch.float32, device=self.device)
            for k in self.dwp_layout:
                self.pool.pop(k + ".dwp", None)          # the private first-pass buffers are no longer needed
            self.realloc_count += 1
            self.dwp_dirty = False
        if self.dwp_arena is not None:
            self.memset(self.dwp_arena)

    def unpack(self, dwp_ptr, lddw, cs, grad, cout, cin, ntaps, win8=None):
        """Queue `grad <- packed gradient` (flushed in batches of UNPACK_MAX launches-worth by unpack_flush)."""
        if "vinet_unpack_wgrad_multi" not in self.lib.fn:
            if win8:
                self.lib.call("vinet_unpack_wgrad_win8", dwp_ptr, lddw, grad.data_ptr(), cout, cin, win8[0], win8[1], self.stream())
            else:
                self.lib.call("vinet_unpack_wgrad", dwp_ptr, lddw, cs, grad.data_ptr(), cout, cin, ntaps, self.stream())
            return
        self.unpack_queue.append((dwp_ptr, lddw, cs, grad, cout, cin, ntaps, win8 or (0, 0)))
        if len(self.unpack_queue) == L.UNPACK_MAX:
            self.unpack_flush()

    def unpack_flush(self):
        q, self.unpack_queue = self.unpack_queue, []
        if not q:
            return
        tab = (L.Unpack * len(q))()
        for i, (dwp_ptr, lddw, cs, grad, cout, cin, ntaps, win8) in enumerate(q):
            e = tab[i]
            e.dwp, e.grad, e.lddw, e.cs, e.Cout, e.Cin, e.ntaps = dwp_ptr, grad.data_ptr(), lddw, cs, cout, cin, ntaps
            e.win8_kh, e.win8_kw = win8
        self.lib.call("vinet_unpack_wgrad_multi", tab, len(q), self.stream())

    # ------------------------------------------------------------------ convolution
    def tma_ok(self, srcs, geom):
        """The TMA-fed kernels need bf16 sources without pending transforms and spatial stride 1."""
        return (self.eng == L.ENGINE_TC and self.use_tma and geom.sh == 1 and geom.sw == 1
                and all(s.xform == L.XF_IDENT or (i == 0 and s.up2) for i, s in enumerate(srcs)))

    # ---- 2x bilinear up-sampling fused into the consumer's input stage (model.py:254 in front of model.py:260-275) ----
    def materialize_up2(self, a):
        """Hi-res copy of a virtual up-sampled activation, for a consumer without a fused input stage (small maps, the
        register-gather tensor-core kernels).  Shares the gradient buffer of `a`."""
        if a.mat is None:
            low_h, low_w = a.H // 2, a.W // 2
            m = Act(self.buf(a.name + ".up", (a.B, a.T, a.H, a.W, a.C), self.tdtype), a.B, a.T, a.H, a.W, a.C)
            m.grad, m.gchoff, m.needs_grad, m.gdt = a.grad, a.gchoff, a.needs_grad, a.gdt
            d = L.Upsample()
            d.z, d.ldz, d.dtype, d.relu, d.B, d.T, d.h, d.w, d.C = a.ptr(), a.ld, self.dt, a.xform & 1, a.B, a.T, low_h, low_w, a.C
            d.u, d.ldu, d.u_dtype = m.ptr(), m.ld, self.dt
            self.call("vinet_upsample_fwd", d)
            a.mat = m
            self.up2_mat_log.append(a.name)
        return a.mat

    def resolve_up2(self, srcs, geom, cout):
        """Sources a convolution can read as they are: a virtual up-sampled source stays virtual when the kernels that will
        serve this convolution (forward AND weight gradient) interpolate in their input stage, else it is materialised."""
        if not any(s.up2 for s in srcs):
            return srcs
        keep = srcs[0].up2 and not any(s.up2 for s in srcs[1:]) and "vinet_conv_up2_fused" in self.lib.fn
        if keep and not self.split:          # (split-precision mode: the operand split interpolates while it splits)
            if self.eng == L.ENGINE_TC and not (self.use_tma and geom.sh == 1 and geom.sw == 1
                                                and all(s.xform == L.XF_IDENT for s in srcs[1:])):
                keep = False
            else:
                g = L.Gather()
                To, Ho, Wo = geom.out_dims(sum(s.T for s in srcs), srcs[0].H, srcs[0].W)
                self._gather_fprop(g, srcs, geom, srcs[0].C, To, Ho, Wo)
                keep = bool(self.lib.fn["vinet_conv_up2_fused"](C.byref(g), cout, self.eng, L.KERNEL_TMA))
        if keep:
            return srcs
        return [self.materialize_up2(s) if s.up2 else s for s in srcs]

    # ---- split-precision operands (parity mode "bf16x3" / "bf16x6") ----
    def split_view(self, name, ptr, ld, dtype, rows, C_, xform=L.XF_IDENT, scale=None, shift=None, up=None):
        """bf16 expansion planes [rows, C] of an fp32 view (after its pending transform): x ~= part0 + part1 (+ part2).
        up = (h, w): the view is a low-res tensor read through the 2x up-sampling (xform has XF_UP2); rows are hi-res rows."""
        parts = [self.buf("%s.%d" % (name, i), (rows, C_), torch.bfloat16) for i in range(self.split)]
        d = L.Split()
        d.x, d.ld, d.dtype, d.rows, d.C = ptr, ld, dtype, rows, C_
        d.scale, d.shift, d.xform, d.nparts, d.ldo = _ptr(scale), _ptr(shift), xform, self.split, C_
        if up is not None:
            d.up_h, d.up_w = up
        for i, t in enumerate(parts):
            d.part[i] = t.data_ptr()
        self.call("vinet_split_bf16", d)
        return parts

    # tcgen05.mma adds into its fp32 TMEM accumulator with truncation: a GEMM of K elements carries a bias of about
    # (K/16) * 2^-24 relative (measured through the goldens: 1e-4 on the K = 22,464 decoder conv).  The parity mode therefore
    # issues the leading (hi x hi) term in tap chunks of at most SPLIT_KCHUNK reduction elements, each starting from a zero
    # accumulator and added to the fp32 output by the epilogue's round-to-nearest add.  The other terms are 2^-8 smaller.
    SPLIT_KCHUNK = 256

    def term_launches(self, taps, cs, chunk_channels=True):
        """[(operand part, weight part, tap subset, first channel, channels)] of one GEMM, in issue order (the first launch
        stores, the rest accumulate).  Chunks of the leading term are whole taps, or 64-aligned channel ranges of one tap."""
        if not self.split:
            return [(0, 0, taps, 0, cs)]
        out = []
        kc = self.SPLIT_KCHUNK
        for pa, pw in self.terms:
            if (pa, pw) != (0, 0) or len(taps) * cs <= kc:
                out.append((pa, pw, taps, 0, cs))
            elif cs <= kc or not chunk_channels:
                per = max(1, kc // cs)
                out += [(0, 0, taps[i:i + per], 0, cs) for i in range(0, len(taps), per)]
            else:
                out += [(0, 0, [t], c0, min(kc, cs - c0)) for t in taps for c0 in range(0, cs, kc)]
        return out

    def split_sources(self, name, srcs):
        """[part][source] operand Acts of a convolution: plain bf16 NDHWC tensors the TMA-fed kernels can fetch."""
        out = [[] for _ in range(self.split)]
        for si, a in enumerate(srcs):
            if isinstance(a, WinAct):
                rows, C_ = a.B * a.T * a.H * a.Wp, 8
                planes = self.split_view("%s.sp%d" % (name, si), a.buf.data_ptr(), 8, self.dt, rows, C_)
                for i, t in enumerate(planes):
                    assert a.cpp == 8
                    out[i].append(WinAct(t.view(a.B, a.T, a.H, a.Wp, 8), a.B, a.T, a.H, a.W, a.wl, a.Wp))
            else:
                planes = self.split_view("%s.sp%d" % (name, si), a.ptr(), a.ld, self.dt, a.rows, a.C, a.xform, a.scale, a.shift,
                                         up=(a.H // 2, a.W // 2) if a.up2 else None)
                for i, t in enumerate(planes):
                    out[i].append(Act(t.view(a.B, a.T, a.H, a.W, a.C), a.B, a.T, a.H, a.W, a.C))
        return out

    def conv(self, name, srcs, w, geom, out, bias=None, cin_real=None, ep=None, wgrad_split=None, stats=None):
        """out <- raw conv of the (virtual T-concat of) srcs; returns a backward closure taking dY.
        stats: device pointer of the layer's statistics slice (stats_slice): the TMA-fed kernels add the per-channel sum and sum
        of squares of the raw output to it from their fp32 accumulators (BatchNorm batch statistics without a read pass).
        ep = (scale, shift, act): fused per-channel epilogue (inference-mode BatchNorm folding).
        wgrad_split = [(parameter name, out channels)]: `w` is a concatenation of several parameters along Cout; the
        weight gradient is unpacked straight into each member's gradient tensor.
        Split-precision parity mode: every GEMM below is issued once per (operand part, weight part) term into the same fp32
        output (first term stores, the others accumulate); the kernels are the ones the bf16 mode runs."""
        srcs = self.resolve_up2(srcs, geom, w.shape[0])
        if self.split:
            assert ep is None, "the split-precision mode accumulates raw terms: no non-linear epilogue"
            psrcs = self.split_sources(name, srcs)
        else:
            psrcs = [srcs]
        terms = self.terms
        a0 = psrcs[0][0]
        cs = a0.C
        Cout = w.shape[0]
        win = isinstance(a0, WinAct)
        tma = win or self.tma_ok(psrcs[0], geom)
        kern = L.KERNEL_TMA if tma else L.KERNEL_GATHER
        layout = ((L.KLAYOUT_WIN8 if a0.cpp == 8 else L.KLAYOUT_WIN4) if win else (L.KLAYOUT_TAP64 if tma else L.KLAYOUT_DENSE))
        To, Ho, Wo = geom.out_dims(sum(s.T for s in srcs), a0.H, a0.W)
        assert (To, Ho, Wo, Cout) == (out.T, out.H, out.W, out.C), (name, (To, Ho, Wo, Cout), (out.T, out.H, out.W, out.C))
        assert w.shape[1] == (cin_real or cs), name
        taps = _taps(geom.kt, geom.kh, geom.kw)
        nreal = len(taps)
        if win:
            # stem: one 64-wide K block per kernel row dh holds the (dw, c) window of 8 pixels x 8 channels;
            # output column wo reads padded pixels wo*sw .. wo*sw+7, i.e. consecutive windows overlap
            assert len(srcs) == 1 and geom.kt == 1 and geom.st == 1 and geom.kw <= 8 and geom.pw == a0.wl and w.shape[1] <= 8
            assert (Wo - 1) * geom.sw + 64 // a0.cpp <= a0.Wp
            taps, cs = [(0, dh, 0) for dh in range(geom.kh)], 64

        def fill_gather(g, ss, tp=None, c0=0, c=None):
            tp = taps if tp is None else tp
            if not win:
                if c is not None and c != cs:
                    ss = [a.slice(c0, c) for a in ss]
                return self._gather_fprop(g, ss, geom, cs if c is None else c, To, Ho, Wo, tp)
            self._fill_win_gather(g, ss[0], geom, To, Ho, Wo, tp, 64 // ss[0].cpp)

        cin_r = w.shape[1]
        flops = 2.0 * a0.B * To * Ho * Wo * nreal * cin_r * Cout
        # forward of the stem on the 4-channel copy of the clip: windows of 8 pixels x 4 channels (Cs = 32) that start
        # 4*sw elements apart, read in place by the streaming kernel (csrc/conv_stream.cu, conv_gemm_stream_win4)
        win4 = False
        if win and a0.cpp == 4 and w.shape[1] <= 4 and geom.sw == 2 and a0.Wp % 2 == 0 and "vinet_conv_win4_fused" in self.lib.fn:
            g4 = L.Gather()
            self._fill_win_gather(g4, a0, geom, To, Ho, Wo, taps, 8)
            win4 = bool(self.lib.fn["vinet_conv_win4_fused"](C.byref(g4), Cout))
        for ti, (pa, pw, tp, c0, cc) in enumerate(self.term_launches(taps, cs, chunk_channels=not win)):
            d = L.Conv()
            d.kernel = kern
            if win4:
                self._fill_win_gather(d.g, a0, geom, To, Ho, Wo, tp, 8)
                cc = 32
            else:
                fill_gather(d.g, psrcs[pa], tp, c0, cc)
            lay = layout
            d.out[0], d.ldo[0], d.out_T[0] = out.ptr(), out.ld, To
            d.out[1], d.ldo[1], d.out_T[1] = None, 0, 0
            d.out_dtype, d.accumulate = self.dt, (0 if ti == 0 else 1)
            d.ep_scale, d.ep_shift, d.ep_act = None, (_ptr(bias) if ti == 0 else None), L.ACT_NONE
            if stats is not None:
                assert kern == L.KERNEL_TMA and not self.split and bias is None and ep is None
                d.stats = stats
            if ep is not None:
                assert bias is None
                d.ep_scale, d.ep_shift, d.ep_act = _ptr(ep[0]), _ptr(ep[1]), ep[2]
            wp, block_n, n_tiles, k_blocks = self.packed_weight(w, name, L.GATHER_FPROP, tp, cc, Cout, lay,
                                                                tiling=self.conv_tiling(d, Cout), part=pw,
                                                                cslice=None if (cc == cs or win4) else (c0, cc))
            d.w, d.N, d.block_n, d.n_tiles, d.k_blocks = wp.data_ptr(), Cout, block_n, n_tiles, k_blocks
            self.timed(name, "fprop", flops * len(tp) * (cs if win4 else cc) / (len(taps) * cs),
                       lambda: self.lib.call("vinet_conv_gemm", C.byref(d), self.eng, self.stream()))
        if not self.record:
            return None

        def backward(dy, lddy):
            rows = a0.B * To * Ho * Wo
            # operand form of dY: the tensor itself, or its bf16 expansion planes in the split-precision mode
            if self.split:
                planes = self.split_view("dyp.%d" % (rows * Cout), dy, lddy, self.dt, rows, Cout)
                dyp = [(t.data_ptr(), Cout) for t in planes]
            else:
                dyp = [(dy, lddy)]
            # ---- weight gradient
            csk = round_up(cs, 64) if tma else cs          # TAP64 rows of the packed gradient
            ktot = len(taps) * csk
            lddw = round_up(Cout, 64)
            dwp = self.dwp_buf(name, round_up(ktot, 128), lddw)
            for pa, pd in terms:
                wg = L.Wgrad()
                wg.kernel = kern
                fill_gather(wg.g, psrcs[pa])
                wg.dy, wg.lddy, wg.dy_dtype, wg.N, wg.dwp, wg.lddw = dyp[pd][0], dyp[pd][1], self.opdt, Cout, dwp.data_ptr(), lddw
                if self.eng == L.ENGINE_TC:
                    tiles = cdiv(ktot, 128) * self.tiling(Cout)[1]
                    chunks = cdiv(rows, 64)
                else:
                    tiles = cdiv(ktot, 64) * cdiv(Cout, 64)
                    chunks = cdiv(rows, 16)
                wg.splits = max(1, min(cdiv(2 * 148, tiles), cdiv(chunks, 4)))
                self.timed(name, "wgrad", flops, lambda: self.lib.call("vinet_conv_wgrad", C.byref(wg), self.eng, self.stream()))
            if wgrad_split is not None:
                off = 0
                for pname, c in wgrad_split:      # member columns [off, off+c) of the packed gradient
                    gwm = self.grad_tensor(pname, w[off:off + c])
                    self.unpack(dwp.data_ptr() + 4 * off, lddw, csk, gwm, c, w.shape[1], len(taps))
                    self.param_grads[pname] = gwm
                    off += c
            else:
                gw = self.grad_tensor(name + ".weight", w)
                if win:
                    self.unpack(dwp.data_ptr(), lddw, 0 if a0.cpp == 8 else a0.cpp, gw, Cout, w.shape[1], 0, win8=(geom.kh, geom.kw))
                else:
                    self.unpack(dwp.data_ptr(), lddw, csk, gw, Cout, w.shape[1], len(taps))
                self.param_grads[name + ".weight"] = gw
            if bias is not None:
                gb = self.grad_tensor(name + ".bias", bias)
                ws = self.buf("colsum.ws", (1024,), torch.float64)
                self.lib.call("vinet_colsum", dy, lddy, self.dt, rows, Cout, ws.data_ptr(), gb.data_ptr(), self.stream())
                self.param_grads[name + ".bias"] = gb
            # ---- data gradient, one launch per temporal phase of the transposed convolution
            if not any(s.needs_grad for s in srcs):
                return
            Ti = sum(s.T for s in srcs)
            gdt = srcs[0].gdt
            assert all(s.gdt == gdt for s in srcs)
            phases = []
            for rho in range(geom.st):
                dts = [dt for dt in range(geom.kt) if dt % geom.st == rho]
                t0 = (rho - geom.pt) % geom.st
                frames = len(range(t0, Ti, geom.st))
                if dts and frames:
                    phases.append((dts, t0, frames))
            if len(phases) < min(geom.st, Ti):          # some frames receive no gradient from this conv
                for s in srcs:
                    self.ensure_init(s)
            accmask = sum((0 if self.first_write(s) else 1) << i for i, s in enumerate(srcs))
            allmask = (1 << len(srcs)) - 1
            for dts, t0, frames in phases:
                ptaps = [(dt, b, c) for dt in dts for b in range(geom.kh) for c in range(geom.kw)]
                n = w.shape[1]
                dtma = self.eng == L.ENGINE_TC and self.use_tma and geom.sh == 1 and geom.sw == 1
                for ti, (pd, pw, tp, c0, cc) in enumerate(self.term_launches(ptaps, Cout)):
                    dd = L.Conv()
                    dd.kernel = L.KERNEL_TMA if dtma else L.KERNEL_GATHER
                    g = dd.g
                    g.mode, g.dtype = L.GATHER_DGRAD, self.opdt
                    g.B, g.Tr, g.Hr, g.Wr, g.row_tstep, g.row_toff = a0.B, frames, a0.H, a0.W, geom.st, t0
                    g.Ts, g.Hs, g.Ws, g.Cs, g.ntaps = To, Ho, Wo, cc, len(tp)
                    _fill_taps(g.tap, tp)
                    g.st, g.sh, g.sw, g.pt, g.ph, g.pw = geom.st, geom.sh, geom.sw, geom.pt, geom.ph, geom.pw
                    g.src[0].ptr, g.src[0].scale, g.src[0].shift = dyp[pd][0] + 2 * c0 * (1 if self.split else 0), None, None
                    g.src[0].ld, g.src[0].T, g.src[0].xform = dyp[pd][1], To, L.XF_IDENT
                    g.src[1].ptr, g.src[1].T = None, 0
                    for i, s in enumerate(srcs):
                        dd.out[i], dd.ldo[i], dd.out_T[i] = s.gptr(), s.ldg, s.T
                    if len(srcs) == 1:
                        dd.out[1], dd.ldo[1], dd.out_T[1] = None, 0, 0
                    dd.out_dtype, dd.accumulate = gdt, (accmask if ti == 0 else allmask)
                    dd.ep_scale, dd.ep_shift, dd.ep_act = None, None, L.ACT_NONE
                    wpd, bn_, nt_, kb_ = self.packed_weight(w, name, L.GATHER_DGRAD, tp, cc, n,
                                                            L.KLAYOUT_TAP64 if dtma else L.KLAYOUT_DENSE,
                                                            tiling=self.conv_tiling(dd, n), part=pw,
                                                            cslice=None if cc == Cout else (c0, cc))
                    dd.w, dd.N, dd.block_n, dd.n_tiles, dd.k_blocks = wpd.data_ptr(), n, bn_, nt_, kb_
                    dflops = 2.0 * a0.B * frames * a0.H * a0.W * len(tp) * cc * n
                    self.timed(name, "dgrad", dflops, lambda: self.lib.call("vinet_conv_gemm", C.byref(dd), self.eng, self.stream()))
        return backward

    # ------------------------------------------------------------------ conv + BatchNorm (+ReLU pending)
    def conv_bn(self, name_conv, name_bn, srcs, w, bn, geom, out, cin_real=None, defer_to=None):
        """BasicConv3d / one half of SepConv3d (model_utils.py:128-160).
        Training: raw conv -> batch statistics -> finalize (scale/shift, running stats) -> materialise
        relu(scale*y+shift) into `out` (possibly a channel slice of a Mixed concat buffer).  Consumers then read
        plain activations, which is what lets the TMA-fed kernels fetch them.  Pure inference (no tape, running
        statistics): scale/shift/ReLU are folded into the conv epilogue and nothing else is launched.
        out=None, defer_to=name: the caller's only consumer applies pending transforms itself (a max-pool): in the bf16 training
        mode the layer is NOT materialised - the returned activation is the raw output tagged scale/shift/ReLU (statistics from
        the conv epilogue, one tiny finalise launch), which saves a full read + write pass (the stem's second half: 1.4 GB)."""
        if out is None:
            a0 = srcs[0]
            To, Ho, Wo = geom.out_dims(sum(s.T for s in srcs), a0.H, a0.W)
            Cn = w.shape[0]
            deferred = (self.training and not self.split and self.epi_stats and self.tma_ok(srcs, geom) and "vinet_bn_finalize" in self.lib.fn
                        and not os.environ.get("VINET_NO_DEFER_BN"))
            if deferred:
                raw = Act(self.buf(name_conv + ".raw", (a0.B, To, Ho, Wo, Cn), self.tdtype), a0.B, To, Ho, Wo, Cn)
                stats = (self.stats_slice(Cn), Cn)
                conv_bwd = self.conv(name_conv, srcs, w, geom, raw, cin_real=cin_real, stats=stats[0])
                st = self.buf(name_bn + ".stat", (2, Cn), torch.float32)      # mean, invstd
                ss = self.buf(name_bn + ".ss", (2, Cn), torch.float32)        # scale, shift
                fin = L.BnFinalize()
                fin.rows, fin.C, fin.gamma, fin.beta = raw.rows, Cn, bn.weight.data_ptr(), bn.bias.data_ptr()
                fin.eps, fin.momentum, fin.training = bn.eps, bn.momentum, 1
                fin.running_mean, fin.running_var = bn.running_mean.data_ptr(), bn.running_var.data_ptr()
                fin.scale, fin.shift, fin.mean, fin.invstd = ss[0].data_ptr(), ss[1].data_ptr(), st[0].data_ptr(), st[1].data_ptr()
                fin.sums = stats[0]
                self.call("vinet_bn_finalize", fin)
                self.bn_counters.append(bn.num_batches_tracked)
                out = Act(raw.buf, a0.B, To, Ho, Wo, Cn, 0, L.XF_AFFINE_RELU, ss[0], ss[1])
                if self.record:
                    self.want_grad(out, defer_to)
                    self._bn_tape(name_bn, bn, raw, out, conv_bwd, None, st, ss)
                return out
            out = self.new_act(defer_to, a0.B, To, Ho, Wo, Cn)
        Cn = out.C
        if not self.training and not self.record and not self.split:
            ss = self._bn_finalize(name_bn, bn, out.rows, Cn, None)[1]
            out.xform, out.scale, out.shift = L.XF_IDENT, None, None
            self.conv(name_conv, srcs, w, geom, out, cin_real=cin_real, ep=(ss[0], ss[1], L.ACT_RELU))
            return out
        raw = Act(self.buf(name_conv + ".raw", (out.B, out.T, out.H, out.W, Cn), self.tdtype), out.B, out.T, out.H, out.W, Cn)
        st = None
        if self.epi_stats and (isinstance(srcs[0], WinAct) or self.tma_ok(srcs, geom)) and (self.epi_stats_1x1 or geom.kt * geom.kh * geom.kw > 1):
            st = (self.stats_slice(Cn), Cn)
        conv_bwd = self.conv(name_conv, srcs, w, geom, raw, cin_real=cin_real, stats=st[0] if st else None)
        self._bn_tail(name_bn, bn, raw, out, conv_bwd, None, stats=st)
        return out

    def _bn_finalize(self, name_bn, bn, rows, Cn, raw, apply=None):
        """Batch statistics of `raw` (training) or the running statistics -> (mean/invstd, scale/shift) buffers.
        apply = a filled BnApply descriptor: also materialise relu(scale*y+shift); training mode does all three steps in ONE
        cooperative launch (vinet_bn_fwd_fused)."""
        st = self.buf(name_bn + ".stat", (2, Cn), torch.float32)      # mean, invstd
        ss = self.buf(name_bn + ".ss", (2, Cn), torch.float32)        # scale, shift
        fin = L.BnFinalize()
        fin.rows, fin.C, fin.gamma, fin.beta = rows, Cn, bn.weight.data_ptr(), bn.bias.data_ptr()
        fin.eps, fin.momentum = bn.eps, bn.momentum
        fin.running_mean, fin.running_var = bn.running_mean.data_ptr(), bn.running_var.data_ptr()
        fin.training = 1 if self.training else 0
        fin.scale, fin.shift, fin.mean, fin.invstd = ss[0].data_ptr(), ss[1].data_ptr(), st[0].data_ptr(), st[1].data_ptr()
        if apply is not None:
            apply.scale, apply.shift = ss[0].data_ptr(), ss[1].data_ptr()
        done = False
        if self.training:
            sums = self.buf(name_bn + ".sums", (2 * Cn + 4,), torch.float64, zero=True)    # [2][C] sums + ticket / flag / departures
            sd = L.BnStats()
            sd.y, sd.ld, sd.dtype, sd.rows, sd.C, sd.sums = raw.ptr(), raw.ld, self.dt, rows, Cn, sums.data_ptr()
            fin.sums = sums.data_ptr()
            # one cooperative launch pays off for very large tensors only (measured, tools/bn_bench.py): small layers are
            # latency-bound and two plain launches pipeline better than one launch with a grid-wide wait
            if apply is not None and "vinet_bn_fwd_fused" in self.lib.fn and rows * Cn * raw.buf.element_size() >= (400 << 20):
                self.lib.call("vinet_bn_fwd_fused", C.byref(sd), C.byref(fin), C.byref(apply), self.stream())
                done = True
            elif "vinet_bn_stats_finalize" in self.lib.fn:      # one launch: the last block finalises
                self.lib.call("vinet_bn_stats_finalize", C.byref(sd), C.byref(fin), self.stream())
            else:
                self.call("vinet_bn_stats", sd)
                self.call("vinet_bn_finalize", fin)
        else:
            self.call("vinet_bn_finalize", fin)
        if apply is not None and not done:
            self.call("vinet_bn_apply", apply)
        if self.training:
            self.bn_counters.append(bn.num_batches_tracked)   # += 1 for all layers in one launch (end_forward)
        return st, ss

    # ---- multi-layer BatchNorm launches: layers that are ready at the same time share ONE launch per pass ----
    def bn_begin(self):
        if self.training and "vinet_bn_stats_finalize_multi" in self.lib.fn:
            self._bn_pending = []

    def bn_end(self):
        self.bn_flush()
        self._bn_pending = None

    def bn_flush(self):
        pend = self._bn_pending
        if not pend:
            return
        self._bn_pending = []
        # one launch handles layers whose incoming gradients share a storage type (AViNet keeps the gradient of the last
        # Mixed block's output in fp32 for the audio-visual fusion while its inner activations use bf16 gradients)
        groups = {}
        for mbr in pend:
            groups.setdefault(mbr[3].gdt, []).append(mbr)
        for members in groups.values():
            for i in range(0, len(members), 4):
                self._bn_batch(members[i:i + 4])

    def _bn_fwd_multi(self, members):
        """Training forward of n <= 4 BatchNorm(+ReLU) layers: (statistics ->) finalise -> materialise.  Members whose statistics
        came out of the convolution epilogue need ONE launch (vinet_bn_apply_stats_multi); the others a statistics + finalise
        launch and an apply launch.  Returns [(mean/invstd buffer, scale/shift buffer)] per member."""
        n = len(members)
        sds, fins, aps, keep = (L.BnStats * n)(), (L.BnFinalize * n)(), (L.BnApply * n)(), []
        epi = all(mbr[6] is not None for mbr in members)
        for i, (name_bn, bn, raw, out, conv_bwd, dy_slot, stats) in enumerate(members):
            Cn, rows = out.C, out.rows
            st = self.buf(name_bn + ".stat", (2, Cn), torch.float32)
            ss = self.buf(name_bn + ".ss", (2, Cn), torch.float32)
            sd, fin, ap = sds[i], fins[i], aps[i]
            if epi:
                fin.sums, fin.sq_stride = stats
            else:
                sums = self.buf(name_bn + ".sums", (2 * Cn + 4,), torch.float64, zero=True)
                sd.y, sd.ld, sd.dtype, sd.rows, sd.C, sd.sums = raw.ptr(), raw.ld, self.dt, rows, Cn, sums.data_ptr()
                fin.sums, fin.sq_stride = sums.data_ptr(), 0
            fin.rows, fin.C, fin.gamma, fin.beta = rows, Cn, bn.weight.data_ptr(), bn.bias.data_ptr()
            fin.eps, fin.momentum, fin.training = bn.eps, bn.momentum, 1
            fin.running_mean, fin.running_var = bn.running_mean.data_ptr(), bn.running_var.data_ptr()
            fin.scale, fin.shift, fin.mean, fin.invstd = ss[0].data_ptr(), ss[1].data_ptr(), st[0].data_ptr(), st[1].data_ptr()
            ap.y, ap.ldy, ap.dtype, ap.rows, ap.C, ap.relu = raw.ptr(), raw.ld, self.dt, rows, Cn, 1
            ap.scale, ap.shift, ap.out, ap.ldo, ap.out_dtype = ss[0].data_ptr(), ss[1].data_ptr(), out.ptr(), out.ld, self.dt
            out.xform, out.scale, out.shift = L.XF_IDENT, None, None
            self.bn_counters.append(bn.num_batches_tracked)
            keep.append((st, ss))
        if epi:
            self.lib.call("vinet_bn_apply_stats_multi", fins, aps, n, self.stream())
        else:
            self.lib.call("vinet_bn_stats_finalize_multi", sds, fins, n, self.stream())
            self.lib.call("vinet_bn_apply_multi", aps, n, self.stream())
        return keep

    def _bn_batch(self, members):
        """members: [(name_bn, bn, raw, out, conv_bwd, dy_slot, stats)] - small train-mode layers: one (statistics +) finalise +
        apply pass; backward one reduce launch + one apply launch, then each member's convolution backward."""
        n = len(members)
        keep = self._bn_fwd_multi(members)
        if not self.record:
            return

        def backward():
            bs, outs = (L.BnBwd * n)(), []
            for i, (name_bn, bn, raw, out, conv_bwd, dy_slot, _) in enumerate(members):
                Cn, rows = out.C, out.rows
                st, ss = keep[i]
                bsums = self.buf(name_bn + ".bsums", (2 * Cn + 4,), torch.float64, zero=True)
                dgamma, dbeta = self.grad_tensor(name_bn + ".weight", bn.weight), self.grad_tensor(name_bn + ".bias", bn.bias)
                if dy_slot is None:
                    # members of one launch need distinct dY buffers
                    dy = self.buf("dy.%d.%d" % (rows * Cn, i), (rows, Cn), self.tdtype)
                    dy_ptr, lddy = dy.data_ptr(), Cn
                else:
                    dy_ptr, lddy = dy_slot
                b = bs[i]
                b.g, b.ldg, b.y, b.ldy, b.dtype, b.rows, b.C, b.relu = out.gptr(), out.ldg, raw.ptr(), raw.ld, self.dt, rows, Cn, 1
                b.scale, b.shift, b.mean, b.invstd = ss[0].data_ptr(), ss[1].data_ptr(), st[0].data_ptr(), st[1].data_ptr()
                b.gamma, b.sums, b.dgamma, b.dbeta = bn.weight.data_ptr(), bsums.data_ptr(), dgamma.data_ptr(), dbeta.data_ptr()
                b.dy, b.lddy, b.dy_dtype, b.training, b.g_dtype = dy_ptr, lddy, self.dt, 1, out.gdt
                self.param_grads[name_bn + ".weight"] = dgamma
                self.param_grads[name_bn + ".bias"] = dbeta
                outs.append((conv_bwd, dy_ptr, lddy))
            assert len({mbr[3].gdt for mbr in members}) == 1
            self.lib.call("vinet_bn_bwd_multi", bs, n, self.stream())
            todo = [o for o in reversed(outs) if o[0] is not None]
            if len(todo) > 1:
                self.fork()
            for i, (conv_bwd, dy_ptr, lddy) in enumerate(todo):      # independent convolutions: alternate the two streams
                self.on_side(i % 2 == 1)
                conv_bwd(dy_ptr, lddy)
            self.join()
        self.tape.append(backward)

    def _bn_tail(self, name_bn, bn, raw, out, conv_bwd, dy_slot, stats=None):
        """BatchNorm + ReLU of the raw conv output `raw` into `out`, and its backward.  The backward either
        hands dY to `conv_bwd` (own conv) or writes it into `dy_slot` = (ptr, ld) of a fused group's dY buffer.
        stats = (device pointer, sum-of-squares stride): batch statistics already accumulated by the convolution epilogue."""
        Cn, rows = out.C, out.rows
        if self._bn_pending is not None and self.training and rows * Cn * raw.buf.element_size() < (24 << 20):
            self._bn_pending.append((name_bn, bn, raw, out, conv_bwd, dy_slot, stats))      # small layer: wait for its launch mates
            return
        if stats is not None:          # large layer, statistics from the conv epilogue: ONE finalise + apply launch
            ((st, ss),) = self._bn_fwd_multi([(name_bn, bn, raw, out, conv_bwd, dy_slot, stats)])
        else:
            out.xform, out.scale, out.shift = L.XF_IDENT, None, None
            ap = L.BnApply()
            ap.y, ap.ldy, ap.dtype, ap.rows, ap.C, ap.relu = raw.ptr(), raw.ld, self.dt, rows, Cn, 1
            ap.out, ap.ldo, ap.out_dtype = out.ptr(), out.ld, self.dt
            st, ss = self._bn_finalize(name_bn, bn, rows, Cn, raw, apply=ap)
        if self.record:
            self._bn_tape(name_bn, bn, raw, out, conv_bwd, dy_slot, st, ss)

    def _bn_tape(self, name_bn, bn, raw, out, conv_bwd, dy_slot, st, ss):
        """Backward of one (large) BatchNorm + ReLU layer: reduce + apply (one cooperative launch above 24 MB), then the conv's."""
        Cn, rows = out.C, out.rows
        training = self.training

        def backward():
            bsums = self.buf(name_bn + ".bsums", (2 * Cn + 4,), torch.float64, zero=True)        # [2][C] sums + ticket / flag / departures
            dgamma, dbeta = self.grad_tensor(name_bn + ".weight", bn.weight), self.grad_tensor(name_bn + ".bias", bn.bias)
            if dy_slot is None:
                dy = self.buf("dy.%d" % (rows * Cn), (rows, Cn), self.tdtype)
                dy_ptr, lddy = dy.data_ptr(), Cn
            else:
                dy_ptr, lddy = dy_slot
            b = L.BnBwd()
            b.g, b.ldg, b.y, b.ldy, b.dtype, b.rows, b.C, b.relu = out.gptr(), out.ldg, raw.ptr(), raw.ld, self.dt, rows, Cn, 1
            b.scale, b.shift, b.mean, b.invstd = ss[0].data_ptr(), ss[1].data_ptr(), st[0].data_ptr(), st[1].data_ptr()
            b.gamma, b.sums, b.dgamma, b.dbeta = bn.weight.data_ptr(), bsums.data_ptr(), dgamma.data_ptr(), dbeta.data_ptr()
            b.dy, b.lddy, b.dy_dtype, b.training = dy_ptr, lddy, self.dt, 1 if training else 0
            b.g_dtype = out.gdt
            if "vinet_bn_bwd_fused" in self.lib.fn and rows * Cn * raw.buf.element_size() >= (24 << 20):   # reduce + apply in one cooperative launch
                self.call("vinet_bn_bwd_fused", b)
            else:
                self.call("vinet_bn_bwd_reduce", b)
                self.call("vinet_bn_bwd_apply", b)
            self.param_grads[name_bn + ".weight"] = dgamma
            self.param_grads[name_bn + ".bias"] = dbeta
            if conv_bwd is not None:
                conv_bwd(dy_ptr, lddy)
        self.tape.append(backward)

    def conv_bn_group(self, gname, x, members):
        """Horizontally fused 1x1x1 BasicConv3d's that read the same input (Mixed_* branch0 / branch1.0 /
        branch2.0, model_utils.py:182-185): ONE GEMM with the concatenated weights computes all raw outputs
        (x is read once instead of three times), each member keeps its own BatchNorm; backward runs ONE
        weight-gradient and ONE data-gradient GEMM over the concatenated dY.
        members: [(name_conv, name_bn, conv_weight, bn, out_act)]."""
        if not self.training and not self.record and not self.split:      # inference: per-member convs with folded BatchNorm
            for nc, nb, w, bn, out in members:
                self.conv_bn(nc, nb, [x], w, bn, _G1, out)
            return
        couts = [w.shape[0] for _, _, w, _, _ in members]
        tot = sum(couts)
        key = gname + ".wcat"
        vers = tuple(_wstamp(w) for _, _, w, _, _ in members)
        hit = self.wcache.get(key)
        if hit is None or None in vers or hit[0] != vers or self._repack_all or any(a is not b[2] for a, b in zip(hit[2], members)):
            with torch.no_grad():
                if hit is not None and all(a is b[2] for a, b in zip(hit[2], members)):
                    wcat = torch.cat([w.detach() for _, _, w, _, _ in members], 0, out=hit[1])     # same storage: packs point at it
                else:
                    wcat = torch.cat([w.detach() for _, _, w, _, _ in members], 0).contiguous()
            self.wcache[key] = (vers, wcat, [w for _, _, w, _, _ in members])
        wcat = self.wcache[key][1]
        rawcat = Act(self.buf(gname + ".raw", (x.B, x.T, x.H, x.W, tot), self.tdtype), x.B, x.T, x.H, x.W, tot)
        gst = self.stats_slice(tot) if (self.epi_stats and self.tma_ok([x], _G1) and self.epi_stats_1x1) else None
        conv_bwd = self.conv(gname, [x], wcat, _G1, rawcat, stats=gst,
                             wgrad_split=[(nc + ".weight", c) for (nc, _, _, _, _), c in zip(members, couts)])
        dycat = None
        if self.record:
            dycat = self.buf(gname + ".dy", (rawcat.rows, tot), self.tdtype)

            def group_backward():
                conv_bwd(dycat.data_ptr(), tot)      # unpacks the weight gradient straight into the members' tensors
            self.tape.append(group_backward)      # appended first => runs after every member's BatchNorm backward
        off = 0
        for (nc, nb, w, bn, out), c in zip(members, couts):
            slot = None if dycat is None else (dycat.data_ptr() + off * dycat.element_size(), tot)
            self._bn_tail(nb, bn, rawcat.slice(off, c), out, None, slot, stats=None if gst is None else (gst + 8 * off, tot))
            off += c

    # ------------------------------------------------------------------ max pooling
    def maxpool(self, name, a, k, s, p):
        To, Ho, Wo = ConvGeom(k, s, p).out_dims(a.T, a.H, a.W)
        out = self.new_act(name, a.B, To, Ho, Wo, a.C)
        d = L.Pool()
        d.x, d.ldx, d.dtype, d.scale, d.shift, d.xform = a.ptr(), a.ld, self.dt, _ptr(a.scale), _ptr(a.shift), a.xform
        d.B, d.Ti, d.Hi, d.Wi, d.C = a.B, a.T, a.H, a.W, a.C
        (d.kt, d.kh, d.kw), (d.st, d.sh, d.sw), (d.pt, d.ph, d.pw) = k, s, p
        d.To, d.Ho, d.Wo, d.out, d.ldo, d.out_dtype = To, Ho, Wo, out.ptr(), out.ld, self.dt
        if self.record and a.needs_grad:       # remember the winning tap per output element for the backward scatter
            d.idx = self.buf(name + ".idx", (a.B, To, Ho, Wo, a.C), torch.uint8).data_ptr()
        self.call("vinet_maxpool_fwd", d)
        if self.record and a.needs_grad:
            def backward():
                # first writer of a's gradient: the gather kernel stores every element; otherwise it accumulates
                key = a.grad.data_ptr()
                d.gin_overwrite = 1 if (key not in self.gwritten and a.gchoff == 0 and a.C == a.ldg) else 0
                if d.gin_overwrite:
                    self.gwritten.add(key)
                else:
                    self.ensure_init(a)
                d.gout, d.ldgo, d.gin, d.ldgi = out.gptr(), out.ldg, a.gptr(), a.ldg
                d.gout_dtype, d.gin_dtype = out.gdt, a.gdt
                self.call("vinet_maxpool_bwd", d)
            self.tape.append(backward)
        return out

    # ------------------------------------------------------------------ decoder: conv -> ReLU -> 2x bilinear
    def conv_relu_up(self, name, srcs, w, geom):
        """Conv3d -> ReLU -> nn.Upsample((1,2,2), 'trilinear') (model.py:256-275).  Returns a VIRTUAL activation: the raw conv output
        z at low resolution, tagged XF_RELU | XF_UP2 with hi-res logical extents.  The next convolution reads it THROUGH the ReLU +
        interpolation in its own input stage (interpolating producer warps of the tcgen05 kernels, the FFMA gather, the operand
        split of the parity mode); only consumers without such a stage make `materialize_up2` write the hi-res tensor.  The
        gradient buffer is hi-res (the consumer's data gradient lands there); vinet_upsample_bwd folds it back onto z."""
        a0 = srcs[0]
        To, Ho, Wo = geom.out_dims(sum(s.T for s in srcs), a0.H, a0.W)
        Cout = w.shape[0]
        z = Act(self.buf(name + ".z", (a0.B, To, Ho, Wo, Cout), self.tdtype), a0.B, To, Ho, Wo, Cout, 0, L.XF_RELU)
        conv_bwd = self.conv(name, srcs, w, geom, z)
        u = Act(z.buf, a0.B, To, 2 * Ho, 2 * Wo, Cout, 0, L.XF_RELU | L.XF_UP2)
        u.up2, u.name = True, name
        if self.record:
            self.want_grad(u, name + ".up")

            def backward():
                d = L.Upsample()
                d.z, d.ldz, d.dtype, d.relu, d.B, d.T, d.h, d.w, d.C = z.ptr(), z.ld, self.dt, 1, a0.B, To, Ho, Wo, Cout
                dz = self.buf("dy.%d" % (z.rows * Cout), (z.rows, Cout), self.tdtype)
                d.gu, d.ldgu, d.dz, d.lddz, d.dz_dtype = u.gptr(), u.ldg, dz.data_ptr(), Cout, self.dt
                d.gu_dtype = u.gdt
                self.call("vinet_upsample_bwd", d)
                conv_bwd(dz.data_ptr(), Cout)
            self.tape.append(backward)
        return u

    # ------------------------------------------------------------------ decoder tail
    def conv_relu(self, name, srcs, w, geom, bias=None):
        """Conv3d (+bias) whose ReLU is applied by the consumer (the head); returns (act, conv backward)."""
        a0 = srcs[0]
        To, Ho, Wo = geom.out_dims(sum(s.T for s in srcs), a0.H, a0.W)
        Cout = w.shape[0]
        h = Act(self.buf(name + ".h", (a0.B, To, Ho, Wo, Cout), self.tdtype), a0.B, To, Ho, Wo, Cout, 0, L.XF_RELU)
        return h, self.conv(name, srcs, w, geom, h, bias=bias)

    def conv_relu_act(self, name, srcs, w, geom):
        """Conv3d (no bias) -> ReLU as a layer of its own (the decoder tail once the (kt,1,1) convolution has been commuted in front of
        the last up-sampling, see model.decoder_plan): the ReLU runs in the convolution's epilogue, so the next convolution reads a
        plain activation (TMA-fetchable); the parity mode sums raw terms and leaves the ReLU pending instead.  Backward masks the
        incoming gradient with (value > 0) - the same mask for the raw and the rectified tensor - and runs the conv backward."""
        a0 = srcs[0]
        To, Ho, Wo = geom.out_dims(sum(s.T for s in srcs), a0.H, a0.W)
        Cout = w.shape[0]
        fuse = not self.split
        a = self.new_act(name + ".a", a0.B, To, Ho, Wo, Cout, xform=L.XF_IDENT if fuse else L.XF_RELU)
        conv_bwd = self.conv(name, srcs, w, geom, a, ep=(None, None, L.ACT_RELU) if fuse else None)
        if self.record:
            def backward():
                dz = self.buf("dy.%d" % (a.rows * Cout), (a.rows, Cout), self.tdtype)
                self.lib.call("vinet_relu_bwd", a.gptr(), a.ldg, a.gdt, a.ptr(), a.ld, self.dt, a.rows, Cout, dz.data_ptr(), Cout,
                              self.dt, self.stream())
                conv_bwd(dz.data_ptr(), Cout)
            self.tape.append(backward)
        return a

    def head(self, name, a, w, b, conv_bwd=None, up2=None):
        """relu? -> Conv3d(C,1,1) + bias -> Sigmoid -> (B,H,W) fp32 (model.py:282-283 + the final view).
        up2 = "pre" | "post": `a` is a LOW-RES raw conv output and the decoder's last 2x up-sampling happens in the head's input
        stage (the hi-res 32-channel tensor - the largest activation of the decoder - is never written).  "pre": relu -> up ->
        head (T = 8 / 16 decoders, model.py:340, 402); "post": up -> relu -> head (T = 32 / 48 with the (kt,1,1) conv commuted)."""
        assert a.T == 1, "the decoder collapses time to one frame before the head"
        H, W = (2 * a.H, 2 * a.W) if up2 else (a.H, a.W)
        rows = a.B * H * W
        out = torch.empty((a.B, H, W), dtype=torch.float32, device=self.device)
        d = L.Head()
        relu = 1 if a.xform == L.XF_RELU else 0
        assert a.xform in (L.XF_RELU, L.XF_IDENT) and up2 in (None, "pre", "post")
        d.x, d.ldx, d.dtype, d.rows, d.C = a.ptr(), a.ld, self.dt, rows, a.C
        d.relu = relu if up2 != "pre" else 0
        if up2:
            assert conv_bwd is not None or not self.record
            d.up2, d.up_h, d.up_w, d.relu_pre = 1, a.H, a.W, (relu if up2 == "pre" else 0)
        d.w, d.b, d.out = w.data_ptr(), b.data_ptr(), out.data_ptr()
        self.call("vinet_head_fwd", d)
        if self.record:
            def backward(gout, _keep=out):       # the descriptor holds a raw pointer to `out` (sigmoid output): keep it alive
                gw, gb = self.grad_tensor(name + ".weight", w, zero=True), self.grad_tensor(name + ".bias", b, zero=True)
                d.gout, d.dw, d.db = gout.data_ptr(), gw.data_ptr(), gb.data_ptr()
                if conv_bwd is not None:        # input is a raw conv output: dx is that conv's dY
                    dx = self.buf("dy.%d" % (rows * a.C), (rows, a.C), self.tdtype)
                    d.dx, d.lddx, d.dx_dtype = dx.data_ptr(), a.C, self.dt
                else:                           # input is a materialised activation: dx is its fp32 gradient
                    assert self.first_write(a), "the head must be the only consumer of its input"
                    d.dx, d.lddx, d.dx_dtype = a.gptr(), a.ldg, a.gdt
                self.call("vinet_head_bwd", d)
                self.param_grads[name + ".weight"] = gw
                self.param_grads[name + ".bias"] = gb
                if conv_bwd is not None and up2:   # dx is the gradient at hi-res: fold it through the transposed interpolation
                    u = L.Upsample()
                    u.z, u.ldz, u.dtype, u.relu, u.B, u.T, u.h, u.w, u.C = a.ptr(), a.ld, self.dt, d.relu_pre, a.B, 1, a.H, a.W, a.C
                    dz = self.buf("dy.%d" % (a.rows * a.C), (a.rows, a.C), self.tdtype)
                    u.gu, u.ldgu, u.gu_dtype, u.dz, u.lddz, u.dz_dtype = dx.data_ptr(), a.C, self.dt, dz.data_ptr(), a.C, self.dt
                    self.call("vinet_upsample_bwd", u)
                    conv_bwd(dz.data_ptr(), a.C)
                elif conv_bwd is not None:
                    conv_bwd(dx.data_ptr(), a.C)
            self.head_backward = backward
        return out

    def mark_backward_point(self, name):
        """Tape marker: when the backward pass reaches it, every gradient of the layers planned AFTER this call has been written
        (the queued weight-gradient unpacks are flushed first); the model's call-back may start communicating them."""
        if not self.record:
            return

        def reached():
            if self.backward_point_cb is not None:
                self.unpack_flush()
                self.backward_point_cb(name)
        self.tape.append(reached)

    # ------------------------------------------------------------------ backward driver
    def backward(self, gout):
        """gout: fp32 (B,H,W) gradient w.r.t. the saliency map. Fills self.param_grads."""
        self.gwritten = set()
        self.unpack_queue = []
        self.arena_bypassed = False
        self.dwp_begin()
        self.head_backward(gout)
        for fn in reversed(self.tape):
            fn()
        self.unpack_flush()
        self.tape = []
        self.consumed_gen = getattr(self, "generation", -1)
        grads, self.param_grads = self.param_grads, {}     # hand over the only references: autograd can adopt the tensors
        return grads
