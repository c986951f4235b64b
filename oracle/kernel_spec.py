"""TEST INFRASTRUCTURE ONLY — executable specification of every kernel behind include/vinet_b200.h.

A slow numpy re-statement of what each C-ABI entry point must compute, operating on HOST memory
through the very same ctypes descriptors (fp32 storage only).  Two uses, both in ``tests/``:
  * ``Spec`` can be injected into ``vinet_b200.engine.Engine(backend=Spec())`` so the whole host-side
    plan (descriptor construction, channel-slice/ld arithmetic, tape order, backward formulas) is
    checked on the CPU against the PyTorch oracle without a GPU;
  * GPU tests run the same descriptor through the CUDA kernel and through this spec and compare.
It is never imported by the product package.
"""
import ctypes as C

import numpy as np

from vinet_b200 import lib as L

EPS = np.float32(2.2204e-16)


def _arr(ptr, n, dtype=np.float32):
    if not ptr or n == 0:
        return None
    ct = {np.float32: C.c_float, np.float64: C.c_double, np.int32: C.c_int32, np.uint8: C.c_uint8, np.uint16: C.c_uint16}[dtype]
    return np.ctypeslib.as_array((ct * int(n)).from_address(int(ptr)))


def _rows_view(ptr, rows, ld, c, dtype=L.F32):
    """[rows, c] strided view of a channel slice whose row stride is ld (bf16 storage: a widened fp32 COPY, read-only use)."""
    if dtype == L.BF16:
        a = _arr(ptr, (rows - 1) * ld + c, np.uint16)
        v = np.lib.stride_tricks.as_strided(a, shape=(rows, c), strides=(ld * 2, 2))
        return (v.astype(np.uint32) << 16).view(np.float32)
    a = _arr(ptr, (rows - 1) * ld + c)
    return np.lib.stride_tricks.as_strided(a, shape=(rows, c), strides=(ld * 4, 4))


def _bf16_bits(x):
    """fp32 array -> bf16 bit patterns (round to nearest even), as uint16."""
    u = np.ascontiguousarray(x, np.float32).view(np.uint32).astype(np.uint64)
    return (((u + 0x7FFF + ((u >> 16) & 1)) >> 16) & 0xFFFF).astype(np.uint16)


def _bf16_round(x):
    return (_bf16_bits(x).astype(np.uint32) << 16).view(np.float32)


def _xform(v, xf, scale, shift, c0, c):
    if xf & 2:
        v = v * _arr(scale, c0 + c)[c0:] + _arr(shift, c0 + c)[c0:]
    if xf & 1:
        v = np.maximum(v, 0)
    return v.astype(np.float32)


def _up_axis(Y, n):
    """Taps of hi-res coordinates Y on a low-res axis of length n: nn.Upsample(scale_factor=2, align_corners=False) =
    (i0, i1, weight of i1), in the kernel's float32 arithmetic (model.py:254)."""
    src = np.maximum((Y.astype(np.float32) + np.float32(0.5)) * np.float32(0.5) - np.float32(0.5), np.float32(0))
    i0 = src.astype(np.int64)
    return i0, np.minimum(i0 + 1, n - 1), (src - i0.astype(np.float32)).astype(np.float32)


def _up2_rows(low, n, Y, X, relu):
    """low: [N, h, w, C] fp32; returns relu? -> 2x bilinear values at frames n, hi-res pixels (Y, X): [len, C] fp32, blended in the
    order of the CUDA implementation (up2.cuh: x first, then y)."""
    h, w = low.shape[1], low.shape[2]
    y0, y1, ly = _up_axis(Y, h)
    x0, x1, lx = _up_axis(X, w)
    f = (lambda v: np.maximum(v, 0)) if relu else (lambda v: v)
    a, b, e0, e1 = f(low[n, y0, x0]), f(low[n, y0, x1]), f(low[n, y1, x0]), f(low[n, y1, x1])
    lx, ly = lx[:, None], ly[:, None]
    one = np.float32(1)
    top = (one - lx) * a + lx * b
    bot = (one - lx) * e0 + lx * e1
    return ((one - ly) * top + ly * bot).astype(np.float32)


def _taps(g):
    return [(g.tap[i][0], g.tap[i][1], g.tap[i][2]) for i in range(g.ntaps)]


def _row_coords(g):
    rows = g.B * g.Tr * g.Hr * g.Wr
    r = np.arange(rows)
    w = r % g.Wr; r //= g.Wr
    h = r % g.Hr; r //= g.Hr
    tr = r % g.Tr; b = r // g.Tr
    return b, tr * g.row_tstep + g.row_toff, h, w


def gather_matrix(g):
    """A[rows, ntaps*Cs] of the implicit GEMM described by a vinet_gather_t."""
    b, t, h, w = _row_coords(g)
    rows = b.shape[0]
    A = np.zeros((rows, g.ntaps * g.Cs), np.float32)
    T0 = g.src[0].T
    for ti, (dt, dh, dw) in enumerate(_taps(g)):
        if g.mode == L.GATHER_FPROP:
            ts, hs, ws = t * g.st - g.pt + dt, h * g.sh - g.ph + dh, w * g.sw - g.pw + dw
            ok = np.ones(rows, bool)
        else:
            nt, nh, nw = t + g.pt - dt, h + g.ph - dh, w + g.pw - dw
            ok = (nt >= 0) & (nh >= 0) & (nw >= 0) & (nt % g.st == 0) & (nh % g.sh == 0) & (nw % g.sw == 0)
            ts, hs, ws = nt // g.st, nh // g.sh, nw // g.sw
        ok &= (ts >= 0) & (ts < g.Ts) & (hs >= 0) & (hs < g.Hs) & (ws >= 0) & (ws < g.Ws)
        for si in (0, 1):
            s = g.src[si]
            if not s.ptr or s.T == 0:
                continue
            sel = ok & ((ts < T0) if si == 0 else (ts >= T0))
            if not sel.any():
                continue
            tl = ts[sel] - (T0 if si else 0)
            if s.xform & L.XF_UP2:      # low-res source [B,T,Hs/2,Ws/2,ld] read through relu? + the 2x bilinear up-sampling
                assert si == 0 and g.mode == L.GATHER_FPROP and not (s.xform & 2) and g.Hs % 2 == 0 and g.Ws % 2 == 0
                lh, lw = g.Hs // 2, g.Ws // 2
                low = _rows_view(s.ptr, g.B * s.T * lh * lw, s.ld, g.Cs, g.dtype).reshape(g.B * s.T, lh, lw, g.Cs)
                v = _up2_rows(low, b[sel] * s.T + tl, hs[sel], ws[sel], bool(s.xform & 1))
                A[np.nonzero(sel)[0], ti * g.Cs:(ti + 1) * g.Cs] = _bf16_round(v) if g.dtype == L.BF16 else v
                continue
            pos = ((b[sel] * s.T + tl) * g.Hs + hs[sel]) * g.Ws + ws[sel]
            src = _rows_view(s.ptr, g.B * s.T * g.Hs * g.Ws, s.ld, g.Cs, g.dtype)
            A[np.nonzero(sel)[0], ti * g.Cs:(ti + 1) * g.Cs] = _xform(src[pos], s.xform, s.scale, s.shift, 0, g.Cs)
    return A


class Spec:
    """Drop-in for vinet_b200.lib.Library on host memory."""

    def __init__(self):
        # names the engine probes with `in lib.fn` to pick its fused / multi-layer BatchNorm paths: the spec implements them by
        # composition of the single-layer specs, so the CPU plan tests walk the same host logic (batching, tape order)
        self.fn = {"vinet_packed_weight_bytes": self._packed_bytes, "vinet_bn_stats_finalize": None, "vinet_bn_fwd_fused": None,
                   "vinet_bn_bwd_fused": None, "vinet_bn_stats_finalize_multi": None, "vinet_bn_apply_multi": None,
                   "vinet_bn_bwd_multi": None, "vinet_conv_up2_fused": self._up2_fused, "vinet_relu_bwd": None}
        self.launches = 0

    def launch_count(self):
        return self.launches

    @staticmethod
    def _up2_fused(g, n, engine, kernel):
        g = g._obj if hasattr(g, "_obj") else g
        return 1 if (g.mode == L.GATHER_FPROP and (g.src[0].xform & ~1) == L.XF_UP2 and g.Hs % 2 == 0 and g.Ws % 2 == 0) else 0

    @staticmethod
    def _packed_bytes(engine, n, block_n, n_tiles, k_blocks):
        return k_blocks * 64 * (-(-n // 64) * 64) * 4

    def call(self, name, *args):
        args = [a._obj if hasattr(a, "_obj") else a for a in args]
        self.launches += 1
        getattr(self, name[len("vinet_"):])(*args)

    # ------------------------------------------------------------------ misc
    def memset_async(self, ptr, value, nbytes, stream):
        C.memset(int(ptr), value, int(nbytes))

    def axpy_f32(self, dst, src, n, acc, stream):
        d, s = _arr(dst, n), _arr(src, n)
        d[:] = d + s if acc else s

    def colsum(self, x, ld, dtype, rows, c, ws, out, stream):
        _arr(out, c)[:] = _rows_view(x, rows, ld, c).astype(np.float64).sum(0)

    # ------------------------------------------------------------------ packing
    def pack_input(self, d, stream):
        n = d.B * d.T * d.H * d.W
        out = _arr(d.out, n * d.cpad).reshape(d.B, d.T, d.H, d.W, d.cpad)
        out[:] = 0
        span = (d.B - 1) * d.sb + (d.C - 1) * d.sc + (d.T - 1) * d.st + (d.H - 1) * d.sh + (d.W - 1) * d.sw + 1
        x = np.lib.stride_tricks.as_strided(_arr(d.x, span), shape=(d.B, d.C, d.T, d.H, d.W),
                                            strides=tuple(4 * s for s in (d.sb, d.sc, d.st, d.sh, d.sw)))
        out[..., :d.C] = np.transpose(x, (0, 2, 3, 4, 1))

    def pack_weights(self, d, stream):
        # the spec keeps ONE packed format (fp32 [k][npad]) for both engines; a TC-engine pack holds term `part` of the
        # weight's bf16 expansion (split-precision parity mode), exactly what the tcgen05 kernels multiply with
        assert d.layout == L.KLAYOUT_DENSE and (d.part == 0 or d.engine == L.ENGINE_TC)
        ldc, kk = (d.ld_cin or d.Cin), d.kt * d.kh * d.kw
        flat = _arr(d.w, ((d.Cout - 1) * ldc + d.Cin) * kk)           # Cin may name a channel slice of a wider tensor
        w = np.lib.stride_tricks.as_strided(flat, shape=(d.Cout, d.Cin, d.kt, d.kh, d.kw),
                                            strides=(ldc * kk * 4, kk * 4, d.kh * d.kw * 4, d.kw * 4, 4))
        if d.engine == L.ENGINE_TC:
            w = w.copy()
            for _ in range(d.part):
                w = w - _bf16_round(w)
            w = _bf16_round(w)
        n = d.Cout if d.mode == L.GATHER_FPROP else d.Cin
        npad = -(-n // 64) * 64
        out = _arr(d.out, d.k_blocks * 64 * npad).reshape(d.k_blocks * 64, npad)
        out[:] = 0
        for ti in range(d.ntaps):
            dt, dh, dw = d.tap[ti][0], d.tap[ti][1], d.tap[ti][2]
            if d.mode == L.GATHER_FPROP:
                out[ti * d.cs:ti * d.cs + d.Cin, :d.Cout] = w[:, :, dt, dh, dw].T
            else:
                out[ti * d.cs:ti * d.cs + d.Cout, :d.Cin] = w[:, :, dt, dh, dw]

    def split_bf16(self, d, stream):
        if d.xform & L.XF_UP2:
            n = d.rows // (4 * d.up_h * d.up_w)
            low = _rows_view(d.x, n * d.up_h * d.up_w, d.ld, d.C, d.dtype).reshape(n, d.up_h, d.up_w, d.C)
            r = np.arange(d.rows)
            v = _up2_rows(low, r // (4 * d.up_h * d.up_w), (r // (2 * d.up_w)) % (2 * d.up_h), r % (2 * d.up_w), bool(d.xform & 1))
        else:
            v = _xform(_rows_view(d.x, d.rows, d.ld, d.C, d.dtype), d.xform, d.scale, d.shift, 0, d.C)
        for p in range(d.nparts):
            bits = _bf16_bits(v)
            out = _arr(d.part[p], (d.rows - 1) * d.ldo + d.C, np.uint16)
            np.lib.stride_tricks.as_strided(out, shape=(d.rows, d.C), strides=(d.ldo * 2, 2))[:] = bits
            v = v - (bits.astype(np.uint32) << 16).view(np.float32)

    def unpack_wgrad(self, dwp, lddw, cs, grad, cout, cin, ntaps, stream):
        rows = -(-(ntaps * cs) // 128) * 128
        p = _arr(dwp, rows * lddw).reshape(rows, lddw)
        g = _arr(grad, cout * cin * ntaps).reshape(cout, cin, ntaps)
        for t in range(ntaps):
            g[:, :, t] = p[t * cs:t * cs + cin, :cout].T

    # ------------------------------------------------------------------ convolution
    def conv_gemm(self, d, engine, stream):
        A = gather_matrix(d.g)
        npad = -(-d.N // 64) * 64
        W = _arr(d.w, d.k_blocks * 64 * npad).reshape(d.k_blocks * 64, npad)
        acc = A @ W[:A.shape[1], :d.N]
        if d.ep_scale:
            acc = acc * _arr(d.ep_scale, d.N)
        if d.ep_shift:
            acc = acc + _arr(d.ep_shift, d.N)
        if d.ep_act == L.ACT_RELU:
            acc = np.maximum(acc, 0)
        elif d.ep_act == L.ACT_SIGMOID:
            acc = 1 / (1 + np.exp(-acc))
        assert d.out_dtype == L.F32
        g = d.g
        b, t, h, w = _row_coords(g)
        for i in (0, 1):
            if not d.out[i]:
                continue
            sel = (t < d.out_T[0]) if i == 0 else (t >= d.out_T[0])
            if not sel.any():
                continue
            tl = t[sel] - (d.out_T[0] if i else 0)
            pos = ((b[sel] * d.out_T[i] + tl) * g.Hr + h[sel]) * g.Wr + w[sel]
            out = _rows_view(d.out[i], g.B * d.out_T[i] * g.Hr * g.Wr, d.ldo[i], d.N)
            out[pos] = (out[pos] + acc[sel]) if (d.accumulate >> i) & 1 else acc[sel]

    def conv_wgrad(self, d, engine, stream):
        A = gather_matrix(d.g)
        rows = A.shape[0]
        dy = _rows_view(d.dy, rows, d.lddy, d.N, d.dy_dtype)
        k = A.shape[1]
        kp = -(-k // 128) * 128
        dwp = _arr(d.dwp, kp * d.lddw).reshape(kp, d.lddw)
        dwp[:k, :d.N] += A.T @ dy

    # ------------------------------------------------------------------ batch norm
    def bn_stats(self, d, stream):
        y = _rows_view(d.y, d.rows, d.ld, d.C).astype(np.float64)
        s = _arr(d.sums, 2 * d.C, np.float64)
        s[:d.C] += y.sum(0)
        s[d.C:] += (y * y).sum(0)

    def bn_finalize(self, d, stream):
        c = d.C
        if d.training:
            s = _arr(d.sums, 2 * c, np.float64)
            mean = s[:c] / d.rows
            var = np.maximum(s[c:] / d.rows - mean * mean, 0)
            s[:] = 0                                   # consumed and cleared
            invstd = 1 / np.sqrt(var + d.eps)
            if d.running_mean:
                unb = var * d.rows / (d.rows - 1) if d.rows > 1 else var
                rm, rv = _arr(d.running_mean, c), _arr(d.running_var, c)
                rm[:] = (1 - d.momentum) * rm + d.momentum * mean
                rv[:] = (1 - d.momentum) * rv + d.momentum * unb
        else:
            mean = _arr(d.running_mean, c).astype(np.float64)
            invstd = 1 / np.sqrt(_arr(d.running_var, c).astype(np.float64) + d.eps)
        sc = _arr(d.gamma, c) * invstd
        _arr(d.scale, c)[:] = sc
        _arr(d.shift, c)[:] = _arr(d.beta, c) - mean * sc
        _arr(d.mean, c)[:] = mean
        _arr(d.invstd, c)[:] = invstd

    def bn_stats_finalize(self, d, f, stream):
        self.bn_stats(d, stream)
        self.bn_finalize(f, stream)

    def bn_fwd_fused(self, d, f, a, stream):
        self.bn_stats_finalize(d, f, stream)
        self.bn_apply(a, stream)

    def bn_bwd_fused(self, d, stream):
        self.bn_bwd_reduce(d, stream)
        self.bn_bwd_apply(d, stream)

    def bn_stats_finalize_multi(self, ds, fs, n, stream):
        for i in range(n):
            self.bn_stats_finalize(ds[i], fs[i], stream)

    def bn_apply_multi(self, aps, n, stream):
        for i in range(n):
            self.bn_apply(aps[i], stream)

    def bn_bwd_multi(self, bs, n, stream):
        for i in range(n):
            self.bn_bwd_fused(bs[i], stream)

    def bn_apply(self, d, stream):
        assert d.dtype == L.F32 and d.out_dtype == L.F32
        y = _rows_view(d.y, d.rows, d.ldy, d.C)
        v = y * _arr(d.scale, d.C) + _arr(d.shift, d.C)
        if d.relu:
            v = np.maximum(v, 0)
        _rows_view(d.out, d.rows, d.ldo, d.C)[:] = v.astype(np.float32)

    def _bn_bwd_terms(self, d):
        c = d.C
        y = _rows_view(d.y, d.rows, d.ldy, c)
        g = _rows_view(d.g, d.rows, d.ldg, c)
        sc, sh = _arr(d.scale, c), _arr(d.shift, c)
        yh = y * sc + sh
        gm = np.where((yh > 0) | (d.relu == 0), g, 0).astype(np.float32)
        yn = (y - _arr(d.mean, c)) * _arr(d.invstd, c)
        return gm, yn, sc

    def bn_bwd_reduce(self, d, stream):
        gm, yn, _ = self._bn_bwd_terms(d)
        _arr(d.dbeta, d.C)[:] = gm.astype(np.float64).sum(0)
        _arr(d.dgamma, d.C)[:] = (gm.astype(np.float64) * yn).sum(0)

    def bn_bwd_apply(self, d, stream):
        gm, yn, sc = self._bn_bwd_terms(d)
        assert d.dy_dtype == L.F32
        dy = _rows_view(d.dy, d.rows, d.lddy, d.C)
        if d.training:
            dy[:] = sc * (gm - _arr(d.dbeta, d.C) / d.rows - yn * _arr(d.dgamma, d.C) / d.rows)
        else:
            dy[:] = sc * gm

    # ------------------------------------------------------------------ pooling
    def _pool_scan(self, d):
        n_in = d.B * d.Ti * d.Hi * d.Wi
        x = _xform(_rows_view(d.x, n_in, d.ldx, d.C), d.xform, d.scale, d.shift, 0, d.C)
        x = x.reshape(d.B, d.Ti, d.Hi, d.Wi, d.C)
        best = np.full((d.B, d.To, d.Ho, d.Wo, d.C), -np.inf, np.float32)
        arg = np.full(best.shape, -1, np.int64)
        bi = np.arange(d.B)[:, None, None, None]
        to, ho, wo = np.arange(d.To)[None, :, None, None], np.arange(d.Ho)[None, None, :, None], np.arange(d.Wo)[None, None, None, :]
        for dt in range(d.kt):
            for dh in range(d.kh):
                for dw in range(d.kw):
                    t, h, w = to * d.st - d.pt + dt, ho * d.sh - d.ph + dh, wo * d.sw - d.pw + dw
                    ok = (t >= 0) & (t < d.Ti) & (h >= 0) & (h < d.Hi) & (w >= 0) & (w < d.Wi)
                    ok = np.broadcast_to(ok, best.shape[:4])
                    tc, hc, wc = np.clip(t, 0, d.Ti - 1), np.clip(h, 0, d.Hi - 1), np.clip(w, 0, d.Wi - 1)
                    v = x[bi, tc, hc, wc]
                    pos = np.broadcast_to(((bi * d.Ti + tc) * d.Hi + hc) * d.Wi + wc, best.shape[:4])[..., None]
                    upd = ok[..., None] & (v > best)
                    best = np.where(upd, v, best)
                    arg = np.where(upd, pos, arg)
        return best, arg

    def maxpool_fwd(self, d, stream):
        best, _ = self._pool_scan(d)
        n_out = d.B * d.To * d.Ho * d.Wo
        _rows_view(d.out, n_out, d.ldo, d.C)[:] = best.reshape(n_out, d.C)

    def maxpool_bwd(self, d, stream):
        _, arg = self._pool_scan(d)
        n_out = d.B * d.To * d.Ho * d.Wo
        g = _rows_view(d.gout, n_out, d.ldgo, d.C)
        gin = _rows_view(d.gin, d.B * d.Ti * d.Hi * d.Wi, d.ldgi, d.C)
        if d.gin_overwrite:
            gin[:] = 0
        arg = arg.reshape(n_out, d.C)
        cc = np.broadcast_to(np.arange(d.C), arg.shape)
        ok = arg >= 0
        np.add.at(gin, (arg[ok], cc[ok]), g[ok])

    # ------------------------------------------------------------------ upsample
    @staticmethod
    def _up_matrix(n):
        """[2n, n] interpolation matrix of the 2x bilinear resize with align_corners=False."""
        m = np.zeros((2 * n, n), np.float32)
        for Y in range(2 * n):
            src = max((Y + 0.5) * 0.5 - 0.5, 0.0)
            i0 = int(src); i1 = min(i0 + 1, n - 1); l1 = src - i0
            m[Y, i0] += 1 - l1
            m[Y, i1] += l1
        return m

    def upsample_fwd(self, d, stream):
        z = _rows_view(d.z, d.B * d.T * d.h * d.w, d.ldz, d.C).reshape(d.B * d.T, d.h, d.w, d.C)
        if d.relu:
            z = np.maximum(z, 0)
        u = np.einsum("Yy,nyxc,Xx->nYXc", self._up_matrix(d.h), z, self._up_matrix(d.w))
        _rows_view(d.u, d.B * d.T * 4 * d.h * d.w, d.ldu, d.C)[:] = u.reshape(-1, d.C)

    def upsample_bwd(self, d, stream):
        gu = _rows_view(d.gu, d.B * d.T * 4 * d.h * d.w, d.ldgu, d.C).reshape(d.B * d.T, 2 * d.h, 2 * d.w, d.C)
        dz = np.einsum("Yy,nYXc,Xx->nyxc", self._up_matrix(d.h), gu, self._up_matrix(d.w)).reshape(-1, d.C)
        if d.relu:
            z = _rows_view(d.z, d.B * d.T * d.h * d.w, d.ldz, d.C)
            dz = np.where(z > 0, dz, 0)
        assert d.dz_dtype == L.F32
        _rows_view(d.dz, d.B * d.T * d.h * d.w, d.lddz, d.C)[:] = dz

    # ------------------------------------------------------------------ head
    def _head_x(self, d):
        if not d.up2:
            return _rows_view(d.x, d.rows, d.ldx, d.C)
        n = d.rows // (4 * d.up_h * d.up_w)
        low = _rows_view(d.x, n * d.up_h * d.up_w, d.ldx, d.C).reshape(n, d.up_h, d.up_w, d.C)
        r = np.arange(d.rows)
        return _up2_rows(low, r // (4 * d.up_h * d.up_w), (r // (2 * d.up_w)) % (2 * d.up_h), r % (2 * d.up_w), bool(d.relu_pre))

    def relu_bwd(self, g, ldg, g_dtype, z, ldz, z_dtype, rows, c, dz, lddz, dz_dtype, stream):
        assert g_dtype == L.F32 and z_dtype == L.F32 and dz_dtype == L.F32
        _rows_view(dz, rows, lddz, c)[:] = np.where(_rows_view(z, rows, ldz, c) > 0, _rows_view(g, rows, ldg, c), 0)

    def head_fwd(self, d, stream):
        x = self._head_x(d)
        if d.relu:
            x = np.maximum(x, 0)
        logit = x @ _arr(d.w, d.C) + (_arr(d.b, 1)[0] if d.b else 0)
        _arr(d.out, d.rows)[:] = 1 / (1 + np.exp(-logit))

    def head_bwd(self, d, stream):
        x = self._head_x(d)
        on = (x > 0) | (d.relu == 0)
        a = np.where(on, x, 0)
        o = _arr(d.out, d.rows)
        dl = _arr(d.gout, d.rows) * o * (1 - o)
        assert d.dx_dtype == L.F32
        _rows_view(d.dx, d.rows, d.lddx, d.C)[:] = np.where(on, dl[:, None] * _arr(d.w, d.C)[None, :], 0)
        _arr(d.dw, d.C)[:] += (dl[:, None].astype(np.float64) * a).sum(0)
        _arr(d.db, 1)[0] += dl.astype(np.float64).sum()

    # ------------------------------------------------------------------ losses (torch autograd of the oracle)
    def _loss(self, d, want_grad):
        import torch
        from . import torch_oracle as O
        s = torch.from_numpy(_arr(d.s, d.B * d.n).reshape(d.B, 1, d.n).copy()).requires_grad_(want_grad)
        g = torch.from_numpy(_arr(d.g, d.B * d.n).reshape(d.B, 1, d.n).copy())
        fn = {L.LOSS_KLDIV: O.kldiv, L.LOSS_CC: O.cc, L.LOSS_SIM: O.similarity, L.LOSS_NSS: O.nss}[d.kind]
        return s, fn(s, g)

    def loss_fwd(self, d, stream):
        _, v = self._loss(d, False)
        _arr(d.out, 1)[0] = v.item()

    def loss_bwd(self, d, stream):
        import torch
        with torch.enable_grad():
            s, v = self._loss(d, True)
            (gr,) = torch.autograd.grad(v, s)
        _arr(d.grad_s, d.B * d.n)[:] = gr.numpy().reshape(-1) * _arr(d.gout, 1)[0]


# ----------------------------------------------------------------------------- audio branch (AViNet)
def _t(ptr, shape):
    import torch
    n = int(np.prod(shape))
    return torch.from_numpy(_arr(ptr, n).reshape(shape))


def _conv1d_fwd(self, d, stream):
    import torch.nn.functional as F
    x, w = _t(d.x, (d.B, d.Cin, d.Lin)), _t(d.w, (d.Cout, d.Cin, d.k))
    b = _t(d.bias, (d.Cout,)) if d.bias else None
    _t(d.y, (d.B, d.Cout, d.Lout))[:] = F.conv1d(x, w, b, d.stride, d.pad)


def _conv1d_bwd(self, d, stream):
    import torch
    import torch.nn.functional as F
    x = _t(d.x, (d.B, d.Cin, d.Lin)).clone().requires_grad_(True)
    w = _t(d.w, (d.Cout, d.Cin, d.k)).clone().requires_grad_(True)
    b = _t(d.bias, (d.Cout,)).clone().requires_grad_(True)
    with torch.enable_grad():
        y = F.conv1d(x, w, b, d.stride, d.pad)
        gx, gw, gb = torch.autograd.grad(y, (x, w, b), _t(d.dy, (d.B, d.Cout, d.Lout)))
    if d.dx:
        _t(d.dx, (d.B, d.Cin, d.Lin))[:] = gx
    _t(d.dw, (d.Cout, d.Cin, d.k))[:] = gw
    if d.dbias:
        _t(d.dbias, (d.Cout,))[:] = gb


def _bn1d(d, y, train_stats):
    import torch
    import torch.nn.functional as F
    g, b = _t(d.gamma, (d.C,)), _t(d.beta, (d.C,))
    if d.training:
        mean, var = train_stats if train_stats else (y.mean((0, 2)), y.var((0, 2), unbiased=False))
    else:
        mean, var = _t(d.running_mean, (d.C,)), _t(d.running_var, (d.C,))
    invstd = 1 / torch.sqrt(var + d.eps)
    o = F.relu((y - mean[None, :, None]) * invstd[None, :, None] * g[None, :, None] + b[None, :, None])
    if d.pool > 1:
        o = F.max_pool1d(o, d.pool, d.pool)
    return o, mean, var, invstd


def _bn1d_fwd(self, d, stream):
    import torch
    y = _t(d.y, (d.B, d.C, d.L))
    o, mean, var, invstd = _bn1d(d, y, None)
    _t(d.out, (d.B, d.C, d.L // d.pool))[:] = o
    _t(d.mean, (d.C,))[:] = mean
    _t(d.invstd, (d.C,))[:] = invstd
    if d.training and d.running_mean:
        n = d.B * d.L
        rm, rv = _t(d.running_mean, (d.C,)), _t(d.running_var, (d.C,))
        rm[:] = (1 - d.momentum) * rm + d.momentum * mean
        rv[:] = (1 - d.momentum) * rv + d.momentum * var * n / (n - 1)


def _bn1d_bwd(self, d, stream):
    import torch
    import torch.nn.functional as F
    y = _t(d.y, (d.B, d.C, d.L)).clone().requires_grad_(True)
    g = _t(d.gamma, (d.C,)).clone().requires_grad_(True)
    b = _t(d.beta, (d.C,)).clone().requires_grad_(True)
    with torch.enable_grad():
        if d.training:
            o = F.relu(F.batch_norm(y, None, None, g, b, True, 0.0, d.eps))
        else:
            o = F.relu(F.batch_norm(y, _t(d.running_mean, (d.C,)), _t(d.running_var, (d.C,)), g, b, False, 0.0, d.eps))
        if d.pool > 1:
            o = F.max_pool1d(o, d.pool, d.pool)
        gy, gg, gb = torch.autograd.grad(o, (y, g, b), _t(d.gout, (d.B, d.C, d.L // d.pool)))
    _t(d.dy, (d.B, d.C, d.L))[:] = gy
    _t(d.dgamma, (d.C,))[:] = gg
    _t(d.dbeta, (d.C,))[:] = gb


def _av_inputs(d):
    import torch
    y0 = torch.from_numpy(_xform(_rows_view(d.y0, d.B * 4 * 7 * 12, d.ld, d.C), d.xform, d.scale, d.shift, 0, d.C))
    return y0.reshape(d.B, 4, 7, 12, d.C).permute(0, 4, 1, 2, 3).contiguous()


def _avfuse_fwd(self, d, stream):
    import torch
    import torch.nn.functional as F
    y0 = _av_inputs(d)
    v = F.max_pool3d(y0, (4, 1, 1), (2, 1, 2)).flatten(2)
    _t(d.vbuf, (d.B, d.C, 42))[:] = v
    o = F.bilinear(v, _t(d.audio, (d.B, d.C, 3)), _t(d.w, (336, 42, 3)), _t(d.bias, (336,)))     # (B,C,336)
    assert d.out_dtype == L.F32
    _rows_view(d.out, d.B * 336, d.ldo, d.C)[:] = o.permute(0, 2, 1).reshape(d.B * 336, d.C).numpy()


def _avfuse_bwd(self, d, stream):
    import torch
    import torch.nn.functional as F
    y0 = _av_inputs(d).requires_grad_(True)
    a = _t(d.audio, (d.B, d.C, 3)).clone().requires_grad_(True)
    w = _t(d.w, (336, 42, 3)).clone().requires_grad_(True)
    b = _t(d.bias, (336,)).clone().requires_grad_(True)
    go = torch.from_numpy(_rows_view(d.gout, d.B * 336, d.ldgo, d.C).copy()).reshape(d.B, 336, d.C).permute(0, 2, 1)
    with torch.enable_grad():
        o = F.bilinear(F.max_pool3d(y0, (4, 1, 1), (2, 1, 2)).flatten(2), a, w, b)
        gy, ga, gw, gb = torch.autograd.grad(o, (y0, a, w, b), go)
    gy0 = _rows_view(d.gy0, d.B * 4 * 7 * 12, d.ldgy0, d.C)
    gy0 += gy.permute(0, 2, 3, 4, 1).reshape(-1, d.C).numpy()
    _t(d.gaudio, (d.B, d.C, 3))[:] = ga
    _t(d.dw, (336, 42, 3))[:] = gw
    _t(d.dbias, (336,))[:] = gb


Spec.conv1d_fwd, Spec.conv1d_bwd = _conv1d_fwd, _conv1d_bwd
Spec.bn1d_fwd, Spec.bn1d_bwd = _bn1d_fwd, _bn1d_bwd
Spec.avfuse_fwd, Spec.avfuse_bwd = _avfuse_fwd, _avfuse_bwd


# ----------------------------------------------------------------------------- transformer fusion variants (csrc/xfmr.cu)
def _strided(ptr, dtype, n0, s0, n1, s1, off):
    """(n0, n1) matrix whose element (i, j) is flat[off + i*s0 + j*s1] (a widened COPY; strides may be 0)."""
    i = off + np.arange(n0, dtype=np.int64)[:, None] * s0 + np.arange(n1, dtype=np.int64)[None, :] * s1
    hi = int(i.max()) + 1
    if dtype == L.BF16:
        flat = (_arr(ptr, hi, np.uint16).astype(np.uint32) << 16).view(np.float32)
    else:
        flat = _arr(ptr, hi)
    return flat[i], i


def _bgemm(self, d, stream):
    for b1 in range(d.nb1):
        for b2 in range(d.nb2):
            acc = np.zeros((d.M, d.N), np.float64)
            if d.K > 0:
                A, _ = _strided(d.A, d.a_dtype, d.M, d.sAm, d.K, d.sAk, b1 * d.sAb1 + b2 * d.sAb2)
                if d.a_scale:
                    if d.a_xf_on_m:
                        A = A * _arr(d.a_scale, d.M)[:, None] + _arr(d.a_shift, d.M)[:, None]
                    else:
                        A = A * _arr(d.a_scale, d.K)[None, :] + _arr(d.a_shift, d.K)[None, :]
                if d.a_relu:
                    A = np.maximum(A, 0)
                Bm, _ = _strided(d.B, d.b_dtype, d.N, d.sBn, d.K, d.sBk, b1 * d.sBb1 + b2 * d.sBb2)
                acc = A.astype(np.float64) @ Bm.astype(np.float64).T
            v = np.float64(d.alpha) * acc
            if d.bias1:
                v = v + _strided(d.bias1, L.F32, d.M, d.s1m, d.N, d.s1n, 0)[0]
            if d.bias2:
                v = v + _strided(d.bias2, L.F32, d.M, d.s2m, d.N, d.s2n, b1 * d.s2b1 + b2 * d.s2b2)[0]
            if d.relu:
                v = np.maximum(v, 0)
            _, ci = _strided(d.C, L.F32, d.M, d.sCm, d.N, d.sCn, b1 * d.sCb1 + b2 * d.sCb2)
            assert d.c_dtype == L.F32, "the numpy spec keeps fp32 storage"
            flat = _arr(d.C, int(ci.max()) + 1)
            flat[ci] = (flat[ci] + v if (d.accumulate & 1) else v).astype(np.float32)


def _softmax_fwd(self, s, rows, n, stream):
    a = _arr(s, rows * n).reshape(rows, n)
    e = np.exp(a - a.max(1, keepdims=True))
    a[:] = e / e.sum(1, keepdims=True)


def _softmax_bwd(self, p, dp, rows, n, stream):
    P, G = _arr(p, rows * n).reshape(rows, n), _arr(dp, rows * n).reshape(rows, n)
    G[:] = P * (G - (G * P).sum(1, keepdims=True))


def _mix32(x):
    x = x.astype(np.uint64)
    m = np.uint64(0xFFFFFFFF)
    x ^= x >> np.uint64(16); x = (x * np.uint64(0x7FEB352D)) & m
    x ^= x >> np.uint64(15); x = (x * np.uint64(0x846CA68B)) & m
    x ^= x >> np.uint64(16)
    return x


def dropout_keep(n, p, seed, counter, salt):
    """The keep mask of csrc/xfmr.cu drop_keep for elements 0..n-1 (n < 2^32)."""
    m = 0xFFFFFFFF
    key = ((seed & m) ^ ((counter * 0x85EBCA6B) & m) ^ ((salt * 0xC2B2AE35) & m)) & m
    rot = ((key << 7) | (key >> 25)) & m
    i = np.arange(n, dtype=np.uint64)
    h = _mix32((_mix32(i ^ np.uint64(key)) + np.uint64(rot)) & np.uint64(m))
    u = (h >> np.uint64(8)).astype(np.float32) * np.float32(1.0 / 16777216.0)
    return u >= np.float32(p)


def _dropout_fwd(self, x, y, mask, n, p, rng, salt, stream):
    st = _arr(rng, 4, np.int32).view(np.int64)
    keep = dropout_keep(n, p, int(st[0]), int(st[1]), int(salt))
    _arr(mask, n, np.uint8)[:] = keep
    xv = _arr(x, n).copy()
    _arr(y, n)[:] = np.where(keep, xv * (np.float32(1) / (np.float32(1) - np.float32(p))), np.float32(0))


def _dropout_bwd(self, g, out, mask, relu_ref, n, p, stream):
    v = _arr(g, n).copy()
    if mask:
        v = np.where(_arr(mask, n, np.uint8) != 0, v * (np.float32(1) / (np.float32(1) - np.float32(p))), np.float32(0))
    if relu_ref:
        v = np.where(_arr(relu_ref, n) > 0, v, np.float32(0))
    _arr(out, n)[:] = v


def _rng_advance(self, rng, stream):
    _arr(rng, 4, np.int32).view(np.int64)[1] += 1


def _add_layernorm_fwd(self, d, stream):
    r, n = d.rows, d.n
    z = _arr(d.x, r * n).reshape(r, n).astype(np.float64) + _arr(d.y, r * n).reshape(r, n)
    mean = z.mean(1, keepdims=True)
    rstd = 1.0 / np.sqrt(((z - mean) ** 2).mean(1, keepdims=True) + d.eps)
    _arr(d.z, r * n).reshape(r, n)[:] = z
    st = _arr(d.stat, r * 2).reshape(r, 2)
    st[:, 0:1], st[:, 1:2] = mean, rstd
    _arr(d.out, r * n).reshape(r, n)[:] = (z - mean) * rstd * _arr(d.gamma, n) + _arr(d.beta, n)


def _add_layernorm_bwd(self, d, stream):
    r, n = d.rows, d.n
    z = _arr(d.z, r * n).reshape(r, n).astype(np.float64)
    st = _arr(d.stat, r * 2).reshape(r, 2).astype(np.float64)
    g = _arr(d.gout, r * n).reshape(r, n).astype(np.float64)
    xh = (z - st[:, 0:1]) * st[:, 1:2]
    gy = g * _arr(d.gamma, n)
    _arr(d.dz, r * n).reshape(r, n)[:] = st[:, 1:2] * (gy - gy.mean(1, keepdims=True) - xh * (gy * xh).mean(1, keepdims=True))
    _arr(d.dgamma, n)[:] += (g * xh).sum(0).astype(np.float32)
    _arr(d.dbeta, n)[:] += g.sum(0).astype(np.float32)


Spec.bgemm, Spec.softmax_fwd, Spec.softmax_bwd = _bgemm, _softmax_fwd, _softmax_bwd
Spec.dropout_fwd, Spec.dropout_bwd, Spec.rng_advance = _dropout_fwd, _dropout_bwd, _rng_advance
Spec.add_layernorm_fwd, Spec.add_layernorm_bwd = _add_layernorm_fwd, _add_layernorm_bwd
