"""GPU: every CUDA kernel, called through the C-ABI via the engine, against the numpy kernel spec run on
the CPU with the same seeded inputs (fp32 engine: tight; bf16 tcgen05 engine: bf16-rounding tolerance)."""
import os

import numpy as np
import pytest
import torch

import scenarios as S
from oracle import torch_oracle as O
from oracle.kernel_spec import Spec

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")

CONV_CASES = [dict(), dict(T0=1, T1=4, kt=5, Cin=24, Cout=16, H=4, W=7),
              dict(B=1, T0=4, T1=8, kt=3, Cin=64, Cout=96, H=14, W=24),
              dict(B=2, T0=2, T1=4, kt=3, Cin=192, Cout=288, H=8, W=12)]


def _reference(fn, precision, **kw):
    """fp32 engine: the numpy kernel spec on the CPU.  bf16 tcgen05 engine: the fp32-FFMA engine on the SAME
    bf16-stored tensors (identical rounding points, so only the accumulation order differs) — BN backward
    gradients cancel heavily, which makes a comparison across different storage precisions meaningless."""
    if precision == "fp32":
        return fn("cpu", "fp32", Spec(), **kw), 1e-3
    return fn("cuda", "bf16_simt", **kw), 1e-2


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
@pytest.mark.parametrize("case", range(len(CONV_CASES)))
def test_conv_concat_relu_upsample(precision, case):
    kw = CONV_CASES[case]
    ref, rtol = _reference(S.conv_up, precision, **kw)
    got = S.conv_up("cuda", precision, **kw)
    bad = S.compare(got, ref, rtol)
    assert not bad, bad
    if precision == "bf16":      # forward only: bf16 storage vs the fp32 spec stays within bf16 rounding
        spec = S.conv_up("cpu", "fp32", Spec(), **kw)
        assert not S.compare({"out": got["out"]}, {"out": spec["out"]}, 1e-2)


# a6 (north_star: "trilinear upsample fused into the following conv's input stage"): geometries of the decoder stages
# convtsp2 / convtsp3 / convtsp4.0 / convtsp4.3 (channel counts incl. a ragged last 64-block, T-concat with a skip tensor, maps
# that are not multiples of the 8 x 16 tiles) + a map too small for the halo kernels (falls back to materialising)
UP_CASES = [dict(B=2, T0=1, T1=2, kt=3, h=7, w=12, C0=64, C1=96, C2=48),
            dict(B=1, T0=2, T1=3, kt=5, h=14, w=24, C0=32, C1=480, C2=192),
            dict(B=2, T0=1, T1=1, kt=2, h=11, w=19, C0=16, C1=64, C2=32),
            dict(B=1, T0=2, T1=4, kt=3, h=28, w=20, C0=24, C1=192, C2=64)]


@pytest.mark.parametrize("precision", ["fp32", "bf16", "bf16x6"])
@pytest.mark.parametrize("case", range(len(UP_CASES)))
def test_upsample_fused_into_conv_input_stage(precision, case):
    """The consumer convolution reads relu + 2x bilinear of the LOW-RES tensor in its input stage.  bf16: the interpolating
    producer warps of conv_stream_kernel / conv_wgrad_halo_kernel must reproduce the plan that materialises the hi-res tensor
    BIT FOR BIT (same arithmetic, same bf16 rounding point, same MMA order) - borders, ragged tiles and all; fp32 / bf16x6 are
    held against the numpy spec / PyTorch autograd."""
    kw = UP_CASES[case]
    got = S.up_fused("cuda", precision, **kw)
    assert got["_materialized"] == ["c2"], got["_materialized"]      # only the scenario's explicit output materialisation
    if precision == "bf16":
        assert got["_up2_launches"] == 2, (got["_up2_launches"], got["_kernels"])     # fprop + weight gradient of c2
        assert set(got["_kernels"]) <= {"conv_stream_kernel", "conv_wgrad_halo_kernel", "conv_gemm_tma_kernel", "conv_wgrad_tma_kernel"}
        ref = S.up_fused("cuda", precision, fuse=False, **kw)
        assert ref["_materialized"] == ["c1", "c2"] and ref["_up2_launches"] == 0
        for k in ("out", "dx", "dy"):
            assert torch.equal(got[k], ref[k]), (k, (got[k].float() - ref[k].float()).abs().max())
        # weight gradients accumulate with fp32 atomics over position chunks (order varies run to run)
        dw = S.compare({k: got[k] for k in ("dW1", "dW2")}, {k: ref[k] for k in ("dW1", "dW2")}, 1e-5)
        assert not dw, dw
        simt = S.up_fused("cuda", "bf16_simt", **kw)
        assert not S.compare(got, simt, 1e-2), S.compare(got, simt, 1e-2)
    else:
        ref = S.up_fused_torch(**kw)
        tol = 1e-3 if precision == "fp32" else 2e-4
        assert not S.compare(got, ref, tol), S.compare(got, ref, tol)
        if precision == "fp32" and case == 0:
            spec = S.up_fused("cpu", "fp32", Spec(), **kw)
            assert not S.compare(got, spec, 1e-4), S.compare(got, spec, 1e-4)


def test_upsample_fusion_falls_back_on_small_maps():
    """7 x 12 hi-res maps are below the halo kernels' tile: the engine asks vinet_conv_up2_fused and materialises instead."""
    got = S.up_fused("cuda", "bf16", B=1, T0=1, T1=2, kt=3, h=3, w=6, C0=16, C1=32, C2=16)
    assert got["_materialized"] == ["c1", "c2"] and got["_up2_launches"] == 0
    ref = S.up_fused("cuda", "bf16_simt", B=1, T0=1, T1=2, kt=3, h=3, w=6, C0=16, C1=32, C2=16)
    assert not S.compare(got, ref, 1e-2), S.compare(got, ref, 1e-2)


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
@pytest.mark.parametrize("name", ["3b", "4c", "5c"])
def test_mixed_block(precision, name):
    ref, rtol = _reference(S.mixed, precision, name=name)
    got = S.mixed("cuda", precision, name=name)
    # 11 chained conv/BN/ReLU layers at ~100 positions: ONE ReLU-mask flip between two correct implementations
    # (inputs within 1e-6 of zero) moves a whole column of a weight gradient by 1/sqrt(rows); budget for a few
    bad = S.compare(got, ref, 10 * rtol)
    assert not bad, bad


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_stem_sepconv_and_pool(precision):
    ref, rtol = _reference(S.stem, precision)
    got = S.stem("cuda", precision)
    bad = S.compare(got, ref, 3 * rtol)
    assert not bad, bad


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_audio_branch_and_bilinear_fusion(precision):
    """SoundNet conv1d/BN/pool kernels + the AV fusion kernel vs PyTorch autograd (fp32 audio path)."""
    ref = S.audio_fuse_torch()
    got = S.audio_fuse("cuda", precision)
    # conv biases feed a train-mode BatchNorm: their true gradient is 0 and both sides hold rounding noise
    bad = S.compare(got, ref, 2e-3 if precision == "fp32" else 2e-2, skip=[".bias"] if True else ())
    bad = [b for b in bad if "bilinear.bias" in b or ".bias" not in b]
    assert not bad, bad


def test_losses_against_reference_goldens():
    """north_star: loss scalars within 1e-5 relative of the reference."""
    from vinet_b200 import loss as PL
    z = np.load(os.path.join(GOLD, "losses.npz"))
    for tag in O.LOSS_CASES:
        s, gt, fix = O.make_loss_inputs(tag)
        s = s.cuda().requires_grad_(True)
        for nm, fn, tgt in [("kldiv", PL.kldiv, gt), ("cc", PL.cc, gt), ("sim", PL.similarity, gt), ("nss", PL.nss, fix)]:
            v = fn(s, tgt.cuda())
            (g,) = torch.autograd.grad(v, s)
            ref = float(z[f"{tag}/{nm}"])
            assert abs(v.item() - ref) <= 1e-5 * abs(ref) + 1e-9, (tag, nm, v.item(), ref)
            gg = g.cpu().numpy() if tag == "a" else g.cpu().numpy()[:, ::7, ::5]
            rg = z[f"{tag}/{nm}_grad"]
            assert np.allclose(gg, rg, rtol=2e-3, atol=2e-5 * np.abs(rg).max()), (tag, nm, np.abs(gg - rg).max())


def test_loss_func_contract():
    from vinet_b200 import loss as PL

    class A:
        kldiv, cc, sim, l1 = True, True, True, False
        kldiv_coeff, cc_coeff, sim_coeff, batch_size = 1.0, -1.0, -1.0, 3
    s, gt, _ = O.make_loss_inputs("a")
    out = PL.loss_func(s.cuda(), gt.cuda(), A)
    assert out.shape == (1,) and out.is_cuda
    ref = O.kldiv(s, gt) - O.cc(s, gt) - O.similarity(s, gt)
    assert abs(out.item() - ref.item()) <= 1e-5 * abs(ref.item()) + 1e-6


@pytest.mark.parametrize("shape", [(2, 5, 9, 11, 24), (1, 1, 4, 6, 8), (2, 16, 28, 48, 64)])
def test_maxpool333_frame_walking_kernel_matches_scan_order_kernel(shape):
    """The 3x3x3/s1/p1 fast path (frame-walking, packed bf16) against the generic scan-order kernel: values AND the
    recorded arg-max taps must agree bit for bit, including ties (post-ReLU activations are full of equal zeros)."""
    import ctypes as C
    from vinet_b200 import lib as L
    lib = L.get()
    B, T, H, W, Cn = shape
    g = torch.Generator().manual_seed(5)
    x = torch.randint(-2, 3, (B, T, H, W, Cn), generator=g).float().clamp_min(0).to(torch.bfloat16).cuda()   # many ties
    res = []
    for fast in (1, 0):
        lib.call("vinet_debug_set", 3, fast)
        out = torch.full((B, T, H, W, Cn), float("nan"), dtype=torch.bfloat16, device="cuda")
        idx = torch.full((B, T, H, W, Cn), 255, dtype=torch.uint8, device="cuda")
        d = L.Pool()
        d.x, d.ldx, d.dtype, d.xform = x.data_ptr(), Cn, L.BF16, L.XF_IDENT
        d.B, d.Ti, d.Hi, d.Wi, d.C = B, T, H, W, Cn
        d.kt = d.kh = d.kw = 3
        d.st = d.sh = d.sw = 1
        d.pt = d.ph = d.pw = 1
        d.To, d.Ho, d.Wo, d.out, d.ldo, d.out_dtype, d.idx = T, H, W, out.data_ptr(), Cn, L.BF16, idx.data_ptr()
        lib.call("vinet_maxpool_fwd", C.byref(d), torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        res.append((out.float().cpu(), idx.cpu()))
    lib.call("vinet_debug_set", 3, 1)
    assert torch.equal(res[0][0], res[1][0])
    assert torch.equal(res[0][1], res[1][1])
    ref = torch.nn.functional.max_pool3d(x.float().permute(0, 4, 1, 2, 3), 3, 1, 1).permute(0, 2, 3, 4, 1).cpu()
    assert torch.equal(res[0][0], ref)


@pytest.mark.parametrize("cfg", [((1, 3, 3), (1, 2, 2), (0, 1, 1)), ((3, 3, 3), (2, 2, 2), (1, 1, 1)), ((2, 2, 2), (2, 2, 2), (0, 0, 0)),
                                 ((3, 3, 3), (1, 1, 1), (1, 1, 1))])
@pytest.mark.parametrize("overwrite", [0, 1])
@pytest.mark.parametrize("gdt", ["f32", "bf16"])
def test_maxpool_backward_gather_matches_scatter_and_autograd(cfg, overwrite, gdt):
    """Gather-over-recorded-taps backward vs the atomic scatter kernel vs PyTorch autograd (fp32 gradients: exact up to
    summation order), as first writer (overwrite) and as accumulating consumer."""
    import ctypes as C
    from vinet_b200 import lib as L
    lib = L.get()
    k, s, p = cfg
    B, T, H, W, Cn = 2, 6, 10, 12, 16
    g = torch.Generator().manual_seed(7)
    x = torch.randint(-2, 3, (B, T, H, W, Cn), generator=g).float().clamp_min(0).to(torch.bfloat16).cuda()
    To, Ho, Wo = [(n + 2 * pp - kk) // ss + 1 for n, kk, ss, pp in zip((T, H, W), k, s, p)]
    tdt, ldt = (torch.float32, L.F32) if gdt == "f32" else (torch.bfloat16, L.BF16)
    gout = torch.randn(B, To, Ho, Wo, Cn, generator=g).to(tdt).float().cuda()
    gin0 = torch.randn(B, T, H, W, Cn, generator=g).to(tdt).float().cuda()
    xr = x.float().permute(0, 4, 1, 2, 3).requires_grad_(True)
    torch.nn.functional.max_pool3d(xr, k, s, p).backward(gout.permute(0, 4, 1, 2, 3))
    ref = xr.grad.permute(0, 2, 3, 4, 1) + (0 if overwrite else gin0)
    res = []
    for fast in (3, 5, 9):       # 3: generic gather, 5: atomic scatter, 9: compile-time-specialised gather for every geometry
        lib.call("vinet_debug_set", 3, fast)
        out = torch.empty(B, To, Ho, Wo, Cn, dtype=torch.bfloat16, device="cuda")
        idx = torch.empty(B, To, Ho, Wo, Cn, dtype=torch.uint8, device="cuda")
        gin = gin0.clone().to(tdt)
        gout_d = gout.to(tdt)
        d = L.Pool()
        d.x, d.ldx, d.dtype, d.xform = x.data_ptr(), Cn, L.BF16, L.XF_IDENT
        d.B, d.Ti, d.Hi, d.Wi, d.C = B, T, H, W, Cn
        (d.kt, d.kh, d.kw), (d.st, d.sh, d.sw), (d.pt, d.ph, d.pw) = k, s, p
        d.To, d.Ho, d.Wo, d.out, d.ldo, d.out_dtype, d.idx = To, Ho, Wo, out.data_ptr(), Cn, L.BF16, idx.data_ptr()
        st = torch.cuda.current_stream().cuda_stream
        lib.call("vinet_maxpool_fwd", C.byref(d), st)
        d.gout, d.ldgo, d.gin, d.ldgi, d.gout_dtype, d.gin_dtype, d.gin_overwrite = gout_d.data_ptr(), Cn, gin.data_ptr(), Cn, ldt, ldt, overwrite
        lib.call("vinet_maxpool_bwd", C.byref(d), st)
        torch.cuda.synchronize()
        res.append(gin.float().cpu())
    lib.call("vinet_debug_set", 3, 1)
    # bf16 gradients: up to 27 terms summed and stored in bf16 (|values| reach ~16, where one bf16 ulp is 0.0625)
    rtol, atol = (1e-6, 1e-6) if gdt == "f32" else (2e-2, 1.3e-1)
    assert torch.allclose(res[0], res[1], rtol=rtol, atol=atol)
    assert torch.allclose(res[0], ref.cpu(), rtol=rtol, atol=atol)
    assert torch.allclose(res[0], res[2], rtol=rtol, atol=atol)


SWEEP_CONVS = [  # B, T0, T1, H, W, Cin, Cout, k, stride_t, pad  (geometries the 128x192 .. 448x768 / T=8..48 sweep produces)
    (1, 24, 0, 28, 48, 64, 64, (3, 1, 1), 1, (1, 0, 0)),      # T=48 backbone: 24 frames = three temporal-halo tiles
    (1, 6, 0, 16, 24, 96, 96, (3, 1, 1), 1, (1, 0, 0)),       # 6 frames: one partial temporal-halo tile
    (1, 12, 0, 20, 28, 32, 64, (1, 3, 3), 1, (0, 1, 1)),      # W = 28: ragged halo tiles in both directions
    (1, 6, 12, 14, 24, 64, 64, (3, 3, 3), 3, (0, 1, 1)),      # T=48 decoder: 6 + 12 concatenated frames -> 6
    (1, 2, 8, 56, 96, 32, 32, (5, 3, 3), 5, (0, 1, 1)),       # T=16 decoder: 2 + 8 -> 2
    (2, 3, 0, 64, 96, 64, 32, (2, 3, 3), 2, (0, 1, 1)),       # convtsp4.3 with an odd frame count
    (1, 8, 0, 112, 192, 64, 192, (1, 3, 3), 1, (0, 1, 1)),    # 448x768: base1.3.conv_s on a 112x192 map
    (1, 4, 0, 32, 48, 192, 192, (3, 1, 1), 1, (1, 0, 0)),     # 128x192: base1.3.conv_t on a 32x48 map
    (1, 48, 0, 16, 24, 64, 64, (7, 1, 1), 2, (3, 0, 0)),      # T=48 stem conv_t: 24 output frames = 1.5 strided temporal-halo tiles
    (2, 2, 0, 8, 12, 256, 320, (1, 3, 3), 1, (0, 1, 1)),      # 8-row map: stays on the per-tap kernel
]


@pytest.mark.parametrize("case", range(len(SWEEP_CONVS)))
def test_conv_kernels_match_ffma_engine_on_sweep_geometries(case):
    """Every tcgen05 conv kernel (streaming fprop / dgrad in halo, temporal-halo and frame-walk modes, halo weight gradient,
    per-tap kernels) against the fp32-FFMA engine on identical bf16 inputs: out, dX (both sources of a concat) and dW."""
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
    import diag_tma as D
    c = SWEEP_CONVS[case]
    a = D.run("bf16", True, *c)
    b = D.run("fp32", False, *c)
    for k_ in a:
        ref = b[k_]
        scale = ref.abs().max().item() + 1e-20
        rel = (a[k_] - ref).abs().max().item() / scale
        tol = 2e-3 if k_ == "dW" else 3e-2       # out / dx are stored in bf16; dW is an fp32 sum of bf16 products
        assert rel <= tol, (k_, rel)
