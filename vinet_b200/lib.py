"""ctypes binding of the C-ABI in ``include/vinet_b200.h`` (the drop-in boundary).

The product path has NO fallback: if ``libvinet_b200.so`` is missing or a call fails, this raises.
Structures mirror the header field by field; ``tests/test_abi.py`` checks ``sizeof`` of every struct
against the values compiled into the library is not possible without a compute call, so it checks
that every symbol declared in the header is exported and that the field lists agree with the header.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libvinet_b200.so")

BF16, F32 = 0, 1
XF_IDENT, XF_RELU, XF_AFFINE, XF_AFFINE_RELU = 0, 1, 2, 3
XF_UP2 = 4          # source stored at half resolution, read through the 2x bilinear up-sampling (after the ReLU bit)
GATHER_FPROP, GATHER_DGRAD = 0, 1
ENGINE_TC, ENGINE_SIMT = 0, 1
ACT_NONE, ACT_RELU, ACT_SIGMOID = 0, 1, 2
LOSS_KLDIV, LOSS_CC, LOSS_SIM, LOSS_NSS = 0, 1, 2, 3
KERNEL_GATHER, KERNEL_TMA = 0, 1
KLAYOUT_DENSE, KLAYOUT_TAP64, KLAYOUT_WIN8, KLAYOUT_WIN4 = 0, 1, 2, 3
MAX_TAPS = 64
TC_BLOCK_M, TC_BLOCK_K = 128, 64

_p, _i32, _i64, _f32 = C.c_void_p, C.c_int32, C.c_int64, C.c_float


class Src(C.Structure):
    _fields_ = [("ptr", _p), ("scale", _p), ("shift", _p), ("ld", _i64), ("T", _i32), ("xform", _i32), ("ldh", _i64), ("ldb", _i64)]


class Gather(C.Structure):
    _fields_ = [("mode", _i32), ("dtype", _i32), ("B", _i32), ("Tr", _i32), ("Hr", _i32), ("Wr", _i32),
                ("row_tstep", _i32), ("row_toff", _i32), ("Ts", _i32), ("Hs", _i32), ("Ws", _i32),
                ("Cs", _i32), ("ntaps", _i32), ("st", _i32), ("sh", _i32), ("sw", _i32),
                ("pt", _i32), ("ph", _i32), ("pw", _i32), ("tap", (C.c_int8 * 4) * MAX_TAPS), ("src", Src * 2)]


class Conv(C.Structure):
    _fields_ = [("g", Gather), ("w", _p), ("N", _i32), ("block_n", _i32), ("n_tiles", _i32), ("k_blocks", _i32),
                ("out", _p * 2), ("ldo", _i64 * 2), ("out_T", _i32 * 2), ("out_dtype", _i32), ("accumulate", _i32),
                ("ep_scale", _p), ("ep_shift", _p), ("ep_act", _i32), ("kernel", _i32), ("stats", _p)]


class Wgrad(C.Structure):
    _fields_ = [("g", Gather), ("dy", _p), ("lddy", _i64), ("dy_dtype", _i32), ("N", _i32), ("dwp", _p),
                ("lddw", _i32), ("splits", _i32), ("kernel", _i32)]


class Pack(C.Structure):
    _fields_ = [("w", _p), ("Cout", _i32), ("Cin", _i32), ("kt", _i32), ("kh", _i32), ("kw", _i32), ("cs", _i32),
                ("mode", _i32), ("ntaps", _i32), ("tap", (C.c_int8 * 4) * MAX_TAPS), ("engine", _i32),
                ("block_n", _i32), ("n_tiles", _i32), ("k_blocks", _i32), ("out", _p), ("layout", _i32), ("part", _i32), ("ld_cin", _i32)]


class PackInput(C.Structure):
    _fields_ = [("x", _p), ("sb", _i64), ("sc", _i64), ("st", _i64), ("sh", _i64), ("sw", _i64), ("B", _i32),
                ("C", _i32), ("T", _i32), ("H", _i32), ("W", _i32), ("cpad", _i32), ("out", _p), ("out_dtype", _i32),
                ("wl", _i32), ("Wp", _i32), ("out4", _p)]


class BnStats(C.Structure):
    _fields_ = [("y", _p), ("ld", _i64), ("dtype", _i32), ("rows", _i64), ("C", _i32), ("sums", _p)]


class BnFinalize(C.Structure):
    _fields_ = [("sums", _p), ("rows", _i64), ("C", _i32), ("gamma", _p), ("beta", _p), ("eps", _f32),
                ("momentum", _f32), ("running_mean", _p), ("running_var", _p), ("training", _i32), ("scale", _p),
                ("shift", _p), ("mean", _p), ("invstd", _p), ("sq_stride", _i64)]


class BnApply(C.Structure):
    _fields_ = [("y", _p), ("ldy", _i64), ("dtype", _i32), ("rows", _i64), ("C", _i32), ("relu", _i32), ("scale", _p),
                ("shift", _p), ("out", _p), ("ldo", _i64), ("out_dtype", _i32)]


class BnBwd(C.Structure):
    _fields_ = [("g", _p), ("ldg", _i64), ("y", _p), ("ldy", _i64), ("dtype", _i32), ("rows", _i64), ("C", _i32),
                ("relu", _i32), ("scale", _p), ("shift", _p), ("mean", _p), ("invstd", _p), ("gamma", _p),
                ("sums", _p), ("dgamma", _p), ("dbeta", _p), ("dy", _p), ("lddy", _i64), ("dy_dtype", _i32),
                ("training", _i32), ("g_dtype", _i32)]


class Pool(C.Structure):
    _fields_ = [("x", _p), ("ldx", _i64), ("dtype", _i32), ("scale", _p), ("shift", _p), ("xform", _i32),
                ("B", _i32), ("Ti", _i32), ("Hi", _i32), ("Wi", _i32), ("C", _i32),
                ("kt", _i32), ("kh", _i32), ("kw", _i32), ("st", _i32), ("sh", _i32), ("sw", _i32),
                ("pt", _i32), ("ph", _i32), ("pw", _i32), ("To", _i32), ("Ho", _i32), ("Wo", _i32),
                ("out", _p), ("ldo", _i64), ("out_dtype", _i32), ("gout", _p), ("ldgo", _i64), ("gin", _p),
                ("ldgi", _i64), ("gout_dtype", _i32), ("gin_dtype", _i32), ("idx", _p), ("gin_overwrite", _i32)]


class Upsample(C.Structure):
    _fields_ = [("z", _p), ("ldz", _i64), ("dtype", _i32), ("relu", _i32), ("B", _i32), ("T", _i32), ("h", _i32),
                ("w", _i32), ("C", _i32), ("u", _p), ("ldu", _i64), ("u_dtype", _i32), ("gu", _p), ("ldgu", _i64),
                ("dz", _p), ("lddz", _i64), ("dz_dtype", _i32), ("gu_dtype", _i32)]


class Head(C.Structure):
    _fields_ = [("x", _p), ("ldx", _i64), ("dtype", _i32), ("relu", _i32), ("rows", _i64), ("C", _i32), ("w", _p),
                ("b", _p), ("out", _p), ("gout", _p), ("dx", _p), ("lddx", _i64), ("dx_dtype", _i32), ("dw", _p),
                ("db", _p), ("up2", _i32), ("up_h", _i32), ("up_w", _i32), ("relu_pre", _i32)]


class Loss(C.Structure):
    _fields_ = [("kind", _i32), ("s", _p), ("g", _p), ("B", _i32), ("n", _i32), ("per_sample", _p), ("out", _p),
                ("counter", _p), ("gout", _p), ("grad_s", _p)]


class Conv1d(C.Structure):
    _fields_ = [("x", _p), ("w", _p), ("bias", _p), ("B", _i32), ("Cin", _i32), ("Lin", _i32), ("Cout", _i32),
                ("Lout", _i32), ("k", _i32), ("stride", _i32), ("pad", _i32), ("y", _p), ("dy", _p), ("dx", _p),
                ("dw", _p), ("dbias", _p)]


class Bn1d(C.Structure):
    _fields_ = [("y", _p), ("B", _i32), ("C", _i32), ("L", _i32), ("pool", _i32), ("gamma", _p), ("beta", _p),
                ("eps", _f32), ("momentum", _f32), ("running_mean", _p), ("running_var", _p), ("training", _i32),
                ("mean", _p), ("invstd", _p), ("out", _p), ("gout", _p), ("dy", _p), ("dgamma", _p), ("dbeta", _p)]


class AvFuse(C.Structure):
    _fields_ = [("y0", _p), ("ld", _i64), ("dtype", _i32), ("scale", _p), ("shift", _p), ("xform", _i32),
                ("audio", _p), ("w", _p), ("bias", _p), ("B", _i32), ("C", _i32), ("vbuf", _p), ("out", _p),
                ("ldo", _i64), ("out_dtype", _i32), ("gout", _p), ("ldgo", _i64), ("gy0", _p), ("ldgy0", _i64),
                ("gaudio", _p), ("dw", _p), ("dbias", _p)]


class Split(C.Structure):
    _fields_ = [("x", _p), ("ld", _i64), ("dtype", _i32), ("rows", _i64), ("C", _i32), ("scale", _p), ("shift", _p),
                ("xform", _i32), ("nparts", _i32), ("part", _p * 3), ("ldo", _i64), ("up_h", _i32), ("up_w", _i32)]


class PostProc(C.Structure):
    _fields_ = [("x", _p), ("N", _i32), ("H", _i32), ("W", _i32), ("oh", _i32), ("ow", _i32), ("blur", _i32), ("ws0", _p),
                ("ws1", _p), ("minmax", _p), ("out", _p)]


class PreProc(C.Structure):
    _fields_ = [("frames", _p), ("N", _i32), ("h", _i32), ("w", _i32), ("H", _i32), ("W", _i32), ("xb", _p), ("xk", _p),
                ("xks", _i32), ("yb", _p), ("yk", _p), ("yks", _i32), ("tmp", _p), ("mean", _f32 * 3), ("std", _f32 * 3), ("out", _p)]


UNPACK_MAX = 48


class Unpack(C.Structure):
    _fields_ = [("dwp", _p), ("grad", _p), ("lddw", _i32), ("cs", _i32), ("Cout", _i32), ("Cin", _i32), ("ntaps", _i32),
                ("win8_kh", _i32), ("win8_kw", _i32), ("begin", _i32)]


class Bgemm(C.Structure):
    _fields_ = [("A", _p), ("sAm", _i64), ("sAk", _i64), ("sAb1", _i64), ("sAb2", _i64), ("a_dtype", _i32), ("a_relu", _i32),
                ("a_scale", _p), ("a_shift", _p),
                ("B", _p), ("sBn", _i64), ("sBk", _i64), ("sBb1", _i64), ("sBb2", _i64), ("b_dtype", _i32), ("c_dtype", _i32),
                ("C", _p), ("sCm", _i64), ("sCn", _i64), ("sCb1", _i64), ("sCb2", _i64),
                ("M", _i32), ("N", _i32), ("K", _i32), ("nb1", _i32), ("nb2", _i32), ("alpha", _f32), ("relu", _i32),
                ("bias1", _p), ("s1m", _i64), ("s1n", _i64),
                ("bias2", _p), ("s2m", _i64), ("s2n", _i64), ("s2b1", _i64), ("s2b2", _i64), ("accumulate", _i32), ("a_xf_on_m", _i32)]


class AddLn(C.Structure):
    _fields_ = [("x", _p), ("y", _p), ("z", _p), ("stat", _p), ("gamma", _p), ("beta", _p), ("eps", _f32), ("n", _i32),
                ("rows", _i64), ("out", _p), ("gout", _p), ("dz", _p), ("dgamma", _p), ("dbeta", _p)]


# name -> (restype, argtypes); every function declared in include/vinet_b200.h
_S = C.c_void_p  # stream
SIGNATURES = {
    "vinet_conv_gemm": (C.c_int, [C.POINTER(Conv), _i32, _S]),
    "vinet_conv_tiling": (C.c_int, [C.POINTER(Conv), _i32, C.POINTER(_i32), C.POINTER(_i32)]),
    "vinet_conv_wgrad": (C.c_int, [C.POINTER(Wgrad), _i32, _S]),
    "vinet_conv_up2_fused": (C.c_int, [C.POINTER(Gather), _i32, _i32, _i32]),
    "vinet_conv_win4_fused": (C.c_int, [C.POINTER(Gather), _i32]),
    "vinet_relu_bwd": (C.c_int, [_p, _i64, _i32, _p, _i64, _i32, _i64, _i32, _p, _i64, _i32, _S]),
    "vinet_pack_weights": (C.c_int, [C.POINTER(Pack), _S]),
    "vinet_pack_weights_multi": (C.c_int, [_p, _p, _i32, _i64, _S]),
    "vinet_packed_weight_bytes": (C.c_size_t, [_i32, _i32, _i32, _i32, _i32]),
    "vinet_unpack_wgrad": (C.c_int, [_p, _i32, _i32, _p, _i32, _i32, _i32, _S]),
    "vinet_unpack_wgrad_win8": (C.c_int, [_p, _i32, _p, _i32, _i32, _i32, _i32, _S]),
    "vinet_unpack_wgrad_multi": (C.c_int, [C.POINTER(Unpack), _i32, _S]),
    "vinet_pack_input": (C.c_int, [C.POINTER(PackInput), _S]),
    "vinet_split_bf16": (C.c_int, [C.POINTER(Split), _S]),
    "vinet_saliency_postprocess": (C.c_int, [C.POINTER(PostProc), _S]),
    "vinet_preprocess_frames": (C.c_int, [C.POINTER(PreProc), _S]),
    "vinet_audio_window": (C.c_int, [_p, _i32, _i32, _p, _i32, _S]),
    "vinet_bn_stats": (C.c_int, [C.POINTER(BnStats), _S]),
    "vinet_bn_finalize": (C.c_int, [C.POINTER(BnFinalize), _S]),
    "vinet_bn_stats_finalize": (C.c_int, [C.POINTER(BnStats), C.POINTER(BnFinalize), _S]),
    "vinet_bn_apply": (C.c_int, [C.POINTER(BnApply), _S]),
    "vinet_bn_apply_stats_multi": (C.c_int, [C.POINTER(BnFinalize), C.POINTER(BnApply), _i32, _S]),
    "vinet_bn_bwd_reduce": (C.c_int, [C.POINTER(BnBwd), _S]),
    "vinet_bn_bwd_apply": (C.c_int, [C.POINTER(BnBwd), _S]),
    "vinet_bn_fwd_fused": (C.c_int, [C.POINTER(BnStats), C.POINTER(BnFinalize), C.POINTER(BnApply), _S]),
    "vinet_bn_bwd_fused": (C.c_int, [C.POINTER(BnBwd), _S]),
    "vinet_bn_stats_finalize_multi": (C.c_int, [C.POINTER(BnStats), C.POINTER(BnFinalize), _i32, _S]),
    "vinet_bn_apply_multi": (C.c_int, [C.POINTER(BnApply), _i32, _S]),
    "vinet_bn_bwd_multi": (C.c_int, [C.POINTER(BnBwd), _i32, _S]),
    "vinet_maxpool_fwd": (C.c_int, [C.POINTER(Pool), _S]),
    "vinet_maxpool_bwd": (C.c_int, [C.POINTER(Pool), _S]),
    "vinet_upsample_fwd": (C.c_int, [C.POINTER(Upsample), _S]),
    "vinet_upsample_bwd": (C.c_int, [C.POINTER(Upsample), _S]),
    "vinet_head_fwd": (C.c_int, [C.POINTER(Head), _S]),
    "vinet_head_bwd": (C.c_int, [C.POINTER(Head), _S]),
    "vinet_loss_fwd": (C.c_int, [C.POINTER(Loss), _S]),
    "vinet_loss_bwd": (C.c_int, [C.POINTER(Loss), _S]),
    "vinet_conv1d_fwd": (C.c_int, [C.POINTER(Conv1d), _S]),
    "vinet_conv1d_bwd": (C.c_int, [C.POINTER(Conv1d), _S]),
    "vinet_bn1d_fwd": (C.c_int, [C.POINTER(Bn1d), _S]),
    "vinet_bn1d_bwd": (C.c_int, [C.POINTER(Bn1d), _S]),
    "vinet_avfuse_fwd": (C.c_int, [C.POINTER(AvFuse), _S]),
    "vinet_avfuse_bwd": (C.c_int, [C.POINTER(AvFuse), _S]),
    "vinet_bgemm": (C.c_int, [C.POINTER(Bgemm), _S]),
    "vinet_softmax_fwd": (C.c_int, [_p, _i64, _i32, _S]),
    "vinet_softmax_bwd": (C.c_int, [_p, _p, _i64, _i32, _S]),
    "vinet_dropout_fwd": (C.c_int, [_p, _p, _p, _i64, _f32, _p, C.c_uint32, _S]),
    "vinet_dropout_bwd": (C.c_int, [_p, _p, _p, _p, _i64, _f32, _S]),
    "vinet_rng_advance": (C.c_int, [_p, _S]),
    "vinet_add_layernorm_fwd": (C.c_int, [C.POINTER(AddLn), _S]),
    "vinet_add_layernorm_bwd": (C.c_int, [C.POINTER(AddLn), _S]),
    "vinet_memset_async": (C.c_int, [_p, C.c_int, C.c_size_t, _S]),
    "vinet_axpy_f32": (C.c_int, [_p, _p, _i64, _i32, _S]),
    "vinet_colsum": (C.c_int, [_p, _i64, _i32, _i64, _i32, _p, _p, _S]),
    "vinet_last_error": (C.c_char_p, []),
    "vinet_last_kernel": (C.c_char_p, []),
    "vinet_version": (C.c_char_p, []),
    "vinet_device_info": (C.c_int, [C.POINTER(_i32), C.POINTER(_i32), C.POINTER(_i32)]),
    "vinet_launch_count": (_i64, []),
    "vinet_up2_launch_count": (_i64, []),
    "vinet_abi_sizes": (C.c_int, [C.POINTER(_i64), _i32]),
    "vinet_debug_set": (C.c_int, [_i32, _i32]),
}

# declaration order of the structs in the header (vinet_abi_sizes)
ABI_STRUCTS = [Src, Gather, Conv, Wgrad, Pack, PackInput, BnStats, BnFinalize, BnApply, BnBwd, Pool, Upsample, Head, Loss,
               Conv1d, Bn1d, AvFuse, Split, PostProc, PreProc, Unpack, Bgemm, AddLn]


class VinetError(RuntimeError):
    pass


class Library:
    """Loaded ``libvinet_b200.so``; ``call(name, *args)`` raises VinetError on a non-zero status."""

    def __init__(self, path=LIB_PATH):
        if not os.path.isfile(path):
            raise VinetError(
                "%s not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(there is no CPU or PyTorch fallback for this path)" % path)
        self.path = path
        self.dll = C.CDLL(path)
        self.fn = {}
        for name, (res, args) in SIGNATURES.items():
            f = getattr(self.dll, name)
            f.restype, f.argtypes = res, args
            self.fn[name] = f
        sizes = (_i64 * len(ABI_STRUCTS))()
        n = self.fn["vinet_abi_sizes"](sizes, len(ABI_STRUCTS))
        mine = [C.sizeof(t) for t in ABI_STRUCTS]
        if n != len(ABI_STRUCTS) or list(sizes) != mine:
            raise VinetError("ABI mismatch between lib.py and %s: %s vs %s" % (path, list(sizes), mine))
        for kv in os.environ.get("VINET_DEBUG_SET", "").split(","):      # development knobs, e.g. VINET_DEBUG_SET=4=200
            if "=" in kv:
                self.call("vinet_debug_set", int(kv.split("=")[0]), int(kv.split("=")[1]))

    def call(self, name, *args):
        rc = self.fn[name](*args)
        if rc != 0:
            raise VinetError("%s failed (%d): %s" % (name, rc, self.fn["vinet_last_error"]().decode()))

    def version(self):
        return self.fn["vinet_version"]().decode()

    def launch_count(self):
        return int(self.fn["vinet_launch_count"]())

    def device_info(self):
        a, b, c = _i32(), _i32(), _i32()
        self.call("vinet_device_info", C.byref(a), C.byref(b), C.byref(c))
        return a.value, b.value, c.value


_lib = None


def get():
    global _lib
    if _lib is None:
        _lib = Library()
    return _lib
