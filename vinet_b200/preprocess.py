"""Device-side input pipeline (reference dataloader.py:242-249, 283-300, generate_result.py:77-89).

The reference decodes 32 images per clip on DataLoader workers and runs PIL resize + ToTensor + Normalize per frame on the CPU,
which cannot feed hundreds of clips per second.  Here the workers only decode; the resize (Pillow's antialiased bilinear
resampling reproduced bit for bit, csrc/preproc.cu), the scaling and the normalisation run on the GPU for a whole batch of frames
and write the (N, 3, H, W) fp32 layout that `VideoSaliencyModel` / `SlidingWindowSaliency` consume (a (B,T,3,H,W) clip batch is
`out.view(B, T, 3, H, W)`, whose `.permute(0,2,1,3,4)` is exactly what train.py:205 hands to the model).

    pre = FramePreprocessor((224, 384))
    frames = pre(uint8_frames)              # (N, h, w, 3) uint8 RGB, host (pinned) or device -> (N, 3, 224, 384) fp32 on the device
    audio = audio_window(excerpt)           # (B, n) fp32 -> (B, 1, 70560, 1) Hanning-windowed, centred  (dataloader.py:113-118)
"""
import ctypes as C
import math

import torch

from . import arch
from . import lib as L

MEAN, STD = (0.485, 0.456, 0.406), (0.229, 0.224, 0.225)        # dataloader.py:245-248
_BITS = 22                                                      # Pillow Resample.c: PRECISION_BITS = 32 - 8 - 2


def resample_tables(in_size, out_size):
    """Pillow's bilinear resampling tables for one axis (Resample.c: precompute_coeffs with the bilinear filter, support 1, then
    normalize_coeffs_8bpc): (bounds [out][2] int32 = first source index / tap count, coefficients [out][ksize] int32)."""
    scale = in_size / out_size
    filterscale = max(scale, 1.0)
    support = filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    bounds = torch.zeros((out_size, 2), dtype=torch.int32)
    kk = torch.zeros((out_size, ksize), dtype=torch.int32)
    inv = 1.0 / filterscale
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        xmin = max(int(center - support + 0.5), 0)
        xmax = min(int(center + support + 0.5), in_size) - xmin
        w = [max(1.0 - abs((x + xmin - center + 0.5) * inv), 0.0) for x in range(xmax)]
        ww = sum(w)
        if ww != 0.0:
            w = [v / ww for v in w]
        for x, v in enumerate(w):
            kk[xx, x] = int(-0.5 + v * (1 << _BITS)) if v < 0 else int(0.5 + v * (1 << _BITS))
        bounds[xx, 0], bounds[xx, 1] = xmin, xmax
    return bounds, kk


class FramePreprocessor:
    def __init__(self, out_hw=(224, 384), mean=MEAN, std=STD):
        self.out_hw, self.mean, self.std = tuple(out_hw), tuple(mean), tuple(std)
        self._tables = {}

    def _tabs(self, h, w, device):
        key = (h, w, str(device))
        if key not in self._tables:
            xb, xk = resample_tables(w, self.out_hw[1])
            yb, yk = resample_tables(h, self.out_hw[0])
            self._tables[key] = tuple(t.to(device).contiguous() for t in (xb, xk, yb, yk))
        return self._tables[key]

    def __call__(self, frames, device=None):
        """frames: (N, h, w, 3) uint8 RGB (decoded images), on the host or the device -> (N, 3, H, W) fp32 on the device."""
        assert frames.dtype == torch.uint8 and frames.dim() == 4 and frames.shape[3] == 3, "expected (N, h, w, 3) uint8 RGB frames"
        device = torch.device(device) if device is not None else (frames.device if frames.is_cuda else torch.device("cuda"))
        frames = frames.to(device, non_blocking=True).contiguous()
        n, h, w, _ = frames.shape
        H, W = self.out_hw
        xb, xk, yb, yk = self._tabs(h, w, device)
        tmp = torch.empty((n, h, W, 3), dtype=torch.uint8, device=device)
        out = torch.empty((n, 3, H, W), dtype=torch.float32, device=device)
        d = L.PreProc()
        d.frames, d.N, d.h, d.w, d.H, d.W = frames.data_ptr(), n, h, w, H, W
        d.xb, d.xk, d.xks, d.yb, d.yk, d.yks = xb.data_ptr(), xk.data_ptr(), xk.shape[1], yb.data_ptr(), yk.data_ptr(), yk.shape[1]
        d.tmp, d.out = tmp.data_ptr(), out.data_ptr()
        for i in range(3):
            d.mean[i], d.std[i] = self.mean[i], self.std[i]
        L.get().call("vinet_preprocess_frames", C.byref(d), torch.cuda.current_stream(device).cuda_stream)
        return out


def audio_window(excerpt, total=arch.AUDIO_LEN):
    """dataloader.py:113-118 (`get_audio_feature`): every excerpt row times np.hanning(n), centred in `total` zero samples;
    returns the (B, 1, total, 1) tensor `VideoAudioSaliencyModel.forward` takes."""
    assert excerpt.is_cuda and excerpt.dtype == torch.float32 and excerpt.dim() == 2
    excerpt = excerpt.contiguous()
    b, n = excerpt.shape
    out = torch.empty((b, 1, total, 1), dtype=torch.float32, device=excerpt.device)
    L.get().call("vinet_audio_window", excerpt.data_ptr(), b, n, out.data_ptr(), total, torch.cuda.current_stream(excerpt.device).cuda_stream)
    return out
