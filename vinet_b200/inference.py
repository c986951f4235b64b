"""Sliding-window video inference (reference generate_result.py:55-73, 96-104, and the audio-visual loop of
generate_result_audio_visual.py:86-118, 177-199).

The reference predicts ONE saliency frame per 32-frame clip forward: frame i (i >= L-1) from the clip of frames [i-L+1 .. i], and
the first L-1 frames from the same clips reversed in time (``torch.flip(clip, [2])``, generate_result.py:70-71), so consecutive
clips overlap by L-1 of L frames and every frame is pushed host -> device L times.  Here

  * the pre-processed frames of a video live ONCE in a device buffer ``(N, 3, H, W)``; a batch of B consecutive windows is a
    strided *view* of it (``x[b, c, t] = frames[i0 + b + t, c]``: batch stride = one frame), which the packing kernel of the
    model reads through its strides - no clip is ever materialised, no frame is copied twice;
  * B windows go through one forward of the folded-BatchNorm eval plan, replayed as a CUDA graph (``GraphedForward``);
  * the time-flipped clips of the first L-1 frames are windows of a reversed copy of the first 2L-2 frames;
  * resize to the source resolution + 11x11 Gaussian blur + min-max normalisation to 8 bits (generate_result.py:100-104,
    utils.py:61-78) run on the device (``postprocess``), so only finished 8-bit maps travel back to the host.

    sal = SlidingWindowSaliency(model, clip_len=32, windows_per_batch=8)
    maps = sal(frames)                    # (N, 3, H, W) normalised fp32 frames -> (N, H, W) fp32 saliency in (0, 1)
    png = sal.postprocess(maps, (360, 640))   # (N, 360, 640) uint8, what generate_result.py writes to disk

Audio-visual models (``VideoAudioSaliencyModel`` / ``VideoAudioSaliencyFusionModel``) take the video's sound track as well:
``maps = sal(frames, audio=AudioTrack(wav, sample_rate, fps, n_frames))``; every window gets the Hanning-windowed excerpt the
reference cuts for it (and its time reversal for the flipped clips), built on the device.
"""
import ctypes as C

import torch

from . import lib as L
from .graph import GraphedForward


def window_view(frames, start, count, clip_len):
    """(count, 3, clip_len, H, W) view of a (N, 3, H, W) frame buffer: window b holds frames start+b .. start+b+clip_len-1."""
    n, c, h, w = frames.shape
    assert frames.is_contiguous() and start >= 0 and start + count + clip_len - 1 <= n
    s = frames.stride()
    return torch.as_strided(frames, (count, c, clip_len, h, w), (s[0], s[1], s[0], s[2], s[3]), frames.storage_offset() + start * s[0])


class AudioTrack:
    """Sound track of one video with the reference's per-frame excerpt bounds (generate_result_audio_visual.py:56-66, the same
    arithmetic as dataloader.py:56-70) and per-window features (`get_audio_feature`, :86-118).
    wav: the mono waveform as the reference holds it (``torchaudio.load(normalization=False)[0] * 2**-23``), any device."""

    def __init__(self, wav, sample_rate, fps, n_frames):
        self.wav = wav.reshape(-1).float().contiguous()
        n = self.wav.numel()
        fs, fps = float(sample_rate), float(fps)
        n_samples = fs / fps
        self.starts, self.ends = [0] * (n_frames + 1), [0] * (n_frames + 1)
        for vf in range(1, n_frames + 1):
            centre = (vf - 1) * (1.0 / fps) * fs
            self.starts[vf] = int(max(0, centre - n_samples / 2))
            self.ends[vf] = int(min(n, abs(centre + n_samples / 2)))

    def bounds(self, start_idx, clip_len):
        """[begin, end) sample range of the window whose first frame is `start_idx` (0-based), as `wav[:, s:e+1]` slices it."""
        s = self.starts[start_idx + 1]
        e = self.ends[-1] if start_idx + clip_len >= len(self.ends) else self.ends[start_idx + clip_len]
        return s, max(s, min(e + 1, self.wav.numel()))

    def features(self, start_indices, clip_len, device, flip=False):
        """(len(start_indices), 1, AUDIO_LEN, 1) fp32 on `device`: excerpt x np.hanning, centred in zeros; flip = the time-reversed
        feature the reference feeds with a time-flipped clip (generate_result_audio_visual.py:194-196)."""
        from . import arch
        if self.wav.device != device:
            self.wav = self.wav.to(device)
        out = torch.empty((len(start_indices), 1, arch.AUDIO_LEN, 1), dtype=torch.float32, device=device)
        lib, st = L.get(), torch.cuda.current_stream(device).cuda_stream
        for i, idx in enumerate(start_indices):
            s, e = self.bounds(idx, clip_len)
            assert e - s <= arch.AUDIO_LEN, "audio excerpt longer than the model's window (fps below 10?)"
            lib.call("vinet_audio_window", self.wav.data_ptr() + 4 * s, 1, e - s, out[i].data_ptr(), arch.AUDIO_LEN, st)
        return torch.flip(out, [2]) if flip else out


class SlidingWindowSaliency:
    """One saliency map per frame of a video with the reference's windowing (generate_result.py:55-73)."""

    def __init__(self, model, clip_len=32, windows_per_batch=8, use_graph=True, stem_cache=None):
        assert not model.training, "call model.eval() first (inference uses the running BatchNorm statistics)"
        self.model, self.L, self.B, self.use_graph = model, clip_len, windows_per_batch, use_graph
        # per-frame stem re-use (model.forward_windows): the (1,7,7) stem conv runs once per frame of a batch instead of once per
        # (window, frame); needs the TMA-fed bf16 engine
        self.stem_cache = (getattr(model, "precision", "") == "bf16" and hasattr(model, "forward_windows")) if stem_cache is None else stem_cache
        self._graphs = {}

    def _forward(self, x, a=None):
        """x: (b, 3, L, H, W) strided window view [, a: (b, 1, AUDIO_LEN, 1) audio features] -> (b, H, W)."""
        if not self.use_graph:
            with torch.no_grad():
                if a is not None:
                    return self.model(x, a)
                if self.stem_cache:
                    b, c, t, h, w = x.shape
                    store = torch.empty((b + t - 1, c, h, w), dtype=x.dtype, device=x.device)
                    _fill_store(store, x)
                    return self.model.forward_windows(store, b)
                return self.model(x)
        key = (tuple(x.shape), tuple(x.stride()))
        g = self._graphs.get(key)
        if g is not None and g[0].stale():        # another shape ran through the model since: its buffers were re-allocated
            g = None
        if g is None:
            self._graphs.clear()                   # one live graph: the engine keeps one set of activation buffers per model
            # the static input keeps the caller's strides: an overlapping (batch stride = 1 frame) window view needs a buffer of
            # b + L - 1 frames, not b * L
            b, c, t, h, w = x.shape
            store = torch.empty((b + t - 1, c, h, w), dtype=x.dtype, device=x.device)
            static = window_view(store, 0, b, t)
            assert tuple(static.stride()) == tuple(x.stride())
            g = self._graphs[key] = (_StaticWindows(self.model, store, static, b if self.stem_cache else 0, audio=a), store)
        runner, store = g
        return runner(x, a)

    def refresh(self):
        """Model weights changed: drop the captured graphs."""
        self._graphs = {}
        self.model.invalidate_weight_cache()

    def __call__(self, frames, audio=None):
        """frames: (N, 3, H, W) fp32, pre-processed as by generate_result.py:77-89 (resize, ToTensor, ImageNet normalise), on the
        host or on the device; audio: an AudioTrack for the audio-visual models.  Returns (N, H, W) fp32 saliency maps on the
        device; N >= 2L-1 like the reference requires."""
        Lc, n = self.L, frames.shape[0]
        av = getattr(self.model, "_n_extra", 0) == 1
        assert av == (audio is not None), "audio-visual models need audio=AudioTrack(...), ViNet takes none"
        if n < 2 * Lc - 1:
            raise ValueError("more frames are needed: %d < %d (generate_result.py:55)" % (n, 2 * Lc - 1))
        dev = next(self.model.parameters()).device
        frames = frames.to(dev, non_blocking=True).contiguous()
        out = torch.empty((n,) + tuple(frames.shape[2:]), dtype=torch.float32, device=dev)
        # Every batch has the SAME number of windows (one captured graph, one set of activation buffers): the last batch of a
        # pass is shifted back to be full and recomputes a few windows of the previous one.
        # frames L-1 .. N-1: window ending at the frame
        nwin = n - Lc + 1
        b = min(self.B, Lc - 1, nwin)
        for w0 in range(0, nwin, b):
            w0 = min(w0, nwin - b)
            a = audio.features(range(w0, w0 + b), Lc, dev) if av else None
            out[w0 + Lc - 1:w0 + Lc - 1 + b] = self._forward(window_view(frames, w0, b, Lc), a)
        # frames 0 .. L-2: the clip STARTING at the frame, reversed in time (its last frame is the wanted one):
        # rev[k] = frames[2L-3-k]; the reversed clip of frame j is the window of `rev` starting at k = L-2-j
        rev = torch.flip(frames[:2 * Lc - 2], [0]).contiguous()
        for k0 in range(0, Lc - 1, b):
            k0 = min(k0, Lc - 1 - b)
            a = audio.features([Lc - 2 - (k0 + i) for i in range(b)], Lc, dev, flip=True) if av else None
            pred = self._forward(window_view(rev, k0, b, Lc), a)             # pred[i] belongs to frame j = L-2-(k0+i)
            out[Lc - 2 - k0 - b + 1:Lc - 2 - k0 + 1] = torch.flip(pred, [0])
        return out

    # ------------------------------------------------------------------ device post-processing
    def postprocess(self, maps, size_wh, blur=True):
        """generate_result.py:100-104 on the device: bilinear resize of every (H, W) map to the source resolution (cv2.resize
        semantics), 11x11 Gaussian blur (sigma 2, reflect-101 border), per-map min-max normalisation and rounding to 8 bits
        (utils.img_save(normalize=True)).  maps: (N, H, W) fp32 -> (N, size_wh[1], size_wh[0]) uint8."""
        lib = L.get()
        n, h, w = maps.shape
        ow, oh = int(size_wh[0]), int(size_wh[1])
        maps = maps.contiguous().float()
        out = torch.empty((n, oh, ow), dtype=torch.uint8, device=maps.device)
        ws = torch.empty((2, n, oh, ow), dtype=torch.float32, device=maps.device)
        mm = torch.empty((n, 2), dtype=torch.float32, device=maps.device)
        d = L.PostProc()
        d.x, d.N, d.H, d.W, d.oh, d.ow, d.blur = maps.data_ptr(), n, h, w, oh, ow, 1 if blur else 0
        d.ws0, d.ws1, d.minmax, d.out = ws[0].data_ptr(), ws[1].data_ptr(), mm.data_ptr(), out.data_ptr()
        lib.call("vinet_saliency_postprocess", C.byref(d), torch.cuda.current_stream(maps.device).cuda_stream)
        return out


class _StaticWindows(GraphedForward):
    """GraphedForward whose static input is an overlapping window view of a small frame store: a call copies the b + L - 1
    distinct frames of the batch once (device to device) instead of b * L."""

    def __init__(self, model, store, static_view, stem_windows=0, audio=None):
        self.store = store
        if stem_windows:      # the model consumes the frame store itself (per-frame stem re-use)
            self.model = _WindowsCall(model, stem_windows)
            self.inputs = [store]
        else:
            self.model = model
            self.inputs = [static_view]
        if audio is not None:  # audio-visual models: a second static input, copied per call
            self.inputs.append(audio.clone())
        self._capture(2)

    def stale(self):
        m = self.model.model if isinstance(self.model, _WindowsCall) else self.model
        return self._epoch != tuple(e.realloc_count for e in m.__dict__.get("_engines", {}).values())

    def _capture(self, warmup):
        super()._capture(warmup)
        m = self.model.model if isinstance(self.model, _WindowsCall) else self.model
        self._epoch = tuple(e.realloc_count for e in m.__dict__.get("_engines", {}).values())

    def __call__(self, x, a=None):
        _fill_store(self.store, x)
        if a is not None:
            self.inputs[-1].copy_(a, non_blocking=True)
        self.graph.replay()
        return self.out


class _WindowsCall:
    """Callable adapter: model.forward_windows(store, b) with the call signature GraphedForward replays."""

    def __init__(self, model, windows):
        self.model, self.windows = model, windows

    def __call__(self, store):
        return self.model.forward_windows(store, self.windows)


def _fill_store(store, x):
    """The b + L - 1 distinct frames behind the window view x: frame f is x[0, :, f] for f < L and x[f - L + 1, :, L - 1] after."""
    b, c, t, h, w = x.shape
    store[:t].copy_(x[0].permute(1, 0, 2, 3), non_blocking=True)
    if b > 1:
        store[t:].copy_(x[1:, :, t - 1], non_blocking=True)
