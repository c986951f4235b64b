"""TEST INFRASTRUCTURE ONLY — restatement of the reference's per-frame input transform (/root/reference/dataloader.py:242-249,
generate_result.py:77-89): transforms.Resize((224, 384)) on a PIL image -> ToTensor -> Normalize(ImageNet mean / std), and of
the audio excerpt windowing (dataloader.py:89-122).

The resize arithmetic lives in a third-party dependency that is not under /root/reference: Pillow (`Pillow==7.1.2` in
requirements.txt; 12.x installed here, same algorithm): src/libImaging/Resample.c, bilinear filter with antialiasing,
8-bit fixed-point two-pass convolution (PRECISION_BITS = 22, horizontal pass first, 8-bit intermediate).  Pinned by
tests/test_preprocess_cpu.py against PIL + torchvision themselves."""
import numpy as np

PRECISION_BITS = 32 - 8 - 2
MEAN = np.array([0.485, 0.456, 0.406], np.float32)
STD = np.array([0.229, 0.224, 0.225], np.float32)


def pillow_coeffs(in_size, out_size):
    """Resample.c precompute_coeffs + normalize_coeffs_8bpc for the bilinear filter (support 1): (bounds [out,2], kk [out,ksize])."""
    scale = in_size / out_size
    filterscale = max(scale, 1.0)
    support = 1.0 * filterscale
    ksize = int(np.ceil(support)) * 2 + 1
    bounds = np.zeros((out_size, 2), np.int32)
    kk = np.zeros((out_size, ksize), np.int32)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        xmin = max(int(center - support + 0.5), 0)
        xmax = min(int(center + support + 0.5), in_size) - xmin
        x = np.arange(xmax)
        w = np.maximum(1.0 - np.abs((x + xmin - center + 0.5) * ss), 0.0)
        ww = w.sum()
        if ww != 0.0:
            w = w / ww
        kk[xx, :xmax] = np.where(w < 0, (-0.5 + w * (1 << PRECISION_BITS)).astype(np.int64), (0.5 + w * (1 << PRECISION_BITS)).astype(np.int64))
        bounds[xx] = (xmin, xmax)
    return bounds, kk


def _pass(img, bounds, kk, axis):
    """One 8bpc pass along `axis` of an (h, w, 3) uint8 image."""
    src = np.moveaxis(img.astype(np.int64), axis, 0)
    out = np.empty((bounds.shape[0],) + src.shape[1:], np.uint8)
    for xx, (xmin, xmax) in enumerate(bounds):
        acc = (1 << (PRECISION_BITS - 1)) + np.tensordot(kk[xx, :xmax].astype(np.int64), src[xmin:xmin + xmax], axes=(0, 0))
        out[xx] = np.clip(acc >> PRECISION_BITS, 0, 255)
    return np.moveaxis(out, 0, axis)


def pil_resize_bilinear(img, out_hw):
    """img: (h, w, 3) uint8 -> (H, W, 3) uint8, what PIL.Image.resize((W, H), BILINEAR) returns."""
    h, w = img.shape[:2]
    H, W = out_hw
    if w != W:
        img = _pass(img, *pillow_coeffs(w, W), axis=1)
    if h != H:
        img = _pass(img, *pillow_coeffs(h, H), axis=0)
    return img


def frame_transform(img, out_hw=(224, 384)):
    """(h, w, 3) uint8 RGB -> (3, H, W) fp32: Resize -> ToTensor -> Normalize (dataloader.py:242-249)."""
    r = pil_resize_bilinear(img, out_hw).astype(np.float32) / np.float32(255)
    return ((r - MEAN) / STD).transpose(2, 0, 1).astype(np.float32)


def audio_window(excerpt, total=70560):
    """dataloader.py:113-118: the excerpt times np.hanning(n), centred in a zero buffer of `total` samples."""
    n = excerpt.shape[0]
    out = np.zeros(total, np.float32)
    lo = total // 2 - n // 2
    out[lo:lo + n] = np.hanning(n).astype(np.float32) * excerpt.astype(np.float32)
    return out


def av_excerpt_bounds(n_wav, sample_rate, fps, n_frames):
    """generate_result_audio_visual.py:56-66 (= dataloader.py:56-70): per-frame audio excerpt start / end sample (1-based frames)."""
    fs, fps = float(sample_rate), float(fps)
    n_samples = fs / fps
    starts, ends = np.zeros(n_frames + 1, dtype=int), np.zeros(n_frames + 1, dtype=int)
    for vf in range(1, n_frames + 1):
        starts[vf] = int(max(0, ((vf - 1) * (1.0 / fps) * fs) - n_samples / 2))
        ends[vf] = int(min(n_wav, abs(((vf - 1) * (1.0 / fps) * fs) + n_samples / 2)))
    return starts, ends


def av_audio_feature(wav, starts, ends, start_idx, clip_len, total=70560):
    """generate_result_audio_visual.py:86-118 (`get_audio_feature`) for a 1-D waveform: the excerpt of the window whose first frame
    is `start_idx`, times np.hanning, centred in `total` zeros -> (total,) fp32."""
    s = starts[start_idx + 1]
    e = ends[-1] if start_idx + clip_len >= len(ends) else ends[start_idx + clip_len]
    return audio_window(np.asarray(wav[s:e + 1], np.float32), total)


def sliding_window_reference_av(model, frames, clip_len, wav, starts, ends, frame_ids=None):
    """generate_result_audio_visual.py:177-199 with a callable `model(clip, audio)`: {frame index: (H, W) map} for the frames in
    `frame_ids` (all when None); the first L-1 frames come from the time-flipped clip AND the time-flipped audio feature."""
    import torch
    n = frames.shape[0]
    out = {}
    with torch.no_grad():
        for i in range(clip_len - 1, n):
            j = i - clip_len + 1
            want_fwd = frame_ids is None or i in frame_ids
            want_rev = i < 2 * clip_len - 2 and (frame_ids is None or j in frame_ids)
            if not (want_fwd or want_rev):
                continue
            clip = frames[j:i + 1].unsqueeze(0).permute(0, 2, 1, 3, 4)
            a = torch.from_numpy(av_audio_feature(wav, starts, ends, j, clip_len)).view(1, 1, -1, 1).to(frames.device)
            if want_fwd:
                out[i] = model(clip, a)[0]
            if want_rev:
                out[j] = model(torch.flip(clip, [2]), torch.flip(a, [2]))[0]
    return out
