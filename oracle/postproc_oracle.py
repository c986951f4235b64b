"""TEST INFRASTRUCTURE ONLY — numpy restatement of the reference's saliency-map post-processing
(/root/reference/generate_result.py:96-104 `process`, utils.py:61-78 `blur` / `img_save`):
cv2.resize (bilinear, half-pixel centres, float path) -> cv2.GaussianBlur((11,11), 0) (sigma 2, BORDER_REFLECT_101) ->
torchvision make_grid(normalize=True) ((x - min) / (max - min + 1e-5)) -> round(255 x + 0.5) clamped to uint8.
Checked against cv2 itself by tests/test_inference_cpu.py when cv2 is importable."""
import numpy as np


def resize_bilinear(img, ow, oh):
    h, w = img.shape
    def taps(n_out, n_in):
        f = (np.arange(n_out, dtype=np.float64) + 0.5) * (n_in / n_out) - 0.5      # cv2 computes the source coordinate in double
        i0 = np.floor(f).astype(np.int64)
        fr = (f - i0).astype(np.float32)
        lo = i0 < 0
        i0[lo] = 0; fr[lo] = 0
        hi = i0 >= n_in - 1
        i0[hi] = n_in - 1; fr[hi] = 0
        return i0, np.minimum(i0 + 1, n_in - 1), fr
    x0, x1, fx = taps(ow, w)
    y0, y1, fy = taps(oh, h)
    top = img[y0][:, x0] * (1 - fx) + img[y0][:, x1] * fx
    bot = img[y1][:, x0] * (1 - fx) + img[y1][:, x1] * fx
    return (top * (1 - fy[:, None]) + bot * fy[:, None]).astype(np.float32)


def gaussian_kernel(k=11):
    sigma = 0.3 * ((k - 1) * 0.5 - 1) + 0.8
    x = np.arange(k, dtype=np.float64) - (k - 1) / 2
    g = np.exp(-x * x / (2 * sigma * sigma))
    return (g / g.sum()).astype(np.float32)


def blur11(img):
    g = gaussian_kernel(11)
    p = np.pad(img, 5, mode="reflect")          # numpy 'reflect' == cv2 BORDER_REFLECT_101
    tmp = sum(g[k] * p[5:-5, k:k + img.shape[1]] for k in range(11))
    p = np.pad(tmp, ((5, 5), (0, 0)), mode="reflect")
    return sum(g[k] * p[k:k + img.shape[0], :] for k in range(11)).astype(np.float32)


def to_uint8(img):
    lo, hi = img.min(), img.max()
    v = (np.clip(img, lo, hi) - lo) / (hi - lo + np.float32(1e-5))
    return np.rint(np.clip(v * 255 + 0.5, 0, 255)).astype(np.uint8)


def process(smap, size_wh, blur=True):
    r = resize_bilinear(np.asarray(smap, np.float32), size_wh[0], size_wh[1])
    return to_uint8(blur11(r) if blur else r)


def sliding_window_reference(model, frames, clip_len):
    """generate_result.py:55-73 with a torch module `model` (the oracle): returns (N,H,W) maps, one model call per clip."""
    import torch
    n = frames.shape[0]
    out = torch.zeros((n,) + tuple(frames.shape[2:]))
    with torch.no_grad():
        for i in range(clip_len - 1, n):
            clip = frames[i - clip_len + 1:i + 1].unsqueeze(0).permute(0, 2, 1, 3, 4)
            out[i] = model(clip)[0]
            if i < 2 * clip_len - 2:
                out[i - clip_len + 1] = model(torch.flip(clip, [2]))[0]
    return out
