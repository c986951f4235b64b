"""CPU: SURVEY §8 row f2 — host logic of the sliding-window inference driver (window views over one frame buffer, batching, the
time-flipped clips of the first L-1 frames) on the numpy kernel spec against the reference's per-clip loop
(generate_result.py:55-73) run with the oracle; and the numpy restatement of the post-processing against cv2 itself."""
import numpy as np
import pytest
import torch

from oracle import postproc_oracle as PO
from oracle import torch_oracle as O
from oracle.kernel_spec import Spec
from vinet_b200 import SlidingWindowSaliency, VideoSaliencyModel
from vinet_b200.inference import window_view


def test_window_view_is_a_view_of_the_frame_buffer():
    f = torch.arange(10 * 3 * 2 * 2, dtype=torch.float32).view(10, 3, 2, 2)
    v = window_view(f, 2, 3, 4)
    assert v.shape == (3, 3, 4, 2, 2) and v.data_ptr() == f[2].data_ptr()
    for b in range(3):
        for t in range(4):
            assert torch.equal(v[b, :, t], f[2 + b + t])


def test_sliding_window_matches_reference_loop():
    T, H, W, N = 8, 32, 32, 18
    ref = O.ViNetOracle(T)
    O.randomize_(ref, 12)
    ref.eval()
    m = VideoSaliencyModel(num_clips=T)
    m.load_state_dict(ref.state_dict())
    m.set_precision("fp32")
    m.__dict__["_backend"] = Spec()
    m.eval()
    g = torch.Generator().manual_seed(3)
    frames = torch.randn(N, 3, H, W, generator=g)
    want = PO.sliding_window_reference(ref, frames, T)
    got = SlidingWindowSaliency(m, clip_len=T, windows_per_batch=3, use_graph=False)(frames)
    assert got.shape == (N, H, W)
    assert torch.allclose(got, want, rtol=1e-3, atol=1e-5), (got - want).abs().max()
    with pytest.raises(ValueError):
        SlidingWindowSaliency(m, clip_len=T, use_graph=False)(frames[:2 * T - 2])


def test_postprocess_restatement_matches_cv2():
    cv2 = pytest.importorskip("cv2")
    g = np.random.default_rng(0)
    smap = (1 / (1 + np.exp(-3 * g.standard_normal((56, 96))))).astype(np.float32)
    for size in [(160, 90), (96, 56), (333, 201)]:
        r = cv2.resize(smap, size)
        assert np.allclose(PO.resize_bilinear(smap, *size), r, rtol=1e-5, atol=2e-6)
        b = cv2.GaussianBlur(r, (11, 11), 0)
        assert np.allclose(PO.blur11(r), b, rtol=1e-4, atol=1e-6)
        want = PO.to_uint8(b)
        got = PO.process(smap, size)
        assert np.abs(got.astype(int) - want.astype(int)).max() <= 1


def _reference_get_audio_feature():
    """The reference's OWN `get_audio_feature` (generate_result_audio_visual.py:86-118), executed from the reference tree (the
    module itself imports torchaudio / cv2 at the top, so only the function's source lines are compiled; nothing is copied)."""
    import os
    import sys
    from oracle import ref_loader
    path = os.path.join(ref_loader.REF_DIR, "generate_result_audio_visual.py")
    if not os.path.isfile(path):
        pytest.skip("reference tree not present (GPU box)")
    src = open(path).read().split("\n")
    a = next(i for i, l in enumerate(src) if l.startswith("def get_audio_feature"))
    b = next(i for i in range(a + 1, len(src)) if src[i].startswith("def "))
    ns = {"torch": torch, "np": np, "sys": sys}
    exec(compile("\n".join(src[a:b]), "generate_result_audio_visual.py[get_audio_feature]", "exec"), ns)
    return ns["get_audio_feature"]


def test_audio_track_bounds_and_feature_restatement_match_the_reference():
    """AudioTrack's excerpt arithmetic and the oracle's feature restatement against the executed reference function, over odd /
    even excerpt lengths, the start of the video (clamped at 0) and its last window."""
    from oracle import preproc_oracle as PR
    from vinet_b200 import AudioTrack
    ref_fn = _reference_get_audio_feature()

    class Args:
        clip_size = 32
    for fs, fps, n_frames in [(22050, 15, 70), (16000, 25, 64), (22050, 29.97, 90)]:
        n_wav = int(fs * n_frames / fps) + 7
        wav = np.random.default_rng(int(fs + n_frames)).standard_normal(n_wav).astype(np.float32)
        starts, ends = PR.av_excerpt_bounds(n_wav, fs, fps, n_frames)
        track = AudioTrack(torch.from_numpy(wav), fs, fps, n_frames)
        assert track.starts == list(starts) and track.ends == list(ends)
        data = {"v": {"wav": torch.from_numpy(wav).view(1, -1), "starts": starts, "ends": ends}}
        for idx in [0, 1, 2, 17, n_frames - 33, n_frames - 32]:
            want = ref_fn("v", data, Args, idx).view(-1).numpy()
            got = PR.av_audio_feature(wav, starts, ends, idx, 32)
            assert np.array_equal(got, want), (fs, fps, idx)
            s, e = track.bounds(idx, 32)
            assert (s, e) == (starts[idx + 1], min(ends[min(idx + 32, n_frames)] + 1, n_wav)) and np.count_nonzero(want) <= e - s


def test_audio_visual_driver_pairs_every_window_with_its_excerpt(monkeypatch):
    """Host logic of SlidingWindowSaliency for the audio-visual models (generate_result_audio_visual.py:177-199) with a stub model
    that reports which clip and which audio feature it was given: frame i >= L-1 is predicted from frames i-L+1 .. i with the
    excerpt of window i-L+1; frame j < L-1 from the time-flipped clip j .. j+L-1 with the time-flipped excerpt of window j."""
    from oracle import preproc_oracle as PR
    from vinet_b200 import AudioTrack
    T, N, fs, fps = 8, 21, 8000, 20.0
    wav = np.arange(1, int(fs * N / fps) + 1, dtype=np.float32)          # sample k holds k+1: an excerpt identifies its position
    starts, ends = PR.av_excerpt_bounds(wav.shape[0], fs, fps, N)

    def cpu_features(self, start_indices, clip_len, device, flip=False):  # AudioTrack.features without the CUDA kernel
        f = torch.stack([torch.from_numpy(PR.av_audio_feature(wav, starts, ends, j, clip_len)) for j in start_indices]).view(-1, 1, 70560, 1)
        return torch.flip(f, [2]) if flip else f
    monkeypatch.setattr(AudioTrack, "features", cpu_features)

    class Stub(torch.nn.Module):
        _n_extra = 1

        def __init__(self):
            super().__init__()
            self.w = torch.nn.Parameter(torch.zeros(1))

        def forward(self, x, a):
            out = torch.zeros(x.shape[0], 2, 4)
            out[:, 0, 0], out[:, 0, 1] = x[:, 0, 0, 0, 0], x[:, 0, -1, 0, 0]          # first / last frame of the clip as given
            nz = (a[:, 0, :, 0] != 0)
            first = nz.float().argmax(1)
            last = a.shape[2] - 1 - nz.flip(1).float().argmax(1)
            out[:, 0, 2], out[:, 0, 3] = a[torch.arange(len(a)), 0, first + 1, 0], a[torch.arange(len(a)), 0, last - 1, 0]
            out[:, 1, 0], out[:, 1, 1] = first.float(), last.float()
            return out
    frames = torch.arange(N, dtype=torch.float32).view(N, 1, 1, 1).expand(N, 3, 2, 4).contiguous()
    sal = SlidingWindowSaliency(Stub().eval(), clip_len=T, windows_per_batch=3, use_graph=False)
    got = sal(frames, audio=AudioTrack(torch.from_numpy(wav), fs, fps, N))
    for i in range(N):
        j, flipped = (i - T + 1, False) if i >= T - 1 else (i, True)
        f = PR.av_audio_feature(wav, starts, ends, j, T)
        f = f[::-1] if flipped else f
        nz = np.nonzero(f)[0]
        want_clip = (j + T - 1, j) if flipped else (j, j + T - 1)
        assert (got[i, 0, 0].item(), got[i, 0, 1].item()) == want_clip, i
        assert (got[i, 1, 0].item(), got[i, 1, 1].item()) == (nz[0], nz[-1]), i
        assert got[i, 0, 2].item() == f[nz[0] + 1] and got[i, 0, 3].item() == f[nz[-1] - 1], i
    with pytest.raises(AssertionError):
        sal(frames)
