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
