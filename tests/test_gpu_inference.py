"""GPU: SURVEY §8 row f2 — sliding-window inference through the real kernels (folded-BatchNorm eval plan, overlapping window
views, CUDA-graph replay) against the reference's per-clip loop (generate_result.py:55-73) run with the oracle, and the
device post-processing (generate_result.py:100-104) against its numpy / cv2 restatement."""
import numpy as np
import pytest
import torch

from oracle import postproc_oracle as PO
from oracle import torch_oracle as O
from vinet_b200 import GraphedForward, SlidingWindowSaliency, VideoSaliencyModel

pytestmark = pytest.mark.gpu


def _pair(T, seed, precision):
    ref = O.ViNetOracle(T)
    O.randomize_(ref, seed)
    ref.eval()
    m = VideoSaliencyModel(num_clips=T)
    m.load_state_dict(ref.state_dict())
    return ref, m.cuda().set_precision(precision).eval()


@pytest.mark.parametrize("precision,use_graph", [("fp32", False), ("fp32", True), ("bf16x6", True)])
def test_sliding_window_matches_reference_loop(precision, use_graph):
    T, H, W, N = 8, 64, 96, 21
    ref, m = _pair(T, 13, precision)
    frames = torch.randn(N, 3, H, W, generator=torch.Generator().manual_seed(4))
    want = PO.sliding_window_reference(ref, frames, T)
    sal = SlidingWindowSaliency(m, clip_len=T, windows_per_batch=4, use_graph=use_graph)
    got = sal(frames).cpu()
    assert torch.allclose(got, want, rtol=1e-3, atol=1e-6), ((got - want).abs() / want.abs()).max()
    got2 = sal(frames.cuda()).cpu()                      # second video through the captured graphs, frames already on the device
    assert torch.equal(got, got2)


def test_sliding_window_bf16_throughput_mode_tracks_fp32():
    T, H, W, N = 32, 64, 96, 70
    ref, m = _pair(T, 14, "bf16")
    frames = torch.randn(N, 3, H, W, generator=torch.Generator().manual_seed(5))
    got = SlidingWindowSaliency(m, clip_len=T, windows_per_batch=8)(frames).cpu()
    _, m32 = _pair(T, 14, "fp32")
    want = SlidingWindowSaliency(m32, clip_len=T, windows_per_batch=8, use_graph=False)(frames).cpu()
    assert torch.isfinite(got).all() and (got - want).abs().mean() < 2e-2, (got - want).abs().mean()


def test_graphed_forward_replays_eval_forward():
    T, H, W = 16, 64, 64
    ref, m = _pair(T, 15, "fp32")
    d = O.make_inputs(2, T, H, W, 15)
    x = d["x"].cuda()
    with torch.no_grad():
        want = ref(d["x"])
    fwd = GraphedForward(m, x)
    assert fwd.launches_per_replay > 50
    got = fwd(x).cpu()
    assert torch.allclose(got, want, rtol=1e-3, atol=1e-6)
    x2 = O.make_inputs(2, T, H, W, 16)["x"]
    with torch.no_grad():
        want2 = ref(x2)
    assert torch.allclose(fwd(x2.cuda()).cpu(), want2, rtol=1e-3, atol=1e-6)


@pytest.mark.parametrize("size", [(160, 90), (640, 360), (333, 201)])
def test_postprocess_matches_reference_pipeline(size):
    g = np.random.default_rng(1)
    maps = (1 / (1 + np.exp(-3 * g.standard_normal((3, 56, 96))))).astype(np.float32)
    _, m = _pair(8, 1, "fp32")
    sal = SlidingWindowSaliency(m, clip_len=8)
    got = sal.postprocess(torch.from_numpy(maps).cuda(), size).cpu().numpy()
    assert got.shape == (3, size[1], size[0]) and got.dtype == np.uint8
    for i in range(3):
        want = PO.process(maps[i], size)
        assert np.abs(got[i].astype(int) - want.astype(int)).max() <= 1, i
        assert (got[i] != want).mean() < 0.02
    try:
        import cv2
    except ImportError:
        return
    want = PO.to_uint8(cv2.GaussianBlur(cv2.resize(maps[0], size), (11, 11), 0))
    assert np.abs(got[0].astype(int) - want.astype(int)).max() <= 1


def test_per_frame_stem_reuse_matches_plain_windows():
    """model.forward_windows (stem conv_s once per frame, temporal stem conv over overlapping window views with a one-frame batch
    pitch) must reproduce the plain batched-clip forward of the same bf16 engine bit for bit, eager and graph-replayed."""
    from vinet_b200.inference import window_view
    T, H, W, b = 32, 64, 96, 5
    _, m = _pair(T, 17, "bf16")
    frames = torch.randn(b + T - 1, 3, H, W, generator=torch.Generator().manual_seed(6)).cuda()
    with torch.no_grad():
        want = m(window_view(frames, 0, b, T)).clone()
        got = m.forward_windows(frames, b).clone()
    assert torch.equal(got, want), (got - want).abs().max()
    N = 70
    video = torch.randn(N, 3, H, W, generator=torch.Generator().manual_seed(7))
    a = SlidingWindowSaliency(m, clip_len=T, windows_per_batch=8, stem_cache=True)(video).cpu()
    c = SlidingWindowSaliency(m, clip_len=T, windows_per_batch=8, stem_cache=False)(video).cpu()
    assert torch.equal(a, c), (a - c).abs().max()


def test_audio_visual_sliding_window_matches_reference_loop():
    """generate_result_audio_visual.py:177-199 through the driver: every window gets the excerpt the reference cuts for it
    (Hanning-windowed on the device), the first L-1 frames come from time-flipped clips WITH time-flipped audio, windows are
    batched and graph-replayed.  Checked against the reference's loop (oracle restatement, pinned on the CPU against the reference's
    own get_audio_feature) calling the same model one clip at a time."""
    from oracle import preproc_oracle as PR
    from vinet_b200 import AudioTrack, VideoAudioSaliencyModel
    T, N, fs, fps = 32, 66, 22050, 15.0
    ref = O.AViNetOracle(T)
    O.randomize_(ref, 21)
    m = VideoAudioSaliencyModel(num_clips=T, soundnet_weights=False)
    m.load_state_dict(ref.state_dict())
    m = m.cuda().set_precision("fp32").eval()
    g = torch.Generator().manual_seed(6)
    frames = torch.randn(N, 3, 224, 384, generator=g)
    wav = (0.05 * torch.randn(int(fs * N / fps) + 5, generator=g)).numpy()
    starts, ends = PR.av_excerpt_bounds(wav.shape[0], fs, fps, N)
    track = AudioTrack(torch.from_numpy(wav), fs, fps, N)
    dev = torch.device("cuda", torch.cuda.current_device())
    idx = [0, 3, N - T]
    feat = track.features(idx, T, dev).cpu().numpy()
    rev = track.features(idx, T, dev, flip=True).cpu().numpy()
    for i, j in enumerate(idx):
        want = PR.av_audio_feature(wav, starts, ends, j, T)
        assert np.allclose(feat[i, 0, :, 0], want, rtol=1e-6, atol=1e-9) and np.allclose(rev[i, 0, :, 0], want[::-1], rtol=1e-6, atol=1e-9)
    sal = SlidingWindowSaliency(m, clip_len=T, windows_per_batch=4)
    got = sal(frames, audio=track).cpu()
    assert got.shape == (N, 224, 384) and torch.isfinite(got).all()
    ids = [0, 7, T - 2, T - 1, T + 9, N - 1]            # flipped clips (first L-1 frames), the first forward window, the last frame
    want = PR.sliding_window_reference_av(lambda c, a: m(c.cuda(), a.cuda()).cpu(), frames, T, wav, starts, ends, frame_ids=ids)
    assert sorted(want) == ids
    for i in ids:
        assert torch.allclose(got[i], want[i], rtol=1e-4, atol=1e-6), (i, (got[i] - want[i]).abs().max())
    with pytest.raises(AssertionError):
        sal(frames)                                     # an audio-visual model needs its sound track
