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
