"""GPU: SURVEY §8 row f3 — the device input pipeline against the reference's transform (PIL resize + ToTensor + Normalize):
byte work bit-exact, the normalised fp32 planes bit-exact; audio windowing within float rounding."""
import numpy as np
import pytest
import torch

from oracle import preproc_oracle as PO
from vinet_b200 import FramePreprocessor, audio_window

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("hw", [(360, 640), (224, 384), (100, 150), (480, 853), (37, 61)])
def test_frames_match_the_reference_transform_bit_for_bit(hw):
    g = np.random.default_rng(hw[1])
    frames = g.integers(0, 256, (3,) + hw + (3,), dtype=np.uint8)
    out = FramePreprocessor((224, 384))(torch.from_numpy(frames)).cpu().numpy()
    assert out.shape == (3, 3, 224, 384) and out.dtype == np.float32
    for i in range(3):
        assert np.array_equal(out[i], PO.frame_transform(frames[i])), i
    try:
        from PIL import Image
        from torchvision import transforms as T
    except ImportError:
        return
    t = T.Compose([T.Resize((224, 384)), T.ToTensor(), T.Normalize([0.485, 0.456, 0.406], [0.229, 0.224, 0.225])])
    assert np.array_equal(out[0], t(Image.fromarray(frames[0])).numpy())


def test_clip_batch_layout_feeds_the_model():
    """(B*T, h, w, 3) decoded frames -> (B, T, 3, H, W) -> the permuted (B, 3, T, H, W) view train.py:205 hands to the model."""
    from vinet_b200 import VideoSaliencyModel
    B, T = 2, 8
    frames = torch.from_numpy(np.random.default_rng(0).integers(0, 256, (B * T, 90, 160, 3), dtype=np.uint8))
    x = FramePreprocessor((64, 96))(frames).view(B, T, 3, 64, 96).permute(0, 2, 1, 3, 4)
    m = VideoSaliencyModel(num_clips=T).cuda().eval()
    with torch.no_grad():
        out = m(x)
    assert out.shape == (B, 64, 96) and torch.isfinite(out).all()


@pytest.mark.parametrize("n", [47040, 47041, 1])
def test_audio_window(n):
    ex = np.random.default_rng(n).standard_normal((2, n)).astype(np.float32) * 0.05
    out = audio_window(torch.from_numpy(ex).cuda()).cpu().numpy()
    assert out.shape == (2, 1, 70560, 1)
    for b in range(2):
        assert np.allclose(out[b, 0, :, 0], PO.audio_window(ex[b]), rtol=1e-6, atol=1e-9)
