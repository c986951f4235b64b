"""CPU: SURVEY §8 row f3 — the restated input transform (oracle/preproc_oracle.py) and the product's resampling tables
(vinet_b200/preprocess.py) against PIL + torchvision themselves (the reference's transform, dataloader.py:242-249)."""
import numpy as np
import pytest
import torch

from oracle import preproc_oracle as PO
from vinet_b200.preprocess import resample_tables

SIZES = [(360, 640), (224, 384), (100, 150), (480, 853), (37, 61)]


@pytest.mark.parametrize("hw", SIZES)
def test_oracle_resize_is_pillow_bit_for_bit(hw):
    Image = pytest.importorskip("PIL.Image")
    img = np.random.default_rng(hw[0]).integers(0, 256, hw + (3,), dtype=np.uint8)
    ref = np.asarray(Image.fromarray(img).resize((384, 224), Image.BILINEAR))
    assert np.array_equal(PO.pil_resize_bilinear(img, (224, 384)), ref)


def test_oracle_transform_is_torchvision_bit_for_bit():
    Image = pytest.importorskip("PIL.Image")
    T = pytest.importorskip("torchvision.transforms")
    t = T.Compose([T.Resize((224, 384)), T.ToTensor(), T.Normalize([0.485, 0.456, 0.406], [0.229, 0.224, 0.225])])
    img = np.random.default_rng(1).integers(0, 256, (360, 640, 3), dtype=np.uint8)
    assert np.array_equal(t(Image.fromarray(img)).numpy(), PO.frame_transform(img))


@pytest.mark.parametrize("sizes", [(640, 384), (150, 384), (37, 224), (224, 224), (1080, 224)])
def test_product_tables_equal_the_oracle_tables(sizes):
    b, k = resample_tables(*sizes)
    ob, ok = PO.pillow_coeffs(*sizes)
    assert np.array_equal(b.numpy(), ob) and np.array_equal(k.numpy(), ok)


def test_audio_window_restatement():
    ex = np.random.default_rng(2).standard_normal(47041).astype(np.float32)
    for n in (47040, 47041, 2, 1):
        out = PO.audio_window(ex[:n])
        lo = 70560 // 2 - n // 2
        assert out.shape == (70560,) and np.count_nonzero(out[:lo]) == 0 and np.count_nonzero(out[lo + n:]) == 0
        assert np.allclose(out[lo:lo + n], np.hanning(n).astype(np.float32) * ex[:n])
