"""CPU: SURVEY §8 row f4 — the S3D Kinetics checkpoint remap of the reference's training script (train.py:141-172) against the
drop-in's ``backbone`` attribute tree: (a) the restated remap (oracle/s3d_remap.py), (b) when /root/reference is present
(build container), the reference's OWN source lines executed unchanged with the drop-in model in scope."""
import os

import pytest
import torch

from oracle import ref_loader, s3d_remap
from oracle import torch_oracle as O
from vinet_b200 import VideoSaliencyModel


def _check_loaded(model, ckpt):
    sd = model.backbone.state_dict()
    n = 0
    for k, v in ckpt.items():
        key = s3d_remap.remap_key(k)
        if key in sd and sd[key].shape == v.shape:
            assert torch.equal(sd[key], v.to(sd[key].dtype)), key
            n += 1
    assert n == len(sd) == 462          # every backbone tensor (462 of the 470 state_dict entries) came from the checkpoint


def test_s3d_remap_restated():
    m = VideoSaliencyModel()
    ref = O.ViNetOracle()
    assert list(m.backbone.state_dict().keys()) == list(ref.backbone.state_dict().keys())
    ckpt = s3d_remap.fake_s3d_checkpoint(ref.backbone, seed=3)
    copied, bad_size, bad_name = s3d_remap.load_s3d_weights(m.backbone, ckpt)
    assert len(copied) == 462 and not bad_size and sorted(bad_name) == ["base1.0.conv_extra.weight", "fc.0.weight"]
    _check_loaded(m, ckpt)
    # the oracle (= reference layout) accepts the same file the same way
    c2, _, n2 = s3d_remap.load_s3d_weights(ref.backbone, ckpt)
    assert c2 == copied and n2 == bad_name


@pytest.mark.skipif(not ref_loader.available(), reason="needs /root/reference (build container only)")
def test_reference_train_py_remap_block_runs_unchanged(tmp_path, capsys):
    """Lines 141-172 of the reference's train.py, executed verbatim (read from the reference tree at test time, nothing copied)
    with `model` bound to the drop-in VideoSaliencyModel and `file_weight` to a synthetic S3D checkpoint."""
    src = open(os.path.join(ref_loader.REF_DIR, "train.py")).read().split("\n")
    block = "\n".join(src[140:172])
    assert block.lstrip().startswith("if not (args.use_sound or args.use_vox):") and "model.backbone.load_state_dict(model_dict)" in block
    m = VideoSaliencyModel()
    ckpt = s3d_remap.fake_s3d_checkpoint(O.ViNetOracle().backbone, seed=5)
    f = str(tmp_path / "S3D_kinetics400.pt")
    torch.save(ckpt, f)

    class Args:
        use_sound, use_vox = False, False
    exec(compile(block, "train.py[141:172]", "exec"), {"args": Args, "model": m, "file_weight": f, "torch": torch, "os": os})
    out = capsys.readouterr().out
    assert "loading weight file" in out and " loaded" in out and " name? fc.0.weight" in out and "size?" not in out
    _check_loaded(m, ckpt)
