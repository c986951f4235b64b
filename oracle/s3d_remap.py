"""TEST INFRASTRUCTURE ONLY — restatement of the S3D Kinetics-400 checkpoint remap done by the reference's training script
(/root/reference/train.py:141-172): source keys ``[module.]base.N.rest`` become ``base{g}.{N - sn}.rest`` with
(g, sn) = (1, 0) for N < 5, (2, 5) for 5 <= N < 8, (3, 8) for 8 <= N < 14, (4, 14) for N >= 14; a tensor is copied only when
the destination key exists with the same shape.  The caller then does ``model.backbone.load_state_dict(model_dict)``.
"""
import torch

STAGE_STARTS = [0, 5, 8, 14]          # train.py:151


def remap_key(name):
    if "module" in name:
        name = ".".join(name.split(".")[1:])          # train.py:147-148
    if "base." in name:
        parts = name.split(".")
        bn = int(parts[1])
        g = max(i for i, s in enumerate(STAGE_STARTS) if bn >= s)
        name = "base%d.%d." % (g + 1, bn - STAGE_STARTS[g]) + ".".join(parts[2:])
    return name


def load_s3d_weights(backbone, weight_dict):
    """Returns (copied, skipped_size, skipped_name) key lists, after loading into `backbone` like train.py:145-170."""
    model_dict = backbone.state_dict()
    copied, bad_size, bad_name = [], [], []
    with torch.no_grad():
        for name, param in weight_dict.items():
            key = remap_key(name)
            if key in model_dict:
                if param.size() == model_dict[key].size():
                    model_dict[key].copy_(param)
                    copied.append(key)
                else:
                    bad_size.append(key)
            else:
                bad_name.append(key)
    backbone.load_state_dict(model_dict)
    return copied, bad_size, bad_name


def fake_s3d_checkpoint(backbone, seed=0, prefix="module."):
    """A synthetic checkpoint in the S3D_kinetics400.pt key layout (flat ``base.N``), derived from a backbone's own keys, plus
    the classifier head the real file carries (which the remap must skip) and one tensor of the wrong shape."""
    g = torch.Generator().manual_seed(seed)
    out = {}
    for k, v in backbone.state_dict().items():
        stage, idx, rest = k.split(".", 2)
        n = STAGE_STARTS[int(stage[4:]) - 1] + int(idx)
        t = torch.randn(v.shape, generator=g) if v.is_floating_point() else torch.full_like(v, 7)
        out["%sbase.%d.%s" % (prefix, n, rest)] = t
    out[prefix + "fc.0.weight"] = torch.randn(400, 1024, 1, 1, 1, generator=g)
    first = prefix + "base.0.conv_s.weight"
    out[first.replace("conv_s", "conv_extra")] = torch.randn(3, generator=g)
    return out
