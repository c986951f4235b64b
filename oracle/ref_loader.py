"""TEST INFRASTRUCTURE ONLY — import the UNMODIFIED reference from /root/reference (build container only).

The reference is flat scripts; `model.py:5` imports an uninstalled, unused package (`block`), and
`model.py:224` loads `./soundnet8_final.pth` relative to the cwd.  Nothing is
copied: the modules are imported in place with a stub for `block` and a temporary chdir.
The GPU box has no /root/reference: callers must check `available()` first.
"""
import contextlib
import os
import sys
import types

REF_DIR = os.environ.get("VINET_REFERENCE_DIR", "/root/reference")
# build-time copy of the unmodified reference sources (git-ignored, made by __graft_entry__.build(); travels to the GPU box)
VENDORED_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "baseline", "_ref")


def available():
    return os.path.isfile(os.path.join(REF_DIR, "model.py"))


def use_dir(path):
    """Import the reference from `path` (e.g. baseline/_ref on the GPU box) instead of /root/reference."""
    global REF_DIR
    if os.path.abspath(path) != os.path.abspath(REF_DIR):
        REF_DIR = path
        _cache.clear()
        for name in ("model", "model_utils", "loss"):
            sys.modules.pop(name, None)


_cache = {}


@contextlib.contextmanager
def _cwd(path):
    old = os.getcwd()
    os.chdir(path)
    try:
        yield
    finally:
        os.chdir(old)


def load():
    """Returns (ref_model_module, ref_loss_module)."""
    if "m" in _cache:
        return _cache["m"], _cache["l"]
    if not available():
        raise RuntimeError("reference sources not present at %s" % REF_DIR)
    blk = types.ModuleType("block")
    blk.fusions = types.ModuleType("block.fusions")
    sys.modules.setdefault("block", blk)
    sys.modules.setdefault("block.fusions", blk.fusions)
    sys.path.insert(0, REF_DIR)
    try:
        with _cwd(REF_DIR):
            import importlib
            m = importlib.import_module("model")
            l = importlib.import_module("loss")
    finally:
        sys.path.remove(REF_DIR)
    _cache["m"], _cache["l"] = m, l
    return m, l


def build_vinet(num_clips=32, num_hier=3):
    m, _ = load()
    return m.VideoSaliencyModel(num_clips=num_clips, num_hier=num_hier)


def _with_random_soundnet(m, make, random_soundnet):
    """random_soundnet: the 57 MB soundnet8_final.pth is not vendored; model.py:147,224 load it relative to the cwd, so hand their
    torch.load a freshly initialised SoundNet state_dict instead (synthetic benchmarks and goldens only need the shapes)."""
    import torch
    with _cwd(REF_DIR):
        if random_soundnet and not os.path.isfile("soundnet8_final.pth"):
            real_load = torch.load
            torch.load = lambda *a, **k: m.SoundNet().state_dict()
            try:
                return make()
            finally:
                torch.load = real_load
        return make()


def build_avinet(random_soundnet=False, **kwargs):
    """VideoAudioSaliencyModel(**kwargs) of the unmodified reference (model.py:191), e.g. use_transformer=True."""
    m, _ = load()
    return _with_random_soundnet(m, lambda: m.VideoAudioSaliencyModel(**kwargs), random_soundnet)


def build_fusion(random_soundnet=False, **kwargs):
    """VideoAudioSaliencyFusionModel(**kwargs) of the unmodified reference (model.py:116)."""
    m, _ = load()
    return _with_random_soundnet(m, lambda: m.VideoAudioSaliencyFusionModel(**kwargs), random_soundnet)
