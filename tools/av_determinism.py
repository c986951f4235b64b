"""Development aid: run-to-run difference of two identical eval forwards of the audio-visual models, per precision mode."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from oracle import torch_oracle as O
from vinet_b200 import VideoAudioSaliencyFusionModel, VideoAudioSaliencyModel, VideoSaliencyModel

d = O.make_inputs(1, 32, 224, 384, 5, audio=True)
x, a = d["x"].cuda(), d["audio"].cuda()
def rnd(m, ref):
    """Seeded non-trivial weights (the default initialisation gives a near-constant map that hides everything)."""
    m.load_state_dict(O.randomize_(ref, 3).state_dict())
    return m


for name, make in [("vinet", lambda: rnd(VideoSaliencyModel(), O.ViNetOracle(32))),
                   ("avinet", lambda: rnd(VideoAudioSaliencyModel(soundnet_weights=False), O.AViNetOracle(32))),
                   ("avinet_xf", lambda: rnd(VideoAudioSaliencyModel(use_transformer=True, soundnet_weights=False), O.AViNetOracle(32, use_transformer=True))),
                   ("fusion", lambda: rnd(VideoAudioSaliencyFusionModel(soundnet_weights=False), O.AVFusionOracle()))]:
    for prec in ("fp32", "bf16"):
        torch.manual_seed(0)
        m = make().cuda().set_precision(prec).eval()
        args = (x,) if name == "vinet" else (x, a)
        with torch.no_grad():
            outs = [m(*args).clone() for _ in range(4)]
        e = m._engine_for(x.device)
        aud = e.pool.get("audionet.o7")
        diffs = [(outs[0] - o).abs().max().item() for o in outs[1:]]
        print("%-10s %-5s max|run0 - run_i| = %s   (map range %.3f..%.3f)" % (name, prec, ["%.2e" % v for v in diffs], outs[0].min().item(), outs[0].max().item()))
