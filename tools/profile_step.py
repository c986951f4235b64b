"""Per-kernel device-time table of one step (torch.profiler / CUPTI; concurrent, warm caches).
usage: profile_step.py [B] [vinet|avinet|avinet_xf|fusion] [train|eval]"""
import os
import sys
import collections
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from torch.profiler import profile, ProfilerActivity
from vinet_b200 import VideoAudioSaliencyFusionModel, VideoAudioSaliencyModel, VideoSaliencyModel, kldiv

for kv in os.environ.get("VINET_DEBUG_SET", "").split(","):      # e.g. VINET_DEBUG_SET=3=5 : atomic max-pool backward
    if "=" in kv:
        from vinet_b200 import lib as _L
        _L.get().call("vinet_debug_set", int(kv.split("=")[0]), int(kv.split("=")[1]))
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
which = sys.argv[2] if len(sys.argv) > 2 else "vinet"
mode = sys.argv[3] if len(sys.argv) > 3 else "train"
dev = torch.device("cuda")
torch.manual_seed(0)
model = {"vinet": lambda: VideoSaliencyModel(), "avinet": lambda: VideoAudioSaliencyModel(soundnet_weights=False),
         "avinet_xf": lambda: VideoAudioSaliencyModel(use_transformer=True, soundnet_weights=False),
         "fusion": lambda: VideoAudioSaliencyFusionModel(soundnet_weights=False)}[which]().to(dev)
model.train() if mode == "train" else model.eval()
opt = torch.optim.Adam([p for p in model.parameters() if p.requires_grad], lr=1e-4, fused=True)
x = torch.randn(B, 32, 3, 224, 384, device=dev)
gt = torch.rand(B, 224, 384, device=dev) + 1e-3
extra = [0.05 * torch.randn(B, 1, 70560, 1, device=dev)] if which != "vinet" else []


def step():
    if mode == "eval":
        with torch.no_grad():
            return model(x.permute(0, 2, 1, 3, 4), *extra)
    loss = kldiv(model(x.permute(0, 2, 1, 3, 4), *extra), gt)
    loss.backward()
    opt.step()
    opt.zero_grad(set_to_none=True)


for _ in range(3):
    step()
torch.cuda.synchronize()
import time
t0 = time.perf_counter()
for _ in range(3):
    step()
torch.cuda.synchronize()
print("wall ms/step %.2f" % ((time.perf_counter() - t0) / 3 * 1e3))
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    step()
    torch.cuda.synchronize()
tot = collections.defaultdict(float)
cnt = collections.Counter()
for ev in prof.events():
    if ev.device_type == torch.autograd.DeviceType.CUDA:
        name = ev.name if os.environ.get("VINET_PROFILE_FULLNAMES") and ev.name.startswith("void at::") else re.sub(r"<.*", "", re.sub(r"\(.*", "", ev.name))
        tot[name] += ev.device_time / 1e3 if hasattr(ev, "device_time") else ev.cuda_time / 1e3
        cnt[name] += 1
want = os.environ.get("VINET_PROFILE_LIST")
if want:
    for ev in prof.events():
        if ev.device_type == torch.autograd.DeviceType.CUDA and want in ev.name:
            print("  %-40s %8.1f us" % (re.sub(r"<.*", "", ev.name)[:40], ev.device_time if hasattr(ev, "device_time") else ev.cuda_time))
s = sum(tot.values())
print("device total ms %.2f over %d launches" % (s, sum(cnt.values())))
for k, v in sorted(tot.items(), key=lambda kv: -kv[1])[:40]:
    print("%-70s %5d %9.3f ms %5.1f%%" % (k[:70], cnt[k], v, 100 * v / s))
