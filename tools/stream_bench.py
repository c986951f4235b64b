"""Frames/s of the sliding-window inference driver on a synthetic video, against the reference's
"real-time (60 fps)" claim (README.md:27) and its one-forward-per-frame loop.
usage: stream_bench.py [frames N=256] [windows per batch=8] [precision=bf16]"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from vinet_b200 import SlidingWindowSaliency, VideoSaliencyModel

N = int(sys.argv[1]) if len(sys.argv) > 1 else 256
B = int(sys.argv[2]) if len(sys.argv) > 2 else 8
prec = sys.argv[3] if len(sys.argv) > 3 else "bf16"
torch.manual_seed(0)
m = VideoSaliencyModel().cuda().set_precision(prec).eval()
frames = torch.randn(N, 3, 224, 384).pin_memory()
sal = SlidingWindowSaliency(m, 32, B)
out = sal(frames)                      # warm-up: captures the graphs
torch.cuda.synchronize()
res = {"frames": N, "windows_per_batch": B, "precision": prec, "clip": "32x224x384"}
for what, src in (("frames_from_pinned_host", frames), ("frames_on_device", frames.cuda())):
    t0 = time.perf_counter()
    maps = sal(src)
    png = sal.postprocess(maps, (640, 360))
    host = png.cpu()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    res[what] = {"seconds": dt, "frames_per_s": N / dt, "forward_clips": N - 31 + 31, "includes": "H2D of the frames, all window forwards, "
                 "resize+blur+normalise on the device, D2H of the 8-bit maps"}
# the reference's loop: one forward per output frame, one clip per call (batch 1), no post-processing on the device
one = SlidingWindowSaliency(m, 32, 1)
one(frames[:64])
torch.cuda.synchronize()
t0 = time.perf_counter()
one(frames[:96])
torch.cuda.synchronize()
dt = time.perf_counter() - t0
res["one_window_per_forward"] = {"seconds": dt, "frames_per_s": 96 / dt}
res["reference_claim"] = "real-time (60 fps), README.md:27; 0.016 s/frame on a Titan X (DHF1K leaderboard)"
print(json.dumps(res))
