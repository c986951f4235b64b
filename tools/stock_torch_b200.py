"""Stock PyTorch (ATen + cuDNN) on one B200: the existing Blackwell implementation of this path, timed beside ours.

BASELINE.md §4 step 4 / SURVEY.md §7 step 1: the oracle restatement (oracle/torch_oracle.py: exactly the reference's
torch.nn calls) fwd + kldiv + bwd + fused Adam at batch B of 32x224x384 clips in four modes:
  fp32      TF32 off - the literal reference arithmetic
  tf32      torch.backends.cudnn.allow_tf32 = True
  bf16      torch.autocast(bfloat16)
  bf16_cl   autocast + channels_last_3d weights/activations + cudnn.benchmark
Also times the eval forward (no_grad) per mode.  Writes one JSON document (argv[1], default stdout).
TEST/MEASUREMENT INFRASTRUCTURE: uses oracle/ as the stock-PyTorch model, never the product path.
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

from oracle import torch_oracle as O

T, H, W = 32, 224, 384
FWD_BWD_GFLOP, FWD_GFLOP = 675.006, 229.318


def run_mode(mode, B, steps, warmup):
    torch.backends.cudnn.allow_tf32 = mode != "fp32"
    torch.backends.cuda.matmul.allow_tf32 = mode != "fp32"
    torch.backends.cudnn.benchmark = mode == "bf16_cl"
    dev = torch.device("cuda")
    torch.manual_seed(0)
    m = O.ViNetOracle(T).to(dev).train()
    if mode == "bf16_cl":
        m = m.to(memory_format=torch.channels_last_3d)
    opt = torch.optim.Adam(m.parameters(), lr=1e-4, fused=True)
    g = torch.Generator().manual_seed(1234)
    x = torch.randn(B, T, 3, H, W, generator=g).to(dev).permute(0, 2, 1, 3, 4)
    gt = (torch.rand(B, H, W, generator=g) + 1e-3).to(dev)
    if mode == "bf16_cl":
        x = x.contiguous(memory_format=torch.channels_last_3d)
    cast = torch.autocast("cuda", dtype=torch.bfloat16, enabled=mode.startswith("bf16"))

    def train_step():
        with cast:
            pred = m(x)
        loss = O.kldiv(pred.float(), gt)
        loss.backward()
        opt.step()
        opt.zero_grad(set_to_none=True)
        return loss

    def timed(fn, n):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n

    res = {"mode": mode, "batch": B}
    try:
        for _ in range(warmup):
            train_step()
        ms = timed(train_step, steps)
        res.update({"train_ms_per_step": ms, "train_clips_per_s": B / ms * 1e3, "train_tflops": B * FWD_BWD_GFLOP / ms,
                    "train_peak_mem_gb": torch.cuda.max_memory_allocated() / 2 ** 30})
        m.eval()

        def eval_step():
            with torch.no_grad(), cast:
                return m(x[:1])
        for _ in range(warmup):
            eval_step()
        ms1 = timed(eval_step, steps)
        res.update({"eval_b1_ms": ms1, "eval_b1_clips_per_s": 1e3 / ms1})
    except Exception as ex:          # e.g. out of memory in fp32 at B=8: record and go on
        res["error"] = str(ex)[:300]
    del m, opt
    torch.cuda.empty_cache()
    torch.cuda.reset_peak_memory_stats()
    return res


def main():
    out = sys.argv[1] if len(sys.argv) > 1 else ""
    B = int(sys.argv[2]) if len(sys.argv) > 2 else 8
    t0 = time.time()
    doc = {"what": "stock PyTorch (oracle restatement = the reference's torch.nn calls) on one B200, ViNet 32x224x384, "
                   "fwd + kldiv + bwd + fused Adam", "torch": torch.__version__, "cudnn": torch.backends.cudnn.version(),
           "gpu": torch.cuda.get_device_name(0), "modes": []}
    for mode in ("bf16_cl", "bf16", "tf32", "fp32"):
        doc["modes"].append(run_mode(mode, B, steps=5, warmup=3))
        print(json.dumps(doc["modes"][-1]), flush=True)
    doc["wall_s"] = time.time() - t0
    if out:
        json.dump(doc, open(out, "w"), indent=1)


if __name__ == "__main__":
    main()
