"""bench.py — clips/sec (32x224x384) fwd+bwd(+Adam) of ViNet on N B200s; one JSON line on rank 0.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--batch 8] [--precision bf16]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A step is one training pass over one batch of synthetic clips (BASELINE.json config 2 at N=1: batch 8 of
32x224x384, bf16): forward, kldiv loss, backward, (N>1: ONE NCCL all-reduce of the flat gradient arena), fused Adam.
`value` has inputs resident in HBM; `e2e` goes through the public module API from pinned host buffers with
the H2D copy of the clip + gt and the D2H read of the loss inside the timed region.
`--impl reference` times the reference's own CPU implementation path (the PyTorch oracle restatement,
oracle/torch_oracle.py, pinned to the executed reference) on the host cores, one clip per step.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FWD_BWD_GFLOP = 675.006      # conv MACs x2 per 32x224x384 clip, fwd+bwd (SURVEY.md §8d / BASELINE.md §2)
T, H, W = 32, 224, 384


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        d = json.load(open(p))
        return d.get("bf16_tflops", 1590.0), d.get("bf16_tflops_sustained", 1400.0), d.get("hbm_gbs", 6650.0), "measured"
    return 1590.0, 1400.0, 6650.0, "fallback"


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.stop_flag = index, [], False

    def run(self):
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.samples.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        sm = sorted(float(s[0]) for s in self.samples if s[0].replace(".", "").isdigit())
        mx = [float(s[1]) for s in self.samples if s[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(s) > 3 + i and s[3 + i].lower().startswith("active") for s in self.samples)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(self.samples)}


def cpu_baseline(steps=2, warmup=1):
    """The oracle (PyTorch restatement of the reference path) fwd+bwd on the host cores, B=1."""
    import torch
    from oracle import torch_oracle as O
    torch.set_num_threads(os.cpu_count())
    m = O.ViNetOracle(32).train()
    d = O.make_inputs(1, T, H, W, 0)
    ts = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        for p in m.parameters():
            p.grad = None
        loss = O.kldiv(m(d["x"]), d["gt"])
        loss.backward()
        ts.append(time.perf_counter() - t0)
    ts = sorted(ts[warmup:])
    med = ts[len(ts) // 2]
    return {"value": 1.0 / med, "unit": "clips/s", "cores": os.cpu_count(), "kind": "port",
            "sample": "%d timed fwd+bwd iterations of 1 clip 32x224x384 fp32 (median %.2f s), torch %s, %d threads"
                      % (steps, med, torch.__version__, torch.get_num_threads())}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    t0 = time.perf_counter()
    cb = cpu_baseline(steps=max(1, args.steps), warmup=min(args.warmup, 1))
    ms = 1000.0 / cb["value"]
    line = {"impl": "reference", "metric": "clips/sec (32x224x384) fwd+bwd", "value": cb["value"], "unit": "clips/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "ViNet fwd+bwd, 32x224x384 clips, CPU reference path, 1 clip per step"},
            "cpu_baseline": cb, "e2e": {"value": cb["value"], "unit": "clips/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "wall_s": time.perf_counter() - t0}
    print(json.dumps(line))


def kernel_roofline(model, x, gt, kldiv, torch):
    """Time every conv launch of one step with CUDA events on the launching stream; returns the per-kernel
    table and the dominant (largest-FLOP) launch."""
    eng = model._engine_for(x.device)
    eng.profile = []
    flush = torch.empty(192 * 1024 * 1024, dtype=torch.uint8, device=x.device)
    eng.l2_flush = flush
    loss = kldiv(model(x), gt)
    loss.backward()
    torch.cuda.synchronize()
    rows = []
    for name, kind, flops, e0, e1, kern in eng.profile:
        rows.append({"name": name, "kind": kind, "gflop": flops / 1e9, "ms": e0.elapsed_time(e1), "kernel": kern})
    eng.profile, eng.l2_flush = None, None
    for p in model.parameters():
        p.grad = None
    return rows


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--batch", type=int, default=8, help="clips per GPU (weak scaling)")
    ap.add_argument("--precision", default="bf16")
    ap.add_argument("--no-adam", action="store_true")
    ap.add_argument("--ddp", action="store_true", help="N>1: wrap in DistributedDataParallel instead of the flat-arena all-reduce")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="launch the step eagerly instead of replaying one CUDA graph")
    ap.add_argument("--kernel-table", default="", help="write the per-conv-launch timing table to this JSON file")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    from vinet_b200 import VideoSaliencyModel, kldiv
    from vinet_b200 import lib as L

    rank, world, local = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))
    assert world == args.gpus or world == 1, "launch with torchrun --nproc-per-node == --gpus"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    torch.manual_seed(0)
    model = VideoSaliencyModel().to(dev).set_precision(args.precision).train()
    net = model
    if world > 1 and args.ddp:
        net = torch.nn.parallel.DistributedDataParallel(model, device_ids=[local], gradient_as_bucket_view=True)
    elif world > 1:
        # one process per GPU, same weights everywhere, every gradient written into ONE flat buffer that is averaged
        # with a single NCCL all-reduce per step (model.sync_gradients) - no per-tensor bucket copies
        model.broadcast_parameters(0)
        model.enable_grad_arena()
    # N=1: the whole step is one CUDA graph.  N>1 (flat arena): forward + loss + backward are the graph, the NCCL all-reduce
    # and the optimizer run eagerly after each replay (capturing NCCL inside the step dead-locked at N=2 in round 1).
    use_graph = not args.no_graph and not args.no_adam and not (world > 1 and args.ddp)
    opt = None if args.no_adam else torch.optim.Adam(model.parameters(), lr=1e-4, fused=True, capturable=use_graph and world == 1)
    B = args.batch
    g = torch.Generator().manual_seed(1234 + rank)
    # caller layout: (B,T,3,H,W) memory viewed as (B,3,T,H,W)  (train.py:204-205)
    hx = torch.randn(B, T, 3, H, W, generator=g).pin_memory()
    hgt = (torch.rand(B, H, W, generator=g) + 1e-3).pin_memory()
    dx = hx.to(dev)
    dgt = hgt.to(dev)
    lib = L.get()

    def step(x_btchw, gt):
        pred = net(x_btchw.permute(0, 2, 1, 3, 4))
        loss = kldiv(pred, gt)
        loss.backward()
        if world > 1 and not args.ddp:
            model.sync_gradients()
        if opt is not None:
            opt.step()
            opt.zero_grad(set_to_none=True)
        else:
            for p in model.parameters():
                p.grad = None
        return loss

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = t.item()
        return ms

    graphed = None
    if use_graph:
        # the public GraphedTrainStep API: the whole step (forward, kldiv, backward, fused Adam) captured once as a CUDA
        # graph and replayed; falls back to eager launches if the capture is refused
        try:
            from vinet_b200 import GraphedTrainStep
            graphed = GraphedTrainStep(model, kldiv, opt, dx.permute(0, 2, 1, 3, 4), dgt,
                                       after_backward=model.sync_gradients if world > 1 else None, capture_optimizer=world == 1)
            eager_step = step

            def step(x_btchw, gt):          # noqa: F811
                return graphed(x_btchw.permute(0, 2, 1, 3, 4), gt)
        except Exception as ex:          # pragma: no cover
            sys.stderr.write("bench: CUDA graph capture failed (%s); running eagerly\n" % str(ex)[:200])
            graphed = None
    for _ in range(max(args.warmup, 3)):
        step(dx, dgt)
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    n0 = lib.launch_count()
    ms = timed(lambda: step(dx, dgt), args.steps)
    launches = lib.launch_count() - n0
    if graphed is not None:
        launches = graphed.launches_per_replay * args.steps
    # end-to-end: pinned host -> device every step, loss read back every step
    sink = []

    # The input pipeline is double-buffered the way a training loop's prefetcher is: the H2D copy of step i+1 runs on a copy
    # stream while step i computes.  Every timed step still consumes exactly one freshly copied clip + ground truth (the
    # first copy of the timed region is issued, un-overlapped, inside it) and reads its loss back to the host.
    copy_stream = torch.cuda.Stream(device=dev)
    pending = []

    def issue_copy():
        with torch.cuda.stream(copy_stream):
            x = hx.to(dev, non_blocking=True)
            gt = hgt.to(dev, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(copy_stream)
        pending.append((x, gt, ev))

    def e2e_run(steps):
        for i in range(steps):
            if not pending:
                issue_copy()
            x, gt, ev = pending.pop(0)
            cur = torch.cuda.current_stream()
            cur.wait_event(ev)
            x.record_stream(cur)
            gt.record_stream(cur)
            if i + 1 < steps:
                issue_copy()               # prefetch the next step's input behind this step's kernels
            sink.append(step(x, gt).item())
    e2e_run(2)
    ms_e2e = timed(lambda: e2e_run(args.steps), 1)
    if sampler:
        sampler.stop_flag = True
        sampler.join(timeout=2)

    clips = B * world * args.steps
    value = clips / (ms / 1e3)
    e2e_value = clips / (ms_e2e / 1e3)
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    burst, sustained, hbm, src = peaks()
    table = kernel_roofline(model, dx.permute(0, 2, 1, 3, 4), dgt, kldiv, torch) if world == 1 else []
    roof = None
    if table:
        # dominant kernel = the CUDA kernel with the largest total time over the step's conv launches; its roofline entry
        # is that kernel's largest launch, timed alone (L2 flushed) with CUDA events on the launching stream
        by_kernel = {}
        for r in table:
            by_kernel[r["kernel"]] = by_kernel.get(r["kernel"], 0.0) + r["ms"]
        dom_kernel = max(by_kernel, key=by_kernel.get)
        dom = max((r for r in table if r["kernel"] == dom_kernel), key=lambda r: (r["gflop"], -r["ms"]))
        tot_ms = sum(r["ms"] for r in table)
        key = "%s/%s:%s" % (dom_kernel, dom["kind"], dom["name"])
        traffic = None          # DRAM bytes per launch of that kernel from the committed ncu --set full capture (B=8)
        tk = os.path.join(ROOT, "profiles", "r1_top_kernel.json")
        if os.path.isfile(tk) and B == 8:
            traffic = (json.load(open(tk)).get("captures", {}).get(key) or {}).get("traffic_bytes")
        roof = {"bound": "tensor", "achieved": dom["gflop"] / dom["ms"], "peak": burst, "unit": "TFLOP/s",
                "frac": dom["gflop"] / dom["ms"] / burst, "traffic": traffic, "kernel": key,
                "kernel_share_of_conv_time": by_kernel[dom_kernel] / tot_ms,
                "peak_source": src + " burst bf16 (kernel timed alone, L2 flushed)",
                "algorithmic_gflop_per_launch": dom["gflop"], "launch_ms": dom["ms"],
                "conv_kernels_gflop": sum(r["gflop"] for r in table), "conv_kernels_ms_isolated": tot_ms,
                "conv_kernels_tflops": sum(r["gflop"] for r in table) / tot_ms,
                "conv_ms_by_kernel": {k: round(v, 3) for k, v in sorted(by_kernel.items(), key=lambda kv: -kv[1])},
                "step_tflops": value * FWD_BWD_GFLOP / 1e3, "step_frac_of_sustained": value * FWD_BWD_GFLOP / 1e3 / world / sustained}
        if args.kernel_table:
            json.dump(table, open(args.kernel_table, "w"), indent=0)
    line = {"metric": "clips/sec (32x224x384) fwd+bwd", "value": value, "unit": "clips/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16" if args.precision == "bf16" else "f32", "data": "synthetic",
            "config": {"workload": "ViNet (VideoSaliencyModel) fwd + kldiv + bwd%s, batch %d x 32x224x384 clips per GPU, %s"
                                   % ("" if args.no_adam else " + fused Adam", B, args.precision),
                       "global_batch": B * world, "parallelism": "dp%d" % world,
                       "cuda_graph": graphed is not None, "grad_sync": "none" if world == 1 else ("DistributedDataParallel" if args.ddp else "flat arena, one NCCL all-reduce"),
                       "l2": "inputs (264 MB/clip-batch) and activations (GBs) exceed the 126 MB L2; no explicit flush",
                       "e2e_pipeline": "H2D of step i+1 (pinned host, copy stream) overlaps the kernels of step i; loss.item() every step"},
            "e2e": {"value": e2e_value, "unit": "clips/s", "h2d_bytes_per_step": hx.numel() * 4 + hgt.numel() * 4, "d2h_bytes_per_step": 4,
                    "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": launches, "clocks": sampler.summary() if sampler else None, "roofline": roof}
    if world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline()
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
