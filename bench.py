"""bench.py — clips/sec of the ViNet / AViNet hot path on N B200s; one JSON line on rank 0.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                  [--model vinet|avinet] [--mode train|eval] [--clip-len T] [--height H] [--width W] [--batch B] [--precision bf16]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Default (what the driver runs) = BASELINE.json config 2 at N=1 / config 3 at N>1: ViNet, batch 8 of 32x224x384 clips per GPU, bf16:
forward, kldiv loss, backward, (N>1: ONE NCCL all-reduce of the flat gradient arena), fused Adam — one CUDA-graph replay per step.
Other BASELINE.json configs: `--model avinet` (config 4: 4 clips per rank + audio), `--mode eval` (config 1: inference forward),
`--clip-len / --height / --width` (config 5: the resolution / clip-length sweep, restated per SURVEY §8d to the shapes the
reference supports: H, W multiples of 32, T in {8,16,32,48}).
`value` has inputs resident in HBM; `e2e` goes through the public module API from pinned host buffers with the H2D copy of the
clip (+ gt / audio) and the D2H read of the result (loss scalar, or the saliency maps in eval mode) inside the timed region.
`--impl reference` times the reference's own CPU implementation of the same workload on the host cores: the UNMODIFIED reference
sources when `baseline/_ref` holds them (copied there by `__graft_entry__.build()` in the build container; `kind: "reference"`),
else the oracle restatement pinned to the executed reference (`kind: "port"`).
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

# conv MACs x2 per clip at 224x384 (SURVEY.md §8d / BASELINE.md §2, torch FlopCounterMode on the unmodified reference);
# exactly proportional to H*W
GFLOP = {"fwd": {8: 57.246, 16: 114.486, 32: 229.318, 48: 343.974}, "fwd_bwd": {8: 168.500, 16: 336.983, 32: 675.006, 48: 1012.501}}
AV_EXTRA = {"fwd": 0.190, "fwd_bwd": 0.498}          # SoundNet (bilinear fusion not counted)


def gflop_per_clip(model, mode, T, H, W):
    kind = "fwd_bwd" if mode == "train" else "fwd"
    g = GFLOP[kind][T] * (H * W) / (224.0 * 384.0)
    return g + (AV_EXTRA[kind] if model == "avinet" else 0.0)


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


def workload_name(args, reference=False):
    what = ("ViNet (VideoSaliencyModel)" if args.model == "vinet" else
            {"bilinear": "AViNet (VideoAudioSaliencyModel, SoundNet audio fusion)",
             "transformer": "AViNet (VideoAudioSaliencyModel(use_transformer=True), 3 encoder layers over the 32 channel tokens)",
             "tokens": "VideoAudioSaliencyFusionModel (336 visual + 3 audio tokens x 512 through 3 encoder layers)"}[args.av_fusion])
    if args.mode == "train":
        return "%s fwd + kldiv + bwd%s" % (what, "" if (args.no_adam or reference) else " + fused Adam")
    return "%s eval forward (no_grad%s)" % (what, "" if reference else ", BatchNorm folded")


# ----------------------------------------------------------------------------------------------- reference arm (CPU)
def reference_models(args):
    """(model, kldiv, kind): the unmodified reference from baseline/_ref when present, else the pinned oracle restatement."""
    import torch
    from oracle import ref_loader
    from oracle import torch_oracle as O
    ref_dir = os.path.join(ROOT, "baseline", "_ref")
    if os.path.isfile(os.path.join(ref_dir, "model.py")):
        ref_loader.use_dir(ref_dir)
        if args.model == "vinet":
            m = ref_loader.build_vinet(args.clip_len)
        elif args.av_fusion == "tokens":
            m = ref_loader.build_fusion(random_soundnet=True)
        else:
            m = ref_loader.build_avinet(random_soundnet=True, use_transformer=args.av_fusion == "transformer")
        _, rl = ref_loader.load()
        return m, rl.kldiv, "reference"
    if args.model == "vinet":
        m = O.ViNetOracle(args.clip_len)
    else:
        m = O.AVFusionOracle() if args.av_fusion == "tokens" else O.AViNetOracle(args.clip_len, use_transformer=args.av_fusion == "transformer")
    return m, O.kldiv, "port"


def cpu_baseline(args, steps=2, warmup=1):
    """The reference's CPU path on the host cores for the same workload, one clip per step (a bounded sample)."""
    import torch
    from oracle import torch_oracle as O
    torch.set_num_threads(os.cpu_count())
    m, kld, kind = reference_models(args)
    d = O.make_inputs(1, args.clip_len, args.height, args.width, 0, audio=(args.model == "avinet"))
    inputs = (d["x"],) if args.model == "vinet" else (d["x"], d["audio"])
    train = args.mode == "train"
    m.train() if train else m.eval()
    ts = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        if train:
            for p in m.parameters():
                p.grad = None
            loss = kld(m(*inputs), d["gt"])
            loss.backward()
        else:
            with torch.no_grad():
                m(*inputs)
        ts.append(time.perf_counter() - t0)
    ts = sorted(ts[warmup:])
    med = ts[len(ts) // 2]
    return {"value": 1.0 / med, "unit": "clips/s", "cores": os.cpu_count(), "kind": kind,
            "sample": "%d timed %s iterations of 1 clip %dx%dx%d fp32 (median %.2f s), torch %s, %d threads"
                      % (steps, "fwd+bwd" if train else "eval fwd", args.clip_len, args.height, args.width, med, torch.__version__,
                         torch.get_num_threads())}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    t0 = time.perf_counter()
    cb = cpu_baseline(args, steps=max(1, min(args.steps, 5)), warmup=min(args.warmup, 1))
    ms = 1000.0 / cb["value"]
    line = {"impl": "reference", "metric": metric_name(args), "value": cb["value"], "unit": "clips/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "%s, %dx%dx%d clips, the reference's CPU path, 1 clip per step"
                                   % (workload_name(args, True), args.clip_len, args.height, args.width)},
            "cpu_baseline": cb, "e2e": {"value": cb["value"], "unit": "clips/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "wall_s": time.perf_counter() - t0}
    print(json.dumps(line))


def metric_name(args):
    return "clips/sec (%dx%dx%d) %s" % (args.clip_len, args.height, args.width, "fwd+bwd" if args.mode == "train" else "fwd")


# ----------------------------------------------------------------------------------------------- our arm
def kernel_table(model, inputs, gt, kldiv, torch, train):
    """Time every conv launch of one step alone (L2 flushed before each) with CUDA events on the launching stream."""
    eng = model._engine_for(inputs[0].device)
    eng.profile = []
    eng.l2_flush = torch.empty(192 * 1024 * 1024, dtype=torch.uint8, device=inputs[0].device)
    if train:
        loss = kldiv(model(*inputs), gt)
        loss.backward()
    else:
        with torch.no_grad():
            model(*inputs)
    torch.cuda.synchronize()
    rows = [{"name": n, "kind": k, "gflop": f / 1e9, "ms": e0.elapsed_time(e1), "kernel": kern} for n, k, f, e0, e1, kern in eng.profile]
    eng.profile, eng.l2_flush = None, None
    for p in model.parameters():
        p.grad = None
    return rows


def roofline_entry(table, burst, sustained, src, value, world, gf_clip, B):
    """Dominant kernel = the CUDA kernel with the largest total time over the step's conv launches.  `achieved` is the
    FLOP-weighted rate of that kernel family (sum of algorithmic GFLOP / sum of its isolated launch times); the best single
    launch is kept as a secondary field."""
    by_ms, by_gf = {}, {}
    for r in table:
        by_ms[r["kernel"]] = by_ms.get(r["kernel"], 0.0) + r["ms"]
        by_gf[r["kernel"]] = by_gf.get(r["kernel"], 0.0) + r["gflop"]
    dom = max(by_ms, key=by_ms.get)
    launches = [r for r in table if r["kernel"] == dom]
    best = max(launches, key=lambda r: r["gflop"] / max(r["ms"], 1e-9))
    tot_ms, tot_gf = sum(by_ms.values()), sum(by_gf.values())
    achieved = by_gf[dom] / by_ms[dom]
    traffic = None          # DRAM bytes per launch of the dominant kernel's largest launch, from the committed ncu --set full capture
    key = "%s/%s:%s" % (dom, best["kind"], best["name"])
    for f in ("r2_top_kernel.json", "r1_top_kernel.json"):
        tk = os.path.join(ROOT, "profiles", f)
        if os.path.isfile(tk) and B == 8 and traffic is None:
            traffic = (json.load(open(tk)).get("captures", {}).get(key) or {}).get("traffic_bytes")
    return {"bound": "tensor", "achieved": achieved, "peak": burst, "unit": "TFLOP/s", "frac": achieved / burst, "traffic": traffic,
            "kernel": dom, "how": "FLOP-weighted over the %d launches of the dominant kernel in one step, each timed alone with CUDA "
                                  "events on the launching stream after an L2 flush" % len(launches),
            "kernel_share_of_conv_time": by_ms[dom] / tot_ms,
            "peak_source": src + " burst bf16 (kernels timed alone, L2 flushed)",
            "algorithmic_gflop": by_gf[dom], "kernel_ms": by_ms[dom],
            "best_launch": {"name": "%s:%s" % (best["kind"], best["name"]), "tflops": best["gflop"] / best["ms"],
                            "frac": best["gflop"] / best["ms"] / burst, "gflop": best["gflop"], "ms": best["ms"]},
            "all_conv_kernels": {"gflop": tot_gf, "ms_isolated": tot_ms, "tflops": tot_gf / tot_ms, "frac": tot_gf / tot_ms / burst,
                                 "ms_by_kernel": {k: round(v, 3) for k, v in sorted(by_ms.items(), key=lambda kv: -kv[1])},
                                 "tflops_by_kernel": {k: round(by_gf[k] / by_ms[k], 1) for k in by_ms}},
            "step_tflops": value * gf_clip / 1e3, "step_frac_of_sustained": value * gf_clip / 1e3 / world / sustained}


def parity_probe(torch, precision):
    """Accuracy of the benchmarked precision mode next to the speed: one 32-frame 128x192 clip pair in train mode against an
    fp64 run of the oracle on the same GPU (max relative saliency-map error, kldiv relative error), with stock bf16 autocast
    and the tensor-core parity mode as yardsticks."""
    import copy
    from oracle import torch_oracle as O
    from vinet_b200 import VideoSaliencyModel, kldiv
    T, B, H, W = 32, 2, 128, 192
    ref = O.ViNetOracle(T)
    O.randomize_(ref, 11)
    d = O.make_inputs(B, T, H, W, 11)
    x, gt = d["x"].cuda(), d["gt"].cuda()
    ref = ref.cuda().train()
    ref64 = copy.deepcopy(ref).double()
    with torch.no_grad():
        p64 = ref64(x.double())
        l64 = O.kldiv(p64, gt.double()).item()
        with torch.autocast("cuda", dtype=torch.bfloat16):
            pa = ref(x).float()
        la = O.kldiv(pa, gt).item()
    out = {"shape": "2 x 32x128x192, train-mode BatchNorm, vs fp64 oracle",
           "torch_bf16_autocast": {"map_max_rel": float(((pa - p64).abs() / p64.abs()).max()), "kldiv_rel": abs(la - l64) / abs(l64)}}
    for prec in dict.fromkeys([precision, "bf16x6"]):
        m = VideoSaliencyModel(num_clips=T)
        m.load_state_dict(ref.state_dict())
        m = m.cuda().set_precision(prec).train()
        with torch.no_grad():
            pm = m(x)
            lm = kldiv(pm, gt).item()
        out[prec] = {"map_max_rel": float(((pm - p64).abs() / p64.abs()).max()), "kldiv_rel": abs(lm - l64) / abs(l64)}
        del m
    torch.cuda.empty_cache()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--model", default="vinet", choices=["vinet", "avinet"])
    ap.add_argument("--av-fusion", default="bilinear", choices=["bilinear", "transformer", "tokens"],
                    help="--model avinet: bilinear = VideoAudioSaliencyModel (config 4); transformer = its use_transformer=True "
                         "variant (model.py:239-247); tokens = VideoAudioSaliencyFusionModel (model.py:116-189)")
    ap.add_argument("--mode", default="train", choices=["train", "eval"])
    ap.add_argument("--clip-len", type=int, default=32, choices=[8, 16, 32, 48])
    ap.add_argument("--height", type=int, default=224)
    ap.add_argument("--width", type=int, default=384)
    ap.add_argument("--batch", type=int, default=0, help="clips per GPU (weak scaling); default 8 (ViNet) / 4 (AViNet, config 4)")
    ap.add_argument("--precision", default="bf16")
    ap.add_argument("--no-adam", action="store_true")
    ap.add_argument("--ddp", action="store_true", help="N>1: wrap in DistributedDataParallel instead of the flat-arena all-reduce")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity", action="store_true", help="skip the accuracy probe of the benchmarked precision mode")
    ap.add_argument("--no-graph", action="store_true", help="launch the step eagerly instead of replaying one CUDA graph")
    ap.add_argument("--graph-nccl", type=int, default=1, help="N>1: 1 = the NCCL all-reduce and the optimizer are captured in the "
                    "step's CUDA graph, 0 = they run eagerly after every replay")
    ap.add_argument("--kernel-table", default="", help="write the per-conv-launch timing table to this JSON file")
    args = ap.parse_args()
    assert args.height % 32 == 0 and args.width % 32 == 0, "the reference supports H, W multiples of 32 only (SURVEY fact 8)"
    if args.model == "avinet":
        assert (args.clip_len, args.height, args.width) == (32, 224, 384), "AViNet is hard-wired to 32x224x384 (model.py:230,237)"
    if args.batch <= 0:
        args.batch = 4 if args.model == "avinet" else 8
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    from vinet_b200 import GraphedForward, GraphedTrainStep, VideoAudioSaliencyFusionModel, VideoAudioSaliencyModel, VideoSaliencyModel, kldiv
    from vinet_b200 import lib as L

    rank, world, local = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))
    assert world == args.gpus or world == 1, "launch with torchrun --nproc-per-node == --gpus"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        dist.all_reduce(torch.zeros(1, device=dev))      # communicator set-up and first collective before anything is captured
    torch.manual_seed(0)
    train = args.mode == "train"
    T, H, W, B = args.clip_len, args.height, args.width, args.batch
    if args.model == "vinet":
        model = VideoSaliencyModel(num_clips=T)
    elif args.av_fusion == "tokens":
        model = VideoAudioSaliencyFusionModel(num_clips=T, soundnet_weights=False)
    else:                                                                      # synthetic benchmark: random SoundNet weights
        model = VideoAudioSaliencyModel(use_transformer=args.av_fusion == "transformer", num_clips=T, soundnet_weights=False)
    model = model.to(dev).set_precision(args.precision)
    model.train() if train else model.eval()
    net = model
    if train and world > 1 and args.ddp:
        net = torch.nn.parallel.DistributedDataParallel(model, device_ids=[local], gradient_as_bucket_view=True,
                                                        find_unused_parameters=args.model == "avinet")
    elif train and world > 1:
        # one process per GPU, same weights everywhere, every gradient written into ONE flat buffer that is averaged
        # with a single NCCL all-reduce per step (model.sync_gradients) - no per-tensor bucket copies
        model.broadcast_parameters(0)
        model.enable_grad_arena()
        if not os.environ.get("VINET_NO_OVERLAP_SYNC"):
            model.overlap_gradient_sync()     # the decoder's slice of the arena is reduced while the backbone's backward runs
    use_graph = not args.no_graph and not (train and args.no_adam) and not (world > 1 and args.ddp)
    opt = None
    nccl_in_graph = world > 1 and bool(args.graph_nccl) and use_graph and not args.ddp
    if train and not args.no_adam:
        opt = torch.optim.Adam([p for p in model.parameters() if p.requires_grad], lr=1e-4, fused=True,
                               capturable=use_graph and (world == 1 or nccl_in_graph))
    g = torch.Generator().manual_seed(1234 + rank)
    # caller layout: (B,T,3,H,W) memory viewed as (B,3,T,H,W)  (train.py:204-205)
    hx = torch.randn(B, T, 3, H, W, generator=g).pin_memory()
    hgt = (torch.rand(B, H, W, generator=g) + 1e-3).pin_memory()
    host = [hx, hgt]
    if args.model == "avinet":
        from vinet_b200 import arch
        ha = torch.zeros(B, 1, arch.AUDIO_LEN, 1)
        n = 47040          # 32 frames @ 15 fps of 22.05 kHz audio, Hanning-windowed (dataloader.py:113-118)
        s0 = (arch.AUDIO_LEN - n) // 2
        ha[:, 0, s0:s0 + n, 0] = 0.05 * torch.randn(B, n, generator=g) * torch.hann_window(n)
        host.append(ha.pin_memory())
    dx, dgt = hx.to(dev), hgt.to(dev)
    dextra = [t.to(dev) for t in host[2:]]
    lib = L.get()

    def clip_view(x_btchw):
        return x_btchw.permute(0, 2, 1, 3, 4)

    def eager_step(x_btchw, gt, *extra):
        if not train:
            with torch.no_grad():
                return net(clip_view(x_btchw), *extra)
        pred = net(clip_view(x_btchw), *extra)
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
    step = eager_step

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
        # the public graph APIs: the whole step (forward, kldiv, backward, fused Adam) / the inference forward captured once
        # as a CUDA graph and replayed; falls back to eager launches if the capture is refused
        try:
            if train:
                graphed = GraphedTrainStep(model, kldiv, opt, clip_view(dx), dgt, example_extra=dextra,
                                           after_backward=model.sync_gradients if world > 1 else None,
                                           capture_optimizer=world == 1 or nccl_in_graph)

                def step(x_btchw, gt, *extra):          # noqa: F811
                    return graphed(clip_view(x_btchw), gt, *extra)
            else:
                graphed = GraphedForward(model, clip_view(dx), *dextra)

                def step(x_btchw, gt, *extra):          # noqa: F811
                    return graphed(clip_view(x_btchw), *extra)
        except Exception as ex:          # pragma: no cover
            sys.stderr.write("bench: CUDA graph capture failed (%s); running eagerly\n" % str(ex)[:300])
            graphed, step = None, eager_step
    for _ in range(max(args.warmup, 3)):
        step(dx, dgt, *dextra)
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    n0 = lib.launch_count()
    ms = timed(lambda: step(dx, dgt, *dextra), args.steps)
    launches = lib.launch_count() - n0
    if graphed is not None:
        launches = graphed.launches_per_replay * args.steps
    # ---- end to end: pinned host -> device every step, result read back every step.  The input pipeline is double-buffered the
    # way a training loop's prefetcher is: the H2D copy of step i+1 runs on a copy stream while step i computes.  Every timed
    # step still consumes exactly one freshly copied batch (the first copy of the timed region is issued, un-overlapped, inside
    # it) and reads its result (the loss scalar; in eval mode the B saliency maps) back to the host.
    sink = []
    copy_stream = torch.cuda.Stream(device=dev)
    pending = []
    host_out = torch.empty(B, H, W).pin_memory() if not train else None
    loss_ring = [(torch.empty(1).pin_memory(), torch.cuda.Event()) for _ in range(2)]

    def issue_copy():
        with torch.cuda.stream(copy_stream):
            ts = [t.to(dev, non_blocking=True) for t in host]
            ev = torch.cuda.Event()
            ev.record(copy_stream)
        pending.append((ts, ev))

    def e2e_run(steps):
        for i in range(steps):
            if not pending:
                issue_copy()
            ts, ev = pending.pop(0)
            cur = torch.cuda.current_stream()
            cur.wait_event(ev)
            for t in ts:
                t.record_stream(cur)
            if i + 1 < steps:
                issue_copy()               # prefetch the next step's input behind this step's kernels
            r = step(ts[0], ts[1], *ts[2:])
            if train:
                # the loss of EVERY step is copied to pinned host memory and read by the host - one step late, the way a training
                # loop logs it: the copy of step i is awaited after step i+1 has been launched, so the host never idles the GPU
                slot = loss_ring[i % 2]
                slot[0].copy_(r.detach().reshape(1), non_blocking=True)
                slot[1].record(torch.cuda.current_stream())
                if i > 0:
                    prev = loss_ring[(i - 1) % 2]
                    prev[1].synchronize()
                    sink.append(float(prev[0][0]))
                if i + 1 == steps:
                    slot[1].synchronize()
                    sink.append(float(slot[0][0]))
            else:
                host_out.copy_(r, non_blocking=True)
                torch.cuda.current_stream().synchronize()
                sink.append(float(host_out[0, 0, 0]))
    e2e_run(2)
    ms_e2e = timed(lambda: e2e_run(args.steps), 1)
    if sampler:
        sampler.stop_flag = True
        sampler.join(timeout=2)

    clips = B * world * args.steps
    value = clips / (ms / 1e3)
    e2e_value = clips / (ms_e2e / 1e3)
    if rank != 0:
        return finish(world, nccl_in_graph)
    burst, sustained, hbm, src = peaks()
    gf_clip = gflop_per_clip(args.model, args.mode, T, H, W)
    roof = None
    if world == 1 and args.precision == "bf16":
        table = kernel_table(model, [clip_view(dx)] + dextra, dgt, kldiv, torch, train)
        if table:
            roof = roofline_entry(table, burst, sustained, src, value, world, gf_clip, B)
        if args.kernel_table:
            json.dump(table, open(args.kernel_table, "w"), indent=0)
    else:
        roof = {"bound": "tensor", "achieved": value * gf_clip / 1e3 / world, "peak": sustained, "unit": "TFLOP/s",
                "frac": value * gf_clip / 1e3 / world / sustained, "traffic": None, "kernel": "whole step (per GPU)",
                "peak_source": src + " sustained bf16"}
    h2d = sum(t.numel() * t.element_size() for t in host)
    line = {"metric": metric_name(args), "value": value, "unit": "clips/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16" if args.precision == "bf16" else ("f32" if args.precision == "fp32" else args.precision),
            "data": "synthetic",
            "config": {"workload": "%s, batch %d x %dx%dx%d clips per GPU, %s" % (workload_name(args), B, T, H, W, args.precision),
                       "global_batch": B * world, "parallelism": "dp%d" % world, "gflop_per_clip": gf_clip,
                       "cuda_graph": graphed is not None, "nccl_in_graph": bool(nccl_in_graph and graphed is not None),
                       "grad_sync": "none" if (world == 1 or not train) else ("DistributedDataParallel" if args.ddp else "flat arena, one NCCL all-reduce"),
                       "l2": "inputs (%d MB per batch) and activations (GBs) exceed the 126 MB L2; no explicit flush" % (h2d >> 20),
                       "e2e_pipeline": "H2D of step i+1 (pinned host, copy stream) overlaps the kernels of step i; the result of every step is copied to pinned host memory and read by the host (training: one step late, behind the next launch)"},
            "e2e": {"value": e2e_value, "unit": "clips/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4 if train else B * H * W * 4,
                    "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": launches, "clocks": sampler.summary() if sampler else None, "roofline": roof}
    if world == 1 and not args.no_parity and args.model == "vinet":
        try:
            line["parity"] = parity_probe(torch, args.precision)
        except Exception as ex:          # pragma: no cover - the probe must never cost the bench line
            line["parity"] = {"error": str(ex)[:200]}
    if world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline(args)
    print(json.dumps(line))
    finish(world, nccl_in_graph)


def finish(world, nccl_in_graph):
    """Leave the process group.  With NCCL kernels captured in a live CUDA graph, destroy_process_group() was observed to hang
    (the JSON line was out, the ranks never exited): flush and exit the process directly in that case."""
    import torch.distributed as dist
    sys.stdout.flush()
    sys.stderr.flush()
    if world > 1 and nccl_in_graph:
        os._exit(0)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
