"""CUDA-graph capture of one whole training step (optimizer + step glue).

A step of this path is ~550 kernel launches of a few microseconds each; replaying them as ONE CUDA graph removes the
per-launch CPU work and shortens the gaps between dependent kernels.  Everything the step does is already stream-ordered
C-ABI calls plus PyTorch's fused Adam, so the capture is plain ``torch.cuda.graph``:

    step = GraphedTrainStep(model, vinet_b200.kldiv, optimizer, example_clip, example_gt)
    for clip, gt in loader:
        loss = step(clip, gt)          # device tensor, valid until the next call

The optimizer must be capturable (``torch.optim.Adam(..., fused=True, capturable=True)``).  Shapes are fixed at capture.
Multi-GPU: ``after_backward=model.sync_gradients`` (flat gradient arena).  With ``capture_optimizer=True`` the all-reduce and
the optimizer are captured with the step; with ``capture_optimizer=False`` the graph holds forward + loss + backward and the
all-reduce + optimizer run eagerly after every replay.
"""
import torch


class GraphedTrainStep:
    def __init__(self, model, loss_fn, optimizer, example_x, example_gt, warmup=3, after_backward=None, capture_optimizer=True,
                 example_extra=()):
        assert example_x.is_cuda, "GraphedTrainStep needs CUDA tensors"
        assert warmup >= 1, "at least one eager warm-up step (lazy initialisation must not happen during capture)"
        self.model, self.loss_fn, self.optimizer = model, loss_fn, optimizer
        # after_backward: e.g. model.sync_gradients.  capture_optimizer=False captures forward + loss + backward only and runs
        # after_backward() and optimizer.step() eagerly after every replay (multi-GPU: NCCL stays outside the graph)
        self.after_backward, self.capture_optimizer = after_backward, capture_optimizer
        # static inputs with the caller's memory layout (train.py:205 hands a permuted (B,3,T,H,W) view of BTCHW memory)
        self.x = torch.empty_strided(example_x.shape, example_x.stride(), dtype=example_x.dtype, device=example_x.device)
        self.gt = torch.empty_strided(example_gt.shape, example_gt.stride(), dtype=example_gt.dtype, device=example_gt.device)
        self.x.copy_(example_x)
        self.gt.copy_(example_gt)
        # further model inputs (AViNet: the audio excerpt), also static
        self.extra = [torch.empty_strided(t.shape, t.stride(), dtype=t.dtype, device=t.device).copy_(t) for t in example_extra]
        # warm-up and capture must not change the training state: snapshot parameters, buffers and optimizer state
        saved = [t.detach().clone() for t in list(model.parameters()) + list(model.buffers())]
        fresh_opt = len(optimizer.state) == 0
        saved_opt = None if fresh_opt else [[v.detach().clone() if torch.is_tensor(v) else v for v in st.values()]
                                            for st in optimizer.state.values()]
        cur = torch.cuda.current_stream()
        side = torch.cuda.Stream(device=example_x.device)
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            for _ in range(warmup):
                self._eager_step()
        cur.wait_stream(side)
        torch.cuda.synchronize(example_x.device)
        self.graph = torch.cuda.CUDAGraph()
        optimizer.zero_grad(set_to_none=True)
        from . import lib as _lib
        self._mark_weights_dirty()        # the re-pack of the cached tensor-core weights must be part of the captured step
        n0 = _lib.get().launch_count()
        # thread_local: other threads (the NCCL watchdog of torch.distributed polls CUDA events) may keep calling the CUDA API
        # while this thread captures; the default "global" mode turns those calls into capture errors
        with torch.cuda.graph(self.graph, capture_error_mode="thread_local"):
            self.loss = self._eager_step(zero=False) if capture_optimizer else self._fwd_bwd()
        self.launches_per_replay = _lib.get().launch_count() - n0     # kernels of this library inside one replay
        self._epoch = self._buffers_epoch()
        with torch.no_grad():
            for t, s in zip(list(model.parameters()) + list(model.buffers()), saved):
                t.copy_(s)
            for i, st in enumerate(optimizer.state.values()):
                for j, (k, v) in enumerate(st.items()):
                    if torch.is_tensor(v):
                        v.zero_() if fresh_opt else v.copy_(saved_opt[i][j])
        self._mark_weights_dirty()

    def _fwd_bwd(self):
        loss = self.loss_fn(self.model(self.x, *self.extra), self.gt)
        loss.backward()
        return loss

    def _eager_step(self, zero=True):
        if zero:
            self.optimizer.zero_grad(set_to_none=True)
        loss = self._fwd_bwd()
        if self.after_backward is not None:
            self.after_backward()
        self.optimizer.step()
        return loss

    def _buffers_epoch(self):
        return tuple(e.realloc_count for e in self.model.__dict__.get("_engines", {}).values())

    def _mark_weights_dirty(self):
        # replays update the parameters without bumping Tensor._version: tell the engines that their packed weight copies
        # must be refreshed by the next eager forward (the captured step re-packs by itself)
        for e in self.model.__dict__.get("_engines", {}).values():
            e.weights_dirty = True

    def __call__(self, x, gt, *extra):
        self.x.copy_(x, non_blocking=True)
        self.gt.copy_(gt, non_blocking=True)
        for dst, src in zip(self.extra, extra):
            dst.copy_(src, non_blocking=True)
        if self._epoch != self._buffers_epoch():
            raise RuntimeError("vinet_b200: the model ran with another input shape since this step was captured and its activation "
                               "buffers were re-allocated; capture a new GraphedTrainStep")
        self.graph.replay()
        if not self.capture_optimizer:       # gradients live in static tensors (param.grad) that every replay overwrites
            if self.after_backward is not None:
                self.after_backward()
            self.optimizer.step()
        self._mark_weights_dirty()
        return self.loss


class GraphedForward:
    """Inference forward (``model.eval()``, no autograd) captured once as a CUDA graph and replayed: the ~150 launches of a folded
    BatchNorm forward become one graph launch.  Shapes are fixed at capture; parameters are read at replay time through the
    engine's packed copies, so call ``refresh()`` after loading new weights.

        fwd = GraphedForward(model, example_clip)          # AViNet: GraphedForward(model, example_clip, example_audio)
        saliency = fwd(clip)                               # (B,H,W) fp32, valid until the next call
    """

    def __init__(self, model, example_x, *example_extra, warmup=2):
        assert example_x.is_cuda and not model.training, "GraphedForward needs CUDA tensors and model.eval()"
        self.model = model
        self.inputs = [torch.empty_strided(t.shape, t.stride(), dtype=t.dtype, device=t.device).copy_(t)
                       for t in (example_x,) + tuple(example_extra)]
        self._capture(warmup)

    def _capture(self, warmup):
        cur = torch.cuda.current_stream()
        side = torch.cuda.Stream(device=self.inputs[0].device)
        side.wait_stream(cur)
        with torch.cuda.stream(side), torch.no_grad():
            for _ in range(max(1, warmup)):
                self.model(*self.inputs)
        cur.wait_stream(side)
        torch.cuda.synchronize(self.inputs[0].device)
        from . import lib as _lib
        self.graph = torch.cuda.CUDAGraph()
        n0 = _lib.get().launch_count()
        with torch.no_grad(), torch.cuda.graph(self.graph):
            self.out = self.model(*self.inputs)
        self.launches_per_replay = _lib.get().launch_count() - n0
        self._epoch = tuple(e.realloc_count for e in getattr(self.model, "__dict__", {}).get("_engines", {}).values())

    def refresh(self):
        """Parameters changed (load_state_dict, training steps in between): re-pack and re-capture."""
        self.model.invalidate_weight_cache()
        self._capture(1)

    def stale(self):
        """True when a forward with another input shape re-allocated the engine's activation buffers after the capture."""
        return self._epoch != tuple(e.realloc_count for e in self.model.__dict__.get("_engines", {}).values())

    def __call__(self, x, *extra):
        for dst, src in zip(self.inputs, (x,) + extra):
            dst.copy_(src, non_blocking=True)
        if self.stale():
            self._capture(1)
        self.graph.replay()
        return self.out
