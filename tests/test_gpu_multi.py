"""GPU: the reference's multi-GPU entry (nn.DataParallel, /root/reference/train.py:182-184) and the one-process-per-GPU path
(flat gradient arena + ONE NCCL all-reduce) on the REAL kernels.  The two-device tests skip on a single-GPU box
(run them with `gpurun --gpus 2`); the replica test runs anywhere."""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

from oracle import torch_oracle as O
from vinet_b200 import VideoSaliencyModel, kldiv

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
T, H, W = 8, 64, 96


def _model(seed=0, precision="fp32"):
    ref = O.ViNetOracle(T)
    O.randomize_(ref, seed)
    m = VideoSaliencyModel(num_clips=T)
    m.load_state_dict(ref.state_dict())
    return m.cuda().set_precision(precision).train()


def _direct(x, gt, seed=0):
    m = _model(seed)
    pred = m(x)
    kldiv(pred, gt).backward()
    return pred.detach(), {n: p.grad.clone() for n, p in m.named_parameters()}, {k: v.clone() for k, v in m.state_dict().items()}


def test_data_parallel_replica_trains():
    """What nn.DataParallel does per device: a replica whose `_parameters` are empty and whose weights are broadcast copies
    with a grad_fn.  Forward + backward through the replica must deliver the same output and the same parameter gradients (on
    the ORIGINAL parameters) as the module itself."""
    d = O.make_inputs(2, T, H, W, 0)
    x, gt = d["x"].cuda(), d["gt"].cuda()
    p0, g0, _ = _direct(x, gt)
    m = _model()
    (rep,) = torch.nn.parallel.replicate(m, [torch.cuda.current_device()])
    assert getattr(rep, "_is_replica", False) and len(list(rep.parameters())) == 0
    pred = rep(x)
    kldiv(pred, gt).backward()
    assert torch.allclose(pred, p0, rtol=1e-4, atol=1e-6)
    for n, p in m.named_parameters():
        assert p.grad is not None, n
        assert torch.allclose(p.grad, g0[n], rtol=1e-3, atol=1e-5 * float(g0[n].abs().max()) + 1e-12), n


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_nn_data_parallel_two_gpus_matches_per_shard_runs():
    """train.py:182-184 unchanged: `model = nn.DataParallel(model)`.  One clip per GPU: every replica normalises with its own
    batch statistics (the reference's semantics), so the gathered output equals the two single-clip runs and the gradient of
    the batch-mean loss equals the mean of the per-shard gradients."""
    d = O.make_inputs(2, T, H, W, 0)
    x, gt = d["x"].cuda(0), d["gt"].cuda(0)
    shards = [_direct(x[i:i + 1], gt[i:i + 1]) for i in range(2)]
    m = _model()
    dp = torch.nn.DataParallel(m, device_ids=[0, 1])
    pred = dp(x)
    assert pred.shape == (2, H, W) and pred.device.index == 0
    kldiv(pred, gt).backward()
    for i in range(2):
        assert torch.allclose(pred[i:i + 1], shards[i][0], rtol=1e-4, atol=1e-6), i
    for n, p in m.named_parameters():
        want = 0.5 * (shards[0][1][n] + shards[1][1][n])
        assert torch.allclose(p.grad, want, rtol=2e-3, atol=2e-5 * float(want.abs().max()) + 1e-12), n


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_two_ranks_flat_arena_allreduce_equals_mean_of_shards(precision):
    """SURVEY §4 item 3 on the real kernels: two processes (NCCL), one clip each, flat gradient arena + model.sync_gradients():
    the averaged gradients equal the mean of the two single-GPU shard gradients; BatchNorm statistics stay per rank."""
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", PYTHONPATH=ROOT)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                        "--master-port", "29653", os.path.join(ROOT, "tests", "ddp_gpu_worker.py"), precision],
                       capture_output=True, text=True, timeout=600, env=env)
    print(r.stdout[-1500:])
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "DDP_GPU_WORKER_OK" in r.stdout
