"""CPU, world_size 2 over gloo: the N>1 path of bench.py/train.py — one process per rank, the model wrapped in
DistributedDataParallel, clips sharded on dim 0, per-replica BatchNorm, ONE gradient all-reduce (mean).
The kernels are replaced by the numpy kernel spec (there is no GPU here); what is checked is the host logic:
averaged gradients are identical on both ranks and equal the mean of the two single-rank gradients."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
T, H, W = 8, 64, 64


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _model():
    from oracle import torch_oracle as O
    from oracle.kernel_spec import Spec
    from vinet_b200 import VideoSaliencyModel
    ref = O.ViNetOracle(T)
    O.randomize_(ref, 7)
    m = VideoSaliencyModel(num_clips=T)
    m.load_state_dict(ref.state_dict())
    m.set_precision("fp32")
    m.__dict__["_backend"] = Spec()
    return m.train()


def _local_grads(rank):
    from oracle import torch_oracle as O
    m = _model()
    d = O.make_inputs(1, T, H, W, 100 + rank)
    O.kldiv(m(d["x"]), d["gt"]).backward()
    return {n: p.grad.clone() for n, p in m.named_parameters()}


def _worker_arena(rank, world, port, out):
    """Same job through the flat gradient arena: model.enable_grad_arena() + model.sync_gradients() (bench.py's N>1 path)."""
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.set_num_threads(2)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import torch_oracle as O
    m = _model()
    if rank == 1:                                       # rank 1 starts from different weights: the broadcast must fix that
        with torch.no_grad():
            for p in m.parameters():
                p.add_(1.0)
    m.broadcast_parameters(0)
    m.enable_grad_arena()
    d = O.make_inputs(1, T, H, W, 100 + rank)
    loss = O.kldiv(m(d["x"]), d["gt"])
    loss.backward()
    flat = next(iter(m.__dict__["_arenas"].values()))[0]
    lo, hi = flat.data_ptr(), flat.data_ptr() + flat.numel() * 4
    views = all(lo <= p.grad.data_ptr() < hi for p in m.parameters())      # autograd adopted the arena slices, no copies
    m.sync_gradients()
    grads = {n: p.grad.clone() for n, p in m.named_parameters()}
    torch.save({"grads": grads, "loss": float(loss.detach()), "views": views}, os.path.join(out, "rank%d.pt" % rank))
    dist.barrier()
    dist.destroy_process_group()


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.set_num_threads(2)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import torch_oracle as O
    m = _model()
    ddp = torch.nn.parallel.DistributedDataParallel(m)
    d = O.make_inputs(1, T, H, W, 100 + rank)          # a different clip per rank
    loss = O.kldiv(ddp(d["x"]), d["gt"])
    loss.backward()
    grads = {n: p.grad.clone() for n, p in m.named_parameters()}
    torch.save({"grads": grads, "loss": float(loss.detach())}, os.path.join(out, "rank%d.pt" % rank))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(900)
def test_ddp_gloo_world2_gradients_are_the_mean_of_the_shards(tmp_path):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    r0 = torch.load(os.path.join(tmp_path, "rank0.pt"))
    r1 = torch.load(os.path.join(tmp_path, "rank1.pt"))
    g0, g1 = _local_grads(0), _local_grads(1)
    assert r0["loss"] != r1["loss"]
    n_checked = 0
    for n in g0:
        a, b = r0["grads"][n], r1["grads"][n]
        assert torch.equal(a, b), n                     # every rank holds the same all-reduced gradient
        want = 0.5 * (g0[n] + g1[n])
        assert torch.allclose(a, want, rtol=1e-4, atol=1e-6 * (want.abs().max().item() + 1e-12)), n
        n_checked += 1
    assert n_checked == len(g0) >= 230                  # every weight / bias of the T=8 model (T=32: 239)


@pytest.mark.timeout(900)
def test_flat_arena_world2_one_allreduce_gives_the_mean_of_the_shards(tmp_path):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    port = _free_port()
    mp.spawn(_worker_arena, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    r0 = torch.load(os.path.join(tmp_path, "rank0.pt"))
    r1 = torch.load(os.path.join(tmp_path, "rank1.pt"))
    assert r0["views"] and r1["views"]
    g0, g1 = _local_grads(0), _local_grads(1)
    for n in g0:
        a, b = r0["grads"][n], r1["grads"][n]
        assert torch.equal(a, b), n
        want = 0.5 * (g0[n] + g1[n])
        assert torch.allclose(a, want, rtol=1e-4, atol=1e-6 * (want.abs().max().item() + 1e-12)), n
