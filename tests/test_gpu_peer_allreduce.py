"""The one collective of the path on real peer memory: drt_comm_* (one-shot all-reduce of grad_V over NVLink,
csrc/peer_allreduce.cuh) under torch.distributed, one process per GPU.  Needs >= 2 GPUs (skipped otherwise);
the host-side sharding / fallback logic is covered on CPU by tests/test_dist_cpu.py."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _contribution(rank, it, n, device):
    k = torch.arange(n, device=device, dtype=torch.float64)
    return torch.sin(k * (0.37 + rank) + it) * (1.0 + rank) + 1e-3 * it


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), DRT_ALLREDUCE="peer")   # force the kernel at any world size
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    from drt_b200 import dist as ddist
    worst, used_peer = 0.0, True
    for n in (1, 7, 75378, 300000):                     # 75 378 = 25 126 vertices x 3 (C4)
        for it in range(12):                            # both epoch parities, many times
            g = _contribution(rank, it, n, dev)
            expect = torch.zeros_like(g)
            for r in range(world):                      # the kernel adds in rank order: bit-exact expectation
                expect += _contribution(r, it, n, dev)
            if n == 75378 and it % 3 == 0:
                g3 = g.view(-1, 3).clone()
                out, loss = ddist.allreduce_grad(g3, torch.tensor([float(rank + it)], dtype=torch.float64, device=dev))
                assert loss.item() == sum(float(r + it) for r in range(world))
                got = out.reshape(-1)
            else:
                got = ddist.allreduce_grad(g)[0]
            worst = max(worst, (got - expect).abs().max().item())
    p = ddist.peer_allreduce(1, dev)
    used_peer = p is not None and p.ok
    timed_out = p.timed_out() if p is not None else False
    # the torch.distributed route gives the same sums (to rounding)
    os.environ["DRT_ALLREDUCE"] = "nccl"
    g = _contribution(rank, 99, 75378, dev)
    ref = sum(_contribution(r, 99, 75378, dev) for r in range(world))
    nccl_err = (ddist.allreduce_grad(g)[0] - ref).abs().max().item()
    if rank == 0:
        q.put((worst, used_peer, timed_out, nccl_err))
    dist.barrier()
    dist.destroy_process_group()


def test_peer_allreduce_two_or_more_gpus():
    n_gpu = torch.cuda.device_count()
    if n_gpu < 2:
        pytest.skip("needs >= 2 GPUs on one box")
    world = min(n_gpu, 8)   # dist.PEER_DEFAULT_MAX_WORLD: every GPU of one box
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p_ in procs:
        p_.start()
    worst, used_peer, timed_out, nccl_err = q.get(timeout=300)
    for p_ in procs:
        p_.join(timeout=120)
        assert p_.exitcode == 0
    assert used_peer, "the peer-memory path was not taken (IPC unavailable?)"
    assert not timed_out
    assert worst == 0.0            # rank-order sum: every rank gets exactly these bits
    assert nccl_err < 1e-12
