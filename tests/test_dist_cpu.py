"""world_size-2 gloo test of the host-side multi-GPU logic (view sharding + gradient all-reduce).
The per-rank gradients come from the CPU oracle here; the CUDA path is covered by -m gpu tests."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from conftest import load_mesh
    from drt_b200 import dist as ddist, views
    from oracle import oracle
    v, f = load_mesh("hand_vh")
    cams = views.turntable_cameras(v, 40, 40, 6)
    m = oracle.OracleMesh(v, f)

    def grad_of(view_ids):
        g = np.zeros_like(v)
        loss = 0.0
        for k in view_ids:
            o, d = views.generate_ray(40, 40, cams[k][3], cams[k][2])
            o, d = o.numpy(), d.numpy()
            qq = m.trace_fwd(o, d, 1.4723)
            gd = 2 * qq["out_dir"] * qq["mask"]
            loss += float((qq["out_dir"] ** 2).sum())
            g += m.trace_bwd(o, d, qq["tri1"], qq["tri2"], None, gd, 1.4723)
        return g, loss

    mine = ddist.shard_views(6, rank, world)
    g, loss = grad_of(mine)
    gt, lt = ddist.allreduce_grad(torch.from_numpy(g.copy()), torch.tensor([loss], dtype=torch.float64))
    g2 = torch.from_numpy(g.copy())
    ddist.allreduce_grad(g2)
    if rank == 0:
        g_all, loss_all = grad_of(range(6))
        q.put((mine, float(np.abs(gt.numpy() - g_all).max() / np.abs(g_all).max()), abs(lt.item() - loss_all),
               float((g2 - gt).abs().max())))
    dist.barrier()
    dist.destroy_process_group()


def test_view_sharding_and_grad_allreduce_gloo():
    from drt_b200 import dist as ddist
    assert ddist.shard_views(72, 3, 8) == list(range(3, 72, 8))
    assert sorted(sum((ddist.shard_views(7, r, 2) for r in range(2)), [])) == list(range(7))
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p_ in procs:
        p_.start()
    mine, rel, dl, d2 = q.get(timeout=180)
    for p_ in procs:
        p_.join(timeout=60)
        assert p_.exitcode == 0
    assert mine == [0, 2, 4]
    assert rel < 1e-12 and dl < 1e-9 and d2 == 0.0


def test_balanced_view_sharding():
    from drt_b200 import dist as ddist
    rng = np.random.default_rng(3)
    costs = (1.0 + 0.5 * np.sin(np.arange(72) / 72 * 4 * np.pi) + 0.1 * rng.random(72)).tolist()   # coverage varies with the azimuth
    for world in (1, 2, 4, 8):
        parts = [ddist.shard_views_balanced(costs, r, world) for r in range(world)]
        assert sorted(sum(parts, [])) == list(range(72))                      # a partition
        load = [sum(costs[k] for k in p) for p in parts]
        rr = [sum(costs[k] for k in ddist.shard_views(72, r, world)) for r in range(world)]
        assert max(load) <= max(rr) + 1e-12                                   # never worse than round-robin here
        assert max(load) - min(load) <= max(costs)                            # LPT bound
    assert ddist.shard_views_balanced([1, 1, 1, 1], 1, 2) == [1, 3]           # ties: stable and deterministic


def test_allreduce_is_identity_without_process_group():
    from drt_b200 import dist as ddist
    g = torch.ones(4, 3, dtype=torch.float64)
    out, loss = ddist.allreduce_grad(g, torch.tensor([2.0]))
    assert out is g and loss.item() == 2.0
