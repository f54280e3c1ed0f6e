"""Multi-GPU plumbing: one process per GPU, views sharded, mesh/BVH replicated, ONE collective per
step -- all-reduce(SUM) of the vertex gradient.  The reference is single-GPU
(optix_extend.cpp:10); views are independent in the forward pass and the backward pass is a sum
over rays, so sharding rays and summing grad_V is exact up to float64 summation order."""
import torch
import torch.distributed as dist


def shard_views(n_views, rank, world):
    """View k -> rank k mod world (round-robin keeps neighbouring azimuths on different GPUs, so
    every rank sees a similar mix of coverages)."""
    return list(range(rank, n_views, world))


def allreduce_grad(grad_V, loss=None, group=None):
    """In-place SUM all-reduce of grad_V [V,3] float64 (+ optionally a loss scalar riding along in
    the same message).  Returns (grad_V, loss)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return grad_V, loss
    if loss is None:
        dist.all_reduce(grad_V, op=dist.ReduceOp.SUM, group=group)
        return grad_V, None
    buf = torch.cat([grad_V.reshape(-1), loss.reshape(-1).to(grad_V.dtype)])
    dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=group)
    grad_V.copy_(buf[:grad_V.numel()].view_as(grad_V))
    return grad_V, buf[grad_V.numel():].clone()
