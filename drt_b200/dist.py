"""Multi-GPU plumbing: one process per GPU, views sharded, mesh/BVH replicated, ONE collective per
step -- all-reduce(SUM) of the vertex gradient.  The reference is single-GPU
(optix_extend.cpp:10); views are independent in the forward pass and the backward pass is a sum
over rays, so sharding rays and summing grad_V is exact up to float64 summation order.

The collective itself is the library's one-shot peer-memory all-reduce (drt_comm_*: one kernel over NVLink /
NVSwitch, csrc/peer_allreduce.cuh) when every rank of the group sits on the same host and could open its peers'
IPC handles; otherwise torch.distributed's all_reduce (NCCL on GPUs, gloo on CPU tensors).  torch.distributed is
the rendezvous either way (it carries the 64-byte IPC handles)."""
import ctypes as C
import os
import socket

import torch
import torch.distributed as dist


def shard_views(n_views, rank, world):
    """View k -> rank k mod world (round-robin keeps neighbouring azimuths on different GPUs, so
    every rank sees a similar mix of coverages)."""
    return list(range(rank, n_views, world))


def shard_views_balanced(costs, rank, world):
    """Views -> ranks by estimated cost instead of round-robin: longest-processing-time-first (views sorted by cost,
    each given to the least loaded rank; ties -> lower view / lower rank), deterministic, so every rank computes the same
    assignment from the same costs.  The all-reduce phase of a step is the ranks' arrival skew; the cost of
    a view is dominated by the rays that hit the object, so `costs` can be the number of measured pixels per view
    (len(view.targets)) or last iteration's valid-path counts.  Returns this rank's views in ascending order."""
    costs = [float(c) for c in costs]
    order = sorted(range(len(costs)), key=lambda k: (-costs[k], k))
    load = [0.0] * world
    mine = []
    for k in order:
        r = min(range(world), key=lambda j: (load[j], j))
        load[r] += costs[k]
        if r == rank:
            mine.append(k)
    return sorted(mine)


class PeerAllReduce:
    """In-place SUM all-reduce of float64 CUDA tensors of up to `max_doubles` elements through drt_comm_*.
    Collective constructor: every rank of `group` must create it at the same point."""

    def __init__(self, max_doubles, device, group=None):
        from . import _lib
        self._lib, self.device, self.group = _lib, torch.device(device), group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.capacity = int(max_doubles)
        self._h = None
        ok, err = 1, ""
        try:
            h = C.c_void_p()
            _lib.call("drt_comm_create", self.device.index or 0, self.rank, self.world, self.capacity, C.byref(h))
            self._h = h
            mine = C.create_string_buffer(64)
            _lib.call("drt_comm_handle", h, mine)
            mine = (socket.gethostname(), bytes(mine.raw))
        except Exception as e:  # no GPU / no IPC: agree on it with the others below
            ok, err, mine = 0, str(e), (socket.gethostname(), b"\0" * 64)
        everyone = [None] * self.world
        dist.all_gather_object(everyone, mine, group=group)
        if ok and len({h for h, _ in everyone}) == 1:
            try:
                _lib.call("drt_comm_connect", self._h, b"".join(b for _, b in everyone))
            except Exception as e:
                ok, err = 0, str(e)
        else:
            ok = 0
        flag = torch.tensor([ok], dtype=torch.int32, device=self.device)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)   # all ranks or none
        self.ok = bool(flag.item())
        self.error = err
        if not self.ok and self._h is not None:
            self.close()

    def __call__(self, t):
        if not (self.ok and t.is_cuda and t.dtype == torch.float64 and t.is_contiguous() and t.numel() <= self.capacity):
            return False
        self._lib.call("drt_comm_allreduce_sum_f64", self._h, C.c_void_p(t.data_ptr()), t.numel(),
                       C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream))
        return True

    def timed_out(self):
        """synchronises; True if a wait for a peer ever ran into the kernel's spin limit (results are then invalid)"""
        if not self.ok:
            return False
        out = C.c_int(0)
        self._lib.call("drt_comm_status", self._h, C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream), C.byref(out))
        return bool(out.value)

    def close(self):
        if self._h is not None:
            self._lib.load().drt_comm_destroy(self._h)
            self._h = None
        self.ok = False

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


_peer = {}
PEER_DEFAULT_MAX_WORLD = 8


def peer_allreduce(numel, device, group=None):
    """The process-wide PeerAllReduce for tensors of up to `numel` doubles (created collectively on first use).
    DRT_ALLREDUCE=nccl disables it, =peer forces it; by default it is used for all the GPUs of one box (up to 8 ranks:
    bit-exact rank-order sums; the all-reduce phase of a step is dominated by the ranks' arrival skew, not by the
    collective) and NCCL is used above or across hosts."""
    mode = os.environ.get("DRT_ALLREDUCE", "").lower()
    if mode in ("nccl", "torch") or (mode != "peer" and dist.get_world_size(group) > PEER_DEFAULT_MAX_WORLD):
        return None
    key = (id(group), str(device))
    p = _peer.get(key)
    if p is None or (p.ok and p.capacity < numel):
        if p is not None:
            p.close()
        p = _peer[key] = PeerAllReduce(max(int(numel), 1 << 16), device, group)
    return p if p.ok else None


def allreduce_grad(grad_V, loss=None, group=None):
    """In-place SUM all-reduce of grad_V [V,3] float64 (+ optionally a loss scalar riding along in
    the same message).  Returns (grad_V, loss)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return grad_V, loss
    buf = grad_V if loss is None else torch.cat([grad_V.reshape(-1), loss.reshape(-1).to(grad_V.dtype)])
    done = False
    if buf.is_cuda and buf.dtype == torch.float64:
        buf = buf if buf.is_contiguous() else buf.contiguous()
        p = peer_allreduce(buf.numel(), buf.device, group)
        done = p is not None and p(buf)
    if not done:
        dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=group)
    if loss is None:
        if buf is not grad_V:
            grad_V.copy_(buf.view_as(grad_V))
        return grad_V, None
    grad_V.copy_(buf[:grad_V.numel()].view_as(grad_V))
    return grad_V, buf[grad_V.numel():].clone()
