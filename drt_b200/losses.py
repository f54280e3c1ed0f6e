"""Loss_calculator.ray_loss (reference optim.py:91-108) as one autograd.Function on the fused kernels
(SURVEY.md 8(f) N2): forward = drt_trace_fwd + drt_ray_loss_grad (loss value and d loss/d out_dir in one
pass over the valid paths), backward = drt_trace_bwd scaled by the upstream scalar.  out_ori is detached in
the reference (optim.py:100), so only out_dir carries gradient."""
import ctypes as C

import torch

from . import _lib, optix
from . import DiffRender as _R

_ptr = optix._ptr


class RayLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, vertices, origin, ray_dir, screen, valid, mesh, int_ior, ext_ior):
        dev = mesh.device
        V = vertices.detach().contiguous()
        o, d = origin.detach().contiguous(), ray_dir.detach().contiguous()
        scr = screen.detach().contiguous()
        if not (V.dtype == o.dtype == d.dtype == scr.dtype == torch.float64):
            raise TypeError("ray_loss works in float64 like the reference (captured_data.py:9)")
        if o.shape != d.shape or o.shape != scr.shape or o.dim() != 2 or o.shape[1] != 3:
            raise ValueError("origin, ray_dir and screen must all be [N,3]")
        val = None
        if valid is not None:
            if valid.shape != (o.shape[0],):
                raise ValueError("valid must be [N]")
            val = valid.to(torch.bool).contiguous()
        n = o.shape[0]
        out_ori = torch.empty((n, 3), dtype=torch.float64, device=dev)
        out_dir = torch.empty((n, 3), dtype=torch.float64, device=dev)
        mask = torch.empty((n, 3), dtype=torch.bool, device=dev)
        rec = torch.empty((max(n, 1), 4), dtype=torch.int32, device=dev)
        rec_count = torch.empty(1, dtype=torch.int32, device=dev)
        g_dir = torch.empty((n, 3), dtype=torch.float64, device=dev)
        loss = torch.zeros(1, dtype=torch.float64, device=dev)
        st = optix._stream_ptr(dev)
        _lib.call("drt_trace_fwd", mesh._h, _ptr(V), _ptr(o), _ptr(d), n, float(ext_ior), float(int_ior), _ptr(out_ori),
                  _ptr(out_dir), _ptr(mask), _ptr(rec), _ptr(rec_count), C.c_void_p(0), st)
        # loss and d loss/d out_dir over the compact records of the valid paths only (rows of other rays stay
        # unwritten: the backward kernel reads exactly the recorded rows)
        _lib.call("drt_ray_loss_grad_rec", _ptr(out_ori), _ptr(out_dir), _ptr(scr), _ptr(val), _ptr(rec), _ptr(rec_count), n,
                  _ptr(g_dir), _ptr(loss), st)
        ctx.mesh, ctx.iors = mesh, (float(ext_ior), float(int_ior))
        ctx.save_for_backward(V, o, d, rec, rec_count, g_dir)
        return loss[0]

    @staticmethod
    def backward(ctx, g_loss):
        V, o, d, rec, rec_count, g_dir = ctx.saved_tensors
        grad_V = torch.zeros_like(V)
        mesh = ctx.mesh
        _lib.call("drt_trace_bwd", mesh._h, _ptr(V), _ptr(o), _ptr(d), o.shape[0], ctx.iors[0], ctx.iors[1], _ptr(rec),
                  _ptr(rec_count), C.c_void_p(0), _ptr(g_dir), _ptr(grad_V), optix._stream_ptr(mesh.device))
        return grad_V * g_loss, None, None, None, None, None, None, None


def ray_loss(scene, origin, ray_dir, screen, valid=None):
    """sum over valid & traced rays of || out_dir - normalize(screen - out_ori) ||^2  (optim.py:96-106)."""
    return RayLoss.apply(scene.vertices, origin, ray_dir, screen, valid, scene.optix_mesh, _R.intIOR, _R.extIOR)
