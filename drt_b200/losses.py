"""Loss_calculator.ray_loss (reference optim.py:91-108) as one autograd.Function on the fused kernels
(SURVEY.md 8(f) N2).  out_ori is detached in the reference (optim.py:100), so only out_dir carries gradient.

`ray_loss` / `ray_loss_view` = drt_ray_loss_step: the forward wavefront, the loss and the vertex gradient in ONE
library call with no dense per-ray output at all (rays that miss write nothing, d loss/d out_dir is never
stored); the gradient is computed together with the loss and scaled by the upstream scalar in backward().
The ray origin may be one row per ray (the reference layout), an expanded / single-row tensor (a pinhole view
has ONE origin, captured_data.py:38) or one row per view; the screen targets may be the reference's dense
(screen_pixel, valid) pair or `SparseTargets` (only the measured pixels, captured_data.py:104).

`ray_loss_rec` = the earlier three-call route (drt_trace_fwd -> drt_ray_loss_grad_rec -> drt_trace_bwd), kept
for A/B measurements: it still materialises out_ori / out_dir / mask / d loss/d out_dir."""
import ctypes as C

import torch

from . import _lib, optix
from . import DiffRender as _R

_ptr = optix._ptr


class RayLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, vertices, origin, ray_dir, screen, valid, mesh, int_ior, ext_ior):
        dev = mesh.device
        optix.check_on(dev, vertices=vertices, origin=origin, ray_dir=ray_dir, screen=screen, valid=valid)
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


def ray_loss_rec(scene, origin, ray_dir, screen, valid=None):
    """The three-call route (dense out_ori/out_dir/mask and d loss/d out_dir in memory); same value and gradient."""
    return RayLoss.apply(scene.vertices, origin, ray_dir, screen, valid, scene.optix_mesh, _R.intIOR, _R.extIOR)


class SparseTargets:
    """Screen targets of the measured pixels only: `idx` int32 [n] ray indices, strictly ascending, `xyz` float64
    [n,3] screen points.  Equivalent to the reference's dense pair with valid = False everywhere else
    (captured_data.py:101-104: valid = screen_pixel[:,0] != 0)."""

    def __init__(self, idx, xyz):
        if idx.dtype != torch.int32 or idx.dim() != 1:
            raise TypeError("SparseTargets.idx must be int32 [n]")
        if xyz.dtype != torch.float64 or xyz.shape != (idx.shape[0], 3):
            raise TypeError("SparseTargets.xyz must be float64 [n,3]")
        self.idx, self.xyz = idx.contiguous(), xyz.contiguous()

    @staticmethod
    def from_dense(screen, valid=None):
        """valid rows of a dense (screen_pixel [N,3], valid [N]) pair -> SparseTargets on the same device."""
        if valid is None:
            valid = screen[:, 0] != 0  # captured_data.py:104
        idx = torch.nonzero(valid, as_tuple=False).reshape(-1)
        return SparseTargets(idx.to(torch.int32), screen[idx].to(torch.float64))

    def to(self, device, non_blocking=False):
        return SparseTargets(self.idx.to(device, non_blocking=non_blocking), self.xyz.to(device, non_blocking=non_blocking))

    def pin_memory(self):
        return SparseTargets(self.idx.pin_memory(), self.xyz.pin_memory())

    def __len__(self):
        return self.idx.shape[0]


def origin_rows(origin, n_rays):
    """-> (rows float64 [r,3] contiguous, rays_per_origin).  Accepts the reference's [N,3] layout, an expanded
    (stride-0) tensor as captured_data.generate_ray returns it (captured_data.py:38), or r rows with r | N."""
    if origin.dim() == 1:
        origin = origin.reshape(1, 3)
    if origin.dim() != 2 or origin.shape[1] != 3:
        raise ValueError(f"origin must be [N,3], [r,3] or [3], got {tuple(origin.shape)}")
    r = origin.shape[0]
    if r == n_rays and n_rays > 1 and origin.stride(0) == 0:
        return origin[:1].contiguous(), max(n_rays, 1)
    if r == n_rays or n_rays == 0:
        return origin.contiguous(), 1
    if r < 1 or n_rays % r:
        raise ValueError(f"{r} origin rows do not divide {n_rays} rays")
    return origin.contiguous(), n_rays // r


class RayLossStep(torch.autograd.Function):
    """loss and d loss/d vertices from ONE drt_ray_loss_step call (six launches, no dense per-ray output)."""

    @staticmethod
    def forward(ctx, vertices, origin, rpo, ray_dir, screen, valid, targets, mesh, int_ior, ext_ior, n_paths, ev_after_fwd, image_size,
                tile_beams=None):
        dev = mesh.device
        optix.check_on(dev, vertices=vertices, origin=origin, ray_dir=ray_dir, screen=screen, valid=valid, n_paths=n_paths, tile_beams=tile_beams,
                       **({"targets.idx": targets.idx, "targets.xyz": targets.xyz} if targets is not None else {}))
        V = vertices.detach().contiguous()
        o, d = origin.detach(), ray_dir.detach().contiguous()
        if not (V.dtype == o.dtype == d.dtype == torch.float64):
            raise TypeError("ray_loss works in float64 like the reference (captured_data.py:9)")
        if d.dim() != 2 or d.shape[1] != 3:
            raise ValueError("ray_dir must be [N,3]")
        if V.shape[0] != mesh.n_verts:
            raise ValueError(f"vertices has {V.shape[0]} rows, the mesh was built with {mesh.n_verts}")
        n = d.shape[0]
        scr = val = idx = xyz = None
        n_tgt = 0
        if targets is not None:
            mode, idx, xyz, n_tgt = 1, targets.idx, targets.xyz, len(targets)
        else:
            mode = 0
            scr = screen.detach().contiguous()
            if scr.dtype != torch.float64 or scr.shape != d.shape:
                raise ValueError("screen must be float64 [N,3]")
            if valid is not None:
                if valid.shape != (n,):
                    raise ValueError("valid must be [N]")
                val = valid.to(torch.bool).contiguous()
        if tile_beams is not None and (tile_beams.dtype != torch.float32 or not tile_beams.is_contiguous()
                                       or tile_beams.numel() != _lib.load().drt_tile_beams_floats(n)):
            raise ValueError("tile_beams must be the float32 buffer prepare_tile_beams returned for this batch")
        need_grad = ctx.needs_input_grad[0]
        loss = torch.zeros(1, dtype=torch.float64, device=dev)
        grad_V = torch.zeros_like(V) if need_grad else None
        _lib.call("drt_ray_loss_step_beams", mesh._h, _ptr(V), _ptr(o), int(rpo), _ptr(d), n, float(ext_ior), float(int_ior), mode,
                  _ptr(scr), _ptr(val), _ptr(idx), _ptr(xyz), n_tgt, int(image_size[1]) if image_size else 0,
                  int(image_size[0]) if image_size else 0, _ptr(tile_beams), _ptr(loss), _ptr(grad_V), _ptr(n_paths),
                  C.c_void_p(ev_after_fwd or 0), optix._stream_ptr(dev))
        st = torch.cuda.current_stream(dev)
        for t in (V, o, d, scr, val, idx, xyz, tile_beams):  # consumed asynchronously on the stream
            if t is not None:
                t.record_stream(st)
        ctx.save_for_backward(grad_V)
        return loss[0]

    @staticmethod
    def backward(ctx, g_loss):
        (grad_V,) = ctx.saved_tensors
        return (grad_V * g_loss if grad_V is not None else None,) + (None,) * 13


def prepare_tile_beams(origin, ray_dir, image_size=None):
    """Per-tile direction intervals of a ray batch (drt_tile_beams), to be computed ONCE per view set and passed to every later
    `ray_loss(..., tile_beams=...)` on the same rays: the view sets of DRT are fixed for a whole optimisation
    (captured_data.py:94-108, optim.py:95), and what the entry query's beam culling needs from the rays does not depend on the
    mesh.  `origin`, `ray_dir`, `image_size` exactly as they will be given to `ray_loss`.  -> float32 device tensor (1.5 B/ray)."""
    rows, rpo = origin_rows(origin, ray_dir.shape[0])
    d = ray_dir.detach().contiguous()
    optix.check_on(d.device, origin=rows, ray_dir=d)
    if not d.is_cuda or rows.dtype != torch.float64 or d.dtype != torch.float64 or d.dim() != 2 or d.shape[1] != 3:
        raise TypeError("prepare_tile_beams needs float64 CUDA rays [N,3]")
    n = d.shape[0]
    beams = torch.empty(_lib.load().drt_tile_beams_floats(n), dtype=torch.float32, device=d.device)
    with torch.cuda.device(d.device):
        _lib.call("drt_tile_beams", _ptr(rows), int(rpo), _ptr(d), n, int(image_size[1]) if image_size else 0,
                  int(image_size[0]) if image_size else 0, _ptr(beams), optix._stream_ptr(d.device))
    st = torch.cuda.current_stream(d.device)
    rows.record_stream(st)
    d.record_stream(st)
    return beams


def ray_loss(scene, origin, ray_dir, screen=None, valid=None, targets=None, n_paths=None, ev_after_fwd=None, image_size=None,
             tile_beams=None):
    """sum over valid & traced rays of || out_dir - normalize(screen - out_ori) ||^2  (optim.py:96-106).

    Either the reference's dense pair (`screen` [N,3], `valid` [N] or None) or `targets` = SparseTargets.
    `origin`: [N,3], expanded/[1,3] (one origin for all rays) or [r,3] with r | N (ray i starts at row i // (N/r)).
    `n_paths`: optional int32[1] device tensor receiving the number of valid two-bounce paths.
    `image_size` = (resy, resx): optional hint that the rays are whole images in scanline order (captured_data.py:26-31);
    the entry query then works on 32-pixel tiles (4x8, else 8x4).  Same results either way.
    `tile_beams`: optional buffer from `prepare_tile_beams(origin, ray_dir, image_size)` for these very rays (fixed view sets):
    the beam pass then skips its scan of all ray directions.  Same results either way."""
    if (screen is None) == (targets is None):
        raise ValueError("give either screen (+valid) or targets")
    rows, rpo = origin_rows(origin, ray_dir.shape[0])
    return RayLossStep.apply(scene.vertices, rows, rpo, ray_dir, screen, valid, targets, scene.optix_mesh, _R.intIOR, _R.extIOR,
                             n_paths, ev_after_fwd, image_size, tile_beams)


def ray_loss_view(scene, view):
    """`view` = captured_data.CompactView (one origin row, ray_dir, sparse targets) on the scene's device."""
    return ray_loss(scene, view.origin, view.ray_dir, targets=view.targets, image_size=getattr(view, "image_size", None),
                    tile_beams=getattr(view, "tile_beams", None))


class SilhouetteLoss(torch.autograd.Function):
    """Loss_calculator.vh_loss (optim.py:67-80) for a list of views from ONE call of drt_silhouette_loss (up to 8 views per launch):
    value and vertex gradient, no intermediate tensors, no host synchronisation."""

    @staticmethod
    def forward(ctx, vertices, mesh, Edges, E2F32, views, resy, resx, detach_depth, n_samples):
        dev = mesh.device
        V = vertices.detach().contiguous()
        optix.check_on(dev, vertices=V, Edges=Edges, E2F=E2F32, n_samples=n_samples)
        if V.dtype != torch.float64 or Edges.dtype != torch.long or E2F32.dtype != torch.int32:
            raise TypeError("silhouette_loss needs float64 vertices, Edges long [E,2] and the int32 E2F table")
        keep, cols = [], [[] for _ in range(6)]
        for mask, camera_M, origin in views:
            R, K, R_inv, K_inv = (m.detach().to(torch.float64).contiguous() for m in camera_M)
            o = origin.detach().to(torch.float64).contiguous()
            mk = mask.detach().to(torch.float64).contiguous()
            optix.check_on(dev, R=R, K=K, R_inverse=R_inv, K_inverse=K_inv, origin=o, mask=mk)
            if mk.numel() != int(resy) * int(resx) or o.shape != (3,) or R.shape != (4, 4) or K.shape != (3, 3) or R_inv.shape != (4, 4) or K_inv.shape != (3, 3):
                raise ValueError("mask must hold resy*resx values, origin [3], camera_M = (R [4,4], K [3,3], R_inverse [4,4], K_inverse [3,3])")
            for c, t in zip(cols, (R, K, R_inv, K_inv, o, mk)):
                c.append(t.data_ptr())
            keep += [R, K, R_inv, K_inv, o, mk]
        n = len(views)
        arrays = [(C.c_void_p * max(n, 1))(*c) for c in cols]
        loss = torch.zeros(1, dtype=torch.float64, device=dev)
        grad_V = torch.zeros_like(V) if ctx.needs_input_grad[0] else None
        _lib.call("drt_silhouette_loss", mesh._h, _ptr(V), _ptr(Edges), _ptr(E2F32), Edges.shape[0], n, *arrays, int(resx), int(resy),
                  int(bool(detach_depth)), _ptr(loss), _ptr(grad_V), _ptr(n_samples), optix._stream_ptr(dev))
        st = torch.cuda.current_stream(dev)
        for t in [V] + keep:
            t.record_stream(st)
        ctx.save_for_backward(grad_V)
        return loss[0]

    @staticmethod
    def backward(ctx, g_loss):
        (grad_V,) = ctx.saved_tensors
        return (grad_V * g_loss if grad_V is not None else None,) + (None,) * 8


def silhouette_loss(scene, mask, camera_M=None, origin=None, detach_depth=True, n_samples=None):
    """`(mask.view(resy, resx)[index[:,1], index[:,0]] - output).abs().sum()` over the silhouette-edge samples (optim.py:74-79:
    silhouette_edge -> primary_visibility -> the L1 term), fused.  One view: `mask` = its silhouette image [resy*resx], `camera_M`,
    `origin` [3] = the camera centre.  Several views (optim.py:72 sums 8 per iteration): `mask` = a list of (mask, camera_M, origin)
    triples -- they run in one launch.  `n_samples`: optional int32[1] device tensor, incremented by the number of samples used."""
    views = mask if camera_M is None else [(mask, camera_M, origin)]
    Edges, _ = scene._edges()
    return SilhouetteLoss.apply(scene.vertices, scene.optix_mesh, Edges, scene._E2F32, list(views), _R.resy, _R.resx, detach_depth, n_samples)


class DihedralLoss(torch.autograd.Function):
    """Loss_calculator.sm_loss (optim.py:82-89) from one launch of drt_dihedral_loss."""

    @staticmethod
    def forward(ctx, vertices, E2F32):
        V = vertices.detach().contiguous()
        optix.check_on(V.device, E2F=E2F32)
        if V.dtype != torch.float64 or not V.is_cuda or E2F32.dtype != torch.int32:
            raise TypeError("smoothness_loss needs float64 CUDA vertices and the int32 E2F table")
        loss = torch.zeros(1, dtype=torch.float64, device=V.device)
        grad_V = torch.zeros_like(V) if ctx.needs_input_grad[0] else None
        with torch.cuda.device(V.device):
            _lib.call("drt_dihedral_loss", _ptr(V), _ptr(E2F32), E2F32.shape[0], _ptr(loss), _ptr(grad_V), optix._stream_ptr(V.device))
        V.record_stream(torch.cuda.current_stream(V.device))
        ctx.save_for_backward(grad_V)
        return loss[0]

    @staticmethod
    def backward(ctx, g_loss):
        (grad_V,) = ctx.saved_tensors
        return (grad_V * g_loss if grad_V is not None else None), None


def smoothness_loss(scene):
    """sum over edges of -log(1 + cos(dihedral angle))  (optim.py:85-87), fused with its gradient."""
    scene._edges()
    return DihedralLoss.apply(scene.vertices, scene._E2F32)
