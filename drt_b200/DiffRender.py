"""B200-native `DiffRender`: same Scene contract as the reference module of this name, with the ray
path (reference DiffRender.py:386-392, 420-438, 492-546 and optix_extend.cpp) replaced by
libdrt_b200's fused CUDA kernels.  `import drt_b200.DiffRender as Render` is the drop-in for
`import DiffRender as Render` (optim.py:6): module globals intIOR / resy / resx / device / Float are
assignable after import (optim.py:178-182) and read at call time.

No PyTorch op graph on the ray path: Scene.render_transparent is ONE autograd.Function whose
forward is one kernel launch and whose backward is one kernel launch.
"""
import ctypes as C

import numpy as np
import torch

from . import _lib, silhouette, trimesh_lite
from . import optix  # the plugin module the reference JIT-builds (DiffRender.py:5-6)

debug = False
# render resolution (DiffRender.py:16-17)
resy = 960
resx = 1280
Float = torch.float64
device = "cuda"
extIOR, intIOR = 1.00029, 1.5  # DiffRender.py:21

_ptr = optix._ptr


class Ray:
    """SoA ray set + original pixel index (reference DiffRender.py:269-283)."""

    def __init__(self, origin, direction, ray_ind=None):
        self.origin = origin
        self.direction = direction
        self.ray_ind = torch.arange(len(origin), device=origin.device) if ray_ind is None else ray_ind

    def select(self, mask):
        return Ray(self.origin[mask], self.direction[mask], self.ray_ind[mask])

    def __len__(self):
        return len(self.ray_ind)


class RefractTrace(torch.autograd.Function):
    """(vertices, origin, ray_dir) -> (out_ori, out_dir, mask); d/d vertices only.

    forward  = drt_trace_fwd : Q1 -> refract -> Q2 -> refract -> Q3, one launch
    backward = drt_trace_bwd : replay of the cached hit records + analytic Jacobian + scatter-add
    """

    @staticmethod
    def forward(ctx, vertices, origin, ray_dir, mesh, int_ior, ext_ior):
        dev = mesh.device
        optix.check_on(dev, vertices=vertices, origin=origin, ray_dir=ray_dir)
        V = vertices.detach()
        if V.dtype != torch.float64 or origin.dtype != torch.float64 or ray_dir.dtype != torch.float64:
            raise TypeError("render_transparent works in float64 like the reference (DiffRender.py:19)")
        V = V.contiguous()
        o = origin.detach().contiguous()
        d = ray_dir.detach().contiguous()
        if o.dim() != 2 or o.shape[1] != 3 or o.shape != d.shape:
            raise ValueError(f"origin/ray_dir must both be [N,3], got {tuple(o.shape)} and {tuple(d.shape)}")
        if V.shape[0] != mesh.n_verts:
            raise ValueError(f"vertices has {V.shape[0]} rows, the mesh was built with {mesh.n_verts}")
        n = o.shape[0]
        out_ori = torch.empty((n, 3), dtype=torch.float64, device=dev)
        out_dir = torch.empty((n, 3), dtype=torch.float64, device=dev)
        mask = torch.empty((n, 3), dtype=torch.bool, device=dev)
        # compact hit records (ray, tri1, tri2, 0) of the valid paths + their count, for backward
        need_rec = vertices.requires_grad
        rec = torch.empty((max(n, 1), 4), dtype=torch.int32, device=dev) if need_rec else None
        rec_count = torch.empty(1, dtype=torch.int32, device=dev) if need_rec else None
        st = optix._stream_ptr(dev)
        _lib.call("drt_trace_fwd", mesh._h, _ptr(V), _ptr(o), _ptr(d), n, float(ext_ior), float(int_ior), _ptr(out_ori),
                  _ptr(out_dir), _ptr(mask), _ptr(rec), _ptr(rec_count), C.c_void_p(0), st)
        ctx.mesh = mesh
        ctx.iors = (float(ext_ior), float(int_ior))
        ctx.save_for_backward(V, o, d, rec, rec_count)
        ctx.mark_non_differentiable(mask)
        ctx.set_materialize_grads(False)
        return out_ori, out_dir, mask

    @staticmethod
    def backward(ctx, g_ori, g_dir, _g_mask):
        V, o, d, rec, rec_count = ctx.saved_tensors
        grad_V = torch.zeros_like(V)
        if (g_ori is None and g_dir is None) or rec is None:
            return grad_V, None, None, None, None, None
        n = o.shape[0]
        if g_dir is None:
            g_dir = torch.zeros_like(o)
        g_dir = g_dir.contiguous()
        g_ori = None if g_ori is None else g_ori.contiguous()
        mesh = ctx.mesh
        optix.check_on(mesh.device, grad_out_dir=g_dir, grad_out_ori=g_ori)
        _lib.call("drt_trace_bwd", mesh._h, _ptr(V), _ptr(o), _ptr(d), n, ctx.iors[0], ctx.iors[1], _ptr(rec),
                  _ptr(rec_count), _ptr(g_ori), _ptr(g_dir), _ptr(grad_V), optix._stream_ptr(mesh.device))
        return grad_V, None, None, None, None, None


class RefractTraceSmooth(torch.autograd.Function):
    """OPTIONAL non-parity mode: (vertices, normals, origin, ray_dir) -> (out_ori, out_dir, mask) with the shading normal
    interpolated from the vertex normals (the code the reference keeps commented out, DiffRender.py:107-114).
    forward = drt_trace_fwd_smooth, backward = drt_trace_bwd_smooth: d/d vertices (through the hit distances) and
    d/d normals (the Jacobian w.r.t. the interpolated normals), both scatter-added per vertex."""

    @staticmethod
    def forward(ctx, vertices, normals, origin, ray_dir, mesh, int_ior, ext_ior):
        dev = mesh.device
        optix.check_on(dev, vertices=vertices, normals=normals, origin=origin, ray_dir=ray_dir)
        if any(t.dtype != torch.float64 for t in (vertices, normals, origin, ray_dir)):
            raise TypeError("render_transparent works in float64 like the reference (DiffRender.py:19)")
        V, VN = vertices.detach().contiguous(), normals.detach().contiguous()
        o, d = origin.detach().contiguous(), ray_dir.detach().contiguous()
        if o.dim() != 2 or o.shape[1] != 3 or o.shape != d.shape:
            raise ValueError(f"origin/ray_dir must both be [N,3], got {tuple(o.shape)} and {tuple(d.shape)}")
        if V.shape[0] != mesh.n_verts or VN.shape != V.shape:
            raise ValueError(f"vertices / normals must be [{mesh.n_verts},3], got {tuple(V.shape)} and {tuple(VN.shape)}")
        n = o.shape[0]
        out_ori = torch.empty((n, 3), dtype=torch.float64, device=dev)
        out_dir = torch.empty((n, 3), dtype=torch.float64, device=dev)
        mask = torch.empty((n, 3), dtype=torch.bool, device=dev)
        rec = torch.empty((max(n, 1), 4), dtype=torch.int32, device=dev)
        rec_count = torch.empty(1, dtype=torch.int32, device=dev)
        _lib.call("drt_trace_fwd_smooth", mesh._h, _ptr(V), _ptr(VN), _ptr(o), _ptr(d), n, float(ext_ior), float(int_ior), _ptr(out_ori),
                  _ptr(out_dir), _ptr(mask), _ptr(rec), _ptr(rec_count), optix._stream_ptr(dev))
        ctx.mesh = mesh
        ctx.iors = (float(ext_ior), float(int_ior))
        ctx.save_for_backward(V, VN, o, d, rec, rec_count)
        ctx.mark_non_differentiable(mask)
        ctx.set_materialize_grads(False)
        return out_ori, out_dir, mask

    @staticmethod
    def backward(ctx, g_ori, g_dir, _g_mask):
        V, VN, o, d, rec, rec_count = ctx.saved_tensors
        grad_V, grad_VN = torch.zeros_like(V), torch.zeros_like(VN)
        if g_ori is None and g_dir is None:
            return grad_V, grad_VN, None, None, None, None, None
        if g_dir is None:
            g_dir = torch.zeros_like(o)
        g_dir = g_dir.contiguous()
        g_ori = None if g_ori is None else g_ori.contiguous()
        mesh = ctx.mesh
        optix.check_on(mesh.device, grad_out_dir=g_dir, grad_out_ori=g_ori)
        _lib.call("drt_trace_bwd_smooth", mesh._h, _ptr(V), _ptr(VN), _ptr(o), _ptr(d), o.shape[0], ctx.iors[0], ctx.iors[1], _ptr(rec),
                  _ptr(rec_count), _ptr(g_ori), _ptr(g_dir), _ptr(grad_V), _ptr(grad_VN), optix._stream_ptr(mesh.device))
        return grad_V, grad_VN, None, None, None, None, None


class PlaneHit(torch.autograd.Function):
    """OPTIONAL: (out_ori, out_dir, mask) -> points where the exit rays of the valid paths meet a background plane
    (drt_plane_hit / drt_plane_hit_bwd).  Not part of the reference's loss."""

    @staticmethod
    def forward(ctx, out_ori, out_dir, mask, plane_point, plane_normal):
        dev = out_ori.device
        o, d = out_ori.detach().contiguous(), out_dir.detach().contiguous()
        m = mask.contiguous()
        if m.dim() != 2 or m.shape[1] != 3 or o.shape != d.shape or o.shape[0] != m.shape[0] or o.dtype != torch.float64 or not o.is_cuda:
            raise ValueError("plane_hit takes render_transparent's outputs: float64 [N,3] CUDA out_ori / out_dir and bool [N,3] mask")
        plane = (C.c_double * 6)(*[float(x) for x in plane_point], *[float(x) for x in plane_normal])
        n = o.shape[0]
        pts = torch.empty((n, 3), dtype=torch.float64, device=dev)
        front = torch.empty(n, dtype=torch.bool, device=dev)
        with torch.cuda.device(dev):
            _lib.call("drt_plane_hit", _ptr(o), _ptr(d), _ptr(m), n, plane, _ptr(pts), _ptr(front), optix._stream_ptr(dev))
        ctx.plane = plane
        ctx.save_for_backward(o, d, m)
        ctx.mark_non_differentiable(front)
        return pts, front

    @staticmethod
    def backward(ctx, g_pts, _g_front):
        o, d, m = ctx.saved_tensors
        g_o, g_d = torch.empty_like(o), torch.empty_like(d)
        with torch.cuda.device(o.device):
            _lib.call("drt_plane_hit_bwd", _ptr(o), _ptr(d), _ptr(m), o.shape[0], ctx.plane, _ptr(g_pts.contiguous()), _ptr(g_o), _ptr(g_d),
                      optix._stream_ptr(o.device))
        return g_o, g_d, None, None, None


def vertex_normals(vertices, faces):
    """Scene.init_VN (DiffRender.py:319-336) with JIT_corner_angles (:165-187): vertex normal = normalised sum of the
    adjacent unit face normals weighted by the (detached) corner angle; differentiable w.r.t. `vertices` through the face
    normals, like the reference (`triangles` is not detached there, the weights are)."""
    tri = vertices[faces]
    u, v, w = tri[:, 1] - tri[:, 0], tri[:, 2] - tri[:, 0], tri[:, 2] - tri[:, 1]
    face_n = torch.linalg.cross(u, v, dim=1)
    face_n = face_n / face_n.norm(dim=1, keepdim=True)
    with torch.no_grad():
        un, vn, wn = (x / x.norm(dim=1, keepdim=True) for x in (u, v, w))
        a0 = torch.acos((un * vn).sum(1).clamp(-1, 1))
        a1 = torch.acos((-un * wn).sum(1).clamp(-1, 1))
        angles = torch.stack([a0, a1, np.pi - a0 - a1], dim=1)                  # [F,3]
    vert_n = torch.zeros_like(vertices).index_add(0, faces.reshape(-1), (angles.unsqueeze(2) * face_n.unsqueeze(1)).reshape(-1, 3))
    return vert_n / vert_n.norm(dim=1, keepdim=True)


class Scene:
    """Same public surface as the reference Scene (DiffRender.py:298-546) for the methods optim.py
    drives: ctor, update_mesh, update_verticex, render_transparent, render_mask, optix_intersect,
    silhouette_edge, primary_visibility, dihedral_angle; attributes vertices, faces, mesh, mean_len."""

    def __init__(self, mesh_path=None, cuda_device=0, vertices=None, faces=None):
        self.cuda_device = cuda_device
        self.optix_mesh = optix.optix_mesh(cuda_device)
        self.refit = False  # True: update_verticex refits the BVH instead of rebuilding it
        # OPTIONAL non-parity mode: shade with normals interpolated from the vertex normals instead of the
        # flat face normal the reference uses (DiffRender.py:103-104 live, :107-114 commented out)
        self.smooth_normals = False
        self._mesh, self._mesh_dirty = None, False
        if mesh_path is not None:
            self.update_mesh(mesh_path)
        elif vertices is not None:
            self.set_mesh(vertices, faces)

    @property
    def _dev(self):
        return self.optix_mesh.device

    # DiffRender.py:303-317
    def update_mesh(self, mesh_path):
        mesh = trimesh_lite.load(mesh_path, process=False)
        assert mesh.is_watertight
        self.set_mesh(mesh.vertices, mesh.faces, mesh)

    @property
    def mesh(self):
        """The host-side mesh (`scene.mesh.export(path)`, optim.py:50,226) with CURRENT vertices.  The reference copies
        the vertices to the host on every update_verticex (DiffRender.py:381); here the D2H copy happens on first
        access after an update, so an unchanged optim.py exports the optimised mesh and the hot loop pays nothing."""
        if self._mesh_dirty:
            self._mesh.vertices = self.vertices.detach().cpu().numpy()
            self._mesh_dirty = False
        return self._mesh

    @mesh.setter
    def mesh(self, m):
        self._mesh = m
        self._mesh_dirty = False

    def set_mesh(self, vertices, faces, mesh=None):
        self.mesh = mesh if mesh is not None else trimesh_lite.TriMesh(
            vertices.detach().cpu().numpy() if isinstance(vertices, torch.Tensor) else vertices,
            faces.detach().cpu().numpy() if isinstance(faces, torch.Tensor) else faces)
        self.vertices = torch.as_tensor(np.asarray(self.mesh.vertices), dtype=Float).to(self._dev)
        self.faces = torch.as_tensor(np.asarray(self.mesh.faces), dtype=torch.long).to(self._dev)
        # the float32 cast of DiffRender.py:311 happens inside the library (drt_bvh_build_f64)
        self.optix_mesh.update_mesh(self.faces.to(torch.int32), self.vertices.detach())
        # a corrupt face list must not render silently (the library clamps bad indices so that no kernel can fault);
        # checked at load time only -- the read-back synchronises, so it stays off the per-iteration path
        bad = self.optix_mesh.bad_indices()
        if bad:
            raise ValueError(f"{bad} face indices are outside [0, {self.vertices.shape[0]}): corrupt mesh")
        self._edges_ready = False

    # DiffRender.py:378-384 (without the per-iteration D2H copy and the dead init_VN)
    def update_verticex(self, vertices):
        self.optix_mesh.update_vert(vertices.detach(), refit=self.refit)
        self.vertices = vertices
        self._mesh_dirty = True

    @property
    def triangles(self):
        return self.vertices[self.faces]

    def sync_mesh(self):
        """Kept for callers of the earlier API: `scene.mesh` itself is always current now."""
        return self.mesh

    # DiffRender.py:386-392
    def optix_intersect(self, ray):
        o = ray.origin.detach().to(torch.float32)
        d = ray.direction.detach().to(torch.float32)
        T, faces_ind = self.optix_mesh.intersect(torch.cat([o, d], dim=1))
        return faces_ind.to(torch.long), T > 0

    # DiffRender.py:420-432
    def render_transparent(self, origin, ray_dir):
        self.optix_mesh.set_image_size(resy, resx)  # module globals, assigned by the caller like optim.py:179-180
        if self.smooth_normals:
            return RefractTraceSmooth.apply(self.vertices, self.init_VN(), origin, ray_dir, self.optix_mesh, intIOR, extIOR)
        return RefractTrace.apply(self.vertices, origin, ray_dir, self.optix_mesh, intIOR, extIOR)

    # DiffRender.py:319-336 (dead in the reference's live path, F2/F12; only the smooth-normal mode needs it)
    def init_VN(self):
        self.normals = vertex_normals(self.vertices, self.faces)
        return self.normals

    def render_background(self, origin, ray_dir, plane_point, plane_normal):
        """OPTIONAL (north-star wording, not in the reference: SURVEY.md F5): the two-bounce path followed by the
        intersection of the exit ray with a background plane -> (points [N,3], front bool[N])."""
        out_ori, out_dir, mask = self.render_transparent(origin, ray_dir)
        return PlaneHit.apply(out_ori, out_dir, mask, plane_point, plane_normal)

    # DiffRender.py:434-438
    def render_mask(self, origin, ray_dir):
        _, hitted = self.optix_intersect(Ray(origin, ray_dir))
        return hitted.to(Float)

    # ---- silhouette / smoothness side of the Scene contract -------------
    def _edges(self):
        if not self._edges_ready:  # DiffRender.py:338-355, built lazily (only vh_loss / sm_loss need it)
            self.Edges, self.E2F, self._mean_len = silhouette.build_edge_tables(self.mesh, self.faces, self._dev)
            self._E2F32 = self.E2F.to(torch.int32).contiguous()   # what drt_silhouette_classify reads
            self._edges_ready = True
        return self.Edges, self.E2F

    @property
    def mean_len(self):  # optim.py:129
        self._edges()
        return self._mean_len

    def dihedral_angle(self):  # DiffRender.py:440-443 (cosine per edge; optim.py:85)
        _, E2F = self._edges()
        return silhouette.dihedral_cos(self.vertices, E2F)

    def silhouette_edge(self, origin):  # DiffRender.py:445-457 -> drt_silhouette_classify
        Edges, _ = self._edges()
        return silhouette.silhouette_edges(self.vertices, Edges, self._E2F32, origin)

    def primary_visibility(self, silhouette_edge, camera_M, origin, detach_depth=False):  # DiffRender.py:459-479 -> drt_silhouette_sample
        return silhouette.primary_visibility(self.optix_mesh, self.vertices, silhouette_edge, camera_M, origin, resy, resx, detach_depth)
