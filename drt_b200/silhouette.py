"""Silhouette-edge visibility sampling on the fused kernels of csrc/silhouette.cuh.

Same quantities as the reference's edge machinery -- init_edge (DiffRender.py:338-355), edge_face_norm (:150-163),
silhouette_edge (:445-457), primary_visibility (:459-479), primary_edge_sample (:189-267), dihedral_angle (:440-443).
Edge classification is one kernel over the E2F table; projection + the two probe rays per silhouette edge + the in-image
filter are one kernel that walks the BVH itself; the backward is one kernel.  (A PyTorch restatement of the same functions,
used by the CPU tests, lives in oracle/silhouette_torch.py.)
"""
import numpy as np
import torch

from . import trimesh_lite


def build_edge_tables(mesh, faces_t, device):
    """-> (Edges long[E,2] unique undirected edges, E2F long[E,2,3] the two faces on each edge,
    mean_len float) -- DiffRender.py:342-355."""
    v = mesh.vertices
    e = mesh.edges
    mean_len = float(np.linalg.norm(v[e[:, 0]] - v[e[:, 1]], axis=1).mean())
    es = mesh.edges_sorted
    pairs = trimesh_lite.group_rows_pairs(es)
    edges = es[pairs[:, 0]]
    e2f_index = mesh.edges_face[pairs]
    Edges = torch.as_tensor(edges, dtype=torch.long, device=device)
    E2F = faces_t[torch.as_tensor(e2f_index, dtype=torch.long, device=device)]
    return Edges, E2F, mean_len


def _face_normals(vertices, tri):
    a, b, c = vertices[tri[:, 0]], vertices[tri[:, 1]], vertices[tri[:, 2]]
    n = torch.linalg.cross(b - a, c - a, dim=1)
    return n / n.norm(dim=1, keepdim=True)


def edge_face_norm(vertices, E2F):
    """Unit normals of the two faces adjacent to every edge (DiffRender.py:150-163)."""
    return _face_normals(vertices, E2F[:, 0]), _face_normals(vertices, E2F[:, 1])


def dihedral_cos(vertices, E2F):
    """cos of the dihedral angle per edge (DiffRender.py:440-443), differentiable."""
    n1, n2 = edge_face_norm(vertices, E2F)
    return (n1 * n2).sum(dim=1)


def silhouette_edges(vertices, Edges, E2F32, origin):
    """Scene.silhouette_edge (DiffRender.py:445-457) on drt_silhouette_classify: edges whose two faces face opposite ways as
    seen from `origin`, in ascending edge order like the reference's boolean-mask index."""
    import ctypes as C

    from . import _lib, optix
    if origin.dim() != 1 or origin.shape[0] != 3:
        raise ValueError("origin must be a [3] tensor (DiffRender.py:446)")
    V = vertices.detach().contiguous()
    dev = V.device
    optix.check_on(dev, Edges=Edges, E2F=E2F32, origin=origin)
    if V.dtype != torch.float64 or E2F32.dtype != torch.int32:
        raise TypeError("silhouette_edges works on float64 vertices and the int32 E2F table")
    o = origin.detach().to(torch.float64).contiguous()
    n_e = E2F32.shape[0]
    flags = torch.empty(n_e, dtype=torch.bool, device=dev)
    with torch.cuda.device(dev):
        _lib.call("drt_silhouette_classify", optix._ptr(V), optix._ptr(E2F32), n_e, optix._ptr(o), optix._ptr(flags), optix._stream_ptr(dev))
    return Edges[flags]


class EdgeVisibility(torch.autograd.Function):
    """Scene.primary_visibility + primary_edge_sample (DiffRender.py:459-479, 189-267) as two kernels: forward
    drt_silhouette_sample (projection, midpoint sample, the two probe rays through the BVH, in-image filter), backward
    drt_silhouette_backward (the reference's hand-written dE_pos chained through the projection into the vertices)."""

    @staticmethod
    def forward(ctx, vertices, mesh, sil_edges, camera_M, origin, resy, resx, detach_depth):
        from . import _lib, optix
        dev = mesh.device
        R, K, R_inv, K_inv = (m.detach().to(torch.float64).contiguous() for m in camera_M)
        V = vertices.detach().contiguous()
        E = sil_edges.contiguous()
        o = origin.detach().to(torch.float64).contiguous()
        optix.check_on(dev, vertices=V, silhouette_edge=E, R=R, K=K, R_inverse=R_inv, K_inverse=K_inv, origin=o)
        if V.dtype != torch.float64 or E.dtype != torch.long or E.dim() != 2 or E.shape[1] != 2:
            raise TypeError("primary_visibility needs float64 vertices and silhouette_edge long [k,2]")
        if R.shape != (4, 4) or K.shape != (3, 3) or R_inv.shape != (4, 4) or K_inv.shape != (3, 3) or o.shape != (3,):
            raise ValueError("camera_M must be (R [4,4], K [3,3], R_inverse [4,4], K_inverse [3,3]) and origin [3]")
        k = E.shape[0]
        index_xy = torch.empty((k, 2), dtype=torch.long, device=dev)
        f = torch.empty(k, dtype=torch.float64, device=dev)
        keep = torch.empty(k, dtype=torch.bool, device=dev)
        _lib.call("drt_silhouette_sample", mesh._h, optix._ptr(V), optix._ptr(E), k, optix._ptr(R), optix._ptr(K), optix._ptr(R_inv),
                  optix._ptr(K_inv), optix._ptr(o), int(resx), int(resy), optix._ptr(index_xy), optix._ptr(f), optix._ptr(keep),
                  optix._stream_ptr(dev))
        kept = torch.nonzero(keep, as_tuple=False).reshape(-1)     # ascending edge slot: the order of the reference's mask index
        index = index_xy[kept]
        output = torch.full((kept.shape[0],), 0.5, dtype=torch.float32, device=dev)   # DiffRender.py:240
        ctx.save_for_backward(V, E, R, K, f, kept)
        ctx.detach_depth = bool(detach_depth)
        ctx.mark_non_differentiable(index)
        return index, output

    @staticmethod
    def backward(ctx, _g_index, g_output):
        from . import _lib, optix
        V, E, R, K, f, kept = ctx.saved_tensors
        grad_V = torch.zeros_like(V)
        if g_output is not None and kept.shape[0]:
            g = g_output.detach().to(torch.float32).contiguous()
            with torch.cuda.device(V.device):
                _lib.call("drt_silhouette_backward", optix._ptr(V), optix._ptr(E), optix._ptr(R), optix._ptr(K), int(ctx.detach_depth),
                          optix._ptr(f), optix._ptr(kept), optix._ptr(g), kept.shape[0], optix._ptr(grad_V), optix._stream_ptr(V.device))
        return grad_V, None, None, None, None, None, None, None


def primary_visibility(mesh, vertices, silhouette_edge, camera_M, origin, resy, resx, detach_depth=False):
    """-> (index long[m,2] pixel (x,y) of the kept samples, output float32[m] = 0.5 with the reference's gradient)."""
    return EdgeVisibility.apply(vertices, mesh, silhouette_edge, camera_M, origin, resy, resx, detach_depth)
