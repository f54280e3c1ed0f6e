"""Silhouette-edge visibility sampling on top of the closest-hit query.

Same quantities as the reference's edge machinery -- init_edge (DiffRender.py:338-355), edge_face_norm
(:150-163), silhouette_edge (:445-457), primary_visibility (:459-479), primary_edge_sample (:189-267),
dihedral_angle (:440-443) -- written against drt_b200.trimesh_lite and Scene.optix_intersect, i.e. the
two rays per silhouette edge go through drt_closest_hit (un-normalised directions are fine there).
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


def silhouette_edges(vertices, Edges, E2F, origin):
    """Edges whose two faces face opposite ways as seen from `origin` (DiffRender.py:445-457)."""
    assert origin.dim() == 1
    v = vertices.detach()
    n1, n2 = edge_face_norm(v, E2F)
    d1 = (n1 * (origin - v[E2F[:, 0, 0]])).sum(dim=1)
    d2 = (n2 * (origin - v[E2F[:, 1, 0]])).sum(dim=1)
    return Edges[torch.logical_xor(d1 > 0, d2 > 0)]


class EdgeSample(torch.autograd.Function):
    """One sample per silhouette edge: the edge midpoint in pixels; two probe rays one pixel either
    side of the edge decide which side is covered.  Hand-written backward (DiffRender.py:189-267)."""

    @staticmethod
    def forward(ctx, E_pos, intersect_fn, camera_M, ray_origin, ray_cls):
        assert ray_origin.dim() == 1
        n = E_pos.shape[0]
        _, _, R_inv, K_inv = camera_M
        a, b = E_pos[:, 0], E_pos[:, 1]                      # [n,2] pixel positions of the edge ends
        mid = 0.5 * (a + b)
        nrm = torch.stack((a[:, 1] - b[:, 1], b[:, 0] - a[:, 0]), dim=1)   # edge normal in the image
        unit = nrm / nrm.norm(dim=1, keepdim=True)
        probes = torch.cat((mid + unit, mid - unit), dim=0)  # [2n,2]: upper side then lower side
        ones = torch.ones((2 * n, 1), dtype=E_pos.dtype, device=E_pos.device)
        cam = torch.cat((probes, ones), dim=1) @ K_inv.T     # pixel at z = 1
        world = torch.cat((cam, ones), dim=1) @ R_inv.T
        direction = world[:, :3] - ray_origin.view(1, 3)     # NOT normalised, like the reference (:222)
        _, hit = intersect_fn(ray_cls(ray_origin.view(1, 3).expand_as(direction), direction))
        cover = hit.to(E_pos.dtype)
        f = cover[:n] - cover[n:]
        # dE[i, endpoint, coord] = -nrm[i, coord]  (DiffRender.py:243-249)
        dE = torch.stack((torch.stack((-nrm[:, 0], -nrm[:, 0]), dim=1), torch.stack((-nrm[:, 1], -nrm[:, 1]), dim=1)), dim=2)
        dE = dE * f.view(-1, 1, 1)
        valid = f.abs() > 1e-5
        index = mid[valid].to(torch.long)
        output = 0.5 * torch.ones(index.shape[0], dtype=torch.float32, device=E_pos.device)
        ctx.mark_non_differentiable(index)
        ctx.save_for_backward(dE, valid)
        return index, output

    @staticmethod
    def backward(ctx, _g_index, g_output):
        dE, valid = ctx.saved_tensors
        g = dE.clone()
        g[valid] = g[valid] * g_output.view(-1, 1, 1).to(g.dtype)
        return g, None, None, None, None


def primary_visibility(vertices, silhouette_edge, camera_M, origin, intersect_fn, ray_cls, resy, resx, detach_depth=False):
    """Project the silhouette edges, sample them, drop samples outside the image (DiffRender.py:459-479).
    -> (index long[m,2] pixel (x,y), output float[m])."""
    R, K, _, _ = camera_M
    V = vertices[silhouette_edge.reshape(-1)]
    ones = torch.ones((V.shape[0], 1), dtype=V.dtype, device=V.device)
    cam = R @ torch.cat((V, ones), dim=1).T                   # [4,2n]
    xyz = cam[:3]
    if detach_depth:
        xyz = torch.cat((xyz[:2], xyz[2:3].detach()), dim=0)
    pix = K @ xyz
    E_pos = (pix[:2] / pix[2]).T.reshape(-1, 2, 2)
    index, output = EdgeSample.apply(E_pos, intersect_fn, camera_M, origin, ray_cls)
    keep = (index[:, 0] < resx - 1) & (index[:, 1] < resy - 1) & (index[:, 0] >= 0) & (index[:, 1] >= 0)
    return index[keep], output[keep]
