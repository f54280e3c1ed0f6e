"""PyTorch restatement of the reference's silhouette-edge sampling -- TEST INFRASTRUCTURE ONLY.

silhouette_edge (DiffRender.py:445-457), primary_visibility (:459-479) and primary_edge_sample (:189-267) as plain torch ops on
top of ANY intersect callable with Scene.optix_intersect's contract, checked against outputs of the unmodified reference
(tests/golden/silhouette_hand_vh.npz, made by oracle/make_golden.py).  The product path is drt_b200/silhouette.py on the fused
kernels of csrc/silhouette.cuh; nothing under drt_b200/ imports this module.
"""
import torch

from drt_b200.silhouette import edge_face_norm


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
