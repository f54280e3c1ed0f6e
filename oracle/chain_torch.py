"""PyTorch restatement of the reference's refraction chain -- TEST INFRASTRUCTURE / BASELINE ONLY.

The reference implements render_transparent as ~500 PyTorch ops recorded by autograd
(DiffRender.py:420-432, 492-546).  This module restates that op chain (float64,
boolean-mask compaction between stages, index_put scatter at the end, autograd for the backward)
on top of ANY intersect callable with the plugin's contract (optix_extend.cpp:29-57), so that
  * the C oracle has an independent second restatement to be checked against (CPU), and
  * bench.py can time "the reference's approach on the same B200" with the query served by
    drt_closest_hit -- the R-GPU baseline of BASELINE.md -- without /root/reference on the box.
Nothing under drt_b200/ imports it.
"""
import torch

EXT_IOR = 1.00029  # DiffRender.py:21


def _dot(a, b):  # DiffRender.py:23-29
    return a[:, 0] * b[:, 0] + a[:, 1] * b[:, 1] + a[:, 2] * b[:, 2]


def _query(intersect, o, d):
    """Scene.optix_intersect (DiffRender.py:386-392): float32 cast, one [N,6] buffer, T>0 test."""
    T, ids = intersect(torch.cat([o.detach().float(), d.detach().float()], dim=1))
    return ids.long(), T > 0


def vertex_normals(vertices, faces):
    """Scene.init_VN + JIT_corner_angles (DiffRender.py:319-336, 165-187): sparse [V,F] matrix of detached corner angles times
    the unit face normals, row-normalised.  Only the OPTIONAL smooth-normal mode uses it (dead in the live path, F2)."""
    tri = vertices[faces]
    u, v, w = tri[:, 1] - tri[:, 0], tri[:, 2] - tri[:, 0], tri[:, 2] - tri[:, 1]
    face_n = torch.linalg.cross(u, v, dim=1)
    face_n = face_n / face_n.norm(dim=1, keepdim=True)
    un, vn, wn = (x / x.norm(dim=1, keepdim=True) for x in (u, v, w))
    ang = torch.empty_like(tri[:, :, 0])
    ang[:, 0] = torch.acos(_dot(un, vn).clamp(-1, 1))
    ang[:, 1] = torch.acos(_dot(-un, wn).clamp(-1, 1))
    ang[:, 2] = torch.pi - ang[:, 0] - ang[:, 1]
    row = faces.reshape(-1)
    col = torch.arange(len(faces), device=faces.device).unsqueeze(1).expand(-1, 3).reshape(-1)
    M = torch.sparse_coo_tensor(torch.stack((row, col)), ang.detach().reshape(-1), (len(vertices), len(faces)))
    vert_n = torch.sparse.mm(M, face_n)
    return vert_n / vert_n.norm(dim=1, keepdim=True)


def plane_hit(out_ori, out_dir, mask, plane_point, plane_normal):
    """OPTIONAL background-plane step (not in the reference): x = o + s d with s = ((p0 - o).n)/(d.n);
    zeros where the path is invalid, the ray is parallel to the plane or the plane is behind it."""
    p0 = torch.as_tensor(plane_point, dtype=out_ori.dtype, device=out_ori.device)
    n = torch.as_tensor(plane_normal, dtype=out_ori.dtype, device=out_ori.device)
    dn = out_dir @ n
    ok = mask[:, 0] & (dn != 0)
    s = ((p0 - out_ori) @ n) / torch.where(ok, dn, torch.ones_like(dn))
    ok = ok & (s > 0)
    pts = torch.where(ok.unsqueeze(1), out_ori + s.unsqueeze(1) * out_dir, torch.zeros_like(out_ori))
    return pts, ok


def _surface(vertices, faces, o, d, tri_ids, int_ior, ext_ior, normals=None):
    """JIT_Dintersect + refract_ray for rays that hit (DiffRender.py:64-121, 503-535).
    -> (keep mask, new origin, new direction).  normals != None: the commented-out interpolation of :107-114."""
    tri = vertices[faces[tri_ids]]                      # [n,3,3] gather: the differentiable link to the vertices
    a0, a1, a2 = tri[:, 0], tri[:, 1], tri[:, 2]
    e1, e2 = a1 - a0, a2 - a0
    pvec = torch.linalg.cross(d, e2, dim=1)
    inv_det = 1.0 / _dot(e1, pvec)
    qvec = torch.linalg.cross(o - a0, e1, dim=1)
    t = _dot(e2, qvec) * inv_det
    n = torch.linalg.cross(e1, e2, dim=1)
    n = n / n.norm(dim=1, keepdim=True)                 # flat face normal (:103-104)
    if normals is not None:                             # :107-114 as written there (u, v detached)
        u = (_dot(o - a0, pvec) * inv_det).detach()
        v = (_dot(d, qvec) * inv_det).detach()
        nn = normals[faces[tri_ids]]
        n = (1 - u - v).reshape((-1, 1)) * nn[:, 0] + u.reshape((-1, 1)) * nn[:, 1] + v.reshape((-1, 1)) * nn[:, 2]
        n = n / n.norm(p=2, dim=1, keepdim=True)
    wo = -d
    cos_i = _dot(wo, n).clamp(-1, 1)
    entering = cos_i > 0
    eta_i = torch.where(entering, torch.full_like(t, ext_ior), torch.full_like(t, int_ior))
    eta_t = torch.where(entering, torch.full_like(t, int_ior), torch.full_like(t, ext_ior))
    n = torch.where(entering.unsqueeze(1), n, -n)
    cos_i = torch.where(entering, cos_i, -cos_i)
    # FrDielectric (:51-61): only the total-internal-reflection flag is used (:526)
    sin_t = torch.sqrt((1 - cos_i * cos_i).clamp(0, 1)) * eta_i / eta_t
    keep = ~(sin_t >= 1)
    # Refract (:35-49): cosThetaT from sin2ThetaI (sic), then renormalise
    eta = (eta_i / eta_t).unsqueeze(1)
    c = _dot(n, wo).unsqueeze(1)
    s2 = (1 - c * c).clamp(min=0)
    c_t = torch.sqrt(1 - s2.clamp(max=1))
    w = eta * -wo + (eta * c - c_t) * n
    w = w / w.norm(dim=1, keepdim=True)
    new_o = o + t.unsqueeze(1) * d + 1e-5 * w            # :528-532
    return keep, new_o, w


def render_transparent(vertices, faces, origin, ray_dir, intersect, int_ior, ext_ior=EXT_IOR, normals=None):
    """(out_ori, out_dir, mask[N,3]) with autograd history back to `vertices` (and `normals` in the optional smooth mode)."""
    n = origin.shape[0]
    idx = torch.arange(n, device=origin.device)
    o, d = origin, ray_dir
    for _ in range(2):                                   # trace2 (:537-546)
        ids, hit = _query(intersect, o, d)
        idx, o, d, ids = idx[hit], o[hit], d[hit], ids[hit]
        keep, o, d = _surface(vertices, faces, o, d, ids, int_ior, ext_ior, normals)
        idx, o, d = idx[keep], o[keep], d[keep]
    _, hit = _query(intersect, o, d)                     # third query: any further surface rejects the path (:425-427)
    idx, o, d = idx[~hit], o[~hit], d[~hit]
    out_ori = torch.zeros_like(origin).index_put((idx,), o)
    out_dir = torch.zeros_like(origin).index_put((idx,), d)
    mask = torch.zeros(origin.shape, dtype=torch.bool, device=origin.device)
    mask[idx] = True
    return out_ori, out_dir, mask
