"""Synthetic meshes for the benchmark configs (BASELINE.md "Configs").

None of this exists in the reference (its meshes come from captured visual hulls, data/*.ply);
the generators only provide watertight inputs of the named sizes.
"""
import numpy as np


def icosahedron(radius=1.9):
    """C1: regular icosahedron, 12 vertices / 20 faces, outward winding."""
    p = (1.0 + 5.0 ** 0.5) / 2.0
    v = np.array(
        [[-1, p, 0], [1, p, 0], [-1, -p, 0], [1, -p, 0], [0, -1, p], [0, 1, p], [0, -1, -p], [0, 1, -p],
         [p, 0, -1], [p, 0, 1], [-p, 0, -1], [-p, 0, 1]], dtype=np.float64)
    v *= radius / np.linalg.norm(v[0])
    f = np.array(
        [[0, 11, 5], [0, 5, 1], [0, 1, 7], [0, 7, 10], [0, 10, 11], [1, 5, 9], [5, 11, 4], [11, 10, 2],
         [10, 7, 6], [7, 1, 8], [3, 9, 4], [3, 4, 2], [3, 2, 6], [3, 6, 8], [3, 8, 9], [4, 9, 5],
         [2, 4, 11], [6, 2, 10], [8, 6, 7], [9, 8, 1]], dtype=np.int64)
    return v, f


def tetrahedron():
    """The 4-face solid of the reference-generated known-answer vector."""
    v = np.array([[0, 0, 0], [4, 0, 0], [0, 4, 0], [0, 0, 4]], dtype=np.float64)
    f = np.array([[0, 2, 1], [0, 1, 3], [0, 3, 2], [1, 2, 3]], dtype=np.int64)
    return v, f


def subdivide(vertices, faces, jitter=0.0, seed=0):
    """One 1->4 midpoint subdivision (C4: horse_vh 12 562 -> 50 248 triangles).  Optional seeded
    jitter of the new vertices along the face normal keeps the four children from being coplanar."""
    v = np.asarray(vertices, dtype=np.float64)
    f = np.asarray(faces, dtype=np.int64)
    e = np.concatenate([f[:, [0, 1]], f[:, [1, 2]], f[:, [2, 0]]], axis=0)
    es = np.sort(e, axis=1)
    uniq, inv = np.unique(es, axis=0, return_inverse=True)
    inv = inv.reshape(-1)
    mid = 0.5 * (v[uniq[:, 0]] + v[uniq[:, 1]])
    if jitter > 0:
        rng = np.random.default_rng(seed)
        fn = np.cross(v[f[:, 1]] - v[f[:, 0]], v[f[:, 2]] - v[f[:, 0]])
        fn /= np.linalg.norm(fn, axis=1, keepdims=True)
        en = np.zeros_like(mid)
        np.add.at(en, inv, np.concatenate([fn, fn, fn], axis=0))
        en /= np.maximum(np.linalg.norm(en, axis=1, keepdims=True), 1e-30)
        mid = mid + en * rng.uniform(-jitter, jitter, size=(len(mid), 1))
    nf = len(f)
    m01, m12, m20 = (inv[:nf] + len(v)), (inv[nf:2 * nf] + len(v)), (inv[2 * nf:] + len(v))
    nfaces = np.concatenate([
        np.stack([f[:, 0], m01, m20], axis=1), np.stack([m01, f[:, 1], m12], axis=1),
        np.stack([m20, m12, f[:, 2]], axis=1), np.stack([m01, m12, m20], axis=1)], axis=0)
    return np.concatenate([v, mid], axis=0), nfaces


def displaced_torus(nu=400, nv=250, R=60.0, r=25.0, sigma=0.5, seed=0):
    """C5: torus grid nu x nv quads -> 2*nu*nv triangles (default 200 000), radial noise sigma mm."""
    rng = np.random.default_rng(seed)
    u = np.arange(nu) * (2 * np.pi / nu)
    w = np.arange(nv) * (2 * np.pi / nv)
    uu, ww = np.meshgrid(u, w, indexing="ij")
    rr = r + sigma * rng.standard_normal(uu.shape)
    x = (R + rr * np.cos(ww)) * np.cos(uu)
    z = (R + rr * np.cos(ww)) * np.sin(uu)
    y = rr * np.sin(ww) + r + 2.0  # sits on y ~ 0 like the turntable objects
    v = np.stack([x, y, z], axis=-1).reshape(-1, 3)
    i = np.arange(nu)[:, None]
    j = np.arange(nv)[None, :]
    a = (i * nv + j).reshape(-1)
    b = (((i + 1) % nu) * nv + j).reshape(-1)
    c = (((i + 1) % nu) * nv + (j + 1) % nv).reshape(-1)
    d = (i * nv + (j + 1) % nv).reshape(-1)
    f = np.concatenate([np.stack([a, d, c], axis=1), np.stack([a, c, b], axis=1)], axis=0)
    return v.astype(np.float64), f.astype(np.int64)


def is_watertight(faces):
    """Every undirected edge shared by exactly two faces with opposite orientation
    (the property DiffRender.py:305 asserts through trimesh)."""
    f = np.asarray(faces, dtype=np.int64)
    e = np.concatenate([f[:, [0, 1]], f[:, [1, 2]], f[:, [2, 0]]], axis=0)
    key = np.sort(e, axis=1)
    _, inv, cnt = np.unique(key, axis=0, return_inverse=True, return_counts=True)
    if not (cnt == 2).all():
        return False
    sign = np.where(e[:, 0] < e[:, 1], 1, -1)
    tot = np.zeros(len(cnt), dtype=np.int64)
    np.add.at(tot, inv.reshape(-1), sign)
    return bool((tot == 0).all())
