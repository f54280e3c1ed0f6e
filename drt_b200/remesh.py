"""Stand-in for the reference's MeshLab bridge (optim.py:12-56: `Meshlabserver.remesh` = "Remeshing: Isotropic Explicit
Remeshing", 3 iterations, TargetLen = remesh_len, executed by the external `meshlabserver` binary between optimisation passes).

MeshLab is not part of this image, so this is NOT that filter; it is the part of it the coarse-to-fine schedule of optim.py needs
(`remesh_len` shrinks from start_len to end_len over the passes, optim.py:192): every edge longer than 4/3 of the target length is
split at its midpoint, faces are re-triangulated conformingly (1, 2 or 3 split edges per face), and the vertices are relaxed
tangentially.  Watertightness and orientation are preserved by construction (the reference asserts `mesh.is_watertight` after
every reload, DiffRender.py:305).  Edges shorter than the target are left alone (no collapse step): a mesh is only ever refined.

    remesh(vertices, faces, target_len, iterations=3) -> (vertices, faces)
    Remesher().remesh(scene, remesh_len)   # same call as optim.py:198 (`meshlabserver.remesh(scene, remesh_len)`)
"""
import os
import tempfile

import numpy as np


def _edge_table(faces):
    """-> (uniq [E,2] sorted vertex pairs, fe [F,3] edge index of the edges (v0v1, v1v2, v2v0) of every face)"""
    e = np.concatenate([faces[:, [0, 1]], faces[:, [1, 2]], faces[:, [2, 0]]], axis=0)
    uniq, inv = np.unique(np.sort(e, axis=1), axis=0, return_inverse=True)
    return uniq, inv.reshape(3, -1).T


def split_long_edges(vertices, faces, max_len):
    """One refinement sweep: split every edge longer than `max_len` at its midpoint and re-triangulate each face by the number
    of its split edges (conforming: both faces on an edge see the same midpoint).  -> (vertices, faces, n_split)"""
    v = np.asarray(vertices, dtype=np.float64)
    f = np.asarray(faces, dtype=np.int64)
    uniq, fe = _edge_table(f)
    length = np.linalg.norm(v[uniq[:, 0]] - v[uniq[:, 1]], axis=1)
    split = length > max_len
    n_split = int(split.sum())
    if n_split == 0:
        return v, f, 0
    mid_id = np.full(len(uniq), -1, dtype=np.int64)
    mid_id[split] = len(v) + np.arange(n_split)
    v = np.concatenate([v, 0.5 * (v[uniq[split, 0]] + v[uniq[split, 1]])], axis=0)
    m = mid_id[fe]                       # [F,3] midpoint vertex of edge k of the face, or -1
    k = (m >= 0).sum(axis=1)
    out = [f[k == 0]]
    # one split edge: rotate the face so that the split edge is (a, b); -> (a, m, c), (m, b, c)
    for r in range(3):
        sel = (k == 1) & (m[:, r] >= 0)
        a, b, c, mm = f[sel, r], f[sel, (r + 1) % 3], f[sel, (r + 2) % 3], m[sel, r]
        out += [np.stack([a, mm, c], 1), np.stack([mm, b, c], 1)]
    # two split edges: rotate so that the UNSPLIT edge is (c, a), i.e. edges (a,b) and (b,c) are split;
    # -> (m_ab, b, m_bc) and the quad (a, m_ab, m_bc, c) cut along its shorter diagonal
    for r in range(3):
        sel = (k == 2) & (m[:, (r + 2) % 3] < 0)
        a, b, c = f[sel, r], f[sel, (r + 1) % 3], f[sel, (r + 2) % 3]
        mab, mbc = m[sel, r], m[sel, (r + 1) % 3]
        out.append(np.stack([mab, b, mbc], 1))
        d1 = np.linalg.norm(v[a] - v[mbc], axis=1) <= np.linalg.norm(v[mab] - v[c], axis=1)
        out += [np.where(d1[:, None], np.stack([a, mab, mbc], 1), np.stack([a, mab, c], 1)),
                np.where(d1[:, None], np.stack([a, mbc, c], 1), np.stack([mab, mbc, c], 1))]
    # three split edges: the regular 1 -> 4 split
    sel = k == 3
    a, b, c = f[sel, 0], f[sel, 1], f[sel, 2]
    mab, mbc, mca = m[sel, 0], m[sel, 1], m[sel, 2]
    out += [np.stack([a, mab, mca], 1), np.stack([mab, b, mbc], 1), np.stack([mca, mbc, c], 1), np.stack([mab, mbc, mca], 1)]
    return v, np.concatenate(out, axis=0), n_split


def tangential_smooth(vertices, faces, strength=0.5):
    """Move every vertex towards the centroid of its neighbours, within its tangent plane (MeshLab's "Smooth Step")."""
    v = np.asarray(vertices, dtype=np.float64)
    f = np.asarray(faces, dtype=np.int64)
    uniq, _ = _edge_table(f)
    acc = np.zeros_like(v)
    cnt = np.zeros(len(v))
    np.add.at(acc, uniq[:, 0], v[uniq[:, 1]])
    np.add.at(acc, uniq[:, 1], v[uniq[:, 0]])
    np.add.at(cnt, uniq[:, 0], 1.0)
    np.add.at(cnt, uniq[:, 1], 1.0)
    fn = np.cross(v[f[:, 1]] - v[f[:, 0]], v[f[:, 2]] - v[f[:, 0]])          # area-weighted face normals
    vn = np.zeros_like(v)
    for c in range(3):
        np.add.at(vn, f[:, c], fn)
    vn /= np.maximum(np.linalg.norm(vn, axis=1, keepdims=True), 1e-300)
    delta = acc / np.maximum(cnt, 1.0)[:, None] - v
    delta -= vn * (delta * vn).sum(axis=1, keepdims=True)
    return v + strength * delta


def remesh(vertices, faces, target_len, iterations=3, smooth=0.5):
    """Refine until no edge is longer than 4/3 * target_len (at most `iterations` sweeps x 3 split rounds, like the filter's
    "Iterations" parameter), relaxing tangentially after each sweep."""
    v, f = np.asarray(vertices, dtype=np.float64), np.asarray(faces, dtype=np.int64)
    for _ in range(iterations):
        total = 0
        for _ in range(3):
            v, f, n = split_long_edges(v, f, 4.0 / 3.0 * target_len)
            total += n
            if n == 0:
                break
        if total == 0:
            break
        if smooth > 0:
            v = tangential_smooth(v, f, smooth)
    return v, f


class Remesher:
    """Same call shape as optim.py:12-56 `Meshlabserver`: `remesh(scene, remesh_len)` exports the scene's CURRENT mesh, remeshes
    it and reloads it through `scene.update_mesh(path)` (optim.py:50-52), so the Scene goes through the same code path."""

    def __init__(self, tmp_path=None):
        self.tmp_path = tmp_path or tempfile.gettempdir()
        pid = os.getpid()
        self.ply_path = os.path.join(self.tmp_path, f"temp_{pid}.ply")
        self.remeshply_path = os.path.join(self.tmp_path, f"remesh_{pid}.ply")

    def remesh(self, scene, remesh_len):
        from . import plyio
        scene.mesh.export(self.ply_path)                               # optim.py:50
        v, f = plyio.read_ply(self.ply_path)
        v, f = remesh(v, f, remesh_len)                                 # optim.py:51 (the external meshlabserver call)
        plyio.write_ply(self.remeshply_path, v, f)
        scene.update_mesh(self.remeshply_path)                          # optim.py:52
