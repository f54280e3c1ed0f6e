"""The remesh stand-in for the reference's MeshLab bridge (optim.py:12-56): refinement to a target edge length keeps the
mesh watertight and consistently oriented (DiffRender.py:305 asserts watertightness after every reload)."""
import numpy as np

from conftest import load_mesh
from drt_b200 import meshgen, plyio, remesh


def _signed_volume(v, f):
    a, b, c = v[f[:, 0]], v[f[:, 1]], v[f[:, 2]]
    return float((a * np.cross(b, c)).sum() / 6.0)


def _edge_lengths(v, f):
    e = np.concatenate([f[:, [0, 1]], f[:, [1, 2]], f[:, [2, 0]]], axis=0)
    return np.linalg.norm(v[e[:, 0]] - v[e[:, 1]], axis=1)


def test_split_patterns_keep_the_mesh_closed_and_oriented():
    v, f = meshgen.icosahedron()
    rng = np.random.default_rng(0)
    v = v + 0.15 * rng.standard_normal(v.shape)      # unequal edges, so that a threshold splits SOME edges of a face
    vol = _signed_volume(v, f)
    for _ in range(6):                       # random thresholds hit faces with 1, 2 and 3 split edges
        lim = 0.999 * np.quantile(_edge_lengths(v, f), rng.uniform(0.2, 0.9))
        v, f, n = remesh.split_long_edges(v, f, lim)
        assert n > 0 and meshgen.is_watertight(f)
        assert abs(_signed_volume(v, f) - vol) < 1e-9 * abs(vol)      # midpoint splits do not move the surface
        assert f.min() == 0 and f.max() == len(v) - 1


def test_remesh_reaches_the_target_length_on_a_reference_mesh():
    v, f = load_mesh("hand_vh")
    vol = _signed_volume(v, f)
    target = 2.0                                                       # mean edge of hand_vh is 3.32
    v2, f2 = remesh.remesh(v, f, target, iterations=3)
    assert meshgen.is_watertight(f2) and len(f2) > 2 * len(f)
    assert _edge_lengths(v2, f2).max() <= 4.0 / 3.0 * target * 1.05     # tangential relaxation may stretch an edge slightly
    assert abs(_signed_volume(v2, f2) - vol) < 0.02 * abs(vol) and _signed_volume(v2, f2) * vol > 0
    v3, f3 = remesh.remesh(v2, f2, target, iterations=3, smooth=0.0)
    assert len(f3) <= 1.05 * len(f2)                                    # already at the target: (almost) nothing left to split


class _Scene:
    def __init__(self, v, f):
        from drt_b200 import trimesh_lite
        self.mesh = trimesh_lite.TriMesh(v, f)
        self.loaded = None

    def update_mesh(self, path):
        self.loaded = plyio.read_ply(path)


def test_remesher_has_the_call_shape_of_meshlabserver(tmp_path):
    v, f = meshgen.icosahedron()
    sc = _Scene(v, f)
    remesh.Remesher(str(tmp_path)).remesh(sc, remesh_len=0.8)            # optim.py:198
    v2, f2 = sc.loaded
    assert meshgen.is_watertight(f2) and len(f2) > len(f)
    assert _edge_lengths(v2, f2).max() <= 4.0 / 3.0 * 0.8 * 1.1
