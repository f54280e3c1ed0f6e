"""The handful of trimesh features Scene needs (DiffRender.py:303-309, 338-355, 357-376;
optim.py:50,226), on numpy only: load/export PLY, watertightness, edge tables, vertex neighbours."""
import numpy as np

from . import plyio
from .meshgen import is_watertight


class TriMesh:
    def __init__(self, vertices, faces):
        self.vertices = np.asarray(vertices, dtype=np.float64)
        self.faces = np.asarray(faces, dtype=np.int64)
        self._cache = {}

    @property
    def is_watertight(self):
        return is_watertight(self.faces)

    @property
    def edges(self):
        """Directed edges, face-major: (f0:v0v1, f0:v1v2, f0:v2v0, f1:...), like trimesh."""
        f = self.faces
        return f[:, [0, 1, 1, 2, 2, 0]].reshape(-1, 2)

    @property
    def edges_sorted(self):
        return np.sort(self.edges, axis=1)

    @property
    def edges_face(self):
        return np.repeat(np.arange(len(self.faces)), 3)

    @property
    def vertex_neighbors(self):
        if "vn" not in self._cache:
            e = np.unique(self.edges_sorted, axis=0)
            both = np.concatenate([e, e[:, ::-1]], axis=0)
            order = np.argsort(both[:, 0], kind="stable")
            both = both[order]
            splits = np.searchsorted(both[:, 0], np.arange(1, len(self.vertices)))
            self._cache["vn"] = [a.tolist() for a in np.split(both[:, 1], splits)]
        return self._cache["vn"]

    def export(self, path):
        plyio.write_ply(path, self.vertices, self.faces)
        return path


def load(path, process=False):
    v, f = plyio.read_ply(path)
    return TriMesh(v, f)


def group_rows_pairs(edges_sorted):
    """trimesh.grouping.group_rows(edges_sorted, require_count=2): indices [E,2] of the two directed
    edges that share each undirected edge (DiffRender.py:349)."""
    e = np.asarray(edges_sorted)
    key = e[:, 0].astype(np.int64) * (int(e.max()) + 1) + e[:, 1]
    order = np.argsort(key, kind="stable")
    ks = key[order]
    start = np.nonzero(np.concatenate([[True], ks[1:] != ks[:-1]]))[0]
    count = np.diff(np.concatenate([start, [len(ks)]]))
    sel = start[count == 2]
    return np.stack([order[sel], order[sel + 1]], axis=1)
