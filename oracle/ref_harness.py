"""Runs the UNMODIFIED reference (/root/reference/DiffRender.py) on CPU -- build-container only.

TEST INFRASTRUCTURE ONLY.  /root/reference does not exist on the GPU box, so nothing that runs
there imports this module; it is used by oracle/make_golden.py to produce tests/golden/ and by
the `-m "not gpu"` test that re-validates the oracle when the reference tree is present.

Recipe: stub the modules the reference imports but does not need on this path
(`config`, `trimesh`, `imageio`), intercept torch.utils.cpp_extension.load (DiffRender.py:5-6) so
that `optix.optix_mesh` is a stand-in with the four methods of optix_extend.cpp:77-83, and
construct Scene without its trimesh-based constructor (DiffRender.py:299-317).
The stand-in's intersect is the oracle's brute-force closest hit (oracle/drt_oracle.c), or any
callable given as `intersect_fn` (Tier-B parity: feed the ids of the intersector under test).
"""
import os
import sys
import types
import warnings

import numpy as np
import torch

REF_ROOT = os.environ.get("DRT_REFERENCE_ROOT", "/root/reference")


def available():
    return os.path.isfile(os.path.join(REF_ROOT, "DiffRender.py"))


class StandInOptixMesh:
    """Same surface as optix_mesh (optix_extend.cpp:6-83)."""
    intersect_fn = None  # optional override: f(ray6 np.float32[N,6]) -> (T f32[N], ID i32[N])
    calls = None         # when a list, every (T, ID) is appended

    def __init__(self, cuda_device=0):
        self.F = self.V = self.mesh = None

    def update_mesh(self, F, V):
        assert F.shape[1] == 3 and V.shape[1] == 3
        self.F, self.V = F, V
        self._rebuild()

    def update_vert(self, V):
        assert V.shape[1] == 3
        self.V = V
        self._rebuild()

    def _rebuild(self):
        from . import oracle
        self.mesh = oracle.OracleMesh(self.V.detach().cpu().numpy().astype(np.float64), self.F.cpu().numpy())

    def intersect(self, Ray):
        assert Ray.shape[1] == 6 and Ray.dtype == torch.float32
        r = Ray.detach().cpu().numpy()
        if StandInOptixMesh.intersect_fn is not None:
            T, ID = StandInOptixMesh.intersect_fn(r)
        else:
            T, ID = self.mesh.closest_hit(r, use_bvh=False)
        if StandInOptixMesh.calls is not None:
            StandInOptixMesh.calls.append((T.copy(), ID.copy()))
        return [torch.from_numpy(np.asarray(T, np.float32)), torch.from_numpy(np.asarray(ID, np.int32))]


_R = None


def load_reference():
    """-> the reference's DiffRender module object, imported unmodified."""
    global _R
    if _R is not None:
        return _R
    if not available():
        raise RuntimeError(f"reference tree not found at {REF_ROOT}")
    import torch.utils.cpp_extension as ce
    saved = ce.load
    ce.load = lambda **kw: types.SimpleNamespace(optix_mesh=StandInOptixMesh)
    cfg = types.ModuleType("config")
    cfg.optix_include = cfg.optix_ld = ""
    stubs = {"config": cfg, "trimesh": types.ModuleType("trimesh"), "imageio": types.ModuleType("imageio")}
    prev = {k: sys.modules.get(k) for k in stubs}
    sys.modules.update(stubs)
    sys.path.insert(0, REF_ROOT)
    try:
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            import DiffRender as R  # noqa: N811
    finally:
        sys.path.remove(REF_ROOT)
        ce.load = saved
        for k, v in prev.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    R.device = "cpu"
    _R = R
    return R


def make_scene(vertices, faces, int_ior, requires_grad=True):
    """Reference Scene over (vertices, faces) without trimesh."""
    R = load_reference()
    R.intIOR = float(int_ior)  # optim.py:178
    s = R.Scene.__new__(R.Scene)
    s.optix_mesh = StandInOptixMesh(0)
    s.faces = torch.as_tensor(np.asarray(faces), dtype=torch.long)
    s.vertices = torch.tensor(np.asarray(vertices), dtype=torch.float64, requires_grad=requires_grad)
    s.triangles = s.vertices[s.faces]
    s.normals = torch.zeros_like(s.vertices)  # dead on this path (DiffRender.py:65, 497)
    s.optix_mesh.update_mesh(s.faces.int(), s.vertices.detach().float())  # DiffRender.py:311-313
    return s


def render_transparent(vertices, faces, origin, ray_dir, int_ior, g_ori=None, g_dir=None):
    """Reference forward (+ backward of sum(out_ori*g_ori + out_dir*g_dir) when g_dir is given).
    -> dict(out_ori, out_dir, mask[N] bool, grad_V | None)"""
    s = make_scene(vertices, faces, int_ior)
    o = torch.as_tensor(np.asarray(origin), dtype=torch.float64)
    d = torch.as_tensor(np.asarray(ray_dir), dtype=torch.float64)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        out_ori, out_dir, mask = s.render_transparent(o, d)
        gV = None
        if g_dir is not None:
            L = (out_dir * torch.as_tensor(np.asarray(g_dir))).sum()
            if g_ori is not None:
                L = L + (out_ori * torch.as_tensor(np.asarray(g_ori))).sum()
            L.backward()
            gV = s.vertices.grad.detach().numpy().copy()
    return dict(out_ori=out_ori.detach().numpy(), out_dir=out_dir.detach().numpy(),
                mask=mask[:, 0].numpy().copy(), grad_V=gV)
