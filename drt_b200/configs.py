"""The benchmark / parity configurations of BASELINE.json as concrete synthetic inputs
(BASELINE.md "Configs").  Meshes of C2-C4 are the reference's data/*.ply
converted to npz fixtures by oracle/make_golden.py (tests/golden/meshes)."""
import os

import numpy as np

from . import meshgen, views

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
MESH_DIR = os.path.join(_ROOT, "tests", "golden", "meshes")
INT_IOR = 1.4723   # config.py:22
EXT_IOR = 1.00029  # DiffRender.py:21


def load_mesh(name):
    z = np.load(os.path.join(MESH_DIR, name + ".npz"))
    return z["vertices"].astype(np.float64), z["faces"].astype(np.int64)


CONFIGS = {
    # name: (mesh builder, resy, resx, n_views, description)
    "C1": dict(mesh=lambda: meshgen.icosahedron(), resy=64, resx=64, n_views=1, view0=5,
               desc="icosahedron (20 tris), 1 view 64x64"),
    "C2": dict(mesh=lambda: load_mesh("hand_vh"), resy=512, resx=512, n_views=1, view0=5,
               desc="hand_vh (4390 tris), 1 view 512x512"),
    "C3": dict(mesh=lambda: load_mesh("mouse_vh"), resy=720, resx=960, n_views=72, view0=0,
               desc="mouse_vh (9246 tris), 72 views 960x720"),
    "C4": dict(mesh=lambda: meshgen.subdivide(*load_mesh("horse_vh"), jitter=0.05, seed=0), resy=720, resx=960,
               n_views=72, view0=0, desc="horse_vh 1->4 subdivided (50248 tris), 72 views 960x720"),
    "C5": dict(mesh=lambda: meshgen.displaced_torus(), resy=1080, resx=1920, n_views=256, view0=0,
               desc="displaced torus (200000 tris), 256 views 1920x1080"),
}


def make(name):
    """-> dict(vertices, faces, cams, resy, resx, n_views, desc).  cams = the turntable of 72 (C5: 256)
    views; a config with n_views=1 uses cams[view0]."""
    c = CONFIGS[name]
    v, f = c["mesh"]()
    total = 256 if name == "C5" else 72
    cams = views.turntable_cameras(v, c["resy"], c["resx"], total)
    if c["n_views"] == 1:
        cams = [cams[c["view0"]]]
    return dict(name=name, vertices=v, faces=f, cams=cams, resy=c["resy"], resx=c["resx"], n_views=len(cams),
                desc=c["desc"])


def perturbed_target_mesh(vertices, seed=1, scale=0.3):
    """Seeded smooth-ish perturbation standing in for the unknown true shape: the synthetic
    `screen_pixel` targets of ray_loss (optim.py:96-106) come from tracing this mesh."""
    rng = np.random.default_rng(seed)
    v = np.asarray(vertices, dtype=np.float64)
    ctr = v.mean(0)
    k = rng.normal(size=(3, 3)) * 0.02
    return v + scale * np.sin((v - ctr) @ k)
