"""Robustness of the query stage: the conservative FMA slab test + exact float64 triangle test must
reproduce the brute-force oracle bit for bit on hostile inputs -- triangle soups with degenerate and
duplicated triangles, axis-parallel rays lying in face planes, tiny / huge coordinate scales and ray
origins far outside the range the build-time box inflation covers (the additive-slack path)."""
import numpy as np
import pytest
import torch

from conftest import load_mesh
from oracle import oracle

pytestmark = pytest.mark.gpu


def _gpu_hits(dev, v, f, ray):
    from drt_b200 import optix
    om = optix.optix_mesh(dev.index or 0)
    om.update_mesh(torch.tensor(f, dtype=torch.int32, device=dev), torch.tensor(v, dtype=torch.float32, device=dev))
    T, I = om.intersect(torch.tensor(ray, dtype=torch.float32, device=dev))
    return T.cpu().numpy(), I.cpu().numpy()


def _check(dev, v, f, ray, min_hit_frac=0.0):
    v = np.asarray(v, np.float32).astype(np.float64)  # what both sides see
    ray = np.asarray(ray, np.float32)
    T0, I0 = oracle.OracleMesh(v, f).closest_hit(ray, use_bvh=False)
    T, I = _gpu_hits(dev, v, f, ray)
    assert np.array_equal(I, I0), f"{(I != I0).sum()} of {len(I)} ids differ"
    assert np.array_equal(T, T0)
    assert (I0 >= 0).mean() >= min_hit_frac
    return I0


def test_triangle_soup_with_degenerate_and_duplicate_triangles(cuda_device):
    rng = np.random.default_rng(0)
    v = rng.uniform(-1, 1, size=(300, 3))
    f = rng.integers(0, 300, size=(800, 3))
    f[:40, 2] = f[:40, 1]                     # zero-area: two equal indices
    f[40:60] = f[60:80]                       # exact duplicates: ties must resolve to the lowest id
    v[f[100, 2]] = 0.5 * (v[f[100, 0]] + v[f[100, 1]])   # collinear vertices
    o = rng.uniform(-3, 3, size=(20000, 3))
    d = rng.normal(size=(20000, 3))
    I0 = _check(cuda_device, v, f, np.concatenate([o, d], 1), 0.05)
    assert not np.isin(I0, np.arange(60, 80)).any()       # the higher-numbered copy of a duplicate never wins


def test_axis_parallel_rays_in_face_planes(cuda_device):
    """A regular grid mesh in the plane z=0 and boxes with coordinates on exact float values: rays with zero
    direction components, origins exactly on box planes, rays grazing shared edges and vertices."""
    n = 16
    xs = np.arange(n + 1, dtype=np.float64)
    X, Y = np.meshgrid(xs, xs, indexing="ij")
    v = np.stack([X.ravel(), Y.ravel(), np.zeros(X.size)], 1)
    idx = lambda i, j: i * (n + 1) + j  # noqa: E731
    f = []
    for i in range(n):
        for j in range(n):
            f += [[idx(i, j), idx(i + 1, j), idx(i + 1, j + 1)], [idx(i, j), idx(i + 1, j + 1), idx(i, j + 1)]]
    f = np.array(f)
    rays = []
    for i in range(n + 1):
        for j in range(n + 1):
            rays.append([i, j, 5, 0, 0, -1])            # straight down onto every vertex (up to 6 triangles tie)
            rays.append([i + 0.5, j, 3, 0, 0, -2])      # onto edge midpoints, un-normalised
            rays.append([i, j + 0.25, -4, 0, 0, 1])     # from below
    for k in range(n + 1):
        rays.append([-3, k, 0, 1, 0, 0])                # IN the plane z=0 along a grid line: coplanar, det == 0
        rays.append([k, -3, 0, 0, 1, 0])
        rays.append([-3, k + 0.5, 1e-3, 1, 0, -1e-4])   # grazing
    I0 = _check(cuda_device, v, f, np.array(rays, np.float64), 0.3)
    assert (I0[0: 3 * (n + 1) ** 2: 3] >= 0).all()   # every ray aimed straight at a grid vertex hits


@pytest.mark.parametrize("scale", [1e-4, 1.0, 1e4])
def test_coordinate_scale_invariance(cuda_device, scale):
    v, f = load_mesh("hand_vh")
    rng = np.random.default_rng(2)
    ctr = 0.5 * (v.min(0) + v.max(0))
    o = ctr + rng.normal(size=(8000, 3)) * 150
    d = (v[rng.integers(0, len(v), 8000)] + rng.normal(size=(8000, 3))) - o
    ray = np.concatenate([o * scale, d * scale], 1)
    I0 = _check(cuda_device, v * scale, f, ray, 0.3)
    if scale != 1.0:  # power-of-ten scaling is not exact in binary, so only the hit statistics are comparable
        assert abs((I0 >= 0).mean() - (_check(cuda_device, v, f, np.concatenate([o, d], 1)) >= 0).mean()) < 0.01


def test_far_origins_take_the_additive_slack_path(cuda_device):
    """|origin| > 64 * max|coordinate|: the build-time inflation no longer covers the FMA rounding and the per-ray
    slack E must keep the box test conservative -- including near-axis-parallel directions."""
    v, f = load_mesh("hand_vh")
    rng = np.random.default_rng(3)
    pmax = np.abs(v).max()
    tgt = v[rng.integers(0, len(v), 6000)] + rng.normal(size=(6000, 3)) * 0.5
    dirs = rng.normal(size=(6000, 3))
    dirs[:2000, 1] *= 1e-6                                # nearly perpendicular to y
    dirs[2000:3000, 0] = 0.0                              # exactly axis-parallel in x
    dirs /= np.linalg.norm(dirs, axis=1, keepdims=True)
    dist = rng.uniform(100, 5000, size=(6000, 1)) * pmax  # 1e2 .. 5e3 scene sizes away
    o = tgt - dirs * dist
    ray = np.concatenate([o, dirs], 1)
    assert (np.abs(o).max(1) > 64 * pmax).all()
    _check(cuda_device, v, f, ray, 0.1)


def test_single_and_two_triangle_meshes_and_empty_slots(cuda_device):
    v = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0], [1, 1, 0.5]], np.float64)
    rng = np.random.default_rng(4)
    o = rng.uniform(-1, 2, size=(3000, 3)) + [0, 0, 2]
    d = rng.normal(size=(3000, 3)) * 0.3 + [0, 0, -1]
    ray = np.concatenate([o, d], 1)
    _check(cuda_device, v, np.array([[0, 1, 2]]), ray, 0.02)
    _check(cuda_device, v, np.array([[0, 1, 2], [1, 3, 2]]), ray, 0.05)
    _check(cuda_device, v, np.array([[0, 1, 2], [1, 3, 2], [0, 1, 3]]), ray, 0.05)


def test_cooperative_build_and_no_beam_give_the_same_answers(cuda_device):
    """The defaults are the 19-launch LBVH build and the beam-culled entry query; the parity suite must also hold with the
    single cooperative build kernel and without beam culling -- run in a subprocess, the switches are read once."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, DRT_COOP_BUILD="1", DRT_BEAM="0")
    r = subprocess.run([sys.executable, "-m", "pytest", "-q", "-x", "-m", "gpu", "-p", "no:cacheprovider",
                        os.path.join(root, "tests", "test_gpu_parity.py"), os.path.join(root, "tests", "test_gpu_loss_step.py"),
                        "-k", "closest_hit or refit or loss_step_vs_oracle or edge_cases"],
                       cwd=root, env=env, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    # the fused step in 3 forced lanes (internal streams, a lane boundary inside an image) on a whole benchmark view
    env = dict(os.environ, DRT_LANES="3")
    r = subprocess.run([sys.executable, "-m", "pytest", "-q", "-x", "-m", "gpu", "-p", "no:cacheprovider",
                        os.path.join(root, "tests", "test_gpu_headline_parity.py"), "-k", "full_c4_view"],
                       cwd=root, env=env, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
