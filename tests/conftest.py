import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLD = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


def load_mesh(name):
    """Mesh fixtures converted from the reference's data/<name>.ply by oracle/make_golden.py."""
    z = np.load(os.path.join(GOLD, "meshes", name + ".npz"))
    return z["vertices"].astype(np.float64), z["faces"].astype(np.int64)


def load_chain_case(case):
    """-> dict with vertices, faces, origin, ray_dir, g_ori, g_dir, int_ior, valid_idx, out_ori, out_dir,
    grad_V, grad_V_dir_only (reference-generated; rays/upstream gradients are regenerated and
    checksummed)."""
    from drt_b200 import views
    z = dict(np.load(os.path.join(GOLD, f"chain_{case}.npz")))
    if "origin" not in z:
        resy, resx = (int(x) for x in z["res"])
        o, d = views.generate_ray(resy, resx, z["cam_K_inv"], z["cam_R_inv"])
        z["origin"], z["ray_dir"] = o.numpy(), d.numpy()
    rng = np.random.default_rng(int(z["seed"]))
    z["g_ori"] = rng.standard_normal(z["origin"].shape)
    z["g_dir"] = rng.standard_normal(z["origin"].shape)
    chk = np.array([z["origin"].sum(), z["ray_dir"].sum(), z["g_ori"].sum(), z["g_dir"].sum()])
    assert np.allclose(chk, z["ray_checksum"], rtol=1e-12, atol=1e-9), "golden inputs drifted"
    z["int_ior"] = float(z["int_ior"])
    return z


CHAIN_CASES = ["tetra", "icosa_c1", "hand_vh_160", "mouse_vh_96x128", "hand_vh_perturbed_96"]


def grad_rel_err(g, ref):
    """SURVEY.md 8(c) Tier B metric: worst per-vertex ||dg||/||ref|| over vertices with
    ||ref|| > 1e-9*max, and global max-abs / max-abs."""
    nr = np.linalg.norm(ref, axis=1)
    sel = nr > 1e-9 * nr.max()
    per_vertex = (np.linalg.norm(g - ref, axis=1)[sel] / nr[sel]).max() if sel.any() else 0.0
    glob = np.abs(g - ref).max() / max(np.abs(ref).max(), 1e-300)
    return per_vertex, glob


@pytest.fixture(scope="session")
def cuda_device():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")


@pytest.fixture(params=["staged", "direct"])
def fwd_route(request, cuda_device):
    """drt_ray_loss_step has two forward routes chosen by batch size (drt_tuning_set "direct_max_rays"): the staged wavefront the
    benchmark times and the one-thread-per-path kernel for single views.  Tests that use this fixture run on BOTH at every size."""
    from drt_b200 import _lib
    lib = _lib.load()
    old = lib.drt_tuning_get(b"direct_max_rays")
    _lib.call("drt_tuning_set", b"direct_max_rays", 0 if request.param == "staged" else 2_000_000_000)
    yield request.param
    _lib.call("drt_tuning_set", b"direct_max_rays", old)
