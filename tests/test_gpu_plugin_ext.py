"""B1 literally: the reference's pybind plugin class `optix.optix_mesh` (optix_extend.cpp:77-83) as a COMPILED torch
extension over the C ABI (drt_b200/csrc/optix_extend_b200.cpp = the binding INTEGRATION.md section 4 shows), driven exactly
the way DiffRender.py drives the reference plugin (DiffRender.py:311-313 update_mesh, :379-380 update_vert, :386-392 intersect)."""
import numpy as np
import pytest
import torch

from conftest import load_mesh
from oracle import oracle
from test_gpu_parity import _random_rays

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def plugin():
    from drt_b200 import build
    return build.load_torch_plugin()


def test_compiled_plugin_closest_hit_bit_exact_vs_brute_force(cuda_device, plugin):
    v, f = load_mesh("hand_vh")
    ray = _random_rays(v, 20000, 1)
    extra = _random_rays(v, 4000, 2)
    extra[:2000, 3:] = v[np.arange(2000) % len(v)].astype(np.float32) - extra[:2000, :3]   # aimed at vertices
    e = f[np.arange(2000) % len(f)]
    extra[2000:, 3:] = (0.5 * (v[e[:, 0]] + v[e[:, 1]])).astype(np.float32) - extra[2000:, :3]  # and at edge midpoints
    ray = np.concatenate([ray, extra], 0)
    T0, I0 = oracle.OracleMesh(v, f).closest_hit(ray, use_bvh=False)
    om = plugin.optix_mesh(cuda_device.index or 0)
    F = torch.tensor(f, dtype=torch.int32, device=cuda_device)
    om.update_mesh(F, torch.tensor(v, dtype=torch.float32, device=cuda_device))          # DiffRender.py:311-313
    T, I = om.intersect(torch.tensor(ray, device=cuda_device))                            # DiffRender.py:389
    assert T.stride() == (2,) and I.stride() == (2,) and I.dtype == torch.int32           # {float t; int id} records
    assert np.array_equal(I.cpu().numpy(), I0) and np.array_equal(T.cpu().numpy(), T0)
    # update_vert = new positions, same faces, full rebuild (DiffRender.py:379-380)
    v2 = v + 0.3 * np.sin(v[:, [1, 2, 0]] * 0.1)
    om.update_vert(torch.tensor(v2, dtype=torch.float32, device=cuda_device))
    T2, I2 = om.intersect(torch.tensor(ray, device=cuda_device))
    T3, I3 = oracle.OracleMesh(v2, f).closest_hit(ray, use_bvh=False)
    assert np.array_equal(I2.cpu().numpy(), I3) and np.array_equal(T2.cpu().numpy(), T3)


def test_compiled_plugin_matches_ctypes_binding_and_rejects_bad_arguments(cuda_device, plugin):
    from drt_b200 import optix
    v, f = load_mesh("mouse_vh")
    ray = torch.tensor(_random_rays(v, 5000, 7), device=cuda_device)
    F = torch.tensor(f, dtype=torch.int32, device=cuda_device)
    V = torch.tensor(v, dtype=torch.float32, device=cuda_device)
    a, b = plugin.optix_mesh(cuda_device.index or 0), optix.optix_mesh(cuda_device.index or 0)
    a.update_mesh(F, V)
    b.update_mesh(F, V)
    (Ta, Ia), (Tb, Ib) = a.intersect(ray), b.intersect(ray)
    assert torch.equal(Ia, Ib) and torch.equal(Ta, Tb)
    with pytest.raises(RuntimeError):
        a.intersect(ray[:, :5].contiguous())          # assert(Ray.size(1) == 6), optix_extend.cpp:31
    with pytest.raises(RuntimeError):
        a.update_mesh(F.to(torch.int64), V)
    with pytest.raises(RuntimeError):
        a.update_vert(V.cpu())
