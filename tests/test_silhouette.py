"""N1: silhouette edge sampling against the reference's own silhouette_edge /
primary_visibility / primary_edge_sample outputs (tests/golden/silhouette_hand_vh.npz)."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLD, load_mesh
from drt_b200 import silhouette, trimesh_lite, views
from oracle import oracle


class _Ray:
    def __init__(self, origin, direction):
        self.origin, self.direction = origin, direction


def _golden():
    return np.load(os.path.join(GOLD, "silhouette_hand_vh.npz"))


def _camera(v, z, device):
    resy, resx = (int(x) for x in z["res"])
    cams = views.turntable_cameras(v, resy, resx, 72)
    cam = tuple(torch.tensor(m, device=device) for m in cams[int(z["view"])])
    return cam, resy, resx


def _check(z, sil, index, output, grad, dih):
    assert np.array_equal(sil, z["sil_edges"])
    assert np.array_equal(index, z["index"]) and np.allclose(output, z["output"])
    assert np.abs(grad - z["grad_V"]).max() <= 1e-9 * np.abs(z["grad_V"]).max()
    assert np.allclose(dih, z["dihedral_cos"], atol=1e-12)


def test_silhouette_host_logic_with_oracle_intersector():
    z = _golden()
    v, f = load_mesh("hand_vh")
    cam, resy, resx = _camera(v, z, "cpu")
    mesh = trimesh_lite.TriMesh(v, f)
    V = torch.tensor(v, requires_grad=True)
    F = torch.tensor(f)
    Edges, E2F, mean_len = silhouette.build_edge_tables(mesh, F, "cpu")
    assert Edges.shape == (len(f) * 3 // 2, 2) and 2.5 < mean_len < 4.0
    om = oracle.OracleMesh(v, f)

    def intersect(ray):
        r6 = torch.cat([ray.origin.float(), ray.direction.float()], dim=1).numpy()
        T, ID = om.closest_hit(r6)
        return torch.from_numpy(ID.astype(np.int64)), torch.from_numpy(T > 0)

    origin = cam[2][:3, 3].clone()
    sil = silhouette.silhouette_edges(V, Edges, E2F, origin)
    index, output = silhouette.primary_visibility(V, sil, cam, origin, intersect, _Ray, resy, resx, detach_depth=True)
    (output * torch.tensor(z["weights"], dtype=output.dtype)).sum().backward()
    dih = silhouette.dihedral_cos(V.detach(), E2F).numpy()
    _check(z, sil.numpy(), index.numpy(), output.detach().numpy(), V.grad.numpy(), dih)


@pytest.mark.gpu
def test_silhouette_through_scene_on_gpu(cuda_device):
    import drt_b200.DiffRender as R
    z = _golden()
    v, f = load_mesh("hand_vh")
    cam, resy, resx = _camera(v, z, cuda_device)
    R.resy, R.resx = resy, resx
    sc = R.Scene(vertices=v, faces=f, cuda_device=cuda_device.index or 0)
    V = sc.vertices.clone().requires_grad_(True)
    sc.update_verticex(V)
    origin = cam[2][:3, 3].clone()
    sil = sc.silhouette_edge(origin)
    index, output = sc.primary_visibility(sil, cam, origin, detach_depth=True)
    (output * torch.tensor(z["weights"], dtype=output.dtype, device=cuda_device)).sum().backward()
    _check(z, sil.cpu().numpy(), index.cpu().numpy(), output.detach().cpu().numpy(), V.grad.cpu().numpy(),
           sc.dihedral_angle().detach().cpu().numpy())
    assert abs(sc.mean_len - 3.32) < 0.05  # hand_vh mean edge 3.32
