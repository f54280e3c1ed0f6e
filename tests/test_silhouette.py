"""N1: silhouette edge sampling against the reference's own silhouette_edge /
primary_visibility / primary_edge_sample outputs (tests/golden/silhouette_hand_vh.npz): the PyTorch restatement
(oracle/silhouette_torch.py) on CPU, the fused kernels (csrc/silhouette.cuh) through Scene on the GPU."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLD, load_mesh
from drt_b200 import silhouette, trimesh_lite, views
from oracle import oracle, silhouette_torch


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
    sil = silhouette_torch.silhouette_edges(V, Edges, E2F, origin)
    index, output = silhouette_torch.primary_visibility(V, sil, cam, origin, intersect, _Ray, resy, resx, detach_depth=True)
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


@pytest.mark.gpu
@pytest.mark.parametrize("mesh,res,view_ids", [("mouse_vh", (240, 320), (0, 23, 50)), ("C4", (720, 960), (7, 31))])
def test_fused_silhouette_kernels_vs_torch_restatement(cuda_device, mesh, res, view_ids):
    """drt_silhouette_classify / _sample / _backward against the PyTorch restatement (itself golden-checked against the
    reference above) on bigger meshes and several views: same silhouette edges in the same order, same sample pixels,
    vertex gradient <= 1e-9; detach_depth on and off (DiffRender.py:468-469)."""
    import drt_b200.DiffRender as R
    from drt_b200 import configs
    if mesh == "C4":
        v, f = configs.make("C4")["vertices"], configs.make("C4")["faces"]
    else:
        v, f = load_mesh(mesh)
    resy, resx = res
    R.resy, R.resx = resy, resx
    sc = R.Scene(vertices=v, faces=f, cuda_device=cuda_device.index or 0)
    cams = views.turntable_cameras(v, resy, resx, 72)
    g = torch.Generator(device="cpu").manual_seed(4)
    for vid in view_ids:
        cam = tuple(torch.tensor(m, device=cuda_device) for m in cams[vid])
        origin = cam[2][:3, 3].clone()
        for detach in (True, False):
            V = sc.vertices.detach().clone().requires_grad_(True)
            sc.update_verticex(V)
            sil = sc.silhouette_edge(origin)
            index, output = sc.primary_visibility(sil, cam, origin, detach_depth=detach)
            w = torch.randn(output.shape[0], generator=g).to(cuda_device)
            (output * w).sum().backward()
            V2 = sc.vertices.detach().clone().requires_grad_(True)
            Edges, E2F = sc._edges()
            sil2 = silhouette_torch.silhouette_edges(V2, Edges, E2F, origin)
            index2, output2 = silhouette_torch.primary_visibility(V2, sil2, cam, origin, sc.optix_intersect, R.Ray, resy, resx, detach_depth=detach)
            (output2 * w).sum().backward()
            assert torch.equal(sil, sil2) and sil.shape[0] > 100
            assert torch.equal(index, index2) and torch.equal(output, output2) and index.shape[0] > 50
            ga, gb = V.grad.cpu().numpy(), V2.grad.cpu().numpy()
            assert np.abs(ga - gb).max() <= 1e-9 * np.abs(gb).max(), (vid, detach)


@pytest.mark.gpu
@pytest.mark.parametrize("detach", [True, False])
def test_fused_silhouette_and_smoothness_losses_vs_the_drop_in_path(cuda_device, detach):
    """losses.silhouette_loss / smoothness_loss (one launch each, no sync) against the same terms written as optim.py writes
    them (optim.py:74-79, 85-87) on the drop-in Scene methods: value <= 1e-12, vertex gradient <= 1e-9."""
    import drt_b200.DiffRender as R
    from drt_b200 import losses, synthetic_data
    v, f = load_mesh("mouse_vh")
    resy, resx = 240, 320
    R.resy, R.resx, R.intIOR = resy, resx, 1.4723
    data = synthetic_data.SyntheticData(v + 0.8 * np.sin(v[:, [1, 2, 0]] * 0.15), f, resy, resx, n_views=6, num_view=6, int_ior=1.4723)
    sc = R.Scene(vertices=v, faces=f, cuda_device=cuda_device.index or 0)
    for vid in (0, 2, 5):
        _, _, sil, origin, _, cam = data.get_view(vid)
        Va = sc.vertices.detach().clone().requires_grad_(True)
        sc.update_verticex(Va)
        edges = sc.silhouette_edge(origin[0])
        index, output = sc.primary_visibility(edges, cam, origin[0], detach_depth=detach)
        ref = (sil.view(resy, resx)[index[:, 1], index[:, 0]] - output).abs().sum() + 0.3 * (-torch.log(1 + sc.dihedral_angle())).sum()
        ref.backward()
        Vb = sc.vertices.detach().clone().requires_grad_(True)
        sc.update_verticex(Vb)
        n = torch.zeros(1, dtype=torch.int32, device=cuda_device)
        got = losses.silhouette_loss(sc, sil, cam, origin[0], detach_depth=detach, n_samples=n) + 0.3 * losses.smoothness_loss(sc)
        got.backward()
        assert int(n.item()) == index.shape[0] and index.shape[0] > 50
        assert abs(got.item() - ref.item()) <= 1e-12 * abs(ref.item())
        ga, gb = Va.grad.cpu().numpy(), Vb.grad.cpu().numpy()
        assert np.abs(ga - gb).max() <= 1e-9 * np.abs(ga).max()
    # several views in ONE call (optim.py:72 sums 8 per iteration) = the sum of the single-view calls; 10 views = two launches
    vids = [0, 1, 2, 3, 4, 5, 0, 2, 4, 1]
    batch = [data.get_view(k) for k in vids]
    Vc = sc.vertices.detach().clone().requires_grad_(True)
    sc.update_verticex(Vc)
    one = sum(losses.silhouette_loss(sc, sil, cam, origin[0], detach_depth=detach) for _, _, sil, origin, _, cam in batch)
    one.backward()
    Vd = sc.vertices.detach().clone().requires_grad_(True)
    sc.update_verticex(Vd)
    n = torch.zeros(1, dtype=torch.int32, device=cuda_device)
    many = losses.silhouette_loss(sc, [(sil, cam, origin[0]) for _, _, sil, origin, _, cam in batch], detach_depth=detach, n_samples=n)
    many.backward()
    assert abs(many.item() - one.item()) <= 1e-12 * abs(one.item()) and int(n.item()) > 500
    assert np.abs(Vc.grad.cpu().numpy() - Vd.grad.cpu().numpy()).max() <= 1e-9 * Vc.grad.abs().max().item()
