"""B2 -- the Scene contract an UNCHANGED optim.py relies on, replayed call for call:

    scene = Render.Scene(path)                                   optim.py:175
    Meshlabserver.remesh: scene.mesh.export(ply) -> [meshlab] -> scene.update_mesh(remeshed)   optim.py:46-52
    init_vertices = scene.vertices; parameter; SGD              optim.py:164-171
    per iteration: scene.update_verticex(init + parameter) -> all_loss -> backward -> step      optim.py:199-217
    scene.mesh.export(result)                                    optim.py:226

The reference refreshes `scene.mesh.vertices` on every update_verticex (DiffRender.py:381); an export that wrote the
initial vertices would silently restart every remesh pass from the visual hull.
"""
import os
import shutil

import numpy as np
import pytest
import torch


class _StubOptixMesh:
    """CPU stand-in for the plugin object: only what Scene needs to keep its host-side bookkeeping going."""

    def __init__(self, cuda_device=0):
        self.device = torch.device("cpu")
        self.calls = []

    def update_mesh(self, F, V):
        self.calls.append(("update_mesh", tuple(F.shape), tuple(V.shape)))

    def update_vert(self, V, refit=False):
        self.calls.append(("update_vert", tuple(V.shape), refit))

    def bad_indices(self):
        return 0


def _replay(Render, path, tmp_path, n_pass, n_iter, loss_fn, lr=0.05):
    """optim.py:145-219 with MeshLab replaced by a file copy (the remesher is an external binary, absent here)."""
    from drt_b200 import plyio
    scene = Render.Scene(str(path))
    exported = []
    for i_pass in range(n_pass):
        ply, remeshed = tmp_path / f"temp_{i_pass}.ply", tmp_path / f"remesh_{i_pass}.ply"
        scene.mesh.export(str(ply))                                      # optim.py:50
        shutil.copy(ply, remeshed)                                       # optim.py:51 (meshlabserver)
        exported.append(plyio.read_ply(str(ply))[0])
        scene.update_mesh(str(remeshed))                                 # optim.py:52
        init_vertices = scene.vertices                                   # optim.py:166
        parameter = torch.zeros(init_vertices.shape, dtype=torch.float64, requires_grad=True, device=init_vertices.device)
        parameter.register_hook(lambda g: torch.nan_to_num(g, nan=0.0).clamp(-1.0, 1.0))   # optim.py:155-162, 168
        opt = torch.optim.SGD([parameter], lr=lr, momentum=0.9, nesterov=True)
        for _ in range(n_iter):
            opt.zero_grad()
            vertices = init_vertices + parameter                         # optim.py:202
            scene.update_verticex(vertices)                              # optim.py:203
            loss_fn(scene).backward()
            opt.step()
    final = tmp_path / "recons.ply"
    scene.mesh.export(str(final))                                        # optim.py:226
    return scene, exported, plyio.read_ply(str(final))[0]


def test_mesh_export_follows_update_verticex_cpu_stub(tmp_path, monkeypatch):
    import drt_b200.DiffRender as Render
    from drt_b200 import meshgen, optix, plyio
    monkeypatch.setattr(optix, "optix_mesh", _StubOptixMesh)
    v, f = meshgen.icosahedron()
    path = tmp_path / "ico_vh.ply"
    plyio.write_ply(str(path), v, f)

    def loss_fn(scene):  # anything with a gradient w.r.t. the vertices handed to update_verticex
        return (scene.vertices - 1.0).pow(2).sum()

    scene, exported, final = _replay(Render, path, tmp_path, n_pass=2, n_iter=3, loss_fn=loss_fn)
    assert np.allclose(exported[0], v.astype(np.float32), atol=1e-6)              # pass 0 starts from the loaded mesh
    assert np.abs(exported[1] - exported[0]).max() > 1e-3                          # pass 1 starts from the OPTIMISED mesh
    now = scene.vertices.detach().numpy()
    assert np.abs(final - now).max() <= 1e-6 * max(1.0, np.abs(now).max())         # final export = current vertices (PLY is float32)
    assert np.abs(final - exported[1]).max() > 1e-3
    # the D2H refresh is lazy: nothing is copied until .mesh is read
    scene.update_verticex(scene.vertices.detach() + 1.0)
    assert scene._mesh_dirty
    assert np.allclose(scene.mesh.vertices, scene.vertices.detach().numpy())
    assert not scene._mesh_dirty
    # one build for the ctor + one per remesh pass; one vertex update per iteration (+ the one above)
    names = [c[0] for c in scene.optix_mesh.calls]
    assert names.count("update_mesh") == 1 + 2 and names.count("update_vert") == 2 * 3 + 1


def test_set_mesh_rejects_corrupt_faces(monkeypatch):
    import drt_b200.DiffRender as Render
    from drt_b200 import meshgen, optix

    class Bad(_StubOptixMesh):
        def bad_indices(self):
            return 2

    monkeypatch.setattr(optix, "optix_mesh", Bad)
    v, f = meshgen.icosahedron()
    with pytest.raises(ValueError, match="face indices"):
        Render.Scene(vertices=v, faces=f)


@pytest.mark.gpu
def test_optim_py_call_sequence_on_gpu(tmp_path, cuda_device):
    """The same replay on the real Scene: ray loss through render_transparent exactly as optim.py:91-108 writes it, a
    silhouette term (optim.py:67-80) and the smoothness term (optim.py:82-89); exported vertices are the optimised ones."""
    import drt_b200.DiffRender as Render
    from drt_b200 import configs, plyio, synthetic_data
    v, f = configs.load_mesh("hand_vh")
    path = tmp_path / "hand_vh.ply"
    plyio.write_ply(str(path), v, f)
    v = plyio.read_ply(str(path))[0]  # float32-rounded, what Scene(path) sees
    Render.intIOR = 1.4723
    target = configs.perturbed_target_mesh(v, scale=2.0)
    data = synthetic_data.SyntheticData(target, f, 120, 160, n_views=8, num_view=8, int_ior=Render.intIOR)
    Render.resy, Render.resx = data.resy, data.resx
    ray_view, silh_view = data.ray_view_generator(), data.silh_view_generator()

    def all_loss(scene):
        target_px, valid, mask, origin, ray_dir, camera_M = data.get_view(next(ray_view))
        out_ori, out_dir, render_mask = scene.render_transparent(origin, ray_dir)
        tg = target_px - out_ori.detach()
        tg = tg / tg.norm(dim=1, keepdim=True)
        ray_loss = ((out_dir - tg)[valid * render_mask[:, 0]]).pow(2).sum()
        _, _, sil, origin, _, cam = data.get_view(next(silh_view))
        edges = scene.silhouette_edge(origin[0])
        index, output = scene.primary_visibility(edges, cam, origin[0], detach_depth=True)
        vh_loss = (sil.view(data.resy, data.resx)[index[:, 1], index[:, 0]] - output).abs().sum()
        sm_loss = (-torch.log(1 + scene.dihedral_angle())).sum()
        return 40 * 217.5 / data.resy / data.resy * ray_loss + 2e-3 * 217.5 / data.resy * vh_loss + 0.08 * scene.mean_len / 10 * sm_loss

    scene, exported, final = _replay(Render, path, tmp_path, n_pass=2, n_iter=4, loss_fn=all_loss, lr=0.5)
    assert np.abs(exported[0] - v).max() <= 1e-5
    moved = np.abs(exported[1] - exported[0]).max()
    assert moved > 1e-3, moved
    assert moved > 1e-3, "pass 1 was remeshed from the INITIAL mesh: scene.mesh did not follow update_verticex"
    now = scene.vertices.detach().cpu().numpy()
    assert np.abs(final - now).max() <= 1e-5 * max(1.0, np.abs(now).max())
    assert np.abs(final - exported[1]).max() > 1e-4
    assert scene.optix_mesh.info()["builds"] >= 2 * 4  # a rebuild per update_verticex, like DiffRender.py:380
    assert os.path.getsize(tmp_path / "recons.ply") > 0


@pytest.mark.gpu
def test_coarse_to_fine_passes_with_the_remesh_stand_in(tmp_path, cuda_device):
    """optim.py:189-217 with `meshlabserver.remesh(scene, remesh_len)` served by drt_b200.remesh.Remesher: every pass exports the
    optimised mesh, refines it to a shorter target edge length and reloads it through Scene.update_mesh (rebuilding the BVH and
    the edge tables); the ray loss keeps working on the refined mesh."""
    import drt_b200.DiffRender as Render
    from drt_b200 import configs, losses, meshgen, plyio, remesh, synthetic_data
    v, f = configs.load_mesh("hand_vh")
    path = tmp_path / "hand_vh.ply"
    plyio.write_ply(str(path), v, f)
    Render.intIOR = 1.4723
    data = synthetic_data.SyntheticData(configs.perturbed_target_mesh(v, scale=1.5), f, 120, 160, n_views=6, num_view=6, int_ior=Render.intIOR)
    Render.resy, Render.resx = data.resy, data.resx
    scene = Render.Scene(str(path))
    remesher = remesh.Remesher(str(tmp_path))
    ray_view = data.ray_view_generator()
    n_faces, losses_seen = [scene.faces.shape[0]], []
    for remesh_len in (3.0, 2.2):                                       # interp_R(start_len, end_len, ...) of optim.py:192
        remesher.remesh(scene, remesh_len)                              # optim.py:198
        assert scene.mesh.is_watertight
        n_faces.append(scene.faces.shape[0])
        init_vertices = scene.vertices
        parameter = torch.zeros_like(init_vertices, requires_grad=True)
        parameter.register_hook(lambda g: torch.nan_to_num(g, nan=0.0).clamp(-1.0, 1.0))
        opt = torch.optim.SGD([parameter], lr=0.05, momentum=0.9, nesterov=True)
        for _ in range(3):
            opt.zero_grad()
            scene.update_verticex(init_vertices + parameter)
            loss = losses.ray_loss_view(scene, data.get_view_compact(next(ray_view)))
            loss.backward()
            assert torch.isfinite(loss) and parameter.grad.abs().max() > 0
            losses_seen.append(loss.item())
            opt.step()
        # the export the next pass starts from holds the optimised vertices of THIS pass
        exported = plyio.read_ply(scene.mesh.export(str(tmp_path / "check.ply")))[0]
        now = scene.vertices.detach().cpu().numpy()
        assert exported.shape == now.shape and np.abs(exported - now).max() <= 1e-5 * max(1.0, np.abs(now).max())
    assert n_faces[0] < n_faces[1] < n_faces[2]
    assert scene.optix_mesh.info()["n_faces"] == n_faces[2] and scene.dihedral_angle().shape[0] == n_faces[2] * 3 // 2
