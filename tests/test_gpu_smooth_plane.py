"""The two OPTIONAL, non-parity modes BASELINE.json's north-star names and the reference does not run,
through the C ABI (drt_trace_fwd_smooth / drt_trace_bwd_smooth, drt_plane_hit / drt_plane_hit_bwd):
  * smooth-normal mode -- the interpolation the reference keeps commented out in JIT_Dintersect (DiffRender.py:107-114) --
    against its PyTorch restatement with autograd (oracle/chain_torch.py) on the oracle's hit ids; vertex normals =
    Scene.init_VN (DiffRender.py:319-336), pinned to the reference by tests/test_oracle_golden.py;
  * the background-plane step against the same expression under torch autograd.
Parity with the reference is "unpinned" for both by construction: the reference never executes them."""
import numpy as np
import pytest
import torch

from conftest import grad_rel_err, load_mesh
from oracle import chain_torch, oracle

pytestmark = pytest.mark.gpu

INT_IOR = 1.4723


def _view(v, res, k):
    from drt_b200 import views
    cams = views.turntable_cameras(v, res[0], res[1], 72)
    return views.generate_ray(res[0], res[1], cams[k][3], cams[k][2])


@pytest.mark.parametrize("mesh,res,k", [("hand_vh", (96, 128), 11), ("mouse_vh", (120, 104), 40)])
def test_smooth_normal_mode_vs_torch_restatement(cuda_device, mesh, res, k):
    import drt_b200.DiffRender as R
    v, f = load_mesh(mesh)
    o, d = _view(v, res, k)
    rng = np.random.default_rng(5)
    g_ori, g_dir = torch.tensor(rng.standard_normal(o.shape)), torch.tensor(rng.standard_normal(o.shape))

    # CPU: the restated op chain with interpolated normals, autograd through init_VN back to the vertices
    m = oracle.OracleMesh(v, f)

    def intersect(ray6):
        T, ID = m.closest_hit(ray6.numpy(), use_bvh=True)
        return torch.from_numpy(T), torch.from_numpy(ID)

    Vc = torch.tensor(v, requires_grad=True)
    Fc = torch.tensor(f.astype(np.int64))
    VNc = chain_torch.vertex_normals(Vc, Fc)
    VNc.retain_grad()
    oo_c, od_c, mk_c = chain_torch.render_transparent(Vc, Fc, o, d, intersect, INT_IOR, normals=VNc)
    ((oo_c * g_ori).sum() + (od_c * g_dir).sum()).backward()

    # GPU: Scene in smooth-normal mode
    R.intIOR = INT_IOR
    R.resy, R.resx = res
    sc = R.Scene(vertices=v, faces=f, cuda_device=cuda_device.index or 0)
    sc.smooth_normals = True
    Vg = sc.vertices.detach().clone().requires_grad_(True)
    sc.update_verticex(Vg)
    oo, od, mk = sc.render_transparent(o.to(cuda_device), d.to(cuda_device))
    sc.normals.retain_grad()
    ((oo * g_ori.to(cuda_device)).sum() + (od * g_dir.to(cuda_device)).sum()).backward()

    mk_g, mk_r = mk[:, 0].cpu().numpy(), mk_c[:, 0].numpy()
    assert mk_r.sum() > 200
    # a ray whose refraction is within rounding of total internal reflection may be decided differently by the two evaluation
    # orders of the interpolated normal: allow a handful, compare the rest
    differ = mk_g != mk_r
    assert differ.sum() <= 2, int(differ.sum())
    both = mk_g & mk_r
    assert np.abs(oo.detach().cpu().numpy()[both] - oo_c.detach().numpy()[both]).max() < 1e-10
    assert np.abs(od.detach().cpu().numpy()[both] - od_c.detach().numpy()[both]).max() < 1e-12
    assert (oo.detach().cpu().numpy()[~mk_g] == 0).all() and (od.detach().cpu().numpy()[~mk_g] == 0).all()
    if not differ.any():
        # Jacobian w.r.t. the interpolated (vertex) normals, and the total vertex gradient (normals chained through init_VN)
        pv, gl = grad_rel_err(sc.normals.grad.cpu().numpy(), VNc.grad.numpy())
        assert pv < 1e-8 and gl < 1e-10, ("normals", pv, gl)
        pv, gl = grad_rel_err(Vg.grad.cpu().numpy(), Vc.grad.numpy())
        assert pv < 1e-8 and gl < 1e-10, ("vertices", pv, gl)
    # the default (flat normals, the reference's live behaviour) is a different result and is untouched by the switch
    sc.smooth_normals = False
    with torch.no_grad():
        od_flat = sc.render_transparent(o.to(cuda_device), d.to(cuda_device))[1]
    assert (od_flat - od.detach()).abs().max() > 1e-3


def test_smooth_entry_points_reject_bad_arguments(cuda_device):
    import drt_b200.DiffRender as R
    from drt_b200 import _lib
    v, f = load_mesh("hand_vh")
    sc = R.Scene(vertices=v, faces=f, cuda_device=cuda_device.index or 0)
    o, d = _view(v, (16, 16), 3)
    with pytest.raises(ValueError):
        R.RefractTraceSmooth.apply(sc.vertices, sc.vertices[:-1], o.to(cuda_device), d.to(cuda_device), sc.optix_mesh, INT_IOR, 1.00029)
    with pytest.raises(TypeError):
        R.RefractTraceSmooth.apply(sc.vertices, sc.vertices.float(), o.to(cuda_device), d.to(cuda_device), sc.optix_mesh, INT_IOR, 1.00029)
    lib = _lib.load()
    assert lib.drt_trace_fwd_smooth(sc.optix_mesh._h, sc.vertices.data_ptr(), None, o.to(cuda_device).data_ptr(), d.to(cuda_device).data_ptr(), 4,
                                    1.00029, INT_IOR, None, None, None, None, None, None) != 0
    assert lib.drt_plane_hit(None, None, None, 4, None, None, None, None) != 0


def test_background_plane_step_vs_torch(cuda_device):
    import drt_b200.DiffRender as R
    v, f = load_mesh("hand_vh")
    res = (96, 128)
    o, d = _view(v, res, 30)
    R.intIOR = INT_IOR
    R.resy, R.resx = res
    sc = R.Scene(vertices=v, faces=f, cuda_device=cuda_device.index or 0)
    with torch.no_grad():
        oo, od, mk = sc.render_transparent(o.to(cuda_device), d.to(cuda_device))
    valid = mk[:, 0]
    assert int(valid.sum()) > 200
    # a plane behind the object as seen from the camera (most exit rays reach it), tilted; and one the rays never reach
    ctr = torch.tensor(0.5 * (v.min(0) + v.max(0)))
    view_dir = (ctr - o[0]) / (ctr - o[0]).norm()
    p0 = (ctr + 120.0 * view_dir).tolist()
    nrm = (view_dir + torch.tensor([0.1, -0.05, 0.02], dtype=torch.float64)).tolist()
    g = torch.randn(oo.shape, dtype=torch.float64, generator=torch.Generator().manual_seed(2))
    a, b = oo.clone().requires_grad_(True), od.clone().requires_grad_(True)
    pts, front = R.PlaneHit.apply(a, b, mk, p0, nrm)
    (pts * g.to(cuda_device)).sum().backward()
    ac, bc = oo.cpu().clone().requires_grad_(True), od.cpu().clone().requires_grad_(True)
    pts_c, front_c = chain_torch.plane_hit(ac, bc, mk.cpu(), p0, nrm)
    (pts_c * g).sum().backward()
    assert torch.equal(front.cpu(), front_c) and 0.5 * int(valid.sum()) < int(front.sum()) <= int(valid.sum())
    assert (pts.detach().cpu() - pts_c.detach()).abs().max() < 1e-10
    assert (a.grad.cpu() - ac.grad).abs().max() < 1e-10 * max(1.0, ac.grad.abs().max().item())
    assert (b.grad.cpu() - bc.grad).abs().max() < 1e-10 * max(1.0, bc.grad.abs().max().item())
    assert (pts.detach()[~front] == 0).all() and (a.grad[~front] == 0).all()
    # points lie on the plane
    on = ((pts.detach()[front].cpu() - torch.tensor(p0, dtype=torch.float64)) @ torch.tensor(nrm, dtype=torch.float64)).abs().max()
    assert on < 1e-9
    # the plane on the camera's side is behind every exit ray's origin or in front of few: never counted when s <= 0
    p_back = (ctr - 400.0 * view_dir).tolist()
    pts_b, front_b = R.PlaneHit.apply(oo, od, mk, p_back, view_dir.tolist())
    pts_bc, front_bc = chain_torch.plane_hit(oo.cpu(), od.cpu(), mk.cpu(), p_back, view_dir.tolist())
    assert torch.equal(front_b.cpu(), front_bc) and (pts_b.cpu() - pts_bc).abs().max() < 1e-10
    # Scene.render_background is the composition; gradients reach the vertices
    Vg = sc.vertices.detach().clone().requires_grad_(True)
    sc.update_verticex(Vg)
    pts2, front2 = sc.render_background(o.to(cuda_device), d.to(cuda_device), p0, nrm)
    assert torch.equal(front2, front) and torch.equal(pts2.detach(), pts.detach())
    (pts2 * g.to(cuda_device)).sum().backward()
    assert torch.isfinite(Vg.grad).all() and Vg.grad.abs().max() > 0
