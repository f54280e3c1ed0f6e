"""Parity ON THE CONFIGS THE NUMBERS ARE QUOTED ON (BASELINE.json configs 4 and 5): the C4 mesh (horse_vh 1->4 subdivided,
50 248 triangles) and the C5 torus (200 000 triangles -- the deepest tree and the most Morton-key collisions) against the
CPU oracle, through the C ABI:
  * the >= 64 k-ray stratified sample of the config's own rays that the roofline denominator is frozen on
    (views.stratified_sample = tools/freeze_counters.py) through drt_closest_hit, drt_trace_fwd and drt_ray_loss_step;
  * one FULL C4 view (960x720 = 691 200 rays, pixel tiles on) through the fused step bench.py times.
Bars: hit ids, T, mask and the exit rays bit-exact; loss <= 1e-12 relative; vertex gradient <= 1e-9 per vertex.
Reference lines: DiffRender.py:386-392 (query), :420-432 (render_transparent), optim.py:96-106 (ray_loss), :210 (backward).
"""
import numpy as np
import pytest
import torch

from conftest import grad_rel_err
from oracle import oracle

pytestmark = [pytest.mark.gpu, pytest.mark.usefixtures("fwd_route")]


def _cfg(name):
    from drt_b200 import configs
    return configs, configs.make(name)


def _oracle_step(m, o, d, screen, valid, int_ior):
    """optim.py:96-106 + :210 on the oracle's render_transparent / backward"""
    q = m.trace_fwd(o, d, int_ior)
    use = q["mask"][:, 0] & valid
    tg = screen - q["out_ori"]
    with np.errstate(invalid="ignore", divide="ignore"):
        tg = tg / np.linalg.norm(tg, axis=1, keepdims=True)
    diff = np.where(use[:, None], q["out_dir"] - tg, 0.0)
    loss = float((diff[use] ** 2).sum())
    gV = m.trace_bwd(o, d, q["tri1"], q["tri2"], None, 2.0 * diff, int_ior)
    return q, loss, gV


def _seeded_targets(q, seed):
    """a measured screen point for ~85 % of the valid paths and a few rays that are not traced: optim.py:96, captured_data.py:104"""
    rng = np.random.default_rng(seed)
    n = len(q["out_ori"])
    screen = q["out_ori"] + 95.0 * q["out_dir"] + 0.5 * rng.standard_normal((n, 3))
    valid = (q["mask"][:, 0] & (rng.random(n) < 0.85)) | (rng.random(n) < 0.01)
    return np.ascontiguousarray(screen * valid[:, None]), valid


@pytest.mark.parametrize("name", ["C4", "C5"])
def test_headline_config_sample_vs_oracle(cuda_device, name):
    import drt_b200.DiffRender as R
    from drt_b200 import losses, views
    configs, cfg = _cfg(name)
    R.intIOR = configs.INT_IOR
    o, d, stride = views.stratified_sample(cfg, 65536, device=cuda_device)
    assert len(o) >= 65536
    m = oracle.OracleMesh(cfg["vertices"], cfg["faces"])
    sc = R.Scene(vertices=cfg["vertices"], faces=cfg["faces"], cuda_device=cuda_device.index or 0)
    og, dg = torch.tensor(o, device=cuda_device), torch.tensor(d, device=cuda_device)

    # ---- optix_mesh.intersect (optix_extend.cpp:29-57) on the float32 casts: T and ID bit-exact --------------
    ray6 = torch.cat([og.float(), dg.float()], dim=1)
    T, ID = sc.optix_mesh.intersect(ray6)
    T_ref, ID_ref = m.closest_hit(ray6.cpu().numpy())
    assert np.array_equal(ID.cpu().numpy(), ID_ref)
    assert np.array_equal(T.cpu().numpy().view(np.uint32), T_ref.view(np.uint32))
    assert 0.05 < (ID_ref >= 0).mean() < 0.6

    # ---- render_transparent (DiffRender.py:420-432): mask and exit rays bit-exact ---------------------------
    q = m.trace_fwd(o, d, configs.INT_IOR)
    with torch.no_grad():
        oo, od, mk = sc.render_transparent(og, dg)
    assert np.array_equal(mk.cpu().numpy(), q["mask"])
    assert np.array_equal(oo.cpu().numpy().view(np.uint64), q["out_ori"].view(np.uint64))
    assert np.array_equal(od.cpu().numpy().view(np.uint64), q["out_dir"].view(np.uint64))
    assert q["mask"][:, 0].sum() > 1000

    # ---- the fused step (optim.py:96-106 + :210): loss and gradient ------------------------------------------
    screen, valid = _seeded_targets(q, seed=11)
    _, ref_loss, ref_g = _oracle_step(m, o, d, screen, valid, configs.INT_IOR)
    V = sc.vertices.detach().clone().requires_grad_(True)
    sc.update_verticex(V)
    n_paths = torch.zeros(1, dtype=torch.int32, device=cuda_device)
    tg = losses.SparseTargets.from_dense(torch.tensor(screen, device=cuda_device), torch.tensor(valid, device=cuda_device))
    loss = losses.ray_loss(sc, og, dg, targets=tg, n_paths=n_paths)
    loss.backward()
    assert int(n_paths.item()) == int(q["mask"][:, 0].sum())
    assert abs(loss.item() - ref_loss) <= 1e-12 * abs(ref_loss), (loss.item(), ref_loss)
    pv, gl = grad_rel_err(V.grad.cpu().numpy(), ref_g)
    assert pv < 1e-9 and gl < 1e-11, (pv, gl)


def test_full_c4_view_vs_oracle(cuda_device):
    """One whole view of the benchmark workload exactly as bench.py feeds it: one origin row, ray_dir [691 200, 3], sparse
    targets, the image-size hint that turns a warp's batch into a pixel tile."""
    import drt_b200.DiffRender as R
    from drt_b200 import losses, views
    configs, cfg = _cfg("C4")
    R.intIOR = configs.INT_IOR
    resy, resx = cfg["resy"], cfg["resx"]
    cam = cfg["cams"][17]
    o_t, d_t = views.generate_ray(resy, resx, cam[3], cam[2])
    o, d = o_t.numpy(), d_t.numpy()
    m = oracle.OracleMesh(cfg["vertices"], cfg["faces"])
    q = m.trace_fwd(o, d, configs.INT_IOR)
    screen, valid = _seeded_targets(q, seed=5)
    _, ref_loss, ref_g = _oracle_step(m, o, d, screen, valid, configs.INT_IOR)

    sc = R.Scene(vertices=cfg["vertices"], faces=cfg["faces"], cuda_device=cuda_device.index or 0)
    R.resy, R.resx = resy, resx
    og, dg = o_t.to(cuda_device), d_t.to(cuda_device)
    with torch.no_grad():
        oo, od, mk = sc.render_transparent(og, dg)      # dense route, 8x4 tiles (drt_trace_fwd)
    assert np.array_equal(mk.cpu().numpy(), q["mask"])
    assert np.array_equal(oo.cpu().numpy().view(np.uint64), q["out_ori"].view(np.uint64))
    assert np.array_equal(od.cpu().numpy().view(np.uint64), q["out_dir"].view(np.uint64))
    # entry-hit ids of all 691 200 primary rays
    ray6 = torch.cat([og.float(), dg.float()], dim=1)
    T, ID = sc.optix_mesh.intersect(ray6)
    T_ref, ID_ref = m.closest_hit(ray6.cpu().numpy())
    ids = ID.cpu().numpy()
    assert np.array_equal(ids, ID_ref) and np.array_equal(T.cpu().numpy().view(np.uint32), T_ref.view(np.uint32))
    assert np.array_equal(ids >= 0, q["stage"] >= 1)
    ok = q["mask"][:, 0]                      # the oracle's forward records tri1 for the valid paths
    assert np.array_equal(ids[ok], q["tri1"][ok])

    tg = losses.SparseTargets.from_dense(torch.tensor(screen, device=cuda_device), torch.tensor(valid, device=cuda_device))
    for image_size in ((resy, resx), None):             # 4x8 pixel tiles / scanline batches
        V = sc.vertices.detach().clone().requires_grad_(True)
        sc.update_verticex(V)
        n_paths = torch.zeros(1, dtype=torch.int32, device=cuda_device)
        loss = losses.ray_loss(sc, og[:1], dg, targets=tg, n_paths=n_paths, image_size=image_size)
        loss.backward()
        assert int(n_paths.item()) == int(q["mask"][:, 0].sum())
        assert abs(loss.item() - ref_loss) <= 1e-12 * abs(ref_loss), (image_size, loss.item(), ref_loss)
        pv, gl = grad_rel_err(V.grad.cpu().numpy(), ref_g)
        assert pv < 1e-9 and gl < 1e-11, (image_size, pv, gl)
