"""The fused ray-loss step (drt_ray_loss_step / drt_b200.losses.ray_loss) against
  * the CPU oracle: forward chain + the reference's ray_loss expression (optim.py:96-106) in numpy + the
    oracle's analytic backward,
  * the reference's expression evaluated by torch autograd on top of render_transparent,
  * its own equivalent input layouts (dense / sparse targets, per-ray / shared / per-view origins),
and drt_generate_rays against captured_data.generate_ray's torch evaluation."""
import ctypes as C
import os

import numpy as np
import pytest
import torch

from conftest import grad_rel_err, load_mesh
from oracle import oracle

pytestmark = [pytest.mark.gpu, pytest.mark.usefixtures("fwd_route")]

INT_IOR = 1.4723


def _scene(v, f, dev):
    import drt_b200.DiffRender as R
    R.intIOR = INT_IOR
    return R, R.Scene(vertices=v, faces=f, cuda_device=dev.index or 0)


def _oracle_loss_and_grad(v, f, o, d, screen, valid):
    """optim.py:96-106 on the oracle's render_transparent + the oracle's backward"""
    m = oracle.OracleMesh(v, f)
    q = m.trace_fwd(o, d, INT_IOR)
    use = q["mask"][:, 0] & valid
    tg = screen - q["out_ori"]
    with np.errstate(invalid="ignore", divide="ignore"):
        tg = tg / np.linalg.norm(tg, axis=1, keepdims=True)
    diff = np.where(use[:, None], q["out_dir"] - tg, 0.0)
    loss = float((diff[use] ** 2).sum())
    gV = m.trace_bwd(o, d, q["tri1"], q["tri2"], None, 2.0 * diff, INT_IOR)
    return loss, gV, int(q["mask"][:, 0].sum())


def _views_of(v, res, ks):
    from drt_b200 import views
    cams = views.turntable_cameras(v, res[0], res[1], 72)
    rays = [views.generate_ray(res[0], res[1], cams[k][3], cams[k][2]) for k in ks]
    return torch.cat([r[0] for r in rays]), torch.cat([r[1] for r in rays]), [r[0][:1] for r in rays]


def _targets(sc, o, d, seed, keep=0.8):
    g = torch.Generator(device="cpu").manual_seed(seed)
    with torch.no_grad():
        oo, od, mk = sc.render_transparent(o, d)
    noise = torch.randn(o.shape, generator=g, dtype=torch.float64).to(o.device)
    screen = (oo + 90.0 * od + 0.7 * noise).contiguous()
    valid = (torch.rand(len(o), generator=g) < keep).to(o.device) & mk[:, 0] | (torch.rand(len(o), generator=g) < 0.02).to(o.device)
    screen = screen * valid[:, None]
    return screen, valid, mk


@pytest.mark.parametrize("mesh,res,ks", [("hand_vh", (96, 128), (11,)), ("mouse_vh", (120, 104), (3, 40, 57))])
def test_loss_step_vs_oracle_and_every_input_layout(cuda_device, fwd_route, mesh, res, ks):
    from drt_b200 import losses
    v, f = load_mesh(mesh)
    R, sc = _scene(v, f, cuda_device)
    o_cpu, d_cpu, cam_o = _views_of(v, res, ks)
    o, d = o_cpu.to(cuda_device), d_cpu.to(cuda_device)
    screen, valid, mk = _targets(sc, o, d, seed=3)
    ref_loss, ref_g, ref_paths = _oracle_loss_and_grad(v, f, o_cpu.numpy(), d_cpu.numpy(), screen.cpu().numpy(), valid.cpu().numpy())
    assert ref_paths == int(mk[:, 0].sum().item()) and ref_paths > 100

    def run(origin, **kw):
        V = sc.vertices.detach().clone().requires_grad_(True)
        sc.update_verticex(V)
        n_paths = torch.zeros(1, dtype=torch.int32, device=cuda_device)
        loss = losses.ray_loss(sc, origin, d, n_paths=n_paths, **kw)
        (2.5 * loss).backward()
        return loss.item(), V.grad.cpu().numpy() / 2.5, int(n_paths.item())

    sparse = losses.SparseTargets.from_dense(screen, valid)
    assert len(sparse) == int(valid.sum().item())
    per_view = torch.cat(cam_o).to(cuda_device)                       # one origin row per view
    layouts = {
        "dense targets, origin per ray": run(o, screen=screen, valid=valid),
        "sparse targets, origin per ray": run(o, targets=sparse),
        "sparse targets, origin per view": run(per_view, targets=sparse),
        "dense targets, origin per view": run(per_view, screen=screen, valid=valid),
        # image size known: the entry query walks 32-pixel tiles (4x8, else 8x4) instead of 32x1 strips -- same paths, same numbers
        "tiles, sparse targets, origin per view": run(per_view, targets=sparse, image_size=res),
        "tiles, dense targets, origin per ray": run(o, screen=screen, valid=valid, image_size=res),
        "untileable image size hint is ignored": run(per_view, targets=sparse, image_size=(res[0] + 1, res[1])),
    }
    if len(ks) == 1:
        layouts["expanded origin"] = run(per_view.expand(len(d), 3), targets=sparse)
    # per-tile direction intervals prepared once for the (fixed) rays instead of being re-derived in every step: same culling
    counts = {}
    for name, origin, size in (("tiles, origin per view", per_view, res), ("strips, origin per view", per_view, None), ("tiles, origin per ray", o, res)):
        beams = losses.prepare_tile_beams(origin, d, size)
        layouts["prepared beams, " + name] = run(origin, targets=sparse, image_size=size, tile_beams=beams)
        counts[name] = sc.optix_mesh.last_counts()
        run(origin, targets=sparse, image_size=size)
        assert sc.optix_mesh.last_counts() == counts[name], name
        if fwd_route == "staged" and os.environ.get("DRT_BEAM") != "0":  # the direct route has no beam pass; DRT_BEAM=0 switches it off
            assert 0 < counts[name]["tiles_kept"] < counts[name]["tiles"]
    # a buffer prepared for another tile map does not match the call's signature: ignored, the rays are scanned as usual
    layouts["stale beams are ignored"] = run(per_view, targets=sparse, image_size=res, tile_beams=losses.prepare_tile_beams(per_view, d, None))
    assert sc.optix_mesh.last_counts() == counts["tiles, origin per view"]
    for name, (loss, g, n_paths) in layouts.items():
        assert n_paths == ref_paths, name
        assert abs(loss - ref_loss) <= 1e-12 * abs(ref_loss), (name, loss, ref_loss)
        pv, gl = grad_rel_err(g, ref_g)
        assert pv < 1e-9 and gl < 1e-11, (name, pv, gl)
    # the three-call route (dense out_ori/out_dir/g_out_dir in memory) gives the same numbers
    V = sc.vertices.detach().clone().requires_grad_(True)
    sc.update_verticex(V)
    old = losses.ray_loss_rec(sc, o, d, screen, valid)
    old.backward()
    assert abs(old.item() - ref_loss) <= 1e-12 * abs(ref_loss)
    pv, gl = grad_rel_err(V.grad.cpu().numpy(), ref_g)
    assert pv < 1e-9 and gl < 1e-11


def test_loss_step_without_grad_and_valid_none(cuda_device):
    """vertices without requires_grad -> loss value only (no gradient kernel work); valid=None = all true."""
    from drt_b200 import losses
    v, f = load_mesh("hand_vh")
    R, sc = _scene(v, f, cuda_device)
    o_cpu, d_cpu, _ = _views_of(v, (80, 80), (25,))
    o, d = o_cpu.to(cuda_device), d_cpu.to(cuda_device)
    with torch.no_grad():
        oo, od, mk = sc.render_transparent(o, d)
    screen = oo + 50.0 * od + 0.3
    a = losses.ray_loss(sc, o, d, screen=screen)
    assert not a.requires_grad
    tg = screen - oo
    tg = tg / tg.norm(dim=1, keepdim=True)
    ref = ((od - tg)[mk[:, 0]]).pow(2).sum()
    assert abs(a.item() - ref.item()) <= 1e-12 * abs(ref.item())


def test_loss_step_degenerate_batches(cuda_device):
    from drt_b200 import losses
    v, f = load_mesh("hand_vh")
    R, sc = _scene(v, f, cuda_device)
    V = sc.vertices.detach().clone().requires_grad_(True)
    sc.update_verticex(V)
    empty = torch.zeros((0, 3), dtype=torch.float64, device=cuda_device)
    z = losses.ray_loss(sc, empty, empty, screen=empty)
    z.backward()
    assert z.item() == 0.0 and not V.grad.any()
    # all rays miss
    o = torch.tensor([[0.0, 0.0, 1e4]], dtype=torch.float64, device=cuda_device).expand(4096, 3)
    d = torch.tensor([[0.0, 0.0, 1.0]], dtype=torch.float64, device=cuda_device).repeat(4096, 1)
    n_paths = torch.ones(1, dtype=torch.int32, device=cuda_device)
    m = losses.ray_loss(sc, o, d, targets=losses.SparseTargets(torch.zeros(0, dtype=torch.int32, device=cuda_device),
                                                              torch.zeros((0, 3), dtype=torch.float64, device=cuda_device)), n_paths=n_paths)
    assert m.item() == 0.0 and n_paths.item() == 0
    # empty mesh
    e = R.Scene(vertices=np.zeros((0, 3)), faces=np.zeros((0, 3), dtype=np.int64), cuda_device=cuda_device.index or 0)
    assert losses.ray_loss(e, o, d, screen=d).item() == 0.0
    with pytest.raises(ValueError):
        losses.ray_loss(sc, o[:7], d, screen=d)          # 7 origin rows do not divide 4096 rays
    with pytest.raises(ValueError):
        losses.ray_loss(sc, o, d)                        # no targets at all


def test_loss_step_large_batch_matches_three_call_route(cuda_device):
    """1.3 M rays (several persistent batches per warp, merge and no-merge scatter both reachable)."""
    from drt_b200 import losses, views
    v, f = load_mesh("mouse_vh")
    R, sc = _scene(v, f, cuda_device)
    cams = views.turntable_cameras(v, 1100, 1200, 72)
    o, d = views.generate_ray(1100, 1200, cams[50][3], cams[50][2], device=cuda_device)
    screen, valid, mk = _targets(sc, o, d, seed=9, keep=0.9)
    Va = sc.vertices.detach().clone().requires_grad_(True)
    sc.update_verticex(Va)
    a = losses.ray_loss_rec(sc, o, d, screen, valid)
    a.backward()
    Vb = sc.vertices.detach().clone().requires_grad_(True)
    sc.update_verticex(Vb)
    b = losses.ray_loss(sc, o[:1], d, targets=losses.SparseTargets.from_dense(screen, valid))
    b.backward()
    assert abs(a.item() - b.item()) <= 1e-12 * abs(a.item())
    pv, gl = grad_rel_err(Vb.grad.cpu().numpy(), Va.grad.cpu().numpy())
    assert pv < 1e-10 and gl < 1e-12, (pv, gl)


def test_generate_rays_kernel_vs_generate_ray(cuda_device):
    """drt_generate_rays vs the torch evaluation of captured_data.generate_ray (captured_data.py:23-40): unit
    directions to 4 ulp (the reference leaves the summation order of its two matmuls to the BLAS), same origin."""
    from drt_b200 import views
    v, _ = load_mesh("hand_vh")
    for (resy, resx, k) in ((96, 128, 7), (1, 1, 0), (33, 5, 60)):
        cam = views.turntable_cameras(v, resy, resx, 72)[k]
        o_ref, d_ref = views.generate_ray(resy, resx, cam[3], cam[2])
        o, d = views.generate_ray_device(resy, resx, cam[3], cam[2], cuda_device)
        assert o.shape == (1, 3) and torch.equal(o.cpu(), o_ref[:1])
        err = (d.cpu() - d_ref).abs().max().item()
        assert err <= 4 * np.finfo(np.float64).eps, err
        assert (d.norm(dim=1) - 1).abs().max().item() <= 4.5e-16   # 2 ulp


def test_compact_view_through_the_loader(cuda_device):
    """SyntheticData.get_view_compact + losses.ray_loss_view == the reference-layout view through ray_loss."""
    from drt_b200 import losses
    from drt_b200.synthetic_data import SyntheticData
    v, f = load_mesh("hand_vh")
    data = SyntheticData(v * 1.01, f, 72, 96, n_views=4, num_view=4, cuda_device=cuda_device.index or 0, int_ior=INT_IOR)
    R, sc = _scene(v, f, cuda_device)
    for k in range(4):
        screen, valid, _, origin, ray_dir, _ = data.get_view(k)
        cv = data.get_view_compact(k)
        assert cv.origin.shape == (1, 3) and cv.h2d_bytes() < 0.45 * 73 * len(ray_dir)
        a = losses.ray_loss(sc, origin, ray_dir, screen=screen, valid=valid)
        b = losses.ray_loss_view(sc, cv)
        assert a.item() > 0 and abs(a.item() - b.item()) <= 1e-13 * a.item()
    # several views as one batch (CompactView.concat): the sum of the per-view losses, and the same gradient
    V = sc.vertices.detach().clone().requires_grad_(True)
    sc.update_verticex(V)
    total = sum(losses.ray_loss_view(sc, data.get_view_compact(k)) for k in (0, 2, 3))
    total.backward()
    g_views = V.grad.clone()
    V.grad = None
    batch = data.compact_batch((0, 2, 3)).to(cuda_device)
    assert batch.origin.shape == (3, 3) and batch.ray_dir.shape[0] == 3 * 72 * 96
    both = losses.ray_loss_view(sc, batch)
    both.backward()
    assert abs(both.item() - total.item()) <= 1e-13 * total.item()
    pv, gl = grad_rel_err(V.grad.cpu().numpy(), g_views.cpu().numpy())
    assert pv < 1e-10 and gl < 1e-12, (pv, gl)


def test_loss_step_cabi_errors(cuda_device):
    from drt_b200 import _lib
    lib = _lib.load()
    v, f = load_mesh("hand_vh")
    R, sc = _scene(v, f, cuda_device)
    h = sc.optix_mesh._h
    t = torch.zeros((8, 3), dtype=torch.float64, device=cuda_device)
    p = lambda x: C.c_void_p(x.data_ptr())  # noqa: E731
    V = sc.vertices.detach()
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    err = lambda: lib.drt_last_error().decode()  # noqa: E731
    assert lib.drt_ray_loss_step(h, p(V), p(t), 0, p(t), 8, 1.0, 1.5, 0, p(t), None, None, None, 0, 0, 0, p(t), None, None, None, st) == 1 and "rays_per_origin" in err()
    assert lib.drt_ray_loss_step(h, p(V), p(t), 1, p(t), 8, 1.0, 1.5, 2, p(t), None, None, None, 0, 0, 0, p(t), None, None, None, st) == 1 and "target_mode" in err()
    assert lib.drt_ray_loss_step(h, p(V), p(t), 1, p(t), 8, 1.0, 1.5, 0, None, None, None, None, 0, 0, 0, p(t), None, None, None, st) == 1 and "screen" in err()
    assert lib.drt_ray_loss_step(h, p(V), p(t), 1, p(t), 8, 1.0, 1.5, 1, None, None, None, None, 3, 0, 0, p(t), None, None, None, st) == 1 and "sparse" in err()
    assert lib.drt_ray_loss_step(h, p(V), p(t), 1, p(t), 8, 1.0, 1.5, 0, p(t), None, None, None, 0, 0, 0, None, None, None, None, st) == 1
    assert lib.drt_ray_loss_step(h, p(V), p(t), 1, p(t), 8, 1.0, 1.5, 0, p(t), None, None, None, 0, 0, 0, p(t), None, None, None, st) == 0
    assert lib.drt_generate_rays(-1, 4, p(t), p(t), p(t), p(t), st) == 1
    torch.cuda.synchronize()


def test_tile_beams_hold_the_direction_intervals_of_each_tile(cuda_device):
    """drt_tile_beams against numpy: per 4x8 pixel tile the min / max of the float32-cast directions and the common origin."""
    from drt_b200 import losses
    v, _ = load_mesh("hand_vh")
    res = (64, 96)
    o_cpu, d_cpu, cam_o = _views_of(v, res, (5, 50))
    d = d_cpu.to(cuda_device)
    per_view = torch.cat(cam_o).to(cuda_device)
    n_tiles = len(d) // 32
    b = losses.prepare_tile_beams(per_view, d, res).cpu().numpy().reshape(3, n_tiles + 1, 4)
    hdr = b[0, 0].view(np.int32)
    assert hdr[0] == 0x4D414542 and hdr[1] == len(d) and hdr[2] == res[1] and hdr[3] == ((res[0] * res[1]) << 3 | 2)
    d32 = d_cpu.numpy().astype(np.float32).reshape(2, res[0] // 8, 8, res[1] // 4, 4, 3)
    assert np.array_equal(b[0, 1:, :3], d32.min(axis=(2, 4)).reshape(-1, 3))
    assert np.array_equal(b[1, 1:, :3], d32.max(axis=(2, 4)).reshape(-1, 3))
    assert (b[0, 1:, 3].view(np.uint32) == 3).all()                       # has rays | one origin
    o32 = np.repeat(torch.cat(cam_o).numpy().astype(np.float32), n_tiles // 2, axis=0)
    assert np.array_equal(np.stack([b[1, 1:, 3], b[2, 1:, 0], b[2, 1:, 1]], axis=1), o32)
    # rays with their own origin rows that differ inside a tile: no beam for those tiles
    o_var = o_cpu.clone()
    o_var[::977] += 0.25
    bv = losses.prepare_tile_beams(o_var.to(cuda_device), d, res).cpu().numpy().reshape(3, n_tiles + 1, 4)
    flags = bv[0, 1:, 3].view(np.uint32)
    assert (flags == 1).sum() > 0 and ((flags == 1) | (flags == 3)).all()
