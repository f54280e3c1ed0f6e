"""GPU parity: libdrt_b200 (through its C ABI / the Scene mirror) against the CPU oracle and the
reference-generated golden vectors.  Bit-exact for hit ids and masks; endpoints 1e-5 and gradients
1e-4 relative are the north-star bars, the asserts below are far tighter."""
import numpy as np
import pytest
import torch

from conftest import CHAIN_CASES, grad_rel_err, load_chain_case, load_mesh
from oracle import oracle

pytestmark = pytest.mark.gpu

INT_IOR = 1.4723


def _scene(v, f, dev):
    import drt_b200.DiffRender as R
    R.intIOR = INT_IOR
    return R, R.Scene(vertices=v, faces=f, cuda_device=dev.index or 0)


def _random_rays(v, n, seed, unnormalised=True):
    rng = np.random.default_rng(seed)
    ctr = 0.5 * (v.min(0) + v.max(0))
    ext = np.linalg.norm(v.max(0) - v.min(0))
    o = ctr + rng.normal(size=(n, 3)) * ext
    tgt = v[rng.integers(0, len(v), n)] + rng.normal(size=(n, 3)) * 0.02 * ext
    d = tgt - o
    if not unnormalised:
        d /= np.linalg.norm(d, axis=1, keepdims=True)
    return np.concatenate([o, d], 1).astype(np.float32)


@pytest.mark.parametrize("mesh", ["hand_vh", "mouse_vh"])
def test_closest_hit_bit_exact_vs_brute_force(cuda_device, mesh):
    from drt_b200 import optix
    v, f = load_mesh(mesh)
    ray = _random_rays(v, 20000, 1)
    # grazing / edge-on rays: aim exactly at vertices and edge midpoints
    extra = _random_rays(v, 4000, 2)
    extra[:2000, 3:] = v[np.arange(2000) % len(v)].astype(np.float32) - extra[:2000, :3]
    e = f[np.arange(2000) % len(f)]
    extra[2000:, 3:] = (0.5 * (v[e[:, 0]] + v[e[:, 1]])).astype(np.float32) - extra[2000:, :3]
    ray = np.concatenate([ray, extra], 0)
    m = oracle.OracleMesh(v, f)
    T0, I0 = m.closest_hit(ray, use_bvh=False)
    om = optix.optix_mesh(cuda_device.index or 0)
    om.update_mesh(torch.tensor(f, dtype=torch.int32, device=cuda_device), torch.tensor(v, dtype=torch.float32, device=cuda_device))
    T, I = om.intersect(torch.tensor(ray, device=cuda_device))
    assert T.stride() == (2,) and I.stride() == (2,) and I.dtype == torch.int32  # reference hit-buffer layout
    T, I = T.cpu().numpy(), I.cpu().numpy()
    assert np.array_equal(I, I0), f"{(I != I0).sum()} of {len(I)} ids differ"
    assert np.array_equal(T, T0)
    assert ((T > 0) == (I >= 0)).all() and (I >= 0).mean() > 0.3


@pytest.mark.parametrize("case", CHAIN_CASES)
def test_trace_fwd_bwd_vs_golden_and_oracle(cuda_device, case):
    z = load_chain_case(case)
    R, sc = _scene(z["vertices"], z["faces"], cuda_device)
    R.intIOR = z["int_ior"]
    V = torch.tensor(z["vertices"], dtype=torch.float64, device=cuda_device, requires_grad=True)
    sc.update_verticex(V)
    o = torch.tensor(z["origin"], device=cuda_device)
    d = torch.tensor(z["ray_dir"], device=cuda_device)
    out_ori, out_dir, mask = sc.render_transparent(o, d)
    assert mask.dtype == torch.bool and mask.shape == o.shape
    mk = mask.cpu().numpy()
    assert (mk[:, 0] == mk[:, 1]).all() and (mk[:, 0] == mk[:, 2]).all()
    idx = np.nonzero(mk[:, 0])[0]
    assert np.array_equal(idx, z["valid_idx"])
    oo, od = out_ori.detach().cpu().numpy(), out_dir.detach().cpu().numpy()
    # golden = the unmodified reference (DiffRender.py) on CPU
    assert np.abs(oo[idx] - z["out_ori"]).max() < 1e-10
    assert np.abs(od[idx] - z["out_dir"]).max() < 1e-12
    assert not oo[~mk[:, 0]].any() and not od[~mk[:, 0]].any()
    # oracle: same rounding by construction -> bit-exact
    m = oracle.OracleMesh(z["vertices"], z["faces"])
    q = m.trace_fwd(z["origin"], z["ray_dir"], z["int_ior"])
    assert np.array_equal(oo, q["out_ori"]) and np.array_equal(od, q["out_dir"])
    # backward, both upstream gradients
    L = (out_ori * torch.tensor(z["g_ori"], device=cuda_device)).sum() + (out_dir * torch.tensor(z["g_dir"], device=cuda_device)).sum()
    L.backward()
    g = V.grad.cpu().numpy()
    pv, gl = grad_rel_err(g, z["grad_V"])
    assert pv < 1e-8 and gl < 1e-10, (pv, gl)
    # ray_loss-shaped: out_ori detached -> g_out_ori is None (optim.py:100)
    V.grad = None
    out_ori, out_dir, mask = sc.render_transparent(o, d)
    (out_dir * torch.tensor(z["g_dir"], device=cuda_device)).sum().backward()
    pv, gl = grad_rel_err(V.grad.cpu().numpy(), z["grad_V_dir_only"])
    assert pv < 1e-8 and gl < 1e-10, (pv, gl)


def test_full_view_vs_oracle_hand_c2_subsample(cuda_device):
    """C2 geometry (hand_vh) at 256x256 with every query stage checked: validity, records, outputs."""
    from drt_b200 import views
    v, f = load_mesh("hand_vh")
    R, sc = _scene(v, f, cuda_device)
    cams = views.turntable_cameras(v, 256, 256, 72)
    m = oracle.OracleMesh(v, f)
    for k in (0, 31):
        o, d = views.generate_ray(256, 256, cams[k][3], cams[k][2], device=cuda_device)
        out_ori, out_dir, mask = sc.render_transparent(o, d)
        q = m.trace_fwd(o.cpu().numpy(), d.cpu().numpy(), INT_IOR)
        assert np.array_equal(mask.cpu().numpy(), q["mask"])
        assert np.array_equal(out_ori.cpu().numpy(), q["out_ori"]) and np.array_equal(out_dir.cpu().numpy(), q["out_dir"])
        assert 0.05 < q["mask"][:, 0].mean() < 0.5
        hm = sc.render_mask(o, d).cpu().numpy()
        assert np.array_equal(hm > 0, q["stage"] >= 1)


def test_refit_equals_rebuild_and_perturbed_vertices(cuda_device):
    from drt_b200 import views
    v, f = load_mesh("mouse_vh")
    R, sc = _scene(v, f, cuda_device)
    rng = np.random.default_rng(3)
    v2 = v + rng.normal(scale=0.05, size=v.shape)
    cams = views.turntable_cameras(v, 128, 128, 72)
    o, d = views.generate_ray(128, 128, cams[7][3], cams[7][2], device=cuda_device)
    V2 = torch.tensor(v2, device=cuda_device)
    outs = []
    for refit in (False, True):
        sc.refit = refit
        sc.update_verticex(V2)
        outs.append([t.cpu().numpy() for t in sc.render_transparent(o, d)])
    for a, b in zip(*outs):
        assert np.array_equal(a, b)
    info = sc.optix_mesh.info()
    assert info["refits"] == 1 and info["builds"] == 2 and info["n_faces"] == len(f)
    q = oracle.OracleMesh(v2, f).trace_fwd(o.cpu().numpy(), d.cpu().numpy(), INT_IOR)
    assert np.array_equal(outs[0][2], q["mask"]) and np.array_equal(outs[0][0], q["out_ori"])


def test_edge_cases(cuda_device):
    from drt_b200 import optix, _lib
    dev = cuda_device
    om = optix.optix_mesh(dev.index or 0)
    ray = torch.tensor([[0, 0, 5, 0, 0, -1.0]], dtype=torch.float32, device=dev)
    with pytest.raises(_lib.DrtError):  # reference: assert(builded), optix_extend.cpp:30
        om.intersect(ray)
    with pytest.raises(ValueError):     # reference: assert(Ray.size(1) == 6)
        om.update_mesh(torch.zeros((1, 3), dtype=torch.int32, device=dev), torch.zeros((3, 2), device=dev))
    # empty mesh: everything misses
    om.update_mesh(torch.zeros((0, 3), dtype=torch.int32, device=dev), torch.zeros((0, 3), dtype=torch.float32, device=dev))
    T, I = om.intersect(ray)
    assert T.item() == -1 and I.item() == -1
    # zero rays
    T, I = om.intersect(torch.zeros((0, 6), dtype=torch.float32, device=dev))
    assert T.numel() == 0 and I.numel() == 0
    # single triangle, hit / miss / parallel / behind / on-plane ray
    om.update_mesh(torch.tensor([[0, 1, 2]], dtype=torch.int32, device=dev),
                   torch.tensor([[0, 0, 0], [1, 0, 0], [0, 1, 0]], dtype=torch.float32, device=dev))
    rays = torch.tensor([[0.2, 0.2, 1, 0, 0, -1], [2, 2, 1, 0, 0, -1], [0.2, 0.2, 1, 1, 0, 0], [0.2, 0.2, -1, 0, 0, -1],
                         [0.2, 0.2, 1, 0, 0, -4], [-1, 0.2, 0, 1, 0, 0]], dtype=torch.float32, device=dev)
    T, I = om.intersect(rays)
    assert I.tolist() == [0, -1, -1, -1, 0, -1]
    assert T.tolist()[0] == 1.0 and T.tolist()[4] == 0.25  # t in units of |direction|
    # out-of-range face indices are clamped and counted instead of faulting
    om.update_mesh(torch.tensor([[0, 1, 7]], dtype=torch.int32, device=dev),
                   torch.tensor([[0, 0, 0], [1, 0, 0], [0, 1, 0]], dtype=torch.float32, device=dev))
    assert om.bad_indices() == 1
    # two triangles sharing an edge, ray through the shared edge: lowest id wins
    om.update_mesh(torch.tensor([[0, 1, 2], [2, 1, 3]], dtype=torch.int32, device=dev),
                   torch.tensor([[0, 0, 0], [1, 0, 0], [0, 1, 0], [1, 1, 0]], dtype=torch.float32, device=dev))
    T, I = om.intersect(torch.tensor([[0.5, 0.5, 1, 0, 0, -1], [0.75, 0.75, 2, 0, 0, -1]], dtype=torch.float32, device=dev))
    assert I.tolist() == [0, 1]


def test_all_rays_miss_and_zero_rays_through_scene(cuda_device):
    v, f = load_mesh("hand_vh")
    R, sc = _scene(v, f, cuda_device)
    V = sc.vertices.clone().requires_grad_(True)
    sc.update_verticex(V)
    o = torch.tensor([[1e3, 1e3, 1e3]] * 8, dtype=torch.float64, device=cuda_device)
    d = torch.tensor([[0, 0, 1.0]] * 8, dtype=torch.float64, device=cuda_device)
    out_ori, out_dir, mask = sc.render_transparent(o, d)
    assert not mask.any() and not out_ori.any() and not out_dir.any()
    (out_dir.sum() + out_ori.sum()).backward()
    assert not V.grad.any()
    out_ori, out_dir, mask = sc.render_transparent(o[:0], d[:0])
    assert out_ori.shape == (0, 3) and mask.shape == (0, 3)


def test_ray_loss_grad_kernel_matches_torch(cuda_device):
    """drt_ray_loss_grad against the reference's ray_loss expression (optim.py:99-106) in torch."""
    import ctypes as C
    from drt_b200 import _lib, views
    v, f = load_mesh("hand_vh")
    R, sc = _scene(v, f, cuda_device)
    cams = views.turntable_cameras(v, 96, 96, 72)
    o, d = views.generate_ray(96, 96, cams[3][3], cams[3][2], device=cuda_device)
    out_ori, out_dir, mask = sc.render_transparent(o, d)
    g = torch.Generator(device="cpu").manual_seed(0)
    screen = (out_ori + out_dir * 50 + torch.randn(o.shape, generator=g, dtype=torch.float64).to(cuda_device)).contiguous()
    valid = (torch.rand(len(o), generator=g) > 0.2).to(cuda_device)
    od = out_dir.clone().requires_grad_(True)
    target = screen - out_ori.detach()
    target = target / target.norm(dim=1, keepdim=True)
    vm = valid * mask[:, 0]
    loss = (od - target)[vm].pow(2).sum()
    loss.backward()
    g_dir = torch.empty_like(out_dir)
    lsum = torch.zeros(1, dtype=torch.float64, device=cuda_device)
    p = lambda t: C.c_void_p(t.data_ptr())  # noqa: E731
    _lib.call("drt_ray_loss_grad", p(out_ori), p(out_dir), p(mask), p(screen), p(valid), len(o), p(g_dir), p(lsum),
              C.c_void_p(torch.cuda.current_stream().cuda_stream))
    assert torch.allclose(g_dir, od.grad, rtol=1e-13, atol=1e-15)
    assert abs(lsum.item() - loss.item()) <= 1e-12 * max(1.0, abs(loss.item()))


def test_wavefront_path_large_batch_vs_oracle(cuda_device):
    """> 2^20 rays takes the five-kernel wavefront (smaller batches take the one-launch kernel):
    1.32 M rays of mouse_vh, masks and outputs bit-exact against the oracle, gradients to 1e-10."""
    from drt_b200 import views
    v, f = load_mesh("mouse_vh")
    R, sc = _scene(v, f, cuda_device)
    V = sc.vertices.clone().requires_grad_(True)
    sc.update_verticex(V)
    cams = views.turntable_cameras(v, 1100, 1200, 72)
    o, d = views.generate_ray(1100, 1200, cams[50][3], cams[50][2], device=cuda_device)
    assert len(o) > (1 << 20)
    out_ori, out_dir, mask = sc.render_transparent(o, d)
    rng = np.random.default_rng(5)
    g_ori, g_dir = rng.standard_normal(o.shape), rng.standard_normal(o.shape)
    ((out_ori * torch.tensor(g_ori, device=cuda_device)).sum() + (out_dir * torch.tensor(g_dir, device=cuda_device)).sum()).backward()
    m = oracle.OracleMesh(v, f)
    on, dn = o.cpu().numpy(), d.cpu().numpy()
    q = m.trace_fwd(on, dn, INT_IOR)
    assert np.array_equal(mask.cpu().numpy(), q["mask"])
    assert np.array_equal(out_ori.detach().cpu().numpy(), q["out_ori"]) and np.array_equal(out_dir.detach().cpu().numpy(), q["out_dir"])
    gq = m.trace_bwd(on, dn, q["tri1"], q["tri2"], g_ori, g_dir, INT_IOR)
    pv, gl = grad_rel_err(V.grad.cpu().numpy(), gq)
    assert pv < 1e-9 and gl < 1e-11, (pv, gl)
    hm = sc.render_mask(o, d).cpu().numpy()
    assert np.array_equal(hm > 0, q["stage"] >= 1)


@pytest.mark.parametrize("kernel", ["wavefront", "simple", "onelaunch", "bwd-merge", "bwd-nomerge"])
def test_every_forward_kernel_variant_passes_the_parity_suite(kernel):
    """The kernel choice is by batch size; force each variant (five-launch wavefront, one-thread-per-path
    megakernel, cooperative single-launch wavefront) over the whole golden/oracle suite."""
    import os
    import subprocess
    import sys
    env = dict(os.environ, DRT_ONE_LAUNCH="1" if kernel == "onelaunch" else "0")
    if kernel in ("wavefront", "simple", "onelaunch"):
        env["DRT_FWD_KERNEL"] = "wavefront" if kernel == "onelaunch" else kernel
    else:  # both variants of the backward scatter (the default picks one by rays per vertex)
        env["DRT_BWD_MERGE"] = "1" if kernel == "bwd-merge" else "0"
    if os.environ.get("DRT_PARITY_CHILD"):
        pytest.skip("already inside the forced-kernel child run")
    env["DRT_PARITY_CHILD"] = "1"
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.abspath(__file__), "-q", "-x", "-m", "gpu", "-k",
                        "golden or full_view or refit or all_rays_miss or large_batch"], env=env, capture_output=True, text=True)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]


def test_fused_ray_loss_matches_reference_expression(cuda_device):
    """drt_b200.losses.ray_loss against the reference's ray_loss expression (optim.py:96-106) evaluated
    with torch autograd on top of render_transparent."""
    from drt_b200 import losses, views
    v, f = load_mesh("hand_vh")
    R, sc = _scene(v, f, cuda_device)
    cams = views.turntable_cameras(v, 128, 128, 72)
    o, d = views.generate_ray(128, 128, cams[20][3], cams[20][2], device=cuda_device)
    g = torch.Generator(device="cpu").manual_seed(1)
    with torch.no_grad():
        oo, od, mk = sc.render_transparent(o, d)
    screen = (oo + od * 80 + 0.5 * torch.randn(o.shape, generator=g, dtype=torch.float64).to(cuda_device)).contiguous()
    valid = (torch.rand(len(o), generator=g) > 0.1).to(cuda_device)
    Va = sc.vertices.clone().requires_grad_(True)
    sc.update_verticex(Va)
    out_ori, out_dir, mask = sc.render_transparent(o, d)
    target = screen - out_ori.detach()
    target = target / target.norm(dim=1, keepdim=True)
    ref = ((out_dir - target)[valid * mask[:, 0]]).pow(2).sum()
    (3.0 * ref).backward()
    Vb = sc.vertices.detach().clone().requires_grad_(True)
    sc.update_verticex(Vb)
    mine = losses.ray_loss(sc, o, d, screen, valid)
    (3.0 * mine).backward()
    assert abs(mine.item() - ref.item()) <= 1e-12 * abs(ref.item())
    pv, gl = grad_rel_err(Vb.grad.cpu().numpy(), Va.grad.cpu().numpy())
    assert pv < 1e-10 and gl < 1e-12, (pv, gl)


@pytest.mark.parametrize("res", [(96, 128), (1100, 1200)])
def test_image_size_hint_changes_nothing_but_the_batching(cuda_device, res):
    """Render.resy / resx (DiffRender.py:16-17, optim.py:179-180) let the entry query walk 32-pixel tiles (4x8, else 8x4): outputs,
    masks and gradients are those of the scanline batching (small batch: one-launch kernel, large: wavefront)."""
    from drt_b200 import views
    v, f = load_mesh("mouse_vh")
    R, sc = _scene(v, f, cuda_device)
    cam = views.turntable_cameras(v, res[0], res[1], 72)[31]
    o, d = views.generate_ray(res[0], res[1], cam[3], cam[2], device=cuda_device)
    g = torch.randn(o.shape, dtype=torch.float64, generator=torch.Generator().manual_seed(2)).to(cuda_device)
    outs = []
    old = (R.resy, R.resx)
    try:
        for hint in ((res[0] + 1, res[1]), res):       # first an untileable size (ignored), then the real one
            R.resy, R.resx = hint
            V = sc.vertices.detach().clone().requires_grad_(True)
            sc.update_verticex(V)
            oo, od, mk = sc.render_transparent(o, d)
            (od * g).sum().backward()
            outs.append((oo.detach().clone(), od.detach().clone(), mk.clone(), V.grad.clone(), sc.render_mask(o, d)))
    finally:
        R.resy, R.resx = old
    a, b = outs
    assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1]) and torch.equal(a[2], b[2]) and torch.equal(a[4], b[4])
    assert a[2].any()
    pv, gl = grad_rel_err(b[3].cpu().numpy(), a[3].cpu().numpy())
    assert pv < 1e-10 and gl < 1e-12, (pv, gl)
