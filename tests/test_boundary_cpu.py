"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports every symbol
include/drt_b200.h declares; the host mirror fails loudly without a GPU; PLY / mesh helpers."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "drt_b200.h")).read()
    return sorted(set(re.findall(r"DRT_API[^;(]*?\b(drt_\w+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from drt_b200 import _lib, build
    so = build.build()
    lib = ctypes.CDLL(so)
    decl = _declared_symbols()
    assert len(decl) >= 13
    for s in decl:
        assert hasattr(lib, s), f"{s} declared in include/drt_b200.h but not exported"
    assert sorted(_lib.SIGNATURES) == decl, "drt_b200/_lib.py binds a different set than the header declares"
    lib.drt_version.restype = ctypes.c_int
    assert lib.drt_version() >= 1000


def test_header_cites_reference_for_each_entry_point():
    src = open(os.path.join(ROOT, "include", "drt_b200.h")).read()
    for fn in ("drt_bvh_create", "drt_bvh_build", "drt_bvh_update_vert", "drt_bvh_set_image_size", "drt_closest_hit", "drt_trace_fwd",
               "drt_trace_bwd", "drt_ray_loss_grad", "drt_ray_loss_step", "drt_ray_loss_step_beams", "drt_tile_beams", "drt_generate_rays",
               "drt_comm_create", "drt_trace_fwd_smooth", "drt_trace_bwd_smooth", "drt_plane_hit", "drt_plane_hit_bwd", "drt_tuning_set"):
        i = src.index(f" {fn}(")
        assert re.search(r"(optix_extend\.cpp|DiffRender\.py|optim\.py|captured_data\.py):\d+", src[max(0, i - 3500):i]), fn


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_no_silent_cpu_fallback():
    import drt_b200.DiffRender as R
    from drt_b200 import _lib
    with pytest.raises(_lib.DrtError):
        R.Scene(vertices=np.zeros((3, 3)), faces=np.array([[0, 1, 2]]))


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "drt_b200")
    for dp, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dp, fn)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, re.M), f"{fn} imports the oracle"
                assert "liboracle" not in txt and "oracle.py" not in txt, fn


def test_ply_roundtrip_and_mesh_tables(tmp_path):
    from drt_b200 import meshgen, plyio, trimesh_lite
    v, f = meshgen.icosahedron()
    p = str(tmp_path / "i.ply")
    plyio.write_ply(p, v, f)
    m = trimesh_lite.load(p)
    assert m.is_watertight and m.faces.shape == (20, 3) and np.abs(m.vertices - v).max() < 1e-6
    assert m.edges.shape == (60, 2) and len(np.unique(m.edges_sorted, axis=0)) == 30
    assert all(len(n) == 5 for n in m.vertex_neighbors)
    pairs = trimesh_lite.group_rows_pairs(m.edges_sorted)
    assert pairs.shape == (30, 2)
    assert (m.edges_sorted[pairs[:, 0]] == m.edges_sorted[pairs[:, 1]]).all()
    assert not meshgen.is_watertight(f[:-1])
    v2, f2 = meshgen.subdivide(v, f)
    assert f2.shape == (80, 3) and meshgen.is_watertight(f2)


def test_generate_ray_convention():
    """captured_data.py:23-40: integer pixel coordinates, origin = camera centre, unit directions;
    the principal ray goes through the object centre."""
    from drt_b200 import meshgen, views
    v, _ = meshgen.icosahedron()
    cams = views.turntable_cameras(v, 64, 64, 72)
    R, K, R_inv, K_inv = cams[9]
    o, d = views.generate_ray(64, 64, K_inv, R_inv)
    assert o.shape == (4096, 3) and torch.allclose(d.norm(dim=1), torch.ones(4096, dtype=torch.float64))
    assert torch.equal(o[0], o[-1])
    ctr = 0.5 * (v.min(0) + v.max(0))
    c = d[32 * 64 + 32].numpy()
    want = (ctr - o[0].numpy()) / np.linalg.norm(ctr - o[0].numpy())
    assert np.allclose(c, want, atol=1e-12)
    assert np.allclose(R @ R_inv, np.eye(4), atol=1e-12) and np.allclose(K @ K_inv, np.eye(3), atol=1e-12)
    # pixel (x, y) -> row-major index y*resx + x; x grows to the right of the image
    px = (K @ (R[:3, :3] @ (o[0].numpy() + 5 * d[10].numpy()) + R[:3, 3]))
    assert np.allclose(px[:2] / px[2], [10, 0], atol=1e-9)


def test_captured_data_loader_schema_and_soft_mask(tmp_path):
    """N4: the captured-set schema (captured_data.py:94-108) read from an .npz stand-in; process_mask against
    cv2's distance transform, which is what the reference calls (captured_data.py:12-20)."""
    import cv2
    from drt_b200 import captured_data as cd, meshgen, views
    rng = np.random.default_rng(0)
    M = np.zeros((40, 50), np.uint8)
    cv2.circle(M, (25, 18), 11, 255, -1)
    M[30:36, 5:20] = 255
    m1 = M.copy() // 255
    ref = (cv2.distanceTransform(m1, cv2.DIST_L2, 0) - 0).clip(0, 1) - (cv2.distanceTransform(1 - m1, cv2.DIST_L2, 0) - 1).clip(0, 1)
    ref = (ref + 1) / 2
    ref[-1] = 0.5
    got = cd.process_mask(M)
    assert got.shape == M.shape and np.abs(got - ref).max() < 1e-5 and got.min() == 0 and got.max() == 1
    # a 3-view pinhole set, 12x16 pixels
    v, _ = meshgen.icosahedron()
    cams = views.turntable_cameras(v, 12, 16, 3)
    screen = rng.normal(size=(3, 12 * 16, 3))
    screen[:, ::5] = 0
    masks = (rng.uniform(size=(3, 12, 16)) > 0.5).astype(np.uint8)
    masks[:, 0, 0] = 1
    p = str(tmp_path / "set.npz")
    np.savez(p, cam_proj=np.stack([c[0] for c in cams]), cam_k=cams[0][1], screen_position=screen, mask=masks)
    d = cd.Data_Redmi({"num_view": 3, "name": "horse"}, path=p, res=(12, 16))
    d.device = "cpu"
    assert len(d.Views) == 3
    scr, valid, mask, origin, ray_dir, cam = d.get_view(1)
    assert scr.shape == (192, 3) and valid.dtype == torch.bool and not valid[::5].any() and valid[1]
    o_ref, d_ref = views.generate_ray(12, 16, cams[1][3], cams[1][2])
    assert torch.allclose(ray_dir, d_ref) and torch.allclose(origin, o_ref) and mask.shape == (12, 16)
    assert torch.allclose(cam[0] @ cam[2], torch.eye(4, dtype=torch.float64), atol=1e-12)


def test_compact_view_and_origin_layout_host_logic():
    """CompactView / SparseTargets / origin_rows (the lossless compact host format of the fused ray-loss step)."""
    import torch
    from drt_b200 import losses, views
    from drt_b200.captured_data import CompactView
    cam = views.turntable_cameras(np.array([[-1.0, -1, -1], [1, 1, 1]]), 6, 8, 72)[5]
    o, d = views.generate_ray(6, 8, cam[3], cam[2])
    screen = torch.zeros_like(d)
    screen[[3, 17, 40]] = torch.tensor([[1.0, 2, 3], [4, 5, 6], [7, 8, 9]], dtype=torch.float64)
    valid = screen[:, 0] != 0
    cv = CompactView.from_reference_view((screen, valid, None, o, d, None))
    assert cv.origin.shape == (1, 3) and torch.equal(cv.origin[0], o[0])
    assert cv.targets.idx.tolist() == [3, 17, 40] and cv.targets.idx.dtype == torch.int32
    assert torch.equal(cv.targets.xyz, screen[[3, 17, 40]])
    assert cv.h2d_bytes() == 24 + 48 * 24 + 3 * 28
    both = CompactView.concat([cv, cv])
    assert both.origin.shape == (2, 3) and both.ray_dir.shape == (96, 3)
    assert both.targets.idx.tolist() == [3, 17, 40, 51, 65, 88] and losses.origin_rows(both.origin, 96)[1] == 48
    o2 = o.clone()
    o2[5, 0] += 1e-9                                      # calibrated per-pixel origins stay per ray
    assert CompactView.from_reference_view((screen, valid, None, o2, d, None)).origin.shape == (48, 3)
    assert losses.origin_rows(o, 48)[1] == 1
    assert losses.origin_rows(o[:1].expand(48, 3), 48)[1] == 48
    assert losses.origin_rows(o[:4], 48)[1] == 12 and losses.origin_rows(o[0], 48)[1] == 48
    with pytest.raises(ValueError):
        losses.origin_rows(o[:5], 48)
    with pytest.raises(TypeError):
        losses.SparseTargets(torch.zeros(3, dtype=torch.int64), torch.zeros((3, 3), dtype=torch.float64))
