"""Generates tests/golden/ from the UNMODIFIED reference (build container only; needs /root/reference).

    python -m oracle.make_golden

TEST INFRASTRUCTURE ONLY.  Outputs (all small, committed):
  meshes/<name>.npz    float32 vertices + int32 faces of the reference's data/<name>.ply, the inputs
                       BASELINE.json's configs name (C2 hand_vh, C3 mouse_vh, C4 horse_vh)
  kat_functions.npz    Refract / FrDielectric / JIT_Dintersect known answers (DiffRender.py:35-121)
  chain_<case>.npz     Scene.render_transparent forward + autograd backward (DiffRender.py:420-432,
                       optim.py:210) on seeded rays, with the brute-force stand-in intersector
"""
import os
import sys
import warnings

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from drt_b200 import meshgen, plyio, views  # noqa: E402
from oracle import ref_harness  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
INT_IOR = 1.4723  # config.py:22


def save_meshes():
    os.makedirs(os.path.join(GOLD, "meshes"), exist_ok=True)
    for name in ("hand_vh", "mouse_vh", "horse_vh"):
        v, f = plyio.read_ply(os.path.join(ref_harness.REF_ROOT, "data", name + ".ply"))
        np.savez_compressed(os.path.join(GOLD, "meshes", name + ".npz"), vertices=v.astype(np.float32),
                            faces=f.astype(np.int32))


def kat_functions():
    R = ref_harness.load_reference()
    t = lambda a: torch.tensor(np.asarray(a, dtype=np.float64))  # noqa: E731
    out = {}
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        th = np.radians([0.0, 10.0, 40.0, 80.0, 89.9])
        wo = np.stack([np.sin(th), 0 * th, np.cos(th)], 1)
        n = np.tile([0.0, 0.0, 1.0], (len(th), 1))
        tir, wt = R.Refract(t(wo), t(n), t(np.full(len(th), 1.00029 / 1.5)))
        out.update(refract_in_wo=wo, refract_in_eta=1.00029 / 1.5, refract_in_tir=tir.numpy(), refract_in_wt=wt.numpy())
        th = np.radians([10.0, 30.0, 41.0, 41.9, 45.0])
        wo = np.stack([np.sin(th), 0 * th, np.cos(th)], 1)
        tir, wt = R.Refract(t(wo), t(n), t(np.full(len(th), 1.5 / 1.00029)))
        ftir, fr = R.FrDielectric(t(np.cos(th)), t(np.full(len(th), 1.5)), t(np.full(len(th), 1.00029)))
        out.update(refract_out_wo=wo, refract_out_eta=1.5 / 1.00029, refract_out_tir=tir.numpy(),
                   refract_out_wt=wt.numpy(), fr_cos=np.cos(th), fr_tir=ftir.numpy(), fr_R=fr.numpy())
        rng = np.random.default_rng(7)
        o = rng.normal(size=(16, 3)) + [0, 0, 5]
        d = rng.normal(size=(16, 3)) * 0.2 + [0, 0, -1]
        d /= np.linalg.norm(d, axis=1, keepdims=True)
        tri = rng.normal(size=(16, 3, 3)) * 2
        o[0], d[0] = (0.5, 0.7, 5), np.array([0.1, -0.2, -1]) / np.linalg.norm([0.1, -0.2, -1])
        tri[0] = [(0, 0, 0), (2, 0, 0.5), (0, 3, 0.25)]
        u, v, tt, nn = R.JIT_Dintersect(t(o), t(d), t(tri), t(np.zeros((16, 3, 3))))
        out.update(di_o=o, di_d=d, di_tri=tri, di_u=u.numpy(), di_v=v.numpy(), di_t=tt.numpy(), di_n=nn.numpy())
    np.savez_compressed(os.path.join(GOLD, "kat_functions.npz"), **out)


def upstream(seed, shape):
    """Seeded upstream gradients; regenerated (not stored) by the tests, guarded by a checksum."""
    rng = np.random.default_rng(seed)
    return rng.standard_normal(shape), rng.standard_normal(shape)


def chain_case(case, v, f, o, d, int_ior, seed, cam=None):
    g_ori, g_dir = upstream(seed, o.shape)
    r = ref_harness.render_transparent(v, f, o, d, int_ior, g_ori, g_dir)
    idx = np.nonzero(r["mask"])[0]
    assert len(idx) >= 4, (case, len(idx))
    # zeros everywhere else is part of the contract (DiffRender.py:421-423)
    inv = np.ones(len(o), bool)
    inv[idx] = False
    assert not r["out_ori"][inv].any() and not r["out_dir"][inv].any()
    # a second backward with the ray_loss-shaped upstream gradient: grad_out_ori = 0 (optim.py:100)
    r2 = ref_harness.render_transparent(v, f, o, d, int_ior, None, g_dir)
    np.savez_compressed(
        os.path.join(GOLD, f"chain_{case}.npz"), vertices=v, faces=f.astype(np.int32), int_ior=int_ior, seed=seed,
        # rays: stored explicitly for hand-made cases, else as the camera they are generated from
        **(dict(origin=o, ray_dir=d) if cam is None else dict(cam_R_inv=cam[0], cam_K_inv=cam[1], res=np.array(cam[2]))),
        ray_checksum=np.array([o.sum(), d.sum(), g_ori.sum(), g_dir.sum()]),
        valid_idx=idx.astype(np.int64), out_ori=r["out_ori"][idx], out_dir=r["out_dir"][idx],
        grad_V=r["grad_V"], grad_V_dir_only=r2["grad_V"])
    print(f"{case}: {len(o)} rays, {len(idx)} valid, |grad_V|max {np.abs(r['grad_V']).max():.3g}")


def silhouette_case():
    """N1: the reference's silhouette_edge + primary_visibility + primary_edge_sample (DiffRender.py:445-479,
    189-267) on hand_vh, with edge tables from drt_b200.trimesh_lite (trimesh itself is absent) and the
    brute-force stand-in intersector."""
    from drt_b200 import silhouette, trimesh_lite
    R = ref_harness.load_reference()
    v, f = plyio.read_ply(os.path.join(ref_harness.REF_ROOT, "data", "hand_vh.ply"))
    s = ref_harness.make_scene(v, f, INT_IOR)
    mesh = trimesh_lite.TriMesh(v, f)
    s.Edges, s.E2F, _ = silhouette.build_edge_tables(mesh, s.faces, "cpu")
    resy, resx = 240, 320
    R.resy, R.resx = resy, resx
    cams = views.turntable_cameras(v, resy, resx, 72)
    Rm, K, R_inv, K_inv = (torch.tensor(m) for m in cams[11])
    origin = R_inv[:3, 3].clone()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        sil = s.silhouette_edge(origin)
        index, output = s.primary_visibility(sil, (Rm, K, R_inv, K_inv), origin, detach_depth=True)
        w = torch.tensor(np.random.default_rng(9).standard_normal(len(output)), dtype=output.dtype)
        (output * w).sum().backward()
        dih = R.dot(*R.edge_face_norm(s.vertices.detach(), s.E2F)).numpy()
    np.savez_compressed(os.path.join(GOLD, "silhouette_hand_vh.npz"), view=11, res=np.array([resy, resx]),
                        sil_edges=sil.numpy(), index=index.numpy(), output=output.detach().numpy(), weights=w.numpy(),
                        grad_V=s.vertices.grad.numpy(), dihedral_cos=dih)
    print(f"silhouette: {len(sil)} silhouette edges, {len(index)} samples in view, |grad|max {s.vertices.grad.abs().max():.3g}")


def vertex_normals_case():
    """Scene.init_VN (DiffRender.py:319-336) of the unmodified reference on hand_vh: the vertex normals and the gradient of a
    seeded linear functional of them w.r.t. the vertices -- pins oracle/chain_torch.vertex_normals and
    drt_b200.DiffRender.vertex_normals (the input of the optional smooth-normal mode)."""
    R = ref_harness.load_reference()
    v, f = plyio.read_ply(os.path.join(ref_harness.REF_ROOT, "data", "hand_vh.ply"))
    s = ref_harness.make_scene(v, f, INT_IOR)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        s.init_VN()
        w = torch.tensor(np.random.default_rng(21).standard_normal(s.normals.shape))
        (s.normals * w).sum().backward()
    np.savez_compressed(os.path.join(GOLD, "vertex_normals_hand_vh.npz"), normals=s.normals.detach().numpy(), weights=w.numpy(),
                        grad_V=s.vertices.grad.numpy())
    print(f"vertex normals: {tuple(s.normals.shape)}, |grad|max {s.vertices.grad.abs().max():.3g}")


def view_rays(v, resy, resx, k, n_views=72):
    cams = views.turntable_cameras(v, resy, resx, n_views)
    _, _, R_inv, K_inv = cams[k]
    o, d = views.generate_ray(resy, resx, K_inv, R_inv)
    return o.numpy(), d.numpy(), (R_inv, K_inv, (resy, resx))


def main():
    if not ref_harness.available():
        sys.exit("reference tree not present; golden vectors can only be regenerated in the build container")
    save_meshes()
    kat_functions()
    silhouette_case()
    vertex_normals_case()
    # App. B tetrahedron (5 rays, one of them a miss)
    v, f = meshgen.tetrahedron()
    o = np.array([(0.6, 0.7, 9), (0.9, 0.5, 9), (0.4, 1.1, 9), (1.2, 0.3, 9), (3.9, 3.9, 9)], float)
    d = np.tile(np.array([0.02, 0.01, -1.0]) / np.linalg.norm([0.02, 0.01, -1.0]), (5, 1))
    chain_case("tetra", v, f, o, d, 1.5, 1)
    # C1: icosahedron, one 64x64 view
    v, f = meshgen.icosahedron()
    o, d, cam = view_rays(v, 64, 64, 5)
    chain_case("icosa_c1", v, f, o, d, INT_IOR, 2, cam)
    # C2-like: hand_vh, 160x160 view (brute force over 4 390 triangles keeps this to seconds)
    v, f = plyio.read_ply(os.path.join(ref_harness.REF_ROOT, "data", "hand_vh.ply"))
    o, d, cam = view_rays(v, 160, 160, 5)
    chain_case("hand_vh_160", v, f, o, d, INT_IOR, 3, cam)
    # C3-like: mouse_vh, 128x96 view, another azimuth
    v, f = plyio.read_ply(os.path.join(ref_harness.REF_ROOT, "data", "mouse_vh.ply"))
    o, d, cam = view_rays(v, 96, 128, 23)
    chain_case("mouse_vh_96x128", v, f, o, d, INT_IOR, 4, cam)
    # vertices that are NOT fp32-representable (after an optimiser step: optim.py:202-203)
    v, f = plyio.read_ply(os.path.join(ref_harness.REF_ROOT, "data", "hand_vh.ply"))
    v = v + np.random.default_rng(5).normal(scale=0.05, size=v.shape)
    o, d, cam = view_rays(v, 96, 96, 40)
    chain_case("hand_vh_perturbed_96", v, f, o, d, INT_IOR, 5, cam)


if __name__ == "__main__":
    main()
