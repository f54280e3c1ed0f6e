"""Inputs of tools/warp_sim: the benchmark mesh and the ray lists of the three query stages of a few views, in
the order the kernels see them (Q1: scanline; Q2/Q3: the survivors in Q1 order).  The stage-to-stage refraction
is the chain of DiffRender.py:503-535 restated in numpy (developer tool; no parity claim)."""
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from drt_b200 import configs, views  # noqa: E402

out = sys.argv[1] if len(sys.argv) > 1 else "/tmp/wsim"
cfgname = sys.argv[2] if len(sys.argv) > 2 else "C4"
view_ids = [int(x) for x in sys.argv[3].split(",")] if len(sys.argv) > 3 else [0, 23, 47]
SIM = os.environ.get("WARP_SIM", "/tmp/warp_sim")
os.makedirs(out, exist_ok=True)
cfg = configs.make(cfgname)
V = np.asarray(cfg["vertices"], dtype=np.float64)
F = np.asarray(cfg["faces"], dtype=np.int32)
with open(f"{out}/mesh.bin", "wb") as f:
    np.array([len(V), len(F)], np.int32).tofile(f)
    V.astype(np.float32).tofile(f)
    F.tofile(f)


def write_rays(path, o, d):
    with open(path, "wb") as f:
        np.array([len(o), 0], np.int32).tofile(f)
        o.astype(np.float32).tofile(f)
        d.astype(np.float32).tofile(f)


def hits(rays, mode="closest"):
    env = dict(os.environ, SIM_HITS_ONLY="1")
    subprocess.check_call([SIM, f"{out}/mesh.bin", rays, mode, f"{out}/hits.bin"], env=env, stdout=subprocess.DEVNULL)
    n = np.fromfile(rays, np.int32, 1)[0]
    raw = open(f"{out}/hits.bin", "rb").read()
    return np.frombuffer(raw[:4 * n], np.int32).copy(), np.frombuffer(raw[4 * n:], np.float64).copy()


def refract(o, d, tri, ext=1.00029, inn=configs.INT_IOR):
    a0, a1, a2 = V[F[tri, 0]], V[F[tri, 1]], V[F[tri, 2]]
    Nn = np.cross(a1 - a0, a2 - a0)
    n = Nn / np.linalg.norm(Nn, axis=1, keepdims=True)
    t = ((a0 - o) * Nn).sum(1) / (d * Nn).sum(1)
    c0 = -(d * n).sum(1)
    ent = c0 > 0
    npr = np.where(ent[:, None], n, -n)
    c = np.abs(c0)
    etaI, etaT = np.where(ent, ext, inn), np.where(ent, inn, ext)
    tir = np.sqrt(np.clip(1 - c * c, 0, 1)) * etaI / etaT >= 1
    eta = etaI / etaT
    w = eta[:, None] * d + ((eta * c - c)[:, None]) * npr     # tan-law Refract (DiffRender.py:39-47)
    wt = w / np.linalg.norm(w, axis=1, keepdims=True)
    x = o + t[:, None] * d
    return x + 1e-5 * wt, wt, ~tir


TILE = os.environ.get("TILE")   # e.g. "8x4": reorder every view into tiles of 32 pixels (x fastest inside the tile)


def tile_order(resy, resx, tw, th):
    ys, xs = np.meshgrid(np.arange(resy), np.arange(resx), indexing="ij")
    key = ((ys // th) * (resx // tw) + xs // tw) * (tw * th) + (ys % th) * tw + xs % tw
    return np.argsort(key.reshape(-1), kind="stable")


o = np.concatenate([views.generate_ray(cfg["resy"], cfg["resx"], cfg["cams"][k][3], cfg["cams"][k][2])[0].numpy() for k in view_ids])
d = np.concatenate([views.generate_ray(cfg["resy"], cfg["resx"], cfg["cams"][k][3], cfg["cams"][k][2])[1].numpy() for k in view_ids])
if TILE:
    tw, th = (int(x) for x in TILE.split("x"))
    n_pix = cfg["resy"] * cfg["resx"]
    perm = np.concatenate([k * n_pix + tile_order(cfg["resy"], cfg["resx"], tw, th) for k in range(len(view_ids))])
    o, d = o[perm], d[perm]
write_rays(f"{out}/q1.bin", o, d)
id1, _ = hits(f"{out}/q1.bin")
h = id1 >= 0
o1, d1, ok1 = refract(o[h], d[h], id1[h])
o1, d1 = o1[ok1], d1[ok1]
write_rays(f"{out}/q2.bin", o1, d1)
id2, _ = hits(f"{out}/q2.bin")
h2 = id2 >= 0
o2, d2, ok2 = refract(o1[h2], d1[h2], id2[h2])
write_rays(f"{out}/q3.bin", o2[ok2], d2[ok2])
print(f"{cfgname}: {len(o)} primary rays, {h.sum()} hit, {len(o1)} refracted, {h2.sum()} exit hits, {ok2.sum()} to the occlusion query")
