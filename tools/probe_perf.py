"""Scratch timing probe (not the bench): fwd/bwd kernel times on a few configs."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import drt_b200.DiffRender as R
from drt_b200 import meshgen, views

def load(name):
    z = np.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "meshes", name + ".npz"))
    return z["vertices"].astype(np.float64), z["faces"].astype(np.int64)

def run(tag, v, f, resy, resx, nviews, reps=5):
    dev = torch.device("cuda:0")
    R.intIOR = 1.4723
    sc = R.Scene(vertices=v, faces=f)
    V = sc.vertices.clone().requires_grad_(True)
    cams = views.turntable_cameras(v, resy, resx, 72)[:nviews]
    o, d = views.view_batch(cams, resy, resx, device=dev)
    n = len(o)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    tb = tf = tw = 0
    for r in range(reps + 2):
        V.grad = None
        ev[0].record()
        sc.update_verticex(V)
        ev[1].record()
        out_ori, out_dir, mask = sc.render_transparent(o, d)
        ev[2].record()
        g = torch.ones_like(out_dir)
        torch.cuda.synchronize()
        e4 = torch.cuda.Event(enable_timing=True); e5 = torch.cuda.Event(enable_timing=True)
        e4.record()
        out_dir.backward(g)
        e5.record()
        torch.cuda.synchronize()
        if r >= 2:
            tb += ev[0].elapsed_time(ev[1]); tf += ev[1].elapsed_time(ev[2]); tw += e4.elapsed_time(e5)
    tb, tf, tw = tb / reps, tf / reps, tw / reps
    cov = mask[:, 0].float().mean().item()
    print(f"{tag}: {len(f)} tris, {n} rays, valid {cov:.3f} | build {tb:.3f} ms  fwd {tf:.3f} ms  bwd {tw:.3f} ms | "
          f"{n / (tf + tw) / 1e6:.1f} Mrays/s fwd+bwd, fwd only {n / tf / 1e6:.1f} Mrays/s", flush=True)

if __name__ == "__main__":
    v, f = load("hand_vh"); run("C2 hand 512x512", v, f, 512, 512, 1)
    v, f = load("mouse_vh"); run("C3 mouse 960x720 x8", v, f, 720, 960, 8)
    v, f = load("horse_vh"); v, f = meshgen.subdivide(v, f, 0.05); run("C4 horse50k 960x720 x8", v, f, 720, 960, 8)
    v, f = meshgen.displaced_torus(); run("C5 torus200k 1920x1080 x2", v, f, 1080, 1920, 2)
