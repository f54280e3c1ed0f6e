"""Scratch: fwd/bwd timing for one config under the current DRT_FWD_* env."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import drt_b200.DiffRender as R
from drt_b200 import configs, views

def run(name, nviews, reps=5):
    dev = torch.device("cuda:0")
    cfg = configs.make(name)
    R.intIOR = configs.INT_IOR
    sc = R.Scene(vertices=cfg["vertices"], faces=cfg["faces"])
    V = sc.vertices.clone().requires_grad_(True)
    sc.update_verticex(V)
    cams = cfg["cams"][:: max(1, len(cfg["cams"]) // nviews)][:nviews]
    o, d = views.view_batch(cams, cfg["resy"], cfg["resx"], device=dev)
    n = len(o)
    tf = tw = 0.0
    for r in range(reps + 2):
        V.grad = None
        e = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        e[0].record()
        out_ori, out_dir, mask = sc.render_transparent(o, d)
        e[1].record()
        g = torch.ones_like(out_dir)
        e[2].record()
        out_dir.backward(g)
        e[3].record()
        torch.cuda.synchronize()
        if r >= 2:
            tf += e[0].elapsed_time(e[1]); tw += e[2].elapsed_time(e[3])
    tf /= reps; tw /= reps
    print(f"{os.environ.get('DRT_FWD_KERNEL','persistent')}/{os.environ.get('DRT_FWD_THRESH','-')} {name} x{nviews}: {n} rays valid {mask[:,0].float().mean().item():.3f} "
          f"fwd {tf:.3f} ms ({n/tf/1e6:.2f} Grays/s)  bwd {tw:.3f} ms | checksum {out_dir.sum().item():.12g} {V.grad.abs().sum().item():.12g}", flush=True)

if __name__ == "__main__":
    for a in sys.argv[1:]:
        name, nv = a.split(":")
        run(name, int(nv))
