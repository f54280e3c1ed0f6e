"""Freezes the roofline denominator: canonical-LBVH node/triangle visit counts per primary ray,
measured by the CPU oracle on a stratified sample (>= 64k rays) of each config's own rays
(SURVEY.md 8(d)).  Writes profiles/canonical_counters.json.   python tools/freeze_counters.py [C2 C3 ...]"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from drt_b200 import configs, views
from oracle import oracle

OUT = os.path.join(ROOT, "profiles", "canonical_counters.json")


sample_rays = views.stratified_sample  # shared with tests/test_gpu_headline_parity.py


def main():
    names = sys.argv[1:] or ["C2", "C3", "C4", "C5"]
    res = json.load(open(OUT)) if os.path.exists(OUT) else {}
    for name in names:
        t0 = time.time()
        cfg = configs.make(name)
        o, d, stride = sample_rays(cfg)
        m = oracle.OracleMesh(cfg["vertices"], cfg["faces"])
        q = m.trace_fwd(o, d, configs.INT_IOR)
        n = len(o)
        c = q["counters"].astype(float) / n
        st = q["stage"]
        nodes, tris = c[0] + c[2] + c[4], c[1] + c[3] + c[5]
        valid = float((st == 5).mean())
        B_trav = 32.0 * nodes + 36.0 * tris
        rec = dict(
            desc=cfg["desc"], n_tris=int(len(cfg["faces"])), n_verts=int(len(cfg["vertices"])), sample_rays=int(n), sample_stride=int(stride),
            q1_hit_frac=float((st >= 1).mean()), q2_rays_frac=float((st >= 2).mean()), q3_rays_frac=float((st >= 4).mean()),
            valid_frac=valid,
            nodes_per_ray=dict(q1=c[0], q2=c[2], q3=c[4], total=nodes), tris_per_ray=dict(q1=c[1], q2=c[3], q3=c[5], total=tris),
            bytes_per_ray=dict(
                io_fwd=48 + 51, io_bwd=48 + 48, rec_fwd=8, rec_bwd=8, trav=B_trav, vtx_fwd=valid * 144.0, vtx_bwd=valid * 288.0,
                fwd=48 + 51 + 8 + B_trav + valid * 144.0, bwd=48 + 48 + 8 + valid * 288.0,
                total=195 + 16 + B_trav + valid * 288.0),
        )
        res[name] = rec
        print(name, json.dumps(rec["nodes_per_ray"]), json.dumps(rec["tris_per_ray"]), f"valid {valid:.3f}",
              f"B={rec['bytes_per_ray']['total']:.0f} B/ray ({time.time() - t0:.1f}s)", flush=True)
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    json.dump(res, open(OUT, "w"), indent=1)


if __name__ == "__main__":
    main()
