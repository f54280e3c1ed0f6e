"""Builds libdrt_b200 variants with different -D switches for A/B runs on the GPU box (DRT_B200_LIB=<path> selects one).
   python tools/build_variants.py name:-DDRT_X=1,-DDRT_Y=2 ...   -> drt_b200/_C/variants/lib_<name>.so"""
import os, subprocess, sys
from concurrent.futures import ThreadPoolExecutor
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from drt_b200 import build

def one(spec):
    name, _, defs = spec.partition(":")
    out = os.path.join(build.OUT_DIR, "variants", f"lib_{name}.so")
    os.makedirs(os.path.dirname(out), exist_ok=True)
    env = dict(os.environ); env.pop("CC", None); env.pop("CXX", None)
    cmd = ["/usr/local/cuda/bin/nvcc"] + build.NVCC_FLAGS + [d for d in defs.split(",") if d] + ["-o", out, os.path.join(build.SRC, "capi.cu")]
    r = subprocess.run(cmd, env=env, capture_output=True, text=True)
    if r.returncode:
        sys.stderr.write(r.stderr)
        raise SystemExit(f"variant {name} failed")
    return out

if __name__ == "__main__":
    with ThreadPoolExecutor(4) as ex:
        for p in ex.map(one, sys.argv[1:]):
            print(p)
