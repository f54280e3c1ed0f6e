"""Writes the stratified ray sample of a config for tools/stepsim/stepsim.c and runs it:  python tools/stepsim/run.py C4"""
import os, struct, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np
from drt_b200 import configs, views
here = os.path.dirname(os.path.abspath(__file__))
name = sys.argv[1] if len(sys.argv) > 1 else "C4"
cfg = configs.make(name)
o, d, _ = views.stratified_sample(cfg, 65536)
path = os.path.join(here, "rays.bin")
with open(path, "wb") as f:
    f.write(struct.pack("3i", len(cfg["vertices"]), len(cfg["faces"]), len(o)))
    f.write(np.ascontiguousarray(cfg["vertices"], np.float64).tobytes())
    f.write(np.ascontiguousarray(cfg["faces"], np.int32).tobytes())
    f.write(np.ascontiguousarray(o).tobytes()); f.write(np.ascontiguousarray(d).tobytes())
exe = os.path.join(here, "stepsim")
subprocess.check_call(["gcc", "-O2", "-ffp-contract=off", "-o", exe, os.path.join(here, "stepsim.c"), "-lm"])
subprocess.check_call([exe, path])
