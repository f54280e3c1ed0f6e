#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_loss_step.py tests/test_gpu_parity.py -x -q -m gpu -k "layout or image_size or large_batch" > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log; tail -3 gpurun_out/pytest_gpu.log
DRT_TILE_SHAPE=4x8 timeout 300 python -m pytest tests/test_gpu_loss_step.py tests/test_gpu_parity.py -x -q -m gpu -k "layout or image_size" > gpurun_out/pytest_gpu_4x8.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu_4x8.log; tail -3 gpurun_out/pytest_gpu_4x8.log
B="python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --no-ref-chain-gpu"
run() { local name=$1; shift; env "$@" timeout 200 $B > gpurun_out/sweep_$name.json 2> gpurun_out/sweep_$name.err; }
rm -f gpurun_out/sweep_*
run default
run t4x8            DRT_TILE_SHAPE=4x8
run t16x2           DRT_TILE_SHAPE=16x2
run rg16            DRT_R_GRID=16
run rg32            DRT_R_GRID=32
run rg4             DRT_R_GRID=4
run vq1_2           DRT_VOTE_Q1=2
run vq1_8           DRT_VOTE_Q1=8
run vq2_8           DRT_VOTE_Q2=8
run vq2_2           DRT_VOTE_Q2=2
run vq3_8           DRT_VOTE_Q3=8
run t4x8_rg32       DRT_TILE_SHAPE=4x8 DRT_R_GRID=32
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/sweep_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1]); p = d["phases_ms"]
        print("%-24s step %.3f ms  build %.3f fwd %.3f  bwd %.3f  loss %.6f" % (f[17:-5], d["ms_per_step"], p["bvh_build"], p["fwd"], p["bwd"], d["loss"]))
    except Exception as e:
        print(f, "ERR", e, open(f[:-4] + "err").read()[-300:])
PY
