#!/bin/bash
# GPU pass 7: parity suite (dense-path tiles, raw prmt), sweeps of residency / refill / backward launch bounds
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -6 gpurun_out/pytest_gpu.log
B="python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --no-ref-chain-gpu"
run() { local name=$1; shift; env "$@" timeout 200 $B > gpurun_out/sweep_$name.json 2> gpurun_out/sweep_$name.err; }
rm -f gpurun_out/sweep_*
B4=DRT_B200_LIB=$PWD/drt_b200/_C/variants/libdrt_b200_bwd4.so
B2=DRT_B200_LIB=$PWD/drt_b200/_C/variants/libdrt_b200_bwd2.so
run default
run m10             DRT_Q_MINB=10
run m7              DRT_Q_MINB=7
run t24             DRT_FWD_THRESH=24
run t16_q2          DRT_THRESH_Q2=16
run v6              DRT_VOTE=6
run v2              DRT_VOTE=2
run bwd4            $B4
run bwd2            $B2
run bwd4_merge      $B4 DRT_BWD_MERGE=1
env timeout 200 $B --loss-path rec > gpurun_out/sweep_recpath.json 2> gpurun_out/sweep_recpath.err
env timeout 200 $B --loss-path dense > gpurun_out/sweep_densepath.json 2> gpurun_out/sweep_densepath.err
env DRT_TILE=0 timeout 200 $B --loss-path rec > gpurun_out/sweep_recpath_scan.json 2> gpurun_out/sweep_recpath_scan.err
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/sweep_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1]); p = d["phases_ms"]
        print("%-24s step %.3f ms  build %.3f fwd %.3f  bwd %.3f  loss %.6f" % (f[17:-5], d["ms_per_step"], p["bvh_build"], p["fwd"], p["bwd"], d["loss"]))
    except Exception as e:
        print(f, "ERR", e, open(f[:-4] + "err").read()[-300:])
PY
