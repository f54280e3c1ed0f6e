#!/bin/bash
# A/B sweep of library variants / env knobs on the GPU box: tools/gpu_sweep.sh <tag> "<name>|<env assignments>|<lib or ->" ...
tag=$1; shift
mkdir -p gpurun_out
for spec in "$@"; do
  IFS='|' read -r name envs lib <<< "$spec"
  libenv=""
  [ "$lib" != "-" ] && [ -n "$lib" ] && libenv="DRT_B200_LIB=$PWD/drt_b200/_C/variants/lib_$lib.so"
  env $envs $libenv python bench.py --steps ${STEPS:-10} --warmup 3 --no-e2e --no-cpu-baseline --no-ref-chain-gpu --no-iteration ${BENCH_ARGS} > gpurun_out/${tag}_$name.json 2> gpurun_out/${tag}_$name.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/${tag}_$name.json"))
    p=d["phases_ms"]; pc=d.get("parity_check") or {}
    print("%-18s step %.3f  build %.3f fwd %.3f bwd %.3f  parity %s  counts %s" % ("$name", d["ms_per_step"], p["bvh_build"], p["fwd"], p["bwd"], pc.get("ok"), d.get("stage_counts_rank0")))
except Exception as e:
    print("$name FAILED", e); print(open("gpurun_out/${tag}_$name.err").read()[-800:])
PY
done
