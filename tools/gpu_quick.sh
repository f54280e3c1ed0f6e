#!/bin/bash
# quick validation: the GPU parity suite + one resident bench line per loss route
mkdir -p gpurun_out
timeout 600 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log; tail -4 gpurun_out/pytest_gpu.log
B="python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu-baseline --no-ref-chain-gpu"
timeout 200 $B > gpurun_out/quick_step.json 2> gpurun_out/quick_step.err
timeout 200 $B --loss-path rec > gpurun_out/quick_rec.json 2> gpurun_out/quick_rec.err
timeout 200 $B --config C3 > gpurun_out/quick_c3.json 2> gpurun_out/quick_c3.err
python - <<'PY'
import json
for f in ("step", "rec", "c3"):
    try:
        d = json.loads(open(f"gpurun_out/quick_{f}.json").read().strip().splitlines()[-1]); p = d["phases_ms"]
        print("%-6s %.3f Grays/s step %.3f ms  build %.3f fwd %.3f bwd %.3f launches %d" % (f, d["value"] / 1e9, d["ms_per_step"], p["bvh_build"], p["fwd"], p["bwd"], d["gpu_launches"]))
    except Exception as e:
        print(f, "ERR", e, open(f"gpurun_out/quick_{f}.err").read()[-400:])
PY
