#!/bin/bash
# round 2b: prepared tile beams A/B (C4 72 views, 9 views, C2) + the loss-step tests
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_loss_step.py tests/test_gpu_headline_parity.py -x -q -m gpu > gpurun_out/r2b_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2b_pytest.log; tail -5 gpurun_out/r2b_pytest.log
B="python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu-baseline --no-ref-chain-gpu --no-iteration"
for tb in prepared inline; do
  timeout 200 $B --tile-beams $tb > gpurun_out/r2b_c4_$tb.json 2> gpurun_out/r2b_c4_$tb.err
  timeout 200 $B --tile-beams $tb --views 9 > gpurun_out/r2b_c4v9_$tb.json 2> gpurun_out/r2b_c4v9_$tb.err
  timeout 200 $B --tile-beams $tb --config C2 > gpurun_out/r2b_c2_$tb.json 2> gpurun_out/r2b_c2_$tb.err
  timeout 200 $B --tile-beams $tb --config C3 > gpurun_out/r2b_c3_$tb.json 2> gpurun_out/r2b_c3_$tb.err
done
python - <<'PY'
import json
for f in ("c4", "c4v9", "c2", "c3"):
  for tb in ("prepared", "inline"):
    try:
        d = json.loads(open(f"gpurun_out/r2b_{f}_{tb}.json").read().strip().splitlines()[-1]); p = d["phases_ms"]
        print("%-5s %-8s %.3f Grays/s step %.3f ms  build %.3f fwd %.3f bwd %.3f launches %d parity %s" % (f, tb, d["value"] / 1e9, d["ms_per_step"], p["bvh_build"], p["fwd"], p["bwd"], d["gpu_launches"], (d.get("parity_check") or {}).get("ok")))
    except Exception as e:
        print(f, tb, "ERR", e, open(f"gpurun_out/r2b_{f}_{tb}.err").read()[-400:])
PY
