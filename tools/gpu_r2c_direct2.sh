#!/bin/bash
# direct route: residency variants, and the 1.23 M-ray single view of bench.py's optim_iteration with the route forced either way
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_loss_step.py tests/test_gpu_headline_parity.py -x -q -m gpu > gpurun_out/r2c_direct_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/r2c_direct_pytest.log; tail -3 gpurun_out/r2c_direct_pytest.log
for cfg in "C2|--config C2" "C3v1|--config C3 --views 1" "C4v1|--views 1"; do
  IFS='|' read -r name args <<< "$cfg"
  BENCH_ARGS="$args" STEPS=20 bash tools/gpu_sweep.sh r2cd2_$name "direct|DRT_DIRECT_MAX=2000000000|-" "direct_minb6|DRT_DIRECT_MAX=2000000000|d6" "direct_minb8|DRT_DIRECT_MAX=2000000000|d8"
done
for m in 0 2000000000; do
  DRT_DIRECT_MAX=$m python bench.py --config C2 --steps 5 --warmup 3 --no-e2e --no-cpu-baseline --no-ref-chain-gpu --no-parity-check > gpurun_out/r2cd2_iter_$m.json 2> gpurun_out/r2cd2_iter_$m.err
  python -c "
import json; d=json.load(open('gpurun_out/r2cd2_iter_$m.json')); print('direct_max $m', d['optim_iteration']['phases_ms'], d['optim_iteration'].get('fused_losses',{}) )"
done
