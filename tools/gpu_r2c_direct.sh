#!/bin/bash
# round-2 session 3: the one-thread-per-path forward for small batches (DRT_DIRECT_MAX) against the staged wavefront
mkdir -p gpurun_out
DRT_DIRECT_MAX=2000000000 timeout 500 python -m pytest tests/test_gpu_loss_step.py tests/test_gpu_headline_parity.py -x -q -m gpu > gpurun_out/r2c_direct_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/r2c_direct_pytest.log; tail -3 gpurun_out/r2c_direct_pytest.log
for cfg in "C2|--config C2" "C3v1|--config C3 --views 1" "C4v1|--views 1" "C4v3|--views 3" "C4v9|--views 9"; do
  IFS='|' read -r name args <<< "$cfg"
  BENCH_ARGS="$args" STEPS=20 bash tools/gpu_sweep.sh r2cd_$name "staged|DRT_DIRECT_MAX=0|-" "direct|DRT_DIRECT_MAX=2000000000|-"
done
