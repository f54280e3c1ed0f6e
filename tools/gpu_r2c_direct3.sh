#!/bin/bash
# re-measure the route crossover with the default (non-cooperative) drain
mkdir -p gpurun_out
for cfg in "C2|--config C2" "C3v1|--config C3 --views 1" "C4v1|--views 1" "C4v2|--views 2" "C4v3|--views 3" "C4v9|--views 9"; do
  IFS='|' read -r name args <<< "$cfg"
  echo "== $name"
  BENCH_ARGS="$args --no-parity-check" STEPS=20 bash tools/gpu_sweep.sh r2cd3_$name "staged_l2|DRT_DIRECT_MAX=0|-" "staged_l4|DRT_DIRECT_MAX=0 DRT_LANES=4|-" "direct|DRT_DIRECT_MAX=2000000000|-"
done
DRT_DIRECT_MAX=0 python bench.py --config C2 --steps 5 --warmup 3 --no-e2e --no-cpu-baseline --no-ref-chain-gpu --no-parity-check > gpurun_out/r2cd3_iter_0.json 2> gpurun_out/r2cd3_iter_0.err
python -c "
import json; d=json.load(open('gpurun_out/r2cd3_iter_0.json')); print('direct_max 0 fused', d['optim_iteration']['fused']['phases_ms'])"
