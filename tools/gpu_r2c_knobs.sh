#!/bin/bash
# the remaining runtime knobs re-checked after the vote re-tune (C4 72 views)
mkdir -p gpurun_out
BENCH_ARGS="--no-parity-check" STEPS=15 bash tools/gpu_sweep.sh r2ck7 "default||-" "q2t28|DRT_THRESH_Q2=28|-" "q2t24|DRT_THRESH_Q2=24|-" "q3t24|DRT_THRESH_Q3=24|-" "q23t16|DRT_THRESH_Q2=16 DRT_THRESH_Q3=16|-" "tile8x4|DRT_TILE_SHAPE=8x4|-" "lanes3|DRT_LANES=3|-" "lanes6|DRT_LANES=6|-" "vote0|DRT_VOTE=0|-"
