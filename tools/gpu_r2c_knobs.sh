#!/bin/bash
# vote every 8 steps, threshold 1: queue depth, full-queue walk, neighbours
mkdir -p gpurun_out
BENCH_ARGS="--no-parity-check" STEPS=15 DRT_VOTE=1 bash tools/gpu_sweep.sh r2ck6 "ve8|DRT_VOTE=1|ve8" "ve8d2|DRT_VOTE=1|ve8d2" "ve8d4|DRT_VOTE=1|ve8d4" "ve8fq0|DRT_VOTE=1|ve8fq0" "ve7|DRT_VOTE=1|ve7" "ve9|DRT_VOTE=1|ve9" "ve8d4_v2|DRT_VOTE=2|ve8d4"
