#!/bin/bash
# vote threshold on the final kernels with the four-lane default (C4, 72 views; C3)
mkdir -p gpurun_out
BENCH_ARGS="--no-parity-check" STEPS=15 bash tools/gpu_sweep.sh r2ck "default||-" "vote2|DRT_VOTE=2|-" "vote3|DRT_VOTE=3|-" "default_again||-" "vote3_again|DRT_VOTE=3|-" "q2v3|DRT_VOTE_Q2=3|-"
BENCH_ARGS="--no-parity-check --config C3" STEPS=15 bash tools/gpu_sweep.sh r2ck_c3 "default||-" "vote3|DRT_VOTE=3|-" "vote2|DRT_VOTE=2|-"
BENCH_ARGS="--no-parity-check --views 9" STEPS=20 bash tools/gpu_sweep.sh r2ck_v9 "default||-" "vote3|DRT_VOTE=3|-"
