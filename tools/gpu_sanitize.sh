#!/bin/bash
# compute-sanitizer over smoke() (dense path + fused ray-loss step, 4096 rays) and a small fused-step parity subset
mkdir -p gpurun_out
S='import __graft_entry__ as g; g.smoke()'
timeout 240 compute-sanitizer --tool memcheck --error-exitcode 9 python -c "$S" > gpurun_out/sanitizer_memcheck_smoke.txt 2>&1; echo "memcheck exit $?" >> gpurun_out/sanitizer_memcheck_smoke.txt
timeout 240 compute-sanitizer --tool racecheck --error-exitcode 9 python -c "$S" > gpurun_out/sanitizer_racecheck_smoke.txt 2>&1; echo "racecheck exit $?" >> gpurun_out/sanitizer_racecheck_smoke.txt
timeout 240 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_loss_step.py -q -x -m gpu -k "degenerate or generate_rays or cabi or (every_input_layout and hand_vh)" > gpurun_out/sanitizer_memcheck_step.txt 2>&1; echo "memcheck exit $?" >> gpurun_out/sanitizer_memcheck_step.txt
tail -3 gpurun_out/sanitizer_memcheck_smoke.txt gpurun_out/sanitizer_racecheck_smoke.txt gpurun_out/sanitizer_memcheck_step.txt
timeout 200 python -m pytest tests/test_optimize_loop.py -q -m gpu 2>&1 | tail -2
