#!/bin/bash
# final check of the round: the exact commands the driver runs at round end (GPU tests, smoke, default bench, reference arm)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log; tail -5 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
timeout 300 python bench.py --impl reference > gpurun_out/bench_reference_arm.json 2> gpurun_out/bench_reference_arm.err; cut -c1-200 gpurun_out/bench_reference_arm.json
timeout 600 python bench.py > gpurun_out/bench_c4_n1.json 2> gpurun_out/bench_c4_n1.err; tail -c 400 gpurun_out/bench_c4_n1.err; cut -c1-400 gpurun_out/bench_c4_n1.json
