#!/bin/bash
# first GPU pass of a session: parity tests, smoke, bench (new step path + A/B against the three-call route), ncu passes
mkdir -p gpurun_out
set -x
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
timeout 600 python bench.py > gpurun_out/bench_step.json 2> gpurun_out/bench_step.err; tail -c 1500 gpurun_out/bench_step.err
timeout 300 python bench.py --loss-path rec --no-e2e --no-cpu-baseline --no-ref-chain-gpu > gpurun_out/bench_rec.json 2> gpurun_out/bench_rec.err
for c in C3 C2; do timeout 300 python bench.py --config $c --no-e2e --no-cpu-baseline --no-ref-chain-gpu > gpurun_out/bench_step_$c.json 2>> gpurun_out/bench_step.err; done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-ref-chain-gpu > gpurun_out/ncu_launch_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"ls_|sort_|fit_" -o gpurun_out/prof_step -f \
    python bench.py --steps 1 --warmup 3 --views 8 --no-e2e --no-cpu-baseline --no-ref-chain-gpu > gpurun_out/ncu_full_bench.log 2>&1
ls -la gpurun_out
