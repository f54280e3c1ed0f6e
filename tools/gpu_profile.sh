#!/bin/bash
# gpurun -- 'bash tools/gpu_profile.sh': parity suite, smoke, the bench lines of every config, A/B of the loss routes, and the two ncu
# passes (launch list over 2 timed steps; --set full of one 8-view step) whose summaries tools/summarize_profiles.py writes to profiles/
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
timeout 600 python bench.py > gpurun_out/bench_c4_n1.json 2> gpurun_out/bench_c4_n1.err; tail -c 600 gpurun_out/bench_c4_n1.err
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_reference_arm.json 2> gpurun_out/bench_reference_arm.err
B="python bench.py --no-e2e --no-cpu-baseline --no-ref-chain-gpu"
timeout 300 $B --config C3 > gpurun_out/bench_c3_n1.json 2> gpurun_out/bench_c3.err
timeout 300 $B --config C2 > gpurun_out/bench_c2_n1.json 2> gpurun_out/bench_c2.err
timeout 300 $B --config C5 --views 32 > gpurun_out/bench_c5_32views_n1.json 2> gpurun_out/bench_c5.err
DRT_BWD_MERGE=1 timeout 300 $B --steps 10 > gpurun_out/sweep_bwd_merge1.json 2> /dev/null
DRT_BWD_MERGE=0 timeout 300 $B --steps 10 > gpurun_out/sweep_bwd_merge0.json 2> /dev/null
timeout 300 $B --steps 10 --loss-path rec > gpurun_out/sweep_rec_path.json 2> /dev/null
timeout 300 $B --steps 10 --loss-path dense > gpurun_out/sweep_dense_path.json 2> /dev/null
timeout 300 $B --steps 10 --refit > gpurun_out/sweep_refit.json 2> /dev/null
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 400 --csv --log-file gpurun_out/launches.csv \
    $B --steps 2 --warmup 3 > gpurun_out/ncu_launch_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"ls_|sort_|fit_" -o gpurun_out/prof_step -f \
    $B --steps 1 --warmup 3 --views 8 > gpurun_out/ncu_full_bench.log 2>&1
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/bench_c*.json") + glob.glob("gpurun_out/sweep_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1]); p = d["phases_ms"]
        e = d.get("e2e") or {}
        print("%-26s %.3f Grays/s step %.3f ms  build %.3f fwd %.3f bwd %.3f frac %.3f e2e %.3f G" % (f[11:-5], d["value"] / 1e9, d["ms_per_step"], p["bvh_build"], p["fwd"], p["bwd"], d["roofline"]["frac"], e.get("value", 0) / 1e9))
    except Exception as ex:
        print(f, "ERR", ex)
PY
