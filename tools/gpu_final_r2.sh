#!/bin/bash
# gpurun -- 'bash tools/gpu_final_r2.sh': the round-2 record on ONE B200 -- parity suite, smoke, the bench lines of every config,
# the reference arm, and the two ncu passes (launch list over 2 timed steps; --set full of one 72-view step) whose summaries
# tools/summarize_profiles.py turns into profiles/r02_*.  Numbers printed under ncu are never bench values.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r02_pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02_pytest_gpu.log
tail -3 gpurun_out/r02_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke.log 2>&1; tail -2 gpurun_out/r02_smoke.log
timeout 900 python bench.py > gpurun_out/r02_bench_c4_n1.json 2> gpurun_out/r02_bench_c4_n1.err; tail -c 400 gpurun_out/r02_bench_c4_n1.err
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r02_bench_reference_arm.json 2> gpurun_out/r02_bench_reference_arm.err
B="python bench.py --no-e2e --no-cpu-baseline --no-ref-chain-gpu --no-iteration"
timeout 300 $B --config C3 > gpurun_out/r02_bench_c3_n1.json 2> gpurun_out/r02_bench_c3.err
timeout 300 $B --config C2 > gpurun_out/r02_bench_c2_n1.json 2> gpurun_out/r02_bench_c2.err
timeout 300 $B --config C5 --views 32 > gpurun_out/r02_bench_c5_32views_n1.json 2> gpurun_out/r02_bench_c5.err
timeout 300 $B --views 9 > gpurun_out/r02_bench_c4_9views_n1.json 2> /dev/null
timeout 300 $B --views 1 > gpurun_out/r02_bench_c4_1view_n1.json 2> /dev/null
timeout 300 $B --config C3 --views 1 > gpurun_out/r02_bench_c3_1view_n1.json 2> /dev/null
timeout 300 $B --refit > gpurun_out/r02_bench_c4_refit.json 2> /dev/null
DRT_BEAM=0 timeout 300 $B > gpurun_out/r02_bench_c4_nobeam.json 2> /dev/null
DRT_LANES_BIG=1 timeout 300 $B > gpurun_out/r02_bench_c4_one_lane.json 2> /dev/null
P="$B --no-parity-check --graph off"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 600 --csv --log-file gpurun_out/r02_launches.csv \
    $P --steps 2 --warmup 3 > gpurun_out/r02_ncu_launch.log 2>&1
# the --set full capture runs the step on ONE lane (DRT_LANES_BIG=1): the same kernels, one instance of each instead of four
DRT_LANES_BIG=1 timeout 1200 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"ls_|sort_|fit_|emit_" -o gpurun_out/r02_prof_step -f \
    $P --steps 1 --warmup 3 > gpurun_out/r02_ncu_full.log 2>&1
ls -la gpurun_out/r02_prof_step.ncu-rep
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r02_bench_c*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1]); p = d["phases_ms"]
        e = d.get("e2e") or {}
        r = d.get("roofline") or {}
        print("%-34s %.3f Grays/s step %.3f ms  build %.3f fwd %.3f bwd %.3f frac %.3f issue_frac %s e2e %.3f G" % (
            f[15:-5], d["value"] / 1e9, d["ms_per_step"], p["bvh_build"], p["fwd"], p["bwd"], r.get("frac", 0), r.get("issue_frac"), e.get("value", 0) / 1e9))
    except Exception as ex:
        print(f, "ERR", ex)
PY
