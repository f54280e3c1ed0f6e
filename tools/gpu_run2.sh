#!/bin/bash
# second GPU pass: full parity suite, bench with bucketed targets + chunked e2e, scheduling-policy sweep
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
timeout 600 python bench.py > gpurun_out/bench_step.json 2> gpurun_out/bench_step.err; tail -c 800 gpurun_out/bench_step.err
B="python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --no-ref-chain-gpu"
run() { # name, env...
    local name=$1; shift
    env "$@" timeout 200 $B > gpurun_out/sweep_$name.json 2> gpurun_out/sweep_$name.err
}
D4=DRT_B200_LIB=$PWD/drt_b200/_C/variants/libdrt_b200_defer4.so
D3=DRT_B200_LIB=$PWD/drt_b200/_C/variants/libdrt_b200_defer3.so
run base            DRT_VOTE=0
run v8              DRT_VOTE=8
run v8_t16          DRT_VOTE=8 DRT_FWD_THRESH=16
run v8_t8           DRT_VOTE=8 DRT_FWD_THRESH=8
run v16_t16         DRT_VOTE=16 DRT_FWD_THRESH=16
run v4_t8           DRT_VOTE=4 DRT_FWD_THRESH=8
run v0_t16          DRT_VOTE=0 DRT_FWD_THRESH=16
run d4_base         $D4 DRT_VOTE=0
run d4_v8_t8        $D4 DRT_VOTE=8 DRT_FWD_THRESH=8
run d4_v8_t16       $D4 DRT_VOTE=8 DRT_FWD_THRESH=16
run d4_v16_t16      $D4 DRT_VOTE=16 DRT_FWD_THRESH=16
run d3_v8_t8        $D3 DRT_VOTE=8 DRT_FWD_THRESH=8
run v8_t8_q2only    DRT_VOTE_Q2=8 DRT_THRESH_Q2=8
run v8_t8_q1only    DRT_VOTE_Q1=8 DRT_THRESH_Q1=8
run v8_t8_q3only    DRT_VOTE_Q3=8 DRT_THRESH_Q3=8
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/sweep_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1]); p = d["phases_ms"]
        print("%-22s step %.3f ms  fwd %.3f  bwd %.3f  loss %.6f" % (f[17:-5], d["ms_per_step"], p["fwd"], p["bwd"], d["loss"]))
    except Exception as e:
        print(f, "ERR", e)
PY
