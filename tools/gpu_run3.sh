#!/bin/bash
# third GPU pass: parity suite on the staged kernels, then the scheduling-policy sweep with bulk-copy staged refill
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
B="python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --no-ref-chain-gpu"
run() { # name, env...
    local name=$1; shift
    env "$@" timeout 200 $B > gpurun_out/sweep_$name.json 2> gpurun_out/sweep_$name.err
}
rm -f gpurun_out/sweep_*
D4=DRT_B200_LIB=$PWD/drt_b200/_C/variants/libdrt_b200_defer4.so
D3=DRT_B200_LIB=$PWD/drt_b200/_C/variants/libdrt_b200_defer3.so
run s0_v0_t32       DRT_STAGED=0 DRT_LS_VOTE=0
run s0_v8_t32       DRT_STAGED=0
run s1_v8_t32
run s1_v0_t32       DRT_LS_VOTE=0
run s1_v8_t24       DRT_LS_THRESH=24
run s1_v8_t16       DRT_LS_THRESH=16
run s1_v8_t8        DRT_LS_THRESH=8
run s1_v4_t8        DRT_LS_THRESH=8 DRT_LS_VOTE=4
run s1_v16_t16      DRT_LS_THRESH=16 DRT_LS_VOTE=16
run s1_v0_t16       DRT_LS_THRESH=16 DRT_LS_VOTE=0
run s0_v8_t16       DRT_STAGED=0 DRT_LS_THRESH=16
run s1_v8_t16_q1    DRT_LS_THRESH_Q1=16
run s1_v8_t16_q2    DRT_LS_THRESH_Q2=16
run s1_v8_t16_q3    DRT_LS_THRESH_Q3=16
run s1_v8_t8_q2     DRT_LS_THRESH_Q2=8
run d4_s1_v8_t32    $D4
run d4_s1_v8_t16    $D4 DRT_LS_THRESH=16
run d4_s1_v8_t8     $D4 DRT_LS_THRESH=8
run d3_s1_v8_t16    $D3 DRT_LS_THRESH=16
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/sweep_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1]); p = d["phases_ms"]
        print("%-22s step %.3f ms  fwd %.3f  bwd %.3f  loss %.6f" % (f[17:-5], d["ms_per_step"], p["fwd"], p["bwd"], d["loss"]))
    except Exception as e:
        print(f, "ERR", e, open(f[:-4] + "err").read()[-300:])
PY
