#!/bin/bash
# GPU pass 4: parity suite on the quantised 32-byte nodes, A/B against the 64-byte float nodes
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -6 gpurun_out/pytest_gpu.log
B="python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --no-ref-chain-gpu"
run() { local name=$1; shift; env "$@" timeout 200 $B > gpurun_out/sweep_$name.json 2> gpurun_out/sweep_$name.err; }
rm -f gpurun_out/sweep_*
N64=DRT_B200_LIB=$PWD/drt_b200/_C/variants/libdrt_b200_node64.so
D3=DRT_B200_LIB=$PWD/drt_b200/_C/variants/libdrt_b200_q_defer3.so
run q32_v0
run q32_v8          DRT_VOTE=8
run q32_v8_m7       DRT_VOTE=8 DRT_Q_MINB=7
run q32_v8_m6       DRT_VOTE=8 DRT_Q_MINB=6
run q32_v8_m10      DRT_VOTE=8 DRT_Q_MINB=10
run q32_v12         DRT_VOTE=12
run q32_v4          DRT_VOTE=4
run q32_d3_v8       $D3 DRT_VOTE=8
run n64_v0          $N64
run n64_v8          $N64 DRT_VOTE=8
run q32_v8_rec      DRT_VOTE=8 BENCH_EXTRA=1
env DRT_VOTE=8 timeout 200 $B --loss-path rec > gpurun_out/sweep_q32_v8_recpath.json 2> gpurun_out/sweep_q32_v8_recpath.err
env DRT_VOTE=8 timeout 200 $B --config C3 > gpurun_out/sweep_q32_v8_C3.json 2> gpurun_out/sweep_q32_v8_C3.err
env DRT_VOTE=8 timeout 200 $B --config C2 > gpurun_out/sweep_q32_v8_C2.json 2> gpurun_out/sweep_q32_v8_C2.err
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/sweep_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1]); p = d["phases_ms"]
        print("%-22s step %.3f ms  build %.3f fwd %.3f  bwd %.3f  loss %.6f" % (f[17:-5], d["ms_per_step"], p["bvh_build"], p["fwd"], p["bwd"], d["loss"]))
    except Exception as e:
        print(f, "ERR", e, open(f[:-4] + "err").read()[-300:])
PY
