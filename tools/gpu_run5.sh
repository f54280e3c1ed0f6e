#!/bin/bash
# GPU pass 5: parity suite with tile batches, sweep tile x vote x defer, C3 A/B of the node layouts
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -6 gpurun_out/pytest_gpu.log
B="python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --no-ref-chain-gpu"
run() { local name=$1; shift; env "$@" timeout 200 $B > gpurun_out/sweep_$name.json 2> gpurun_out/sweep_$name.err; }
rm -f gpurun_out/sweep_*
N64=DRT_B200_LIB=$PWD/drt_b200/_C/variants/libdrt_b200_node64.so
D3=DRT_B200_LIB=$PWD/drt_b200/_C/variants/libdrt_b200_q_defer3.so
D4=DRT_B200_LIB=$PWD/drt_b200/_C/variants/libdrt_b200_q_defer4.so
run scan_v0         DRT_TILE=0
run tile_v0
run tile_v4         DRT_VOTE=4
run tile_v8         DRT_VOTE=8
run tile_d3_v4      $D3 DRT_VOTE=4
run tile_d3_v8      $D3 DRT_VOTE=8
run tile_d3_v12     $D3 DRT_VOTE=12
run tile_d4_v8      $D4 DRT_VOTE=8
run tile_d4_v12     $D4 DRT_VOTE=12
run tile_d3_v8_t24  $D3 DRT_VOTE=8 DRT_FWD_THRESH=24
run tile_n64_v8     $N64 DRT_VOTE=8
for c in C3 C2; do
  env DRT_VOTE=8 timeout 200 $B --config $c > gpurun_out/sweep_${c}_q32_tile_v8.json 2> gpurun_out/sweep_${c}_q32_tile_v8.err
  env DRT_VOTE=8 $N64 timeout 200 $B --config $c > gpurun_out/sweep_${c}_n64_tile_v8.json 2> gpurun_out/sweep_${c}_n64_tile_v8.err
  env DRT_VOTE=8 $D3 timeout 200 $B --config $c > gpurun_out/sweep_${c}_q32_d3_tile_v8.json 2> gpurun_out/sweep_${c}_q32_d3_tile_v8.err
  env DRT_TILE=0 $N64 timeout 200 $B --config $c > gpurun_out/sweep_${c}_n64_scan_v0.json 2> gpurun_out/sweep_${c}_n64_scan_v0.err
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/sweep_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1]); p = d["phases_ms"]
        print("%-24s step %.3f ms  build %.3f fwd %.3f  bwd %.3f  loss %.6f" % (f[17:-5], d["ms_per_step"], p["bvh_build"], p["fwd"], p["bwd"], d["loss"]))
    except Exception as e:
        print(f, "ERR", e, open(f[:-4] + "err").read()[-300:])
PY
