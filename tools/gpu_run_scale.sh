#!/bin/bash
# gpurun --gpus N -- 'bash tools/gpu_run_scale.sh N': the driver's multi-GPU launch of bench.py on N GPUs of one box (peer-memory all-reduce test first)
N=${1:-2}
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_peer_allreduce.py -q -m gpu > gpurun_out/r02_pytest_peer_n$N.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02_pytest_peer_n$N.log; tail -2 gpurun_out/r02_pytest_peer_n$N.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 400 $TR --master-port 29511 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/r02_bench_c4_n$N.json 2> gpurun_out/r02_bench_c4_n$N.err; tail -c 300 gpurun_out/r02_bench_c4_n$N.err
python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/r02_bench_c4_n$N.json").read().strip().splitlines()[-1])
    e = d.get("e2e") or {}
    print("N=$N %.3f G rays/s  step %.3f ms" % (d["value"] / 1e9, d["ms_per_step"]), d["phases_ms"], d["config"].get("allreduce"), d["config"].get("launch"),
          "e2e %.3f G" % (e.get("value", 0) / 1e9), d.get("allreduce_check"), (d.get("shard") or {}).get("rank_compute_ms"))
except Exception as ex:
    print("ERR", ex)
PY
