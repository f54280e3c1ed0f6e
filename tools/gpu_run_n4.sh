#!/bin/bash
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_peer_allreduce.py -q -m gpu > gpurun_out/pytest_peer.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_peer.log; tail -15 gpurun_out/pytest_peer.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29511 bench.py --gpus 4 --steps 20 --warmup 3 --no-e2e > gpurun_out/bench_c4_n4_peer.json 2> gpurun_out/bench_c4_n4_peer.err; tail -c 600 gpurun_out/bench_c4_n4_peer.err
DRT_ALLREDUCE=nccl timeout 300 $TR --master-port 29512 bench.py --gpus 4 --steps 20 --warmup 3 --no-e2e > gpurun_out/bench_c4_n4_nccl.json 2> gpurun_out/bench_c4_n4_nccl.err
python - <<'PY'
import json
for f in ("peer", "nccl"):
    try:
        d = json.loads(open(f"gpurun_out/bench_c4_n4_{f}.json").read().strip().splitlines()[-1])
        print(f, "%.3f G  step %.3f ms" % (d["value"] / 1e9, d["ms_per_step"]), d["phases_ms"], d["config"].get("allreduce"), "grad_norm", d["grad_norm"], "loss", d["loss"])
    except Exception as e:
        print(f, "ERR", e)
PY
