#!/bin/bash
# round-2 session 3: A/B of the cooperative leaf drain (DRT_COOP_DRAIN) and the leaf fall-through (DRT_LEAF_FALL); parity first
mkdir -p gpurun_out
V=$PWD/drt_b200/_C/variants
DRT_B200_LIB=$V/lib_coopfall.so timeout 500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_headline_parity.py tests/test_gpu_loss_step.py -x -q -m gpu > gpurun_out/r2c_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/r2c_pytest.log; tail -3 gpurun_out/r2c_pytest.log
STEPS=10 bash tools/gpu_sweep.sh r2c "base||base" "coop||coop" "fall||fall" "coopfall||coopfall" "coop_v6|DRT_VOTE=6|coop" "coop_v8|DRT_VOTE=8|coop" \
   "coopfall_v6|DRT_VOTE=6|coopfall" "coopd4||coopd4" "coopd4_v8|DRT_VOTE=8|coopd4" "coopfalld4_v6|DRT_VOTE=6|coopfalld4"
