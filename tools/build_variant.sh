#!/bin/bash
# builds an A/B variant of libdrt_b200.so with extra -D flags: tools/build_variant.sh <name> [-DDRT_DEFER=4 ...]
# -> drt_b200/_C/variants/libdrt_b200_<name>.so ; select it with DRT_B200_LIB=<path> (drt_b200/_lib.py)
set -e
cd "$(dirname "$0")/.."
name=$1; shift
mkdir -p drt_b200/_C/variants
env -u CC -u CXX /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 \
    -Xcompiler -fPIC,-fvisibility=hidden -shared "$@" -o drt_b200/_C/variants/libdrt_b200_$name.so drt_b200/csrc/capi.cu
echo drt_b200/_C/variants/libdrt_b200_$name.so
