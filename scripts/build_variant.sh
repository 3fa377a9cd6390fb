#!/bin/sh
# Development aid: build the library with extra nvcc flags into scripts/_prof/libmyo_<name>.so   usage: build_variant.sh name -DFLAG=1 ...
set -e
D="$(cd "$(dirname "$0")" && pwd)"
C="$D/../myochallenge_b200/csrc"
NAME="$1"; shift
mkdir -p "$D/_prof"
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC "$@" \
  -shared -o "$D/_prof/libmyo_$NAME.so" "$C/myo_model.cpp" "$C/myo_pack.cpp" "$C/myo_kernels.cu" "$C/myo_policy.cu" "$C/myo_rollout.cu" "$C/myo_ppo.cu" "$C/myo_util.cu" -lcudart -lcublas
