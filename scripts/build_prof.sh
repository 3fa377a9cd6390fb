#!/bin/sh
# Development aid: build the library with -DMYO_PROFILE (per-phase cycle counters) into scripts/_prof/
set -e
D="$(cd "$(dirname "$0")" && pwd)"
C="$D/../myochallenge_b200/csrc"
mkdir -p "$D/_prof"
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -DMYO_PROFILE \
  -shared -o "$D/_prof/libmyo_prof.so" "$C/myo_model.cpp" "$C/myo_pack.cpp" "$C/myo_kernels.cu" "$C/myo_policy.cu" "$C/myo_rollout.cu" "$C/myo_ppo.cu" "$C/myo_lstm_seq.cu" "$C/myo_util.cu" -prec-div=false -prec-sqrt=false -ftz=true -lcudart -lcublas
