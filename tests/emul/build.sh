#!/bin/sh
# TEST INFRASTRUCTURE: host (single-lane) build of the kernel sources, see emul.cpp
set -e
D="$(cd "$(dirname "$0")" && pwd)"
C="$D/../../myochallenge_b200/csrc"
g++ -O2 -std=c++17 -fPIC -shared -I "$D/fake" -o "$D/libmyo_emul.so" "$D/emul.cpp" "$C/myo_model.cpp" "$C/myo_pack.cpp"
