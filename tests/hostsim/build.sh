#!/bin/bash
# TEST INFRASTRUCTURE: links libbm's own objects (built by ../../build.sh for sm_100a) against tests/hostsim/fake_cudart.cpp
# instead of libcudart -> tests/hostsim/_build/libbm_hostsim.so: the host side of the library runs without a GPU
# (kernel launches are recorded and skipped; see fake_cudart.cpp for what is checked).
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
ROOT="$(cd "$HERE/../.." && pwd)"
OBJ="$ROOT/boltzmann-machines_b200/build"
OUT="$HERE/_build"
ls "$OBJ"/*.o > /dev/null
mkdir -p "$OUT"
g++ -O1 -g -std=c++17 -fPIC -I/usr/local/cuda/include -c "$HERE/fake_cudart.cpp" -o "$OUT/fake_cudart.o"
g++ -O2 -g -std=c++17 -fPIC -I/usr/local/cuda/include -c "$HERE/kernels_cpu.cpp" -o "$OUT/kernels_cpu.o"
g++ -shared -Wl,-Bsymbolic -o "$OUT/libbm_hostsim.so" "$OBJ"/*.o "$OUT/fake_cudart.o" "$OUT/kernels_cpu.o" -ldl -lpthread -lrt -lstdc++
g++ -O2 -g -std=c++17 -fPIC -shared -o "$OUT/libfakenccl.so" "$HERE/fake_nccl.cpp" -lrt -lpthread
echo "built $OUT/libbm_hostsim.so"
