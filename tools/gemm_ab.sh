#!/bin/bash
# On the GPU box: time variants of the GEMM kernel against each other on the same box (the box copy of the tree is
# scratch).  Usage: GEMM_FLAVORS_ONLY=ff1 tools/gemm_ab.sh "name|extra nvcc flags" ...
cd "$(dirname "$0")/.."
mkdir -p build/ab
NV="nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -Isdp-net_b200/csrc -Iinclude"
for v in "$@"; do
  IFS='|' read -r name flags <<< "$v"
  $NV $flags -c sdp-net_b200/csrc/gemm_tc.cu -o build/obj/gemm_tc.o 2> build/ab/$name.log || { echo "$name: nvcc failed"; tail -5 build/ab/$name.log; continue; }
  nvcc -shared -gencode arch=compute_100a,code=sm_100a -o sdp-net_b200/lib/libsdpnet_b200.so build/obj/*.o
  echo "== $name"
  timeout 300 python tools/gemm_flavors.py 1024 ${GEMM_SECS:-1.0} 2>&1 | tail -14
done
