#!/bin/bash
# On the GPU box: time variants of the attention kernel against each other on the same box (the box copy of the tree is
# scratch).  Usage: tools/attn_ab.sh "name|source|extra nvcc flags" ...
cd "$(dirname "$0")/.."
NV="nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -Isdp-net_b200/csrc -Xptxas -v"
for v in "$@"; do
  IFS='|' read -r name src flags <<< "$v"
  $NV $flags -c ${src:-sdp-net_b200/csrc/attention_tc.cu} -o build/obj/attention_tc.o 2> build/ab/$name.log || { echo "$name: nvcc failed"; tail -5 build/ab/$name.log; continue; }
  nvcc -shared -gencode arch=compute_100a,code=sm_100a -o sdp-net_b200/lib/libsdpnet_b200.so build/obj/*.o
  echo "$name: $(grep -A3 'attention_tc5_kernelILi96ELi261' build/ab/$name.log | grep -o 'Used [0-9]* registers\|[0-9]* bytes spill stores' | tr '\n' ' ') $(timeout 120 python tools/attn_probe.py 2>&1 | grep -E 'XL attention|EXC|nan [1-9]' | tr '\n' ' ')"
done
