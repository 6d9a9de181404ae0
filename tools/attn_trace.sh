#!/bin/bash
# On the GPU box: rebuild the attention kernel with its cycle trace compiled in (-DAT_TRACE=1, block 0 prints
# (role, warp, code, clock64) per trace point), run one XL-shaped launch, leave the trace in gpurun_out/$1.
# The box copy of the tree is scratch: the shipped library is not touched.
cd "$(dirname "$0")/.."
set -e
OUT=gpurun_out/${1:-attn_trace.txt}
NV="nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC"
$NV -DAT_TRACE=1 -c sdp-net_b200/csrc/attention_tc.cu -o build/obj/attention_tc.o
nvcc -shared -gencode arch=compute_100a,code=sm_100a -o sdp-net_b200/lib/libsdpnet_b200.so build/obj/*.o
python - > $OUT <<'PY'
import torch, sdpnet_b200 as sdp
B, S, h, d = 1024, 261, 8, 96
qkv = torch.randn(B, S, 3 * h * d, device="cuda").bfloat16()
out = torch.empty(B, S, h * d, device="cuda", dtype=torch.bfloat16)
sdp.ops.attention(qkv, out, h, None, None, None, None)
torch.cuda.synchronize()
PY
wc -l $OUT
