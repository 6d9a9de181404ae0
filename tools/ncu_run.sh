#!/bin/bash
# usage (under gpurun): bash tools/ncu_run.sh <kernel-regex> <outname> [B]
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
REGEX=${1:-"ln_dwconv|attention_bf16|gemm_bf16_tc|ln_rows"}
NAME=${2:-prof}
B=${3:-128}
python tools/ncu_probe.py $B 2 > gpurun_out/${NAME}_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"$REGEX" -o gpurun_out/$NAME -f python tools/ncu_probe.py $B 1 > gpurun_out/${NAME}_ncu.log 2>&1
echo "rc=$?"; cat gpurun_out/${NAME}_plain.log; tail -5 gpurun_out/${NAME}_ncu.log
