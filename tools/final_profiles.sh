#!/bin/bash
# On the GPU box: the ncu captures behind profiles/<tag>_* (run only after the same commands have exited 0 without ncu).
#   1. --set full over ONE whole XL forward (third of three), exported as raw CSV (the .ncu-rep is too large to bring back)
#   2. launch list (gpu__time_duration) of a short bench run
#   3. --set full + source of one XL-shaped attention launch, exported as raw and source CSV
cd "$(dirname "$0")/.."
TAG=${1:-r02f}
OUT=gpurun_out
mkdir -p $OUT
NCU="ncu --clock-control none"
python tools/step_ncu_probe.py > $OUT/${TAG}_step_plain.log 2>&1 || { echo "step probe failed"; tail -5 $OUT/${TAG}_step_plain.log; exit 1; }
L=$(grep -o "launches [0-9]*" $OUT/${TAG}_step_plain.log | tail -1 | cut -d' ' -f2)
echo "launches per forward: $L"
timeout 1200 $NCU --set full -s $((2 * L)) -c $L -o /tmp/${TAG}_step -f python tools/step_ncu_probe.py > $OUT/${TAG}_step_ncu.log 2>&1
ncu -i /tmp/${TAG}_step.ncu-rep --page raw --csv > $OUT/${TAG}_step_raw.csv 2>/dev/null
ls -la /tmp/${TAG}_step.ncu-rep $OUT/${TAG}_step_raw.csv
timeout 900 $NCU --metrics gpu__time_duration.sum -c 2400 --csv --log-file $OUT/${TAG}_launches.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-comparators > $OUT/${TAG}_launches_ncu.log 2>&1
export ATTN_PROBE_XL_ONLY=1
timeout 600 $NCU --set full --import-source on -k regex:attention_tc5 -s 3 -c 1 -o /tmp/${TAG}_attn -f python tools/attn_probe.py > $OUT/${TAG}_attn_ncu.log 2>&1
ncu -i /tmp/${TAG}_attn.ncu-rep --page source --csv > $OUT/${TAG}_attn_source.csv 2>/dev/null
ncu -i /tmp/${TAG}_attn.ncu-rep --page raw --csv > $OUT/${TAG}_attn_raw.csv 2>/dev/null
cp /tmp/${TAG}_attn.ncu-rep $OUT/ 2>/dev/null
ls -la $OUT | grep $TAG
