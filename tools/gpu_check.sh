#!/bin/bash
# Run on the GPU box (via gpurun): every test group in its own process with a timeout, so one
# faulting kernel cannot hide the others; logs land in gpurun_out/.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
OUT=gpurun_out
: > $OUT/summary.txt
nvidia-smi --query-gpu=name,driver_version,clocks.max.sm,memory.total --format=csv,noheader | tee -a $OUT/summary.txt
run() { # name timeout cmd...
  local name=$1 to=$2; shift 2
  local t0=$(date +%s)
  timeout -k 10 "$to" "$@" > "$OUT/$name.log" 2>&1
  local rc=$?
  echo "[$name] rc=$rc $(( $(date +%s) - t0 ))s :: $(tail -n 1 "$OUT/$name.log" | cut -c1-160)" | tee -a $OUT/summary.txt
}
PT="python -m pytest -q -p no:cacheprovider -m gpu --timeout 600"
GROUPS_DEFAULT="smoke gemm_bf16 gemm_fp32 act rows dwconv attention preprocess parity bench"
for g in ${@:-$GROUPS_DEFAULT}; do
  case $g in
    smoke)     run smoke 600 python -c "import __graft_entry__ as g; g.smoke()" ;;
    gemm_bf16) run k_gemm_bf16 600 $PT tests/test_gpu_kernels.py -k "gemm_bf16 or misaligned" ;;
    gemm_fp32) run k_gemm_fp32 300 $PT tests/test_gpu_kernels.py -k "gemm_fp32" ;;
    act)       run k_act 300 $PT tests/test_gpu_kernels.py -k "activation or fast_gelu or launch_counter" ;;
    rows)      run k_rows 300 $PT tests/test_gpu_kernels.py -k "layernorm or pool_ln or bridges or im2col or embed_tokens" ;;
    dwconv)    run k_dwconv 300 $PT tests/test_gpu_kernels.py -k "ln_dwconv" ;;
    attention) run k_attention 300 $PT tests/test_gpu_kernels.py -k "attention" ;;
    preprocess) run k_preprocess 300 $PT tests/test_gpu_preprocess.py ;;
    parity)    run parity 1500 $PT tests/test_gpu_parity.py ;;
    bench)     run bench 900 python bench.py --steps 5 --warmup 3 ;;
    bench_s)   run bench_s 600 python bench.py --steps 5 --warmup 3 --config S --no-cpu-baseline ;;
    bench_m)   run bench_m 600 python bench.py --steps 5 --warmup 3 --config M --no-cpu-baseline ;;
    refarm)    run refarm 600 python bench.py --impl reference --steps 2 --warmup 1 ;;
    full)      run full_gpu_suite 2400 $PT tests -x ;;
  esac
done
echo "==== failures ===="
grep -h -E "^(FAILED|ERROR)|Error|error:|assert " $OUT/k_*.log $OUT/parity.log $OUT/smoke.log 2>/dev/null | cut -c1-220 | head -60
echo "==== bench ===="
tail -n 3 $OUT/bench.log 2>/dev/null | cut -c1-3000
