#!/bin/bash
# One GPU-box call: parity tests, smoke, bench line, ncu launch list + full captures of the heaviest kernels.
# Everything lands in gpurun_out/ (merged back into the repo's gpurun_out/ by gpurun).
set -u
TAG=${1:-r01c}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/${TAG}_smi.txt 2>&1
echo "== variant tests" ; date +%T
timeout 600 python -m pytest tests/test_gpu_variants.py -m gpu -q -p no:cacheprovider > $OUT/${TAG}_tests_variants.log 2>&1
tail -5 $OUT/${TAG}_tests_variants.log
echo "== parity tests" ; date +%T
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider --deselect tests/test_gpu_variants.py > $OUT/${TAG}_tests.log 2>&1
tail -5 $OUT/${TAG}_tests.log
echo "== smoke" ; date +%T
timeout 300 python __graft_entry__.py smoke > $OUT/${TAG}_smoke.log 2>&1 ; tail -2 $OUT/${TAG}_smoke.log
echo "== bench" ; date +%T
timeout 600 python bench.py --steps 20 --warmup 5 > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err
cut -c1-900 $OUT/${TAG}_bench.json
echo "== ncu launch list" ; date +%T
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/${TAG}_launches.csv \
    python tools/ncu_step.py 3 > $OUT/${TAG}_ncu_list.log 2>&1
wc -l $OUT/${TAG}_launches.csv
echo "== ncu full captures" ; date +%T
cap() {   # name, kernel regex (matched against the demangled name incl. template arguments), skip
  timeout 240 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:$2" -s $3 -c 1 \
      -f -o $OUT/${TAG}_$1 python tools/ncu_step.py 2 > $OUT/${TAG}_ncu_$1.log 2>&1
  ls -la $OUT/${TAG}_$1.ncu-rep 2>/dev/null
}
cap gemm_bnf 'gemm_tf32_kernel<.*256, .*0, .*1, .*256' 1
cap attention_bwd 'attention_bwd_mma' 1
cap conv_fwd 'conv_temporal_fwd_mma' 1
cap bn_elu 'bn_elu_apply' 1
date +%T
echo "== done"
