#!/bin/bash
# One GPU-box call: parity tests, smoke, bench line, reference arm, ncu launch list, CUPTI timeline of the captured step.
# Everything lands in gpurun_out/ (merged back into the repo's gpurun_out/ by gpurun).
set -u
TAG=${1:-r01z}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/${TAG}_smi.txt 2>&1
echo "== tests" ; date +%T
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider > $OUT/${TAG}_tests.log 2>&1
tail -4 $OUT/${TAG}_tests.log
echo "== smoke" ; date +%T
timeout 300 python __graft_entry__.py smoke > $OUT/${TAG}_smoke.log 2>&1 ; tail -2 $OUT/${TAG}_smoke.log
echo "== bench" ; date +%T
timeout 600 python bench.py --steps 20 --warmup 5 > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err
cut -c1-700 $OUT/${TAG}_bench.json
echo "== reference arm" ; date +%T
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/${TAG}_bench_ref.json 2> $OUT/${TAG}_bench_ref.err
cut -c1-400 $OUT/${TAG}_bench_ref.json
echo "== ncu launch list" ; date +%T
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu > $OUT/${TAG}_ncu_list.log 2>&1
wc -l $OUT/${TAG}_launches.csv
echo "== ncu full capture of the fused conv kernels" ; date +%T
timeout 400 ncu --set full --clock-control none -k regex:conv_tc -c 8 -f -o $OUT/${TAG}_conv \
    python tools/ncu_step.py 2 > $OUT/${TAG}_ncu_conv.log 2>&1
ls -la $OUT/${TAG}_conv.ncu-rep
echo "== timeline" ; date +%T
timeout 300 python tools/trace_step.py $OUT/${TAG}_trace.json 2>&1 | tail -1
date +%T
echo "== done"
