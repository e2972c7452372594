set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider -x > gpurun_out/r01l_tests.log 2>&1; tail -3 gpurun_out/r01l_tests.log
python tools/quick_ms.py 2>&1 | tail -1
EEGB200_CONV_BWD=1 python tools/quick_ms.py 2>&1 | tail -1
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/r01l_bench.json 2> gpurun_out/r01l_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r01l_bench.json'))
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline'])
for k in d['top_kernels'][:6]: print(round(k['ms_per_launch']*1000,1), k['launches_per_step'], k['kernel'])
PY
