set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider -x > gpurun_out/r01k_tests.log 2>&1; tail -3 gpurun_out/r01k_tests.log
python tools/quick_ms.py 2>&1 | tail -1
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/r01k_bench.json 2> gpurun_out/r01k_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r01k_bench.json'))
print(d['value'], d['ms_per_step'], d['e2e']['value'])
for k in d['top_kernels'][:40]:
    if 'attention' in k['kernel'] or 'conv' in k['kernel'] or 'colsum' in k['kernel'] or 'layernorm' in k['kernel']: print(round(k['ms_per_launch']*1000,1), k['launches_per_step'], k['kernel'])
PY
