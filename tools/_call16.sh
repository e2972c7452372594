set -u
mkdir -p gpurun_out
EEGB200_ATTN_TC=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/r01q_bench.json 2> gpurun_out/r01q_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r01q_bench.json'))
print(d['value'], d['ms_per_step'], d['e2e']['value'])
for k in d['top_kernels'][:40]:
    if 'attention' in k['kernel'] or 'N=768' in k['kernel']: print(round(k['ms_per_launch']*1000,1), k['launches_per_step'], k['kernel'])
PY
