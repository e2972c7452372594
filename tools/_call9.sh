set -u
mkdir -p gpurun_out
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu > gpurun_out/r01j_bench.json 2> gpurun_out/r01j_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r01j_bench.json'))
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['passes_trials_s'])
for k in d['top_kernels'][:14]: print(round(k['ms_per_launch']*1000,1), k['launches_per_step'], k['kernel'])
PY
