set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider -x > gpurun_out/r01i_tests.log 2>&1; tail -4 gpurun_out/r01i_tests.log
python tools/quick_ms.py 2>&1 | tail -1
EEGB200_CONV_XP=2 python tools/quick_ms.py 2>&1 | tail -1
echo "== error budget XP=3"; python tools/gpu_error_budget.py 2>&1 | grep -E "^out|train"
echo "== error budget XP=2"; EEGB200_CONV_XP=2 python tools/gpu_error_budget.py 2>&1 | grep -E "^y1|^out|train"
echo "== error budget XP=1"; EEGB200_CONV_XP=1 python tools/gpu_error_budget.py 2>&1 | grep -E "^y1|^out|train"
timeout 300 python tools/trace_step.py gpurun_out/r01i_trace.json 2>&1 | tail -1
