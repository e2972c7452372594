set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r01h_tests.log 2>&1; tail -4 gpurun_out/r01h_tests.log
python tools/quick_ms.py 2>&1 | tail -1
EEGB200_WGRAD_SPLIT=0 python tools/quick_ms.py 2>&1 | tail -1
timeout 300 python tools/trace_step.py gpurun_out/r01h_trace.json 2>&1 | tail -1
