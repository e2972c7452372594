set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r01g_tests.log 2>&1; tail -6 gpurun_out/r01g_tests.log
python tools/quick_ms.py 2>&1 | tail -1
EEGB200_DH0_FUSED=0 python tools/quick_ms.py 2>&1 | tail -1
EEGB200_ACC_OVERLAP=0 python tools/quick_ms.py 2>&1 | tail -1
EEGB200_DH0_FUSED=0 EEGB200_ACC_OVERLAP=0 python tools/quick_ms.py 2>&1 | tail -1
timeout 300 python tools/trace_step.py gpurun_out/r01g_trace.json 2>&1 | tail -1
