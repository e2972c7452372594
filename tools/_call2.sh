set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_variants.py tests/test_gpu_parity.py -m gpu -q -p no:cacheprovider -k "loops or evaluate" > gpurun_out/r01d_tests.log 2>&1; tail -5 gpurun_out/r01d_tests.log
echo "== probe cache on"; timeout 300 python tools/e2e_probe.py 2>&1 | tail -12
echo "== probe cache off"; EEGB200_STEP_CACHE=0 PASSES=4 timeout 300 python tools/e2e_probe.py 2>&1 | tail -6
echo "== probe cache on, no callback"; CB=0 PASSES=4 timeout 300 python tools/e2e_probe.py 2>&1 | tail -6
