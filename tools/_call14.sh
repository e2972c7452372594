set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -p no:cacheprovider -x -k "train_step or dropout or graph" > gpurun_out/r01o_tests.log 2>&1; tail -3 gpurun_out/r01o_tests.log
python tools/quick_ms.py 2>&1 | tail -1
