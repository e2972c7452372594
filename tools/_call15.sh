set -u
mkdir -p gpurun_out
EEGB200_ATTN_TC=1 timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -p no:cacheprovider -x -k "forward_eval_stages or golden or dropout_tensorcore or train_step_grad" > gpurun_out/r01p_tests.log 2>&1; tail -15 gpurun_out/r01p_tests.log | cut -c1-220
EEGB200_ATTN_TC=1 timeout 120 python tools/quick_ms.py 2>&1 | tail -2
