set -u
mkdir -p gpurun_out
echo "== probe ring"; PASSES=6 timeout 300 python tools/e2e_probe.py 2>&1 | tail -14
echo "== probe ring, no callback"; CB=0 PASSES=4 timeout 300 python tools/e2e_probe.py 2>&1 | tail -8
timeout 600 python -m pytest tests -m gpu -q -p no:cacheprovider -k "loops or resident or graph" > gpurun_out/r01e_tests.log 2>&1; tail -4 gpurun_out/r01e_tests.log
