set -u
mkdir -p gpurun_out
echo "== probe"; PASSES=4 timeout 300 python tools/e2e_probe.py 2>&1 | tail -8
echo "== trace"; timeout 300 python tools/trace_step.py gpurun_out/r01f_trace.json 2>&1 | tail -3
