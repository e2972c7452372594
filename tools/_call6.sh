timeout 500 python tools/gemm_sweep.py 2>&1 | tail -20
