"""Cycles per tcgen05.mma (kind::tf32, K = 8) for small shapes, chained into one accumulator (csrc/debug_probe.cu).
    python tools/gpu_mma_cost.py"""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from eeg_image_decode_b200 import _lib  # noqa: E402

L = _lib.lib()
L.eegb200_debug_umma_cost.argtypes = [ctypes.c_int] * 5 + [ctypes.c_void_p, ctypes.c_void_p]
out = torch.zeros(2, dtype=torch.int64, device="cuda")
print("M   N    major  k_steps   cycles/MMA (complete)  cycles/MMA (issue)   floor M*N/256... (N/2 at M=128)")
for mn in (0, 1):
    for M in (128, 64):
        for N in (32, 48, 96, 160, 192, 256):
            for ks in (4,):
                for rep in range(2):
                    _lib.check(L.eegb200_debug_umma_cost(M, N, mn, 256, ks, _lib.ptr(out), _lib.stream_ptr()), "umma_cost")
                    torch.cuda.synchronize()
                a, b = out.tolist()
                print(f"{M:3d} {N:4d}   {'MN' if mn else 'K '}    {ks:3d}       {a / 256:8.1f}               {b / 256:8.1f}          {N / 2:6.0f}")
