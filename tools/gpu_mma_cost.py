"""Cycles per tcgen05.mma (kind::tf32, K = 8) for small shapes, chained into one accumulator, alone and with background
activity on the same SM (shared-memory stores / a second issuing thread / tcgen05.ld traffic): csrc/debug_probe.cu.
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
BG = {0: "alone", 1: "+smem stores", 2: "+2nd issuer", 4: "+tcgen05.ld", 7: "+all three", 8: "+commit/4 MMAs",
      16: "+12 warps polling", 17: "+polling+stores", 32: "alternating shapes", 66: "+2nd issuer, other shape",
      128: "+12 warps LDS/STS", 128 + 66: "+2nd other shape +LDS/STS", 256: "+12 warps st+fence.proxy.async", 768: "+12 warps st+fence, paced 200ns"}
print("M    N  major  background                       cycles/MMA (complete)  (issue)")
for M, N, mn in ((128, 48, 0), (64, 48, 1), (64, 32, 1), (128, 96, 0)):
    for bg in (0, 32, 66, 128, 256, 768):
        for rep in range(2):
            _lib.check(L.eegb200_debug_umma_cost(M, N, mn, 512, bg, _lib.ptr(out), _lib.stream_ptr()), "umma_cost")
            torch.cuda.synchronize()
        a, b = out.tolist()
        print(f"{M:3d} {N:4d}   {'MN' if mn else 'K '}   {BG[bg]:30s}   {a / 512:8.1f}            {b / 512:8.1f}")
