"""Checks the closed-form shared-memory layouts of docs/ROUND2_CONV_TCGEN05.md against what TMA actually writes
(eegb200_debug_tma_tile): every element of a [128 x 32] operand k-block must sit at the byte offset the formula gives.
    python tools/gpu_tma_layout_probe.py      (needs a B200)"""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from eeg_image_decode_b200 import _lib  # noqa: E402


def k_major_off(r, k):
    return (r // 8) * 1024 + (r % 8) * 128 + (((k // 4) ^ (r % 8)) * 16) + (k % 4) * 4


def mn_major_off(m, k):
    return (m // 32) * 4096 + (k // 4) * 512 + (k % 4) * 128 + ((((m % 32) // 8) ^ (k % 4)) * 32) + (m % 8) * 4


L = _lib.lib()
L.eegb200_debug_tma_tile.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]
ok = True
for mn_major, name, off in ((0, "K-major SWIZZLE_128B", k_major_off), (1, "MN-major SWIZZLE_128B_ATOM_32B", mn_major_off)):
    ld = 128
    src = torch.arange(128 * ld, dtype=torch.float32, device="cuda").reshape(128, ld)      # value = linear index
    out = torch.empty(4096, dtype=torch.float32, device="cuda")
    _lib.check(L.eegb200_debug_tma_tile(_lib.ptr(src), ld, mn_major, _lib.ptr(out), _lib.stream_ptr()), "debug_tma_tile")
    img = out.cpu()
    bad = 0
    for r in range(128):
        for k in range(32):
            want = float(k * ld + r if mn_major else r * ld + k)
            if img[off(r, k) // 4].item() != want:
                bad += 1
    print(f"{name}: {4096 - bad} of 4096 elements where the formula says" + ("" if bad == 0 else "  <-- FORMULA WRONG"))
    ok &= bad == 0
sys.exit(0 if ok else 1)
