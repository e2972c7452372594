"""Which TMEM lanes receive the 64 rows of an M = 64 tcgen05.mma accumulator (cta_group::1)?  (csrc/debug_probe.cu)
    python tools/gpu_m64_probe.py      (needs a B200)"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from eeg_image_decode_b200 import _lib  # noqa: E402

L = _lib.lib()
out = torch.zeros(128, device="cuda")
_lib.check(L.eegb200_debug_umma_m64(_lib.ptr(out), _lib.stream_ptr()), "debug_umma_m64")
v = out.cpu().tolist()
rows = {}
for lane, x in enumerate(v):
    if x > 0:
        rows[int(round(x)) - 1] = lane
print("lane -> value:", [int(x) for x in v])
print("row -> lane  :", [rows.get(r, -1) for r in range(64)])
assumed = [32 * (r // 16) + (r % 16) for r in range(64)]
print("ASSUMED MAPPING (rows 16i..16i+15 -> lanes 32i..32i+15):", "CONFIRMED" if [rows.get(r, -1) for r in range(64)] == assumed else "WRONG")
