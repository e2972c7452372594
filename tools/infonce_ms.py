"""Cost of one rank's share of the symmetric InfoNCE loss + gradient at the data-parallel problem sizes: B local rows
against N = world * B gathered targets (image + text), no communication (what an 8-rank step adds over a 1-rank step
in pure compute).     python tools/infonce_ms.py [N ...]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from eeg_image_decode_b200 import _lib  # noqa: E402
from eeg_image_decode_b200.loss import _InfoNCE, fused_contrastive  # noqa: E402

B, D = 1024, 1024
dev = torch.device("cuda")
g = torch.Generator().manual_seed(0)
eeg = torch.randn(B, D, generator=g).to(dev)
scale = torch.tensor(2.6593, device=dev)
for N in [int(a) for a in sys.argv[1:]] or [1024, 2048, 8192]:
    img = torch.nn.functional.normalize(torch.randn(N, D, generator=g), dim=-1).to(dev)
    txt = torch.nn.functional.normalize(torch.randn(N, D, generator=g), dim=-1).to(dev)
    nce = _InfoNCE()
    for _ in range(3):
        fused_contrastive(nce, eeg, img, txt, scale, 0.99, row_offset=0, need_grad=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        fused_contrastive(nce, eeg, img, txt, scale, 0.99, row_offset=0, need_grad=True)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    _lib.prof_enable(True)
    fused_contrastive(nce, eeg, img, txt, scale, 0.99, row_offset=0, need_grad=True)
    torch.cuda.synchronize()
    rep = _lib.prof_report()
    _lib.prof_enable(False)
    print(f"B={B} N={N}: {ms * 1000:.0f} us per loss+grad (eager launches)")
    for k, v in rep.items():
        print("     ", k, v)
