"""A few eager training steps at B=1024 for ncu captures (kernel filter / launch list chosen on the ncu command line)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from eeg_image_decode_b200 import _lib
from eeg_image_decode_b200.atms import ATMS
from eeg_image_decode_b200.train import GraphedTrainStep, StepEngine
torch.manual_seed(0)
B = 1024
dev = torch.device("cuda")
m = ATMS().to(dev).train()
eng = StepEngine(m, torch.optim.AdamW(m.parameters(), lr=3e-4))
g = torch.Generator().manual_seed(1)
x = torch.randn(B, 63, 250, generator=g).to(dev)
img = torch.nn.functional.normalize(torch.randn(B, 1024, generator=g), dim=-1).to(dev)
txt = torch.nn.functional.normalize(torch.randn(B, 1024, generator=g), dim=-1).to(dev)
lab = torch.randint(0, 1654, (B,), generator=g).to(dev)
gal = torch.nn.functional.normalize(torch.randn(1654, 1024, generator=g), dim=-1).to(dev)
sid = torch.full((B,), 8, device=dev)
step = GraphedTrainStep(eng, gal, False, enabled=False)
for i in range(int(sys.argv[1]) if len(sys.argv) > 1 else 3):
    step(x, sid, img, txt, lab)
torch.cuda.synchronize()
print("done", _lib.launch_count())
