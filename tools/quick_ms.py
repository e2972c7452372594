"""ms per captured training step at B=1024 (device-resident inputs), for quick A/B runs of env toggles."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from eeg_image_decode_b200.atms import ATMS
from eeg_image_decode_b200.train import GraphedTrainStep, StepEngine
torch.manual_seed(0)
B = 1024
dev = torch.device("cuda")
m = ATMS().to(dev).train()
eng = StepEngine(m, torch.optim.AdamW(m.parameters(), lr=3e-4))
g = torch.Generator().manual_seed(1)
xs = [torch.randn(B, 63, 250, generator=g).to(dev) for _ in range(4)]
img = torch.nn.functional.normalize(torch.randn(B, 1024, generator=g), dim=-1).to(dev)
txt = torch.nn.functional.normalize(torch.randn(B, 1024, generator=g), dim=-1).to(dev)
lab = torch.randint(0, 1654, (B,), generator=g).to(dev)
gal = torch.nn.functional.normalize(torch.randn(1654, 1024, generator=g), dim=-1).to(dev)
sid = torch.full((B,), 8, device=dev)
step = GraphedTrainStep(eng, gal, False)
for i in range(8):
    step(xs[i % 4], sid, img, txt, lab)
torch.cuda.synchronize()
best = 1e9
for rep in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(30):
        loss, feats, ok = step(xs[i % 4], sid, img, txt, lab)
    e1.record()
    torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1) / 30)
print(f"{' '.join(k + '=' + v for k, v in os.environ.items() if k.startswith('EEGB200_')) or 'default'}: {best:.4f} ms/step "
      f"({B / best:.1f} k trials/s) loss {loss[0].item():.4f} graph={step.graph is not None}")
