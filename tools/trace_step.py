"""Kernel timeline of the captured training step (torch.profiler / CUPTI): per-stream start/end of every kernel in a few
graph replays, written as a compact JSON for critical-path analysis (nsys is not in the image)."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from torch.profiler import ProfilerActivity, profile
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
step = GraphedTrainStep(eng, gal, False)
for i in range(6):
    step(x, sid, img, txt, lab)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for i in range(3):
        step(x, sid, img, txt, lab)
    torch.cuda.synchronize()
out = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/trace.json"
prof.export_chrome_trace(out + ".full")
ev = json.load(open(out + ".full"))["traceEvents"]
k = [{"name": e["name"][:90], "ts": e["ts"], "dur": e["dur"], "stream": e.get("args", {}).get("stream"),
      "grid": e.get("args", {}).get("grid"), "block": e.get("args", {}).get("block")}
     for e in ev if e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset")]
k.sort(key=lambda e: e["ts"])
json.dump(k, open(out, "w"))
os.remove(out + ".full")
print("kernels", len(k), "span ms", (k[-1]["ts"] + k[-1]["dur"] - k[0]["ts"]) / 1e3)
