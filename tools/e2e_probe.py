"""Diagnostic: per-pass wall time of train_model() from pinned host batches (the bench's e2e leg), with allocator counters."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests", "golden")):
    sys.path.insert(0, p)
import torch
from eeg_image_decode_b200.atms import ATMS
from eeg_image_decode_b200.train import train_model
import bench

B, K, NB = 1024, int(os.environ.get("K", "20")), 4
dev = torch.device("cuda")
torch.manual_seed(0)
model = ATMS().to(dev).train()
opt = torch.optim.AdamW(model.parameters(), lr=3e-4)
g = torch.Generator().manual_seed(1)
pin = lambda t: t.pin_memory()
eeg = [pin(torch.randn(B, 63, 250, generator=g)) for _ in range(NB)]
img = [pin(torch.nn.functional.normalize(torch.randn(B, 1024, generator=g), dim=-1)) for _ in range(NB)]
txt = [pin(torch.nn.functional.normalize(torch.randn(B, 1024, generator=g), dim=-1)) for _ in range(NB)]
lab = [pin(torch.randint(0, 1654, (B,), generator=g)) for _ in range(NB)]
gal = torch.nn.functional.normalize(torch.randn(1654, 1024, generator=g), dim=-1)
img_all, txt_all = gal.repeat_interleave(10, dim=0), gal
loader = bench.PinnedLoader([eeg[i % NB] for i in range(K)], [lab[i % NB] for i in range(K)], [txt[i % NB] for i in range(K)],
                            [img[i % NB] for i in range(K)])
reads = []
cb = (lambda i, l: reads.append(float(l[0]))) if os.environ.get("CB", "1") == "1" else None
keys = ("num_alloc_retries", "num_device_alloc", "num_device_free", "reserved_bytes.all.current", "allocated_bytes.all.current")
train_model("sub-08", model, bench.PinnedLoader(eeg[:1], lab[:1], txt[:1], img[:1]), opt, dev, txt_all, img_all, bench.Cfg(), step_callback=cb)
torch.cuda.synchronize()
from eeg_image_decode_b200 import train as _T
_orig_call = _T.GraphedTrainStep.__call__
_evs = []
def _timed_call(self, *a):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    r = _orig_call(self, *a)
    e1.record()
    _evs.append((e0, e1, time.perf_counter()))
    return r
_T.GraphedTrainStep.__call__ = _timed_call
for rep in range(int(os.environ.get("PASSES", "8"))):
    _evs.clear()
    s0 = torch.cuda.memory_stats()
    t0 = time.perf_counter()
    train_model("sub-08", model, loader, opt, dev, txt_all, img_all, bench.Cfg(), step_callback=cb)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    s1 = torch.cuda.memory_stats()
    gs = next(iter(model._gstep_cache.values())) if model._gstep_cache else None
    print(f"pass {rep}: {1e3 * dt:7.1f} ms  {B * K / dt / 1e3:6.1f} k trials/s  replays={gs.replays if gs else None} calls={gs.calls if gs else None} "
          + " ".join(f"{k.split('.')[0]}={s1.get(k, 0) - (s0.get(k, 0) if 'current' not in k else 0)}" for k in keys), flush=True)
    if rep >= 1:
        dur = [a.elapsed_time(b) for a, b, _ in _evs]
        gap = [_evs[i][1].elapsed_time(_evs[i + 1][0]) for i in range(len(_evs) - 1)]
        host = [1e3 * (_evs[i + 1][2] - _evs[i][2]) for i in range(len(_evs) - 1)]
        print(f"        step (copies+replay) ms: mean {sum(dur)/len(dur):.3f} max {max(dur):.3f} | GPU gap between steps ms: mean "
              f"{sum(gap)/len(gap):.3f} max {max(gap):.3f} | host iteration ms: mean {sum(host)/len(host):.3f} max {max(host):.3f}", flush=True)
