"""Bring-up check of the fused tcgen05 conv stack (csrc/conv_tc.cu), forward and backward, against the fp64 / fp32 CPU
oracle: conv-stack stages (debug stores on), the embedding, and every parameter gradient of one train step (dropout
off), at even, ragged and multi-wave tile counts.      python tools/gpu_conv_tc_check.py [--big]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import torch  # noqa: E402
import recipe  # noqa: E402
from oracle import atms_oracle as O  # noqa: E402
from eeg_image_decode_b200 import _lib  # noqa: E402
from eeg_image_decode_b200.atms import ATMS  # noqa: E402
from eeg_image_decode_b200.train import StepEngine  # noqa: E402

NOISE = ("enc_eeg.0.tsconv.0.bias", "enc_eeg.0.tsconv.4.bias", "encoder.encoder.attn_layers.0.attention.key_projection.bias")


def rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


ok = True
_lib.lib().eegb200_set_debug_stores(1)
for train in (False, True):
    for B, seed in ((3, 11), (8, 21), (64, 13), (7, 5)):
        sd = recipe.make_state_dict()
        x = recipe.make_eeg(B, seed=seed)
        sid = torch.full((B,), 8)
        ref = O.atms_forward(sd, x, sid, train=train, dtype=torch.float64)
        m = ATMS()
        m.load_state_dict(sd)
        m = m.cuda()
        m.dropout_p = [0.0] * 8
        m.train(train)
        out = m.encode(x.cuda(), sid.cuda(), train=train, seed=1)
        torch.cuda.synchronize()
        res = {
            "y1": rel(m.ws_tensor("y1").reshape(B, 36, 63, 40).permute(0, 3, 2, 1), ref["y1"]),
            "a1": rel(m.ws_tensor("a1").reshape(B, 36, 63, 40).permute(0, 3, 2, 1), ref["a1"]),
            "y2": rel(m.ws_tensor("y2").reshape(B, 36, 40).permute(0, 2, 1), ref["y2"].reshape(B, 40, 36)),
            "out": rel(out, ref["out"]),
        }
        good = res["y1"] < 1e-3 and res["a1"] < 2e-3 and res["y2"] < 2e-3 and res["out"] < 1e-3
        ok &= good
        print(f"fwd train={int(train)} B={B:3d}: " + " ".join(f"{k}={v:.2e}" for k, v in res.items()) + ("  ok" if good else "  MISMATCH"), flush=True)
_lib.lib().eegb200_set_debug_stores(0)

sizes = [(3, 11), (8, 21), (7, 5), (64, 13), (100, 3)] + ([(1024, 1234)] if "--big" in sys.argv else [])
for B, seed in sizes:
    sd = recipe.make_state_dict()
    x = recipe.make_eeg(B, seed=seed)
    sid = torch.full((B,), 8)
    img = recipe.make_targets(B, seed=seed, tag="img")
    txt = recipe.make_targets(B, seed=seed, tag="txt")
    lo, grads, r = O.train_step(sd, {}, x, sid, img, txt, 1)
    m = ATMS()
    m.load_state_dict(recipe.make_state_dict())
    m = m.cuda().train()
    m.dropout_p = [0.0] * 8
    loss, feats = StepEngine(m, None).step(x.cuda(), sid.cuda(), img.cuda(), txt.cuda(), use_shared=False)
    torch.cuda.synchronize()
    errs = {k: rel(m.grad_view(k), g) for k, g in grads.items() if g is not None and k not in NOISE}
    conv = {k.replace("enc_eeg.0.", ""): v for k, v in errs.items() if k.startswith("enc_eeg.0.tsconv")}
    worst = max(errs, key=errs.get)
    good = max(errs.values()) < 3e-2 and abs(loss[0].item() - lo.item()) < 3e-3 * abs(lo.item())
    noise = max(m.grad_view(k).abs().max().item() for k in NOISE)
    good = good and noise < 1e-2
    ok &= good
    print(f"bwd B={B:4d}: loss {loss[0].item():.5f} (oracle {lo.item():.5f}) conv grads " +
          " ".join(f"{k}={v:.1e}" for k, v in conv.items()) + f" | worst {worst}={errs[worst]:.1e} noise={noise:.1e}" +
          ("  ok" if good else "  MISMATCH"), flush=True)
print("CONV_TC CHECK", "PASS" if ok else "FAIL")
sys.exit(0 if ok else 1)
