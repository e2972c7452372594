"""Round-2 bring-up check of the experimental fused tcgen05 conv forward (csrc/conv_tc.cu): runs the encoder with
EEGB200_CONV_TC=1 (set here, before the library is loaded) and compares the conv-stack stages and the embedding against
the fp64 oracle, eval and train mode, even and ragged tile counts.   python tools/gpu_conv_tc_check.py"""
import os
import sys

os.environ["EEGB200_CONV_TC"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import torch  # noqa: E402
import recipe  # noqa: E402
from oracle import atms_oracle as O  # noqa: E402
from eeg_image_decode_b200.atms import ATMS  # noqa: E402


def rel(a, b):
    a, b = a.double().cpu(), b.double()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


ok = True
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
        print(f"train={int(train)} B={B:3d}: " + " ".join(f"{k}={v:.2e}" for k, v in res.items()) + ("  ok" if good else "  MISMATCH"))
print("CONV_TC CHECK", "PASS" if ok else "FAIL")
sys.exit(0 if ok else 1)
