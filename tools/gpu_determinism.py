"""Run the same train-mode forward twice (and once more after dirtying the workspace) and report the first stage that differs."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import torch, recipe
from eeg_image_decode_b200.atms import ATMS
B = 32
m = ATMS(); m.load_state_dict(recipe.make_state_dict()); m = m.cuda().train(); m.dropout_p = [0.0] * 8
x = recipe.make_eeg(B, seed=91).cuda(); sid = torch.full((B,), 8).cuda()
names = ["h0", "qkv", "attn_o", "r1", "x1", "ffn_u", "ffn_h", "r2", "x3", "y1", "a1", "y2", "feat", "z1", "z2"]
def run():
    out = m.encode(x, sid, train=True, seed=1).clone()
    snap = {n: m.ws_tensor(n).clone() for n in names}
    snap["out"] = out
    snap["bn1_sums"] = m.ws_tensor("bn1_sums").clone(); snap["bn2_sums"] = m.ws_tensor("bn2_sums").clone()
    return snap
a = run(); b = run()
m.workspace(B).fill_(77)       # dirty the whole workspace (0x4D4D4D4D floats ~ 2.1e8)
c = run()
for n in names + ["bn1_sums", "bn2_sums", "out"]:
    d1 = (a[n].double() - b[n].double()).abs().max().item(); d2 = (a[n].double() - c[n].double()).abs().max().item()
    print(f"{n:10s} run1-vs-run2 {d1:.3e}   run1-vs-dirty {d2:.3e}   scale {a[n].abs().max().item():.3e}")
