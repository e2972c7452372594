"""Per-stage relative error of the product (tcgen05 TF32) path against the fp64 oracle, eval mode."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import torch, recipe
from oracle import atms_oracle as O
from eeg_image_decode_b200.atms import ATMS

B = 64
sd = recipe.make_state_dict()
x = recipe.make_eeg(B, seed=13)
sid = torch.full((B,), 8)
ref = O.atms_forward(sd, x, sid, dtype=torch.float64)
m = ATMS(); m.load_state_dict(sd); m = m.cuda().eval()
out = m.encode(x.cuda(), sid.cuda(), train=False)
def rel(a, b):
    a = a.double().cpu(); b = b.double()
    return ((a - b).norm() / b.norm()).item(), ((a - b).norm(dim=-1) / b.norm(dim=-1).clamp_min(1e-30)).max().item()
tok = lambda t: t.reshape(B, 64, -1)[:, :, :250]
print("h0", rel(tok(m.ws_tensor("h0")), ref["h0"]))
print("attn_o", rel(m.ws_tensor("attn_o").reshape(B, 64, 4, 64)[..., :62].reshape(B, 64, 248), ref["attn_o"]))
print("x1", rel(tok(m.ws_tensor("x1")), ref["x1"]))
print("ffn_u", rel(m.ws_tensor("ffn_u").reshape(B, 64, 256), ref["ffn_u"]))
print("x3", rel(tok(m.ws_tensor("x3")), ref["x3"]))
print("y1", rel(m.ws_tensor("y1").reshape(B, 36, 63, 40).permute(0, 3, 2, 1), ref["y1"]))
print("y2", rel(m.ws_tensor("y2").reshape(B, 36, 40).permute(0, 2, 1), ref["y2"].reshape(B, 40, 36)))
print("feat", rel(m.ws_tensor("feat"), ref["feat"]))
print("z1", rel(m.ws_tensor("z1"), ref["z1"]))
print("out (global, worst row)", rel(out, ref["out"]))

# ---- train mode (batch-statistics BatchNorm), small batch, dropout off: the tightest parity case in the test-suite ----
for Bt, seed in ((8, 21), (8, 41), (16, 61), (64, 13)):
    xt = recipe.make_eeg(Bt, seed=seed)
    sidt = torch.full((Bt,), 8)
    reft = O.atms_forward(recipe.make_state_dict(), xt, sidt, train=True, dtype=torch.float64)
    mt = ATMS(); mt.load_state_dict(recipe.make_state_dict()); mt = mt.cuda().train(); mt.dropout_p = [0.0] * 8
    outt = mt.encode(xt.cuda(), sidt.cuda(), train=True, seed=1)
    print(f"train B={Bt} seed={seed}: x3", rel(mt.ws_tensor("x3").reshape(Bt, 64, -1)[:, :, :250], reft["x3"]),
          "feat", rel(mt.ws_tensor("feat"), reft["feat"]), "out", rel(outt, reft["out"]))
