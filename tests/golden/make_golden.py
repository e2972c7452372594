"""Generate the committed golden fixtures by running the UNMODIFIED reference (CPU, fp32).

Run in the build container only:   python tests/golden/make_golden.py
Writes tests/golden/*.npz.  Inputs/weights come from tests/golden/recipe.py so that the tests can
rebuild them without the reference.  torch version is recorded in each file.
"""
import os
import random
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import recipe  # noqa: E402
from ref_import import import_reference  # noqa: E402

torch.set_num_threads(8)
torch.use_deterministic_algorithms(False)
R = import_reference()


def ref_model(seed=0, p_drop=None):
    m = R.ATMS()
    sd = recipe.make_state_dict(seed)
    missing = m.load_state_dict(sd, strict=True)
    assert not missing.missing_keys and not missing.unexpected_keys
    if p_drop is not None:
        for mod in m.modules():
            if isinstance(mod, torch.nn.Dropout):
                mod.p = p_drop
    return m


def save(name, **arrs):
    out = {}
    for k, v in arrs.items():
        if isinstance(v, torch.Tensor):
            v = v.detach().cpu().numpy()
        out[k] = np.asarray(v)
    out["torch_version"] = np.asarray(torch.__version__)
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path) // 1024, "KiB")


def capture(m):
    """forward hooks on the reference submodules -> stage outputs."""
    cap = {}
    hs = []
    hs.append(m.encoder.enc_embedding.register_forward_hook(lambda mod, i, o: cap.__setitem__("h0", o.detach())))
    hs.append(m.encoder.encoder.attn_layers[0].attention.register_forward_hook(
        lambda mod, i, o: cap.__setitem__("attn_out", o[0].detach())))
    hs.append(m.encoder.encoder.attn_layers[0].norm1.register_forward_hook(lambda mod, i, o: cap.__setitem__("x1", o.detach())))
    hs.append(m.encoder.register_forward_hook(lambda mod, i, o: cap.__setitem__("enc", o.detach())))
    hs.append(m.enc_eeg[0].tsconv[1].register_forward_hook(lambda mod, i, o: cap.__setitem__("y1", o.detach())))
    hs.append(m.enc_eeg[0].tsconv[4].register_forward_hook(lambda mod, i, o: cap.__setitem__("y2", o.detach())))
    hs.append(m.enc_eeg.register_forward_hook(lambda mod, i, o: cap.__setitem__("feat", o.detach())))
    hs.append(m.proj_eeg[0].register_forward_hook(lambda mod, i, o: cap.__setitem__("z1", o.detach().clone())))
    return cap, hs


# ---------------------------------------------------------------- A/B: eval forward
def gen_eval_forward():
    m = ref_model().eval()
    cap, hs = capture(m)
    x = recipe.make_eeg(2, seed=11)
    sid = torch.tensor([8, 3])
    with torch.no_grad():
        out = m(x, sid)
    save("eval_forward_b2", out=out, sid=sid, **{k: v for k, v in cap.items()})
    for h in hs:
        h.remove()
    # odd batch + shared-token branch (an id >= 10 switches the WHOLE batch to the shared token)
    x = recipe.make_eeg(3, seed=12)
    sid = torch.tensor([10, 2, 5])
    with torch.no_grad():
        out = m(x, sid)
    save("eval_forward_b3_shared", out=out, sid=sid)
    # larger batch, output only (kept small: 64 x 1024 floats)
    x = recipe.make_eeg(64, seed=13)
    sid = torch.full((64,), 8)
    with torch.no_grad():
        out = m(x, sid)
    save("eval_forward_b64", out=out, sid=sid)


# ---------------------------------------------------------------- C: train step, dropout p = 0
def gen_train_step():
    B = 8
    m = ref_model(p_drop=0.0).train()
    cap, hs = capture(m)
    opt = torch.optim.AdamW(m.parameters(), lr=3e-4)
    x = recipe.make_eeg(B, seed=21)
    sid = torch.full((B,), 8)
    img = recipe.make_targets(B, seed=21, tag="img")
    txt = recipe.make_targets(B, seed=21, tag="txt")
    arrs = {}
    for step in (1, 2):
        opt.zero_grad()
        out = m(x, sid).float()
        il = m.loss_func(out, img, m.logit_scale)
        tl = m.loss_func(out, txt, m.logit_scale)
        loss = 0.99 * il + 0.01 * tl
        loss.backward()
        if step == 1:
            arrs["out1"] = out.detach().clone()
            arrs["y1_1"] = cap["y1"].clone()
            arrs["y2_1"] = cap["y2"].clone()
            arrs["feat_1"] = cap["feat"].clone()
            for k, p in m.named_parameters():
                if p.grad is None:
                    arrs["gradnone/" + k] = np.asarray(1)
                else:
                    arrs["graddig/" + k] = recipe.digest(p.grad)
                    if p.grad.numel() <= 4096:
                        arrs["grad/" + k] = p.grad.detach().clone()
        arrs[f"loss{step}"] = loss.detach()
        arrs[f"img_loss{step}"] = il.detach()
        arrs[f"txt_loss{step}"] = tl.detach()
        opt.step()
        for k, v in m.state_dict().items():
            if v.dtype.is_floating_point and not k.endswith(".pe"):
                arrs[f"paramdig{step}/" + k] = recipe.digest(v)
        for k in ("enc_eeg.0.tsconv.2.running_mean", "enc_eeg.0.tsconv.2.running_var",
                  "enc_eeg.0.tsconv.5.running_mean", "enc_eeg.0.tsconv.5.running_var"):
            arrs[f"bn{step}/" + k] = m.state_dict()[k].clone()
    save("train_step_b8", **arrs)
    for h in hs:
        h.remove()


# ---------------------------------------------------------------- D: ClipLoss
def gen_cliploss():
    g = torch.Generator().manual_seed(5)
    arrs = {}
    for B in (1, 5, 16):
        E = (torch.randn(B, 1024, generator=g) * 1.0).requires_grad_(True)
        T = recipe.make_targets(B, seed=30 + B)
        s = torch.tensor(2.659, requires_grad=True)
        loss = R.ClipLoss()(E, T, s)
        loss.backward()
        arrs[f"E{B}"], arrs[f"T{B}"] = E.detach(), T
        arrs[f"loss{B}"], arrs[f"dE{B}"], arrs[f"ds{B}"] = loss.detach(), E.grad, s.grad
    save("cliploss", **arrs)


# ---------------------------------------------------------------- E: train_model / evaluate_model host semantics
class _Loader:
    """Mimics DataLoader over the 6-tuples of eegdatasets_leaveone.py:375."""

    def __init__(self, eeg, labels, txt, img, bs):
        self.eeg, self.labels, self.txt, self.img, self.bs = eeg, labels, txt, img, bs

    def __iter__(self):
        n = self.eeg.shape[0] // self.bs * self.bs
        for i in range(0, n, self.bs):
            sl = slice(i, i + self.bs)
            yield (self.eeg[sl], self.labels[sl], ["t"] * self.bs, self.txt[sl], ["i"] * self.bs, self.img[sl])


class _Cfg:
    epochs = 1
    insubject = True
    encoder_type = "ATMS"


def gen_loops():
    n_cls, n_per = 40, 10
    m = ref_model(p_drop=0.0)
    opt = torch.optim.AdamW(m.parameters(), lr=3e-4)
    n = 24
    eeg = recipe.make_eeg(n, seed=41)
    labels = recipe.make_labels(n, n_cls, seed=41)
    img_all = recipe.make_targets(n_cls * n_per, seed=41, tag="img_all")      # (n_cls*10, 1024); train_model uses [::10]
    txt_all = recipe.make_targets(n_cls, seed=41, tag="txt_all")
    img = img_all[labels * n_per]
    txt = txt_all[labels]
    loader = _Loader(eeg, labels, txt, img, 8)
    avg_loss, acc, feats = R.train_model("sub-08", m, loader, opt, torch.device("cpu"), txt_all, img_all, _Cfg())
    arrs = dict(train_avg_loss=np.float64(avg_loss), train_acc=np.float64(acc), train_feats=feats.detach())

    # evaluate_model: 20 test trials over 200 classes, batch_size 1 like the reference's test loader
    n_cls_t = 200
    m.eval()
    teeg = recipe.make_eeg(20, seed=42)
    tlabels = recipe.make_labels(20, n_cls_t, seed=42)
    timg_all = recipe.make_targets(n_cls_t, seed=42, tag="timg")
    ttxt_all = recipe.make_targets(n_cls_t, seed=42, tag="ttxt")
    # make the task non-trivial but learnable-free: plant the model's own embedding direction for half of the classes
    with torch.no_grad():
        e = m(teeg, torch.full((20,), 8))
    timg_all = timg_all.clone()
    for i in range(0, 20, 2):
        timg_all[tlabels[i]] = torch.nn.functional.normalize(e[i] + 3.0 * timg_all[tlabels[i]] * e[i].norm(), dim=-1)
    tl = _Loader(teeg, tlabels, ttxt_all[tlabels], timg_all[tlabels], 1)
    for k in (200, 100, 50, 10, 4, 2):
        random.seed(1000 + k)
        loss, acc, top5 = R.evaluate_model("sub-08", m, tl, torch.device("cpu"), ttxt_all, timg_all, k, _Cfg())
        arrs[f"eval_k{k}"] = np.asarray([loss, acc, top5], dtype=np.float64)
    arrs["timg_all"] = timg_all
    save("loops", **arrs)


# ---------------------------------------------------------------- F: joint-subject variant (per-subject value embeddings)
def _set_dropout(m, p):
    for mod in m.modules():
        if isinstance(mod, torch.nn.Dropout):
            mod.p = p


def gen_joint():
    from ref_import import import_reference_joint
    J = import_reference_joint()

    def joint_model():
        m = J.ATMS(joint_train=True)
        r = m.load_state_dict(recipe.make_joint_state_dict(), strict=True)
        assert not r.missing_keys and not r.unexpected_keys
        return m

    arrs = {}
    # eval forward, mixed subjects in one batch (Embed.py:144 picks the Linear per trial)
    m = joint_model().eval()
    cap = {}
    h = m.encoder.enc_embedding.register_forward_hook(lambda mod, i, o: cap.__setitem__("h0", o.detach()))
    x = recipe.make_eeg(5, seed=51)
    sid = torch.tensor([3, 0, 3, 9, 0])
    with torch.no_grad():
        out = m(x, sid)
    arrs["eval_out"], arrs["eval_sid"], arrs["eval_h0"] = out, sid, cap["h0"]
    h.remove()
    # two AdamW steps, dropout p = 0, mixed subjects: only the value embeddings of subjects in the batch get gradients
    B = 8
    m = joint_model().train()
    _set_dropout(m, 0.0)
    opt = torch.optim.AdamW(m.parameters(), lr=3e-4)
    x = recipe.make_eeg(B, seed=52)
    sid = torch.tensor([2, 2, 5, 5, 5, 2, 7, 7])
    img = recipe.make_targets(B, seed=52, tag="img")
    txt = recipe.make_targets(B, seed=52, tag="txt")
    arrs["train_sid"] = sid
    for step in (1, 2):
        opt.zero_grad()
        out = m(x, sid).float()
        loss = 0.99 * m.loss_func(out, img, m.logit_scale) + 0.01 * m.loss_func(out, txt, m.logit_scale)
        loss.backward()
        if step == 1:
            arrs["out1"] = out.detach().clone()
            for k, p in m.named_parameters():
                if p.grad is None:
                    arrs["gradnone/" + k] = np.asarray(1)
                else:
                    arrs["graddig/" + k] = recipe.digest(p.grad)
        arrs[f"loss{step}"] = loss.detach()
        opt.step()
        for k, v in m.state_dict().items():
            if v.dtype.is_floating_point and not k.endswith(".pe"):
                arrs[f"paramdig{step}/" + k] = recipe.digest(v)
    # the script's own train_model / evaluate_model (every trial carries the id parsed from `sub`)
    n_cls, n_per = 40, 10
    m = joint_model()
    _set_dropout(m, 0.0)
    opt = torch.optim.AdamW(m.parameters(), lr=3e-4)
    n = 16
    eeg = recipe.make_eeg(n, seed=53)
    labels = recipe.make_labels(n, n_cls, seed=53)
    img_all = recipe.make_targets(n_cls * n_per, seed=53, tag="img_all")
    txt_all = recipe.make_targets(n_cls, seed=53, tag="txt_all")
    loader = _Loader(eeg, labels, txt_all[labels], img_all[labels * n_per], 8)
    avg_loss, acc, feats = J.train_model("sub-03", m, loader, opt, torch.device("cpu"), txt_all, img_all, _Cfg())
    arrs.update(train_avg_loss=np.float64(avg_loss), train_acc=np.float64(acc), train_feats=feats.detach())
    for k, v in m.state_dict().items():
        if v.dtype.is_floating_point and not k.endswith(".pe"):
            arrs["loop_paramdig/" + k] = recipe.digest(v)
    n_cls_t = 200
    teeg = recipe.make_eeg(10, seed=54)
    tlabels = recipe.make_labels(10, n_cls_t, seed=54)
    timg_all = recipe.make_targets(n_cls_t, seed=54, tag="timg")
    ttxt_all = recipe.make_targets(n_cls_t, seed=54, tag="ttxt")
    tl = _Loader(teeg, tlabels, ttxt_all[tlabels], timg_all[tlabels], 1)
    for k in (200, 10):
        random.seed(2000 + k)
        arrs[f"eval_k{k}"] = np.asarray(J.evaluate_model("sub-03", m, tl, torch.device("cpu"), ttxt_all, timg_all, k, _Cfg()),
                                        dtype=np.float64)
    save("joint", **arrs)


# ---------------------------------------------------------------- G: reconstruction-training variant (MSE + InfoNCE)
def gen_reconstruction():
    from ref_import import import_reference_reconstruction
    G = import_reference_reconstruction()

    def model():
        m = G.ATMS()
        r = m.load_state_dict(recipe.make_state_dict(), strict=True)
        assert not r.missing_keys and not r.unexpected_keys
        _set_dropout(m, 0.0)
        return m

    arrs = {}
    # one step of the loss mix of ATMS_reconstruction.py:224-228 (alpha = 0.90) with gradient digests
    B = 8
    m = model().train()
    x = recipe.make_eeg(B, seed=61)
    sid = torch.full((B,), 8)
    img = recipe.make_targets(B, seed=61, tag="img")
    out = m(x, sid).float()
    mse = torch.nn.MSELoss()(out, img)
    il = m.loss_func(out, img, m.logit_scale)
    loss = 0.90 * mse * 10 + (1 - 0.90) * il * 10
    loss.backward()
    arrs.update(step_out=out.detach(), step_loss=loss.detach(), step_mse=mse.detach(), step_img_loss=il.detach())
    for k, p in m.named_parameters():
        if p.grad is not None:
            arrs["graddig/" + k] = recipe.digest(p.grad)
    # the script's own train_model / evaluate_model
    n_cls, n_per = 40, 10
    m = model()
    opt = torch.optim.AdamW(m.parameters(), lr=3e-4)
    n = 24
    eeg = recipe.make_eeg(n, seed=62)
    labels = recipe.make_labels(n, n_cls, seed=62)
    img_all = recipe.make_targets(n_cls * n_per, seed=62, tag="img_all")
    txt_all = recipe.make_targets(n_cls, seed=62, tag="txt_all")
    loader = _Loader(eeg, labels, txt_all[labels], img_all[labels * n_per], 8)
    avg_loss, acc, feats = G.train_model("sub-08", m, loader, opt, torch.device("cpu"), txt_all, img_all, _Cfg())
    arrs.update(train_avg_loss=np.float64(avg_loss), train_acc=np.float64(acc), train_feats=feats.detach())
    for k, v in m.state_dict().items():
        if v.dtype.is_floating_point and not k.endswith(".pe"):
            arrs["loop_paramdig/" + k] = recipe.digest(v)
    n_cls_t = 200
    teeg = recipe.make_eeg(10, seed=63)
    tlabels = recipe.make_labels(10, n_cls_t, seed=63)
    timg_all = recipe.make_targets(n_cls_t, seed=63, tag="timg")
    ttxt_all = recipe.make_targets(n_cls_t, seed=63, tag="ttxt")
    tl = _Loader(teeg, tlabels, ttxt_all[tlabels], timg_all[tlabels], 1)
    random.seed(3000)
    arrs["eval_k200"] = np.asarray(G.evaluate_model("sub-08", m, tl, torch.device("cpu"), ttxt_all, timg_all, 200, _Cfg()),
                                   dtype=np.float64)
    save("reconstruction", **arrs)


if __name__ == "__main__":
    gens = {"eval_forward": gen_eval_forward, "train_step": gen_train_step, "cliploss": gen_cliploss, "loops": gen_loops,
            "joint": gen_joint, "reconstruction": gen_reconstruction}
    for name in (sys.argv[1:] or list(gens)):
        gens[name]()
