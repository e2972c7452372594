"""GPU parity of the joint-subject and reconstruction-training variants through the C ABI,
against the CPU oracle and the golden vectors produced by the unmodified reference scripts
(Retrieval/ATMS_retrieval_joint_train.py, Generation/ATMS_reconstruction.py).  Tolerances as in test_gpu_parity.py:
exact-fp32 verification backend ~1e-5, product path (tcgen05 TF32) embeddings within 1e-3 relative."""
import os
import random

import pytest
import torch

import recipe
from oracle import atms_oracle as O
from test_gpu_parity import NOISE_GRADS, _Cfg, _Loader, load, make_model, rel_l2, rows_rel, tok

pytestmark = pytest.mark.gpu
VP = O.JOINT_VALUE_PREFIX


@pytest.fixture(scope="module")
def lib():
    from eeg_image_decode_b200 import _lib
    _lib.lib()
    return _lib


def make_joint(p_drop=None):
    from eeg_image_decode_b200.joint import ATMS
    m = ATMS(joint_train=True)
    r = m.load_state_dict(recipe.make_joint_state_dict(), strict=True)
    assert not r.missing_keys and not r.unexpected_keys
    m = m.cuda()
    if p_drop is not None:
        m.dropout_p = [p_drop] * 8
    return m


# ---------------------------------------------------------------- joint-subject variant
@pytest.mark.parametrize("backend", [1, 0])
def test_joint_eval_forward_mixed_subjects(lib, backend):
    lib.set_gemm_backend(backend)
    try:
        g = load("joint")
        m = make_joint().eval()
        sid = torch.as_tensor(g["eval_sid"])                  # [3, 0, 3, 9, 0]: the batch is regrouped by subject inside
        x = recipe.make_eeg(5, seed=51)
        out = m.encode(x.cuda(), sid.cuda(), train=False)
        assert rows_rel(out, torch.as_tensor(g["eval_out"])) < (2e-5 if backend == 1 else 1e-3)
        # the workspace holds the regrouped batch: row j is trial order[j]
        order = sorted(range(5), key=sid.tolist().__getitem__)
        h0 = tok(m.ws_tensor("h0"), 5).cpu()
        ref = torch.as_tensor(g["eval_h0"])[order]
        assert (h0 - ref).abs().max().item() < (3e-5 if backend == 1 else 4e-3) * max(1.0, ref.abs().max().item())
        assert m._last_subjects == [0, 3, 9]
        # a batch of one subject takes the no-regroup path and must agree with the mixed batch row by row
        x3 = x[[0, 2]].cuda()
        out3 = m.encode(x3, torch.tensor([3, 3]).cuda(), train=False, known_subject=3)
        assert rows_rel(out3, out[[0, 2]]) < 1e-5
    finally:
        lib.set_gemm_backend(0)


def test_joint_unknown_subject_raises_keyerror(lib):
    from eeg_image_decode_b200.joint import train_model
    m = make_joint().eval()
    with pytest.raises(KeyError):       # the reference: self.value_embedding['10'] (Embed.py:144)
        m.encode(recipe.make_eeg(2).cuda(), torch.tensor([1, 10]).cuda(), train=False)
    with pytest.raises(KeyError):
        train_model("sub-10", m, [], None, torch.device("cuda"), torch.zeros(4, 1024), torch.zeros(40, 1024), _Cfg())


@pytest.mark.parametrize("backend", [1, 0])
def test_joint_train_steps_gradients_and_update(lib, backend):
    from eeg_image_decode_b200.train import StepEngine
    lib.set_gemm_backend(backend)
    try:
        g = load("joint")
        B = 8
        x = recipe.make_eeg(B, seed=52)
        sid = torch.as_tensor(g["train_sid"])                # [2, 2, 5, 5, 5, 2, 7, 7]
        img = recipe.make_targets(B, seed=52, tag="img")
        txt = recipe.make_targets(B, seed=52, tag="txt")
        sd = recipe.make_joint_state_dict()
        opt_state = {}
        m = make_joint(p_drop=0.0).train()
        eng = StepEngine(m, None)
        tol_l = 1e-4 if backend == 1 else 3e-3
        for step in (1, 2):
            lo, grads, r = O.train_step(sd, opt_state, x, sid, img, txt, step)
            loss, feats = eng.step(x.cuda(), sid.cuda(), img.cuda(), txt.cuda(), use_shared=False)
            assert abs(loss[0].item() - float(g[f"loss{step}"])) < tol_l * abs(float(g[f"loss{step}"])) * step
            assert abs(loss[0].item() - lo.item()) < tol_l * abs(lo.item()) * step
            if step > 1:
                continue
            assert rows_rel(feats, torch.as_tensor(g["out1"])) < (5e-5 if backend == 1 else 1e-3)   # caller's trial order
            tol_g = 2e-3 if backend == 1 else 3e-2
            for k, gr in grads.items():
                if gr is None or k in NOISE_GRADS:
                    continue
                e = rel_l2(m.grad_view(k), gr)
                assert e < tol_g, f"grad {k}: rel l2 {e} (backend {backend})"
            assert m._last_subjects == [2, 5, 7]
            # subjects outside the batch: zero gradient, parameters bit-identical after the update (AdamW skips them)
            init = recipe.make_joint_state_dict()
            new = m.state_dict()
            for sj in (0, 1, 3, 4, 6, 8, 9):
                assert m.grad_view(f"{VP}{sj}.weight").abs().max().item() == 0.0
                assert torch.equal(new[f"{VP}{sj}.weight"].cpu(), init[f"{VP}{sj}.weight"])
                assert torch.equal(new[f"{VP}{sj}.bias"].cpu(), init[f"{VP}{sj}.bias"])
            for sj in (2, 5, 7):
                k = f"{VP}{sj}.weight"
                d = (new[k].cpu() - sd[k]).abs()
                assert int((d > 1e-5).sum().item()) <= max(2, int((1e-3 if backend == 1 else 2e-2) * d.numel())), k
                assert not torch.equal(new[k].cpu(), init[k])
        assert m._adam_steps["ve2"] == 2 and m._adam_steps["ve0"] == 0 and m._adam_steps["main"] == 2
    finally:
        lib.set_gemm_backend(0)


def test_joint_autograd_bridge(lib):
    """reference-style usage on the joint model: grads of absent subjects stay None"""
    g = load("joint")
    B = 8
    m = make_joint(p_drop=0.0).train()
    x = recipe.make_eeg(B, seed=52).cuda()
    sid = torch.as_tensor(g["train_sid"]).cuda()
    img = recipe.make_targets(B, seed=52, tag="img").cuda()
    txt = recipe.make_targets(B, seed=52, tag="txt").cuda()
    out = m(x, sid).float()
    loss = 0.99 * m.loss_func(out, img, m.logit_scale) + 0.01 * m.loss_func(out, txt, m.logit_scale)
    loss.backward()
    assert abs(loss.item() - float(g["loss1"])) < 3e-3 * abs(float(g["loss1"]))
    named = dict(m.named_parameters())
    for k in named:
        if ("gradnone/" + k) in g:
            assert named[k].grad is None, k
    for k in (f"{VP}2.weight", f"{VP}5.bias", f"{VP}7.weight", "proj_eeg.0.weight", "logit_scale"):
        dig = recipe.digest(named[k].grad.cpu())
        ref = torch.as_tensor(g["graddig/" + k])
        assert abs(dig[0] - ref[0]).item() < 3e-2 * ref[0].item(), k


def test_joint_train_model_and_evaluate_model_loops(lib):
    from eeg_image_decode_b200.joint import evaluate_model, train_model
    g = load("joint")
    n_cls, n_per, n = 40, 10, 16
    m = make_joint(p_drop=0.0)
    opt = torch.optim.AdamW(m.parameters(), lr=3e-4)
    eeg = recipe.make_eeg(n, seed=53)
    labels = recipe.make_labels(n, n_cls, seed=53)
    img_all = recipe.make_targets(n_cls * n_per, seed=53, tag="img_all")
    txt_all = recipe.make_targets(n_cls, seed=53, tag="txt_all")
    loader = _Loader(eeg, labels, txt_all[labels], img_all[labels * n_per], 8)
    avg_loss, acc, feats = train_model("sub-03", m, loader, opt, torch.device("cuda"), txt_all, img_all, _Cfg())
    assert abs(avg_loss - float(g["train_avg_loss"])) < 5e-3 * float(g["train_avg_loss"])
    assert rows_rel(feats[:8], torch.as_tensor(g["train_feats"])[:8]) < 1e-3
    assert rows_rel(feats, torch.as_tensor(g["train_feats"])) < 2e-2
    assert abs(acc - float(g["train_acc"])) <= 1.0 / n + 1e-9
    # only subject 3's value embedding moved; its optimiser state is published, the others' is not
    init = recipe.make_joint_state_dict()
    new = m.state_dict()
    assert not torch.equal(new[f"{VP}3.weight"].cpu(), init[f"{VP}3.weight"])
    assert torch.equal(new[f"{VP}4.weight"].cpu(), init[f"{VP}4.weight"])
    named = dict(m.named_parameters())
    assert named[f"{VP}3.weight"] in opt.state and named[f"{VP}4.weight"] not in opt.state
    for k in (f"{VP}3.weight", "proj_eeg.0.weight"):
        dig = recipe.digest(new[k].cpu())
        assert (dig - torch.as_tensor(g["loop_paramdig/" + k]))[3:].abs().max().item() <= 2.5 * 3e-4 * 2, k
    teeg = recipe.make_eeg(10, seed=54)
    tlabels = recipe.make_labels(10, 200, seed=54)
    timg_all = recipe.make_targets(200, seed=54, tag="timg")
    ttxt_all = recipe.make_targets(200, seed=54, tag="ttxt")
    tl = _Loader(teeg, tlabels, ttxt_all[tlabels], timg_all[tlabels], 1)
    for k in (200, 10):
        random.seed(2000 + k)
        loss, a, t5 = evaluate_model("sub-03", m, tl, torch.device("cuda"), ttxt_all, timg_all, k, _Cfg())
        ref = g[f"eval_k{k}"]
        assert abs(loss - ref[0]) < 1e-3
        assert abs(a - ref[1]) <= 0.1 + 1e-9, (k, a, ref)
        assert abs(t5 - ref[2]) <= 0.1 + 1e-9, (k, t5, ref)


def test_joint_cuda_graph_step_matches_eager(lib):
    from eeg_image_decode_b200.train import GraphedTrainStep, StepEngine
    B = 8
    xs = [recipe.make_eeg(B, seed=90 + i).cuda() for i in range(4)]
    sid = torch.full((B,), 6).cuda()
    img = recipe.make_targets(B, seed=90, tag="img").cuda()
    txt = recipe.make_targets(B, seed=90, tag="txt").cuda()
    lab = recipe.make_labels(B, 50, seed=90).cuda()
    gal = recipe.make_targets(50, seed=90, tag="gal").cuda()
    outs = {}
    for graphed in (False, True):
        m = make_joint(p_drop=0.0).train()
        gs = GraphedTrainStep(StepEngine(m, None), gal, use_shared=False, enabled=graphed, known_subject=6)
        losses = [gs(xs[i], sid, img, txt, lab)[0][0].item() for i in range(4)]
        assert (gs.graph is not None) == graphed
        outs[graphed] = (losses, dict(m._adam_steps))
    assert outs[True][1] == outs[False][1] and outs[True][1]["ve6"] == 4 and outs[True][1]["ve5"] == 0
    for a, b in zip(outs[True][0], outs[False][0]):
        assert abs(a - b) < 3e-3 * abs(b), outs


# ---------------------------------------------------------------- reconstruction-training variant
def test_mse_kernel(lib):
    gen = torch.Generator().manual_seed(4)
    for B, n_total in ((1, 1), (7, 7), (64, 256)):
        e = torch.randn(B, 1024, generator=gen)
        t = torch.randn(B, 1024, generator=gen)
        d0 = torch.randn(B, 1024, generator=gen)
        want = 9.0 * ((e.double() - t.double()) ** 2).sum() / (n_total * 1024)
        want_g = d0.double() + 9.0 * 0.5 * 2.0 * (e.double() - t.double()) / (n_total * 1024)
        loss = torch.tensor([0.25, 0.0, 0.0], device="cuda")
        d = d0.cuda()
        lib.mse(e.cuda(), t.cuda(), n_total, 9.0, 0.5, loss=loss[0:1], loss_term=loss[2:3], d_eeg=d)
        assert abs(loss[0].item() - 0.25 - want.item()) < 1e-5 * max(1.0, want.item())
        assert abs(loss[2].item() - want.item()) < 1e-5 * max(1.0, want.item())
        assert (d.cpu().double() - want_g).abs().max().item() < 1e-6
        lib.mse(e.cuda(), t.cuda(), n_total, 9.0, 1.0, loss=loss[1:2])          # loss only
        assert abs(loss[1].item() - want.item()) < 1e-5 * max(1.0, want.item())


@pytest.mark.parametrize("backend", [1, 0])
def test_reconstruction_step_matches_reference(lib, backend):
    from eeg_image_decode_b200.train import StepEngine
    lib.set_gemm_backend(backend)
    try:
        g = load("reconstruction")
        B = 8
        x = recipe.make_eeg(B, seed=61)
        sid = torch.full((B,), 8)
        img = recipe.make_targets(B, seed=61, tag="img")
        sd = recipe.make_state_dict()
        lo, grads, r = O.train_step(sd, {}, x, sid, img, img, 1, alpha=0.90, variant="reconstruction")
        m = make_model(p_drop=0.0).train()
        eng = StepEngine(m, None, 0.90, "reconstruction")
        loss, feats = eng.step(x.cuda(), sid.cuda(), img.cuda(), None, use_shared=False)
        tol_l = 1e-4 if backend == 1 else 3e-3
        assert rows_rel(feats, torch.as_tensor(g["step_out"])) < (5e-5 if backend == 1 else 1e-3)
        assert abs(loss[0].item() - float(g["step_loss"])) < tol_l * abs(float(g["step_loss"]))
        assert abs(loss[1].item() - float(g["step_img_loss"])) < tol_l * abs(float(g["step_img_loss"]))
        assert abs(loss[2].item() - 9.0 * float(g["step_mse"])) < tol_l * 9.0 * float(g["step_mse"])
        tol_g = 2e-3 if backend == 1 else 3e-2
        for k, gr in grads.items():
            if gr is None or k in NOISE_GRADS:
                continue
            e = rel_l2(m.grad_view(k), gr)
            assert e < tol_g, f"grad {k}: rel l2 {e} (backend {backend})"
            dig = recipe.digest(m.grad_view(k).cpu())
            ref = torch.as_tensor(g["graddig/" + k])
            assert abs(dig[0] - ref[0]).item() < tol_g * ref[0].item() + 1e-7, k
    finally:
        lib.set_gemm_backend(0)


def test_reconstruction_loops_and_feature_export(lib, tmp_path):
    from eeg_image_decode_b200.reconstruction import evaluate_model, get_eegfeatures, train_model
    g = load("reconstruction")
    n_cls, n_per, n = 40, 10, 24
    m = make_model(p_drop=0.0)
    opt = torch.optim.AdamW(m.parameters(), lr=3e-4)
    eeg = recipe.make_eeg(n, seed=62)
    labels = recipe.make_labels(n, n_cls, seed=62)
    img_all = recipe.make_targets(n_cls * n_per, seed=62, tag="img_all")
    txt_all = recipe.make_targets(n_cls, seed=62, tag="txt_all")
    loader = _Loader(eeg, labels, txt_all[labels], img_all[labels * n_per], 8)
    avg_loss, acc, feats = train_model("sub-08", m, loader, opt, torch.device("cuda"), txt_all, img_all, _Cfg())
    assert abs(avg_loss - float(g["train_avg_loss"])) < 5e-3 * float(g["train_avg_loss"])
    assert rows_rel(feats[:8], torch.as_tensor(g["train_feats"])[:8]) < 1e-3
    assert rows_rel(feats, torch.as_tensor(g["train_feats"])) < 2e-2
    assert abs(acc - float(g["train_acc"])) <= 1.0 / n + 1e-9
    for k in ("proj_eeg.0.weight", "enc_eeg.0.tsconv.4.weight"):
        dig = recipe.digest(m.state_dict()[k].cpu())
        assert (dig - torch.as_tensor(g["loop_paramdig/" + k]))[3:].abs().max().item() <= 2.5 * 3e-4 * 3, k
    teeg = recipe.make_eeg(10, seed=63)
    tlabels = recipe.make_labels(10, 200, seed=63)
    timg_all = recipe.make_targets(200, seed=63, tag="timg")
    ttxt_all = recipe.make_targets(200, seed=63, tag="ttxt")
    tl = _Loader(teeg, tlabels, ttxt_all[tlabels], timg_all[tlabels], 1)
    random.seed(3000)
    loss, a, t5 = evaluate_model("sub-08", m, tl, torch.device("cuda"), ttxt_all, timg_all, 200, _Cfg())
    ref = g["eval_k200"]
    assert abs(loss - ref[0]) < 5e-3 * abs(ref[0]) + 1e-4          # the MSE term dominates (B=1: ClipLoss == 0)
    assert abs(a - ref[1]) <= 0.1 + 1e-9 and abs(t5 - ref[2]) <= 0.1 + 1e-9, (a, t5, ref)
    # embedding export consumed by the diffusion prior / SDXL stage
    big = _Loader(teeg, tlabels, ttxt_all[tlabels], timg_all[tlabels], 5)
    random.seed(1)
    l2, a2, last_labels, ft = get_eegfeatures("sub-08", m, big, torch.device("cuda"), ttxt_all, timg_all, 200,
                                              out_dir=str(tmp_path))
    assert ft.shape == (10, 1024) and not ft.is_cuda and torch.equal(last_labels, tlabels[5:])
    saved = torch.load(os.path.join(str(tmp_path), "ATM_S_eeg_features_sub-08.pt"))
    assert torch.equal(saved, ft)
    direct = m.encode(teeg.cuda(), torch.full((10,), 8).cuda(), train=False)
    assert rows_rel(ft, direct) < 1e-4       # batch 5 vs 10: the (tile, part) partial sums of Y2 are red.add-ed in a different order
    want = O.reconstruction_loss(ft[:5], timg_all[tlabels][:5], m.logit_scale.detach().cpu(), 0.9)
    want2 = O.reconstruction_loss(ft[5:], timg_all[tlabels][5:], m.logit_scale.detach().cpu(), 0.9)
    assert abs(l2 - 0.5 * (want.item() + want2.item())) < 3e-3 * abs(l2)


@pytest.mark.gpu
def test_autograd_bridge_refuses_stale_activations():
    """the saved activations live in one workspace per batch size: backward() through a forward that a later forward has
    overwritten must raise instead of silently differentiating the wrong batch"""
    from eeg_image_decode_b200.atms import ATMS
    torch.manual_seed(0)
    m = ATMS().cuda().train()
    x = torch.randn(4, 63, 250, device="cuda")
    sid = torch.full((4,), 8, device="cuda")
    out1 = m(x, sid)
    out2 = m(x * 0.5, sid)
    with pytest.raises(RuntimeError, match="overwritten by a later forward"):
        out1.sum().backward()
    out2.sum().backward()                      # the latest forward still works
    assert m.enc_eeg[0].projection[0].weight.grad is not None


@pytest.mark.gpu
def test_graphed_step_captures_the_loader_batch_size_not_a_ragged_one():
    """a ragged batch arriving when the capture would happen (third call) must run eagerly and leave the capture to the
    next full batch; targets of the captured step live in the loss workspace (rounded in place) and still give the same
    loss as the eager step"""
    from eeg_image_decode_b200.atms import ATMS
    from eeg_image_decode_b200.train import GraphedTrainStep, StepEngine
    torch.manual_seed(0)
    g = torch.Generator().manual_seed(5)
    gal = torch.nn.functional.normalize(torch.randn(50, 1024, generator=g), dim=-1).cuda()

    def batch(B):
        x = torch.randn(B, 63, 250, generator=g).cuda()
        img = torch.nn.functional.normalize(torch.randn(B, 1024, generator=g), dim=-1).cuda()
        txt = torch.nn.functional.normalize(torch.randn(B, 1024, generator=g), dim=-1).cuda()
        return x, torch.full((B,), 8, device="cuda"), img, txt, torch.randint(0, 50, (B,), generator=g).cuda()

    losses = {}
    for graphed in (False, True):
        m = ATMS()
        m.load_state_dict(recipe.make_state_dict())
        m = m.cuda().train()
        m.dropout_p = [0.0] * 8
        gs = GraphedTrainStep(StepEngine(m, torch.optim.AdamW(m.parameters(), lr=3e-4)), gal, use_shared=False, enabled=graphed)
        g.manual_seed(5)
        out = []
        for B in (8, 8, 5, 8, 8, 5, 8):
            loss, feats, _ = gs(*batch(B))
            out.append(loss[0].item())
        if graphed:
            assert gs.graph is not None and gs.B == 8 and gs.replays >= 2
        losses[graphed] = out
    for a, b in zip(losses[False], losses[True]):
        assert abs(a - b) < 2e-3 * abs(a), (losses[False], losses[True])
