"""Host-side plumbing of the Python mirror, exercised WITHOUT a GPU: the C-ABI entry points are replaced by recording
stubs that read / write the raw buffers they are handed (host memory here), so the tests see exactly what the library
would see -- group tables, trial order, pointer tables, AdamW segments.  No arithmetic of the product is under test
here (that is tests/test_gpu_*.py, against the oracle on a real B200); this guards the glue: trial regrouping of the
joint-subject model, gradient routing, optimiser-segment bookkeeping, the reconstruction loss composition."""
import ctypes

import numpy as np
import pytest
import torch

import recipe


def _view(ptr, n, ctype=ctypes.c_float, dtype=np.float32):
    return np.frombuffer((ctype * n).from_address(ptr), dtype=dtype)


@pytest.fixture
def fake(monkeypatch):
    from eeg_image_decode_b200 import _lib
    calls = []
    monkeypatch.setattr(torch.Tensor, "is_cuda", property(lambda s: True), raising=False)
    monkeypatch.setattr(_lib, "stream_ptr", lambda: ctypes.c_void_p(0))
    monkeypatch.setattr(torch.cuda, "is_current_stream_capturing", lambda: False)

    def groups_of(io):
        if not io.joint_value_w:
            return None
        return ([io.group_offsets[i] for i in range(io.n_groups + 1)], [io.group_subject[i] for i in range(io.n_groups)])

    def atms_forward(io, phases=7, device=None):
        x = _view(io.x, io.B * 63 * 250).reshape(io.B, 63, 250)
        sid = _view(io.subject_ids, io.B, ctypes.c_int64, np.int64)
        out = _view(io.out, io.B * 1024).reshape(io.B, 1024)
        out[:] = x[:, 0, :1] + 100.0 * sid[:, None]          # row b identifies (trial, subject id) it was computed from
        calls.append(("fwd", io.B, phases, io.train, groups_of(io), sid.copy()))

    def atms_backward(io, d_out, grads, phases=7, device=None):
        d = None if d_out is None else _view(d_out.data_ptr(), io.B * 1024).reshape(io.B, 1024).copy()
        if io.joint_value_w:
            for sj in groups_of(io)[1]:                       # mark the gradient buffers the library would write
                _view(io.joint_value_dw[sj], 1)[0] += 1.0
                _view(io.joint_value_db[sj], 1)[0] += 1.0
        else:
            _view(grads[0], 1)[0] += 1.0
        calls.append(("bwd", io.B, phases, groups_of(io), d))

    def infonce(io, phases, device=None):
        if phases & 2:
            _view(io.loss, 3)[:] = (1.0 * io.w_img + 2.0 * io.w_txt, 1.0, 2.0 if io.tgt_txt else 0.0)
            if io.d_eeg:
                e = _view(io.eeg, io.B * io.D).reshape(io.B, io.D)
                _view(io.d_eeg, io.B * io.D).reshape(io.B, io.D)[:] = e     # "gradient" = the embedding row itself
        calls.append(("infonce", io.B, io.N, phases, round(io.w_img, 6), round(io.w_txt, 6), bool(io.tgt_txt)))

    def mse(eeg, tgt, n_total_rows, weight, grad_out, loss=None, loss_term=None, d_eeg=None):
        if loss is not None:
            loss += 7.0
        if loss_term is not None:
            loss_term += 7.0
        calls.append(("mse", tuple(eeg.shape), n_total_rows, round(weight, 6), grad_out, d_eeg is not None))

    def adamw_step(p, g, m, v, n, lr, b1, b2, eps, wd, step):
        calls.append(("adamw", p.data_ptr(), int(n), int(step)))

    monkeypatch.setattr(_lib, "atms_forward", atms_forward)
    monkeypatch.setattr(_lib, "atms_backward", atms_backward)
    monkeypatch.setattr(_lib, "infonce", infonce)
    monkeypatch.setattr(_lib, "mse", mse)
    monkeypatch.setattr(_lib, "adamw_step", adamw_step)
    return calls


def _trial_x(ids):
    x = torch.zeros(len(ids), 63, 250)
    x[:, 0, 0] = torch.arange(len(ids), dtype=torch.float32)      # trial number in element [b,0,0]
    return x


def test_joint_encode_regroups_trials_and_restores_order(fake):
    from eeg_image_decode_b200.joint import ATMS
    m = ATMS(joint_train=True).eval()
    sid = torch.tensor([3, 0, 3, 9, 0])
    out = m.encode(_trial_x(sid), sid, train=False)
    kind, B, phases, train, groups, seen_sid = fake[-1]
    assert (kind, B, phases, train) == ("fwd", 5, 7, 0)
    assert groups == ([0, 2, 4, 5], [0, 3, 9])                    # what eegb200_atms_forward is handed
    assert seen_sid.tolist() == [0, 0, 3, 3, 9]
    # the embeddings come back in the caller's trial order, each computed from its own trial and subject
    assert out[:, 0].tolist() == [0 + 300.0, 1 + 0.0, 2 + 300.0, 3 + 900.0, 4 + 0.0]
    assert m._last_subjects == [0, 3, 9]
    # gradient rows are regrouped the same way
    m.train()
    out = m.encode(_trial_x(sid), sid, train=True, seed=1)
    d = torch.arange(5, dtype=torch.float32)[:, None].expand(5, 1024).contiguous()
    m.backprop(d)
    assert fake[-1][0] == "bwd" and fake[-1][4][:, 0].tolist() == [1.0, 4.0, 0.0, 2.0, 3.0]
    # only the value embeddings of subjects 0, 3, 9 were handed out for writing
    touched = [sj for sj in range(10) if m.grad_view(f"encoder.enc_embedding.value_embedding.{sj}.weight").flatten()[0] != 0]
    assert touched == [0, 3, 9]


def test_joint_pointer_tables_address_the_right_parameters(fake):
    from eeg_image_decode_b200.joint import ATMS
    m = ATMS(joint_train=True)
    P, G, Bf, J = m._pointers()
    named = dict(m.named_parameters())
    for sj in range(10):
        w = named[f"encoder.enc_embedding.value_embedding.{sj}.weight"]
        b = named[f"encoder.enc_embedding.value_embedding.{sj}.bias"]
        assert J[0][sj] == w.data_ptr() and J[1][sj] == b.data_ptr()
        assert J[2][sj] == m.grad_view(f"encoder.enc_embedding.value_embedding.{sj}.weight").data_ptr()
        assert J[3][sj] == m.grad_view(f"encoder.enc_embedding.value_embedding.{sj}.bias").data_ptr()
    assert P[4] == named["encoder.encoder.attn_layers.0.attention.query_projection.weight"].data_ptr()
    assert all(P[i] for i in range(len(P)))                       # slots 0/1 hold valid placeholders


def test_joint_step_updates_only_present_subjects(fake):
    from eeg_image_decode_b200.joint import ATMS
    from eeg_image_decode_b200.train import StepEngine, publish_optimizer_state
    m = ATMS(joint_train=True).train()
    sid = torch.tensor([2, 2, 5, 5, 5, 2, 7, 7])
    img = recipe.make_targets(8, tag="img")
    txt = recipe.make_targets(8, tag="txt")
    opt = torch.optim.AdamW(m.parameters(), lr=3e-4)
    eng = StepEngine(m, opt)
    loss, feats = eng.step(_trial_x(sid), sid, img, txt, use_shared=False)
    assert abs(loss[0].item() - (0.99 * 1.0 + 0.01 * 2.0)) < 1e-6
    assert feats[:, 0].tolist() == [200.0, 201.0, 502.0, 503.0, 504.0, 205.0, 706.0, 707.0]
    bwd = [c for c in fake if c[0] == "bwd"][-1]
    assert bwd[3] == ([0, 3, 6, 8], [2, 5, 7])
    # d loss / d embedding (the stub returns the embedding itself) reaches the backward in regrouped order
    assert bwd[4][:, 0].tolist() == [200.0, 201.0, 205.0, 502.0, 503.0, 504.0, 706.0, 707.0]
    steps = [c for c in fake if c[0] == "adamw"]
    segs = m.adam_segments(False, [2, 5, 7])
    base = m.flat_params.data_ptr()
    assert [(c[1] - base) // 4 for c in steps] == [s[1] for s in segs] and [c[2] for c in steps] == [s[2] for s in segs]
    assert [s[0] for s in segs] == ["main", "table", "ve2", "ve5", "ve7"]
    assert m._adam_steps["ve5"] == 1 and m._adam_steps["ve4"] == 0 and m._adam_steps["shared"] == 0
    publish_optimizer_state(m, opt)
    named = dict(m.named_parameters())
    assert named["encoder.enc_embedding.value_embedding.5.bias"] in opt.state
    assert named["encoder.enc_embedding.value_embedding.4.bias"] not in opt.state
    assert named["encoder.enc_embedding.subject_embedding.shared_embedding"] not in opt.state


def test_joint_step_with_foreign_optimizer_sets_grads_like_autograd(fake):
    from eeg_image_decode_b200.joint import ATMS
    from eeg_image_decode_b200.train import StepEngine
    m = ATMS(joint_train=True).train()
    opt = torch.optim.SGD(m.parameters(), lr=0.0)
    sid = torch.tensor([1, 1, 4])
    StepEngine(m, opt).step(_trial_x(sid), sid, recipe.make_targets(3, tag="img"), recipe.make_targets(3, tag="txt"),
                            use_shared=False)
    named = dict(m.named_parameters())
    assert named["encoder.enc_embedding.value_embedding.1.weight"].grad is not None
    assert named["encoder.enc_embedding.value_embedding.4.bias"].grad is not None
    assert named["encoder.enc_embedding.value_embedding.0.weight"].grad is None
    assert named["encoder.enc_embedding.subject_embedding.shared_embedding"].grad is None
    assert named["subject_wise_linear.3.weight"].grad is None


def test_joint_autograd_bridge_grad_none_for_absent_subjects(fake):
    from eeg_image_decode_b200.joint import ATMS
    m = ATMS(joint_train=True).train()
    sid = torch.tensor([6, 1, 6])
    out = m(_trial_x(sid), sid)
    assert out[:, 0].tolist() == [600.0, 101.0, 602.0]
    out.sum().backward()
    named = dict(m.named_parameters())
    got = sorted(k for k, p in named.items() if p.grad is not None and ".value_embedding." in k)
    assert got == sorted(f"encoder.enc_embedding.value_embedding.{s}.{t}" for s in (1, 6) for t in ("weight", "bias"))
    assert named["proj_eeg.0.weight"].grad is not None and named["subject_wise_linear.0.weight"].grad is None
    assert named["encoder.enc_embedding.subject_embedding.subject_embedding.weight"].grad is not None
    assert named["encoder.enc_embedding.subject_embedding.shared_embedding"].grad is None


def test_plain_model_step_and_autograd_bridge_unchanged(fake):
    from eeg_image_decode_b200.atms import ATMS
    from eeg_image_decode_b200.train import StepEngine
    m = ATMS().train()
    sid = torch.full((4,), 8)
    loss, feats = StepEngine(m, None).step(_trial_x(sid), sid, recipe.make_targets(4, tag="img"),
                                           recipe.make_targets(4, tag="txt"), use_shared=False)
    assert [c[4] for c in fake if c[0] == "fwd"] == [None]           # no group table for the plain model
    steps = [c for c in fake if c[0] == "adamw"]
    assert [c[2] for c in steps] == [m._n_main, 2500] and m._adam_steps == {"main": 1, "table": 1, "shared": 0}
    out = m(_trial_x(sid), sid)
    out.sum().backward()
    named = dict(m.named_parameters())
    assert named["encoder.enc_embedding.value_embedding.weight"].grad is not None
    assert named["encoder.enc_embedding.subject_embedding.shared_embedding"].grad is None


def test_reconstruction_loss_composition(fake):
    """alpha*10*MSE + (1-alpha)*10*ClipLoss(img): weights handed to eegb200_infonce / eegb200_mse
    (Generation/ATMS_reconstruction.py:198, 227-228)"""
    from eeg_image_decode_b200.atms import ATMS
    from eeg_image_decode_b200.train import StepEngine
    m = ATMS().train()
    sid = torch.full((4,), 8)
    img = recipe.make_targets(4, tag="img")
    loss, feats = StepEngine(m, None, 0.90, "reconstruction").step(_trial_x(sid), sid, img, None, use_shared=False)
    nce = [c for c in fake if c[0] == "infonce"][-1]
    assert nce[4:] == (1.0, 0.0, False)                  # (1 - 0.9) * 10 on the image ClipLoss, no text target
    ms = [c for c in fake if c[0] == "mse"][-1]
    assert ms == ("mse", (4, 1024), 4, 9.0, 1.0, True)   # 0.9 * 10 on the MSE, mean over the 4 rows, gradient wanted
    assert abs(loss[0].item() - (1.0 + 7.0)) < 1e-6 and abs(loss[2].item() - 7.0) < 1e-6
    # evaluation-time composition: alpha = 0.99, no gradient
    eng = StepEngine(m, None, 0.99, "reconstruction")
    loss, d_e, d_s = eng.loss_and_grad(feats, img, None, need_grad=False)
    assert d_e is None and [c for c in fake if c[0] == "mse"][-1] == ("mse", (4, 1024), 4, 9.9, 1.0, False)
    assert [c for c in fake if c[0] == "infonce"][-1][4] == round(0.01 * 10, 6)
