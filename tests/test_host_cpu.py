"""CPU-side checks (no GPU): the C-ABI library loads and exports every declared symbol, the Python mirror of the
reference interface has the reference's state_dict, and the product path refuses to run without CUDA."""
import ctypes
import os
import re

import pytest
import torch

import recipe

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    from eeg_image_decode_b200 import _lib
    L = _lib.lib()
    hdr = open(os.path.join(ROOT, "include", "eegdecode_b200.h")).read()
    names = set(re.findall(r"\b(eegb200_[a-z0-9_]+)\s*\(", hdr))
    assert len(names) >= 14
    for n in sorted(names):
        assert hasattr(L, n), f"{n} declared in the header but not exported"
    assert L.eegb200_abi_version() == _lib.ABI_VERSION == 2
    assert _lib.atms_workspace_bytes(4) > 0
    assert _lib.infonce_workspace_bytes(8, 8, 1024, 2) > 0


def test_state_dict_matches_reference_layout():
    from eeg_image_decode_b200.atms import ATMS
    m = ATMS()
    sd = m.state_dict()
    assert list(sd.keys()) == list(recipe.STATE_SHAPES.keys())
    for k, shp in recipe.STATE_SHAPES.items():
        assert tuple(sd[k].shape) == tuple(shp), k
    assert sum(p.numel() for p in m.parameters()) == 3202413
    r = m.load_state_dict(recipe.make_state_dict(), strict=True)
    assert not r.missing_keys and not r.unexpected_keys
    assert abs(m.logit_scale.item() - 2.6593) < 1e-3


def test_no_cpu_fallback():
    from eeg_image_decode_b200.atms import ATMS
    from eeg_image_decode_b200.loss import ClipLoss
    m = ATMS().eval()
    with pytest.raises(RuntimeError):
        m(torch.randn(2, 63, 250), torch.tensor([1, 2]))
    with pytest.raises(RuntimeError):
        ClipLoss()(torch.randn(4, 1024), torch.randn(4, 1024), torch.tensor(2.0))


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "eeg_image_decode_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f), errors="replace").read()
                assert "import oracle" not in src and "from oracle" not in src and "oracle/" not in src, f


@pytest.mark.skipif(not os.path.isdir("/root/reference/Retrieval"), reason="live reference not mounted")
def test_default_init_is_rng_identical_to_reference():
    from ref_import import import_reference
    from eeg_image_decode_b200.atms import ATMS
    R = import_reference()
    torch.manual_seed(7)
    ref = R.ATMS().state_dict()
    torch.manual_seed(7)
    ours = ATMS().state_dict()
    assert list(ref.keys()) == list(ours.keys())
    for k in ref:
        assert torch.equal(ref[k], ours[k]), k
    # a reference checkpoint loads strictly, and ours loads into the reference
    R.ATMS().load_state_dict(ours, strict=True)


# ---------------------------------------------------------------- joint-subject / reconstruction variants (host logic)
def test_joint_model_state_dict_and_segments():
    from eeg_image_decode_b200.joint import ATMS
    m = ATMS(joint_train=True)
    shapes = recipe.joint_state_shapes()
    sd = m.state_dict()
    assert list(sd.keys()) == list(shapes.keys())
    for k, shp in shapes.items():
        assert tuple(sd[k].shape) == tuple(shp), k
    r = m.load_state_dict(recipe.make_joint_state_dict(), strict=True)
    assert not r.missing_keys and not r.unexpected_keys
    # plain constructor call of that script == single value embedding, ten subject_wise_linear layers
    p = ATMS()
    assert "encoder.enc_embedding.value_embedding.weight" in p.state_dict() and "subject_wise_linear.9.bias" in p.state_dict()
    # AdamW segments: always-trained prefix, table or shared token, one segment per subject present in the batch
    segs = m.adam_segments(False, [7, 2, 2])
    assert [s[0] for s in segs] == ["main", "table", "ve2", "ve7"]
    named = dict(m.named_parameters())
    for name, off, n in segs[2:]:
        sj = name[2:]
        w, b = named[f"encoder.enc_embedding.value_embedding.{sj}.weight"], named[f"encoder.enc_embedding.value_embedding.{sj}.bias"]
        base = m.flat_params.data_ptr()
        assert w.data_ptr() == base + 4 * off and b.data_ptr() + 4 * 250 == base + 4 * (off + n)
    assert m.adam_segment_of("encoder.enc_embedding.value_embedding.4.bias") == "ve4"
    assert m.adam_segment_of("proj_eeg.0.weight") == "main"
    # every trainable tensor on the hot path lies inside the gradient arena; the never-used ones do not
    assert m._offs["subject_wise_linear.0.weight"] >= m._n_hot
    assert m._offs["encoder.enc_embedding.value_embedding.9.bias"] + 250 <= m._n_hot


def test_joint_grouping_by_subject():
    from eeg_image_decode_b200.joint import ATMS
    m = ATMS(joint_train=True)
    x = torch.arange(5, dtype=torch.float32).reshape(5, 1, 1).expand(5, 63, 250).contiguous()
    sid = torch.tensor([3, 0, 3, 9, 0])
    xs, sids, perm, groups = m._group_by_subject(x, sid, None)
    assert perm.tolist() == [1, 4, 0, 2, 3] and sids.tolist() == [0, 0, 3, 3, 9]
    assert xs[:, 0, 0].tolist() == [1.0, 4.0, 0.0, 2.0, 3.0]
    assert groups == [(0, 0), (2, 3), (4, 9)]
    # already ordered / single subject: no permutation, no copy
    _, _, perm, groups = m._group_by_subject(x, torch.tensor([1, 1, 4, 4, 8]), None)
    assert perm is None and groups == [(0, 1), (2, 4), (4, 8)]
    _, _, perm, groups = m._group_by_subject(x, sid, 6)          # caller vouches for the subject: ids are not read
    assert perm is None and groups == [(0, 6)]
    with pytest.raises(KeyError):       # Embed.py:144: self.value_embedding['10']
        m._group_by_subject(x, torch.tensor([1, 10, 2, 3, 4]), None)
    with pytest.raises(KeyError):
        m._group_by_subject(x, sid, 10)


@pytest.mark.skipif(not os.path.isdir("/root/reference/Retrieval"), reason="live reference not mounted")
def test_joint_default_init_is_rng_identical_to_reference():
    from ref_import import import_reference_joint
    from eeg_image_decode_b200.joint import ATMS
    J = import_reference_joint()
    for jt in (True, False):
        torch.manual_seed(11)
        ref = J.ATMS(joint_train=jt).state_dict()
        torch.manual_seed(11)
        ours = ATMS(joint_train=jt).state_dict()
        assert list(ref.keys()) == list(ours.keys())
        for k in ref:
            assert torch.equal(ref[k], ours[k]), k
        J.ATMS(joint_train=jt).load_state_dict(ours, strict=True)


def test_variant_modules_mirror_reference_signatures():
    import inspect
    from eeg_image_decode_b200 import joint, reconstruction
    assert list(inspect.signature(joint.ATMS.__init__).parameters)[1:] == ["sequence_length", "num_subjects", "joint_train"]
    for mod in (joint, reconstruction):
        assert list(inspect.signature(mod.train_model).parameters)[:8] == [
            "sub", "eeg_model", "dataloader", "optimizer", "device", "text_features_all", "img_features_all", "config"]
        assert list(inspect.signature(mod.evaluate_model).parameters) == [
            "sub", "eeg_model", "dataloader", "device", "text_features_all", "img_features_all", "k", "config"]
    assert list(inspect.signature(reconstruction.get_eegfeatures).parameters)[:7] == [
        "sub", "eegmodel", "dataloader", "device", "text_features_all", "img_features_all", "k"]
    with pytest.raises(RuntimeError):       # no CPU fallback in the variants either
        reconstruction.train_model("sub-08", reconstruction.ATMS(), [], None, "cpu", torch.zeros(4, 1024), torch.zeros(40, 1024), None)


def test_ctypes_structs_match_the_c_header(tmp_path):
    """the Python binding's struct layouts are checked against the header with the C compiler (sizeof / offsetof)"""
    import shutil
    import subprocess
    from eeg_image_decode_b200 import _lib
    cc = shutil.which("gcc") or shutil.which("cc")
    if cc is None:
        pytest.skip("no C compiler")
    structs = {"eegb200_atms_io": _lib.AtmsIO, "eegb200_infonce_io": _lib.InfoNceIO, "eegb200_gemm_desc": _lib.GemmDesc}
    lines = []
    for cname, st in structs.items():
        lines.append(f'printf("{cname} %zu\\n", sizeof({cname}));')
        for fname, _ in st._fields_:
            lines.append(f'printf("{cname}.{fname} %zu\\n", offsetof({cname}, {fname}));')
    src = tmp_path / "layout.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "%s"\nint main(void){%s return 0;}\n'
                   % (os.path.join(ROOT, "include", "eegdecode_b200.h"), "".join(lines)))
    exe = tmp_path / "layout"
    subprocess.run([cc, str(src), "-o", str(exe)], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split("\n")
    got = {l.split()[0]: int(l.split()[1]) for l in out if l.strip()}
    for cname, st in structs.items():
        assert got[cname] == ctypes.sizeof(st), cname
        for fname, _ in st._fields_:
            assert got[f"{cname}.{fname}"] == getattr(st, fname).offset, f"{cname}.{fname}"


def test_fused_adamw_is_used_only_for_plain_single_group_adamw():
    """the fused kernel must not silently replace other optimizer semantics"""
    from eeg_image_decode_b200.atms import ATMS
    from eeg_image_decode_b200.train import fused_optimizer_ok
    m = ATMS()
    assert fused_optimizer_ok(m, None)
    assert fused_optimizer_ok(m, torch.optim.AdamW(m.parameters(), lr=3e-4))
    assert not fused_optimizer_ok(m, torch.optim.SGD(m.parameters(), lr=1e-2))
    assert not fused_optimizer_ok(m, torch.optim.AdamW(m.parameters(), lr=3e-4, amsgrad=True))
    assert not fused_optimizer_ok(m, torch.optim.AdamW(m.parameters(), lr=3e-4, maximize=True))
    ps = list(m.parameters())
    assert not fused_optimizer_ok(m, torch.optim.AdamW([{"params": ps[:10]}, {"params": ps[10:], "lr": 1e-5}]))
    assert not fused_optimizer_ok(m, torch.optim.AdamW(ps[:-3], lr=3e-4))            # hot parameters missing from the group
    m.proj_eeg[0].weight.requires_grad_(False)
    assert not fused_optimizer_ok(m, torch.optim.AdamW(m.parameters(), lr=3e-4))     # a frozen hot parameter


def test_optimizer_state_is_imported_and_tied_to_the_optimizer_object():
    """resuming from optimizer.load_state_dict() continues the moments / bias-correction step; a new optimizer object
    starts from zero; .float() / .to() keep the moments"""
    from eeg_image_decode_b200.atms import ATMS
    from eeg_image_decode_b200.train import adopt_optimizer, publish_optimizer_state
    m = ATMS()
    named = dict(m.named_parameters())
    hot = m._hot_order()
    # a "checkpoint": a stock AdamW that has taken 3 steps on the same parameters
    donor = torch.optim.AdamW(m.parameters(), lr=0.0)
    for _ in range(3):
        for n in hot:
            named[n].grad = torch.full_like(named[n], 0.5)
        donor.step()
    ckpt = donor.state_dict()
    opt = torch.optim.AdamW(m.parameters(), lr=3e-4)
    adopt_optimizer(m, opt)
    assert m._adam_steps["main"] == 0 and float(m._adam_m.abs().sum()) == 0.0
    opt.load_state_dict(ckpt)
    adopt_optimizer(m, opt)
    w = "proj_eeg.0.weight"
    o = m._offs[w]
    assert m._adam_steps["main"] == 3 and m._adam_steps["table"] == 3
    assert torch.allclose(m._adam_m[o:o + 8], donor.state[named[w]]["exp_avg"].reshape(-1)[:8])
    assert torch.allclose(m._adam_v[o:o + 8], donor.state[named[w]]["exp_avg_sq"].reshape(-1)[:8])
    # published views alias the arenas: adopting again is a no-op
    publish_optimizer_state(m, opt)
    assert opt.state[named[w]]["exp_avg"].data_ptr() == m._adam_m[o:].data_ptr()
    before = m._adam_m.clone()
    adopt_optimizer(m, opt)
    assert torch.equal(before, m._adam_m) and m._adam_steps["main"] == 3
    # .float() rebuilds the arenas: the moments and step counters survive
    m.float()
    assert torch.equal(before, m._adam_m) and m._adam_steps["main"] == 3
    # a different optimizer object does not inherit them
    adopt_optimizer(m, torch.optim.AdamW(m.parameters(), lr=3e-4))
    assert m._adam_steps["main"] == 0 and float(m._adam_m.abs().sum()) == 0.0


def test_infonce_rejects_strided_or_non_fp32_inputs():
    """raw pointers cross the C ABI: the loss front end refuses anything but dense fp32 matrices (no silent garbage)"""
    import pytest
    import torch
    from eeg_image_decode_b200.loss import _InfoNCE
    e = torch.zeros(4, 8)
    s = torch.tensor(1.0)
    with pytest.raises(RuntimeError, match="contiguous 2-D float32"):
        _InfoNCE().run(e, torch.zeros(8, 4).t(), None, s, 1.0, 0.0, 0, False)
    with pytest.raises(RuntimeError, match="contiguous 2-D float32"):
        _InfoNCE().run(e.half(), torch.zeros(4, 8), None, s, 1.0, 0.0, 0, False)
