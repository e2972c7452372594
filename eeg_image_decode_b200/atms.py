"""ATM-S encoder behind the reference's Python surface.

Drop-in for ``ATMS`` of Retrieval/ATMS_retrieval.py:171-191: same constructor, ``forward(x, subject_ids)``,
``.logit_scale``, ``.loss_func`` and a ``state_dict`` with the reference's keys/shapes, so a
reference ``.pth`` loads with ``strict=True`` and vice versa.  The sub-modules below are parameter
containers only: all arithmetic happens in libeegdecode_b200.so (hand-written sm_100a kernels); there is no
PyTorch-eager or CPU path -- calling the model on CPU tensors raises.

Parameters are views into one flat fp32 arena (``ATMS.flat_params``) so that the fused AdamW kernel and the
data-parallel gradient all-reduce work on a single buffer.
"""
from __future__ import annotations

import ctypes
import math
from typing import Dict, List, Optional

import numpy as np
import torch
import torch.nn as nn

from . import _lib
from .loss import ClipLoss

N_SUBJECT_ROWS = 10   # iTransformer(num_subjects=10) default (ATMS_retrieval.py:62)
JOINT_VALUE_PREFIX = "encoder.enc_embedding.value_embedding."   # + "<subject>.weight" / ".bias" (Embed.py:128-130)


# ------------------------------------------------------------------------------------------------
# parameter containers (names and nesting give the reference state_dict keys)
# ------------------------------------------------------------------------------------------------
class _Holder(nn.Module):
    def forward(self, *a, **k):  # pragma: no cover
        raise RuntimeError("this sub-module only holds parameters; call ATMS.forward (CUDA kernels) instead")


class _PositionalEmbedding(_Holder):          # Embed.py:8-26
    def __init__(self, d_model, max_len=5000):
        super().__init__()
        pe = torch.zeros(max_len, d_model).float()
        position = torch.arange(0, max_len).float().unsqueeze(1)
        div_term = (torch.arange(0, d_model, 2).float() * -(math.log(10000.0) / d_model)).exp()
        pe[:, 0::2] = torch.sin(position * div_term)
        pe[:, 1::2] = torch.cos(position * div_term)
        self.register_buffer("pe", pe.unsqueeze(0))


class _TimeFeatureEmbedding(_Holder):         # Embed.py:96-106 (instantiated, never used: x_mark is None)
    def __init__(self, d_model):
        super().__init__()
        self.embed = nn.Linear(4, d_model, bias=False)


class _SubjectEmbedding(_Holder):             # Embed.py:109-121
    def __init__(self, num_subjects, d_model):
        super().__init__()
        self.subject_embedding = nn.Embedding(num_subjects, d_model)
        self.shared_embedding = nn.Parameter(torch.randn(1, d_model))
        self.mask_embedding = nn.Parameter(torch.randn(1, d_model))


class _DataEmbedding(_Holder):                # Embed.py:124-139
    def __init__(self, c_in, d_model, num_subjects, joint_train=False):
        super().__init__()
        if joint_train:     # one value embedding per subject (Embed.py:127-130)
            self.value_embedding = nn.ModuleDict({str(sj): nn.Linear(c_in, d_model) for sj in range(num_subjects)})
        else:
            self.value_embedding = nn.Linear(c_in, d_model)
        self.position_embedding = _PositionalEmbedding(d_model)
        self.temporal_embedding = _TimeFeatureEmbedding(d_model)
        self.subject_embedding = _SubjectEmbedding(num_subjects, d_model)
        self.mask_token = nn.Parameter(torch.randn(1, d_model))


class _AttentionLayer(_Holder):               # SelfAttention_Family.py:179-192
    def __init__(self, d_model, n_heads):
        super().__init__()
        dk = d_model // n_heads
        self.query_projection = nn.Linear(d_model, dk * n_heads)
        self.key_projection = nn.Linear(d_model, dk * n_heads)
        self.value_projection = nn.Linear(d_model, dk * n_heads)
        self.out_projection = nn.Linear(dk * n_heads, d_model)


class _EncoderLayer(_Holder):                 # Transformer_EncDec.py:27-37
    def __init__(self, d_model, n_heads, d_ff):
        super().__init__()
        self.attention = _AttentionLayer(d_model, n_heads)
        self.conv1 = nn.Conv1d(d_model, d_ff, 1)
        self.conv2 = nn.Conv1d(d_ff, d_model, 1)
        self.norm1 = nn.LayerNorm(d_model)
        self.norm2 = nn.LayerNorm(d_model)


class _Encoder(_Holder):                      # Transformer_EncDec.py:54-59
    def __init__(self, d_model, n_heads, d_ff):
        super().__init__()
        self.attn_layers = nn.ModuleList([_EncoderLayer(d_model, n_heads, d_ff)])
        self.norm = nn.LayerNorm(d_model)


class _ITransformer(_Holder):                 # ATMS_retrieval.py:61-85
    def __init__(self, seq_len=250, d_model=250, n_heads=4, d_ff=256, num_subjects=N_SUBJECT_ROWS, joint_train=False):
        super().__init__()
        self.enc_embedding = _DataEmbedding(seq_len, d_model, num_subjects, joint_train)
        self.encoder = _Encoder(d_model, n_heads, d_ff)


class _PatchEmbedding(_Holder):               # ATMS_retrieval.py:97-116
    def __init__(self, emb_size=40):
        super().__init__()
        self.tsconv = nn.Sequential(
            nn.Conv2d(1, 40, (1, 25), stride=(1, 1)), nn.Identity(), nn.BatchNorm2d(40), nn.Identity(),
            nn.Conv2d(40, 40, (63, 1), stride=(1, 1)), nn.BatchNorm2d(40), nn.Identity(), nn.Identity())
        self.projection = nn.Sequential(nn.Conv2d(40, emb_size, (1, 1), stride=(1, 1)), nn.Identity())


class _Fn(_Holder):
    def __init__(self, fn):
        super().__init__()
        self.fn = fn


# ------------------------------------------------------------------------------------------------
class ATMS(nn.Module):
    """``ATMS(num_channels=63, sequence_length=250, num_subjects=2, num_features=64, num_latents=1024, num_blocks=1)``

    ``normalize`` (keyword only, default False = the reference, which feeds the un-normalised LayerNorm output to the
    logits, ATMS_retrieval.py:182-191): L2-normalise the returned embedding (forward and backward), the CLIP-style
    variant BASELINE.json's north_star describes."""

    def __init__(self, num_channels=63, sequence_length=250, num_subjects=2, num_features=64, num_latents=1024,
                 num_blocks=1, *, normalize=False, _joint_train=False):
        super().__init__()
        self.normalize = bool(normalize)
        if num_channels != 63 or sequence_length != 250 or num_latents != 1024:
            raise ValueError("the sm_100a kernels are specialised for the reference geometry: 63 channels x 250 samples -> 1024")
        # per-subject value embeddings (the model of Retrieval/ATMS_retrieval_joint_train.py; see joint.py)
        self.joint_train = bool(_joint_train)
        self.encoder = _ITransformer(joint_train=self.joint_train)
        self.subject_wise_linear = nn.ModuleList([nn.Linear(250, sequence_length) for _ in range(num_subjects)])
        self.enc_eeg = nn.Sequential(_PatchEmbedding(), nn.Identity())
        self.proj_eeg = nn.Sequential(
            nn.Linear(1440, 1024),
            _Fn(nn.Sequential(nn.Identity(), nn.Linear(1024, 1024), nn.Identity())),
            nn.LayerNorm(1024))
        self.logit_scale = nn.Parameter(torch.ones([]) * np.log(1 / 0.07))
        self.loss_func = ClipLoss()
        # runtime state (not part of the state_dict)
        self.dropout_p = list(_lib.REF_DROPOUT_P)   # index = eegb200_dropout_site
        self._seed_gen = torch.Generator().manual_seed(torch.initial_seed() & 0x7FFFFFFF)
        self._ws: Dict[int, torch.Tensor] = {}
        self._last = None
        self._last_subjects = None      # joint-subject model: subjects of the latest encode()
        self._flatten()

    # ---------------------------------------------------------------- flat arena
    def _p_names(self) -> List[str]:
        """state_dict names behind the C-ABI parameter slots (enum eegb200_param).  The joint-subject model has no single
        value embedding: slots 0/1 then carry subject 0's tensors as valid-but-unused placeholders."""
        if not self.joint_train:
            return list(_lib.P_NAMES)
        names = list(_lib.P_NAMES)
        names[0], names[1] = JOINT_VALUE_PREFIX + "0.weight", JOINT_VALUE_PREFIX + "0.bias"
        return names

    def _joint_names(self):
        return [(f"{JOINT_VALUE_PREFIX}{sj}.weight", f"{JOINT_VALUE_PREFIX}{sj}.bias") for sj in range(N_SUBJECT_ROWS)]

    def _hot_order(self) -> List[str]:
        skip = {_lib.P_SUBJ_TABLE, _lib.P_SUBJ_SHARED} | ({0, 1} if self.joint_train else set())
        names = [n for i, n in enumerate(_lib.P_NAMES) if i not in skip]
        names += ["logit_scale", _lib.P_NAMES[_lib.P_SUBJ_TABLE], _lib.P_NAMES[_lib.P_SUBJ_SHARED]]
        if self.joint_train:      # one AdamW segment per subject, weight then bias
            for w, b in self._joint_names():
                names += [w, b]
        return names

    def _flatten(self) -> None:
        """(re)build the flat arenas and point every parameter at its slice"""
        named = dict(self.named_parameters())
        hot = self._hot_order()
        cold = [n for n in named if n not in hot]
        dev = named["logit_scale"].device
        offs, off = {}, 0
        for n in hot + cold:
            offs[n] = off
            off += (named[n].numel() + 63) // 64 * 64
        flat = torch.zeros(off, dtype=torch.float32, device=dev)
        for n in hot + cold:
            p = named[n]
            sl = flat[offs[n]:offs[n] + p.numel()].view(p.shape)
            sl.copy_(p.data)
            p.data = sl
        self.flat_params = flat
        self._offs = offs
        self._n_main = offs[_lib.P_NAMES[_lib.P_SUBJ_TABLE]]            # always-trained prefix (incl. logit_scale)
        self._n_hot = offs[cold[0]] if cold else off
        self.flat_grads = torch.zeros(self._n_hot, dtype=torch.float32, device=dev)
        # fused-AdamW moments survive .to() / .cuda() / .float() (the arena layout depends on the parameter names only);
        # they belong to ONE optimizer object (train.py::adopt_optimizer resets / imports them when it changes)
        old_m, old_v = getattr(self, "_adam_m", None), getattr(self, "_adam_v", None)
        if old_m is not None and old_m.numel() == self._n_hot:
            self._adam_m, self._adam_v = old_m.to(dev), old_v.to(dev)
        else:
            self._adam_m = None
            self._adam_v = None
            self._adam_steps = {"main": 0, "table": 0, "shared": 0}
            if self.joint_train:
                self._adam_steps.update({f"ve{sj}": 0 for sj in range(N_SUBJECT_ROWS)})
            self._adam_owner = None
        self._ptr_cache = None
        self._ws = {}
        self._ws_pinned = set()         # batch sizes whose workspace a captured CUDA graph points into
        self._gstep_cache = {}          # captured training steps, reused across train_model() calls (train.py)
        # device counter mixed into the dropout seed by the kernels (0 unless a captured CUDA graph advances it)
        self._seed_ctr = torch.zeros(1, dtype=torch.int64, device=dev)

    def adam_segments(self, use_shared: bool, subjects=None):
        """(name, offset, length) slices of the flat arena that received a gradient this step -- torch.optim skips
        parameters whose grad is None: the unused half of {subject table, shared token} and, in the joint-subject model,
        the value embeddings of subjects absent from the batch"""
        o_tab = self._offs[_lib.P_NAMES[_lib.P_SUBJ_TABLE]]
        o_sh = self._offs[_lib.P_NAMES[_lib.P_SUBJ_SHARED]]
        segs = [("main", 0, self._n_main)]
        segs.append(("shared", o_sh, 250) if use_shared else ("table", o_tab, N_SUBJECT_ROWS * 250))
        if self.joint_train:
            jn = self._joint_names()
            for sj in sorted(set(int(v) for v in (subjects if subjects is not None else range(N_SUBJECT_ROWS)))):
                w, b = jn[sj]
                segs.append((f"ve{sj}", self._offs[w], self._offs[b] + 250 - self._offs[w]))
        return segs

    def adam_segment_of(self, name: str) -> str:
        if name == _lib.P_NAMES[_lib.P_SUBJ_TABLE]:
            return "table"
        if name == _lib.P_NAMES[_lib.P_SUBJ_SHARED]:
            return "shared"
        if self.joint_train and name.startswith(JOINT_VALUE_PREFIX):
            return "ve" + name[len(JOINT_VALUE_PREFIX):].split(".")[0]
        return "main"

    def _apply(self, fn, recurse=True):
        r = super()._apply(fn, recurse)
        self._flatten()
        return r

    def grad_view(self, name: str) -> torch.Tensor:
        p = dict(self.named_parameters())[name]
        o = self._offs[name]
        return self.flat_grads[o:o + p.numel()].view(p.shape)

    def _pointers(self):
        if self._ptr_cache is None:
            named = dict(self.named_parameters())
            bufs = dict(self.named_buffers())
            pn = self._p_names()
            for n in pn:
                if not named[n].is_cuda:
                    raise RuntimeError("ATMS lives on %s: this implementation is CUDA-only (sm_100a kernels, no CPU "
                                       "fallback); call .to('cuda') first" % named[n].device)
            P = _lib.PtrArrayP(*[named[n].data_ptr() for n in pn])
            G = _lib.PtrArrayP(*[self.flat_grads.data_ptr() + 4 * self._offs[n] for n in pn])
            Bf = _lib.PtrArrayB(*[bufs[n].data_ptr() for n in _lib.BUF_NAMES])
            J = None
            if self.joint_train:
                arr = ctypes.c_void_p * N_SUBJECT_ROWS
                jn = self._joint_names()
                gp = self.flat_grads.data_ptr()
                J = (arr(*[named[w].data_ptr() for w, _ in jn]), arr(*[named[b].data_ptr() for _, b in jn]),
                     arr(*[gp + 4 * self._offs[w] for w, _ in jn]), arr(*[gp + 4 * self._offs[b] for _, b in jn]))
            self._ptr_cache = (P, G, Bf, J)
        return self._ptr_cache

    def workspace(self, B: int) -> torch.Tensor:
        ws = self._ws.get(B)
        # the size depends on the library state (fused conv path vs. verification backend / debug stage stores)
        if ws is not None and ws.numel() < _lib.atms_workspace_bytes(B) and B not in self._ws_pinned:
            ws = None
        if ws is None:
            ws = torch.empty(_lib.atms_workspace_bytes(B), dtype=torch.uint8, device=self.flat_params.device)
            # keep only the latest batch size resident, plus the ones a live CUDA graph was captured on
            self._ws = {b: t for b, t in self._ws.items() if b in self._ws_pinned}
            self._ws[B] = ws
        return ws

    # ---------------------------------------------------------------- engine-level API (no autograd)
    def _make_io(self, x, subject_ids, train: bool, seed: int, out, groups=None):
        P, G, Bf, J = self._pointers()
        B = x.shape[0]
        ws = self.workspace(B)
        io = _lib.AtmsIO()
        io.params = ctypes.cast(P, ctypes.POINTER(ctypes.c_void_p))
        io.buffers = ctypes.cast(Bf, ctypes.POINTER(ctypes.c_void_p))
        io.x = x.data_ptr()
        io.subject_ids = subject_ids.data_ptr()
        io.B = B
        io.n_subjects = N_SUBJECT_ROWS
        io.train = int(train)
        io.update_running_stats = int(train)
        io.seed = seed
        self._p_arr = _lib.FloatArrayS(*self.dropout_p)
        io.dropout_p = ctypes.cast(self._p_arr, ctypes.POINTER(ctypes.c_float))
        io.workspace = ws.data_ptr()
        io.workspace_bytes = ws.numel()
        io.out = out.data_ptr()
        io.seed_offset_dev = self._seed_ctr.data_ptr() if self._seed_ctr.is_cuda else None
        if self.joint_train:
            # groups: [(first trial, subject)], trials of one subject contiguous (see _group_by_subject)
            offs = (ctypes.c_int32 * (len(groups) + 1))(*[g[0] for g in groups], B)
            subj = (ctypes.c_int32 * len(groups))(*[g[1] for g in groups])
            pp = ctypes.POINTER(ctypes.c_void_p)
            io.joint_value_w, io.joint_value_b = ctypes.cast(J[0], pp), ctypes.cast(J[1], pp)
            io.joint_value_dw, io.joint_value_db = ctypes.cast(J[2], pp), ctypes.cast(J[3], pp)
            io.group_offsets, io.group_subject, io.n_groups = offs, subj, len(groups)
            self._grp_arrs = (offs, subj)       # keep the host arrays alive as long as the io
        return io

    def _group_by_subject(self, x, subject_ids, known_subject: Optional[int]):
        """joint-subject model: order the batch so that trials of one subject are contiguous (one grouped GEMM per
        subject instead of the reference's per-trial Python loop, Embed.py:144).  Reads the ids on the host like the
        reference's per-trial ``.item()`` unless the caller states the (single) subject of the batch.
        Returns (x, subject_ids, perm or None, [(first trial, subject)])."""
        B = x.shape[0]
        if known_subject is not None:
            ids = [int(known_subject)] * B
        else:
            ids = [int(v) for v in subject_ids.tolist()]
        for v in ids:
            if not 0 <= v < N_SUBJECT_ROWS:
                raise KeyError(str(v))          # the reference: self.value_embedding[str(subject_id.item())]
        order = sorted(range(B), key=ids.__getitem__)      # stable
        perm = None
        if order != list(range(B)):
            perm = torch.tensor(order, dtype=torch.long, device=x.device)
            x = x.index_select(0, perm)
            subject_ids = subject_ids.index_select(0, perm)
        groups = []
        for pos, i in enumerate(order):
            if not groups or groups[-1][1] != ids[i]:
                groups.append((pos, ids[i]))
        return x, subject_ids, perm, groups

    def _check_inputs(self, x, subject_ids):
        if not (torch.is_tensor(x) and x.is_cuda):
            raise RuntimeError("ATMS.forward expects a CUDA tensor (no CPU fallback on this path)")
        if x.dim() != 3 or x.shape[1] != 63 or x.shape[2] != 250:
            raise RuntimeError(f"expected EEG of shape [B,63,250], got {tuple(x.shape)}")
        x = x.contiguous().float()
        subject_ids = subject_ids.to(device=x.device, dtype=torch.int64).contiguous()
        if subject_ids.numel() != x.shape[0]:
            raise RuntimeError("subject_ids must have one entry per trial")
        return x, subject_ids

    def next_seed(self) -> int:
        return int(torch.randint(0, 2 ** 62, (1,), generator=self._seed_gen).item())

    def encode(self, x, subject_ids, train: Optional[bool] = None, seed: Optional[int] = None,
               phases: int = _lib.PHASE_ALL, out: Optional[torch.Tensor] = None,
               known_subject: Optional[int] = None) -> torch.Tensor:
        """forward through the CUDA kernels; keeps what the backward needs in the workspace.
        ``known_subject`` (joint-subject model only): the caller guarantees every trial carries this id, which saves
        the host read of ``subject_ids``."""
        x, subject_ids = self._check_inputs(x, subject_ids)
        train = self.training if train is None else train
        if seed is None:
            seed = self.next_seed() if train else 0
        perm, groups = None, None
        if self.joint_train:
            x, subject_ids, perm, groups = self._group_by_subject(x, subject_ids, known_subject)
            if perm is not None and phases != _lib.PHASE_ALL:
                raise NotImplementedError("phase-split (data-parallel) steps of the joint-subject model need batches that are "
                                          "already ordered by subject")
        user_out = out
        if out is None or perm is not None:
            out = torch.empty(x.shape[0], 1024, device=x.device, dtype=torch.float32)
        raw = out
        if self.normalize:      # the kernels write the LayerNorm output here; `out` receives its L2-normalised rows
            raw = torch.empty_like(out)
        io = self._make_io(x, subject_ids, train, seed, raw, groups)
        _lib.atms_forward(io, phases, x.device)
        if self.normalize and (phases & _lib.PHASE_C):
            self._norms = torch.empty(out.shape[0], device=out.device, dtype=torch.float32)
            _lib.l2norm_forward(raw, out, self._norms)
            self._norm_y = out
        if train and (phases & _lib.PHASE_C):
            for bn in (self.enc_eeg[0].tsconv[2], self.enc_eeg[0].tsconv[5]):
                bn.num_batches_tracked.add_(1)
        self._last = (io, x, subject_ids, out, perm, getattr(self, "_grp_arrs", None))
        self._fwd_gen = getattr(self, "_fwd_gen", 0) + 1      # which forward owns the saved activations (autograd bridge)
        self._last_subjects = sorted({g[1] for g in groups}) if groups else None
        if perm is not None:          # hand the embeddings back in the caller's trial order
            res = user_out if user_out is not None else torch.empty_like(out)
            res.index_copy_(0, perm, out)
            return res
        return out

    def backprop(self, d_out: torch.Tensor, phases: int = _lib.PHASE_ALL) -> None:
        """accumulates d loss / d params into ``flat_grads`` (views: ``grad_view(name)``)"""
        if self._last is None:
            raise RuntimeError("backprop() needs a preceding train-mode encode()")
        io, perm = self._last[0], self._last[4]
        G = self._pointers()[1]
        if d_out is not None:
            d_out = d_out.contiguous() if perm is None else d_out.index_select(0, perm)
            if self.normalize:
                d_raw = torch.empty_like(d_out)
                _lib.l2norm_backward(self._norm_y, self._norms, d_out.float(), d_raw)
                d_out = d_raw
        _lib.atms_backward(io, d_out, ctypes.cast(G, ctypes.POINTER(ctypes.c_void_p)), phases, self.flat_params.device)

    def zero_flat_grads(self) -> None:
        self.flat_grads.zero_()

    def ws_tensor(self, name: str) -> torch.Tensor:
        io = self._last[0]
        return _lib.ws_tensor(self.workspace(io.B), io.B, name)

    # ---------------------------------------------------------------- drop-in forward (autograd)
    def _autograd_names(self) -> List[str]:
        if not self.joint_train:
            return list(_lib.P_NAMES)
        return [n for i, n in enumerate(_lib.P_NAMES) if i > 1] + [n for wb in self._joint_names() for n in wb]

    def forward(self, x, subject_ids):
        if torch.is_grad_enabled() and self.training:
            named = dict(self.named_parameters())
            plist = [named[n] for n in self._autograd_names()]
            return _ATMSFunction.apply(self, x, subject_ids, *plist)
        return self.encode(x, subject_ids, train=self.training)


class _ATMSFunction(torch.autograd.Function):
    """autograd bridge: lets reference-style code do ``loss.backward()`` through the CUDA backward."""

    @staticmethod
    def forward(ctx, model, x, subject_ids, *params):
        out = model.encode(x, subject_ids, train=True)
        ctx.model = model
        ctx.gen = model._fwd_gen
        ctx.use_shared = bool((subject_ids >= N_SUBJECT_ROWS).any().item()) or bool((subject_ids < 0).any().item())
        ctx.subjects = model._last_subjects
        return out

    @staticmethod
    def backward(ctx, d_out):
        model = ctx.model
        if ctx.gen != model._fwd_gen:
            # the saved activations live in ONE workspace per batch size: a later forward (train or eval) has replaced them
            raise RuntimeError("eeg_image_decode_b200: backward() through an ATMS forward whose activations were "
                               "overwritten by a later forward of the same model; call backward() before the next forward")
        model.flat_grads.zero_()
        model.backprop(d_out.contiguous())
        tab, sh = _lib.P_NAMES[_lib.P_SUBJ_TABLE], _lib.P_NAMES[_lib.P_SUBJ_SHARED]
        grads = []
        for n in model._autograd_names():
            if (n == tab and ctx.use_shared) or (n == sh and not ctx.use_shared):
                grads.append(None)
            elif n.startswith(JOINT_VALUE_PREFIX) and model.joint_train and \
                    int(n[len(JOINT_VALUE_PREFIX):].split(".")[0]) not in ctx.subjects:
                grads.append(None)          # subjects absent from the batch: grad stays None like in the reference
            else:
                grads.append(model.grad_view(n).clone())
        return (None, None, None, *grads)
