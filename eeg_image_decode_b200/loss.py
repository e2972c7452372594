"""Contrastive (InfoNCE) loss behind the reference's ``ClipLoss`` signature (models/loss.py:78-141).

``ClipLoss()(eeg_features, target_features, logit_scale)`` -> 0-d loss, differentiable w.r.t. the first
argument and ``logit_scale`` (the targets are the frozen CLIP embeddings at every reference call site,
ATMS_retrieval.py:229-230; no gradient is produced for them).  ``logit_scale`` is used raw, like the
reference.  With ``world_size > 1`` the targets are all-gathered (NCCL) and every rank evaluates only its
own B_local x N row block; the column log-sum-exp statistics are exchanged with one small all-gather.  This
is the reference's ``local_loss=False, gather_with_grad=False`` result (loss.py:59-73, 113-120): same loss
value on every rank, same gradient w.r.t. the local embeddings.

``fused_contrastive`` evaluates both targets (image, text) of the training step in one pass.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch
import torch.nn as nn

from . import _lib


def _dist_ready(world_size: int) -> bool:
    return world_size > 1 and torch.distributed.is_available() and torch.distributed.is_initialized()


class _InfoNCE:
    """workspace + two-phase call into eegb200_infonce; shared by ClipLoss and the fused train step."""

    def __init__(self):
        # one workspace per problem shape, never replaced: a captured CUDA graph may hold pointers into any of them
        # (e.g. the full-batch shape) while a ragged last batch runs eagerly with another
        self._ws_by_key = {}

    def workspace(self, B, N, D, nt, device):
        key = (B, N, D, nt, device)
        ws = self._ws_by_key.get(key)
        if ws is None:
            if len(self._ws_by_key) >= 8 and not torch.cuda.is_current_stream_capturing():
                self._ws_by_key.pop(next(k for k in self._ws_by_key if k != getattr(self, "_graph_key", None)))
            ws = torch.empty(_lib.infonce_workspace_bytes(B, N, D, nt), dtype=torch.uint8, device=device)
            self._ws_by_key[key] = ws
        if torch.cuda.is_current_stream_capturing():
            self._graph_key = key           # never evicted
        return ws

    def target_slots(self, B, N, D, nt, device):
        """[nt, N, D] fp32 view of the operand the logits GEMM reads inside the workspace of this shape: targets gathered
        (TF32-rounded) straight into it and passed to run() are used in place (eegb200_infonce_target_offset)"""
        ws = self.workspace(B, N, D, nt, device)
        off = _lib.infonce_target_offset(B, N, D, nt)
        return ws[off:off + nt * N * D * 4].view(torch.float32).view(nt, N, D)

    def run(self, eeg, tgt_img, tgt_txt, logit_scale, w_img, w_txt, row_offset, need_grad, grad_out=1.0, group=None,
            world_size=1):
        """targets are the GLOBAL (already gathered) [N,D] matrices.  Returns (loss[3] device, d_eeg, d_scale)."""
        for name, t in (("eeg", eeg), ("tgt_img", tgt_img), ("tgt_txt", tgt_txt)):
            # raw device pointers cross the C ABI: anything but a dense fp32 matrix would be read as garbage
            if t is not None and (t.dtype != torch.float32 or not t.is_contiguous() or t.dim() != 2):
                raise RuntimeError(f"InfoNCE: {name} must be a contiguous 2-D float32 CUDA tensor, got {t.dtype} "
                                   f"{tuple(t.shape)} strides {t.stride()}")
        B, D = eeg.shape
        N = tgt_img.shape[0]
        nt = 2 if tgt_txt is not None else 1
        ws = self.workspace(B, N, D, nt, eeg.device)
        dev = eeg.device
        col_stats = torch.empty(2, nt * N, device=dev, dtype=torch.float32)
        loss = torch.zeros(3, device=dev, dtype=torch.float32)
        d_eeg = torch.empty_like(eeg) if need_grad else None
        d_scale = torch.zeros((), device=dev, dtype=torch.float32) if need_grad else None
        io = _lib.InfoNceIO()
        io.eeg, io.tgt_img = eeg.data_ptr(), tgt_img.data_ptr()
        io.tgt_txt = tgt_txt.data_ptr() if tgt_txt is not None else None
        io.B, io.N, io.D, io.row_offset = B, N, D, row_offset
        io.logit_scale = logit_scale.data_ptr()
        io.w_img, io.w_txt, io.grad_out = w_img, w_txt, grad_out
        io.workspace, io.workspace_bytes = ws.data_ptr(), ws.numel()
        io.col_stats = col_stats.data_ptr()
        io.col_parts, io.n_parts = None, 1
        io.loss = loss.data_ptr()
        io.d_eeg = d_eeg.data_ptr() if need_grad else None
        io.d_logit_scale = d_scale.data_ptr() if need_grad else None
        if _dist_ready(world_size):
            _lib.infonce(io, _lib.PHASE_A, dev)
            parts = torch.empty(world_size, 2, nt * N, device=dev, dtype=torch.float32)
            torch.distributed.all_gather_into_tensor(parts, col_stats, group=group)
            io.col_parts, io.n_parts = parts.data_ptr(), world_size
            _lib.infonce(io, _lib.PHASE_B, dev)
            self._keep = parts
        else:
            _lib.infonce(io, _lib.PHASE_A | _lib.PHASE_B, dev)
        return loss, d_eeg, d_scale


def _prep(t: torch.Tensor, what: str) -> torch.Tensor:
    if not (torch.is_tensor(t) and t.is_cuda):
        raise RuntimeError(f"ClipLoss: {what} must be a CUDA tensor (this path has no CPU implementation)")
    return t.detach().contiguous().float()


def gather_targets(t: torch.Tensor, world_size: int, group=None) -> torch.Tensor:
    """NCCL all-gather of the (B_local, D) target block -> (N, D); rank order == row order (loss.py:60-72)."""
    if not _dist_ready(world_size):
        return t
    out = torch.empty(world_size * t.shape[0], t.shape[1], device=t.device, dtype=t.dtype)
    torch.distributed.all_gather_into_tensor(out, t.contiguous(), group=group)
    return out


class _ClipLossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, eeg, tgt, logit_scale, mod):
        e = _prep(eeg, "image_features")
        t_local = _prep(tgt, "text_features")
        if e.shape != t_local.shape or e.dim() != 2:
            raise RuntimeError(f"ClipLoss expects two [B,D] tensors, got {tuple(eeg.shape)} and {tuple(tgt.shape)}")
        s = logit_scale.detach().float().reshape(()).contiguous()
        ws = mod.world_size if _dist_ready(mod.world_size) else 1
        t_all = gather_targets(t_local, ws)
        need = eeg.requires_grad or logit_scale.requires_grad
        loss, d_e, d_s = mod._engine.run(e, t_all, None, s, 1.0, 0.0, mod.rank * e.shape[0] if ws > 1 else 0, need,
                                         world_size=ws)
        total = loss[0].clone()
        if ws > 1:   # each rank holds its share; the reference returns the full loss on every rank
            torch.distributed.all_reduce(total)
        ctx.save_for_backward(d_e if need else None, d_s if need else None)
        return total

    @staticmethod
    def backward(ctx, g):
        d_e, d_s = ctx.saved_tensors
        if d_e is None:
            return None, None, None, None
        return d_e * g, None, (d_s * g).reshape(()), None


class ClipLoss(nn.Module):
    """Same constructor as the reference (models/loss.py:79-98)."""

    def __init__(self, local_loss=False, gather_with_grad=False, cache_labels=False, rank=0, world_size=1,
                 use_horovod=False):
        super().__init__()
        if use_horovod:
            raise NotImplementedError("horovod is not part of the B200 build; use torch.distributed (NCCL)")
        if local_loss or gather_with_grad:
            raise NotImplementedError("only the reference default (local_loss=False, gather_with_grad=False) is implemented")
        self.local_loss = local_loss
        self.gather_with_grad = gather_with_grad
        self.cache_labels = cache_labels
        self.rank = rank
        self.world_size = world_size
        self.use_horovod = use_horovod
        self._engine = _InfoNCE()

    def forward(self, image_features, text_features, logit_scale):
        if not torch.is_tensor(logit_scale):
            logit_scale = torch.tensor(float(logit_scale), device=image_features.device)
        return _ClipLossFn.apply(image_features, text_features, logit_scale, self)


def fused_contrastive(engine: _InfoNCE, eeg, img_all, txt_all, logit_scale, alpha=0.99, row_offset=0, need_grad=True,
                      world_size=1, group=None) -> Tuple[torch.Tensor, Optional[torch.Tensor], Optional[torch.Tensor]]:
    """alpha * ClipLoss(eeg, img) + (1-alpha) * ClipLoss(eeg, txt)   (ATMS_retrieval.py:229-234) in one pass.
    Returns (loss[3]: mix/img/txt shares of this rank, d_eeg, d_logit_scale)."""
    return engine.run(eeg, img_all, txt_all, logit_scale, float(alpha), float(1.0 - alpha), row_offset, need_grad,
                      world_size=world_size, group=group)
