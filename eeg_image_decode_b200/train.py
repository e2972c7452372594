"""Training / evaluation step functions with the reference's signatures
(Retrieval/ATMS_retrieval.py:199-254 ``train_model``, :258-362 ``evaluate_model``, :364-512 ``main_train_loop``).

Differences that are not observable through the return values:
  * loss and accuracy are accumulated on the device and read back once per call instead of two ``.item()``
    syncs per step (:238, :250);
  * the optimiser update runs as one fused AdamW kernel over the flat parameter arena when the caller passes a
    ``torch.optim.AdamW`` (hyper-parameters are read from its ``param_groups``; moments are exposed through
    ``optimizer.state`` as views);
  * with torch.distributed initialised (one process per GPU) the step is data-parallel: targets are
    all-gathered for the global-batch InfoNCE, BatchNorm statistics and gradients are all-reduced (SUM), so W
    ranks x B_local behave like one process at batch W*B_local.
"""
from __future__ import annotations

import random
import re
from typing import Optional

import torch

from . import _lib
from .atms import ATMS, N_SUBJECT_ROWS
from .loss import _InfoNCE, fused_contrastive, gather_targets


def extract_id_from_string(s):
    match = re.search(r"\d+$", s)
    if match:
        return int(match.group())
    return None


def _world():
    if torch.distributed.is_available() and torch.distributed.is_initialized():
        return torch.distributed.get_world_size(), torch.distributed.get_rank()
    return 1, 0


# ------------------------------------------------------------------------------------------------
# fused AdamW over the flat arena
# ------------------------------------------------------------------------------------------------
def _adam_hparams(optimizer):
    if optimizer is None:
        return dict(lr=3e-4, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2)
    g = optimizer.param_groups[0]
    return dict(lr=float(g["lr"]), betas=tuple(g["betas"]), eps=float(g["eps"]), weight_decay=float(g["weight_decay"]))


def fused_optimizer_ok(model: ATMS, optimizer) -> bool:
    """The fused AdamW kernel implements exactly torch.optim.AdamW with ONE parameter group that covers every trainable
    parameter of the hot path and default flags.  Anything else (several groups, amsgrad / maximize, frozen or
    missing parameters, another optimizer class) keeps the caller's semantics through optimizer.step()."""
    if optimizer is None:
        return True
    if type(optimizer).__name__ not in ("AdamW", "FusedAdamW") or len(optimizer.param_groups) != 1:
        return False
    g = optimizer.param_groups[0]
    if g.get("amsgrad", False) or g.get("maximize", False) or g.get("differentiable", False):
        return False
    if not all(isinstance(g.get(k), (int, float)) for k in ("lr", "eps", "weight_decay")):
        return False          # tensor learning rates (capturable mode) are not baked into the kernel arguments
    covered = {id(p) for p in g["params"]}
    named = dict(model.named_parameters())
    return all(id(named[n]) in covered and named[n].requires_grad for n in model._hot_order())


def adopt_optimizer(model: ATMS, optimizer) -> None:
    """Tie the flat moment arenas to ``optimizer``: a NEW optimizer object starts from zero moments (it must not inherit
    another one's), and state that did not come from the arenas -- ``optimizer.load_state_dict(checkpoint)``, steps taken
    through ``optimizer.step()`` -- is imported (exp_avg, exp_avg_sq, step) so that resuming continues where the
    checkpoint stopped instead of silently restarting the moments and the bias correction."""
    if optimizer is None:
        return
    if model._adam_m is None:
        model._adam_m = torch.zeros_like(model.flat_grads)
        model._adam_v = torch.zeros_like(model.flat_grads)
    if getattr(model, "_adam_owner", None) != id(optimizer):
        model._adam_m.zero_()
        model._adam_v.zero_()
        model._adam_steps = {k: 0 for k in model._adam_steps}
        model._adam_owner = id(optimizer)
    named = dict(model.named_parameters())
    for n in model._hot_order():
        p = named[n]
        st = optimizer.state.get(p)
        if not st or "exp_avg" not in st:
            continue
        o = model._offs[n]
        if st["exp_avg"].data_ptr() == model._adam_m[o:].data_ptr():
            continue                                   # already a view of the arena (publish_optimizer_state)
        model._adam_m[o:o + p.numel()].copy_(st["exp_avg"].reshape(-1).to(model._adam_m.device, torch.float32))
        model._adam_v[o:o + p.numel()].copy_(st["exp_avg_sq"].reshape(-1).to(model._adam_v.device, torch.float32))
        model._adam_steps[model.adam_segment_of(n)] = int(float(st["step"]))


def fused_adamw_step(model: ATMS, optimizer=None, use_shared: bool = False, device_steps=None, subjects=None) -> None:
    """torch.optim.AdamW semantics on model.flat_params / model.flat_grads.  Parameters that received no gradient
    this step (the unused half of {subject table, shared token}; in the joint-subject model the value embeddings of
    subjects outside ``subjects``; the never-used cold parameters) are skipped, as torch.optim does for ``grad is None``."""
    hp = _adam_hparams(optimizer)
    if model._adam_m is None:
        model._adam_m = torch.zeros_like(model.flat_grads)
        model._adam_v = torch.zeros_like(model.flat_grads)
    for name, off, n in model.adam_segments(use_shared, subjects):
        if device_steps is not None:
            # CUDA-graph path: the step number lives in device memory and is advanced inside the captured graph
            device_steps[name].add_(1)
            _lib.adamw_step_dev(model.flat_params[off:], model.flat_grads[off:], model._adam_m[off:], model._adam_v[off:], n,
                                hp["lr"], hp["betas"][0], hp["betas"][1], hp["eps"], hp["weight_decay"], device_steps[name])
            continue
        model._adam_steps[name] += 1
        _lib.adamw_step(model.flat_params[off:], model.flat_grads[off:], model._adam_m[off:], model._adam_v[off:], n,
                        hp["lr"], hp["betas"][0], hp["betas"][1], hp["eps"], hp["weight_decay"], model._adam_steps[name])


def publish_optimizer_state(model: ATMS, optimizer) -> None:
    """expose the fused moments through optimizer.state (views) so optimizer.state_dict() stays meaningful"""
    if optimizer is None or model._adam_m is None:
        return
    named = dict(model.named_parameters())
    for n in model._hot_order():
        p = named[n]
        steps = model._adam_steps[model.adam_segment_of(n)]
        if steps == 0:
            continue
        o = model._offs[n]
        optimizer.state[p] = {
            "step": torch.tensor(float(steps)),
            "exp_avg": model._adam_m[o:o + p.numel()].view(p.shape),
            "exp_avg_sq": model._adam_v[o:o + p.numel()].view(p.shape),
        }


# ------------------------------------------------------------------------------------------------
# one contrastive step (forward, 2x InfoNCE, backward, optimiser), optionally data-parallel
# ------------------------------------------------------------------------------------------------
class _PeerSum:
    """symmetric NVLink buffer + device-side sequence counter behind eegb200_peer_sum_f64 (one per process).
    EEGB200_PEER_SYNCBN=0 keeps the SyncBatchNorm exchange on NCCL."""
    _instance = None
    _failed = False

    @classmethod
    def create(cls, device, rank, world):
        import os
        if cls._instance is not None or cls._failed:
            return cls._instance
        if (os.environ.get("EEGB200_PEER_SYNCBN", "1") == "0" or device.type != "cuda" or world > 16
                or not torch.distributed.is_initialized() or torch.distributed.get_backend() != "nccl"):
            cls._failed = True
            return None
        try:
            import ctypes
            import torch.distributed._symmetric_memory as symm_mem
            nbytes = _lib.peer_sum_buffer_bytes()
            with torch.cuda.device(device):
                buf = symm_mem.empty(nbytes, dtype=torch.uint8, device=device)
                buf.zero_()
                hdl = symm_mem.rendezvous(buf, torch.distributed.group.WORLD)
                ptrs = [int(x) for x in hdl.buffer_ptrs]
                if len(ptrs) != world or any(x == 0 for x in ptrs):
                    raise RuntimeError("symmetric memory: peer pointers unavailable")
                self = cls()
                self.buf, self.hdl, self.rank, self.world = buf, hdl, rank, world
                self.ptr_array = (ctypes.c_void_p * world)(*ptrs)
                self.seq = torch.zeros(1, dtype=torch.int64, device=device)
                self.err = torch.zeros(1, dtype=torch.int32, device=device)
                torch.cuda.synchronize(device)
            torch.distributed.barrier()          # every rank's buffer is zeroed before anyone pushes into it
            ok = torch.ones(1, device=device)
        except Exception as e:                    # no fabric / IPC support: fall back to NCCL on every rank alike
            import warnings
            warnings.warn(f"eeg_image_decode_b200: NVLink peer SyncBN exchange unavailable ({e!r}); using NCCL")
            self, ok = None, torch.zeros(1, device=device)
        torch.distributed.all_reduce(ok, op=torch.distributed.ReduceOp.MIN)      # all ranks or none
        if ok.item() < 1:
            cls._failed = True
            return None
        cls._instance = self
        return self

    def sum_(self, t):
        _lib.peer_sum_f64(t, self.ptr_array, self.rank, self.world, self.seq, self.err)

    def check(self):
        """host-side check after a synchronisation point: did every exchange complete?"""
        if int(self.err.item()) != 0:
            raise RuntimeError("eegb200_peer_sum_f64: a rank never arrived at a SyncBatchNorm exchange")


class StepEngine:
    """``variant``: "retrieval" -- alpha*ClipLoss(img) + (1-alpha)*ClipLoss(txt), alpha = 0.99 (ATMS_retrieval.py:229-234);
    "reconstruction" -- alpha*10*MSE(eeg, img) + (1-alpha)*10*ClipLoss(img), alpha = 0.90
    (Generation/ATMS_reconstruction.py:198, 224-228)."""

    def __init__(self, model: ATMS, optimizer=None, alpha: float = 0.99, variant: str = "retrieval"):
        if variant not in ("retrieval", "reconstruction"):
            raise ValueError(f"unknown step variant {variant!r}")
        self.model = model
        self.optimizer = optimizer
        self.alpha = alpha
        self.variant = variant
        self.nce = _InfoNCE()
        self.world, self.rank = _world()
        self.fused = fused_optimizer_ok(model, optimizer)
        if self.fused:
            adopt_optimizer(model, optimizer)
        self._peer = None
        if self.world > 1:
            self._peer = _PeerSum.create(model.flat_params.device, self.rank, self.world)

    def _allreduce(self, t):
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.SUM)

    def check_collectives(self):
        """call at a host synchronisation point: raises if a peer-memory exchange gave up waiting for a rank"""
        if self._peer is not None:
            self._peer.check()

    def _allreduce_stats(self, t):
        """SyncBatchNorm sums (80 doubles): one-shot exchange over NVLink peer memory when available (csrc/peer_sum.cu),
        else NCCL"""
        if self._peer is not None and t.dtype == torch.float64 and t.numel() <= 256:
            self._peer.sum_(t)
        else:
            self._allreduce(t)

    def _allreduce_pair(self, a, b):
        """two all-reduces as ONE NCCL group launch (the encoder part of the gradient arena and the subject table sit on
        either side of the bucket that is already in flight)"""
        cm = getattr(torch.distributed, "_coalescing_manager", None)
        if cm is not None and a.is_cuda and not getattr(self, "_no_coalesce", False):
            try:
                with cm(device=a.device):
                    torch.distributed.all_reduce(a)
                    torch.distributed.all_reduce(b)
                return
            except Exception:            # older / different process-group back ends: fall back for good
                if torch.cuda.is_current_stream_capturing():
                    raise
                self._no_coalesce = True
        self._allreduce(a)
        self._allreduce(b)

    def gather_async(self, img_feat, txt_feat):
        """start the all-gather of the (input) target blocks; it overlaps the encoder forward.  Returns what
        loss_and_grad(gathered=...) consumes."""
        W = self.world
        if W == 1:
            return None
        # Each rank rounds its own block to TF32 (what the loss kernel would otherwise do to all W blocks) and the
        # all-gather lands directly in the logits GEMM's operand inside the loss workspace: no copy of the 2 x N x D targets
        B, D = img_feat.shape
        nt = 2 if self.variant == "retrieval" else 1
        slots = self.nce.target_slots(B, W * B, D, nt, img_feat.device)
        out = []
        for i, t in enumerate((img_feat, txt_feat if nt == 2 else None)):
            if t is None:
                out.append((None, None))
                continue
            r = torch.empty_like(t, memory_format=torch.contiguous_format)
            _lib.tf32_round(t.contiguous(), r)
            out.append((slots[i], torch.distributed.all_gather_into_tensor(slots[i], r, async_op=True)))
        return out

    def loss_and_grad(self, feats, img_feat, txt_feat, need_grad=True, gathered=None):
        """this rank's share of the step loss ([3] device tensor: total, ClipLoss(img), second term) and d loss / d feats,
        d loss / d logit_scale.  Targets are the LOCAL rows; the global-batch gather happens here (or was started by
        gather_async)."""
        m, W = self.model, self.world
        if gathered is not None:
            for _, h in gathered:
                if h is not None:
                    h.wait()
            img_all = gathered[0][0]
        else:
            img_all = gather_targets(img_feat, W)
        row0 = self.rank * feats.shape[0]
        if self.variant == "retrieval":
            txt_all = gathered[1][0] if gathered is not None else gather_targets(txt_feat, W)
            return fused_contrastive(self.nce, feats, img_all, txt_all, m.logit_scale.detach(), self.alpha,
                                     row_offset=row0, need_grad=need_grad, world_size=W)
        w_clip, w_mse = (1.0 - self.alpha) * 10.0, self.alpha * 10.0
        loss, d_e, d_s = self.nce.run(feats, img_all, None, m.logit_scale.detach(), w_clip, 0.0, row0, need_grad,
                                      world_size=W)
        # nn.MSELoss() is the mean over the whole (global) batch; each rank adds the share of its own rows
        _lib.mse(feats, img_feat.contiguous(), img_all.shape[0], w_mse, 1.0, loss=loss[0:1], loss_term=loss[2:3], d_eeg=d_e)
        return loss, d_e, d_s

    def step(self, eeg, subject_ids, img_feat, txt_feat, use_shared: bool, seed: Optional[int] = None, device_steps=None,
             known_subject: Optional[int] = None, after_forward=None, before_update=None):
        """returns (loss[3] device tensor -- this rank's share, embeddings [B,1024]).  ``known_subject``: see ATMS.encode.
        ``after_forward(feats)`` / ``before_update()``: optional hooks right after the embeddings exist and right before the
        optimiser touches the parameters (GraphedTrainStep scores the train accuracy on a side stream between them)."""
        m = self.model
        W = self.world
        m.zero_flat_grads()
        seed = m.next_seed() if seed is None else seed
        gathered = None
        if W > 1:
            gathered = self.gather_async(img_feat, txt_feat)   # inputs only: travels while the encoder runs
            seed ^= (self.rank + 1) * 0x9E3779B97F4A7C15 & 0x3FFFFFFFFFFFFFFF     # decorrelate the ranks' dropout masks
            # SyncBN: all-reduce the batch statistics between the forward phases
            m.encode(eeg, subject_ids, train=True, seed=seed, phases=_lib.PHASE_A, known_subject=known_subject)
            self._allreduce_stats(m.ws_tensor("bn1_sums"))
            out = m._last[3]
            self._phase(_lib.PHASE_B, fwd=True, batch_scale=W)
            self._allreduce_stats(m.ws_tensor("bn2_sums"))
            self._phase(_lib.PHASE_C, fwd=True, batch_scale=W)
            for bn in (m.enc_eeg[0].tsconv[2], m.enc_eeg[0].tsconv[5]):
                bn.num_batches_tracked.add_(1)
            feats = out
        else:
            feats = m.encode(eeg, subject_ids, train=True, seed=seed, known_subject=known_subject)
        if after_forward is not None:
            after_forward(feats)
        loss, d_e, d_s = self.loss_and_grad(feats, img_feat, txt_feat, gathered=gathered)
        m.grad_view("logit_scale").add_(d_s)
        if W > 1:
            m.backprop(d_e, phases=_lib.PHASE_A)
            # Gradient all-reduce in two buckets.  After phase A the projector, the conv head and logit_scale are final:
            # that contiguous tail of the arena is 2.6 M of the 3.2 M parameters (10 MB) and its all-reduce (NCCL's own
            # stream) hides behind phases B and C; only the encoder / conv part (2.5 MB) is exposed at the end.
            o_tail = m._offs["enc_eeg.0.projection.0.weight"]
            o_tab = m._offs[_lib.P_NAMES[_lib.P_SUBJ_TABLE]]
            h_tail = torch.distributed.all_reduce(m.flat_grads[o_tail:o_tab], async_op=True)
            self._allreduce_stats(m.ws_tensor("bn2_bwd_sums"))
            self._phase(_lib.PHASE_B, fwd=False, batch_scale=W)
            self._allreduce_stats(m.ws_tensor("bn1_bwd_sums"))
            self._phase(_lib.PHASE_C, fwd=False, batch_scale=W)
            self._allreduce_pair(m.flat_grads[:o_tail], m.flat_grads[o_tab:])
            h_tail.wait()
        else:
            m.backprop(d_e)
        subjects = m._last_subjects          # joint-subject model: whose value embeddings got a gradient
        if W > 1 and subjects is not None and known_subject is None:
            raise NotImplementedError("data-parallel steps of the joint-subject model need known_subject (every rank must "
                                      "update the same value embeddings)")
        if before_update is not None:
            before_update()
        if self.fused:
            fused_adamw_step(m, self.optimizer, use_shared, device_steps, subjects)
        else:
            self._generic_optimizer_step(use_shared, subjects)
        return loss, feats

    def _phase(self, phase, fwd, batch_scale):
        # the C side derives the BatchNorm element count from its local B; under SyncBN the reduced sums cover
        # W*B samples, so the count must be scaled.  B is patched in the io for the BN-count only.
        m = self.model
        io = m._last[0]
        dev = m.flat_params.device
        if fwd:
            _lib.atms_forward(io, phase | (_BN_SCALE_SHIFT(batch_scale)), dev)
        else:
            G = m._pointers()[1]
            import ctypes
            _lib.atms_backward(io, None, ctypes.cast(G, ctypes.POINTER(ctypes.c_void_p)),
                               phase | (_BN_SCALE_SHIFT(batch_scale)), dev)

    def _generic_optimizer_step(self, use_shared, subjects=None):
        named = dict(self.model.named_parameters())
        live = {seg[0] for seg in self.model.adam_segments(use_shared, subjects)}
        for n in self.model._hot_order():
            on = self.model.adam_segment_of(n) in live and named[n].requires_grad
            named[n].grad = self.model.grad_view(n) if on else None
        self.optimizer.step()


class GraphedTrainStep:
    """The whole training step (forward, 2x InfoNCE, backward, AdamW, train-accuracy scoring: ~90 kernel launches)
    captured once into a CUDA graph and replayed.  The first two calls run eagerly (they create workspaces, optimiser
    state and kernel attributes); the third call captures -- capture does not execute -- and replays.  Everything that
    changes from step to step lives in device memory: inputs are copied into static buffers, the dropout seed is
    `seed + device counter`, the AdamW step numbers are device counters advanced inside the graph.
    Data-parallel steps (NCCL all-gather / all-reduce inside the step) are captured as well.  Falls back to eager
    execution for foreign optimisers and odd batch sizes."""

    def __init__(self, eng: "StepEngine", gallery, use_shared: bool, enabled: bool = True, known_subject: Optional[int] = None):
        import os
        self.eng, self.gallery, self.use_shared = eng, gallery, use_shared
        # joint-subject model: the graph bakes in which value embedding is used, so the caller must pin the subject
        self.known_subject = known_subject
        if eng.model.joint_train and known_subject is None:
            enabled = False
        self.enabled = enabled and os.environ.get("EEGB200_CUDA_GRAPH", "1") != "0"
        self.dp_ok = os.environ.get("EEGB200_CUDA_GRAPH_DP", "1") != "0"
        self.acc_overlap = os.environ.get("EEGB200_ACC_OVERLAP", "1") != "0"
        self._side = None
        self.graph = None
        self.calls = 0
        self.B = None
        self.launches_per_replay = 0
        self.replays = 0

    def _body(self, eeg, sid, img, txt, labels, device_steps):
        # Train accuracy (ATMS_retrieval.py:241-250): argmax over the class-prototype gallery.  The reference scores after
        # optimizer.step() with the updated logit_scale, but argmax is invariant under a positive scale and the embeddings
        # are the pre-update ones (:223), so the 3xTF32 scoring GEMM (~105 us) runs on a side stream right after the
        # forward, overlapped with the backward, instead of extending the critical path after AdamW; it is joined before
        # the optimiser writes logit_scale.  (A non-positive logit_scale would flip the ranking: init is ln(1/0.07) = 2.66
        # and the reference would equally break there.)  EEGB200_ACC_OVERLAP=0 restores the serial order.
        model = self.eng.model
        if not self.acc_overlap:
            loss, feats = self.eng.step(eeg, sid, img, txt, self.use_shared, device_steps=device_steps,
                                        known_subject=self.known_subject)
            r = _lib.retrieval(feats, self.gallery, model.logit_scale.detach(), labels=labels, want_top5=False)
            return loss, feats, r["correct"]
        main = torch.cuda.current_stream()
        if self._side is None:
            self._side = torch.cuda.Stream(device=eeg.device)
        side, box = self._side, {}

        def after_forward(feats):
            side.wait_stream(main)
            with torch.cuda.stream(side):
                box["r"] = _lib.retrieval(feats, self.gallery, model.logit_scale.detach(), labels=labels, want_top5=False)

        def before_update():
            main.wait_stream(side)

        loss, feats = self.eng.step(eeg, sid, img, txt, self.use_shared, device_steps=device_steps,
                                    known_subject=self.known_subject, after_forward=after_forward, before_update=before_update)
        return loss, feats, box["r"]["correct"]

    def __call__(self, eeg, sid, img, txt, labels):
        eng = self.eng
        m = eng.model
        B = eeg.shape[0]
        # data-parallel steps are captured too (NCCL collectives are graph-capturable); EEGB200_CUDA_GRAPH_DP=0 opts out
        ok = self.enabled and eng.fused and (eng.world == 1 or self.dp_ok)
        if self.calls == 0:
            self._nominal_B = B             # the loader's batch size; a ragged batch must not become the captured shape
        if not ok or (self.graph is None and (self.calls < 2 or B != self._nominal_B)) or (self.B is not None and B != self.B):
            self.calls += 1
            return self._body(eeg, sid, img, txt, labels, None)
        if self.graph is None:
            self.B = B
            m._ws_pinned.add(B)          # the graph bakes in pointers into this workspace: it must outlive other batch sizes
            self._ws_keep = m.workspace(B)
            self.s = [t.clone() for t in (eeg, sid, img, txt, labels)]
            # single GPU, retrieval loss: the static target buffers ARE the logits GEMM's operand inside the loss workspace
            # and the per-step copy-in rounds to TF32 on the way (one kernel instead of a memcpy + the loss's own
            # round-and-copy pass per target)
            self._tgt_in_place = (eng.world == 1 and eng.variant == "retrieval" and img.dtype == torch.float32
                                  and txt.dtype == torch.float32 and img.dim() == 2 and img.shape == txt.shape)
            if self._tgt_in_place:
                slots = eng.nce.target_slots(B, B, img.shape[1], 2, img.device)
                for i, t in ((2, img), (3, txt)):
                    _lib.tf32_round(t.contiguous(), slots[i - 2])
                    self.s[i] = slots[i - 2]
            self.dev_steps = {k: torch.tensor([v], dtype=torch.int64, device=eeg.device) for k, v in m._adam_steps.items()}
            self._dev_host = dict(m._adam_steps)     # what the device counters hold, mirrored on the host
            torch.cuda.synchronize()
            n0 = _lib.launch_count()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self.out = self._body(*self.s, self.dev_steps)
                m._seed_ctr.add_(1)
            self.launches_per_replay = _lib.launch_count() - n0
            self.graph = g
        else:
            for i, (dst, src) in enumerate(zip(self.s, (eeg, sid, img, txt, labels))):
                if self._tgt_in_place and i in (2, 3):
                    _lib.tf32_round(src.contiguous(), dst)
                else:
                    dst.copy_(src, non_blocking=True)
        # steps taken outside this graph (a ragged last batch run eagerly, another captured step on the same model) moved
        # the host-side AdamW step numbers: bring the device counters back in line before replaying
        for name, v in m._adam_steps.items():
            if self._dev_host.get(name) != v:
                self.dev_steps[name].fill_(v)
        self.graph.replay()
        self.replays += 1
        for name, _, _ in m.adam_segments(self.use_shared, [self.known_subject] if self.known_subject is not None else None):
            m._adam_steps[name] += 1
        self._dev_host = dict(m._adam_steps)
        return self.out


class _StageSlot:
    """one set of device-side input buffers (EEG, labels, text / image targets) of the host->device staging ring"""

    def __init__(self, device):
        self.device = device
        self.bufs = {}
        self.free_ev = None

    def _buf(self, name, src, dtype):
        key = (name, tuple(src.shape))
        b = self.bufs.get(key)
        if b is None:
            self.bufs = {k: v for k, v in self.bufs.items() if k[0] != name}       # one shape per input is kept
            b = torch.empty(tuple(src.shape), dtype=dtype, device=self.device)
            self.bufs[key] = b
        b.copy_(src, non_blocking=True)
        return b

    def load(self, eeg, labels, txt, img):
        return (self._buf("eeg", eeg, torch.float32), self._buf("labels", labels, labels.dtype),
                self._buf("txt", txt, torch.float32), self._buf("img", img, torch.float32))


class _StagingRing:
    def __init__(self, device, n_slots=3):
        self.device = device
        self.copy_stream = torch.cuda.Stream(device=device)
        self.slots = [_StageSlot(device) for _ in range(n_slots)]
        self.i = 0
        self._sid = {}

    def next_slot(self):
        self.i = (self.i + 1) % len(self.slots)
        return self.slots[self.i]

    def subject_ids(self, batch_size, subject_id):
        key = (int(batch_size), int(subject_id))
        t = self._sid.get(key)
        if t is None:
            if len(self._sid) > 16:
                self._sid.clear()
            t = torch.full((batch_size,), subject_id, dtype=torch.long, device=self.device)
            self._sid[key] = t
        return t


def _staging_ring(model, device) -> _StagingRing:
    ring = model.__dict__.get("_stage_ring")
    if ring is None or ring.device != device:
        ring = _StagingRing(device)
        model.__dict__["_stage_ring"] = ring
    return ring


def _device_gallery(model, img_features_all, device):
    """``img_features_all[::10]`` as a contiguous fp32 device tensor.  main_train_loop hands the same table to every
    epoch: the strided gather + upload (~3 ms for 1654 x 1024) is done once per table (keyed by storage, shape and the
    tensor's in-place version counter) instead of once per train_model() call."""
    key = (img_features_all.data_ptr(), img_features_all._version, tuple(img_features_all.shape), img_features_all.dtype,
           str(img_features_all.device), str(device))
    hit = model.__dict__.get("_gallery_cache")
    if hit is not None and hit[0] == key:
        return hit[1]
    g = (img_features_all[::10]).to(device).float().contiguous()
    model.__dict__["_gallery_cache"] = (key, g, img_features_all)      # keeps the source alive: its data_ptr stays unique
    return g


def _cached_graphed_step(model: ATMS, optimizer, alpha, variant, use_shared, known_subject, gallery) -> "GraphedTrainStep":
    """The captured step survives across train_model() calls (one per epoch in main_train_loop): everything the graph
    bakes in is part of the key -- optimiser object and hyper-parameters, loss variant, token branch, dropout rates,
    gallery shape -- and the gallery itself is copied into the graph's static buffer.  Without this every epoch paid two
    eager steps plus a re-capture (~9 ms, 12 % of a 20-step epoch at B = 1024)."""
    world, _ = _world()
    fused = fused_optimizer_ok(model, optimizer)
    if fused:
        adopt_optimizer(model, optimizer)                 # imports a freshly loaded optimizer.state_dict() before stepping
    hp = _adam_hparams(optimizer) if fused else None      # foreign optimisers step eagerly: nothing of theirs is baked in
    hp_key = (hp["lr"], tuple(hp["betas"]), hp["eps"], hp["weight_decay"]) if hp else None
    key = (id(optimizer), type(optimizer).__name__, variant, float(alpha), bool(use_shared), known_subject,
           tuple(gallery.shape), tuple(model.dropout_p), hp_key, world)
    import os
    gstep = model._gstep_cache.get(key) if os.environ.get("EEGB200_STEP_CACHE", "1") != "0" else None
    if gstep is None:
        model._gstep_cache.clear()          # one live graph per model: drop the previous configuration
        eng = StepEngine(model, optimizer, alpha, variant)
        gstep = GraphedTrainStep(eng, gallery.clone(), use_shared, known_subject=known_subject)
        model._gstep_cache[key] = gstep
    else:
        gstep.gallery.copy_(gallery)
    return gstep


def _BN_SCALE_SHIFT(world: int) -> int:
    """phase-mask bits 8..15 carry the SyncBN world size (0/1 = local statistics)"""
    return (int(world) & 0xFF) << 8


# ------------------------------------------------------------------------------------------------
def train_model(sub, eeg_model, dataloader, optimizer, device, text_features_all, img_features_all, config, *,
                step_callback=None):
    """One epoch.  Returns (average_loss, accuracy, features[n_seen,1024]) like ATMS_retrieval.py:199-254.  Under
    torch.distributed every rank feeds its shard (equal batch sizes on all ranks): loss and accuracy are the global-batch
    values on every rank, the features are the rank's own rows.
    ``step_callback(step_index, loss[3] on the HOST)`` (optional, keyword only) receives every step's loss (mix, image,
    text share), read back from the device like the reference does (:238) but delivered one step late so that the
    read-back does not stall the next launch.
    Also serves the joint-subject model (Retrieval/ATMS_retrieval_joint_train.py:201-254, same body): every trial then
    goes through the value embedding of the subject parsed from ``sub``."""
    return _train_epoch(sub, eeg_model, dataloader, optimizer, device, text_features_all, img_features_all, config,
                        variant="retrieval", alpha=0.99, step_callback=step_callback)


def _train_epoch(sub, eeg_model, dataloader, optimizer, device, text_features_all, img_features_all, config, *,
                 variant, alpha, step_callback=None):
    eeg_model.train()
    device = torch.device(device)
    if device.type != "cuda":
        raise RuntimeError("train_model: this implementation runs on CUDA only (no CPU fallback)")
    # (:201 moves text_features_all to the device as well; nothing in the step reads it -- the text logits are
    # commented out in the reference, :240-245 -- so that transfer is skipped)
    img_features_all = _device_gallery(eeg_model, img_features_all, device)         # :202 class-prototype gallery [::10]
    subject_id = extract_id_from_string(sub)
    use_shared = subject_id is None or subject_id >= N_SUBJECT_ROWS or subject_id < 0
    known_subject = None
    if eeg_model.joint_train:
        if use_shared:
            raise KeyError(str(subject_id))     # the reference: self.value_embedding[str(subject_id.item())] (Embed.py:144)
        known_subject = subject_id
    gstep = _cached_graphed_step(eeg_model, optimizer, alpha, variant, use_shared, known_subject, img_features_all)
    eng = gstep.eng
    loss_acc = torch.zeros(3, device=device)
    correct = torch.zeros(1, device=device, dtype=torch.int32)
    total = 0
    features_list = []
    n_batches = 0
    # host->device copies of batch i+1 run on a side stream while batch i computes (the reference copies synchronously
    # at the top of every step, :210-213), into a small ring of device buffers that lives on the model: no allocator
    # traffic in the loop (fresh `.to(device)` tensors per step cost ~6 cudaMalloc calls per 20-step epoch and the
    # reserved pool grew by 240 MB per epoch, measured on B200)
    ring = _staging_ring(eeg_model, device)
    copy_stream = ring.copy_stream
    main_stream = torch.cuda.current_stream(device)
    copy_stream.wait_stream(main_stream)

    def stage(batch):
        eeg_data, labels, text, text_features, img, img_features = batch
        if eeg_data.is_cuda:      # device-resident loader (data.py): nothing to stage
            return (eeg_data, labels.to(device), text_features.to(device).float(), img_features.to(device).float()), None, None
        slot = ring.next_slot()
        with torch.cuda.stream(copy_stream):
            if slot.free_ev is not None:
                copy_stream.wait_event(slot.free_ev)        # the step that read this slot last has been enqueued before
            t = slot.load(eeg_data, labels, text_features, img_features)
            ev = torch.cuda.Event()
            ev.record(copy_stream)
        return t, ev, slot

    # per-step loss read-back (step_callback): the value is copied to pinned host memory right after the step and handed
    # to the callback one step later, so the host never stalls the launch of the next step
    pending = []            # (step index, pinned tensor, event)
    pin_ring = torch.empty(8, 3, dtype=torch.float32, pin_memory=True) if step_callback is not None else None

    def flush(keep: int):
        while len(pending) > keep:
            i_, host_, ev_ = pending.pop(0)
            ev_.synchronize()
            step_callback(i_, host_.clone())

    it = iter(dataloader)
    nxt = next(it, None)
    staged = stage(nxt) if nxt is not None else None
    batch_idx = -1
    while staged is not None:
        batch_idx += 1
        (eeg_data, labels, text_features, img_features), ev, slot = staged
        nxt = next(it, None)
        staged = stage(nxt) if nxt is not None else None
        if ev is not None:
            main_stream.wait_event(ev)
        batch_size = eeg_data.size(0)
        subject_ids = ring.subject_ids(batch_size, subject_id if subject_id is not None else -1)
        # step + train accuracy against the 1654-way prototype gallery (:241-250, scored with the post-update logit_scale)
        loss, eeg_features, n_ok = gstep(eeg_data, subject_ids, img_features, text_features, labels)
        if slot is not None:
            slot.free_ev = torch.cuda.Event()
            slot.free_ev.record(main_stream)
        loss_acc += loss
        features_list.append(eeg_features.clone())
        correct += n_ok
        total += batch_size
        n_batches += 1
        if step_callback is not None:
            host = pin_ring[batch_idx % 8]          # at most 2 copies are in flight (flush(keep=1) below)
            host.copy_(loss, non_blocking=True)
            ev_l = torch.cuda.Event()
            ev_l.record(main_stream)
            pending.append((batch_idx, host, ev_l))
            flush(keep=1)
    if step_callback is not None:
        flush(keep=0)
    if n_batches == 0:
        raise RuntimeError("train_model: empty dataloader")
    if eng.world > 1:
        # loss shares sum to the global-batch loss; accuracy is reported over the global batch too (the features returned
        # below stay this rank's rows)
        counts = torch.stack([correct[0].to(torch.float64), torch.tensor(float(total), device=device, dtype=torch.float64)])
        torch.distributed.all_reduce(loss_acc)
        torch.distributed.all_reduce(counts)
        correct, total = counts[0:1], int(counts[1].item())
        eng.check_collectives()
    publish_optimizer_state(eeg_model, optimizer if eng.fused else None)
    average_loss = float(loss_acc[0].item()) / n_batches
    accuracy = int(correct.item()) / total
    return average_loss, accuracy, torch.cat(features_list, dim=0)


def evaluate_model(sub, eeg_model, dataloader, device, text_features_all, img_features_all, k, config):
    """k-way zero-shot retrieval.  Returns (average_loss, accuracy, top5_acc) like ATMS_retrieval.py:258-362.
    The candidate sets are drawn with the same ``random.sample`` calls, in the same order, as the reference."""
    return _evaluate_epoch(sub, eeg_model, dataloader, device, text_features_all, img_features_all, k, config,
                           variant="retrieval", alpha=0.99)


EVAL_CHUNK = 1024       # trials per eval-mode forward launch chain


def _drain(dataloader):
    """the loader's batches as lists (EEG, labels, text features, image features); nothing is computed yet"""
    eeg, labels, txt, img = [], [], [], []
    for (eeg_data, lab, _text, text_features, _img, img_features) in dataloader:
        eeg.append(eeg_data)
        labels.append(lab)
        txt.append(text_features)
        img.append(img_features)
    return eeg, labels, txt, img


def _encode_batches(eeg_model, eeg_list, device, subject_id, known_subject):
    """Eval-mode embeddings of every trial of a drained loader, EVAL_CHUNK trials per forward instead of one forward
    per loader batch (the reference's test loader has batch size 1: 200 B=1 forwards per evaluate_model call,
    ATMS_retrieval.py:264-279).  Eval-mode BatchNorm uses the running statistics, so a trial's embedding does not depend
    on which other trials share its launch."""
    n = sum(int(e.size(0)) for e in eeg_list)
    feats = torch.empty(n, 1024, device=device, dtype=torch.float32)
    sid = subject_id if subject_id is not None else -1
    i0, cur, cur_n = 0, [], 0

    def flush():
        nonlocal i0, cur, cur_n
        if not cur:
            return
        x = cur[0].to(device) if len(cur) == 1 else torch.cat([e.to(device) for e in cur], dim=0)
        ids = torch.full((cur_n,), sid, dtype=torch.long, device=device)
        eeg_model.encode(x, ids, train=False, known_subject=known_subject, out=feats[i0:i0 + cur_n])
        i0, cur, cur_n = i0 + cur_n, [], 0

    for e in eeg_list:
        if cur_n + int(e.size(0)) > EVAL_CHUNK:
            flush()
        cur.append(e)
        cur_n += int(e.size(0))
    flush()
    return feats


def _eval_losses(eng, feats, sizes, img_list, txt_list, device):
    """sum over the loader's batches of the per-batch loss (the reference averages per-batch losses, :282-293)"""
    total = torch.zeros(3, device=device)
    if all(b == 1 for b in sizes):
        # batch size 1 (the reference's test loader): ClipLoss is exactly 0 -- cross entropy of a 1x1 logit matrix
        # (models/loss.py:137-140); only the MSE term of the reconstruction variant remains, summed by one kernel
        if eng.variant == "reconstruction":
            img = torch.cat([t.to(device).float() for t in img_list], dim=0).contiguous()
            _lib.mse(feats, img, 1, eng.alpha * 10.0, 1.0, loss=total[0:1], loss_term=total[2:3], d_eeg=None)
        return total
    off = 0
    for b, img, txt in zip(sizes, img_list, txt_list):
        loss, _, _ = eng.loss_and_grad(feats[off:off + b], img.to(device).float().contiguous(),
                                       txt.to(device).float().contiguous() if txt is not None else None, need_grad=False)
        total += loss
        off += b
    return total


def _evaluate_epoch(sub, eeg_model, dataloader, device, text_features_all, img_features_all, k, config, *, variant, alpha):
    """shared by the retrieval script (loss alpha*ClipLoss(img) + (1-alpha)*ClipLoss(txt), ATMS_retrieval.py:290-293) and
    the reconstruction script (alpha*10*MSE + (1-alpha)*10*ClipLoss(img), ATMS_reconstruction.py:283-286); the k-way
    scoring below is identical in both (:295-357 / :290-349).
    The reference's loop (one B=1 forward, one loss, k-way scoring per loader batch) is run as: drain the loader, draw the
    candidate sets (same ``random.sample`` calls in the same order), ONE batched forward, ONE scoring launch."""
    eeg_model.eval()
    device = torch.device(device)
    if device.type != "cuda":
        raise RuntimeError("evaluate_model: this implementation runs on CUDA only (no CPU fallback)")
    if k not in (2, 4, 10, 50, 100, 200):
        print("Error.")
    text_features_all = text_features_all.to(device).float()
    img_features_all = img_features_all.to(device).float().contiguous()
    n_cls = text_features_all.size(0)
    all_labels = set(range(n_cls))
    eng = StepEngine(eeg_model, None, alpha, variant)
    eng.world, eng.rank = 1, 0          # evaluation is per process (the reference evaluates on one device), no collectives
    subject_id = extract_id_from_string(sub)
    known_subject = None
    if eeg_model.joint_train:
        if subject_id is None or not 0 <= subject_id < N_SUBJECT_ROWS:
            raise KeyError(str(subject_id))
        known_subject = subject_id
    eeg_list, label_batches, txt_list, img_list = _drain(dataloader)
    if not eeg_list:
        raise RuntimeError("evaluate_model: empty dataloader")
    sizes = [int(e.size(0)) for e in eeg_list]
    # host side: the reference's per-sample candidate draw (:297-300).  For every k other than 200 it draws a second list
    # AFTER the candidate features were gathered (:323-325 and :340-341); the scores still belong to the first list and
    # the true label is last in both, so the second draw only advances the RNG
    label_list, sel_rows = [], []
    for lab in label_batches:
        for label in lab.tolist():
            possible_classes = list(all_labels - {label})
            selected_classes = random.sample(possible_classes, k - 1) + [label]
            if k in (2, 4, 10, 50, 100):
                random.sample(possible_classes, k - 1)
            sel_rows.append(selected_classes)
            label_list.append(label)
    with torch.no_grad():
        feats = _encode_batches(eeg_model, eeg_list, device, subject_id, known_subject)
        total_loss = _eval_losses(eng, feats, sizes, img_list, txt_list if variant == "retrieval" else [None] * len(sizes), device)
        sel = torch.tensor(sel_rows, dtype=torch.int32)
        r = _lib.retrieval(feats, img_features_all, eeg_model.logit_scale.detach(), sel=sel, want_top5=(k >= 50))
    top1 = r["top1"].tolist()
    top5 = r["top5"].tolist() if k >= 50 else None
    correct = top5_correct_count = 0
    for i, label in enumerate(label_list):
        t1 = top1[i]
        if not 0 <= t1 < k:
            raise RuntimeError(f"evaluate_model: invalid argmax {t1} for trial {i} (non-finite scores? logit_scale = "
                               f"{float(eeg_model.logit_scale)})")
        if sel_rows[i][t1] == label:
            correct += 1
        if top5 is not None and label in [sel_rows[i][j] for j in top5[i] if 0 <= j < k]:
            top5_correct_count += 1
    total = len(label_list)
    average_loss = float(total_loss[0].item()) / len(sizes)
    return average_loss, correct / total, top5_correct_count / total


def main_train_loop(sub, current_time, eeg_model, train_dataloader, test_dataloader, optimizer, device,
                    text_features_train_all, text_features_test_all, img_features_train_all, img_features_test_all,
                    config, logger=None):
    """epoch loop of ATMS_retrieval.py:364-512 without the plotting / wandb side effects (``logger`` is any object
    with ``log(dict)``; checkpoints are the caller's business: ``eeg_model.state_dict()`` is reference-compatible)."""
    results = []
    for epoch in range(config.epochs):
        train_loss, train_accuracy, _ = train_model(sub, eeg_model, train_dataloader, optimizer, device,
                                                    text_features_train_all, img_features_train_all, config=config)
        ev = {}
        for k in (200, 2, 4, 10, 50, 100):
            ev[k] = evaluate_model(sub, eeg_model, test_dataloader, device, text_features_test_all,
                                   img_features_test_all, k=k, config=config)
        test_loss, test_accuracy, top5_acc = ev[200]
        epoch_results = {
            "epoch": epoch + 1, "test_loss": test_loss, "test_accuracy": test_accuracy, "v2_acc": ev[2][1],
            "v4_acc": ev[4][1], "v10_acc": ev[10][1], "top5_acc": top5_acc, "v50_acc": ev[50][1],
            "v100_acc": ev[100][1], "v50_top5_acc": ev[50][2], "v100_top5_acc": ev[100][2],
        }
        results.append(epoch_results)
        if logger is not None and hasattr(logger, "log"):
            logger.log({"Train Loss": train_loss, "Train Accuracy": train_accuracy, "Test Loss": test_loss,
                        "Test Accuracy": test_accuracy, "v2 Accuracy": ev[2][1], "v4 Accuracy": ev[4][1],
                        "v10 Accuracy": ev[10][1], "Epoch": epoch})
        print(f"Epoch {epoch + 1}/{config.epochs} - Train Loss: {train_loss:.4f}, Train Accuracy: {train_accuracy:.4f}, "
              f"Test Loss: {test_loss:.4f}, Test Accuracy: {test_accuracy:.4f}, Top5 Accuracy: {top5_acc:.4f}")
    return results
