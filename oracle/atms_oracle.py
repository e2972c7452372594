"""CPU oracle for the ATM-S contrastive hot path.  TEST INFRASTRUCTURE ONLY.

This is a plain-torch (CPU, fp32 or fp64) restatement of the reference algorithm.  Only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import it; the product package ``eeg_image_decode_b200`` never does (it fails loudly
when the CUDA library is missing instead of falling back to this file).

Parity status: PINNED.  The reference ships no tests or golden vectors, so the
oracle is pinned against outputs of the reference itself, imported unmodified from
``/root/reference`` in the build container by ``tests/golden/make_golden.py``; the resulting
fixtures live in ``tests/golden/*.npz`` and ``tests/test_oracle_golden.py`` checks this file
against them (and against the live reference when it is mounted).

Every function cites the reference lines it restates (paths relative to /root/reference).
All tensors are keyed by the reference ``state_dict`` names so a reference checkpoint can be
fed in unchanged.

Dropout: the reference draws masks from torch's global Philox stream, which a fused kernel cannot
reproduce; the oracle therefore takes the *keep masks* explicitly (``masks[site]`` is a {0,1}
tensor; the kept values are scaled by 1/(1-p) exactly like ``nn.Dropout``).  ``masks=None``
means eval mode / p=0.
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import torch
import torch.nn.functional as F

# ----- fixed hyper-parameters of ATM-S (Retrieval/ATMS_retrieval.py:44-59, 97-167) -----
N_CH = 63          # EEG channels = tokens
N_T = 250          # time points = d_model
N_HEAD = 4
D_HEAD = 62        # 250 // 4  (SelfAttention_Family.py:184-185)
D_FF = 256
N_FILT = 40
K_TEMP = 25
K_POOL = 51
S_POOL = 5
N_POOL = 36        # (250-25+1 - 51)//5 + 1
D_FEAT = 1440
D_OUT = 1024
P_DROP_TR = 0.25   # Config.dropout
P_DROP_CONV = 0.5  # PatchEmbedding Dropout(0.5)
P_DROP_PROJ = 0.5  # Proj_eeg drop_proj
EPS = 1e-5
BN_MOMENTUM = 0.1

K = {  # short aliases -> reference state_dict keys
    "Wv": "encoder.enc_embedding.value_embedding.weight",
    "bv": "encoder.enc_embedding.value_embedding.bias",
    "pe": "encoder.enc_embedding.position_embedding.pe",
    "subj": "encoder.enc_embedding.subject_embedding.subject_embedding.weight",
    "shared": "encoder.enc_embedding.subject_embedding.shared_embedding",
    "Wq": "encoder.encoder.attn_layers.0.attention.query_projection.weight",
    "bq": "encoder.encoder.attn_layers.0.attention.query_projection.bias",
    "Wk": "encoder.encoder.attn_layers.0.attention.key_projection.weight",
    "bk": "encoder.encoder.attn_layers.0.attention.key_projection.bias",
    "Wvv": "encoder.encoder.attn_layers.0.attention.value_projection.weight",
    "bvv": "encoder.encoder.attn_layers.0.attention.value_projection.bias",
    "Wo": "encoder.encoder.attn_layers.0.attention.out_projection.weight",
    "bo": "encoder.encoder.attn_layers.0.attention.out_projection.bias",
    "W1": "encoder.encoder.attn_layers.0.conv1.weight",
    "b1": "encoder.encoder.attn_layers.0.conv1.bias",
    "W2": "encoder.encoder.attn_layers.0.conv2.weight",
    "b2": "encoder.encoder.attn_layers.0.conv2.bias",
    "g1": "encoder.encoder.attn_layers.0.norm1.weight",
    "be1": "encoder.encoder.attn_layers.0.norm1.bias",
    "g2": "encoder.encoder.attn_layers.0.norm2.weight",
    "be2": "encoder.encoder.attn_layers.0.norm2.bias",
    "gf": "encoder.encoder.norm.weight",
    "bef": "encoder.encoder.norm.bias",
    "Wt": "enc_eeg.0.tsconv.0.weight",
    "bt": "enc_eeg.0.tsconv.0.bias",
    "bn1_w": "enc_eeg.0.tsconv.2.weight",
    "bn1_b": "enc_eeg.0.tsconv.2.bias",
    "bn1_rm": "enc_eeg.0.tsconv.2.running_mean",
    "bn1_rv": "enc_eeg.0.tsconv.2.running_var",
    "Ws": "enc_eeg.0.tsconv.4.weight",
    "bs": "enc_eeg.0.tsconv.4.bias",
    "bn2_w": "enc_eeg.0.tsconv.5.weight",
    "bn2_b": "enc_eeg.0.tsconv.5.bias",
    "bn2_rm": "enc_eeg.0.tsconv.5.running_mean",
    "bn2_rv": "enc_eeg.0.tsconv.5.running_var",
    "Wc": "enc_eeg.0.projection.0.weight",
    "bc": "enc_eeg.0.projection.0.bias",
    "Wp1": "proj_eeg.0.weight",
    "bp1": "proj_eeg.0.bias",
    "Wp2": "proj_eeg.1.fn.1.weight",
    "bp2": "proj_eeg.1.fn.1.bias",
    "gp": "proj_eeg.2.weight",
    "bep": "proj_eeg.2.bias",
    "logit_scale": "logit_scale",
}

DROPOUT_SITES = {  # site -> (logical mask shape without batch, p)
    "embed": ((64, N_T), P_DROP_TR),        # Embed.py:162
    "attn": ((N_HEAD, 64, 64), P_DROP_TR),  # SelfAttention_Family.py:71
    "res1": ((64, N_T), P_DROP_TR),         # Transformer_EncDec.py:45
    "ffn1": ((64, D_FF), P_DROP_TR),        # Transformer_EncDec.py:48 (token-major view of (B,256,64))
    "ffn2": ((64, N_T), P_DROP_TR),         # Transformer_EncDec.py:49
    "conv": ((N_FILT, N_POOL), P_DROP_CONV),  # ATMS_retrieval.py:108
    "proj": ((D_OUT,), P_DROP_PROJ),        # ATMS_retrieval.py:164
}


def _drop(x, masks, site):
    if masks is None or site not in masks or masks[site] is None:
        return x
    p = DROPOUT_SITES[site][1]
    return x * masks[site].to(x.dtype) * (1.0 / (1.0 - p))


def positional_embedding(n_tok: int, d_model: int = N_T, dtype=torch.float32) -> torch.Tensor:
    """sin/cos table, rows indexed by token (= channel) position.  Embed.py:8-26."""
    pe = torch.zeros(n_tok, d_model, dtype=torch.float32)
    position = torch.arange(0, n_tok).float().unsqueeze(1)
    div_term = (torch.arange(0, d_model, 2).float() * -(math.log(10000.0) / d_model)).exp()
    pe[:, 0::2] = torch.sin(position * div_term)
    pe[:, 1::2] = torch.cos(position * div_term)
    return pe.to(dtype)


JOINT_VALUE_PREFIX = "encoder.enc_embedding.value_embedding."   # + "<subject>.weight" / ".bias" (ModuleDict, Embed.py:128-130)


def joint_value_keys(sd) -> list:
    """state_dict keys of the per-subject value embeddings of the joint-train variant, [] for the plain model"""
    out = []
    for k in sd:
        if k.startswith(JOINT_VALUE_PREFIX):
            tail = k[len(JOINT_VALUE_PREFIX):].split(".")
            if len(tail) == 2 and tail[0].isdigit():
                out.append(k)
    return out


def value_embedding(sd, P, x, subject_ids, dtype):
    """DataEmbedding value path (Embed.py:142-146).  Plain model: one Linear(250,250) over the time axis.
    joint_train=True (model of Retrieval/ATMS_retrieval_joint_train.py:173-176): every trial goes through the Linear of
    its own subject, ``self.value_embedding[str(subject_id.item())](x[i])``; an id without an entry is the reference's
    KeyError."""
    if "Wv" in P:
        return F.linear(x, P["Wv"], P["bv"])
    rows = []
    for i, sj in enumerate(subject_ids.tolist()):
        kw, kb = f"{JOINT_VALUE_PREFIX}{sj}.weight", f"{JOINT_VALUE_PREFIX}{sj}.bias"
        if kw not in sd:
            raise KeyError(str(sj))
        rows.append(F.linear(x[i], sd[kw].to(dtype), sd[kb].to(dtype)))
    return torch.stack(rows)


def atms_forward(sd: Dict[str, torch.Tensor], x: torch.Tensor, subject_ids: torch.Tensor,
                 train: bool = False, masks: Optional[Dict[str, torch.Tensor]] = None,
                 dtype=torch.float32, update_running_stats: bool = False) -> Dict[str, torch.Tensor]:
    """ATMS.forward (Retrieval/ATMS_retrieval.py:182-191) with every stage output returned.

    ``train`` selects BatchNorm batch statistics (``nn.Module.train()``); dropout is active only
    through ``masks``.  Returns a dict of intermediates; ``out`` is the (B,1024) embedding.
    """
    P = {a: sd[k].to(dtype) for a, k in K.items() if k in sd}
    x = x.to(dtype)
    B = x.shape[0]
    r: Dict[str, torch.Tensor] = {}

    # ---- DataEmbedding.forward  (Embed.py:141-162) ----
    v = value_embedding(sd, P, x, subject_ids, dtype)       # :143-146  (B,63,250), Linear over the time axis
    v = v + P["pe"][:, :N_CH]                               # :149  pe indexed by channel-token
    # SubjectEmbedding.forward (Embed.py:116-121): ANY id >= num_embeddings -> shared token for the whole batch
    n_subj = P["subj"].shape[0]
    if bool(torch.any(subject_ids >= n_subj)):
        tok = P["shared"].expand(B, 1, -1)
    else:
        tok = P["subj"][subject_ids].unsqueeze(1)
    h0 = torch.cat([tok, v], dim=1)                         # :160  (B,64,250)
    h0 = _drop(h0, masks, "embed")                          # :162
    r["h0"] = h0

    # ---- AttentionLayer.forward + FullAttention.forward (SelfAttention_Family.py:194-213, 56-75) ----
    q = F.linear(h0, P["Wq"], P["bq"]).view(B, 64, N_HEAD, D_HEAD)
    k = F.linear(h0, P["Wk"], P["bk"]).view(B, 64, N_HEAD, D_HEAD)
    vv = F.linear(h0, P["Wvv"], P["bvv"]).view(B, 64, N_HEAD, D_HEAD)
    r["q"], r["k"], r["v"] = q, k, vv
    scores = torch.einsum("blhe,bshe->bhls", q, k)          # :59
    A = torch.softmax(scores * (1.0 / math.sqrt(D_HEAD)), dim=-1)   # :71, scale = 1/sqrt(E)
    r["attn_p"] = A
    A = _drop(A, masks, "attn")
    o = torch.einsum("bhls,bshd->blhd", A, vv).reshape(B, 64, N_HEAD * D_HEAD)   # :72, :211
    r["attn_o"] = o
    a = F.linear(o, P["Wo"], P["bo"])                       # :213

    # ---- EncoderLayer.forward (Transformer_EncDec.py:39-51) ----
    x1 = F.layer_norm(h0 + _drop(a, masks, "res1"), (N_T,), P["g1"], P["be1"], EPS)      # :45-47
    r["x1"] = x1
    u = F.linear(x1, P["W1"].squeeze(-1), P["b1"])          # :48 Conv1d k=1 over the feature axis == per-token Linear
    r["ffn_u"] = u
    hf = _drop(F.gelu(u), masks, "ffn1")                    # exact erf GELU (F.gelu default)
    y = _drop(F.linear(hf, P["W2"].squeeze(-1), P["b2"]), masks, "ffn2")   # :49
    x2 = F.layer_norm(x1 + y, (N_T,), P["g2"], P["be2"], EPS)             # :51
    r["x2"] = x2
    # ---- Encoder.forward final norm (Transformer_EncDec.py:77-78) + iTransformer slice (ATMS_retrieval.py:91) ----
    x3 = F.layer_norm(x2, (N_T,), P["gf"], P["bef"], EPS)
    r["x3"] = x3
    enc = x3[:, :N_CH, :]                                   # keeps [subject token, ch0..ch61]
    r["enc"] = enc

    # ---- PatchEmbedding.forward (ATMS_retrieval.py:97-125) ----
    c = F.conv2d(enc.unsqueeze(1), P["Wt"], P["bt"])        # :102  (B,40,63,226)
    c = F.avg_pool2d(c, (1, K_POOL), (1, S_POOL))           # :103  (B,40,63,36)
    r["y1"] = c
    new_stats = {}
    if train:
        m1 = c.mean(dim=(0, 2, 3))
        v1 = c.var(dim=(0, 2, 3), unbiased=False)
        n1 = c.numel() // N_FILT
        new_stats["bn1"] = (m1, v1 * n1 / max(n1 - 1, 1))
    else:
        m1, v1 = P["bn1_rm"], P["bn1_rv"]
    c = (c - m1[None, :, None, None]) / torch.sqrt(v1[None, :, None, None] + EPS)
    c = c * P["bn1_w"][None, :, None, None] + P["bn1_b"][None, :, None, None]   # :104
    c = F.elu(c)                                            # :105
    r["a1"] = c
    c = F.conv2d(c, P["Ws"], P["bs"])                       # :106  (B,40,1,36)
    r["y2"] = c
    if train:
        m2 = c.mean(dim=(0, 2, 3))
        v2 = c.var(dim=(0, 2, 3), unbiased=False)
        n2 = c.numel() // N_FILT
        new_stats["bn2"] = (m2, v2 * n2 / max(n2 - 1, 1))
    else:
        m2, v2 = P["bn2_rm"], P["bn2_rv"]
    c = (c - m2[None, :, None, None]) / torch.sqrt(v2[None, :, None, None] + EPS)
    c = c * P["bn2_w"][None, :, None, None] + P["bn2_b"][None, :, None, None]   # :107
    c = F.elu(c)                                            # :108
    if masks is not None and masks.get("conv") is not None:
        c = _drop(c.squeeze(2), masks, "conv").unsqueeze(2)  # :109
    c = F.conv2d(c, P["Wc"], P["bc"])                       # :112  1x1
    feat = c.squeeze(2).permute(0, 2, 1).reshape(B, D_FEAT)  # :114 'b e h w -> b (h w) e', :145 flatten
    r["feat"] = feat

    # ---- Proj_eeg (ATMS_retrieval.py:157-167) ----
    z1 = F.linear(feat, P["Wp1"], P["bp1"])
    r["z1"] = z1
    z2 = z1 + _drop(F.linear(F.gelu(z1), P["Wp2"], P["bp2"]), masks, "proj")   # ResidualAdd :128-137
    out = F.layer_norm(z2, (D_OUT,), P["gp"], P["bep"], EPS)
    r["out"] = out

    if train:
        r["bn_batch_stats"] = new_stats
        if update_running_stats:
            for name, (m, vu) in new_stats.items():
                rm, rv = sd[K[name + "_rm"]], sd[K[name + "_rv"]]
                rm.mul_(1 - BN_MOMENTUM).add_(m.to(rm.dtype) * BN_MOMENTUM)
                rv.mul_(1 - BN_MOMENTUM).add_(vu.to(rv.dtype) * BN_MOMENTUM)
    return r


def clip_loss(eeg: torch.Tensor, tgt: torch.Tensor, logit_scale: torch.Tensor) -> torch.Tensor:
    """ClipLoss.forward at world_size == 1 (models/loss.py:121-141).  logit_scale is used RAW."""
    logits_per_image = logit_scale * eeg @ tgt.T            # :122
    logits_per_text = logit_scale * tgt @ eeg.T             # :123
    labels = torch.arange(eeg.shape[0], device=eeg.device)  # :128
    return (F.cross_entropy(logits_per_image, labels) + F.cross_entropy(logits_per_text, labels)) / 2


def clip_loss_global(eeg_chunks, tgt_chunks, logit_scale):
    """world_size > 1, local_loss=False (models/loss.py:103-120): every rank builds the full N x N
    problem from the gathered chunks.  The loss VALUE is the single-process loss at batch N."""
    return clip_loss(torch.cat(list(eeg_chunks), 0), torch.cat(list(tgt_chunks), 0), logit_scale)


def contrastive_loss(eeg, img, txt, logit_scale, alpha: float = 0.99):
    """train_model loss mix (Retrieval/ATMS_retrieval.py:229-234)."""
    return alpha * clip_loss(eeg, img, logit_scale) + (1 - alpha) * clip_loss(eeg, txt, logit_scale)


def reconstruction_loss(eeg, img, logit_scale, alpha: float = 0.90):
    """loss of the reconstruction-training variant (Generation/ATMS_reconstruction.py:198-201, 224-228; evaluate_model
    uses the same expression with alpha = 0.99, :259, :283-286): alpha*10*MSE(eeg, img) + (1-alpha)*10*ClipLoss(eeg, img).
    The text ClipLoss is computed by the reference but not used."""
    return alpha * F.mse_loss(eeg, img) * 10 + (1 - alpha) * clip_loss(eeg, img, logit_scale) * 10


def train_accuracy_counts(eeg, gallery, labels, logit_scale):
    """train_model accuracy bookkeeping (ATMS_retrieval.py:241-250): argmax over the gallery."""
    logits = logit_scale * eeg @ gallery.T
    return int((torch.argmax(logits, dim=1) == labels).sum().item())


def retrieval_topk(eeg, gallery_sel, logit_scale, k_top: int = 5):
    """evaluate_model inner scoring (ATMS_retrieval.py:306-320) for a batch of queries with
    per-query candidate sets ``gallery_sel`` (Q, k, D).  Returns (top1 index, top-5 indices) into
    the candidate list, torch.argmax / torch.topk semantics."""
    logits = logit_scale * torch.einsum("qd,qkd->qk", eeg, gallery_sel)
    top1 = torch.argmax(logits, dim=1)
    kk = min(k_top, logits.shape[1])
    top5 = torch.topk(logits, kk, dim=1, largest=True).indices
    return top1, top5


def adamw_step(p, g, m, v, step: int, lr=3e-4, b1=0.9, b2=0.999, eps=1e-8, wd=1e-2):
    """torch.optim.AdamW single-tensor update (torch/optim/adamw.py -> adam.py `_single_tensor_adam`
    with decoupled weight decay), constructed at ATMS_retrieval.py:548 with defaults.  In place."""
    p.mul_(1 - lr * wd)
    m.lerp_(g, 1 - b1)
    v.mul_(b2).addcmul_(g, g, value=1 - b2)
    bc1 = 1 - b1 ** step
    bc2 = 1 - b2 ** step
    denom = (v.sqrt() / math.sqrt(bc2)).add_(eps)
    p.addcdiv_(m, denom, value=-(lr / bc1))


TRAINED_KEYS = [k for a, k in K.items() if a not in ("pe", "bn1_rm", "bn1_rv", "bn2_rm", "bn2_rv")]


def train_step(sd, opt_state, x, subject_ids, img, txt, step: int, masks=None, lr=3e-4,
               dtype=torch.float32, alpha=0.99, variant: str = "retrieval"):
    """One body of the hot loop (ATMS_retrieval.py:215-237): forward, loss mix, backward, AdamW.

    ``sd`` (reference-keyed tensors) and ``opt_state`` ({key: (m, v)}) are updated in place.
    Parameters whose grad is None (unused this step) are skipped like torch.optim does.
    Returns (loss, grads dict, forward intermediates)."""
    leaves = {}
    sd_g = dict(sd)
    for k in TRAINED_KEYS + joint_value_keys(sd):
        if k in sd:
            t = sd[k].detach().to(dtype).clone().requires_grad_(True)
            leaves[k] = t
            sd_g[k] = t
    r = atms_forward(sd_g, x, subject_ids, train=True, masks=masks, dtype=dtype)
    if variant == "reconstruction":     # Generation/ATMS_reconstruction.py:224-228 (alpha = 0.90 there)
        loss = reconstruction_loss(r["out"], img.to(dtype), sd_g["logit_scale"], alpha)
    else:
        loss = contrastive_loss(r["out"], img.to(dtype), txt.to(dtype), sd_g["logit_scale"], alpha)
    grads_list = torch.autograd.grad(loss, list(leaves.values()), allow_unused=True)
    grads = {k: g for k, g in zip(leaves.keys(), grads_list)}
    # BN running statistics (momentum 0.1, unbiased variance)
    for name, (mm, vu) in r["bn_batch_stats"].items():
        rm, rv = sd[K[name + "_rm"]], sd[K[name + "_rv"]]
        rm.mul_(1 - BN_MOMENTUM).add_(mm.detach().to(rm.dtype) * BN_MOMENTUM)
        rv.mul_(1 - BN_MOMENTUM).add_(vu.detach().to(rv.dtype) * BN_MOMENTUM)
    with torch.no_grad():
        for k, g in grads.items():
            if g is None:
                continue
            if k not in opt_state:
                opt_state[k] = (torch.zeros_like(sd[k]), torch.zeros_like(sd[k]))
            mm, vv = opt_state[k]
            adamw_step(sd[k], g.to(sd[k].dtype), mm, vv, step, lr=lr)
    return loss.detach(), grads, r
