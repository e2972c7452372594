"""Deterministic synthetic weights / inputs shared by the golden generator and the tests.

Weights are NOT taken from the reference's constructor RNG stream (that cannot be reproduced on
the GPU box, where /root/reference does not exist); they come from per-key seeded generators so
both sides can rebuild them bit-identically with nothing but torch.
"""
import hashlib
import math

import torch

# reference state_dict keys -> shapes (probed from the reference)
STATE_SHAPES = {
    "logit_scale": (),
    "encoder.enc_embedding.mask_token": (1, 250),
    "encoder.enc_embedding.value_embedding.weight": (250, 250),
    "encoder.enc_embedding.value_embedding.bias": (250,),
    "encoder.enc_embedding.position_embedding.pe": (1, 5000, 250),
    "encoder.enc_embedding.temporal_embedding.embed.weight": (250, 4),
    "encoder.enc_embedding.subject_embedding.shared_embedding": (1, 250),
    "encoder.enc_embedding.subject_embedding.mask_embedding": (1, 250),
    "encoder.enc_embedding.subject_embedding.subject_embedding.weight": (10, 250),
    "encoder.encoder.attn_layers.0.attention.query_projection.weight": (248, 250),
    "encoder.encoder.attn_layers.0.attention.query_projection.bias": (248,),
    "encoder.encoder.attn_layers.0.attention.key_projection.weight": (248, 250),
    "encoder.encoder.attn_layers.0.attention.key_projection.bias": (248,),
    "encoder.encoder.attn_layers.0.attention.value_projection.weight": (248, 250),
    "encoder.encoder.attn_layers.0.attention.value_projection.bias": (248,),
    "encoder.encoder.attn_layers.0.attention.out_projection.weight": (250, 248),
    "encoder.encoder.attn_layers.0.attention.out_projection.bias": (250,),
    "encoder.encoder.attn_layers.0.conv1.weight": (256, 250, 1),
    "encoder.encoder.attn_layers.0.conv1.bias": (256,),
    "encoder.encoder.attn_layers.0.conv2.weight": (250, 256, 1),
    "encoder.encoder.attn_layers.0.conv2.bias": (250,),
    "encoder.encoder.attn_layers.0.norm1.weight": (250,),
    "encoder.encoder.attn_layers.0.norm1.bias": (250,),
    "encoder.encoder.attn_layers.0.norm2.weight": (250,),
    "encoder.encoder.attn_layers.0.norm2.bias": (250,),
    "encoder.encoder.norm.weight": (250,),
    "encoder.encoder.norm.bias": (250,),
    "subject_wise_linear.0.weight": (250, 250),
    "subject_wise_linear.0.bias": (250,),
    "subject_wise_linear.1.weight": (250, 250),
    "subject_wise_linear.1.bias": (250,),
    "enc_eeg.0.tsconv.0.weight": (40, 1, 1, 25),
    "enc_eeg.0.tsconv.0.bias": (40,),
    "enc_eeg.0.tsconv.2.weight": (40,),
    "enc_eeg.0.tsconv.2.bias": (40,),
    "enc_eeg.0.tsconv.2.running_mean": (40,),
    "enc_eeg.0.tsconv.2.running_var": (40,),
    "enc_eeg.0.tsconv.2.num_batches_tracked": (),
    "enc_eeg.0.tsconv.4.weight": (40, 40, 63, 1),
    "enc_eeg.0.tsconv.4.bias": (40,),
    "enc_eeg.0.tsconv.5.weight": (40,),
    "enc_eeg.0.tsconv.5.bias": (40,),
    "enc_eeg.0.tsconv.5.running_mean": (40,),
    "enc_eeg.0.tsconv.5.running_var": (40,),
    "enc_eeg.0.tsconv.5.num_batches_tracked": (),
    "enc_eeg.0.projection.0.weight": (40, 40, 1, 1),
    "enc_eeg.0.projection.0.bias": (40,),
    "proj_eeg.0.weight": (1024, 1440),
    "proj_eeg.0.bias": (1024,),
    "proj_eeg.1.fn.1.weight": (1024, 1024),
    "proj_eeg.1.fn.1.bias": (1024,),
    "proj_eeg.2.weight": (1024,),
    "proj_eeg.2.bias": (1024,),
}


def _gen(tag: str, seed: int) -> torch.Generator:
    h = int.from_bytes(hashlib.sha256(f"{seed}:{tag}".encode()).digest()[:7], "little")
    return torch.Generator().manual_seed(h)


def positional_table() -> torch.Tensor:
    pe = torch.zeros(5000, 250)
    position = torch.arange(0, 5000).float().unsqueeze(1)
    div_term = (torch.arange(0, 250, 2).float() * -(math.log(10000.0) / 250)).exp()
    pe[:, 0::2] = torch.sin(position * div_term)
    pe[:, 1::2] = torch.cos(position * div_term)
    return pe.unsqueeze(0)


def joint_state_shapes(num_subjects: int = 10) -> dict:
    """state_dict of ATMS(joint_train=True) in Retrieval/ATMS_retrieval_joint_train.py:173-176 (probed): the value
    embedding is a ModuleDict of per-subject Linear(250,250) and there are num_subjects subject_wise_linear layers"""
    shapes = {}
    for key, shape in STATE_SHAPES.items():
        if key.startswith("encoder.enc_embedding.value_embedding."):
            if key.endswith("weight"):
                for sj in range(num_subjects):
                    shapes[f"encoder.enc_embedding.value_embedding.{sj}.weight"] = (250, 250)
                    shapes[f"encoder.enc_embedding.value_embedding.{sj}.bias"] = (250,)
            continue
        if key.startswith("subject_wise_linear."):
            if key == "subject_wise_linear.0.weight":
                for sj in range(num_subjects):
                    shapes[f"subject_wise_linear.{sj}.weight"] = (250, 250)
                    shapes[f"subject_wise_linear.{sj}.bias"] = (250,)
            continue
        shapes[key] = shape
    return shapes


def make_joint_state_dict(seed: int = 0, num_subjects: int = 10) -> dict:
    return make_state_dict(seed, joint_state_shapes(num_subjects))


def make_state_dict(seed: int = 0, shapes: dict = None) -> dict:
    """Reference-keyed fp32 state_dict with non-trivial affine/BN values."""
    sd = {}
    for key, shape in (shapes or STATE_SHAPES).items():
        g = _gen(key, seed)
        if key.endswith("num_batches_tracked"):
            sd[key] = torch.tensor(0, dtype=torch.long)
        elif key == "logit_scale":
            sd[key] = torch.tensor(math.log(1 / 0.07), dtype=torch.float32)
        elif key.endswith("position_embedding.pe"):
            sd[key] = positional_table()
        elif key.endswith("running_mean"):
            sd[key] = 0.1 * torch.randn(shape, generator=g)
        elif key.endswith("running_var"):
            sd[key] = 0.6 + 0.8 * torch.rand(shape, generator=g)
        elif ("norm" in key or "tsconv.2" in key or "tsconv.5" in key or key.startswith("proj_eeg.2")) \
                and key.endswith("weight") and len(shape) == 1:
            sd[key] = 1.0 + 0.1 * torch.randn(shape, generator=g)
        elif key.endswith("bias"):
            sd[key] = 0.05 * torch.randn(shape, generator=g)
        elif "embedding" in key and ("shared" in key or "mask" in key or key.endswith("subject_embedding.weight")):
            sd[key] = torch.randn(shape, generator=g)
        elif key.endswith("mask_token"):
            sd[key] = torch.randn(shape, generator=g)
        else:  # Linear / Conv weights: U(-1/sqrt(fan_in), 1/sqrt(fan_in)) like torch's default
            fan_in = 1
            for s in shape[1:]:
                fan_in *= s
            bound = 1.0 / math.sqrt(max(fan_in, 1))
            sd[key] = (torch.rand(shape, generator=g) * 2 - 1) * bound
    return sd


def make_eeg(batch: int, seed: int = 1234) -> torch.Tensor:
    return torch.randn(batch, 63, 250, generator=_gen("eeg", seed))


def make_targets(n: int, seed: int = 1234, tag: str = "img") -> torch.Tensor:
    t = torch.randn(n, 1024, generator=_gen(tag, seed))
    return torch.nn.functional.normalize(t, dim=-1)


def make_labels(n: int, n_cls: int, seed: int = 1234) -> torch.Tensor:
    return torch.randint(0, n_cls, (n,), generator=_gen("labels", seed))


def digest(t: torch.Tensor, n: int = 96) -> torch.Tensor:
    """Small fingerprint of a tensor: [l2 norm, sum, abs-max, n strided samples]."""
    f = t.detach().double().flatten()
    if f.numel() == 0:
        return torch.zeros(3 + n, dtype=torch.float64)
    idx = torch.linspace(0, f.numel() - 1, n).long()
    return torch.cat([torch.stack([f.norm(), f.sum(), f.abs().max()]), f[idx]])
