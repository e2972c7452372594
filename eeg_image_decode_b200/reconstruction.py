"""Reconstruction-training variant and embedding export.

Drop-in for ``train_model`` / ``evaluate_model`` of Generation/ATMS_reconstruction.py (:193-249, :251-317) -- the ATM-S
encoder trained with ``alpha*10*MSE(eeg, img) + (1-alpha)*10*ClipLoss(eeg, img)`` (alpha 0.90 in training, :198, :224-228;
0.99 in evaluation, :259, :283-286) so that the embedding regresses onto the un-normalised CLIP image embedding the
frozen diffusion prior / SDXL + IP-Adapter stage consumes -- and for the notebooks' ``get_eegfeatures``
(Generation/Generation_metrics_sub8.ipynb, cell defining it) which dumps the eval-mode embeddings to
``ATM_S_eeg_features_{sub}.pt``.  The model class is the same ``ATMS`` (ATMS_reconstruction.py:162-183 is identical to
the retrieval one).  The MSE term runs as ``eegb200_mse`` right after the InfoNCE kernels (include/eegdecode_b200.h);
everything else is the shared step engine (CUDA-graph captured, data-parallel capable).
"""
from __future__ import annotations

import os
import random

import torch

from . import _lib
from .atms import ATMS, N_SUBJECT_ROWS  # noqa: F401
from .train import StepEngine, _drain, _encode_batches, _eval_losses, _evaluate_epoch, _train_epoch, extract_id_from_string


def train_model(sub, eeg_model, dataloader, optimizer, device, text_features_all, img_features_all, config, *,
                step_callback=None):
    """One epoch of ATMS_reconstruction.py:193-249.  Returns (average_loss, accuracy, features[n_seen,1024])."""
    return _train_epoch(sub, eeg_model, dataloader, optimizer, device, text_features_all, img_features_all, config,
                        variant="reconstruction", alpha=0.90, step_callback=step_callback)


def _export(sub, eeg_model, dataloader, device, text_features_all, img_features_all, k, alpha, keep_features):
    eeg_model.eval()
    device = torch.device(device)
    if device.type != "cuda":
        raise RuntimeError("this implementation runs on CUDA only (no CPU fallback)")
    text_features_all = text_features_all.to(device).float()
    img_features_all = img_features_all.to(device).float().contiguous()
    all_labels = set(range(text_features_all.size(0)))
    eng = StepEngine(eeg_model, None, alpha, "reconstruction")
    eng.world, eng.rank = 1, 0          # export / evaluation is per process, no collectives
    subject_id = extract_id_from_string(sub)
    eeg_list, label_batches, _txt, img_list = _drain(dataloader)
    if not eeg_list:
        raise RuntimeError("get_eegfeatures: empty dataloader")
    sizes = [int(e.size(0)) for e in eeg_list]
    # host side: the notebook's per-trial candidate draw (one draw, top-1 only, any k), in loader order
    label_list, sel_rows = [], []
    for lab in label_batches:
        for label in lab.tolist():
            possible_classes = list(all_labels - {label})
            sel_rows.append(random.sample(possible_classes, k - 1) + [label])
            label_list.append(label)
    with torch.no_grad():
        # one batched eval-mode forward per 1024 trials (66 160 trials per subject in the notebooks) instead of one per batch
        feats = _encode_batches(eeg_model, eeg_list, device, subject_id, None)
        total_loss = _eval_losses(eng, feats, sizes, img_list, [None] * len(sizes), device)
        correct = 0
        for i0 in range(0, len(label_list), 4096):          # scoring in slabs: the logits buffer is [Q, n_classes]
            sel = torch.tensor(sel_rows[i0:i0 + 4096], dtype=torch.int32)
            r = _lib.retrieval(feats[i0:i0 + 4096], img_features_all, eeg_model.logit_scale.detach(), sel=sel, want_top5=False)
            for j, t1 in enumerate(r["top1"].tolist()):
                correct += int(sel_rows[i0 + j][t1] == label_list[i0 + j])
    average_loss = float(total_loss[0].item()) / len(sizes)
    return average_loss, correct, len(label_list), label_batches[-1], (feats if keep_features else None)


def evaluate_model(sub, eeg_model, dataloader, device, text_features_all, img_features_all, k, config):
    """k-way zero-shot retrieval with the reconstruction loss mix at alpha = 0.99 (ATMS_reconstruction.py:251-353).
    Returns (average_loss, accuracy, top5_acc); candidate draws as in the reference (same body as the retrieval script)."""
    return _evaluate_epoch(sub, eeg_model, dataloader, device, text_features_all, img_features_all, k, config,
                           variant="reconstruction", alpha=0.99)


def get_eegfeatures(sub, eegmodel, dataloader, device, text_features_all, img_features_all, k, *, save_features=True,
                    out_dir="."):
    """Embedding export of the generation notebooks (``get_eegfeatures(sub, eegmodel, dataloader, device,
    text_features_all, img_features_all, k)``): eval-mode embeddings of every trial, k-way accuracy, loss mix with
    alpha = 0.9.  Returns (average_loss, accuracy, labels of the last batch, features.cpu()) and writes
    ``ATM_S_eeg_features_{sub}.pt`` (the file the diffusion prior / SDXL stage loads) unless ``save_features=False``."""
    average_loss, correct, total, labels, feats = _export(sub, eegmodel, dataloader, device, text_features_all,
                                                            img_features_all, k, 0.9, True)
    features_tensor = feats.cpu()
    if save_features:
        print("features_tensor", features_tensor.shape)
        torch.save(features_tensor, os.path.join(out_dir, f"ATM_S_eeg_features_{sub}.pt"))
    return average_loss, correct / total, labels, features_tensor
