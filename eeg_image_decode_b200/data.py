"""HBM-resident data path: replaces ``DataLoader(EEGDataset(...))`` for the hot loop.

One THINGS-EEG2 subject is 66160 x 63 x 250 fp32 = 4.17 GB (all ten: 41.7 GB) -- it fits in a B200's 180 GB next to the
(16540, 1024) image and (1654, 1024) text CLIP features, so the per-step synchronous, unpinned host->device copy of the
reference (Retrieval/ATMS_retrieval.py:210-213) disappears: shuffling and the index -> (text, image) feature gathers
run on the device.  The index arithmetic restates ``EEGDataset.__getitem__`` (Retrieval/eegdatasets_leaveone.py:326-375,
``pictures is None`` branch) and the batches are the same 6-tuples, so ``train_model`` / ``evaluate_model`` take this
loader unchanged.  Loading the pickled EEG and computing CLIP features stay the reference's business.
"""
from __future__ import annotations

from typing import Iterator, List, Optional, Sequence

import torch


class ResidentEEGData:
    def __init__(self, eeg: torch.Tensor, labels: torch.Tensor, text_features: torch.Tensor, img_features: torch.Tensor,
                 train: bool = True, n_cls: Optional[int] = None, text: Optional[Sequence[str]] = None,
                 img: Optional[Sequence[str]] = None, device="cuda"):
        dev = torch.device(device)
        if eeg.dim() != 3 or eeg.shape[1] != 63 or eeg.shape[2] != 250:
            raise ValueError(f"expected EEG of shape [N,63,250], got {tuple(eeg.shape)}")
        self.data = eeg.to(dev, dtype=torch.float32).contiguous()
        self.labels = labels.to(dev, dtype=torch.long)
        self.text_features = text_features.to(dev).float().contiguous()
        self.img_features = img_features.to(dev).float().contiguous()
        self.train = bool(train)
        self.n_cls = int(n_cls) if n_cls is not None else (1654 if train else 200)
        self.text = list(text) if text is not None else None
        self.img = list(img) if img is not None else None
        self.device = dev

    def __len__(self) -> int:
        return self.data.shape[0]

    def feature_indices(self, index: torch.Tensor):
        """(text_index, img_index) of EEGDataset.__getitem__ (eegdatasets_leaveone.py:333-348), vectorised.
        train: 10 images x 4 repetitions per class; test: one (80-repetition average) trial per class."""
        if self.train:
            per_sub = self.n_cls * 10 * 4
            r = index % per_sub
            return r // (10 * 4), r // 4
        per_sub = self.n_cls * 1 * 80
        r = index % per_sub
        return r, r

    def batch(self, index: torch.Tensor):
        """the reference's 6-tuple (x, label, text, text_features, img, img_features) for a vector of sample indices"""
        index = index.to(self.device, dtype=torch.long)
        ti, ii = self.feature_indices(index)
        text = [self.text[i] for i in ti.tolist()] if self.text is not None else None
        img = [self.img[i] for i in ii.tolist()] if self.img is not None else None
        return (self.data.index_select(0, index), self.labels.index_select(0, index), text,
                self.text_features.index_select(0, ti), img, self.img_features.index_select(0, ii))

    def loader(self, batch_size: int, shuffle: bool = True, drop_last: bool = True,
               generator: Optional[torch.Generator] = None) -> "ResidentLoader":
        return ResidentLoader(self, batch_size, shuffle, drop_last, generator)


class ResidentLoader:
    """iterable with DataLoader semantics (fresh permutation per epoch, optional drop_last)"""

    def __init__(self, ds: ResidentEEGData, batch_size: int, shuffle: bool, drop_last: bool,
                 generator: Optional[torch.Generator]):
        self.ds, self.batch_size, self.shuffle, self.drop_last, self.generator = ds, int(batch_size), shuffle, drop_last, generator

    def __len__(self) -> int:
        n = len(self.ds)
        return n // self.batch_size if self.drop_last else (n + self.batch_size - 1) // self.batch_size

    def __iter__(self) -> Iterator:
        n = len(self.ds)
        if self.shuffle:
            g = self.generator
            perm = torch.randperm(n, generator=g, device=g.device if g is not None else "cpu").to(self.ds.device)
        else:
            perm = torch.arange(n, device=self.ds.device)
        end = n - n % self.batch_size if self.drop_last else n
        for i in range(0, end, self.batch_size):
            yield self.ds.batch(perm[i:i + self.batch_size])
