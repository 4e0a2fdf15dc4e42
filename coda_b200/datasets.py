"""Datasets: the reference loader contract (coda/datasets.py:4-23) plus shard-aware variants."""
from __future__ import annotations

import os

import torch

from .synth import shard_range, synth


class Dataset:
    """(H, N, C) post-softmax scores from ``filepath`` (+ optional ``*_labels.pt``), forced to fp32
    (coda/datasets.py:12-23)."""

    def __init__(self, filepath, device):
        self.device = device
        self.preds = torch.load(filepath, map_location=device).float().contiguous()
        print("Loaded preds of shape", self.preds.shape)
        self.labels = None
        label_p = filepath.replace(".pt", "_labels.pt")
        if os.path.exists(label_p):
            self.labels = torch.load(label_p, map_location=device)
            print("Loaded labels of shape", self.labels.shape)
        else:
            print("Did not load labels.")


class TensorDataset:
    """Wrap tensors already in memory.  ``n_offset``/``n_global`` describe an N-axis shard."""

    def __init__(self, preds, labels=None, n_offset=0, n_global=None):
        self.preds, self.labels, self.device = preds, labels, preds.device
        self.n_offset = n_offset
        self.n_global = preds.shape[1] if n_global is None else n_global


class SyntheticDataset(TensorDataset):
    """This rank's shard of the synthetic task; labels are replicated (N int64)."""

    def __init__(self, H, N, C, seed=0, device="cuda", dense=False, rank=0, world=1, generator_device=None):
        lo, hi = shard_range(N, rank, world)
        gdev = generator_device or device
        preds, _ = synth(H, N, C, seed, device=gdev, dense=dense, n_lo=lo, n_hi=hi)
        _, labels = synth(H, N, C, seed, device=gdev, dense=dense, want_preds=False)
        super().__init__(preds.to(device), labels, n_offset=lo, n_global=N)
        self.labels_host = labels.cpu()
