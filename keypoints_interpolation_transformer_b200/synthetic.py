"""Seeded synthetic batches in the dataloader's layout: smooth keypoint
trajectories, ~2 % exact zeros, missing blocks drawn with the reference's policy and AUTSL statistics
(dataset_config.json:5-9), hold-filled as dataloader.py:421-434, SOS frame prepended (:482-493).
Host-side generator for benchmarks / examples (there is no dataset or network in this environment)."""
import random as _pyrandom

import numpy as np
import torch

from . import missing


def synthetic_batch(B, T, K, seed=42, zero_frac=0.02, smooth=True, dataset="AUTSL"):
    """-> (inputs [B,T+1,K,2], gt [B,T,K,2], mask [B,T+1]) fp32 CPU tensors."""
    rs = np.random.RandomState(seed)
    pr = _pyrandom.Random(seed)
    t = np.arange(T, dtype=np.float64).reshape(1, T, 1, 1) / T
    if smooth:
        f = rs.uniform(0.5, 3.0, size=(B, 1, K, 2))
        ph = rs.uniform(0.0, 1.0, size=(B, 1, K, 2))
        gt = (0.5 + 0.3 * np.sin(2 * np.pi * (f * t + ph))).astype(np.float32)
    else:
        gt = rs.uniform(0.0, 1.0, size=(B, T, K, 2)).astype(np.float32)
    if zero_frac > 0:
        gt[rs.uniform(size=(B, T, K)) < zero_frac] = 0.0
    inputs = np.empty((B, T + 1, K, 2), dtype=np.float32)
    mask = np.zeros((B, T + 1), dtype=np.float32)
    inputs[:, 0] = 1.0
    for b in range(B):
        src, m = missing.draw_sources(T, False, dataset, rng=pr, nprng=rs, config=missing.DATASET_CONFIG)
        inputs[b, 1:] = gt[b][src]
        mask[b, 1:] = m
    return torch.from_numpy(inputs), torch.from_numpy(gt), torch.from_numpy(mask)
