"""Batch sources with the reference's batch-dict contract (model_training/dataset.py:100-159).

The hdf5 reader of the reference is host-side data loading and stays out of this package's scope
(dataset.py OUT OF SCOPE, "next" row N3); when the reference's own ``dataset``
module and ``h5py`` are importable the trainer uses them unchanged.  For benchmarks and tests this
module provides a synthetic source producing the same dictionary:

    input_features [B, T, 512] f32, seq_class_ids [B, S] int, n_time_steps [B], phone_seq_lens [B],
    day_indicies [B], transcriptions [B, 1], block_nums [B], trial_nums [B]
"""
from __future__ import annotations

import numpy as np
import torch
from torch.utils.data import Dataset


class SyntheticBrainToTextDataset(Dataset):
    """A learnable synthetic corpus: each phoneme class has a fixed random signature in channel space;
    a trial is a phoneme sequence rendered as bumps of its signatures plus noise, so that a model
    trained on it reaches a low phoneme error rate -- used to compare training runs of two
    implementations on equal terms (there is no real data offline)."""

    def __init__(self, *, n_batches, batch_size=64, days_per_batch=4, n_days=45, neural_dim=512, n_classes=41,
                 T=400, min_len=6, max_len=14, frames_per_phone=(14, 24), noise=0.6, seed=0, split="train",
                 ragged=True):
        self.n_batches, self.batch_size, self.days_per_batch, self.n_days = n_batches, batch_size, days_per_batch, n_days
        self.D, self.C, self.T = neural_dim, n_classes, T
        self.min_len, self.max_len, self.fpp, self.noise = min_len, max_len, frames_per_phone, noise
        self.seed, self.split, self.ragged = seed, split, ragged
        g = np.random.RandomState(1234)                          # signatures shared by train and val
        self.signatures = g.randn(n_classes, neural_dim).astype(np.float32) * 0.8
        self.day_gain = 1.0 + 0.15 * g.randn(n_days, neural_dim).astype(np.float32)
        self.day_offset = 0.2 * g.randn(n_days, neural_dim).astype(np.float32)

    def __len__(self):
        return self.n_batches

    def _trial(self, rng, day):
        S = rng.randint(self.min_len, self.max_len + 1)
        labels = rng.randint(1, self.C, size=S)
        x = np.zeros((self.T, self.D), dtype=np.float32)
        t = rng.randint(2, 8)
        for ph in labels:
            dur = rng.randint(self.fpp[0], self.fpp[1] + 1)
            if t + dur >= self.T - 2:
                break
            env = np.hanning(dur + 2)[1:-1].astype(np.float32)[:, None]
            x[t:t + dur] += env * self.signatures[ph][None, :]
            t += dur + rng.randint(0, 4)
        n_steps = min(self.T, t + rng.randint(4, 12)) if self.ragged else self.T
        x[:n_steps] += self.noise * rng.randn(n_steps, self.D).astype(np.float32)
        x[:n_steps] = x[:n_steps] * self.day_gain[day] + self.day_offset[day]
        x[n_steps:] = 0
        return x, labels, n_steps

    def __getitem__(self, idx):
        rng = np.random.RandomState((self.seed * 1000003 + idx * 7919 + (0 if self.split == "train" else 99991)) % (2 ** 31))
        if self.split == "train":
            days = rng.choice(self.n_days, size=self.days_per_batch, replace=False)
            per = self.batch_size // self.days_per_batch
            day_list = np.repeat(days, per)[:self.batch_size]
        else:
            day_list = np.full((self.batch_size,), idx % self.n_days)          # validation batches are single-day
        xs, labs, nst = [], [], []
        for d in day_list:
            x, l, n = self._trial(rng, int(d))
            xs.append(x); labs.append(l); nst.append(n)
        Tm = max(nst)
        Sm = max(len(l) for l in labs)
        feats = np.stack([x[:Tm] for x in xs])
        lab = np.zeros((len(labs), Sm), dtype=np.int64)
        for i, l in enumerate(labs):
            lab[i, :len(l)] = l
        return {
            "input_features": torch.from_numpy(feats),
            "seq_class_ids": torch.from_numpy(lab),
            "n_time_steps": torch.tensor(nst, dtype=torch.int64),
            "phone_seq_lens": torch.tensor([len(l) for l in labs], dtype=torch.int64),
            "day_indicies": torch.from_numpy(day_list.astype(np.int64)),
            "transcriptions": torch.zeros((len(labs), 1), dtype=torch.int64),
            "block_nums": torch.zeros(len(labs), dtype=torch.int64),
            "trial_nums": torch.arange(len(labs), dtype=torch.int64),
        }
