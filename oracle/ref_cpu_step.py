"""CPU training step of the reference for bench.py's `--impl reference` arm and `cpu_baseline` leg.
TEST/BENCH INFRASTRUCTURE ONLY -- nothing in the product package imports this.

Two flavours, chosen at run time:
  kind "reference": the UNMODIFIED reference modules `rnn_model.py` (GRUDecoder) and `data_augmentations.py`
      (gauss_smooth), staged by `__graft_entry__.build()` from /root/reference/model_training into
      oracle/_ref/model_training/ (git-ignored, travels to the GPU box), driven by the statement sequence of the
      reference's training loop (rnn_trainer.py:511-558: noise augmentation, smoothing, model, log_softmax, CTCLoss
      mean, backward, clip_grad_norm_, AdamW with the three parameter groups of rnn_trainer.py:259-292).  The trainer
      class itself needs omegaconf / h5py / the Dryad data and cannot run offline, so its loop body is what is timed.
  kind "port": oracle/torch_cpu_port.py (same library calls, restated) when the staged files are absent.
fp32 on CPU: `torch.autocast(device_type="cuda")` is a no-op for CPU tensors, which is exactly what the reference does
when it falls back to CPU (rnn_trainer.py:98-107).
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref", "model_training")


def have_reference() -> bool:
    return os.path.exists(os.path.join(REF_DIR, "rnn_model.py")) and os.path.exists(os.path.join(REF_DIR, "data_augmentations.py"))


class RefStep:
    """model + optimizer + one-step callable on CPU tensors."""

    def __init__(self, cfg, *, lr=5e-3 * 0.5, white_std=1.0, offset_std=0.2, clip=10.0, threads=None):
        self.cores = threads or os.cpu_count() or 1
        torch.set_num_threads(self.cores)
        self.white_std, self.offset_std, self.clip = white_std, offset_std, clip
        self.ps, self.st = cfg["patch_size"], cfg["patch_stride"]
        torch.manual_seed(0)
        if have_reference():
            if REF_DIR not in sys.path:
                sys.path.insert(0, REF_DIR)
            from rnn_model import GRUDecoder                      # the reference, unmodified
            from data_augmentations import gauss_smooth           # the reference, unmodified
            self.kind = "reference"
            self.model = GRUDecoder(neural_dim=cfg["neural_dim"], n_units=cfg["n_units"], n_days=cfg["n_days"], n_classes=cfg["n_classes"],
                                    rnn_dropout=cfg["rnn_dropout"], input_dropout=cfg["input_dropout"], n_layers=cfg["n_layers"],
                                    patch_size=cfg["patch_size"], patch_stride=cfg["patch_stride"])
            self._smooth = lambda f: gauss_smooth(inputs=f, device="cpu", smooth_kernel_std=2, smooth_kernel_size=100)
        else:
            if HERE not in sys.path:
                sys.path.insert(0, HERE)
            from torch_cpu_port import PortModel, smooth_same
            self.kind = "port"
            self.model = PortModel(**cfg)
            self._smooth = smooth_same
        named = list(self.model.named_parameters())
        bias = [p for n, p in named if "gru.bias" in n or "out.bias" in n]
        day = [p for n, p in named if "day_" in n]
        other = [p for n, p in named if "day_" not in n and "gru.bias" not in n and "out.bias" not in n]
        # rnn_trainer.py:270-290 (fused=True is a CUDA-only flag: the CPU fallback of the reference cannot use it either)
        self.opt = torch.optim.AdamW([{"params": bias, "weight_decay": 0}, {"params": day, "lr": lr, "weight_decay": 0}, {"params": other}],
                                     lr=lr, betas=(0.9, 0.999), eps=0.1, weight_decay=1e-3)
        self.ctc = torch.nn.CTCLoss(blank=0, reduction="none", zero_infinity=False)

    def step(self, x, n_steps, labels, lens, days, cut=0):
        self.model.train()
        self.opt.zero_grad()
        f = x
        if self.white_std > 0:
            f = f + torch.randn(f.shape) * self.white_std
        if self.offset_std > 0:
            f = f + torch.randn((f.shape[0], 1, f.shape[2])) * self.offset_std
        if cut > 0:
            f = f[:, cut:, :]
            n_steps = n_steps - cut
        f = self._smooth(f)
        adj = ((n_steps - self.ps) / self.st + 1).to(torch.int32)
        logits = self.model(f, days)
        loss = self.ctc(torch.permute(logits.log_softmax(2), [1, 0, 2]), labels, adj, lens)
        loss = torch.mean(loss)
        loss.backward()
        if self.clip > 0:
            torch.nn.utils.clip_grad_norm_(self.model.parameters(), max_norm=self.clip, error_if_nonfinite=True, foreach=True)
        self.opt.step()
        return float(loss.detach())
