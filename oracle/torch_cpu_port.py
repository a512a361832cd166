"""CPU baseline for bench.py: the reference's training step restated on PyTorch-CPU.  TEST/BENCH
INFRASTRUCTURE ONLY (imported by bench.py's cpu_baseline / --impl reference legs and by tests).

The reference's arithmetic for this path lives in third-party PyTorch (torch.nn.GRU -> MKL/oneDNN on
CPU, torch.nn.CTCLoss, torch.optim.AdamW); the reference file itself cannot travel to the GPU box, so
this module restates its call sequence with the same library calls:
  model  : rnn_model.py:88-134   (day einsum + Softsign + Dropout, unfold patches, nn.GRU, Linear)
  step   : rnn_trainer.py:527-558 (noise, gauss smoothing, log_softmax, CTCLoss mean, backward,
                                   clip_grad_norm_, AdamW with the three param groups)
fp32 on CPU: torch.autocast(device_type='cuda') is a no-op for CPU tensors, exactly what happens when
the reference falls back to CPU (rnn_trainer.py:98-107).  Checked against the imported reference by
tests/test_oracle_golden.py via the golden vectors (same logits / loss).
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F

from gru_ctc_oracle import gauss_taps


class PortModel(torch.nn.Module):
    def __init__(self, neural_dim=512, n_units=768, n_days=45, n_classes=41, n_layers=5, patch_size=14, patch_stride=4,
                 rnn_dropout=0.4, input_dropout=0.2):
        super().__init__()
        self.patch_size, self.patch_stride, self.n_layers, self.n_units = patch_size, patch_stride, n_layers, n_units
        self.day_weights = torch.nn.ParameterList([torch.nn.Parameter(torch.eye(neural_dim)) for _ in range(n_days)])
        self.day_biases = torch.nn.ParameterList([torch.nn.Parameter(torch.zeros(1, neural_dim)) for _ in range(n_days)])
        self.drop = torch.nn.Dropout(input_dropout)
        self.gru = torch.nn.GRU(neural_dim * max(patch_size, 1), n_units, n_layers, dropout=rnn_dropout, batch_first=True)
        for name, p in self.gru.named_parameters():
            if "weight_hh" in name:
                torch.nn.init.orthogonal_(p)
            if "weight_ih" in name:
                torch.nn.init.xavier_uniform_(p)
        self.out = torch.nn.Linear(n_units, n_classes)
        torch.nn.init.xavier_uniform_(self.out.weight)
        self.h0 = torch.nn.Parameter(torch.nn.init.xavier_uniform_(torch.zeros(1, 1, n_units)))

    def forward(self, x, day_idx):
        w = torch.stack([self.day_weights[int(i)] for i in day_idx], 0)
        b = torch.cat([self.day_biases[int(i)] for i in day_idx], 0).unsqueeze(1)
        x = F.softsign(torch.einsum("btd,bdk->btk", x, w) + b)
        x = self.drop(x)
        if self.patch_size > 0:
            u = x.permute(0, 2, 1).unfold(2, self.patch_size, self.patch_stride)   # [B, D, T', P]
            x = u.permute(0, 2, 3, 1).reshape(x.size(0), u.size(2), -1)
        h = self.h0.expand(self.n_layers, x.shape[0], self.n_units).contiguous()
        y, _ = self.gru(x, h)
        return self.out(y)

    def load_numpy(self, params):
        sd = {k: torch.from_numpy(np.asarray(v, dtype=np.float32)).reshape(self.state_dict()[k].shape) for k, v in params.items()}
        self.load_state_dict(sd)


def smooth_same(x, std=2.0, size=100):
    k = torch.from_numpy(gauss_taps(std, size))
    C = x.shape[2]
    return F.conv1d(x.permute(0, 2, 1), k.view(1, 1, -1).repeat(C, 1, 1), padding="same", groups=C).permute(0, 2, 1)


def make_optimizer(model, lr=5e-3, lr_day=5e-3, wd=1e-3, wd_day=0.0, betas=(0.9, 0.999), eps=0.1):
    bias = [p for n, p in model.named_parameters() if "gru.bias" in n or "out.bias" in n]
    day = [p for n, p in model.named_parameters() if "day_" in n]
    other = [p for n, p in model.named_parameters() if "day_" not in n and "gru.bias" not in n and "out.bias" not in n]
    return torch.optim.AdamW([{"params": bias, "weight_decay": 0}, {"params": day, "lr": lr_day, "weight_decay": wd_day},
                              {"params": other}], lr=lr, betas=betas, eps=eps, weight_decay=wd)


def train_step(model, opt, x, n_steps, labels, lens, days, *, white_std=1.0, offset_std=0.2, cut=0, clip=10.0):
    """rnn_trainer.py:511-558 on CPU tensors.  Returns the scalar loss."""
    model.train()
    opt.zero_grad()
    f = x
    if white_std > 0:
        f = f + torch.randn(f.shape) * white_std
    if offset_std > 0:
        f = f + torch.randn(f.shape[0], 1, f.shape[2]) * offset_std
    if cut > 0:
        f = f[:, cut:, :]
        n_steps = n_steps - cut
    f = smooth_same(f)
    adj = ((n_steps - model.patch_size) / model.patch_stride + 1).to(torch.int32)
    logits = model(f, days)
    loss = F.ctc_loss(logits.log_softmax(2).permute(1, 0, 2), labels, adj, lens, blank=0, reduction="none", zero_infinity=False)
    loss = loss.mean()
    loss.backward()
    if clip > 0:
        torch.nn.utils.clip_grad_norm_(model.parameters(), max_norm=clip, error_if_nonfinite=True, foreach=True)
    opt.step()
    return float(loss.detach())
