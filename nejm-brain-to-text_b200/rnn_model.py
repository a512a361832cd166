"""Drop-in for the reference's ``model_training/rnn_model.py``: same class name, constructor,
``forward`` signature, attribute names and ``state_dict`` keys -- backed by the sm_100a engine.

Reference: rnn_model.py:4-134 (GRUDecoder).  The trainer filters parameters by the substrings
'gru.bias', 'out.bias', 'day_', 'gru', 'day' (rnn_trainer.py:146-148,249-254,267-269), so the names
are part of the API: day_weights.{i}, day_biases.{i}, gru.weight_ih_l{k}, gru.weight_hh_l{k},
gru.bias_ih_l{k}, gru.bias_hh_l{k}, out.weight, out.bias, h0.

All parameters are views into ONE flat fp32 buffer (so the gradient all-reduce and the fused
clip+AdamW kernel see a single array).  There is no CPU execution path: calling the module on CPU
tensors raises.
"""
from __future__ import annotations

from typing import Optional

import torch
from torch import nn

from . import _native as N
from . import engine as E


class _ParamGroup(nn.Module):
    """Name holder so that parameters appear as ``gru.*`` / ``out.*`` in the state_dict."""

    def forward(self, *a, **k):  # pragma: no cover
        raise RuntimeError("parameter container; the computation lives in GRUDecoder.forward")


class _GRUDecoderFn(torch.autograd.Function):
    """Autograd bridge: forward/backward run entirely in the native engine."""

    @staticmethod
    def forward(ctx, model, x, day_idx, states, seed, *params):
        eng = model._engine_for(x.shape[0], x.shape[1], training=True)
        logits, hidden = eng.forward(x, day_idx, training=True, smooth_mode=0, seed=seed, states=states, want_hidden=True)
        ctx.model, ctx.eng = model, eng
        ctx.mark_non_differentiable(hidden)
        return logits, hidden

    @staticmethod
    def backward(ctx, dlogits, _dhidden):
        model, eng = ctx.model, ctx.eng
        eng.set_dlogits(dlogits)
        eng.backward()
        touched = eng.touched_days()
        flags = touched.cpu().tolist() if model.n_days > 0 else []
        grads = []
        for name, p in model._named_flat:
            if not p.requires_grad:
                grads.append(None)
                continue
            if name.startswith("day_") and flags[int(name.split(".")[1])] <= 0:
                grads.append(None)              # day layer not in this batch: grad stays None, AdamW skips it
                continue
            off, n = model._slots[name]
            # a copy: autograd may keep what is returned here as p.grad, and the engine's buffer is rewritten by the next backward
            grads.append(eng.grads[off:off + n].view(p.shape).clone())
        return (None, None, None, None, None) + tuple(grads)


class GRUDecoder(nn.Module):
    '''
    Defines the GRU decoder

    This class combines day-specific input layers, a GRU, and an output classification layer
    '''

    def __init__(self, neural_dim, n_units, n_days, n_classes, rnn_dropout=0.0, input_dropout=0.0, n_layers=5,
                 patch_size=0, patch_stride=0):
        super().__init__()
        self.neural_dim = neural_dim
        self.n_units = n_units
        self.n_classes = n_classes
        self.n_layers = n_layers
        self.n_days = n_days
        self.rnn_dropout = rnn_dropout
        self.input_dropout = input_dropout
        self.patch_size = patch_size
        self.patch_stride = patch_stride
        self.input_size = neural_dim * patch_size if patch_size > 0 else neural_dim

        self._cfg = E.make_config(neural_dim, n_units, n_layers, n_days, n_classes, patch_size, patch_stride, rnn_dropout,
                                  input_dropout)
        layout = E.param_layout(self._cfg)
        self._slots = {name: (off, rows * cols) for name, off, rows, cols in layout}
        flat = torch.zeros(E.param_elems(self._cfg), dtype=torch.float32)

        # --- initial values: the same init calls, in the same order, as rnn_model.py:50-86, so that a given
        #     torch seed yields the same weights as the reference module.
        init = {}
        for i in range(n_days):
            init[f"day_weights.{i}"] = torch.eye(neural_dim)
        for i in range(n_days):
            init[f"day_biases.{i}"] = torch.zeros(1, neural_dim)
        gru = nn.GRU(input_size=self.input_size, hidden_size=n_units, num_layers=n_layers, dropout=rnn_dropout,
                     batch_first=True, bidirectional=False)
        for name, param in gru.named_parameters():
            if "weight_hh" in name:
                nn.init.orthogonal_(param)
            if "weight_ih" in name:
                nn.init.xavier_uniform_(param)
        for name, param in gru.named_parameters():
            init["gru." + name] = param.detach()
        out = nn.Linear(n_units, n_classes)
        nn.init.xavier_uniform_(out.weight)
        init["out.weight"], init["out.bias"] = out.weight.detach(), out.bias.detach()
        init["h0"] = nn.init.xavier_uniform_(torch.zeros(1, 1, n_units))
        shapes = {k: tuple(v.shape) for k, v in init.items()}
        for name, (off, n) in self._slots.items():
            flat[off:off + n] = init[name].reshape(-1)
        self._flat = flat

        # --- parameters as views, registered under the reference's names and in its order
        def view(name):
            off, n = self._slots[name]
            return nn.Parameter(self._flat[off:off + n].view(shapes[name]))

        self.day_weights = nn.ParameterList([view(f"day_weights.{i}") for i in range(n_days)])
        self.day_biases = nn.ParameterList([view(f"day_biases.{i}") for i in range(n_days)])
        self.gru = _ParamGroup()
        for l in range(n_layers):
            for kind in ("weight_ih", "weight_hh", "bias_ih", "bias_hh"):
                self.gru.register_parameter(f"{kind}_l{l}", view(f"gru.{kind}_l{l}"))
        self.out = _ParamGroup()
        self.out.register_parameter("weight", view("out.weight"))
        self.out.register_parameter("bias", view("out.bias"))
        self.h0 = view("h0")
        self._shapes = shapes
        self._named_flat = [(n, p) for n, p in self.named_parameters()]
        self._engine: Optional[E.Engine] = None
        self._flat_grads: Optional[torch.Tensor] = None
        self.weights_synced = False
        # True only while the fused clip+AdamW kernel owns the updates (it refreshes the bf16 operand copies itself);
        # otherwise a foreign optimizer may have changed the parameters in place, so forward() re-derives them.
        self.fused_updates = False

    # ------------------------------------------------------------------ storage management
    def _rebind(self):
        for name, p in self._named_flat:
            off, n = self._slots[name]
            p.data = self._flat[off:off + n].view(self._shapes[name])
            p.grad = None
        self._engine = None
        self.weights_synced = False

    def _apply(self, fn, recurse=True):
        new = fn(self._flat)
        if new.dtype != torch.float32:
            raise TypeError("GRUDecoder keeps fp32 master weights; the bf16 tensor-core operands are derived internally")
        self._flat = new.contiguous()
        self._rebind()
        return self

    @property
    def flat_parameters(self) -> torch.Tensor:
        return self._flat

    def _engine_for(self, B: int, T: int, training: bool) -> E.Engine:
        if not self._flat.is_cuda:
            raise N.B2TError("GRUDecoder (b2t_b200) runs on a CUDA sm_100a device only: call .to('cuda') first; there is no CPU path")
        e = self._engine
        if e is None or B > e.max_batch or T > e.max_T or (training and not e.training_capable):
            mb = max(B, e.max_batch if e else 0)
            mt = max(T, e.max_T if e else 0)
            tr = training or (e.training_capable if e else False)
            old = e if (e is not None and e.training_capable) else None
            self._engine = None
            e = E.Engine(self._cfg, self._flat, max_batch=mb, max_T=mt, max_label_len=500, training=tr,
                         flat_grads=old.grads if old is not None else None)
            if old is not None:
                e.adopt_optimizer_state(old)       # moments and the per-segment AdamW step counters
            self._engine = e
            self.weights_synced = True
        return e

    def engine(self, B: int, T: int, training: bool = True) -> E.Engine:
        """The native engine sized for (B, T) -- used by the fused trainer path."""
        e = self._engine_for(B, T, training)
        if not self.weights_synced:
            e.refresh_weights()
            self.weights_synced = True
        return e

    def load_state_dict(self, state_dict, strict=True, assign=False):
        r = super().load_state_dict(state_dict, strict=strict, assign=False)
        self.weights_synced = False
        return r

    # ------------------------------------------------------------------ forward
    def forward(self, x, day_idx, states=None, return_state=False):
        '''
        x        (tensor)  - batch of examples (trials) of shape: (batch_size, time_series_length, neural_dim)
        day_idx  (tensor)  - tensor which is a list of day indexs corresponding to the day of each example in the batch x.
        '''
        if not x.is_cuda:
            raise N.B2TError("GRUDecoder (b2t_b200) got a CPU tensor; this implementation has no CPU path")
        B, T, _ = x.shape
        train = self.training and torch.is_grad_enabled()
        eng = self._engine_for(B, T, training=train)
        # parameters may have been changed by a foreign optimizer / load_state_dict: refresh the bf16 copies
        if not (self.fused_updates and self.weights_synced):
            eng.refresh_weights()
            self.weights_synced = True
        if train:
            import os
            # ranks of a data-parallel job share the torch seed (identical weight init): mix the rank into the dropout stream
            seed = (int(torch.randint(0, 2 ** 62, (1,)).item()) ^ (int(os.environ.get("RANK", "0")) * 0x9E3779B97F4A7C15)) & (2 ** 63 - 1)
            logits, hidden = _GRUDecoderFn.apply(self, x, day_idx, states, seed, *[p for _, p in self._named_flat])
        else:
            logits, hidden = eng.forward(x, day_idx, training=False, smooth_mode=0, states=states, want_hidden=True)
        if return_state:
            return logits, hidden
        return logits
