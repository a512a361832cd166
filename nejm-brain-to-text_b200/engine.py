"""Python handle on the native engine: owns the flat torch buffers and calls the C ABI.

PyTorch is used here for device memory and streams only; all arithmetic happens in
libb2t_b200.so (hand-written sm_100a kernels).
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Dict, List, Optional, Tuple

import torch

from . import _native as N


def make_config(neural_dim, n_units, n_layers, n_days, n_classes, patch_size, patch_stride, rnn_dropout=0.0,
                input_dropout=0.0) -> N.Config:
    return N.Config(int(neural_dim), int(n_units), int(n_layers), int(n_days), int(n_classes), int(patch_size),
                    int(patch_stride), float(rnn_dropout), float(input_dropout))


def param_layout(cfg: N.Config) -> List[Tuple[str, int, int, int]]:
    """[(state_dict name, element offset, rows, cols)] of the flat parameter buffer."""
    n = N.check(N.lib.b2t_param_segments(C.byref(cfg)), "b2t_param_segments")
    out = []
    name = C.create_string_buffer(64)
    off, rows, cols = C.c_longlong(), C.c_longlong(), C.c_longlong()
    for i in range(n):
        N.check(N.lib.b2t_param_segment(C.byref(cfg), i, name, 64, C.byref(off), C.byref(rows), C.byref(cols)),
                "b2t_param_segment")
        out.append((name.value.decode(), off.value, rows.value, cols.value))
    return out


def param_elems(cfg: N.Config) -> int:
    return N.check(N.lib.b2t_param_elems(C.byref(cfg)), "b2t_param_elems")


def grad_elems(cfg: N.Config) -> int:
    return N.check(N.lib.b2t_grad_elems(C.byref(cfg)), "b2t_grad_elems")


def flat_from_state_dict(cfg: N.Config, sd) -> torch.Tensor:
    """Pack a reference-style state_dict (tensors or arrays) into the flat fp32 layout (CPU tensor)."""
    flat = torch.zeros(param_elems(cfg), dtype=torch.float32)
    for name, off, rows, cols in param_layout(cfg):
        key = name if name in sd else next((k for k in sd if k.endswith(name) and k[:-len(name)] in ("_orig_mod.", "module.")), None)
        if key is None:
            raise KeyError(f"state_dict has no entry for {name}")
        v = torch.as_tensor(sd[key]).detach().to(torch.float32).reshape(-1)
        if v.numel() != rows * cols:
            raise ValueError(f"{name}: expected {rows}x{cols} elements, got {v.numel()}")
        flat[off:off + rows * cols] = v
    return flat


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


class Engine:
    """One engine = one (config, max batch, max T) workspace on one GPU."""

    def __init__(self, cfg: N.Config, flat_params: torch.Tensor, *, max_batch: int, max_T: int, max_label_len: int = 500,
                 training: bool = True, flat_grads: Optional[torch.Tensor] = None):
        if not flat_params.is_cuda:
            raise N.B2TError("the b2t_b200 engine runs on CUDA (sm_100a) only; there is no CPU path")
        assert flat_params.dtype == torch.float32 and flat_params.is_contiguous()
        self.cfg = cfg
        self.device = flat_params.device
        self.training_capable = bool(training)
        self.max_batch, self.max_T, self.max_label_len = int(max_batch), int(max_T), int(max_label_len)
        self.params = flat_params
        self.n_params = param_elems(cfg)
        assert flat_params.numel() == self.n_params, (flat_params.numel(), self.n_params)
        with torch.cuda.device(self.device):
            if training:
                self.grads = flat_grads if flat_grads is not None else torch.zeros(grad_elems(cfg), device=self.device)
                self.exp_avg = torch.zeros(self.n_params, device=self.device)
                self.exp_avg_sq = torch.zeros(self.n_params, device=self.device)
            else:
                self.grads = self.exp_avg = self.exp_avg_sq = None
            nbytes = N.check(N.lib.b2t_workspace_bytes(C.byref(cfg), self.max_batch, self.max_T, self.max_label_len,
                                                       int(training)), "b2t_workspace_bytes")
            self.workspace = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
            h = N.lib.b2t_engine_create(C.byref(cfg), self.max_batch, self.max_T, self.max_label_len, int(training),
                                        _ptr(self.params), _ptr(self.grads), _ptr(self.exp_avg), _ptr(self.exp_avg_sq),
                                        _ptr(self.workspace), nbytes)
            if not h:
                raise N.B2TError("b2t_engine_create failed: " + N.last_error())
            self.handle = h
            self.refresh_weights()
        self._day_buf = torch.zeros(self.max_batch, dtype=torch.int32, device=self.device)
        self.stats = torch.zeros(2, device=self.device)
        self.last_Tp = 0
        self.last_B = 0

    def __del__(self):
        h = getattr(self, "handle", None)
        if h and N is not None and getattr(N, "lib", None) is not None:     # module globals may already be gone at interpreter exit
            N.lib.b2t_engine_destroy(h)
            self.handle = None

    # ------------------------------------------------------------------ optimizer step counters
    def steps_tensor(self) -> torch.Tensor:
        """Per-segment AdamW step counters (device int32, aliasing the engine's own memory; b2t_step_counters)."""
        n = len(param_layout(self.cfg))
        ptr = N.lib.b2t_step_counters(self.handle)

        class _Wrap:
            pass
        w = _Wrap()
        w.__cuda_array_interface__ = {"shape": (n,), "typestr": "<i4", "data": (int(ptr), False), "version": 2}
        return torch.as_tensor(w, device=self.device)

    def adopt_optimizer_state(self, old: "Engine"):
        """Carry the AdamW state of ``old`` (moments AND per-segment step counters) into this engine.  Used when the model
        regrows its engine for a larger batch or a longer trial: resetting the step counters while the moments stay warm
        would change the bias correction (bc1 = 1 - beta1**step) and with it the update size."""
        self.exp_avg.copy_(old.exp_avg)
        self.exp_avg_sq.copy_(old.exp_avg_sq)
        self.steps_tensor().copy_(old.steps_tensor())

    # ------------------------------------------------------------------ weights
    def refresh_weights(self):
        N.check(N.lib.b2t_refresh_weights(self.handle, _stream()), "b2t_refresh_weights")

    # ------------------------------------------------------------------ forward
    def forward(self, x: torch.Tensor, day_idx, *, training: bool = False, smooth_mode: int = 0, smooth_std: float = 2.0,
                smooth_size: int = 100, cut: int = 0, white_noise_std: float = 0.0, offset_noise_std: float = 0.0,
                white_noise: Optional[torch.Tensor] = None, offset_noise: Optional[torch.Tensor] = None, seed: int = 0,
                states: Optional[torch.Tensor] = None, want_logits: bool = True, want_hidden: bool = False):
        assert x.is_cuda and x.dim() == 3 and x.shape[2] == self.cfg.neural_dim
        x = x.contiguous().float()
        B, T, _ = x.shape
        if torch.is_tensor(day_idx):
            self._day_buf[:B].copy_(day_idx.to(torch.int32), non_blocking=True)
        else:
            self._day_buf[:B].copy_(torch.tensor([int(d) for d in day_idx], dtype=torch.int32), non_blocking=True)
        ntaps = 9  # only used for the T' preview below; the library recomputes the taps
        a = N.ForwardArgs()
        a.x = x.data_ptr(); a.B = B; a.T = T; a.day_idx = self._day_buf.data_ptr()
        a.training = int(training); a.smooth_mode = int(smooth_mode); a.smooth_std = float(smooth_std)
        a.smooth_size = int(smooth_size); a.cut = int(cut)
        a.white_noise_std = float(white_noise_std); a.offset_noise_std = float(offset_noise_std)
        a.white_noise = _ptr(white_noise.contiguous().float()) if white_noise is not None else None
        a.offset_noise = _ptr(offset_noise.contiguous().float()) if offset_noise is not None else None
        a.seed = int(seed) & 0xFFFFFFFFFFFFFFFF
        if states is not None:
            states = states.contiguous().float()
            a.states = states.data_ptr()
        logits = None
        hidden = torch.empty((self.cfg.n_layers, B, self.cfg.n_units), device=self.device) if want_hidden else None
        Tp = N.lib.b2t_output_frames(C.byref(self.cfg), T, int(smooth_mode), self._ntaps(smooth_mode, smooth_std, smooth_size),
                                     int(cut) if training else 0)
        if want_logits:
            logits = torch.empty((B, max(Tp, 1), self.cfg.n_classes), device=self.device)
            a.logits_out = logits.data_ptr()
        if want_hidden:
            a.hidden_out = hidden.data_ptr()
        rc = N.check(N.lib.b2t_forward(self.handle, C.byref(a), _stream()), "b2t_forward")
        assert rc == Tp, (rc, Tp)
        self.last_Tp, self.last_B = Tp, B
        self._keep = (x, white_noise, offset_noise, states)   # keep inputs alive until the stream has consumed them
        return logits, hidden

    _taps_cache: Dict[Tuple[float, int], int] = {}

    @classmethod
    def _ntaps(cls, mode, std, size) -> int:
        if mode == 0:
            return 0
        key = (float(std), int(size))
        if key not in cls._taps_cache:
            import math
            radius = int(4.0 * std + 0.5)
            w = [math.exp(-0.5 / (std * std) * i * i) for i in range(-radius, radius + 1)]
            s = sum(w)
            cls._taps_cache[key] = sum(1 for i, v in enumerate(w) if 0 <= size // 2 - radius + i < size and v / s > 0.01)
        return cls._taps_cache[key]

    # ------------------------------------------------------------------ loss / backward / optimizer
    @staticmethod
    def _trim_labels(labels: torch.Tensor, tgt_len: torch.Tensor, max_target_len: Optional[int]) -> torch.Tensor:
        """Labels arrive zero-padded to a fixed width (the reference dataset pads seq_class_ids to 500); the CTC kernels
        pick their variant from the padded width, so cut the padding down to the longest target of the batch.  The width
        is taken from ``max_target_len`` (known on the host, no device sync) or from a CPU ``tgt_len``."""
        if max_target_len is None and not tgt_len.is_cuda and tgt_len.numel() > 0:
            max_target_len = int(tgt_len.max())
        if max_target_len is not None:
            w = max(1, min(int(max_target_len), labels.shape[1]))
            if w < labels.shape[1]:
                labels = labels[:, :w]
        return labels

    def ctc_loss(self, labels: torch.Tensor, in_len: torch.Tensor, tgt_len: torch.Tensor, *, grad_scale: float,
                 want_grad: bool = True, max_target_len: Optional[int] = None) -> torch.Tensor:
        labels = self._trim_labels(labels, tgt_len, max_target_len)
        labels = labels.to(device=self.device, dtype=torch.int32).contiguous()
        in_len = in_len.to(device=self.device, dtype=torch.int32).contiguous()
        tgt_len = tgt_len.to(device=self.device, dtype=torch.int32).contiguous()
        loss = torch.empty(labels.shape[0], device=self.device)
        N.check(N.lib.b2t_ctc_loss(self.handle, labels.data_ptr(), labels.shape[1], in_len.data_ptr(), tgt_len.data_ptr(),
                                   float(grad_scale), loss.data_ptr(), int(want_grad), _stream()), "b2t_ctc_loss")
        self._keep_ctc = (labels, in_len, tgt_len)
        return loss

    def set_dlogits(self, dlogits: torch.Tensor):
        d = dlogits.contiguous().float()
        assert d.shape == (self.last_B, self.last_Tp, self.cfg.n_classes), (d.shape, self.last_B, self.last_Tp)
        N.check(N.lib.b2t_set_dlogits(self.handle, d.data_ptr(), _stream()), "b2t_set_dlogits")
        self._keep_dl = d

    def backward(self):
        N.check(N.lib.b2t_backward(self.handle, _stream()), "b2t_backward")

    def reserve_comm_sms(self, n_sms: int):
        """Data parallel: the tail of backward leaves n_sms SMs to the gradient all-reduce (b2t_set_comm_sms)."""
        N.check(N.lib.b2t_set_comm_sms(self.handle, int(n_sms)), "b2t_set_comm_sms")

    def all_reduce_grads(self, group=None):
        """Data-parallel gradient exchange: SUM all-reduce of the flat gradient buffer (gradients + day-touched flags) on a side
        stream behind the engine's gradient-bucket events; the current stream continues only when everything has been reduced.
        Call right after backward().  One collective by default; bucket by bucket (overlapping the tail of backward) on request."""
        import torch.distributed as dist
        if getattr(self, "_comm_stream", None) is None:
            self._comm_stream = torch.cuda.Stream(device=self.device)
        n = N.check(N.lib.b2t_grad_buckets(self.handle), "b2t_grad_buckets")
        off, cnt = C.c_longlong(), C.c_longlong()
        works = []
        # Default: ONE collective over the whole buffer, issued behind the last bucket.  Measured on B200 boxes (weak scaling, ms per step,
        # one collective vs bucket-by-bucket overlap with the backward tail): N=4 3.30 vs 3.62, N=8 3.62 vs 3.99 -- the bucketed
        # collectives and the tail GEMMs slow each other down more than the overlap saves.  B2T_DP_BUCKETS=bucketed selects the overlap.
        if os.environ.get("B2T_DP_BUCKETS", "single") != "bucketed":
            lo, hi = None, 0
            with torch.cuda.stream(self._comm_stream):
                for i in range(n):
                    N.check(N.lib.b2t_grad_bucket(self.handle, i, C.byref(off), C.byref(cnt)), "b2t_grad_bucket")
                    N.check(N.lib.b2t_grad_bucket_wait(self.handle, i, self._comm_stream.cuda_stream), "b2t_grad_bucket_wait")
                    lo = off.value if lo is None else min(lo, off.value); hi = max(hi, off.value + cnt.value)
                dist.all_reduce(self.grads[lo:hi], group=group, async_op=True).wait()
            return
        with torch.cuda.stream(self._comm_stream):
            pend = None                                  # [offset, count, bucket id, largest member] not yet issued
            def flush():
                if pend is not None:
                    works.append(dist.all_reduce(self.grads[pend[0]:pend[0] + pend[1]], group=group, async_op=True))
            for i in range(n):
                k = N.check(N.lib.b2t_grad_bucket(self.handle, i, C.byref(off), C.byref(cnt)), "b2t_grad_bucket")
                o, c = off.value, cnt.value
                # neighbouring small buckets (the layers above layer 0's input weights and the head become final together when the
                # weight gradients are batched) travel as one collective: each all-reduce costs ~25 us of latency on its own
                merge = pend is not None and max(c, pend[3]) < (6 << 20) and (pend[0] + pend[1] == o or o + c == pend[0])
                if not merge:
                    flush()
                    pend = None
                N.check(N.lib.b2t_grad_bucket_wait(self.handle, i, self._comm_stream.cuda_stream), "b2t_grad_bucket_wait")
                pend = [min(pend[0], o), pend[1] + c, k, max(pend[3], c)] if merge else [o, c, k, c]
            flush()
        for w in works:
            w.wait()                                     # current stream waits for the collective

    def optimizer_step(self, lr, weight_decay, beta1, beta2, eps, max_grad_norm) -> torch.Tensor:
        a = N.AdamWArgs()
        for i in range(3):
            a.lr[i] = float(lr[i]); a.weight_decay[i] = float(weight_decay[i])
        a.beta1, a.beta2, a.eps, a.max_grad_norm = float(beta1), float(beta2), float(eps), float(max_grad_norm)
        N.check(N.lib.b2t_optimizer_step(self.handle, C.byref(a), self.stats.data_ptr(), _stream()), "b2t_optimizer_step")
        return self.stats

    def greedy_edit(self, labels, in_len, tgt_len, max_target_len: Optional[int] = None):
        labels = self._trim_labels(labels, tgt_len, max_target_len)
        labels = labels.to(device=self.device, dtype=torch.int32).contiguous()
        in_len = in_len.to(device=self.device, dtype=torch.int32).contiguous()
        tgt_len = tgt_len.to(device=self.device, dtype=torch.int32).contiguous()
        B = labels.shape[0]
        dec = torch.empty((B, self.last_Tp), dtype=torch.int32, device=self.device)
        dlen = torch.empty(B, dtype=torch.int32, device=self.device)
        ed = torch.empty(B, dtype=torch.int32, device=self.device)
        N.check(N.lib.b2t_greedy_edit(self.handle, labels.data_ptr(), labels.shape[1], in_len.data_ptr(), tgt_len.data_ptr(),
                                      dec.data_ptr(), dlen.data_ptr(), ed.data_ptr(), _stream()), "b2t_greedy_edit")
        return dec, dlen, ed

    def debug_buffer(self, name: str, layer: int = 0) -> torch.Tensor:
        """Test hook (b2t_debug_buffer): an internal activation buffer of the last forward/backward as a flat device tensor
        aliasing the engine's workspace (bf16 or fp32 depending on the buffer, see include/b2t_b200.h)."""
        ptr, n = C.c_void_p(), C.c_longlong()
        N.check(N.lib.b2t_debug_buffer(self.handle, name.encode(), int(layer), C.byref(ptr), C.byref(n)), "b2t_debug_buffer")
        f32 = name in ("gx", "dY", "logits")

        class _Wrap:
            pass
        w = _Wrap()
        w.__cuda_array_interface__ = {"shape": (n.value,), "typestr": "<f4" if f32 else "<i2", "data": (int(ptr.value), False), "version": 2}
        t = torch.as_tensor(w, device=self.device)
        return t if f32 else t.view(torch.bfloat16)

    def touched_days(self) -> torch.Tensor:
        return self.grads[self.n_params:self.n_params + self.cfg.n_days]


def gemm_bf16(A: torch.Tensor, B: torch.Tensor, *, a_mn=False, b_mn=False, out_bf16=False, bias=None) -> torch.Tensor:
    """Test hook: C = A B^T through the tcgen05 GEMM.  A: [M,K] (or [K,M] when a_mn), B: [N,K] (or [K,N] when b_mn)."""
    assert A.dtype == torch.bfloat16 and B.dtype == torch.bfloat16 and A.is_cuda
    A = A.contiguous(); B = B.contiguous()
    M, K = (A.shape[1], A.shape[0]) if a_mn else A.shape
    Nn = B.shape[1] if b_mn else B.shape[0]
    out = torch.empty((M, Nn), device=A.device, dtype=torch.bfloat16 if out_bf16 else torch.float32)
    N.check(N.lib.b2t_gemm_bf16(A.data_ptr(), B.data_ptr(), out.data_ptr(), M, Nn, K, int(a_mn), int(b_mn), int(out_bf16),
                                _ptr(bias), _stream()), "b2t_gemm_bf16")
    return out
