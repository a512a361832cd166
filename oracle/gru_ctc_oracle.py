"""CPU oracle for the GRU -> CTC training / inference path.  TEST INFRASTRUCTURE ONLY.

This file is a numpy restatement of the reference algorithm.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
leg may import it; the product package never does (it fails loudly when the CUDA
library is missing).

Parity pin: every function here is checked against the *imported* reference code
(``/root/reference/model_training/rnn_model.py``, ``data_augmentations.py``) plus
``torch.nn.CTCLoss`` / ``torch.autograd`` / ``torch.optim.AdamW`` by
``oracle/gen_golden.py`` (run in the build container, where /root/reference exists);
the resulting input/output vectors are committed under ``tests/golden/`` and
re-checked by ``tests/test_oracle_golden.py`` on every run.  The reference itself
ships no tests for this path, so those generated vectors are the pin.

Reference call sites restated (file:line under /root/reference/model_training):
  gauss taps ............ data_augmentations.py:19-24
  gauss_smooth .......... data_augmentations.py:27-37
  transform_data ........ rnn_trainer.py:436-484
  adjusted_lens ......... rnn_trainer.py:532
  day layer + softsign .. rnn_model.py:95-99
  input dropout ......... rnn_model.py:102-103
  patch unfold .......... rnn_model.py:106-119
  h0 expand ............. rnn_model.py:122-123
  GRU (torch.nn.GRU) .... rnn_model.py:65-72,126   gate order r,z,n
  head .................. rnn_model.py:129
  log_softmax + CTC ..... rnn_trainer.py:538-545   blank=0, reduction none -> mean
  clip_grad_norm_ ....... rnn_trainer.py:550-555
  AdamW / LambdaLR ...... rnn_trainer.py:259-292, 294-363
  greedy decode + PER ... rnn_trainer.py:724-736, 764
"""
from __future__ import annotations

import math
import numpy as np

NEG_INF = -np.inf


# ----------------------------------------------------------------------------
# smoothing (data_augmentations.py:6-37)
# ----------------------------------------------------------------------------
def gauss_taps(std: float = 2.0, size: int = 100) -> np.ndarray:
    """Impulse response of scipy.ndimage.gaussian_filter1d(sigma=std) on a unit impulse
    of length `size`, thresholded at > 0.01 and renormalised (data_augmentations.py:19-24).
    scipy uses truncate=4.0 -> radius int(4*std+0.5); weights exp(-0.5 x^2/std^2)/sum in
    float64, result cast to the float32 input dtype."""
    radius = int(4.0 * float(std) + 0.5)
    x = np.arange(-radius, radius + 1, dtype=np.float64)
    w = np.exp(-0.5 / (float(std) * float(std)) * x * x)
    w = w / w.sum()
    imp = np.zeros(size, dtype=np.float64)
    c = size // 2
    lo = c - radius
    for i, wi in enumerate(w):
        j = lo + i
        if 0 <= j < size:
            imp[j] += wi
    imp = imp.astype(np.float32)
    k = imp[imp > 0.01]
    k = (k / np.sum(k)).astype(np.float32)
    return k


def gauss_smooth(x: np.ndarray, std: float = 2.0, size: int = 100, padding: str = "same") -> np.ndarray:
    """Depthwise cross-correlation along time with the taps above.
    x: [B, T, C] -> [B, T, C] ('same', zero padded) or [B, T-K+1, C] ('valid')."""
    k = gauss_taps(std, size).astype(np.float64)
    K = len(k)
    B, T, C = x.shape
    xd = x.astype(np.float64)
    if padding == "same":
        left = (K - 1) // 2
        right = K - 1 - left
        xp = np.concatenate([np.zeros((B, left, C)), xd, np.zeros((B, right, C))], axis=1)
        To = T
    elif padding == "valid":
        xp = xd
        To = T - K + 1
    else:
        raise ValueError(padding)
    out = np.zeros((B, To, C), dtype=np.float64)
    for j in range(K):
        out += k[j] * xp[:, j:j + To, :]
    return out.astype(np.float32)


def transform_data(x, n_time_steps, *, mode, white_noise=None, offset_noise=None, cut=0,
                   white_noise_std=1.0, constant_offset_std=0.2, smooth=True,
                   smooth_kernel_std=2.0, smooth_kernel_size=100):
    """rnn_trainer.py:436-484 with the random draws injected (white_noise: [B,T,C] ~N(0,1),
    offset_noise: [B,1,C] ~N(0,1), cut: int) so that the CUDA path can be fed the same draws."""
    x = x.astype(np.float32).copy()
    n = np.asarray(n_time_steps).copy()
    if mode == "train":
        if white_noise is not None and white_noise_std > 0:
            x = x + white_noise.astype(np.float32) * np.float32(white_noise_std)
        if offset_noise is not None and constant_offset_std > 0:
            x = x + offset_noise.astype(np.float32) * np.float32(constant_offset_std)
        if cut > 0:
            x = x[:, cut:, :]
            n = n - cut
    if smooth:
        x = gauss_smooth(x, smooth_kernel_std, smooth_kernel_size, "same")
    return x, n


def adjusted_lens(n_time_steps, patch_size=14, patch_stride=4):
    """rnn_trainer.py:532 -- float divide then truncation to int32."""
    n = np.asarray(n_time_steps, dtype=np.float32)
    return ((n - patch_size) / patch_stride + 1).astype(np.int32)


# ----------------------------------------------------------------------------
# model forward / backward (rnn_model.py:88-134)
# ----------------------------------------------------------------------------
def sigmoid(x):
    return 1.0 / (1.0 + np.exp(-x))


def softsign(x):
    return x / (1.0 + np.abs(x))


def unfold_patches(xd, patch_size, patch_stride):
    """[B,T,D] -> [B,T',patch*D], element (p*D + d) = xd[b, t'*stride + p, d] (rnn_model.py:106-119)."""
    B, T, D = xd.shape
    if patch_size <= 0:
        return xd
    Tp = (T - patch_size) // patch_stride + 1
    out = np.empty((B, Tp, patch_size * D), dtype=xd.dtype)
    for p in range(patch_size):
        out[:, :, p * D:(p + 1) * D] = xd[:, p:p + (Tp - 1) * patch_stride + 1:patch_stride, :]
    return out


def fold_patches_grad(dX, T, D, patch_size, patch_stride):
    B, Tp, _ = dX.shape
    if patch_size <= 0:
        return dX
    dxd = np.zeros((B, T, D), dtype=dX.dtype)
    for p in range(patch_size):
        dxd[:, p:p + (Tp - 1) * patch_stride + 1:patch_stride, :] += dX[:, :, p * D:(p + 1) * D]
    return dxd


class Params(dict):
    """state_dict-keyed numpy parameters: day_weights.{i}, day_biases.{i}, gru.weight_ih_l{k},
    gru.weight_hh_l{k}, gru.bias_ih_l{k}, gru.bias_hh_l{k}, out.weight, out.bias, h0."""

    @property
    def n_layers(self):
        return sum(1 for k in self if k.startswith("gru.weight_hh_l"))

    @property
    def n_units(self):
        return self["gru.weight_hh_l0"].shape[1]


def forward(params: Params, x, day_idx, *, patch_size=14, patch_stride=4, states=None,
            in_mask=None, layer_masks=None, dtype=np.float64, keep_cache=False):
    """GRUDecoder.forward.  x: [B,T,D] (already smoothed), day_idx: [B].
    in_mask: optional [B,T,D] multiplicative dropout mask (already scaled by 1/keep).
    layer_masks: optional list (n_layers-1) of [B,T',H] scaled masks applied to the output of
    layers 0..L-2 (torch.nn.GRU dropout semantics).  Returns logits [B,T',C], hidden [L,B,H]."""
    f = dtype
    B, T, D = x.shape
    L = params.n_layers
    H = params.n_units
    xd_pre = np.empty((B, T, D), dtype=f)
    for b in range(B):
        d = int(day_idx[b])
        xd_pre[b] = x[b].astype(f) @ params[f"day_weights.{d}"].astype(f) + params[f"day_biases.{d}"].astype(f)
    xd = softsign(xd_pre)
    if in_mask is not None:
        xd = xd * in_mask.astype(f)
    X = unfold_patches(xd, patch_size, patch_stride)
    Tp = X.shape[1]
    if states is None:
        states = np.broadcast_to(params["h0"].astype(f).reshape(1, 1, H), (L, B, H)).copy()
    cache = {"x": x, "xd_pre": xd_pre, "in_mask": in_mask, "layers": [], "states": states,
             "T": T, "D": D, "layer_masks": layer_masks}
    hidden = np.empty((L, B, H), dtype=f)
    inp = X
    for l in range(L):
        Wih = params[f"gru.weight_ih_l{l}"].astype(f)
        Whh = params[f"gru.weight_hh_l{l}"].astype(f)
        bih = params[f"gru.bias_ih_l{l}"].astype(f)
        bhh = params[f"gru.bias_hh_l{l}"].astype(f)
        gx = inp @ Wih.T + bih                     # [B,T',3H]
        h = states[l].astype(f)
        out = np.empty((B, Tp, H), dtype=f)
        R = np.empty((B, Tp, H), dtype=f); Z = np.empty_like(R); N = np.empty_like(R)
        HN = np.empty_like(R); HP = np.empty_like(R)
        for t in range(Tp):
            gh = h @ Whh.T + bhh
            r = sigmoid(gx[:, t, :H] + gh[:, :H])
            z = sigmoid(gx[:, t, H:2 * H] + gh[:, H:2 * H])
            hn = gh[:, 2 * H:]
            n = np.tanh(gx[:, t, 2 * H:] + r * hn)
            HP[:, t] = h
            h = (1.0 - z) * n + z * h
            out[:, t] = h
            R[:, t] = r; Z[:, t] = z; N[:, t] = n; HN[:, t] = hn
        hidden[l] = h
        lay = {"inp": inp, "R": R, "Z": Z, "N": N, "HN": HN, "HP": HP, "out": out}
        if layer_masks is not None and l < L - 1:
            inp = out * layer_masks[l].astype(f)
        else:
            inp = out
        cache["layers"].append(lay)
    logits = inp @ params["out.weight"].astype(f).T + params["out.bias"].astype(f)
    cache["top"] = inp
    if keep_cache:
        return logits, hidden, cache
    return logits, hidden


def backward(params: Params, cache, dlogits, day_idx, *, patch_size=14, patch_stride=4, dtype=np.float64):
    """Manual BPTT matching torch.autograd on the reference module.  Returns grads dict with the
    same keys as params (day layers never touched are absent, mirroring grad=None)."""
    f = dtype
    L = params.n_layers
    H = params.n_units
    B, Tp, C = dlogits.shape
    g = {}
    top = cache["top"]
    dl2 = dlogits.reshape(B * Tp, C).astype(f)
    g["out.weight"] = dl2.T @ top.reshape(B * Tp, H)
    g["out.bias"] = dl2.sum(0)
    dout = dlogits.astype(f) @ params["out.weight"].astype(f)        # grad wrt (masked) top-layer output
    dh0 = np.zeros((H,), dtype=f)
    for l in range(L - 1, -1, -1):
        lay = cache["layers"][l]
        if cache["layer_masks"] is not None and l < L - 1:
            dout = dout * cache["layer_masks"][l].astype(f)
        Wih = params[f"gru.weight_ih_l{l}"].astype(f)
        Whh = params[f"gru.weight_hh_l{l}"].astype(f)
        R, Z, N, HN, HP = lay["R"], lay["Z"], lay["N"], lay["HN"], lay["HP"]
        dGx = np.empty((B, Tp, 3 * H), dtype=f)
        dGh = np.empty((B, Tp, 3 * H), dtype=f)
        dh = np.zeros((B, H), dtype=f)
        for t in range(Tp - 1, -1, -1):
            dht = dh + dout[:, t]
            r, z, n, hn, hp = R[:, t], Z[:, t], N[:, t], HN[:, t], HP[:, t]
            dn = dht * (1.0 - z)
            dz = dht * (hp - n)
            dn_pre = dn * (1.0 - n * n)
            dz_pre = dz * z * (1.0 - z)
            dr = dn_pre * hn
            dr_pre = dr * r * (1.0 - r)
            dGx[:, t, :H] = dr_pre; dGx[:, t, H:2 * H] = dz_pre; dGx[:, t, 2 * H:] = dn_pre
            dGh[:, t, :H] = dr_pre; dGh[:, t, H:2 * H] = dz_pre; dGh[:, t, 2 * H:] = dn_pre * r
            dh = dht * z + dGh[:, t] @ Whh
        dh0 += dh.sum(0)
        inp = lay["inp"]
        K = inp.shape[2]
        g[f"gru.weight_ih_l{l}"] = dGx.reshape(B * Tp, 3 * H).T @ inp.reshape(B * Tp, K)
        g[f"gru.bias_ih_l{l}"] = dGx.reshape(B * Tp, 3 * H).sum(0)
        g[f"gru.weight_hh_l{l}"] = dGh.reshape(B * Tp, 3 * H).T @ HP.reshape(B * Tp, H)
        g[f"gru.bias_hh_l{l}"] = dGh.reshape(B * Tp, 3 * H).sum(0)
        dout = dGx @ Wih                                              # grad wrt this layer's input
    g["h0"] = dh0.reshape(1, 1, H)
    T, D = cache["T"], cache["D"]
    dxd = fold_patches_grad(dout, T, D, patch_size, patch_stride)
    if cache["in_mask"] is not None:
        dxd = dxd * cache["in_mask"].astype(f)
    pre = cache["xd_pre"]
    dpre = dxd / (1.0 + np.abs(pre)) ** 2
    x = cache["x"].astype(f)
    for b in range(B):
        d = int(day_idx[b])
        kw, kb = f"day_weights.{d}", f"day_biases.{d}"
        gw = x[b].T @ dpre[b]
        gb = dpre[b].sum(0, keepdims=True)
        g[kw] = g.get(kw, 0) + gw
        g[kb] = g.get(kb, 0) + gb
    return g


# ----------------------------------------------------------------------------
# CTC (torch.nn.CTCLoss(blank=0, reduction='none', zero_infinity=False))
# ----------------------------------------------------------------------------
def log_softmax(x, axis=-1):
    m = np.max(x, axis=axis, keepdims=True)
    y = x - m
    return y - np.log(np.sum(np.exp(y), axis=axis, keepdims=True))


def _lse2(a, b):
    m = np.maximum(a, b)
    with np.errstate(invalid="ignore"):
        out = m + np.log(np.exp(a - m) + np.exp(b - m))
    return np.where(np.isneginf(m), NEG_INF, out)


def _lse3(a, b, c):
    return _lse2(_lse2(a, b), c)


def ctc_loss_and_grad(logits, targets, input_lengths, target_lengths, *, blank=0, want_grad=True,
                      dtype=np.float64):
    """logits: [B,T,C] raw (log_softmax is applied here, rnn_trainer.py:539); targets: [B,Smax].
    Returns (loss[B], dlogits[B,T,C]) where dlogits is d(mean_b loss_b)/dlogits, i.e. includes the
    1/B of torch.mean (rnn_trainer.py:545).  Frames t >= input_length get zero grad."""
    f = dtype
    B, T, C = logits.shape
    lp = log_softmax(logits.astype(f), -1)
    losses = np.empty((B,), dtype=f)
    dlogits = np.zeros((B, T, C), dtype=f)
    for b in range(B):
        S = int(target_lengths[b]); Tb = int(input_lengths[b])
        lab = np.asarray(targets[b][:S], dtype=np.int64)
        Lp = 2 * S + 1
        ext = np.full((Lp,), blank, dtype=np.int64)
        ext[1::2] = lab
        skip = np.zeros((Lp,), dtype=bool)      # may take the s-2 transition
        for s in range(2, Lp):
            skip[s] = ext[s] != blank and ext[s] != ext[s - 2]
        alpha = np.full((Tb, Lp), NEG_INF, dtype=f)
        if Tb > 0:
            alpha[0, 0] = lp[b, 0, blank]
            if Lp > 1:
                alpha[0, 1] = lp[b, 0, ext[1]]
        for t in range(1, Tb):
            a = alpha[t - 1]
            a1 = np.concatenate([[NEG_INF], a[:-1]])
            a2 = np.concatenate([[NEG_INF, NEG_INF], a[:-2]])
            a2 = np.where(skip, a2, NEG_INF)
            alpha[t] = _lse3(a, a1, a2) + lp[b, t, ext]
        if Tb > 0:
            ll = _lse2(alpha[Tb - 1, Lp - 1], alpha[Tb - 1, Lp - 2] if Lp > 1 else NEG_INF)
        else:
            ll = 0.0 if S == 0 else NEG_INF
        losses[b] = -ll
        if not want_grad or Tb == 0:
            continue
        beta = np.full((Tb, Lp), NEG_INF, dtype=f)
        beta[Tb - 1, Lp - 1] = lp[b, Tb - 1, blank]
        if Lp > 1:
            beta[Tb - 1, Lp - 2] = lp[b, Tb - 1, ext[Lp - 2]]
        skip_fwd = np.zeros((Lp,), dtype=bool)  # s may jump to s+2
        skip_fwd[:-2] = skip[2:]
        for t in range(Tb - 2, -1, -1):
            bt = beta[t + 1]
            b1 = np.concatenate([bt[1:], [NEG_INF]])
            b2 = np.concatenate([bt[2:], [NEG_INF, NEG_INF]])
            b2 = np.where(skip_fwd, b2, NEG_INF)
            beta[t] = _lse3(bt, b1, b2) + lp[b, t, ext]
        ab = alpha + beta                        # emission counted twice
        for t in range(Tb):
            lcab = np.full((C,), NEG_INF, dtype=f)
            for s in range(Lp):
                lcab[ext[s]] = _lse2(lcab[ext[s]], ab[t, s])
            with np.errstate(invalid="ignore", over="ignore"):
                occ = np.exp(lcab - ll - lp[b, t])
            occ = np.where(np.isneginf(lcab), 0.0, occ)
            dlogits[b, t] = (np.exp(lp[b, t]) - occ) / B
    return losses, dlogits


# ----------------------------------------------------------------------------
# optimizer (rnn_trainer.py:259-363, 550-558)
# ----------------------------------------------------------------------------
def lr_lambda(step, min_lr_ratio, decay_steps, warmup_steps):
    """rnn_trainer.py:306-325."""
    if step < warmup_steps:
        return float(step) / float(max(1, warmup_steps))
    if step < decay_steps:
        progress = float(step - warmup_steps) / float(max(1, decay_steps - warmup_steps))
        cosine = 0.5 * (1 + math.cos(math.pi * progress))
        return max(min_lr_ratio, min_lr_ratio + (1 - min_lr_ratio) * cosine)
    return min_lr_ratio


def param_group(name):
    """rnn_trainer.py:267-269 -> 'bias' | 'day' | 'other'."""
    if "gru.bias" in name or "out.bias" in name:
        return "bias"
    if "day_" in name:
        return "day"
    return "other"


def clip_grad_norm(grads, max_norm):
    """torch.nn.utils.clip_grad_norm_: returns (total_norm, clipped grads)."""
    tot = math.sqrt(sum(float(np.sum(np.asarray(v, dtype=np.float64) ** 2)) for v in grads.values()))
    coef = min(1.0, max_norm / (tot + 1e-6))
    return tot, {k: np.asarray(v) * coef for k, v in grads.items()}


def adamw_step(params, grads, state, *, step, lr, betas=(0.9, 0.999), eps=0.1, weight_decay=0.0):
    """One AdamW update of a single tensor group; `step` is 1-based; state: dict name -> (m, v)."""
    b1, b2 = betas
    for k, gk in grads.items():
        p = params[k].astype(np.float64)
        gk = np.asarray(gk, dtype=np.float64).reshape(p.shape)
        m, v = state.get(k, (np.zeros_like(p), np.zeros_like(p)))
        p = p * (1.0 - lr * weight_decay)
        m = b1 * m + (1 - b1) * gk
        v = b2 * v + (1 - b2) * gk * gk
        bc1 = 1 - b1 ** step
        bc2 = 1 - b2 ** step
        denom = np.sqrt(v) / math.sqrt(bc2) + eps
        p = p - (lr / bc1) * m / denom
        params[k] = p.astype(np.float32)
        state[k] = (m, v)


# ----------------------------------------------------------------------------
# greedy decode + edit distance (rnn_trainer.py:724-736)
# ----------------------------------------------------------------------------
def greedy_decode(logits_bt, length):
    ids = np.argmax(logits_bt[:length], axis=-1)
    out = []
    prev = None
    for i in ids:
        if prev is None or i != prev:
            out.append(int(i))
        prev = i
    return [i for i in out if i != 0]


def edit_distance(a, b):
    a = list(a); b = list(b)
    prev = list(range(len(b) + 1))
    for i in range(1, len(a) + 1):
        cur = [i] + [0] * len(b)
        for j in range(1, len(b) + 1):
            cur[j] = min(prev[j] + 1, cur[j - 1] + 1, prev[j - 1] + (a[i - 1] != b[j - 1]))
        prev = cur
    return prev[len(b)]


def rearrange_speech_logits(logits):
    """evaluate_model_helpers.py:79-83: [BLANK, phones..., SIL] -> [BLANK, SIL, phones...]."""
    return np.concatenate((logits[..., 0:1], logits[..., -1:], logits[..., 1:-1]), axis=-1)
