"""Drop-in for ``model_training/data_augmentations.py`` (gauss_smooth) on the sm_100a library.

Reference: data_augmentations.py:6-37.  Same signature; ``inputs`` must be a CUDA tensor [B, T, N];
returns a tensor of the same dtype ([B, T, N] for padding='same', [B, T-K+1, N] for 'valid').
"""
import torch

from . import _native as N


def gauss_smooth(inputs, device, smooth_kernel_std=2, smooth_kernel_size=100, padding='same'):
    if not inputs.is_cuda:
        raise N.B2TError("gauss_smooth (b2t_b200) needs a CUDA tensor; there is no CPU path")
    if padding not in ('same', 'valid'):
        raise ValueError(f"padding must be 'same' or 'valid', got {padding!r}")
    x = inputs.contiguous().float()
    B, T, C = x.shape
    mode = 1 if padding == 'same' else 2
    out = torch.empty((B, T, C), device=x.device, dtype=torch.float32)
    t_out = N.check(N.lib.b2t_gauss_smooth(x.data_ptr(), B, T, C, float(smooth_kernel_std), int(smooth_kernel_size), mode,
                                           out.data_ptr(), torch.cuda.current_stream().cuda_stream), "b2t_gauss_smooth")
    if t_out != T:
        out = out.view(-1)[:B * t_out * C].view(B, t_out, C)
    return out.to(inputs.dtype)
