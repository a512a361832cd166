"""ctypes binding of libb2t_b200.so (the C ABI declared in include/b2t_b200.h).

There is deliberately no fallback: if the library is missing or fails to load, importing this
module raises, and every op of the package is unusable.  The product path never touches oracle/.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("B2T_LIB_PATH") or os.path.join(_HERE, "libb2t_b200.so")     # B2T_LIB_PATH: A/B builds of the same ABI


class B2TError(RuntimeError):
    pass


class Config(C.Structure):
    _fields_ = [("neural_dim", C.c_int), ("n_units", C.c_int), ("n_layers", C.c_int), ("n_days", C.c_int),
                ("n_classes", C.c_int), ("patch_size", C.c_int), ("patch_stride", C.c_int),
                ("rnn_dropout", C.c_float), ("input_dropout", C.c_float)]


class ForwardArgs(C.Structure):
    _fields_ = [("x", C.c_void_p), ("B", C.c_int), ("T", C.c_int), ("day_idx", C.c_void_p), ("training", C.c_int),
                ("smooth_mode", C.c_int), ("smooth_std", C.c_float), ("smooth_size", C.c_int), ("cut", C.c_int),
                ("white_noise_std", C.c_float), ("offset_noise_std", C.c_float), ("white_noise", C.c_void_p),
                ("offset_noise", C.c_void_p), ("seed", C.c_ulonglong), ("states", C.c_void_p),
                ("logits_out", C.c_void_p), ("hidden_out", C.c_void_p)]


class AdamWArgs(C.Structure):
    _fields_ = [("lr", C.c_float * 3), ("weight_decay", C.c_float * 3), ("beta1", C.c_float), ("beta2", C.c_float),
                ("eps", C.c_float), ("max_grad_norm", C.c_float)]


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} not found: build it with `python nejm-brain-to-text_b200/build.py` "
            "(nvcc, sm_100a).  There is no CPU or PyTorch fallback for this package.")
    lib = C.CDLL(LIB_PATH)
    vp, ci, cf, ll = C.c_void_p, C.c_int, C.c_float, C.c_longlong
    cfgp = C.POINTER(Config)
    sig = {
        "b2t_last_error": (C.c_char_p, []),
        "b2t_version": (ci, []),
        "b2t_launch_count": (ll, []),
        "b2t_param_segments": (ci, [cfgp]),
        "b2t_param_segment": (ci, [cfgp, ci, C.c_char_p, ci, C.POINTER(ll), C.POINTER(ll), C.POINTER(ll)]),
        "b2t_param_elems": (ll, [cfgp]),
        "b2t_grad_elems": (ll, [cfgp]),
        "b2t_workspace_bytes": (ll, [cfgp, ci, ci, ci, ci]),
        "b2t_engine_create": (vp, [cfgp, ci, ci, ci, ci, vp, vp, vp, vp, vp, ll]),
        "b2t_engine_destroy": (None, [vp]),
        "b2t_refresh_weights": (ci, [vp, vp]),
        "b2t_forward": (ci, [vp, C.POINTER(ForwardArgs), vp]),
        "b2t_output_frames": (ci, [cfgp, ci, ci, ci, ci]),
        "b2t_gauss_smooth": (ci, [vp, ci, ci, ci, cf, ci, ci, vp, vp]),
        "b2t_ctc_loss": (ci, [vp, vp, ci, vp, vp, cf, vp, ci, vp]),
        "b2t_ctc_loss_tbc": (ci, [vp, ci, ci, ci, vp, ci, vp, vp, cf, vp, vp, vp, ll, vp]),
        "b2t_ctc_workspace_bytes": (ll, [ci, ci, ci]),
        "b2t_set_dlogits": (ci, [vp, vp, vp]),
        "b2t_backward": (ci, [vp, vp]),
        "b2t_optimizer_step": (ci, [vp, C.POINTER(AdamWArgs), vp, vp]),
        "b2t_grad_buckets": (ci, [vp]),
        "b2t_set_comm_sms": (ci, [vp, ci]),
        "b2t_grad_bucket": (ci, [vp, ci, C.POINTER(ll), C.POINTER(ll)]),
        "b2t_grad_bucket_wait": (ci, [vp, ci, vp]),
        "b2t_step_counters": (vp, [vp]),
        "b2t_debug_set_trace": (ci, [vp, vp]),
        "b2t_debug_buffer": (ci, [vp, C.c_char_p, ci, C.POINTER(vp), C.POINTER(ll)]),
        "b2t_debug_timeline": (ci, [ci]),
        "b2t_debug_dump_timeline": (ci, [C.c_char_p, ci]),
        "b2t_greedy_edit": (ci, [vp, vp, ci, vp, vp, vp, vp, vp, vp]),
        "b2t_gemm_bf16": (ci, [vp, vp, vp, ci, ci, ci, ci, ci, ci, vp, vp]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    return lib, sig


lib, SIGNATURES = _load()


def last_error() -> str:
    return lib.b2t_last_error().decode("utf-8", "replace")


def check(rc: int, what: str) -> int:
    if rc < 0:
        raise B2TError(f"{what} failed ({rc}): {last_error()}")
    return rc


def optional_symbols():
    """Decoder entry points are bound lazily by lm_decoder.py (they live in the same library)."""
    return lib
