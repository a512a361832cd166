"""Drop-in for the ``lm_decoder`` pybind module (language_model/runtime/server/x86/python/lm_decoder.cc:51-75)
used by ``language-model-standalone.py``: same class names, constructor argument order and call protocol

    opts = DecodeOptions(max_active, min_active, beam, lattice_beam, acoustic_scale, blank_skip_threshold, length_penalty, nbest)
    res  = DecodeResource(fst_path, lm_fst_path, rescore_lm_fst_path, dict_path, unit_path)
    dec  = BrainSpeechDecoder(res, opts)
    dec.Reset(); DecodeNumpy(dec, logits, log_priors, blank_penalty); dec.FinishDecoding(); dec.result()

backed by the GPU token-passing decoder in libb2t_b200.so (C ABI b2t_decoder_*).  Differences: errors raise
Python exceptions instead of aborting the process through glog; ``DecodeBatch`` is an extension that decodes many
utterances concurrently (one CTA each).
"""
from __future__ import annotations

import ctypes as C
from typing import List

import numpy as np

from . import _native as N


class _Opts(C.Structure):
    _fields_ = [("max_active", C.c_int), ("min_active", C.c_int), ("beam", C.c_float), ("lattice_beam", C.c_float),
                ("acoustic_scale", C.c_float), ("blank_skip_threshold", C.c_float), ("length_penalty", C.c_float), ("nbest", C.c_int)]


_lib = N.lib
_vp, _ci, _cf = C.c_void_p, C.c_int, C.c_float
_lib.b2t_decoder_last_error.restype = C.c_char_p
_lib.b2t_decoder_create.restype = _vp
_lib.b2t_decoder_create.argtypes = [C.c_char_p, C.c_char_p, C.POINTER(_Opts), _ci, _ci]
_lib.b2t_decoder_destroy.argtypes = [_vp]
_lib.b2t_decoder_destroy.restype = None
_lib.b2t_decoder_set_options.argtypes = [_vp, C.POINTER(_Opts)]
_lib.b2t_decoder_reset.argtypes = [_vp, _ci]
_lib.b2t_decoder_set_strict_order.argtypes = [_vp, _ci]
_lib.b2t_decoder_decode_logits.argtypes = [_vp, _ci, _vp, _vp, _ci, _ci, _cf]
_lib.b2t_decoder_decode_logprobs.argtypes = [_vp, _ci, _vp, _ci, _ci]
_lib.b2t_decoder_finish.argtypes = [_vp, _ci]
_lib.b2t_decoder_rescore.argtypes = [_vp, _ci]
_lib.b2t_decoder_set_rescore_lms.argtypes = [_vp, C.c_char_p, C.c_char_p]
_lib.b2t_decoder_num_results.argtypes = [_vp, _ci]
_lib.b2t_decoder_get_result.argtypes = [_vp, _ci, _ci, C.POINTER(_cf), C.POINTER(_cf), C.c_char_p, _ci]
_lib.b2t_decoder_decode_batch.argtypes = [_vp, _vp, _vp, _ci, _ci, _ci, _cf, _ci]
_lib.b2t_decoder_stats.argtypes = [_vp, _ci, C.POINTER(_ci), C.POINTER(C.c_longlong), C.POINTER(C.c_longlong), C.POINTER(C.c_double)]
_lib.b2t_decoder_tokens_per_frame.argtypes = [_vp, _ci, _vp, _ci]
_lib.b2t_decoder_debug_frame_tokens.argtypes = [_vp, _ci, _ci, _vp, _vp, _ci]


def _check(rc, what):
    if rc < 0:
        raise N.B2TError(f"{what} failed ({rc}): {_lib.b2t_decoder_last_error().decode('utf-8', 'replace')}")
    return rc


class DecodeOptions:
    def __init__(self, max_active, min_active, beam, lattice_beam, acoustic_scale, blank_skip_threshold, length_penalty, nbest):
        self.c = _Opts(int(max_active), int(min_active), float(beam), float(lattice_beam), float(acoustic_scale),
                       float(blank_skip_threshold), float(length_penalty), int(nbest))


class DecodeResource:
    def __init__(self, fst_path, lm_fst_path, rescore_lm_fst_path, dict_path, unit_path):
        self.fst_path, self.lm_fst_path, self.rescore_lm_fst_path = fst_path, lm_fst_path, rescore_lm_fst_path
        self.dict_path, self.unit_path = dict_path, unit_path


class DecodeResult:
    __slots__ = ("ac_score", "lm_score", "sentence")

    def __init__(self, ac_score, lm_score, sentence):
        self.ac_score, self.lm_score, self.sentence = ac_score, lm_score, sentence

    def __repr__(self):
        return f"DecodeResult(ac_score={self.ac_score:.4f}, lm_score={self.lm_score:.4f}, sentence={self.sentence!r})"


class BrainSpeechDecoder:
    def __init__(self, resource: DecodeResource, opts: DecodeOptions, max_frames: int = 1024, max_slots: int = 1, strict_order=None):
        self._resource, self._opts = resource, opts        # the reference keeps both alive through shared_ptr
        self.max_slots = int(max_slots)
        self._h = _lib.b2t_decoder_create(resource.fst_path.encode(), resource.dict_path.encode(), C.byref(opts.c), int(max_frames),
                                          self.max_slots)
        if not self._h:
            raise N.B2TError("BrainSpeechDecoder: " + _lib.b2t_decoder_last_error().decode("utf-8", "replace"))
        if resource.lm_fst_path and resource.rescore_lm_fst_path:     # lattice LM rescoring (brain_speech_decoder.h:57-79)
            _check(_lib.b2t_decoder_set_rescore_lms(self._h, resource.lm_fst_path.encode(), resource.rescore_lm_fst_path.encode()), "DecodeResource LMs")
        if strict_order is not None:
            self.set_strict_order(strict_order)

    def set_strict_order(self, on: bool):
        """Reproduce Kaldi's serial token-list order and online cutoff tightening (results equal the reference's also when
        max_active binds); off = the faster two-pass search.  Between utterances only."""
        _check(_lib.b2t_decoder_set_strict_order(self._h, 1 if on else 0), "set_strict_order")

    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            _lib.b2t_decoder_destroy(h)
            self._h = None

    # --- pybind surface (lm_decoder.cc:63-71)
    def SetOpt(self, opts: DecodeOptions):
        self._opts = opts
        _check(_lib.b2t_decoder_set_options(self._h, C.byref(opts.c)), "SetOpt")

    def Decode(self, logp, slot: int = 0):
        """logp: [T, C] float32 log-probabilities (numpy array or CPU torch tensor)."""
        a = np.ascontiguousarray(logp.numpy() if hasattr(logp, "numpy") else logp, dtype=np.float32)
        if a.ndim != 2:
            raise ValueError("Decode expects a [T, C] matrix")
        _check(_lib.b2t_decoder_decode_logprobs(self._h, slot, a.ctypes.data, a.shape[0], a.shape[1]), "Decode")

    def Rescore(self, slot: int = 0):
        _check(_lib.b2t_decoder_rescore(self._h, slot), "Rescore")

    def Reset(self, slot: int = 0):
        _check(_lib.b2t_decoder_reset(self._h, slot), "Reset")

    def FinishDecoding(self, slot: int = 0):
        _check(_lib.b2t_decoder_finish(self._h, slot), "FinishDecoding")

    def DecodedSomething(self, slot: int = 0) -> bool:
        r = self.result(slot)
        return len(r) > 0 and len(r[0].sentence) > 0

    def result(self, slot: int = 0) -> List[DecodeResult]:
        n = _check(_lib.b2t_decoder_num_results(self._h, slot), "result")
        out, buf = [], C.create_string_buffer(1 << 16)
        ac, lm = _cf(), _cf()
        for i in range(n):
            _check(_lib.b2t_decoder_get_result(self._h, slot, i, C.byref(ac), C.byref(lm), buf, 1 << 16), "result")
            out.append(DecodeResult(ac.value, lm.value, buf.value.decode("utf-8", "replace")))
        return out

    # --- extensions
    def DecodeBatch(self, logits, lens=None, blank_penalty: float = 0.0, finish: bool = True):
        """logits: [N, T, C] float32 (N <= max_slots); decodes all utterances concurrently; read result(slot=n)."""
        a = np.ascontiguousarray(logits, dtype=np.float32)
        Nn, T, Cc = a.shape
        ln = np.full((Nn,), T, dtype=np.int32) if lens is None else np.ascontiguousarray(lens, dtype=np.int32)
        _check(_lib.b2t_decoder_decode_batch(self._h, a.ctypes.data, ln.ctypes.data, Nn, T, Cc, float(blank_penalty), int(finish)), "DecodeBatch")

    def stats(self, slot: int = 0):
        fr, tk, lk, ms = _ci(), C.c_longlong(), C.c_longlong(), C.c_double()
        _check(_lib.b2t_decoder_stats(self._h, slot, C.byref(fr), C.byref(tk), C.byref(lk), C.byref(ms)), "stats")
        return {"frames": fr.value, "tokens": tk.value, "links": lk.value, "kernel_ms": ms.value}

    def tokens_per_frame(self, slot: int = 0):
        a = np.zeros(8192, dtype=np.int32)
        n = _check(_lib.b2t_decoder_tokens_per_frame(self._h, slot, a.ctypes.data, 8192), "tokens_per_frame")
        return a[:n]

    def debug_frame_tokens(self, frame_plus_one: int, slot: int = 0, cap: int = 1 << 20):
        """Test hook: (states, costs) of one frame's tokens in pool order (= Kaldi's list order in strict mode); before FinishDecoding."""
        st = np.zeros(cap, dtype=np.int32); co = np.zeros(cap, dtype=np.float32)
        n = _check(_lib.b2t_decoder_debug_frame_tokens(self._h, slot, frame_plus_one, st.ctypes.data, co.ctypes.data, cap), "debug_frame_tokens")
        return st[:n], co[:n]


def DecodeNumpy(decoder: BrainSpeechDecoder, logits, log_priors, blank_penalty, slot: int = 0):
    """lm_decoder.cc:14-37."""
    x = np.ascontiguousarray(logits, dtype=np.float32)            # py::array::forcecast
    pr = np.asarray(log_priors, dtype=np.float32)
    if x.ndim != 2 or pr.ndim != 2:
        raise ValueError("DecodeNumpy expects 2-D logits and log_priors")
    # the reference subtracts with torch broadcasting (lm_decoder.cc:30), so a [1, C] prior is legal; the C ABI wants [T, C]
    try:
        pr = np.ascontiguousarray(np.broadcast_to(pr, x.shape), dtype=np.float32)
    except ValueError:
        raise ValueError(f"DecodeNumpy: log_priors of shape {pr.shape} do not broadcast to the logits {x.shape}") from None
    _check(_lib.b2t_decoder_decode_logits(decoder._h, slot, x.ctypes.data, pr.ctypes.data, x.shape[0], x.shape[1], float(blank_penalty)),
           "DecodeNumpy")


def DecodeNumpyLogProbs(decoder: BrainSpeechDecoder, logp, slot: int = 0):
    """lm_decoder.cc:39-49."""
    decoder.Decode(np.ascontiguousarray(logp, dtype=np.float32), slot)


_lib.b2t_prefix_last_error.restype = C.c_char_p
_lib.b2t_prefix_beam_search.argtypes = [_vp, _vp, _ci, _ci, _ci, _ci, _ci, _ci, _ci, _vp, _vp, _vp, _vp, _vp, _vp]


def ctc_prefix_beam_search(logp, lens=None, blank=0, first_beam_size=10, second_beam_size=10, max_len=None):
    """CtcPrefixBeamSearch over a batch: logp [N, T, C] (or [T, C]) float32 log-probabilities.
    Returns, per utterance, a list of (token ids, score, viterbi score, viterbi times), best first."""
    a = np.ascontiguousarray(logp, dtype=np.float32)
    if a.ndim == 2:
        a = a[None]
    Nn, T, Cc = a.shape
    ln = np.full((Nn,), T, dtype=np.int32) if lens is None else np.ascontiguousarray(lens, dtype=np.int32)
    ml = int(max_len or max(T, 1))
    ids = np.zeros((Nn, second_beam_size, ml), np.int32); ol = np.zeros((Nn, second_beam_size), np.int32)
    sc = np.zeros((Nn, second_beam_size), np.float32); vt = np.zeros_like(sc); tm = np.zeros_like(ids); nh = np.zeros((Nn,), np.int32)
    rc = _lib.b2t_prefix_beam_search(a.ctypes.data, ln.ctypes.data, Nn, T, Cc, int(blank), int(first_beam_size), int(second_beam_size), ml,
                                     ids.ctypes.data, ol.ctypes.data, sc.ctypes.data, vt.ctypes.data, tm.ctypes.data, nh.ctypes.data)
    if rc < 0:
        raise N.B2TError(f"ctc_prefix_beam_search failed ({rc}): {_lib.b2t_prefix_last_error().decode()}")
    return [[(ids[n, r, :ol[n, r]].tolist(), float(sc[n, r]), float(vt[n, r]), tm[n, r, :ol[n, r]].tolist()) for r in range(nh[n])]
            for n in range(Nn)]
