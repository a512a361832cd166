"""ctypes access to the C++ decoder oracle (oracle/_build/libdecoder_oracle.so) for the tests."""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_lib = None


def oracle_lib():
    global _lib
    if _lib is None:
        subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle")])
        lib = C.CDLL(os.path.join(ROOT, "oracle", "_build", "libdecoder_oracle.so"))
        lib.orc_create.restype = C.c_void_p
        lib.orc_create.argtypes = [C.c_char_p, C.c_char_p, C.c_int, C.c_int, C.c_float, C.c_float, C.c_float, C.c_float, C.c_float, C.c_int,
                                   C.c_char_p, C.c_int]
        lib.orc_prefix_create.restype = C.c_void_p
        lib.orc_prefix_create.argtypes = [C.c_int, C.c_int, C.c_int]
        for n in ("orc_destroy", "orc_reset", "orc_finish", "orc_prefix_destroy", "orc_prefix_reset"):
            getattr(lib, n).argtypes = [C.c_void_p]
            getattr(lib, n).restype = None
        lib.orc_decode_logprobs.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int]
        lib.orc_decode_logits.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_float]
        lib.orc_num_results.argtypes = [C.c_void_p]
        lib.orc_get_result.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_float), C.POINTER(C.c_float), C.c_char_p, C.c_int]
        lib.orc_tokens_per_frame.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        lib.orc_graph_info.argtypes = [C.c_void_p, C.POINTER(C.c_longlong), C.POINTER(C.c_longlong)]
        lib.orc_prefix_search.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int]
        lib.orc_prefix_num.argtypes = [C.c_void_p]
        lib.orc_prefix_get.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.POINTER(C.c_float), C.POINTER(C.c_float), C.c_void_p, C.c_int]
        _lib = lib
    return _lib


class OracleDecoder:
    def __init__(self, fst, words, max_active=7000, min_active=200, beam=17.0, lattice_beam=8.0, acoustic_scale=0.325,
                 blank_skip=1.0, length_penalty=0.0, nbest=100):
        self.lib = oracle_lib()
        err = C.create_string_buffer(256)
        self.h = self.lib.orc_create(fst.encode(), words.encode(), max_active, min_active, beam, lattice_beam, acoustic_scale, blank_skip,
                                     length_penalty, nbest, err, 256)
        if not self.h:
            raise RuntimeError(err.value.decode())

    def __del__(self):
        if getattr(self, "h", None):
            self.lib.orc_destroy(self.h)
            self.h = None

    def reset(self):
        self.lib.orc_reset(self.h)

    def decode_logprobs(self, lp):
        lp = np.ascontiguousarray(lp, dtype=np.float32)
        self.lib.orc_decode_logprobs(self.h, lp.ctypes.data, lp.shape[0], lp.shape[1])

    def decode_logits(self, logits, log_priors=None, blank_penalty=0.0):
        x = np.ascontiguousarray(logits, dtype=np.float32)
        pr = None if log_priors is None else np.ascontiguousarray(log_priors, dtype=np.float32)
        self.lib.orc_decode_logits(self.h, x.ctypes.data, None if pr is None else pr.ctypes.data, x.shape[0], x.shape[1], blank_penalty)

    def finish(self):
        self.lib.orc_finish(self.h)

    def results(self):
        out = []
        buf = C.create_string_buffer(1 << 16)
        ac, lm = C.c_float(), C.c_float()
        for i in range(self.lib.orc_num_results(self.h)):
            self.lib.orc_get_result(self.h, i, C.byref(ac), C.byref(lm), buf, 1 << 16)
            out.append((ac.value, lm.value, buf.value.decode()))
        return out

    def tokens_per_frame(self):
        a = np.zeros(4096, dtype=np.int32)
        n = self.lib.orc_tokens_per_frame(self.h, a.ctypes.data, 4096)
        return a[:n]


def prefix_search(logp, first_beam=10, second_beam=10, blank=0):
    lib = oracle_lib()
    h = lib.orc_prefix_create(blank, first_beam, second_beam)
    lp = np.ascontiguousarray(logp, dtype=np.float32)
    lib.orc_prefix_search(h, lp.ctypes.data, lp.shape[0], lp.shape[1])
    out = []
    ids = np.zeros(4096, dtype=np.int32); times = np.zeros(4096, dtype=np.int32)
    sc, vt = C.c_float(), C.c_float()
    for i in range(lib.orc_prefix_num(h)):
        n = lib.orc_prefix_get(h, i, ids.ctypes.data, 4096, C.byref(sc), C.byref(vt), times.ctypes.data, 4096)
        out.append((ids[:n].tolist(), sc.value, vt.value, times[:n].tolist()))
    lib.orc_prefix_destroy(h)
    return out
