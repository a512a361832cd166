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
        lib.orc_token_list.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        lib.orc_graph_info.argtypes = [C.c_void_p, C.POINTER(C.c_longlong), C.POINTER(C.c_longlong)]
        lib.orc_prefix_search.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int]
        lib.orc_rescore.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p, C.c_char_p, C.c_int]
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

    def rescore(self, lm_fst, rescore_lm_fst):
        err = C.create_string_buffer(256)
        if self.lib.orc_rescore(self.h, lm_fst.encode(), rescore_lm_fst.encode(), err, 256) != 0:
            raise RuntimeError("rescore: " + err.value.decode())

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

    def token_list(self, cap=1 << 20):
        """(states, costs) of the current token list, in Kaldi's list order."""
        st = np.zeros(cap, dtype=np.int32); co = np.zeros(cap, dtype=np.float32)
        n = self.lib.orc_token_list(self.h, st.ctypes.data, co.ctypes.data, cap)
        return st[:n], co[:n]


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


# ---------------------------------------------------------------------------------------------------------------
# The reference's shipped 1-gram graph (language_model/pretrained_language_models/openwebtext_1gram_lm_sil).  It is reference
# DATA, not source: __graft_entry__.build() stages it under oracle/_ref/ (git-ignored, travels to the GPU box) when
# /root/reference is present; nothing here reads /root/reference at test time.
REAL_GRAPH_DIR = os.path.join(ROOT, "oracle", "_ref", "openwebtext_1gram_lm_sil")


def real_graph():
    fst, words = os.path.join(REAL_GRAPH_DIR, "TLG.fst"), os.path.join(REAL_GRAPH_DIR, "words.txt")
    return (fst, words) if os.path.exists(fst) and os.path.exists(words) else None


def read_fst(path):
    """OpenFST binary 'vector' FST with 'standard' arcs -> (start, finals[ns], offsets[ns + 1], arcs[na] structured array)."""
    import struct
    buf = open(path, "rb").read()
    pos = 0

    def take(fmt):
        nonlocal pos
        v = struct.unpack_from(fmt, buf, pos)
        pos += struct.calcsize(fmt)
        return v if len(v) > 1 else v[0]

    def take_str():
        nonlocal pos
        n = take("<i")
        s = buf[pos:pos + n].decode()
        pos += n
        return s

    assert take("<i") == 2125659606, "not an OpenFST binary file"
    ft, at = take_str(), take_str()
    assert ft == "vector" and at == "standard", (ft, at)
    _version, flags = take("<i"), take("<i")
    _props, start, ns, na = take("<Q"), take("<q"), take("<q"), take("<q")
    for bit in (1, 2):                      # embedded symbol tables
        if flags & bit:
            take("<i"); take_str(); take("<q")
            size = take("<q")
            for _ in range(size):
                take_str(); take("<q")
    arc_dt = np.dtype([("il", "<i4"), ("ol", "<i4"), ("w", "<f4"), ("next", "<i4")])
    finals = np.empty(ns, np.float32); off = np.zeros(ns + 1, np.int64)
    chunks = []
    for s in range(ns):
        fw, cnt = struct.unpack_from("<fq", buf, pos)
        pos += 12
        finals[s] = fw; off[s + 1] = off[s] + cnt
        chunks.append(np.frombuffer(buf, arc_dt, cnt, pos))
        pos += 16 * cnt
    arcs = np.concatenate(chunks) if chunks else np.empty(0, arc_dt)
    assert len(arcs) == na or na <= 0
    return int(start), finals, off, arcs


def random_walk_utterance(graph, rng, n_words=3, T=95, C=41, peak=7.0, noise=0.8, max_arcs=400):
    """a seeded random walk from the start state until `n_words` words were emitted and a final state
    is reached; the visited input labels (1 = blank, 2 = SIL, 3.. = phones; 0 = epsilon) are rendered as peaky posteriors.
    Returns (logits [T, C], word ids) or None when the walk does not fit into T frames."""
    start, finals, off, arcs = graph
    s, labels, words = start, [], []
    for _ in range(max_arcs):
        if len(words) >= n_words and np.isfinite(finals[s]):
            break
        a = arcs[off[s]:off[s + 1]]
        if len(a) == 0:
            return None
        cand = a[a["il"] != 1] if np.any(a["il"] != 1) else a          # do not idle on blank self loops
        if len(words) >= n_words:                                       # head for a final state: prefer non-word arcs
            c2 = cand[cand["ol"] == 0]
            cand = c2 if len(c2) else cand
        k = cand[rng.randint(len(cand))]
        if k["il"] > 1:
            labels.append(int(k["il"]) - 1)                             # graph ilabel -> logit column
        if k["ol"] != 0:
            words.append(int(k["ol"]))
        s = int(k["next"])
    else:
        return None
    x = noise * rng.randn(T, C).astype(np.float32)
    x[:, 0] += 2.0
    t, prev = int(rng.randint(1, 3)), None
    for cls in labels:
        if prev == cls:
            t += 1                                                      # CTC needs a blank between repeated labels
        d = int(rng.randint(2, 4))
        if t + d >= T - 1:
            return None
        x[t:t + d, cls] += peak
        t += d + int(rng.randint(0, 2))
        prev = cls
    return x, words


def cmp_strict(dec_results, ref_results, acoustic_scale, tol=1e-3):
    """Integer work: 1-best, the n-best set and the scores.  Only hypotheses that tie (total score) with the LAST kept entry
    may differ: which of several equal-score homophones makes the cut is the n-shortest-paths heap order in the reference
    (fst::ShortestPath) and creation order here and in the oracle."""
    assert len(dec_results) == len(ref_results), (len(dec_results), len(ref_results))
    assert dec_results[0].sentence == ref_results[0][2]
    ours = {r.sentence: (r.ac_score, r.lm_score) for r in dec_results}
    ref = {r[2]: (r[0], r[1]) for r in ref_results}
    total = lambda v: v[1] + acoustic_scale * v[0]
    worst = min(total(v) for v in ref.values())
    cut = lambda d: {k for k, v in d.items() if total(v) > worst + tol * max(1.0, abs(worst))}
    assert cut(ours) == cut(ref), (sorted(cut(ours) - cut(ref)), sorted(cut(ref) - cut(ours)))
    assert all(abs(total(v) - worst) <= 2 * tol * max(1.0, abs(worst)) for k, v in ours.items() if k not in ref), "a non-tied hypothesis differs"
    for k in set(ours) & set(ref):
        assert abs(ours[k][0] - ref[k][0]) < tol * max(1.0, abs(ours[k][0])), k
        assert abs(ours[k][1] - ref[k][1]) < tol * max(1.0, abs(ours[k][1])), k
