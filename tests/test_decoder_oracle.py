"""CPU tests pinning the decoder oracle: the reference's golden 3x3 prefix-search vector, graph I/O,
and invariants of the WFST search (which the reference itself never tests: parity unpinned)."""
import math
import os
import sys

import numpy as np
import pytest

import decoder_util as D

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import make_toy_tlg as TLG  # noqa: E402


def test_prefix_beam_golden_3x3():
    """ctc_prefix_beam_search_test.cc:18-59 (first_beam = second_beam = 3)."""
    data = np.log(np.array([0.25, 0.40, 0.35, 0.40, 0.35, 0.25, 0.10, 0.50, 0.40], dtype=np.float32).reshape(3, 3))
    res = D.prefix_search(data, first_beam=3, second_beam=3)
    assert [r[0] for r in res] == [[2, 1], [1, 2], [1]]
    for got, want in zip([math.exp(r[1]) for r in res], [0.2185, 0.1550, 0.1525]):
        assert abs(got - want) < 1e-6 * want * 4                       # EXPECT_FLOAT_EQ (4 ulp)
    for got, want in zip([math.exp(r[2]) for r in res], [0.07, 0.064, 0.07]):
        assert abs(got - want) < 1e-6 * want * 4
    assert [r[3] for r in res] == [[0, 2], [0, 2], [2]]


@pytest.fixture(scope="module")
def toy(tmp_path_factory):
    d = tmp_path_factory.mktemp("tlg")
    fst, words = str(d / "TLG.fst"), str(d / "words.txt")
    info = TLG.build(fst, words, n_words=60, seed=1)
    return fst, words, info


def test_wfst_decodes_rendered_sentence(toy):
    fst, words, info = toy
    seq = [3, 17, 42, 8]
    logits = TLG.render_logits([info["prons"][w] for w in seq], T=90, seed=3)
    dec = D.OracleDecoder(fst, words, nbest=20, acoustic_scale=0.6)
    dec.decode_logits(logits, None, 0.0)
    partial = dec.results()
    assert len(partial) == 1                                            # 1-best after every Decode() chunk
    dec.finish()
    res = dec.results()
    want = " ".join(info["words"][w].lower() for w in seq)
    assert res[0][2] == want
    assert partial[0][2] == want or len(partial[0][2]) > 0
    tot = [-(a * 0.6) - l for a, l, _ in res]                           # total cost = graph + scaled acoustic
    assert all(tot[i] <= tot[i + 1] + 1e-4 for i in range(len(tot) - 1))  # best first
    assert len({s for _, _, s in res}) == len(res)                      # distinct word sequences
    assert tot[-1] - tot[0] <= 8.0 + 1e-3                               # within lattice_beam
    # nbest == 1 (back-pointer best path with final costs) agrees with the head of the n-best list
    d1 = D.OracleDecoder(fst, words, nbest=1, acoustic_scale=0.6)
    d1.decode_logits(logits, None, 0.0)
    d1.finish()
    r1 = d1.results()
    assert r1[0][2] == res[0][2] and abs(r1[0][0] - res[0][0]) < 1e-3 and abs(r1[0][1] - res[0][1]) < 1e-3


def test_wfst_chunked_equals_whole(toy):
    fst, words, info = toy
    logits = TLG.render_logits([info["prons"][w] for w in [5, 6, 7]], T=70, seed=5)
    a = D.OracleDecoder(fst, words, nbest=10, acoustic_scale=0.6)
    a.decode_logits(logits, None, 1.0)
    a.finish()
    b = D.OracleDecoder(fst, words, nbest=10, acoustic_scale=0.6)
    for i in range(0, 70, 16):
        b.decode_logits(logits[i:i + 16], None, 1.0)
    b.finish()
    assert a.results() == b.results()
    b.reset()
    b.decode_logits(logits, None, 1.0)
    b.finish()
    assert a.results() == b.results()                                   # Reset() restores a clean decoder


def test_blank_skipping_changes_frames_not_result(toy):
    fst, words, info = toy
    logits = TLG.render_logits([info["prons"][w] for w in [11, 12]], T=80, seed=7, peak=9.0, noise=0.3)
    blank_frames = logits.argmax(1) == 0
    logits[blank_frames, 0] += 10.0                                     # confident blanks: exp(logp[0]) > 0.9 on those frames
    full = D.OracleDecoder(fst, words, nbest=1, acoustic_scale=0.6, blank_skip=1.0)
    full.decode_logits(logits, None, 0.0); full.finish()
    skip = D.OracleDecoder(fst, words, nbest=1, acoustic_scale=0.6, blank_skip=0.9)
    skip.decode_logits(logits, None, 0.0); skip.finish()
    assert len(skip.tokens_per_frame()) < len(full.tokens_per_frame()) == 80
    assert skip.results()[0][2] == full.results()[0][2]


@pytest.mark.skipif(D.real_graph() is None, reason="the shipped 1-gram graph is staged under oracle/_ref/ by __graft_entry__.build() in the build container")
def test_reads_shipped_1gram_graph():
    base = D.REAL_GRAPH_DIR
    dec = D.OracleDecoder(base + "/TLG.fst", base + "/words.txt", nbest=5, max_active=2000)
    import ctypes as C
    ns, na = C.c_longlong(), C.c_longlong()
    start = dec.lib.orc_graph_info(dec.h, C.byref(ns), C.byref(na))
    assert (ns.value, na.value, start) == (179946, 704714, 0)           # SURVEY.md section 2 fixture facts
    rng = np.random.RandomState(0)
    x = rng.randn(40, 41).astype(np.float32)
    x[:, 0] += 3
    for t, c in [(5, 9), (6, 9), (9, 22), (12, 30), (13, 30), (16, 1), (20, 15), (24, 33)]:
        x[t, c] += 9
    dec.decode_logits(x, np.zeros_like(x), math.log(90.0))
    dec.finish()
    res = dec.results()
    assert len(res) >= 1 and all(isinstance(s, str) for _, _, s in res) and res[0][2] == res[0][2].lower()


@pytest.mark.skipif(D.real_graph() is None, reason="the shipped 1-gram graph is staged under oracle/_ref/ by __graft_entry__.build() in the build container")
def test_python_reader_and_random_walks_on_shipped_graph():
    """The Python reader of the OpenFST file agrees with the C++ one, and seeded random walks give
    in-vocabulary utterances the oracle decodes back when the beam is wide and the posteriors are clean."""
    fst, words = D.real_graph()
    g = D.read_fst(fst)
    assert (g[0], len(g[1]), len(g[3])) == (0, 179946, 704714)
    assert int(g[3]["il"].max()) == 41 and int((g[3]["il"] == 0).sum()) >= 0
    vocab = {}
    for line in open(words):
        w, i = line.split()
        vocab[int(i)] = w
    rng = np.random.RandomState(1)
    hits = total = 0
    dec = D.OracleDecoder(fst, words, 3000, 200, 17.0, 8.0, 0.6, 1.0, 0.0, 1)
    for _ in range(4):
        u = None
        while u is None:
            u = D.random_walk_utterance(g, rng, n_words=2, peak=10.0, noise=0.3)
        x, wd = u
        dec.reset(); dec.decode_logits(x, np.zeros_like(x), math.log(2.0)); dec.finish()
        got = dec.results()[0][2].split()
        want = [vocab[w].lower() for w in wd]
        total += len(want)
        hits += sum(1 for a, b in zip(got, want) if a == b)
    assert hits >= total // 2, (hits, total)        # homophones / alternative segmentations are legitimate misses
