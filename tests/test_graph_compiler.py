"""ARPA + lexicon -> TLG compiler (nejm-brain-to-text_b200/graph_compiler.py) checked with the decoder oracle: the best path
through the compiled graph for a cleanly rendered word sequence is that sequence, and its graph cost equals the back-off
n-gram cost of the sentence computed by an independent scorer plus the optional-silence costs."""
import importlib.util
import math
import os
import sys

import numpy as np
import pytest

import decoder_util as D

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import make_toy_tlg as TLG  # noqa: E402

spec = importlib.util.spec_from_file_location("graph_compiler", os.path.join(ROOT, "nejm-brain-to-text_b200", "graph_compiler.py"))
GC = importlib.util.module_from_spec(spec)
spec.loader.exec_module(GC)

ARPA = """\\data\\
ngram 1=7
ngram 2=8
ngram 3=4

\\1-grams:
-1.2 </s>
-99 <s> -0.4
-0.9 alpha -0.3
-0.8 beta -0.25
-1.1 gamma -0.2
-1.3 delta
-1.0 beater -0.1

\\2-grams:
-0.5 <s> alpha -0.2
-0.7 <s> beta
-0.4 alpha beta -0.15
-0.9 alpha gamma
-0.6 beta gamma -0.1
-0.3 gamma </s>
-0.8 beta </s>
-0.5 beater alpha

\\3-grams:
-0.2 <s> alpha beta
-0.35 alpha beta gamma
-0.15 beta gamma </s>
-0.6 alpha beta </s>

\\end\\
"""
PHONES = [f"P{i}" for i in range(39)]
LEXICON = {"alpha": ["P0 P1 P2"], "beta": ["P3 P4"], "gamma": ["P5 P6 P7 P8"], "delta": ["P9 P9 P10"], "beater": ["P3 P4 P11"]}


@pytest.fixture(scope="module")
def compiled(tmp_path_factory):
    d = tmp_path_factory.mktemp("gc")
    arpa, lex = str(d / "lm.arpa"), str(d / "lexicon.txt")
    open(arpa, "w").write(ARPA)
    with open(lex, "w") as f:
        for w, prons in LEXICON.items():
            for pr in prons:
                f.write(f"{w} {pr}\n")
    fst, words = str(d / "TLG.fst"), str(d / "words.txt")
    info = GC.compile_to_files(arpa, lex, PHONES, fst, words)
    order, grams = GC.parse_arpa(arpa)
    return fst, words, info, GC.NgramLM(order, grams)


def test_scorer_backoff():
    order, grams = 3, {("a",): (-1.0, -0.5), ("b",): (-2.0, 0.0), ("a", "b"): (-0.3, 0.0)}
    lm = GC.NgramLM(order, grams)
    assert abs(lm.cost(("a",), "b") - 0.3 * GC.LN10) < 1e-9
    assert abs(lm.cost(("b",), "a") - (1.0 + 0.0) * GC.LN10) < 1e-9          # back-off weight of "b" is 0
    assert abs(lm.cost(("x", "a"), "a") - (0.5 + 1.0) * GC.LN10) < 1e-9      # bow(a) + P(a)


@pytest.mark.parametrize("sentence", [["alpha", "beta", "gamma"], ["beta", "gamma"], ["alpha", "gamma"], ["delta", "alpha", "beta"],
                                      ["beater", "alpha", "beta"], ["gamma", "delta", "delta"]])
def test_best_path_cost_equals_ngram_cost(compiled, sentence):
    fst, words, info, lm = compiled
    assert info["order"] == 3 and info["n_words"] == 5
    ids = {p: 3 + i for i, p in enumerate(PHONES)}
    prons = [[ids[p] for p in LEXICON[w][0].split()] for w in sentence]
    logits = TLG.render_logits(prons, T=90, seed=1, peak=12.0, noise=0.1)
    dec = D.OracleDecoder(fst, words, 7000, 200, 30.0, 10.0, 1.0, 1.0, 0.0, 5)
    dec.decode_logits(logits, np.zeros_like(logits), 0.0)
    dec.finish()
    res = dec.results()
    assert res and res[0][2].split() == sentence, res[:2]
    want = lm.sentence_cost(sentence) + (len(sentence) + 1) * math.log(2.0)      # optional silence: ln 2 at the start and after each word
    assert abs(-res[0][1] - want) < 1e-3 * max(1.0, want), (-res[0][1], want)


@pytest.mark.gpu
def test_compiled_graph_gpu_vs_oracle(compiled, pkg):
    import b2t_pkg
    LM = b2t_pkg.submodule("lm_decoder")
    fst, words, info, lm = compiled
    ids = {p: 3 + i for i, p in enumerate(PHONES)}
    opts = (7000, 200, 17.0, 8.0, 0.6, 1.0, 0.0, 20)
    dec = LM.BrainSpeechDecoder(LM.DecodeResource(fst, "", "", words, ""), LM.DecodeOptions(*opts), max_frames=128)
    for k, sentence in enumerate((["alpha", "beta", "gamma"], ["beater", "alpha"], ["delta", "gamma"])):
        prons = [[ids[p] for p in LEXICON[w][0].split()] for w in sentence]
        logits = TLG.render_logits(prons, T=90, seed=3 + k, peak=7.0, noise=1.0)
        ref = D.OracleDecoder(fst, words, *opts)
        ref.decode_logits(logits, np.zeros_like(logits), math.log(3.0)); ref.finish()
        dec.Reset()
        LM.DecodeNumpy(dec, logits, np.zeros_like(logits), math.log(3.0))
        dec.FinishDecoding()
        ours, r = dec.result(), ref.results()
        assert [x.sentence for x in ours][:1] == [x[2] for x in r][:1]
        assert {x.sentence for x in ours} == {x[2] for x in r}
        byref = {x[2]: x for x in r}
        assert all(abs(x.lm_score - byref[x.sentence][1]) < 1e-3 * max(1.0, abs(x.lm_score)) for x in ours)


def test_unigram_lm_pron_probs_and_missing_words(tmp_path):
    """A 1-gram LM (the shipped openwebtext graph is one), lexiconp-style pronunciation probabilities, a second pronunciation,
    and an LM word without a lexicon entry (dropped from the graph)."""
    arpa = tmp_path / "uni.arpa"
    arpa.write_text("\\data\\\nngram 1=5\n\n\\1-grams:\n-0.7 </s>\n-99 <s>\n-0.5 red\n-0.9 green\n-1.5 nolex\n\n\\end\\\n")
    lex = tmp_path / "lexiconp.txt"
    lex.write_text("red 0.8 P0 P1\nred 0.2 P2 P1 P3\ngreen 1.0 P4 P5 P6\n")
    fst, words = str(tmp_path / "TLG.fst"), str(tmp_path / "words.txt")
    info = GC.compile_to_files(str(arpa), str(lex), PHONES, fst, words)
    assert info["order"] == 1 and info["n_words"] == 2
    assert "nolex" not in open(words).read()
    lm = GC.NgramLM(*GC.parse_arpa(str(arpa)))
    ids = {p: 3 + i for i, p in enumerate(PHONES)}
    for sentence, prons, pron_cost in ((["red", "green"], [["P0", "P1"], ["P4", "P5", "P6"]], -math.log(0.8)),
                                       (["green", "red"], [["P4", "P5", "P6"], ["P2", "P1", "P3"]], -math.log(0.2))):
        logits = TLG.render_logits([[ids[p] for p in pr] for pr in prons], T=70, seed=2, peak=12.0, noise=0.1)
        dec = D.OracleDecoder(fst, words, 7000, 200, 30.0, 10.0, 1.0, 1.0, 0.0, 3)
        dec.decode_logits(logits, np.zeros_like(logits), 0.0)
        dec.finish()
        res = dec.results()
        assert res[0][2].split() == sentence
        want = lm.sentence_cost(sentence) + 3 * math.log(2.0) + pron_cost
        assert abs(-res[0][1] - want) < 1e-3 * want, (-res[0][1], want)


# ---------------------------------------------------------------------------------------------------------------------
# Rescore() restatement of the oracle (brain_speech_decoder.cc:47-101): groundwork for the next row N1.
ARPA_NEW = ARPA.replace("-0.4 alpha beta -0.15", "-1.4 alpha beta -0.15").replace("-0.9 alpha gamma", "-0.1 alpha gamma") \
               .replace("-0.35 alpha beta gamma", "-1.9 alpha beta gamma").replace("-1.3 delta", "-0.6 delta")


def _fst_min_cost(order, grams, words):
    """Independent checker: cheapest path of `words` through the back-off acceptor, epsilon (back-off) arcs free to take at any
    time, final cost included -- what composition + determinisation of a one-path lattice with G yields."""
    start, arcs, backoff, final = GC.build_g(order, grams)

    def closure(st):
        work = list(st)
        while work:
            h = work.pop()
            if h in backoff:
                c, hb = backoff[h]
                if st[h] + c < st.get(hb, math.inf):
                    st[hb] = st[h] + c
                    work.append(hb)
        return st

    cur = closure({start: 0.0})
    for w in words:
        nxt = {}
        for h, c in cur.items():
            for ww, cost, hn in arcs.get(h, ()):
                if ww == w and c + cost < nxt.get(hn, math.inf):
                    nxt[hn] = c + cost
        if not nxt:
            return math.inf
        cur = closure(nxt)
    return min((c + final[h] for h, c in cur.items() if h in final), default=math.inf)


def test_rescore_restatement(compiled, tmp_path):
    fst, words, info, lm = compiled
    order, grams_old = GC.parse_arpa(os.path.join(os.path.dirname(fst), "lm.arpa"))
    arpa_new = tmp_path / "new.arpa"
    arpa_new.write_text(ARPA_NEW)
    _, grams_new = GC.parse_arpa(str(arpa_new))
    word_ids = {}
    for line in open(words):
        w, i = line.split()
        if int(i) > 0:
            word_ids[w] = int(i)
    g_old, g_new = str(tmp_path / "G.fst"), str(tmp_path / "G_new.fst")
    GC.write_g_fst(order, grams_old, word_ids, g_old)
    GC.write_g_fst(order, grams_new, word_ids, g_new)
    ids = {p: 3 + i for i, p in enumerate(PHONES)}
    sentence = ["alpha", "beta", "gamma"]
    logits = TLG.render_logits([[ids[p] for p in LEXICON[w][0].split()] for w in sentence], T=90, seed=5, peak=5.0, noise=1.2)
    dec = D.OracleDecoder(fst, words, 7000, 200, 20.0, 8.0, 0.5, 1.0, 0.0, 50)
    dec.decode_logits(logits, np.zeros_like(logits), 0.0)
    dec.finish()
    first = dec.results()
    assert len(first) >= 5 and first[0][2].split() == sentence
    # (1) rescoring with the LM the graph was built from changes nothing but float noise
    dec.rescore(g_old, g_old)
    same = dec.results()
    assert [r[2] for r in same] == [r[2] for r in first]
    assert all(abs(a[1] - b[1]) < 1e-4 * max(1.0, abs(b[1])) and a[0] == b[0] for a, b in zip(same, first))
    # (2) new LM: graph' = g - c_old(W) + c_new(W) for every first-pass W, re-ranked by graph' + acoustic
    dec2 = D.OracleDecoder(fst, words, 7000, 200, 20.0, 8.0, 0.5, 1.0, 0.0, 10 ** 6)      # all distinct sequences within the lattice beam
    dec2.decode_logits(logits, np.zeros_like(logits), 0.0)
    dec2.finish()
    every = dec2.results()
    expect = []
    for ac, lmv, sent in every:
        ws = sent.split()
        c_old, c_new = _fst_min_cost(order, grams_old, ws), _fst_min_cost(order, grams_new, ws)
        if math.isfinite(c_old) and math.isfinite(c_new):
            g = -lmv - c_old + c_new
            expect.append((g + (-ac * 0.5), g, ac, sent))                     # total cost with the acoustic cost unscaled (ac = -a / acoustic_scale)
    expect.sort(key=lambda e: (e[0], e[1]))
    dec = D.OracleDecoder(fst, words, 7000, 200, 20.0, 8.0, 0.5, 1.0, 0.0, 50)    # a fresh first pass (Rescore follows FinishDecoding once)
    dec.decode_logits(logits, np.zeros_like(logits), 0.0)
    dec.finish()
    n_first = len(dec.results())
    dec.rescore(g_old, g_new)
    got = dec.results()
    assert len(got) == min(n_first, len(expect))
    assert [r[2] for r in got] == [e[3] for e in expect[:len(got)]]
    for r, e in zip(got, expect):
        assert abs(-r[1] - e[1]) < 1e-3 * max(1.0, abs(e[1])) and abs(r[0] - e[2]) < 1e-4 * max(1.0, abs(e[2]))
    assert [r[2] for r in got] != [r[2] for r in first[:len(got)]]            # the new LM really re-ranks this example


def test_product_rescore_core_matches_oracle(compiled, tmp_path, pkg):
    """b2t_lm_rescore_sequences (host core of Rescore() in the product library, no GPU involved) against the oracle's Rescore()
    on the same first-pass sequences and the same LM acceptors."""
    import ctypes as C
    fst, words, info, lm = compiled
    order, grams_old = GC.parse_arpa(os.path.join(os.path.dirname(fst), "lm.arpa"))
    arpa_new = tmp_path / "new.arpa"
    arpa_new.write_text(ARPA_NEW)
    _, grams_new = GC.parse_arpa(str(arpa_new))
    word_ids = {}
    for line in open(words):
        w, i = line.split()
        if int(i) > 0:
            word_ids[w.lower()] = int(i)
    g_old, g_new = str(tmp_path / "G.fst"), str(tmp_path / "G_new.fst")
    GC.write_g_fst(order, grams_old, {w: i for w, i in word_ids.items()}, g_old)
    GC.write_g_fst(order, grams_new, {w: i for w, i in word_ids.items()}, g_new)
    ids = {p: 3 + i for i, p in enumerate(PHONES)}
    sentence = ["beater", "alpha", "beta"]
    logits = TLG.render_logits([[ids[p] for p in LEXICON[w][0].split()] for w in sentence], T=90, seed=8, peak=5.0, noise=1.2)
    scale = 0.5

    def first_pass(nbest):
        d = D.OracleDecoder(fst, words, 7000, 200, 20.0, 8.0, scale, 1.0, 0.0, nbest)
        d.decode_logits(logits, np.zeros_like(logits), 0.0)
        d.finish()
        return d

    every = first_pass(10 ** 6).results()                              # all distinct sequences within the lattice beam
    ref = first_pass(30)
    keep = len(ref.results())
    ref.rescore(g_old, g_new)
    want = ref.results()
    seqs = [[word_ids[w] for w in s.split()] for _, _, s in every]
    flat = np.array([i for s in seqs for i in s], dtype=np.int32)
    lens = np.array([len(s) for s in seqs], dtype=np.int32)
    graph = np.array([-lmv for _, lmv, _ in every], dtype=np.float32)
    acoustic = np.array([-ac * scale for ac, _, _ in every], dtype=np.float32)
    order_out = np.zeros(keep, dtype=np.int32); graph_out = np.zeros(keep, dtype=np.float32)
    lib = C.CDLL(os.path.join(ROOT, "nejm-brain-to-text_b200", "libb2t_b200.so"))
    lib.b2t_lm_rescore_sequences.argtypes = [C.c_char_p, C.c_char_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
    m = lib.b2t_lm_rescore_sequences(g_old.encode(), g_new.encode(), len(seqs), flat.ctypes.data, lens.ctypes.data, graph.ctypes.data,
                                     acoustic.ctypes.data, keep, order_out.ctypes.data, graph_out.ctypes.data)
    assert m == len(want) > 5
    assert [every[order_out[i]][2] for i in range(m)] == [r[2] for r in want]
    assert all(abs(-graph_out[i] - want[i][1]) < 1e-4 * max(1.0, abs(want[i][1])) for i in range(m))
    assert lib.b2t_lm_rescore_sequences(b"/nonexistent.fst", g_new.encode(), 0, None, None, None, None, 1, order_out.ctypes.data, graph_out.ctypes.data) < 0


@pytest.mark.gpu
@pytest.mark.parametrize("nbest", [30, 1])
def test_decoder_rescore_gpu_vs_oracle(compiled, tmp_path, pkg, nbest):
    """BrainSpeechDecoder.Rescore() end to end (brain_speech_decoder.cc:61-101): DecodeResource carries the LM the graph was built from
    and the rescoring LM; after FinishDecoding the GPU decoder's pruned lattice is re-scored and the list re-ranked.  Checker: the
    oracle's Facade::Rescore on the same posteriors and acceptors."""
    import b2t_pkg
    LM = b2t_pkg.submodule("lm_decoder")
    fst, words, info, lm = compiled
    order, grams_old = GC.parse_arpa(os.path.join(os.path.dirname(fst), "lm.arpa"))
    arpa_new = tmp_path / "new.arpa"
    arpa_new.write_text(ARPA_NEW)
    _, grams_new = GC.parse_arpa(str(arpa_new))
    word_ids = {}
    for line in open(words):
        w, i = line.split()
        if int(i) > 0:
            word_ids[w] = int(i)
    g_old, g_new = str(tmp_path / "G.fst"), str(tmp_path / "G_new.fst")
    GC.write_g_fst(order, grams_old, word_ids, g_old)
    GC.write_g_fst(order, grams_new, word_ids, g_new)
    ids = {p: 3 + i for i, p in enumerate(PHONES)}
    opts = (7000, 200, 20.0, 8.0, 0.5, 1.0, 0.0, nbest)
    for seed, sentence in ((5, ["alpha", "beta", "gamma"]), (8, ["beater", "alpha", "beta"])):
        logits = TLG.render_logits([[ids[p] for p in LEXICON[w][0].split()] for w in sentence], T=90, seed=seed, peak=5.0, noise=1.2)
        ref = D.OracleDecoder(fst, words, *opts)
        ref.decode_logits(logits, np.zeros_like(logits), 0.0)
        ref.finish()
        first = ref.results()
        ref.rescore(g_old, g_new)
        want = ref.results()
        dec = LM.BrainSpeechDecoder(LM.DecodeResource(fst, g_old, g_new, words, ""), LM.DecodeOptions(*opts), max_frames=128)
        dec.Reset()
        LM.DecodeNumpy(dec, logits, np.zeros_like(logits), 0.0)
        dec.FinishDecoding()
        assert [r.sentence for r in dec.result()] == [r[2] for r in first]
        dec.Rescore()
        got = dec.result()
        assert len(got) == len(want) >= 1
        assert [r.sentence for r in got] == [r[2] for r in want]
        for a, b in zip(got, want):
            assert abs(a.lm_score - b[1]) < 1e-3 * max(1.0, abs(b[1])) and abs(a.ac_score - b[0]) < 1e-3 * max(1.0, abs(b[0]))
        # a second utterance through the same decoder object
        dec.Reset()
    # rescoring with the LM the graph was built from leaves the first pass as it was
    dec = LM.BrainSpeechDecoder(LM.DecodeResource(fst, g_old, g_old, words, ""), LM.DecodeOptions(*opts), max_frames=128)
    LM.DecodeNumpy(dec, logits, np.zeros_like(logits), 0.0)
    dec.FinishDecoding()
    before = [(r.sentence, r.lm_score) for r in dec.result()]
    dec.Rescore()
    after = [(r.sentence, r.lm_score) for r in dec.result()]
    assert [s for s, _ in before] == [s for s, _ in after] and all(abs(a[1] - b[1]) < 1e-3 * max(1.0, abs(b[1])) for a, b in zip(after, before))


@pytest.mark.gpu
@pytest.mark.parametrize("order", [3, 5])
def test_strict_order_on_compiled_ngram_graphs(order, tmp_path, pkg):
    """Graphs compiled by graph_compiler from a synthetic 3-/5-gram ARPA (back-off arcs = input epsilons, so the serial
    epsilon-closure replay is exercised), max_active binding: strict mode == oracle (per-frame token counts, 1-best, n-best)."""
    import b2t_pkg
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import make_synth_lm as SL
    LM = b2t_pkg.submodule("lm_decoder")
    li = SL.build(str(tmp_path), order=order, n_words=300, n_sent=4000, seed=3)
    fst, words = str(tmp_path / "TLG.fst"), str(tmp_path / "words.txt")
    GC.compile_to_files(li["arpa"], li["lexicon"], li["phones"], fst, words)
    widx = {w: i for i, w in enumerate(li["words"])}
    sents = [s[1:-1] for s in li["corpus"] if 2 <= len(s) - 2 <= 4][:3]
    for max_active in (100, 500):
        opts = (max_active, 50, 17.0, 8.0, 0.5, 1.0, 0.0, 30)
        dec = LM.BrainSpeechDecoder(LM.DecodeResource(fst, "", "", words, ""), LM.DecodeOptions(*opts), max_frames=128, strict_order=True)
        ref = D.OracleDecoder(fst, words, *opts)
        bound = 0
        for n, s in enumerate(sents):
            logits = TLG.render_logits([li["prons"][widx[w]] for w in s], T=95, seed=40 + n, noise=1.5)
            ref.reset(); ref.decode_logits(logits, np.zeros_like(logits), math.log(7.0)); ref.finish()
            dec.Reset()
            LM.DecodeNumpy(dec, logits, np.zeros_like(logits), math.log(7.0))
            dec.FinishDecoding()
            assert np.array_equal(dec.tokens_per_frame(), ref.tokens_per_frame()), (order, max_active, n)
            D.cmp_strict(dec.result(), ref.results(), opts[4])
            bound += int(ref.tokens_per_frame().max() > max_active)
        assert bound > 0
