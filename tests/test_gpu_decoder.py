"""GPU parity tests of the WFST decoder (lm_decoder drop-in) against the C++ oracle on generated TLG graphs.

Parity definition (SURVEY.md section 7 / DESIGN.md): integer outputs -- the 1-best word sequence, the n-best SET of
word sequences -- identical; scores within 1e-3; per-frame token counts identical when max_active is not binding
(when it is, Kaldi's result depends on hash iteration order and only the 1-best / scores are compared)."""
import math
import os
import sys

import numpy as np
import pytest

import decoder_util as D

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import make_toy_tlg as TLG  # noqa: E402


@pytest.fixture(scope="module")
def LM(pkg):
    import b2t_pkg
    return b2t_pkg.submodule("lm_decoder")


@pytest.fixture(scope="module")
def graph(tmp_path_factory):
    d = tmp_path_factory.mktemp("tlg")
    fst, words = str(d / "TLG.fst"), str(d / "words.txt")
    info = TLG.build(fst, words, n_words=300, seed=2)
    return fst, words, info


def _ours(LM, fst, words, opts, **kw):
    return LM.BrainSpeechDecoder(LM.DecodeResource(fst, "", "", words, ""), LM.DecodeOptions(*opts), **kw)


def _cmp(ours, ref, tol=1e-3):
    assert len(ours) == len(ref), (len(ours), len(ref))
    assert ours[0].sentence == ref[0][2]
    assert {r.sentence for r in ours} == {r[2] for r in ref}
    byref = {r[2]: r for r in ref}
    for r in ours:
        assert abs(r.ac_score - byref[r.sentence][0]) < tol * max(1.0, abs(r.ac_score)), r
        assert abs(r.lm_score - byref[r.sentence][1]) < tol * max(1.0, abs(r.lm_score)), r


@pytest.mark.parametrize("seed", [0, 1, 2, 3])
def test_nbest_matches_oracle(LM, graph, seed):
    fst, words, info = graph
    rng = np.random.RandomState(seed)
    seq = rng.randint(0, 300, size=rng.randint(2, 6))
    logits = TLG.render_logits([info["prons"][w] for w in seq], T=110, seed=seed, noise=1.2)
    opts = (7000, 200, 14.0, 8.0, 0.6, 1.0, 0.0, 50)
    ref = D.OracleDecoder(fst, words, *opts)
    ref.decode_logits(logits, np.zeros_like(logits), math.log(3.0))
    ref.finish()
    dec = _ours(LM, fst, words, opts, max_frames=128)
    dec.Reset()
    LM.DecodeNumpy(dec, logits, np.zeros_like(logits), math.log(3.0))
    partial = dec.result()
    assert len(partial) == 1 and dec.DecodedSomething()
    dec.FinishDecoding()
    _cmp(dec.result(), ref.results())
    if ref.tokens_per_frame().max() <= 7000:                       # max_active not binding: token sets must agree exactly
        assert np.array_equal(dec.tokens_per_frame(), ref.tokens_per_frame())


def test_one_best_chunked_and_reset(LM, graph):
    fst, words, info = graph
    logits = TLG.render_logits([info["prons"][w] for w in [10, 20, 30]], T=100, seed=9)
    opts = (3000, 200, 12.0, 6.0, 0.6, 1.0, 0.0, 1)
    ref = D.OracleDecoder(fst, words, *opts)
    lp = logits - np.log(np.exp(logits).sum(1, keepdims=True))
    ref.decode_logprobs(lp.astype(np.float32)); ref.finish()
    dec = _ours(LM, fst, words, opts, max_frames=128)
    for rep in range(2):                                           # second pass exercises Reset()
        dec.Reset()
        for i in range(0, 100, 32):
            LM.DecodeNumpyLogProbs(dec, lp[i:i + 32].astype(np.float32))
        dec.FinishDecoding()
        _cmp(dec.result(), ref.results())


def test_blank_skip_and_length_penalty(LM, graph):
    fst, words, info = graph
    logits = TLG.render_logits([info["prons"][w] for w in [40, 41]], T=90, seed=4, peak=9.0, noise=0.3)
    logits[logits.argmax(1) == 0, 0] += 10.0
    opts = (7000, 200, 14.0, 8.0, 0.6, 0.9, -0.5, 20)
    ref = D.OracleDecoder(fst, words, *opts)
    ref.decode_logits(logits, None, 0.0); ref.finish()
    dec = _ours(LM, fst, words, opts, max_frames=128)
    LM.DecodeNumpy(dec, logits, np.zeros_like(logits), 0.0)
    dec.FinishDecoding()
    _cmp(dec.result(), ref.results())
    assert len(dec.tokens_per_frame()) == len(ref.tokens_per_frame()) < 90


def test_max_active_binding_one_best(LM, graph):
    fst, words, info = graph
    logits = TLG.render_logits([info["prons"][w] for w in [7, 8, 9, 10]], T=120, seed=11, noise=1.5)
    opts = (300, 50, 16.0, 6.0, 0.5, 1.0, 0.0, 10)                 # tiny max_active: pruning is order dependent in Kaldi
    ref = D.OracleDecoder(fst, words, *opts)
    ref.decode_logits(logits, None, 0.0); ref.finish()
    dec = _ours(LM, fst, words, opts, max_frames=128)
    LM.DecodeNumpy(dec, logits, np.zeros_like(logits), 0.0)
    dec.FinishDecoding()
    ours, r = dec.result(), ref.results()
    assert ours[0].sentence == r[0][2]
    assert abs(ours[0].ac_score - r[0][0]) < 1e-3 * abs(r[0][0]) and abs(ours[0].lm_score - r[0][1]) < 1e-3 * max(1.0, abs(r[0][1]))


def test_batch_decode_equals_single(LM, graph):
    fst, words, info = graph
    opts = (7000, 200, 14.0, 8.0, 0.6, 1.0, 0.0, 10)
    N, T = 12, 100
    rng = np.random.RandomState(5)
    batch = np.stack([TLG.render_logits([info["prons"][w] for w in rng.randint(0, 300, size=3)], T=T, seed=100 + n) for n in range(N)])
    dec = _ours(LM, fst, words, opts, max_frames=128, max_slots=N)
    dec.DecodeBatch(batch, blank_penalty=math.log(2.0))
    single = _ours(LM, fst, words, opts, max_frames=128)
    for n in range(N):
        single.Reset()
        LM.DecodeNumpy(single, batch[n], np.zeros_like(batch[n]), math.log(2.0))
        single.FinishDecoding()
        a, b = dec.result(slot=n), single.result()
        assert [r.sentence for r in a] == [r.sentence for r in b]
        assert all(abs(x.ac_score - y.ac_score) < 1e-4 and abs(x.lm_score - y.lm_score) < 1e-4 for x, y in zip(a, b))


def test_errors_are_exceptions(LM, graph, pkg):
    fst, words, info = graph
    with pytest.raises(pkg._native.B2TError):
        _ours(LM, "/nonexistent/TLG.fst", words, (7000, 200, 17.0, 8.0, 0.3, 1.0, 0.0, 10))
    dec = _ours(LM, fst, words, (7000, 200, 17.0, 8.0, 0.3, 1.0, 0.0, 10), max_frames=16)
    with pytest.raises(pkg._native.B2TError):
        LM.DecodeNumpy(dec, np.zeros((40, 41), np.float32), np.zeros((40, 41), np.float32), 0.0)   # longer than max_frames
    with pytest.raises(pkg._native.B2TError):
        dec.Rescore()


def test_prefix_beam_golden_and_oracle(LM):
    """The reference's golden 3x3 vector (ctc_prefix_beam_search_test.cc:18-59) and random parity vs the oracle."""
    data = np.log(np.array([0.25, 0.40, 0.35, 0.40, 0.35, 0.25, 0.10, 0.50, 0.40], dtype=np.float32).reshape(3, 3))
    res = LM.ctc_prefix_beam_search(data, first_beam_size=3, second_beam_size=3)[0]
    assert [r[0] for r in res] == [[2, 1], [1, 2], [1]]
    for got, want in zip([math.exp(r[1]) for r in res], [0.2185, 0.1550, 0.1525]):
        assert abs(got - want) < 4e-6 * want
    for got, want in zip([math.exp(r[2]) for r in res], [0.07, 0.064, 0.07]):
        assert abs(got - want) < 4e-6 * want
    assert [r[3] for r in res] == [[0, 2], [0, 2], [2]]
    rng = np.random.RandomState(0)
    x = rng.randn(6, 60, 41).astype(np.float32) * 2.0
    x[..., 0] += 2.5
    lp = x - np.log(np.exp(x).sum(-1, keepdims=True))
    lens = np.array([60, 45, 60, 10, 0, 33], dtype=np.int32)
    ours = LM.ctc_prefix_beam_search(lp, lens=lens)
    for n in range(6):
        ref = D.prefix_search(lp[n, :lens[n]]) if lens[n] > 0 else [([], 0.0, 0.0, [])]
        assert [r[0] for r in ours[n]] == [r[0] for r in ref]                        # identical hypotheses, same order
        assert all(abs(a[1] - b[1]) < 1e-4 and abs(a[2] - b[2]) < 1e-4 for a, b in zip(ours[n], ref))
        assert [r[3] for r in ours[n]] == [r[3] for r in ref]


def test_gpu_lattice_pruning_equals_host_pruning(LM, graph, monkeypatch):
    """lattice_prune_kernel (extra costs + compaction on the device) against the host restatement of FinalizeDecoding:
    identical n-best lists, scores included, on a wide beam where most of the lattice is pruned away."""
    fst, words, info = graph
    opts = (7000, 200, 17.0, 8.0, 0.325, 1.0, 0.0, 100)
    N, T = 6, 100
    rng = np.random.RandomState(21)
    batch = np.stack([TLG.render_logits([info["prons"][w] for w in rng.randint(0, 300, size=rng.randint(2, 5))], T=T, seed=300 + n, noise=1.3)
                      for n in range(N)])
    res = {}
    for mode in ("0", "1"):
        monkeypatch.setenv("B2T_DECODER_HOST_PRUNE", mode)
        dec = _ours(LM, fst, words, opts, max_frames=128, max_slots=N)
        dec.DecodeBatch(batch, blank_penalty=math.log(90.0))
        res[mode] = [[(r.sentence, r.ac_score, r.lm_score) for r in dec.result(slot=n)] for n in range(N)]
        single = _ours(LM, fst, words, opts, max_frames=128)       # the single-utterance entry points take the same path
        LM.DecodeNumpy(single, batch[0], np.zeros_like(batch[0]), math.log(90.0))
        single.FinishDecoding()
        single.FinishDecoding()                                    # idempotent
        assert [(r.sentence, r.ac_score, r.lm_score) for r in single.result()] == res[mode][0]
    assert all(len(r) > 1 for r in res["0"])
    assert res["0"] == res["1"]
    # 1-best: back-pointer walk on the device (partial result after every chunk, final result with final costs) vs host walk
    one = {}
    for mode in ("0", "1"):
        monkeypatch.setenv("B2T_DECODER_HOST_PRUNE", mode)
        dec = _ours(LM, fst, words, (7000, 200, 17.0, 8.0, 0.325, 1.0, 0.0, 1), max_frames=128)
        got = []
        for n in range(3):
            dec.Reset()
            lp = batch[n] - np.log(np.exp(batch[n]).sum(1, keepdims=True))
            for i in range(0, T, 25):
                LM.DecodeNumpyLogProbs(dec, lp[i:i + 25].astype(np.float32))
                got.append([(r.sentence, r.ac_score, r.lm_score) for r in dec.result()])
            dec.FinishDecoding()
            got.append([(r.sentence, r.ac_score, r.lm_score) for r in dec.result()])
        one[mode] = got
    assert one["0"] == one["1"] and all(len(g) == 1 for g in one["0"])


def test_prefix_beam_wide(LM):
    """Beam sizes at the kernel's limits (second beam 64, first beam = all classes) against the oracle."""
    rng = np.random.RandomState(4)
    x = rng.randn(3, 40, 41).astype(np.float32) * 1.5
    x[..., 0] += 2.0
    lp = x - np.log(np.exp(x).sum(-1, keepdims=True))
    for fb, sb in ((41, 16), (10, 64)):
        ours = LM.ctc_prefix_beam_search(lp, first_beam_size=fb, second_beam_size=sb)
        for n in range(3):
            ref = D.prefix_search(lp[n], fb, sb)
            assert [r[0] for r in ours[n]] == [r[0] for r in ref]
            assert all(abs(a[1] - b[1]) < 1e-4 and abs(a[2] - b[2]) < 1e-4 for a, b in zip(ours[n], ref))


@pytest.mark.parametrize("sb", [100, 500, 512])
def test_prefix_beam_sweep_widths(LM, sb):
    """BASELINE.json configs[4] sweeps the beam width to 500: second beams beyond the old limit of 64 (the emulated unordered_map
    grows through 257 and 541 buckets), hypotheses / scores / Viterbi times against the oracle."""
    rng = np.random.RandomState(40 + sb)
    x = rng.randn(2, 60, 41).astype(np.float32) * 1.2
    x[..., 0] += 1.5
    lp = x - np.log(np.exp(x).sum(-1, keepdims=True))
    ours = LM.ctc_prefix_beam_search(lp, first_beam_size=10, second_beam_size=sb)
    for n in range(2):
        ref = D.prefix_search(lp[n], 10, sb)
        assert len(ours[n]) == len(ref) and len(ref) > 64
        # hypotheses with exactly equal scores come out in creation order here and in std::sort's (unstable) order in the
        # reference: compare per hypothesis, and the order through the scores
        so = {tuple(r[0]): r for r in ours[n]}
        sr = {tuple(r[0]): r for r in ref}
        assert set(so) == set(sr)
        for k, r in sr.items():
            assert so[k][1] == pytest.approx(r[1], abs=1e-4) and so[k][2] == pytest.approx(r[2], abs=1e-4) and so[k][3] == r[3], k
        assert all(a[1] >= b[1] for a, b in zip(ours[n], ours[n][1:]))
        assert [r[0] for r in ours[n][:20]] == [r[0] for r in ref[:20]] or len({round(r[1], 6) for r in ref[:21]}) < 21


def test_prefix_beam_width_500_full_first_beam(LM):
    """"for prefix search first_beam = min(41, w), second_beam = w".  At w = 500 the 21 000 candidates of a
    frame exceed shared memory and live in global memory; same hypotheses, scores and times as the oracle."""
    rng = np.random.RandomState(77)
    x = rng.randn(1, 14, 41).astype(np.float32) * 1.2
    x[..., 0] += 1.0
    lp = x - np.log(np.exp(x).sum(-1, keepdims=True))
    ours = LM.ctc_prefix_beam_search(lp, first_beam_size=41, second_beam_size=500)[0]
    ref = D.prefix_search(lp[0], 41, 500)
    assert len(ours) == len(ref) == 500
    so = {tuple(r[0]): r for r in ours}
    sr = {tuple(r[0]): r for r in ref}
    worst = min(r[1] for r in ref)
    cut = lambda d: {k for k, v in d.items() if v[1] > worst + 1e-4}          # (ties with the last kept hypothesis may differ)
    assert cut(so) == cut(sr)
    for k in cut(sr):
        assert so[k][1] == pytest.approx(sr[k][1], abs=1e-4) and so[k][2] == pytest.approx(sr[k][2], abs=1e-4) and so[k][3] == sr[k][3], k


def test_prefix_beam_rejects_beams_beyond_the_caps(LM):
    lp = np.log(np.full((1, 4, 41), 1.0 / 41, dtype=np.float32))
    with pytest.raises(Exception, match="beam sizes"):
        LM.ctc_prefix_beam_search(lp, first_beam_size=10, second_beam_size=513)


@pytest.mark.skipif(D.real_graph() is None, reason="the shipped 1-gram graph is staged under oracle/_ref/ by __graft_entry__.build()")
@pytest.mark.parametrize("max_active,n_utt", [(7000, 2), (500, 5)])
def test_shipped_1gram_graph_vs_oracle(LM, max_active, n_utt):
    """The reference's own decoding graph (openwebtext 1-gram TLG, 179 946 states) at its shipped decoder settings
    (LM/README.md: beam 17, lattice_beam 8, acoustic_scale 0.325, blank penalty log 90, n-best 100): utterances rendered from
    random walks through the graph, GPU decoder against the oracle."""
    fst, words = D.real_graph()
    g = D.read_fst(fst)
    rng = np.random.RandomState(7)
    utts = []
    while len(utts) < n_utt:
        u = D.random_walk_utterance(g, rng, n_words=int(rng.randint(1, 4)), peak=9.0, noise=0.6)
        if u is not None:
            utts.append(u[0])
    batch = np.stack(utts)
    opts = (max_active, 200, 17.0, 8.0, 0.325, 1.0, 0.0, 100)
    dec = _ours(LM, fst, words, opts, max_frames=128, max_slots=n_utt)
    dec.DecodeBatch(batch, blank_penalty=math.log(90.0))
    ref = D.OracleDecoder(fst, words, *opts)
    for n in range(n_utt):
        ref.reset(); ref.decode_logits(batch[n], np.zeros_like(batch[n]), math.log(90.0)); ref.finish()
        ours, r = dec.result(slot=n), ref.results()
        # max_active is binding on this graph (thousands of tokens per frame), where Kaldi's own pruning depends on hash order
        # (decode parity note), and the vocabulary is full of homophones with near-equal scores at the n-best cut:
        # the 1-best and the scores of the shared hypotheses must agree, the two n-best sets must overlap almost entirely
        byref = {x[2]: x for x in r}
        byours = {x.sentence: x for x in ours}
        shared = [x for x in ours if x.sentence in byref]
        info = (ours[0], r[0], len(shared), len(ours), len(r))
        if max_active >= 7000:
            assert ours[0].sentence == r[0][2], info
        else:       # heavily binding: either decoder's best hypothesis must at least be a hypothesis of the other, at the same score
            assert ours[0].sentence in byref and r[0][2] in byours, info
        assert len(shared) >= 0.8 * max(len(ours), len(r)), info
        for x in shared:
            assert abs(x.ac_score - byref[x.sentence][0]) < 1e-3 * max(1.0, abs(x.ac_score))
            assert abs(x.lm_score - byref[x.sentence][1]) < 1e-3 * max(1.0, abs(x.lm_score))


REF_LOGITS = os.path.join(ROOT, "oracle", "_ref", "test_logits.npy")


@pytest.mark.skipif(not os.path.exists(REF_LOGITS), reason="test_logits.npy (input fixture of the reference's x86/python/test.py) is staged under oracle/_ref/ by build()")
def test_prefix_beam_on_reference_logits(LM):
    """The reference's own handwriting logits [32, 3469, 32] (x86/python/test.py:30-36, columns rearranged as there): long
    sequences (3 469 frames, prefixes of several hundred tokens) through the prefix beam search, GPU vs oracle."""
    logits = np.load(REF_LOGITS)
    logits = logits[:, :, [31] + [26, 27, 30, 29, 28] + list(range(26))]
    x = logits[:3, :1200].astype(np.float32)
    lp = x - x.max(-1, keepdims=True)
    lp = lp - np.log(np.exp(lp).sum(-1, keepdims=True))
    lens = np.array([1200, 777, 1200], dtype=np.int32)
    ours = LM.ctc_prefix_beam_search(lp, lens=lens)
    for n in range(3):
        ref = D.prefix_search(lp[n, :lens[n]])
        assert len(ours[n]) == len(ref)
        # The tail of these beams holds several hypotheses with EXACTLY the same score (tokens at the probability floor); which of
        # them survive the second-beam cut is decided by std::nth_element / std::sort in the reference and by creation order in the
        # kernel.  Everything strictly above the cut score must agree exactly: hypotheses, scores, Viterbi scores and Viterbi times
        # (the times depend on the iteration order of the reference's unordered_map, which the kernel reproduces: UMapOrder).
        cut = ref[-1][1]
        k = sum(1 for r in ref if r[1] > cut + 1e-6)
        assert k >= 5
        assert [r[0] for r in ours[n][:k]] == [r[0] for r in ref[:k]]
        assert all(abs(a[1] - b[1]) < 1e-3 * max(1.0, abs(b[1])) and abs(a[2] - b[2]) < 1e-3 * max(1.0, abs(b[2])) for a, b in zip(ours[n][:k], ref[:k]))
        assert [r[3] for r in ours[n][:k]] == [r[3] for r in ref[:k]]
        assert all(abs(a[1] - cut) < 1e-3 * abs(cut) for a in ours[n][k:])          # the rest are other members of the tie
        assert len(ours[n][0][0]) > 20


# ---------------------------------------------------------------------------------------------- strict serial-order mode
@pytest.mark.parametrize("max_active", [100, 500, 7000])
def test_strict_order_toy_graph(LM, graph, max_active):
    """max_active binding (noisy posteriors on the generated graph): Kaldi's result depends on its token-list order and on the
    online next_cutoff tightening (lattice-faster-decoder.cc:785-822); the strict mode reproduces both, so per-frame token
    counts, 1-best and the n-best set equal the oracle's."""
    fst, words, info = graph
    opts = (max_active, 20 if max_active <= 100 else 200, 14.0, 8.0, 0.6, 1.0, 0.0, 50)
    dec = _ours(LM, fst, words, opts, max_frames=128, strict_order=True)
    ref = D.OracleDecoder(fst, words, *opts)
    bound = 0
    for seed in range(4):
        rng = np.random.RandomState(100 + seed)
        seq = rng.randint(0, 300, size=rng.randint(2, 6))
        logits = TLG.render_logits([info["prons"][w] for w in seq], T=110, seed=seed, noise=2.0)
        ref.reset(); ref.decode_logits(logits, np.zeros_like(logits), math.log(3.0)); ref.finish()
        dec.Reset()
        LM.DecodeNumpy(dec, logits, np.zeros_like(logits), math.log(3.0))
        dec.FinishDecoding()
        assert np.array_equal(dec.tokens_per_frame(), ref.tokens_per_frame()), (seed, dec.tokens_per_frame()[:12], ref.tokens_per_frame()[:12])
        D.cmp_strict(dec.result(), ref.results(), opts[4])
        bound += int(ref.tokens_per_frame().max() > max_active)
    if max_active <= 500:
        assert bound > 0, "the case is meant to exercise a binding max_active"


def test_strict_order_switch_and_chunks(LM, graph):
    """Switching the order between utterances and feeding in chunks gives the same strict result."""
    fst, words, info = graph
    opts = (300, 100, 14.0, 8.0, 0.6, 1.0, 0.0, 20)
    logits = TLG.render_logits([info["prons"][w] for w in [5, 50, 150]], T=100, seed=3, noise=2.0)
    lp = (logits - np.log(np.exp(logits).sum(1, keepdims=True))).astype(np.float32)
    ref = D.OracleDecoder(fst, words, *opts)
    ref.decode_logprobs(lp); ref.finish()
    dec = _ours(LM, fst, words, opts, max_frames=128)
    dec.Reset(); LM.DecodeNumpyLogProbs(dec, lp); dec.FinishDecoding()
    fast = [r.sentence for r in dec.result()]
    dec.set_strict_order(True)
    dec.Reset()
    for i in range(0, 100, 32):
        LM.DecodeNumpyLogProbs(dec, lp[i:i + 32])
    dec.FinishDecoding()
    assert np.array_equal(dec.tokens_per_frame(), ref.tokens_per_frame())
    D.cmp_strict(dec.result(), ref.results(), opts[4])
    dec.set_strict_order(False)
    dec.Reset(); LM.DecodeNumpyLogProbs(dec, lp); dec.FinishDecoding()
    assert [r.sentence for r in dec.result()] == fast


@pytest.mark.skipif(D.real_graph() is None, reason="the shipped 1-gram graph is staged under oracle/_ref/ by __graft_entry__.build()")
@pytest.mark.parametrize("max_active,n_utt", [(7000, 2), (500, 4), (100, 4)])
def test_strict_order_shipped_1gram_graph(LM, max_active, n_utt):
    """The reference's own graph at its shipped settings, where max_active binds: strict mode == oracle (counts, 1-best, n-best)."""
    fst, words = D.real_graph()
    g = D.read_fst(fst)
    rng = np.random.RandomState(7)
    utts = []
    while len(utts) < n_utt:
        u = D.random_walk_utterance(g, rng, n_words=int(rng.randint(1, 4)), peak=9.0, noise=0.6)
        if u is not None:
            utts.append(u[0])
    batch = np.stack(utts)
    opts = (max_active, 200 if max_active > 200 else 50, 17.0, 8.0, 0.325, 1.0, 0.0, 100)
    dec = _ours(LM, fst, words, opts, max_frames=128, max_slots=n_utt, strict_order=True)
    dec.DecodeBatch(batch, blank_penalty=math.log(90.0))
    for n in range(n_utt):
        ref = D.OracleDecoder(fst, words, *opts)       # every slot is a decoder of its own (the HashList size is decoder history)
        ref.decode_logits(batch[n], np.zeros_like(batch[n]), math.log(90.0)); ref.finish()
        assert np.array_equal(dec.tokens_per_frame(slot=n), ref.tokens_per_frame()), n
        D.cmp_strict(dec.result(slot=n), ref.results(), opts[4])
