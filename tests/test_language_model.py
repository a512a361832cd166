"""CPU tests of the decode-side callers (language_model.py): n-best post-processing against vectors produced by the UNMODIFIED
reference language-model-standalone.py (oracle/gen_nbest_golden.py), and the Redis wire protocol over the in-process loopback with
the reference's client-helper call sequence (evaluate_model.py:186-236)."""
import json
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def LM(pkg):
    import b2t_pkg
    return b2t_pkg.submodule("language_model")


@pytest.fixture(scope="module")
def golden():
    return json.load(open(os.path.join(ROOT, "tests", "golden", "nbest_postproc.json")))


def _fake_llm(hyps, length_penalty):            # the stand-in scorer of oracle/gen_nbest_golden.py
    out = []
    for h in hyps:
        s = 0.0
        for w in h.split():
            s -= 1.0 + (sum(ord(c) for c in w) % 17) / 5.0
        out.append(s - len(h.split()) * length_penalty)
    return out


def test_string_differences_and_punctuation(LM, golden):
    for d in golden["diffs"]:
        cost, path, hl = LM.get_string_differences(d["cue"], d["out"])
        assert cost == d["cost"] and list(path) == d["path"] and [list(x) for x in hl] == d["highlight"], d
    for s, want in golden["punct"]:
        assert LM.remove_punctuation(s) == want


def test_augment_nbest_and_fusion_match_reference(LM, golden):
    for c in golden["cases"]:
        aug = LM.augment_nbest([list(x) for x in c["nbest"]], top_candidates_to_augment=c["top"], acoustic_scale=c["acoustic_scale"],
                               score_penalty_percent=c["penalty"])
        assert [a[0] for a in aug] == [a[0] for a in c["augmented"]]                  # same candidates, same order
        assert np.allclose([a[1:] for a in aug], [a[1:] for a in c["augmented"]], rtol=1e-12, atol=0)
        best, nb, conf = LM.fuse_nbest_scores(aug, c["acoustic_scale"], c["length_penalty"], c["alpha"], _fake_llm, returnConfidence=True,
                                              current_context_str=c["context"])
        assert best == c["best"] and abs(conf - c["confidence"]) < 1e-12
        assert len(nb) == len(c["nbest_out"])
        for a, b in zip(nb, c["nbest_out"]):
            fa, fb = a.split(';'), b.split(';')
            assert fa[0] == fb[0] and np.allclose([float(x) for x in fa[1:]], [float(x) for x in fb[1:]], rtol=1e-12)
    # without an LLM the fused total is acoustic_scale * ac + ngram (do_opt = 0 branch, language-model-standalone.py:633-645)
    best, nb = LM.fuse_nbest_scores([["a b", -10.0, -3.0], ["a c", -9.0, -5.0]], 0.5, 0.0, 0.0, None)
    assert best == "a b" and nb[0].endswith(str(0.5 * -10.0 + -3.0))


class _FakeResult:
    def __init__(self, s, a, l):
        self.sentence, self.ac_score, self.lm_score = s, a, l


class _FakeDecoder:
    """Decoder double with the BrainSpeechDecoder call protocol (no GPU in this test)."""
    def __init__(self):
        self.frames, self.calls = 0, []

    def Reset(self):
        self.frames = 0; self.calls.append("reset")

    def Decode(self, logp, slot=0):
        self.frames += logp.shape[0]; self.calls.append("decode")

    def FinishDecoding(self):
        self.calls.append("finish")

    def Rescore(self):
        self.calls.append("rescore")

    def SetOpt(self, o):
        self.calls.append("setopt")

    def result(self):
        if self.frames == 0:
            return []
        return [_FakeResult(f"i want {self.frames} waters", -20.0, -5.0), _FakeResult(f"i need {self.frames} waters", -21.0, -5.5),
                _FakeResult(f"i want {self.frames} water", -22.0, -6.0)]


def test_loopback_wire_protocol(LM, pkg, monkeypatch):
    import b2t_pkg
    H = b2t_pkg.submodule("evaluate_model_helpers")
    lmd = b2t_pkg.submodule("lm_decoder")
    dec = _FakeDecoder()
    monkeypatch.setattr(lmd, "DecodeNumpy", lambda d, logits, pri, bp, slot=0: d.Decode(logits))
    srv = LM.LanguageModelServer(decoder=dec, nbest=100, acoustic_scale=0.3)
    r = LM.LoopbackRedis()
    r.flushall()
    r.serve(srv)
    try:
        last = {k: H.get_current_redis_time_ms(r) for k in ("partial", "final", "reset", "update")}
        last["reset"] = H.reset_remote_language_model(r, last["reset"])
        last["update"] = H.update_remote_lm_params(r, last["update"], acoustic_scale=0.325, blank_penalty=9.0, alpha=0.5)
        assert srv.p["acoustic_scale"] == 0.325 and srv.p["blank_penalty"] == 9.0 and "setopt" in dec.calls
        logits = np.zeros((30, 41), np.float32)
        last["partial"], partial = H.send_logits_to_remote_lm(r, "remote_lm_input", "remote_lm_output_partial", last["partial"], logits)
        assert partial == "i want 30 waters"
        last["partial"], partial = H.send_logits_to_remote_lm(r, "remote_lm_input", "remote_lm_output_partial", last["partial"], logits[:12])
        assert partial == "i want 42 waters"
        last["final"], out = H.finalize_remote_lm(r, "remote_lm_output_final", last["final"])
        assert out["candidate_sentences"][0] == "i want 42 waters"
        assert "i need 42 water" in out["candidate_sentences"]                      # produced by augment_nbest (word swap)
        assert all(abs(t - (0.325 * a + n)) < 1e-9 for t, a, n in zip(out["candidate_total_scores"], out["candidate_acoustic_scores"],
                                                                      out["candidate_ngram_scores"]))
        assert dec.calls.index("finish") > dec.calls.index("decode")
        assert r.xlen("remote_lm_args") >= 2 and r.xlen("remote_lm_done_finalizing") == 1
        last["reset"] = H.reset_remote_language_model(r, last["reset"])
        assert dec.frames == 0
    finally:
        r.stop_serving()
