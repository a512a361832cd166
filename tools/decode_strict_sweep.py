"""Width sweep of the WFST decoder: default (two-pass) search, strict serial-order search and the oracle (single-threaded C++ restatement
of the reference decoder) -- 1-best agreement, n-best set agreement, per-frame token counts, time per batch.  Markdown to stdout."""
import math, os, sys, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "tools"))
import numpy as np
import b2t_pkg
import decoder_util as D
import make_toy_tlg as TLG
import make_synth_lm as SL
LM = b2t_pkg.submodule("lm_decoder"); GC = b2t_pkg.submodule("graph_compiler")


def sweep(name, fst, words, batch, widths, bp):
    n = len(batch)
    print(f"\n## {name}: {n} utterances x {batch.shape[1]} frames, beam 17, lattice_beam 8, acoustic_scale 0.325, blank penalty log(90), n-best 100\n")
    print("| max_active | default: ms per batch | strict: ms per batch | oracle: ms per trial | 1-best = oracle: default / strict | n-best set = oracle: default / strict | tokens per frame = oracle: default / strict |")
    print("|---:|---:|---:|---:|---|---|---|")
    for ma in widths:
        opts = (ma, min(200, ma), 17.0, 8.0, 0.325, 1.0, 0.0, 100)
        res, tpf, ms = {}, {}, {}
        for mode in ("default", "strict"):
            dec = LM.BrainSpeechDecoder(LM.DecodeResource(fst, "", "", words, ""), LM.DecodeOptions(*opts), max_frames=128, max_slots=n, strict_order=(mode == "strict"))
            dec.DecodeBatch(batch[:2], blank_penalty=bp)
            dec2 = LM.BrainSpeechDecoder(LM.DecodeResource(fst, "", "", words, ""), LM.DecodeOptions(*opts), max_frames=128, max_slots=n, strict_order=(mode == "strict"))
            dec2.DecodeBatch(batch[:1], blank_penalty=bp); del dec2          # (allocation warm-up on a throw-away decoder: slots keep HashList history)
            del dec
            dec = LM.BrainSpeechDecoder(LM.DecodeResource(fst, "", "", words, ""), LM.DecodeOptions(*opts), max_frames=128, max_slots=n, strict_order=(mode == "strict"))
            t0 = time.perf_counter()
            dec.DecodeBatch(batch, blank_penalty=bp)
            ms[mode] = (time.perf_counter() - t0) * 1e3
            res[mode] = [dec.result(slot=i) for i in range(n)]
            tpf[mode] = [dec.tokens_per_frame(slot=i).tolist() for i in range(n)]
            del dec
        refs, rtpf = [], []
        t1 = time.perf_counter()
        for i in range(n):
            ref = D.OracleDecoder(fst, words, *opts)
            ref.decode_logits(batch[i], np.zeros_like(batch[i]), bp); ref.finish()
            refs.append(ref.results()); rtpf.append(ref.tokens_per_frame().tolist())
        dto = (time.perf_counter() - t1) / n * 1e3
        def agree(mode):
            one = sum(1 for a, b in zip(res[mode], refs) if (a[0].sentence if a else "") == (b[0][2] if b else ""))
            sets = sum(1 for a, b in zip(res[mode], refs) if {x.sentence for x in a} == {x[2] for x in b})
            tok = sum(1 for a, b in zip(tpf[mode], rtpf) if a == b)
            return one, sets, tok
        d1, ds, dt_ = agree("default"); s1, ss, st = agree("strict")
        print(f"| {ma} | {ms['default']:.1f} | {ms['strict']:.1f} | {dto:.1f} | {d1}/{n} / {s1}/{n} | {ds}/{n} / {ss}/{n} | {dt_}/{n} / {st}/{n} |", flush=True)


def main():
    bp = math.log(90.0)
    d = tempfile.mkdtemp()
    fst, words = os.path.join(d, "TLG.fst"), os.path.join(d, "words.txt")
    info = TLG.build(fst, words, n_words=1000, seed=5, bigram_frac=0.02)
    rng = np.random.RandomState(3)
    truth = [rng.randint(0, 1000, size=rng.randint(2, 4)).tolist() for _ in range(16)]
    batch = np.stack([TLG.render_logits([info["prons"][w] for w in truth[i]], T=95, seed=500 + i, noise=1.0) for i in range(16)])
    print("# WFST decoder width sweep: default search, strict serial-order search, oracle (B200)")
    print("\n(n-best sets are compared exactly here; the tests allow hypotheses that tie with the last kept entry to differ)")
    sweep(f"generated TLG ({info['n_states']} states)", fst, words, batch, (10, 100, 500, 7000), bp)
    dd = os.path.join(d, "lm3")
    li = SL.build(dd, order=3, n_words=1000, n_sent=20000, seed=3)
    gfst, gwords = os.path.join(dd, "TLG.fst"), os.path.join(dd, "words.txt")
    gi = GC.compile_to_files(li["arpa"], li["lexicon"], li["phones"], gfst, gwords)
    widx = {w: i for i, w in enumerate(li["words"])}
    rs = np.random.RandomState(11)
    sents = [s[1:-1] for s in li["corpus"] if 2 <= len(s) - 2 <= 4]
    sents = [sents[i] for i in rs.choice(len(sents), size=8, replace=False)]
    ub = np.stack([TLG.render_logits([li["prons"][widx[w]] for w in s], T=95, seed=700 + i, noise=1.0) for i, s in enumerate(sents)])
    sweep(f"3-gram graph compiled by graph_compiler.py ({gi['n_states']} states)", gfst, gwords, ub, (100, 500, 7000), bp)


if __name__ == "__main__":
    main()
