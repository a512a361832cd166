"""Inference-side measurements (BASELINE.json configs[3], configs[4]) on one B200; writes a markdown report to stdout.

  * GRU logits: one trial at a time (runSingleDecodingStep shape: [1, 400, 512], 'valid' smoothing -> T' = 95) and 64 trials per call
  * WFST token-passing decoder (lm_decoder drop-in) on a generated TLG graph: batch of 64 utterances per call, max_active sweep
    (the "beam width" of the lattice decoder), n-best 100, against the single-threaded C++ restatement of the reference decoder
    (oracle/decoder_oracle.cpp): latency per trial, 1-best agreement, WER against the rendered word sequences
  * LM-free CTC prefix beam search: (first_beam, second_beam) sweep against the same oracle

The real 3-/5-gram graphs of the reference are a download (not in the repository); the graph here is generated
(tools/make_toy_tlg.py: unigram back-off state + bigram histories over a random lexicon)."""
import math, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "tools"))
import numpy as np
import torch
import b2t_pkg, bench
import decoder_util as D
import make_toy_tlg as TLG

E = b2t_pkg.submodule("engine"); LM = b2t_pkg.submodule("lm_decoder")
from torch_cpu_port import PortModel
import gru_ctc_oracle as O


def wer(ref, hyp):
    return O.edit_distance(ref, hyp), len(ref)


def main():
    class Out(list):
        def append(self, x):
            print(x, flush=True)                     # progressive output: a cut-off run still leaves its finished rows
    out = Out()
    only = set(sys.argv[1:])                         # optional section filter: gru / wfst / prefix
    torch.manual_seed(0)
    cfg = E.make_config(**dict(bench.CFG, rnn_dropout=0.0, input_dropout=0.0))
    flat = E.flat_from_state_dict(cfg, PortModel(**bench.CFG).state_dict()).cuda()
    # ---- GRU logits
    out.append("## GRU logits (5 x 768 GRU, 512 features x 400 bins, 'valid' smoothing, T' = 95; random weights)\n")
    out.append("| batch per call | ms per call | trials/s | note |\n|---:|---:|---:|---|")
    for Bq in ((1, 64) if not only or "gru" in only else ()):
        eng = E.Engine(cfg, flat, max_batch=Bq, max_T=400, max_label_len=1, training=False)
        x = torch.randn(Bq, 400, 512, device="cuda"); days = torch.zeros(Bq, dtype=torch.int32)
        for _ in range(5):
            eng.forward(x, days, training=False, smooth_mode=2)
        torch.cuda.synchronize()
        n = 30
        t0 = time.perf_counter()
        for _ in range(n):
            lg, _ = eng.forward(x, days, training=False, smooth_mode=2)
            lg_host = lg.float().cpu().numpy()                       # the reference returns host numpy (implicit sync)
        dt = (time.perf_counter() - t0) / n
        out.append(f"| {Bq} | {dt * 1e3:.3f} | {Bq / dt:.0f} | host wall clock incl. D2H of the logits ([{Bq}, {lg.shape[1]}, 41] f32) |")
        del eng
    # ---- graph + utterances
    import tempfile
    d = tempfile.mkdtemp()
    fst, words = os.path.join(d, "TLG.fst"), os.path.join(d, "words.txt")
    info = TLG.build(fst, words, n_words=1000, seed=5, bigram_frac=0.02)
    Nutt, T = 16, 95
    rng = np.random.RandomState(3)
    truth = [rng.randint(0, 1000, size=rng.randint(2, 4)).tolist() for _ in range(Nutt)]
    batch = np.stack([TLG.render_logits([info["prons"][w] for w in truth[n]], T=T, seed=500 + n, noise=1.0) for n in range(Nutt)])
    out.append(f"\n## WFST decoder, generated TLG ({info['n_states']} states, {info['n_arcs']} arcs), {Nutt} utterances x {T} frames per call, "
               "beam 17, lattice_beam 8, acoustic_scale 0.325, blank penalty log(90), n-best 100\n")
    out.append("| max_active | ours: ms per batch (GPU search + host n-best) | of which GPU search kernel | ours: ms per trial | oracle (1 CPU thread): ms per trial | speed-up | 1-best agreement | WER ours | WER oracle | mean n-best size ours / oracle |")
    out.append("|---:|---:|---:|---:|---:|---:|---:|---:|---:|---|")
    bp = math.log(90.0)
    for ma in ((10, 20, 50, 100, 200, 500, 2000, 7000) if not only or "wfst" in only else ()):
        opts = (ma, min(200, ma), 17.0, 8.0, 0.325, 1.0, 0.0, 100)
        dec = LM.BrainSpeechDecoder(LM.DecodeResource(fst, "", "", words, ""), LM.DecodeOptions(*opts), max_frames=128, max_slots=Nutt)
        dec.DecodeBatch(batch[:2], blank_penalty=bp)                 # warm-up (allocations, graph upload)
        t0 = time.perf_counter()
        dec.DecodeBatch(batch, blank_penalty=bp)
        dt = time.perf_counter() - t0
        kms = max(dec.stats(slot=n)["kernel_ms"] for n in range(Nutt))
        ours = [dec.result(slot=n) for n in range(Nutt)]
        ref = D.OracleDecoder(fst, words, *opts)
        refs = []
        t1 = time.perf_counter()
        for n in range(Nutt):
            ref.reset(); ref.decode_logits(batch[n], np.zeros_like(batch[n]), bp); ref.finish()
            refs.append(ref.results())
        dto = (time.perf_counter() - t1) / Nutt
        agree = sum(1 for a, b in zip(ours, refs) if (a[0].sentence if a else "") == (b[0][2] if b else ""))
        def werf(hyps):
            e = l = 0
            for n, h in enumerate(hyps):
                ref_words = [info["words"][w].lower() for w in truth[n]]
                de, dl = wer(ref_words, h.split())
                e += de; l += dl
            return e / max(l, 1)
        w_o = werf([a[0].sentence if a else "" for a in ours]); w_r = werf([b[0][2] if b else "" for b in refs])
        out.append(f"| {ma} | {dt * 1e3:.1f} | {kms:.1f} | {dt * 1e3 / Nutt:.3f} | {dto * 1e3:.2f} | {dto / (dt / Nutt):.1f}x | {agree}/{Nutt} | {w_o:.3f} | {w_r:.3f} | "
                   f"{np.mean([len(a) for a in ours]):.1f} / {np.mean([len(b) for b in refs]):.1f} |")
        del dec
    # ---- throughput with many utterances in flight (no oracle: it would take minutes)
    if not only or "wfst" in only:
        out.append("\n## WFST decoder throughput, 64 utterances per call (same graph and settings, n-best 100)\n")
        out.append("| max_active | ms per batch | trials/s | GPU search kernel ms | GPU share |\n|---:|---:|---:|---:|---:|")
        rng2 = np.random.RandomState(8)
        big = np.stack([TLG.render_logits([info["prons"][w] for w in rng2.randint(0, 1000, size=rng2.randint(2, 4))], T=T, seed=900 + n, noise=1.0) for n in range(64)])
        for ma in (200, 7000):
            opts = (ma, min(200, ma), 17.0, 8.0, 0.325, 1.0, 0.0, 100)
            dec = LM.BrainSpeechDecoder(LM.DecodeResource(fst, "", "", words, ""), LM.DecodeOptions(*opts), max_frames=128, max_slots=64)
            dec.DecodeBatch(big, blank_penalty=bp)
            t0 = time.perf_counter()
            dec.DecodeBatch(big, blank_penalty=bp)
            dt = time.perf_counter() - t0
            kms = dec.stats(slot=0)["kernel_ms"]
            out.append(f"| {ma} | {dt * 1e3:.1f} | {64 / dt:.0f} | {kms:.1f} | {kms / (dt * 1e3):.2f} |")
            del dec
    # ---- synthetic 3-gram / 5-gram graphs compiled by nejm-brain-to-text_b200/graph_compiler.py (configs[4]: width sweep with n-gram LMs)
    if not only or "lm" in only:
        import importlib.util
        import make_synth_lm as SL
        spec = importlib.util.spec_from_file_location("graph_compiler", os.path.join(ROOT, "nejm-brain-to-text_b200", "graph_compiler.py"))
        GC = importlib.util.module_from_spec(spec); spec.loader.exec_module(GC)
        out.append("\n## Width sweep on synthetic n-gram graphs (1000-word vocabulary, 20 000-sentence corpus; ARPA -> TLG by graph_compiler.py), "
                   "8 utterances sampled from the corpus, shipped decoder settings, n-best 100\n")
        out.append("| LM | graph states / arcs | max_active | ours: ms per trial | oracle: ms per trial | speed-up | 1-best agreement | WER ours | WER oracle |")
        out.append("|---|---|---:|---:|---:|---:|---:|---:|---:|")
        for order in (3, 5):
            dd = os.path.join(d, f"lm{order}")
            li = SL.build(dd, order=order, n_words=1000, n_sent=20000, seed=order)
            gfst, gwords = os.path.join(dd, "TLG.fst"), os.path.join(dd, "words.txt")
            gi = GC.compile_to_files(li["arpa"], li["lexicon"], li["phones"], gfst, gwords)
            widx = {w: i for i, w in enumerate(li["words"])}
            rs = np.random.RandomState(11)
            sents = [s[1:-1] for s in li["corpus"] if 2 <= len(s) - 2 <= 4]
            sents = [sents[i] for i in rs.choice(len(sents), size=8, replace=False)]
            ub = np.stack([TLG.render_logits([li["prons"][widx[w]] for w in s], T=T, seed=700 + n, noise=1.0) for n, s in enumerate(sents)])
            for ma in (10, 100, 500, 7000):
                opts = (ma, min(200, ma), 17.0, 8.0, 0.325, 1.0, 0.0, 100)
                dec = LM.BrainSpeechDecoder(LM.DecodeResource(gfst, "", "", gwords, ""), LM.DecodeOptions(*opts), max_frames=128, max_slots=8)
                dec.DecodeBatch(ub[:2], blank_penalty=bp)
                t0 = time.perf_counter()
                dec.DecodeBatch(ub, blank_penalty=bp)
                dt = (time.perf_counter() - t0) / 8
                ours = [dec.result(slot=n) for n in range(8)]
                ref = D.OracleDecoder(gfst, gwords, *opts)
                t1 = time.perf_counter()
                refs = []
                for n in range(8):
                    ref.reset(); ref.decode_logits(ub[n], np.zeros_like(ub[n]), bp); ref.finish()
                    refs.append(ref.results())
                dto = (time.perf_counter() - t1) / 8
                agree = sum(1 for a, b in zip(ours, refs) if (a[0].sentence if a else "") == (b[0][2] if b else ""))
                def werl(hyps):
                    e = l = 0
                    for n, h in enumerate(hyps):
                        de, dl = wer(sents[n], h.split()); e += de; l += dl
                    return e / max(l, 1)
                out.append(f"| {order}-gram | {gi['n_states']} / {gi['n_arcs']} | {ma} | {dt * 1e3:.2f} | {dto * 1e3:.2f} | {dto / dt:.1f}x | {agree}/8 | "
                           f"{werl([a[0].sentence if a else '' for a in ours]):.3f} | {werl([b[0][2] if b else '' for b in refs]):.3f} |")
                del dec
    # ---- prefix beam
    out.append(f"\n## LM-free CTC prefix beam search, {Nutt} utterances x {T} frames per call (log-softmax of the same logits)\n")
    out.append("| first_beam x second_beam | ours: ms per batch | ours: ms per trial | oracle (1 CPU thread): ms per trial | identical hypothesis lists |")
    out.append("|---|---:|---:|---:|---:|")
    lp = batch - np.log(np.exp(batch).sum(-1, keepdims=True))
    for fb, sb in (((10, 10), (10, 20), (10, 50), (10, 64), (41, 16)) if not only or "prefix" in only else ()):
        LM.ctc_prefix_beam_search(lp, first_beam_size=fb, second_beam_size=sb)     # warm-up (device workspace)
        t0 = time.perf_counter()
        ours = LM.ctc_prefix_beam_search(lp, first_beam_size=fb, second_beam_size=sb)
        dt = time.perf_counter() - t0
        t1 = time.perf_counter()
        refs = [D.prefix_search(lp[n], fb, sb) for n in range(Nutt)]
        dto = (time.perf_counter() - t1) / Nutt
        same = sum(1 for a, b in zip(ours, refs) if [r[0] for r in a] == [r[0] for r in b])
        out.append(f"| {fb} x {sb} | {dt * 1e3:.1f} | {dt * 1e3 / Nutt:.3f} | {dto * 1e3:.2f} | {same}/{Nutt} |")
    

if __name__ == "__main__":
    main()
