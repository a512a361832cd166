"""Quick timing probe of the WFST decoder host phases (B2T_DECODER_TIMING=1) at one max_active; few utterances."""
import math, os, sys, time, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "tools"))
os.environ["B2T_DECODER_TIMING"] = "1"
import numpy as np
import b2t_pkg
import decoder_util as D
import make_toy_tlg as TLG
LM = b2t_pkg.submodule("lm_decoder")
ma = int(sys.argv[1]) if len(sys.argv) > 1 else 7000
N = int(sys.argv[2]) if len(sys.argv) > 2 else 4
d = tempfile.mkdtemp()
fst, words = os.path.join(d, "TLG.fst"), os.path.join(d, "words.txt")
info = TLG.build(fst, words, n_words=1000, seed=5, bigram_frac=0.02)
rng = np.random.RandomState(3)
truth = [rng.randint(0, 1000, size=rng.randint(2, 4)).tolist() for _ in range(N)]
batch = np.stack([TLG.render_logits([info["prons"][w] for w in truth[n]], T=95, seed=500 + n, noise=1.0) for n in range(N)])
opts = (ma, min(200, ma), 17.0, 8.0, 0.325, 1.0, 0.0, 100)
dec = LM.BrainSpeechDecoder(LM.DecodeResource(fst, "", "", words, ""), LM.DecodeOptions(*opts), max_frames=128, max_slots=N)
dec.DecodeBatch(batch[:1], blank_penalty=math.log(90.0))
t0 = time.perf_counter()
dec.DecodeBatch(batch, blank_penalty=math.log(90.0))
dt = time.perf_counter() - t0
print(f"max_active {ma}: {N} utterances in {dt * 1e3:.1f} ms; search kernel {dec.stats()['kernel_ms']:.1f} ms", flush=True)
ref = D.OracleDecoder(fst, words, *opts)
t1 = time.perf_counter()
same = 0
for n in range(N):
    ref.reset(); ref.decode_logits(batch[n], np.zeros_like(batch[n]), math.log(90.0)); ref.finish()
    r = ref.results(); o = dec.result(slot=n)
    same += int([x.sentence for x in o][:1] == [x[2] for x in r][:1] and {x.sentence for x in o} == {x[2] for x in r})
print(f"oracle: {(time.perf_counter() - t1) / N * 1e3:.1f} ms per utterance; identical 1-best and n-best set: {same}/{N}")
