"""Short workload for ncu captures of the decode-side kernels: WFST search + lattice pruning (4 utterances, max_active 7000,
generated 170 k-state graph) and the prefix beam search (16 utterances, 10 x 10)."""
import math, os, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "tools"))
import numpy as np
import b2t_pkg
import make_toy_tlg as TLG
LM = b2t_pkg.submodule("lm_decoder")
d = tempfile.mkdtemp()
fst, words = os.path.join(d, "TLG.fst"), os.path.join(d, "words.txt")
info = TLG.build(fst, words, n_words=1000, seed=5, bigram_frac=0.02)
rng = np.random.RandomState(3)
N = 4
batch = np.stack([TLG.render_logits([info["prons"][w] for w in rng.randint(0, 1000, size=3)], T=95, seed=500 + n, noise=1.0) for n in range(N)])
dec = LM.BrainSpeechDecoder(LM.DecodeResource(fst, "", "", words, ""), LM.DecodeOptions(7000, 200, 17.0, 8.0, 0.325, 1.0, 0.0, 100), max_frames=128, max_slots=N)
for _ in range(2):
    dec.DecodeBatch(batch, blank_penalty=math.log(90.0))
big = np.stack([TLG.render_logits([info["prons"][w] for w in rng.randint(0, 1000, size=3)], T=95, seed=600 + n, noise=1.0) for n in range(16)])
lp = big - np.log(np.exp(big).sum(-1, keepdims=True))
for _ in range(2):
    LM.ctc_prefix_beam_search(lp, first_beam_size=10, second_beam_size=10)
print("done", len(dec.result(slot=0)))
