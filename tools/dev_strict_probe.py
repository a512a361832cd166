"""Bring-up: strict-order decoder vs the oracle, frame by frame (token list order and costs)."""
import math, os, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "tools"))
import numpy as np
import b2t_pkg
import decoder_util as D
import make_toy_tlg as TLG
LM = b2t_pkg.submodule("lm_decoder")
max_active = int(os.environ.get("MA", "100")); seed = int(os.environ.get("SEED", "1"))
d = tempfile.mkdtemp()
fst, words = os.path.join(d, "TLG.fst"), os.path.join(d, "words.txt")
info = TLG.build(fst, words, n_words=300, seed=2)
opts = (max_active, 20 if max_active <= 100 else 200, 14.0, 8.0, 0.6, 1.0, 0.0, 50)
rng = np.random.RandomState(100 + seed)
seq = rng.randint(0, 300, size=rng.randint(2, 6))
logits = TLG.render_logits([info["prons"][w] for w in seq], T=110, seed=seed, noise=2.0)
x = logits - logits.max(1, keepdims=True)
lp = (x - np.log(np.exp(x).sum(1, keepdims=True))).astype(np.float32)
lp[:, 0] -= math.log(3.0)
dec = LM.BrainSpeechDecoder(LM.DecodeResource(fst, "", "", words, ""), LM.DecodeOptions(*opts), max_frames=128, strict_order=True)
ref = D.OracleDecoder(fst, words, *opts)
dec.Reset(); ref.reset()
s0, c0 = dec.debug_frame_tokens(0) if False else (None, None)
for t in range(lp.shape[0]):
    LM.DecodeNumpyLogProbs(dec, lp[t:t + 1]); ref.decode_logprobs(lp[t:t + 1])
    so, co = dec.debug_frame_tokens(t + 1); sr, cr = ref.token_list()
    same_set = set(so.tolist()) == set(sr.tolist())
    same_order = len(so) == len(sr) and np.array_equal(so, sr)
    cost_ok = same_order and np.array_equal(co, cr)
    print(t, len(so), len(sr), "set", same_set, "order", same_order, "cost", cost_ok)
    if not same_order:
        k = next((i for i in range(min(len(so), len(sr))) if so[i] != sr[i]), min(len(so), len(sr)))
        print("first order difference at", k, so[max(0, k - 3):k + 5], sr[max(0, k - 3):k + 5])
        print("only ours", sorted(set(so.tolist()) - set(sr.tolist()))[:20], "only ref", sorted(set(sr.tolist()) - set(so.tolist()))[:20])
        if not same_set or int(os.environ.get("STOP", "1")):
            break
