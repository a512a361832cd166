"""Generate tests/golden/nbest_postproc.json from the UNMODIFIED reference language-model-standalone.py (build container only):
get_string_differences, remove_punctuation, augment_nbest and the score fusion of gpt2_lm_decode (with rescore_with_gpt2 replaced
by a deterministic stand-in scorer, so that no 6.7 B-parameter model is needed)."""
import importlib.util, json, os, sys, types
import numpy as np
for name in ("redis", "lm_decoder"):
    sys.modules.setdefault(name, types.ModuleType(name))
spec = importlib.util.spec_from_file_location("lm_standalone", "/root/reference/language_model/language-model-standalone.py")
R = importlib.util.module_from_spec(spec)
spec.loader.exec_module(R)

def fake_llm(hyps, length_penalty):
    """Deterministic stand-in for the LLM log-likelihood: depends on the words only."""
    out = []
    for h in hyps:
        s = 0.0
        for w in h.split():
            s -= 1.0 + (sum(ord(c) for c in w) % 17) / 5.0
        out.append(s - len(h.split()) * length_penalty)
    return out
R.rescore_with_gpt2 = lambda model, tok, dev, hyps, lp: fake_llm(hyps, lp)

rng = np.random.RandomState(0)
vocab = ["i", "you", "we", "want", "need", "water", "coffee", "the", "a", "please", "now", "today", "go", "home", "there", "their", "they're", "to", "too", "two"]
def sent(n): return " ".join(vocab[i] for i in rng.randint(0, len(vocab), size=n))
cases = []
for trial in range(12):
    n = int(rng.randint(2, 7))
    base = sent(n).split()
    nbest = []
    for k in range(int(rng.randint(3, 14))):
        w = list(base)
        for _ in range(int(rng.randint(0, 3))):
            w[int(rng.randint(0, n))] = vocab[int(rng.randint(0, len(vocab)))]
        if rng.rand() < 0.2:
            w = w[:-1]
        nbest.append([" ".join(w) + (" " if rng.rand() < 0.3 else ""), float(-rng.rand() * 40 - 5), float(-rng.rand() * 20 - 2)])
    top = int(rng.choice([2, 5, 20])); asc = float(rng.choice([0.3, 0.325, 0.5])); pen = float(rng.choice([0.01, 0.05]))
    aug = R.augment_nbest([list(x) for x in nbest], top_candidates_to_augment=top, acoustic_scale=asc, score_penalty_percent=pen)
    alpha = float(rng.choice([0.0, 0.55, 1.0])); lp = float(rng.choice([0.0, 0.2])); ctx = str(rng.choice(["", "hello there"]))
    best, nb_out, conf = R.gpt2_lm_decode(None, None, "cpu", [list(x) for x in aug], asc, lp, alpha, returnConfidence=True, current_context_str=ctx)
    cases.append({"nbest": nbest, "top": top, "acoustic_scale": asc, "penalty": pen, "augmented": aug, "alpha": alpha, "length_penalty": lp,
                  "context": ctx, "best": best, "nbest_out": nb_out, "confidence": float(conf)})
diffs = []
for _ in range(40):
    a, b = sent(int(rng.randint(1, 7))), sent(int(rng.randint(1, 7)))
    if rng.rand() < 0.5:
        bw = a.split(); bw[int(rng.randint(0, len(bw)))] = "zebra"; b = " ".join(bw)
    cost, path, hl = R.get_string_differences(a, b)
    diffs.append({"cue": a, "out": b, "cost": int(cost), "path": [p if isinstance(p, str) else int(p) for p in path], "highlight": [list(map(int, x)) for x in hl]})
punct = ["Hello, World!", "it's  a -- test- case.", "  What?  I'm 'fine' ", "a-b c - d", "NO. 1 choice!!"]
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden", "nbest_postproc.json")
json.dump({"cases": cases, "diffs": diffs, "punct": [[s, R.remove_punctuation(s)] for s in punct]}, open(OUT, "w"))
print("wrote", OUT, len(cases), len(diffs))
