"""Synthetic n-gram language model + lexicon for decoder sweeps (the reference's 3-/5-gram models are downloads).

A random 'language' (Zipfian vocabulary, first-order Markov word generator so that higher-order n-grams carry information)
is sampled into a corpus; n-gram counts up to `order` become an ARPA file with absolute-discounting style probabilities and
back-off weights (normalised per history).  Pronunciations are random phone strings.

    python tools/make_synth_lm.py out_dir [order] [n_words] [n_sentences]
"""
import math, os, sys
from collections import Counter, defaultdict

import numpy as np


def build(out_dir, order=3, n_words=1000, n_sent=20000, seed=0, n_phones=39):
    rng = np.random.RandomState(seed)
    os.makedirs(out_dir, exist_ok=True)
    words = [f"w{i}" for i in range(n_words)]
    prons, seen = [], set()
    while len(prons) < n_words:
        pr = tuple(int(p) for p in rng.randint(0, n_phones, size=rng.randint(2, 6)))
        if pr not in seen:
            seen.add(pr); prons.append(pr)
    zipf = 1.0 / np.arange(1, n_words + 1) ** 1.1
    zipf /= zipf.sum()
    succ = [rng.choice(n_words, size=12, p=zipf) for _ in range(n_words)]     # every word prefers a few successors
    corpus = []
    for _ in range(n_sent):
        L = rng.randint(2, 8)
        s = [int(rng.choice(n_words, p=zipf))]
        while len(s) < L:
            s.append(int(succ[s[-1]][rng.randint(12)]) if rng.rand() < 0.7 else int(rng.choice(n_words, p=zipf)))
        corpus.append(["<s>"] + [words[i] for i in s] + ["</s>"])
    counts = [Counter() for _ in range(order + 1)]
    for sent in corpus:
        for n in range(1, order + 1):
            for i in range(len(sent) - n + 1):
                ng = tuple(sent[i:i + n])
                if n > 1 or ng[0] != "<s>":
                    counts[n][ng] += 1
    counts[1][("<s>",)] = len(corpus)
    D = 0.5
    # probabilities: absolute discounting with interpolation expressed in back-off form
    ctx_total, ctx_types = [defaultdict(int) for _ in range(order + 1)], [defaultdict(int) for _ in range(order + 1)]
    for n in range(1, order + 1):
        for ng, c in counts[n].items():
            if n == 1 and ng[0] == "<s>":
                continue
            ctx_total[n][ng[:-1]] += c
            ctx_types[n][ng[:-1]] += 1
    logp, bow = [dict() for _ in range(order + 1)], [dict() for _ in range(order + 1)]
    uni_total = ctx_total[1][()]
    for ng, c in counts[1].items():
        logp[1][ng] = -99.0 if ng[0] == "<s>" else math.log10(c / uni_total)
    for n in range(2, order + 1):
        for ng, c in counts[n].items():
            if c < (2 if n >= 3 else 1):
                continue
            logp[n][ng] = math.log10(max(c - D, 0.1) / ctx_total[n][ng[:-1]])
    # back-off mass per history: 1 - sum of the kept higher-order probabilities, divided by the lower-order mass of the same words
    by_hist = [defaultdict(list) for _ in range(order + 1)]
    for n in range(2, order + 1):
        for ng, lp in logp[n].items():
            by_hist[n][ng[:-1]].append((ng[-1], lp))

    def lower_prob(h, w):                    # back-off probability of w after history h (h one shorter than the n-gram's)
        n = len(h) + 1
        ng = h + (w,)
        if ng in logp[n]:
            return 10 ** logp[n][ng]
        if not h:
            return 1e-9
        return 10 ** bow[n - 1].get(h, 0.0) * lower_prob(h[1:], w)

    for n in range(1, order):
        for h, lst in by_hist[n + 1].items():
            if h not in logp[n]:
                logp[n][h] = -5.0            # a context must exist as an n-gram of its own order
        for h, lst in sorted(by_hist[n + 1].items(), key=lambda kv: len(kv[0])):
            kept = sum(10 ** lp for _, lp in lst)
            low = sum(lower_prob(h[1:], w) for w, _ in lst)
            bow[n][h] = math.log10(max(1.0 - kept, 1e-4) / max(1.0 - low, 1e-4))
    arpa = os.path.join(out_dir, f"lm{order}.arpa")
    with open(arpa, "w") as f:
        f.write("\\data\\\n")
        for n in range(1, order + 1):
            f.write(f"ngram {n}={len(logp[n])}\n")
        for n in range(1, order + 1):
            f.write(f"\n\\{n}-grams:\n")
            for ng, lp in logp[n].items():
                b = bow[n].get(ng) if n < order else None
                f.write(f"{lp:.6f} {' '.join(ng)}" + (f" {b:.6f}" if b is not None else "") + "\n")
        f.write("\n\\end\\\n")
    lex = os.path.join(out_dir, "lexicon.txt")
    phones = [f"P{i}" for i in range(n_phones)]
    with open(lex, "w") as f:
        for w, pr in zip(words, prons):
            f.write(w + " " + " ".join(phones[p] for p in pr) + "\n")
    return {"arpa": arpa, "lexicon": lex, "phones": phones, "words": words, "prons": [[3 + p for p in pr] for pr in prons], "corpus": corpus,
            "ngrams": [len(logp[n]) for n in range(1, order + 1)]}


if __name__ == "__main__":
    out = sys.argv[1]
    info = build(out, int(sys.argv[2]) if len(sys.argv) > 2 else 3, int(sys.argv[3]) if len(sys.argv) > 3 else 1000,
                 int(sys.argv[4]) if len(sys.argv) > 4 else 20000)
    print({k: v for k, v in info.items() if k in ("arpa", "lexicon", "ngrams")})
