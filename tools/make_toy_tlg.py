"""Build a small TLG-style decoding graph (OpenFST binary, vector/standard) for tests and decode benchmarks.

There is no OpenFST here and the reference's 3-/5-gram graphs are not available offline, so this writes the
composed graph directly instead of composing T, L and G:
  * tokens: ilabel 1 = <blk>, 2 = SIL, 3..41 = phones (tokens.txt order of ctc_compile_dict_token.sh:65);
  * every word is a chain of phone states in the "corrected" CTC topology (ctc_token_fst_corrected.py:42-57):
    enter a phone by emitting it, self-loop on repeats, optional blank state between phones (mandatory when two
    consecutive phones are equal), every state has a self-loop;
  * the word label sits on the first arc of the chain (as after determinisation of L o G), an optional SIL
    follows each word (make_lexicon_fst.pl sil_prob);
  * G is a back-off bigram: history states h(w), arcs w:w/-ln P(w|h) for seen bigrams, and an INPUT-EPSILON
    back-off arc to the unigram state (arpa-lm-compiler.cc:162-239), which exercises ProcessNonemitting;
  * final states = LM history states with the </s> cost as final weight.
"""
import struct

import numpy as np


def write_fst(path, start, finals, arcs_by_state):
    n = len(arcs_by_state)
    narcs = sum(len(a) for a in arcs_by_state)
    with open(path, "wb") as f:
        f.write(struct.pack("<i", 2125659606))
        for s in (b"vector", b"standard"):
            f.write(struct.pack("<i", len(s)) + s)
        f.write(struct.pack("<iiQqqq", 2, 0, 0, start, n, narcs))
        for s in range(n):
            f.write(struct.pack("<fq", finals[s], len(arcs_by_state[s])))
            for (il, ol, w, nx) in arcs_by_state[s]:
                f.write(struct.pack("<iifi", il, ol, w, nx))


def build(path_fst, path_words, n_words=200, seed=0, n_phones=39, min_len=2, max_len=6, bigram_frac=0.05, sil_prob=0.5):
    rng = np.random.RandomState(seed)
    BLK, SIL = 1, 2
    words, prons = [], []
    seen = set()
    while len(words) < n_words:
        L = rng.randint(min_len, max_len + 1)
        pr = tuple(int(p) for p in rng.randint(3, 3 + n_phones, size=L))
        if pr in seen:
            continue
        seen.add(pr)
        words.append(f"W{len(words)}"); prons.append(pr)
    uni = rng.dirichlet(np.ones(n_words) * 0.7)
    arcs, finals = [], []

    def new_state(final=float("inf")):
        arcs.append([]); finals.append(final)
        return len(arcs) - 1

    # LM history states: 0 = unigram (back-off) state, 1 + w = history "w"
    uni_state = new_state(final=float(-np.log(0.1)))
    hist = [new_state(final=float(-np.log(0.1))) for _ in range(n_words)]
    start = uni_state
    arcs[uni_state].append((BLK, 0, 0.0, uni_state))          # leading blanks / silence
    arcs[uni_state].append((SIL, 0, 0.0, uni_state))

    def add_word(src, w, cost):
        """chain for word w leaving LM state src with LM cost `cost`, ending in hist[w]"""
        pr = prons[w]
        prev, prev_phone = src, None
        for i, ph in enumerate(pr):
            st = new_state()
            ol = w + 1 if i == 0 else 0
            c = cost if i == 0 else 0.0
            if prev_phone is not None and prev_phone != ph:
                arcs[prev].append((ph, ol, c, st))             # direct phone -> different phone
            if i == 0:
                arcs[prev].append((ph, ol, c, st))
            else:
                b = new_state()                                # optional (or mandatory) blank between phones
                arcs[prev].append((BLK, 0, 0.0, b))
                arcs[b].append((BLK, 0, 0.0, b))
                arcs[b].append((ph, ol, c, st))
            arcs[st].append((ph, 0, 0.0, st))                  # repeats of the same phone
            prev, prev_phone = st, ph
        end = hist[w]
        # word end: optional silence / blank, then the LM history state (reached with an input-epsilon arc)
        tail = new_state()
        arcs[prev].append((BLK, 0, 0.0, tail))
        arcs[prev].append((SIL, 0, float(-np.log(sil_prob)), tail))
        arcs[tail].append((BLK, 0, 0.0, tail))
        arcs[tail].append((SIL, 0, 0.0, tail))
        arcs[tail].append((0, 0, 0.0, end))
        arcs[prev].append((0, 0, float(-np.log(1.0 - sil_prob)), end))

    for w in range(n_words):
        add_word(uni_state, w, float(-np.log(uni[w])))
    n_big = max(1, int(bigram_frac * n_words))
    for h in range(n_words):
        nxt = rng.choice(n_words, size=n_big, replace=False)
        p = rng.dirichlet(np.ones(n_big)) * 0.6
        for w, pw in zip(nxt, p):
            add_word(hist[h], int(w), float(-np.log(pw)))
        arcs[hist[h]].append((0, 0, float(-np.log(0.4)), uni_state))   # back-off (input epsilon)
    for s in range(len(arcs)):                                          # the shipped TLG has a self-loop on every state
        if not any(nx == s for (_, _, _, nx) in arcs[s]) and any(il != 0 for (il, _, _, _) in arcs[s]):
            pass
    write_fst(path_fst, start, finals, arcs)
    with open(path_words, "w") as f:
        f.write("<eps> 0\n")
        for i, w in enumerate(words):
            f.write(f"{w} {i + 1}\n")
    return {"n_states": len(arcs), "n_arcs": sum(len(a) for a in arcs), "words": words, "prons": prons}


def render_logits(prons_seq, T, C=41, seed=0, peak=6.0, noise=1.0, frames_per_phone=(2, 4)):
    """Peaky synthetic posteriors for a word sequence given as LM-order class ids (0 blank, 1 SIL, 2.. phones)."""
    rng = np.random.RandomState(seed)
    x = noise * rng.randn(T, C).astype(np.float32)
    t = rng.randint(1, 3)
    x[:, 0] += 2.0
    prev = None
    for pr in prons_seq:
        for ph in pr:
            cls = ph - 1                      # graph ilabel -> logit column (ilabel = column + 1)
            if prev == cls:
                t += 1                        # CTC needs a blank between repeated labels
            d = rng.randint(frames_per_phone[0], frames_per_phone[1] + 1)
            if t + d >= T:
                return x
            x[t:t + d, cls] += peak
            t += d + rng.randint(0, 2)
            prev = cls
        if rng.rand() < 0.5 and t + 2 < T:
            x[t:t + 2, 1] += peak             # SIL between words
            t += 2
            prev = 1
    return x


if __name__ == "__main__":
    import sys
    info = build(sys.argv[1], sys.argv[2], n_words=int(sys.argv[3]) if len(sys.argv) > 3 else 200)
    print({k: v for k, v in info.items() if k.startswith("n_")})
