"""ARPA n-gram + pronunciation lexicon -> decoding graph (TLG) for the WFST decoder, built directly.

The reference builds its graphs with SRILM + OpenFST command-line tools (language_model/tools/fst/make_tlg.sh:29-46):
    G   = arpa2fst(lm.arpa)                                   (kaldi/lm/arpa-lm-compiler.cc)
    L   = make_lexicon_fst.pl --pron-probs lexiconp.txt 0.5 SIL  (optional silence after every word and at the start)
    T   = ctc_token_fst_corrected.py                          (CTC topology: blank loop, phone entry emits, phone self loop)
    TLG = T o min(det(L o G))
None of those tools exists offline, and general composition / determinisation is not needed for this family of machines:
every factor has a fixed shape, so the composed graph can be written down state by state.

    G states   : n-gram histories.  Word arc h --w/-ln P(w|h)--> longest existing suffix of h+w;  back-off arc
                 h --eps/-ln bow(h)--> h[1:];  </s> becomes a final cost;  <s> only names the start history.
    L o G      : per history, a prefix tree (trie) over the pronunciations of the words that have an explicit n-gram from
                 that history (deterministic on phones by construction); the word label sits on the LAST phone arc of the
                 pronunciation; after a word: optional silence (eps / -ln(1-p_sil)  or  SIL / -ln p_sil) into the trie root
                 of the next history.
    T o (L o G): every LG state q is split by the token state it was entered with: (0, q) after blank / at the start, (p, q)
                 after phone p.  Blank:  (t, q) --blk--> (0, q).  Phone loop: (p, q) --p--> (p, q).  LG arc q --j:w/c--> q' gives
                 (t, q) --j:w/c--> (j, q') for t != j (a repeated phone has to pass through blank, as in CTC).  LG epsilon arcs
                 (back-off, optional silence) are kept as input-epsilon arcs out of (0, q) and (p, q).
The result accepts the same label sequences with the same total path costs as the reference's TLG; state numbering and
the position of the costs along a path differ (no weight pushing / minimisation), which only changes where beam pruning bites.
Because OpenFST is not available here, agreement with the reference's compiled graphs is not pinned; tests check path costs
against an independent n-gram scorer (tests/test_graph_compiler.py).

Labels follow the reference's units: 0 = <eps>, 1 = <blk>, 2 = SIL, 3.. = phones (graph input label = logit column + 1).
"""
from __future__ import annotations

import math
import struct
from collections import defaultdict

LN10 = math.log(10.0)
BLK, SIL = 1, 2


# ------------------------------------------------------------------------------------------------ ARPA
def parse_arpa(path):
    """-> (order, {ngram tuple: (log10 prob, log10 back-off or 0.0)})"""
    grams, order, cur = {}, 0, 0
    with open(path, encoding="utf-8") as f:
        for raw in f:
            line = raw.strip()
            if not line or line.startswith("ngram ") or line == "\\data\\":
                continue
            if line == "\\end\\":
                break
            if line.startswith("\\") and line.endswith("-grams:"):
                cur = int(line[1:line.index("-")])
                order = max(order, cur)
                continue
            if cur == 0:
                continue
            parts = line.split()
            lp = float(parts[0])
            words = tuple(parts[1:1 + cur])
            bow = float(parts[1 + cur]) if len(parts) > 1 + cur else 0.0
            grams[words] = (lp, bow)
    return order, grams


class NgramLM:
    """Plain back-off scorer over the parsed ARPA table (also the independent checker used by the tests)."""

    def __init__(self, order, grams):
        self.order, self.grams = order, grams

    def cost(self, history, word):
        """-ln P(word | history) with back-off; history is a tuple of words (most recent last)."""
        h = tuple(history)[-(self.order - 1):] if self.order > 1 else ()
        c = 0.0
        while True:
            g = self.grams.get(h + (word,))
            if g is not None:
                return c - g[0] * LN10
            if not h:
                return math.inf
            hb = self.grams.get(h)
            if hb is not None:
                c -= hb[1] * LN10
            h = h[1:]

    def sentence_cost(self, words):
        h, c = ("<s>",), 0.0
        for w in list(words) + ["</s>"]:
            c += self.cost(h, w)
            h = (h + (w,))[-(self.order - 1):] if self.order > 1 else ()
        return c


# ------------------------------------------------------------------------------------------------ G
def build_g(order, grams):
    """-> (start history, {history: [(word, cost, next history)]}, {history: (back-off cost, shorter history)}, {history: final cost})"""
    # history states (kaldi/lm/arpa-lm-compiler.cc): the empty history plus every n-gram below the maximal order that
    # something is conditioned on
    contexts = {ng[:-1] for ng in grams if len(ng) >= 2}
    states = {()} | {ng for ng in grams if len(ng) < order and ng in contexts and ng[-1] != "</s>"}

    def dest(h):
        h = h[-(order - 1):] if order > 1 else ()
        while h not in states:
            h = h[1:]
        return h

    arcs, backoff, final = defaultdict(list), {}, {}
    for ng, (lp, bow) in grams.items():
        h, w = ng[:-1], ng[-1]
        if h not in states or w == "<s>":
            continue
        cost = -lp * LN10
        if w == "</s>":
            final[h] = cost
        else:
            arcs[h].append((w, cost, dest(ng)))
    for h in states:
        if h:
            g = grams.get(h)
            backoff[h] = (-(g[1] if g else 0.0) * LN10, dest(h[1:]) if len(h) > 1 else ())
    start = ("<s>",) if ("<s>",) in states else ()
    return start, arcs, backoff, final


# ------------------------------------------------------------------------------------------------ TLG
def read_lexicon(path):
    """lines 'WORD ph1 ph2 ...' (optionally 'WORD prob ph1 ...' as in lexiconp.txt) -> {word: [(cost, [phones])]}"""
    lex = defaultdict(list)
    with open(path, encoding="utf-8") as f:
        for line in f:
            p = line.split()
            if len(p) < 2:
                continue
            cost, phones = 0.0, p[1:]
            try:
                prob = float(p[1])
                if len(p) > 2:
                    cost, phones = -math.log(prob), p[2:]
            except ValueError:
                pass
            lex[p[0]].append((cost, phones))
    return lex


def compile_tlg(order, grams, lexicon, phone_ids, word_ids, sil_prob=0.5):
    """-> (start, finals[list], arcs_by_state[list of (ilabel, olabel, cost, next)]) in the decoder's graph conventions."""
    start_h, g_arcs, g_backoff, g_final = build_g(order, grams)
    c_sil = -math.log(sil_prob) if 0.0 < sil_prob < 1.0 else None
    c_nosil = -math.log(1.0 - sil_prob) if c_sil is not None else 0.0

    # ---- L o G: states are ("root", h), ("post", h) and trie nodes ("n", h, prefix); arcs (phone or 0, word or 0, cost, next)
    lg = defaultdict(list)
    lg_final = {}

    def root(h):
        return ("root", h)

    seen = set()
    todo = [start_h]
    while todo:
        h = todo.pop()
        if h in seen:
            continue
        seen.add(h)
        r = root(h)
        lg[r]                                         # make sure the state exists
        if h in g_final:
            lg_final[r] = g_final[h]
        if h in g_backoff:
            bc, hb = g_backoff[h]
            lg[r].append((0, 0, bc, root(hb)))
            todo.append(hb)
        for w, cost, hn in g_arcs.get(h, ()):
            if w not in lexicon or w not in word_ids:
                continue
            todo.append(hn)
            post = ("post", hn)
            if post not in lg:
                lg[post].append((0, 0, c_nosil, root(hn)))
                if c_sil is not None:
                    lg[post].append((SIL, 0, c_sil, root(hn)))
            for pc, phones in lexicon[w]:
                ids = [phone_ids[p] for p in phones]
                node = r
                for k, ph in enumerate(ids):
                    last = k == len(ids) - 1
                    if last:
                        lg[node].append((ph, word_ids[w], cost + pc, post))
                    else:
                        nxt = ("n", h, tuple(ids[:k + 1]))
                        if not any(a[0] == ph and a[3] == nxt for a in lg[node]):
                            lg[node].append((ph, 0, 0.0, nxt))
                        node = nxt
    # optional silence at the very start (make_lexicon_fst.pl: start -> loop directly or through SIL)
    lg_start = ("start",)
    lg[lg_start].append((0, 0, c_nosil, root(start_h)))
    if c_sil is not None:
        lg[lg_start].append((SIL, 0, c_sil, root(start_h)))

    # ---- T o LG: (token state t, q); t = 0 after blank / at the start, t = p after phone p
    index, finals, out = {}, [], []

    def sid(t, q):
        key = (t, q)
        i = index.get(key)
        if i is None:
            i = index[key] = len(out)
            out.append(None)
            finals.append(math.inf)
            work.append(key)
        return i

    work = []
    start = sid(0, lg_start)
    while work:
        t, q = work.pop()
        i = index[(t, q)]
        arcs = [(BLK, 0, 0.0, sid(0, q))]
        if t != 0:
            arcs.append((t, 0, 0.0, i))
        for il, ol, c, qn in lg[q]:
            if il == 0:
                arcs.append((0, ol, c, sid(t, qn)))            # LG epsilon: the token state is unchanged
            elif il != t:
                arcs.append((il, ol, c, sid(il, qn)))
        out[i] = arcs
        if q in lg_final:
            finals[i] = lg_final[q]
    return start, finals, out


def write_fst(path, start, finals, arcs_by_state):
    """OpenFST binary 'vector' / 'standard' (the format fst::Fst<StdArc>::Read and the decoder load)."""
    n = len(arcs_by_state)
    na = sum(len(a) for a in arcs_by_state)
    with open(path, "wb") as f:
        f.write(struct.pack("<i", 2125659606))
        for s in (b"vector", b"standard"):
            f.write(struct.pack("<i", len(s)) + s)
        f.write(struct.pack("<iiQqqq", 2, 0, 0, start, n, na))
        for s in range(n):
            f.write(struct.pack("<fq", finals[s], len(arcs_by_state[s])))
            for (il, ol, w, nx) in arcs_by_state[s]:
                f.write(struct.pack("<iifi", il, ol, w, nx))


def write_g_fst(order, grams, word_ids, path):
    """The n-gram model as the word acceptor `Rescore()` composes lattices with (G.fst / G_no_prune.fst after
    ReadAndPrepareLmFst: back-off arcs are eps:eps, words on both sides, </s> as final costs).  Words without an id are dropped."""
    start_h, arcs, backoff, final = build_g(order, grams)
    ids = {start_h: 0}

    def sid(h):
        if h not in ids:
            ids[h] = len(ids)
        return ids[h]

    todo, out, fin = [start_h], {}, {}
    while todo:
        h = todo.pop()
        i = sid(h)
        if i in out:
            continue
        lst = []
        for w, c, hn in arcs.get(h, ()):
            if w in word_ids:
                lst.append((word_ids[w], word_ids[w], c, sid(hn)))
                todo.append(hn)
        if h in backoff:
            bc, hb = backoff[h]
            lst.append((0, 0, bc, sid(hb)))
            todo.append(hb)
        out[i] = sorted(lst)
        fin[i] = final.get(h, math.inf)
    n = len(ids)
    write_fst(path, 0, [fin.get(i, math.inf) for i in range(n)], [out.get(i, []) for i in range(n)])
    return {"n_states": n, "n_arcs": sum(len(v) for v in out.values())}


def compile_to_files(arpa_path, lexicon_path, phones, out_fst, out_words, sil_prob=0.5):
    """phones: list of phone names in unit order (phones[0] gets label 3).  Writes TLG.fst and words.txt; returns sizes."""
    order, grams = parse_arpa(arpa_path)
    lexicon = read_lexicon(lexicon_path)
    phone_ids = {p: 3 + i for i, p in enumerate(phones)}
    phone_ids["SIL"] = SIL
    vocab = sorted({ng[-1] for ng in grams if ng[-1] not in ("<s>", "</s>") and ng[-1] in lexicon})
    word_ids = {w: i + 1 for i, w in enumerate(vocab)}
    start, finals, arcs = compile_tlg(order, grams, lexicon, phone_ids, word_ids, sil_prob)
    write_fst(out_fst, start, finals, arcs)
    with open(out_words, "w", encoding="utf-8") as f:
        f.write("<eps> 0\n")
        for w, i in word_ids.items():
            f.write(f"{w} {i}\n")
    return {"order": order, "n_states": len(arcs), "n_arcs": sum(len(a) for a in arcs), "n_words": len(vocab)}
