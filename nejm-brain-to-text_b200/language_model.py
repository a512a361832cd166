"""Decode-side callers of the n-gram decoder (rows N4 and B5), mirroring
``language_model/language-model-standalone.py``:

  * n-best post-processing: ``get_string_differences`` (:272-310), ``remove_punctuation`` (:313-324), ``augment_nbest`` (:327-411) and
    the score fusion of ``gpt2_lm_decode`` (:165-251): total = acoustic_scale * ac + (1 - alpha) * ngram + alpha * llm.  The LLM
    itself (OPT-6.7b through HF transformers, :91-161) is library code outside the hot path: it plugs in as a callable
    ``llm_scorer(hypotheses, length_penalty) -> scores``; without one the LLM term is zero, exactly like the reference with do_opt=0.
  * ``LanguageModelServer``: the body of the reference's Redis main loop (:516-790) as methods (reset / decode chunk / finalize /
    update parameters) around the GPU WFST decoder (lm_decoder.BrainSpeechDecoder).
  * ``LoopbackRedis``: an in-process stand-in for the handful of Redis calls both sides use (xadd / xread / get / set / flushall /
    ping / time / xlen) plus ``serve``, which drives a LanguageModelServer from those streams on a thread.  The reference's client
    helpers (evaluate_model_helpers.py:136-297, mirrored in this package) then run unmodified, without a redis-server process,
    over the same stream names and payloads.
"""
from __future__ import annotations

import re
import threading
import time
from typing import Callable, List, Optional, Sequence

import numpy as np


# ------------------------------------------------------------------------------------------------ string alignment
def get_string_differences(cue: str, decoder_output: str):
    """Levenshtein alignment of ``decoder_output`` against ``cue`` with the reference's tie order (insertion, then deletion, then
    substitution; language-model-standalone.py:272-310).  Returns (cost, path, indices_to_highlight); the path holds, per word of
    decoder_output, its index when it matched or 'R' / 'D' (insertions removed)."""
    out_w, cue_w = decoder_output.split(), cue.split()
    n, m = len(out_w), len(cue_w)
    cost = [[0] * (m + 1) for _ in range(n + 1)]
    step = [[None] * (m + 1) for _ in range(n + 1)]          # how (i, j) is reached: 'M', 'I', 'D', 'R'
    for j in range(1, m + 1):
        cost[0][j], step[0][j] = j, 'I'
    for i in range(1, n + 1):
        cost[i][0], step[i][0] = i, 'D'
        for j in range(1, m + 1):
            if out_w[i - 1] == cue_w[j - 1]:
                cost[i][j], step[i][j] = cost[i - 1][j - 1], 'M'
            else:
                ins, dele, sub = cost[i][j - 1], cost[i - 1][j], cost[i - 1][j - 1]
                if ins <= dele and ins <= sub:
                    cost[i][j], step[i][j] = ins + 1, 'I'
                elif dele <= ins and dele <= sub:
                    cost[i][j], step[i][j] = dele + 1, 'D'
                else:
                    cost[i][j], step[i][j] = sub + 1, 'R'
    path = []
    i, j = n, m
    while i > 0 or j > 0:
        s = step[i][j]
        if s == 'M':
            path.append(i - 1); i -= 1; j -= 1
        elif s == 'I':
            path.append('I'); j -= 1
        elif s == 'D':
            path.append('D'); i -= 1
        else:
            path.append('R'); i -= 1; j -= 1
    path.reverse()
    path = [p for p in path if p != 'I']
    hl, cur = [], 0
    for label, word in zip(path, out_w):
        if label in ('R', 'D'):
            hl.append((cur, cur + len(word)))
        cur += len(word) + 1
    return cost[n][m], path, hl


def remove_punctuation(sentence: str) -> str:
    sentence = re.sub(r'[^a-zA-Z\- \']', '', sentence)
    sentence = sentence.replace('- ', ' ').lower()
    sentence = sentence.replace('--', '').lower()
    sentence = sentence.replace(" '", "'").lower()
    sentence = sentence.strip()
    return ' '.join(sentence.split())


# ------------------------------------------------------------------------------------------------ n-best augmentation
def augment_nbest(nbest, top_candidates_to_augment=20, acoustic_scale=0.3, score_penalty_percent=0.01):
    """Enlarge the n-best list by swapping the words in which two equally long candidates differ
    (language-model-standalone.py:327-411).  nbest: [[sentence, ac_score, lm_score], ...]; returns the same, best total first."""
    sent = [x[0].strip() for x in nbest]
    ac = [x[1] for x in nbest]
    lm = [x[2] for x in nbest]
    tot = [acoustic_scale * a + l for a, l in zip(ac, lm)]
    order = np.argsort(tot)[::-1]
    sent = [sent[i] for i in order]; ac = [ac[i] for i in order]; lm = [lm[i] for i in order]; tot = [tot[i] for i in order]
    new_s, new_a, new_l, new_t = [], [], [], []

    def consider(s, i1, i2):
        if s not in sent and s not in new_s:
            ma, ml = np.mean([ac[i1], ac[i2]]), np.mean([lm[i1], lm[i2]])
            new_s.append(s)
            new_a.append(ma - score_penalty_percent * np.abs(ma))
            new_l.append(ml - score_penalty_percent * np.abs(ml))
            new_t.append(acoustic_scale * new_a[-1] + new_l[-1])

    for i1 in range(int(np.min([len(sent) - 1, top_candidates_to_augment]))):
        w1 = sent[i1].split()
        for i2 in range(i1 + 1, int(np.min([len(sent), top_candidates_to_augment]))):
            w2 = sent[i2].split()
            if len(w1) != len(w2):
                continue
            _, p1, _ = get_string_differences(sent[i1], sent[i2])
            _, p2, _ = get_string_differences(sent[i2], sent[i1])
            r1s = [i for i, p in enumerate(p2) if p == 'R']
            r2s = [i for i, p in enumerate(p1) if p == 'R']
            for r1, r2 in zip(r1s, r2s):
                n1, n2 = w1.copy(), w2.copy()
                n1[r1] = w2[r2]
                n2[r2] = w1[r1]
                consider(' '.join(n1), i1, i2)
                consider(' '.join(n2), i1, i2)
    sent += new_s; ac += new_a; lm += new_l; tot += new_t
    order = np.argsort(tot)[::-1]
    return [[sent[i], ac[i], lm[i]] for i in order]


# ------------------------------------------------------------------------------------------------ score fusion
def fuse_nbest_scores(nbest, acoustic_scale, length_penalty, alpha, llm_scorer: Optional[Callable[[List[str], float], Sequence[float]]] = None,
                      returnConfidence=False, current_context_str=None):
    """The candidate clean-up, LLM call and score fusion of gpt2_lm_decode (language-model-standalone.py:165-251).
    ``llm_scorer(hypotheses, length_penalty)`` returns one log-likelihood per hypothesis (the reference's rescore_with_gpt2 bound to
    its model); when it raises, the reference's fall-back ladder applies (five chunks, then zeros); None scores zeros."""
    hyps, acs, old = [], [], []
    for out in nbest:
        hyp = out[0].strip()
        if len(hyp) == 0:
            continue
        if current_context_str is not None and len(current_context_str.split()) > 0:
            hyp = current_context_str + ' ' + hyp
        hyp = hyp.replace('>', '').replace('  ', ' ').replace(' ,', ',').replace(' .', '.').replace(' ?', '?')
        hyps.append(hyp); acs.append(out[1]); old.append(out[2])
    if len(hyps) == 0:
        return ("", []) if not returnConfidence else ("", [], 0.)
    acs, old = np.array(acs), np.array(old)
    if llm_scorer is None:
        new = np.zeros(len(hyps))
    else:
        try:
            new = np.array(llm_scorer(hyps, length_penalty))
        except Exception:  # noqa: BLE001  (VRAM exhaustion in the reference: retry in five chunks, then give up on the LLM term)
            try:
                new, step = [], int(np.ceil(len(hyps) / 5))
                for i in range(0, len(hyps), step):
                    new.extend(llm_scorer(hyps[i:i + step], length_penalty))
                new = np.array(new)
            except Exception:  # noqa: BLE001
                new = np.zeros(len(hyps))
    if current_context_str is not None and len(current_context_str.split()) > 0:
        hyps = [h[(len(current_context_str) + 1):] for h in hyps]
    total = (acoustic_scale * acs) + ((1 - alpha) * old) + (alpha * new)
    best = int(np.argmax(total))
    nbest_out = [';'.join(map(str, [nbest[i][0], nbest[i][1], nbest[i][2], new[i], total[i]]))
                 for i in range(int(np.min((len(nbest), len(new), len(total)))))]
    if not returnConfidence:
        return hyps[best], nbest_out
    t = total - np.max(total)
    pr = np.exp(t)
    return hyps[best], nbest_out, pr[best] / np.sum(pr)


# ------------------------------------------------------------------------------------------------ LM server
class LanguageModelServer:
    """State and handlers of the reference's LM process (language-model-standalone.py:421-790) around one decoder."""

    def __init__(self, lm_path=None, *, decoder=None, max_active=7000, min_active=200, beam=17.0, lattice_beam=8.0, acoustic_scale=0.3,
                 ctc_blank_skip_threshold=1.0, length_penalty=0.0, nbest=100, blank_penalty=90.0, alpha=0.55, do_opt=0, rescore=0,
                 top_candidates_to_augment=20, score_penalty_percent=0.01, llm_scorer=None, max_frames=2048):
        import os
        from . import lm_decoder
        self.lm_decoder = lm_decoder
        self.p = dict(lm_path=lm_path, max_active=int(max_active), min_active=int(min_active), beam=float(beam), lattice_beam=float(lattice_beam),
                      acoustic_scale=float(acoustic_scale), ctc_blank_skip_threshold=float(ctc_blank_skip_threshold),
                      length_penalty=float(length_penalty), nbest=int(nbest), blank_penalty=float(blank_penalty), alpha=float(alpha),
                      do_opt=int(do_opt), rescore=int(rescore), top_candidates_to_augment=int(top_candidates_to_augment),
                      score_penalty_percent=float(score_penalty_percent))
        self.llm_scorer = llm_scorer
        if decoder is None:
            # build_lm_decoder (language-model-standalone.py:18-62): TLG.fst is mandatory, G / G_no_prune optional (rescoring)
            tlg = os.path.join(lm_path, 'TLG.fst')
            if not os.path.exists(tlg):
                raise ValueError('TLG file not found at {}'.format(tlg))
            g = os.path.join(lm_path, 'G.fst')
            g_rescore = os.path.join(lm_path, 'G_no_prune.fst')
            g = g if os.path.exists(g) else ""
            g_rescore = g_rescore if os.path.exists(g_rescore) else ""
            res = lm_decoder.DecodeResource(tlg, g, g_rescore, os.path.join(lm_path, 'words.txt'), "")
            decoder = lm_decoder.BrainSpeechDecoder(res, self._opts(), max_frames=max_frames)
        self.decoder = decoder

    def _opts(self):
        p = self.p
        return self.lm_decoder.DecodeOptions(p['max_active'], p['min_active'], p['beam'], p['lattice_beam'], p['acoustic_scale'],
                                             p['ctc_blank_skip_threshold'], p['length_penalty'], p['nbest'])

    def args(self):
        return dict(self.p)

    def reset(self):
        self.decoder.Reset()

    def decode(self, logits) -> str:
        """One chunk of logits [T, 41] (LM class order); returns the partial best sentence (:760-786)."""
        logits = np.asarray(logits, dtype=np.float32).reshape(-1, 41)
        self.lm_decoder.DecodeNumpy(self.decoder, logits, np.zeros_like(logits), np.log(self.p['blank_penalty']))
        r = self.decoder.result()
        return r[0].sentence if len(r) > 0 else ''

    def finalize(self, current_context_str=''):
        """FinishDecoding (+ Rescore) + n-best augmentation + score fusion (:573-660).  Returns (decoded_final, nbest_strings)."""
        p = self.p
        self.decoder.FinishDecoding()
        if p['rescore']:
            self.decoder.Rescore()
        results = self.decoder.result()
        nbest_out = [[d.sentence, d.ac_score, d.lm_score] for d in results]
        if p['nbest'] > 1 and len(nbest_out) > 0:
            nbest_out = augment_nbest(nbest_out, p['top_candidates_to_augment'], p['acoustic_scale'], p['score_penalty_percent'])
        if p['do_opt'] and self.llm_scorer is not None:
            decoded_final, nbest_redis, _ = fuse_nbest_scores(nbest_out, p['acoustic_scale'], p['length_penalty'], p['alpha'], self.llm_scorer,
                                                              returnConfidence=True, current_context_str=current_context_str)
        elif len(results) > 0:
            decoded_final = results[0].sentence
            nbest_redis = [';'.join(map(str, [s.strip(), a, l, 0.0, p['acoustic_scale'] * a + l])) for s, a, l in nbest_out]
        else:
            decoded_final, nbest_redis = '', []
        return decoded_final, nbest_redis

    def update_params(self, **kw):
        """remote_lm_update_params (:663-738); decoder options go through SetOpt like update_ngram_params (:66-88)."""
        for k, v in kw.items():
            if k in self.p and k != 'lm_path':
                self.p[k] = type(self.p[k])(float(v)) if isinstance(self.p[k], (int, float)) else v
        self.decoder.SetOpt(self._opts())
        return self.args()


# ------------------------------------------------------------------------------------------------ in-process Redis stand-in
class LoopbackRedis:
    """The subset of redis.Redis that evaluate_model.py / evaluate_model_helpers.py / language-model-standalone.py use, in process:
    streams with monotonically increasing ids, blocking xread, a key-value store.  Values come back as bytes like from redis-py."""

    def __init__(self):
        self._cv = threading.Condition()
        self._streams = {}
        self._kv = {}
        self._seq = 0

    @staticmethod
    def _b(v):
        return v if isinstance(v, bytes) else str(v).encode()

    def ping(self):
        return True

    def time(self):
        t = time.time()
        return (int(t), int((t - int(t)) * 1e6))

    def flushall(self):
        with self._cv:
            self._streams.clear(); self._kv.clear()

    def set(self, k, v):
        with self._cv:
            self._kv[k] = self._b(v)

    def get(self, k):
        with self._cv:
            return self._kv.get(k)

    def xlen(self, stream):
        with self._cv:
            return len(self._streams.get(stream, []))

    def xadd(self, stream, fields):
        with self._cv:
            self._seq += 1
            ms = int(time.time() * 1000)
            eid = f"{ms}-{self._seq}".encode()
            self._streams.setdefault(stream, []).append((eid, (ms, self._seq), {self._b(k): self._b(v) for k, v in fields.items()}))
            self._cv.notify_all()
            return eid

    @staticmethod
    def _key(last):
        if isinstance(last, bytes):
            last = last.decode()
        if isinstance(last, str):
            if last == '$':
                return None
            a, _, b = last.partition('-')
            return (int(a), int(b) if b else 0)
        return (int(last), 1 << 62)              # a millisecond timestamp: everything stamped later

    def xread(self, streams, count=None, block=None):
        deadline = None if block is None else time.time() + block / 1000.0
        with self._cv:
            while True:
                out = []
                for name, last in streams.items():
                    k = self._key(last)
                    items = [(eid, data) for eid, key, data in self._streams.get(name, []) if k is not None and key > k]
                    if items:
                        out.append([name.encode() if isinstance(name, str) else name, items[:count] if count else items])
                if out or block is None:
                    return out
                left = deadline - time.time()
                if left <= 0:
                    return []
                self._cv.wait(timeout=left)

    # -------------------------------------------------------------------------------------------- server side
    def serve(self, server: LanguageModelServer, *, input_stream='remote_lm_input', partial_output_stream='remote_lm_output_partial',
              final_output_stream='remote_lm_output_final') -> threading.Thread:
        """Run the reference's main loop (:516-790) against this object on a daemon thread; ``stop_serving()`` ends it."""
        self._stop = threading.Event()
        seen = {k: self.xadd('__boot__', {'k': k}) for k in ('reset', 'finalize', 'update', 'logits')}

        def newest(stream, key):
            r = self.xread({stream: seen[key]}, count=1, block=None)
            if not r:
                return None
            eid, data = r[0][1][-1]
            seen[key] = eid
            return data

        def loop():
            while not self._stop.is_set():
                if self.xlen('remote_lm_args') == 0:
                    self.xadd('remote_lm_args', {k: ('' if v is None else v) for k, v in server.args().items()})
                if newest('remote_lm_reset', 'reset') is not None:
                    server.reset()
                    self.xadd('remote_lm_done_resetting', {'done': 1})
                    continue
                if newest('remote_lm_finalize', 'finalize') is not None:
                    ctx = self.get('contextual_decoding_current_context')
                    ctx = ctx.decode().strip() if ctx is not None else ''
                    final, nbest = server.finalize(ctx)
                    if server.p['nbest'] > 1:
                        self.xadd(final_output_stream, {'lm_response_final': final, 'scoring': ';'.join(nbest), 'context_str': ctx})
                    else:
                        self.xadd(final_output_stream, {'lm_response_final': final})
                    self.xadd('remote_lm_done_finalizing', {'done': 1})
                    continue
                upd = newest('remote_lm_update_params', 'update')
                if upd is not None:
                    args = server.update_params(**{k.decode(): v.decode() for k, v in upd.items()})
                    self.xadd('remote_lm_args', {k: ('' if v is None else v) for k, v in args.items()})
                    self.xadd('remote_lm_done_updating_params', {'done': 1})
                    continue
                r = self.xread({input_stream: seen['logits']}, count=1, block=20)
                if r:
                    eid, data = r[0][1][-1]
                    seen['logits'] = eid
                    partial = server.decode(np.frombuffer(data[b'logits'], dtype=np.float32))
                    self.xadd(partial_output_stream, {'lm_response_partial': partial})

        th = threading.Thread(target=loop, daemon=True)
        th.start()
        self._thread = th
        return th

    def stop_serving(self):
        if getattr(self, '_stop', None) is not None:
            self._stop.set()
            self._thread.join(timeout=5)
