"""Bring-up: prefix beam search at wide second beams vs the oracle."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import b2t_pkg
import decoder_util as D
LM = b2t_pkg.submodule("lm_decoder")
for sb in (100, 200, 300, 400, 500):
    rng = np.random.RandomState(40 + sb)
    x = rng.randn(1, 60, 41).astype(np.float32) * 1.2
    x[..., 0] += 1.5
    lp = x - np.log(np.exp(x).sum(-1, keepdims=True))
    ours = LM.ctc_prefix_beam_search(lp, first_beam_size=10, second_beam_size=sb)[0]
    ref = D.prefix_search(lp[0], 10, sb)
    ids_ok = [r[0] for r in ours] == [r[0] for r in ref]
    k = next((i for i, (a, b) in enumerate(zip(ours, ref)) if a[0] != b[0]), -1)
    sc_ok = all(abs(a[1] - b[1]) < 1e-4 and abs(a[2] - b[2]) < 1e-4 for a, b in zip(ours, ref))
    tm_ok = [r[3] for r in ours] == [r[3] for r in ref]
    kt = next((i for i, (a, b) in enumerate(zip(ours, ref)) if a[3] != b[3]), -1)
    print(sb, "len", len(ours), len(ref), "ids", ids_ok, k, "scores", sc_ok, "times", tm_ok, kt)

    if not ids_ok:
        so = {tuple(r[0]): r for r in ours}; sr = {tuple(r[0]): r for r in ref}
        print("  set equal:", set(so) == set(sr), "only ours", len(set(so) - set(sr)), "only ref", len(set(sr) - set(so)))
        for i in range(max(0, k - 2), min(len(ours), k + 4)):
            print("  ", i, ours[i][0], round(ours[i][1], 5), "|", ref[i][0], round(ref[i][1], 5))
        bad = [(x, so[x][1], sr[x][1]) for x in set(so) & set(sr) if abs(so[x][1] - sr[x][1]) > 1e-4]
        print("  shared hyps with different scores:", len(bad), bad[:3])
