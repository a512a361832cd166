"""Per-step cycle trace of the recurrence kernels at the bench configuration (bring-up / profiling aid)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
import numpy as np, torch
import b2t_pkg, bench
E = b2t_pkg.submodule("engine"); N = b2t_pkg.load()._native
from torch_cpu_port import PortModel
torch.manual_seed(0)
os.environ.setdefault("B2T_REC_CHUNKS", "1")
cfg = E.make_config(**bench.CFG)
flat = E.flat_from_state_dict(cfg, PortModel(**bench.CFG).state_dict()).cuda()
eng = E.Engine(cfg, flat, max_batch=64, max_T=400, max_label_len=64, training=True)
hb = {k: v.cuda() for k, v in bench.synth_batches(1, 1)[0].items()}
in_len = torch.full((64,), 97, dtype=torch.int32)
Tp = 97
trace = torch.zeros(2 * Tp * 8 + 148 * 8, dtype=torch.int64, device="cuda")
N.check(N.lib.b2t_debug_set_trace(eng.handle, trace.data_ptr()), "trace")
for i in range(3):
    eng.forward(hb["x"], hb["days"], training=True, smooth_mode=1, white_noise_std=1.0, offset_noise_std=0.2, seed=i, want_logits=False)
    eng.ctc_loss(hb["labels"], in_len, hb["lens"], grad_scale=1 / 64)
    eng.backward()
torch.cuda.synchronize()
skew = trace.cpu().numpy()[2 * Tp * 8:].reshape(148, 8)
tr = trace.cpu().numpy()[:2 * Tp * 8].reshape(2, Tp, 8)
fw = ["h chunk 0 staged", "all chunks staged", "MMA saw chunk 0", "MMAs issued", "accumulator done", "h_t stored", "-", "gates exchanged"]
bw = ["gather start", "partials gathered", "B operand written", "MMAs issued", "accumulator done", "partials stored"]
bw2 = ["dG chunk 0 staged", "all dG staged", "MMA saw chunk 0", "MMAs issued", "reduce start", "partials reduced", "dG published", "accumulator done"]
use_bwd2 = os.environ.get("B2T_REC_BWD2", "0") != "0"
for which, lab, names, seq in ((0, "FWD layer 0", fw, (0, 2, 3, 4, 7, 5)),
                               (1, "BWD top layer (2-D kernel)", bw2, (6, 0, 2, 3, 7)) if use_bwd2 else (1, "BWD top layer", bw, (0, 1, 2, 3, 4, 5))):
    t = tr[which].astype(np.float64)
    print(lab)
    steps = range(10, 90)
    per = np.mean([t[s + 1, seq[0]] - t[s, seq[0]] for s in steps])
    print(f"  cycles per step: {per:.0f}")
    for a, b in zip(seq[:-1], seq[1:]):
        d = np.mean([t[s, b] - t[s, a] for s in steps])
        print(f"  {names[a]:>20s} -> {names[b]:<20s}: {d:8.0f} cyc")
    d = np.mean([t[s + 1, seq[0]] - t[s, seq[-1]] for s in steps])
    print(f"  {names[seq[-1]]:>20s} -> next {names[seq[0]]:<20s}: {d:8.0f} cyc")
    if which == 0:
        print(f"  (all chunks staged - chunk 0 staged: {np.mean([t[s, 1] - t[s, 0] for s in steps]):.0f} cyc)")
if use_bwd2:
    t = tr[1].astype(np.float64); steps = range(10, 90)
    m = lambda f: np.mean([f(s) for s in steps])
    print(f"  bwd2 detail: acc done(s) -> reduce start(s+1) {m(lambda s: t[s + 1, 4] - t[s, 7]):.0f}; reduce {m(lambda s: t[s + 1, 5] - t[s + 1, 4]):.0f}; "
          f"reduced -> dG published {m(lambda s: t[s + 1, 6] - t[s + 1, 5]):.0f}; all dG staged - chunk 0 staged {m(lambda s: t[s, 1] - t[s, 0]):.0f}")

if use_bwd2 and skew[:, 0].any():
    n = int((skew[:, 0] != 0).sum())
    k = skew[:n].astype(np.float64)
    base = k[:, 0].min()
    print("bwd2 cross-CTA timing of step 40 (ns after the earliest dG publish; 1 ns = 1.965 cycles): CTA = (mb, kq)")
    print("  cta  dG published  all staged  MMAs issued  acc done  [next step: reduce done is slot 4 of step 40 = before publish]")
    for c in range(n):
        print(f"  {c:3d} ({c // 4 % 6},{c % 4})  {k[c, 0] - base:8.0f}  {k[c, 1] - base:8.0f}  {k[c, 2] - base:8.0f}  {k[c, 3] - base:8.0f}   reduce done {k[c, 4] - base:8.0f}")
