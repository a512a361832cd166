"""Per-step cycle trace of the whole-stack recurrence kernels (gru_stack.cuh) at the bench configuration.
   B2T_TRACE_CTA_FWD / B2T_TRACE_CTA_BWD select the traced CTA (24 CTAs per layer at the bench shape)."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import b2t_pkg, bench
E = b2t_pkg.submodule("engine"); N = b2t_pkg.load()._native
torch.manual_seed(0)
cfg = E.make_config(**dict(bench.CFG, rnn_dropout=float(os.environ.get('B2T_TRACE_DROPOUT', bench.CFG['rnn_dropout']))))
flat = (torch.randn(E.param_elems(cfg)) * 0.03).cuda()
eng = E.Engine(cfg, flat, max_batch=64, max_T=400, max_label_len=64, training=True)
hb = {k: v.cuda() for k, v in bench.synth_batches(1, 1)[0].items()}
in_len = torch.full((64,), 97, dtype=torch.int32)
Tp = 97
trace = torch.zeros(2 * Tp * 8 + 128 + 148 * 8, dtype=torch.int64).pin_memory() if os.environ.get("B2T_TRACE_PINNED") else torch.zeros(2 * Tp * 8 + 128 + 148 * 8, dtype=torch.int64, device="cuda")
N.check(N.lib.b2t_debug_set_trace(eng.handle, trace.data_ptr()), "trace")
try:
    for i in range(3):
        eng.forward(hb["x"], hb["days"], training=True, smooth_mode=1, white_noise_std=1.0, offset_noise_std=0.2, seed=i, want_logits=False)
        eng.ctc_loss(hb["labels"], in_len, hb["lens"], grad_scale=1 / 64)
        eng.backward()
    torch.cuda.synchronize()
except Exception as ex:
    print("FAILED:", str(ex).split("\n")[0])
    d = trace[2 * Tp * 8 + 128:].view(148, 8)
    for blk in range(148):
        if d[blk, 0] != 0:
            print(f"block {blk}: code {hex(int(d[blk, 0]))} {d[blk, 1:6].tolist()}")
    print("signaller progress per block:", " ".join(f"{blk}:{int(d[blk, 5]):x}" for blk in range(120)))
    print("last phases (A, B) per block:", " ".join(f"{blk}:{int(d[blk, 6]):x}/{int(d[blk, 7]):x}" for blk in range(120)))
    sys.exit(1)
tr = trace.cpu().numpy()[:2 * Tp * 8].reshape(2, Tp, 8).astype(np.float64)
fw = ["chunk0 staged", "all staged", "MMA saw chunk0", "MMAs issued", "acc done", "h_t stored"]
bw = ["dG chunk0 staged", "all dG staged", "MMA saw chunk0", "MMAs issued", "reduce start", "partials reduced", "dG published", "acc done"]
print("FWD cta", os.environ.get("B2T_TRACE_CTA_FWD", "0"), "(group 0 of the CTA)")
t = tr[0]; steps = range(10, 90)
print(f"  cycles per step: {np.mean([t[s + 1, 5] - t[s, 5] for s in steps]):.0f}   (first..last step total {(t[96, 5] - t[0, 5]):.0f} cycles = {(t[96,5]-t[0,5])/1.965e3:.0f} us)")
for a, b in ((5, 0), (0, 1), (0, 2), (2, 3), (3, 4), (4, 5)):
    nxt = a == 5
    d = np.mean([t[s + 1, b] - t[s, a] if nxt else t[s, b] - t[s, a] for s in steps])
    print(f"  {fw[a]:>16s} -> {('next ' if nxt else '') + fw[b]:<22s}: {d:8.0f}")
print("  per-step deltas (h_t stored), steps 0..96:", " ".join(f"{t[s+1,5]-t[s,5]:.0f}" for s in range(0, 96, 6)))
print("BWD cta", os.environ.get("B2T_TRACE_CTA_BWD", "top layer"), "(group 0 of the CTA)")
t = tr[1]
print(f"  cycles per step: {np.mean([t[s + 1, 6] - t[s, 6] for s in steps]):.0f}   (total {(t[96, 6] - t[0, 6]):.0f} cycles = {(t[96,6]-t[0,6])/1.965e3:.0f} us)")
for a, b, nxt in ((6, 0, False), (0, 1, False), (0, 2, False), (2, 3, False), (3, 7, False), (7, 5, True), (5, 6, False)):
    d = np.mean([t[s + 1, b] - t[s, a] if nxt else t[s, b] - t[s, a] for s in steps])
    print(f"  {bw[a]:>16s} -> {('next ' if nxt else '') + bw[b]:<22s}: {d:8.0f}")
print("  per-step deltas (dG published), steps 0..96:", " ".join(f"{t[s+1,6]-t[s,6]:.0f}" for s in range(0, 96, 6)))

# cross-layer timing: %globaltimer stamps of the first CTA of every layer (kernel start, steps 0 / 1 / T/2 / T-1 stored)
tail = trace.cpu().numpy()[2 * Tp * 8:2 * Tp * 8 + 128].reshape(2, 8, 8).astype(np.float64)
for d, name in ((0, "FWD"), (1, "BWD")):
    t0 = min(tail[d, l, 0] for l in range(5))
    print(name, "per layer (us after the first CTA's start): start, step0, step1, step T/2, step T-1")
    order = range(5) if d == 0 else range(4, -1, -1)
    for l in order:
        print(f"  layer {l}: " + " ".join(f"{(tail[d, l, k] - t0) / 1e3:8.1f}" for k in range(5)))
