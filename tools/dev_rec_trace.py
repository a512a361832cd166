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
trace = torch.zeros(2 * Tp * 8, dtype=torch.int64, device="cuda")
N.check(N.lib.b2t_debug_set_trace(eng.handle, trace.data_ptr()), "trace")
for i in range(3):
    eng.forward(hb["x"], hb["days"], training=True, smooth_mode=1, white_noise_std=1.0, offset_noise_std=0.2, seed=i, want_logits=False)
    eng.ctc_loss(hb["labels"], in_len, hb["lens"], grad_scale=1 / 64)
    eng.backward()
torch.cuda.synchronize()
tr = trace.cpu().numpy().reshape(2, Tp, 8)
fw = ["h chunk 0 staged", "all chunks staged", "MMA saw chunk 0", "MMAs issued", "accumulator done", "h_t stored", "-", "gates exchanged"]
bw = ["gather start", "partials gathered", "B operand written", "MMAs issued", "accumulator done", "partials stored"]
for which, lab, names, seq in ((0, "FWD layer 0", fw, (0, 2, 3, 4, 7, 5)), (1, "BWD top layer", bw, (0, 1, 2, 3, 4, 5))):
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
