"""Host cost of rebuilding the engine's plans when (B, T) changes between steps (ragged real data: every step)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import b2t_pkg, bench
E = b2t_pkg.submodule("engine")
cfg = E.make_config(**bench.CFG)
torch.manual_seed(0)
flat = (torch.randn(E.param_elems(cfg)) * 0.03).cuda()
eng = E.Engine(cfg, flat, max_batch=64, max_T=400, max_label_len=64, training=True)
xs = {T: torch.randn(64, T, 512, device="cuda") for T in (400, 384, 368)}
days = torch.zeros(64, dtype=torch.int32)
def run(Ts, n=30):
    for i in range(6):
        eng.forward(xs[Ts[i % len(Ts)]], days, training=True, smooth_mode=1, want_logits=False)
    torch.cuda.synchronize()
    host = 0.0
    t0 = time.perf_counter()
    for i in range(n):
        h0 = time.perf_counter()
        eng.forward(xs[Ts[i % len(Ts)]], days, training=True, smooth_mode=1, want_logits=False)
        host += time.perf_counter() - h0
    torch.cuda.synchronize()
    return host / n * 1e3, (time.perf_counter() - t0) / n * 1e3
print("same T every step : host %.3f ms per forward call, wall %.3f ms" % run([400]))
print("T changes each step: host %.3f ms per forward call, wall %.3f ms" % run([400, 384, 368]))
