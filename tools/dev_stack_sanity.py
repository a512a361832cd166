"""Bring-up: single-layer stack kernels (no gated GEMMs -> runs under compute-sanitizer, which serialises kernels)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, b2t_pkg
E = b2t_pkg.submodule("engine")
B = int(os.environ.get("SB", "64")); H = int(os.environ.get("SH", "256")); train = os.environ.get("STRAIN", "0") == "1"
Dn = int(os.environ.get("SD", "32")); Tn = int(os.environ.get("ST", "74"))
cfg = E.make_config(Dn, H, 1, 3, 41, 14, 4, 0.0, 0.0)
torch.manual_seed(0)
flat = (torch.randn(E.param_elems(cfg)) * 0.05).cuda()
eng = E.Engine(cfg, flat, max_batch=B, max_T=Tn, max_label_len=8, training=train)
x = torch.randn(B, Tn, Dn, device="cuda")
days = torch.zeros(B, dtype=torch.int32)
N = b2t_pkg.load()._native
dbg = torch.zeros(4096 + 148 * 32 * 4, dtype=torch.int64).pin_memory()
if os.environ.get("SDBG", "0") == "1":
    N.check(N.lib.b2t_debug_set_trace(eng.handle, dbg.data_ptr()), "trace")
lg, _ = eng.forward(x, days, training=train, smooth_mode=1)
torch.cuda.synchronize()
print("forward ok", float(lg.abs().mean()))
if train:
    labels = torch.randint(1, 41, (B, 4), dtype=torch.int32); il = torch.full((B,), lg.shape[1], dtype=torch.int32); tl = torch.full((B,), 3, dtype=torch.int32)
    eng.ctc_loss(labels, il, tl, grad_scale=1.0 / B)
    eng.backward()
    try:
        torch.cuda.synchronize()
    except Exception as ex:
        print("FAILED:", str(ex).split("\n")[0])
        d = dbg[4096:].view(148, 32, 4)
        for blk in range(148):
            for w in range(32):
                if d[blk, w, 0] != 0:
                    print(f"block {blk} warp {w}: {hex(int(d[blk, w, 0]))} {d[blk, w, 1:].tolist()}")
        sys.exit(1)
    print("backward ok", float(eng.grads[:eng.n_params].abs().mean()))
