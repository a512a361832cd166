"""Dump the event timeline of one training step (all side streams) at the bench configuration."""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import torch
import b2t_pkg, bench
E = b2t_pkg.submodule("engine"); N = b2t_pkg.load()._native
from torch_cpu_port import PortModel
torch.manual_seed(0)
cfg = E.make_config(**bench.CFG)
eng = E.Engine(cfg, E.flat_from_state_dict(cfg, PortModel(**bench.CFG).state_dict()).cuda(), max_batch=64, max_T=400, max_label_len=64, training=True)
hb = {k: v.cuda() for k, v in bench.synth_batches(1, 1)[0].items()}
in_len = torch.full((64,), 97, dtype=torch.int32)
def step(i):
    eng.forward(hb["x"], hb["days"], training=True, smooth_mode=1, white_noise_std=1.0, offset_noise_std=0.2, seed=i, want_logits=False)
    eng.ctc_loss(hb["labels"], in_len, hb["lens"], grad_scale=1 / 64)
    eng.backward()
    eng.optimizer_step([1e-3] * 3, [0, 0, 1e-3], 0.9, 0.999, 0.1, 10.0)
for i in range(4): step(i)
torch.cuda.synchronize()
N.lib.b2t_debug_timeline(1)
step(9)
buf = C.create_string_buffer(1 << 16)
N.lib.b2t_debug_dump_timeline(buf, 1 << 16)
N.lib.b2t_debug_timeline(0)
rows = [l.split() for l in buf.value.decode().strip().split("\n")]
print("lane task start_us end_us dur_us")
for lane, name, a, b in sorted(rows, key=lambda r: float(r[2])):
    print(f"{lane} {name:8s} {float(a):8.1f} {float(b):8.1f} {float(b) - float(a):7.1f}")
