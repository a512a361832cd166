"""Short workload for ncu captures: a few layer-0 input-projection GEMMs, then two full training steps."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import torch
import b2t_pkg, bench
E = b2t_pkg.submodule("engine")
from torch_cpu_port import PortModel
torch.manual_seed(0)
a = torch.randn(97 * 64, 7168, device="cuda").to(torch.bfloat16); w = torch.randn(2304, 7168, device="cuda").to(torch.bfloat16)
for _ in range(3):
    E.gemm_bf16(a, w)
cfg = E.make_config(**bench.CFG)
eng = E.Engine(cfg, E.flat_from_state_dict(cfg, PortModel(**bench.CFG).state_dict()).cuda(), max_batch=64, max_T=400, max_label_len=64, training=True)
hb = {k: v.cuda() for k, v in bench.synth_batches(1, 1)[0].items()}
in_len = torch.full((64,), 97, dtype=torch.int32)
for i in range(2):
    eng.forward(hb["x"], hb["days"], training=True, smooth_mode=1, white_noise_std=1.0, offset_noise_std=0.2, seed=i, want_logits=False)
    eng.ctc_loss(hb["labels"], in_len, hb["lens"], grad_scale=1 / 64)
    eng.backward()
    eng.optimizer_step([1e-3] * 3, [0, 0, 1e-3], 0.9, 0.999, 0.1, 10.0)
torch.cuda.synchronize()
print("done")
