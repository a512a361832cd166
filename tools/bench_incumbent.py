"""Context number (not part of bench.py): the reference's own PyTorch GPU path -- cuDNN GRU + ATen CTC + fused AdamW under
bf16 autocast, as rnn_trainer.py:527-558 dispatches it -- on the same B200, same synthetic batch.  This is the GPU incumbent
SURVEY.md section 2b says to beat."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import torch
import torch.nn.functional as F
import bench
from torch_cpu_port import PortModel, smooth_same
torch.backends.cudnn.deterministic = True
torch.set_float32_matmul_precision('high')
torch.manual_seed(0)
m = PortModel(**bench.CFG).cuda()
bias = [p for n, p in m.named_parameters() if "gru.bias" in n or "out.bias" in n]
day = [p for n, p in m.named_parameters() if "day_" in n]
other = [p for n, p in m.named_parameters() if "day_" not in n and "gru.bias" not in n and "out.bias" not in n]
opt = torch.optim.AdamW([{"params": bias, "weight_decay": 0}, {"params": day, "weight_decay": 0}, {"params": other}], lr=2.5e-3, betas=(0.9, 0.999),
                        eps=0.1, weight_decay=1e-3, fused=True)
hb = {k: v.cuda() for k, v in bench.synth_batches(1, 1)[0].items()}
k = torch.from_numpy(__import__("gru_ctc_oracle").gauss_taps(2, 100)).cuda().view(1, 1, -1).repeat(512, 1, 1)
def step():
    opt.zero_grad()
    with torch.autocast(device_type="cuda", dtype=torch.bfloat16):
        f = hb["x"] + torch.randn_like(hb["x"]) + torch.randn(64, 1, 512, device="cuda") * 0.2
        f = F.conv1d(f.permute(0, 2, 1), k, padding="same", groups=512).permute(0, 2, 1)
        adj = ((hb["n_steps"] - 14) / 4 + 1).to(torch.int32)
        logits = m(f, hb["days"].tolist())
        loss = F.ctc_loss(logits.log_softmax(2).permute(1, 0, 2), hb["labels"].long(), adj, hb["lens"].long(), blank=0, reduction="none").mean()
    loss.backward()
    torch.nn.utils.clip_grad_norm_(m.parameters(), 10.0, error_if_nonfinite=True, foreach=True)
    opt.step()
    return loss
for _ in range(5): step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
n = 20
for _ in range(n): step()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / n
print({"incumbent": "torch cuDNN-GRU + ATen CTC + fused AdamW (bf16 autocast, eager)", "ms_per_step": ms, "trials_per_s": 64 / ms * 1e3, "loss": float(step())})
