"""Validation PER of this package's trainer against the reference's own PyTorch GPU path, trained identically (BASELINE.json:
"val PER vs ref").  Both arms see the same synthetic learnable corpus (nejm-brain-to-text_b200/datasets.py; there is no real
data offline), the same batch order, model size (512 features, 5 x 768 GRU, 41 classes), augmentation settings, AdamW groups,
cosine schedule and clipping (rnn_args.yaml values, schedule shortened to the run length).  They differ in what cannot be
shared: the random streams of noise / dropout / initialisation-independent kernels (Philox in our kernels, torch generators in
the reference arm) and bf16 rounding.  The reference arm is the reference's call sequence (rnn_trainer.py:511-558, 653-770)
on torch.nn.GRU / torch CTC / fused AdamW under bf16 autocast (oracle/torch_cpu_port.py restates rnn_model.py; the reference
files themselves do not exist on the GPU box).

    python tools/per_parity.py [n_batches] [val_every] [seed]
"""
import math, os, sys, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np
import torch
import torch.nn.functional as F
import b2t_pkg
import gru_ctc_oracle as O
from torch_cpu_port import PortModel, smooth_same

NB = int(sys.argv[1]) if len(sys.argv) > 1 else 300
VAL = int(sys.argv[2]) if len(sys.argv) > 2 else 100
SEED = int(sys.argv[3]) if len(sys.argv) > 3 else 10
sessions = [f"s{i}" for i in range(45)]
tmp = tempfile.mkdtemp()
args = {
    "mode": "train", "output_dir": os.path.join(tmp, "out"), "checkpoint_dir": os.path.join(tmp, "out", "checkpoint"),
    "save_best_checkpoint": False, "save_all_val_steps": False, "save_final_model": False, "save_val_metrics": False,
    "early_stopping": False, "early_stopping_val_steps": 20, "gpu_number": "0", "seed": SEED, "use_amp": True,
    "init_from_checkpoint": False, "init_checkpoint_path": None,
    "model": {"n_input_features": 512, "n_units": 768, "rnn_dropout": 0.4, "rnn_trainable": True, "n_layers": 5, "patch_size": 14,
              "patch_stride": 4, "input_network": {"input_layer_dropout": 0.2, "input_trainable": True}},
    "num_training_batches": NB, "lr_scheduler_type": "cosine", "lr_max": 0.005, "lr_min": 0.0001, "lr_decay_steps": NB,
    "lr_warmup_steps": max(10, NB // 20), "lr_max_day": 0.005, "lr_min_day": 0.0001, "lr_decay_steps_day": NB, "lr_warmup_steps_day": max(10, NB // 20),
    "beta0": 0.9, "beta1": 0.999, "epsilon": 0.1, "weight_decay": 0.001, "weight_decay_day": 0, "grad_norm_clip_value": 10,
    "batches_per_train_log": 100, "batches_per_val_step": VAL, "log_individual_day_val_PER": False, "log_val_skip_logs": False,
    "save_val_logits": False, "save_val_data": False,
    "dataset": {"data_transforms": {"white_noise_std": 1.0, "constant_offset_std": 0.2, "random_walk_std": 0.0, "random_walk_axis": -1,
                                    "static_gain_std": 0.0, "random_cut": 3, "smooth_kernel_size": 100, "smooth_data": True, "smooth_kernel_std": 2},
                "neural_dim": 512, "batch_size": 64, "n_classes": 41, "days_per_batch": 4, "seed": 1, "num_dataloader_workers": 8,
                "loader_shuffle": False, "sessions": sessions, "dataset_probability_val": [1] * len(sessions),
                "synthetic": {"T": 400, "min_len": 6, "max_len": 14, "noise": 0.6, "val_batches": 4}},
}

# ------------------------------------------------------------------ ours
T = b2t_pkg.submodule("rnn_trainer").BrainToTextDecoder_Trainer
t0 = time.time()
tr = T(args)
stats = tr.train()
ours_per, ours_loss = stats["val_PERs"], stats["val_losses"]
print(f"ours: {time.time() - t0:.1f} s, val PER {['%.4f' % p for p in ours_per]}", flush=True)
train_ds, val_ds = tr.train_dataset, tr.val_dataset
del tr
torch.cuda.empty_cache()

# ------------------------------------------------------------------ reference arm (torch GPU path)
torch.manual_seed(SEED)
np.random.seed(SEED)
m = PortModel(neural_dim=512, n_units=768, n_days=45, n_classes=41, n_layers=5, patch_size=14, patch_stride=4, rnn_dropout=0.4, input_dropout=0.2).cuda()
bias = [p for n, p in m.named_parameters() if "gru.bias" in n or "out.bias" in n]
day = [p for n, p in m.named_parameters() if "day_" in n]
other = [p for n, p in m.named_parameters() if "day_" not in n and "gru.bias" not in n and "out.bias" not in n]
opt = torch.optim.AdamW([{"params": bias, "weight_decay": 0, "group_type": "bias"}, {"params": day, "lr": args["lr_max_day"], "weight_decay": 0, "group_type": "day_layer"},
                         {"params": other, "group_type": "other"}], lr=args["lr_max"], betas=(0.9, 0.999), eps=0.1, weight_decay=0.001, fused=True)

def lr_lambda(step, min_ratio, decay, warm):
    if step < warm:
        return float(step) / float(max(1, warm))
    if step < decay:
        prog = float(step - warm) / float(max(1, decay - warm))
        return max(min_ratio, min_ratio + (1 - min_ratio) * 0.5 * (1 + math.cos(math.pi * prog)))
    return min_ratio
lam = lambda s: lr_lambda(s, args["lr_min"] / args["lr_max"], NB, args["lr_warmup_steps"])
sched = torch.optim.lr_scheduler.LambdaLR(opt, [lam, lam, lam], -1)

def validate():
    m.eval()
    ed = ln = 0
    losses = []
    with torch.no_grad(), torch.autocast(device_type="cuda", dtype=torch.bfloat16):
        for i in range(len(val_ds)):
            b = val_ds[i]
            x = b["input_features"].cuda(); n = b["n_time_steps"].cuda()
            f = F.conv1d(x.permute(0, 2, 1), taps, padding="same", groups=512).permute(0, 2, 1)
            logits = m(f, b["day_indicies"].tolist())
            adj = ((n - 14) / 4 + 1).to(torch.int32)
            loss = F.ctc_loss(logits.float().log_softmax(2).permute(1, 0, 2), b["seq_class_ids"].cuda().long(), adj, b["phone_seq_lens"].cuda().long(), blank=0, reduction="none").mean()
            losses.append(loss.item())
            lg = logits.float().cpu().numpy()
            for k in range(lg.shape[0]):
                dec = O.greedy_decode(lg[k], int(adj[k]))
                true = b["seq_class_ids"][k][: int(b["phone_seq_lens"][k])].tolist()
                ed += O.edit_distance(dec, true); ln += len(true)
    return ed / ln, float(np.mean(losses))

taps = torch.from_numpy(O.gauss_taps(2, 100)).cuda().view(1, 1, -1).repeat(512, 1, 1)
ref_per, ref_loss = [], []
t0 = time.time()
rng = np.random.RandomState(SEED)
from torch.utils.data import DataLoader
for i, b in enumerate(DataLoader(train_ds, batch_size=None, shuffle=False, num_workers=8)):
    m.train()
    opt.zero_grad()
    with torch.autocast(device_type="cuda", dtype=torch.bfloat16):
        x = b["input_features"].cuda(); n = b["n_time_steps"].cuda()
        x = x + torch.randn_like(x) * 1.0 + torch.randn(x.shape[0], 1, x.shape[2], device="cuda") * 0.2
        cut = int(rng.randint(0, 3))
        if cut > 0:
            x = x[:, cut:, :]; n = n - cut
        f = F.conv1d(x.permute(0, 2, 1), taps, padding="same", groups=512).permute(0, 2, 1)
        adj = ((n - 14) / 4 + 1).to(torch.int32)
        logits = m(f, b["day_indicies"].tolist())
        loss = F.ctc_loss(logits.log_softmax(2).permute(1, 0, 2), b["seq_class_ids"].cuda().long(), adj, b["phone_seq_lens"].cuda().long(), blank=0, reduction="none").mean()
    loss.backward()
    torch.nn.utils.clip_grad_norm_(m.parameters(), 10.0, error_if_nonfinite=True, foreach=True)
    opt.step(); sched.step()
    if i % VAL == 0 or i == NB - 1:
        p, l = validate()
        ref_per.append(p); ref_loss.append(l)
print(f"reference arm: {time.time() - t0:.1f} s, val PER {['%.4f' % p for p in ref_per]}", flush=True)
print("\n| after batch | val PER ours | val PER reference arm | difference | val CTC loss ours | val CTC loss reference arm |\n|---:|---:|---:|---:|---:|---:|")
steps = [i for i in range(NB) if i % VAL == 0 or i == NB - 1]
for s, a, b_, la, lb in zip(steps, ours_per, ref_per, ours_loss, ref_loss):
    print(f"| {s} | {a:.4f} | {b_:.4f} | {a - b_:+.4f} | {la:.3f} | {lb:.3f} |")
