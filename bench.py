#!/usr/bin/env python
"""Headline benchmark: trials/sec of the bs-64 GRU+CTC training step (512-feature x 400-bin synthetic
trials, bf16 tensor-core GEMMs) on N B200s of one node.  See the contract in the task statement.

    python bench.py --gpus 1 --steps 20 --warmup 5
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference        # the reference's PyTorch-CPU path (oracle/torch_cpu_port.py) on host cores

One step = augmentation(noise, cut) + Gaussian smoothing + day layer + 5-layer GRU + head + log-softmax/CTC
+ full backward (BPTT) + gradient all-reduce (N>1) + clip + AdamW, on one batch of 64 trials per GPU
(weak scaling: the global batch is 64*N, gradients are mean-reduced over it).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CFG = dict(neural_dim=512, n_units=768, n_layers=5, n_days=45, n_classes=41, patch_size=14, patch_stride=4,
           rnn_dropout=0.4, input_dropout=0.2)
B, T = 64, 400
GRU_GEMM_FLOP_PER_TRIAL = 18.88e9      # BASELINE.md section 2 (training, GRU GEMMs only)
N_ROT = 4                              # distinct input batches rotated through (4 x 52 MB > 126 MB L2)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return d.get("bf16_tflops_sustained", 1400.0), d.get("hbm_gbs", 6650.0), "measured"
    return 1400.0, 6650.0, "fallback"


def synth_batches(seed, n):
    import numpy as np
    import torch
    rng = np.random.RandomState(seed)
    out = []
    for _ in range(n):
        x = torch.from_numpy(rng.randn(B, T, CFG["neural_dim"]).astype("float32"))
        lens = rng.randint(20, 46, size=B)
        labels = np.zeros((B, 45), dtype=np.int32)
        for b in range(B):
            labels[b, :lens[b]] = rng.randint(1, 41, size=lens[b])
        days = np.repeat(rng.choice(CFG["n_days"], size=4, replace=False), 16).astype(np.int32)
        n_steps = np.full((B,), T, dtype=np.int32)
        out.append(dict(x=x, labels=torch.from_numpy(labels), lens=torch.from_numpy(lens.astype(np.int32)),
                        days=torch.from_numpy(days), n_steps=torch.from_numpy(n_steps)))
    return out


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region (B200_PROFILING.md)."""

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.proc:
            self.proc.terminate()
        sm, mx, reasons = [], 0, set()
        for r in self.rows:
            f = [c.strip() for c in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx = max(mx, float(f[1]))
            except ValueError:
                continue
            for name, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons), "samples": len(sm)}


def run_ours(args):
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    import b2t_pkg
    import __graft_entry__
    if rank == 0:
        __graft_entry__.build()
    if world > 1:
        dist.barrier()
    E = b2t_pkg.submodule("engine")
    N = b2t_pkg.load()._native
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    from torch_cpu_port import PortModel
    torch.manual_seed(0)
    init = PortModel(**CFG)                                  # random init of the reference architecture (same init calls)
    cfg = E.make_config(**CFG)
    flat = E.flat_from_state_dict(cfg, init.state_dict()).to(dev)
    eng = E.Engine(cfg, flat, max_batch=B, max_T=T, max_label_len=64, training=True)
    host = synth_batches(1234 + rank, N_ROT)
    for hb in host:
        for k in hb:
            hb[k] = hb[k].pin_memory()
    res = [{k: v.to(dev) for k, v in hb.items()} for hb in host]
    in_len = ((res[0]["n_steps"].float() - CFG["patch_size"]) / CFG["patch_stride"] + 1).to(torch.int32)
    lr = [5e-3 * 0.5] * 3
    wd = [0.0, 0.0, 1e-3]
    gscale = 1.0 / (B * world)
    loss_host = torch.empty(B, pin_memory=True)
    stage = [{k: torch.empty_like(v, device=dev) for k, v in host[0].items()} for _ in range(2)]

    def step(i, from_host):
        if from_host:
            hb, d = host[i % N_ROT], stage[i % 2]
            for k in hb:
                d[k].copy_(hb[k], non_blocking=True)
        else:
            d = res[i % N_ROT]
        eng.forward(d["x"], d["days"], training=True, smooth_mode=1, cut=i % 3, white_noise_std=1.0, offset_noise_std=0.2,
                    seed=1000 + i, want_logits=False)
        loss = eng.ctc_loss(d["labels"], in_len, d["lens"], grad_scale=gscale)
        eng.backward()
        if world > 1:
            dist.all_reduce(eng.grads)                      # ONE NCCL all-reduce: gradients + day-touched flags
        eng.optimizer_step(lr, wd, 0.9, 0.999, 0.1, 10.0)
        if from_host:
            loss_host.copy_(loss, non_blocking=True)
            torch.cuda.current_stream().synchronize()       # the reference reads loss.item() every step (rnn_trainer.py:562)
        return loss

    def timed(n, from_host, sample_clocks=False):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        cs = ClockSampler(local) if sample_clocks else None
        if cs:
            cs.start()
        l0 = N.lib.b2t_launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(n):
            step(i, from_host)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        clocks = cs.stop() if cs else None
        launches = N.lib.b2t_launch_count() - l0
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = t.item()
        return ms, launches, clocks

    for i in range(max(args.warmup, 3)):
        step(i, False)
    ms, launches, clocks = timed(args.steps, False, sample_clocks=True)
    for i in range(2):
        step(i, True)
    ms_e2e, _, _ = timed(args.steps, True)
    final_loss = float(step(0, False).mean().item())

    # dominant kernel (layer-0 input projection GEMM, 205 GFLOP per launch) timed live with CUDA events on the launch stream
    kern = None
    if rank == 0:
        M, K, Nn = 97 * 64, 7168, 2304
        a = torch.randn(M, K, device=dev).to(torch.bfloat16); w = torch.randn(Nn, K, device=dev).to(torch.bfloat16)
        for _ in range(3):
            E.gemm_bf16(a, w)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            E.gemm_bf16(a, w)
        e1.record(); torch.cuda.synchronize()
        kms = e0.elapsed_time(e1) / 10
        kern = {"name": "gemm_bf16_kernel<K,K> 6208x2304x7168 (layer-0 input projection)", "ms": kms,
                "tflops": 2.0 * M * K * Nn / kms / 1e9}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    peak_tf, peak_hbm, src = peaks()
    tps = B * world * args.steps / (ms / 1e3)
    tps_e2e = B * world * args.steps / (ms_e2e / 1e3)
    achieved = tps / world * GRU_GEMM_FLOP_PER_TRIAL / 1e12
    cpu = cpu_baseline(bounded_trials=8, steps=1)
    h2d = sum(v.numel() * v.element_size() for v in host[0].values())
    out = {
        "metric": "trials/sec (512-feat x 400-step) GRU+CTC train", "value": tps, "unit": "trials/s", "n_gpus": world,
        "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": "configs[1]: batch=64/GPU GRU+CTC training step (5x768 GRU, 512 feat x 400 bins, T'=97, labels 20-45)",
                   "global_batch": B * world, "parallelism": f"dp{world}", "l2": f"{N_ROT} rotating 52 MB input batches + ~0.8 GB activation working set per step (> 126 MB L2)",
                   "final_loss": final_loss},
        "e2e": {"value": tps_e2e, "unit": "trials/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": B * 4, "ms_per_step": ms_e2e / args.steps},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": {"bound": "tensor", "achieved": achieved, "peak": peak_tf, "unit": "TFLOP/s", "frac": achieved / peak_tf,
                     "traffic": None, "peak_source": src + " (bf16_tflops_sustained)",
                     "note": "whole-step GRU-GEMM FLOPs (18.88 GFLOP/trial, BASELINE.md) / step time, per GPU",
                     "dominant_kernel": kern},
        "cpu_baseline": cpu,
    }
    print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


def cpu_baseline(bounded_trials=8, steps=1):
    """The reference's PyTorch-CPU training step (oracle/torch_cpu_port.py) on this box's host cores, bounded sample."""
    import torch
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    from torch_cpu_port import PortModel, make_optimizer, train_step
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    torch.manual_seed(0)
    model = PortModel(**CFG)
    opt = make_optimizer(model)
    hb = synth_batches(99, 1)[0]
    nb = bounded_trials
    x, labels, lens, days, n_steps = hb["x"][:nb], hb["labels"][:nb].long(), hb["lens"][:nb].long(), hb["days"][:nb].long(), hb["n_steps"][:nb].long()
    train_step(model, opt, x[:2], n_steps[:2], labels[:2], lens[:2], days[:2])         # warm-up (thread pool, allocator)
    t0 = time.time()
    for _ in range(steps):
        train_step(model, opt, x, n_steps, labels, lens, days)
    dt = time.time() - t0
    return {"value": nb * steps / dt, "unit": "trials/s", "cores": cores, "kind": "port",
            "sample": f"{steps} full training step(s) on {nb} synthetic 512x400 trials, fp32, torch {torch.__version__} CPU, {cores} threads"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return
    nb = 16
    steps = max(1, min(args.steps, 3))
    import torch
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    from torch_cpu_port import PortModel, make_optimizer, train_step
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    torch.manual_seed(0)
    model = PortModel(**CFG)
    opt = make_optimizer(model)
    hb = synth_batches(99, 1)[0]
    x, labels, lens, days, n_steps = hb["x"][:nb], hb["labels"][:nb].long(), hb["lens"][:nb].long(), hb["days"][:nb].long(), hb["n_steps"][:nb].long()
    for _ in range(max(1, min(args.warmup, 2))):
        train_step(model, opt, x[:4], n_steps[:4], labels[:4], lens[:4], days[:4])
    t0 = time.time()
    for _ in range(steps):
        train_step(model, opt, x, n_steps, labels, lens, days)
    dt = time.time() - t0
    v = nb * steps / dt
    sample = f"{steps} training step(s) of {nb} synthetic 512x400 trials (bounded sample of the batch-64 step), fp32, torch CPU, {cores} threads"
    out = {"impl": "reference", "metric": "trials/sec (512-feat x 400-step) GRU+CTC train", "value": v, "unit": "trials/s",
           "n_gpus": world, "steps": steps, "warmup": max(1, min(args.warmup, 2)), "ms_per_step": dt / steps * 1e3,
           "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
           "config": {"workload": "configs[1]: GRU+CTC training step, reference PyTorch-CPU path (oracle/torch_cpu_port.py), bounded 16-trial sample"},
           "cpu_baseline": {"value": v, "unit": "trials/s", "cores": cores, "kind": "port", "sample": sample},
           "e2e": {"value": v, "unit": "trials/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(out))


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    a = ap.parse_args()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
