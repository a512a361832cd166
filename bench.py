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
    """SM clock / throttle reasons sampled DURING the timed region (B200_PROFILING.md clocks line).
    NVML is polled from a thread every ~2 ms (the timed region is ~0.1 s, too short for `nvidia-smi -lms`);
    nvidia-smi is only the fallback when the NVML binding is missing."""
    _REASONS = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4}

    def __init__(self, index):
        self.index, self.sm, self.mx, self.reasons, self.power = index, [], 0, set(), 0.0
        self._stop = threading.Event()
        self._thread = None

    def _nvml_loop(self):
        import pynvml as nv
        nv.nvmlInit()
        h = nv.nvmlDeviceGetHandleByIndex(self.index)
        self.mx = float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))
        while not self._stop.is_set():
            self.sm.append(float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)))
            mask = nv.nvmlDeviceGetCurrentClocksEventReasons(h) if hasattr(nv, "nvmlDeviceGetCurrentClocksEventReasons") \
                else nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
            for name, bit in self._REASONS.items():
                if mask & bit:
                    self.reasons.add(name)
            try:
                self.power = max(self.power, nv.nvmlDeviceGetPowerUsage(h) / 1e3)
            except Exception:
                pass
            time.sleep(0.002)

    def _smi_loop(self):
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), f"--query-gpu={q}", "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=10).stdout
                f = [c.strip() for c in out.strip().split(",")]
                self.sm.append(float(f[0])); self.mx = max(self.mx, float(f[1]))
                for name, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], f[2:6]):
                    if v.lower().startswith("active"):
                        self.reasons.add(name)
            except Exception:
                return

    def _loop(self):
        try:
            self._nvml_loop()
        except Exception:
            self._smi_loop()

    def start(self):
        self._thread = threading.Thread(target=self._loop, daemon=True)
        self._thread.start()
        time.sleep(0.01)                                      # let NVML initialise before the timed region opens

    def stop(self):
        self._stop.set()
        if self._thread:
            self._thread.join(timeout=15)
        sm = sorted(self.sm)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": self.mx or None, "reasons": sorted(self.reasons),
                "samples": len(sm), "power_w_max": self.power or None}


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
    copy_stream = torch.cuda.Stream(device=dev)
    copied = [torch.cuda.Event(), torch.cuda.Event()]

    def prefetch(i):
        """Host -> device copy of step i's batch (pinned memory) on the copy stream, into the staging buffer the step
        before last has finished with (every e2e step ends with a stream synchronize)."""
        hb, d = host[i % N_ROT], stage[i % 2]
        copy_stream.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(copy_stream):
            for k in hb:
                d[k].copy_(hb[k], non_blocking=True)
            copied[i % 2].record(copy_stream)

    def step(i, from_host, last=False):
        if from_host:
            # double-buffered input pipeline: batch i was issued during step i-1 (or just now for the first step); batch i+1 is
            # issued here and overlaps this step's compute.  Every step copies exactly one batch inside the timed region.
            if i == 0:
                prefetch(0)
            if not last:
                prefetch(i + 1)
            d = stage[i % 2]
            torch.cuda.current_stream().wait_event(copied[i % 2])
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
            step(i, from_host, last=(i == n - 1))
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
        step(i, True, last=(i == 1))
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

    # per-kernel view of one more step (CUDA events around every task on its own stream, engine timeline facility): where the
    # step's time goes and what each kernel class achieves against the share of the chip it occupies
    kernels = None
    torch.cuda.synchronize()
    if rank == 0:
        N.lib.b2t_debug_timeline(1)
    step(3, False)                                           # every rank takes part (the step holds a collective when world > 1)
    torch.cuda.synchronize()
    if rank == 0:
        import ctypes
        buf = ctypes.create_string_buffer(1 << 16)
        N.lib.b2t_debug_dump_timeline(buf, 1 << 16)
        N.lib.b2t_debug_timeline(0)
        H, Tp = CFG["n_units"], 97
        agg = {}
        for line in buf.value.decode().strip().split("\n"):
            f = line.split()
            if len(f) != 4:
                continue
            name, dur = f[1], float(f[3]) - float(f[2])
            key = ("gru_rec_bwd_kernel" if name.startswith("RB") else "gru_rec_fwd_kernel" if name.startswith("R") else
                   "gemm_bf16_kernel (L0 input projection, time chunks)" if name.startswith("G0.") else
                   "gemm_bf16_kernel (other)" if name[0] in "GDd" or name in ("day", "head") else name)
            a = agg.setdefault(key, {"launches": 0, "us": 0.0})
            a["launches"] += 1; a["us"] += dur
        rec_flop = 2.0 * B * 3 * H * H * Tp * CFG["n_layers"]              # all layers, all steps, one direction
        kernels = []
        for key, a in sorted(agg.items(), key=lambda kv: -kv[1]["us"]):
            e = {"name": key, "launches": a["launches"], "sum_us": round(a["us"], 1)}
            if key.startswith("gru_rec"):
                e["flop"] = rec_flop
                e["tflops_per_launch_avg"] = rec_flop / a["us"] / 1e6
                e["note"] = "48 CTAs per launch, up to 3 launches concurrent; latency-bound serial chain (97 steps x 5 layers)"
            if key.startswith("gemm_bf16_kernel (L0"):
                e["flop"] = 2.0 * Tp * B * 7168 * 2304
                e["tflops_per_launch_avg"] = e["flop"] / a["us"] / 1e6
                e["note"] = "runs concurrently with recurrence launches, i.e. on a share of the SMs"
            kernels.append(e)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    peak_tf, peak_hbm, src = peaks()
    tps = B * world * args.steps / (ms / 1e3)
    tps_e2e = B * world * args.steps / (ms_e2e / 1e3)
    achieved = tps / world * GRU_GEMM_FLOP_PER_TRIAL / 1e12
    cpu = cpu_baseline(bounded_trials=8, steps=1) if world == 1 else None      # reported at N=1 only (the other ranks' cores are busy at N>1)
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
                     "traffic": 169.5e6, "traffic_note": "dram bytes read+write of ONE launch of the dominant GEMM below (ncu --set full, profiles/r1_ncu_full_summary.md); algorithmic 179 MB",
                     "peak_source": src + " (bf16_tflops_sustained)",
                     "note": "achieved = whole-step GRU-GEMM FLOPs (18.88 GFLOP/trial, BASELINE.md) / step time, per GPU: the fraction of the GRU-GEMM roofline north_star asks for",
                     "dominant_kernel": kern, "kernels": kernels},
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
