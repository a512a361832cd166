#!/usr/bin/env python
"""Headline benchmark: trials/sec of the bs-64 GRU+CTC training step (512-feature x 400-bin synthetic
trials, bf16 tensor-core GEMMs) on N B200s of one node.  See the contract in the task statement.

    python bench.py --gpus 1 --steps 20 --warmup 5
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference        # the reference's own PyTorch-CPU path (oracle/ref_cpu_step.py) on host cores, same batch

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
        self._first = threading.Event()
        self._thread = None

    def _nvml_loop(self):
        import pynvml as nv
        nv.nvmlInit()
        h = nv.nvmlDeviceGetHandleByIndex(self.index)
        self.mx = float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))
        while not self._stop.is_set():
            self.sm.append(float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)))
            self._first.set()
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
                self._first.set()
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
        self._first.wait(timeout=20)                          # NVML init can take 100s of ms in a fresh process: the timed region opens
        self.sm.clear()                                       # only after the sampler runs; samples taken before it are dropped

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
    torch.manual_seed(0)
    # random init of the reference architecture through the package's own GRUDecoder (same init calls, same order as rnn_model.py:50-86)
    model = b2t_pkg.submodule("rnn_model").GRUDecoder(neural_dim=CFG["neural_dim"], n_units=CFG["n_units"], n_days=CFG["n_days"], n_classes=CFG["n_classes"],
                                                      rnn_dropout=CFG["rnn_dropout"], input_dropout=CFG["input_dropout"], n_layers=CFG["n_layers"],
                                                      patch_size=CFG["patch_size"], patch_stride=CFG["patch_stride"])
    cfg = E.make_config(**CFG)
    flat = model.flat_parameters.detach().clone().to(dev)
    del model
    eng = E.Engine(cfg, flat, max_batch=B, max_T=T, max_label_len=64, training=True)
    if world > 1:
        eng.reserve_comm_sms(int(os.environ.get("B2T_COMM_SMS", "0")))   # (only useful with B2T_DP_BUCKETS=bucketed: SMs the backward tail leaves to the overlapping collectives)
    host = synth_batches(1234 + rank, N_ROT)
    for hb in host:
        for k in hb:
            hb[k] = hb[k].pin_memory()
    res = [{k: v.to(dev) for k, v in hb.items()} for hb in host]
    in_len = ((res[0]["n_steps"].float() - CFG["patch_size"]) / CFG["patch_stride"] + 1).to(torch.int32)
    lr = [5e-3 * 0.5] * 3
    wd = [0.0, 0.0, 1e-3]
    strong = args.scaling == "strong"
    Bl = B // world if strong else B                        # trials per GPU: strong scaling keeps the GLOBAL batch at 64 (the reference's batch)
    if strong and (B % world or Bl < 1):
        raise SystemExit("--scaling strong needs a world size that divides 64")
    gscale = 1.0 / (Bl * world)
    loss_host = [torch.empty(Bl, pin_memory=True) for _ in range(2)]
    loss_ready = [torch.cuda.Event(), torch.cuda.Event()]
    stage = [{k: torch.empty_like(v[:Bl], device=dev) for k, v in host[0].items()} for _ in range(2)]
    copy_stream = torch.cuda.Stream(device=dev)
    copied = [torch.cuda.Event(), torch.cuda.Event()]
    if strong:
        in_len = in_len[:Bl].contiguous()
        res = [{k: v[:Bl].contiguous() for k, v in r.items()} for r in res]
        host = [{k: v[:Bl].contiguous().pin_memory() for k, v in hb.items()} for hb in host]

    e2e_losses = []

    def prefetch(i):
        """Host -> device copy of step i's batch (pinned memory) on the copy stream, into the staging buffer the step
        before last has finished with (the copy stream waits for the compute stream's work issued so far)."""
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
                    seed=1000 + i * 7919 + rank, want_logits=False)
        loss = eng.ctc_loss(d["labels"], in_len, d["lens"], grad_scale=gscale, max_target_len=45)
        eng.backward()
        if world > 1:
            if os.environ.get("B2T_BENCH_COMM_SERIAL"):     # diagnostic: no overlap at all (collective issued after backward has drained)
                torch.cuda.current_stream().synchronize()
            eng.all_reduce_grads()                          # the step's gradient all-reduce (gradients + day-touched flags), issued bucket by bucket behind backward
        eng.optimizer_step(lr, wd, 0.9, 0.999, 0.1, 10.0)
        if from_host:
            # the reference reads loss.item() every step (rnn_trainer.py:562): every step's per-trial losses are copied to pinned host
            # memory and read on the host -- one step late, so that the read of step i-1 overlaps the device work of step i
            loss_host[i % 2].copy_(loss, non_blocking=True)
            loss_ready[i % 2].record()
            if i > 0:
                loss_ready[(i - 1) % 2].synchronize()
                e2e_losses.append(float(loss_host[(i - 1) % 2][0]))
            if last:
                loss_ready[i % 2].synchronize()
                e2e_losses.append(float(loss_host[i % 2][0]))
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
    if args.profile_range:
        # ncu --replay-mode app-range: the kernels of the range run concurrently, as they do in production (the per-kernel mode
        # serialises launches, which the whole-stack recurrence and its gated GEMMs cannot survive)
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
        for i in range(args.profile_range):
            step(i, False)
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
        print(json.dumps({"profile_range_steps": args.profile_range}))
        return
    ms, launches, clocks = timed(args.steps, False, sample_clocks=True)
    for i in range(2):
        step(i, True, last=(i == 1))
    ms_e2e, _, _ = timed(args.steps, True)
    final_loss = float(step(0, False).mean().item())

    # exposed communication: the same timed loop without the all-reduce (every rank still applies its local gradients)
    comm_ms = None
    if world > 1:
        def step_nocomm(i):
            d = res[i % N_ROT]
            eng.forward(d["x"], d["days"], training=True, smooth_mode=1, cut=i % 3, white_noise_std=1.0, offset_noise_std=0.2,
                        seed=1000 + i * 7919 + rank, want_logits=False)
            eng.ctc_loss(d["labels"], in_len, d["lens"], grad_scale=gscale, max_target_len=45)
            eng.backward()
            eng.optimizer_step(lr, wd, 0.9, 0.999, 0.1, 10.0)
        dist.barrier(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(args.steps):
            step_nocomm(i)
        e1.record(); torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1)], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        comm_ms = (ms - t.item()) / args.steps

    # layer-0 input projection GEMM (205 GFLOP per launch) timed alone with CUDA events on the launch stream
    gemm_l0 = None
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
        gemm_l0 = {"name": "gemm_bf16_kernel<K,K> 6208x2304x7168 (layer-0 input projection), timed alone", "ms": kms,
                   "tflops": 2.0 * M * K * Nn / kms / 1e9}
        del a, w

    # per-kernel view of one more step (CUDA events around every task on its own stream, engine timeline facility): where the
    # step's time goes and what each kernel class achieves
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
        H, Tp, L = CFG["n_units"], 97, CFG["n_layers"]
        agg = {}
        for line in buf.value.decode().strip().split("\n"):
            f = line.split()
            if len(f) != 4:
                continue
            name, dur = f[1], float(f[3]) - float(f[2])
            key = ("gru_stack_bwd_kernel" if name == "SRB" else "gru_stack_fwd_kernel" if name == "SR" else
                   "gru_rec_bwd_kernel" if name.startswith("RB") else "gru_rec_fwd_kernel" if name.startswith("R") else
                   "gemm_bf16_kernel (L0 input projection)" if name.startswith("G0") else
                   "gemm_bf16_kernel (other)" if name[0] in "GDd" or name in ("day", "head") else name)
            a = agg.setdefault(key, {"launches": 0, "us": 0.0})
            a["launches"] += 1; a["us"] += dur
        rec_flop = 2.0 * Bl * 3 * H * H * Tp * L                            # all layers, all steps, one direction
        kernels = []
        for key, a in sorted(agg.items(), key=lambda kv: -kv[1]["us"]):
            e = {"name": key, "launches": a["launches"], "sum_us": round(a["us"], 1), "avg_us": round(a["us"] / a["launches"], 1)}
            if key.startswith("gru_"):
                e["flop_per_launch"] = rec_flop / a["launches"]
                e["tflops"] = rec_flop / a["us"] / 1e6                      # algorithmic FLOPs of a launch / its average duration
                e["note"] = "serial chain over 97 time steps x 5 layers: latency-bound (publish -> L2 -> stage -> dependent MMAs -> gate math)"
            if key.startswith("gemm_bf16_kernel (L0"):
                e["flop_per_launch"] = 2.0 * Tp * Bl * 7168 * 2304 / a["launches"]
                e["tflops"] = 2.0 * Tp * Bl * 7168 * 2304 / a["us"] / 1e6
            kernels.append(e)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    peak_tf, peak_hbm, src = peaks()
    tps = Bl * world * args.steps / (ms / 1e3)
    tps_e2e = Bl * world * args.steps / (ms_e2e / 1e3)
    whole = tps / world * GRU_GEMM_FLOP_PER_TRIAL / 1e12
    h2d = sum(v.numel() * v.element_size() for v in host[0].values())
    # dominant kernel = the kernel class with the largest share of the step (sum of its launch durations)
    dom = next((k for k in (kernels or []) if "tflops" in k), None)
    extras = {}
    if world == 1 and not args.no_extras:
        del eng
        torch.cuda.empty_cache()
        for name, fn in (("gpu_incumbent", gpu_incumbent), ("decode_config4", decode_config4)):
            try:
                extras[name] = fn(dev)
            except Exception as ex:  # noqa: BLE001 -- context numbers must never take the headline line down
                extras[name] = {"error": f"{type(ex).__name__}: {ex}"}
    cpu = cpu_baseline(steps=1) if world == 1 else None      # reported at N=1 only (the other ranks' cores are busy at N>1)
    roof = {"bound": "tensor", "unit": "TFLOP/s", "peak": peak_tf, "peak_source": src + " (bf16_tflops_sustained)",
            "achieved": dom["tflops"] if dom else whole, "frac": (dom["tflops"] if dom else whole) / peak_tf,
            "dominant_kernel": dom["name"] if dom else None,
            "traffic": None,
            "traffic_note": "per-kernel DRAM bytes cannot be captured for the shipped schedule (ncu's kernel replay serialises launches, the stack "
                            "kernels and their gated GEMMs wait for each other); profiles/r2_ncu_stack_and_step.md holds an app-range capture of the "
                            "whole step (1.94 GB read + 2.21 GB written per step) and a --set full capture of the same kernels with one layer "
                            "(gru_stack_bwd_kernel: 174 MB read + 114 MB written per 247-step launch)",
            "note": "achieved = algorithmic FLOPs of one launch of the kernel with the largest time share / its average launch duration "
                    "(CUDA events on its own stream, live in this run); the launch occupies a share of the SMs, the peak is the whole chip's",
            "whole_step": {"achieved": whole, "frac": whole / peak_tf,
                           "note": "whole-step GRU-GEMM FLOPs (18.88 GFLOP/trial, BASELINE.md) / step time, per GPU: the fraction of the GRU-GEMM roofline north_star asks for"},
            "gemm_l0": dict(gemm_l0, frac_of_burst_peak=None) if gemm_l0 else None,
            "kernels": kernels}
    out = {
        "metric": "trials/sec (512-feat x 400-step) GRU+CTC train", "value": tps, "unit": "trials/s", "n_gpus": world,
        "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True,
        "scaling": args.scaling, "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": f"configs[1]: batch={Bl}/GPU GRU+CTC training step (5x768 GRU, 512 feat x 400 bins, T'=97, labels 20-45)",
                   "global_batch": Bl * world, "parallelism": f"dp{world}", "l2": f"{N_ROT} rotating 52 MB input batches + ~0.8 GB activation working set per step (> 126 MB L2)",
                   "final_loss": final_loss},
        "e2e": {"value": tps_e2e, "unit": "trials/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": Bl * 4, "ms_per_step": ms_e2e / args.steps},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": roof,
        "cpu_baseline": cpu,
    }
    if comm_ms is not None:
        out["exposed_comm_ms_per_step"] = comm_ms
    out.update(extras)
    print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


def gpu_incumbent(dev, steps=10):
    """Context: the reference's GPU code path on the same B200 -- torch.nn.GRU (cuDNN) + cuBLAS + ATen CTC + fused AdamW under
    bf16 autocast, the statement sequence of rnn_model.py:88-134 / rnn_trainer.py:527-558, eager -- on the same synthetic batch.
    Library code end to end: the incumbent SURVEY.md 2b says to beat, not part of the product (and nothing of oracle/ is used)."""
    import math
    import torch
    import torch.nn.functional as F
    torch.backends.cudnn.deterministic = True
    torch.set_float32_matmul_precision("high")
    torch.manual_seed(0)
    D, H, L, P, S = CFG["neural_dim"], CFG["n_units"], CFG["n_layers"], CFG["patch_size"], CFG["patch_stride"]

    class Incumbent(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.day_weights = torch.nn.ParameterList([torch.nn.Parameter(torch.eye(D)) for _ in range(CFG["n_days"])])
            self.day_biases = torch.nn.ParameterList([torch.nn.Parameter(torch.zeros(1, D)) for _ in range(CFG["n_days"])])
            self.drop = torch.nn.Dropout(CFG["input_dropout"])
            self.gru = torch.nn.GRU(D * P, H, L, dropout=CFG["rnn_dropout"], batch_first=True)
            self.out = torch.nn.Linear(H, CFG["n_classes"])
            self.h0 = torch.nn.Parameter(torch.nn.init.xavier_uniform_(torch.zeros(1, 1, H)))

        def forward(self, x, day_idx):
            w = torch.stack([self.day_weights[i] for i in day_idx], 0)
            b = torch.cat([self.day_biases[i] for i in day_idx], 0).unsqueeze(1)
            x = self.drop(F.softsign(torch.einsum("btd,bdk->btk", x, w) + b))
            u = x.permute(0, 2, 1).unfold(2, P, S)
            x = u.permute(0, 2, 3, 1).reshape(x.size(0), u.size(2), -1)
            y, _ = self.gru(x, self.h0.expand(L, x.shape[0], H).contiguous())
            return self.out(y)
    m = Incumbent().to(dev)
    named = list(m.named_parameters())
    bias = [p for n, p in named if "gru.bias" in n or "out.bias" in n]
    day = [p for n, p in named if "day_" in n]
    other = [p for n, p in named if "day_" not in n and "gru.bias" not in n and "out.bias" not in n]
    opt = torch.optim.AdamW([{"params": bias, "weight_decay": 0}, {"params": day, "weight_decay": 0}, {"params": other}], lr=2.5e-3,
                            betas=(0.9, 0.999), eps=0.1, weight_decay=1e-3, fused=True)
    hb = {k: v.to(dev) for k, v in synth_batches(1, 1)[0].items()}
    w = [math.exp(-0.5 * (i / 2.0) ** 2) for i in range(-8, 9)]
    w = [v / sum(w) for v in w]
    taps = [v for v in w if v > 0.01]
    k = torch.tensor([v / sum(taps) for v in taps], device=dev).view(1, 1, -1).repeat(D, 1, 1)      # data_augmentations.py:19-24 (9 taps)
    days = hb["days"].tolist()

    def step():
        opt.zero_grad()
        with torch.autocast(device_type="cuda", dtype=torch.bfloat16):
            f = hb["x"] + torch.randn_like(hb["x"]) + torch.randn(B, 1, D, device=dev) * 0.2
            f = F.conv1d(f.permute(0, 2, 1), k, padding="same", groups=D).permute(0, 2, 1)
            adj = ((hb["n_steps"] - P) / S + 1).to(torch.int32)
            logits = m(f, days)
            loss = F.ctc_loss(logits.log_softmax(2).permute(1, 0, 2), hb["labels"].long(), adj, hb["lens"].long(), blank=0, reduction="none").mean()
        loss.backward()
        torch.nn.utils.clip_grad_norm_(m.parameters(), 10.0, error_if_nonfinite=True, foreach=True)
        opt.step()
        return loss
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        step()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    return {"what": "the reference's GPU code path through torch on this B200: cuDNN GRU + cuBLAS + ATen CTC + fused AdamW, bf16 autocast, eager",
            "ms_per_step": ms, "value": B / ms * 1e3, "unit": "trials/s"}


def decode_config4(dev, n_utt=74, n_dec=2, rounds=3):
    """Context for BASELINE.json configs[3]: inference as one pipeline -- GRU logits for a batch of trials ('valid' smoothing, T'=95,
    host inputs, logits returned to the host like runSingleDecodingStep) followed by the n-gram WFST decode of the batch at the
    reference's shipped decoder settings (max_active 7000, beam 17, lattice_beam 8, n-best 100) on a 3-gram graph compiled by the
    package's graph compiler from a synthetic corpus.  The GRU has random weights (flat posteriors), so the decoder is fed rendered
    in-vocabulary posteriors of the same [n, 95, 41] shape.  Two decoder objects (74 utterance slots each = 148 CTAs, one per SM)
    are driven from two host threads, so that the host part of one batch (lattice read-back, n-best) overlaps the GPU search of
    the other; throughput is trials per wall-clock second over `rounds` batches per decoder."""
    import math
    import tempfile
    import threading
    import numpy as np
    import torch
    import b2t_pkg
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import make_synth_lm as SL
    import make_toy_tlg as TLG
    E = b2t_pkg.submodule("engine"); LM = b2t_pkg.submodule("lm_decoder"); GC = b2t_pkg.submodule("graph_compiler")
    d = tempfile.mkdtemp()
    li = SL.build(d, order=3, n_words=1000, n_sent=20000, seed=3)
    fst, words = os.path.join(d, "TLG.fst"), os.path.join(d, "words.txt")
    gi = GC.compile_to_files(li["arpa"], li["lexicon"], li["phones"], fst, words)
    widx = {w: i for i, w in enumerate(li["words"])}
    rs = np.random.RandomState(11)
    sents = [s[1:-1] for s in li["corpus"] if 2 <= len(s) - 2 <= 4]
    sents = [sents[i] for i in rs.choice(len(sents), size=n_utt, replace=False)]
    post = np.stack([TLG.render_logits([li["prons"][widx[w]] for w in s], T=95, seed=700 + n, noise=1.0) for n, s in enumerate(sents)]).astype(np.float32)
    cfg = E.make_config(**dict(CFG, rnn_dropout=0.0, input_dropout=0.0))
    torch.manual_seed(0)
    flat = (torch.randn(E.param_elems(cfg)) * 0.02).to(dev)
    opts = (7000, 200, 17.0, 8.0, 0.325, 1.0, 0.0, 100)
    bp = math.log(90.0)
    x_host = torch.randn(n_utt, T, CFG["neural_dim"]).pin_memory()
    days = torch.zeros(n_utt, dtype=torch.int32)
    engs = [E.Engine(cfg, flat, max_batch=n_utt, max_T=T, max_label_len=1, training=False) for _ in range(n_dec)]
    decs = [LM.BrainSpeechDecoder(LM.DecodeResource(fst, "", "", words, ""), LM.DecodeOptions(*opts), max_frames=128, max_slots=n_utt) for _ in range(n_dec)]
    streams = [torch.cuda.Stream(device=dev) for _ in range(n_dec)]
    gru_ms = [0.0] * n_dec

    def pipeline(k):
        with torch.cuda.stream(streams[k]):
            t0 = time.perf_counter()
            lg, _ = engs[k].forward(x_host.to(dev, non_blocking=True), days, training=False, smooth_mode=2)
            lg_host = lg.float().cpu().numpy()                       # [n, 95, 41] to the host (evaluate_model_helpers.py:109)
            gru_ms[k] = (time.perf_counter() - t0) * 1e3
        assert lg_host.shape == post.shape
        decs[k].DecodeBatch(post, blank_penalty=bp)

    for k in range(n_dec):
        pipeline(k)                                                  # warm-up (allocations, graph upload)

    def worker(k):
        for _ in range(rounds):
            pipeline(k)
    ths = [threading.Thread(target=worker, args=(k,)) for k in range(n_dec)]
    t0 = time.perf_counter()
    for th in ths:
        th.start()
    for th in ths:
        th.join()
    dt = time.perf_counter() - t0
    hyp = [decs[0].result(slot=n) for n in range(n_utt)]
    # the same batch through the strict serial-order search (Kaldi's token-list order reproduced; parity mode), decode only
    decs[0].set_strict_order(True)
    decs[0].DecodeBatch(post, blank_penalty=bp)
    t1 = time.perf_counter()
    decs[0].DecodeBatch(post, blank_penalty=bp)
    dt_strict = time.perf_counter() - t1
    strict_same = sum(int(bool(a) and bool(b) and a[0].sentence == b[0].sentence) for a, b in zip(hyp, [decs[0].result(slot=n) for n in range(n_utt)]))
    decs[0].set_strict_order(False)
    t1 = time.perf_counter()
    decs[0].DecodeBatch(post, blank_penalty=bp)
    dt_fast1 = time.perf_counter() - t1
    err = tot = 0
    for n, h in enumerate(hyp):
        ref = [w.lower() for w in sents[n]]
        got = h[0].sentence.split() if h else []
        dp = list(range(len(got) + 1))                               # word-level edit distance
        for i in range(1, len(ref) + 1):
            prev, dp[0] = dp[0], i
            for j in range(1, len(got) + 1):
                prev, dp[j] = dp[j], min(dp[j] + 1, dp[j - 1] + 1, prev + (ref[i - 1] != got[j - 1]))
        err += dp[len(got)]; tot += len(ref)
    return {"what": f"{n_dec} x {n_utt} trials in flight: GRU logits (host in, host out) + WFST n-gram decode, shipped settings (max_active 7000, n-best 100), compiled 3-gram graph",
            "graph_states": gi["n_states"], "graph_arcs": gi["n_arcs"], "batches": n_dec * rounds, "ms_total": dt * 1e3, "gru_ms_per_batch": max(gru_ms),
            "value": n_dec * rounds * n_utt / dt, "unit": "trials/s", "ms_per_trial": dt * 1e3 / (n_dec * rounds * n_utt),
            "wer_vs_rendered_transcripts": err / max(tot, 1), "posterior_noise": 1.0,
            "strict_order": {"what": f"one decoder, {n_utt} utterances per call, decode only: strict serial-order search vs the default two-pass search",
                             "strict_trials_per_s": n_utt / dt_strict, "fast_trials_per_s": n_utt / dt_fast1, "same_1best": f"{strict_same}/{n_utt}"}}


def _ref_step_runner():
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    from ref_cpu_step import RefStep
    import torch
    rs = RefStep(CFG)
    hb = synth_batches(99, 1)[0]
    batch = (hb["x"], hb["n_steps"].long(), hb["labels"].long(), hb["lens"].long(), hb["days"].long())
    return rs, batch, torch


def cpu_baseline(steps=1):
    """The reference's own PyTorch-CPU training step (oracle/ref_cpu_step.py: unmodified rnn_model.py + data_augmentations.py when
    staged under oracle/_ref, else the port) on this box's host cores: a bounded sample of `steps` full batch-64 steps."""
    rs, batch, torch = _ref_step_runner()
    x, n_steps, labels, lens, days = batch
    rs.step(x[:4], n_steps[:4], labels[:4], lens[:4], days[:4])       # warm-up (thread pool, allocator)
    t0 = time.time()
    for i in range(steps):
        rs.step(x, n_steps, labels, lens, days, cut=i % 3)
    dt = time.time() - t0
    return {"value": B * steps / dt, "unit": "trials/s", "cores": rs.cores, "kind": rs.kind,
            "sample": f"{steps} full training step(s) of the batch-64 workload (64 synthetic 512x400 trials), fp32, torch {torch.__version__} CPU, {rs.cores} threads"}


def run_reference(args):
    """Reference arm: the reference's CPU implementation of the same step, same batch size, honouring --steps/--warmup (capped so
    that the run ends within a few minutes: a batch-64 step takes ~2 s on 16 cores)."""
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return
    steps = max(1, min(args.steps, 40))
    warm = max(1, min(args.warmup, 5))
    rs, batch, torch = _ref_step_runner()
    x, n_steps, labels, lens, days = batch
    for i in range(warm):
        rs.step(x, n_steps, labels, lens, days, cut=i % 3)
    t0 = time.time()
    for i in range(steps):
        rs.step(x, n_steps, labels, lens, days, cut=i % 3)
    dt = time.time() - t0
    v = B * steps / dt
    sample = f"{steps} full training steps of the batch-64 workload, fp32, torch {torch.__version__} CPU, {rs.cores} threads"
    out = {"impl": "reference", "metric": "trials/sec (512-feat x 400-step) GRU+CTC train", "value": v, "unit": "trials/s",
           "n_gpus": world, "steps": steps, "warmup": warm, "ms_per_step": dt / steps * 1e3,
           "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
           "config": {"workload": "configs[1]: batch=64/GPU GRU+CTC training step (5x768 GRU, 512 feat x 400 bins, T'=97, labels 20-45)",
                      "global_batch": B, "parallelism": "cpu", "impl_note": ("unmodified reference rnn_model.py + data_augmentations.py (oracle/_ref/model_training), "
                                                                               "loop body of rnn_trainer.py:511-558" if rs.kind == "reference" else
                                                                               "oracle/torch_cpu_port.py (reference files not staged)")},
           "cpu_baseline": {"value": v, "unit": "trials/s", "cores": rs.cores, "kind": rs.kind, "sample": sample},
           "e2e": {"value": v, "unit": "trials/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(out))


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak: 64 trials per GPU (default); strong: global batch 64 split over the GPUs (the reference's batch, parity config)")
    ap.add_argument("--profile-range", type=int, default=0,
                    help="run N steps between cudaProfilerStart/Stop and exit (for ncu --replay-mode app-range); not a bench value")
    ap.add_argument("--no-extras", action="store_true", help="skip the context sections (GPU incumbent, config-4 decode pipeline)")
    a = ap.parse_args()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
