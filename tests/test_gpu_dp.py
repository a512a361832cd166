"""Data-parallel training on the CUDA engine, two GPUs of one box (skipped on a single-GPU box; run with `gpurun --gpus 2`).

SURVEY.md section 7 tolerance row "DP k in {2,4,8} vs k = 1 at global batch 64 with host-injected RNG: grads rel <= 1e-5 (fp32
reduce)": every rank runs forward / CTC / backward on its half of one batch with grad_scale = 1 / global batch, the flat gradient
buffer is SUM-all-reduced bucket by bucket behind backward (Engine.all_reduce_grads), and the result must equal the single-GPU
gradient of the whole batch; the touched-day flags must be the union of the ranks' days."""
import os
import sys

import numpy as np
import pytest
import torch

import util

pytestmark = pytest.mark.gpu
ROOT = util.ROOT


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist
    import b2t_pkg
    import gru_ctc_oracle as O
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
    E = b2t_pkg.submodule("engine")
    D, H, L, n_days, B, T = 64, 256, 3, 6, 64, 120
    cfg = E.make_config(D, H, L, n_days, 41, 14, 4, 0.0, 0.0)
    rng = np.random.RandomState(3)
    flat = torch.from_numpy((rng.randn(E.param_elems(cfg)) * 0.05).astype(np.float32))
    x = rng.randn(B, T, D).astype(np.float32)
    n_steps = rng.randint(90, T + 1, size=B); n_steps[0] = T
    lens = rng.randint(2, 8, size=B)
    labels = np.zeros((B, 8), dtype=np.int32)
    for b in range(B):
        labels[b, :lens[b]] = rng.randint(1, 41, size=lens[b])
        x[b, n_steps[b]:] = 0
    days = np.repeat(np.array([0, 2, 3, 5]), B // 4).astype(np.int32)          # rank 0 sees days {0, 2}, rank 1 days {3, 5}
    wn = rng.randn(B, T, D).astype(np.float32); on = rng.randn(B, D).astype(np.float32)
    in_len = O.adjusted_lens(n_steps - 1).astype(np.int32)                      # cut = 1

    def grads_of(sel, scale, reduce):
        eng = E.Engine(cfg, flat.clone().cuda(), max_batch=len(sel), max_T=T, max_label_len=8, training=True)
        c = lambda a: torch.from_numpy(np.ascontiguousarray(a[sel]))
        eng.forward(c(x).cuda(), c(days), training=True, smooth_mode=1, cut=1, white_noise_std=1.0, offset_noise_std=0.2,
                    white_noise=c(wn).cuda(), offset_noise=c(on).cuda())
        eng.ctc_loss(c(labels), c(in_len), c(lens.astype(np.int32)), grad_scale=scale)
        eng.backward()
        if reduce:
            eng.all_reduce_grads()
        torch.cuda.synchronize()
        return eng.grads.detach().cpu().numpy().copy(), eng.n_params

    half = np.arange(B)[rank * (B // world):(rank + 1) * (B // world)]
    g_dp, n_params = grads_of(half, 1.0 / B, True)
    if rank == 0:
        g_one, _ = grads_of(np.arange(B), 1.0 / B, False)
        np.savez(os.path.join(out_dir, "dp.npz"), dp=g_dp, one=g_one, n_params=n_params)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs (gpurun --gpus 2)")
def test_two_gpu_gradients_equal_single_gpu(tmp_path, pkg):
    import torch.multiprocessing as mp
    port = 29500 + os.getpid() % 2000
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    z = np.load(os.path.join(str(tmp_path), "dp.npz"))
    n = int(z["n_params"])
    dp, one = z["dp"], z["one"]
    scale = np.abs(one[:n]).max()
    err = np.abs(dp[:n] - one[:n]).max() / scale
    # the only difference is the association of fp32 sums over the batch (two partial sums added by the collective instead of
    # one accumulation; bias / day gradients by atomics)
    assert err < 2e-5, err
    flags_dp, flags_one = dp[n:n + 6], one[n:n + 6]
    assert (flags_dp > 0).tolist() == (flags_one > 0).tolist() == [True, False, True, True, False, True]
