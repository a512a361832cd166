"""CPU test (gloo, world_size 2) of the data-parallel protocol the trainer and bench.py use on NCCL:
every rank computes gradients of  sum_b loss_b / (B_local * world)  on its shard, ONE all_reduce(SUM) over the
flat buffer [gradients | day-touched flags] follows, and the update skips day layers no rank touched.
The per-rank compute is the numpy oracle here (the CUDA engine needs a GPU); what is under test is the flat
layout, the scaling, the single collective and the union-of-touched-days rule."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import util

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", "oracle"))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _flat_grads(O, params, rest, sel, world, layout, n_params, n_days):
    """oracle gradients of the shard `sel`, packed exactly like the engine's gradient buffer"""
    xs, _ = O.transform_data(rest["x"][sel], rest["n_steps"][sel], mode="val")
    adj = O.adjusted_lens(rest["n_steps"][sel])
    lg, _, cache = O.forward(O.Params(params), xs, rest["days"][sel], keep_cache=True)
    _, dl = O.ctc_loss_and_grad(lg, rest["labels"][sel], adj, rest["lens"][sel])      # already / B_local
    g = O.backward(O.Params(params), cache, dl / world, rest["days"][sel])
    flat = np.zeros(n_params + 64 * ((n_days + 63) // 64), dtype=np.float32)
    for name, off, rows, cols in layout:
        if name in g:
            flat[off:off + rows * cols] = np.asarray(g[name], dtype=np.float32).reshape(-1)
    for d in set(int(v) for v in rest["days"][sel]):
        flat[n_params + d] = 1.0
    return flat


def _worker(rank, world, port, layout, n_params, n_days, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import gru_ctc_oracle as O
    params, grads, p1, rest = util.load_golden("train_ragged.npz")
    B = rest["x"].shape[0]
    sel = np.arange(B)[rank::world]
    flat = torch.from_numpy(_flat_grads(O, params, rest, sel, world, layout, n_params, n_days))
    dist.all_reduce(flat)                                    # the ONE collective of a training step
    if rank == 0:
        np.save(out, flat.numpy())
    dist.destroy_process_group()


def test_dp_allreduce_equals_global_batch(pkg, tmp_path):
    import b2t_pkg
    E = b2t_pkg.submodule("engine")
    import gru_ctc_oracle as O
    params, grads, p1, rest = util.load_golden("train_ragged.npz")
    D, H, L, n_days, B, T = [int(v) for v in rest["cfg"]]
    cfg = E.make_config(D, H, L, n_days, 41, 14, 4)
    layout, n_params = E.param_layout(cfg), E.param_elems(cfg)
    assert E.grad_elems(cfg) == n_params + 64 * ((n_days + 63) // 64)
    out = str(tmp_path / "reduced.npy")
    mp.spawn(_worker, args=(2, _free_port(), layout, n_params, n_days, out), nprocs=2, join=True)
    red = np.load(out)
    # reference: the whole batch on one rank (the golden gradients come from the reference's own step)
    for name, off, rows, cols in layout:
        got = red[off:off + rows * cols]
        if name in grads:
            assert util.rel_err(got, grads[name].reshape(-1)) < 2e-4, name
        else:
            assert np.all(got == 0), name
    touched = red[n_params:n_params + n_days] > 0
    assert sorted(np.nonzero(touched)[0].tolist()) == sorted(set(int(d) for d in rest["days"]))
