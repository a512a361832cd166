"""CPU tests: the oracle restatements against the golden vectors generated from the reference."""
import os
import sys

import numpy as np
import pytest

import util

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "oracle"))
import gru_ctc_oracle as O  # noqa: E402


def test_gauss_taps_and_smoothing():
    z = np.load(os.path.join(util.GOLDEN, "smooth.npz"))
    taps = O.gauss_taps(2, 100)
    assert len(taps) == 9 and np.array_equal(taps, z["taps"])
    assert np.abs(O.gauss_smooth(z["x"], 2, 100, "same") - z["same"]).max() < 2e-6
    assert np.abs(O.gauss_smooth(z["x"], 2, 100, "valid") - z["valid"]).max() < 2e-6
    assert O.gauss_smooth(z["x"], 2, 100, "valid").shape[1] == z["x"].shape[1] - 8


@pytest.mark.parametrize("name", ["train_small.npz", "train_ragged.npz"])
def test_numpy_oracle_full_step(name):
    params, grads, p1, rest = util.load_golden(name)
    P = O.Params(params)
    xs, _ = O.transform_data(rest["x"], rest["n_steps"], mode="val")
    adj = O.adjusted_lens(rest["n_steps"])
    lg, hid, cache = O.forward(P, xs, rest["days"], keep_cache=True)
    assert np.abs(lg - rest["logits"]).max() < 2e-4
    loss, dl = O.ctc_loss_and_grad(lg, rest["labels"], adj, rest["lens"])
    assert util.rel_err(loss, rest["loss_vec"]) < 1e-5
    assert np.abs(dl - rest["dlogits"]).max() < 2e-6
    g = O.backward(P, cache, dl, rest["days"])
    assert set(g) == set(grads)                      # untouched day layers have no gradient (grad=None in the reference)
    for k in grads:
        assert util.rel_err(g[k].reshape(grads[k].shape), grads[k]) < 2e-4, k
    tot, clipped = O.clip_grad_norm(g, 10.0)
    assert abs(tot - float(rest["grad_norm"])) < 1e-3 * tot
    st = {}
    lr = float(rest["lr"])
    for grp, wd in (("bias", 0.0), ("day", 0.0), ("other", 1e-3)):
        O.adamw_step(P, {k: v for k, v in clipped.items() if O.param_group(k) == grp}, st, step=1, lr=lr, eps=0.1, weight_decay=wd)
    for k, v in p1.items():
        assert np.abs(P[k].reshape(v.shape) - v).max() < 2e-6, k
    for b in range(lg.shape[0]):
        dec = O.greedy_decode(lg[b], int(adj[b]))
        assert O.edit_distance(dec, rest["labels"][b][:rest["lens"][b]]) == int(rest["edit_distances"][b])


@pytest.mark.parametrize("name", ["train_small.npz"])
def test_torch_cpu_port_matches_reference(name):
    """bench.py's CPU baseline (oracle/torch_cpu_port.py) reproduces the reference's logits and loss."""
    import torch
    from torch_cpu_port import PortModel, smooth_same
    params, grads, p1, rest = util.load_golden(name)
    D, H, L, n_days, B, T = [int(v) for v in rest["cfg"]]
    m = PortModel(D, H, n_days, 41, L, 14, 4, 0.0, 0.0)
    m.load_numpy(params)
    m.eval()
    with torch.no_grad():
        lg = m(smooth_same(torch.from_numpy(rest["x"])), rest["days"])
    assert np.abs(lg.numpy() - rest["logits"]).max() < 1e-4


def test_lr_schedule_and_groups():
    assert O.lr_lambda(0, 0.02, 120000, 1000) == 0.0
    assert abs(O.lr_lambda(500, 0.02, 120000, 1000) - 0.5) < 1e-12
    assert abs(O.lr_lambda(1000, 0.02, 120000, 1000) - 1.0) < 1e-12
    assert O.lr_lambda(130000, 0.02, 120000, 1000) == 0.02
    assert O.param_group("gru.bias_ih_l0") == "bias" and O.param_group("day_weights.3") == "day" and O.param_group("h0") == "other"
    assert list(O.adjusted_lens([400, 399, 398, 14, 17])) == [97, 97, 97, 1, 1]
