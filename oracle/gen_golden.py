"""Generate golden vectors from the UNMODIFIED reference (run in the build container only).

    python oracle/gen_golden.py            # writes tests/golden/*.npz

Imports /root/reference/model_training/{rnn_model,data_augmentations}.py, runs them on seeded
inputs together with torch.nn.CTCLoss / autograd / clip_grad_norm_ / AdamW exactly as
rnn_trainer.py:527-558 does (fp32 on CPU: autocast(device_type='cuda') is a no-op on CPU), checks
the numpy restatement in oracle/gru_ctc_oracle.py against them, and stores inputs + outputs as
small fixtures.  /root/reference does not exist on the GPU box, so nothing else reads it.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, "/root/reference/model_training")

from rnn_model import GRUDecoder                   # noqa: E402  (the reference)
from data_augmentations import gauss_smooth        # noqa: E402  (the reference)
import gru_ctc_oracle as O                         # noqa: E402

OUT = os.path.join(HERE, "..", "tests", "golden")


def params_from_module(m):
    return O.Params({k: v.detach().numpy().copy() for k, v in m.state_dict().items()})


def gen_smooth():
    g = torch.Generator().manual_seed(0)
    x = torch.randn(3, 50, 16, generator=g)
    same = gauss_smooth(x, "cpu", 2, 100, padding="same").numpy()
    valid = gauss_smooth(x, "cpu", 2, 100, padding="valid").numpy()
    taps = O.gauss_taps(2, 100)
    assert len(taps) == 9, len(taps)
    o_same = O.gauss_smooth(x.numpy(), 2, 100, "same")
    o_valid = O.gauss_smooth(x.numpy(), 2, 100, "valid")
    assert np.abs(o_same - same).max() < 2e-6, np.abs(o_same - same).max()
    assert np.abs(o_valid - valid).max() < 2e-6
    np.savez_compressed(os.path.join(OUT, "smooth.npz"), x=x.numpy(), same=same, valid=valid, taps=taps)
    print("smooth ok; taps", taps)


def make_batch(B, T, D, n_days, S_lo, S_hi, seed, ragged):
    rng = np.random.RandomState(seed)
    x = rng.randn(B, T, D).astype(np.float32)
    if ragged:
        n_steps = rng.randint(int(T * 0.75), T + 1, size=B).astype(np.int64)
        n_steps[0] = T
        for b in range(B):
            x[b, n_steps[b]:] = 0
    else:
        n_steps = np.full((B,), T, dtype=np.int64)
    lens = rng.randint(S_lo, S_hi + 1, size=B).astype(np.int64)
    Smax = int(lens.max())
    labels = np.zeros((B, Smax), dtype=np.int64)
    for b in range(B):
        labels[b, :lens[b]] = rng.randint(1, 41, size=lens[b])
    days = np.repeat(rng.choice(n_days, size=max(1, B // 2), replace=False), 2)[:B].astype(np.int64)
    return x, n_steps, labels, lens, days


def gen_train_step(name, *, D, H, L, n_days, B, T, seed, S_lo, S_hi, ragged):
    """One full reference training step (eval-mode dropout=0 so that it is deterministic):
    smoothing -> model -> log_softmax -> CTCLoss(mean) -> backward -> clip -> AdamW."""
    torch.manual_seed(seed)
    model = GRUDecoder(neural_dim=D, n_units=H, n_days=n_days, n_classes=41, rnn_dropout=0.0,
                       input_dropout=0.0, n_layers=L, patch_size=14, patch_stride=4)
    # perturb the day layers off identity so that their grads are exercised non-trivially
    with torch.no_grad():
        for i in range(n_days):
            model.day_weights[i].add_(0.05 * torch.randn(D, D))
            model.day_biases[i].add_(0.05 * torch.randn(1, D))
    p0 = params_from_module(model)
    x, n_steps, labels, lens, days = make_batch(B, T, D, n_days, S_lo, S_hi, seed + 1, ragged)

    # --- reference step (rnn_trainer.py:520-558), fp32 CPU
    feats = gauss_smooth(torch.from_numpy(x), "cpu", 2, 100)
    adj = ((torch.from_numpy(n_steps) - 14) / 4 + 1).to(torch.int32)
    model.train()
    logits = model(feats, torch.from_numpy(days))
    logits.retain_grad()
    ctc = torch.nn.CTCLoss(blank=0, reduction="none", zero_infinity=False)
    loss_vec = ctc(torch.permute(logits.log_softmax(2), [1, 0, 2]), torch.from_numpy(labels), adj,
                   torch.from_numpy(lens))
    loss = torch.mean(loss_vec)
    loss.backward()
    grads = {k: v.grad.detach().numpy().copy() for k, v in model.named_parameters() if v.grad is not None}
    dlogits = logits.grad.numpy().copy()
    bias_params = [p for n, p in model.named_parameters() if "gru.bias" in n or "out.bias" in n]
    day_params = [p for n, p in model.named_parameters() if "day_" in n]
    other = [p for n, p in model.named_parameters() if "day_" not in n and "gru.bias" not in n and "out.bias" not in n]
    lr = 5e-3 * 0.37
    opt = torch.optim.AdamW([{"params": bias_params, "weight_decay": 0}, {"params": day_params, "lr": lr, "weight_decay": 0},
                             {"params": other}], lr=lr, betas=(0.9, 0.999), eps=0.1, weight_decay=1e-3)
    gn = torch.nn.utils.clip_grad_norm_(model.parameters(), max_norm=10, error_if_nonfinite=True, foreach=True)
    opt.step()
    p1 = params_from_module(model)

    # --- the numpy restatement must agree
    xs, _ = O.transform_data(x, n_steps, mode="val")
    assert np.abs(xs - feats.numpy()).max() < 2e-6
    adj_o = O.adjusted_lens(n_steps)
    assert (adj_o == adj.numpy()).all()
    lg, hid, cache = O.forward(p0, xs, days, keep_cache=True)
    err = np.abs(lg - logits.detach().numpy()).max()
    assert err < 2e-4, err
    lo, dl = O.ctc_loss_and_grad(lg, labels, adj_o, lens)
    assert np.abs(lo - loss_vec.detach().numpy()).max() < 1e-3 * max(1.0, np.abs(lo).max()), (lo, loss_vec)
    assert np.abs(dl - dlogits).max() < 2e-6, np.abs(dl - dlogits).max()
    og = O.backward(p0, cache, dl, days)
    assert set(og) == set(grads), (set(og) ^ set(grads))
    for k in grads:
        e = np.abs(og[k].reshape(grads[k].shape) - grads[k]).max()
        s = np.abs(grads[k]).max() + 1e-8
        assert e < 2e-4 * max(1.0, s), (k, e, s)
    tot, clipped = O.clip_grad_norm(og, 10.0)
    assert abs(tot - float(gn)) < 1e-3 * max(1.0, tot), (tot, float(gn))
    p_new = O.Params({k: v.copy() for k, v in p0.items()})
    st = {}
    for grp, wd in (("bias", 0.0), ("day", 0.0), ("other", 1e-3)):
        sub = {k: v for k, v in clipped.items() if O.param_group(k) == grp}
        O.adamw_step(p_new, sub, st, step=1, lr=lr, eps=0.1, weight_decay=wd)
    for k in p1:
        e = np.abs(p_new[k] - p1[k]).max()
        assert e < 2e-6, (k, e)

    # greedy decode + PER numerator (rnn_trainer.py:724-736)
    import torchaudio.functional as AF
    eds = []
    for b in range(B):
        dec = torch.argmax(logits[b, :adj[b]].detach(), dim=-1)
        dec = torch.unique_consecutive(dec, dim=-1).numpy()
        dec = np.array([i for i in dec if i != 0])
        ed = AF.edit_distance(dec, labels[b, :lens[b]])
        eds.append(ed)
        od = O.greedy_decode(lg[b], int(adj_o[b]))
        assert list(dec) == od
        assert O.edit_distance(od, labels[b, :lens[b]]) == ed

    out = {f"p.{k}": v for k, v in p0.items()}
    out.update({f"g.{k}": v for k, v in grads.items()})
    out.update({f"p1.{k}": v for k, v in p1.items() if np.abs(p1[k] - p0[k]).max() > 0})
    out.update(dict(x=x, n_steps=n_steps, labels=labels, lens=lens, days=days, logits=logits.detach().numpy(),
                    loss_vec=loss_vec.detach().numpy(), dlogits=dlogits, grad_norm=np.float32(float(gn)),
                    lr=np.float64(lr), edit_distances=np.array(eds), hidden=hid.astype(np.float32),
                    cfg=np.array([D, H, L, n_days, B, T])))
    np.savez_compressed(os.path.join(OUT, name), **out)
    print(name, "ok: logits err", err, "loss", float(loss), "gn", float(gn))


def gen_full_forward():
    """Config 1 of BASELINE.json: one 512x400 trial through the full-size model (seed 0), trainer path
    ('same' smoothing, T'=97) and evaluate path ('valid', T'=95).  Weights are NOT stored (177 MB);
    the fixture keeps the input seed, a few logits and the greedy phoneme strings; the test
    regenerates the weights with torch.manual_seed(0) + the same init calls (same torch build)."""
    torch.manual_seed(0)
    model = GRUDecoder(neural_dim=512, n_units=768, n_days=45, n_classes=41, rnn_dropout=0.4,
                       input_dropout=0.2, n_layers=5, patch_size=14, patch_stride=4).eval()
    g = torch.Generator().manual_seed(0)
    x = torch.randn(1, 400, 512, generator=g)
    with torch.no_grad():
        la = model(gauss_smooth(x, "cpu", 2, 100), [0])
        lb, hb = model(gauss_smooth(x, "cpu", 2, 100, padding="valid"), torch.tensor([0]), None, True)
    assert la.shape == (1, 97, 41) and lb.shape == (1, 95, 41)
    p = params_from_module(model)
    xs, _ = O.transform_data(x.numpy(), [400], mode="val")
    lo, _ = O.forward(p, xs, [0], dtype=np.float32)
    assert np.abs(lo - la.numpy()).max() < 5e-4, np.abs(lo - la.numpy()).max()
    np.savez_compressed(os.path.join(OUT, "full_forward.npz"), logits_same=la.numpy(), logits_valid=lb.numpy(),
                        hidden_valid=hb.numpy(), n_params=np.int64(sum(v.size for v in p.values())),
                        w_checksum=np.float64(sum(float(np.abs(v).sum()) for v in p.values())))
    print("full_forward ok", la.shape, lb.shape)


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(8)
    gen_smooth()
    gen_train_step("train_small.npz", D=32, H=64, L=2, n_days=4, B=4, T=62, seed=3, S_lo=2, S_hi=5, ragged=False)
    gen_train_step("train_ragged.npz", D=32, H=128, L=3, n_days=5, B=6, T=90, seed=5, S_lo=1, S_hi=7, ragged=True)
    gen_full_forward()
