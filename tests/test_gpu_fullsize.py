"""GPU parity tests at the BENCHED geometry (D=512, H=768, L=5, B=64, T=400 -> T'=97) and of the stochastic ops.

  * config 1 of BASELINE.json: one 512x400 trial through the full-size model, trainer path ('same' smoothing, T'=97) and
    evaluate path ('valid', T'=95) against tests/golden/full_forward.npz, which oracle/gen_golden.py produced by running the
    UNMODIFIED reference (rnn_model.py + data_augmentations.py) in the build container;
  * one full-size batch-64 training step with HOST-INJECTED white/offset noise and cut in {0,1,2} (rnn_trainer.py:436-484)
    against the reference model in fp32 on the host cores; the gradient tolerance is not a guess: the same step is also run
    through torch's own bf16 GPU path (cuDNN GRU under autocast, the reference's production numerics) and OUR error must stay
    within a small multiple of ITS error against fp32;
  * dropout (input and inter-layer): the masks are recovered from the engine's buffers and injected into the numpy oracle, so
    forward values AND the gradients (i.e. that backward regenerates the very same masks) are checked value by value;
  * optimizer state across an engine regrow (ragged batches make the model rebuild its engine for a longer T).
"""
import json
import os

import numpy as np
import pytest
import torch

import util

pytestmark = pytest.mark.gpu

FULL = dict(neural_dim=512, n_units=768, n_days=45, n_classes=41, n_layers=5, patch_size=14, patch_stride=4)


@pytest.fixture(scope="module")
def mods(pkg):
    import b2t_pkg
    return {m: b2t_pkg.submodule(m) for m in ("rnn_model", "rnn_trainer", "ctc", "data_augmentations", "evaluate_model_helpers", "engine")}


def _report(name, payload):
    """Measured error tables land in gpurun_out/ (copied into profiles/ by hand when they are worth keeping)."""
    d = os.path.join(util.ROOT, "gpurun_out")
    if os.path.isdir(d):
        with open(os.path.join(d, name), "w") as f:
            json.dump(payload, f, indent=1, sort_keys=True)


def _greedy(lg):
    ids = lg.argmax(-1)
    keep = np.concatenate(([True], ids[1:] != ids[:-1]))
    return [int(i) for i in ids[keep] if i != 0]


def test_config1_single_trial_vs_reference_fixture(mods):
    z = np.load(os.path.join(util.GOLDEN, "full_forward.npz"))
    torch.manual_seed(0)
    m = mods["rnn_model"].GRUDecoder(rnn_dropout=0.4, input_dropout=0.2, **FULL).eval()
    # same torch seed + same init calls in the same order => the very weights the reference module had when the fixture was made
    assert sum(p.numel() for p in m.parameters()) == int(z["n_params"]) == 44315177
    chk = sum(float(p.detach().abs().double().sum()) for p in m.parameters())
    assert abs(chk - float(z["w_checksum"])) < 1e-6 * float(z["w_checksum"]), "weights differ from the reference's init: fixture not comparable"
    m.to("cuda")
    g = torch.Generator().manual_seed(0)
    x = torch.randn(1, 400, 512, generator=g).cuda()
    gs = mods["data_augmentations"].gauss_smooth
    rep = {}
    with torch.no_grad():
        la = m(gs(x, "cuda", 2, 100), [0]).cpu().numpy()                                        # trainer path, T' = 97
    assert la.shape == z["logits_same"].shape == (1, 97, 41)
    rep["same_max_abs"] = float(np.abs(la - z["logits_same"]).max())
    assert rep["same_max_abs"] < 5e-2
    same_frames = float((la[0].argmax(-1) == z["logits_same"][0].argmax(-1)).mean())
    rep["same_argmax_agreement"] = same_frames
    assert same_frames >= 0.995
    assert _greedy(la[0]) == _greedy(z["logits_same"][0])                                       # identical greedy phoneme string (PER 0 vs ref)
    # evaluate path: runSingleDecodingStep ('valid' smoothing fused into the input kernel), bf16 input like evaluate_model.py:118
    H = mods["evaluate_model_helpers"]
    args = {"use_amp": True, "dataset": {"data_transforms": {"smooth_kernel_std": 2, "smooth_kernel_size": 100}}}
    lb = H.runSingleDecodingStep(x, 0, m, args, "cuda")
    assert lb.dtype == np.float32 and lb.shape == z["logits_valid"].shape == (1, 95, 41)
    rep["valid_max_abs"] = float(np.abs(lb - z["logits_valid"]).max())
    assert rep["valid_max_abs"] < 5e-2
    assert float((lb[0].argmax(-1) == z["logits_valid"][0].argmax(-1)).mean()) >= 0.995
    assert _greedy(lb[0]) == _greedy(z["logits_valid"][0])
    with torch.no_grad():
        xv = gs(x, "cuda", 2, 100, padding="valid")
        _, hid = m(xv, torch.tensor([0]).cuda(), None, True)
    rep["hidden_valid_max_abs"] = float(np.abs(hid.cpu().numpy() - z["hidden_valid"]).max())
    assert rep["hidden_valid_max_abs"] < 3e-2
    _report("r2_config1_parity.json", rep)


def _bench_like_batch(seed, B=64, T=400, D=512, n_days=45, ragged=True):
    rng = np.random.RandomState(seed)
    x = rng.randn(B, T, D).astype(np.float32)
    n_steps = rng.randint(T * 3 // 4, T + 1, size=B).astype(np.int64) if ragged else np.full((B,), T, np.int64)
    n_steps[0] = T
    for b in range(B):
        x[b, n_steps[b]:] = 0
    lens = rng.randint(20, 46, size=B).astype(np.int64)
    labels = np.zeros((B, 45), dtype=np.int64)
    for b in range(B):
        labels[b, :lens[b]] = rng.randint(1, 41, size=lens[b])
    days = np.repeat(rng.choice(n_days, size=4, replace=False), B // 4).astype(np.int64)
    wn = rng.randn(B, T, D).astype(np.float32)
    on = rng.randn(B, D).astype(np.float32)
    return x, n_steps, labels, lens, days, wn, on


@pytest.fixture(scope="module")
def fullsize_setup(mods):
    """Full-size weights (day layers perturbed off identity), the same weights in the reference model, one bench-like batch."""
    import sys
    sys.path.insert(0, os.path.join(util.ROOT, "oracle"))
    from ref_cpu_step import RefStep
    torch.manual_seed(1)
    m = mods["rnn_model"].GRUDecoder(rnn_dropout=0.0, input_dropout=0.0, **FULL)
    with torch.no_grad():
        for i in range(FULL["n_days"]):
            m.day_weights[i].add_(0.03 * torch.randn(512, 512))
            m.day_biases[i].add_(0.05 * torch.randn(1, 512))
    cfg = dict(FULL, rnn_dropout=0.0, input_dropout=0.0)
    rs = RefStep(cfg)                                            # the unmodified reference module when staged, else the port
    rs.model.load_state_dict({k: v.detach().clone() for k, v in m.state_dict().items()})
    batch = _bench_like_batch(21)
    return m, rs, batch


def _ref_grads(model, feats, days, labels, adj, lens, device, autocast):
    """The reference's statement sequence (rnn_trainer.py:527-547) on `device`; fp32 on the host, bf16 autocast on the GPU."""
    model = model.to(device).train()
    for p in model.parameters():
        p.grad = None
    f = torch.from_numpy(feats).to(device)
    with torch.autocast(device_type="cuda", enabled=autocast, dtype=torch.bfloat16):
        logits = model(f, torch.from_numpy(days).to(device))
        loss_vec = torch.nn.functional.ctc_loss(torch.permute(logits.log_softmax(2), [1, 0, 2]), torch.from_numpy(labels).to(device),
                                                torch.from_numpy(adj).to(device), torch.from_numpy(lens).to(device), blank=0,
                                                reduction="none", zero_infinity=False)
        loss = torch.mean(loss_vec)
    loss.backward()
    grads = {k: p.grad.detach().float().cpu().numpy().copy() for k, p in model.named_parameters() if p.grad is not None}
    return logits.detach().float().cpu().numpy(), loss_vec.detach().float().cpu().numpy(), grads


@pytest.mark.parametrize("cut", [0, 1, 2])
def test_fullsize_train_step_injected_noise(mods, fullsize_setup, cut):
    import gru_ctc_oracle as O
    E = mods["engine"]
    m, rs, (x, n_steps, labels, lens, days, wn, on) = fullsize_setup
    B = x.shape[0]
    # --- checker: the reference model in fp32 on the host, fed the oracle's restatement of transform_data with the same draws
    feats, n_cut = O.transform_data(x, n_steps, mode="train", white_noise=wn, offset_noise=on[:, None, :], cut=cut,
                                    white_noise_std=1.0, constant_offset_std=0.2)
    adj = O.adjusted_lens(n_cut)
    ref_logits, ref_loss, ref_g = _ref_grads(rs.model, feats, days, labels, adj, lens, "cpu", False)
    # --- torch's own bf16 GPU path (cuDNN GRU under autocast) on the same step: how far do the reference's production numerics
    #     sit from fp32?  That sets the scale of the tolerance below.
    torch.backends.cudnn.deterministic = True
    cu_logits, cu_loss, cu_g = _ref_grads(rs.model, feats, days, labels, adj, lens, "cuda", True)
    rs.model.to("cpu")
    # --- ours: default schedule, through the C ABI
    cfg = E.make_config(512, 768, 5, 45, 41, 14, 4, 0.0, 0.0)
    eng = E.Engine(cfg, m.flat_parameters.detach().clone().cuda(), max_batch=B, max_T=400, max_label_len=64, training=True)
    logits, _ = eng.forward(torch.from_numpy(x).cuda(), torch.from_numpy(days.astype(np.int32)), training=True, smooth_mode=1, cut=cut,
                            white_noise_std=1.0, offset_noise_std=0.2, white_noise=torch.from_numpy(wn).cuda(),
                            offset_noise=torch.from_numpy(on).cuda())
    loss = eng.ctc_loss(torch.from_numpy(labels.astype(np.int32)), torch.from_numpy(adj), torch.from_numpy(lens.astype(np.int32)), grad_scale=1.0 / B)
    eng.backward()
    torch.cuda.synchronize()
    lg = logits.cpu().numpy()
    assert lg.shape == ref_logits.shape
    # pad frames (t >= adjusted length) carry no gradient and are excluded from the loss; compare the valid region
    valid = np.arange(lg.shape[1])[None, :] < adj[:, None]
    e_log, e_log_cu = float(np.abs(lg - ref_logits)[valid].max()), float(np.abs(cu_logits - ref_logits)[valid].max())
    assert e_log < max(5e-2, 2.0 * e_log_cu), (e_log, e_log_cu)
    e_loss, e_loss_cu = util.rel_err(loss.cpu().numpy(), ref_loss), util.rel_err(cu_loss, ref_loss)
    assert e_loss < max(1e-2, 2.0 * e_loss_cu), (e_loss, e_loss_cu)
    got = util.unflatten(E, cfg, eng.grads[:eng.n_params])
    table, bad = {}, {}
    for k, g in ref_g.items():
        eo = util.rel_err(got[k].reshape(g.shape), g)
        ec = util.rel_err(cu_g[k], g)
        table[k] = {"ours": eo, "torch_bf16_cudnn": ec}
        if eo > max(3.0 * ec, 1e-2):                              # within 3x of what cuDNN-bf16 itself deviates from fp32 (floor 1 %)
            bad[k] = (round(eo, 4), round(ec, 4))
    _report(f"r2_fullsize_step_cut{cut}.json", {"logits_max_abs": {"ours": e_log, "torch_bf16_cudnn": e_log_cu},
                                                 "loss_rel": {"ours": e_loss, "torch_bf16_cudnn": e_loss_cu}, "grad_rel_to_max": table})
    assert not bad, bad
    touched = sorted(np.nonzero(eng.touched_days().cpu().numpy())[0].tolist())
    assert touched == sorted(set(int(d) for d in days))
    for k in got:
        if k not in ref_g:
            assert np.abs(got[k]).max() == 0.0, k


def test_dropout_masks_forward_and_backward(mods):
    """Input dropout (rnn_model.py:102-103) and inter-layer GRU dropout (nn.GRU dropout=) with the masks the device drew:
    recovered from the engine's own buffers, injected into the numpy oracle, then forward values and every gradient compared.
    A backward pass that regenerated different masks than forward would fail the gradient comparison outright."""
    import gru_ctc_oracle as O
    E = mods["engine"]
    params, grads, p1, rest = util.load_golden("train_ragged.npz")
    D, H, L, n_days, B, T = [int(v) for v in rest["cfg"]]
    p_in, p_rnn = 0.2, 0.4
    cfg = E.make_config(D, H, L, n_days, 41, 14, 4, p_rnn, p_in)
    eng = E.Engine(cfg, util.flat_from_params(E, cfg, params, "cuda"), max_batch=B, max_T=T, max_label_len=16, training=True)
    x = torch.from_numpy(rest["x"]).cuda()
    days = rest["days"]
    logits, _ = eng.forward(x, torch.from_numpy(days.astype(np.int32)), training=True, smooth_mode=1, seed=12345)
    in_len = O.adjusted_lens(rest["n_steps"])
    loss = eng.ctc_loss(torch.from_numpy(rest["labels"].astype(np.int32)), torch.from_numpy(in_len), torch.from_numpy(rest["lens"].astype(np.int32)),
                        grad_scale=1.0 / B)
    eng.backward()
    torch.cuda.synchronize()
    Bp, Tp = (B + 15) // 16 * 16, logits.shape[1]
    xd = eng.debug_buffer("xd").float().view(Bp, T, D)[:B].cpu().numpy()
    in_mask = (xd != 0).astype(np.float32) / (1.0 - p_in)
    frac = float((xd != 0).mean())
    assert abs(frac - (1.0 - p_in)) < 0.02, frac                                  # Bernoulli(keep) within sampling error
    layer_masks = []
    for l in range(L - 1):
        hd = eng.debug_buffer("hdrop", l).float().view(Tp, Bp, H)[:, :B].cpu().numpy()
        hs = eng.debug_buffer("hseq", l).float().view(Tp + 1, Bp, H)[1:, :B].cpu().numpy()
        keep = hd != 0
        assert abs(float(keep.mean()) - (1.0 - p_rnn)) < 0.02
        # kept entries are h / keep (both stored as bf16 roundings of the fp32 value): inverted-dropout scaling
        sel = keep & (np.abs(hs) > 1e-3)
        assert np.abs(hd[sel] * (1.0 - p_rnn) / hs[sel] - 1.0).max() < 2e-2
        layer_masks.append(np.transpose(keep.astype(np.float32) / (1.0 - p_rnn), (1, 0, 2)))    # [B, T', H]
    P = O.Params(params)
    xs, _ = O.transform_data(rest["x"], rest["n_steps"], mode="val")
    ref_logits, _, cache = O.forward(P, xs, days, in_mask=in_mask, layer_masks=layer_masks, keep_cache=True)
    assert np.abs(logits.cpu().numpy() - ref_logits).max() < 5e-2
    # dropout really changed the function: the eval-mode logits differ
    assert np.abs(ref_logits - rest["logits"]).max() > 0.1
    ref_loss, dlog = O.ctc_loss_and_grad(ref_logits, rest["labels"], in_len, rest["lens"])
    assert util.rel_err(loss.cpu().numpy(), ref_loss) < 2e-2
    ref_g = O.backward(P, cache, dlog, days)
    got = util.unflatten(E, cfg, eng.grads[:eng.n_params])
    bad = {k: round(util.rel_err(got[k].reshape(np.asarray(g).shape), g), 4) for k, g in ref_g.items()
           if util.rel_err(got[k].reshape(np.asarray(g).shape), g) >= util.GRAD_TOL}
    assert not bad, bad
    # input-dropout mask seen from the backward side: d(pre-activation) of the day layer is zero exactly where forward dropped
    dpre = eng.debug_buffer("dpre").float().view(Bp, T, D)[:B].cpu().numpy()
    for b in range(B):
        tv = int(rest["n_steps"][b])
        dropped = xd[b, :tv] == 0
        assert np.all(dpre[b, :tv][dropped] == 0)
    # and a different seed draws different masks
    eng.forward(x, torch.from_numpy(days.astype(np.int32)), training=True, smooth_mode=1, seed=999)
    torch.cuda.synchronize()
    xd2 = eng.debug_buffer("xd").float().view(Bp, T, D)[:B].cpu().numpy()
    assert ((xd2 != 0) != (xd != 0)).mean() > 0.1


def test_optimizer_state_survives_engine_regrow(mods):
    """Ragged batches make GRUDecoder rebuild its engine for a longer T: moments AND the per-segment AdamW step counters
    must carry over (resetting the counters while the moments stay warm changes the bias correction ~10x at eps=0.1)."""
    import gru_ctc_oracle as O
    E = mods["engine"]
    params, grads, p1, rest = util.load_golden("train_ragged.npz")
    D, H, L, n_days, B, T = [int(v) for v in rest["cfg"]]

    def run(regrow):
        m = mods["rnn_model"].GRUDecoder(neural_dim=D, n_units=H, n_days=n_days, n_classes=41, n_layers=L, patch_size=14, patch_stride=4)
        m.load_state_dict({k: torch.from_numpy(v).reshape(m.state_dict()[k].shape) for k, v in params.items()})
        m.to("cuda")
        x = torch.from_numpy(rest["x"]).cuda()
        sched = [T - 20, T - 20, T] if regrow else [T, T, T]          # engine sized for the step's T: grows before the third step
        steps = None
        for Tm in sched:
            eng = m.engine(B, Tm, training=True)
            xi = x[:, :T - 20]                                       # identical data in both runs
            n_steps = np.minimum(rest["n_steps"], T - 20)
            eng.forward(xi, torch.from_numpy(rest["days"].astype(np.int32)), training=True, smooth_mode=1)
            eng.ctc_loss(torch.from_numpy(rest["labels"].astype(np.int32)), torch.from_numpy(O.adjusted_lens(n_steps)),
                         torch.from_numpy(rest["lens"].astype(np.int32)), grad_scale=1.0 / B)
            eng.backward()
            eng.optimizer_step([1e-2] * 3, [0.0, 0.0, 1e-3], 0.9, 0.999, 0.1, 10.0)
            steps = eng.steps_tensor().cpu().numpy().copy()
        torch.cuda.synchronize()
        return m.flat_parameters.detach().cpu().numpy().copy(), steps
    p_grow, s_grow = run(True)
    p_flat, s_flat = run(False)
    assert s_grow.max() == 3 and np.array_equal(s_grow, s_flat)
    # bias / day gradients are accumulated with atomics (summation order varies run to run): equality up to fp32 round-off
    assert np.abs(p_grow - p_flat).max() < 1e-5 * max(1.0, np.abs(p_flat).max())
