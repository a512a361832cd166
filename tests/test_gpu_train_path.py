"""GPU parity tests of the training path (GEMM, forward, CTC, backward, optimizer) through the C ABI.

Checker = oracle/gru_ctc_oracle.py (numpy restatement pinned to the reference by tests/golden) and the
golden vectors themselves.  Tolerances: integer outputs bit-exact; CTC (fp32) 1e-5 relative; anything
that passes through bf16 tensor-core GEMMs is compared at bf16 tolerances stated per test.
"""
import numpy as np
import pytest
import torch

import util

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def E(pkg):
    import b2t_pkg
    return b2t_pkg.submodule("engine")


@pytest.mark.parametrize("M,N,K", [(128, 128, 64), (256, 384, 512), (200, 41, 768), (97 * 16, 192, 448), (130, 136, 72),
                                   # 128 x 256 tiles (forced below): few / many contraction slices, ragged row count
                                   (256, 256, 256), (1024, 768, 256), (200, 512, 768), (6208, 768, 2304)])
@pytest.mark.parametrize("a_mn,b_mn", [(False, False), (False, True), (True, True)])
def test_gemm_vs_torch(E, M, N, K, a_mn, b_mn, monkeypatch):
    if N % 256 == 0:
        monkeypatch.setenv("B2T_GEMM_BN", "256")
    g = torch.Generator(device="cuda").manual_seed(M * 7 + N * 3 + K)
    A = torch.randn(M, K, device="cuda", generator=g).to(torch.bfloat16)
    B = torch.randn(N, K, device="cuda", generator=g).to(torch.bfloat16)
    bias = torch.randn(N, device="cuda", generator=g)
    ref = A.float() @ B.float().t() + bias
    Ain = A.t().contiguous() if a_mn else A
    Bin = B.t().contiguous() if b_mn else B
    if (a_mn and M % 8) or (b_mn and N % 8):
        pytest.skip("MN-major operands need a 16-byte aligned leading dimension")
    out = E.gemm_bf16(Ain, Bin, a_mn=a_mn, b_mn=b_mn, bias=bias)
    torch.cuda.synchronize()
    err = (out - ref).abs().max().item()
    assert err < 1e-2 * max(1.0, ref.abs().max().item() / 50), err      # fp32 accumulation of exact bf16 products
    out16 = E.gemm_bf16(Ain, Bin, a_mn=a_mn, b_mn=b_mn, out_bf16=True, bias=bias) if not (a_mn and b_mn) else None
    if out16 is not None:
        assert (out16.float() - ref).abs().max().item() < 0.02 * ref.abs().max().item() + 1e-2


def _engine_from_golden(E, name, training=True):
    params, grads, p1, rest = util.load_golden(name)
    D, H, L, n_days, B, T = [int(v) for v in rest["cfg"]]
    cfg = E.make_config(D, H, L, n_days, 41, 14, 4, 0.0, 0.0)
    flat = util.flat_from_params(E, cfg, params, "cuda")
    eng = E.Engine(cfg, flat, max_batch=B, max_T=T, max_label_len=16, training=training)
    return eng, cfg, params, grads, p1, rest


@pytest.mark.parametrize("name", ["train_small.npz", "train_ragged.npz"])
def test_forward_logits_vs_golden(E, name):
    eng, cfg, params, grads, p1, rest = _engine_from_golden(E, name)
    x = torch.from_numpy(rest["x"]).cuda()
    logits, hidden = eng.forward(x, torch.from_numpy(rest["days"]), training=False, smooth_mode=1, want_hidden=True)
    torch.cuda.synchronize()
    ref = rest["logits"]
    assert logits.shape == ref.shape
    err = np.abs(logits.cpu().numpy() - ref).max()
    assert err < 5e-2, err                               # bf16 operands, fp32 accumulate (SURVEY tolerance table)
    assert np.abs(hidden.cpu().numpy() - rest["hidden"]).max() < 3e-2


@pytest.mark.parametrize("name", ["train_small.npz", "train_ragged.npz"])
def test_ctc_vs_golden(E, pkg, name):
    import ctypes as C
    params, grads, p1, rest = util.load_golden(name)
    N = pkg._native
    logits = torch.from_numpy(rest["logits"]).cuda()                     # [B,T,C]
    B, T, Cc = logits.shape
    tbc = logits.permute(1, 0, 2).contiguous()
    labels = torch.from_numpy(rest["labels"]).to(torch.int32).cuda()
    import gru_ctc_oracle as O
    in_len = torch.from_numpy(O.adjusted_lens(rest["n_steps"])).to(torch.int32).cuda()
    tgt = torch.from_numpy(rest["lens"]).to(torch.int32).cuda()
    ws = torch.empty(N.lib.b2t_ctc_workspace_bytes(T, B, labels.shape[1]), dtype=torch.uint8, device="cuda")
    loss = torch.empty(B, device="cuda")
    dl = torch.empty_like(tbc)
    N.check(N.lib.b2t_ctc_loss_tbc(tbc.data_ptr(), T, B, Cc, labels.data_ptr(), labels.shape[1], in_len.data_ptr(), tgt.data_ptr(),
                                   1.0 / B, loss.data_ptr(), dl.data_ptr(), ws.data_ptr(), ws.numel(),
                                   torch.cuda.current_stream().cuda_stream), "ctc")
    torch.cuda.synchronize()
    assert util.rel_err(loss.cpu().numpy(), rest["loss_vec"]) < 1e-5
    got = dl.permute(1, 0, 2).cpu().numpy()
    assert np.abs(got - rest["dlogits"]).max() < 1e-5 * max(1.0, np.abs(rest["dlogits"]).max() * 10)


@pytest.fixture(params=["default", "chunks3_lanes2", "chunks2_lanes1"])
def rec_schedule(request, monkeypatch):
    """Recurrence schedules: default, and forced time-chunking / lane counts so that the chunk carry of h and dh and the
    multi-stream wave-front are exercised on the small golden shapes too."""
    if request.param == "chunks3_lanes2":
        monkeypatch.setenv("B2T_REC_CHUNKS", "3"); monkeypatch.setenv("B2T_REC_LANES_FWD", "2"); monkeypatch.setenv("B2T_REC_LANES_BWD", "2")
    elif request.param == "chunks2_lanes1":
        monkeypatch.setenv("B2T_REC_CHUNKS", "2"); monkeypatch.setenv("B2T_REC_LANES_FWD", "1"); monkeypatch.setenv("B2T_REC_LANES_BWD", "1")
    return request.param


@pytest.mark.parametrize("name", ["train_small.npz", "train_ragged.npz"])
def test_train_step_vs_golden(E, name, rec_schedule):
    """forward -> CTC -> backward -> clip+AdamW against the reference's own step (golden)."""
    import gru_ctc_oracle as O
    eng, cfg, params, grads, p1, rest = _engine_from_golden(E, name)
    x = torch.from_numpy(rest["x"]).cuda()
    B = x.shape[0]
    logits, _ = eng.forward(x, torch.from_numpy(rest["days"]), training=True, smooth_mode=1)
    in_len = torch.from_numpy(O.adjusted_lens(rest["n_steps"]))
    loss = eng.ctc_loss(torch.from_numpy(rest["labels"]), in_len, torch.from_numpy(rest["lens"]), grad_scale=1.0 / B)
    eng.backward()
    torch.cuda.synchronize()
    assert util.rel_err(loss.cpu().numpy(), rest["loss_vec"]) < 2e-2      # logits carry bf16 error
    got = util.unflatten(E, cfg, eng.grads[:eng.n_params])
    errs = {k: util.rel_err(got[k].reshape(g.shape), g) for k, g in grads.items()}
    print("GRAD_ERR", name, rec_schedule, "max", round(max(errs.values()), 5), max(errs, key=errs.get))
    bad = {k: round(r, 4) for k, r in errs.items() if r >= util.GRAD_TOL}           # bf16 GEMM operands / bf16 activation stash
    assert not bad, bad
    touched = eng.touched_days().cpu().numpy()
    assert sorted(np.nonzero(touched)[0].tolist()) == sorted(set(int(d) for d in rest["days"]))
    for k in got:                                                        # untouched day layers: zero grad, skipped by AdamW
        if k not in grads:
            assert np.abs(got[k]).max() == 0.0, k
    lr = float(rest["lr"])
    stats = eng.optimizer_step([lr] * 3, [0.0, 0.0, 1e-3], 0.9, 0.999, 0.1, 10.0)
    torch.cuda.synchronize()
    assert abs(stats[0].item() - float(rest["grad_norm"])) < 3e-2 * float(rest["grad_norm"])
    newp = util.unflatten(E, cfg, eng.params)
    for k, v in p1.items():
        # after one step the update is lr * m_hat/(sqrt(v_hat)+eps); compare the *delta* at bf16-gradient tolerance
        d_ref = v - params[k]
        d_got = newp[k].reshape(v.shape) - params[k]
        assert np.abs(d_got - d_ref).max() < util.GRAD_TOL * np.abs(d_ref).max() + 1e-7, k
    for k in newp:
        if k not in p1:
            assert np.array_equal(newp[k].reshape(params[k].shape), params[k].astype(np.float32)), k


def test_nonfinite_gradient_norm_leaves_parameters_untouched(E):
    """clip_grad_norm_(error_if_nonfinite=True) raises before optimizer.step() in the reference (rnn_trainer.py:550-558): with a
    non-finite norm the fused clip+AdamW must not touch parameters, moments or step counters, and must report the norm."""
    import gru_ctc_oracle as O
    eng, cfg, params, grads, p1, rest = _engine_from_golden(E, "train_small.npz")
    x = torch.from_numpy(rest["x"]).cuda()
    eng.forward(x, torch.from_numpy(rest["days"]), training=True, smooth_mode=1)
    eng.ctc_loss(torch.from_numpy(rest["labels"]), torch.from_numpy(O.adjusted_lens(rest["n_steps"])), torch.from_numpy(rest["lens"]), grad_scale=1.0 / x.shape[0])
    eng.backward()
    torch.cuda.synchronize()
    good = eng.grads[:eng.n_params].clone()
    before, steps = eng.params.clone(), eng.steps_tensor().clone()
    eng.grads[7] = float("inf")
    stats = eng.optimizer_step([1e-3] * 3, [0.0, 0.0, 1e-3], 0.9, 0.999, 0.1, 10.0)
    torch.cuda.synchronize()
    assert not torch.isfinite(stats[0]).item()
    assert torch.equal(eng.params, before) and torch.equal(eng.steps_tensor(), steps)
    eng.grads[:eng.n_params] = good                                     # the same step with finite gradients goes through
    stats = eng.optimizer_step([1e-3] * 3, [0.0, 0.0, 1e-3], 0.9, 0.999, 0.1, 10.0)
    torch.cuda.synchronize()
    assert torch.isfinite(stats[0]).item() and not torch.equal(eng.params, before)


def test_greedy_edit_bit_exact(E):
    import gru_ctc_oracle as O
    eng, cfg, params, grads, p1, rest = _engine_from_golden(E, "train_ragged.npz", training=False)
    x = torch.from_numpy(rest["x"]).cuda()
    logits, _ = eng.forward(x, torch.from_numpy(rest["days"]), training=False, smooth_mode=1)
    in_len = torch.from_numpy(O.adjusted_lens(rest["n_steps"]))
    dec, dlen, ed = eng.greedy_edit(torch.from_numpy(rest["labels"]), in_len, torch.from_numpy(rest["lens"]))
    torch.cuda.synchronize()
    lg = logits.cpu().numpy()
    for b in range(lg.shape[0]):
        want = O.greedy_decode(lg[b], int(in_len[b]))                    # same logits -> integer pipeline must be identical
        assert dec[b, :dlen[b]].cpu().tolist() == want
        assert int(ed[b]) == O.edit_distance(want, rest["labels"][b][:rest["lens"][b]])


def _random_params(E, cfg, seed):
    rng = np.random.RandomState(seed)
    params = {}
    for name, off, rows, cols in E.param_layout(cfg):
        if name.startswith("day_weights"):
            v = np.eye(rows, cols) + 0.05 * rng.randn(rows, cols)
        elif "bias" in name or name == "h0":
            v = 0.1 * rng.randn(rows, cols)
        else:
            v = rng.randn(rows, cols) / np.sqrt(cols)
        params[name] = v.astype(np.float32)
    return params


@pytest.mark.gpu
@pytest.mark.parametrize("B,bg_cap,chunks,nsub,H", [(64, 64, 3, 1, 192), (64, 32, 2, 0, 192), (64, 32, 2, 1, 192), (64, 32, 1, 2, 192), (64, 16, 1, 1, 192),
                                                    (64, 16, 3, 4, 192), (32, 64, 2, 0, 192), (128, 64, 3, 1, 192), (128, 32, 2, 0, 192), (40, 64, 2, 0, 192),
                                                    (64, 32, 3, 0, 256), (64, 16, 2, 0, 256), (24, 32, 1, 0, 256), (32, 32, 2, 0, 512),
                                                    # whole-stack persistent kernels (gru_stack.cuh; default when H % 256 == 0): bg_cap < 0 selects them
                                                    (64, -1, 0, 0, 256), (128, -1, 0, 0, 256), (24, -1, 0, 0, 256), (40, -1, 0, 0, 256), (96, -1, 0, 0, 256),
                                                    (16, -1, 0, 0, 512), (64, -1, 0, 0, 512)])
def test_wide_batch_train_step_vs_oracle(E, monkeypatch, B, bg_cap, chunks, nsub, H):
    """Batch-group widths 16 / 32 / 64 of the recurrence kernels (forward: trials per CTA = MMA N; backward: partial
    exchange geometry), several time chunks, against the numpy oracle evaluated on the same inputs.  H = 192 runs the
    contraction-split backward kernel (two 128-row output blocks, the second one partial); H = 256 / 512 run the
    two-dimensional one (gru_rec_bwd2.cuh: dG all-gather + 4-way partial reduction)."""
    import gru_ctc_oracle as O
    monkeypatch.setenv("B2T_STACK", "1" if bg_cap < 0 else "0")       # bg_cap >= 0: the per-(layer, time chunk) launches of gru_rec.cuh
    if bg_cap >= 0:
        monkeypatch.setenv("B2T_REC_BG_FWD", str(bg_cap)); monkeypatch.setenv("B2T_REC_BG_BWD", str(bg_cap))
        monkeypatch.setenv("B2T_REC_CHUNKS", str(chunks))
    if nsub:                                                     # batch groups sharing one CTA in the backward recurrence (0 = default choice)
        monkeypatch.setenv("B2T_REC_NSUB_BWD", str(nsub))
    if H % 256 == 0 and bg_cap >= 0:                             # opt-in two-dimensional backward kernel
        monkeypatch.setenv("B2T_REC_BWD2", "1")
    D, L, n_days, T = 32, 3, 4, 74
    cfg = E.make_config(D, H, L, n_days, 41, 14, 4, 0.0, 0.0)
    params = _random_params(E, cfg, 7)
    shapes = {"h0": (1, 1, H)}
    P = O.Params({k: (v.reshape(shapes[k]) if k in shapes else (v.reshape(-1) if "bias" in k and not k.startswith("day") else v))
                  for k, v in params.items()})
    rng = np.random.RandomState(11)
    x = rng.randn(B, T, D).astype(np.float32)
    n_steps = rng.randint(60, T + 1, size=B); n_steps[0] = T
    for b in range(B):
        x[b, n_steps[b]:] = 0
    lens = rng.randint(2, 6, size=B)
    labels = np.zeros((B, 8), dtype=np.int64)
    for b in range(B):
        labels[b, :lens[b]] = rng.randint(1, 41, size=lens[b])
    days = rng.randint(0, n_days - 1, size=B)                    # day n_days-1 stays untouched
    eng = E.Engine(cfg, util.flat_from_params(E, cfg, params, "cuda"), max_batch=B, max_T=T, max_label_len=8, training=True)
    logits, _ = eng.forward(torch.from_numpy(x).cuda(), torch.from_numpy(days.astype(np.int32)), training=True, smooth_mode=1)
    in_len = O.adjusted_lens(n_steps)
    loss = eng.ctc_loss(torch.from_numpy(labels.astype(np.int32)), torch.from_numpy(in_len.astype(np.int32)),
                        torch.from_numpy(lens.astype(np.int32)), grad_scale=1.0 / B)
    eng.backward()
    torch.cuda.synchronize()
    xs, _ = O.transform_data(x, n_steps, mode="val")
    ref_logits, _, cache = O.forward(P, xs, days, keep_cache=True)
    assert np.abs(logits.cpu().numpy() - ref_logits).max() < 5e-2
    ref_loss, dlog = O.ctc_loss_and_grad(ref_logits, labels, in_len, lens)
    assert util.rel_err(loss.cpu().numpy(), ref_loss) < 2e-2
    ref_g = O.backward(P, cache, dlog, days)          # dlog already carries the 1/B of the batch mean
    got = util.unflatten(E, cfg, eng.grads[:eng.n_params])
    bad = {}
    for k, g in ref_g.items():
        r = util.rel_err(got[k].reshape(np.asarray(g).shape), g)
        if r >= util.GRAD_TOL:
            bad[k] = round(r, 4)
    assert not bad, bad
    assert np.abs(got[f"day_weights.{n_days - 1}"]).max() == 0.0
