"""GPU tests of the reference-facing Python API (rnn_model / rnn_trainer / data_augmentations / helpers).

They read like the reference's own usage: build GRUDecoder, call model(x, day_idx), use
torch-style CTC + backward + optimizer, run BrainToTextDecoder_Trainer.train() on a synthetic corpus."""
import os

import numpy as np
import pytest
import torch

import util

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def mods(pkg):
    import b2t_pkg
    return {m: b2t_pkg.submodule(m) for m in ("rnn_model", "rnn_trainer", "ctc", "data_augmentations", "evaluate_model_helpers", "engine")}


def _model_from_golden(mods, name, **kw):
    params, grads, p1, rest = util.load_golden(name)
    D, H, L, n_days, B, T = [int(v) for v in rest["cfg"]]
    m = mods["rnn_model"].GRUDecoder(neural_dim=D, n_units=H, n_days=n_days, n_classes=41, n_layers=L, patch_size=14, patch_stride=4, **kw)
    sd = {k: torch.from_numpy(v).reshape(m.state_dict()[k].shape) for k, v in params.items()}
    m.load_state_dict(sd)
    return m.to("cuda"), params, grads, rest


def test_gauss_smooth_matches_reference(mods):
    z = np.load(os.path.join(util.GOLDEN, "smooth.npz"))
    gs = mods["data_augmentations"].gauss_smooth
    x = torch.from_numpy(z["x"]).cuda()
    assert np.abs(gs(x, "cuda", 2, 100).cpu().numpy() - z["same"]).max() < 2e-6
    v = gs(x, "cuda", 2, 100, padding="valid")
    assert v.shape == z["valid"].shape and np.abs(v.cpu().numpy() - z["valid"]).max() < 2e-6


def test_model_forward_and_state(mods):
    m, params, grads, rest = _model_from_golden(mods, "train_ragged.npz")
    m.eval()
    gs = mods["data_augmentations"].gauss_smooth
    x = gs(torch.from_numpy(rest["x"]).cuda(), "cuda", 2, 100)
    with torch.no_grad():
        logits, hidden = m(x, torch.from_numpy(rest["days"]).cuda(), None, True)
        assert np.abs(logits.cpu().numpy() - rest["logits"]).max() < 5e-2
        assert np.abs(hidden.cpu().numpy() - rest["hidden"]).max() < 3e-2
        # streaming: feeding the returned state back must continue the sequence (rnn_model.py:88,122-126)
        l2 = m(x, list(int(d) for d in rest["days"]), states=hidden)
        assert l2.shape == logits.shape and not torch.allclose(l2, logits)


def test_autograd_path_matches_reference_grads(mods):
    """model(x) -> log_softmax -> ctc_loss -> mean -> backward, exactly the trainer's statement sequence."""
    import gru_ctc_oracle as O
    m, params, grads, rest = _model_from_golden(mods, "train_small.npz")
    m.train()
    gs = mods["data_augmentations"].gauss_smooth
    x = gs(torch.from_numpy(rest["x"]).cuda(), "cuda", 2, 100)
    logits = m(x, torch.from_numpy(rest["days"]).cuda())
    adj = torch.from_numpy(O.adjusted_lens(rest["n_steps"])).cuda()
    loss = mods["ctc"].ctc_loss(torch.permute(logits.log_softmax(2), [1, 0, 2]), torch.from_numpy(rest["labels"]).cuda(), adj,
                                torch.from_numpy(rest["lens"]).cuda())
    assert util.rel_err(loss.detach().cpu().numpy(), rest["loss_vec"]) < 2e-2
    torch.mean(loss).backward()
    got = {n: p.grad for n, p in m.named_parameters()}
    for k, g in grads.items():
        assert got[k] is not None, k
        assert util.rel_err(got[k].cpu().numpy().reshape(g.shape), g) < util.GRAD_TOL, k
    for k, g in got.items():
        if k not in grads:
            assert g is None, k                       # day layers absent from the batch keep grad=None like the reference
    gn = torch.nn.utils.clip_grad_norm_(m.parameters(), max_norm=10, error_if_nonfinite=True, foreach=True)
    assert abs(float(gn) - float(rest["grad_norm"])) < 3e-2 * float(rest["grad_norm"])


@pytest.mark.parametrize("T,B,S,cap", [(40, 5, 9, None), (97, 8, 45, None), (30, 4, 3, None), (170, 3, 80, None), (60, 4, 20, 70)])
def test_ctc_loss_matches_torch(mods, T, B, S, cap):
    """Both CTC kernels (warp-per-trial for 2S+1 <= 128, block-per-trial beyond) against torch in float64.
    Covers empty targets, repeated labels, ragged input lengths and an infeasible-looking short input."""
    g = torch.Generator().manual_seed(T * 31 + S)
    C = 41
    lp = torch.randn(T, B, C, generator=g).log_softmax(2).cuda().requires_grad_()
    Sw = cap or S                                              # padded label width decides which kernel runs
    tg = torch.randint(1, C, (B, Sw), generator=g)
    tg[0, 1:4] = tg[0, 0]                                      # repeated labels need blanks in between
    il = torch.randint(max(2 * S + 2, T // 2), T + 1, (B,), generator=g).clamp(max=T)
    il[0] = T
    tl = torch.randint(1, S + 1, (B,), generator=g)
    tl[-1] = 0                                                 # empty target
    tl[0] = min(S, 6)
    ours = mods["ctc"].ctc_loss(lp, tg.cuda(), il.cuda(), tl.cuda())
    ours.sum().backward()
    # checker in float64 so that the comparison measures OUR fp32 log-space error, not the sum of two fp32 errors
    ref_lp = lp.detach().cpu().double().requires_grad_()
    ref = torch.nn.functional.ctc_loss(ref_lp, tg, il, tl, blank=0, reduction="none", zero_infinity=False)
    ref.sum().backward()
    assert util.rel_err(ours.detach().cpu().numpy(), ref.detach().numpy()) < 1e-5
    # |grad| <= 1; the fp32 recursion's rounding random-walks with the number of frames: 1e-5 up to 64 frames, 2e-5 beyond
    assert np.abs(lp.grad.cpu().numpy() - ref_lp.grad.numpy()).max() < (1e-5 if T <= 64 else 2e-5)


def _trainer_args(tmp, n_batches):
    sessions = [f"s{i}" for i in range(6)]
    return {
        "mode": "train", "output_dir": os.path.join(tmp, "out"), "checkpoint_dir": os.path.join(tmp, "out", "checkpoint"),
        "save_best_checkpoint": True, "save_all_val_steps": False, "save_final_model": False, "save_val_metrics": True,
        "early_stopping": False, "early_stopping_val_steps": 20, "gpu_number": "0", "seed": 10, "use_amp": True,
        "init_from_checkpoint": False, "init_checkpoint_path": None,
        "model": {"n_input_features": 64, "n_units": 128, "rnn_dropout": 0.2, "rnn_trainable": True, "n_layers": 2, "patch_size": 14,
                  "patch_stride": 4, "input_network": {"input_layer_dropout": 0.1, "input_trainable": True}},
        "num_training_batches": n_batches, "lr_scheduler_type": "cosine", "lr_max": 0.01, "lr_min": 0.001, "lr_decay_steps": n_batches,
        "lr_warmup_steps": 10, "lr_max_day": 0.01, "lr_min_day": 0.001, "lr_decay_steps_day": n_batches, "lr_warmup_steps_day": 10,
        "beta0": 0.9, "beta1": 0.999, "epsilon": 0.1, "weight_decay": 0.001, "weight_decay_day": 0, "grad_norm_clip_value": 10,
        "batches_per_train_log": 50, "batches_per_val_step": n_batches - 1, "log_individual_day_val_PER": False, "log_val_skip_logs": False,
        "save_val_logits": False, "save_val_data": False,
        "dataset": {"data_transforms": {"white_noise_std": 0.3, "constant_offset_std": 0.1, "random_walk_std": 0.0, "random_walk_axis": -1,
                                        "static_gain_std": 0.0, "random_cut": 3, "smooth_kernel_size": 100, "smooth_data": True, "smooth_kernel_std": 2},
                    "neural_dim": 64, "batch_size": 16, "n_classes": 41, "days_per_batch": 2, "seed": 1, "num_dataloader_workers": 0,
                    "loader_shuffle": False, "sessions": sessions, "dataset_probability_val": [1] * len(sessions),
                    "synthetic": {"T": 160, "min_len": 3, "max_len": 6, "noise": 0.3, "val_batches": 3}},
    }


def test_trainer_learns_synthetic_corpus(mods, tmp_path):
    T = mods["rnn_trainer"].BrainToTextDecoder_Trainer
    args = _trainer_args(str(tmp_path), 300)
    tr = T(args)
    stats = tr.train()
    assert len(stats["train_losses"]) == 300 and len(stats["val_PERs"]) == 2
    first, last = np.mean(stats["train_losses"][:10]), np.mean(stats["train_losses"][-10:])
    assert last < 0.5 * first, (first, last)
    assert stats["val_PERs"][-1] < stats["val_PERs"][0]
    assert os.path.exists(os.path.join(args["checkpoint_dir"], "best_checkpoint"))
    ck = torch.load(os.path.join(args["checkpoint_dir"], "best_checkpoint"), weights_only=False)
    assert set(ck) == {"model_state_dict", "optimizer_state_dict", "scheduler_state_dict", "val_PER", "val_loss"}
    assert "gru.weight_ih_l0" in ck["model_state_dict"] and len(ck["optimizer_state_dict"]["param_groups"]) == 3
    m = tr.validation(tr.val_loader, return_logits=True)
    for k in ("avg_PER", "avg_loss", "day_PERs", "decoded_seqs", "true_seq", "logits", "n_time_steps"):
        assert k in m


def test_trainer_on_session_files(mods, tmp_path, pkg):
    """The reference's data path: per-session files (here .npz shards of the hdf5 layout) -> BrainToTextDataset with the reference's
    day / trial sampling -> pinned prefetch -> fused training step; labels arrive padded to 500 like the reference's files."""
    import b2t_pkg
    DS = b2t_pkg.submodule("dataset")
    synth = b2t_pkg.submodule("datasets").SyntheticBrainToTextDataset(n_batches=1, batch_size=24, days_per_batch=1, n_days=4, neural_dim=64,
                                                                      T=120, min_len=3, max_len=6, noise=0.3, seed=3)
    args = _trainer_args(str(tmp_path), 120)
    del args["dataset"]["synthetic"]
    args["dataset"]["sessions"] = [f"t15.2023.08.{10 + d}" for d in range(4)]
    args["dataset"]["dataset_probability_val"] = [1] * 4
    args["dataset"]["dataset_dir"] = str(tmp_path / "data")
    args["dataset"]["max_time_steps"] = 128
    args["batches_per_val_step"] = 119
    rng = np.random.RandomState(0)
    for d, sess in enumerate(args["dataset"]["sessions"]):
        os.makedirs(os.path.join(args["dataset"]["dataset_dir"], sess))
        for split, n in (("train", 40), ("val", 10)):
            trials = []
            for t in range(n):
                x, lab, nst = synth._trial(rng, d)
                ids = np.zeros(500, dtype=np.int64); ids[:len(lab)] = lab
                trials.append({"input_features": x[:nst], "seq_class_ids": ids, "transcription": np.zeros(500, dtype=np.int64),
                               "n_time_steps": nst, "seq_len": len(lab), "block_num": 1, "trial_num": t})
            DS.write_session_npz(os.path.join(args["dataset"]["dataset_dir"], sess, f"data_{split}.npz"), trials)
    tr = mods["rnn_trainer"].BrainToTextDecoder_Trainer(args)
    stats = tr.train()
    assert len(stats["train_losses"]) == 120
    assert np.mean(stats["train_losses"][-10:]) < 0.7 * np.mean(stats["train_losses"][:10])
    assert 0.0 <= stats["val_PERs"][-1] <= stats["val_PERs"][0]
    assert os.path.exists(os.path.join(args["output_dir"], "train_val_trials.json"))


def test_run_single_decoding_step(mods):
    m, params, grads, rest = _model_from_golden(mods, "train_small.npz")
    m.eval()
    H = mods["evaluate_model_helpers"]
    args = {"use_amp": True, "dataset": {"data_transforms": {"smooth_kernel_std": 2, "smooth_kernel_size": 100}}}
    x = torch.from_numpy(rest["x"][:1]).cuda().to(torch.bfloat16)
    lg = H.runSingleDecodingStep(x, int(rest["days"][0]), m, args, "cuda")
    Tv = rest["x"].shape[1] - 8
    assert lg.dtype == np.float32 and lg.shape == (1, (Tv - 14) // 4 + 1, 41)
    r = H.rearrange_speech_logits_pt(lg)
    assert np.array_equal(r[..., 1], lg[..., 40]) and np.array_equal(r[..., 2:], lg[..., 1:40])
