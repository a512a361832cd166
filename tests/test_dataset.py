"""CPU tests of the input pipeline (nejm-brain-to-text_b200/dataset.py) against the reference's dataset.py semantics.

tests/golden/dataset_index.json was produced by oracle/gen_dataset_golden.py, which runs the UNMODIFIED reference class
(h5py stubbed: the batch index is pure numpy-RNG work): same seed => the same batch index, draw for draw."""
import importlib.util
import json
import os

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def DS():
    spec = importlib.util.spec_from_file_location("b2t_dataset", os.path.join(ROOT, "nejm-brain-to-text_b200", "dataset.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


@pytest.fixture(scope="module")
def golden():
    g = json.load(open(os.path.join(ROOT, "tests", "golden", "dataset_index.json")))
    g["trial_idx"] = {int(d): v for d, v in g["trial_idx"].items()}
    return g


def test_train_batch_index_matches_reference(DS, golden):
    for case in golden["train"]:
        must = None if case["must_include_days"] is None else list(case["must_include_days"])
        ds = DS.BrainToTextDataset(trial_indicies=golden["trial_idx"], n_batches=6, split="train", batch_size=case["batch_size"],
                                   days_per_batch=case["days_per_batch"], random_seed=case["seed"], must_include_days=must)
        got = [[(int(d), [int(t) for t in ts]) for d, ts in ds.batch_index[i].items()] for i in range(6)]
        want = [[(d, ts) for d, ts in b] for b in case["index"]]
        assert got == want
        for b in got:                                            # the semantics the reference documents
            assert sum(len(ts) for _, ts in b) == case["batch_size"] and len(b) == case["days_per_batch"]


def test_test_batch_index_matches_reference_and_covers_once(DS, golden):
    ds = DS.BrainToTextDataset(trial_indicies=golden["trial_idx"], n_batches=None, split="test", batch_size=16, days_per_batch=None, random_seed=3)
    got = [[(int(d), [int(t) for t in ts]) for d, ts in ds.batch_index[i].items()] for i in range(len(ds))]
    assert got == [[(d, ts) for d, ts in b] for b in golden["test"]]
    seen = [(d, t) for b in got for d, ts in b for t in ts]
    assert len(seen) == len(set(seen)) == sum(len(v["trials"]) for v in golden["trial_idx"].values())
    assert all(len(b) == 1 for b in got)                         # validation batches are single-day


def _make_sessions(DS, tmp, n_days=3, n_trials=12, D=16):
    rng = np.random.RandomState(5)
    paths = []
    for d in range(n_days):
        sess = os.path.join(tmp, f"t15.2023.08.{d + 10}")
        os.makedirs(sess)
        trials = []
        for t in range(n_trials):
            T, S = int(rng.randint(20, 60)), int(rng.randint(2, 9))
            ids = np.zeros(500, dtype=np.int64); ids[:S] = rng.randint(1, 41, size=S)
            tr = np.zeros(500, dtype=np.int64); tr[:5] = [104, 101, 108, 108, 111]
            trials.append({"input_features": rng.randn(T, D).astype(np.float32), "seq_class_ids": ids, "transcription": tr,
                           "n_time_steps": T, "seq_len": S, "block_num": 1 + t // 6, "trial_num": t})
        DS.write_session_npz(os.path.join(sess, "data_train.npz"), trials)
        paths.append(os.path.join(sess, "data_train.hdf5"))      # the reference's file name: resolved to the .npz twin
    return paths


def test_batches_from_npz_shards(DS, tmp_path):
    paths = _make_sessions(DS, str(tmp_path))
    tr, te = DS.train_test_split_indicies(paths, test_percentage=0.25, seed=1, bad_trials_dict={"t15.2023.08.10": {"1": [0, 1]}})
    assert all(len(te[d]["trials"]) == max(1, int(len(tr[d]["trials"]) + len(te[d]["trials"])) // 4) for d in te)
    assert 0 not in tr[0]["trials"] + te[0]["trials"] and 1 not in tr[0]["trials"] + te[0]["trials"]      # bad trials excluded
    assert not set(tr[1]["trials"]) & set(te[1]["trials"])
    ds = DS.BrainToTextDataset(trial_indicies=tr, n_batches=4, split="train", batch_size=8, days_per_batch=2, random_seed=2)
    b = ds[0]
    assert set(b) == {"input_features", "seq_class_ids", "n_time_steps", "phone_seq_lens", "day_indicies", "transcriptions", "block_nums", "trial_nums"}
    assert b["input_features"].shape[0] == 8 and b["input_features"].dtype == torch.float32 and b["input_features"].shape[2] == 16
    assert b["input_features"].shape[1] == int(b["n_time_steps"].max()) and b["seq_class_ids"].shape == (8, 500)
    assert len(set(b["day_indicies"].tolist())) == 2
    for i in range(8):                                           # zero padding behind every trial's own length
        assert torch.all(b["input_features"][i, int(b["n_time_steps"][i]):] == 0)
        assert torch.all(b["seq_class_ids"][i, int(b["phone_seq_lens"][i]):] == 0)
    sub = DS.BrainToTextDataset(trial_indicies=tr, n_batches=1, split="train", batch_size=4, days_per_batch=1, random_seed=2, feature_subset=[0, 3, 5])
    assert sub[0]["input_features"].shape[2] == 3


def test_pinned_loader_yields_the_same_batches(DS, tmp_path):
    paths = _make_sessions(DS, str(tmp_path))
    tr, _ = DS.train_test_split_indicies(paths, test_percentage=0, seed=1)
    ds = DS.BrainToTextDataset(trial_indicies=tr, n_batches=7, split="train", batch_size=8, days_per_batch=2, random_seed=4)
    loader = DS.PinnedBatchLoader(ds, max_T=64, neural_dim=16, depth=2, pin=False)
    n = 0
    for i, b in enumerate(loader):
        ref = ds[i]
        for k in ref:
            assert torch.equal(b[k], ref[k]), k
        n += 1
    assert n == 7
