"""Generate tests/golden/dataset_index.json from the UNMODIFIED reference dataset.py (run in the build container only).

The reference class needs h5py only inside __getitem__; the batch-index construction (numpy global RNG) runs with a stub module.
"""
import json, os, sys, types
import numpy as np
sys.modules.setdefault("h5py", types.ModuleType("h5py"))
sys.path.insert(0, "/root/reference/model_training")
from dataset import BrainToTextDataset          # noqa: E402  (the reference)

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden", "dataset_index.json")
rng = np.random.RandomState(0)
n_days = 7
trial_idx = {d: {"trials": sorted(rng.choice(200, size=int(rng.randint(9, 60)), replace=False).tolist()), "session_path": f"/nope/t15.2023.0{d}/data_train.hdf5"}
             for d in range(n_days)}
cases = []
for (bs, dpb, seed, must) in [(64, 4, 10, None), (32, 3, 1, None), (16, 5, 7, [0, -1]), (10, 3, 3, None)]:
    ds = BrainToTextDataset(trial_indicies=trial_idx, n_batches=6, split="train", batch_size=bs, days_per_batch=dpb, random_seed=seed,
                            must_include_days=None if must is None else list(must))
    cases.append({"batch_size": bs, "days_per_batch": dpb, "seed": seed, "must_include_days": must,
                  "index": [[(int(d), [int(t) for t in ts]) for d, ts in ds.batch_index[i].items()] for i in range(6)]})
te = BrainToTextDataset(trial_indicies=trial_idx, n_batches=None, split="test", batch_size=16, days_per_batch=None, random_seed=3)
test_index = [[(int(d), [int(t) for t in ts]) for d, ts in te.batch_index[i].items()] for i in range(len(te))]
json.dump({"trial_idx": {str(d): v for d, v in trial_idx.items()}, "train": cases, "test": test_index}, open(OUT, "w"))
print("wrote", OUT, len(test_index), "test batches")
