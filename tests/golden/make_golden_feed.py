"""Generates tests/golden/feed_small.npz by running the IMPORTED reference (MovieLensDataset.__getitem__ + collate_fn,
sample_negative_items) on a small synthetic interaction table.  Run here (needs /root/reference):
    python tests/golden/make_golden_feed.py"""
import os, sys
import numpy as np
import pandas as pd
sys.path.insert(0, "/root/reference")
from src.training.datasets.movielens import collate_fn                      # noqa: E402
from src.data.movielens import get_user_positive_items, sample_negative_items  # noqa: E402
import torch                                                                 # noqa: E402

rng = np.random.default_rng(42)
n_users, n_items, n_inter, R = 23, 57, 300, 5
inter = pd.DataFrame({"user_idx": rng.integers(0, n_users - 2, n_inter), "movie_idx": rng.integers(0, n_items, n_inter),
                      "label": rng.integers(0, 2, n_inter).astype(np.float64)})
uf = rng.standard_normal((n_users, 3)).astype(np.float32)
mf = rng.standard_normal((n_items, 20)).astype(np.float32)
pos = get_user_positive_items(inter)
rows = rng.permutation(n_inter)[:64]
np.random.seed(1234)
samples = []
for r in rows:   # MovieLensDataset.__getitem__ (datasets/movielens.py:86-133) with the precomputed matrices given directly
    row = inter.iloc[int(r)]
    u, i = int(row["user_idx"]), int(row["movie_idx"])
    neg = sample_negative_items(u, pos, n_items, R)
    samples.append({"user_idx": u, "user_features": torch.tensor(uf[u], dtype=torch.float32), "pos_item_idx": i,
                    "pos_item_features": torch.tensor(mf[i], dtype=torch.float32),
                    "neg_item_indices": torch.tensor(neg, dtype=torch.long),
                    "neg_item_features": torch.tensor(mf[neg], dtype=torch.float32),
                    "label": torch.tensor(float(row["label"]), dtype=torch.float32)})
batch = collate_fn(samples)
out = {"in_user_idx": inter["user_idx"].to_numpy(), "in_item_idx": inter["movie_idx"].to_numpy(),
       "in_label": inter["label"].to_numpy(), "in_uf": uf, "in_mf": mf, "rows": rows, "n_items": n_items, "R": R, "seed": 1234}
for k, v in batch.items():
    out["out_" + k] = v.numpy()
np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "feed_small.npz"), **out)
print({k: (v.shape if hasattr(v, "shape") else v) for k, v in out.items()})
