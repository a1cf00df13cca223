"""Device-side training feed ("next" row f1): the batches that
`DataLoader(MovieLensDataset(...), batch_size, shuffle, collate_fn=collate_fn)` yields in the reference
(src/training/datasets/movielens.py:22-162, scripts/train_movielens.py:87-120), built on the GPU.

The reference assembles every sample in Python (pandas row lookup, a set difference over all items and
np.random.choice per sample: ~0.45 ms per sample, so the ML-1M epoch is feed-bound); here an epoch is a device-side
permutation plus, per batch, three row gathers and one negative-sampling kernel.  Batches carry the same keys, shapes and
dtypes as collate_fn's; the negatives follow the same distribution (uniform over the user's non-interacted items, without
replacement) from a counter-based generator, not numpy's stream."""
from __future__ import annotations

from typing import Dict, Iterator, List, Mapping, Optional, Sequence

import numpy as np
import torch

from . import kernels as K


def positives_csr(user_positive_items: Mapping[int, Sequence[int]], n_users: int):
    """dict user -> positive items (reference get_user_positive_items) -> (indptr int64 [n_users+1], items int32 sorted)."""
    counts = np.zeros(n_users + 1, dtype=np.int64)
    rows: List[np.ndarray] = [np.empty(0, dtype=np.int32)] * n_users
    for u, items in user_positive_items.items():
        u = int(u)
        if 0 <= u < n_users:
            arr = np.unique(np.asarray(items, dtype=np.int64)).astype(np.int32)  # a set in the reference: dedup, sorted
            rows[u] = arr
            counts[u + 1] = arr.size
    indptr = np.cumsum(counts)
    items = np.concatenate(rows) if n_users > 0 and indptr[-1] > 0 else np.zeros(1, dtype=np.int32)
    return indptr, items


class DeviceInteractionFeed:
    """Iterable of training / validation batches resident on the device.

    user_idx, item_idx [n] int64, labels [n] float: the interactions table (columns user_idx, movie_idx, label);
    user_features [n_users, fu], item_features [n_items, fi] float32: the dataset's precomputed feature matrices;
    user_positive_items: dict user -> items (training only), num_items: size of the item universe negatives come from.
    """

    def __init__(self, user_idx, item_idx, labels, user_features, item_features,
                 user_positive_items: Optional[Mapping[int, Sequence[int]]] = None, num_items: Optional[int] = None,
                 num_negatives: int = 4, batch_size: int = 1024, shuffle: bool = True, is_training: bool = True,
                 seed: Optional[int] = None, device: str = "cuda", drop_last: bool = False):
        dev = torch.device(device)
        if dev.type != "cuda":
            raise RuntimeError("DeviceInteractionFeed builds its batches with CUDA kernels (no CPU path)")
        as_dev = lambda a, dt: torch.as_tensor(np.asarray(a), dtype=dt).to(dev).contiguous()
        self.user_idx = as_dev(user_idx, torch.int64)
        self.item_idx = as_dev(item_idx, torch.int64)
        self.labels = as_dev(labels, torch.float32)
        self.user_features = as_dev(user_features, torch.float32)
        self.item_features = as_dev(item_features, torch.float32)
        self.n = int(self.user_idx.numel())
        self.batch_size, self.shuffle, self.drop_last = int(batch_size), bool(shuffle), bool(drop_last)
        self.is_training = bool(is_training)
        self.num_negatives = int(num_negatives) if is_training else 0
        self.num_items = int(num_items) if num_items is not None else int(self.item_features.shape[0])
        self.seed = int(seed) if seed is not None else int(torch.initial_seed())
        self.epoch = 0
        self._err = torch.zeros((1,), dtype=torch.int32, device=dev)
        self.device = dev
        if self.num_negatives > 0:
            n_users = int(self.user_features.shape[0])
            indptr, items = positives_csr(user_positive_items or {}, n_users)
            smallest_pool = self.num_items - int(np.diff(indptr).max(initial=0))
            if smallest_pool < self.num_negatives:
                # the reference returns a short list here and collate_fn's torch.stack then fails on the ragged batch
                raise ValueError(f"a user has only {smallest_pool} non-interacted items but num_negatives={self.num_negatives}")
            self.pos_indptr = torch.from_numpy(indptr).to(dev)
            self.pos_items = torch.from_numpy(items).to(dev)

    @classmethod
    def from_dataset(cls, dataset, batch_size: int, shuffle: bool = True, **kw) -> "DeviceInteractionFeed":
        """From a reference `MovieLensDataset` (or anything with its attributes): interactions (DataFrame with
        user_idx / movie_idx / label), user_features, movie_features, user_positive_items, num_items, num_negatives,
        is_training."""
        inter = dataset.interactions
        return cls(inter["user_idx"].to_numpy(), inter["movie_idx"].to_numpy(), inter["label"].to_numpy(),
                   dataset.user_features, dataset.movie_features, getattr(dataset, "user_positive_items", None),
                   int(dataset.num_items), int(dataset.num_negatives), batch_size, shuffle, bool(dataset.is_training), **kw)

    def __len__(self) -> int:
        return self.n // self.batch_size if self.drop_last else (self.n + self.batch_size - 1) // self.batch_size

    def __iter__(self) -> Iterator[Dict[str, torch.Tensor]]:
        self.epoch += 1
        if self.shuffle:
            g = torch.Generator(device=self.device).manual_seed((self.seed * 1000003 + self.epoch) & 0x7FFFFFFFFFFFFFFF)
            order = torch.randperm(self.n, device=self.device, generator=g)
        else:
            order = torch.arange(self.n, device=self.device)
        for b in range(len(self)):
            sel = order[b * self.batch_size: min(self.n, (b + 1) * self.batch_size)]
            u = self.user_idx[sel]
            i = self.item_idx[sel]
            batch = {"user_idx": u, "user_features": K.gather_rows(self.user_features, u, self._err),
                     "pos_item_idx": i, "pos_item_features": K.gather_rows(self.item_features, i, self._err)}
            if self.num_negatives > 0:
                neg = K.sample_negatives(u, self.pos_indptr, self.pos_items, self.num_items, self.num_negatives,
                                         self.seed + 0x632BE59BD9B4E019 * self.epoch, b * self.batch_size, self._err)
                nf = K.gather_rows(self.item_features, neg.view(-1), self._err)
                batch["neg_item_indices"] = neg
                batch["neg_item_features"] = nf.view(sel.numel(), self.num_negatives, -1)
            batch["label"] = self.labels[sel]
            yield batch

    def check(self) -> None:
        """One host sync: raise what the reference would have raised while building these batches."""
        code = int(self._err.item())
        if code == 1:
            raise IndexError("feature index out of range while gathering batch rows")
        if code == 2:
            raise ValueError("a user's negative pool is smaller than num_negatives")
