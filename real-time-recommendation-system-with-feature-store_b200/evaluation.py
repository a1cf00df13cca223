"""Drop-in B200 mirror of the offline evaluation path:
`Evaluator.evaluate` / `Evaluator.evaluate_model` and `EvaluationMetrics` (reference src/evaluation/metrics.py:21-71,
:233-399) and the recommendation twin `generate_recommendations` (scripts/evaluate_model.py:162-234).

`evaluate_model` never leaves the GPU between the towers and the metric table: item tower over the catalogue ->
`FlatIPDeviceIndex` (fp32-grade rows) -> user tower per batch -> fused scoring + top-K kernel with the user's training
items excluded by a CSR (the reference's `-inf` mask) -> `b200rec_eval_metrics` (recall / precision / NDCG / hit rate
at every k, MRR, MAP, coverage) on the [Q, K] id matrix.  Only the per-user metric rows (fp64, 22 numbers per user)
come back, because `EvaluationMetrics.per_user_*` are part of the reference's result; their means are taken with
numpy exactly as the reference does.  CUDA only — there is no CPU path.
"""
from __future__ import annotations

import logging
from dataclasses import dataclass, field
from typing import Dict, Iterable, List, Optional, Sequence, Set, Tuple

import numpy as np
import torch

from . import _native as N
from .retrieval import FlatIPDeviceIndex

logger = logging.getLogger("b200rec")


@dataclass
class EvaluationMetrics:
    """Container for evaluation metrics at multiple K values (reference metrics.py:21-71, same fields)."""

    recall: Dict[int, float] = field(default_factory=dict)
    precision: Dict[int, float] = field(default_factory=dict)
    ndcg: Dict[int, float] = field(default_factory=dict)
    hit_rate: Dict[int, float] = field(default_factory=dict)
    mrr: float = 0.0
    map_score: float = 0.0
    coverage: float = 0.0
    per_user_recall: Dict[int, List[float]] = field(default_factory=dict)
    per_user_ndcg: Dict[int, List[float]] = field(default_factory=dict)

    def to_dict(self) -> Dict[str, float]:
        out: Dict[str, float] = {}
        for name, table in (("recall", self.recall), ("precision", self.precision), ("ndcg", self.ndcg),
                            ("hit_rate", self.hit_rate)):
            for k, v in table.items():
                out[f"{name}@{k}"] = v
        out["mrr"] = self.mrr
        out["map"] = self.map_score
        out["coverage"] = self.coverage
        return out

    def __str__(self) -> str:
        bar = "=" * 50
        lines = [bar, "Evaluation Results", bar]
        for k in sorted(self.recall):
            lines += [f"@{k}:", f"  Recall:    {self.recall[k]:.4f}", f"  Precision: {self.precision[k]:.4f}",
                      f"  NDCG:      {self.ndcg[k]:.4f}", f"  Hit Rate:  {self.hit_rate[k]:.4f}"]
        lines += ["-" * 50, f"MRR:      {self.mrr:.4f}", f"MAP:      {self.map_score:.4f}",
                  f"Coverage: {self.coverage:.4f}", bar]
        return "\n".join(lines)


def _csr(lists: Sequence[np.ndarray]) -> Tuple[np.ndarray, np.ndarray]:
    indptr = np.zeros(len(lists) + 1, dtype=np.int64)
    if lists:
        np.cumsum([len(x) for x in lists], out=indptr[1:])
    flat = np.concatenate(lists).astype(np.int64) if indptr[-1] > 0 else np.zeros((1,), dtype=np.int64)
    return indptr, flat


def _dcg_tables(k_max: int) -> Tuple[np.ndarray, np.ndarray]:
    """inv_log2[i] = 1/log2(i+2) and idcg[n] = sum_{i<n} inv_log2[i], built with the reference's own expressions
    (metrics.py:146,150: `1.0 / np.log2(i + 2)`, Python `sum` from 0) so the device results match bit for bit."""
    inv = np.array([1.0 / np.log2(i + 2) for i in range(k_max)], dtype=np.float64)
    idcg = np.zeros(k_max + 1, dtype=np.float64)
    acc = 0
    for i in range(k_max):
        acc = acc + inv[i]
        idcg[i + 1] = acc
    return inv, idcg


def ranking_metrics_device(pred: torch.Tensor, gt_lists: Sequence[np.ndarray], gt_counts: Optional[Sequence[int]],
                           k_values: Sequence[int], num_items: Optional[int] = None,
                           repeat: Optional[torch.Tensor] = None):
    """pred int64 [Q, K] on the device (rows of the catalogue, -1 padding), gt_lists[q] = that user's relevant rows.
    Returns (per_user fp64 [Q, 4*n_k + 2] numpy, device column sums fp64 tensor, distinct recommended rows or None)."""
    if not pred.is_cuda:
        raise RuntimeError("ranking_metrics_device needs the prediction matrix on a CUDA device (no CPU path)")
    assert pred.dtype == torch.int64 and pred.dim() == 2 and pred.stride(1) == 1
    dev = pred.device
    Q, K = pred.shape
    ks = sorted(int(k) for k in k_values)
    indptr, flat = _csr([np.sort(np.asarray(g, dtype=np.int64)) for g in gt_lists])
    inv, idcg = _dcg_tables(max(K, max(ks)))
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    d_indptr, d_flat, d_inv, d_idcg = t(indptr), t(flat), t(inv), t(idcg)
    d_cnt = t(np.asarray(gt_counts, dtype=np.int64)) if gt_counts is not None else None
    ncol = 4 * len(ks) + 2
    per_user = torch.empty((Q, ncol), dtype=torch.float64, device=dev)
    sums = torch.empty((ncol,), dtype=torch.float64, device=dev)
    cov_bits = cov_cnt = None
    if num_items:
        cov_bits = torch.empty(((num_items + 31) // 32,), dtype=torch.int32, device=dev)
        cov_cnt = torch.empty((1,), dtype=torch.int64, device=dev)
    import ctypes
    kv = (ctypes.c_int32 * len(ks))(*ks)
    N.check(N.lib().b200rec_eval_metrics(N.ptr(pred), Q, K, pred.stride(0), N.ptr(repeat),
                                         repeat.stride(0) if repeat is not None else 0, N.ptr(d_indptr), N.ptr(d_flat),
                                         N.ptr(d_cnt), kv, len(ks), N.ptr(d_inv), N.ptr(d_idcg), N.ptr(per_user),
                                         N.ptr(sums), N.ptr(cov_bits), int(num_items or 0), max(ks), N.ptr(cov_cnt),
                                         N.stream()), "eval_metrics")
    covered = int(cov_cnt.item()) if cov_cnt is not None else None
    return per_user.cpu().numpy(), sums, covered


class Evaluator:
    """Evaluates recommendation models (reference metrics.py:233-399): same constructor and methods."""

    def __init__(self, k_values: List[int] = [5, 10, 20, 50, 100], num_items: Optional[int] = None,
                 device: Optional[str] = None):
        self.k_values = sorted(k_values)
        self.num_items = num_items
        self.device = torch.device(device) if device else None

    # ------------------------------------------------------------------ metric table from per-user rows
    def _assemble(self, per_user: np.ndarray, covered: Optional[int]) -> EvaluationMetrics:
        res = EvaluationMetrics()
        have = per_user.shape[0] > 0
        col = lambda c: np.ascontiguousarray(per_user[:, c]).tolist()   # np.mean(list), as the reference
        for j, k in enumerate(self.k_values):
            r, p, n, h = col(4 * j), col(4 * j + 1), col(4 * j + 2), col(4 * j + 3)
            res.recall[k] = np.mean(r) if have else 0.0
            res.precision[k] = np.mean(p) if have else 0.0
            res.ndcg[k] = np.mean(n) if have else 0.0
            res.hit_rate[k] = np.mean(h) if have else 0.0
            res.per_user_recall[k] = r
            res.per_user_ndcg[k] = n
        res.mrr = np.mean(col(4 * len(self.k_values))) if have else 0.0
        res.map_score = np.mean(col(4 * len(self.k_values) + 1)) if have else 0.0
        if self.num_items and covered is not None:
            res.coverage = covered / self.num_items
        return res

    def _device(self) -> torch.device:
        if not torch.cuda.is_available():
            raise RuntimeError("b200rec Evaluator runs on a CUDA (sm_100a) device only; there is no CPU path")
        return self.device or torch.device("cuda", torch.cuda.current_device())

    # ------------------------------------------------------------------ Evaluator.evaluate (metrics.py:240-319)
    def evaluate(self, predictions: Dict[int, List[int]], ground_truth: Dict[int, Set[int]],
                 exclude_items: Optional[Dict[int, Set[int]]] = None) -> EvaluationMetrics:
        """predictions: user -> ranked item ids; ground_truth: user -> relevant ids.  Item ids are used as they come
        (any integers); users without ground truth are skipped, as in the reference."""
        users = [u for u in predictions if u in ground_truth and len(ground_truth[u]) > 0]
        if not users:
            return self._assemble(np.zeros((0, 4 * len(self.k_values) + 2)), 0 if self.num_items else None)
        lists = []
        for u in users:
            items = predictions[u]
            if exclude_items and u in exclude_items:
                ex = exclude_items[u]
                items = [i for i in items if i not in ex]
            lists.append(np.asarray(items, dtype=np.int64))
        K = max(max((len(x) for x in lists), default=1), 1)
        # ids are arbitrary integers here: compact them to rows of one vocabulary so -1 can mean "empty slot"
        vocab = np.unique(np.concatenate(lists + [np.fromiter(ground_truth[u], dtype=np.int64) for u in users]))
        pred = np.full((len(users), K), -1, dtype=np.int64)
        rep = np.zeros((len(users), K), dtype=np.uint8)
        for r, x in enumerate(lists):
            rows = np.searchsorted(vocab, x)
            pred[r, : len(x)] = rows
            _, first = np.unique(rows, return_index=True)
            mask = np.ones(len(x), dtype=np.uint8)
            mask[first] = 0
            rep[r, : len(x)] = mask
        gts = [np.searchsorted(vocab, np.fromiter(ground_truth[u], dtype=np.int64)) for u in users]
        dev = self._device()
        # coverage counts distinct recommended ITEM IDS among the first max(k) of every list (metrics.py:279)
        per_user, _, covered = ranking_metrics_device(torch.from_numpy(pred).to(dev), gts, None, self.k_values,
                                                      len(vocab) if self.num_items else None,
                                                      torch.from_numpy(rep).to(dev) if rep.any() else None)
        return self._assemble(per_user, covered)

    # ------------------------------------------------------------------ Evaluator.evaluate_model (metrics.py:321-399)
    @torch.no_grad()
    def recommend(self, model, test_users: Sequence[int], train_items: Dict[int, Iterable[int]],
                  user_features: np.ndarray, item_features: np.ndarray, top_k: int, batch_size: int = 256,
                  device: Optional[str] = None) -> Tuple[torch.Tensor, FlatIPDeviceIndex]:
        """Top-`top_k` catalogue ROWS for every test user with that user's training items excluded, as one device
        tensor [len(test_users), top_k] (the reference's np.dot + -inf mask + argsort, metrics.py:381-396 and
        scripts/evaluate_model.py:217-232).  `batch_size` users share one launch of the fused kernel."""
        dev = torch.device(device) if device else self._device()
        model.eval()
        items = torch.as_tensor(item_features, dtype=torch.float32).to(dev)
        item_emb = model.get_item_embeddings({"numerical": items, "categorical": {}})
        n_items = item_emb.shape[0]
        index = FlatIPDeviceIndex(item_emb.shape[1], storage="fp32", device=dev)
        index.add(item_emb)
        users = np.asarray(list(test_users), dtype=np.int64)
        out = torch.empty((len(users), top_k), dtype=torch.int64, device=dev)
        uf_all = torch.as_tensor(np.asarray(user_features), dtype=torch.float32)
        for s in range(0, len(users), batch_size):
            batch = users[s:s + batch_size]
            uf = uf_all[torch.from_numpy(batch)].to(dev)
            emb = model.get_user_embeddings({"numerical": uf, "categorical": {}})
            excl = []
            for u in batch.tolist():
                t = np.fromiter(train_items.get(u, ()), dtype=np.int64)
                excl.append(np.unique(t[(t >= 0) & (t < n_items)]))          # `if train_item < len(user_scores)` (:390)
            indptr, rows = _csr(excl)
            q_op = index.prepare_queries(emb, normalize=False)
            index.search_device(q_op, top_k, torch.from_numpy(indptr).to(dev),
                                torch.from_numpy(rows.astype(np.int32)).to(dev),
                                out=(torch.empty((len(batch), top_k), dtype=torch.float32, device=dev), out[s:s + len(batch)]))
        return out, index

    def evaluate_model(self, model, test_users: List[int], test_ground_truth: Dict[int, Set[int]],
                       train_items: Dict[int, Set[int]], user_features: np.ndarray, item_features: np.ndarray,
                       item_ids: List[int], batch_size: int = 256, device: str = "cuda") -> EvaluationMetrics:
        if not str(device).startswith("cuda"):
            raise RuntimeError("b200rec Evaluator.evaluate_model needs device='cuda' (no CPU path)")
        top_k = max(self.k_values)
        rows, _ = self.recommend(model, test_users, train_items, user_features, item_features, top_k, batch_size, device)
        # predictions[user] = [item_ids[row] ...] in the reference; ground truth is in item-id space: translate the
        # ground truth to rows once instead of translating Q x K predictions
        row_of = {int(i): r for r, i in enumerate(item_ids)}
        keep, gts, counts = [], [], []
        for q, u in enumerate(test_users):
            gt = test_ground_truth.get(u)
            if not gt:
                continue                                                     # `if user_id not in ground_truth` / empty
            keep.append(q)
            gts.append(np.asarray([row_of[int(i)] for i in gt if int(i) in row_of], dtype=np.int64))
            counts.append(len(gt))
        if not keep:
            return self._assemble(np.zeros((0, 4 * len(self.k_values) + 2)), 0 if self.num_items else None)
        sel = rows if len(keep) == len(test_users) else rows[torch.as_tensor(keep, device=rows.device)]
        per_user, _, covered = ranking_metrics_device(sel.contiguous(), gts, counts, self.k_values,
                                                      len(item_ids) if self.num_items else None)
        return self._assemble(per_user, covered)
