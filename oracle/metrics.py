"""CPU ORACLE (test infrastructure, NOT product code) — ranking metrics of the offline evaluation.

Restates ``Evaluator.evaluate`` (reference ``src/evaluation/metrics.py:240-319``) and its per-user helpers
(``recall_at_k`` :74-97, ``precision_at_k`` :100-120, ``ndcg_at_k`` :123-158, ``hit_rate_at_k`` :161-179,
``reciprocal_rank`` :182-200, ``average_precision`` :203-231) in plain Python / numpy fp64.  PINNED: the imported
reference itself produced ``tests/golden/metrics_small.npz`` (``tests/golden/make_golden_metrics.py``) and
``tests/test_oracle.py`` requires this restatement to reproduce every number of it exactly.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline leg may import this module.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Set

import numpy as np


def per_user_row(pred: Sequence[int], gt: Set[int], k_values: Sequence[int]) -> List[float]:
    """[recall, precision, ndcg, hit] for every k, then reciprocal rank and average precision (one user)."""
    out: List[float] = []
    for k in k_values:
        top = set(pred[:k])
        hits = len(top & gt)
        recall = hits / len(gt) if len(gt) else 0.0                      # :74-97
        precision = hits / k                                            # :100-120
        dcg = 0.0
        for i, item in enumerate(pred[:k]):                             # :140-146
            if item in gt:
                dcg += 1.0 / np.log2(i + 2)
        ideal = min(len(gt), k)
        idcg = sum(1.0 / np.log2(i + 2) for i in range(ideal))          # :149-150
        ndcg = 0.0 if (len(gt) == 0 or idcg == 0) else dcg / idcg
        out += [recall, precision, ndcg, 1.0 if hits > 0 else 0.0]      # :161-179
    rr = 0.0
    for i, item in enumerate(pred):                                     # :182-200
        if item in gt:
            rr = 1.0 / (i + 1)
            break
    ap, nh = 0.0, 0
    for i, item in enumerate(pred):                                     # :203-231
        if item in gt:
            nh += 1
            ap += nh / (i + 1)
    out += [rr, ap / len(gt) if len(gt) else 0.0]
    return out


def evaluate(predictions: Dict[int, List[int]], ground_truth: Dict[int, Set[int]], k_values: Sequence[int],
             num_items: Optional[int] = None, exclude_items: Optional[Dict[int, Set[int]]] = None):
    """-> (flat metric dict as EvaluationMetrics.to_dict(), per-user matrix [users, 4*n_k+2]) — metrics.py:240-319."""
    ks = sorted(k_values)
    rows, seen = [], set()
    for u, pred in predictions.items():
        if u not in ground_truth:
            continue
        gt = ground_truth[u]
        if exclude_items and u in exclude_items:
            pred = [i for i in pred if i not in exclude_items[u]]
        if len(gt) == 0:
            continue
        seen.update(pred[:max(ks)])
        rows.append(per_user_row(pred, gt, ks))
    mat = np.asarray(rows, dtype=np.float64).reshape(len(rows), 4 * len(ks) + 2)
    mean = lambda c: float(np.mean(mat[:, c].tolist())) if len(rows) else 0.0
    out = {}
    for j, k in enumerate(ks):
        out[f"recall@{k}"], out[f"precision@{k}"] = mean(4 * j), mean(4 * j + 1)
        out[f"ndcg@{k}"], out[f"hit_rate@{k}"] = mean(4 * j + 2), mean(4 * j + 3)
    out["mrr"], out["map"] = mean(4 * len(ks)), mean(4 * len(ks) + 1)
    out["coverage"] = len(seen) / num_items if num_items else 0.0
    return out, mat
