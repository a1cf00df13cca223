"""Generates tests/golden/metrics_small.npz by running the REFERENCE's own Evaluator.evaluate
(/root/reference/src/evaluation/metrics.py:240-319) on synthetic ranked lists: the recommendation lists of
evaltwin_small.npz (distinct ids, as a top-K search returns them) and a second case with repeated ids, ragged lists,
an exclusion dict, users without ground truth and ground-truth ids that are never recommended.
Run here (needs /root/reference):  python tests/golden/make_golden_metrics.py"""
import json, os, sys
import numpy as np
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, "/root/reference")
from src.evaluation.metrics import Evaluator   # noqa: E402

g = np.load(os.path.join(HERE, "evaltwin_small.npz"))
recs, test_users = g["recs"], g["test_users"].tolist()
n_items = int(g["items"].shape[0])
rng = np.random.default_rng(77)
cases = {}

# case A: distinct top-K lists, ground truth partly inside the lists
pred = {int(u): recs[j].tolist() for j, u in enumerate(test_users)}
gt = {}
for j, u in enumerate(test_users):
    n_in = int(rng.integers(0, 6))
    inside = rng.choice(recs[j], size=n_in, replace=False).tolist()
    outside = rng.integers(0, n_items, size=int(rng.integers(0, 5))).tolist()
    gt[int(u)] = set(int(x) for x in inside + outside)
gt[int(test_users[0])] = set()                       # empty ground truth: skipped
del gt[int(test_users[1])]                           # user without ground truth: skipped
ks = [1, 5, 10, 20]
ev = Evaluator(k_values=ks, num_items=n_items)
m = ev.evaluate(pred, gt)
cases["A"] = {"k_values": ks, "num_items": n_items, "pred": {str(k): v for k, v in pred.items()},
              "gt": {str(k): sorted(v) for k, v in gt.items()}, "exclude": None, "metrics": {k: float(v) for k, v in m.to_dict().items()},
              "per_user_recall": {str(k): [float(x) for x in v] for k, v in m.per_user_recall.items()},
              "per_user_ndcg": {str(k): [float(x) for x in v] for k, v in m.per_user_ndcg.items()}}

# case B: repeated ids, ragged lists, exclusion dict, ids far outside any catalogue
pred, gt, excl = {}, {}, {}
for u in range(40):
    L = int(rng.integers(1, 30))
    pred[u] = rng.integers(0, 25, size=L).tolist()   # small id range: many repeats
    gt[u] = set(int(x) for x in rng.integers(0, 25, size=int(rng.integers(1, 6))).tolist() + [10_000 + u])
    if u % 3 == 0:
        excl[u] = set(int(x) for x in rng.integers(0, 25, size=4).tolist())
ks = [3, 5, 50]
ev = Evaluator(k_values=ks, num_items=30)
m = ev.evaluate(pred, gt, excl)
cases["B"] = {"k_values": ks, "num_items": 30, "pred": {str(k): v for k, v in pred.items()},
              "gt": {str(k): sorted(v) for k, v in gt.items()}, "exclude": {str(k): sorted(v) for k, v in excl.items()},
              "metrics": {k: float(v) for k, v in m.to_dict().items()},
              "per_user_recall": {str(k): [float(x) for x in v] for k, v in m.per_user_recall.items()},
              "per_user_ndcg": {str(k): [float(x) for x in v] for k, v in m.per_user_ndcg.items()}}
np.savez_compressed(os.path.join(HERE, "metrics_small.npz"), cases_json=json.dumps(cases))
print({c: cases[c]["metrics"] for c in cases})
