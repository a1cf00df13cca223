"""Generates tests/golden/evaltwin_small.npz by running the REFERENCE's generate_recommendations
(/root/reference/scripts/evaluate_model.py:160-234: np.dot + train-item -inf mask + argsort[::-1][:k]) on synthetic,
tie-free embeddings through an identity 'model'.  Run here (needs /root/reference)."""
import importlib.util, os, sys
import numpy as np
import torch
sys.path.insert(0, "/root/reference")
spec = importlib.util.spec_from_file_location("ref_eval", "/root/reference/scripts/evaluate_model.py")
ref = importlib.util.module_from_spec(spec)
spec.loader.exec_module(ref)


class Identity:   # the towers are not under test here: embeddings in, embeddings out
    def get_item_embeddings(self, f):
        return f["numerical"]

    def get_user_embeddings(self, f):
        return f["numerical"]


rng = np.random.default_rng(7)
n_items, n_users, d, k = 1200, 200, 64, 100
items = rng.standard_normal((n_items, d)).astype(np.float32)
users = rng.standard_normal((n_users, d)).astype(np.float32)
items /= np.linalg.norm(items, axis=1, keepdims=True)
users /= np.linalg.norm(users, axis=1, keepdims=True)
test_users = rng.permutation(n_users)[:170].tolist()
train = {int(u): rng.choice(n_items + 50, size=int(rng.integers(0, 300)), replace=False).tolist() for u in range(0, n_users, 2)}
recs = ref.generate_recommendations(Identity(), test_users, train, users, items, top_k=k, batch_size=64, device="cpu")
out = {"items": items, "users": users, "test_users": np.asarray(test_users), "k": k,
       "train_users": np.asarray(sorted(train)), "train_indptr": np.cumsum([0] + [len(train[u]) for u in sorted(train)]),
       "train_items": np.concatenate([np.asarray(train[u], dtype=np.int64) for u in sorted(train)]),
       "recs": np.asarray([recs[u] for u in test_users], dtype=np.int64)}
np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "evaltwin_small.npz"), **out)
print({a: getattr(b, "shape", b) for a, b in out.items()})
