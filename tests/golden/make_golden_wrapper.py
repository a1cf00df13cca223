"""Generates tests/golden/wrapper_small.npz by running the REFERENCE's own FaissIndex / RetrievalEngine wrapper code
(/root/reference/src/serving/retrieval.py:49-329, 505-692) with a numpy stand-in for the `faiss` module, which cannot be
installed offline: the stand-in supplies only the external arithmetic (IndexFlatIP add / search, normalize_L2) through
oracle/flat_ip.py, so everything the wrapper itself does — casts, 1-D promotion, cosine normalisation, the Flat fallback
below 1024 items, id maps, the filter_ids post-pass with k_search = min(2k, N), incremental add, engine.retrieve — is the
reference's code.  Run here (needs /root/reference):  python tests/golden/make_golden_wrapper.py"""
import json, os, sys, types
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import flat_ip as O   # noqa: E402

faiss = types.ModuleType("faiss")


class _Flat(O.IndexFlatIP):
    is_trained = True

    def train(self, x):
        pass


def _normalize_L2(x):
    x[:] = O.normalize_L2(x)


faiss.IndexFlatIP = _Flat
faiss.IndexFlatL2 = _Flat
faiss.normalize_L2 = _normalize_L2
faiss.METRIC_INNER_PRODUCT, faiss.METRIC_L2 = 0, 1
faiss.index_factory = lambda d, factory, metric=0: _Flat(d)
faiss.get_num_gpus = lambda: 0
sys.modules["faiss"] = faiss
annoy = types.ModuleType("annoy")
annoy.AnnoyIndex = object
sys.modules["annoy"] = annoy
sys.path.insert(0, "/root/reference")
import importlib.util   # noqa: E402
spec = importlib.util.spec_from_file_location("ref_retrieval", "/root/reference/src/serving/retrieval.py")
ref = importlib.util.module_from_spec(spec)
spec.loader.exec_module(ref)

rng = np.random.default_rng(2024)
N, D, Q = 700, 16, 9
emb = rng.standard_normal((N, D)).astype(np.float32)
emb[50] = emb[10]                       # duplicate rows: ties
extra = rng.standard_normal((40, D)).astype(np.float32)
qry = rng.standard_normal((Q, D)).astype(np.float32)
ids = [f"item_{i}" for i in range(N)]
extra_ids = [f"new_{i}" for i in range(40)]
allowed = [f"item_{i}" for i in range(0, N, 3)]
out = {"emb": emb, "extra": extra, "qry": qry, "N": N, "D": D}
cases = {}
for metric in ("cosine", "ip"):
    ix = ref.FaissIndex({"dimension": D, "index_factory": "IVF1024,Flat", "metric": metric})
    ix.build(emb.copy(), ids)
    cases[f"{metric}_k10"] = ix.search(qry.copy(), k=10)
    cases[f"{metric}_1d_k5"] = ix.search(qry[0].copy(), k=5)
    cases[f"{metric}_filter_k7"] = ix.search(qry.copy(), k=7, filter_ids=allowed)
    ix.add(extra.copy(), extra_ids)
    cases[f"{metric}_after_add_k10"] = ix.search(qry.copy(), k=10)
    cases[f"{metric}_size"] = ix.current_size
eng = ref.RetrievalEngine({"index_type": "faiss", "embedding_dim": D, "faiss": {"index_factory": "Flat", "metric": "cosine"}})
eng.build_index(emb.copy(), ids)
r_ids, r_scores, r_metrics = eng.retrieve(qry[:2].copy(), k=6)
cases["engine_retrieve_k6"] = (r_ids, r_scores)
cases["engine_metrics_keys"] = sorted(r_metrics.keys())
out["cases_json"] = json.dumps(cases)
out["allowed_json"] = json.dumps(allowed)
np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "wrapper_small.npz"), **out)
print({k: (type(v).__name__, len(v) if hasattr(v, "__len__") else v) for k, v in cases.items()})
