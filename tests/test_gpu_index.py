"""GPU parity of the DROP-IN index / engine / evaluation classes against the CPU oracle and the committed goldens the
reference itself produced: B200FlatIndex / RetrievalEngine vs
wrapper_small.npz, exact_topk_eval and Evaluator vs evaltwin_small.npz / metrics_small.npz, the default fp32 storage
(3-term operand, KB = 6 at D = 128), the IxFI save/load round trip."""
import json
import os
import struct

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _same(got, want, atol=2e-6, rtol=1e-6):
    """ids identical; scores equal up to fp32 summation order (the goldens are numpy sgemm, ours fp32 FMA chains)."""
    assert got[0] == want[0], "item ids differ"
    assert len(got[1]) == len(want[1])
    for a, b in zip(got[1], want[1]):
        assert len(a) == len(b) and np.allclose(a, b, rtol=rtol, atol=atol)


@pytest.mark.parametrize("metric", ["cosine", "ip"])
def test_b200_flat_index_matches_the_reference_wrapper(metric):
    """Every case of wrapper_small.npz (outputs of the REFERENCE's FaissIndex code, tests/golden/make_golden_wrapper.py):
    batch search, 1-D query, the filter_ids post-pass with k_search = min(2k, N), search after add, current_size."""
    from b200rec.retrieval import B200FlatIndex
    g = np.load(os.path.join(GOLDEN, "wrapper_small.npz"))
    cases, allowed = json.loads(str(g["cases_json"])), json.loads(str(g["allowed_json"]))
    emb, extra, qry, N, D = g["emb"], g["extra"], g["qry"], int(g["N"]), int(g["D"])
    ids = [f"item_{i}" for i in range(N)]
    ix = B200FlatIndex({"dimension": D, "index_factory": "IVF1024,Flat", "metric": metric})
    with pytest.raises(ValueError, match="Index not built yet"):
        ix.search(qry, k=3)
    with pytest.raises(ValueError, match="No index to save"):
        ix.save("/tmp/never")
    ix.build(emb.copy(), ids)
    assert ix.current_size == N and ix.index.ntotal == N and ix.id_map[3] == "item_3"
    _same(ix.search(qry.copy(), k=10), cases[f"{metric}_k10"])
    _same(ix.search(qry[0].copy(), k=5), cases[f"{metric}_1d_k5"])
    _same(ix.search(qry.copy(), k=7, filter_ids=allowed), cases[f"{metric}_filter_k7"])
    assert ix.search(qry[:2].copy(), k=4, filter_ids=[]) == ([[], []], [[], []])      # `item_id in []` is never true
    ix.add(extra.copy(), [f"new_{i}" for i in range(len(extra))])
    _same(ix.search(qry.copy(), k=10), cases[f"{metric}_after_add_k10"])
    assert ix.current_size == cases[f"{metric}_size"]
    # device-resident queries take the same path without the host round trip
    _same(ix.search(torch.from_numpy(qry).cuda(), k=10), cases[f"{metric}_after_add_k10"])


def test_retrieval_engine_matches_the_reference_engine():
    from b200rec.retrieval import RetrievalEngine
    g = np.load(os.path.join(GOLDEN, "wrapper_small.npz"))
    cases = json.loads(str(g["cases_json"]))
    emb, qry, N, D = g["emb"], g["qry"], int(g["N"]), int(g["D"])
    eng = RetrievalEngine({"index_type": "faiss", "embedding_dim": D, "faiss": {"index_factory": "Flat", "metric": "cosine"}})
    eng.build_index(emb.copy(), [f"item_{i}" for i in range(N)])
    ids, scores, metrics = eng.retrieve(qry[:2].copy(), k=6)
    _same((ids, scores), cases["engine_retrieve_k6"])
    assert sorted(metrics.keys()) == cases["engine_metrics_keys"]
    assert metrics["num_results"] == 12 and metrics["cache_hit"] is False
    assert abs(metrics["avg_score"] - np.mean([s for row in scores for s in row])) < 1e-6
    m = eng.get_metrics()
    assert m["total_queries"] == 1 and m["index_size"] == N and abs(m["avg_latency_ms"] - metrics["latency_ms"]) < 1e-6
    eng.update_index(g["extra"].copy(), [f"new_{i}" for i in range(len(g["extra"]))])
    _same(eng.retrieve(qry.copy(), k=10)[:2], cases["cosine_after_add_k10"])
    for bad, msg in (("annoy", "not part of the B200 hot path"), ("nope", "Unknown index type")):
        with pytest.raises(ValueError, match=msg):
            RetrievalEngine({"index_type": bad})


def test_save_load_round_trip_and_ixfi_header(tmp_path):
    """`.faiss` = faiss' IndexFlatIP file ("IxFI"), `.pkl` = id maps (reference retrieval.py:248-299)."""
    from b200rec.retrieval import B200FlatIndex
    rng = np.random.default_rng(3)
    N, D = 900, 24
    emb = rng.standard_normal((N, D)).astype(np.float32)
    qry = rng.standard_normal((11, D)).astype(np.float32)
    ids = [f"m{i}" for i in range(N)]
    a = B200FlatIndex({"dimension": D, "metric": "cosine"})
    a.build(emb, ids)
    before = a.search(qry, k=9)
    a.save(str(tmp_path / "idx" / "flat"))
    raw = open(tmp_path / "idx" / "flat.faiss", "rb").read()
    assert raw[:4] == b"IxFI"
    d, ntotal, d1, d2, trained, metric = struct.unpack("<iqqqBi", raw[4:4 + 33])
    assert (d, ntotal, d1, d2, trained, metric) == (D, N, 1 << 20, 1 << 20, 1, 0)
    assert struct.unpack("<Q", raw[37:45])[0] == N * D and len(raw) == 45 + 4 * N * D
    stored = np.frombuffer(raw[45:], dtype=np.float32).reshape(N, D)
    assert np.allclose(np.linalg.norm(stored, axis=1), 1.0, atol=1e-6)               # rows are saved normalised
    b = B200FlatIndex({"dimension": 1, "metric": "cosine"})
    b.load(str(tmp_path / "idx" / "flat"))
    assert b.current_size == N and b.dimension == D and b.id_map == a.id_map
    after = b.search(qry, k=9)
    assert after[0] == before[0]
    for x, y in zip(after[1], before[1]):
        assert np.allclose(x, y, rtol=0, atol=1e-6)
    b.add(rng.standard_normal((5, D)).astype(np.float32), [f"n{i}" for i in range(5)])
    assert b.current_size == N + 5 and len(b.search(qry, k=3)[0][0]) == 3


@pytest.mark.parametrize("D,N,Q,k", [(128, 20000, 70, 100), (64, 5000, 300, 10), (100, 3000, 5, 50)])
def test_fp32_storage_matches_the_fp32_oracle(D, N, Q, k):
    """The DEFAULT storage of the drop-in class: split-bf16 x3 catalogue rows (ld = 3 * pad64(D): KB = 6 k-blocks at
    D = 128) select k + margin candidates, which are re-scored exactly in fp32 (csrc/rescore.cu), on inputs that are NOT
    bf16-representable.  Scores within 1e-6 of the fp32 oracle; ids equal except inside groups of scores tied within
    1e-6 (fp32 summation order; the north_star tolerance is 1e-5)."""
    from b200rec.retrieval import FlatIPDeviceIndex
    from oracle.flat_ip import IndexFlatIP, normalize_L2
    rng = np.random.default_rng(D + N)
    cat = normalize_L2(rng.standard_normal((N, D)).astype(np.float32))
    qry = normalize_L2(rng.standard_normal((Q, D)).astype(np.float32))
    ix = FlatIPDeviceIndex(D, storage="fp32")
    assert ix.ld == 3 * ((D + 63) // 64 * 64)
    ix.add(cat[: N // 2])
    ix.add(cat[N // 2:])                                                             # incremental add
    Dg, Ig = ix.search(qry, k)
    ref = IndexFlatIP(D)
    ref.add(cat)
    rD, rI = ref.search(qry, k)
    assert Dg.dtype == np.float32 and Ig.dtype == np.int64 and Dg.shape == (Q, k)
    assert np.abs(Dg - rD).max() <= 1e-6
    assert (np.diff(Dg, axis=1) <= 0).all()
    bad = np.argwhere(Ig != rI)
    for q, j in bad:
        exact = float(cat[Ig[q, j]].astype(np.float64) @ qry[q].astype(np.float64))
        assert abs(exact - rD[q, j]) <= 1e-6, f"query {q} rank {j}: id {Ig[q, j]} ({exact}) vs {rI[q, j]} ({rD[q, j]})"
    assert len(bad) <= 0.002 * Ig.size
    # without the fp32 queries attached the same index answers from the 3-product tensor-core scores alone (1e-5 grade)
    q_op = ix.prepare_queries(qry)
    del q_op._b200_f32
    Dc, Ic = ix.search_device(q_op, k)
    assert np.abs(Dc.cpu().numpy() - rD).max() <= 1e-5 and (Ic.cpu().numpy() == rI).mean() >= 0.98
    assert np.allclose(ix.reconstruct_n(0, 7), cat[:7], atol=0)                       # fp32 rows are kept exactly


def test_exact_topk_eval_matches_the_reference_recommendation_twin():
    """exact_topk_eval (fused kernel + exclusion CSR) against evaltwin_small.npz: the ids the REFERENCE's
    generate_recommendations returned (np.dot + -inf mask + argsort, scripts/evaluate_model.py:217-232)."""
    from b200rec.retrieval import FlatIPDeviceIndex, exact_topk_eval
    g = np.load(os.path.join(GOLDEN, "evaltwin_small.npz"))
    items, users, test_users, k = g["items"], g["users"], g["test_users"], int(g["k"])
    train = {int(u): g["train_items"][g["train_indptr"][j]:g["train_indptr"][j + 1]].tolist()
             for j, u in enumerate(g["train_users"].tolist())}
    ix = FlatIPDeviceIndex(items.shape[1], storage="fp32")
    ix.add(items)
    got = exact_topk_eval(torch.from_numpy(users[test_users]).cuda(), ix,
                          {u: [t for t in v if t < len(items)] for u, v in train.items()}, test_users.tolist(), k)
    assert np.array_equal(got, g["recs"])


def _metrics_cases():
    g = np.load(os.path.join(GOLDEN, "metrics_small.npz"))
    cases = json.loads(str(g["cases_json"]))
    for c in cases.values():
        c["pred"] = {int(k): v for k, v in c["pred"].items()}
        c["gt"] = {int(k): set(v) for k, v in c["gt"].items()}
        c["exclude"] = {int(k): set(v) for k, v in c["exclude"].items()} if c["exclude"] else None
    return cases


def test_device_metrics_equal_the_reference_evaluator():
    """Evaluator.evaluate through b200rec_eval_metrics: every aggregate and per-user number equals what the imported
    reference computed (metrics_small.npz), bit for bit (fp64, same accumulation order)."""
    from b200rec.evaluation import Evaluator
    for name, c in _metrics_cases().items():
        m = Evaluator(k_values=c["k_values"], num_items=c["num_items"]).evaluate(c["pred"], c["gt"], c["exclude"])
        got = m.to_dict()
        for key, want in c["metrics"].items():
            assert got[key] == want, (name, key, got[key], want)
        for k in c["k_values"]:
            assert m.per_user_recall[k] == c["per_user_recall"][str(k)]
            assert m.per_user_ndcg[k] == c["per_user_ndcg"][str(k)]


def test_evaluate_model_on_device_equals_oracle_pipeline():
    """Evaluator.evaluate_model (towers -> masked fused top-K -> device metrics) against the oracle pipeline on the same
    weights: numpy towers -> eval twin (np.dot, -inf mask, argsort) -> oracle metrics (reference metrics.py:321-399)."""
    from b200rec.evaluation import Evaluator
    from b200rec.training_utils import create_two_tower_model_for_training
    from oracle import metrics as M
    from oracle.flat_ip import eval_twin_topk
    from oracle.two_tower import TowerOracle
    torch.manual_seed(11)
    rng = np.random.default_rng(11)
    NU, NI, FU, FI = 400, 900, 3, 20
    model = create_two_tower_model_for_training(FU, FI, {"embedding_dim": 32, "hidden_layers": [64, 32]}).cuda()
    model.eval()
    for bn in [m for m in model.modules() if isinstance(m, torch.nn.BatchNorm1d)]:    # non-trivial eval statistics
        bn.running_mean.normal_(0, 0.3)
        bn.running_var.uniform_(0.5, 1.5)
    uf = rng.standard_normal((NU, FU)).astype(np.float32)
    itf = rng.standard_normal((NI, FI)).astype(np.float32)
    test_users = rng.permutation(NU)[:300].tolist()
    train = {u: set(rng.integers(0, NI + 40, size=int(rng.integers(0, 30))).tolist()) for u in range(NU) if u % 5}
    gt = {u: set(rng.integers(0, NI, size=int(rng.integers(0, 8))).tolist()) for u in test_users if u % 7}
    item_ids = list(range(NI))
    ks = [5, 10, 20, 50, 100]
    got = Evaluator(k_values=ks, num_items=NI).evaluate_model(model, test_users, gt, train, uf, itf, item_ids,
                                                             batch_size=128, device="cuda")
    sd = lambda t: {k: v.detach().cpu().numpy() for k, v in t.state_dict().items()}
    ue = TowerOracle(sd(model.user_tower), 2, dtype=np.float32).forward(uf[test_users], training=False)
    ie = TowerOracle(sd(model.item_tower), 2, dtype=np.float32).forward(itf, training=False)
    recs = eval_twin_topk(ue, ie, {u: sorted(v) for u, v in train.items()}, test_users, 100)
    want, _ = M.evaluate({u: [item_ids[i] for i in recs[u]] for u in test_users}, gt, ks, NI)
    for key, v in want.items():
        assert abs(got.to_dict()[key] - v) <= 2e-3 + 2e-2 * abs(v), (key, got.to_dict()[key], v)
    # with identical recommendation lists the tables are identical: feed the device lists to the oracle
    rows, _ = Evaluator(k_values=ks, num_items=NI).recommend(model, test_users, train, uf, itf, 100, 128, "cuda")
    rows = rows.cpu().numpy()
    exact, mat = M.evaluate({u: rows[j].tolist() for j, u in enumerate(test_users)}, gt, ks, NI)
    for key, v in exact.items():
        assert got.to_dict()[key] == v, (key, got.to_dict()[key], v)
    # and the lists themselves agree with the twin except where fp32 rounding of near-equal scores swaps neighbours
    agree = np.mean([rows[j].tolist() == recs[u] for j, u in enumerate(test_users)])
    assert agree >= 0.9
    for j, u in enumerate(test_users):
        assert not (set(rows[j].tolist()) & train.get(u, set())), "a training item was recommended"
