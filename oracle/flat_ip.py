"""CPU ORACLE (test infrastructure, NOT product code) — exact inner-product top-K.

Restates, in numpy, what the reference delegates to ``faiss.IndexFlatIP`` behind
``FaissIndex`` (reference ``src/serving/retrieval.py:70-136`` build, ``:141-197`` search,
``:199-226`` add) and the in-repo exact twin ``scripts/evaluate_model.py:217-232`` /
``src/evaluation/metrics.py:381-396`` (``np.dot`` + ``-inf`` mask + ``argsort[::-1][:k]``).

PARITY UNPINNED at the Faiss boundary: ``faiss-cpu==1.7.4`` (reference ``requirements.txt:13``)
is a third-party wheel that is neither vendored under /root/reference nor installable offline,
and no reference test touches ``src/serving/retrieval.py``.  What is restated here is the
published IndexFlatIP algorithm:
  * scores are fp32 inner products (sgemm in query/database blocks),
  * the result is the k largest, sorted by descending score,
  * labels are int64 insertion-order row numbers,
  * a candidate replaces the current k-th only if STRICTLY greater, so at the k boundary the
    lower row id survives; we make the whole order deterministic: (score desc, row id asc),
  * unfilled slots (k > ntotal) carry label -1 and score -FLT_MAX (``CMin<float>::neutral()``),
  * ``normalize_L2``: x *= 1/sqrt(sum x^2) in fp32, zero rows untouched.
It is pinned instead against the reference's own exact twin (np.dot + argsort) in
``tests/test_oracle.py`` on tie-free data.

Three restatements of the same search live here: ``IndexFlatIP.search`` (the readable CHECKER the parity tests
compare the CUDA path with), ``search_reservoir`` (the CPU BASELINE bench.py times: MKL sgemm tiles + faiss'
reservoir result handler in C, ``oracle/csrc/flat_select.c``) and ``search_blocked`` (the same with torch.topk per
block, the fallback when no C compiler is around); the tests pin the latter two to the checker.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference``
leg may import this module.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

FLT_MAX = np.finfo(np.float32).max


def normalize_L2(x: np.ndarray) -> np.ndarray:
    """faiss.normalize_L2 (in place, fp32).  reference call sites retrieval.py:86,167,214."""
    assert x.dtype == np.float32 and x.ndim == 2
    nrm2 = np.einsum("ij,ij->i", x, x, dtype=np.float32)
    inv = np.ones_like(nrm2)
    nz = nrm2 > 0
    inv[nz] = (1.0 / np.sqrt(nrm2[nz])).astype(np.float32)
    x *= inv[:, None]
    return x


def _select_sorted(scores: np.ndarray, ids: np.ndarray, k: int) -> Tuple[np.ndarray, np.ndarray]:
    """Top-k of one block under the total order (score desc, id asc).  scores [nq, m]."""
    nq, m = scores.shape
    kk = min(k, m)
    if kk < m:
        # argpartition is not tie-stable; widen to every element >= the k-th score, then sort.
        part = np.argpartition(-scores, kk - 1, axis=1)[:, :kk]
        kth = np.take_along_axis(scores, part, axis=1).min(axis=1)
    out_s = np.full((nq, k), -FLT_MAX, dtype=np.float32)
    out_i = np.full((nq, k), -1, dtype=np.int64)
    for q in range(nq):
        if kk < m:
            cand = np.nonzero(scores[q] >= kth[q])[0]
        else:
            cand = np.arange(m)
        order = np.lexsort((ids[cand], -scores[q, cand].astype(np.float64)))[:kk]
        sel = cand[order]
        out_s[q, :kk] = scores[q, sel]
        out_i[q, :kk] = ids[sel]
    return out_s, out_i


def merge_topk(parts_s: Sequence[np.ndarray], parts_i: Sequence[np.ndarray], k: int):
    """k-way merge of per-shard (scores [nq,k'], ids [nq,k']) lists; id -1 entries are padding."""
    s = np.concatenate(parts_s, axis=1)
    i = np.concatenate(parts_i, axis=1)
    nq = s.shape[0]
    out_s = np.full((nq, k), -FLT_MAX, dtype=np.float32)
    out_i = np.full((nq, k), -1, dtype=np.int64)
    for q in range(nq):
        valid = i[q] >= 0
        sq, iq = s[q, valid], i[q, valid]
        order = np.lexsort((iq, -sq.astype(np.float64)))[:k]
        out_s[q, : len(order)] = sq[order]
        out_i[q, : len(order)] = iq[order]
    return out_s, out_i


class IndexFlatIP:
    """faiss.IndexFlatIP restated (reference retrieval.py:98 ctor, :122/:217 add, :171 search)."""

    def __init__(self, d: int, db_block: int = 65536, q_block: int = 1024):
        self.d = int(d)
        self.xb = np.zeros((0, self.d), dtype=np.float32)
        self.db_block = db_block
        self.q_block = q_block
        self.is_trained = True

    @property
    def ntotal(self) -> int:
        return self.xb.shape[0]

    def add(self, x: np.ndarray) -> None:
        x = np.ascontiguousarray(x, dtype=np.float32)
        assert x.ndim == 2 and x.shape[1] == self.d
        self.xb = x if self.ntotal == 0 else np.concatenate([self.xb, x], axis=0)

    def search(self, q: np.ndarray, k: int, exclude: Optional[List[np.ndarray]] = None):
        """Returns (D fp32 [nq,k] descending, I int64 [nq,k]).

        ``exclude`` (optional, one int array per query) restates the eval twin's ``-inf`` masking of
        a user's train items (scripts/evaluate_model.py:225-228): excluded rows can never be returned.
        """
        q = np.ascontiguousarray(q, dtype=np.float32)
        assert q.ndim == 2 and q.shape[1] == self.d
        nq = q.shape[0]
        D = np.full((nq, k), -FLT_MAX, dtype=np.float32)
        I = np.full((nq, k), -1, dtype=np.int64)
        for q0 in range(0, nq, self.q_block):
            q1 = min(nq, q0 + self.q_block)
            run_s = [np.full((q1 - q0, k), -FLT_MAX, dtype=np.float32)]
            run_i = [np.full((q1 - q0, k), -1, dtype=np.int64)]
            for b0 in range(0, self.ntotal, self.db_block):
                b1 = min(self.ntotal, b0 + self.db_block)
                s = q[q0:q1] @ self.xb[b0:b1].T  # fp32 sgemm block
                ids = np.arange(b0, b1, dtype=np.int64)
                if exclude is not None:
                    for r in range(q0, q1):
                        ex = np.asarray(exclude[r], dtype=np.int64)
                        ex = ex[(ex >= b0) & (ex < b1)]
                        s[r - q0, ex - b0] = -np.inf
                bs, bi = _select_sorted(s, ids, k)
                if exclude is not None:
                    bi[np.isneginf(bs)] = -1
                    bs[bi < 0] = -FLT_MAX
                run_s.append(bs)
                run_i.append(bi)
            D[q0:q1], I[q0:q1] = merge_topk(run_s, run_i, k)
        return D, I


def search_blocked(xb, q, k: int, threads: int = 0, db_block: int = 1 << 17, q_block: int = 1024):
    """The same exact search laid out the way faiss-cpu runs IndexFlatIP for nq >= 20 (exhaustive_inner_product_blas:
    sgemm over (query block x database block) + a k-selection per block folded into the running result), on all host
    cores: torch-CPU (MKL) sgemm, torch.topk per block, running merge by a topk over [k running | k new].  This is the
    CPU BASELINE bench.py times (`cpu_baseline`, `--impl reference`); IndexFlatIP.search above is the readable checker
    with the deterministic tie order (this one leaves the order among exactly equal scores to torch.topk).
    xb / q: float32 numpy or torch CPU tensors.  Returns (D, I) numpy."""
    import torch
    if threads > 0:
        torch.set_num_threads(threads)
    xb = torch.as_tensor(xb)
    q = torch.as_tensor(q)
    nq, n = q.shape[0], xb.shape[0]
    D = torch.full((nq, k), -float(FLT_MAX), dtype=torch.float32)
    I = torch.full((nq, k), -1, dtype=torch.int64)
    for q0 in range(0, nq, q_block):
        qs = q[q0:q0 + q_block]
        run_s, run_i = D[q0:q0 + q_block], I[q0:q0 + q_block]
        for b0 in range(0, n, db_block):
            s = qs @ xb[b0:b0 + db_block].T
            kk = min(k, s.shape[1])
            bs, bi = torch.topk(s, kk, dim=1)
            cat_s = torch.cat([run_s, bs], dim=1)
            cat_i = torch.cat([run_i, bi + b0], dim=1)
            top_s, pos = torch.topk(cat_s, k, dim=1)
            run_s, run_i = top_s, torch.gather(cat_i, 1, pos)
        D[q0:q0 + q_block], I[q0:q0 + q_block] = run_s, run_i
    return D.numpy(), I.numpy()


_SELECT_LIB = None


def _select_lib():
    """oracle/_build/libflatselect.so (oracle/csrc/flat_select.c: faiss-cpu's reservoir result handler restated in C),
    built on first use when missing or stale."""
    global _SELECT_LIB
    if _SELECT_LIB is None:
        import ctypes
        from . import build_c
        lib = ctypes.CDLL(build_c.build())
        p, i64, i32 = ctypes.c_void_p, ctypes.c_int64, ctypes.c_int32
        lib.flat_select_add_block.argtypes = [p, i64, i64, i64, i64, i32, p, p, p, p, i32]
        lib.flat_select_add_block.restype = None
        lib.flat_select_finish.argtypes = [i64, i32, p, p, p, p, p, i32]
        lib.flat_select_finish.restype = None
        _SELECT_LIB = lib
    return _SELECT_LIB


def search_reservoir(xb, q, k: int, threads: int = 0, db_block: int = 4096, q_block: int = 1024):
    """Exact flat IP search the way faiss-cpu 1.7.4 runs IndexFlatIP for nq >= 20 and k >= 100
    (`exhaustive_inner_product_blas` + `ReservoirResultHandler`): sgemm over a (query block x database block) tile that
    stays cache resident (torch-CPU / MKL, all host threads), the tile handed to a per-query reservoir of capacity 2k
    that only admits scores above the query's running threshold (oracle/csrc/flat_select.c, OpenMP over the queries).
    About one compare per score next to the 2·D FLOP of the sgemm, so the search runs at the sgemm rate — this is the
    CPU BASELINE bench.py times (`cpu_baseline`, `--impl reference`).  Order (score desc, row asc) like the checker
    IndexFlatIP.search above.  xb / q: float32 numpy or torch CPU tensors.  Returns (D, I) numpy."""
    import torch
    lib = _select_lib()
    if threads > 0:
        torch.set_num_threads(threads)
    nt = threads if threads > 0 else torch.get_num_threads()
    xb = torch.as_tensor(xb)
    q = torch.as_tensor(q)
    assert xb.dtype == torch.float32 and q.dtype == torch.float32 and xb.is_contiguous() and q.is_contiguous()
    nq, n = q.shape[0], xb.shape[0]
    D = np.empty((nq, k), dtype=np.float32)
    I = np.empty((nq, k), dtype=np.int64)
    tile = torch.empty((min(q_block, max(nq, 1)), db_block), dtype=torch.float32)
    for q0 in range(0, nq, q_block):
        qs = q[q0:q0 + q_block]
        m = qs.shape[0]
        res_s = np.empty((m, 2 * k), dtype=np.float32)
        res_i = np.empty((m, 2 * k), dtype=np.int64)
        res_n = np.zeros((m,), dtype=np.int32)
        thr = np.full((m,), -FLT_MAX, dtype=np.float32)
        for b0 in range(0, n, db_block):
            blk = xb[b0:b0 + db_block]
            w = blk.shape[0]
            s = tile[:m, :w]
            torch.matmul(qs, blk.T, out=s)
            lib.flat_select_add_block(s.data_ptr(), s.stride(0), m, w, b0, k, res_s.ctypes.data, res_i.ctypes.data,
                                      res_n.ctypes.data, thr.ctypes.data, nt)
        lib.flat_select_finish(m, k, res_s.ctypes.data, res_i.ctypes.data, res_n.ctypes.data,
                               D[q0:q0 + m].ctypes.data, I[q0:q0 + m].ctypes.data, nt)
    return D, I


class FaissIndexOracle:
    """FaissIndex Flat/cosine-or-IP wrapper semantics (reference retrieval.py:49-226)."""

    def __init__(self, config: Optional[Dict] = None):
        self.config = config or {}
        self.dimension = self.config.get("dimension", 128)
        self.metric = self.config.get("metric", "cosine")
        self.index = None
        self.id_map: Dict[int, str] = {}
        self.reverse_id_map: Dict[str, int] = {}
        self.current_size = 0

    def build(self, embeddings: np.ndarray, ids: List[str]) -> None:  # retrieval.py:70-136
        emb = embeddings.astype(np.float32)
        if self.metric == "cosine":
            normalize_L2(emb)
        self.index = IndexFlatIP(self.dimension)
        self.index.add(emb)
        for i, item_id in enumerate(ids):
            self.id_map[i] = item_id
            self.reverse_id_map[item_id] = i
        self.current_size = len(emb)

    def add(self, embeddings: np.ndarray, ids: List[str]) -> None:  # retrieval.py:199-226
        if self.index is None:
            raise ValueError("Index not built yet")
        emb = embeddings.astype(np.float32)
        if self.metric == "cosine":
            normalize_L2(emb)
        self.index.add(emb)
        for i, item_id in enumerate(ids):
            self.id_map[self.current_size + i] = item_id
            self.reverse_id_map[item_id] = self.current_size + i
        self.current_size += len(emb)

    def search(self, query_embeddings: np.ndarray, k: int = 10, filter_ids=None):  # retrieval.py:141-197
        if self.index is None:
            raise ValueError("Index not built yet")
        q = query_embeddings.astype(np.float32)
        if q.ndim == 1:
            q = q.reshape(1, -1)
        if self.metric == "cosine":
            normalize_L2(q)
        k_search = min(k * 2, self.current_size) if filter_ids else k
        D, I = self.index.search(q, k_search)
        out_ids, out_d = [], []
        for r in range(len(q)):
            ids_r, d_r = [], []
            for j in range(k_search):
                idx = int(I[r, j])
                if idx >= 0 and idx in self.id_map:
                    item_id = self.id_map[idx]
                    if filter_ids is None or item_id in filter_ids:
                        ids_r.append(item_id)
                        d_r.append(float(D[r, j]))
                        if len(ids_r) >= k:
                            break
            out_ids.append(ids_r)
            out_d.append(d_r)
        return out_ids, out_d


def eval_twin_topk(user_emb: np.ndarray, item_emb: np.ndarray, train_items: Dict[int, list],
                   users: Sequence[int], top_k: int) -> Dict[int, list]:
    """The reference's own exact scorer, restated line for line in behaviour:
    scripts/evaluate_model.py:217-232 (np.dot, -inf mask of train items, argsort[::-1][:k])."""
    num_items = item_emb.shape[0]
    scores = np.dot(user_emb, item_emb.T)
    rec = {}
    for j, u in enumerate(users):
        s = scores[j].copy()
        for t in train_items.get(u, []):
            if t < num_items:
                s[t] = -np.inf
        rec[u] = np.argsort(s)[::-1][:top_k].tolist()
    return rec


def bf16_round(x: np.ndarray) -> np.ndarray:
    """Round-to-nearest-even fp32 -> bf16 -> fp32 (the GPU catalogue storage type in BASELINE cfg 3)."""
    x = np.ascontiguousarray(x, dtype=np.float32)
    u = x.view(np.uint32).astype(np.uint64)
    rounded = ((u + 0x7FFF + ((u >> 16) & 1)) >> 16) << 16
    return rounded.astype(np.uint32).view(np.float32)
