"""Drop-in B200 mirror of the exact inner-product path of the reference's `src/serving/retrieval.py`:
IndexBase (:16-46), FaissIndex Flat/cosine branch (:49-329) and RetrievalEngine (:505-692).

`B200FlatIndex.index` is a `FlatIPDeviceIndex`, the faiss-shaped object: `add(x)`, `search(q, k) -> (D, I)` with
faiss.IndexFlatIP's array contract (fp32 scores descending, int64 labels, -1 / -FLT_MAX padding).  The catalogue lives
in HBM as a bf16 tensor-core operand; scoring + top-K run in one fused tcgen05 kernel (csrc/topk.cu).  IVF / L2 /
Annoy / Milvus are outside this hot path and are rejected loudly.
"""
from __future__ import annotations

import hashlib
import logging
import pickle
import struct
import time
from abc import ABC, abstractmethod
from pathlib import Path
from typing import Any, Dict, List, Optional, Tuple

import numpy as np
import torch

from . import kernels as K
from ._native import pad64

logger = logging.getLogger("b200rec")
FLT_MAX = float(np.finfo(np.float32).max)


class IndexBase(ABC):
    """Abstract base class for ANN indices (reference retrieval.py:16-46)."""

    @abstractmethod
    def build(self, embeddings: np.ndarray, ids: List[str]):
        pass

    @abstractmethod
    def search(self, query_embeddings: np.ndarray, k: int = 10) -> Tuple[List[List[str]], List[List[float]]]:
        pass

    @abstractmethod
    def add(self, embeddings: np.ndarray, ids: List[str]):
        pass

    @abstractmethod
    def save(self, path: str):
        pass

    @abstractmethod
    def load(self, path: str):
        pass


class FlatIPDeviceIndex:
    """faiss.IndexFlatIP on one B200: rows appended with `add`, exact `search`.

    storage "bf16": rows rounded to bf16 (BASELINE config 3; scores are exact fp32 sums of the rounded products).
    storage "fp32": rows kept as split-bf16 x3 operands, i.e. fp32-grade products (|err| ~ 1e-6 * |score|), for the
    reference-scale catalogues where ids must match a CPU fp32 search except at ties within 1e-5.
    """

    def __init__(self, d: int, storage: str = "fp32", device: Optional[torch.device] = None, row_offset: int = 0):
        if storage not in ("bf16", "fp32"):
            raise ValueError("storage must be 'bf16' or 'fp32'")
        if not torch.cuda.is_available():
            raise RuntimeError("FlatIPDeviceIndex needs a CUDA (sm_100a) device; there is no CPU search path")
        self.d = int(d)
        self.storage = storage
        self.terms = 1 if storage == "bf16" else 3
        self.kpad = pad64(self.d)
        self.ld = self.terms * self.kpad
        self.device = device or torch.device("cuda", torch.cuda.current_device())
        self.row_offset = int(row_offset)
        self.ntotal = 0
        self.is_trained = True
        self._cat = torch.empty((0, self.ld), dtype=torch.bfloat16, device=self.device)
        self._f32 = torch.empty((0, self.d), dtype=torch.float32, device=self.device) if storage == "fp32" else None
        self._ws: Optional[torch.Tensor] = None

    # ------------------------------------------------------------------ rows
    def _reserve(self, n: int) -> None:
        if n <= self._cat.shape[0]:
            return
        cap = max(n, int(self._cat.shape[0] * 1.5), 1024)
        new = torch.empty((cap, self.ld), dtype=torch.bfloat16, device=self.device)
        new[: self.ntotal] = self._cat[: self.ntotal]
        self._cat = new
        if self._f32 is not None:
            nf = torch.empty((cap, self.d), dtype=torch.float32, device=self.device)
            nf[: self.ntotal] = self._f32[: self.ntotal]
            self._f32 = nf

    def add(self, x, normalize: bool = False) -> None:
        """Append rows (np.ndarray or torch tensor [n, d], any float dtype).  normalize=True applies faiss.normalize_L2."""
        xt = torch.as_tensor(x)
        if xt.dim() != 2 or xt.shape[1] != self.d:
            raise ValueError(f"expected [n, {self.d}] rows, got {tuple(xt.shape)}")
        n = xt.shape[0]
        if n == 0:
            return
        self._reserve(self.ntotal + n)
        step = 1 << 20
        for s in range(0, n, step):
            chunk = xt[s:s + step].to(self.device, dtype=torch.float32, non_blocking=True).contiguous()
            m = chunk.shape[0]
            dst = self._cat[self.ntotal + s: self.ntotal + s + m]
            f32 = self._f32[self.ntotal + s: self.ntotal + s + m] if self._f32 is not None else None
            K.N.check(K.N.lib().b200rec_normalize_rows(K.N.ptr(chunk), m, self.d, chunk.stride(0), int(normalize), 1,
                                                       K.N.ptr(f32), self.d, None, K.N.ptr(dst), self.kpad, self.terms,
                                                       1, K.N.stream()), "normalize_rows")
        self.ntotal += n

    def add_bf16_rows(self, rows: torch.Tensor) -> None:
        """Append rows that already are bf16 operands [n, ld] on the device (bulk loaders; storage 'bf16' only)."""
        assert self.storage == "bf16" and rows.dtype == torch.bfloat16 and rows.shape[1] == self.ld
        if self.ntotal == 0 and rows.is_contiguous() and rows.device == self.device:
            self._cat = rows
        else:
            self._reserve(self.ntotal + rows.shape[0])
            self._cat[self.ntotal: self.ntotal + rows.shape[0]] = rows
        self.ntotal += rows.shape[0]

    def reconstruct_n(self, i0: int = 0, n: Optional[int] = None) -> np.ndarray:
        n = self.ntotal - i0 if n is None else n
        if self._f32 is not None:
            return self._f32[i0:i0 + n].cpu().numpy()
        return self._cat[i0:i0 + n, : self.d].float().cpu().numpy()

    # ------------------------------------------------------------------ search
    def prepare_queries(self, q, normalize: bool = False) -> torch.Tensor:
        qt = torch.as_tensor(q)
        if qt.dim() == 1:
            qt = qt.reshape(1, -1)
        if qt.shape[1] != self.d:
            raise ValueError(f"expected [nq, {self.d}] queries, got {tuple(qt.shape)}")
        qt = qt.to(self.device, dtype=torch.float32, non_blocking=True).contiguous()
        _, _, qop = K.normalize_rows(qt, normalize=normalize, faiss_rule=True, want_f32=False, want_norms=False,
                                     terms=self.terms, side=0)
        return qop

    def _workspace(self, nq: int, k: int) -> torch.Tensor:
        need = K.topk_workspace_bytes(self.ntotal, self.ld, nq, k)
        if need == 0:
            raise RuntimeError(f"b200rec flat_ip_topk: {K.N.last_error()}")
        if self._ws is None or self._ws.numel() < need:
            self._ws = torch.empty((need,), dtype=torch.uint8, device=self.device)
        return self._ws

    def has_sample_pass(self, nq: int, k: int) -> bool:
        return self.ntotal > 0 and K.topk_has_sample(self.ntotal, self.ld, nq, k)

    def sample_device(self, q_op: torch.Tensor, k: int, k_out: Optional[int] = None, shards: int = 1) -> torch.Tensor:
        """Sampling pass only: the k_out largest group maxima per query [nq, k_out] (exchanged between row shards)."""
        return K.topk_sample(self._cat[: self.ntotal], q_op, k, self._workspace(q_op.shape[0], k), k_out, shards)

    def sample_fanout(self, q_op: torch.Tensor, k: int, k_out: int, shards: int, dst_vals) -> None:
        """Sampling pass whose [nq, k_out] maxima are stored to every address in dst_vals (peer gather buffers)."""
        K.topk_sample_fanout(self._cat[: self.ntotal], q_op, k, k_out, shards, dst_vals, self._workspace(q_op.shape[0], k))

    def search_fanout(self, q_op: torch.Tensor, k: int, tau_init, dst_scores, dst_ids) -> None:
        """Local top-k (global row ids) stored to every (dst_scores[d], dst_ids[d]) address (peer gather buffers)."""
        K.flat_ip_topk_fanout(self._cat[: self.ntotal], q_op, k, self.row_offset, tau_init, dst_scores, dst_ids,
                              self._workspace(q_op.shape[0], k))

    def search_device(self, q_op: torch.Tensor, k: int, exclude_indptr=None, exclude_rows=None, tau_init=None,
                      out=None):
        """q_op: bf16 operand [nq, ld] on the device -> (D, I) device tensors (written into `out` when given)."""
        nq = q_op.shape[0]
        if self.ntotal == 0:
            if out is not None:
                out[0].fill_(-FLT_MAX)
                out[1].fill_(-1)
                return out
            return (torch.full((nq, k), -FLT_MAX, dtype=torch.float32, device=self.device),
                    torch.full((nq, k), -1, dtype=torch.int64, device=self.device))
        return K.flat_ip_topk(self._cat[: self.ntotal], q_op, k, self.row_offset, exclude_indptr, exclude_rows,
                              self._workspace(nq, k), tau_init, out)

    def search(self, q, k: int, normalize: bool = False) -> Tuple[np.ndarray, np.ndarray]:
        """faiss contract: (D float32 [nq,k] descending, I int64 [nq,k]) as numpy arrays."""
        d, i = self.search_device(self.prepare_queries(q, normalize), k)
        return d.cpu().numpy(), i.cpu().numpy()


class B200FlatIndex(IndexBase):
    """`FaissIndex` on the exact Flat / inner-product branch (reference retrieval.py:49-329), B200-resident."""

    def __init__(self, config: Optional[Dict[str, Any]] = None):
        self.config = config or {}
        self.dimension = self.config.get("dimension", 128)
        self.index_factory = self.config.get("index_factory", "Flat")
        self.metric = self.config.get("metric", "cosine")
        self.nprobe = self.config.get("nprobe", 20)
        self.use_gpu = True
        self.storage = self.config.get("storage", "fp32")
        if self.metric not in ("cosine", "ip", "inner_product"):
            raise ValueError(f"metric {self.metric!r}: only the cosine / inner-product path is B200-native "
                             "(IndexFlatL2 is outside the hot path)")
        self.index: Optional[FlatIPDeviceIndex] = None
        self.id_map: Dict[int, str] = {}
        self.reverse_id_map: Dict[str, int] = {}
        self.current_size = 0

    def build(self, embeddings: np.ndarray, ids: List[str]):
        n_items = len(embeddings)
        logger.info("Building B200 flat index for %d items...", n_items)
        start = time.time()
        if "IVF" in self.index_factory and n_items >= 1024:
            logger.info("index_factory %r is approximate in the reference; the B200 index is always exact",
                        self.index_factory)
        self.index = FlatIPDeviceIndex(self.dimension, storage=self.storage)
        self.index.add(np.asarray(embeddings, dtype=np.float32), normalize=(self.metric == "cosine"))
        self.id_map = {}
        self.reverse_id_map = {}
        self._id_array = None
        for i, item_id in enumerate(ids):
            self.id_map[i] = item_id
            self.reverse_id_map[item_id] = i
        self.current_size = n_items
        logger.info("Index built in %.2f seconds", time.time() - start)

    def search(self, query_embeddings: np.ndarray, k: int = 10,
               filter_ids: Optional[List[str]] = None) -> Tuple[List[List[str]], List[List[float]]]:
        if self.index is None:
            raise ValueError("Index not built yet")
        if not (torch.is_tensor(query_embeddings) and query_embeddings.is_cuda):   # device tensors skip the host round trip
            query_embeddings = np.asarray(query_embeddings).astype(np.float32)
        if len(query_embeddings.shape) == 1:
            query_embeddings = query_embeddings.reshape(1, -1)
        k_search = min(k * 2, self.current_size) if filter_ids else k
        distances, indices = self.index.search(query_embeddings, k_search, normalize=(self.metric == "cosine"))
        if filter_ids is None and len(self.id_map) == self.current_size:
            # vectorised idx -> id map (the reference walks nq x k entries in Python, retrieval.py:177-195)
            if getattr(self, "_id_array", None) is None or len(self._id_array) != self.current_size:
                self._id_array = np.empty(self.current_size, dtype=object)
                self._id_array[:] = [self.id_map[i] for i in range(self.current_size)]
            valid = (indices >= 0) & (indices < self.current_size)
            names = self._id_array[np.where(valid, indices, 0)]
            if valid.all():
                return names[:, :k].tolist(), distances[:, :k].astype(float).tolist()
            return ([names[i][valid[i]][:k].tolist() for i in range(len(indices))],
                    [distances[i][valid[i]][:k].astype(float).tolist() for i in range(len(indices))])
        allowed = set(filter_ids) if filter_ids is not None else None
        batch_ids, batch_distances = [], []
        for i in range(len(query_embeddings)):
            item_ids, item_distances = [], []
            for j in range(k_search):
                idx = int(indices[i, j])
                if idx >= 0 and idx in self.id_map:
                    item_id = self.id_map[idx]
                    if allowed is None or item_id in allowed:
                        item_ids.append(item_id)
                        item_distances.append(float(distances[i, j]))
                        if len(item_ids) >= k:
                            break
            batch_ids.append(item_ids)
            batch_distances.append(item_distances)
        return batch_ids, batch_distances

    def add(self, embeddings: np.ndarray, ids: List[str]):
        if self.index is None:
            raise ValueError("Index not built yet")
        embeddings = np.asarray(embeddings).astype(np.float32)
        self.index.add(embeddings, normalize=(self.metric == "cosine"))
        for i, item_id in enumerate(ids):
            new_idx = self.current_size + i
            self.id_map[new_idx] = item_id
            self.reverse_id_map[item_id] = new_idx
        self.current_size += len(embeddings)
        logger.info("Added %d items to index. Total size: %d", len(embeddings), self.current_size)

    def update(self, embeddings: np.ndarray, ids: List[str]):
        logger.warning("Flat indices do not support in-place updates. Consider periodic rebuilds.")

    def remove(self, ids: List[str]):
        logger.warning("Flat indices do not support removal. Consider periodic rebuilds.")

    # faiss `IxFI` file layout (faiss/impl/index_write.cpp, 1.7.x): fourcc, d:int32, ntotal:int64, 2x dummy int64,
    # is_trained:uint8, metric:int32 (0 = inner product), n_floats:uint64, raw fp32 rows.
    def save(self, path: str):
        if self.index is None:
            raise ValueError("No index to save")
        path = Path(path)
        path.parent.mkdir(parents=True, exist_ok=True)
        rows = self.index.reconstruct_n(0, self.index.ntotal).astype(np.float32)
        with open(path.with_suffix(".faiss"), "wb") as f:
            f.write(b"IxFI")
            f.write(struct.pack("<iqqqBi", self.index.d, self.index.ntotal, 1 << 20, 1 << 20, 1, 0))
            f.write(struct.pack("<Q", rows.size))
            f.write(rows.tobytes())
        with open(path.with_suffix(".pkl"), "wb") as f:
            pickle.dump({"id_map": self.id_map, "reverse_id_map": self.reverse_id_map,
                         "current_size": self.current_size, "config": self.config}, f)
        logger.info("Saved index to %s", path)

    def load(self, path: str):
        path = Path(path)
        with open(path.with_suffix(".faiss"), "rb") as f:
            if f.read(4) != b"IxFI":
                raise ValueError("only faiss IndexFlatIP ('IxFI') files are supported by the B200 flat index")
            d, ntotal, _, _, _, metric = struct.unpack("<iqqqBi", f.read(struct.calcsize("<iqqqBi")))
            (nfloats,) = struct.unpack("<Q", f.read(8))
            rows = np.frombuffer(f.read(nfloats * 4), dtype=np.float32).reshape(ntotal, d)
        with open(path.with_suffix(".pkl"), "rb") as f:
            data = pickle.load(f)
        self.id_map = data["id_map"]
        self.reverse_id_map = data["reverse_id_map"]
        self._id_array = None
        self.current_size = data["current_size"]
        self.config = data["config"]
        self.dimension = d
        self.index = FlatIPDeviceIndex(d, storage=self.storage)
        self.index.add(rows, normalize=False)  # rows were normalised before they were saved
        logger.info("Loaded index from %s with %d items", path, self.current_size)


class RetrievalEngine:
    """High-level retrieval engine managing the index and the query cache (reference retrieval.py:505-692)."""

    def __init__(self, config: Dict[str, Any]):
        self.config = config
        self.index_type = config.get("index_type", "b200")
        self.top_k = config.get("top_k", 100)
        self.update_interval = config.get("update_interval_seconds", 300)
        self.index = self._create_index()
        self.cache: Dict[str, Any] = {}
        self.cache_ttl = config.get("cache_ttl", 300)
        self.last_cache_clear = time.time()
        self.total_queries = 0
        self.cache_hits = 0
        self.total_latency = 0

    def _create_index(self) -> IndexBase:
        index_config = dict(self.config.get(self.index_type, {}))
        index_config["dimension"] = self.config.get("embedding_dim", 128)
        if self.index_type in ("b200", "faiss"):
            return B200FlatIndex(index_config)
        if self.index_type in ("annoy", "milvus"):
            raise ValueError(f"index type {self.index_type!r} is not part of the B200 hot path (exact flat IP only)")
        raise ValueError(f"Unknown index type: {self.index_type}")

    def build_index(self, embeddings: np.ndarray, ids: List[str]):
        self.index.build(embeddings, ids)
        self.cache.clear()
        logger.info("Built index with %d items", len(embeddings))

    def retrieve(self, query_embeddings: np.ndarray, k: Optional[int] = None, filter_ids: Optional[List[str]] = None,
                 use_cache: bool = True) -> Tuple[List[List[str]], List[List[float]], Dict[str, Any]]:
        start_time = time.time()
        k = k or self.top_k
        cache_key = None
        if use_cache and filter_ids is None:
            cache_key = hashlib.md5(np.asarray(query_embeddings).tobytes()).hexdigest()
            if cache_key in self.cache:
                entry = self.cache[cache_key]
                if time.time() - entry["timestamp"] < self.cache_ttl:
                    self.cache_hits += 1
                    latency = time.time() - start_time
                    self.total_queries += 1
                    self.total_latency += latency
                    return entry["ids"], entry["scores"], {"latency_ms": latency * 1000, "cache_hit": True}
        item_ids, scores = self.index.search(query_embeddings, k, filter_ids)
        if use_cache and cache_key:
            self.cache[cache_key] = {"ids": item_ids, "scores": scores, "timestamp": time.time()}
            if time.time() - self.last_cache_clear > self.cache_ttl:
                self._clear_expired_cache()
        latency = time.time() - start_time
        self.total_queries += 1
        self.total_latency += latency
        flat = [s for lst in scores for s in lst]
        metrics = {"latency_ms": latency * 1000, "cache_hit": False,
                   "num_results": sum(len(ids) for ids in item_ids),
                   "avg_score": float(np.mean(flat)) if flat else float("nan")}
        return item_ids, scores, metrics

    def update_index(self, new_embeddings: np.ndarray, new_ids: List[str]):
        self.index.add(new_embeddings, new_ids)
        self.cache.clear()

    def _clear_expired_cache(self):
        now = time.time()
        expired = [key for key, entry in self.cache.items() if now - entry["timestamp"] > self.cache_ttl]
        for key in expired:
            del self.cache[key]
        self.last_cache_clear = now

    def get_metrics(self) -> Dict[str, Any]:
        return {"total_queries": self.total_queries,
                "avg_latency_ms": self.total_latency / max(self.total_queries, 1) * 1000,
                "cache_hit_rate": self.cache_hits / max(self.total_queries, 1),
                "cache_size": len(self.cache), "index_size": self.index.current_size, "index_type": self.index_type}

    def save(self, path: str):
        self.index.save(path)

    def load(self, path: str):
        self.index.load(path)
        self.cache.clear()


def exact_topk_eval(user_emb: torch.Tensor, item_index: FlatIPDeviceIndex, train_items: Dict[int, List[int]],
                    user_rows: List[int], k: int = 100) -> np.ndarray:
    """The offline-eval twin (scripts/evaluate_model.py:217-232, src/evaluation/metrics.py:381-396): top-k item rows per
    user with that user's train items excluded, through the same fused kernel (exclusion CSR instead of a -inf mask)."""
    lists = [np.unique(np.asarray(train_items.get(u, []), dtype=np.int64)) for u in user_rows]
    indptr = np.zeros(len(lists) + 1, dtype=np.int64)
    np.cumsum([len(x) for x in lists], out=indptr[1:])
    rows = np.concatenate(lists).astype(np.int32) if indptr[-1] > 0 else np.zeros((1,), dtype=np.int32)
    dev = item_index.device
    q_op = item_index.prepare_queries(user_emb, normalize=False)
    _, ids = item_index.search_device(q_op, k, torch.from_numpy(indptr).to(dev), torch.from_numpy(rows).to(dev))
    return ids.cpu().numpy()
