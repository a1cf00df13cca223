"""Request micro-batching for the serving loop ("next" row f4).

`RecommendationService.get_recommendations` (reference src/serving/service.py:183-302) runs, per request, one user-tower
forward on a [1, F] tensor (`_get_user_embedding`, :262-302) and one `retrieval_engine.retrieve(user_embedding, k)`
(:203-207): nq = 1, the regime in which the catalogue scan is pure HBM streaming (0.56 ms for 10 M x 128 on one B200,
whether 1 or 128 queries share it).  `RetrievalBatcher` lets concurrent requests of one asyncio event loop (the
reference's threading model: one uvicorn worker, service.py:372-380) share that scan: requests that arrive within
`max_wait_ms` (or until `max_batch` are pending) are stacked into ONE user-tower forward and ONE exact top-k search on
the device, and every caller gets exactly what the unbatched calls return."""
from __future__ import annotations

import asyncio
import time
from typing import Any, Dict, List, Optional, Tuple

import torch


class RetrievalBatcher:
    """await batcher.recommend(user_features, k) -> (item_ids, scores, metrics) — the retrieval stage of one request.

    model: TwoTowerModel (its get_user_embeddings is called on the stacked features);
    engine: RetrievalEngine (engine.index.search does the batched exact search; its query counters are updated).
    """

    def __init__(self, model, engine, max_batch: int = 256, max_wait_ms: float = 2.0):
        self.model = model
        self.engine = engine
        self.max_batch = int(max_batch)
        self.max_wait = float(max_wait_ms) / 1e3
        self._pending: List[Tuple[Dict[str, Any], int, asyncio.Future]] = []
        self._timer: Optional[asyncio.TimerHandle] = None
        self.batches = 0
        self.requests = 0

    async def recommend(self, user_features: Dict[str, Any], k: int):
        loop = asyncio.get_running_loop()
        fut = loop.create_future()
        self._pending.append((user_features, int(k), fut))
        if len(self._pending) >= self.max_batch:
            self._flush()
        elif self._timer is None:
            self._timer = loop.call_later(self.max_wait, self._flush)
        return await fut

    # ------------------------------------------------------------------ one batch
    @staticmethod
    def _stack(feature_dicts: List[Dict[str, Any]]) -> Dict[str, Any]:
        """[{"numerical": [1,F] or [F], "categorical": {name: [1]}}, ...] -> one dict with [n, F] / [n] tensors."""
        num = torch.cat([torch.as_tensor(f["numerical"]).reshape(1, -1) for f in feature_dicts], dim=0)
        cat: Dict[str, torch.Tensor] = {}
        for name in (feature_dicts[0].get("categorical") or {}):
            cat[name] = torch.cat([torch.as_tensor(f["categorical"][name]).reshape(1) for f in feature_dicts])
        return {"numerical": num, "categorical": cat}

    def _flush(self) -> None:
        if self._timer is not None:
            self._timer.cancel()
            self._timer = None
        batch, self._pending = self._pending[: self.max_batch], self._pending[self.max_batch:]
        if not batch:
            return
        if self._pending:  # more than one batch was waiting: keep draining
            self._timer = asyncio.get_running_loop().call_soon(self._flush)
        t0 = time.perf_counter()
        try:
            feats = self._stack([b[0] for b in batch])
            dev = next(self.model.parameters()).device
            feats = {"numerical": feats["numerical"].to(dev, torch.float32),
                     "categorical": {n: t.to(dev) for n, t in feats["categorical"].items()}}
            with torch.no_grad():
                emb = self.model.get_user_embeddings(feats)                 # ONE tower forward for the whole batch
            k_max = max(b[1] for b in batch)
            ids, scores = self.engine.index.search(emb, k=k_max)           # ONE exact top-k search for the whole batch
            seconds = time.perf_counter() - t0
            latency_ms = seconds * 1e3
            # every request of the batch waited for the one shared search: charge `seconds` to each of them, so the
            # engine's avg_latency_ms stays the per-request latency (RetrievalEngine accounts in seconds)
            self.engine.account(seconds * len(batch), calls=len(batch))
            self.batches += 1
            self.requests += len(batch)
            for row, (_, k, fut) in enumerate(batch):
                if not fut.done():
                    fut.set_result(([ids[row][:k]], [scores[row][:k]],
                                    {"latency_ms": latency_ms, "cache_hit": False, "num_results": len(ids[row][:k]),
                                     "batch_size": len(batch)}))
        except Exception as exc:  # noqa: BLE001 - every waiting request sees the failure, as its own call would have
            for _, _, fut in batch:
                if not fut.done():
                    fut.set_exception(exc)
