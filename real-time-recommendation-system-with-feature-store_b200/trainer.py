"""Mirror of the reference's `src/training/trainers/two_tower.py` (TwoTowerTrainer :25-273) with the optimiser step on
B200 kernels: one flat fp32 parameter / gradient buffer, one fused sum-of-squares + clip coefficient (clip_grad_norm_
:144), one fused Adam with L2-coupled weight decay (:60-64,146); embedding tables that opted into row-sparse training
get a row-sparse Adam on their touched rows.  `train_step` is the step body of `train_epoch` (:98-151) on tensors that
are already on the device — what bench.py times."""
from __future__ import annotations

import logging
import os
import time
from pathlib import Path
from typing import Any, Dict, List, Optional

import torch

from . import kernels as K
from .two_tower import TwoTowerModel

logger = logging.getLogger("b200rec")


class FlatAdam:
    """torch.optim.Adam(lr, betas=(0.9,0.999), eps=1e-8, weight_decay) + clip_grad_norm_(max_norm) on flat buffers."""

    def __init__(self, params, lr: float = 1e-3, weight_decay: float = 1e-5, betas=(0.9, 0.999), eps: float = 1e-8,
                 max_grad_norm: Optional[float] = 1.0):
        params = [p for p in params if p.requires_grad]
        self.params = params                      # torch.optim order (state_dict indices)
        self.dense = [p for p in params if not getattr(p, "_b200_sparse", False)]
        self.sparse = [p for p in params if getattr(p, "_b200_sparse", False)]
        if not self.dense:
            raise ValueError("FlatAdam needs at least one dense parameter")
        dev = self.dense[0].device
        if dev.type != "cuda":
            raise RuntimeError("FlatAdam runs on CUDA parameters only (no CPU path)")
        n = sum(p.numel() for p in self.dense)
        self.flat = torch.empty((n,), dtype=torch.float32, device=dev)
        self.grad = torch.zeros((n,), dtype=torch.float32, device=dev)
        self.m = torch.zeros_like(self.flat)
        self.v = torch.zeros_like(self.flat)
        off = 0
        self._seg = []                            # (offset, numel) of every dense parameter inside the flat buffers
        for p in self.dense:
            k = p.numel()
            self.flat[off:off + k].copy_(p.data.reshape(-1))
            p.data = self.flat[off:off + k].view(p.shape)
            p.grad = self.grad[off:off + k].view(p.shape)
            self._seg.append((off, k))
            off += k
        # torch.optim.Adam skips a parameter whose .grad is None (never reached by backward): no update, no weight decay,
        # no moment decay.  Gradients here are views of one buffer, so "reached by backward" is tracked per segment: an
        # autograd hook marks dense parameters, GatherConcatFn marks tables it scatters into in place.
        self._touched = [False] * len(self.dense)
        self._touch_sticky = False
        # Dense embedding tables (two_tower.py marks them _b200_table): Adam + weight decay moves every row each step, but
        # the step's gradient is non-zero only on the rows the batch touched.  The scatter kernel raises a per-row flag
        # (ops.GatherConcatFn.backward) and b200rec_adam_table / b200rec_table_sumsq read the gradient of flagged rows
        # only (24 B instead of 32 per parameter).  Any other way a table gets its gradient falls back to the flat kernels.
        self._table = [bool(getattr(p, "_b200_table", False)) and p.dim() == 2 and p.shape[1] <= 128
                       and os.environ.get("B200REC_TABLE_ADAM", "1") != "0" for p in self.dense]
        self._flag_ok = [True] * len(self.dense)
        for i, p in enumerate(self.dense):
            p._b200_touch = (self, i)
            # fires after every backward that reaches p, also when the kernels wrote straight into p.grad and handed
            # autograd no tensor: it only records "reached by backward" (flagged=None)
            p.register_post_accumulate_grad_hook(lambda t, _i=i: self._mark(_i, None))
            if self._table[i]:
                p._b200_row_flags = torch.zeros((p.shape[0],), dtype=torch.int32, device=dev)
                # a DEFINED gradient tensor arriving through autograd (any op other than the flagging scatter) was
                # accumulated densely: this step's update of the table must read the whole gradient
                # (autograd also calls tensor hooks with None when a backward returned no tensor for p)
                p.register_hook(lambda g, _i=i: self._mark(_i, False) if g is not None else None)
        self.sparse_state = {id(p): (torch.zeros_like(p.data), torch.zeros_like(p.data)) for p in self.sparse}
        self.lr, self.wd, self.betas, self.eps, self.max_grad_norm = lr, weight_decay, betas, eps, max_grad_norm
        self.step_count = 0
        self._grad_clean = True
        self.clear_grad = True                    # step() zeroes the gradients it consumes (False: keep them readable)
        self._acc = torch.zeros((1,), dtype=torch.float64, device=dev)
        self._coef = torch.ones((1,), dtype=torch.float32, device=dev)
        self.grad_norm = torch.zeros((1,), dtype=torch.float32, device=dev)
        self.param_groups = [{"lr": lr}]
        # CUDA-graph training (TwoTowerTrainer.enable_cuda_graph): step counter, lr and Adam's bias corrections on the device
        self.dev_state = None

    def _mark(self, i: int, flagged: Optional[bool] = False) -> None:
        """Parameter i received a gradient this step.  flagged=True: through the row-flagging scatter kernel;
        False: some other writer (the table's update falls back to the flat kernels); None: no statement."""
        self._touched[i] = True
        if flagged is False:
            self._flag_ok[i] = False

    def _plan(self):
        """(flat [lo, hi) element ranges, table segment indices) of the parameters that received a gradient this step:
        tables whose gradient came through the flagging scatter are updated row-wise, everything else by the flat
        kernels (adjacent segments merged into one launch)."""
        everything = self._touch_sticky or all(self._touched)
        flat, tables = [], []
        for i, (off, k) in enumerate(self._seg):
            if not (everything or self._touched[i]):
                continue
            if self._table[i] and self._touched[i] and self._flag_ok[i] and not self._touch_sticky:
                tables.append(i)
            elif flat and flat[-1][1] == off:
                flat[-1] = (flat[-1][0], off + k)
            else:
                flat.append((off, off + k))
        return flat, tables

    def _table_views(self, i: int):
        off, k = self._seg[i]
        shape = self.dense[i].shape
        return (self.flat[off:off + k].view(shape), self.grad[off:off + k].view(shape), self.m[off:off + k].view(shape),
                self.v[off:off + k].view(shape), self.dense[i]._b200_row_flags)

    def touch_all(self) -> None:
        """Gradients are written straight into .grad by the caller (no autograd): update every parameter."""
        self._touch_sticky = True

    def _runs(self):
        """Contiguous [lo, hi) element ranges of the flat buffer whose parameters received a gradient this step."""
        if self._touch_sticky or all(self._touched):
            return [(0, self.flat.numel())]
        runs = []
        for (off, k), t in zip(self._seg, self._touched):
            if not t:
                continue
            if runs and runs[-1][1] == off:
                runs[-1] = (runs[-1][0], off + k)
            else:
                runs.append((off, off + k))
        return runs

    def _reduce_runs(self):
        """[lo, hi) ranges of the flat gradient buffer that data parallel must all-reduce: everything except the dense
        embedding tables whose rows are exchanged sample-wise (None = the whole buffer)."""
        if not any(getattr(p, "_b200_row_exchange", None) is not None for p in self.dense):
            return None
        runs = []
        for (off, k), p in zip(self._seg, self.dense):
            if getattr(p, "_b200_row_exchange", None) is not None:
                continue
            if runs and runs[-1][1] == off:
                runs[-1] = (runs[-1][0], off + k)
            else:
                runs.append((off, off + k))
        return runs

    def zero_grad(self, set_to_none: bool = False) -> None:
        # step() resets every gradient it consumed (adam_dense clear_grad) and untouched segments were never written:
        # the buffer is already zero unless something accumulated into it since the last step
        if not self._grad_clean:
            self.grad.zero_()
        self._grad_clean = False
        self._touched = [False] * len(self.dense)
        self._flag_ok = [True] * len(self.dense)
        for p in self.sparse:
            p._b200_sparse_grad = None

    def _sparse_lists(self, p):
        lists = getattr(p, "_b200_sparse_grad", None) or []
        dp = getattr(self, "dp", None)
        if dp is not None and lists:
            # data parallel: every replica applies the touched rows of ALL replicas (tables are replicated)
            rows = torch.cat([r for r, _, _ in lists]) if len(lists) > 1 else lists[0][0]
            vals = torch.cat([v for _, v, _ in lists]) if len(lists) > 1 else lists[0][1]
            rows, vals = dp.gather_sparse(rows, vals)
            if os.environ.get("B200REC_EMB_BWD", "atomic") != "sorted":
                from .ops import sparse_slot_map
                return [K.sparse_claim_accumulate(rows, vals, vals.shape[1], p.shape[0], sparse_slot_map(p), 0)]
            return [K.embedding_sparse_grad(rows, vals, vals.shape[1], p.shape[0], 0)]
        if len(lists) <= 1:
            return lists
        # several gathers hit the table this step: concatenate and coalesce again (invalid tail rows are 0 = padding)
        rows = torch.cat([r for r, _, _ in lists])
        vals = torch.cat([v for _, v, _ in lists])
        return [K.embedding_sparse_grad(rows, vals, vals.shape[1], p.shape[0], 0)]

    def enable_device_state(self) -> None:
        if self.dev_state is None:
            dev = self.flat.device
            self.dev_state = {"step": torch.full((1,), self.step_count, dtype=torch.int64, device=dev),
                              "lr": torch.full((1,), float(self.param_groups[0]["lr"]), dtype=torch.float32, device=dev),
                              "hyper": torch.zeros((3,), dtype=torch.float32, device=dev), "lr_host": None}

    def sync_device_state(self) -> None:
        """Before a graph replay: the scheduler may have changed the learning rate and eager steps (ragged batches) may
        have advanced the step counter (tiny fills, outside the graph, only when something changed)."""
        lr = float(self.param_groups[0]["lr"])
        if self.dev_state["lr_host"] != lr:
            self.dev_state["lr"].fill_(lr)
            self.dev_state["lr_host"] = lr
        if self.dev_state.get("step_host") != self.step_count:
            self.dev_state["step"].fill_(self.step_count)
            self.dev_state["step_host"] = self.step_count

    def step(self) -> None:
        self.step_count += 1
        lr = self.param_groups[0]["lr"]
        clip = None
        dp = getattr(self, "dp", None)
        if dp is not None:
            # losses are normalised by the global batch: gradients add up.  Tables whose touched rows were exchanged in
            # the backward (dist.DataParallel) already hold the global sum: only the other segments are all-reduced.
            dp.reduce_dense_grad_(self.grad, self._reduce_runs())
        sparse = [(p, self._sparse_lists(p)) for p in self.sparse]
        runs, tables = self._plan()
        if self.max_grad_norm is not None:
            self._acc.zero_()
            if tables:
                # untouched parameters hold zero gradients: the norm is taken over what received one
                for lo, hi in runs:
                    K.sumsq_(self.grad[lo:hi], self._acc)
                for i in tables:
                    _, g2, _, _, flags = self._table_views(i)
                    K.table_sumsq_(g2, flags, self._acc)
            else:
                K.sumsq_(self.grad, self._acc)
            for _, lists in sparse:
                for _, vals, _ in lists:
                    K.sumsq_(vals.reshape(-1), self._acc)   # rows beyond n are zero-filled by the coalescing kernel
            K.clip_coef(self._acc, float(self.max_grad_norm), self._coef, self.grad_norm)
            clip = self._coef
        self._grad_clean = bool(self.clear_grad)   # every run below clears the gradients it consumes; the rest was never written
        if self.dev_state is not None and getattr(self, "use_device_state", False):
            hyper = self.dev_state["hyper"]   # written by K.train_step_begin at the start of this step
            for i in tables:
                p2, g2, m2, v2, flags = self._table_views(i)
                K.adam_table_(p2, g2, m2, v2, flags, 0.0, self.betas[0], self.betas[1], self.eps, self.wd, 1, clip, hyper,
                              clear_grad=self.clear_grad)
            for lo, hi in runs:
                K.adam_dense_dev_(self.flat[lo:hi], self.grad[lo:hi], self.m[lo:hi], self.v[lo:hi], self.betas[0],
                                  self.betas[1], self.eps, self.wd, hyper, clip, clear_grad=self.clear_grad)
            for p, lists in sparse:
                m, v = self.sparse_state[id(p)]
                for rows, vals, n in lists:
                    K.sparse_adam_dev_(p.data, m, v, rows, vals, n, self.betas[0], self.betas[1], self.eps, hyper, clip)
            return
        for i in tables:
            p2, g2, m2, v2, flags = self._table_views(i)
            K.adam_table_(p2, g2, m2, v2, flags, lr, self.betas[0], self.betas[1], self.eps, self.wd, self.step_count, clip,
                          None, clear_grad=self.clear_grad)
        for lo, hi in runs:
            K.adam_dense_(self.flat[lo:hi], self.grad[lo:hi], self.m[lo:hi], self.v[lo:hi], lr, self.betas[0],
                          self.betas[1], self.eps, self.wd, self.step_count, clip, clear_grad=self.clear_grad)
        for p, lists in sparse:
            m, v = self.sparse_state[id(p)]
            for rows, vals, n in lists:
                K.sparse_adam_(p.data, m, v, rows, vals, n, lr, self.betas[0], self.betas[1], self.eps,
                               self.step_count, clip)

    def _moments(self, p):
        if getattr(p, "_b200_sparse", False):
            return self.sparse_state[id(p)]
        off, k = self._seg[p._b200_touch[1]]
        return self.m[off:off + k].view(p.shape), self.v[off:off + k].view(p.shape)

    def state_dict(self) -> Dict[str, Any]:
        """torch.optim.Adam's state_dict layout (what the reference stores under 'optimizer_state',
        trainers/two_tower.py:209): parameter i of model.parameters() -> {step, exp_avg, exp_avg_sq}, one param group.
        A parameter Adam has never stepped (state-less in torch) still carries its zero moments here."""
        state = {}
        for i, p in enumerate(self.params):
            m, v = self._moments(p)
            state[i] = {"step": torch.tensor(float(self.step_count)), "exp_avg": m.detach().clone(),
                        "exp_avg_sq": v.detach().clone()}
        group = {"lr": self.param_groups[0]["lr"], "betas": tuple(self.betas), "eps": self.eps, "weight_decay": self.wd,
                 "amsgrad": False, "maximize": False, "foreach": None, "capturable": False, "differentiable": False,
                 "fused": None, "decoupled_weight_decay": False, "params": list(range(len(self.params)))}
        return {"state": state, "param_groups": [group]}

    def load_state_dict(self, sd: Dict[str, Any]) -> None:
        """Accepts the layout above, i.e. also a state_dict saved by torch.optim.Adam over the same model (the reference
        trainer's checkpoints).  Parameters missing from `state` (never stepped) keep zero moments."""
        group = sd["param_groups"][0]
        self.param_groups[0]["lr"] = group["lr"]
        self.betas, self.eps, self.wd = tuple(group["betas"]), group["eps"], group["weight_decay"]
        steps = []
        for slot, i in enumerate(group["params"]):
            st = sd["state"].get(i, sd["state"].get(str(i)))
            m, v = self._moments(self.params[slot])
            if st is None:
                m.zero_()
                v.zero_()
                continue
            m.copy_(st["exp_avg"].to(m.device))
            v.copy_(st["exp_avg_sq"].to(v.device))
            steps.append(int(float(st["step"])))
        self.step_count = max(steps) if steps else 0
        if self.dev_state is not None:
            self.dev_state["step_host"] = None
            self.dev_state["lr_host"] = None


class TwoTowerTrainer:
    """Trainer for the Two-Tower model (reference trainers/two_tower.py:25-273): same constructor, same methods, same
    checkpoint keys; the step runs entirely in b200rec kernels."""

    def __init__(self, model: TwoTowerModel, train_loader, val_loader, config: Dict[str, Any], device: str = "cuda"):
        if not str(device).startswith("cuda"):
            raise RuntimeError("b200rec TwoTowerTrainer needs device='cuda' (no CPU path)")
        self.model = model.to(device)
        self.train_loader = train_loader
        self.val_loader = val_loader
        self.config = config
        self.device = device
        self.optimizer = FlatAdam(self.model.parameters(), lr=config.get("learning_rate", 0.001),
                                  weight_decay=config.get("weight_decay", 1e-5),
                                  max_grad_norm=config.get("max_grad_norm", 1.0))
        self.scheduler = torch.optim.lr_scheduler.ReduceLROnPlateau(_LRShim(self.optimizer), mode="min", factor=0.5,
                                                                    patience=2)
        self.early_stopping_patience = config.get("early_stopping_patience", 5)
        self.best_val_loss = float("inf")
        self.patience_counter = 0
        self.checkpoint_dir = Path(config.get("checkpoint_dir", "models/checkpoints"))
        self.checkpoint_dir.mkdir(parents=True, exist_ok=True)
        self.train_losses: List[float] = []
        self.val_losses: List[float] = []

    # ------------------------------------------------------------------ CUDA-graph replay of the hot step
    def enable_cuda_graph(self, warm_steps: int = 2) -> bool:
        """Replay the training step as ONE CUDA graph per input shape (the eager step is ~210 kernel launches and
        host-launch-bound at ~18 us each).  The first `warm_steps` calls of a shape run eagerly, the next one is
        captured; shapes that differ (the ragged last batch) keep running eagerly.  What changes from step to step lives
        on the device: step counter / lr / Adam bias corrections (b200rec_train_step_begin, *_adam_*_dev) and a
        per-step salt of every dropout seed.  Under data parallel the NCCL collectives are part of the captured graph."""
        self._graph_warm = int(warm_steps)
        self._graphs: Dict[Any, Any] = {}
        self._graph_seen: Dict[Any, int] = {}
        self.optimizer.enable_device_state()
        return self._graph_capturable()

    def release_graphs(self) -> None:
        """Drop the captured steps (their NCCL nodes must be gone before the process group is destroyed)."""
        if getattr(self, "_graphs", None):
            torch.cuda.synchronize()
            self._graphs.clear()
            self._graph_seen.clear()

    def _graph_capturable(self) -> bool:
        # NCCL collectives issued through torch.distributed are capturable (the process group's internal stream joins the
        # capture through events); every replica captures and replays the same sequence.  B200REC_DP_GRAPH=0 keeps the
        # data-parallel step eager.
        return getattr(self.model, "dp", None) is None or os.environ.get("B200REC_DP_GRAPH", "1") != "0"

    def _graph_key(self, uf, pf, nf, uc, ic):
        cat = lambda d: tuple((k, tuple(v.shape), v.dtype) for k, v in (d or {}).items())
        return (tuple(uf.shape), tuple(pf.shape), None if nf is None else tuple(nf.shape), cat(uc), cat(ic),
                self.model.training)

    def _graphed_step(self, key, uf, pf, nf, uc, ic):
        entry = self._graphs.get(key)
        if entry is None:
            static = {"uf": uf.clone(), "pf": pf.clone(), "nf": None if nf is None else nf.clone(),
                      "uc": {k: v.clone() for k, v in (uc or {}).items()},
                      "ic": {k: v.clone() for k, v in (ic or {}).items()}}
            opt = self.optimizer
            opt.sync_device_state()
            salt = (int(torch.initial_seed()) * 0x9E3779B97F4A7C15 + 0x51ED27) & 0xFFFFFFFFFFFFFFFF
            graph = torch.cuda.CUDAGraph()
            opt.use_device_state = True
            try:
                with torch.cuda.graph(graph):
                    K.train_step_begin(opt.dev_state["step"], opt.dev_state["lr"], opt.betas[0], opt.betas[1],
                                       opt.dev_state["hyper"], salt or 1)
                    loss = self._step_body(static["uf"], static["pf"], static["nf"], static["uc"] or None,
                                           static["ic"] or None)
            finally:
                opt.use_device_state = False
            opt.step_count -= 1          # the capture ran the Python of one step without executing it
            entry = self._graphs[key] = (graph, static, loss)
        graph, static, loss = entry
        static["uf"].copy_(uf)
        static["pf"].copy_(pf)
        if nf is not None:
            static["nf"].copy_(nf)
        for k, v in (uc or {}).items():
            static["uc"][k].copy_(v)
        for k, v in (ic or {}).items():
            static["ic"][k].copy_(v)
        self.optimizer.sync_device_state()
        if not self.optimizer._grad_clean:       # something accumulated gradients outside a step: the captured step
            self.optimizer.grad.zero_()          # (which starts from a clean buffer) must not add to them
        graph.replay()
        self.optimizer._grad_clean = bool(self.optimizer.clear_grad)
        self.optimizer.step_count += 1
        self.optimizer.dev_state["step_host"] = self.optimizer.step_count   # the graph's first node incremented it
        return loss                      # static tensor: overwritten by the next replay of this shape

    # ------------------------------------------------------------------ the hot step (trainers/two_tower.py:98-146)
    def train_step(self, user_features: torch.Tensor, pos_item_features: torch.Tensor,
                   neg_item_features: Optional[torch.Tensor] = None,
                   user_categorical: Optional[Dict[str, torch.Tensor]] = None,
                   item_categorical: Optional[Dict[str, torch.Tensor]] = None) -> torch.Tensor:
        if getattr(self, "_graphs", None) is not None and self._graph_capturable():
            key = self._graph_key(user_features, pos_item_features, neg_item_features, user_categorical, item_categorical)
            seen = self._graph_seen.get(key, 0)
            self._graph_seen[key] = seen + 1
            if seen >= self._graph_warm:
                return self._graphed_step(key, user_features, pos_item_features, neg_item_features, user_categorical,
                                          item_categorical)
        return self._step_body(user_features, pos_item_features, neg_item_features, user_categorical, item_categorical)

    def _step_body(self, user_features, pos_item_features, neg_item_features=None, user_categorical=None,
                   item_categorical=None) -> torch.Tensor:
        model = self.model
        dp = getattr(model, "dp", None)
        self.optimizer.dp = dp
        self.optimizer.zero_grad()
        # The two towers are independent until the loss: the user tower runs on a side stream next to the item tower
        # (autograd replays each tower's backward on the stream of its forward, so the backward overlaps the same way;
        # every launch of a tower fills less than one wave of the 148 SMs).  Under data parallel only when each tower
        # has its own peer all-reduce context (dist._TowerDP): BatchNorm all-reduces of both towers issued on ONE NCCL
        # communicator from two streams would not be ordered.
        overlap = ((dp is None or getattr(dp, "streams_safe", False))
                   and os.environ.get("B200REC_OVERLAP", "1") != "0")
        main = torch.cuda.current_stream()
        if overlap:
            side = self._side_stream()
            side.wait_stream(main)
            with torch.cuda.stream(side):
                u = model.get_user_embeddings({"numerical": user_features, "categorical": user_categorical or {}})
            u.record_stream(main)
        else:
            u = model.get_user_embeddings({"numerical": user_features, "categorical": user_categorical or {}})
        p = model.get_item_embeddings({"numerical": pos_item_features, "categorical": item_categorical or {}})
        if neg_item_features is not None:
            batch_size, num_neg, feat_dim = neg_item_features.shape
            n = model.get_item_embeddings({"numerical": neg_item_features.view(-1, feat_dim), "categorical": {}})
            explicit_loss = model.contrastive_loss(u, p, n)
            if dp is not None:
                explicit_loss = explicit_loss / dp.world   # this replica's share of the global-batch mean
            in_batch_loss = model.in_batch_negative_loss(u, p)
            loss = 0.7 * explicit_loss + 0.3 * in_batch_loss
        else:
            loss = model.in_batch_negative_loss(u, p)
        if overlap:
            main.wait_stream(side)       # the loss kernels (main stream) read u
        loss.backward()
        if overlap:
            main.wait_stream(side)       # the user tower's gradients were accumulated on the side stream
        self.optimizer.step()
        return dp.global_loss(loss) if dp is not None else loss.detach()

    def _side_stream(self) -> torch.cuda.Stream:
        st = getattr(self, "_side", None)
        if st is None:
            st = self._side = torch.cuda.Stream(device=self.device)
        return st

    def train_epoch(self, epoch: int) -> float:
        self.model.train()
        total = torch.zeros((), dtype=torch.float32, device=self.device)
        num_batches = 0
        for batch in self.train_loader:
            uf = batch["user_features"].to(self.device, non_blocking=True)
            pf = batch["pos_item_features"].to(self.device, non_blocking=True)
            nf = batch["neg_item_features"].to(self.device, non_blocking=True) if "neg_item_features" in batch else None
            total += self.train_step(uf, pf, nf)   # no per-step host sync; one .item() per epoch
            num_batches += 1
        avg_loss = float(total.item()) / num_batches if num_batches > 0 else 0.0
        from . import ops
        ops.check_index_errors(self.device)     # an out-of-range categorical id raises IndexError, as nn.Embedding does
        self.train_losses.append(avg_loss)
        return avg_loss

    @torch.no_grad()
    def validate(self) -> float:
        self.model.eval()
        total = torch.zeros((), dtype=torch.float32, device=self.device)
        num_batches = 0
        for batch in self.val_loader:
            uf = batch["user_features"].to(self.device, non_blocking=True)
            pf = batch["pos_item_features"].to(self.device, non_blocking=True)
            u = self.model.get_user_embeddings({"numerical": uf, "categorical": {}})
            p = self.model.get_item_embeddings({"numerical": pf, "categorical": {}})
            total += self.model.in_batch_negative_loss(u, p)
            num_batches += 1
        dp = getattr(self.model, "dp", None)
        if dp is not None:
            # data parallel: in_batch_negative_loss returns this replica's SHARE of the global-batch loss, and the
            # scheduler / early stopping below must decide identically on every replica (a per-rank value lets the
            # learning rates diverge and replicas leave train() at different epochs, which deadlocks the collectives)
            counts = dp.gather_counts(num_batches)
            if len(set(counts)) != 1:
                raise RuntimeError(f"data-parallel validation needs the same number of batches on every replica, got {counts}")
            total = dp.global_loss(total)
        avg_loss = float(total.item()) / num_batches if num_batches > 0 else 0.0
        self.val_losses.append(avg_loss)
        return avg_loss

    def save_checkpoint(self, epoch: int, is_best: bool = False) -> None:
        checkpoint = {"epoch": epoch, "user_tower_state": self.model.user_tower.state_dict(),
                      "item_tower_state": self.model.item_tower.state_dict(),
                      "temperature": self.model.temperature, "user_bias": self.model.user_bias,
                      "item_bias": self.model.item_bias, "optimizer_state": self.optimizer.state_dict(),
                      "train_losses": self.train_losses, "val_losses": self.val_losses}
        torch.save(checkpoint, self.checkpoint_dir / "two_tower_latest.pth")
        if is_best:
            torch.save(checkpoint, self.checkpoint_dir / "two_tower_best.pth")

    def train(self, num_epochs: int) -> Dict[str, List[float]]:
        for epoch in range(1, num_epochs + 1):
            start = time.time()
            train_loss = self.train_epoch(epoch)
            val_loss = self.validate()
            self.scheduler.step(val_loss)
            logger.info("Epoch %d/%d - train %.4f val %.4f (%.1fs)", epoch, num_epochs, train_loss, val_loss,
                        time.time() - start)
            is_best = val_loss < self.best_val_loss
            if is_best:
                self.best_val_loss = val_loss
                self.patience_counter = 0
            else:
                self.patience_counter += 1
            self.save_checkpoint(epoch, is_best)
            if self.patience_counter >= self.early_stopping_patience:
                logger.info("Early stopping triggered after %d epochs", epoch)
                break
        return {"train_losses": self.train_losses, "val_losses": self.val_losses}


class _LRShim(torch.optim.Optimizer):
    """Lets torch's ReduceLROnPlateau drive FlatAdam's learning rate (it only touches param_groups[*]['lr'])."""

    def __init__(self, flat: FlatAdam):
        self._flat = flat
        super().__init__([torch.nn.Parameter(torch.zeros(1))], {"lr": flat.param_groups[0]["lr"]})
        self.param_groups = flat.param_groups
        for g in self.param_groups:
            g.setdefault("params", [])

    def step(self, closure=None):  # pragma: no cover - never called
        return None
